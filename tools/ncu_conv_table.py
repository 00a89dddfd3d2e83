"""Per-layer table of the sparse convolutions from an ncu launch list (gpu__time_duration.sum of every launch,
cold-cache and serialised) + the pair counts written by tools/profile_forward.py: duration of the pair-GEMM and of the
reduce/epilogue (or the fused stem / persistent kernel), algorithmic gather and scatter bytes and the
GB/s they imply, against the measured HBM peak.
usage: python tools/ncu_conv_table.py launches.csv profile_forward_layers.json [out.json]"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(csv_path, layers_path, out_path=None):
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')) / 1e3) for x in csv.DictReader(lines)]
    meta = json.load(open(layers_path))
    layers = meta['layers']
    conv = [(n, t) for n, t in rows if any(k in n for k in ('k_pairgemm_tc', 'k_pairgemm_tma', 'k_reduce_epilogue', 'k_stem_direct'))]
    per_fwd = 2 * (1 + 12 * 2)                       # two encoders: stem + 12 x (pair-GEMM, reduce)
    tail = 4                                         # the two BEV Conv2d run on the same kernels (pair-GEMM + reduce each) after the encoders
    if len(conv) % (per_fwd + tail) != 0:
        tail = 0                                     # IR_CONV2D=simt
    last = conv[-(per_fwd + tail):len(conv) - tail]  # the encoder layers of the last forward of the run
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    out, i = [], 0
    print(f'{"encoder":9s} {"L":>2s} {"Cin>Cout":>9s} {"K":>2s} {"rows":>6s} {"pairs":>7s} | {"gemm us":>8s} {"gather GB/s":>11s} | {"reduce us":>9s} {"scatter GB/s":>12s} | {"conv GB/s":>9s} {"of HBM":>6s}')
    tg = tr = tb = 0.0
    for L in layers:
        if L['layer'] == 0:
            name, t_red = last[i]; i += 1
            t_gemm = 0.0
            assert 'k_stem_direct' in name, name
        else:
            (n1, t_gemm), (n2, t_red) = last[i], last[i + 1]; i += 2
            assert 'pairgemm' in n1 and 'reduce' in n2, (n1, n2)
        G, S, W = L['bytes_gather'], L['bytes_scatter'], L['bytes_weights']
        g_gbs = (G + W) / (t_gemm * 1e-6) / 1e9 if t_gemm else float('nan')
        s_gbs = S / (t_red * 1e-6) / 1e9
        c_gbs = (G + S + W) / ((t_gemm + t_red) * 1e-6) / 1e9
        tg += t_gemm; tr += t_red; tb += G + S + W
        out.append(dict(L, gemm_us=round(t_gemm, 2), reduce_us=round(t_red, 2), gather_GBs=round(g_gbs, 1) if t_gemm else None,
                        scatter_GBs=round(s_gbs, 1), conv_GBs=round(c_gbs, 1), conv_frac_of_hbm=round(c_gbs / peak, 3)))
        print(f'{L["encoder"]:9s} {L["layer"]:2d} {L["cin"]:4d}>{L["cout"]:<4d} {L["K"]:2d} {L["rows_out"]:6d} {L["pairs"]:7d} | {t_gemm:8.2f} '
              f'{g_gbs:11.0f} | {t_red:9.2f} {s_gbs:12.0f} | {c_gbs:9.0f} {c_gbs / peak:6.2f}')
    tot = tb / ((tg + tr) * 1e-6) / 1e9
    print(f'all layers: pair-GEMM {tg:.1f} us + reduce/stem {tr:.1f} us = {tg + tr:.1f} us for {tb / 1e6:.1f} MB (G+S+W) -> {tot:.0f} GB/s = {tot / peak:.3f} of the measured HBM peak ({peak} GB/s)')
    if out_path:
        json.dump(dict(source=os.path.basename(csv_path), peak_gbs=peak, total_us=tg + tr, gemm_us=tg, reduce_us=tr, bytes=tb,
                       conv_GBs=tot, conv_frac=tot / peak, layers=out), open(out_path, 'w'), indent=1)


if __name__ == '__main__':
    main(*sys.argv[1:4])
