"""Isolated timing of the sparse-conv kernels (pair-GEMM + reduce) on the bench workload's rulebooks.
usage: python tools/bench_spconv.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import bench
from instancerefer_b200 import ops, synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = sys.argv[2] if len(sys.argv) > 2 else None
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
dev = 'cuda'
cases = {}
# instance encoder maps (voxelised candidates) and scene maps
pts = torch.from_numpy(np.stack(b['instance_points'][0], 0)).to(dev)
cand = torch.arange(32, dtype=torch.int32, device=dev)
ws_i = ops.EncoderWorkspace(ops.round_rows(32 * 1024), dev)
ops.encoder_reset(ws_i); ops.voxelize(pts, cand, 0.02, ws_i); ops.encoder_build_maps(ws_i)
ws_s = ops.EncoderWorkspace(ops.round_rows(b['lidar_coords'].shape[0]), dev)
ops.encoder_build_maps(ws_s, torch.from_numpy(b['lidar_coords']).to(dev))
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ch = [32, 64, 128, 128, 128]
rows = []
for name, ws in (('inst', ws_i), ('scene', ws_s)):
    nl = ws.nlvl().cpu().tolist(); kc = ws.kcount().cpu().numpy()
    for lvl in range(1, 5):
        for kind in ('k2', 'k3'):
            if kind == 'k2':
                cin, cout, K = ch[lvl - 1], ch[lvl], 8
                (in_idx, slot), cnt, n_out, n_in, P = ws.k2(lvl - 1), ws.kcount()[5 + lvl - 1], ws.nlvl()[lvl:lvl + 1], nl[lvl - 1], int(kc[5 + lvl - 1][:8].sum())
            else:
                cin, cout, K = ch[lvl], ch[lvl], 27
                (in_idx, slot), cnt, n_out, n_in, P = ws.k3(lvl), ws.kcount()[lvl], ws.nlvl()[lvl:lvl + 1], nl[lvl], int(kc[lvl][:27].sum())
            tag = f'{name}.L{lvl}.{kind} {cin}->{cout} N={nl[lvl]} P={P}'
            if only and only not in tag: continue
            F = torch.randn(max(n_in, 1), cin, device=dev)
            W = torch.randn(K, cin, cout, device=dev) / (cin * K) ** 0.5
            wprep = ops.spconv_wprep(W)
            sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
            out = torch.empty(ws.n_max, cout, device=dev)
            T = ws.T()
            from instancerefer_b200 import _lib
            import ctypes
            lib = _lib.load()
            res = {}
            dbgs = [int(x) for x in os.environ.get('IR_DBG', '0').split(',')]
            for mode, use_tc, dbg in [('tc', True, d) for d in dbgs] + [('simt', False, 0)]:
                lib.ir_debug_set(dbg)
                for cold in ((False, True) if dbg == 0 else (False,)):
                    lib.ir_profile_enable(1)
                    for r in range(reps + 3):
                        if cold: flush.zero_()
                        ops.spconv_layer(F, in_idx, slot, cnt, n_out, ws.n_max, W, sc, sh, None, True, T, out, wprep=wprep, use_tc=use_tc)
                    torch.cuda.synchronize()
                    cap = 512
                    gm, rm = (ctypes.c_float * cap)(), (ctypes.c_float * cap)()
                    meta, nout = (ctypes.c_int32 * (4 * cap))(), ctypes.c_int32(0)
                    _lib.call('ir_profile_read', gm, rm, meta, cap, ctypes.byref(nout))
                    lib.ir_profile_enable(0)
                    g = np.median([gm[i] for i in range(3, nout.value)]) * 1e3
                    rd = np.median([rm[i] for i in range(3, nout.value)]) * 1e3
                    res[(mode if dbg == 0 else f'tc{dbg}', cold)] = (g, rd)
            by = P * (4 * cin + 4) + P * 4 * cout + 4 * K * cin * cout
            g, rd = res[('tc', False)]
            print(f'{tag:52s} tc warm gemm {g:6.1f} us ({by / g / 1e3:6.0f} GB/s, {2 * P * cin * cout / g / 1e6:5.1f} TF) reduce {rd:5.1f} | '
                  f'tc cold {res[("tc", True)][0]:6.1f}/{res[("tc", True)][1]:5.1f} | simt warm {res[("simt", False)][0]:6.1f} | ' +
                  ' '.join(f'{k[0]}={v[0]:.1f}' for k, v in res.items() if k[0].startswith('tc') and k[0] != 'tc'), flush=True)
            lib.ir_debug_set(0)
