"""Persistent encoder kernel vs per-layer launches on the bench workload's two encoders: output agreement, graph-replay
timing of the feature pass, and the per-phase time stamps the persistent kernel leaves in its sync area.
usage: python tools/bench_encoder.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from instancerefer_b200 import ops, synthetic, _lib
from instancerefer_b200.instancerefer import InstanceRefer

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = 'cuda'
lib = _lib.load()
model = InstanceRefer(7, bench.make_args())
model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)
model = model.to(dev).eval()
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
pts = torch.from_numpy(np.stack(b['instance_points'][0], 0)).to(dev)
cand = torch.arange(32, dtype=torch.int32, device=dev)
ws_i = ops.EncoderWorkspace(ops.round_rows(32 * 1024), dev)
ops.encoder_reset(ws_i); ops.voxelize(pts, cand, 0.02, ws_i); ops.encoder_build_maps(ws_i)
lc = torch.from_numpy(b['lidar_coords']).to(dev); lf = torch.from_numpy(b['lidar_feats']).to(dev)
ws_s = ops.EncoderWorkspace(ops.round_rows(lc.shape[0]), dev)
ops.encoder_build_maps(ws_s, lc)
torch.cuda.synchronize()
pa, ps = model.attribute.net.prepared()['params'], model.scene.net.prepared()['params']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(mode, which):
    lib.ir_encoder_mode_set(mode)
    oa = torch.zeros(ws_i.n_max, 128, device=dev); os_ = torch.zeros(ws_s.n_max, 128, device=dev)
    def go():
        if which == 'pair':
            ops.encoder_features_pair(pa, ws_i, None, oa, ps, ws_s, lf, os_)
        elif which == 'inst':
            ops.encoder_features(pa, ws_i, None, oa)
        else:
            ops.encoder_features(ps, ws_s, lf, os_)
    go(); go(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        go()
    ts = []
    for cold in (False, True):
        ms = 0.0
        for _ in range(reps):
            if cold: flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ts.append(ms / reps * 1e3)
    n_i, n_s = int(ws_i.nlvl()[4]), int(ws_s.nlvl()[4])
    return ts, oa[:n_i].clone(), os_[:n_s].clone()

def stamps(ws):
    L = ws.layout
    raw = ws.buf[L.off_sync:L.off_sync + 512].view(torch.int64).cpu().numpy()
    st = raw[32:64]
    t0 = st[31]
    return [(int(st[p]) - int(t0)) / 1e3 for p in range(25)]

if os.environ.get('IR_TMA', '0') == '1':
    for which in ('inst', 'scene'):
        outs = {}
        for gm in ('ldg', 'tma'):
            ops.set_gather_mode(gm)
            ts, oa, os_ = run(0, which)
            outs[gm] = oa if which == 'inst' else os_
            print(f'{which:5s} layers gather={gm}  warm {ts[0]:7.1f} us  cold {ts[1]:7.1f} us', flush=True)
        print(f'   max|tma-ldg| = {float((outs["tma"] - outs["ldg"]).abs().max()):.3e}  (max|out| {float(outs["ldg"].abs().max()):.3f})', flush=True)
    ops.set_gather_mode('ldg')
    sys.exit(0)

res = {}
for which in ('inst', 'scene', 'pair'):
    for mode in (0, 1):
        ts, oa, os_ = run(mode, which)
        res[(which, mode)] = (ts, oa, os_)
        print(f'{which:5s} mode={"persist" if mode else "layers "}  warm {ts[0]:7.1f} us  cold {ts[1]:7.1f} us', flush=True)
        if mode == 1:
            ws = ws_s if which == 'scene' else ws_i
            s = stamps(ws)
            print('   phase end times (us since first ticket):', ' '.join(f'{x:.1f}' for x in s), flush=True)
    a, bb = res[(which, 0)], res[(which, 1)]
    if which in ('inst', 'pair'):
        print(f'   inst  out max|persist-layers| = {float((a[1] - bb[1]).abs().max()):.3e}  (max|out| {float(a[1].abs().max()):.3f})')
    if which in ('scene', 'pair'):
        print(f'   scene out max|persist-layers| = {float((a[2] - bb[2]).abs().max()):.3e}  (max|out| {float(a[2].abs().max()):.3f})')

# ---- per-item stamps of one persistent launch (instance encoder)
if os.environ.get('IR_ITEMS', '1') != '0':
    import ctypes
    dbg = torch.zeros(8 * 8192, dtype=torch.int64, device=dev)
    lib.ir_encoder_mode_set(1)
    oa = torch.zeros(ws_i.n_max, 128, device=dev)
    lib.ir_encoder_persist_debug(ctypes.c_void_p(dbg.data_ptr()))
    ops.encoder_features(pa, ws_i, None, oa)
    torch.cuda.synchronize()
    lib.ir_encoder_persist_debug(None)
    d = dbg.cpu().numpy().reshape(-1, 8)
    np.save(os.path.join(ROOT, 'gpurun_out', 'persist_items.npy'), d[d[:, 0] > 0])
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    ph = (d[:, 7] >> 32).astype(int)
    cta = (d[:, 7] & 0xffffffff).astype(int)
    print('occupancy (runtime):', lib.ir_encoder_persist_occupancy(), ' distinct CTAs with items:', len(set(cta.tolist())),
          ' max items of one CTA in a phase:', max(np.bincount(cta[ph == p]).max() for p in range(25) if (ph == p).any()))
    print('per-phase item stats (us): n, first decode, [median: weights, wait-exit, work-end, cta-done, published] last published')
    for p in range(25):
        m = d[ph == p]
        if not len(m): continue
        rel = lambda c: (np.median(m[:, c][m[:, c] > 0]) - t0) / 1e3 if (m[:, c] > 0).any() else float('nan')
        dur = lambda a, b: np.median((m[:, b] - m[:, a])[(m[:, a] > 0) & (m[:, b] > 0)]) / 1e3 if ((m[:, a] > 0) & (m[:, b] > 0)).any() else float('nan')
        print(f'  ph {p:2d} n={len(m):4d} first {(m[:, 0].min() - t0) / 1e3:7.1f} | dec->wts {dur(0, 1):5.1f} wts->dep {dur(1, 2):5.1f} dec->dep {dur(0, 2):5.1f} '
              f'dep->work {dur(2, 3):5.1f} work->cta {dur(3, 4):5.1f} cta->pub {dur(4, 5):5.1f} | dep-exit med {rel(2):7.1f} max {(m[:, 2].max() - t0) / 1e3:7.1f} '
              f'work-end med {rel(3):7.1f} max {(m[:, 3].max() - t0) / 1e3:7.1f} pub max {(m[:, 5].max() - t0) / 1e3:7.1f}')
