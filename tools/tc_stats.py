"""Per-role cycle counters of CTA 0 of the tcgen05 pair-GEMM (tuning aid)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np, torch
import bench
from instancerefer_b200 import ops, synthetic, _lib
lib = _lib.load()
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
dev = 'cuda'
pts = torch.from_numpy(np.stack(b['instance_points'][0], 0)).to(dev)
cand = torch.arange(32, dtype=torch.int32, device=dev)
ws = ops.EncoderWorkspace(ops.round_rows(32 * 1024), dev)
ops.encoder_reset(ws); ops.voxelize(pts, cand, 0.02, ws); ops.encoder_build_maps(ws)
lvl = 2; cin = cout = 128; K = 27
(in_idx, slot), cnt, n_out = ws.k3(lvl), ws.kcount()[lvl], ws.nlvl()[lvl:lvl + 1]
n = int(ws.nlvl()[lvl]); print('N', n, 'counts', ws.kcount()[lvl][:27].cpu().tolist())
F = torch.randn(n, cin, device=dev); W = torch.randn(K, cin, cout, device=dev) / 60
wprep = ops.spconv_wprep(W); out = torch.empty(ws.n_max, cout, device=dev)
sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
for flags in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '8,15').split(',')]:
    lib.ir_debug_set(flags)
    for _ in range(3):
        ops.spconv_layer(F, in_idx, slot, cnt, n_out, ws.n_max, W, sc, sh, None, True, ws.T(), out, wprep=wprep, use_tc=True)
    torch.cuda.synchronize()
    st = (ctypes.c_longlong * 64)()
    lib.ir_debug_stats(st)
    s = list(st)
    print(f'flags={flags}: kernel {s[48]} cyc')
    for g in range(3):
        print(f'  producer g{g}: items {s[8*g+5]} wdone_wait {s[8*g+4]} wait_empty {s[8*g]} store {s[8*g+1]} fence+arrive {s[8*g+2]} issue_loads {s[8*g+3]}')
    print(f'  mma: items {s[36]} wait_full {s[32]} wait_tempty {s[33]} issue+commit {s[34]} total {s[35]}')
    print(f'  epilogue: weights->tmem {s[40]} wait_tfull {s[41]} work {s[42]} loop {s[43]}')
lib.ir_debug_set(0)
