"""Branch-level GPU timeline of one CUDA-graph replay of InstanceRefer.forward on the bench workload
(stamps = one-thread kernels writing %globaltimer on each branch's stream).
usage: python tools/timeline.py [encoder_mode 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from instancerefer_b200 import ops, synthetic, _lib

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model, dev = bench._forward_setup(0)
_lib.load().ir_encoder_mode_set(mode)
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
d = bench._resident_dict(b, dev)
for _ in range(3):
    model(dict(d))
torch.cuda.synchronize()
tl = ops.Timeline(dev)
ops.TIMELINE = tl
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    model(dict(d))
tl.frozen = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
acc = None
R = 20
for _ in range(R):
    flush.zero_()
    g.replay()
    torch.cuda.synchronize()
    r = tl.read()
    acc = [x[1] for x in r] if acc is None else [a + x[1] for a, x in zip(acc, r)]
print(f'encoder mode {mode}: mean over {R} replays (us since the first stamp)')
for (l, _), a in sorted(zip(r, acc), key=lambda q: q[1]):
    print(f'  {a / R:8.1f}  {l}')
