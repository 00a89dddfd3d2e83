"""Dev probe: graph-replay step time of the forward vs the CTA caps of the conv kernels (ir_tune_set)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from instancerefer_b200 import _lib, synthetic
model, dev = bench._forward_setup(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d = bench._resident_dict(synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD), dev)
for pg, rd in ((296, 1184), (148, 1184), (148, 592), (222, 888), (296, 592), (148, 296), (296, 1184)):
    _lib.call('ir_tune_set', pg, rd)
    ms = bench._time_graph(model, d, 60, 5, flush)
    print(f'pairgemm_ctas={pg} reduce_ctas={rd}: {ms * 1e3:.1f} us/step', flush=True)
