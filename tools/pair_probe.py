"""Dev probe: graph-replay step time with the two encoders on separate chains vs shared launches (pair mode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from instancerefer_b200 import synthetic
model, dev = bench._forward_setup(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d = bench._resident_dict(synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD), dev)
for pair in (False, True, False, True):
    model.pair_encoders = pair
    ms = bench._time_graph(model, d, 50, 5, flush)
    print(f'pair_encoders={pair}: {ms * 1e3:.1f} us/step', flush=True)
