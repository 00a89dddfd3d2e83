"""Branch-level GPU timeline of ONE replay of the captured training iteration (train_graph.GraphedTrainStep) on the
bench's configs[2] batch: stamps are one-thread kernels writing %globaltimer on each branch's stream, captured with the
step.  Shows which chain of the four-stream graph is the critical path.
usage: python tools/train_timeline.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from instancerefer_b200 import ops, synthetic
from instancerefer_b200.instancerefer import InstanceRefer
from instancerefer_b200.optim import FlatAdam
from instancerefer_b200.train_graph import GraphedTrainStep

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
ops.check_device(0)
model = InstanceRefer(7, bench.make_args())
model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)
model = model.to(dev).train()
opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
h = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in bench.train_batches(0, n=1)[0].items()}
stepper = GraphedTrainStep(model, opt, synthetic.SyntheticConfig(), depth=1)
stepper(h)['result'].get()                 # first sight: eager
tl = ops.Timeline(dev)
ops.TIMELINE = tl
stepper(h)['result'].get()                 # capture (stamps recorded as graph nodes) + first replay
tl.frozen = True
acc, R = None, 20
for _ in range(R):
    stepper(h)['result'].get()
    torch.cuda.synchronize()
    r = tl.read()
    acc = [x[1] for x in r] if acc is None else [a + x[1] for a, x in zip(acc, r)]
print(f'captured training iteration, mean over {R} replays (us since the first stamp)')
for (l, _), a in sorted(zip(r, acc), key=lambda q: q[1]):
    print(f'  {a / R:8.1f}  {l}')
