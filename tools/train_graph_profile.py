"""ncu launch list of the captured training iteration (train_graph.GraphedTrainStep) on the bench's configs[2] batch:
three replays between profiler start/stop.
usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
           python tools/train_graph_profile.py"""
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
for _ in range(4):                         # eager, capture + replay, two more replays
    stepper(h)['result'].get()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(3):
    stepper(h)['result'].get()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('replays', stepper.replays, 'launches per replay (this library)', stepper.launches_replayed // stepper.replays)
