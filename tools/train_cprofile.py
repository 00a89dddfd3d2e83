"""Dev aid: cProfile of the host side of the training step (top cumulative / self times)."""
import os, sys, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from instancerefer_b200 import SparseTensor, synthetic
from instancerefer_b200.instancerefer import InstanceRefer
from instancerefer_b200.loss_helper import get_loss, stash_host_labels
from instancerefer_b200.optim import FlatAdam

model = InstanceRefer(7, bench.make_args()); model.load_state_dict(synthetic.make_state_dict(123, model=model)); model = model.cuda().train()
opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)
cfg = synthetic.SyntheticConfig()
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
hosts = [{k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in bench.train_batches(0)]

def step(i):
    h = hosts[i % 4]
    d = stash_host_labels(dict(h))
    d = {k: (v.to('cuda', non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
    d['lidar'] = SparseTensor(d.pop('lidar_feats'), d.pop('lidar_coords'))
    opt.zero_grad()
    d = get_loss(model(d), cfg)
    d['loss'].backward()
    opt.step()
    return float(d['loss'].detach())

for i in range(12): step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(40): step(i)
torch.cuda.synchronize()
pr.disable()
for key in ('cumulative', 'tottime'):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
