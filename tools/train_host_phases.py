"""Dev aid: host-side wall time of each phase of the training step (no profiler), vs GPU event time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from instancerefer_b200 import SparseTensor, synthetic, training as T
from instancerefer_b200.candidates import CandidatePack, target_classes
from instancerefer_b200.instancerefer import InstanceRefer
from instancerefer_b200.loss_helper import get_loss
from instancerefer_b200.optim import FlatAdam

model = InstanceRefer(7, bench.make_args()); model.load_state_dict(synthetic.make_state_dict(123, model=model)); model = model.cuda().train()
opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)
cfg = synthetic.SyntheticConfig()
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
hosts = [{k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in bench.train_batches(0)]
acc = {}
def tick(name, t0):
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()

def step(i, rec):
    a = model.args
    t = time.perf_counter()
    h = hosts[i % 4]
    d = {k: (v.to('cuda', non_blocking=True) if torch.is_tensor(v) else v) for k, v in h.items()}
    d['lidar'] = SparseTensor(d.pop('lidar_feats'), d.pop('lidar_coords'))
    opt.zero_grad()
    if rec: t = tick('h2d+zero', t)
    pack = CandidatePack(d, target_classes(d, a), 'cuda')
    if rec: t = tick('candidate pack', t)
    pa, ps = T.prepare_encoder_maps(model, d, pack)
    if rec: t = tick('maps + sync', t)
    d = T.lang_forward_train(model.lang, d)
    d['_ir_candidates'] = pack
    if rec: t = tick('lang', t)
    d = T.attribute_forward_train(model.attribute, d, pack, pa)
    if rec: t = tick('attribute', t)
    d = T.relation_forward_train(model.relation, d, pack)
    if rec: t = tick('relation', t)
    d = T.scene_forward_train(model.scene, d, pack, ps)
    if rec: t = tick('scene', t)
    d = get_loss(d, cfg)
    if rec: t = tick('loss', t)
    d['loss'].backward()
    if rec: t = tick('backward', t)
    opt.step()
    if rec: t = tick('optim', t)
    l = float(d['loss'].detach())
    if rec: t = tick('final sync', t)

for i in range(5): step(i, False)
torch.cuda.synchronize()
N = 30
t0 = time.perf_counter()
for i in range(N): step(i, True)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / N * 1e3
print(f'wall {tot:.2f} ms/step')
for k, v in acc.items(): print(f'  {k:16s} {v / N * 1e3:6.2f} ms')
