"""Dev aid: encoder train-mode forward/backward on the GPU vs a torch-CPU replica that uses the SAME
device rulebooks; reports the first activation / gradient that disagrees."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import __graft_entry__ as g
g.build()
import weights
from conftest import make_args
from instancerefer_b200 import ops, synthetic, training as T
from instancerefer_b200.instancerefer import InstanceRefer
from instancerefer_b200.candidates import CandidatePack, target_classes
from instancerefer_b200 import SparseTensor
from test_gpu_train import pairs_of_map, conv_ref

args = make_args()
sd = weights.make_state_dict(123)
model = InstanceRefer(7, args); model.load_state_dict(sd); model = model.cuda().train()
b = synthetic.make_batch(41, batch_size=4, num_points=8000, n_inst=12, n_cand=[6, 2, 1, 5], n_tokens=[9, 14, 3, 20])
dd = synthetic.to_data_dict(b, SparseTensor, 'cuda')
pack = CandidatePack(dd, target_classes(dd, args), 'cuda')
net = model.attribute.net
ws = net.workspace(pack.M * 1024, 'cuda')
ops.encoder_reset(ws); ops.voxelize(pack.points, pack.cand_rows, 0.02, ws)

acts = []
orig_apply = T.SparseConvBN.apply
def spy(*a):
    out = orig_apply(*a); out.retain_grad(); acts.append(out); return out
T.SparseConvBN.apply = spy
f4, G = T.encoder_forward_train(net, ws)
gen = torch.Generator().manual_seed(0)
wgt = torch.randn(f4.shape, generator=gen)
(f4 * wgt.cuda()).sum().backward()
torch.cuda.synchronize()
print('rows per level', G.n, 'pairs k3', [int(G.kcount[l].sum()) for l in range(5)], 'k2', [int(G.kcount[5 + l].sum()) for l in range(4)])

# CPU replica
layers = net._layers()
x = ws.feat0(7)[:G.n[0]].cpu()
cacts = []
def cbr(x, idx, kind, level, relu=True, resid=None):
    conv, bn = layers[idx]
    ii, sl, cnt, n_in, n_out, _, _ = G.map(kind, level)
    w = conv.kernel.detach().cpu().requires_grad_(True)
    ga, be = bn.weight.detach().cpu().requires_grad_(True), bn.bias.detach().cpu().requires_grad_(True)
    y = conv_ref(x, w, pairs_of_map(cnt, ii, sl, n_out), n_out)
    y = torch.nn.functional.batch_norm(y, None, None, ga, be, True, 0.1, bn.eps)
    if resid is not None: y = y + resid
    if relu: y = torch.relu(y)
    y.retain_grad(); cacts.append((y, w, ga, be)); return y
h = cbr(x, 0, 'k3', 0)
for s in range(1, 5):
    li = 1 + 3 * (s - 1)
    h = cbr(h, li, 'k2', s - 1); yy = cbr(h, li + 1, 'k3', s); h = cbr(yy, li + 2, 'k3', s, True, h)
(h * wgt).sum().backward()
for i, (a, (c, w, ga, be)) in enumerate(zip(acts, cacts)):
    conv, bn = layers[i]
    r = lambda d, ref: float((d.cpu() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    print('layer %2d  act %.1e  dact %.1e  dW %.1e  dgamma %.1e dbeta %.1e' % (
        i, r(a.detach(), c.detach()), r(a.grad, c.grad), r(conv.kernel.grad, w.grad), r(bn.weight.grad, ga.grad), r(bn.bias.grad, be.grad)))
