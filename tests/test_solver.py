"""Solver / checkpoint drop-in: optimizer state in torch.optim.Adam's format, the
reference's schedules and file names; a short training run on the GPU with resume."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import make_args
from instancerefer_b200 import synthetic


def small_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))


def test_flat_adam_state_dict_is_torch_adam_compatible():
    from instancerefer_b200.optim import FlatAdam
    m1, m2 = small_model(), small_model()
    ref = torch.optim.Adam(m2.parameters(), lr=2e-3, weight_decay=1e-5)
    m2(torch.randn(5, 6)).sum().backward()
    ref.step()
    opt = FlatAdam(m1, lr=1e-3)
    opt.load_state_dict(ref.state_dict())                                  # torch -> flat
    assert opt.step_count == 1 and opt.lr == 2e-3 and opt.weight_decay == 1e-5
    for i, (p, o) in enumerate(zip(opt.params, opt.offsets)):
        assert torch.equal(opt.exp_avg[o:o + p.numel()].view_as(p), ref.state_dict()['state'][i]['exp_avg'])
    again = torch.optim.Adam(small_model().parameters())
    again.load_state_dict(opt.state_dict())                                # flat -> torch
    a, b = again.state_dict(), ref.state_dict()
    for i in b['state']:
        assert torch.equal(a['state'][i]['exp_avg_sq'], b['state'][i]['exp_avg_sq']) and float(a['state'][i]['step']) == 1.0
    assert a['param_groups'][0]['lr'] == 2e-3


def test_schedules_match_reference_formulas():
    from instancerefer_b200.solver import Solver
    m = small_model()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    s = Solver(m, None, {}, opt, 'unit', lr_decay_step=[15, 20], lr_decay_rate=0.1, bn_decay_step=20, bn_decay_rate=0.5,
               output_root=os.path.join('/tmp', 'ir_solver_unit'))
    sched = torch.optim.lr_scheduler.MultiStepLR(torch.optim.Adam(small_model().parameters(), lr=1e-3), [15, 20], 0.1)
    for e in range(25):
        assert abs(s._lr_at(e) - sched.get_last_lr()[0]) < 1e-12
        sched.optimizer.step()
        sched.step()
    assert s._bn_momentum_at(0) == 0.5 and s._bn_momentum_at(45) == 0.125 and s._bn_momentum_at(10 ** 4) == 0.001
    s._apply_schedules(21)
    assert abs(opt.param_groups[0]['lr'] - 1e-5) < 1e-15 and m[1].momentum == 0.25


def test_saved_model_loads_into_reference(tmp_path, state_dict, args):
    """model_last.pth written from the drop-in loads STRICT into the reference's own InstanceRefer."""
    import ref_harness as H
    if not H.available():
        pytest.skip('/root/reference not present (GPU box)')
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.solver import Solver
    m = InstanceRefer(7, args)
    m.load_state_dict(state_dict, strict=True)
    s = Solver(m, None, {}, torch.optim.Adam(m.parameters()), 'ckpt', output_root=str(tmp_path))
    s._finish(3)
    ref_model, _ = H.build_reference_model()
    ref_model.load_state_dict(torch.load(tmp_path / 'ckpt' / 'model_last.pth'), strict=True)
    ck = torch.load(tmp_path / 'ckpt' / 'checkpoint.tar', weights_only=False)
    assert ck['epoch'] == 3 and set(ck) == {'epoch', 'model_state_dict', 'optimizer_state_dict'}
    torch.optim.Adam(ref_model.parameters()).load_state_dict(ck['optimizer_state_dict'])


@pytest.mark.gpu
@pytest.mark.parametrize('graph_train', [False, True])
def test_solver_trains_and_resumes(tmp_path, lib_built, state_dict, args, graph_train):
    """graph_train: the same epochs with every training iteration going through train_graph.GraphedTrainStep (first sight
    of a signature eager, then captured / replayed; the BN-momentum schedule changes the signature every epoch)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import train_ref
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.optim import FlatAdam
    from instancerefer_b200.solver import Solver

    def loader(seeds):
        out = []
        for sd in seeds:
            b = synthetic.make_batch(sd, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9])
            d = synthetic.to_data_dict(b, SparseTensor, 'cpu')
            d['unique_multiple'] = torch.zeros(2, dtype=torch.int64)
            out.append(d)
        return out

    class Loader(list):                     # fresh dicts every epoch (the forward mutates its input dict)
        def __iter__(self):
            return iter([dict(d) for d in list.__iter__(self)])

    model = InstanceRefer(7, args)
    model.load_state_dict(state_dict, strict=True)
    model = model.cuda()
    opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)
    dl = {'train': Loader(loader([5, 6, 7])), 'val': Loader(loader([8]))}
    s = Solver(model, train_ref.SyntheticConfig(), dl, opt, 'run', lr_decay_step=[1], lr_decay_rate=0.1, bn_decay_step=1,
               bn_decay_rate=0.5, output_root=str(tmp_path), graph_train=graph_train)
    best = s(3, verbose=2)
    if graph_train:
        assert s._graph_step.replays == 6 and s._graph_step.eager_steps == 3
    root = tmp_path / 'run'
    assert all((root / f).is_file() for f in ('model.pth', 'model_last.pth', 'checkpoint.tar', 'log.txt'))
    tr = [h['loss'] for h in s.history['train']]
    assert tr[-1] < tr[0] and np.isfinite(tr).all() and best['epoch'] >= 1
    assert abs(opt.lr - 1e-4) < 1e-12 and model.scene.cls[1].momentum == 0.125
    fresh = InstanceRefer(7, args)
    fresh.load_state_dict(torch.load(root / 'model_last.pth'), strict=True)
    # resume: a new solver continues at epoch 3 with the saved Adam moments
    model2 = InstanceRefer(7, args).cuda()
    opt2 = FlatAdam(model2, lr=1e-3, weight_decay=1e-5)
    s2 = Solver(model2, train_ref.SyntheticConfig(), dl, opt2, 'run', output_root=str(tmp_path))
    assert s2.resume() == 3 and opt2.step_count == opt.step_count == 9
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.flat, opt.flat)
    s2(4, verbose=0)
    assert len(s2.history['train']) == 1 and np.isfinite(s2.history['train'][0]['loss'])


def test_flat_adam_is_a_torch_optimizer_under_the_reference_scheduler():
    """lib/solver.py:119-123 builds MultiStepLR(optimizer, [15, 20], 0.1) on whatever scripts/train.py:93 made;
    FlatAdam must be accepted there and the lr the scheduler writes into the param group is the lr ``step`` reads."""
    from instancerefer_b200.optim import FlatAdam
    m = small_model()
    opt = FlatAdam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 1
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [15, 20], 0.1)
    ref = torch.optim.lr_scheduler.MultiStepLR(torch.optim.Adam(small_model().parameters(), lr=1e-3), [15, 20], 0.1)
    for e in range(25):
        assert abs(opt.param_groups[0]['lr'] - ref.get_last_lr()[0]) < 1e-15 and opt.lr == opt.param_groups[0]['lr']
        opt._opt_called = True                      # no CUDA here: the schedulers only need the step order flag
        ref.optimizer.step()
        sched.step()
        ref.step()
    assert abs(opt.lr - 1e-5) < 1e-15
    opt.lr = 3e-4                                   # attribute form used by instancerefer_b200.solver
    assert opt.param_groups[0]['lr'] == 3e-4
    # a parameter that never received a gradient owns no state (torch.optim.Adam semantics)
    sd = opt.state_dict()
    assert sd['state'] == {} and sd['param_groups'][0]['params'] == list(range(len(opt.params)))
    with pytest.raises(ValueError):
        FlatAdam([dict(params=list(small_model().parameters()))])


@pytest.mark.gpu
def test_reference_shaped_solver_loop_matches_torch_adam(lib_built):
    """zero_grad / backward / step / scheduler.step exactly as lib/solver.py:200-205 + :119-125 drive an optimizer,
    FlatAdam against torch.optim.Adam on the same model (one parameter never receives a gradient)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from instancerefer_b200.optim import FlatAdam

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(3)
            self.a, self.b = torch.nn.Linear(6, 8), torch.nn.Linear(8, 3)
            self.unused = torch.nn.Linear(4, 4)     # in parameters(), never in the graph: grad stays None

        def forward(self, x):
            return self.b(torch.relu(self.a(x)))

    m1, m2 = Net().cuda(), Net().cuda()
    o1 = FlatAdam(m1.parameters(), lr=1e-2, weight_decay=1e-3)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-2, weight_decay=1e-3)
    s1 = torch.optim.lr_scheduler.MultiStepLR(o1, [2, 4], 0.1)
    s2 = torch.optim.lr_scheduler.MultiStepLR(o2, [2, 4], 0.1)
    unused0 = m1.unused.weight.detach().clone()
    g = torch.Generator().manual_seed(0)
    for epoch in range(6):
        for _ in range(3):
            x = torch.randn(16, 6, generator=g).cuda()
            for m, o in ((m1, o1), (m2, o2)):
                o.zero_grad()
                m(x).square().mean().backward()
                o.step()
        s1.step()
        s2.step()
        assert abs(o1.param_groups[0]['lr'] - o2.param_groups[0]['lr']) < 1e-15
    for (n, p), q in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), n
    assert torch.equal(m1.unused.weight, unused0)               # no weight decay / moment update without a gradient
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert set(sd1['state']) == set(sd2['state']) == {0, 1, 2, 3}
    for i in sd2['state']:
        assert torch.allclose(sd1['state'][i]['exp_avg_sq'].cpu(), sd2['state'][i]['exp_avg_sq'].cpu(), rtol=1e-4, atol=1e-10)


@pytest.mark.gpu
def test_eval_after_flat_adam_steps_uses_the_updated_weights(lib_built, state_dict, args):
    """FlatAdam.step writes the parameters through a raw-pointer kernel (no tensor version bump): the eval-mode
    prepared copies (folded BN, repacked / transposed weights) and captured graphs must not survive it.  Train two
    steps, evaluate, and compare with a freshly built model loaded from the trained state_dict."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import train_ref
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.graphed import GraphedInstanceRefer
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.loss_helper import get_loss

    def batch(seed, dev='cuda'):
        b = synthetic.make_batch(seed, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9])
        return synthetic.to_data_dict(b, SparseTensor, dev)

    keys = ('lang_scores', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores')
    model = InstanceRefer(7, args)
    model.load_state_dict(state_dict, strict=True)
    model = model.cuda().eval()
    before = {k: model(batch(11))[k].clone() for k in keys}     # fills every eval-mode cache with the initial weights
    runner = GraphedInstanceRefer(model)
    g0 = {k: v.clone() for k, v in runner(batch(11, 'cpu'))['host_scores'].items()}
    from instancerefer_b200.optim import FlatAdam
    opt = FlatAdam(model, lr=1e-2)
    model.train()
    for seed in (5, 6):
        opt.zero_grad()
        get_loss(model(batch(seed)), train_ref.SyntheticConfig())['loss'].backward()
        opt.step()
    model.eval()
    after = {k: model(batch(11))[k].clone() for k in keys}
    fresh = InstanceRefer(7, args)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, strict=True)
    fresh = fresh.cuda().eval()
    want = {k: fresh(batch(11))[k] for k in keys}
    for k in keys:
        assert float((after[k] - want[k]).abs().max()) < 1e-5, k
    assert any(float((after[k] - before[k]).abs().max()) > 1e-3 for k in keys)      # the weights did move
    g1 = runner(batch(11, 'cpu'))['host_scores']                                    # graphs re-captured, not replayed stale
    for k in ('attribute_scores', 'relation_scores', 'scene_scores'):
        assert float((g1[k] - want[k].cpu()).abs().max()) < 1e-5, k
        assert float((g1[k] - g0[k]).abs().max()) > 1e-4, k


def test_flat_adam_buckets_follow_gradient_availability(args):
    """All-reduce buckets of the data-parallel step are cut where gradients become final together: per sub-module,
    each sparse encoder split into deep (stage3/4, finished first, handed over early) and shallow stages; a
    plain parameter list falls back to size-based ranges.  Buckets tile the flat buffer without gaps."""
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.optim import FlatAdam
    model = InstanceRefer(7, args)
    names = [n for n, _ in model.named_parameters()]
    opt = FlatAdam(model, lr=1e-3)
    spans = [(names[lo], names[hi - 1]) for lo, hi, _, _ in opt.buckets]
    assert opt.buckets[0][2] == 0 and opt.buckets[-1][3] == opt.numel
    assert all(a[3] == b[2] and a[1] == b[0] for a, b in zip(opt.buckets, opt.buckets[1:]))
    firsts = [s[0] for s in spans]
    assert firsts[0].startswith('lang.') and any(f.startswith('attribute.net.stem') for f in firsts)
    for mod in ('attribute', 'scene'):
        b = [i for i, (lo, hi, _, _) in enumerate(opt.buckets) if names[lo].startswith(f'{mod}.net.stage3.0')]
        assert len(b) == 1, spans
        lo, hi, _, _ = opt.buckets[b[0]]
        assert all(n.startswith((f'{mod}.net.stage3', f'{mod}.net.stage4')) for n in names[lo:hi]) and hi - lo == 18
    assert 5 <= opt.n_buckets <= 8, spans
    by_size = FlatAdam(list(InstanceRefer(7, args).parameters()), lr=1e-3)
    assert by_size.n_buckets == 4 and by_size.buckets[-1][3] == by_size.numel
