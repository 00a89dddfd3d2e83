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
def test_solver_trains_and_resumes(tmp_path, lib_built, state_dict, args):
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
               bn_decay_rate=0.5, output_root=str(tmp_path))
    best = s(3, verbose=2)
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
