"""GPU parity tests of the training step: every backward / loss / optimiser
kernel through the C ABI against a plain torch fp32 CPU reference of the same op, then one whole
training iteration (train-mode forward, get_loss, backward, Adam) against the CPU oracle
(oracle/train_ref.py) and the committed golden fixtures of the reference run verbatim.

Tolerances: index tables bit-exact; forward values / loss terms 1e-4; gradients 1e-3 of the
tensor's scale (see tests/test_oracle.py::check_train_against_golden for why)."""
import glob
import os

import numpy as np
import pytest
import torch

import model_ref
import train_ref
from instancerefer_b200 import synthetic
from test_gpu_parity import rand_coords
from test_oracle import (ZERO_GRAD, assert_grads_agree, check_train_against_golden, load_case,
                         train_golden_cases)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(lib_built):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from instancerefer_b200 import ops as _ops
    _ops.check_device()
    return _ops


def pairs_of_map(cnt, in_idx, slot, n_out):
    """host (k, in, out) triples of a device rulebook."""
    cnt = cnt.cpu().numpy()
    in_idx = in_idx.cpu().numpy()
    slot = slot.cpu().numpy()
    res = []
    for k in range(in_idx.shape[0]):
        o = np.nonzero(slot[k, :n_out] >= 0)[0]
        res.append((in_idx[k, slot[k, o]].astype(np.int64), o.astype(np.int64)))
        assert o.size == cnt[k]
    return res


def conv_ref(x, w, pairs, n_out):
    out = torch.zeros(n_out, w.shape[2])
    for k, (i, o) in enumerate(pairs):
        if i.size:
            out = out.index_add(0, torch.from_numpy(o), x[torch.from_numpy(i)] @ w[k])
    return out


@pytest.fixture(scope='module')
def graph(ops):
    from instancerefer_b200.training import EncoderGraph
    rng = np.random.default_rng(3)
    C0 = rand_coords(rng, 6000, -30, 40, 3)
    ws = ops.EncoderWorkspace(ops.round_rows(C0.shape[0]), 'cuda')
    ops.encoder_build_maps(ws, torch.from_numpy(C0).cuda())
    return EncoderGraph(ws)


@pytest.mark.parametrize('kind,level', [('k3', 0), ('k3', 2), ('k2', 0), ('k2', 2)])
def test_rulebook_transpose_bit_exact(ops, graph, kind, level):
    ii, sl, cnt, n_in, n_out, _, n_out_dev = graph.map(kind, level)
    out_idx, slot_in = ops.rulebook_transpose(ii, sl, n_out_dev, n_out)
    torch.cuda.synchronize()
    out_idx, slot_in = out_idx.cpu().numpy(), slot_in.cpu().numpy()
    ii_h = ii.cpu().numpy()
    for k, (i, o) in enumerate(pairs_of_map(cnt, ii, sl, n_out)):
        c = i.size
        pos = slot_in[k, i]
        assert (pos >= 0).all() and np.array_equal(out_idx[k, pos], o) and np.array_equal(ii_h[k, pos], i)
        mask = np.ones(n_in, bool)
        mask[i] = False
        assert (slot_in[k, :n_in][mask] == -1).all() and c == int(cnt[k])


@pytest.mark.parametrize('kind,level,cin,cout', [('k3', 0, 7, 32), ('k2', 0, 32, 64), ('k3', 1, 64, 64),
                                                 ('k2', 1, 64, 128), ('k3', 2, 128, 128), ('k2', 3, 128, 128)])
def test_spconv_dgrad_wgrad(ops, graph, kind, level, cin, cout):
    ii, sl, cnt, n_in, n_out, n_in_dev, n_out_dev = graph.map(kind, level)
    K = ii.shape[0]
    g = torch.Generator().manual_seed(level * 7 + cin)
    x = torch.randn(n_in, cin, generator=g, requires_grad=True)
    w = (torch.randn(K, cin, cout, generator=g) / (cin * K) ** 0.5).requires_grad_(True)
    dy = torch.randn(n_out, cout, generator=g) * 1e-3          # gradient-like magnitudes
    pairs = pairs_of_map(cnt, ii, sl, n_out)
    conv_ref(x, w, pairs, n_out).backward(dy)
    out_idx, slot_in = graph.transposed(kind, level)
    dW = ops.spconv_wgrad(x.detach().cuda(), dy.cuda(), ii, out_idx, cnt)
    assert float((dW.cpu() - w.grad).abs().max()) < 1e-4 * float(w.grad.abs().max()) + 1e-9
    if cin >= 64:                                   # the same wgrad on tcgen05 (MN-major gathered operands, range-scaled dy)
        for mag in (1.0, 1e-6):
            dys = (dy * mag).cuda()
            dW2 = ops.spconv_wgrad(x.detach().cuda(), dys, ii, out_idx, cnt, dy_absmax=dys.abs().max().reshape(1), use_tc=True)
            err = float((dW2.cpu() - w.grad * mag).abs().max())
            assert err < 1e-4 * float(w.grad.abs().max()) * mag, (mag, err)
    if cin >= 32:
        dx = torch.empty(n_in, cin, device='cuda')
        wt = w.detach().transpose(1, 2).contiguous().cuda()
        ops.spconv_layer(dy.cuda(), out_idx, slot_in, cnt, n_in_dev, n_in, wt, None, None, None, False,
                         graph.ws.T(), dx)
        assert float((dx.cpu() - x.grad).abs().max()) < 1e-4 * float(x.grad.abs().max()) + 1e-9
        # the same dgrad on tcgen05 with the range-scaled split-fp16 operands, at gradient-like magnitudes
        for mag in (1.0, 1e-7):
            dys = (dy * mag).cuda()
            amax = dys.abs().max().reshape(1)
            dx2 = torch.empty(n_in, cin, device='cuda')
            ops.spconv_layer_scaled(dys, amax, out_idx, slot_in, cnt, n_in_dev, n_in, wt, None, graph.ws.T(), dx2)
            err = float((dx2.cpu() - x.grad * mag).abs().max())
            assert err < 1e-4 * float(x.grad.abs().max()) * mag, (mag, err)


@pytest.mark.parametrize('n,C,relu,resid', [(5000, 32, True, False), (777, 128, True, True), (3, 256, True, False),
                                            (4784, 128, False, False), (64, 64, False, True)])
def test_bn_train_fwd_bwd(ops, n, C, relu, resid):
    g = torch.Generator().manual_seed(n + C)
    x = (torch.randn(n, C, generator=g) * 1.7 + 0.4).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.1).requires_grad_(True)
    r = torch.randn(n, C, generator=g).requires_grad_(True) if resid else None
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = torch.nn.functional.batch_norm(x, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    if resid:
        y = y + r
    if relu:
        y = torch.relu(y)
    dy = torch.randn(n, C, generator=g)
    y.backward(dy)
    rm_d, rv_d = rm.cuda(), rv.cuda()
    yd, mean, rstd = ops.bn_train_fwd(x.detach().cuda(), gamma.detach().cuda(), beta.detach().cuda(),
                                      r.detach().cuda() if resid else None, relu, 1e-5, 0.1, rm_d, rv_d)
    assert float((yd.cpu() - y.detach()).abs().max()) < 1e-4
    assert float((rm_d.cpu() - rm_ref).abs().max()) < 1e-5 and float((rv_d.cpu() - rv_ref).abs().max()) < 1e-4
    dx, dres, dg, db = ops.bn_train_bwd(dy.cuda(), yd, x.detach().cuda(), mean, rstd, gamma.detach().cuda(), relu, resid)
    sc = lambda t: float(t.abs().max()) + 1e-6
    assert float((dx.cpu() - x.grad).abs().max()) < 1e-4 * sc(x.grad)
    assert float((dg.cpu() - gamma.grad).abs().max()) < 1e-4 * sc(gamma.grad)
    assert float((db.cpu() - beta.grad).abs().max()) < 1e-4 * sc(beta.grad)
    if resid:
        assert float((dres.cpu() - r.grad).abs().max()) < 1e-5 * sc(r.grad)


def test_segmax_bwd(ops):
    g = torch.Generator().manual_seed(5)
    n, C, nseg = 900, 128, 7
    f = torch.relu(torch.randn(n, C, generator=g)).requires_grad_(True)
    b = torch.sort(torch.randint(0, nseg, (n,), generator=g))[0]
    coords = torch.zeros(n, 4, dtype=torch.int32)
    coords[:, 3] = b.int()
    pooled = torch.stack([f[b == s].max(0)[0] for s in range(nseg)], 0)
    dp = torch.randn(nseg, C, generator=g)
    pooled.backward(dp)
    n_dev = torch.tensor([n], dtype=torch.int32, device='cuda')
    fd, cd = f.detach().cuda(), coords.cuda()
    pd = ops.segmax(fd, cd, n_dev, n, nseg)
    assert torch.equal(pd.cpu(), pooled.detach())
    df = ops.segmax_bwd(fd, cd, n_dev, n, nseg, pd, dp.cuda()).cpu()
    # ties only occur at 0 after the ReLU, where the reference passes no gradient through the ReLU
    # either; compare where the max is positive and check the totals everywhere
    pos = (pooled.detach()[b] > 0)
    assert torch.equal(df[pos], f.grad[pos])
    assert float((df.sum(0) - f.grad.sum(0)).abs().max()) < 1e-5


def test_cross_entropy_and_region_label(ops):
    g = torch.Generator().manual_seed(9)
    z = torch.randn(16, 18, generator=g, requires_grad=True)
    y = torch.randint(0, 18, (16,), generator=g)
    loss = torch.nn.functional.cross_entropy(z, y)
    loss.backward()
    l, dl = ops.cross_entropy(z.detach().cuda(), y.cuda())
    assert abs(float(l) - float(loss)) < 1e-5 and float((dl.cpu() - z.grad).abs().max()) < 1e-6
    pmin = torch.rand(64, 3, generator=g) - 1
    pmax = pmin + torch.rand(64, 3, generator=g) * 8 + 1
    c = pmin + (pmax - pmin) * torch.rand(64, 3, generator=g)
    for cast in (lambda t: t, lambda t: t.double()):
        want = train_ref.scene_region_label(cast(c), cast(pmin), cast(pmax))
        got = ops.region_label(cast(c).cuda(), cast(pmin).cuda(), cast(pmax).cuda())
        assert torch.equal(got.cpu(), want)
    assert len(set(got.cpu().tolist())) == 9


def test_ref_loss_kernel(ops):
    rng = np.random.default_rng(4)
    counts = [5, 1, 0, 40, 2, 3]
    B = len(counts)
    gt = np.concatenate([rng.uniform(0, 5, (B, 3)), rng.uniform(0.4, 1.2, (B, 3)), rng.uniform(-0.3, 0.3, (B, 1))], 1)
    pred = []
    for b, c in enumerate(counts):
        o = np.concatenate([rng.uniform(0, 5, (c, 3)), rng.uniform(0.4, 1.2, (c, 3)), np.zeros((c, 1))], 1)
        if c and b != 4:
            o[rng.integers(0, c)] = gt[b] * np.array([1, 1, 1, 1.05, 0.95, 1, 0])   # one good match (scene 4: none)
        pred.append(o)
    M = sum(c for c in counts if c >= 2)
    s = [torch.randn(M, generator=torch.Generator().manual_seed(i)).requires_grad_(True) for i in range(3)]
    out = dict(lang_scores=torch.zeros(B, 18), seg_scores=torch.zeros(B, 9), pred_obb_batch=pred,
               attribute_scores=s[0], relation_scores=s[1], scene_scores=s[2])

    class Cfg:
        def param2obb_batch(self, *a):
            return gt
    data = dict(object_cat=np.zeros(B, np.int64), ref_center_label=np.zeros((B, 3), np.float32),
                point_min=np.zeros((B, 3), np.float32), point_max=np.ones((B, 3), np.float32),
                ref_heading_class_label=np.zeros(B, np.int64), ref_heading_residual_label=np.zeros(B, np.float32),
                ref_size_class_label=np.zeros(B, np.int64), ref_size_residual_label=np.zeros((B, 3), np.float32))
    L = train_ref.get_loss(out, data, Cfg())
    L['ref_loss'].sum().backward()
    obb_ofs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    score_ofs, a = [], 0
    for c in counts:
        score_ofs.append(a if c >= 2 else -1)
        a += c if c >= 2 else 0
    label, loss_scene, dscore, iou_max = ops.ref_loss(
        torch.from_numpy(np.concatenate(pred, 0)).cuda(), torch.from_numpy(obb_ofs).cuda(), torch.from_numpy(gt).cuda(),
        torch.tensor(score_ofs, dtype=torch.int32).cuda(), *[t.detach().cuda() for t in s])
    assert abs(float(loss_scene.sum()) / B - float(L['ref_loss'])) < 1e-5 * max(1.0, float(L['ref_loss']))
    assert np.array_equal(label.cpu().numpy(), np.concatenate([l for l in L['cluster_label']]))
    assert float((dscore.cpu() / B - s[0].grad).abs().max()) < 1e-5
    assert float(iou_max[2]) == 0.0 and float(iou_max[4]) < 0.2 and float(loss_scene[4]) == 0.0


def test_adam_step(ops):
    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(100003, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-5)
    p, m, v = p0.cuda(), torch.zeros(100003, device='cuda'), torch.zeros(100003, device='cuda')
    for step in range(1, 4):
        gr = torch.randn(100003, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))
        ref.grad = gr.clone()
        opt.step()
        ops.adam_step(p, (gr * 2).cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step, grad_scale=0.5)
    assert float((p.cpu() - ref.detach()).abs().max()) < 2e-6


# ----------------------------------------------------------------------------- dense operators

def _check(fn_gpu, fn_ref, inputs, tol=1e-4, nondiff=()):
    """fn_gpu(*cuda tensors) and fn_ref(*cpu tensors) -> tensor or tuple; compares outputs and the
    gradients of every floating input (not in `nondiff`) under a random upstream gradient."""
    g = torch.Generator().manual_seed(1234)
    cpu = [t.clone().requires_grad_(True) if (torch.is_tensor(t) and t.is_floating_point() and i not in nondiff) else t
           for i, t in enumerate(inputs)]
    gpu = [t.detach().cuda().requires_grad_(t.requires_grad) if torch.is_tensor(t) else t for t in cpu]
    o_ref, o_gpu = fn_ref(*cpu), fn_gpu(*gpu)
    o_ref = o_ref if isinstance(o_ref, tuple) else (o_ref,)
    o_gpu = o_gpu if isinstance(o_gpu, tuple) else (o_gpu,)
    loss_r = loss_g = 0
    for a, b in zip(o_ref, o_gpu):
        assert a.shape == b.shape, (a.shape, b.shape)
        assert float((a.detach() - b.detach().cpu()).abs().max()) < tol * max(1.0, float(a.abs().max())), 'forward'
        w = torch.randn(a.shape, generator=g)
        loss_r = loss_r + (a * w).sum()
        loss_g = loss_g + (b * w.cuda()).sum()
    loss_r.backward()
    loss_g.backward()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(cpu, gpu)):
        if torch.is_tensor(a) and a.requires_grad:
            assert b.grad is not None, i
            err = float((a.grad - b.grad.cpu()).abs().max())
            assert err < tol * max(float(a.grad.abs().max()), 1e-3), (i, err, float(a.grad.abs().max()))


def R(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed + sum(shape))) * scale


@pytest.mark.parametrize('M,K,N,relu', [(5, 300, 256, True), (320, 256, 768, False), (70, 75, 128, True), (1, 128, 9, False)])
def test_linear_fn(ops, M, K, N, relu):
    from instancerefer_b200.training import Linear
    Fn = torch.nn.functional
    _check(lambda x, W, b: Linear.apply(x, W, b, relu),
           lambda x, W, b: torch.relu(Fn.linear(x, W, b)) if relu else Fn.linear(x, W, b),
           [R(M, K), R(N, K, scale=K ** -0.5), R(N, scale=0.1)])


def test_gemm_variants(ops):
    A, B = R(37, 50).cuda(), R(50, 29, seed=1).cuda()
    assert float((ops.gemm(A, B) - A @ B).abs().max()) < 1e-4
    assert float((ops.gemm(A.t().contiguous(), B, ta=True) - A @ B).abs().max()) < 1e-4
    assert float((ops.gemm(A, B.t().contiguous(), tb=True) - A @ B).abs().max()) < 1e-4
    out = torch.ones(37, 29, device='cuda')
    assert float((ops.gemm(A, B, out=out, accumulate=True) - (A @ B + 1)).abs().max()) < 1e-4
    x = R(1000, 70).cuda()
    assert float((ops.colsum(x) - x.sum(0)).abs().max()) < 1e-3


@pytest.mark.parametrize('M,N,relu', [(13, 256, True), (700, 128, True), (3, 128, False)])
def test_layernorm_fn(ops, M, N, relu):
    from instancerefer_b200.training import LayerNormAct
    Fn = torch.nn.functional
    ref = lambda x, g, b: torch.relu(Fn.layer_norm(x, (N,), g, b, 1e-5)) if relu else Fn.layer_norm(x, (N,), g, b, 1e-5)
    _check(lambda x, g, b: LayerNormAct.apply(x, g, b, 1e-5, relu), ref, [R(M, N, scale=2.0), R(N) * 0.2 + 1, R(N, seed=3) * 0.1])


def test_batchnorm_fn(ops):
    from instancerefer_b200.training import BatchNormAct
    bn_g, bn_r = torch.nn.BatchNorm1d(128).cuda().train(), torch.nn.BatchNorm1d(128).train()
    _check(lambda x, g, b: BatchNormAct.apply(x, g, b, bn_g, True),
           lambda x, g, b: torch.relu(torch.nn.functional.batch_norm(x, bn_r.running_mean, bn_r.running_var, g, b, True, 0.1, 1e-5)),
           [R(40, 128, scale=1.5) + 0.3, R(128) * 0.2 + 1, R(128, seed=3) * 0.1])
    assert float((bn_g.running_var.cpu() - bn_r.running_var).abs().max()) < 1e-5 and int(bn_g.num_batches_tracked) == 1


@pytest.mark.parametrize('norm,M,K,N1,N2', [('bn', 5, 256, 256, 256), ('ln', 37, 128, 128, 128), ('bn', 2, 128, 128, 9), ('ln', 700, 128, 256, 256)])
def test_fused_mlp_head(ops, norm, M, K, N1, N2):
    """ir_mlp_head_train_fwd/_bwd (one call per direction) against torch's Linear-norm-ReLU-Linear."""
    from instancerefer_b200.training import MLPHead
    Fn = torch.nn.functional
    nm_g = (torch.nn.BatchNorm1d(N1) if norm == 'bn' else torch.nn.LayerNorm(N1)).cuda().train()
    rm, rv = torch.zeros(N1), torch.ones(N1)

    def ref(x, w1, b1, g, be, w2, b2):
        h = Fn.linear(x, w1, b1)
        h = Fn.batch_norm(h, rm, rv, g, be, True, 0.1, 1e-5) if norm == 'bn' else Fn.layer_norm(h, (N1,), g, be, 1e-5)
        return Fn.linear(torch.relu(h), w2, b2)
    _check(lambda x, w1, b1, g, be, w2, b2: MLPHead.apply(x, w1, b1, g, be, w2, b2, nm_g, 0.0), ref,
           [R(M, K), R(N1, K, scale=K ** -0.5), R(N1, scale=0.1), R(N1) * 0.2 + 1, R(N1, seed=3) * 0.1,
            R(N2, N1, scale=N1 ** -0.5), R(N2, scale=0.1)],
           # BatchNorm over TWO rows: x_hat = +-1 and d x_hat / dx ~ 2 / |x1 - x2|, so the summation order of the GEMM in
           # front of it (torch on the CPU vs the skinny kernel) is amplified in the input gradient
           tol=1e-3 if (norm == 'bn' and M <= 2) else 2e-4, nondiff=(2,) if norm == 'bn' else ())
    if norm == 'bn':
        assert float((nm_g.running_var.cpu() - rv).abs().max()) < 1e-5 and int(nm_g.num_batches_tracked) == 1
    # dropout inside the fused head: the kept fraction and the mask reuse in backward
    x = torch.randn(64, K, device='cuda', requires_grad=True)
    w1, w2 = torch.randn(N1, K, device='cuda') * K ** -0.5, torch.eye(N1, device='cuda')[:min(N1, N2)]
    if N2 == N1:
        y = MLPHead.apply(x, w1, torch.zeros(N1, device='cuda'), torch.ones(N1, device='cuda'), torch.ones(N1, device='cuda'),
                          w2, torch.zeros(N1, device='cuda'), nm_g, 0.15)
        frac = float((y == 0).float().mean())            # beta = 1 keeps most activations positive: zeros are drops
        assert 0.10 < frac < 0.30, frac


def test_dropout_fn(ops):
    from instancerefer_b200.training import Dropout
    x = torch.randn(200, 128, device='cuda', requires_grad=True)
    y = Dropout.apply(x, 0.15)
    keep = (y != 0)
    assert abs(float(keep.float().mean()) - 0.85) < 0.02
    assert torch.allclose(y[keep], x.detach()[keep] / 0.85)
    y.sum().backward()
    assert torch.allclose(x.grad, keep.float() / 0.85)
    assert not torch.equal(Dropout.apply(x, 0.15) != 0, keep)            # a fresh mask per call


@pytest.mark.parametrize('mode', [0, 1])
def test_l2norm_and_match_fn(ops, mode):
    from instancerefer_b200.training import L2Norm, Match
    Fn = torch.nn.functional
    counts = [4, 0, 3, 6]
    seg = torch.tensor(sum([[b] * c for b, c in enumerate(counts)], []), dtype=torch.int32)
    ofs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32)
    a, p = R(13, 128), R(4, 128, seed=2)
    if mode == 0:
        ref = lambda a, p: (Fn.normalize(a, p=2, dim=1) * p[seg.long()]).sum(1)
    else:
        ref = lambda a, p: Fn.cosine_similarity(a, p[seg.long()], dim=1)
    _check(lambda a, p: Match.apply(a, p, seg.cuda(), ofs.cuda(), mode), ref, [a, p])
    _check(lambda x: L2Norm.apply(x), lambda x: Fn.normalize(x, p=2, dim=1), [R(9, 256)])


def test_conv3x3_fn(ops):
    from instancerefer_b200.training import Conv3x3
    ref = lambda x, w, b: torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, b).permute(0, 2, 3, 1)
    _check(lambda x, w, b: Conv3x3.apply(x, w, b), ref, [R(2, 15, 25, 128), R(128, 128, 3, 3, scale=0.03), R(128, scale=0.1)])


def test_bev_fn(ops):
    from instancerefer_b200.training import BEV
    rng = np.random.default_rng(8)
    n, B = 300, 2
    c = np.stack([rng.integers(-2, 17, n) * 16, rng.integers(-1, 27, n) * 16, rng.integers(0, 6, n) * 16,
                  rng.integers(0, B, n)], 1).astype(np.int32)              # some rows outside the crop, duplicates
    coords = torch.from_numpy(c)
    n_dev = torch.tensor([n], dtype=torch.int32)

    def ref(f, kern):
        cc = coords.long()
        keep = ((cc[:, :3] >= 0) & (cc[:, :3] < torch.tensor([240, 400, 80]))).all(1)
        fz = torch.einsum('nc,nco->no', f[keep], kern[cc[keep, 2] // 16])
        flat = cc[keep, 3] * 375 + (cc[keep, 0] // 16) * 25 + cc[keep, 1] // 16
        return torch.zeros(B * 375, 128).index_add(0, flat, fz)
    _check(lambda f, kern: BEV.apply(f, kern, coords.cuda(), n_dev.cuda(), n, B), ref, [R(n, 128), R(5, 128, 128, scale=0.09)])


def test_scene_attention_fn(ops):
    from instancerefer_b200.training import SceneAttention

    def ref(f, q):
        a = torch.softmax(torch.bmm(f, q.unsqueeze(2)).squeeze(2) / 128 ** 0.5, dim=1)
        return (f * a.unsqueeze(2)).sum(1)
    _check(lambda f, q: SceneAttention.apply(f, q)[0], ref, [R(3, 231, 128), R(3, 128, seed=5, scale=3.0)])


def test_token_attention_fn(ops):
    from instancerefer_b200.training import TokenAttention
    lengths = torch.tensor([7, 12, 1, 12])
    B, L = 4, 12
    mask = (torch.arange(L)[None] < lengths[:, None]).float()

    def ref(feats, embed, fcw, fcb):
        a = torch.softmax(feats @ fcw.t() + fcb, dim=1) * mask[:, :, None]
        a = a / a.sum(1, keepdim=True)
        return torch.einsum('blh,ble->hbe', a, embed)
    feats = R(B, L, 256) * mask[:, :, None]
    _check(lambda f, e, w, b: TokenAttention.apply(f, e, lengths.cuda(), w, b)[0], ref,
           [feats, R(B, L, 256, seed=2), R(4, 256, scale=0.3), R(4, scale=0.1)],
           nondiff=(3,))          # d/d fcb is identically zero (softmax shift invariance): noise on both sides


def test_gru_layer_fn(ops):
    from instancerefer_b200.training import GRULayer
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    B, L, H = 5, 9, 128
    lengths = torch.tensor([9, 4, 1, 7, 9])

    def ref(xp, whh, bhh):
        # nn.GRU with identity input weights is not expressible; run the cell recurrence directly
        out = torch.zeros(B, L, 2 * H)
        xp = xp.view(B, L, 2, 3 * H)
        for b in range(B):
            n = int(lengths[b])
            for d in range(2):
                h = torch.zeros(H)
                for t in (range(n - 1, -1, -1) if d else range(n)):
                    h = model_ref.gru_cell(xp[b, t, d], h, whh[d], bhh[d])
                    out[b, t, d * H:(d + 1) * H] = h
        return out.view(B * L, 2 * H)
    _check(lambda xp, whh, bhh: GRULayer.apply(xp, whh, bhh, lengths.cuda(), B, L), ref,
           [R(B * L, 6 * H), R(2, 3 * H, H, scale=H ** -0.5), R(2, 3 * H, scale=0.1)])


def test_edge_ops(ops):
    from instancerefer_b200.training import EdgeConcat, EdgeMax
    S, nq, k, F, ncls = 20, 6, 4, 25, 18
    g = torch.Generator().manual_seed(3)
    x, xyz = torch.randn(S, F, generator=g), torch.randn(S, 3, generator=g)
    qidx = torch.tensor([0, 3, 4, 9, 15, 19], dtype=torch.int32)
    nbr = torch.randint(0, S, (nq, k), generator=g).int()
    nbr[2, 2:] = -1
    valid = (nbr >= 0)
    j = nbr.clamp(min=0).long()
    i = qidx.long()[:, None].expand(-1, k)
    w_in = ops.edge_inputs(x.cuda(), xyz.cuda(), qidx.cuda(), nbr.cuda(), ncls).cpu()
    want = torch.cat([xyz[j] - xyz[i], x[i][..., -ncls:], x[j][..., -ncls:]], -1) * valid[..., None]
    assert torch.equal(w_in.view(nq, k, -1), want)
    _check(lambda w: EdgeConcat.apply(w, x.cuda(), xyz.cuda(), qidx.cuda(), nbr.cuda(), ncls),
           lambda w: (torch.cat([x[i], w.view(nq, k, F), x[j]], -1) * valid[..., None]).view(nq * k, 3 * F), [R(nq * k, F)])
    _check(lambda m: EdgeMax.apply(m, nbr.cuda()),
           lambda m: torch.where(valid[..., None], m.view(nq, k, -1), torch.full((nq, k, 128), float('-inf'))).max(1)[0],
           [R(nq * k, 128)])


def test_encoder_train_matches_torch_replica(ops, lib_built, state_dict, args):
    """Train-mode encoder forward + backward (13 x [conv, BN, (+skip), ReLU]) against a torch-CPU
    replica driven by the SAME device rulebooks and the same upstream gradient: every activation,
    activation gradient and parameter gradient, layer by layer, to 1e-4 of its scale."""
    from instancerefer_b200 import SparseTensor, training as T
    from instancerefer_b200.candidates import CandidatePack, target_classes
    from instancerefer_b200.instancerefer import InstanceRefer
    model = InstanceRefer(7, args)
    model.load_state_dict(state_dict, strict=True)
    model = model.cuda().train()
    b = synthetic.make_batch(41, batch_size=3, num_points=8000, n_inst=12, n_cand=[6, 2, 5], n_tokens=[9, 14, 3])
    dd = synthetic.to_data_dict(b, SparseTensor, 'cuda')
    pack = CandidatePack(dd, target_classes(dd, args), 'cuda')
    net = model.attribute.net
    ws = net.workspace(pack.M * 1024, 'cuda')
    ops.encoder_reset(ws)
    ops.voxelize(pack.points, pack.cand_rows, 0.02, ws)
    acts, orig = [], T.SparseConvBN.apply

    def spy(*a):
        out = orig(*a)
        out.retain_grad()
        acts.append(out)
        return out
    T.SparseConvBN.apply = spy
    os.environ['IR_TRAIN_ENCODER'] = 'layers'
    try:
        f4, G = T.encoder_forward_train(net, ws)
    finally:
        T.SparseConvBN.apply = orig
        del os.environ['IR_TRAIN_ENCODER']
    assert len(acts) == 13
    wgt = torch.randn(f4.shape, generator=torch.Generator().manual_seed(0))
    (f4 * wgt.cuda()).sum().backward()
    torch.cuda.synchronize()
    layers = net._layers()
    cacts = []

    def cbr(x, idx, kind, level, relu=True, resid=None):
        conv, bn = layers[idx]
        ii, sl, cnt, n_in, n_out, _, _ = G.map(kind, level)
        w = conv.kernel.detach().cpu().requires_grad_(True)
        ga, be = bn.weight.detach().cpu().requires_grad_(True), bn.bias.detach().cpu().requires_grad_(True)
        y = torch.nn.functional.batch_norm(conv_ref(x, w, pairs_of_map(cnt, ii, sl, n_out), n_out), None, None,
                                           ga, be, True, 0.1, bn.eps)
        y = y + resid if resid is not None else y
        y = torch.relu(y) if relu else y
        y.retain_grad()
        cacts.append((y, w, ga, be))
        return y
    h = cbr(ws.feat0(7)[:G.n[0]].cpu(), 0, 'k3', 0)
    for s in range(1, 5):
        li = 1 + 3 * (s - 1)
        h = cbr(h, li, 'k2', s - 1)
        h = cbr(cbr(h, li + 1, 'k3', s), li + 2, 'k3', s, True, h)
    (h * wgt).sum().backward()
    rel = lambda d, ref: float((d.cpu() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
    for i, (a, (c, w, ga, be)) in enumerate(zip(acts, cacts)):
        conv, bn = layers[i]
        errs = (rel(a.detach(), c.detach()), rel(a.grad, c.grad), rel(conv.kernel.grad, w.grad),
                rel(bn.weight.grad, ga.grad), rel(bn.bias.grad, be.grad))
        assert max(errs) < 1e-4, (i, errs)


@pytest.mark.parametrize('which,dgrad,fmode', [('attribute', 'simt', 'fused'), ('scene', 'simt', 'fused'), ('attribute', 'tc', 'fused'),
                                               ('scene', 'tc', 'fused'), ('attribute', 'tc', 'graph'), ('scene', 'tc', 'graph')])
def test_fused_encoder_matches_per_layer_path(ops, lib_built, state_dict, args, which, dgrad, fmode):
    """ir_encoder_train_forward/backward (one call per direction; launched eagerly or replayed from CUDA graphs
    in capacity mode) == the 13 per-layer autograd nodes.  The graph mode runs three iterations (eager, capture,
    replay) on different inputs and compares the last."""
    from instancerefer_b200 import SparseTensor, training as T
    from instancerefer_b200.candidates import CandidatePack, target_classes
    from instancerefer_b200.instancerefer import InstanceRefer
    res = {}
    for mode, seed in (('layers', 17), (fmode, 15), (fmode, 16), (fmode, 17)) if fmode == 'graph' else (('layers', 17), (fmode, 17)):
        b = synthetic.make_batch(seed, batch_size=2, num_points=9000, n_inst=10, n_cand=[4, 3], n_tokens=[5, 6])
        if mode == 'layers' or seed != 16 and seed != 17 or fmode != 'graph':
            model = InstanceRefer(7, args)
            model.load_state_dict(state_dict, strict=True)
            model = model.cuda().train()
        else:                                   # graph mode, later iterations: same model object (cached graphs), fresh weights / stats
            model.load_state_dict(state_dict, strict=True)
            model.zero_grad()
        dd = synthetic.to_data_dict(b, SparseTensor, 'cuda')
        os.environ['IR_TRAIN_ENCODER'] = mode
        os.environ['IR_DGRAD'] = os.environ['IR_WGRAD'] = dgrad
        try:
            if which == 'attribute':
                pack = CandidatePack(dd, target_classes(dd, args), 'cuda')
                net = model.attribute.net
                ws = net.workspace(pack.M * 1024, 'cuda')
                ops.encoder_reset(ws)
                ops.voxelize(pack.points, pack.cand_rows, 0.02, ws)
                f4, G = T.encoder_forward_train(net, ws)
            else:
                net = model.scene.net
                F0, C0 = dd['lidar'].F.float().contiguous(), dd['lidar'].C.int().contiguous()
                ws = net.workspace(F0.shape[0], 'cuda')
                f4, G = T.encoder_forward_train(net, ws, F0, C0)
            wgt = torch.randn(f4.shape, generator=torch.Generator().manual_seed(1)).cuda()
            (f4 * wgt).sum().backward()
            torch.cuda.synchronize()
        finally:
            del os.environ['IR_TRAIN_ENCODER'], os.environ['IR_DGRAD'], os.environ['IR_WGRAD']
        res[mode] = (f4.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()},
                     {k: v.clone() for k, v in net.state_dict().items() if 'running' in k or 'tracked' in k})
    # capacity mode partitions the BatchNorm reductions differently: same sums, different fp32 order
    assert torch.equal(res['layers'][0], res[fmode][0]) if fmode != 'graph' else float((res['layers'][0] - res[fmode][0]).abs().max()) < 1e-4
    for k, g in res['layers'][1].items():
        e = float((g - res[fmode][1][k]).abs().max())
        # wgrad sums with float atomics; the tcgen05 dgrad carries ~2^-22 of each layer's largest gradient
        assert e <= (1e-5 if dgrad == 'simt' else 3e-4) * float(g.abs().max()) + 1e-9, (k, e)
    for k, v in res['layers'][2].items():
        assert torch.equal(v, res[fmode][2][k]) if fmode != 'graph' else float((v.float() - res[fmode][2][k].float()).abs().max()) < 1e-5, k


# ----------------------------------------------------------------------------- whole iteration

def make_train_model(state_dict, args):
    from instancerefer_b200.instancerefer import InstanceRefer
    m = InstanceRefer(7, args)
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0                      # parity runs take Dropout as identity on both sides
    return m


def run_train_step(model, batch):
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss
    model.zero_grad()
    dd = model(synthetic.to_data_dict(batch, SparseTensor, 'cuda'))
    dd = get_loss(dd, train_ref.SyntheticConfig())
    dd['loss'].backward()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu())
             for k, p in model.named_parameters()}
    return dd, grads


@pytest.mark.parametrize('path', train_golden_cases())
def test_train_step_matches_golden(lib_built, state_dict, args, path):
    z, b = load_case(path)
    model = make_train_model(state_dict, args)
    dd, grads = run_train_step(model, b)
    check_train_against_golden(z, {k: dd[k].detach().cpu().reshape(-1)[0] for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss')}, grads)
    sd = model.state_dict()
    for k in z.files:
        if k.startswith('bn/'):
            assert np.abs(sd[k[3:]].cpu().numpy() - z[k]).max() < 1e-5, k


def test_train_step_matches_oracle(lib_built, state_dict, args):
    b = synthetic.make_batch(41, batch_size=5, num_points=8000, n_inst=12, n_cand=[6, 2, 1, 5, 0], n_tokens=[9, 14, 3, 20, 1])
    model = make_train_model(state_dict, args)
    dd, grads = run_train_step(model, b)
    r = train_ref.train_step(state_dict, model_ref.data_from_batch(b), args)
    for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss'):
        assert abs(float(dd[k].reshape(-1)[0]) - float(r[k].reshape(-1)[0])) < 1e-4 * max(1.0, abs(float(r[k].reshape(-1)[0]))), k
    for k in ('lang_scores', 'obj_feats', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores'):
        assert float((dd[k].detach().cpu() - r['outputs'][k]).abs().max()) < 1e-4, k
    assert all(np.array_equal(a.cpu().numpy(), np.asarray(c, np.float32)) for a, c in zip(dd['cluster_label'], r['cluster_label']))
    scale = max(float(g.abs().max()) for g in r['grads'].values())
    assert_grads_agree({k: (grads[k], g) for k, g in r['grads'].items()}, scale)
    new = train_ref.updated_running_stats(state_dict, r['bn_stats'])
    sd = model.state_dict()
    for k, v in new.items():
        assert float((sd[k].cpu().float() - v.float()).abs().max()) < 1e-4, k


@pytest.mark.parametrize('drop', [0.0, 0.3])
def test_fused_scene_tail_matches_per_op_path(lib_built, state_dict, args, drop):
    """ir_scene_tail_train_fwd/_bwd (one call per direction) against the same chain as separate autograd nodes
    (IR_TRAIN_SCENE=ops), whole iteration: outputs, every gradient and the running statistics; with Dropout on, both
    paths draw the same counter-based mask."""
    import instancerefer_b200.training as T
    b = synthetic.make_batch(43, batch_size=3, num_points=8000, n_inst=10, n_cand=[4, 1, 6], n_tokens=[7, 12, 2])
    res = {}
    for mode in ('ops', 'fused'):
        os.environ['IR_TRAIN_SCENE'] = mode
        os.environ['IR_TRAIN_ENCODER'] = 'fused'          # eager encoder: bit-identical f4 for both runs
        try:
            torch.manual_seed(5)
            T._dropout_calls[0] = 0
            model = make_train_model(state_dict, args)
            model.scene.vis_emb_fc[3].p = drop
            dd, grads = run_train_step(model, b)
            res[mode] = (dd['scene_scores'].detach().cpu(), dd['seg_scores'].detach().cpu(), float(dd['loss']), grads,
                         {k: v.cpu() for k, v in model.state_dict().items() if 'running' in k or 'tracked' in k})
        finally:
            del os.environ['IR_TRAIN_SCENE'], os.environ['IR_TRAIN_ENCODER']
    a, f = res['ops'], res['fused']
    # same kernels in the same order; what differs run to run is the summation order of the split-K atomics in the
    # small GEMMs (then amplified by BatchNorm over 3 rows and, rarely, a ReLU flip): the project's fp32 tolerances
    assert float((a[0] - f[0]).abs().max()) < 2e-3 and float((a[1] - f[1]).abs().max()) < 2e-3, 'scores'
    assert abs(a[2] - f[2]) < 1e-4 * max(1.0, abs(a[2])), (a[2], f[2])
    scale = max(float(g.abs().max()) for g in a[3].values())
    assert_grads_agree({k: (f[3][k], a[3][k]) for k in a[3]}, scale)
    for k in a[4]:
        assert torch.allclose(a[4][k].float(), f[4][k].float(), atol=1e-5, rtol=1e-4), k


def test_flat_adam_matches_oracle_update(lib_built, state_dict, args):
    """Two iterations with the flat-buffer Adam against oracle forward/backward + adam_update."""
    from instancerefer_b200.optim import FlatAdam
    b = synthetic.make_batch(5, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9])
    model = make_train_model(state_dict, args)
    opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)
    sd = {k: v.clone() for k, v in state_dict.items()}
    st = {}
    for it in range(2):
        opt.zero_grad()
        run_train_step_noreset(model, b)
        opt.step()
        r = train_ref.train_step(sd, model_ref.data_from_batch(b), args)
        params = {k: sd[k] for k in r['grads']}
        sd.update(train_ref.adam_update(params, r['grads'], st, lr=1e-3, weight_decay=1e-5))
        sd.update(train_ref.updated_running_stats(sd, r['bn_stats']))
    torch.cuda.synchronize()
    # Adam normalises the step to ~lr per entry, so entries whose gradient is numerical noise move by
    # up to lr in either direction on both sides: compare the bulk, not the max
    cur = dict(model.named_parameters())
    for k in r['grads']:
        if k in ZERO_GRAD:
            continue
        d = (cur[k].detach().cpu() - sd[k]).abs()
        assert float(d.mean()) < 5e-4 and float(d.max()) < 4.1e-3, (k, float(d.mean()), float(d.max()))


def run_train_step_noreset(model, batch):
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss
    dd = get_loss(model(synthetic.to_data_dict(batch, SparseTensor, 'cuda')), train_ref.SyntheticConfig())
    dd['loss'].backward()
    return dd


def test_get_eval_matches_oracle(lib_built, state_dict, args):
    """Drop-in get_loss + get_eval on the CUDA outputs vs the oracle's (scenes with 0 / 1 / many candidates)."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.eval_helper import get_eval
    from instancerefer_b200.loss_helper import get_loss
    b = synthetic.make_batch(91, batch_size=4, num_points=5000, n_inst=9, n_cand=[5, 2, 1, 0], n_tokens=[11, 4, 8, 5])
    model = make_train_model(state_dict, args).eval()
    dd = synthetic.to_data_dict(b, SparseTensor, 'cuda')
    dd['unique_multiple'] = torch.tensor([1, 0, 0, 1], device='cuda')
    with torch.no_grad():
        dd = get_eval(get_loss(model(dd), train_ref.SyntheticConfig()), train_ref.SyntheticConfig())
    data = model_ref.data_from_batch(b)
    out = model_ref.forward(state_dict, data, args)
    L = train_ref.get_loss(out, data)
    e = train_ref.get_eval(out, data, L['cluster_label'])
    for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss'):
        assert abs(float(dd[k].reshape(-1)[0]) - float(L[k].reshape(-1)[0])) < 1e-4 * max(1.0, abs(float(L[k].reshape(-1)[0]))), k
    assert dd['ref_acc'] == e['ref_acc'] and np.allclose(dd['ref_iou'], e['ref_iou'], atol=1e-12)
    assert dd['ref_iou_rate_0.25'] == e['ref_iou_rate_0.25'] and dd['ref_iou_rate_0.5'] == e['ref_iou_rate_0.5']
    assert float(dd['lang_acc']) == float(e['lang_acc']) and dd['ref_others_mask'] == e['ref_others_mask']
    assert dd['ref_multiple_mask'] == [1, 0, 0, 1]
    for key, want in (('pred_bboxes', e['pred_box_min_max']), ('gt_bboxes', e['gt_box_min_max'])):
        got = np.stack([np.stack([p.min(0), p.max(0)]) for p in dd[key]])
        assert np.abs(got - np.stack(want)).max() < 1e-12, key
    assert all(np.array_equal(a.cpu().numpy(), np.asarray(c, np.float32)) for a, c in zip(dd['cluster_label'], L['cluster_label']))


def test_full_size_directional_derivative(lib_built, state_dict, args):
    """Size-independent property at BASELINE configs[2] size (2 scenes x 40k points, 32 instances): the
    analytic gradient of the whole training step agrees with a central finite difference of the loss along
    a random parameter direction (fp32 forward, piecewise-linear network: 10% tolerance) — catches indexing /
    grid-size errors that small parity cases cannot."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss
    b = synthetic.make_batch(3000, batch_size=2, num_points=40000, n_inst=32, n_cand=[7, 5], n_tokens=20)
    model = make_train_model(state_dict, args)
    cfg = train_ref.SyntheticConfig()

    def loss_only():
        with torch.no_grad():
            dd = get_loss(model(synthetic.to_data_dict(b, SparseTensor, 'cuda')), cfg)
        return float(dd['loss'].double())

    dd, grads = run_train_step(model, b)
    params = dict(model.named_parameters())
    g = torch.Generator().manual_seed(7)
    # direction restricted to the sparse-conv kernels and heads' matrices (the bulk of the 8 M parameters)
    dirs = {k: torch.randn(p.shape, generator=g).cuda() * float(p.detach().abs().mean()) for k, p in params.items() if p.dim() >= 2}
    analytic = sum(float((grads[k].cuda().double() * d.double()).sum()) for k, d in dirs.items())
    eps = 2e-3
    with torch.no_grad():
        for k, d in dirs.items():
            params[k].add_(d, alpha=eps)
        lp = loss_only()
        for k, d in dirs.items():
            params[k].add_(d, alpha=-2 * eps)
        lm = loss_only()
        for k, d in dirs.items():
            params[k].add_(d, alpha=eps)
    numeric = (lp - lm) / (2 * eps)
    assert abs(numeric - analytic) < 0.1 * max(abs(analytic), abs(numeric)) + 1e-3, (numeric, analytic)


def test_module_forwards_dispatch_in_train_mode(lib_built, state_dict, args):
    """The drop-in route that keeps the reference's own InstanceRefer shell calls the four modules one by one
    (models/instancerefer.py:56-70): in train mode each module's forward must take the training path and the
    chain must give the same scores and gradients as the package's own shell."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss
    b = synthetic.make_batch(5, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9])
    res = []
    for chain in (False, True):
        model = make_train_model(state_dict, args)
        dd = synthetic.to_data_dict(b, SparseTensor, 'cuda')
        if chain:
            for m in (model.lang, model.attribute, model.relation, model.scene):
                dd = m(dd)
        else:
            dd = model(dd)
        dd = get_loss(dd, train_ref.SyntheticConfig())
        dd['loss'].backward()
        torch.cuda.synchronize()
        res.append((dd, {k: p.grad.clone() for k, p in model.named_parameters()}))
    # (split-K partial sums are added with float atomics, and the region classifier's BatchNorm over 2 samples
    #  amplifies the last-bit differences: the forward tolerance of the task, not bitwise equality)
    for k in ('attribute_scores', 'relation_scores', 'scene_scores', 'lang_scores', 'seg_scores', 'loss'):
        assert float((res[0][0][k].detach() - res[1][0][k].detach()).abs().max()) < 1e-4, k
    scale = max(float(g.abs().max()) for g in res[0][1].values())
    for k, g in res[0][1].items():
        assert float((g - res[1][1][k]).abs().max()) < 2e-3 * max(float(g.abs().max()), 1e-3 * scale), k


def test_train_step_with_predicted_language_class(lib_built, state_dict):
    """use_gt_lang: False — the candidate filter uses argmax(lang_scores) (models/attribute_module.py:93-97)."""
    from conftest import make_args
    a = make_args(use_gt_lang=False)
    b = synthetic.make_batch(41, batch_size=3, num_points=9000, n_inst=36, n_cand=[2, 2, 2], n_tokens=[9, 14, 3])
    # whatever class the language head predicts must have candidates: every scene holds each of the 18 classes twice
    for sc in b['instance_class']:
        sc[:] = [i % 18 for i in range(len(sc))]
    model = make_train_model(state_dict, a)
    dd, grads = run_train_step(model, b)
    r = train_ref.train_step(state_dict, model_ref.data_from_batch(b), a)
    assert dd['num_filtered_objs'] == r['outputs']['num_filtered_objs']
    for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss'):
        assert abs(float(dd[k].detach().reshape(-1)[0]) - float(r[k].reshape(-1)[0])) < 1e-4 * max(1.0, abs(float(r[k].reshape(-1)[0]))), k
    scale = max(float(g.abs().max()) for g in r['grads'].values())
    assert_grads_agree({k: (grads[k], g) for k, g in r['grads'].items()}, scale)


def test_gradient_accumulation_over_two_backwards(lib_built, state_dict, args):
    """Two forward/backward passes without zeroing (gradient accumulation) must give the sum of the two gradients
    — guards the static buffers of the graph-replayed encoder passes against aliasing with .grad."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss
    bs = [synthetic.make_batch(s_, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9]) for s_ in (5, 6)]
    cfg = train_ref.SyntheticConfig()

    def grads_of(model, batches, zero_between):
        out = []
        model.zero_grad()
        for b in batches:
            if zero_between:
                model.zero_grad()
            get_loss(model(synthetic.to_data_dict(b, SparseTensor, 'cuda')), cfg)['loss'].backward()
            if zero_between:
                out.append({k: p.grad.clone() for k, p in model.named_parameters()})
        torch.cuda.synchronize()
        return out if zero_between else {k: p.grad.clone() for k, p in model.named_parameters()}

    model = make_train_model(state_dict, args)
    for bn in [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]:
        bn.momentum = 0.0                                   # keep the running statistics fixed across the passes
    grads_of(model, bs, True)                                # warm-up: eager pass, then graph capture
    single = grads_of(model, bs, True)
    acc = grads_of(model, bs, False)
    scale = max(float(g.abs().max()) for g in acc.values())
    for k, g in acc.items():
        want = single[0][k] + single[1][k]
        assert float((g - want).abs().max()) < 2e-3 * max(float(want.abs().max()), 1e-3 * scale), k


def test_graphed_train_step_matches_eager_step(lib_built, state_dict, args):
    """train_graph.GraphedTrainStep (forward + get_loss + backward + Adam replayed from one CUDA graph, capacity-mode
    encoders, device-side Adam scalars) against the eager iteration of lib/solver.py:195-205 from IDENTICAL state:
    the optimiser / BatchNorm state of the eager model is copied into the graphed one before every compared step, so
    the comparison sees one step's arithmetic (Adam amplifies rounding noise of near-zero gradients to ±lr, which
    would otherwise accumulate).  Steps 1-2 run eagerly inside the wrapper (first sight of each signature), 3-4 are
    the capture + first replay, 5-8 pure replays with a learning-rate change in between (the scheduler's effect)."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loss_helper import get_loss, stash_host_labels
    from instancerefer_b200.optim import FlatAdam
    from instancerefer_b200.train_graph import GraphedTrainStep
    cfg = train_ref.SyntheticConfig()
    hosts = []
    for s_, nc in ((5, [4, 3, 2, 5]), (6, [2, 5, 3, 3])):
        b = synthetic.make_batch(s_, batch_size=4, num_points=6000, n_inst=8, n_cand=nc, n_tokens=[6, 9, 4, 7])
        hosts.append({k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if isinstance(v, np.ndarray) else v)
                      for k, v in b.items()})
    ma, mb = make_train_model(state_dict, args), make_train_model(state_dict, args)
    oa, ob = FlatAdam(ma, lr=1e-3, weight_decay=1e-5), FlatAdam(mb, lr=1e-3, weight_decay=1e-5)
    stepper = GraphedTrainStep(mb, ob, cfg, min_hits=1)
    bufs = lambda m: [v for k, v in sorted(m.state_dict().items()) if 'running' in k or 'tracked' in k]

    def eager(h):
        d = stash_host_labels(dict(h))
        d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}
        d['lidar'] = SparseTensor(d.pop('lidar_feats'), d.pop('lidar_coords'))
        oa.zero_grad()
        d = get_loss(ma(d), cfg)
        d['loss'].backward()
        oa.step()
        return d

    for it in range(8):
        h = hosts[it % 2]
        if it == 6:
            oa.param_groups[0]['lr'] = ob.param_groups[0]['lr'] = 2.5e-4
        ob.flat.copy_(oa.flat); ob.exp_avg.copy_(oa.exp_avg); ob.exp_avg_sq.copy_(oa.exp_avg_sq)
        ob.step_count, ob.has_state = oa.step_count, list(oa.has_state)
        for x, y in zip(bufs(mb), bufs(ma)):
            x.copy_(y)
        before = oa.flat.clone()
        da, db = eager(h), stepper(h)
        torch.cuda.synchronize()
        print(it, {k: (float(da[k].detach().reshape(-1)[0]), float(db[k].detach().reshape(-1)[0])) for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss')},
              {k: float((da[k] - db[k]).abs().max()) for k in ('attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores', 'lang_scores', 'obj_feats')})
        for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss'):
            # typically 1e-6; BatchNorm1d over the two language rows of this batch can amplify a rounding difference
            va, vb = float(da[k].detach().reshape(-1)[0]), float(db[k].detach().reshape(-1)[0])
            assert abs(va - vb) < 1e-3 * max(1.0, abs(va)), (it, k, va, vb)
        for k in ('attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores', 'lang_scores'):
            assert float((da[k] - db[k]).abs().max()) < 2e-3, (it, k)
        assert all(torch.equal(x, y) for x, y in zip(da['cluster_label'], db['cluster_label']))
        r = db['result'].get()                                # the async read-back of the same scalars
        assert abs(r['loss'] - float(db['loss'].detach().reshape(-1)[0])) < 1e-6 * max(1.0, abs(r['loss'])) and 0 <= r['seg_acc'] <= 1
        ga = {k: oa.grad_views[i].detach().cpu() for i, (k, _) in enumerate(ma.named_parameters())}
        gb = {k: ob.grad_views[i].detach().cpu() for i, (k, _) in enumerate(mb.named_parameters())}
        scale = max(float(g.abs().max()) for g in ga.values())
        try:
            # capacity mode partitions the BatchNorm sums differently, so activations differ in the last bit; the heads'
            # BatchNorm1d over the few scene rows of a batch (x_hat = +-1 for two rows, rstd up to 1/sqrt(eps))
            # amplifies that into percent-level differences of whole branches' gradients in some runs: every tensor
            # must agree loosely (cosine > 0.999), the tight band is asked of the majority only
            assert_grads_agree({k: (gb[k], ga[k]) for k in ga}, scale, tight=1e-2, frac=0.5)
        except AssertionError as e:
            raise AssertionError(f'step {it}: {str(e)[:600]}')
        assert oa.step_count == ob.step_count == it + 1
        # one Adam step from the same state: the bulk of the update must coincide (entries whose gradient is rounding
        # noise move by ~lr in either direction on both sides)
        ua, ub = oa.flat - before, ob.flat - before
        lr = oa.param_groups[0]['lr']
        assert float(ua.abs().max()) <= 4 * lr and float(ub.abs().max()) <= 4 * lr
        assert float((ua - ub).abs().mean()) < 0.02 * lr, (it, float((ua - ub).abs().mean()))
        for x, y in zip(bufs(mb), bufs(ma)):
            assert torch.allclose(x.float(), y.float(), atol=1e-5, rtol=1e-4), it
    assert stepper.eager_steps == 2 and stepper.replays == 6 and len(stepper.cache) == 2


def test_graphed_train_step_draws_fresh_dropout_masks(lib_built, state_dict, args):
    """A replayed step must not repeat the dropout mask baked in at capture: with lr = 0 (weights and moments
    frozen... Adam still moves nothing) the same batch replayed twice gives different losses when Dropout is on,
    identical ones when it is off."""
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.optim import FlatAdam
    from instancerefer_b200.train_graph import GraphedTrainStep
    b = synthetic.make_batch(5, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9])
    h = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    for drop_on in (True, False):
        m = InstanceRefer(7, args)
        m.load_state_dict(state_dict, strict=True)
        m = m.cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                mod.momentum = 0.0
            if isinstance(mod, torch.nn.Dropout) and not drop_on:
                mod.p = 0.0
        opt = FlatAdam(m, lr=0.0)
        stepper = GraphedTrainStep(m, opt, train_ref.SyntheticConfig())
        losses = [float(stepper(h)['loss']) for _ in range(5)][1:]             # first sight of a signature runs eagerly
        assert stepper.replays == 4 and stepper.eager_steps == 1
        if drop_on:
            assert len(set(losses)) == 4, losses
        else:
            assert max(losses) - min(losses) < 1e-4 * abs(losses[0]), losses
