"""GPU parity tests (B200): every CUDA op through the C ABI against the CPU oracle on the same seeded
inputs.  Integer / index work is compared bit-exactly (rulebooks after canonicalisation inside each
kernel offset); fp32 values within 1e-4 abs (BASELINE.json north_star)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import model_ref
import sparse_ref as R
from instancerefer_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def ops(lib_built):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from instancerefer_b200 import ops as _ops
    _ops.check_device()
    return _ops


def rand_coords(rng, n, lo, hi, nb):
    pts = rng.integers(lo, hi, (n * 2, 4))
    pts[:, 3] = rng.integers(0, nb, n * 2)
    pts = pts[np.sort(np.unique(pts, axis=0, return_index=True)[1])][:n]
    pts = pts[np.argsort(pts[:, 3], kind='stable')]            # batch ids sorted like a collated batch
    return pts.astype(np.int32)


def rulebook_from_ws(cnt, in_idx, slot, n_out, K):
    """-> per k: sorted array of (out,in) pairs reconstructed from count / in_idx / slot tables."""
    out = []
    cnt = cnt.cpu().numpy()
    in_idx = in_idx.cpu().numpy()
    slot = slot.cpu().numpy()[:, :n_out]
    for k in range(K):
        o = np.nonzero(slot[k] >= 0)[0]
        pos = slot[k, o]
        assert o.size == cnt[k] and (np.sort(pos) == np.arange(cnt[k])).all()
        out.append(np.stack([o, in_idx[k, pos]], 1))
    return out


def rulebook_from_oracle(ii, oo, kofs):
    out = []
    for k in range(len(kofs) - 1):
        seg = slice(kofs[k], kofs[k + 1])
        a = np.stack([oo[seg], ii[seg]], 1)
        out.append(a[np.argsort(a[:, 0], kind='stable')])
    return out


# ----------------------------------------------------------------------------- coordinates

@pytest.mark.parametrize('n,lo,hi,nb', [(1, 0, 4, 1), (700, -20, 20, 2), (30000, -40, 60, 3)])
def test_maps_bit_exact(ops, n, lo, hi, nb):
    rng = np.random.default_rng(n)
    C0 = rand_coords(rng, n, lo, hi, nb)
    n = C0.shape[0]
    ws = ops.EncoderWorkspace(ops.round_rows(n), 'cuda')
    ops.encoder_build_maps(ws, torch.from_numpy(C0).cuda())
    torch.cuda.synchronize()
    nl = ws.nlvl().cpu().tolist()
    C, s = C0, 1
    kc = ws.kcount()
    for l in range(5):
        assert nl[l] == C.shape[0]
        if l > 0:
            assert np.array_equal(ws.coords(l)[:nl[l]].cpu().numpy(), C)      # first-occurrence order
        ii, oo, kofs = R.build_kmap(C, C, 3, s)
        got = rulebook_from_ws(kc[l], *ws.k3(l), C.shape[0], 27)
        for a, b in zip(got, rulebook_from_oracle(ii, oo, kofs)):
            assert np.array_equal(a, b)
        if l < 4:
            Cn, _ = R.downsample_coords(C, s)
            ii, oo, kofs = R.build_kmap(C, Cn, 2, s)
            got = rulebook_from_ws(kc[5 + l], *ws.k2(l), Cn.shape[0], 8)
            for a, b in zip(got, rulebook_from_oracle(ii, oo, kofs)):
                assert np.array_equal(a, b)
            C, s = Cn, s * 2


def test_voxelize_bit_exact(ops):
    rng = np.random.default_rng(3)
    n_inst, ppi = 6, 1024
    pts = rng.uniform(-0.6, 0.9, (n_inst, ppi, 7)).astype(np.float32)
    pts[:, 100:200] = pts[:, 0:100]                                      # exact duplicates
    cand = np.array([4, 0, 5, 2], np.int32)
    ws = ops.EncoderWorkspace(ops.round_rows(len(cand) * ppi), 'cuda')
    ops.encoder_reset(ws)
    ops.voxelize(torch.from_numpy(pts).cuda(), torch.from_numpy(cand).cuda(), 0.02, ws)
    torch.cuda.synchronize()
    cl, fl = [], []
    for m in cand:
        c, f = R.sparse_quantize(pts[m][:, :3], pts[m], np.array([0.02] * 3))
        cl.append(c)
        fl.append(f)
    Cr, Fr = R.sparse_collate(cl, fl)
    n = int(ws.nlvl()[0])
    assert n == Cr.shape[0]
    assert np.array_equal(ws.coords(0)[:n].cpu().numpy(), Cr.numpy())
    assert np.array_equal(ws.feat0(7)[:n].cpu().numpy(), Fr.numpy())


# ----------------------------------------------------------------------------- sparse conv

def _layer_case(ops, rng, n, cin, cout, ks, use_tc):
    C0 = rand_coords(rng, n, 0, 24, 2)
    n = C0.shape[0]
    ws = ops.EncoderWorkspace(ops.round_rows(n), 'cuda')
    ops.encoder_build_maps(ws, torch.from_numpy(C0).cuda())
    if ks == 3:
        Cout_np, (in_idx, slot), cnt, n_out = C0, ws.k3(0), ws.kcount()[0], ws.nlvl()[0:1]
    else:
        Cout_np, _ = R.downsample_coords(C0, 1)
        (in_idx, slot), cnt, n_out = ws.k2(0), ws.kcount()[5], ws.nlvl()[1:2]
    F = torch.from_numpy(rng.normal(size=(n, cin)).astype(np.float32))
    W = torch.from_numpy((rng.normal(size=(ks ** 3, cin, cout)) / np.sqrt(cin * ks ** 3)).astype(np.float32))
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    shift = torch.from_numpy(rng.normal(size=cout).astype(np.float32) * 0.1)
    n_o = Cout_np.shape[0]
    resid = torch.from_numpy(rng.normal(size=(n_o, cout)).astype(np.float32))
    ii, oo, kofs = R.build_kmap(C0, Cout_np, ks, 1)
    want = torch.relu(R.spconv(F, W, ii, oo, kofs, n_o) * scale + shift + resid)
    out = torch.empty(ws.n_max, cout, device='cuda')
    Wd = W.cuda()
    wprep = ops.spconv_wprep(Wd) if use_tc else None
    ops.spconv_layer(F.cuda(), in_idx, slot, cnt, n_out, ws.n_max, Wd, scale.cuda(), shift.cuda(),
                     torch.cat([resid, torch.zeros(ws.n_max - n_o, cout)]).cuda(), True, ws.T(), out,
                     wprep=wprep, use_tc=use_tc)
    torch.cuda.synchronize()
    return out[:n_o].cpu(), want


@pytest.mark.parametrize('cin,cout,ks', [(7, 32, 3), (32, 64, 2), (64, 64, 3), (64, 128, 2), (128, 128, 3), (128, 128, 2)])
def test_spconv_layer_simt(ops, cin, cout, ks):
    got, want = _layer_case(ops, np.random.default_rng(cin + cout + ks), 3000, cin, cout, ks, False)
    assert float((got - want).abs().max()) < 2e-5


@pytest.mark.parametrize('cin,cout,ks', [(32, 64, 2), (64, 64, 3), (64, 128, 2), (128, 128, 3), (128, 128, 2)])
def test_spconv_layer_tcgen05(ops, cin, cout, ks):
    """tcgen05 split-fp16 (hi/lo operands, fp32 accumulation) pair-GEMM vs the oracle (fp32-level accuracy required)."""
    got, want = _layer_case(ops, np.random.default_rng(cin + cout + ks), 5000, cin, cout, ks, True)
    assert float((got - want).abs().max()) < 2e-5


@pytest.mark.parametrize('use_tc', [False, True])
def test_encoder_features(ops, state_dict, use_tc):
    from instancerefer_b200.basic_blocks import SparseConvEncoder
    rng = np.random.default_rng(11)
    b = synthetic.make_batch(9, batch_size=2, num_points=8000, n_inst=6, n_cand=3, n_tokens=4)
    C0, F0 = b['lidar_coords'], b['lidar_feats']
    enc = SparseConvEncoder(7)
    enc.use_tc = use_tc
    enc.load_state_dict({k[len('scene.net.'):]: v for k, v in state_dict.items() if k.startswith('scene.net.')})
    enc = enc.cuda().eval()
    ws = enc.workspace(C0.shape[0], 'cuda')
    f4, c4, n4 = enc.encode(ws, torch.from_numpy(F0).cuda(), torch.from_numpy(C0).cuda())
    torch.cuda.synchronize()
    sd = {k: v.float() for k, v in state_dict.items()}
    Fr, Cr, s = model_ref.encoder_forward(sd, 'scene.net', torch.from_numpy(F0), torch.from_numpy(C0))
    n = int(n4)
    assert n == Fr.shape[0] and s == 16
    assert np.array_equal(c4[:n].cpu().numpy(), Cr.numpy())
    assert float((f4[:n].cpu() - Fr).abs().max()) < TOL


@pytest.mark.parametrize('mode', ['layers', 'persist'])
@pytest.mark.parametrize('gain', [1e5, 1e-6])
def test_encoder_range_guard(ops, state_dict, gain, mode):
    """Split-fp16 rule GEMM with activations far outside fp16's comfortable range (|x| ~ 1e5 would overflow the hi part,
    |x| ~ 1e-6 would lose the lo part to subnormals): every layer's epilogue records max|out| and the next layer's gather
    is scaled by a power of two (exact), so the tcgen05 path must still agree with the exact fp32 SIMT kernel."""
    from instancerefer_b200.basic_blocks import SparseConvEncoder
    b = synthetic.make_batch(19, batch_size=1, num_points=6000, n_inst=6, n_cand=3, n_tokens=4)
    C0, F0 = torch.from_numpy(b['lidar_coords']).cuda(), torch.from_numpy(b['lidar_feats']).cuda()
    outs = {}
    try:
        for use_tc in (True, False):
            enc = SparseConvEncoder(7)
            enc.use_tc = use_tc
            enc.load_state_dict({k[len('scene.net.'):]: v for k, v in state_dict.items() if k.startswith('scene.net.')})
            with torch.no_grad():
                bn = enc.stem[0].net[1]                     # stem BN gain: everything downstream scales with it
                bn.weight.mul_(gain)
                bn.bias.mul_(gain)
            enc = enc.cuda().eval()
            ops.set_encoder_mode(mode if use_tc else 'layers')
            ws = enc.workspace(C0.shape[0], 'cuda')
            f4, _, n4 = enc.encode(ws, F0, C0)
            outs[use_tc] = f4[:int(n4)].clone()
    finally:
        ops.set_encoder_mode('layers')
    ref = outs[False]
    assert torch.isfinite(outs[True]).all() and float(ref.abs().max()) > 0
    assert float((outs[True] - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize('which', ['single', 'pair'])
def test_persistent_encoder_matches_per_layer_launches(ops, state_dict, which):
    """ir_encoder_mode_set(1): all 13 layers of one / both encoders in one persistent launch (ticketed pair-GEMM and
    reduce items) against the per-layer launch chain, same rulebooks; two launches in a row (the kernel resets its own
    ticket / phase counters)."""
    from instancerefer_b200.basic_blocks import SparseConvEncoder
    encs, wss, f0s = [], [], []
    for seed, key in ((21, 'scene.net.'), (22, 'attribute.net.')):
        b = synthetic.make_batch(seed, batch_size=2, num_points=7000, n_inst=6, n_cand=3, n_tokens=4)
        C0, F0 = torch.from_numpy(b['lidar_coords']).cuda(), torch.from_numpy(b['lidar_feats']).cuda()
        enc = SparseConvEncoder(7)
        enc.load_state_dict({k[len(key):]: v for k, v in state_dict.items() if k.startswith(key)})
        enc = enc.cuda().eval()
        ws = ops.EncoderWorkspace(ops.round_rows(C0.shape[0]), 'cuda')
        ops.encoder_build_maps(ws, C0)
        encs.append(enc); wss.append(ws); f0s.append(F0)

    def run(mode):
        ops.set_encoder_mode(mode)
        outs = [torch.zeros(w.n_max, 128, device='cuda') for w in wss]
        for _ in range(2):
            if which == 'pair':
                ops.encoder_features_pair(encs[0].prepared()['params'], wss[0], f0s[0], outs[0],
                                          encs[1].prepared()['params'], wss[1], f0s[1], outs[1])
            else:
                ops.encoder_features(encs[0].prepared()['params'], wss[0], f0s[0], outs[0])
        torch.cuda.synchronize()
        return [o[:int(w.nlvl()[4])].clone() for o, w in zip(outs, wss)]
    try:
        want, got = run('layers'), run('persist')
    finally:
        ops.set_encoder_mode('layers')
    for g, w in list(zip(got, want))[:2 if which == 'pair' else 1]:
        assert float(w.abs().max()) > 0 and float((g - w).abs().max()) < 2e-6


def test_segmax(ops):
    rng = np.random.default_rng(5)
    F = torch.from_numpy(rng.normal(size=(777, 128)).astype(np.float32))
    C = torch.zeros(777, 4, dtype=torch.int32)
    C[:, 3] = torch.from_numpy(np.sort(rng.integers(0, 9, 777)).astype(np.int32))
    n_dev = torch.tensor([700], dtype=torch.int32).cuda()
    got = ops.segmax(F.cuda(), C.cuda(), n_dev, 777, 9).cpu()
    want = R.global_max_pool(F[:700], C[:700, 3])
    assert torch.equal(got[:want.shape[0]], want)


# ----------------------------------------------------------------------------- scene head

def test_conv2d_on_the_rule_gemm_matches_simt_and_torch(ops):
    """ir_conv2d_3x3_tc: the BEV Conv2d as a sparse conv with a closed-form dense-grid rulebook on the tcgen05 pair-GEMM
    (range-scaled) against the SIMT direct kernel and torch's conv2d, with BN-fold + ReLU and with plain bias."""
    g = torch.Generator().manual_seed(4)
    B, H, W, Cc = 2, 15, 25, 128
    x = torch.randn(B, H, W, Cc, generator=g).cuda() * 3.0
    w = (torch.randn(Cc, Cc, 3, 3, generator=g) / (9 * Cc) ** 0.5).cuda()
    bias, sc, sh = torch.randn(Cc, generator=g).cuda(), (torch.rand(Cc, generator=g) + 0.5).cuda(), torch.randn(Cc, generator=g).cuda()
    wp = w.permute(2, 3, 1, 0).contiguous()                        # [ky][kx][Cin][Cout]
    for relu, s_, b_ in ((True, sc, sh), (False, None, None)):
        want = torch.nn.functional.conv2d(x.cpu().permute(0, 3, 1, 2), w.cpu(), bias.cpu()).permute(0, 2, 3, 1)   # fp32 on the CPU
        if s_ is not None:
            want = want * s_.cpu() + b_.cpu()
        if relu:
            want = want.relu()
        want = want.cuda()
        simt = ops.conv2d_3x3(x, wp, bias, s_, b_, relu)
        amax = x.abs().max().reshape(1).float()
        shift = (s_ * bias + b_) if s_ is not None else bias
        out_amax = torch.zeros(1, device='cuda')
        tc = ops.conv2d_3x3_tc(x, wp.view(9, Cc, Cc), s_, shift.contiguous(), relu, in_absmax=amax, out_absmax=out_amax)
        assert float((simt - want).abs().max()) < 1e-4
        assert float((tc - want).abs().max()) < 1e-4 and tc.shape == want.shape
        assert abs(float(out_amax) - float(tc.abs().max())) < 1e-6
    big = ops.conv2d_3x3_tc(x * 1e5, wp.view(9, Cc, Cc), None, bias, False, in_absmax=(x * 1e5).abs().max().reshape(1))
    want = torch.nn.functional.conv2d((x * 1e5).cpu().permute(0, 3, 1, 2), w.cpu(), bias.cpu()).permute(0, 2, 3, 1).cuda()
    assert torch.isfinite(big).all() and float((big - want).abs().max()) <= 1e-4 * float(want.abs().max())


def test_bev_conv_attention(ops, state_dict):
    rng = np.random.default_rng(6)
    sd = {k: v.float() for k, v in state_dict.items()}
    n, B = 600, 2
    C = rand_coords(rng, n, -2, 30, B)
    C[:, :3] *= 16
    C[:, 2] = (C[:, 2] // 16 % 7 - 1) * 16                                 # z in {-16..80}: some cropped
    C = C[np.sort(np.unique(C, axis=0, return_index=True)[1])]
    n = C.shape[0]
    F = torch.from_numpy(np.abs(rng.normal(size=(n, 128))).astype(np.float32))
    bn = 'scene.to_bev.2'
    scale = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + 1e-5)
    shift = sd[bn + '.bias'] - sd[bn + '.running_mean'] * scale
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    bev = ops.bev(F.cuda(), torch.from_numpy(C).cuda(), n_dev, n, 16, sd['scene.to_bev.1.kernel'].cuda(),
                  scale.cuda(), shift.cuda(), B)
    want = model_ref.bev_forward(sd, 'scene.to_bev', F, torch.from_numpy(C), 16)
    want = torch.relu(model_ref._bn(sd, bn, want, False))
    assert float((bev.permute(0, 3, 1, 2).cpu() - want).abs().max()) < TOL
    # conv2d x2
    w1, b1 = sd['scene.vis_emb_fc.0.weight'], sd['scene.vis_emb_fc.0.bias']
    s1 = sd['scene.vis_emb_fc.1.weight'] / torch.sqrt(sd['scene.vis_emb_fc.1.running_var'] + 1e-5)
    h1 = sd['scene.vis_emb_fc.1.bias'] - sd['scene.vis_emb_fc.1.running_mean'] * s1
    x = ops.conv2d_3x3(bev, w1.permute(2, 3, 1, 0).contiguous().cuda(), b1.cuda(), s1.cuda(), h1.cuda(), True)
    xr = torch.relu(model_ref._bn(sd, 'scene.vis_emb_fc.1', torch.nn.functional.conv2d(want, w1, b1), False))
    assert float((x.permute(0, 3, 1, 2).cpu() - xr).abs().max()) < TOL
    q = torch.from_numpy(rng.normal(size=(B, 128)).astype(np.float32))
    att, sf = ops.scene_attention(x.view(B, -1, 128), q.cuda())
    feats = xr.reshape(B, 128, -1).permute(0, 2, 1)
    a = torch.softmax(torch.bmm(feats, q.unsqueeze(2)).squeeze(2) / np.sqrt(128), 1)
    assert float((att.cpu() - a).abs().max()) < 1e-5
    assert float((sf.cpu() - (feats * a.unsqueeze(2)).sum(1)).abs().max()) < TOL


# ----------------------------------------------------------------------------- language

def test_linear_and_gru(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 300, generator=g)
    W = torch.randn(256, 300, generator=g) / 17
    b = torch.randn(256, generator=g)
    y = ops.linear(x.cuda(), W.cuda(), b.cuda(), relu=True).cpu()
    assert float((y - torch.relu(x @ W.t() + b)).abs().max()) < 1e-5
    gru = torch.nn.GRU(256, 128, num_layers=1, batch_first=True, bidirectional=True)
    B, L = 4, 13
    lens = torch.tensor([13, 1, 7, 10])
    xin = torch.randn(B, L, 256, generator=g)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xin, lens, batch_first=True, enforce_sorted=False)
    want, _ = torch.nn.utils.rnn.pad_packed_sequence(gru(packed)[0], batch_first=True)
    wih = torch.cat([gru.weight_ih_l0, gru.weight_ih_l0_reverse], 0).detach()
    bih = torch.cat([gru.bias_ih_l0, gru.bias_ih_l0_reverse], 0).detach()
    whh = torch.stack([gru.weight_hh_l0, gru.weight_hh_l0_reverse], 0).detach().contiguous()
    bhh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse], 0).detach().contiguous()
    xp = ops.linear(xin.view(B * L, 256).cuda(), wih.cuda(), bih.cuda())
    out = ops.gru_layer(xp, whh.cuda(), bhh.cuda(), lens.cuda(), B, L).cpu()
    assert float((out - want.detach()).abs().max()) < 1e-5


def test_lang_module_vs_oracle(gpu_model, state_dict):
    b = synthetic.make_batch(41, batch_size=3, num_points=3000, n_inst=4, n_cand=2, n_tokens=[5, 17, 1])
    d = dict(lang_feat=torch.from_numpy(b['lang_feat']).cuda(), lang_len=torch.from_numpy(b['lang_len']).cuda())
    out = gpu_model.lang(d)
    ref = model_ref.lang_forward({k: v.float() for k, v in state_dict.items()}, model_ref.data_from_batch(b))
    for k in ('lang_feat', 'atten_attr', 'atten_rel', 'atten_scene', 'lang_attr_feats', 'lang_cls_feats',
              'lang_rel_feats', 'lang_scene_feats', 'lang_scores'):
        assert float((out[k].cpu() - ref[k]).abs().max()) < 1e-5, k


# ----------------------------------------------------------------------------- heads / relation

@pytest.mark.parametrize('norm,mode', [(1, 1), (2, 2), (2, 3), (1, 0), (0, 0)])
def test_mlp_head(ops, norm, mode):
    g = torch.Generator().manual_seed(norm * 10 + mode)
    M, K, N1, N2, S = 19, 128, 256, 256, 3
    x = torch.randn(M, K, generator=g)
    W1, b1 = torch.randn(N1, K, generator=g) / 11, torch.randn(N1, generator=g)
    W2, b2 = torch.randn(N2, N1, generator=g) / 16, torch.randn(N2, generator=g)
    gam, beta = torch.rand(N1, generator=g) + 0.5, torch.randn(N1, generator=g) * 0.1
    partner = torch.randn(S, N2, generator=g)
    seg = torch.randint(0, S, (M,), generator=g, dtype=torch.int32)
    h = x @ W1.t() + b1
    if norm == 2:
        h = torch.nn.functional.layer_norm(h, (N1,), gam, beta, 1e-5)
    elif norm == 1:
        h = h * gam + beta
    yr = torch.relu(h) @ W2.t() + b2
    y, score = ops.mlp_head(x.cuda(), W1.t().contiguous().cuda(), b1.cuda(), norm, gam.cuda(), beta.cuda(),
                            W2.t().contiguous().cuda(), b2.cuda(),
                            mode, partner=partner.cuda(), seg=seg.cuda())
    if mode == 0:
        assert float((y.cpu() - yr).abs().max()) < 2e-5
    elif mode == 1:
        assert float((y.cpu() - torch.nn.functional.normalize(yr, dim=1)).abs().max()) < 1e-5
    elif mode == 2:
        want = (torch.nn.functional.normalize(yr, dim=1) * partner[seg.long()]).sum(1)
        assert float((score.cpu() - want).abs().max()) < 2e-5
    else:
        want = torch.nn.functional.cosine_similarity(yr, partner[seg.long()], dim=1)
        assert float((score.cpu() - want).abs().max()) < 1e-5


def test_knn_edgeconv(ops, state_dict):
    rng = np.random.default_rng(8)
    sd = {k: v.float() for k, v in state_dict.items()}
    sizes = [5, 40, 9]                                                   # 5 < k exercises short segments
    S = sum(sizes)
    xyz = torch.from_numpy(rng.integers(0, 6, (S, 3)).astype(np.float32) * 0.5)   # many exact ties
    cls = torch.from_numpy(rng.integers(0, 18, S))
    feats = torch.cat([xyz, torch.from_numpy(rng.normal(size=(S, 4)).astype(np.float32)),
                       torch.nn.functional.one_hot(cls, 18).float()], 1)
    bidx = torch.cat([torch.full((n,), i) for i, n in enumerate(sizes)])
    fidx = torch.from_numpy(np.sort(rng.choice(S, 20, replace=False)))
    tr = {}
    want = model_ref.edgeconv_forward(sd, 'relation.gcn', xyz, bidx, fidx, feats, 8, 18, tr)
    seg_ofs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32).cuda()
    nbr = ops.knn(xyz.cuda(), seg_ofs, fidx.int().cuda(), bidx[fidx].int().cuda(), 8).cpu()
    flat = [(q, int(j)) for q in range(nbr.shape[0]) for j in nbr[q] if j >= 0]
    assert flat == list(zip(tr['knn_row'].tolist(), tr['knn_col'].tolist()))     # bit-exact incl. tie order
    t = lambda k: sd[k].t().contiguous().cuda()
    c = lambda k: sd[k].contiguous().cuda()
    p = 'relation.gcn.'
    out = ops.edgeconv(feats.cuda(), xyz.cuda(), fidx.int().cuda(), nbr.cuda(), 18,
                       t(p + 'weight.0.weight'), c(p + 'weight.0.bias'), t(p + 'weight.2.weight'), c(p + 'weight.2.bias'),
                       t(p + 'mlp.0.weight'), c(p + 'mlp.0.bias'), t(p + 'mlp.2.weight'), c(p + 'mlp.2.bias')).cpu()
    assert float((out - want).abs().max()) < 2e-5


# ----------------------------------------------------------------------------- whole forward

OUT_KEYS = ['lang_feat', 'atten_attr', 'atten_rel', 'atten_scene', 'lang_cls_feats', 'lang_attr_feats',
            'lang_rel_feats', 'lang_scene_feats', 'lang_scores', 'obj_feats', 'attribute_scores',
            'relation_scores', 'scene_scores', 'seg_scores', 'vis_atten']


def _run(gpu_model, b):
    from instancerefer_b200 import SparseTensor
    out = gpu_model(synthetic.to_data_dict(b, SparseTensor, 'cuda'))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('path', sorted(p for p in glob.glob(os.path.join(GOLD, 'golden_*.npz')) if 'golden_train_' not in p and 'golden_prepare_' not in p))
def test_forward_vs_golden(gpu_model, path):
    """Committed outputs of the reference's models/*.py executed verbatim (oracle/make_golden.py)."""
    z = np.load(path)
    cfg = json.loads(bytes(z['config_json']).decode())
    shift, seed = cfg.pop('shift', None), cfg.pop('seed')
    b = synthetic.make_batch(seed, **cfg)
    if shift is not None:
        b = synthetic.shift_batch(b, shift)
    out = _run(gpu_model, b)
    for k in OUT_KEYS:
        assert float(np.abs(out[k].cpu().numpy() - z[k]).max()) < TOL, k
    assert out['num_filtered_objs'] == z['num_filtered_objs'].tolist()
    for i, o in enumerate(out['pred_obb_batch']):
        assert np.array_equal(np.asarray(o).reshape(-1, 7), z[f'pred_obb_{i}'].reshape(-1, 7))


def test_forward_c2_vs_oracle(gpu_model, state_dict, args):
    """BASELINE.json configs[1]: 1 scene, 40k points, 32 candidates, 20 tokens."""
    b = synthetic.make_batch(123, batch_size=1, num_points=40000, n_inst=32, n_cand=32, n_tokens=20)
    out = _run(gpu_model, b)
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), args)
    for k in OUT_KEYS:
        assert float((out[k].cpu() - ref[k]).abs().max()) < TOL, k
    s = ref['attribute_scores'] + ref['relation_scores'] + ref['scene_scores']
    assert int(out['ref_pred'][0]) == int(s.argmax())
    assert float((out['ref_probs'].cpu() - torch.softmax(s, 0)).abs().max()) < TOL
    out2 = _run(gpu_model, b)                                            # deterministic: bitwise equal
    for k in ('obj_feats', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores'):
        assert torch.equal(out[k], out2[k]), k


def test_forward_c4_relation_stress(gpu_model, state_dict, args):
    """BASELINE.json configs[3] shape (64-instance scenes), reduced batch so the oracle stays fast."""
    b = synthetic.make_batch(7, batch_size=3, num_points=12000, n_inst=64, n_cand=64, n_tokens=[8, 20, 14])
    out = _run(gpu_model, b)
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), args)
    for k in ('relation_scores', 'attribute_scores', 'scene_scores', 'obj_feats'):
        assert float((out[k].cpu() - ref[k]).abs().max()) < TOL, k


def test_streams_and_graph_replay_bitwise(gpu_model):
    """The multi-stream schedule and the CUDA-graph replay run the same kernels on the same data:
    results must be bitwise identical to the plain sequential chain."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.graphed import GraphedInstanceRefer
    keys = ('lang_scores', 'obj_feats', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores',
            'vis_atten', 'ref_probs')
    runner = GraphedInstanceRefer(gpu_model)
    for seed in (301, 302, 303):                       # same shape signature -> one capture, two replays
        b = synthetic.make_batch(seed, batch_size=2, num_points=9000, n_inst=12, n_cand=[5, 4], n_tokens=[9, 6])
        gpu_model.concurrent = False
        seq = _run(gpu_model, b)
        gpu_model.concurrent = True
        con = _run(gpu_model, b)
        gpu_model.pair_encoders = True                 # both encoders through shared launches
        pair = _run(gpu_model, b)
        gpu_model.pair_encoders = False
        for k in keys:
            assert torch.equal(seq[k], pair[k]), ('paired encoders', k)
        gr = runner(synthetic.to_data_dict(b, SparseTensor, 'cpu'))      # host inputs, staged by the runner
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(seq[k], con[k]), ('streams', k)
            assert torch.equal(seq[k], gr[k]), ('graph', k)
        assert gr['num_filtered_objs'] == seq['num_filtered_objs']
    assert len(runner.cache) == 1
    # asynchronous double-buffered form: two steps in flight, host score copies owned by the handle
    bs = [synthetic.make_batch(s_, batch_size=2, num_points=9000, n_inst=12, n_cand=[5, 4], n_tokens=[9, 6])
          for s_ in (311, 312, 313)]
    hs = [runner.submit(synthetic.to_data_dict(b, SparseTensor, 'cpu')) for b in bs[:2]]
    res = [hs[0].result()]
    hs.append(runner.submit(synthetic.to_data_dict(bs[2], SparseTensor, 'cpu')))
    res += [hs[1].result(), hs[2].result()]
    gpu_model.concurrent = False
    for b, r in zip(bs, res):
        want = _run(gpu_model, b)
        for k in ('attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores', 'lang_scores', 'ref_probs'):
            assert torch.equal(r['host_scores'][k], want[k].cpu()), ('pipelined', k)
        assert torch.equal(r['host_scores']['ref_pred'], want['ref_pred'].cpu())
    gpu_model.concurrent = True


def test_graph_replay_survives_workspace_growth(lib_built, state_dict, args):
    """Encoder workspaces only grow and are never freed: a graph captured for a small scene must still replay
    correctly after a larger scene forced a bigger workspace (and vice versa)."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.graphed import GraphedInstanceRefer
    from instancerefer_b200.instancerefer import InstanceRefer
    m = InstanceRefer(7, args)
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda().eval()
    runner = GraphedInstanceRefer(m)
    small = [synthetic.make_batch(s_, batch_size=1, num_points=5000, n_inst=6, n_cand=4, n_tokens=7) for s_ in (401, 402)]
    big = synthetic.make_batch(403, batch_size=1, num_points=30000, n_inst=24, n_cand=20, n_tokens=7)
    keys = ('attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores', 'lang_scores')
    order = [small[0], big, small[1], big, small[0]]
    got = []
    for b in order:
        r = runner(synthetic.to_data_dict(b, SparseTensor, 'cpu'))
        got.append({k: r['host_scores'][k].clone() for k in keys})
    assert len(runner.cache) == 2 and len(m.scene.net._ws) == 2
    ref = InstanceRefer(7, args)
    ref.load_state_dict(state_dict, strict=True)
    ref = ref.cuda().eval()
    ref.concurrent = False
    for b, g in zip(order, got):
        want = _run(ref, b)
        for k in keys:
            assert torch.equal(g[k], want[k].cpu()), k


def test_forward_predicted_language_class(lib_built, state_dict):
    """use_gt_lang: False — candidates filtered by argmax(lang_scores) (models/attribute_module.py:93-97)."""
    from conftest import make_args
    from instancerefer_b200.instancerefer import InstanceRefer
    a = make_args(use_gt_lang=False)
    m = InstanceRefer(7, a)
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda().eval()
    b = synthetic.make_batch(61, batch_size=2, num_points=5000, n_inst=18, n_cand=3, n_tokens=[6, 8])
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), a)
    pred = ref['lang_scores'].argmax(1).tolist()
    for i in range(2):                       # make the predicted class own >= 2 instances in every scene
        b['instance_class'][i][:3] = [pred[i]] * 3
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), a)
    out = _run(m, b)
    assert out['num_filtered_objs'] == ref['num_filtered_objs']
    for k in ('lang_scores', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores'):
        assert float((out[k].cpu() - ref[k]).abs().max()) < TOL, k


def test_forward_c5_large_scene(gpu_model, state_dict, args):
    """BASELINE.json configs[4] sweep point: 120k points, 64 instances (two row buckets above the bench)."""
    b = synthetic.make_batch(77, batch_size=1, num_points=120000, n_inst=64, n_cand=64, n_tokens=20,
                             room=(11.5, 14.0, 3.0))
    out = _run(gpu_model, b)
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), args)
    for k in ('obj_feats', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores', 'vis_atten'):
        assert float((out[k].cpu() - ref[k]).abs().max()) < TOL, k


def test_encoder_sparse_tensor_api(ops, state_dict):
    """SparseConvEncoder.forward(SparseTensor) -> SparseTensor at stride 16 (API fidelity)."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.basic_blocks import SparseConvEncoder
    b = synthetic.make_batch(5, batch_size=1, num_points=3000, n_inst=4, n_cand=2, n_tokens=4)
    enc = SparseConvEncoder(7)
    enc.load_state_dict({k[len('attribute.net.'):]: v for k, v in state_dict.items() if k.startswith('attribute.net.')})
    enc = enc.cuda().eval()
    y = enc(SparseTensor(torch.from_numpy(b['lidar_feats']), torch.from_numpy(b['lidar_coords'])).cuda())
    sd = {k: v.float() for k, v in state_dict.items()}
    Fr, Cr, s = model_ref.encoder_forward(sd, 'attribute.net', torch.from_numpy(b['lidar_feats']),
                                          torch.from_numpy(b['lidar_coords']))
    assert y.s == 16 and np.array_equal(y.C.cpu().numpy(), Cr.numpy())
    assert float((y.F.cpu() - Fr).abs().max()) < TOL


# ----------------------------------------------------------------------------- BASELINE configs[3] / [4] shapes

def test_relation_stress_c4(gpu_model, state_dict, args):
    """configs[3]: 64-instance scenes, batch 32 (S = 2048 instances, 2048 queries, k = 8) through the
    relation module alone against the oracle (kNN edges bit-exact, scores 1e-4)."""
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.candidates import KEY
    b = synthetic.make_batch(700, batch_size=32, num_points=6000, n_inst=64, n_cand=64, n_tokens=5)
    d = synthetic.to_data_dict(b, SparseTensor, 'cuda')
    lang = torch.randn(32, 256, generator=torch.Generator().manual_seed(1))
    d['lang_rel_feats'] = lang.cuda()
    with torch.no_grad():
        d = gpu_model.relation(d)
    torch.cuda.synchronize()
    data = model_ref.data_from_batch(b)
    cands, _, _ = model_ref.candidate_lists(data, data['object_cat'])
    trace = {}
    sd = {k: v.float() if v.is_floating_point() else v for k, v in state_dict.items()}
    ref = model_ref.relation_forward(sd, data, {'lang_rel_feats': lang}, cands, args, trace=trace)
    assert d['relation_scores'].shape == (2048,)
    assert float((d['relation_scores'].cpu() - ref['relation_scores']).abs().max()) < TOL
    nbr = d['_ir_knn'].cpu().numpy()
    assert np.array_equal(nbr.reshape(-1), trace['knn_col'])                 # every query has k = 8 neighbours here


def test_forward_large_sweep_point(gpu_model, state_dict, args):
    """configs[4], one large point of the sweep (120k points, 64 instances = 64 candidates): whole forward
    against the oracle."""
    from instancerefer_b200 import SparseTensor
    b = synthetic.make_batch(900, batch_size=1, num_points=120000, n_inst=64, n_cand=64, n_tokens=20,
                             room=(11.9, 10.4, 3.0))
    with torch.no_grad():
        out = gpu_model(synthetic.to_data_dict(b, SparseTensor, 'cuda'))
    torch.cuda.synchronize()
    ref = model_ref.forward(state_dict, model_ref.data_from_batch(b), args)
    for k in ('lang_scores', 'obj_feats', 'attribute_scores', 'relation_scores', 'scene_scores', 'seg_scores'):
        assert float((out[k].cpu() - ref[k]).abs().max()) < TOL, k
