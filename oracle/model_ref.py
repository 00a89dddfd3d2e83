"""ORACLE (test infrastructure, never shipped): functional CPU restatement of the reference's
hot path ``InstanceRefer.forward`` (models/instancerefer.py:37-70), eval or train mode, fp32 torch
on CPU, on top of ``oracle/sparse_ref.py`` for the third-party sparse/graph operators.

It is written against a *state_dict with the reference's key names* so the same weights drive the
reference (oracle/ref_harness.py), this restatement and the CUDA path.  Each function cites the
reference lines it follows.  It travels to the GPU box (where /root/reference does not exist) and is
the checker for ``-m gpu`` tests, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline.

PARITY STATUS: the reference has no tests/golden vectors for this path ("parity
unpinned" by the reference itself).  This file is pinned by tests/test_oracle_vs_reference.py
(reference files executed verbatim over the shim, in the dev container) and by the committed
fixtures tests/golden/*.npz generated from that verbatim run (oracle/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as Fn

try:
    from . import sparse_ref as R
except ImportError:                      # when oracle/ itself is on sys.path
    import sparse_ref as R


# ----------------------------------------------------------------------------- small helpers

def _lin(sd, p, x):
    return Fn.linear(x, sd[p + '.weight'], sd[p + '.bias'])


BN_RECORD = None      # dict -> train-mode forwards store {prefix: (batch mean, unbiased var)} for the running-stat update


def _bn(sd, p, x, train, eps=1e-5):
    """nn.BatchNorm{1d,2d} forward on (N,C) or (B,C,H,W); batch stats (biased var) in train."""
    w, b = sd[p + '.weight'], sd[p + '.bias']
    if train:
        dims = [0] if x.dim() == 2 else [0, 2, 3]
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if BN_RECORD is not None:
            n = x.numel() // x.shape[1]
            BN_RECORD[p] = (mean.detach().clone(), (var.detach() * (n / max(n - 1, 1))).clone())
    else:
        mean, var = sd[p + '.running_mean'], sd[p + '.running_var']
    shp = (1, -1) if x.dim() == 2 else (1, -1, 1, 1)
    return (x - mean.view(shp)) / torch.sqrt(var.view(shp) + eps) * w.view(shp) + b.view(shp)


def _ln(sd, p, x, eps=1e-5):
    return Fn.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], eps)


# ----------------------------------------------------------------------------- language

def gru_cell(x_proj, h, w_hh, b_hh):
    """one GRU step, gate order [r,z,n]; n = tanh(Wx+b_in + r*(Wh+b_hn)); h'=(1-z)n+zh."""
    hp = h @ w_hh.t() + b_hh
    H = h.shape[-1]
    r = torch.sigmoid(x_proj[..., :H] + hp[..., :H])
    z = torch.sigmoid(x_proj[..., H:2 * H] + hp[..., H:2 * H])
    n = torch.tanh(x_proj[..., 2 * H:] + r * hp[..., 2 * H:])
    return (1 - z) * n + z * h


def bigru(sd, p, x, lengths, hidden=128, layers=2):
    """packed 2-layer bidirectional GRU (models/lang_module.py:22-28,53-57): per-sample length,
    reverse direction starts at each sample's own last token, padded outputs are zeros."""
    B, L, _ = x.shape
    inp = x
    for l in range(layers):
        outs = []
        for suffix, rev in (('', False), ('_reverse', True)):
            w_ih, w_hh = sd[f'{p}.weight_ih_l{l}{suffix}'], sd[f'{p}.weight_hh_l{l}{suffix}']
            b_ih, b_hh = sd[f'{p}.bias_ih_l{l}{suffix}'], sd[f'{p}.bias_hh_l{l}{suffix}']
            xp = inp @ w_ih.t() + b_ih                      # hoisted input projection
            out = torch.zeros(B, L, hidden)
            for b in range(B):
                n = int(lengths[b])
                h = torch.zeros(hidden)
                steps = range(n - 1, -1, -1) if rev else range(n)
                rows = {}
                for t in steps:
                    h = gru_cell(xp[b, t], h, w_hh, b_hh)
                    rows[t] = h
                if n:
                    out[b, :n] = torch.stack([rows[t] for t in range(n)], 0)
            outs.append(out)
        inp = torch.cat(outs, -1)
    return inp


def lang_forward(sd, data, train=False):
    """models/lang_module.py:51-108 (dropout p taken as 0: eval mode or shared-mask parity)."""
    p = 'lang'
    feats_in = data['lang_feat']
    length = data['lang_len'].long()
    embed = torch.relu(_lin(sd, p + '.word_projection.3',
                            torch.relu(_lin(sd, p + '.word_projection.0', feats_in))))   # :33-37,52
    Lmax = int(length.max())
    feats = bigru(sd, p + '.gru', embed[:, :Lmax], length)                                # :53-58
    out = {'lang_feat': feats}
    mask = (torch.arange(Lmax)[None, :] < length[:, None]).float()                       # :60,127-139
    for fc, att_key, emb_key in (('fc_a', 'atten_attr', 'lang_attr_feats'),
                                 ('fc_cls', None, 'lang_cls_feats'),
                                 ('fc_rel', 'atten_rel', 'lang_rel_feats'),
                                 ('fc_scene', 'atten_scene', 'lang_scene_feats')):
        a = _lin(sd, f'{p}.{fc}', feats).squeeze(2)                                        # :61
        a = torch.softmax(a, dim=1) * mask                                                 # :62-63
        a = a / a.sum(1, keepdim=True)                                                     # :64
        out[emb_key] = torch.bmm(a.unsqueeze(1), embed[:, :Lmax]).squeeze(1)               # :65 (embed, not feats)
        if att_key:
            out[att_key] = a
    out['lang_scores'] = _lin(sd, p + '.lang_cls.0', out['lang_cls_feats'])                # :105-106
    return out


# ----------------------------------------------------------------------------- sparse encoder

ENC_STAGES = ((1, 32, 64), (2, 64, 128), (3, 128, 128), (4, 128, 128))


def encoder_forward(sd, p, F, C, train=False, trace=None):
    """SparseConvEncoder / BEVEncoder forward (models/basic_blocks.py:59-95,136-171):
    stem k3 7->32; 4x [k2 s2 down (BN,ReLU), Residual(k3,BN,ReLU,k3,BN) + identity, ReLU].
    F (N,7) fp32, C (N,4) int [x,y,z,b].  -> (F4, C4, stride 16).  ``trace`` (dict) receives the
    per-stride coords and rulebooks for bit-exact index-table parity."""
    Cn = C.numpy() if torch.is_tensor(C) else np.asarray(C)
    s = 1
    maps = {}

    def conv(Fx, Cx, key, ks, stride_now, down):
        kname = ('k2s2_' if down else 'k3_') + str(stride_now)
        if kname not in maps:
            if down:
                Co, _ = R.downsample_coords(Cx, stride_now)
            else:
                Co = Cx
            maps[kname] = (R.build_kmap(Cx, Co, ks, stride_now), Co)
            if trace is not None:
                trace[kname] = dict(in_idx=maps[kname][0][0], out_idx=maps[kname][0][1],
                                    kofs=maps[kname][0][2], coords_out=np.asarray(Co))
        (ii, oo, kofs), Co = maps[kname]
        return R.spconv(Fx, sd[key], ii, oo, kofs, Co.shape[0]), Co

    def cbr(Fx, Cx, base, ks, stride_now, down, relu=True):
        y, Co = conv(Fx, Cx, base + '.0.kernel', ks, stride_now, down)                    # :10-25
        y = _bn(sd, base + '.1', y, train)
        return (torch.relu(y) if relu else y), Co

    x, Cn = cbr(F, Cn, f'{p}.stem.0.net', 3, s, False)
    if trace is not None:
        trace['feat_stem'] = x
    for (i, cin, cout) in ENC_STAGES:
        x, Cn = cbr(x, Cn, f'{p}.stage{i}.0.net', 2, s, True)
        s *= 2
        base = f'{p}.stage{i}.1.net'                                                       # :28-56
        y, _ = conv(x, Cn, base + '.0.kernel', 3, s, False)
        y = torch.relu(_bn(sd, base + '.1', y, train))
        y, _ = conv(y, Cn, base + '.3.kernel', 3, s, False)
        y = _bn(sd, base + '.4', y, train)
        x = torch.relu(y + x)                                                              # identity skip
        if trace is not None:
            trace[f'feat_stage{i}'] = x
    return x, torch.from_numpy(np.asarray(Cn).astype(np.int32)), s


# ----------------------------------------------------------------------------- attribute

def candidate_lists(data, lang_cls_pred):
    """class filter shared by attribute/relation/scene (models/attribute_module.py:42-81):
    per scene the instance ids whose class == target; scenes with <2 candidates contribute no
    tensors but keep their pred_obbs."""
    cands, pred_obb_batch, num_filtered = [], [], []
    for i, classes in enumerate(data['instance_class']):
        tgt = int(lang_cls_pred[i])
        ids = [j for j, c in enumerate(classes) if int(c) == tgt]
        num_filtered.append(len(ids))
        pred_obb_batch.append(np.asarray([data['instance_obbs'][i][j] for j in ids]))
        cands.append(ids if len(ids) >= 2 else [])
    return cands, pred_obb_batch, num_filtered


def attribute_forward(sd, data, lang, args, train=False, trace=None):
    """models/attribute_module.py:83-131."""
    p = 'attribute'
    lf = _lin(sd, p + '.lang_emb_fc.0', lang['lang_attr_feats'])
    lf = torch.relu(_bn(sd, p + '.lang_emb_fc.1', lf, train))
    lf = Fn.normalize(_lin(sd, p + '.lang_emb_fc.3', lf), p=2, dim=1)                      # :88-90
    pred = lang['lang_scores'].argmax(1) if not args.use_gt_lang else data['object_cat']   # :93-97
    cands, pred_obb_batch, num_filtered = candidate_lists(data, pred)
    voxel = np.array([args.voxel_size_ap] * 3)
    cl, fl = [], []
    for i, ids in enumerate(cands):
        for j in ids:
            pc = data['instance_points'][i][j]
            c, f = R.sparse_quantize(pc[:, :3], pc, voxel)                                # :65-69
            cl.append(c)
            fl.append(f)
    C, F = R.sparse_collate(cl, fl)                                                        # :101
    if trace is not None:
        trace['vox_coords'], trace['vox_feats'] = C.numpy().copy(), F.numpy().copy()
    F4, C4, _ = encoder_forward(sd, p + '.net', F, C, train, trace)                       # :104
    obj = R.global_max_pool(F4, C4[:, 3])                                                  # :105
    v = _lin(sd, p + '.vis_emb_fc.0', obj)
    v = torch.relu(_ln(sd, p + '.vis_emb_fc.1', v))
    v = Fn.normalize(_lin(sd, p + '.vis_emb_fc.3', v), p=2, dim=1)                         # :108-113
    rep = torch.cat([lf[i:i + 1].repeat(len(ids), 1) for i, ids in enumerate(cands) if ids], 0)
    return dict(obj_feats=obj, attribute_scores=(v * rep).sum(1),                          # :126
                pred_obb_batch=pred_obb_batch, num_filtered_objs=num_filtered), cands


# ----------------------------------------------------------------------------- relation

def relation_inputs(data, cands, num_classes=18):
    """models/relation_module.py:38-78: for scenes with >=2 candidates every instance -> 25-d
    [obb centre(3), mean rgb+height(4), onehot(18)]; float64 row cast to fp32 (:94)."""
    feats, batch_index, filtered_index, centres = [], [], [], []
    eye = np.eye(num_classes)
    for i, ids in enumerate(cands):
        if not ids:
            continue
        idset = set(ids)
        for j, pc in enumerate(data['instance_points'][i]):
            m = pc.mean(0)
            m[:3] = data['instance_obbs'][i][j][:3]
            feats.append(np.concatenate([m, eye[int(data['instance_class'][i][j])]], -1))
            if j in idset:
                filtered_index.append(len(batch_index))
            batch_index.append(i)
            centres.append(np.asarray(data['instance_obbs'][i][j][:3]))
    return (torch.tensor(np.asarray(feats), dtype=torch.float32),
            torch.tensor(batch_index, dtype=torch.long),
            torch.tensor(filtered_index, dtype=torch.long),
            torch.tensor(np.asarray(centres), dtype=torch.float32))


def edgeconv_forward(sd, p, support_xyz, batch_index, filtered_index, feats, k, ncls=18, trace=None):
    """DynamicEdgeConv (models/basic_blocks.py:98-133)."""
    qxyz, qb, qf = support_xyz[filtered_index], batch_index[filtered_index], feats[filtered_index]
    row, col = R.knn(support_xyz, qxyz, k, batch_index, qb)                                # :120
    if trace is not None:
        trace['knn_row'], trace['knn_col'] = row.numpy(), col.numpy()
    x_i, x_j = qf[row], feats[col]
    w_in = torch.cat([support_xyz[col] - qxyz[row], x_i[:, -ncls:], x_j[:, -ncls:]], -1)   # :131
    w = _lin(sd, p + '.weight.2', torch.relu(_lin(sd, p + '.weight.0', w_in)))
    e = torch.cat([x_i, w, x_j], 1)                                                        # :132
    msg = _lin(sd, p + '.mlp.2', torch.relu(_lin(sd, p + '.mlp.0', e)))
    return R.scatter_max(msg, row, qf.shape[0])                                            # aggr='max'


def relation_forward(sd, data, lang, cands, args, train=False, trace=None):
    """models/relation_module.py:80-107."""
    p = 'relation'
    lf = _lin(sd, p + '.lang_emb_fc.0', lang['lang_rel_feats'])
    lf = torch.relu(_bn(sd, p + '.lang_emb_fc.1', lf, train))
    lf = _lin(sd, p + '.lang_emb_fc.4', lf)                                                # :82
    feats, bidx, fidx, sxyz = relation_inputs(data, cands, args.num_classes)
    g = edgeconv_forward(sd, p + '.gcn', sxyz, bidx, fidx, feats, args.k, args.num_classes, trace)
    if trace is not None:
        trace['rel_feats_in'], trace['gcn_out'] = feats, g
    v = _lin(sd, p + '.vis_emb_fc.0', g)
    v = _lin(sd, p + '.vis_emb_fc.4', torch.relu(_ln(sd, p + '.vis_emb_fc.1', v)))         # :101
    rep = torch.cat([lf[i:i + 1].repeat(len(ids), 1) for i, ids in enumerate(cands) if ids], 0)
    return dict(relation_scores=Fn.cosine_similarity(v, rep, dim=1))                       # :103


# ----------------------------------------------------------------------------- scene

BEV_MAX = (240, 400, 80)      # models/scene_module.py:22


def bev_forward(sd, p, F4, C4, stride, batch_size_hint=None):
    """SparseCrop + ToDenseBEVConvolution (models/basic_blocks.py:174-243): keep 0<=xyz<BEV_MAX;
    f' = f @ kernel[z//s]; dense[b, x//s, y//s] += f' (duplicates SUM); -> (B,128,15,25)."""
    C4 = C4.long()
    lim = torch.tensor(BEV_MAX)
    keep = ((C4[:, :3] >= 0) & (C4[:, :3] < lim)).all(-1)                                   # :179
    F4, C4 = F4[keep], C4[keep]
    kern = sd[p + '.1.kernel']
    f = torch.einsum('nc,nco->no', F4, kern[C4[:, 2] // stride])                            # :231-232
    H, W = BEV_MAX[0] // 16, BEV_MAX[1] // 16
    B = int(C4[:, 3].max()) + 1                                                            # :235
    flat = C4[:, 3] * (H * W) + (C4[:, 0] // stride) * W + (C4[:, 1] // stride)             # :236-237
    dense = torch.zeros(B * H * W, f.shape[1]).index_add(0, flat, f)                        # to_dense sums
    return dense.view(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def scene_forward(sd, data, lang, attr, cands, args, train=False, trace=None):
    """models/scene_module.py:60-108."""
    p = 'scene'
    B = data['point_min'].shape[0]                                                         # :62-63
    F4, C4, s = encoder_forward(sd, p + '.net', data['lidar_F'], data['lidar_C'], train, trace)
    bev = bev_forward(sd, p + '.to_bev', F4, C4, s)
    bev = torch.relu(_bn(sd, p + '.to_bev.2', bev, train))                                  # :25-30
    if trace is not None:
        trace['bev'] = bev
    x = Fn.conv2d(bev, sd[p + '.vis_emb_fc.0.weight'], sd[p + '.vis_emb_fc.0.bias'])
    x = torch.relu(_bn(sd, p + '.vis_emb_fc.1', x, train))
    x = Fn.conv2d(x, sd[p + '.vis_emb_fc.4.weight'], sd[p + '.vis_emb_fc.4.bias'])          # :71
    h, w = x.shape[-2:]
    feats = x.reshape(B, 128, -1).permute(0, 2, 1)                                         # :74
    lf = _lin(sd, p + '.lang_emb_fc.0', lang['lang_scene_feats'])
    lf = _lin(sd, p + '.lang_emb_fc.4', torch.relu(_ln(sd, p + '.lang_emb_fc.1', lf))).unsqueeze(2)
    atten = torch.softmax((torch.bmm(feats, lf) / math.sqrt(feats.shape[2])).squeeze(2), dim=1)  # :77-80
    scene_feats = (feats * atten.unsqueeze(2)).sum(1)                                      # :83
    c = _lin(sd, p + '.cls.0', scene_feats)
    seg = _lin(sd, p + '.cls.3', torch.relu(_bn(sd, p + '.cls.1', c, train)))               # :84
    rep = torch.cat([scene_feats[i:i + 1].repeat(len(ids), 1) for i, ids in enumerate(cands) if ids], 0)
    o = _lin(sd, p + '.vis_emb_fc1.0', attr['obj_feats'])
    o = _lin(sd, p + '.vis_emb_fc1.4', torch.relu(_ln(sd, p + '.vis_emb_fc1.1', o)))        # :103
    return dict(vis_atten=atten.reshape(B, h, w), seg_scores=seg,
                scene_scores=Fn.cosine_similarity(o, rep, dim=1))                           # :104


# ----------------------------------------------------------------------------- whole forward

def forward(sd, data, args, train=False, trace=None, keep_grad=False):
    """InstanceRefer.forward (models/instancerefer.py:56-70).  ``data``: lang_feat (B,126,300),
    lang_len (B,), object_cat (B,), lidar_F (N,7), lidar_C (N,4), point_min (B,3), and the host
    lists instance_points / instance_obbs / instance_class.  Returns the written dict entries."""
    if not keep_grad:
        sd = {k: (v.detach().float() if v.is_floating_point() else v) for k, v in sd.items()}
    out = {}
    lang = lang_forward(sd, data, train)
    out.update(lang)
    attr, cands = attribute_forward(sd, data, lang, args, train,
                                    None if trace is None else trace.setdefault('attribute', {}))
    out.update(attr)
    out.update(relation_forward(sd, data, lang, cands, args, train,
                                None if trace is None else trace.setdefault('relation', {})))
    out.update(scene_forward(sd, data, lang, attr, cands, args, train,
                             None if trace is None else trace.setdefault('scene', {})))
    return out


def data_from_batch(batch):
    """numpy batch from instancerefer_b200.synthetic.make_batch -> oracle input dict."""
    d = dict(batch)
    d['lidar_F'] = torch.from_numpy(batch['lidar_feats'])
    d['lidar_C'] = torch.from_numpy(batch['lidar_coords'])
    for k in ('lang_feat', 'lang_len', 'object_cat', 'point_min'):
        d[k] = torch.from_numpy(np.asarray(batch[k]))
    return d
