"""Train-mode forward of the hot path with gradients: what
``lib/solver.py:196-205`` gets from the reference modules when ``model.train()`` is set — batch
statistics BatchNorm everywhere, Dropout, and outputs that carry autograd history so the caller's
``loss.backward()`` / optimiser step work unchanged.

Every differentiable operator is a ``torch.autograd.Function`` whose forward AND backward are calls
into the CUDA library (ops.*); torch only chains them (plumbing).  Sparse encoders:
conv = pair-GEMM + reduce (the eval kernels, identity epilogue) -> train-mode BN kernel (+ residual,
ReLU); backward = BN backward, dgrad (the same pair-GEMM/reduce on the transposed rulebook with
W^T) and wgrad (per-offset gathered outer products).
"""
import math

import torch
from torch.autograd import Function

from . import ops


# ----------------------------------------------------------------------------- sparse encoder

class EncoderGraph:
    """Rulebooks of one encoder pass (inside its workspace) + host row counts, and the transposed
    rulebooks the backward needs (built lazily, once per map per step)."""

    def __init__(self, ws):
        self.ws = ws
        self.nlvl_dev = ws.nlvl()
        self.n = [int(v) for v in self.nlvl_dev.tolist()]          # one small D2H per encoder pass
        self.kcount = ws.kcount()
        self._t = {}

    def map(self, kind, level):
        """-> in_idx, slot, count, n_in, n_out, n_in_dev, n_out_dev ; kind 'k3' (level->level) or
        'k2' (level -> level+1)."""
        if kind == 'k3':
            ii, sl = self.ws.k3(level)
            return ii, sl, self.kcount[level], self.n[level], self.n[level], \
                self.nlvl_dev[level:level + 1], self.nlvl_dev[level:level + 1]
        ii, sl = self.ws.k2(level)
        return ii, sl, self.kcount[5 + level], self.n[level], self.n[level + 1], \
            self.nlvl_dev[level:level + 1], self.nlvl_dev[level + 1:level + 2]

    def transposed(self, kind, level):
        key = (kind, level)
        if key not in self._t:
            ii, sl, _, _, n_out, _, n_out_dev = self.map(kind, level)
            self._t[key] = ops.rulebook_transpose(ii, sl, n_out_dev, n_out)
        return self._t[key]


class SparseConvBN(Function):
    """spnn.Conv3d -> spnn.BatchNorm (train) [-> + residual] [-> ReLU]
    (models/basic_blocks.py:10-25,28-56)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, resid, G, kind, level, relu, bn, use_tc):
        ii, sl, count, n_in, n_out, _, n_out_dev = G.map(kind, level)
        K, cin, cout = weight.shape
        x = x.contiguous()
        w = weight.detach().contiguous()
        y = torch.empty(n_out, cout, dtype=torch.float32, device=x.device)
        tc = bool(use_tc and cin >= 32)
        ops.spconv_layer(x, ii, sl, count, n_out_dev, n_out, w, None, None, None, False, G.ws.T(), y,
                         wprep=w if tc else None, use_tc=tc)
        mom = bn.momentum if bn.momentum is not None else 0.1
        out, mean, rstd = ops.bn_train_fwd(y, gamma.detach(), beta.detach(), resid, relu, bn.eps, mom,
                                           bn.running_mean, bn.running_var)
        bn.num_batches_tracked += 1
        ctx.G, ctx.kind, ctx.level, ctx.relu, ctx.has_resid = G, kind, level, relu, resid is not None
        ctx.save_for_backward(x, w, gamma.detach(), y, out, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, gamma, y, out, mean, rstd = ctx.saved_tensors
        G = ctx.G
        ii, sl, count, n_in, n_out, n_in_dev, _ = G.map(ctx.kind, ctx.level)
        dy, dres, dgamma, dbeta = ops.bn_train_bwd(dout.contiguous(), out, y, mean, rstd, gamma, ctx.relu,
                                                   ctx.has_resid)
        out_idx, slot_in = G.transposed(ctx.kind, ctx.level)
        dW = ops.spconv_wgrad(x, dy, ii, out_idx, count)
        dx = None
        if ctx.needs_input_grad[0]:
            K, cin, cout = w.shape
            wt = w.transpose(1, 2).contiguous()                       # (K, Cout, Cin)
            dx = torch.empty(n_in, cin, dtype=torch.float32, device=x.device)
            # dgrad = the forward pipeline on the transposed rulebook (exact fp32 SIMT pair-GEMM:
            # gradients span too many orders of magnitude for the split-fp16 tensor-core operands)
            ops.spconv_layer(dy, out_idx, slot_in, count, n_in_dev, n_in, wt, None, None, None, False,
                             G.ws.T(), dx, wprep=None, use_tc=False)
        return dx, dW, dgamma, dbeta, dres, None, None, None, None, None, None


class SegMax(Function):
    """spnn.GlobalMaxPooling (models/attribute_module.py:105)."""

    @staticmethod
    def forward(ctx, feats, coords, n_dev, n_rows, n_seg):
        pooled = ops.segmax(feats, coords, n_dev, n_rows, n_seg)
        ctx.save_for_backward(feats, coords, n_dev, pooled)
        ctx.n_rows, ctx.n_seg = n_rows, n_seg
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        feats, coords, n_dev, pooled = ctx.saved_tensors
        return ops.segmax_bwd(feats, coords, n_dev, ctx.n_rows, ctx.n_seg, pooled, dpooled.contiguous()), \
            None, None, None, None


def encoder_forward_train(net, ws, feats0=None, coords0=None):
    """Train-mode pass of a SparseConvEncoder / BEVEncoder over a workspace whose level 0 is
    already voxelised (feats0 None) or given by (feats0, coords0).
    -> (F4 (n4,128) with autograd history, EncoderGraph)."""
    ops.encoder_build_maps(ws, coords0)
    G = EncoderGraph(ws)
    if feats0 is None:
        feats0 = ws.feat0(net.input_dim)[:G.n[0]].clone()
    tc = net.use_tc
    layers = net._layers()

    def cbr(x, idx, kind, level, relu=True, resid=None):
        conv, bn = layers[idx]
        return SparseConvBN.apply(x, conv.kernel, bn.weight, bn.bias, resid, G, kind, level, relu, bn, tc)

    x = cbr(feats0, 0, 'k3', 0)
    for s in range(1, 5):
        li = 1 + 3 * (s - 1)
        x = cbr(x, li, 'k2', s - 1)
        y = cbr(x, li + 1, 'k3', s)
        x = cbr(y, li + 2, 'k3', s, relu=True, resid=x)             # relu(bn(conv(y)) + x)
    return x, G


# ----------------------------------------------------------------------------- helpers (heads)

def dropout(x, module):
    """nn.Dropout in train mode (p may be patched to 0 for parity tests)."""
    return torch.nn.functional.dropout(x, module.p, True) if module.p > 0 else x


def repeat_by_scene(x, pack):
    """row i of the per-scene matrix repeated for each of that scene's candidates
    (models/attribute_module.py:116-125)."""
    return x.index_select(0, pack.cand_scene.long())


# ----------------------------------------------------------------------------- module forwards
# Dense parts below marked (torch) still run on torch/cuBLAS/cuDNN kernels under autograd in this
# round; DESIGN.md §0 lists them.  The sparse encoders, pooling, loss and optimiser are CUDA-library
# kernels in both directions.

def lang_forward_train(m, data_dict):
    """models/lang_module.py:51-108 in train mode."""
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    x, length = data_dict['lang_feat'], data_dict['lang_len']
    len_cpu = length.detach().to('cpu')
    L = int(len_cpu.max())
    B = x.shape[0]
    e = m.word_projection(x[:, :L].float())                                     # (torch)
    packed = pack_padded_sequence(e, len_cpu, batch_first=True, enforce_sorted=False)
    feats, _ = pad_packed_sequence(m.gru(packed)[0], batch_first=True)          # (torch, cuDNN GRU)
    data_dict['lang_feat'] = feats
    mask = (torch.arange(L, device=x.device)[None, :] < length.to(x.device)[:, None]).float()
    fcw = torch.cat([m.fc_a.weight, m.fc_cls.weight, m.fc_rel.weight, m.fc_scene.weight], 0)   # (4,256)
    fcb = torch.cat([m.fc_a.bias, m.fc_cls.bias, m.fc_rel.bias, m.fc_scene.bias], 0)
    a = torch.softmax(feats @ fcw.t() + fcb, dim=1) * mask[:, :, None]          # (B,L,4)
    a = a / a.sum(1, keepdim=True)
    pooled = torch.einsum('blh,bld->hbd', a, e)                                 # (4,B,256)
    data_dict['atten_attr'], data_dict['atten_rel'], data_dict['atten_scene'] = a[..., 0], a[..., 2], a[..., 3]
    data_dict['lang_attr_feats'], data_dict['lang_cls_feats'] = pooled[0], pooled[1]
    data_dict['lang_rel_feats'], data_dict['lang_scene_feats'] = pooled[2], pooled[3]
    if m.use_lang_classifier:
        data_dict['lang_scores'] = m.lang_cls(data_dict['lang_cls_feats'])
    return data_dict


def attribute_forward_train(m, data_dict, pack):
    """models/attribute_module.py:83-131 in train mode."""
    dev = pack.points.device
    lang = torch.nn.functional.normalize(m.lang_emb_fc(data_dict['lang_attr_feats']), p=2, dim=1)   # (torch)
    data_dict['num_filtered_objs'] = pack.num_filtered
    data_dict['pred_obb_batch'] = pack.pred_obb_batch
    ws = m.net.workspace(pack.M * pack.points.shape[1], dev)
    ops.encoder_reset(ws)
    ops.voxelize(pack.points, pack.cand_rows, float(m.voxel_size[0]), ws)
    f4, G = encoder_forward_train(m.net, ws)
    obj = SegMax.apply(f4, ws.coords(4), G.nlvl_dev[4:5], G.n[4], pack.M)
    data_dict['obj_feats'] = obj
    vis = torch.nn.functional.normalize(m.vis_emb_fc(obj), p=2, dim=1)                               # (torch)
    data_dict['attribute_scores'] = (vis * repeat_by_scene(lang, pack)).sum(1)
    return data_dict


def relation_forward_train(m, data_dict, pack):
    """models/relation_module.py:80-107 in train mode; the graph (kNN) comes from the CUDA library,
    the edge MLPs run under autograd (torch)."""
    dev = pack.points.device
    lang = m.lang_emb_fc(data_dict['lang_rel_feats'])
    mean = ops.instance_mean(pack.points)
    ncls = m.args.num_classes
    onehot = (pack.centres_cls[:, 3:4] == torch.arange(ncls, device=dev, dtype=torch.float32)[None, :]).float()
    xyz = pack.centres_cls[:, :3].contiguous()
    feats = torch.cat([xyz, mean[:, 3:], onehot], 1).contiguous()
    nbr = ops.knn(xyz, pack.inst_ofs, pack.cand_rows, pack.cand_seg, m.gcn.k)              # (M,k), -1 padded
    q = pack.cand_rows.long()
    valid = nbr >= 0
    j = nbr.clamp(min=0).long()
    x_i = feats[q][:, None, :].expand(-1, nbr.shape[1], -1)
    x_j = feats[j]
    w_in = torch.cat([xyz[j] - xyz[q][:, None, :], x_i[..., -ncls:], x_j[..., -ncls:]], -1)
    w = m.gcn.weight(w_in)
    msg = m.gcn.mlp(torch.cat([x_i, w, x_j], -1))                                           # (M,k,128)
    msg = torch.where(valid[..., None], msg, torch.full_like(msg, float('-inf')))
    g = msg.max(1)[0]
    vis = m.vis_emb_fc(g)
    data_dict['relation_scores'] = torch.nn.functional.cosine_similarity(vis, repeat_by_scene(lang, pack), dim=1)
    return data_dict


def scene_forward_train(m, data_dict, pack):
    """models/scene_module.py:60-108 in train mode."""
    lidar = data_dict['lidar']
    dev = pack.points.device
    B = data_dict['point_min'].shape[0]
    F0 = lidar.F.to(dev, torch.float32).contiguous()
    C0 = lidar.C.to(dev, torch.int32).contiguous()
    ws = m.net.workspace(F0.shape[0], dev)
    f4, G = encoder_forward_train(m.net, ws, F0, C0)
    c4 = ws.coords(4)[:G.n[4]].long()
    # SparseCrop + ToDenseBEVConvolution (models/basic_blocks.py:174-243)                   (torch)
    keep = ((c4[:, 0] >= 0) & (c4[:, 0] < 240) & (c4[:, 1] >= 0) & (c4[:, 1] < 400) & (c4[:, 2] >= 0) & (c4[:, 2] < 80))
    kern = m.to_bev[1].kernel
    z = (c4[:, 2] // 16).clamp(0, kern.shape[0] - 1)
    fz = torch.zeros(f4.shape[0], kern.shape[2], device=dev)
    for zi in range(kern.shape[0]):
        sel = (z == zi) & keep
        fz = fz + torch.where(sel[:, None], f4 @ kern[zi], torch.zeros_like(fz))
    flat = (c4[:, 3] * 375 + (c4[:, 0] // 16).clamp(0, 14) * 25 + (c4[:, 1] // 16).clamp(0, 24)).clamp(0, B * 375 - 1)
    dense = torch.zeros(B * 375, kern.shape[2], device=dev).index_add(0, flat, fz * keep[:, None].float())
    bev = dense.view(B, 15, 25, -1).permute(0, 3, 1, 2)
    bev = torch.relu(m.to_bev[2](bev))
    x = m.vis_emb_fc(bev)                                                                   # (torch, cuDNN)
    feats = x.reshape(B, m.h_dim, -1).permute(0, 2, 1)
    lang = m.lang_emb_fc(data_dict['lang_scene_feats']).unsqueeze(2)
    atten = torch.softmax((torch.bmm(feats, lang) / math.sqrt(feats.shape[2])).squeeze(2), dim=1)
    data_dict['vis_atten'] = atten.reshape(B, x.shape[2], x.shape[3])
    scene_feats = (feats * atten.unsqueeze(2)).sum(1)
    data_dict['seg_scores'] = m.cls(scene_feats)
    obj = m.vis_emb_fc1(data_dict['obj_feats'])
    data_dict['scene_scores'] = torch.nn.functional.cosine_similarity(obj, repeat_by_scene(scene_feats, pack), dim=1)
    return data_dict
