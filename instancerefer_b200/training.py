"""Train-mode forward of the hot path with gradients: what
``lib/solver.py:196-205`` gets from the reference modules when ``model.train()`` is set — batch
statistics BatchNorm everywhere, Dropout, and outputs that carry autograd history so the caller's
``loss.backward()`` / optimiser step work unchanged.

Every differentiable operator is a ``torch.autograd.Function`` whose forward AND backward are calls
into the CUDA library (ops.*); torch only chains them (plumbing: views, cat/stack of parameters,
gradient accumulation).  Sparse encoders:
conv = pair-GEMM + reduce (the eval kernels, identity epilogue) -> train-mode BN kernel (+ residual,
ReLU); backward = BN backward, dgrad (the same pair-GEMM/reduce on the transposed rulebook with
W^T) and wgrad (per-offset gathered outer products).
"""
import os

import torch
from torch.autograd import Function

from . import ops


# ----------------------------------------------------------------------------- sparse encoder

class EncoderGraph:
    """Rulebooks of one encoder pass (inside its workspace) + host row counts, and the transposed
    rulebooks the backward needs (built lazily, once per map per step)."""

    def __init__(self, ws, n=None):
        self.ws = ws
        self.nlvl_dev = ws.nlvl()
        self.n = [int(v) for v in (n if n is not None else self.nlvl_dev.tolist())]   # host row counts (one small D2H)
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


class EncoderTrain(Function):
    """All 13 layers of an encoder as ONE library call per direction (ir_encoder_train_forward /
    ir_encoder_train_backward): activations stay in an arena owned by this node of the autograd graph."""

    @staticmethod
    def forward(ctx, feats0, net, G, *params):
        import ctypes as C
        import os
        from . import _lib
        ws, dev = G.ws, G.ws.buf.device
        layers = net._layers()
        P = _lib.EncoderTrainParams()
        dgrad_tc = os.environ.get('IR_DGRAD', 'tc') != 'simt'
        wgrad_tc = os.environ.get('IR_WGRAD', 'tc') != 'simt'
        P.cin, P.eps = net.input_dim, layers[0][1].eps
        P.use_tc = (1 | (2 if dgrad_tc else 0) | (4 if wgrad_tc else 0)) if net.use_tc else 0
        keep = []
        for i, (conv, bn) in enumerate(layers):
            w, g, b = (t.detach().contiguous() for t in params[3 * i:3 * i + 3])
            keep += [w, g, b]
            P.weight[i], P.gamma[i], P.beta[i] = w.data_ptr(), g.data_ptr(), b.data_ptr()
            P.running_mean[i], P.running_var[i] = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            P.momentum[i] = bn.momentum if bn.momentum is not None else 0.1
        n_lvl = (C.c_int32 * 5)(*G.n)
        lay = _lib.EncoderTrainLayout()
        _lib.call("ir_encoder_train_layout", ws.n_max, n_lvl, net.input_dim, C.byref(lay))
        arena = torch.empty(lay.total_bytes, dtype=torch.uint8, device=dev)
        f0 = feats0.contiguous() if feats0 is not None else None
        _lib.call("ir_encoder_train_forward", C.byref(P), ops._p(f0), ws.ptr, ws.n_max, n_lvl, ops._p(arena), ops._stream())
        torch._foreach_add_([bn.num_batches_tracked for _, bn in layers], 1)
        ctx.state = (P, keep, n_lvl, arena, f0, G, [tuple(t.shape) for t in params])
        o = lay.off_out[12]
        return arena[o:o + G.n[4] * 128 * 4].view(torch.float32).view(G.n[4], 128)

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        from . import _lib
        P, keep, n_lvl, arena, f0, G, shapes = ctx.state
        ws = G.ws
        sizes = [(int(torch.Size(sh).numel()) + 63) // 64 * 64 for sh in shapes]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=arena.device)
        grads, o = [], 0
        for sh, sz in zip(shapes, sizes):
            grads.append(flat[o:o + torch.Size(sh).numel()].view(sh))
            o += sz
        Gr = _lib.EncoderTrainGrads()
        for i in range(13):
            Gr.dweight[i], Gr.dgamma[i], Gr.dbeta[i] = (grads[3 * i + q].data_ptr() for q in range(3))
        _lib.call("ir_encoder_train_backward", C.byref(P), ops._p(f0), ws.ptr, ws.n_max, n_lvl, ops._p(arena),
                  ops._p(dout.contiguous(), torch.float32), C.byref(Gr), ops._stream())
        return (None, None, None, *grads)


# Set by optim.FlatAdam when it overlaps the data-parallel all-reduce with the backward: callable(params, grads) -> bool,
# handed gradients that are final although the autograd node producing them has not returned yet.
EARLY_GRAD_HOOK = None


class _EncoderGraphState:
    """Static buffers + captured CUDA graphs of one encoder at one workspace capacity (n_max bucket)."""


class EncoderTrainGraphed(Function):
    """EncoderTrain with the ~65 forward / ~150 backward launches of an encoder replayed from two CUDA graphs.
    The library calls run in capacity mode (n_lvl = NULL: buffers laid out for n_max rows, row counts read on
    the device), so the launch sequence depends only on the workspace capacity and the graphs are captured once
    per (encoder, n_max bucket) and reused for every batch that fits the bucket.  Inputs that live at changing
    addresses (scene features, the upstream gradient) are copied into static buffers first; parameter gradients
    are returned as views of a static buffer that the next backward of this encoder overwrites."""

    @staticmethod
    def _state(net, G, params, f0_cols):
        import ctypes as C
        import os
        from . import _lib
        ws = G.ws
        layers = net._layers()
        moms = tuple(float(bn.momentum if bn.momentum is not None else 0.1) for _, bn in layers)
        dgrad_tc = os.environ.get('IR_DGRAD', 'tc') != 'simt'
        wgrad_tc = os.environ.get('IR_WGRAD', 'tc') != 'simt'
        use_tc = (1 | (2 if dgrad_tc else 0) | (4 if wgrad_tc else 0)) if net.use_tc else 0
        key = (ws.n_max, ws.buf.data_ptr(), f0_cols, moms, use_tc, tuple(t.data_ptr() for t in params))
        cache = net.__dict__.setdefault('_train_graphs', {})
        st = cache.get(key)
        if st is not None:
            return st
        if len(cache) >= 4:
            cache.clear()
        dev = ws.buf.device
        st = _EncoderGraphState()
        P = _lib.EncoderTrainParams()
        P.cin, P.use_tc, P.eps = net.input_dim, use_tc, layers[0][1].eps
        for i, (conv, bn) in enumerate(layers):
            w, g, b = params[3 * i:3 * i + 3]
            assert w.is_contiguous() and g.is_contiguous() and b.is_contiguous()
            P.weight[i], P.gamma[i], P.beta[i] = w.data_ptr(), g.data_ptr(), b.data_ptr()
            P.running_mean[i], P.running_var[i] = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            P.momentum[i] = moms[i]
        lay = _lib.EncoderTrainLayout()
        _lib.call("ir_encoder_train_layout", ws.n_max, None, net.input_dim, C.byref(lay))
        st.P, st.lay = P, lay
        st.arena = torch.empty(lay.total_bytes, dtype=torch.uint8, device=dev)
        st.f0 = torch.zeros(ws.n_max, f0_cols, dtype=torch.float32, device=dev) if f0_cols else None
        st.dout = torch.zeros(ws.n_max, 128, dtype=torch.float32, device=dev)
        shapes = [tuple(t.shape) for t in params]
        sizes = [(int(torch.Size(sh).numel()) + 63) // 64 * 64 for sh in shapes]
        st.gflat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        st.gslices, o = [], 0
        for sh, sz in zip(shapes, sizes):
            st.gslices.append((o, int(torch.Size(sh).numel()), sh))
            o += sz
        Gr = _lib.EncoderTrainGrads()
        base = st.gflat.data_ptr()
        for i in range(13):
            Gr.dweight[i], Gr.dgamma[i], Gr.dbeta[i] = (base + 4 * st.gslices[3 * i + q][0] for q in range(3))
        st.Gr = Gr
        st.fwd_graph = st.bwd_graph = None
        st.fwd_runs = st.bwd_runs = 0
        o4 = lay.off_out[12]
        st.f4 = st.arena[o4:o4 + ws.n_max * 128 * 4].view(torch.float32).view(ws.n_max, 128)
        cache[key] = st
        return st

    @staticmethod
    def _run(st, which, ws):
        """First call eager (also initialises one-time kernel attributes), second call captures, then replay."""
        import ctypes as C
        from . import _lib

        def launch():
            f0 = ops._p(st.f0)
            if which == 'fwd':
                _lib.call("ir_encoder_train_forward", C.byref(st.P), f0, ws.ptr, ws.n_max, None, ops._p(st.arena), ops._stream())
            elif which == 'bwd':
                _lib.call("ir_encoder_train_backward", C.byref(st.P), f0, ws.ptr, ws.n_max, None, ops._p(st.arena),
                          ops._p(st.dout), C.byref(st.Gr), ops._stream())
            else:                                            # ('range', stage_hi, stage_lo): whole-step capture only
                _lib.call("ir_encoder_train_backward_range", C.byref(st.P), f0, ws.ptr, ws.n_max, None, ops._p(st.arena),
                          ops._p(st.dout), C.byref(st.Gr), which[1], which[2], ops._stream())
        if torch.cuda.is_current_stream_capturing():          # whole-step capture (train_graph.GraphedTrainStep)
            launch()
            return
        runs = st.fwd_runs if which == 'fwd' else st.bwd_runs
        graph = st.fwd_graph if which == 'fwd' else st.bwd_graph
        if graph is not None:
            graph.replay()
        elif runs == 0:
            launch()
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                launch()
            g.replay()
            if which == 'fwd':
                st.fwd_graph = g
            else:
                st.bwd_graph = g
        if which == 'fwd':
            st.fwd_runs += 1
        else:
            st.bwd_runs += 1

    @staticmethod
    def forward(ctx, feats0, net, G, *params):
        ws = G.ws
        st = EncoderTrainGraphed._state(net, G, params, feats0.shape[1] if feats0 is not None else 0)
        if feats0 is not None:
            st.f0[:feats0.shape[0]].copy_(feats0)
        EncoderTrainGraphed._run(st, 'fwd', ws)
        torch._foreach_add_([bn.num_batches_tracked for _, bn in net._layers()], 1)
        ctx.st, ctx.ws, ctx.n4, ctx.params = st, ws, G.n[4], params
        return st.f4[:G.n[4]]

    @staticmethod
    def backward(ctx, dout):
        st = ctx.st
        ops.stamp(f'bwd:enc{st.P.cin}:start')
        st.dout[:ctx.n4].copy_(dout)
        if EARLY_GRAD_HOOK is not None and torch.cuda.is_current_stream_capturing():
            # data parallel, captured step: stages 4..3 first (layers 7..12, 2/3 of the parameters, small levels), hand
            # their gradients to the optimiser so that their all-reduce runs beside the large shallow levels
            EncoderTrainGraphed._run(st, ('range', 4, 3), ctx.ws)
            lo = 3 * 7
            EARLY_GRAD_HOOK(ctx.params[lo:], [st.gflat[o:o + n].view(sh) for o, n, sh in st.gslices[lo:]])
            EncoderTrainGraphed._run(st, ('range', 2, 0), ctx.ws)
        else:
            EncoderTrainGraphed._run(st, 'bwd', ctx.ws)
        ops.stamp(f'bwd:enc{st.P.cin}:end')
        # one copy of the static gradient buffer per backward (a single D2D kernel), returned as views: autograd
        # adopts them as .grad without per-parameter kernels, and nothing the caller holds aliases the buffer the
        # next replay overwrites (gradient accumulation over several backwards stays correct)
        flat = st.gflat.clone()
        return (None, None, None, *(flat[o:o + n].view(sh) for o, n, sh in st.gslices))


def encoder_forward_train(net, ws, feats0=None, coords0=None, G=None):
    """Train-mode pass of a SparseConvEncoder / BEVEncoder over a workspace whose level 0 is
    already voxelised (feats0 None) or given by (feats0, coords0).  ``G``: maps already built and
    row counts already read back (the model forward builds both encoders' maps first and reads
    their counts with ONE D2H copy).
    -> (F4 (n4,128) with autograd history, EncoderGraph).  IR_TRAIN_ENCODER = graph (default: one-call passes
    replayed from CUDA graphs) | fused (one-call passes, launched eagerly) | layers (13 SparseConvBN autograd
    nodes; what the unit tests dissect)."""
    import os
    if G is None:
        ops.encoder_build_maps(ws, coords0)
        G = EncoderGraph(ws)
    layers = net._layers()
    mode = os.environ.get('IR_TRAIN_ENCODER', 'graph')
    if mode != 'layers':
        flat = [t for conv, bn in layers for t in (conv.kernel, bn.weight, bn.bias)]
        if mode == 'graph' and all(t.is_contiguous() for t in flat):
            return EncoderTrainGraphed.apply(feats0, net, G, *flat), G
        return EncoderTrain.apply(feats0, net, G, *flat), G
    if feats0 is None:
        feats0 = ws.feat0(net.input_dim)[:G.n[0]].clone()
    tc = net.use_tc

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


# ----------------------------------------------------------------------------- dense operators

class Linear(Function):
    """nn.Linear (+ReLU): y = act(x W^T + b), W (N,K) in the reference layout."""

    @staticmethod
    def forward(ctx, x, W, b, relu=False):
        x = x.contiguous()
        y = ops.gemm(x, W.detach(), tb=True, bias=None if b is None else b.detach(), relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x, W.detach(), y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        g = ops.relu_bwd(dy.contiguous(), y) if ctx.relu else dy.contiguous()
        dx = ops.gemm(g, W) if ctx.needs_input_grad[0] else None          # (M,N) @ (N,K)
        dW = ops.gemm(g, x, ta=True)                                      # g^T @ x -> (N,K)
        db = ops.colsum(g) if ctx.needs_input_grad[2] else None
        return dx, dW, db, None


class LayerNormAct(Function):
    """nn.LayerNorm (+ReLU)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, relu):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, gamma.detach(), beta.detach(), eps, relu)
        ctx.relu = relu
        ctx.save_for_backward(x, gamma.detach(), y, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(dy.contiguous(), y, x, gamma, mean, rstd, ctx.relu)
        return dx, dg, db, None, None


class BatchNormAct(Function):
    """nn.BatchNorm1d / BatchNorm2d (NHWC rows) in train mode (+ReLU): batch statistics, running
    statistics updated in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu):
        x = x.contiguous()
        mom = bn.momentum if bn.momentum is not None else 0.1
        y, mean, rstd = ops.bn_train_fwd(x, gamma.detach(), beta.detach(), None, relu, bn.eps, mom,
                                         bn.running_mean, bn.running_var)
        bn.num_batches_tracked += 1
        ctx.relu = relu
        ctx.save_for_backward(x, gamma.detach(), y, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y, mean, rstd = ctx.saved_tensors
        dx, _, dg, db = ops.bn_train_bwd(dy.contiguous(), y, x, mean, rstd, gamma, ctx.relu, False)
        return dx, dg, db, None, None


_dropout_calls = [0]


class Dropout(Function):
    @staticmethod
    def forward(ctx, x, p):
        _dropout_calls[0] += 1
        seed = (torch.initial_seed() * 0x9E3779B1 + _dropout_calls[0] * 0x85EBCA6B) & (2 ** 63 - 1)
        y, mask = ops.dropout_fwd(x.contiguous(), p, seed)
        ctx.p = p
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        return ops.dropout_bwd(dy.contiguous(), mask, ctx.p), None


def dropout(x, module):
    """nn.Dropout in train mode (p = 0, as parity tests set it, is the identity)."""
    return Dropout.apply(x, float(module.p)) if module.p > 0 else x


class L2Norm(Function):
    """F.normalize(x, p=2, dim=1)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.l2norm_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.l2norm_bwd(dy.contiguous(), x)


class Match(Function):
    """score[r] of candidate row r against its scene's partner row: mode 0 = normalise + dot
    (models/attribute_module.py:113-126), mode 1 = cosine similarity (relation_module.py:103,
    scene_module.py:104)."""

    @staticmethod
    def forward(ctx, a, partner, seg, row_ofs, mode):
        a, partner = a.contiguous(), partner.contiguous()
        ctx.mode = mode
        ctx.save_for_backward(a, partner, seg, row_ofs)
        return ops.match_fwd(a, partner, seg, mode)

    @staticmethod
    def backward(ctx, ds):
        a, partner, seg, row_ofs = ctx.saved_tensors
        da, dp = ops.match_bwd(ds.contiguous(), a, partner, seg, row_ofs, ctx.mode)
        return da, dp, None, None, None


class Conv3x3(Function):
    """nn.Conv2d(C, Cout, 3) (valid) on NHWC activations as im2col + GEMM; weight in the reference
    layout (Cout, Cin, 3, 3)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        B, H, W, Cc = x.shape
        wp = weight.detach().permute(2, 3, 1, 0).reshape(9 * Cc, -1).contiguous()      # [ky][kx][Cin] x Cout
        col = ops.im2col_3x3(x)
        y = ops.gemm(col, wp, bias=bias.detach())
        ctx.shape = (B, H, W, Cc)
        ctx.save_for_backward(col, wp)
        return y.view(B, H - 2, W - 2, -1)

    @staticmethod
    def backward(ctx, dy):
        col, wp = ctx.saved_tensors
        B, H, W, Cc = ctx.shape
        g = dy.contiguous().view(-1, wp.shape[1])
        dx = ops.col2im_3x3(ops.gemm(g, wp, tb=True), B, H, W, Cc) if ctx.needs_input_grad[0] else None
        dwp = ops.gemm(col, g, ta=True)                                               # (9C, Cout)
        dW = dwp.view(3, 3, Cc, -1).permute(3, 2, 0, 1).contiguous()
        return dx, dW, ops.colsum(g)


class BEV(Function):
    """SparseCrop + ToDenseBEVConvolution (models/basic_blocks.py:174-243) -> (B*375, 128) raw sums."""

    @staticmethod
    def forward(ctx, f4, kernel, coords, n_dev, n_rows, B):
        f4 = f4.contiguous()
        k = kernel.detach().contiguous()
        dense, cell = ops.bev_raw(f4, coords, n_dev, n_rows, 16, k, B)
        ctx.n_rows = n_rows
        ctx.save_for_backward(f4, k, coords, n_dev, cell)
        return dense

    @staticmethod
    def backward(ctx, dd):
        f4, k, coords, n_dev, cell = ctx.saved_tensors
        df, dk = ops.bev_bwd(dd.contiguous(), f4, coords, cell, n_dev, ctx.n_rows, 16, k)
        return df, dk, None, None, None, None


class SceneAttention(Function):
    """softmax_cells(<f, q>/sqrt(C)) weighted sum (models/scene_module.py:73-83) -> (scene_feats, atten)."""

    @staticmethod
    def forward(ctx, feats, q):
        feats, q = feats.contiguous(), q.contiguous()
        atten, sf = ops.scene_attention(feats, q)
        ctx.save_for_backward(feats, q, atten)
        ctx.mark_non_differentiable(atten)
        return sf, atten

    @staticmethod
    def backward(ctx, dsf, _da):
        feats, q, atten = ctx.saved_tensors
        return ops.scene_attention_bwd(feats, q, atten, dsf.contiguous())


class TokenAttention(Function):
    """Four masked attention poolings (models/lang_module.py:60-83) -> (pooled (4,B,E), atten (4,B,L))."""

    @staticmethod
    def forward(ctx, feats, embed, lengths, fcw, fcb):
        feats, embed = feats.contiguous(), embed.contiguous()
        atten, pooled = ops.token_attention(feats, embed, lengths, fcw.detach().contiguous(), fcb.detach().contiguous())
        ctx.save_for_backward(feats, embed, lengths, fcw.detach().contiguous(), fcb.detach().contiguous(), atten)
        ctx.mark_non_differentiable(atten)
        return pooled, atten

    @staticmethod
    def backward(ctx, dpooled, _da):
        feats, embed, lengths, fcw, fcb, atten = ctx.saved_tensors
        df, de, dw, db = ops.token_attention_bwd(feats, embed, lengths, fcw, fcb, atten, dpooled.contiguous())
        return df, de, None, dw, db


class GRULayer(Function):
    """One packed bidirectional GRU layer (models/lang_module.py:53-57): xproj (B*L, 2*3H) hoisted input
    projections, whh (2,3H,H), bhh (2,3H) -> out (B*L, 2H)."""

    @staticmethod
    def forward(ctx, xproj, whh, bhh, lengths, B, L):
        xproj = xproj.contiguous()
        w, b = whh.detach().contiguous(), bhh.detach().contiguous()
        out = ops.gru_layer(xproj, w, b, lengths, B, L)
        ctx.BL = (B, L)
        ctx.save_for_backward(xproj, w, b, lengths, out)
        return out.view(B * L, -1)

    @staticmethod
    def backward(ctx, dout):
        xproj, w, b, lengths, out = ctx.saved_tensors
        B, L = ctx.BL
        dxp, dw, db = ops.gru_layer_bwd(xproj, w, b, lengths, out, dout.contiguous(), B, L)
        return dxp, dw, db, None, None, None


class EdgeConcat(Function):
    """e_in = [x_i, w, x_j] per edge (models/basic_blocks.py:132); only w carries a gradient."""

    @staticmethod
    def forward(ctx, w, x, xyz, qidx, nbr, ncls):
        ctx.F = x.shape[1]
        ctx.save_for_backward(nbr)
        return ops.edge_inputs(x, xyz, qidx, nbr, ncls, w=w.contiguous())

    @staticmethod
    def backward(ctx, de):
        F = ctx.F
        (nbr,) = ctx.saved_tensors
        return de[:, F:2 * F] * (nbr.reshape(-1, 1) >= 0), None, None, None, None, None


class EdgeMax(Function):
    """max aggregation over the incoming edges of each query (MessagePassing aggr='max')."""

    @staticmethod
    def forward(ctx, msg, nbr):
        out, arg = ops.edge_max_fwd(msg.contiguous(), nbr)
        ctx.k = nbr.shape[1]
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        return ops.edge_max_bwd(dout.contiguous(), arg, ctx.k), None


class MLPHead(Function):
    """Linear -> {BatchNorm1d | LayerNorm} -> ReLU [-> Dropout] -> Linear of the reference's nn.Sequential heads
    as ONE library call per direction (ir_mlp_head_train_fwd / _bwd); intermediates stay in an arena owned by this
    node of the autograd graph."""

    @staticmethod
    def forward(ctx, x, w1, b1, g, beta, w2, b2, nm, drop_p):
        import ctypes as C
        from . import _lib
        x = x.contiguous()
        M, K = x.shape
        N1, N2 = w1.shape[0], w2.shape[0]
        H = _lib.MlpHead()
        H.M, H.K, H.N1, H.N2 = M, K, N1, N2
        H.norm = 2 if isinstance(nm, torch.nn.LayerNorm) else 1
        H.eps, H.drop_p = nm.eps, float(drop_p)
        keep = [t.detach().contiguous() for t in (w1, b1, g, beta, w2, b2)]
        H.w1, H.b1, H.gamma, H.beta, H.w2, H.b2 = (t.data_ptr() for t in keep)
        if H.norm == 1:
            H.momentum = nm.momentum if nm.momentum is not None else 0.1
            H.running_mean, H.running_var = nm.running_mean.data_ptr(), nm.running_var.data_ptr()
        if drop_p > 0:
            _dropout_calls[0] += 1
            H.seed = (torch.initial_seed() * 0x9E3779B1 + _dropout_calls[0] * 0x85EBCA6B) & (2 ** 63 - 1)
        arena = torch.empty(_lib.load().ir_mlp_head_arena_bytes(M, N1), dtype=torch.uint8, device=x.device)
        y = torch.empty(M, N2, dtype=torch.float32, device=x.device)
        _lib.call("ir_mlp_head_train_fwd", C.byref(H), ops._p(x), ops._p(arena), ops._p(y), ops._stream())
        if H.norm == 1:
            nm.num_batches_tracked += 1
        ctx.state = (H, keep, arena, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C
        from . import _lib
        H, keep, arena, x = ctx.state
        dev = x.device
        M, K, N1, N2 = H.M, H.K, H.N1, H.N2
        sizes = [N1 * K, N1, N1, N1, N2 * N1, N2]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 63) // 64 * 64
        flat = torch.empty(o, dtype=torch.float32, device=dev)
        dw1, db1, dg, dbeta, dw2, db2 = (flat[a:a + n] for a, n in zip(offs, sizes))
        dx = torch.empty(M, K, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        _lib.call("ir_mlp_head_train_bwd", C.byref(H), ops._p(x), ops._p(arena), ops._p(dy.contiguous(), torch.float32), ops._p(dx),
                  ops._p(dw1), ops._p(db1), ops._p(dg), ops._p(dbeta), ops._p(dw2), ops._p(db2), ops._stream())
        return dx, dw1.view(N1, K), db1, dg, dbeta, dw2.view(N2, N1), db2, None, None


class EdgeConvTrain(Function):
    """DynamicEdgeConv (edge-weight MLP, message MLP, max aggregation) as ONE library call per direction.
    Parameter order: weight.0.{weight,bias}, weight.2.{weight,bias}, mlp.0.{weight,bias}, mlp.2.{weight,bias}."""

    @staticmethod
    def forward(ctx, feats, xyz, qidx, nbr, ncls, *params):
        import ctypes as C
        from . import _lib
        keep = [t.detach().contiguous() for t in params]
        P = _lib.EdgeConvParams()
        P.nq, P.k, P.F, P.ncls = nbr.shape[0], nbr.shape[1], feats.shape[1], ncls
        P.H1, P.Fout = keep[0].shape[0], keep[6].shape[0]
        P.ww1, P.bw1, P.ww2, P.bw2, P.wm1, P.bm1, P.wm2, P.bm2 = (t.data_ptr() for t in keep)
        dev = feats.device
        arena = torch.empty(_lib.load().ir_edgeconv_train_arena_bytes(C.byref(P)), dtype=torch.uint8, device=dev)
        out = torch.empty(P.nq, P.Fout, dtype=torch.float32, device=dev)
        _lib.call("ir_edgeconv_train_fwd", C.byref(P), ops._p(feats, torch.float32), ops._p(xyz, torch.float32),
                  ops._p(qidx, torch.int32), ops._p(nbr, torch.int32), ops._p(arena), ops._p(out), ops._stream())
        ctx.state = (P, keep, arena, [tuple(t.shape) for t in params])
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        from . import _lib
        P, keep, arena, shapes = ctx.state
        numel = [int(torch.Size(sh).numel()) for sh in shapes]
        offs, o = [], 0
        for n in numel:
            offs.append(o)
            o += (n + 63) // 64 * 64
        flat = torch.empty(o, dtype=torch.float32, device=arena.device)
        G = _lib.EdgeConvGrads()
        G.dww1, G.dbw1, G.dww2, G.dbw2, G.dwm1, G.dbm1, G.dwm2, G.dbm2 = (flat.data_ptr() + 4 * a for a in offs)
        ops.stamp('bwd:edgeconv:start')
        _lib.call("ir_edgeconv_train_bwd", C.byref(P), ops._p(arena), ops._p(dout.contiguous(), torch.float32), C.byref(G), ops._stream())
        ops.stamp('bwd:edgeconv:end')
        return (None, None, None, None, None, *(flat[a:a + n].view(sh) for a, n, sh in zip(offs, numel, shapes)))


class SceneTailTrain(Function):
    """models/scene_module.py:25-38,70-71 in train mode as ONE library call per direction (ir_scene_tail_train_fwd /
    _bwd): BEV -> BN -> ReLU -> Conv3x3 -> BN -> ReLU -> Dropout -> Conv3x3.  Parameter order: to_bev.1.kernel,
    to_bev.2.{weight,bias}, vis_emb_fc.0.{weight,bias}, vis_emb_fc.1.{weight,bias}, vis_emb_fc.4.{weight,bias}."""

    @staticmethod
    def forward(ctx, f4, coords, n_dev, n_rows, B, bn0, bn1, drop_p, *params):
        import ctypes as C
        from . import _lib
        f4 = f4.contiguous()
        dev = f4.device
        keep = [t.detach().contiguous() for t in params]
        P = _lib.SceneTail()
        P.B, P.n_rows, P.eps, P.drop_p = B, int(n_rows), bn0.eps, float(drop_p)
        P.mom0 = bn0.momentum if bn0.momentum is not None else 0.1
        P.mom1 = bn1.momentum if bn1.momentum is not None else 0.1
        P.kernel, P.g0, P.be0, P.w1, P.b1, P.g1, P.be1, P.w2, P.b2 = (t.data_ptr() for t in keep)
        P.rm0, P.rv0 = bn0.running_mean.data_ptr(), bn0.running_var.data_ptr()
        P.rm1, P.rv1 = bn1.running_mean.data_ptr(), bn1.running_var.data_ptr()
        if drop_p > 0:
            _dropout_calls[0] += 1
            P.seed = (torch.initial_seed() * 0x9E3779B1 + _dropout_calls[0] * 0x85EBCA6B) & (2 ** 63 - 1)
        arena = torch.empty(_lib.load().ir_scene_tail_arena_bytes(int(n_rows), B), dtype=torch.uint8, device=dev)
        out = torch.empty(B, 11, 21, 128, dtype=torch.float32, device=dev)
        _lib.call("ir_scene_tail_train_fwd", C.byref(P), ops._p(f4, torch.float32), ops._p(coords, torch.int32),
                  ops._p(n_dev, torch.int32), ops._p(arena), ops._p(out), ops._stream())
        bn0.num_batches_tracked += 1
        bn1.num_batches_tracked += 1
        ctx.state = (P, keep, arena, f4, coords, n_dev, [tuple(t.shape) for t in params])
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        from . import _lib
        P, keep, arena, f4, coords, n_dev, shapes = ctx.state
        ops.stamp('bwd:scene_tail:start')
        numel = [int(torch.Size(sh).numel()) for sh in shapes]
        offs, o = [], 0
        for n in numel:
            offs.append(o)
            o += (n + 63) // 64 * 64
        flat = torch.empty(o, dtype=torch.float32, device=f4.device)
        G = _lib.SceneTailGrads()
        G.dkernel, G.dg0, G.dbe0, G.dw1, G.db1, G.dg1, G.dbe1, G.dw2, G.db2 = (flat.data_ptr() + 4 * a for a in offs)
        df4 = torch.empty_like(f4)
        _lib.call("ir_scene_tail_train_bwd", C.byref(P), ops._p(f4), ops._p(coords), ops._p(n_dev), ops._p(arena),
                  ops._p(dout.contiguous(), torch.float32), ops._p(df4), C.byref(G), ops._stream())
        ops.stamp('bwd:scene_tail:end')
        return (df4, None, None, None, None, None, None, None, *(flat[a:a + n].view(sh) for a, n, sh in zip(offs, numel, shapes)))


def mlp_head(seq, x, norm_idx, last_idx, drop_idx=None):
    """Linear -> {BatchNorm1d | LayerNorm} -> ReLU [-> Dropout] -> Linear of the reference's
    nn.Sequential heads (e.g. models/relation_module.py:13-25).  IR_TRAIN_HEADS=ops runs it as separate
    autograd nodes (Linear / norm / Dropout / Linear: what the per-op tests cover)."""
    import os
    nm = seq[norm_idx]
    if os.environ.get('IR_TRAIN_HEADS', 'fused') != 'ops':
        p = float(seq[drop_idx].p) if drop_idx is not None else 0.0
        return MLPHead.apply(x, seq[0].weight, seq[0].bias, nm.weight, nm.bias, seq[last_idx].weight, seq[last_idx].bias, nm, p)
    h = Linear.apply(x, seq[0].weight, seq[0].bias, False)
    if isinstance(nm, torch.nn.LayerNorm):
        h = LayerNormAct.apply(h, nm.weight, nm.bias, nm.eps, True)
    else:
        h = BatchNormAct.apply(h, nm.weight, nm.bias, nm, True)
    if drop_idx is not None:
        h = dropout(h, seq[drop_idx])
    return Linear.apply(h, seq[last_idx].weight, seq[last_idx].bias, False)


# ----------------------------------------------------------------------------- module forwards

class LangTrain(Function):
    """The whole language encoder in train mode as ONE library call per direction (ir_lang_train_fwd / _bwd).
    Parameter order: word_projection.{0,3}.{weight,bias}; for layer 0,1 / direction fwd,rev: weight_ih, weight_hh,
    bias_ih, bias_hh; fc_a, fc_cls, fc_rel, fc_scene {weight,bias}; lang_cls.0.{weight,bias}."""

    @staticmethod
    def forward(ctx, x, lengths, B, L, drop_p, *params):
        import ctypes as C
        from . import _lib
        x = x.contiguous()
        dev = x.device
        keep = [t.detach().contiguous() for t in params]
        P = _lib.LangParams()
        P.B, P.L, P.E_in, P.D, P.H, P.n_cls = B, L, x.shape[1], 256, 128, keep[28].shape[0]
        P.drop_p = float(drop_p)
        if drop_p > 0:
            _dropout_calls[0] += 1
            P.seed = (torch.initial_seed() * 0x9E3779B1 + _dropout_calls[0] * 0x85EBCA6B) & (2 ** 63 - 1)
        P.w0, P.b0, P.w3, P.b3 = (t.data_ptr() for t in keep[0:4])
        for l in range(2):
            for d in range(2):
                wih, whh, bih, bhh = keep[4 + 8 * l + 4 * d: 8 + 8 * l + 4 * d]
                P.wih[l][d], P.whh[l][d], P.bih[l][d], P.bhh[l][d] = wih.data_ptr(), whh.data_ptr(), bih.data_ptr(), bhh.data_ptr()
        for h in range(4):
            P.fcw[h], P.fcb[h] = keep[20 + 2 * h].data_ptr(), keep[21 + 2 * h].data_ptr()
        P.wc, P.bc = keep[28].data_ptr(), keep[29].data_ptr()
        lib = _lib.load()
        arena = torch.empty(lib.ir_lang_train_arena_bytes(C.byref(P)), dtype=torch.uint8, device=dev)
        pooled = torch.empty(4, B, 256, dtype=torch.float32, device=dev)
        scores = torch.empty(B, P.n_cls, dtype=torch.float32, device=dev)
        _lib.call("ir_lang_train_fwd", C.byref(P), ops._p(x), ops._p(lengths, torch.int64), ops._p(arena), ops._p(pooled),
                  ops._p(scores), ops._stream())
        of, oa = C.c_int64(0), C.c_int64(0)
        _lib.call("ir_lang_train_view", C.byref(P), C.byref(of), C.byref(oa))
        feats = arena[of.value:of.value + B * L * 256 * 4].view(torch.float32).view(B, L, 256)
        atten = arena[oa.value:oa.value + 4 * B * L * 4].view(torch.float32).view(4, B, L)
        ctx.state = (P, keep, arena, x, lengths, [tuple(t.shape) for t in params])
        ctx.save_for_backward(pooled)        # an OUTPUT: kept as a plain attribute it would tie node and tensor into a cycle
        ctx.mark_non_differentiable(feats, atten)
        return pooled, scores, feats, atten

    @staticmethod
    def backward(ctx, dpooled, dscores, _df, _da):
        import ctypes as C
        from . import _lib
        P, keep, arena, x, lengths, shapes = ctx.state
        ops.stamp('bwd:lang:start')
        (pooled,) = ctx.saved_tensors
        dev = x.device
        H3 = 3 * 128
        # flat gradient buffer; the per-direction bias gradients of a layer and the four fc_* gradients are
        # contiguous blocks that the library fills with one column sum each
        numel = [int(torch.Size(sh).numel()) for sh in shapes]
        order = list(range(30))
        offs, o = {}, 0

        def seg(i):
            nonlocal o
            offs[i] = o
            o += (numel[i] + 63) // 64 * 64
        for i in (0, 1, 2, 3):
            seg(i)
        for l in range(2):
            base = 4 + 8 * l
            seg(base + 0); seg(base + 1); seg(base + 4); seg(base + 5)               # wih, whh (fwd / rev)
            offs[base + 2] = o; offs[base + 6] = o + H3; o += 2 * H3 + 0               # bih fwd | rev contiguous
            o = (o + 63) // 64 * 64
            offs[base + 3] = o; offs[base + 7] = o + H3; o += 2 * H3
            o = (o + 63) // 64 * 64
        for h in range(4):                                                           # fc weights (4,D) contiguous
            offs[20 + 2 * h] = o + 256 * h
        o += 4 * 256
        for h in range(4):                                                           # fc biases (4) contiguous
            offs[21 + 2 * h] = o + h
        o += 64
        seg(28); seg(29)
        flat = torch.empty(o, dtype=torch.float32, device=dev)
        ptr = lambda i: flat.data_ptr() + 4 * offs[i]
        G = _lib.LangGrads()
        G.dw0, G.db0, G.dw3, G.db3 = ptr(0), ptr(1), ptr(2), ptr(3)
        for l in range(2):
            base = 4 + 8 * l
            G.dwih[l][0], G.dwhh[l][0], G.dwih[l][1], G.dwhh[l][1] = ptr(base), ptr(base + 1), ptr(base + 4), ptr(base + 5)
            G.dbih[l], G.dbhh[l] = ptr(base + 2), ptr(base + 3)
        G.dfcw, G.dfcb, G.dwc, G.dbc = ptr(20), ptr(21), ptr(28), ptr(29)
        ds = dscores.contiguous() if dscores is not None else None
        _lib.call("ir_lang_train_bwd", C.byref(P), ops._p(x), ops._p(lengths, torch.int64), ops._p(arena), ops._p(pooled),
                  ops._p(dpooled.contiguous(), torch.float32), ops._p(ds), C.byref(G), ops._stream())
        grads = [flat[offs[i]:offs[i] + numel[i]].view(shapes[i]) for i in order]
        ops.stamp('bwd:lang:end')
        return (None, None, None, None, None, *grads)


def lang_params(m):
    g = m.gru
    ps = [m.word_projection[0].weight, m.word_projection[0].bias, m.word_projection[3].weight, m.word_projection[3].bias]
    for l in (0, 1):
        for sfx in ('', '_reverse'):
            ps += [getattr(g, f'weight_ih_l{l}{sfx}'), getattr(g, f'weight_hh_l{l}{sfx}'),
                   getattr(g, f'bias_ih_l{l}{sfx}'), getattr(g, f'bias_hh_l{l}{sfx}')]
    for fc in (m.fc_a, m.fc_cls, m.fc_rel, m.fc_scene):
        ps += [fc.weight, fc.bias]
    ps += [m.lang_cls[0].weight, m.lang_cls[0].bias]
    return ps


def lang_forward_train(m, data_dict):
    """models/lang_module.py:51-108 in train mode: word MLP (+Dropout), 2-layer packed biGRU, four
    masked attention pools over the PROJECTED embeddings, classifier.  One library call per direction
    (LangTrain); IR_TRAIN_LANG=ops runs the chain as separate autograd nodes (what the per-op tests cover)."""
    import os
    x, length = data_dict['lang_feat'], data_dict['lang_len']
    dev = x.device
    host = data_dict.get('_ir_host_labels', {}).get('lang_len')
    L = data_dict.get('_ir_lang_len_max')
    if L is None:
        L = int((host if host is not None else length.detach().to('cpu')).max())
    B = x.shape[0]
    len_dev = length.to(dev, torch.int64).contiguous()
    wp = m.word_projection
    if os.environ.get('IR_TRAIN_LANG', 'fused') != 'ops' and m.use_lang_classifier and m.use_bidir and m.hidden_size == 128:
        pooled, scores, feats, atten = LangTrain.apply(x[:, :L].float().reshape(B * L, -1), len_dev, B, L, float(wp[2].p),
                                                       *lang_params(m))
        data_dict['lang_feat'] = feats
        data_dict['atten_attr'], data_dict['atten_rel'], data_dict['atten_scene'] = atten[0], atten[2], atten[3]
        data_dict['lang_attr_feats'], data_dict['lang_cls_feats'] = pooled[0], pooled[1]
        data_dict['lang_rel_feats'], data_dict['lang_scene_feats'] = pooled[2], pooled[3]
        data_dict['lang_scores'] = scores
        return data_dict
    h = Linear.apply(x[:, :L].float().reshape(B * L, -1), wp[0].weight, wp[0].bias, True)
    h = dropout(h, wp[2])
    e = Linear.apply(h, wp[3].weight, wp[3].bias, True)                                     # (B*L, 256)
    g = m.gru
    h = e
    for l in (0, 1):
        wih = torch.cat([getattr(g, f'weight_ih_l{l}'), getattr(g, f'weight_ih_l{l}_reverse')], 0)
        bih = torch.cat([getattr(g, f'bias_ih_l{l}'), getattr(g, f'bias_ih_l{l}_reverse')], 0)
        whh = torch.stack([getattr(g, f'weight_hh_l{l}'), getattr(g, f'weight_hh_l{l}_reverse')], 0)
        bhh = torch.stack([getattr(g, f'bias_hh_l{l}'), getattr(g, f'bias_hh_l{l}_reverse')], 0)
        xp = Linear.apply(h, wih, bih, False)                                               # hoisted input GEMM
        h = GRULayer.apply(xp, whh, bhh, len_dev, B, L)
    feats = h.view(B, L, -1)
    data_dict['lang_feat'] = feats
    fcw = torch.cat([m.fc_a.weight, m.fc_cls.weight, m.fc_rel.weight, m.fc_scene.weight], 0)
    fcb = torch.cat([m.fc_a.bias, m.fc_cls.bias, m.fc_rel.bias, m.fc_scene.bias], 0)
    pooled, atten = TokenAttention.apply(feats, e.view(B, L, -1), len_dev, fcw, fcb)
    data_dict['atten_attr'], data_dict['atten_rel'], data_dict['atten_scene'] = atten[0], atten[2], atten[3]
    data_dict['lang_attr_feats'], data_dict['lang_cls_feats'] = pooled[0], pooled[1]
    data_dict['lang_rel_feats'], data_dict['lang_scene_feats'] = pooled[2], pooled[3]
    if m.use_lang_classifier:
        data_dict['lang_scores'] = Linear.apply(pooled[1], m.lang_cls[0].weight, m.lang_cls[0].bias, False)
    return data_dict


def prepare_attribute_maps(model, pack):
    """Coordinate phase of the instance encoder: voxelise the candidates @2 cm, levels + kernel maps -> workspace."""
    am = model.attribute
    ws_a = am.net.workspace(pack.M * pack.points.shape[1], pack.points.device)
    ops.encoder_reset(ws_a)
    ops.voxelize(pack.points, pack.cand_rows, float(am.voxel_size[0]), ws_a)
    ops.encoder_build_maps(ws_a)
    return ws_a


def prepare_scene_maps(model, data_dict, dev):
    """Coordinate phase of the scene encoder: hash the loader's 5 cm voxels, levels + kernel maps."""
    lidar = data_dict['lidar']
    F0 = lidar.F.to(dev, torch.float32).contiguous()
    C0 = lidar.C.to(dev, torch.int32).contiguous()
    ws_s = model.scene.net.workspace(F0.shape[0], dev)
    ops.encoder_build_maps(ws_s, C0, data_dict.get('_ir_lidar_rows'))
    return ws_s, F0, C0


def prepare_encoder_maps(model, data_dict, pack):
    """Coordinate phase of BOTH sparse encoders, then ONE D2H copy of the ten level row counts — the only host
    synchronisation of the train-mode forward after the token-length read.  ``_ir_capacity`` in the dict (set by
    train_graph.GraphedTrainStep): no read-back at all — every buffer downstream is laid out for the workspace
    capacity and the row counts are read on the device, which is what makes the step capturable."""
    ws_a = prepare_attribute_maps(model, pack)
    ws_s, F0, C0 = prepare_scene_maps(model, data_dict, pack.points.device)
    if data_dict.get('_ir_capacity'):
        n = [ws_a.n_max] * 5 + [ws_s.n_max] * 5
    else:
        n = torch.cat([ws_a.nlvl(), ws_s.nlvl()]).tolist()
    return (ws_a, EncoderGraph(ws_a, n[:5])), (ws_s, EncoderGraph(ws_s, n[5:]), F0, C0)


def attribute_encode_train(m, data_dict, pack, prepared=None):
    """models/attribute_module.py:83-111 in train mode: the language-independent half (encoder, pooling, visual head)."""
    dev = pack.points.device
    data_dict['num_filtered_objs'] = pack.num_filtered
    data_dict['pred_obb_batch'] = pack.pred_obb_batch
    if prepared is None:
        ws = m.net.workspace(pack.M * pack.points.shape[1], dev)
        ops.encoder_reset(ws)
        ops.voxelize(pack.points, pack.cand_rows, float(m.voxel_size[0]), ws)
        f4, G = encoder_forward_train(m.net, ws)
    else:
        ws, G = prepared
        f4, G = encoder_forward_train(m.net, ws, G=G)
    obj = SegMax.apply(f4, ws.coords(4), G.nlvl_dev[4:5], G.n[4], pack.M)
    data_dict['obj_feats'] = obj
    data_dict['_ir_attr_vis'] = mlp_head(m.vis_emb_fc, obj, 1, 3)
    return data_dict


def attribute_match_train(m, data_dict, pack):
    """models/attribute_module.py:113-131: language embedding + normalised dot product."""
    lang = L2Norm.apply(mlp_head(m.lang_emb_fc, data_dict['lang_attr_feats'], 1, 3))
    data_dict['attribute_scores'] = Match.apply(data_dict.pop('_ir_attr_vis'), lang, pack.cand_scene, pack.scene_ofs(), 0)
    return data_dict


def attribute_forward_train(m, data_dict, pack, prepared=None):
    """models/attribute_module.py:83-131 in train mode."""
    return attribute_match_train(m, attribute_encode_train(m, data_dict, pack, prepared), pack)


def relation_encode_train(m, data_dict, pack):
    """models/relation_module.py:80-100 in train mode: kNN graph, edge MLPs as edge-batched GEMMs, max aggregation,
    visual head (the language-independent half)."""
    dev = pack.points.device
    mean = ops.instance_mean(pack.points)
    ncls = m.args.num_classes
    onehot = (pack.centres_cls[:, 3:4] == torch.arange(ncls, device=dev, dtype=torch.float32)[None, :]).float()
    xyz = pack.centres_cls[:, :3].contiguous()
    feats = torch.cat([xyz, mean[:, 3:], onehot], 1).contiguous()
    gcn = m.gcn
    nbr = ops.knn(xyz, pack.inst_ofs, pack.cand_rows, pack.cand_seg, gcn.k)                  # (M,k), -1 padded
    if os.environ.get('IR_TRAIN_EDGECONV', 'fused') != 'ops':
        g = EdgeConvTrain.apply(feats, xyz, pack.cand_rows, nbr, ncls, gcn.weight[0].weight, gcn.weight[0].bias,
                                gcn.weight[2].weight, gcn.weight[2].bias, gcn.mlp[0].weight, gcn.mlp[0].bias,
                                gcn.mlp[2].weight, gcn.mlp[2].bias)
    else:
        w_in = ops.edge_inputs(feats, xyz, pack.cand_rows, nbr, ncls)
        w = Linear.apply(Linear.apply(w_in, gcn.weight[0].weight, gcn.weight[0].bias, True),
                         gcn.weight[2].weight, gcn.weight[2].bias, False)
        e_in = EdgeConcat.apply(w, feats, xyz, pack.cand_rows, nbr, ncls)
        msg = Linear.apply(Linear.apply(e_in, gcn.mlp[0].weight, gcn.mlp[0].bias, True),
                           gcn.mlp[2].weight, gcn.mlp[2].bias, False)
        g = EdgeMax.apply(msg, nbr)
    data_dict['_ir_rel_vis'] = mlp_head(m.vis_emb_fc, g, 1, 4, 3)
    return data_dict


def relation_match_train(m, data_dict, pack):
    """models/relation_module.py:101-107: language embedding + cosine match."""
    lang = mlp_head(m.lang_emb_fc, data_dict['lang_rel_feats'], 1, 4, 3)
    data_dict['relation_scores'] = Match.apply(data_dict.pop('_ir_rel_vis'), lang, pack.cand_scene, pack.scene_ofs(), 1)
    return data_dict


def relation_forward_train(m, data_dict, pack):
    """models/relation_module.py:80-107 in train mode."""
    return relation_match_train(m, relation_encode_train(m, data_dict, pack), pack)


def scene_encode_train(m, data_dict, pack, prepared=None):
    """models/scene_module.py:60-71 in train mode: scene encoder + BEV + the two Conv2d blocks (language-independent)."""
    dev = pack.points.device
    B = data_dict['point_min'].shape[0]
    if prepared is None:
        lidar = data_dict['lidar']
        F0 = lidar.F.to(dev, torch.float32).contiguous()
        C0 = lidar.C.to(dev, torch.int32).contiguous()
        ws = m.net.workspace(F0.shape[0], dev)
        f4, G = encoder_forward_train(m.net, ws, F0, C0)
    else:
        ws, G, F0, C0 = prepared
        f4, G = encoder_forward_train(m.net, ws, F0, C0, G=G)
    bn = m.to_bev[2]
    ve = m.vis_emb_fc
    if os.environ.get('IR_TRAIN_SCENE', 'fused') != 'ops' and G.n[4] > 0:
        x = SceneTailTrain.apply(f4, ws.coords(4), G.nlvl_dev[4:5], G.n[4], B, bn, ve[1], float(ve[3].p),
                                 m.to_bev[1].kernel, bn.weight, bn.bias, ve[0].weight, ve[0].bias,
                                 ve[1].weight, ve[1].bias, ve[4].weight, ve[4].bias)       # (B,11,21,128)
    else:                                                # separate autograd nodes (what the per-op tests cover)
        dense = BEV.apply(f4, m.to_bev[1].kernel, ws.coords(4), G.nlvl_dev[4:5], G.n[4], B)     # (B*375,128)
        x = BatchNormAct.apply(dense, bn.weight, bn.bias, bn, True).view(B, 15, 25, -1)          # NHWC
        x = Conv3x3.apply(x, ve[0].weight, ve[0].bias)
        x = BatchNormAct.apply(x.reshape(-1, x.shape[-1]), ve[1].weight, ve[1].bias, ve[1], True).view(x.shape)
        x = dropout(x, ve[3])
        x = Conv3x3.apply(x, ve[4].weight, ve[4].bias)                                           # (B,11,21,128)
    data_dict['_ir_scene_map'] = x
    return data_dict


def scene_match_train(m, data_dict, pack):
    """models/scene_module.py:73-108: language query, attention over the BEV cells, region classifier, cosine match
    of the candidates' pooled features against the attended scene feature."""
    x = data_dict.pop('_ir_scene_map')
    B, h, w = x.shape[0], x.shape[1], x.shape[2]
    q = mlp_head(m.lang_emb_fc, data_dict['lang_scene_feats'], 1, 4, 3)
    scene_feats, atten = SceneAttention.apply(x.view(B, h * w, -1), q)
    data_dict['vis_atten'] = atten.view(B, h, w)
    data_dict['seg_scores'] = mlp_head(m.cls, scene_feats, 1, 3)
    obj = mlp_head(m.vis_emb_fc1, data_dict['obj_feats'], 1, 4, 3)
    data_dict['scene_scores'] = Match.apply(obj, scene_feats, pack.cand_scene, pack.scene_ofs(), 1)
    return data_dict


def scene_forward_train(m, data_dict, pack, prepared=None):
    """models/scene_module.py:60-108 in train mode."""
    return scene_match_train(m, scene_encode_train(m, data_dict, pack, prepared), pack)
