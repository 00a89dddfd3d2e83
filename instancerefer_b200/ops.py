"""torch-tensor front end of the C ABI: pointer extraction, workspace ownership, stream plumbing.
PyTorch is used only for device memory and streams; all arithmetic happens in the CUDA library."""
import ctypes as C

import torch

from . import _lib
from ._lib import EncoderLayout, EncoderParams, call


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_dev = torch.cuda.current_device


def _stream():
    return _raw_stream(_cur_dev())          # int -> c_void_p by argtypes


def _p(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous() and (dtype is None or t.dtype == dtype)):
        if not t.is_cuda:
            raise _lib.IrError("expected a CUDA tensor (there is no CPU path)")
        if not t.is_contiguous():
            raise _lib.IrError("expected a contiguous tensor")
        raise _lib.IrError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()            # argtypes are c_void_p: a plain int converts without a ctypes object


_checked_devices = set()


_MODE_APPLIED = [False]


def check_device(device_index=None):
    """Raise unless the current device is a B200-class (sm_100) GPU.  Checked once per device."""
    if not _checked_devices and not torch.cuda.is_available():
        raise _lib.IrError("no CUDA device: instancerefer_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device_index is None else device_index
    if idx not in _checked_devices:
        call("ir_check_device", idx)
        _checked_devices.add(idx)
    if not _MODE_APPLIED[0]:
        _MODE_APPLIED[0] = True
        import os
        if os.environ.get('IR_ENCODER') in ('layers', 'persist'):
            set_encoder_mode(os.environ['IR_ENCODER'])
        if os.environ.get('IR_GATHER') in ('ldg', 'tma'):
            set_gather_mode(os.environ['IR_GATHER'])
        if os.environ.get('IR_TUNE'):                     # "pairgemm_ctas,reduce_ctas" (ir_tune_set)
            a, b = (int(x) for x in os.environ['IR_TUNE'].split(','))
            call("ir_tune_set", a, b)


# ----------------------------------------------------------------------------- encoder workspace

def round_rows(n):
    """Row-capacity bucket: workspaces are cached per bucket so allocations stay stable."""
    return max(8192, (int(n) + 8191) // 8192 * 8192)


class EncoderWorkspace:
    """Caller-owned scratch for one sparse-encoder pass (hash tables, level coords, rulebooks,
    activations, pair products).  ``view_*`` helpers expose the device-side tables for tests."""

    def __init__(self, n_max, device):
        self.n_max = int(n_max)
        self.layout = EncoderLayout()
        call("ir_encoder_layout", self.n_max, C.byref(self.layout))
        self.buf = torch.empty(self.layout.total_bytes, dtype=torch.uint8, device=device)
        self.ptr = C.c_void_p(self.buf.data_ptr())

    def _view(self, off, nbytes, dtype):
        return self.buf[off:off + nbytes].view(dtype)

    def nlvl(self):
        return self._view(self.layout.off_nlvl, 32, torch.int32)[:5]

    def kcount(self):
        return self._view(self.layout.off_kcount, 9 * 32 * 4, torch.int32).view(9, 32)

    def coords(self, level):
        return self._view(self.layout.off_coords[level], self.n_max * 16, torch.int32).view(self.n_max, 4)

    def feat0(self, fdim):
        return self._view(self.layout.off_feat0, self.n_max * 8 * 4, torch.float32)[: self.n_max * fdim].view(self.n_max, fdim)

    def k3(self, level):
        L = self.layout
        return (self._view(L.off_k3_in[level], 27 * self.n_max * 4, torch.int32).view(27, self.n_max),
                self._view(L.off_k3_slot[level], 27 * self.n_max * 4, torch.int32).view(27, self.n_max))

    def k2(self, level):
        L = self.layout
        return (self._view(L.off_k2_in[level], 8 * self.n_max * 4, torch.int32).view(8, self.n_max),
                self._view(L.off_k2_slot[level], 8 * self.n_max * 4, torch.int32).view(8, self.n_max))

    def T(self):
        return self._view(self.layout.off_T, 27 * self.n_max * 128 * 4, torch.float32)


def set_encoder_mode(mode):
    """'layers' (default): one pair-GEMM + one reduce launch per conv layer; 'persist': all 13 layers of an encoder (or of
    both, ``encoder_features_pair``) in one persistent launch.  Environment: IR_ENCODER=layers|persist."""
    call("ir_encoder_mode_set", {'layers': 0, 'persist': 1}[mode])


def set_gather_mode(mode):
    """'ldg' (16-byte loads by the producer warps) or 'tma' (cp.async.bulk.tensor tile::gather4 into a raw stage) for the
    forward pair-GEMM.  Environment: IR_GATHER=ldg|tma."""
    call("ir_gather_mode_set", {'ldg': 0, 'tma': 1}[mode])


TIMELINE = None          # tools/timeline.py sets this to a Timeline; None = no stamps (zero overhead)


class Timeline:
    """Branch-level GPU timeline: ``stamp(label)`` queues a one-thread kernel that writes the GPU timer on the current
    stream (also under stream capture); ``read()`` -> [(label, us since the first stamp)]."""

    def __init__(self, device, cap=256):
        self.buf = torch.zeros(cap, dtype=torch.int64, device=device)
        self.labels = []
        self.frozen = False

    def stamp(self, label):
        if self.frozen:                       # replay / repeated call: indices are already assigned
            return
        call("ir_debug_stamp", C.c_void_p(self.buf.data_ptr()), len(self.labels), _stream())
        self.labels.append(label)

    def read(self):
        v = self.buf[:len(self.labels)].cpu().numpy()
        t0 = int(v.min())
        return [(l, (int(x) - t0) / 1e3) for l, x in zip(self.labels, v)]


def stamp(label):
    if TIMELINE is not None:
        TIMELINE.stamp(label)


def encoder_reset(ws):
    call("ir_encoder_reset", ws.ptr, ws.n_max, _stream())


def voxelize(pts, cand, voxel, ws):
    """pts (n_inst, ppi, fdim) fp32 cuda; cand (M,) int32 cuda.  Result stays in ``ws`` level 0."""
    n_inst, ppi, fdim = pts.shape
    call("ir_voxelize", _p(pts, torch.float32), _p(cand, torch.int32), cand.numel(), ppi, fdim,
         float(voxel), ws.ptr, ws.n_max, _stream())


def encoder_build_maps(ws, coords0=None, n0_dev=None):
    """coords0 (n,4) int32; with n0_dev (device int32[1]) only the first *n0_dev rows are live."""
    if coords0 is None:
        call("ir_encoder_build_maps", None, 0, None, ws.ptr, ws.n_max, _stream())
    else:
        call("ir_encoder_build_maps", _p(coords0, torch.int32), coords0.shape[0],
             _p(n0_dev, torch.int32) if n0_dev is not None else None, ws.ptr, ws.n_max, _stream())


def make_encoder_params(cin, weights, bn_scale, bn_shift, wprep=None, use_tc=False):
    """Returns (struct, keepalive list).  13 tensors each, layer order of the header."""
    P = EncoderParams()
    P.cin = cin
    P.use_tc = 1 if use_tc else 0
    keep = []
    for i in range(_lib.ENC_LAYERS):
        P.weight[i] = weights[i].data_ptr()
        P.bn_scale[i] = bn_scale[i].data_ptr()
        P.bn_shift[i] = bn_shift[i].data_ptr()
        P.wprep[i] = wprep[i].data_ptr() if (wprep is not None and wprep[i] is not None) else None
        keep += [weights[i], bn_scale[i], bn_shift[i]]
        if wprep is not None:
            keep.append(wprep[i])
    return P, keep


def encoder_features(params, ws, feats0, out):
    call("ir_encoder_features", C.byref(params), _p(feats0, torch.float32), ws.ptr, ws.n_max,
         _p(out, torch.float32), _stream())


def encoder_features_pair(pa, wsa, feats0a, outa, pb, wsb, feats0b, outb):
    """Both encoders of the model through shared launches (see ir_encoder_features_pair)."""
    call("ir_encoder_features_pair", C.byref(pa), _p(feats0a, torch.float32), wsa.ptr, wsa.n_max, _p(outa, torch.float32),
         C.byref(pb), _p(feats0b, torch.float32), wsb.ptr, wsb.n_max, _p(outb, torch.float32), _stream())


def spconv_wprep(weight):
    """(K,Cin,Cout) fp32 -> tcgen05 operand image tensor (see ir_spconv_prepare_weights)."""
    K, cin, cout = weight.shape
    n = _lib.load().ir_spconv_wprep_floats(K, cin, cout)
    out = torch.empty(n, dtype=torch.float32, device=weight.device)
    call("ir_spconv_prepare_weights", _p(weight.contiguous(), torch.float32), K, cin, cout, _p(out), _stream())
    return out


def spconv_layer(feat_in, in_idx, slot, count, n_out_dev, n_max, weight, scale, shift, resid, relu,
                 T, out, wprep=None, use_tc=False):
    K, cin, cout = weight.shape
    call("ir_spconv_layer", _p(feat_in, torch.float32), cin, cout, K, _p(in_idx, torch.int32),
         in_idx.shape[1], _p(slot, torch.int32), _p(count, torch.int32), _p(n_out_dev, torch.int32),
         n_max, _p(weight, torch.float32), _p(wprep), 1 if use_tc else 0, _p(scale), _p(shift),
         _p(resid), 1 if relu else 0, _p(T, torch.float32), _p(out, torch.float32), _stream())


def segmax(feats, coords, n_dev, n_max, n_seg):
    Cc = feats.shape[1]
    scratch = torch.empty(n_seg * Cc, dtype=torch.int32, device=feats.device)
    out = torch.empty(n_seg, Cc, dtype=torch.float32, device=feats.device)
    call("ir_segmax", _p(feats, torch.float32), _p(coords, torch.int32), _p(n_dev, torch.int32), n_max,
         Cc, n_seg, _p(scratch), _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- scene head

def bev(feats, coords, n_dev, n_max, stride, kernel, scale, shift, B, absmax=None):
    """absmax: optional zeroed device float receiving max|out| (range scale of the tcgen05 Conv2d behind it)."""
    dev = feats.device
    tmp = torch.empty(n_max, 128, dtype=torch.float32, device=dev)
    cell = torch.empty(n_max, dtype=torch.int32, device=dev)
    out = torch.empty(B, 15, 25, 128, dtype=torch.float32, device=dev)
    call("ir_bev", _p(feats, torch.float32), _p(coords, torch.int32), _p(n_dev, torch.int32), n_max, stride,
         _p(kernel, torch.float32), _p(scale), _p(shift), B, _p(tmp), _p(cell), _p(out), _p(absmax), _stream())
    return out


_GRID_RULEBOOKS = {}


def grid_rulebook(B, H, W, device):
    """Closed-form rulebook of a 3x3 valid convolution over a dense (B,H,W) grid in the sparse-conv layout: tap
    k = ky*3+kx of output pixel o = (b,y,x) reads input pixel (b, y+ky, x+kx); every output has all nine pairs."""
    key = (B, H, W, str(device))
    rb = _GRID_RULEBOOKS.get(key)
    if rb is None:
        Ho, Wo = H - 2, W - 2
        b, y, x = torch.meshgrid(torch.arange(B), torch.arange(Ho), torch.arange(Wo), indexing='ij')
        base = (b * H * W + y * W + x).reshape(-1)
        n_out = base.numel()
        in_idx = torch.stack([base + ky * W + kx for ky in range(3) for kx in range(3)]).to(torch.int32)
        slot = torch.arange(n_out, dtype=torch.int32).repeat(9, 1)
        rb = dict(in_idx=in_idx.contiguous().to(device), slot=slot.contiguous().to(device),
                  count=torch.full((9,), n_out, dtype=torch.int32, device=device),
                  n_out_dev=torch.tensor([n_out], dtype=torch.int32, device=device), n_out=n_out, Ho=Ho, Wo=Wo)
        _GRID_RULEBOOKS[key] = rb
    return rb


def conv2d_3x3_tc(x, w9, scale, shift, relu, in_absmax=None, out_absmax=None):
    """x (B,H,W,128) NHWC fp32; w9 (9,128,128) = [ky][kx][Cin][Cout] (16-byte aligned); y = act(scale*conv + shift)."""
    B, H, W, Cc = x.shape
    rb = grid_rulebook(B, H, W, x.device)
    T = torch.empty(9 * rb['n_out'], Cc, dtype=torch.float32, device=x.device)
    out = torch.empty(B, rb['Ho'], rb['Wo'], Cc, dtype=torch.float32, device=x.device)
    call("ir_conv2d_3x3_tc", _p(x, torch.float32), _p(rb['in_idx'], torch.int32), _p(rb['slot'], torch.int32),
         _p(rb['count'], torch.int32), _p(rb['n_out_dev'], torch.int32), rb['n_out'], _p(w9, torch.float32), _p(scale),
         _p(shift), 1 if relu else 0, _p(in_absmax), _p(out_absmax), _p(T), _p(out), _stream())
    return out


def conv2d_3x3(x, wpack, bias, scale, shift, relu):
    B, H, W, Cc = x.shape
    out = torch.empty(B, H - 2, W - 2, Cc, dtype=torch.float32, device=x.device)
    call("ir_conv2d_3x3", _p(x, torch.float32), B, H, W, Cc, _p(wpack, torch.float32), _p(bias), _p(scale),
         _p(shift), 1 if relu else 0, _p(out), _stream())
    return out


def scene_attention(feats, q):
    B, ncell, Cc = feats.shape
    atten = torch.empty(B, ncell, dtype=torch.float32, device=feats.device)
    sf = torch.empty(B, Cc, dtype=torch.float32, device=feats.device)
    call("ir_scene_attention", _p(feats, torch.float32), _p(q, torch.float32), B, ncell, Cc, _p(atten), _p(sf), _stream())
    return atten, sf


# ----------------------------------------------------------------------------- language

def linear(x, W, b, relu=False):
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    call("ir_linear", _p(x, torch.float32), M, K, _p(W, torch.float32), _p(b), N, 1 if relu else 0, _p(y), _stream())
    return y


def gru_layer(xproj, whh, bhh, lengths, B, L, H=128):
    out = torch.empty(B, L, 2 * H, dtype=torch.float32, device=xproj.device)
    call("ir_gru_layer", _p(xproj, torch.float32), _p(whh, torch.float32), _p(bhh, torch.float32),
         _p(lengths, torch.int64), B, L, H, _p(out), _stream())
    return out


def token_attention(feats, embed, lengths, fcw, fcb):
    B, L, D = feats.shape
    E = embed.shape[-1]
    atten = torch.empty(4, B, L, dtype=torch.float32, device=feats.device)
    pooled = torch.empty(4, B, E, dtype=torch.float32, device=feats.device)
    call("ir_token_attention", _p(feats, torch.float32), _p(embed, torch.float32), embed.shape[1] * E,
         _p(lengths, torch.int64), _p(fcw, torch.float32), _p(fcb, torch.float32), B, L, D, E,
         _p(atten), _p(pooled), _stream())
    return atten, pooled


# ----------------------------------------------------------------------------- heads / relation

NORM_NONE, NORM_AFFINE, NORM_LAYER = 0, 1, 2
MODE_RAW, MODE_L2, MODE_DOT, MODE_COS = 0, 1, 2, 3


def mlp_head(x, W1, b1, norm, g, beta, W2, b2, mode, partner=None, seg=None, want_y=False):
    """W1 (K,N1), W2 (N1,N2): (in,out) layout, i.e. transposed nn.Linear weights."""
    M, K = x.shape
    N1, N2 = W1.shape[1], W2.shape[1]
    dev = x.device
    y = torch.empty(M, N2, dtype=torch.float32, device=dev) if (mode < 2 or want_y) else None
    score = torch.empty(M, dtype=torch.float32, device=dev) if mode >= 2 else None
    call("ir_mlp_head", _p(x, torch.float32), M, K, _p(W1, torch.float32), _p(b1), N1, norm, _p(g), _p(beta),
         _p(W2, torch.float32), _p(b2), N2, mode, _p(partner), _p(seg, torch.int32) if seg is not None else None,
         _p(y), _p(score), _stream())
    return y, score


def candidate_softmax(sa, sr, ss, seg_ofs):
    n_seg = seg_ofs.numel() - 1
    prob = torch.empty_like(sa)
    arg = torch.empty(n_seg, dtype=torch.int32, device=sa.device)
    call("ir_candidate_softmax", _p(sa, torch.float32), _p(sr, torch.float32), _p(ss, torch.float32),
         _p(seg_ofs, torch.int32), n_seg, _p(prob), _p(arg), _stream())
    return prob, arg


def instance_mean(pts):
    n_inst, ppi, fdim = pts.shape
    mean = torch.empty(n_inst, fdim, dtype=torch.float32, device=pts.device)
    call("ir_instance_mean", _p(pts, torch.float32), n_inst, ppi, fdim, _p(mean), _stream())
    return mean


def knn(xyz, seg_ofs, qidx, qseg, k):
    nq = qidx.numel()
    nbr = torch.empty(nq, k, dtype=torch.int32, device=xyz.device)
    call("ir_knn", _p(xyz, torch.float32), _p(seg_ofs, torch.int32), _p(qidx, torch.int32),
         _p(qseg, torch.int32), nq, k, _p(nbr), _stream())
    return nbr


def edgeconv(x, xyz, qidx, nbr, ncls, Ww1, bw1, Ww2, bw2, Wm1, bm1, Wm2, bm2):
    """Weights in (in,out) layout (transposed nn.Linear weights)."""
    nq, k = nbr.shape
    F = x.shape[1]
    Fout = Wm2.shape[1]
    out = torch.empty(nq, Fout, dtype=torch.float32, device=x.device)
    call("ir_edgeconv", _p(x, torch.float32), _p(xyz, torch.float32), _p(qidx, torch.int32),
         _p(nbr, torch.int32), nq, k, F, ncls, _p(Ww1), _p(bw1), _p(Ww2), _p(bw2), _p(Wm1), _p(bm1),
         _p(Wm2), _p(bm2), Fout, _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- training step

def rulebook_transpose(in_idx, slot, n_out_dev, n_rows_out):
    """(K,seg_cap) forward rulebook -> (out_idx, slot_in) of the transposed map (dgrad / wgrad)."""
    K, seg_cap = in_idx.shape
    out_idx = torch.empty_like(in_idx)
    slot_in = torch.empty_like(slot)
    call("ir_rulebook_transpose", _p(in_idx, torch.int32), _p(slot, torch.int32), K, seg_cap,
         _p(n_out_dev, torch.int32), int(n_rows_out), _p(out_idx), _p(slot_in), _stream())
    return out_idx, slot_in


def spconv_wgrad(x, dy, in_idx, out_idx, count, dy_absmax=None, use_tc=False):
    K, seg_cap = in_idx.shape
    cin, cout = x.shape[1], dy.shape[1]
    dW = torch.empty(K, cin, cout, dtype=torch.float32, device=x.device)
    if dy_absmax is not None:
        call("ir_spconv_wgrad_scaled", _p(x, torch.float32), cin, _p(dy, torch.float32), _p(dy_absmax, torch.float32), cout, K,
             _p(in_idx, torch.int32), _p(out_idx, torch.int32), _p(count, torch.int32), seg_cap, 1 if use_tc else 0, _p(dW), _stream())
        return dW
    call("ir_spconv_wgrad", _p(x, torch.float32), cin, _p(dy, torch.float32), cout, K, _p(in_idx, torch.int32),
         _p(out_idx, torch.int32), _p(count, torch.int32), seg_cap, _p(dW), _stream())
    return dW


BN_PARTS = 128              # ir_bn_scratch_floats(C) = BN_PARTS * 2 * C


def bn_train_fwd(x, gamma, beta, resid, relu, eps, momentum, running_mean, running_var):
    """x (n,C) fp32 -> (y, mean, rstd); running statistics are updated in place."""
    n, Cc = x.shape
    dev = x.device
    scratch = torch.empty(BN_PARTS * 2 * Cc, dtype=torch.float32, device=dev)
    mean = torch.empty(Cc, dtype=torch.float32, device=dev)
    rstd = torch.empty(Cc, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    call("ir_bn_train_fwd", _p(x, torch.float32), None, n, Cc, _p(gamma, torch.float32), _p(beta, torch.float32),
         _p(resid), 1 if relu else 0, float(eps), float(momentum), _p(running_mean), _p(running_var),
         _p(scratch), _p(mean), _p(rstd), _p(y), _stream())
    return y, mean, rstd


def bn_train_bwd(dy, y, x, mean, rstd, gamma, relu, want_resid, absmax=None):
    n, Cc = x.shape
    dev = x.device
    scratch = torch.empty(BN_PARTS * 2 * Cc, dtype=torch.float32, device=dev)
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_resid else None
    dgamma = torch.empty(Cc, dtype=torch.float32, device=dev)
    dbeta = torch.empty(Cc, dtype=torch.float32, device=dev)
    call("ir_bn_train_bwd", _p(dy, torch.float32), _p(y), _p(x, torch.float32), None, n, Cc, _p(mean), _p(rstd),
         _p(gamma, torch.float32), 1 if relu else 0, _p(scratch), _p(dx), _p(dres), _p(dgamma), _p(dbeta), _p(absmax),
         _stream())
    return dx, dres, dgamma, dbeta


def segmax_bwd(feats, coords, n_dev, n_rows, n_seg, pooled, dpooled):
    Cc = feats.shape[1]
    arg = torch.empty(n_seg * Cc, dtype=torch.int32, device=feats.device)
    dfeats = torch.empty_like(feats)
    call("ir_segmax_bwd", _p(feats, torch.float32), _p(coords, torch.int32), _p(n_dev, torch.int32), int(n_rows),
         Cc, n_seg, _p(pooled, torch.float32), _p(dpooled, torch.float32), _p(arg), _p(dfeats), _stream())
    return dfeats


def cross_entropy(logits, labels):
    """-> (loss (1,), dlogits) : mean CE over rows and its gradient."""
    B, N = logits.shape
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    call("ir_cross_entropy", _p(logits, torch.float32), _p(labels, torch.int64), B, N, _p(loss), _p(dl), _stream())
    return loss, dl


def region_label(ref_center, point_min, point_max):
    f32 = all(t.dtype == torch.float32 for t in (ref_center, point_min, point_max))
    B = ref_center.shape[0]
    label = torch.empty(B, dtype=torch.int64, device=ref_center.device)
    c, lo, hi = (t.to(torch.float64).contiguous() for t in (ref_center, point_min, point_max))   # kept alive
    call("ir_region_label", _p(c), _p(lo), _p(hi), B, 1 if f32 else 0, _p(label), _stream())
    return label


def ref_loss(pred_obb, obb_ofs, gt_obb, score_ofs, sa, sr, ss, margin=0.2, gamma=5.0, iou_thresh=0.2):
    """-> (label (Nc,), loss_scene (B,), dscore (M,), iou_max (B,))."""
    B = gt_obb.shape[0]
    dev = sa.device
    label = torch.empty(pred_obb.shape[0], dtype=torch.float32, device=dev)
    loss_scene = torch.empty(B, dtype=torch.float32, device=dev)
    dscore = torch.zeros_like(sa)
    iou_max = torch.empty(B, dtype=torch.float32, device=dev)
    call("ir_ref_loss", _p(pred_obb, torch.float64), _p(obb_ofs, torch.int32), _p(gt_obb, torch.float64),
         _p(score_ofs, torch.int32), B, _p(sa, torch.float32), _p(sr, torch.float32), _p(ss, torch.float32),
         float(margin), float(gamma), float(iou_thresh), _p(label), _p(loss_scene), _p(dscore), _p(iou_max), _stream())
    return label, loss_scene, dscore, iou_max


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
              block_skip=None):
    """block_skip: uint8 per 64-float block, 1 = leave parameter and moments untouched (no gradient this step)."""
    call("ir_adam_step", _p(params, torch.float32), _p(grads, torch.float32), _p(exp_avg, torch.float32),
         _p(exp_avg_sq, torch.float32), params.numel(), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), int(step), float(grad_scale), _p(block_skip, torch.uint8), _stream())


def adam_hyper(lr, beta1, beta2, step, out):
    """out: pinned/host float32[3] <- (lr, 1-beta1^step, sqrt(1-beta2^step)), the host arithmetic of ir_adam_step."""
    call("ir_adam_hyper", float(lr), float(beta1), float(beta2), int(step), C.c_void_p(out.data_ptr()))


def adam_step_dev(params, grads, exp_avg, exp_avg_sq, hyper_dev, beta1, beta2, eps, weight_decay, grad_scale=1.0,
                  block_skip=None):
    """adam_step with lr and the bias corrections read from hyper_dev (device float32[3]): graph-replayable."""
    call("ir_adam_step_dev", _p(params, torch.float32), _p(grads, torch.float32), _p(exp_avg, torch.float32),
         _p(exp_avg_sq, torch.float32), params.numel(), _p(hyper_dev, torch.float32), float(beta1), float(beta2),
         float(eps), float(weight_decay), float(grad_scale), _p(block_skip, torch.uint8), _stream())


def dropout_seed_step(step_dev):
    """Device uint64 (as an int64 tensor) folded into every later dropout seed on the device, or None: off."""
    call("ir_dropout_seed_step", _p(step_dev, torch.int64) if step_dev is not None else None)


# ----------------------------------------------------------------------------- dense training-step operators

def gemm(A, B, ta=False, tb=False, bias=None, relu=False, out=None, accumulate=False):
    """out (M,N) = op(A) op(B) [+bias] [+out] [relu]; A, B 2-D fp32 with unit inner stride."""
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    N = B.shape[0] if tb else B.shape[1]
    assert (B.shape[1] if tb else B.shape[0]) == K and A.stride(1) == 1 and B.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    call("ir_gemm", M, N, K, C.c_void_p(A.data_ptr()), A.stride(0), 1 if ta else 0, C.c_void_p(B.data_ptr()),
         B.stride(0), 1 if tb else 0, _p(out, torch.float32), N, _p(bias), 1 if relu else 0,
         1 if accumulate else 0, _stream())
    return out


def colsum(x):
    M, N = x.shape
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    call("ir_colsum", _p(x, torch.float32), M, N, _p(out), _stream())
    return out


def relu_bwd(dy, y):
    dx = torch.empty_like(dy)
    call("ir_relu_bwd", _p(dy, torch.float32), _p(y, torch.float32), dy.numel(), _p(dx), _stream())
    return dx


def dropout_fwd(x, p, seed):
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    call("ir_dropout_fwd", _p(x, torch.float32), x.numel(), float(p), C.c_uint64(seed & (2 ** 64 - 1)), _p(y), _p(mask), _stream())
    return y, mask


def dropout_bwd(dy, mask, p):
    dx = torch.empty_like(dy)
    call("ir_dropout_bwd", _p(dy, torch.float32), _p(mask, torch.uint8), dy.numel(), float(p), _p(dx), _stream())
    return dx


def layernorm_fwd(x, gamma, beta, eps, relu):
    M, N = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, dtype=torch.float32, device=x.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    call("ir_layernorm_fwd", _p(x, torch.float32), M, N, _p(gamma, torch.float32), _p(beta, torch.float32), float(eps),
         1 if relu else 0, _p(y), _p(mean), _p(rstd), _stream())
    return y, mean, rstd


def layernorm_bwd(dy, y, x, gamma, mean, rstd, relu):
    M, N = x.shape
    dx = torch.empty_like(x)
    dg = torch.empty(N, dtype=torch.float32, device=x.device)
    db = torch.empty(N, dtype=torch.float32, device=x.device)
    call("ir_layernorm_bwd", _p(dy, torch.float32), _p(y), _p(x, torch.float32), M, N, _p(gamma, torch.float32), _p(mean),
         _p(rstd), 1 if relu else 0, _p(dx), _p(dg), _p(db), _stream())
    return dx, dg, db


def l2norm_fwd(x):
    y = torch.empty_like(x)
    call("ir_l2norm_fwd", _p(x, torch.float32), x.shape[0], x.shape[1], _p(y), _stream())
    return y


def l2norm_bwd(dy, x):
    dx = torch.empty_like(x)
    call("ir_l2norm_bwd", _p(dy, torch.float32), _p(x, torch.float32), x.shape[0], x.shape[1], _p(dx), _stream())
    return dx


def match_fwd(a, partner, seg, mode):
    M, N = a.shape
    score = torch.empty(M, dtype=torch.float32, device=a.device)
    call("ir_match_fwd", _p(a, torch.float32), _p(partner, torch.float32), _p(seg, torch.int32), M, N, mode, _p(score), _stream())
    return score


def match_bwd(dscore, a, partner, seg, row_ofs, mode):
    M, N = a.shape
    da = torch.empty_like(a)
    dp = torch.empty_like(partner)
    call("ir_match_bwd", _p(dscore, torch.float32), _p(a, torch.float32), _p(partner, torch.float32), _p(seg, torch.int32),
         _p(row_ofs, torch.int32), M, N, partner.shape[0], mode, _p(da), _p(dp), _stream())
    return da, dp


def im2col_3x3(x):
    B, H, W, Cc = x.shape
    col = torch.empty(B * (H - 2) * (W - 2), 9 * Cc, dtype=torch.float32, device=x.device)
    call("ir_im2col_3x3", _p(x, torch.float32), B, H, W, Cc, _p(col), _stream())
    return col


def col2im_3x3(dcol, B, H, W, Cc):
    din = torch.empty(B, H, W, Cc, dtype=torch.float32, device=dcol.device)
    call("ir_col2im_3x3", _p(dcol, torch.float32), B, H, W, Cc, _p(din), _stream())
    return din


def bev_raw(feats, coords, n_dev, n_rows, stride, kernel, B):
    """ir_bev without BN/ReLU (train mode) -> (dense (B*375,128), cell (n_rows,))."""
    dev = feats.device
    tmp = torch.empty(max(n_rows, 1), 128, dtype=torch.float32, device=dev)
    cell = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
    out = torch.empty(B * 375, 128, dtype=torch.float32, device=dev)
    call("ir_bev", _p(feats, torch.float32), _p(coords, torch.int32), _p(n_dev, torch.int32), n_rows, stride,
         _p(kernel, torch.float32), None, None, B, _p(tmp), _p(cell), _p(out), None, _stream())
    return out, cell


def bev_bwd(ddense, feats, coords, cell, n_dev, n_rows, stride, kernel):
    df = torch.empty_like(feats)
    dk = torch.empty_like(kernel)
    call("ir_bev_bwd", _p(ddense, torch.float32), _p(feats, torch.float32), _p(coords, torch.int32), _p(cell, torch.int32),
         _p(n_dev, torch.int32), n_rows, stride, _p(kernel, torch.float32), kernel.shape[0], _p(df), _p(dk), _stream())
    return df, dk


def scene_attention_bwd(feats, q, atten, dscene):
    B, ncell, Cc = feats.shape
    df = torch.empty_like(feats)
    dq = torch.empty_like(q)
    call("ir_scene_attention_bwd", _p(feats, torch.float32), _p(q, torch.float32), _p(atten, torch.float32),
         _p(dscene, torch.float32), None, B, ncell, Cc, _p(df), _p(dq), _stream())
    return df, dq


def token_attention_bwd(feats, embed, lengths, fcw, fcb, atten, dpooled):
    B, L, D = feats.shape
    E = embed.shape[-1]
    dev = feats.device
    dfeats = torch.empty_like(feats)
    dembed = torch.empty(B, L, E, dtype=torch.float32, device=dev)
    dw = torch.empty(B, 4 * D, dtype=torch.float32, device=dev)
    db = torch.empty(B, 4, dtype=torch.float32, device=dev)
    call("ir_token_attention_bwd", _p(feats, torch.float32), _p(embed, torch.float32), embed.shape[1] * E,
         _p(lengths, torch.int64), _p(fcw, torch.float32), _p(fcb, torch.float32), _p(atten, torch.float32),
         _p(dpooled, torch.float32), B, L, D, E, _p(dfeats), _p(dembed), _p(dw), _p(db), _stream())
    return dfeats, dembed, colsum(dw).view(4, D), colsum(db)


def gru_layer_bwd(xproj, whh, bhh, lengths, out, dout, B, L, H=128):
    dev = xproj.device
    dxp = torch.empty_like(xproj)
    dhp = torch.empty(B * L, 2 * 3 * H, dtype=torch.float32, device=dev)
    hprev = torch.empty(B * L, 2 * H, dtype=torch.float32, device=dev)
    call("ir_gru_layer_bwd", _p(xproj, torch.float32), _p(whh, torch.float32), _p(bhh, torch.float32),
         _p(lengths, torch.int64), _p(out, torch.float32), _p(dout, torch.float32), B, L, H, _p(dxp), _p(dhp), _p(hprev), _stream())
    dw = torch.empty(2, 3 * H, H, dtype=torch.float32, device=dev)
    for d in (0, 1):            # dW_hh[d] = dhp_d^T @ hprev_d over all (sample, step) rows
        gemm(dhp[:, d * 3 * H:(d + 1) * 3 * H], hprev[:, d * H:(d + 1) * H], ta=True, out=dw[d])
    return dxp, dw, colsum(dhp).view(2, 3 * H)


def edge_inputs(x, xyz, qidx, nbr, ncls, w=None):
    nq, k = nbr.shape
    F = x.shape[1]
    D = 3 * F if w is not None else 3 + 2 * ncls
    out = torch.empty(nq * k, D, dtype=torch.float32, device=x.device)
    call("ir_edge_inputs", _p(x, torch.float32), _p(xyz, torch.float32), _p(qidx, torch.int32), _p(nbr, torch.int32),
         nq, k, F, ncls, _p(w), None if w is not None else _p(out), _p(out) if w is not None else None, _stream())
    return out


def edge_max_fwd(msg, nbr):
    nq, k = nbr.shape
    Cc = msg.shape[1]
    out = torch.empty(nq, Cc, dtype=torch.float32, device=msg.device)
    arg = torch.empty(nq, Cc, dtype=torch.int32, device=msg.device)
    call("ir_edge_max_fwd", _p(msg, torch.float32), _p(nbr, torch.int32), nq, k, Cc, _p(out), _p(arg), _stream())
    return out, arg


def edge_max_bwd(dout, arg, k):
    nq, Cc = dout.shape
    dmsg = torch.empty(nq * k, Cc, dtype=torch.float32, device=dout.device)
    call("ir_edge_max_bwd", _p(dout, torch.float32), _p(arg, torch.int32), nq, k, Cc, _p(dmsg), _stream())
    return dmsg


def ref_eval(pred_obb, obb_ofs, gt_obb, score_ofs, sa, sr, ss, label):
    """-> (pred_idx (B,) i32, ref_acc (B,) f32, iou (B,) f64, pred_corners (B,8,3) f64, gt_corners (B,8,3) f64)."""
    B = gt_obb.shape[0]
    dev = sa.device
    pred_idx = torch.empty(B, dtype=torch.int32, device=dev)
    ref_acc = torch.empty(B, dtype=torch.float32, device=dev)
    iou = torch.empty(B, dtype=torch.float64, device=dev)
    pc = torch.empty(B, 8, 3, dtype=torch.float64, device=dev)
    gc = torch.empty(B, 8, 3, dtype=torch.float64, device=dev)
    call("ir_ref_eval", _p(pred_obb, torch.float64), _p(obb_ofs, torch.int32), _p(gt_obb, torch.float64),
         _p(score_ofs, torch.int32), B, _p(sa.contiguous(), torch.float32), _p(sr.contiguous(), torch.float32),
         _p(ss.contiguous(), torch.float32), _p(label, torch.float32), _p(pred_idx), _p(ref_acc), _p(iou), _p(pc), _p(gc), _stream())
    return pred_idx, ref_acc, iou, pc, gc


def spconv_layer_scaled(feat_in, absmax, in_idx, slot, count, n_out_dev, n_rows, weight, resid, T, out, use_tc=True):
    """ir_spconv_layer_scaled: conv without BN epilogue; tcgen05 pair-GEMM with the input rows range-scaled
    by ``absmax`` (device scalar max|feat_in|) — the dgrad of the training step."""
    K, cin, cout = weight.shape
    call("ir_spconv_layer_scaled", _p(feat_in, torch.float32), _p(absmax, torch.float32), cin, cout, K,
         _p(in_idx, torch.int32), in_idx.shape[1], _p(slot, torch.int32), _p(count, torch.int32),
         _p(n_out_dev, torch.int32), int(n_rows), _p(weight, torch.float32), 1 if use_tc else 0, _p(resid),
         _p(T, torch.float32), _p(out, torch.float32), _stream())
