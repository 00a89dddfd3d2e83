"""Building blocks with the reference's names, constructor signatures and state_dict layout
(models/basic_blocks.py), executing through the CUDA library.

The ``nn.Module`` tree only *holds parameters* in the reference's layout (so checkpoints load
strict); the arithmetic is done by ``SparseConvEncoder.encode`` / ``DynamicEdgeConv.forward`` /
``ToDenseBEVConvolution`` via instancerefer_b200.ops (hash build -> kernel maps -> pair-GEMM ->
reduce with fused BN/ReLU/residual).  These classes hold the eval-mode entry points (BatchNorm folded);
in train mode the modules' ``forward`` dispatch to instancerefer_b200.training (batch statistics, autograd)."""
import math
import os

import torch
import torch.nn as nn

from . import ops
from .sparse_tensor import SparseTensor


_WEIGHTS_EPOCH = [0]


def bump_weights_epoch():
    """Called by whoever writes parameters behind torch's back (``FlatAdam.step`` updates the flat parameter buffer
    through a raw-pointer kernel, which does not move any tensor ``_version``): every prepared copy and every captured
    graph that baked one in becomes stale."""
    _WEIGHTS_EPOCH[0] += 1


def weights_epoch():
    return _WEIGHTS_EPOCH[0]


class PrepCache:
    """Kernel-friendly copies of the parameters (folded BN, repacked weights), rebuilt only when a
    parameter/buffer changed (tensor ``_version``), moved, or the global weights epoch advanced (raw-pointer
    optimiser updates, see ``bump_weights_epoch``)."""

    def _prep_tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _prep_key(self):
        ts = self.__dict__.get('_prep_ts')
        if ts is None:                      # module structure is fixed after construction
            ts = self._prep_tensors()
            self.__dict__['_prep_ts'] = ts
        return [_WEIGHTS_EPOCH[0]] + [(t.data_ptr(), t._version) for t in ts]

    def prepared(self):
        key = self._prep_key()
        if getattr(self, '_prep_cache_key', None) != key:
            with torch.no_grad():
                self._prep_cache = self._prepare()
            self._prep_cache_key = key
        return self._prep_cache


def fold_bn(bn):
    """eval BatchNorm -> per-channel (scale, shift): y = x*scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    return scale.float().contiguous(), shift.float().contiguous()


def require_eval(module):
    if module.training:
        raise RuntimeError(
            "instancerefer_b200: this helper folds BatchNorm running statistics (eval mode); in train mode go "
            "through the module's forward(), which dispatches to instancerefer_b200.training")


class Conv3d(nn.Module):
    """Parameter holder for spnn.Conv3d: ``kernel`` (K,inc,outc), no bias,
    init U(-s,s), s = 1/sqrt(inc*K)."""

    def __init__(self, inc, outc, kernel_size=3, stride=1, dilation=1, transpose=False):
        super().__init__()
        assert dilation == 1 and not transpose
        self.inc, self.outc, self.kernel_size, self.stride = inc, outc, kernel_size, stride
        K = kernel_size ** 3
        self.kernel = nn.Parameter(torch.zeros(K, inc, outc))
        std = 1. / math.sqrt(inc * K)
        self.kernel.data.uniform_(-std, std)


class ReLU(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()


class BasicConvolutionBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, dilation=1, transpose=False):
        super().__init__()
        self.net = nn.Sequential(Conv3d(inc, outc, ks, stride, dilation, transpose),
                                 nn.BatchNorm1d(outc), ReLU(True))


class ResidualBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
        super().__init__()
        assert inc == outc and stride == 1, "the reference only instantiates identity-skip blocks"
        self.net = nn.Sequential(Conv3d(inc, outc, ks, stride, dilation), nn.BatchNorm1d(outc), ReLU(True),
                                 Conv3d(outc, outc, ks, 1, dilation), nn.BatchNorm1d(outc))
        self.downsample = nn.Sequential()
        self.relu = ReLU(True)


class SparseConvEncoder(nn.Module, PrepCache):
    """stem k3 in->32; 4 x [k2 s2 down, residual k3 k3] to 64,128,128,128
    (models/basic_blocks.py:59-95).  ``use_tc``: tcgen05 split-fp16 (hi/lo, fp32 accumulate) pair-GEMM (default) or the exact
    SIMT fp32 kernel."""

    use_tc = os.environ.get('IR_SPCONV', 'tc') != 'simt'

    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = input_dim
        self.stem = nn.Sequential(BasicConvolutionBlock(input_dim, 32, 3))
        self.stage1 = nn.Sequential(BasicConvolutionBlock(32, 64, ks=2, stride=2), ResidualBlock(64, 64, 3))
        self.stage2 = nn.Sequential(BasicConvolutionBlock(64, 128, ks=2, stride=2), ResidualBlock(128, 128, 3))
        self.stage3 = nn.Sequential(BasicConvolutionBlock(128, 128, ks=2, stride=2), ResidualBlock(128, 128, 3))
        self.stage4 = nn.Sequential(BasicConvolutionBlock(128, 128, ks=2, stride=2), ResidualBlock(128, 128, 3))
        self._ws = {}

    def _layers(self):
        cached = self.__dict__.get('_layer_list')
        if cached is not None:
            return cached
        out = [(self.stem[0].net[0], self.stem[0].net[1])]
        for st in (self.stage1, self.stage2, self.stage3, self.stage4):
            out.append((st[0].net[0], st[0].net[1]))
            out.append((st[1].net[0], st[1].net[1]))
            out.append((st[1].net[3], st[1].net[4]))
        self.__dict__['_layer_list'] = out          # module structure is fixed after construction
        return out

    def _prepare(self):
        ws, sc, sh, wp = [], [], [], []
        for conv, bn in self._layers():
            w = conv.kernel.detach().float().contiguous()
            s, b = fold_bn(bn)
            ws.append(w)
            sc.append(s)
            sh.append(b)
            wp.append(ops.spconv_wprep(w) if (self.use_tc and w.shape[1] >= 32) else None)
        params, keep = ops.make_encoder_params(self.input_dim, ws, sc, sh, wp, self.use_tc)
        return dict(params=params, keep=keep)

    def workspace(self, n_rows, device):
        """Smallest cached workspace whose capacity covers ``n_rows`` (grown, never shrunk or freed: captured CUDA
        graphs — inference replay, training passes — hold pointers into it, and a stable capacity keeps them valid
        from batch to batch)."""
        n_max = ops.round_rows(n_rows)
        dev = str(device)
        fit = [k for k in self._ws if k[1] == dev and k[0] >= n_max]
        if not fit:
            self._ws[(n_max, dev)] = ops.EncoderWorkspace(n_max, device)
            fit = [(n_max, dev)]
        self.__dict__['_last_ws'] = self._ws[min(fit)]
        return self._last_ws

    def encode(self, ws, feats0=None, coords0=None, n0_dev=None):
        """Run maps + 13 layers.  Level 0 either comes from ``ops.voxelize`` (already in ``ws``) or
        from (feats0, coords0).  -> (F4 padded (n_max,128), C4 padded (n_max,4), n4 device int)."""
        require_eval(self)
        prep = self.prepared()
        ops.encoder_build_maps(ws, coords0, n0_dev)
        ops.stamp('enc:maps')
        out = torch.empty(ws.n_max, 128, dtype=torch.float32, device=ws.buf.device)
        ops.encoder_features(prep['params'], ws, feats0, out)
        return out, ws.coords(4), ws.nlvl()[4:5]

    def build_maps(self, ws, coords0=None, n0_dev=None):
        """Coordinate half of ``encode``: levels 1-4 and the nine kernel maps."""
        require_eval(self)
        ops.encoder_build_maps(ws, coords0, n0_dev)

    def forward(self, x):
        """SparseTensor -> SparseTensor at stride 16 (API fidelity; synchronises to size the result)."""
        F = x.F.float().contiguous()
        Cc = x.C.int().contiguous()
        ws = self.workspace(F.shape[0], F.device)
        f4, c4, n4 = self.encode(ws, F, Cc)
        n = int(n4.item())
        return SparseTensor(f4[:n].clone(), c4[:n].clone(), x.s * 16)


class BEVEncoder(SparseConvEncoder):
    """Identical topology (models/basic_blocks.py:136-171)."""


class DynamicEdgeConv(nn.Module, PrepCache):
    """kNN graph (query = candidates, support = all instances of the scene) + EdgeConv with max
    aggregation (models/basic_blocks.py:98-133)."""

    def __init__(self, F_in, F_out, k=6, num_classes=18):
        super().__init__()
        self.k, self.num_classes, self.F_in, self.F_out = k, num_classes, F_in, F_out
        self.mlp = nn.Sequential(nn.Linear(3 * F_in, F_out), nn.ReLU(), nn.Linear(F_out, F_out))
        self.weight = nn.Sequential(nn.Linear(3 + num_classes + num_classes, 64), nn.ReLU(), nn.Linear(64, F_in))

    def _prepare(self):
        t = lambda lin: lin.weight.detach().float().t().contiguous()      # (in,out) layout
        b = lambda lin: lin.bias.detach().float().contiguous()
        return dict(Ww1=t(self.weight[0]), bw1=b(self.weight[0]), Ww2=t(self.weight[2]), bw2=b(self.weight[2]),
                    Wm1=t(self.mlp[0]), bm1=b(self.mlp[0]), Wm2=t(self.mlp[2]), bm2=b(self.mlp[2]))

    def forward(self, support_xyz, seg_ofs, filtered_index, query_seg, features):
        """support_xyz (S,3), features (S,F_in) fp32; seg_ofs (n_scene+1,) int32 row offsets of each
        scene; filtered_index (M,) int32 query rows; query_seg (M,) int32 scene of each query."""
        p = self.prepared()
        nbr = ops.knn(support_xyz, seg_ofs, filtered_index, query_seg, self.k)
        out = ops.edgeconv(features, support_xyz, filtered_index, nbr, self.num_classes,
                           p['Ww1'], p['bw1'], p['Ww2'], p['bw2'], p['Wm1'], p['bm1'], p['Wm2'], p['bm2'])
        return out, nbr


class SparseCrop(nn.Module):
    """Holder: the crop bounds are fused into the BEV kernel (0 <= xyz < (240,400,80))."""

    def __init__(self, loc_min=None, loc_max=None):
        super().__init__()
        self.loc_min, self.loc_max = loc_min, loc_max


class ToDenseBEVConvolution(nn.Module):
    """Holder for ``kernel`` (n_kernels, Cin, Cout), init U(+-1/sqrt(Cin))
    (models/basic_blocks.py:195-221)."""

    def __init__(self, in_channels, out_channels, shape=(15, 25, 5), offset=(0, 0, 0), z_dim=2, use_bias=False):
        super().__init__()
        assert not use_bias
        self.in_channels, self.out_channels, self.z_dim = in_channels, out_channels, z_dim
        self.n_kernels = int(shape[z_dim])
        self.kernel = nn.Parameter(torch.zeros(self.n_kernels, in_channels, out_channels))
        std = 1. / math.sqrt(in_channels)
        self.kernel.data.uniform_(-std, std)
