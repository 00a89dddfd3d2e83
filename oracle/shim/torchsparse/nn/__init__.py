"""ORACLE shim: torchsparse.nn.{Conv3d, BatchNorm, ReLU, GlobalMaxPooling} (Appendix A)."""
import math
import numpy as np
import torch
import torch.nn as nn

from torchsparse import SparseTensor
import sparse_ref as R


class Conv3d(nn.Module):
    """kernel (K,inc,outc), no bias, init U(-s,s) with s=1/sqrt(inc*K).
    stride 1: out coords = in coords; stride 2 (ks 2): out coords = downsampled."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1,
                 bias=False, transpose=False):
        super().__init__()
        assert not transpose and dilation == 1 and not bias
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        K = kernel_size ** 3
        shape = (K, in_channels, out_channels) if K > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        std = 1. / math.sqrt(in_channels * K)
        self.kernel.data.uniform_(-std, std)

    def forward(self, x):
        ks, s = self.kernel_size, self.stride
        if ks == 1 and s == 1:
            out = SparseTensor(x.F @ self.kernel, x.C, x.s)
            out.coord_maps, out.kernel_maps = x.coord_maps, x.kernel_maps
            return out
        key = 'k%s_os%d_s%d_d%d' % (ks, x.s, s, self.dilation)
        x.check()
        if key not in x.kernel_maps:
            Cin = x.C.numpy()
            if s == 1:
                Cout = Cin
            else:
                new_s = x.s * s
                if new_s not in x.coord_maps:
                    co, _ = R.downsample_coords(Cin, x.s)
                    x.coord_maps[new_s] = torch.from_numpy(co)
                Cout = x.coord_maps[new_s].numpy()
            x.kernel_maps[key] = R.build_kmap(Cin, Cout, ks, x.s) + (Cout.shape[0],)
        ii, oo, kofs, n_out = x.kernel_maps[key]
        F = R.spconv(x.F, self.kernel, ii, oo, kofs, n_out)
        out = SparseTensor(F, x.C if s == 1 else x.coord_maps[x.s * s], x.s * s)
        out.coord_maps, out.kernel_maps = x.coord_maps, x.kernel_maps
        return out


class BatchNorm(nn.BatchNorm1d):
    def forward(self, x):
        out = SparseTensor(super().forward(x.F), x.C, x.s)
        out.coord_maps, out.kernel_maps = x.coord_maps, x.kernel_maps
        return out


class ReLU(nn.ReLU):
    def forward(self, x):
        out = SparseTensor(super().forward(x.F), x.C, x.s)
        out.coord_maps, out.kernel_maps = x.coord_maps, x.kernel_maps
        return out


class GlobalMaxPooling(nn.Module):
    def forward(self, x):
        return R.global_max_pool(x.F, x.C[:, 3])
