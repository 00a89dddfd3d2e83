"""ORACLE shim: torchsparse.SparseTensor (<=1.2 API: F, C [x,y,z,batch] int32, integer s,
coord_maps, kernel_maps, a + b).  Used at models/basic_blocks.py:55,175-182,227-229,
lib/dataset.py:261, models/attribute_module.py:70."""
import numpy as np
import torch


class SparseTensor:
    def __init__(self, feats, coords, cur_tensor_stride=1):
        self.F = feats
        self.C = coords
        self.s = cur_tensor_stride
        self.coord_maps = {}
        self.kernel_maps = {}

    def check(self):
        if self.s not in self.coord_maps:
            self.coord_maps[self.s] = self.C

    def cuda(self, *a, **k):
        return self.to("cuda")

    def to(self, device, *a, **k):
        # CPU oracle: device moves are no-ops when CUDA is absent / when asked for cpu
        if torch.is_tensor(self.F) and (str(device) == "cpu" or torch.cuda.is_available()):
            self.F = self.F.to(device)
            self.C = self.C.to(device)
        return self

    def detach(self):
        self.F = self.F.detach()
        return self

    def __add__(self, other):
        t = SparseTensor(self.F + other.F, self.C, self.s)
        t.coord_maps = self.coord_maps
        t.kernel_maps = self.kernel_maps
        return t
