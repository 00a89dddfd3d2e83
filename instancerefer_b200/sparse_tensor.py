"""Sparse voxel tensor container with the fields the reference uses from
``torchsparse.SparseTensor`` (<=1.2): ``F`` (N,C) float features, ``C`` (N,4) int32
``[x, y, z, batch]`` (batch LAST), integer stride ``s``, ``.cuda()/.to()`` and ``+``
(models/basic_blocks.py:55,175-182,227-229; lib/dataset.py:261; models/attribute_module.py:70)."""
import torch


class SparseTensor:
    def __init__(self, feats, coords, cur_tensor_stride=1):
        self.F = feats
        self.C = coords
        self.s = cur_tensor_stride
        self.coord_maps = {}
        self.kernel_maps = {}

    def cuda(self, *a, **k):
        return self.to('cuda')

    def to(self, device, *a, **k):
        self.F = torch.as_tensor(self.F).to(device)
        self.C = torch.as_tensor(self.C).to(device)
        return self

    def detach(self):
        self.F = self.F.detach()
        return self

    def __add__(self, other):
        t = SparseTensor(self.F + other.F, self.C, self.s)
        t.coord_maps, t.kernel_maps = self.coord_maps, self.kernel_maps
        return t
