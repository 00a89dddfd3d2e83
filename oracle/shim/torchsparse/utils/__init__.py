"""ORACLE shim: torchsparse.utils.{sparse_quantize, sparse_collate_tensors, sparse_collate_fn}."""
import numpy as np
import torch

from torchsparse import SparseTensor
import sparse_ref as R

sparse_quantize = R.sparse_quantize


def sparse_collate_tensors(sparse_tensors):
    c, f = R.sparse_collate([t.C for t in sparse_tensors], [t.F for t in sparse_tensors])
    return SparseTensor(f, c, sparse_tensors[0].s)


def sparse_collate_fn(batch):
    if isinstance(batch[0], dict):
        out = {}
        for name in batch[0].keys():
            v0 = batch[0][name]
            if isinstance(v0, dict):
                out[name] = sparse_collate_fn([s[name] for s in batch])
            elif isinstance(v0, np.ndarray):
                out[name] = torch.stack([torch.from_numpy(s[name]) for s in batch], 0)
            elif torch.is_tensor(v0):
                out[name] = torch.stack([s[name] for s in batch], 0)
            elif isinstance(v0, SparseTensor):
                out[name] = sparse_collate_tensors([s[name] for s in batch])
            else:
                out[name] = [s[name] for s in batch]
        return out
    return batch
