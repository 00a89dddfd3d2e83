"""ORACLE shim: torch_geometric.nn.{MessagePassing, knn} (PyG 1.6.x / torch_cluster 1.5.x).
Call sites: models/basic_blocks.py:98,100,120,125."""
import torch
import torch.nn as nn

import sparse_ref as R

knn = R.knn


class MessagePassing(nn.Module):
    """flow source->target: edge_index[0]=j (source), [1]=i (target); x=(x_src,x_dst);
    out = per-target max of message(...) over incoming edges, 0 if none."""

    def __init__(self, aggr='add'):
        super().__init__()
        assert aggr == 'max'

    def propagate(self, edge_index, x, pos):
        j, i = edge_index[0], edge_index[1]
        msg = self.message(x_i=x[1][i], x_j=x[0][j], pos_i=pos[1][i], pos_j=pos[0][j])
        return R.scatter_max(msg, i, x[1].shape[0])
