"""Scene sharding across GPUs: eval-mode referrals are independent, so a collated
batch splits into contiguous blocks of scenes, one block per rank, with NO data-path collective; only
the per-candidate score vectors (variable length per rank) are gathered back in scene order.
``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

from .sparse_tensor import SparseTensor

PER_SCENE_LISTS = ('instance_points', 'instance_obbs', 'instance_class')


def scene_range(n_scenes, rank, world):
    """Contiguous block [lo, hi) of scenes owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_data_dict(data_dict, rank, world):
    """Restrict a collated forward-input dict to this rank's scenes.  (B,...) tensors and per-scene lists
    are sliced; the batched ``lidar`` keeps the rows whose batch index falls in the block and renumbers
    them from 0 (row order preserved)."""
    B = len(data_dict['instance_points'])
    lo, hi = scene_range(B, rank, world)
    out = {}
    for k, v in data_dict.items():
        if k == 'lidar':
            b = v.C[:, 3]
            keep = (b >= lo) & (b < hi)
            C = v.C[keep].clone()
            C[:, 3] -= lo
            out[k] = SparseTensor(v.F[keep], C, v.s)
        elif k in PER_SCENE_LISTS or (isinstance(v, list) and len(v) == B):
            out[k] = v[lo:hi]
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_variable(t, group=None):
    """all_gather of tensors whose first dimension differs per rank; returns the rank-order concatenation."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def gather_outputs(out, keys=('lang_scores', 'seg_scores', 'attribute_scores', 'relation_scores',
                              'scene_scores', 'obj_feats'), group=None):
    """Reassemble per-scene (B,..) and per-candidate (M,..) outputs in the reference's scene order."""
    return {k: gather_variable(out[k].contiguous(), group) for k in keys if k in out}
