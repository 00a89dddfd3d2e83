"""Packed input format: what the reference's ``collate_fn`` + per-sample CPU
``sparse_quantize`` (lib/dataset.py:207-261,456-469) hand to the forward, re-laid for one H2D copy.

``collate_packed(samples)`` takes the per-sample dicts of the reference's ``__getitem__`` (it reads only
``point_clouds``, ``instance_points/_obbs/_class``, ``lang_feat``, ``lang_len``, ``object_cat``,
``point_min/_max`` and the ``ref_*`` labels) and returns

  * ``point_clouds``    (B, P, 7) fp32 pinned — the scene is voxelised @5 cm ON THE GPU in ``to_forward_dict``
                        (``ir_voxelize_points``: first point wins, first-occurrence order, batch index = scene),
                        replacing the loader's numpy ``sparse_quantize`` and torchsparse's collate;
  * ``packed_points``   (sum n_b, 1024, 7) fp32 pinned + ``inst_ofs`` (B+1): every instance of every scene in
                        ONE buffer; ``instance_points`` stays available as zero-copy numpy views, so the dict
                        still satisfies the reference's list-of-arrays contract;
  * stacked tensors for everything else.

Host only (numpy / torch CPU) except ``to_forward_dict`` which runs the voxeliser on the device."""
import numpy as np
import torch

from . import ops
from .sparse_tensor import SparseTensor

STACK_KEYS = ('lang_feat', 'lang_len', 'object_cat', 'point_min', 'point_max', 'ref_center_label',
              'ref_size_residual_label', 'ref_heading_class_label', 'ref_heading_residual_label',
              'ref_size_class_label', 'unique_multiple')
PACKED = '_ir_packed'


def collate_packed(samples, pin=True):
    B = len(samples)
    counts = [len(s['instance_points']) for s in samples]
    ofs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ppi, fdim = samples[0]['instance_points'][0].shape
    packed = torch.empty(int(ofs[-1]), ppi, fdim, dtype=torch.float32)
    clouds = torch.empty((B,) + tuple(samples[0]['point_clouds'].shape), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        packed, clouds = packed.pin_memory(), clouds.pin_memory()
    pk, cl = packed.numpy(), clouds.numpy()
    for b, s in enumerate(samples):
        cl[b] = s['point_clouds']
        for j, p in enumerate(s['instance_points']):
            pk[ofs[b] + j] = p
    out = dict(point_clouds=clouds, packed_points=packed, inst_ofs=ofs,
               instance_points=[[pk[ofs[b] + j] for j in range(counts[b])] for b in range(B)],   # zero-copy views
               instance_obbs=[list(s['instance_obbs']) for s in samples],
               instance_class=[list(s['instance_class']) for s in samples])
    for k in STACK_KEYS:
        if k in samples[0]:
            out[k] = torch.from_numpy(np.stack([np.asarray(s[k]) for s in samples], 0))
    return out


def voxelize_scenes(point_clouds, voxel):
    """(B,P,7) fp32 CUDA -> SparseTensor with F (V,7), C (V,4) int32 [x,y,z,b]: the loader's
    sparse_quantize(pc[:, :3], pc, voxel) per scene + batch-index collate, on the device."""
    B, P, fdim = point_clouds.shape
    dev = point_clouds.device
    n = B * P
    lib = ops._lib.load()
    scratch = torch.empty(lib.ir_voxelize_points_scratch_bytes(n), dtype=torch.uint8, device=dev)
    coords = torch.empty(n, 4, dtype=torch.int32, device=dev)
    feats = torch.empty(n, fdim, dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    cloud = torch.arange(B, dtype=torch.int32, device=dev)
    ops.call("ir_voxelize_points", ops._p(point_clouds.contiguous(), torch.float32), ops._p(cloud), B, P, fdim, float(voxel),
             ops._p(scratch), ops._p(coords), ops._p(feats), ops._p(count), ops._stream())
    v = int(count.item())
    return SparseTensor(feats[:v], coords[:v])


def to_forward_dict(batch, device, voxel_size_glp=0.05):
    """Packed batch -> the dict ``InstanceRefer.forward`` consumes: tensors moved with non-blocking
    copies, ``lidar`` built on the GPU, the packed instance buffer attached for the candidate pack."""
    d = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) and k not in ('packed_points',) else v)
         for k, v in batch.items() if k != 'inst_ofs'}
    d['lidar'] = voxelize_scenes(d.pop('point_clouds'), voxel_size_glp)
    d[PACKED] = dict(points=batch['packed_points'], ofs=batch['inst_ofs'])
    d.pop('packed_points', None)
    return d
