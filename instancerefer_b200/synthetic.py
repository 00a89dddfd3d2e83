"""Seeded synthetic ScanRefer-shaped inputs.

Mirrors what the reference's loader hands to ``InstanceRefer.forward`` — the dict built in
``lib/dataset.py:255-298`` and batched by ``collate_fn`` (``lib/dataset.py:456-469``) — without
ScanNet/ScanRefer/GloVe data: an axis-aligned room with its min corner at the origin (so the
reference's ``SparseCrop`` keeps everything, ``models/scene_module.py:22-23``), box-shaped
objects standing on the floor, exactly ``num_points`` points with
``[x,y,z, r,g,b, height]`` features (``lib/dataset.py:105,121-123``), 1024-point instance
samples (``lib/dataset.py:224``) and GloVe-scaled language features (``lib/dataset.py:73-86``).

numpy only; no CUDA, no oracle imports.  The scene voxelisation at 5 cm is the *loader's* job
in the reference (host numpy, ``lib/dataset.py:255-261``); ``quantize_first`` restates it here so
the generated ``lidar`` tensor looks like the loader's output.
"""
import json as _json
import os

import numpy as np

MAX_DES_LEN = 126          # lib/config.py:74
NUM_CLASSES = 18           # config/InstanceRefer.yaml:8
LABEL_KEYS = ('ref_center_label', 'ref_size_residual_label', 'ref_heading_class_label',
              'ref_heading_residual_label', 'ref_size_class_label')


def quantize_first(xyz, feats, voxel):
    """floor(xyz / voxel) in float64; one row per voxel = first point in input order,
    rows kept in first-occurrence order.  -> (coords int32 (V,3), feats[idx])."""
    disc = np.floor(xyz.astype(np.float64) / voxel).astype(np.int64)
    key = ((disc[:, 0] + 32768) << 32) | ((disc[:, 1] + 32768) << 16) | (disc[:, 2] + 32768)
    _, first = np.unique(key, return_index=True)
    first.sort()
    return disc[first].astype(np.int32), feats[first]


def _box_surface(rng, n, centre, size):
    """n points on the 5 visible faces (top + 4 sides) of an axis-aligned box."""
    sx, sy, sz = size
    areas = np.array([sx * sy, sx * sz, sx * sz, sy * sz, sy * sz])
    face = rng.choice(5, size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, (n, 3)) * size
    u[face == 0, 2] = 0.5 * sz
    u[face == 1, 1] = -0.5 * sy
    u[face == 2, 1] = 0.5 * sy
    u[face == 3, 0] = -0.5 * sx
    u[face == 4, 0] = 0.5 * sx
    return u + centre


def _room_surface(rng, n, room):
    lx, ly, lz = room
    areas = np.array([lx * ly, lx * lz, lx * lz, ly * lz, ly * lz])
    face = rng.choice(5, size=n, p=areas / areas.sum())
    p = rng.uniform(0.0, 1.0, (n, 3)) * np.array(room)
    p[face == 0, 2] = 0.0
    p[face == 1, 1] = 0.0
    p[face == 2, 1] = ly - 1e-3
    p[face == 3, 0] = 0.0
    p[face == 4, 0] = lx - 1e-3
    return p


def make_scene(seed, num_points=40000, n_inst=32, n_cand=32, n_tokens=20,
               room=(8.0, 6.0, 3.0), target_class=None):
    """One (scene, utterance) sample in the loader's per-sample format."""
    rng = np.random.default_rng(seed)
    room = np.asarray(room, np.float64)
    per_obj = max(16, min(int(0.8 * num_points / max(n_inst, 1)), 2048))
    n_room = num_points - per_obj * n_inst
    assert n_room > 0
    pts = [_room_surface(rng, n_room, room)]
    labels = [np.zeros(n_room, np.int64)]
    for j in range(n_inst):
        size = rng.uniform(0.4, 1.2, 3)
        centre = np.array([rng.uniform(0.5 * size[0], room[0] - 0.5 * size[0]),
                           rng.uniform(0.5 * size[1], room[1] - 0.5 * size[1]),
                           0.5 * size[2]])
        pts.append(_box_surface(rng, per_obj, centre, size))
        labels.append(np.full(per_obj, j + 1, np.int64))
    xyz = np.concatenate(pts, 0)
    labels = np.concatenate(labels, 0)
    perm = rng.permutation(num_points)
    xyz, labels = xyz[perm], labels[perm]
    rgb = rng.uniform(-0.5, 0.5, (num_points, 3))
    height = xyz[:, 2] - np.percentile(xyz[:, 2], 0.99)
    pc = np.concatenate([xyz, rgb, height[:, None]], 1).astype(np.float32)

    if target_class is None:
        target_class = int(rng.integers(0, NUM_CLASSES))
    others = [c for c in range(NUM_CLASSES) if c != target_class]
    cls = np.array([target_class] * n_cand + [others[int(rng.integers(0, 17))]
                                              for _ in range(n_inst - n_cand)], np.int64)
    cls = cls[rng.permutation(n_inst)] if n_inst else cls

    inst_pts, inst_obbs = [], []
    for j in range(n_inst):
        x = pc[labels == j + 1]
        p = x[:, :3]
        centre = 0.5 * (p.min(0) + p.max(0))
        size = p.max(0) - p.min(0)
        inst_obbs.append(np.concatenate([centre, size, np.array([0])]).astype(np.float64))
        choice = rng.choice(x.shape[0], 1024, replace=x.shape[0] < 1024)
        inst_pts.append(np.ascontiguousarray(x[choice]))

    lang = np.zeros((MAX_DES_LEN, 300), np.float32)
    lang[:n_tokens] = rng.normal(0.0, 0.4, (n_tokens, 300)).astype(np.float32)
    # training labels (lib/dataset.py:263-298): the referred object is one of the target-class
    # instances; its box is the instance OBB, slightly perturbed so the IoU argmax is non-trivial.
    # Drawn last so the forward inputs above are unchanged for a given seed.
    cand_ids = [j for j in range(n_inst) if cls[j] == target_class]
    ref_j = cand_ids[int(rng.integers(0, len(cand_ids)))] if cand_ids else 0
    ref_obb = inst_obbs[ref_j] if n_inst else np.zeros(7)
    ref_center = (ref_obb[:3] + rng.normal(0.0, 0.03, 3)).astype(np.float32)
    ref_size = (ref_obb[3:6] * rng.uniform(0.9, 1.1, 3)).astype(np.float32)
    return dict(point_cloud=pc, instance_points=inst_pts, instance_obbs=inst_obbs,
                instance_class=[int(c) for c in cls], object_cat=np.int64(target_class),
                lang_feat=lang, lang_len=np.int64(n_tokens),
                point_min=pc[:, :3].min(0), point_max=pc[:, :3].max(0),
                ref_center_label=ref_center, ref_size_residual_label=ref_size,
                ref_heading_class_label=np.int64(0), ref_heading_residual_label=np.float32(0),
                ref_size_class_label=np.int64(0), ref_obj_index=np.int64(ref_j))


def make_batch(seed, batch_size=1, voxel_size_glp=0.05, n_cand=32, n_tokens=20, **kw):
    """Batch in the collated layout.  ``n_cand`` / ``n_tokens`` may be ints or per-scene lists.
    Returns plain numpy + nested lists; ``to_data_dict`` turns it into the forward's dict."""
    scenes = []
    for b in range(batch_size):
        nc = n_cand[b] if isinstance(n_cand, (list, tuple)) else n_cand
        nt = n_tokens[b] if isinstance(n_tokens, (list, tuple)) else n_tokens
        scenes.append(make_scene(seed + b, n_cand=nc, n_tokens=nt, **kw))
    coords, feats = [], []
    for b, s in enumerate(scenes):
        c, f = quantize_first(s['point_cloud'][:, :3], s['point_cloud'], voxel_size_glp)
        coords.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int32)], 1))
        feats.append(f)
    return dict(
        lidar_coords=np.concatenate(coords, 0).astype(np.int32),
        lidar_feats=np.concatenate(feats, 0).astype(np.float32),
        lang_feat=np.stack([s['lang_feat'] for s in scenes], 0),
        lang_len=np.stack([s['lang_len'] for s in scenes], 0),
        object_cat=np.stack([s['object_cat'] for s in scenes], 0),
        point_min=np.stack([s['point_min'] for s in scenes], 0),
        point_max=np.stack([s['point_max'] for s in scenes], 0),
        **{k: np.stack([s[k] for s in scenes], 0) for k in LABEL_KEYS},
        instance_points=[s['instance_points'] for s in scenes],
        instance_obbs=[s['instance_obbs'] for s in scenes],
        instance_class=[s['instance_class'] for s in scenes],
    )


def shift_batch(batch, shift):
    """Translate a whole batch (points, obbs, lidar re-quantised) — used to exercise negative
    coordinates, which the reference's SparseCrop silently drops."""
    shift = np.asarray(shift, np.float32)
    out = dict(batch)
    out['instance_points'] = [[np.concatenate([p[:, :3] + shift, p[:, 3:]], 1).astype(np.float32)
                               for p in scene] for scene in batch['instance_points']]
    out['instance_obbs'] = [[np.concatenate([o[:3] + shift.astype(np.float64), o[3:]]) for o in scene]
                            for scene in batch['instance_obbs']]
    c = batch['lidar_coords'].copy()
    c[:, :3] += np.round(shift / 0.05).astype(np.int32)
    f = batch['lidar_feats'].copy()
    f[:, :3] += shift
    out['lidar_coords'], out['lidar_feats'] = c, f
    out['point_min'] = batch['point_min'] + shift
    out['point_max'] = batch['point_max'] + shift
    if 'ref_center_label' in batch:
        out['ref_center_label'] = batch['ref_center_label'] + shift
    return out


def to_data_dict(batch, sparse_tensor_cls, device='cpu'):
    """Build the dict ``InstanceRefer.forward`` consumes (after ``lib/solver.py:242-245`` moved
    the whitelisted tensors to ``device``)."""
    import torch
    lidar = sparse_tensor_cls(torch.from_numpy(batch['lidar_feats']).to(device),
                              torch.from_numpy(batch['lidar_coords']).to(device))
    return dict(
        lidar=lidar,
        lang_feat=torch.from_numpy(batch['lang_feat']).to(device),
        lang_len=torch.from_numpy(batch['lang_len']).to(device),
        object_cat=torch.from_numpy(batch['object_cat']).to(device),
        point_min=torch.from_numpy(batch['point_min']).to(device),
        point_max=torch.from_numpy(batch['point_max']).to(device),
        **{k: torch.from_numpy(np.asarray(batch[k])).to(device) for k in LABEL_KEYS if k in batch},
        instance_points=batch['instance_points'],
        instance_obbs=batch['instance_obbs'],
        instance_class=batch['instance_class'],
    )


# ----------------------------------------------------------------------------- synthetic weights / dataset config

def make_state_dict(seed=123, spec=None, gain=1.0, model=None):
    """Deterministic, machine-independent random-init state_dict with the reference's key names and shapes
    (there is no network for checkpoints).  ``spec`` = [(key, shape, dtype-string)] — e.g.
    tests/golden/state_dict_spec.json, captured from the reference model — or derived from ``model``.
    Inits mirror the reference's (torchsparse conv U(+-1/sqrt(Cin*K)), Linear/Conv2d/GRU
    U(+-1/sqrt(fan_in))); BN statistics are randomised so eval-mode BN is non-trivial.
    Every tensor has its own generator seeded from (seed, crc32(key)), so the same dict can be rebuilt
    anywhere (reference run, CPU oracle, CUDA drop-in)."""
    import math
    import zlib

    import torch
    if spec is None:
        spec = [(k, list(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    sd = {}
    for key, shape, dtype in spec:
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        if dtype == 'torch.int64':
            sd[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'running_mean':
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == 'running_var':
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == 'kernel':                       # (K,Cin,Cout) sparse conv / (5,128,128) BEV
            bound = 1.0 / math.sqrt(shape[1] * (shape[0] if shape[0] in (8, 27) else 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound * gain
        elif len(shape) >= 2:                        # Linear / Conv2d / GRU matrices
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            if '.gru.' in key:
                fan_in = 128
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in) * gain
        elif leaf == 'weight':                       # BN / LN scale
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        else:                                        # biases
            t = torch.randn(shape, generator=g) * 0.05
        sd[key] = t.float()
    return sd


class SyntheticConfig:
    """Stand-in for the reference's ScannetDatasetConfig (which needs ScanNet meta files) with the one method
    ``get_loss`` / ``get_eval`` call (data/scannet/model_util_scannet.py:174-181): ONE size class with zero
    mean size and ONE heading bin, i.e. box size = size residual, heading = heading residual."""

    def param2obb_batch(self, center, heading_class, heading_residual, size_class, size_residual):
        obb = np.zeros((heading_class.shape[0], 7))
        obb[:, 0:3] = center
        obb[:, 3:6] = size_residual
        obb[:, 6] = heading_residual * -1
        return obb


# ----------------------------------------------------------------------------- synthetic scan on disk (scan preprocessing)

RAW_LABELS = ['chair', 'table', 'wall', 'floor', 'cabinet', 'sofa', 'door', 'window']


def synth_scan(seed, n_verts=3000, n_faces=5600, n_objects=9, n_props=7, n_segments=None):
    """Seeded scan contents in ScanNet's shapes: vertices with colour, a face list with repeated vertices, an
    over-segmentation, labelled objects (one segment listed by two objects), an axis alignment
    and PointGroup proposals that overlap."""
    rng = np.random.default_rng(seed)
    n_segments = n_segments or max(n_objects * 4, n_verts // 40)
    xyz = (rng.random((n_verts, 3)) * [8.0, 6.0, 3.0] - [4.0, 3.0, 0.2]).astype(np.float32)
    rgb = rng.integers(0, 256, (n_verts, 3)).astype(np.uint8)
    faces = rng.integers(0, n_verts, (n_faces, 3)).astype(np.int32)
    if n_faces:
        faces[: min(5, n_faces)] = faces[0]                      # degenerate / repeated faces
        faces[min(7, n_faces - 1)] = [3, 3, 3]                   # zero-area face
    seg_indices = rng.integers(0, n_segments, n_verts).astype(np.int64) * 3 + 1     # sparse segment ids
    seg_ids = np.unique(seg_indices)
    rng.shuffle(seg_ids)
    per = np.array_split(seg_ids[: max(1, int(len(seg_ids) * 0.8))], n_objects)
    groups = []
    for o in range(n_objects):
        segs = [int(s) for s in per[o]]
        if o == n_objects - 2 and o > 0 and groups[0]['segments']:
            segs.append(groups[0]['segments'][0])               # a segment claimed by two objects: the later one wins
        groups.append({'objectId': o, 'label': RAW_LABELS[int(rng.integers(0, len(RAW_LABELS)))], 'segments': segs})
    th = float(rng.random() * 6.28)
    matrix = np.array([[np.cos(th), -np.sin(th), 0, rng.normal()], [np.sin(th), np.cos(th), 0, rng.normal()],
                       [0, 0, 1, rng.normal() * 0.1], [0, 0, 0, 1]], np.float64)
    masks = (rng.random((n_props, n_verts)) < 0.15).astype(np.uint8)
    cls = rng.integers(3, 40, n_props).astype(np.int64)
    return {'xyz': xyz, 'rgb': rgb, 'faces': faces, 'seg_indices': seg_indices, 'seg_groups': groups, 'matrix': matrix,
            'masks': masks, 'cls': cls}


def write_label_map(path):
    """A cut-down scannetv2-labels.combined.tsv with the two columns the preprocessing reads."""
    nyu = {'chair': 5, 'table': 7, 'wall': 1, 'floor': 2, 'cabinet': 3, 'sofa': 6, 'door': 8, 'window': 9}
    with open(path, 'w') as f:
        f.write('id\traw_category\tcategory\tnyu40id\n')
        for i, (k, v) in enumerate(nyu.items()):
            f.write(f'{i + 1}\t{k}\t{k}\t{v}\n')
    return nyu


def write_scan(root, scan, s, split='val', with_labels=True):
    """Write the files the reference reads for one scan (prepare_data.py:167-176, :38-47):
    <root>/scans/<scan>/<scan>_vh_clean_2.ply, .aggregation.json, _vh_clean_2.0.010000.segs.json, <scan>.txt and
    <root>/PointGroupInst/<train|val|test>/<scan>.txt + one mask file per proposal.  Returns the directory dict."""
    d = os.path.join(root, 'scans', scan)
    os.makedirs(d, exist_ok=True)
    n, nf = s['xyz'].shape[0], s['faces'].shape[0]
    head = ('ply\nformat binary_little_endian 1.0\ncomment synthetic\n'
            f'element vertex {n}\nproperty float x\nproperty float y\nproperty float z\n'
            'property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n'
            f'element face {nf}\nproperty list uchar int vertex_indices\nend_header\n')
    vt = np.zeros(n, dtype=[('x', '<f4'), ('y', '<f4'), ('z', '<f4'), ('red', 'u1'), ('green', 'u1'), ('blue', 'u1'), ('alpha', 'u1')])
    vt['x'], vt['y'], vt['z'] = s['xyz'].T
    vt['red'], vt['green'], vt['blue'] = s['rgb'].T
    vt['alpha'] = 255
    ft = np.zeros(nf, dtype=[('n', 'u1'), ('v', '<i4', (3,))])
    ft['n'] = 3
    ft['v'] = s['faces']
    with open(os.path.join(d, scan + '_vh_clean_2.ply'), 'wb') as f:
        f.write(head.encode('ascii'))
        f.write(vt.tobytes())
        f.write(ft.tobytes())
    if with_labels:
        with open(os.path.join(d, scan + '.aggregation.json'), 'w') as f:
            _json.dump({'sceneId': scan, 'segGroups': [dict(g, id=g['objectId']) for g in s['seg_groups']]}, f)
    with open(os.path.join(d, scan + '_vh_clean_2.0.010000.segs.json'), 'w') as f:
        _json.dump({'sceneId': scan, 'segIndices': [int(x) for x in s['seg_indices']]}, f)
    with open(os.path.join(d, scan + '.txt'), 'w') as f:
        if s.get('matrix') is not None:
            f.write('axisAlignment = ' + ' '.join(repr(float(x)) for x in s['matrix'].reshape(-1)) + '\n')
        f.write('colorHeight = 968\nnumDepthFrames = 100\n')
    pg = os.path.join(root, 'PointGroupInst', split)
    os.makedirs(os.path.join(pg, 'predicted_masks'), exist_ok=True)
    with open(os.path.join(pg, scan + '.txt'), 'w') as f:
        for i, c in enumerate(s['cls']):
            rel = f'predicted_masks/{scan}_{i:03d}.txt'
            f.write(f'{rel} {int(c)} {0.5 + 0.01 * i:.4f}\n')
            np.savetxt(os.path.join(pg, rel), s['masks'][i], fmt='%d')
    return {'scannet': os.path.join(root, 'scans'), 'pointgroup': os.path.join(root, 'PointGroupInst')}
