"""ORACLE (test infrastructure, CPU numpy): restatement of the reference's scan preprocessing,
``data/scannet/prepare_data.py:30-216`` with ``data/scannet/scannet_utils.py:18-44,97-116`` and
``data/scannet/load_scannet_data.py:17-47``.  Only ``tests/`` and ``oracle/make_golden.py``
import this file; the product path is ``instancerefer_b200/prepare_data.py`` + ``csrc/prepare.cu``.

Pinned: ``oracle/make_golden.py --prepare`` runs the reference's own ``export`` / ``export_one_scan`` verbatim on a
synthetic scan written to disk (its ``plyfile`` import is served by a reader stub, see ``ref_harness``) and stores
inputs and outputs as ``tests/golden/golden_prepare_*.npz``; ``tests/test_oracle.py`` checks this file against them.

Everything is written on arrays (the parsed contents of the scan's files); ``instancerefer_b200.synthetic.write_scan``
is the synthetic scan generator that produces those files in ScanNet's formats.
"""
import numpy as np

OBJ_CLASS_IDS = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 24, 25, 26, 27, 28,
                          29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40])      # prepare_data.py:248-250
MAX_NUM_POINT = 50000                                                            # prepare_data.py:251


def _unit(a):
    """scannet_utils.py:18-24 — in place, fp32, eps added to the length."""
    lens = np.sqrt(a[:, 0] ** 2 + a[:, 1] ** 2 + a[:, 2] ** 2)
    for c in range(3):
        a[:, c] /= (lens + 1e-8)
    return a


def mesh_normals(xyz, faces):
    """scannet_utils.py:26-44.  The reference accumulates with ``normals[faces[:, c]] += n``, a buffered fancy-index
    update: where several faces name the same vertex in corner role c only the LAST face's normal is added.  Stated
    here explicitly through the index of that last face."""
    xyz = np.asarray(xyz, np.float32)
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    tri = xyz[faces]
    n = _unit(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]).astype(np.float32))
    out = np.zeros_like(xyz)
    for c in range(3):
        last = np.full(xyz.shape[0], -1, np.int64)
        last[faces[:, c]] = np.arange(faces.shape[0])           # ascending: the highest face index stays
        has = last >= 0
        out[has] = out[has] + n[last[has]]
    return _unit(out)


def vertices_from_ply(vertex, faces):
    """scannet_utils.py:97-116: (n,9) fp32 = xyz, rgb, normal.  vertex: (n,>=6) array x y z r g b."""
    v = np.zeros((vertex.shape[0], 9), np.float32)
    v[:, :6] = vertex[:, :6]
    v[:, 6:] = mesh_normals(v[:, :3].copy(), faces)
    return v


def align(vertices, matrix):
    """prepare_data.py:60-66."""
    pts = np.ones((vertices.shape[0], 4))
    pts[:, :3] = vertices[:, :3]
    pts = pts @ np.asarray(matrix, np.float64).reshape(4, 4).T
    out = vertices.copy()
    out[:, :3] = pts[:, :3]
    return out


def segment_tables(seg_groups, label_map, n_table):
    """prepare_data.py:73-90 + load_scannet_data.py:17-32 as dense per-segment tables: semantic label and 1-based
    object id of every segment id (0 = unannotated).  Groups are applied label by label in order of first appearance
    (labels) and in file order (objects); a segment listed twice keeps the later assignment."""
    seg_label = np.zeros(n_table, np.int64)
    seg_object = np.zeros(n_table, np.int64)
    by_label = {}
    for g in seg_groups:
        by_label.setdefault(g['label'], []).extend(g['segments'])
    for label, segs in by_label.items():
        seg_label[np.asarray(segs, np.int64)] = label_map[label]
    objects = {}
    for g in seg_groups:
        objects[g['objectId'] + 1] = list(g['segments'])
    for oid, segs in objects.items():
        seg_object[np.asarray(segs, np.int64)] = oid
    obj_label = {oid: int(seg_label[segs[0]]) for oid, segs in objects.items()}   # label of the object's first vertex
    return seg_label, seg_object, obj_label


def boxes(vertices, instance_ids, obj_label):
    """prepare_data.py:92-131: (n_objects, 8) fp64 rows (cx,cy,cz,dx,dy,dz,label,obj_id-1), fp32 arithmetic."""
    out = np.zeros((len(obj_label), 8))
    for oid, label in obj_label.items():
        pc = vertices[instance_ids == oid, :3]
        if len(pc) == 0:
            continue
        lo, hi = pc.min(0), pc.max(0)
        out[oid - 1] = np.concatenate([((lo + hi) / np.float32(2)).astype(np.float64), (hi - lo).astype(np.float64),
                                       [label, oid - 1]])
    return out


def pointgroup_labels(masks, cls, n):
    """prepare_data.py:141-148: proposals applied in list order, later ones overwrite."""
    label = np.zeros(n, np.uint32)
    inst = np.zeros(n, np.uint32)
    for i, (m, c) in enumerate(zip(masks, cls)):
        inst[m != 0] = i + 1
        label[m != 0] = int(c)
    return label, inst


def export(vertex, faces, matrix, seg_indices, seg_groups, label_map, masks, cls):
    """prepare_data.py:30-164 on parsed arrays -> the same 8-tuple."""
    mesh = vertices_from_ply(vertex, faces)
    aligned = align(mesh, matrix) if matrix is not None else mesh
    n = mesh.shape[0]
    if seg_groups is not None:
        seg_indices = np.asarray(seg_indices, np.int64)
        n_tab = int(max(seg_indices.max(initial=-1), max((max(g['segments']) for g in seg_groups if g['segments']), default=-1))) + 1
        seg_label, seg_object, obj_label = segment_tables(seg_groups, label_map, n_tab)
        label_ids = seg_label[seg_indices].astype(np.uint32)
        instance_ids = seg_object[seg_indices].astype(np.uint32)
        bb = boxes(mesh, instance_ids, obj_label)
        bb_al = boxes(aligned, instance_ids, obj_label)
    else:                                                       # test scans: placeholders (prepare_data.py:132-139)
        label_ids = np.zeros(n, np.uint32)
        instance_ids = np.zeros(n, np.uint32)
        bb, bb_al = np.zeros((1, 8)), np.zeros((1, 8))
    label_pg, inst_pg = pointgroup_labels(masks, cls, n)
    return mesh, aligned, label_ids, instance_ids, bb, bb_al, label_pg, inst_pg


def export_one_scan(exported, donotcare=(), choices=None):
    """prepare_data.py:167-216 after ``export``: class filter of vertices and boxes; ``choices`` stands for the
    reference's ``np.random.choice(N, MAX_NUM_POINT, replace=False)`` when N > MAX_NUM_POINT."""
    mesh, aligned, sem, ins, bb, bb_al, sem_pg, ins_pg = exported
    mask = np.logical_not(np.isin(sem, np.asarray(donotcare)))
    mesh, aligned, sem, ins = mesh[mask], aligned[mask], sem[mask], ins[mask]
    if bb.shape[0] > 1:
        keep = np.isin(bb[:, -2], OBJ_CLASS_IDS)
        bb, bb_al = bb[keep], bb_al[keep]
    if choices is not None:
        mesh, aligned, sem, ins, sem_pg, ins_pg = (a[choices] for a in (mesh, aligned, sem, ins, sem_pg, ins_pg))
    return {'vert': mesh, 'aligned_vert': aligned, 'sem_label': sem, 'ins_label': ins, 'sem_label_pg': sem_pg,
            'ins_label_pg': ins_pg, 'bbox': bb, 'aligned_bbox': bb_al}


# synthetic scan generator (shared with the tests and tools): lives in the package, re-exported here
from instancerefer_b200.synthetic import RAW_LABELS, synth_scan, write_label_map, write_scan   # noqa: E402,F401
