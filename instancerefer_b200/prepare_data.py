"""Scan preprocessing on the GPU — drop-in for the reference's ``data/scannet/prepare_data.py``.

Same command line (``--split --scannet_path --pointgroupinst_path --output_path``), same input files (ScanNet
``*_vh_clean_2.ply``, ``*.aggregation.json``, ``*_vh_clean_2.0.010000.segs.json``, ``<scan>.txt``, the label map TSV,
PointGroup's per-scan proposal list + mask files) and the same eight ``.npy`` files per scan, bit for bit
(``*_vert / _aligned_vert / _sem_label / _ins_label / _sem_label_pg / _ins_label_pg / _bbox / _aligned_bbox``).

Host side (here): file parsing and the small per-object dictionaries.  Device side (``csrc/prepare.cu`` through the
C ABI): vertex normals, axis alignment, per-vertex labels, per-object boxes, proposal labels, the class filter and the
row gathers.  There is no CPU implementation of the array work: without the CUDA library the calls raise.

Differences from the reference, all on the host side: PLY files are read by ``read_ply`` (no ``plyfile`` dependency);
``export`` takes ``split`` as an argument instead of a module global; the sub-sampling draw of ``export_one_scan``
(`np.random.choice`, prepare_data.py:204) is made by the caller-supplied ``rng`` (default ``np.random``, as there).
"""
import argparse
import csv
import datetime
import json
import os

import numpy as np
import torch

from . import _lib

OBJ_CLASS_IDS = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 24, 25, 26, 27, 28,
                          29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40])   # wall, floor, ceiling excluded (:248-250)
DONOTCARE_CLASS_IDS = np.array([])                                            # prepare_data.py:247
MAX_NUM_POINT = 50000                                                         # prepare_data.py:251

_PLY_TYPES = {'char': 'i1', 'uchar': 'u1', 'short': 'i2', 'ushort': 'u2', 'int': 'i4', 'uint': 'u4', 'float': 'f4',
              'double': 'f8', 'int8': 'i1', 'uint8': 'u1', 'int16': 'i2', 'uint16': 'u2', 'int32': 'i4',
              'uint32': 'u4', 'float32': 'f4', 'float64': 'f8'}


# ----------------------------------------------------------------------------- file parsing (host)

def read_ply(path):
    """Vertex table (structured array with x y z red green blue ...) and triangle list (f,3) int32 of a PLY file —
    what data/scannet/scannet_utils.py:97-113 takes from ``plyfile``.  binary_little_endian and ascii."""
    with open(path, 'rb') as f:
        raw = f.read()
    end = raw.index(b'end_header')
    end = raw.index(b'\n', end) + 1
    lines = [l.strip() for l in raw[:end].decode('ascii').splitlines()]
    if not lines or lines[0] != 'ply':
        raise ValueError(f'{path}: not a PLY file')
    fmt = [l.split()[1] for l in lines if l.startswith('format')][0]
    elements, cur = [], None
    for l in lines:
        t = l.split()
        if t[:1] == ['element']:
            cur = {'name': t[1], 'count': int(t[2]), 'props': []}
            elements.append(cur)
        elif t[:1] == ['property']:
            cur['props'].append(t[1:])
    if fmt not in ('binary_little_endian', 'ascii'):
        raise ValueError(f'{path}: unsupported PLY format {fmt}')
    out, off = {}, end
    tokens = raw[end:].split() if fmt == 'ascii' else None
    tpos = 0
    for e in elements:
        is_list = [p[0] == 'list' for p in e['props']]
        if not any(is_list):
            dt = np.dtype([(p[-1], '<' + _PLY_TYPES[p[0]]) for p in e['props']])
            if fmt == 'ascii':
                k = len(e['props'])
                vals = np.array(tokens[tpos:tpos + k * e['count']], dtype=np.float64).reshape(e['count'], k)
                tpos += k * e['count']
                arr = np.zeros(e['count'], dt)
                for j, p in enumerate(e['props']):
                    arr[p[-1]] = vals[:, j]
            else:
                arr = np.frombuffer(raw, dt, e['count'], off)
                off += dt.itemsize * e['count']
            out[e['name']] = arr
        else:
            if len(e['props']) != 1:
                raise ValueError(f'{path}: element {e["name"]}: a list property next to others is not supported')
            _, ct, it, _name = e['props'][0]
            if fmt == 'ascii':
                rows = []
                for _ in range(e['count']):
                    k = int(tokens[tpos])
                    rows.append([int(x) for x in tokens[tpos + 1:tpos + 1 + k]])
                    tpos += 1 + k
                if any(len(r) != 3 for r in rows):
                    raise ValueError(f'{path}: only triangle meshes are supported')
                out[e['name']] = np.asarray(rows, np.int32).reshape(-1, 3)
            else:
                dt = np.dtype([('n', '<' + _PLY_TYPES[ct]), ('v', '<' + _PLY_TYPES[it], (3,))])
                arr = np.frombuffer(raw, dt, e['count'], off)
                off += dt.itemsize * e['count']
                if e['count'] and not np.all(arr['n'] == 3):
                    raise ValueError(f'{path}: only triangle meshes are supported')
                out[e['name']] = np.ascontiguousarray(arr['v']).astype(np.int32)
    return out['vertex'], out.get('face', np.zeros((0, 3), np.int32))


def read_label_mapping(filename, label_from='raw_category', label_to='nyu40id'):
    """data/scannet/scannet_utils.py:56-65."""
    mapping = {}
    with open(filename) as f:
        for row in csv.DictReader(f, delimiter='\t'):
            mapping[row[label_from]] = int(row[label_to])
    first = next(iter(mapping))
    try:
        int(first)
        mapping = {int(k): v for k, v in mapping.items()}
    except ValueError:
        pass
    return mapping


def read_aggregation(filename):
    """data/scannet/load_scannet_data.py:17-32: the annotated objects in file order, ids made 1-based."""
    with open(filename) as f:
        groups = json.load(f)['segGroups']
    return [{'objectId': g['objectId'], 'label': g['label'], 'segments': list(g['segments'])} for g in groups]


def read_segmentation(filename):
    """data/scannet/load_scannet_data.py:35-47 as one array: segment id of every vertex."""
    with open(filename) as f:
        return np.asarray(json.load(f)['segIndices'], np.int64)


def read_axis_alignment(meta_file):
    """prepare_data.py:52-57: the last ``axisAlignment`` line of <scan>.txt as 16 floats, or None."""
    m = None
    with open(meta_file) as f:
        for line in f:
            if 'axisAlignment' in line:
                m = [float(x) for x in line.split('=', 1)[1].split()]
    return None if m is None else np.array(m, np.float64).reshape(4, 4)


def read_mask(path):
    """One PointGroup proposal mask (a text column of 0/1, `np.loadtxt` in prepare_data.py:146) as a bool array.
    Files of single digits, one per line, are decoded from the raw bytes; anything else goes through np.loadtxt."""
    with open(path, 'rb') as f:
        raw = np.frombuffer(f.read(), np.uint8)
    if raw.size and raw.size % 2 == 0 and np.all(raw[1::2] == 10):
        d = raw[0::2]
        if np.all((d >= 48) & (d <= 57)):
            return d != 48
    return np.atleast_1d(np.loadtxt(path)) != 0


def read_pointgroup(pointgroup_file, scene, split):
    """prepare_data.py:38-47,144-148: (masks (n_inst, n_verts) uint8, cls (n_inst,) int32) of the scan's proposals.
    ``train`` scans are looked up under train/ then val/, all others under test/ — as the reference does."""
    dirs = [pointgroup_file + '/train/', pointgroup_file + '/val/'] if split == 'train' else [pointgroup_file + '/test/']
    for i, d in enumerate(dirs):
        if os.path.isfile(d + scene + '.txt') or i == len(dirs) - 1:
            temp_dir = d
            break
    masks, cls = [], []
    with open(temp_dir + scene + '.txt') as f:
        for line in f:
            line = line.rstrip('\n')
            if not line:
                continue
            txt_path, c, _ = line.split(' ')
            masks.append(read_mask(os.path.join(temp_dir, txt_path)))
            cls.append(int(c))
    if not masks:
        return np.zeros((0, 0), np.uint8), np.zeros(0, np.int32)
    return np.stack(masks).astype(np.uint8), np.asarray(cls, np.int32)


def segment_tables(seg_groups, label_map, n_table):
    """prepare_data.py:73-90 as dense per-segment tables (semantic label, 1-based object id; 0 = unannotated) and the
    per-object label.  Labels are applied in order of first appearance, objects in file order; a segment listed twice
    keeps the later assignment.  Unknown labels / segment ids raise KeyError / IndexError like the reference."""
    seg_label = np.zeros(n_table, np.int32)
    seg_object = np.zeros(n_table, np.int32)
    by_label, objects = {}, {}
    for g in seg_groups:
        by_label.setdefault(g['label'], []).extend(g['segments'])
        objects[g['objectId'] + 1] = g['segments']
    for label, segs in by_label.items():
        seg_label[np.asarray(segs, np.int64)] = label_map[label]
    for oid, segs in objects.items():
        seg_object[np.asarray(segs, np.int64)] = oid
    n_obj = len(objects)
    if sorted(objects) != list(range(1, n_obj + 1)):
        raise ValueError('object ids must be 1..n_objects (prepare_data.py:112)')
    obj_label = np.array([seg_label[objects[o][0]] for o in range(1, n_obj + 1)], np.int32)
    return seg_label, seg_object, obj_label


# ----------------------------------------------------------------------------- device work (C ABI)

def _dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda() if not torch.is_tensor(a) else a


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class ScanArrays:
    """The eight arrays of one scan on the device, with the reference's dtypes (uint32 labels are carried as int32
    storage of the same bits; boxes fp64)."""

    def __init__(self, vert, aligned, sem, ins, bbox, aligned_bbox, sem_pg, ins_pg):
        self.vert, self.aligned, self.sem, self.ins = vert, aligned, sem, ins
        self.bbox, self.aligned_bbox, self.sem_pg, self.ins_pg = bbox, aligned_bbox, sem_pg, ins_pg

    def numpy(self):
        u = lambda t: t.cpu().numpy().view(np.uint32)
        return (self.vert.cpu().numpy(), self.aligned.cpu().numpy(), u(self.sem), u(self.ins), self.bbox.cpu().numpy(),
                self.aligned_bbox.cpu().numpy(), u(self.sem_pg), u(self.ins_pg))


def export_arrays(vertex, faces, matrix, seg_indices, seg_groups, label_map, masks, cls):
    """prepare_data.py:30-164 from parsed inputs, all array work on the device -> ScanArrays.
    vertex: structured PLY vertex table or (n,>=6) array x y z r g b; faces (f,3); matrix 4x4 or None; seg_indices
    (n,) + seg_groups (list of dicts) or None for unlabelled scans; masks (n_inst,n) / cls (n_inst,)."""
    if not torch.cuda.is_available():
        raise _lib.IrError('instancerefer_b200.prepare_data needs a CUDA device (no CPU fallback)')
    if getattr(vertex, 'dtype', None) is not None and vertex.dtype.names:
        v6 = np.stack([vertex[k].astype(np.float32) for k in ('x', 'y', 'z', 'red', 'green', 'blue')], 1)
    else:
        v6 = np.asarray(vertex, np.float32)[:, :6]
    n = v6.shape[0]
    faces = np.asarray(faces, np.int32).reshape(-1, 3)
    if faces.size and (faces.min() < 0 or faces.max() >= n):
        raise IndexError('face index out of range')
    n_obj = len(seg_groups) if seg_groups is not None else 0
    dev = torch.device('cuda')
    v9 = torch.zeros(n, 9, dtype=torch.float32, device=dev)
    v9[:, :6] = torch.from_numpy(v6).to(dev)
    scratch = torch.empty(_lib.load().ir_prepare_scratch_bytes(n, faces.shape[0], n_obj), dtype=torch.uint8, device=dev)
    d_faces = _dev(faces, np.int32)
    _lib.call('ir_mesh_normals', _p(v9), n, _p(d_faces), faces.shape[0], _p(scratch), _stream())
    if matrix is not None:
        aligned = torch.empty_like(v9)
        m = np.ascontiguousarray(np.asarray(matrix, np.float64).reshape(16))
        _lib.call('ir_align_vertices', _p(v9), n, m.ctypes.data, _p(aligned), _stream())
    else:
        aligned = v9                                                  # "No axis alignment matrix found" (:67-69)
    sem = torch.zeros(n, dtype=torch.int32, device=dev)
    ins = torch.zeros(n, dtype=torch.int32, device=dev)
    if seg_groups is not None:
        seg_indices = np.asarray(seg_indices, np.int64)
        if seg_indices.shape[0] != n:
            raise ValueError('segIndices and the mesh disagree on the number of vertices')
        top = max((max(g['segments']) for g in seg_groups if g['segments']), default=-1)
        n_tab = int(max(seg_indices.max(initial=-1), top)) + 1
        present = np.zeros(n_tab, bool)
        present[seg_indices] = True
        for g in seg_groups:
            if not present[np.asarray(g['segments'], np.int64)].all():
                raise KeyError('aggregation names a segment without vertices')       # seg_to_verts[seg] (:80)
        seg_label, seg_object, obj_label = segment_tables(seg_groups, label_map, n_tab)
        d_seg, d_sl, d_so = _dev(seg_indices, np.int32), _dev(seg_label, np.int32), _dev(seg_object, np.int32)
        _lib.call('ir_vertex_labels', _p(d_seg), n, _p(d_sl), _p(d_so), n_tab, _p(sem), _p(ins), _stream())
        d_ol = _dev(obj_label, np.int32)
        bbox = torch.empty(n_obj, 8, dtype=torch.float64, device=dev)
        _lib.call('ir_instance_boxes', _p(v9), _p(ins), n, n_obj, _p(d_ol), _p(scratch), _p(bbox), _stream())
        if aligned is v9:
            aligned_bbox = bbox.clone()
        else:
            aligned_bbox = torch.empty_like(bbox)
            _lib.call('ir_instance_boxes', _p(aligned), _p(ins), n, n_obj, _p(d_ol), _p(scratch), _p(aligned_bbox), _stream())
    else:                                                             # placeholders for test scans (:132-139)
        bbox = torch.zeros(1, 8, dtype=torch.float64, device=dev)
        aligned_bbox = torch.zeros(1, 8, dtype=torch.float64, device=dev)
    sem_pg = torch.zeros(n, dtype=torch.int32, device=dev)
    ins_pg = torch.zeros(n, dtype=torch.int32, device=dev)
    masks = np.asarray(masks, np.uint8)
    if masks.shape[0]:
        if masks.shape[1] != n:
            raise ValueError('proposal masks and the mesh disagree on the number of vertices')
        d_m, d_c = _dev(masks, np.uint8), _dev(np.asarray(cls), np.int32)
        _lib.call('ir_pointgroup_labels', _p(d_m), _p(d_c), masks.shape[0], n, _p(sem_pg), _p(ins_pg), _stream())
    return ScanArrays(v9, aligned, sem, ins, bbox, aligned_bbox, sem_pg, ins_pg)


def select_rows(t, idx, count_dev=None, m=None):
    """t[idx[:m]] on the device (ir_gather_rows); idx int64 on the device."""
    m = idx.shape[0] if m is None else m
    flat = t.contiguous().view(t.shape[0], -1)
    out = torch.empty((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _lib.call('ir_gather_rows', _p(flat), flat.shape[1] * flat.element_size(), _p(idx), _p(count_dev), m, _p(out), _stream())
    return out


def keep_index(sem, donotcare):
    """Ascending indices of the vertices whose label is not in ``donotcare`` (prepare_data.py:185) -> (idx, count)."""
    n = sem.shape[0]
    dev = sem.device
    dc = torch.tensor([int(x) for x in np.asarray(donotcare).reshape(-1)], dtype=torch.int32, device=dev)
    scratch = torch.empty(_lib.load().ir_prepare_scratch_bytes(n, 0, 0), dtype=torch.uint8, device=dev)
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.call('ir_keep_index', _p(sem), n, _p(dc) if dc.numel() else None, dc.numel(), _p(scratch), _p(idx), _p(cnt), _stream())
    return idx, int(cnt.item())


def filter_scan(A, donotcare=DONOTCARE_CLASS_IDS, obj_class_ids=OBJ_CLASS_IDS, max_num_point=MAX_NUM_POINT, choices=None,
                rng=None):
    """prepare_data.py:185-212 on a ScanArrays -> dict of the eight numpy arrays ready to save.  ``choices`` overrides
    the random sub-sample (tests); otherwise it is drawn like the reference when more than ``max_num_point`` remain."""
    vert, aligned, sem, ins = A.vert, A.aligned, A.sem, A.ins
    if np.asarray(donotcare).size:
        idx, m = keep_index(A.sem, donotcare)
        vert, aligned, sem, ins = (select_rows(t, idx, m=m) for t in (vert, aligned, sem, ins))
    bbox, aligned_bbox = A.bbox.cpu().numpy(), A.aligned_bbox.cpu().numpy()
    if bbox.shape[0] > 1:
        keep = np.isin(bbox[:, -2], obj_class_ids)
        bbox, aligned_bbox = bbox[keep], aligned_bbox[keep]
    sem_pg, ins_pg = A.sem_pg, A.ins_pg
    N = vert.shape[0]
    if choices is None and N > max_num_point:
        choices = (rng or np.random).choice(N, max_num_point, replace=False)
    if choices is not None:
        ci = torch.from_numpy(np.asarray(choices, np.int64)).to(vert.device)
        vert, aligned, sem, ins, sem_pg, ins_pg = (select_rows(t, ci) for t in (vert, aligned, sem, ins, sem_pg, ins_pg))
    u = lambda t: t.cpu().numpy().view(np.uint32)
    return {'vert': vert.cpu().numpy(), 'aligned_vert': aligned.cpu().numpy(), 'sem_label': u(sem), 'ins_label': u(ins),
            'sem_label_pg': u(sem_pg), 'ins_label_pg': u(ins_pg), 'bbox': bbox, 'aligned_bbox': aligned_bbox}


# ----------------------------------------------------------------------------- the reference's entry points

def export(mesh_file, agg_file, seg_file, meta_file, label_map_file, output_file=None, pointgroup_file=None,
           split='train', _arrays=False):
    """prepare_data.py:30-164.  Returns (mesh_vertices, aligned_vertices, label_ids, instance_ids, instance_bboxes,
    aligned_instance_bboxes, label_ids_pg, instance_ids_pg) as numpy arrays of the reference's dtypes."""
    scene = meta_file.split('/')[-1].split('.')[0]
    masks, cls = read_pointgroup(pointgroup_file, scene, split)
    label_map = read_label_mapping(label_map_file, label_from='raw_category', label_to='nyu40id')
    vertex, faces = read_ply(mesh_file)
    matrix = read_axis_alignment(meta_file)
    if matrix is None:
        print('No axis alignment matrix found')
    if os.path.isfile(agg_file):
        groups, seg = read_aggregation(agg_file), read_segmentation(seg_file)
    else:
        print('use placeholders')
        groups, seg = None, None
    A = export_arrays(vertex, faces, matrix, seg, groups, label_map, masks, cls)
    out = A.numpy()
    if output_file is not None:
        names = ('vert', 'aligned_vert', 'sem_label', 'ins_label', 'bbox', 'aligned_bbox', 'sem_label_pg', 'ins_label_pg')
        for name, a in zip(names, out):
            # the reference writes instance_bboxes under BOTH box names at this point (:161-162); kept
            np.save(output_file + '_' + name + '.npy', out[4] if name == 'aligned_bbox' else a)
    return (A, out) if _arrays else out


def export_one_scan(scan_name, output_filename_prefix, scannet_dir, pointgroup_dir, label_map_file, split='train',
                    donotcare=DONOTCARE_CLASS_IDS, obj_class_ids=OBJ_CLASS_IDS, max_num_point=MAX_NUM_POINT, rng=None):
    """prepare_data.py:167-216: the eight .npy files of one scan."""
    d = os.path.join(scannet_dir, scan_name)
    A, _ = export(os.path.join(d, scan_name + '_vh_clean_2.ply'), os.path.join(d, scan_name + '.aggregation.json'),
                  os.path.join(d, scan_name + '_vh_clean_2.0.010000.segs.json'), os.path.join(d, scan_name + '.txt'),
                  label_map_file, None, pointgroup_dir, split, _arrays=True)
    out = filter_scan(A, donotcare, obj_class_ids, max_num_point, rng=rng)
    if A.bbox.shape[0] > 1:
        print('Num of instances: ', len(np.unique(out['ins_label'])))
        print('Num of care instances: ', out['bbox'].shape[0])
    else:
        print('No semantic/instance annotation for test scenes')
    print('Shape of points: {}'.format(out['vert'].shape))
    for k, a in out.items():
        np.save(output_filename_prefix + '_' + k + '.npy', a)
    return out


def shard_scans(scan_names, rank, world):
    """Scans are independent: rank r of `world` processes takes every world-th scan (no collective)."""
    return list(scan_names)[rank::world]


def batch_export(scan_names, output_folder, rank=0, world=1, **kw):
    """prepare_data.py:219-234; under torchrun (RANK / WORLD_SIZE) each process converts its own share of the scans."""
    if not os.path.exists(output_folder):
        print('Creating new data folder: {}'.format(output_folder))
        os.makedirs(output_folder, exist_ok=True)
    for scan_name in shard_scans(scan_names, rank, world):
        print(scan_name)
        print('-' * 20 + 'begin')
        print(datetime.datetime.now())
        export_one_scan(scan_name, os.path.join(output_folder, scan_name), **kw)
        print('-' * 20 + 'done')


def parse_args(argv=None):
    parser = argparse.ArgumentParser('Data Preparision')
    parser.add_argument('--split', type=str, default='train', choices=['train', 'val', 'test'])
    parser.add_argument('--scannet_path', type=str, default='data/scannet/scans/')
    parser.add_argument('--pointgroupinst_path', type=str, default='PointGroupInst/')
    parser.add_argument('--output_path', type=str, default='pointgroup_data')
    parser.add_argument('--meta_path', type=str, default='meta_data', help='directory of scannetv2_<split>.txt and the label map')
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    names = sorted(line.rstrip() for line in open(os.path.join(args.meta_path, 'scannetv2_%s.txt' % args.split)))
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)) % torch.cuda.device_count())
    batch_export(names, args.output_path, rank=rank, world=world, scannet_dir=args.scannet_path, pointgroup_dir=args.pointgroupinst_path,
                 label_map_file=os.path.join(args.meta_path, 'scannetv2-labels.combined.tsv'), split=args.split)


if __name__ == '__main__':
    main()
