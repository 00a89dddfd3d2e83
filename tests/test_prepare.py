"""Scan preprocessing: host parsers on CPU; the CUDA path against the oracle, the golden fixtures
of the reference run verbatim, and through the files (synthetic scan on disk -> the eight .npy files)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import prepare_ref as PR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KEYS = ('vert', 'aligned_vert', 'sem_label', 'ins_label', 'sem_label_pg', 'ins_label_pg', 'bbox', 'aligned_bbox')
NYU = {'chair': 5, 'table': 7, 'wall': 1, 'floor': 2, 'cabinet': 3, 'sofa': 6, 'door': 8, 'window': 9}


def cases():
    return sorted(glob.glob(os.path.join(GOLD, 'golden_prepare_*.npz')))


def load_case(path):
    z = np.load(path)
    cfg = json.loads(bytes(z['config_json']).decode())
    s = {k[3:]: z[k] for k in z.files if k.startswith('in/') and k != 'in/seg_groups_json'}
    s['seg_groups'] = json.loads(bytes(z['in/seg_groups_json']).decode())
    return z, cfg, s


# ----------------------------------------------------------------------------- host side (CPU)

def test_host_parsers_read_what_write_scan_wrote(tmp_path):
    from instancerefer_b200 import prepare_data as P
    s = PR.synth_scan(seed=5, n_verts=400, n_faces=700, n_objects=4, n_props=3)
    tsv = str(tmp_path / 'labels.tsv')
    assert PR.write_label_map(tsv) == NYU
    dirs = PR.write_scan(str(tmp_path), 'scene0002_00', s, split='val')
    d = os.path.join(dirs['scannet'], 'scene0002_00')
    vertex, faces = P.read_ply(os.path.join(d, 'scene0002_00_vh_clean_2.ply'))
    assert np.array_equal(np.stack([vertex['x'], vertex['y'], vertex['z']], 1), s['xyz'])
    assert np.array_equal(np.stack([vertex['red'], vertex['green'], vertex['blue']], 1), s['rgb'])
    assert np.array_equal(faces, s['faces']) and faces.dtype == np.int32
    assert np.array_equal(P.read_segmentation(os.path.join(d, 'scene0002_00_vh_clean_2.0.010000.segs.json')), s['seg_indices'])
    groups = P.read_aggregation(os.path.join(d, 'scene0002_00.aggregation.json'))
    assert groups == s['seg_groups']
    assert np.array_equal(P.read_axis_alignment(os.path.join(d, 'scene0002_00.txt')), s['matrix'])
    assert P.read_label_mapping(tsv) == NYU
    masks, cls = P.read_pointgroup(dirs['pointgroup'], 'scene0002_00', 'train')       # train -> falls back to val/
    assert np.array_equal(masks, s['masks']) and np.array_equal(cls, s['cls'])
    with pytest.raises(FileNotFoundError):
        P.read_pointgroup(dirs['pointgroup'], 'scene0002_00', 'test')


def test_mask_reader_and_scan_sharding(tmp_path):
    from instancerefer_b200 import prepare_data as P
    m = (np.random.RandomState(0).rand(500) < 0.3).astype(np.int64)
    np.savetxt(tmp_path / 'a.txt', m, fmt='%d')
    np.savetxt(tmp_path / 'b.txt', m.astype(np.float64))                  # '0.000000e+00' lines: the loadtxt route
    np.savetxt(tmp_path / 'c.txt', m * 12, fmt='%d')                      # two-digit entries: the loadtxt route
    for name in ('a.txt', 'b.txt', 'c.txt'):
        assert np.array_equal(P.read_mask(str(tmp_path / name)), m != 0), name
    names = [f'scene{i:04d}_00' for i in range(11)]
    parts = [P.shard_scans(names, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == names and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert P.shard_scans(names, 0, 1) == names


def test_ascii_ply(tmp_path):
    from instancerefer_b200 import prepare_data as P
    p = tmp_path / 'a.ply'
    p.write_text('ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n'
                 'property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n'
                 'element face 1\nproperty list uchar int vertex_indices\nend_header\n'
                 '0 0 0 1 2 3 255\n1 0 0 4 5 6 255\n0 1 0.5 7 8 9 255\n3 0 1 2\n')
    v, f = P.read_ply(str(p))
    assert v['z'][2] == np.float32(0.5) and v['blue'].tolist() == [3, 6, 9] and f.tolist() == [[0, 1, 2]]


def test_segment_tables_match_oracle():
    from instancerefer_b200 import prepare_data as P
    s = PR.synth_scan(seed=8, n_verts=900, n_faces=10, n_objects=7, n_props=1)
    n_tab = int(s['seg_indices'].max()) + 1
    a = P.segment_tables(s['seg_groups'], NYU, n_tab)
    b = PR.segment_tables(s['seg_groups'], NYU, n_tab)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[2].tolist() == [b[2][o] for o in range(1, 8)]
    with pytest.raises(KeyError):
        P.segment_tables([{'objectId': 0, 'label': 'unknown thing', 'segments': [1]}], NYU, 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_prepare_has_no_cpu_fallback():
    from instancerefer_b200 import _lib, prepare_data as P
    s = PR.synth_scan(seed=1, n_verts=50, n_faces=60, n_objects=2, n_props=1)
    with pytest.raises(_lib.IrError):
        P.export_arrays(np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1), s['faces'], s['matrix'], s['seg_indices'],
                        s['seg_groups'], NYU, s['masks'], s['cls'])


# ----------------------------------------------------------------------------- device side

def run_gpu(s, labelled=True, **kw):
    from instancerefer_b200 import prepare_data as P
    vertex = np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1)
    A = P.export_arrays(vertex, s['faces'], s['matrix'], s['seg_indices'] if labelled else None,
                        s['seg_groups'] if labelled else None, NYU, s['masks'], s['cls'])
    return A, P.filter_scan(A, **kw)


def assert_same(out, ref, aligned_tol=False):
    """Bit-exact, dtypes included.  The aligned coordinates are an fp64 dot product rounded to fp32: the device sums
    it in a fixed FMA order, numpy's BLAS in its own, so an fp32 rounding tie may fall the other way — at most one
    fp32 ulp on at most 1e-5 of the values (none observed)."""
    for k in KEYS:
        a, b = out[k], ref[k]
        assert a.dtype == b.dtype and a.shape == b.shape, k
        if k in ('aligned_vert', 'aligned_bbox') and not np.array_equal(a, b):
            bad = a != b
            assert bad.mean() <= 1e-5 and np.abs(a - b).max() <= 4 * np.spacing(np.abs(b).max()), k
        else:
            assert np.array_equal(a, b), k


@pytest.mark.gpu
@pytest.mark.parametrize('path', cases())
def test_prepare_matches_reference_fixture(lib_built, path):
    z, cfg, s = load_case(path)
    _, out = run_gpu(s, cfg['labelled'])
    assert_same(out, {k: z['out/' + k] for k in KEYS})


@pytest.mark.gpu
@pytest.mark.parametrize('seed,n,nf,no,npr', [(21, 5000, 9000, 12, 8), (22, 1, 0, 1, 1), (23, 64, 3, 2, 0), (24, 2049, 4000, 40, 20)])
def test_prepare_matches_oracle(lib_built, seed, n, nf, no, npr):
    s = PR.synth_scan(seed=seed, n_verts=n, n_faces=nf, n_objects=no, n_props=npr, n_segments=max(no * 3, n // 30 + 1))
    vertex = np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1)
    ref = PR.export_one_scan(PR.export(vertex, s['faces'], s['matrix'], s['seg_indices'], s['seg_groups'], NYU, s['masks'], s['cls']))
    _, out = run_gpu(s)
    assert_same(out, ref)


@pytest.mark.gpu
def test_prepare_filters_and_subsample(lib_built):
    """DONOTCARE filter (order-preserving compaction on the device) and the fixed-size sub-sample, against the oracle
    with the same kept classes and the same draw; without an alignment matrix the aligned arrays equal the raw ones."""
    s = PR.synth_scan(seed=31, n_verts=7000, n_faces=12000, n_objects=10, n_props=5)
    s['matrix'] = None
    vertex = np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1)
    ex = PR.export(vertex, s['faces'], None, s['seg_indices'], s['seg_groups'], NYU, s['masks'], s['cls'])
    A, out = run_gpu(s)
    assert_same(out, PR.export_one_scan(ex))
    assert np.array_equal(out['vert'], out['aligned_vert']) and np.array_equal(out['bbox'], out['aligned_bbox'])
    from instancerefer_b200 import prepare_data as P
    idx, m = P.keep_index(A.sem, [1, 2, 0])
    sem = ex[2]
    want = np.nonzero(~np.isin(sem, [1, 2, 0]))[0]
    assert m == len(want) and np.array_equal(idx[:m].cpu().numpy(), want)
    got = P.select_rows(A.vert, idx, m=m).cpu().numpy()
    assert np.array_equal(got, ex[0][want])
    ch = np.random.RandomState(3).choice(7000, 2500, replace=False)
    out2 = P.filter_scan(A, choices=ch)
    assert_same(out2, PR.export_one_scan(ex, choices=ch))
    out3 = P.filter_scan(A, max_num_point=3000, rng=np.random.RandomState(4))
    assert out3['vert'].shape == (3000, 9) and out3['ins_label_pg'].shape == (3000,)
    ch3 = np.random.RandomState(4).choice(7000, 3000, replace=False)
    assert np.array_equal(out3['vert'], ex[0][ch3])


@pytest.mark.gpu
def test_prepare_full_size_scan(lib_built):
    """A ScanNet-sized scan (150 k vertices, 300 k faces, 60 objects, 80 proposals) against the oracle."""
    s = PR.synth_scan(seed=41, n_verts=150000, n_faces=300000, n_objects=60, n_props=80)
    vertex = np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1)
    ex = PR.export(vertex, s['faces'], s['matrix'], s['seg_indices'], s['seg_groups'], NYU, s['masks'], s['cls'])
    ch = np.random.RandomState(0).choice(150000, 50000, replace=False)
    A, out = run_gpu(s, choices=ch)
    assert_same(out, PR.export_one_scan(ex, choices=ch))


@pytest.mark.gpu
def test_prepare_through_files(lib_built, tmp_path):
    """export_one_scan from the scan's files to the eight .npy files, against the golden case the reference produced
    from the same files' contents."""
    from instancerefer_b200 import prepare_data as P
    z, cfg, s = load_case(os.path.join(GOLD, 'golden_prepare_a.npz'))
    tsv = str(tmp_path / 'labels.tsv')
    PR.write_label_map(tsv)
    dirs = PR.write_scan(str(tmp_path), 'scene0000_00', s, split=cfg['folder'])
    prefix = str(tmp_path / 'scene0000_00')
    P.export_one_scan('scene0000_00', prefix, dirs['scannet'], dirs['pointgroup'], tsv, split=cfg['split'])
    assert_same({k: np.load(f'{prefix}_{k}.npy') for k in KEYS}, {k: z['out/' + k] for k in KEYS})
