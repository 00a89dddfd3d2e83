"""Packed loader format: host layout on CPU; on the GPU the device voxeliser against the
loader's numpy sparse_quantize restatement (bit-exact) and the forward from a packed batch against the
forward from the reference-format dict (identical)."""
import numpy as np
import pytest
import torch

from instancerefer_b200 import synthetic


def samples(seed=3, B=3):
    out = []
    for b in range(B):
        s = synthetic.make_scene(seed + b, num_points=5000, n_inst=6 + b, n_cand=3 + b % 2, n_tokens=5 + b)
        s['point_clouds'] = s.pop('point_cloud')
        out.append(s)
    return out


def test_collate_packed_layout():
    from instancerefer_b200.loader import collate_packed
    ss = samples()
    b = collate_packed(ss, pin=False)
    assert b['point_clouds'].shape == (3, 5000, 7) and b['packed_points'].shape == (6 + 7 + 8, 1024, 7)
    assert b['inst_ofs'].tolist() == [0, 6, 13, 21]
    base = b['packed_points'].numpy()
    for i, s in enumerate(ss):
        assert np.array_equal(b['point_clouds'][i].numpy(), s['point_clouds'])
        for j, p in enumerate(s['instance_points']):
            v = b['instance_points'][i][j]
            assert np.array_equal(v, p) and np.shares_memory(v, base)              # zero-copy views
        assert b['instance_class'][i] == s['instance_class']
    assert b['lang_feat'].shape == (3, 126, 300) and b['lang_len'].tolist() == [5, 6, 7]
    assert b['ref_center_label'].shape == (3, 3) and b['object_cat'].dtype == torch.int64


@pytest.mark.gpu
def test_device_voxeliser_bit_exact(lib_built):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from instancerefer_b200.loader import voxelize_scenes
    rng = np.random.default_rng(0)
    pc = np.concatenate([rng.uniform(-2, 6, (4, 20000, 3)), rng.uniform(-0.5, 0.5, (4, 20000, 4))], -1).astype(np.float32)
    st = voxelize_scenes(torch.from_numpy(pc).cuda(), 0.05)
    C, F = [], []
    for b in range(4):
        c, f = synthetic.quantize_first(pc[b, :, :3], pc[b], 0.05)
        C.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int32)], 1))
        F.append(f)
    assert np.array_equal(st.C.cpu().numpy(), np.concatenate(C, 0))
    assert np.array_equal(st.F.cpu().numpy(), np.concatenate(F, 0))


@pytest.mark.gpu
def test_forward_from_packed_batch_is_identical(gpu_model):
    from instancerefer_b200 import SparseTensor
    from instancerefer_b200.loader import collate_packed, to_forward_dict
    ss = samples(seed=11, B=2)
    packed = collate_packed(ss)
    # reference-format dict of the same samples (loader-side numpy voxelisation)
    C, F = [], []
    for b, s in enumerate(ss):
        c, f = synthetic.quantize_first(s['point_clouds'][:, :3], s['point_clouds'], 0.05)
        C.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int32)], 1))
        F.append(f)
    ref = dict(lidar=SparseTensor(torch.from_numpy(np.concatenate(F, 0)).cuda(), torch.from_numpy(np.concatenate(C, 0)).cuda()),
               lang_feat=packed['lang_feat'].cuda(), lang_len=packed['lang_len'].cuda(), object_cat=packed['object_cat'].cuda(),
               point_min=packed['point_min'].cuda(), instance_points=[s['instance_points'] for s in ss],
               instance_obbs=[s['instance_obbs'] for s in ss], instance_class=[s['instance_class'] for s in ss])
    with torch.no_grad():
        a = gpu_model(to_forward_dict(packed, 'cuda'))
        b = gpu_model(ref)
    torch.cuda.synchronize()
    for k in ('attribute_scores', 'relation_scores', 'scene_scores', 'lang_scores', 'seg_scores', 'obj_feats'):
        assert torch.equal(a[k], b[k]), k
