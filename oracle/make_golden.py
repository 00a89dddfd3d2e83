"""ORACLE tooling (dev container only): run the reference's models/*.py VERBATIM (over the shim)
and commit small golden fixtures under tests/golden/.

    python oracle/make_golden.py

Writes  state_dict_spec.json  (reference key names / shapes / dtypes, 260 entries)
        golden_<case>.npz     (generator config + every tensor the reference forward writes)
Weights come from oracle/weights.make_state_dict(seed) loaded STRICT into the reference model, so
the fixtures also pin checkpoint-key compatibility."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, HERE)

import ref_harness as H          # noqa: E402
import weights as W              # noqa: E402
from instancerefer_b200 import synthetic as S   # noqa: E402

GOLD = os.path.join(HERE, '..', 'tests', 'golden')

CASES = {
    # BASELINE.json configs[0]-like plumbing case: 8 instances, 10 tokens
    'c1_small': dict(seed=11, batch_size=1, num_points=4000, n_inst=8, n_cand=8, n_tokens=10),
    # ragged batch: scene 1 has a single candidate (skipped in the score vectors, Appendix B.1)
    'ragged_b3': dict(seed=21, batch_size=3, num_points=6000, n_inst=10, n_cand=[4, 1, 3],
                      n_tokens=[7, 12, 3]),
    # negative coordinates (dropped by SparseCrop, Appendix B.5) + a short kNN segment (n_inst<k)
    'shifted_b2': dict(seed=31, batch_size=2, num_points=5000, n_inst=5, n_cand=[3, 5],
                       n_tokens=[5, 9], shift=(-1.0, -2.0, 0.0)),
}

KEYS = ['lang_feat', 'atten_attr', 'atten_rel', 'atten_scene', 'lang_cls_feats', 'lang_attr_feats',
        'lang_rel_feats', 'lang_scene_feats', 'lang_scores', 'obj_feats', 'attribute_scores',
        'relation_scores', 'scene_scores', 'seg_scores', 'vis_atten']


def make_case_batch(cfg):
    cfg = dict(cfg)
    shift = cfg.pop('shift', None)
    seed = cfg.pop('seed')
    b = S.make_batch(seed, **cfg)
    if shift is not None:
        b = S.shift_batch(b, shift)
    return b


def main():
    model, args = H.build_reference_model()
    spec = [(k, list(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, 'state_dict_spec.json'), 'w') as f:
        json.dump(spec, f)
    sd = W.make_state_dict(123, spec)
    model.load_state_dict(sd, strict=True)
    ST = H.shim_sparse_tensor()
    for name, cfg in CASES.items():
        b = make_case_batch(cfg)
        out = H.run_reference(model, S.to_data_dict(b, ST))
        arrs = {k: out[k].detach().numpy() for k in KEYS}
        arrs['num_filtered_objs'] = np.asarray(out['num_filtered_objs'], np.int64)
        for i, o in enumerate(out['pred_obb_batch']):
            arrs[f'pred_obb_{i}'] = np.asarray(o, np.float64)
        arrs['config_json'] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLD, f'golden_{name}.npz'), **arrs)
        print(name, {k: arrs[k].shape for k in ('attribute_scores', 'obj_feats')})


TRAIN_CASES = {
    'train_b2': dict(seed=5, batch_size=2, num_points=6000, n_inst=8, n_cand=[4, 3], n_tokens=[6, 9]),
    'train_ragged_b3': dict(seed=21, batch_size=3, num_points=6000, n_inst=10, n_cand=[4, 1, 3], n_tokens=[7, 12, 3]),
}


def main_train():
    """Golden fixtures of one training iteration (reference models + lib/loss_helper.get_loss run
    verbatim, Dropout p=0, oracle/train_ref.SyntheticConfig as the dataset config): loss terms, the
    L2 norm of every parameter gradient, full gradients of the small tensors, updated BN statistics."""
    import train_ref as T
    spec = W.load_spec()
    ST = H.shim_sparse_tensor()
    for name, cfg in TRAIN_CASES.items():
        model, args = H.build_reference_model()
        model.load_state_dict(W.make_state_dict(123, spec), strict=True)
        b = make_case_batch(cfg)
        out = H.run_reference_train_step(model, S.to_data_dict(b, ST), T.SyntheticConfig())
        arrs = {k: out[k].detach().numpy().reshape(-1) for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss')}
        names, norms = [], []
        for k, p in model.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            names.append(k)
            norms.append(float(g.double().norm()))
            if g.numel() <= 4096:
                arrs['grad/' + k] = g.numpy()
        arrs['grad_names'] = np.frombuffer(json.dumps(names).encode(), dtype=np.uint8)
        arrs['grad_norms'] = np.asarray(norms, np.float64)
        for k, v in model.state_dict().items():
            if k.endswith(('running_mean', 'running_var')) and v.numel() <= 64:
                arrs['bn/' + k] = v.numpy()
        arrs['config_json'] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLD, f'golden_{name}.npz'), **arrs)
        print(name, float(out['loss']), len(names))


PREPARE_CASES = {                       # name -> (synth_scan kwargs, split the reference is told, directory written)
    'prepare_a': (dict(seed=11, n_verts=3000, n_faces=5600, n_objects=9, n_props=7), 'train', 'val'),
    'prepare_b': (dict(seed=12, n_verts=700, n_faces=900, n_objects=4, n_props=3), 'val', 'test'),
    'prepare_nolabels': (dict(seed=13, n_verts=500, n_faces=800, n_objects=3, n_props=2), 'test', 'test'),
}


def main_prepare():
    """Golden fixtures of the scan preprocessing: the reference's data/scannet/prepare_data.py export_one_scan run
    verbatim on synthetic scan files (scratch under oracle/_ref/, git-ignored); inputs + the eight saved arrays."""
    import shutil
    from oracle import prepare_ref as PR
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'prepare_tmp')
    shutil.rmtree(root, ignore_errors=True)
    os.makedirs(root)
    tsv = os.path.join(root, 'labels.tsv')
    PR.write_label_map(tsv)
    for name, (kw, split, folder) in PREPARE_CASES.items():
        s = PR.synth_scan(**kw)
        labelled = 'nolabels' not in name
        dirs = PR.write_scan(root, 'scene0000_00', s, split=folder, with_labels=labelled)
        out = H.run_reference_prepare(dirs, 'scene0000_00', split, tsv, os.path.join(root, name))
        arrs = {'out/' + k: v for k, v in out.items()}
        for k in ('xyz', 'rgb', 'faces', 'seg_indices', 'matrix', 'masks', 'cls'):
            arrs['in/' + k] = s[k]
        arrs['in/seg_groups_json'] = np.frombuffer(json.dumps(s['seg_groups']).encode(), dtype=np.uint8)
        arrs['config_json'] = np.frombuffer(json.dumps({'synth': kw, 'split': split, 'folder': folder, 'labelled': labelled}).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLD, f'golden_{name}.npz'), **arrs)
        print(name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})
        shutil.rmtree(os.path.join(root, 'scans'))
        shutil.rmtree(os.path.join(root, 'PointGroupInst'))
    shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    if '--prepare' in sys.argv:
        main_prepare()
        sys.exit(0)
    if '--train' in sys.argv:
        main_train()
        sys.exit(0)
    main()
