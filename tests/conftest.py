import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def make_args(**kw):
    """config/InstanceRefer.yaml keys consumed on the hot path (flattened like lib/config.py:24-26)."""
    d = dict(language_module='lang_module', attribute_module='attribute_module',
             relation_module='relation_module', scene_module='scene_module', num_classes=18,
             use_bidir=True, voxel_size_ap=0.02, voxel_size_glp=0.05, k=8, use_gt_lang=True)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture(scope='session')
def args():
    return make_args()


@pytest.fixture(scope='session')
def lib_built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope='session')
def state_dict():
    import weights
    return weights.make_state_dict(123)


@pytest.fixture(scope='session')
def gpu_model(lib_built, state_dict, args):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from instancerefer_b200.instancerefer import InstanceRefer
    m = InstanceRefer(7, args)
    m.load_state_dict(state_dict, strict=True)
    return m.cuda().eval()
