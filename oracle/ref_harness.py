"""ORACLE (test infrastructure): run the reference's own ``models/*.py`` VERBATIM on CPU over the
import-name shim in ``oracle/shim``.  Works only where
``/root/reference`` exists (this container) — it is what pins ``oracle/model_ref.py`` and what
``oracle/make_golden.py`` uses to write ``tests/golden/*.npz``.  Nothing on the GPU box may call it.

The reference files are never edited; the environment differences are patched around them:
  * unconditional ``.cuda()`` (models/lang_module.py:60, relation_module.py:94-98,
    attribute_module.py:101)                      -> no-op on a CUDA-less host
  * ``torch.tensor(..., device='cuda')`` (models/scene_module.py:22-23)  -> device dropped
  * ``torch.cuda.IntTensor`` (models/basic_blocks.py:206)               -> torch.IntTensor
  * ``torch.cuda.sparse.FloatTensor`` (models/basic_blocks.py:238)      -> sparse_coo_tensor
  * ``lib/config.py`` (easydict, argv, os.listdir at import)            -> bypassed; args built
    from config/InstanceRefer.yaml flattened exactly like lib/config.py:24-26.
"""
import os
import sys
import types
import importlib
import contextlib

import torch
import yaml

REF = os.environ.get("IR_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REF, "models", "instancerefer.py"))


def load_args(**overrides):
    """YAML sections discarded, keys flattened onto one namespace (lib/config.py:21-26)."""
    with open(os.path.join(REF, "config", "InstanceRefer.yaml")) as f:
        cfg = yaml.safe_load(f)
    flat = {}
    for _, sec in cfg.items():
        flat.update(sec)
    flat.update(overrides)
    return types.SimpleNamespace(**flat)


@contextlib.contextmanager
def cpu_patches():
    """Neutralise the CUDA-isms listed in the module docstring for the duration of the block."""
    if torch.cuda.is_available():
        yield
        return
    saved = (torch.Tensor.cuda, torch.tensor, getattr(torch.cuda, "IntTensor", None),
             torch.cuda.sparse.FloatTensor if hasattr(torch.cuda, "sparse") else None)
    orig_tensor = torch.tensor

    def tensor_nodev(*a, **k):
        k.pop("device", None)
        return orig_tensor(*a, **k)

    def sparse_float(idx, val, size):
        return torch.sparse_coo_tensor(idx, val, size)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.tensor = tensor_nodev
    torch.cuda.IntTensor = torch.IntTensor
    if not hasattr(torch.cuda, "sparse"):
        torch.cuda.sparse = types.SimpleNamespace()
    torch.cuda.sparse.FloatTensor = sparse_float
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.tensor = saved[0], saved[1]
        if saved[2] is not None:
            torch.cuda.IntTensor = saved[2]
        if saved[3] is not None:
            torch.cuda.sparse.FloatTensor = saved[3]


_paths_done = False


def _setup_paths():
    global _paths_done
    if _paths_done:
        return
    for p in (os.path.join(REF, "models"), REF, os.path.join(HERE, "shim"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    _paths_done = True


def shim_sparse_tensor():
    _setup_paths()
    return importlib.import_module("torchsparse").SparseTensor


def build_reference_model(args=None, input_feature_dim=7, seed=123):
    """InstanceRefer(7, args) exactly as scripts/train.py:72-80 does (CPU, shimmed deps)."""
    assert available(), "reference tree not present"
    _setup_paths()
    args = args or load_args()
    for name in ('lang_module', 'attribute_module', 'relation_module', 'scene_module'):
        m = sys.modules.get(name)                       # a drop-in with the same bare name may be cached
        if m is not None and not os.path.abspath(getattr(m, '__file__', '')).startswith(REF):
            del sys.modules[name]
    with cpu_patches():
        mod = importlib.import_module("models.instancerefer")
        torch.manual_seed(seed)
        model = mod.InstanceRefer(input_feature_dim=input_feature_dim, args=args)
    return model, args


def randomize_bn_stats(model, seed=7):
    """Non-trivial eval-mode BN: running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5),
    affine weight ~ U(0.8,1.2), bias ~ N(0,0.05)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    return model


def run_reference(model, data_dict, train=False):
    with cpu_patches():
        model.train(train)
        with torch.set_grad_enabled(train):
            return model(data_dict)


def run_reference_train_step(model, data_dict, config):
    """Reference forward in train mode (Dropout p forced to 0: torch's mask stream cannot be
    reproduced elsewhere), the reference's own lib/loss_helper.get_loss VERBATIM, backward.
    -> data_dict with 'loss', 'ref_loss', 'lang_loss', 'seg_loss'; grads sit on model params."""
    _setup_paths()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.zero_grad()
    out = run_reference(model, data_dict, train=True)
    with cpu_patches():
        orig_ones, orig_zeros = torch.ones, torch.zeros
        torch.ones = lambda *a, **k: (k.pop('device', None), orig_ones(*a, **k))[1]      # lib/loss_helper.py:140
        torch.zeros = lambda *a, **k: (k.pop('device', None), orig_zeros(*a, **k))[1]
        try:
            lh = importlib.import_module('lib.loss_helper')
            out = lh.get_loss(out, config)
        finally:
            torch.ones, torch.zeros = orig_ones, orig_zeros
    out['loss'].backward()
    return out


# ----------------------------------------------------------------------------- scan preprocessing (§8(f)-4)

def _plyfile_stub():
    """The reference's scannet_utils.py imports ``plyfile`` (absent here) and exits without it.  This stand-in serves
    ``PlyData.read`` for the binary little-endian files ``oracle.prepare_ref.write_scan`` writes: ``['vertex']`` with
    ``.count`` / ``.data`` (structured array) and ``['face'].data`` (records whose first field is the index list) — the
    only members data/scannet/scannet_utils.py:97-116 touches."""
    import numpy as np

    class _El:
        def __init__(self, data):
            self.data, self.count = data, len(data)

    class PlyData(dict):
        @staticmethod
        def read(f):
            raw = f.read()
            end = raw.index(b'end_header\n') + len(b'end_header\n')
            head = raw[:end].decode('ascii').split('\n')
            nv = int([l for l in head if l.startswith('element vertex')][0].split()[-1])
            nf = int([l for l in head if l.startswith('element face')][0].split()[-1])
            vt = np.dtype([('x', '<f4'), ('y', '<f4'), ('z', '<f4'), ('red', 'u1'), ('green', 'u1'), ('blue', 'u1'), ('alpha', 'u1')])
            ft = np.dtype([('n', 'u1'), ('v', '<i4', (3,))])
            v = np.frombuffer(raw, vt, nv, end)
            fc = np.frombuffer(raw, ft, nf, end + nv * vt.itemsize)
            out = PlyData()
            out['vertex'] = _El(v)
            out['face'] = _El([(np.array(r['v']),) for r in fc])
            return out

    m = types.ModuleType('plyfile')
    m.PlyData, m.PlyElement = PlyData, object
    return m


def run_reference_prepare(dirs, scan, split, label_map_file, out_prefix, donotcare=()):
    """data/scannet/prepare_data.py ``export_one_scan`` VERBATIM on the scan files under ``dirs`` (written by
    ``oracle.prepare_ref.write_scan``); the module-level names its ``__main__`` block would set are set here.
    Returns the eight saved arrays."""
    import numpy as np
    d = os.path.join(REF, 'data', 'scannet')
    sys.modules['plyfile'] = _plyfile_stub()
    sys.path.insert(0, d)
    try:
        for name in ('scannet_utils', 'load_scannet_data', 'prepare_data'):
            sys.modules.pop(name, None)
        mod = importlib.import_module('prepare_data')
        PR = importlib.import_module('oracle.prepare_ref')
        mod.split, mod.SCANNET_DIR, mod.POINTGROUP_DIR = split, dirs['scannet'], dirs['pointgroup']
        mod.LABEL_MAP_FILE, mod.DONOTCARE_CLASS_IDS = label_map_file, np.array(list(donotcare))
        mod.OBJ_CLASS_IDS, mod.MAX_NUM_POINT = PR.OBJ_CLASS_IDS, PR.MAX_NUM_POINT
        with contextlib.redirect_stdout(open(os.devnull, 'w')):
            mod.export_one_scan(scan, out_prefix)
    finally:
        sys.path.remove(d)
        for name in ('scannet_utils', 'load_scannet_data', 'prepare_data', 'plyfile'):
            sys.modules.pop(name, None)
    keys = ('vert', 'aligned_vert', 'sem_label', 'ins_label', 'sem_label_pg', 'ins_label_pg', 'bbox', 'aligned_bbox')
    return {k: np.load(f'{out_prefix}_{k}.npy') for k in keys}
