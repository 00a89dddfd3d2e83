"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, ctypes signatures cover the header, the workspace layout is sane, the Python modules keep
the reference's state_dict layout, and the product path fails loudly without a GPU (no fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'instancerefer_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(ir_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_header_symbol(lib_built):
    from instancerefer_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/instancerefer_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == names            # ctypes table mirrors the header one to one
    assert _lib.load().ir_version() == 100


def header_prototypes():
    """name -> number of parameters, parsed from the header's prototypes."""
    txt = open(os.path.join(ROOT, 'include', 'instancerefer_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(ir_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;', txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ('', 'void') else len(args.split(','))
    return out


def test_ctypes_argument_counts_match_the_header():
    """Every ctypes signature has exactly as many arguments as its C prototype (a mismatch would shift every
    following argument silently)."""
    from instancerefer_b200 import _lib
    protos = header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], (name, len(argtypes), protos[name])


def test_struct_sizes_match_the_library(lib_built):
    """The ctypes mirrors of the ABI structs have the layout the compiler gave the C structs (probed through
    behaviour: layouts written by the library are self-consistent and arena sizes are positive)."""
    from instancerefer_b200 import _lib
    lay = _lib.EncoderTrainLayout()
    n = (ctypes.c_int32 * 5)(4000, 2000, 900, 300, 80)
    assert _lib.call('ir_encoder_train_layout', 8192, n, 7, ctypes.byref(lay)) == 0
    offs = [*lay.off_y, *lay.off_out, *lay.off_mean, *lay.off_rstd, lay.off_bn_scratch, *lay.off_tr_out, *lay.off_tr_slot,
            *lay.off_grad, lay.off_wt, lay.off_absmax]
    assert len(set(offs)) == len(offs) and all(0 <= o < lay.total_bytes and o % 256 == 0 for o in offs)
    cap = _lib.EncoderTrainLayout()
    assert _lib.call('ir_encoder_train_layout', 8192, None, 7, ctypes.byref(cap)) == 0 and cap.total_bytes > lay.total_bytes
    P = _lib.LangParams()
    P.B, P.L, P.E_in, P.D, P.H, P.n_cls = 2, 9, 300, 256, 128, 18
    a = _lib.load().ir_lang_train_arena_bytes(ctypes.byref(P))
    of, oa = ctypes.c_int64(0), ctypes.c_int64(0)
    assert _lib.call('ir_lang_train_view', ctypes.byref(P), ctypes.byref(of), ctypes.byref(oa)) == 0
    assert 0 < of.value < oa.value < a
    assert _lib.load().ir_mlp_head_arena_bytes(64, 256) > 64 * 256 * 4 * 5
    assert _lib.load().ir_bn_scratch_floats(128) == 128 * 2 * 128
    assert _lib.load().ir_scene_tail_arena_bytes(200, 4) > 4 * 299 * 1152 * 4 * 2
    assert ctypes.sizeof(_lib.SceneTail) == 8 + 8 + 16 + 8 + 13 * 8 and ctypes.sizeof(_lib.SceneTailGrads) == 72


def test_encoder_layout_host_only(lib_built):
    from instancerefer_b200 import _lib
    L = _lib.EncoderLayout()
    assert _lib.call('ir_encoder_layout', 32768, ctypes.byref(L)) == 0
    assert L.cap == 65536 and L.total_bytes == _lib.load().ir_encoder_workspace_bytes(32768)
    offs = [L.off_nlvl, L.off_kcount, L.off_scan, L.off_keys, L.off_vals, *L.off_coords, L.off_pslot,
            *L.off_k3_in, *L.off_k3_slot, *L.off_k2_in, *L.off_k2_slot, L.off_feat0, *L.off_feat, L.off_T]
    assert len(set(offs)) == len(offs) and all(o % 1024 == 0 and 0 <= o < L.total_bytes for o in offs)
    with pytest.raises(_lib.IrError):
        _lib.call('ir_encoder_layout', 0, ctypes.byref(L))


def test_state_dict_layout_matches_reference(args):
    from instancerefer_b200.instancerefer import InstanceRefer
    spec = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_spec.json')))
    sd = InstanceRefer(7, args).state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, dtype in spec:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert sum(v.numel() for k, v in sd.items() if v.is_floating_point() and 'running' not in k) == 8021176


def test_dropin_module_names_resolve():
    import importlib
    import sys
    d = os.path.join(ROOT, 'instancerefer_b200', 'dropin')
    sys.path.insert(0, d)
    try:
        for name, cls in (('lang_module', 'LangModule'), ('attribute_module', 'AttributeModule'),
                          ('relation_module', 'RelationModule'), ('scene_module', 'SceneModule')):
            sys.modules.pop(name, None)
            assert hasattr(importlib.import_module(name), cls)
    finally:
        sys.path.remove(d)
        for name in ('lang_module', 'attribute_module', 'relation_module', 'scene_module'):
            sys.modules.pop(name, None)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib_built, args, state_dict):
    from instancerefer_b200 import SparseTensor, _lib, synthetic
    from instancerefer_b200.instancerefer import InstanceRefer
    m = InstanceRefer(7, args).eval()
    b = synthetic.make_batch(1, batch_size=1, num_points=2000, n_inst=4, n_cand=4, n_tokens=5)
    with pytest.raises(_lib.IrError):
        m(synthetic.to_data_dict(b, SparseTensor, 'cpu'))


def test_eval_only_helpers_refuse_train_mode(args):
    """The eval-mode helpers fold BatchNorm running statistics; in train mode they must fail loudly instead of
    silently using stale statistics (the modules' forward() dispatches to the training path)."""
    from instancerefer_b200.basic_blocks import SparseConvEncoder, require_eval
    enc = SparseConvEncoder(7)
    enc.train()
    with pytest.raises(RuntimeError):
        require_eval(enc)


def test_training_path_has_no_cpu_fallback(args):
    """Train-mode forward on a CUDA-less host raises (no torch/CPU fallback for the training step either)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from instancerefer_b200 import SparseTensor, _lib, synthetic
    from instancerefer_b200.instancerefer import InstanceRefer
    m = InstanceRefer(7, args).train()
    b = synthetic.make_batch(3, batch_size=1, num_points=2000, n_inst=4, n_cand=3, n_tokens=4)
    with pytest.raises(_lib.IrError):
        m(synthetic.to_data_dict(b, SparseTensor, 'cpu'))
