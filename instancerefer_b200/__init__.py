"""instancerefer_b200 — B200-native (sm_100a) implementation of InstanceRefer's hot path
(models/instancerefer.py forward): hand-written CUDA kernels behind a C ABI
(include/instancerefer_b200.h) with drop-in Python modules that keep the reference's module
names, constructor signatures, data_dict keys and state_dict layout.

Drop-in use under the reference's scripts: put ``instancerefer_b200/dropin`` on ``sys.path`` ahead
of the reference's ``models/`` (or set the YAML keys to ``instancerefer_b200.lang_module`` etc.).
"""
from .sparse_tensor import SparseTensor  # noqa: F401

__all__ = ["SparseTensor"]
