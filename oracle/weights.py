"""ORACLE helper: deterministic, machine-independent state_dict with the reference's key names
and shapes (tests/golden/state_dict_spec.json, captured from the reference model built verbatim in
the dev container).  The same dict is loaded (strict) into the reference, handed to
oracle/model_ref.py and loaded (strict) into the CUDA drop-in, so golden outputs travel without
shipping 32 MB of weights.  Inits mirror the reference's (torchsparse conv U(+-1/sqrt(Cin*K)),
Linear/Conv2d/GRU U(+-1/sqrt(fan_in))); BN statistics are randomised so eval-mode BN is
non-trivial."""
import json
import math
import os
import zlib

import torch

SPEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'state_dict_spec.json')


def load_spec(path=SPEC):
    with open(path) as f:
        return json.load(f)


def make_state_dict(seed=123, spec=None, gain=1.0):
    spec = spec or load_spec()
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
            for s in shape[1:]:
                fan_in *= s
            if '.gru.' in key:
                fan_in = 128
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in) * gain
        elif leaf == 'weight':                       # BN / LN scale
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        else:                                        # biases
            t = torch.randn(shape, generator=g) * 0.05
        sd[key] = t.float()
    return sd
