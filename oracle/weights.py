"""ORACLE helper: deterministic, machine-independent state_dict with the reference's key names
and shapes (tests/golden/state_dict_spec.json, captured from the reference model built verbatim in
the dev container).  The same dict is loaded (strict) into the reference, handed to
oracle/model_ref.py and loaded (strict) into the CUDA drop-in, so golden outputs travel without
shipping 32 MB of weights.  Inits mirror the reference's (torchsparse conv U(+-1/sqrt(Cin*K)),
Linear/Conv2d/GRU U(+-1/sqrt(fan_in))); BN statistics are randomised so eval-mode BN is
non-trivial."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

SPEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'state_dict_spec.json')


def load_spec(path=SPEC):
    with open(path) as f:
        return json.load(f)


def make_state_dict(seed=123, spec=None, gain=1.0):
    """The generator itself lives with the synthetic-input helpers of the package
    (instancerefer_b200.synthetic.make_state_dict: the product's bench needs random-init weights too and may
    not import the oracle); here it is bound to the reference's key/shape spec."""
    from instancerefer_b200.synthetic import make_state_dict as gen
    return gen(seed, spec or load_spec(), gain)
