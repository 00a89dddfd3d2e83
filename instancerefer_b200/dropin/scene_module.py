"""Drop-in shim: put this directory on sys.path AHEAD of the reference's models/ and the reference's
own models/instancerefer.py (importlib.import_module('scene_module'), models/instancerefer.py:20-34) loads the
B200 implementation with no YAML change."""
from instancerefer_b200.scene_module import *  # noqa: F401,F403
