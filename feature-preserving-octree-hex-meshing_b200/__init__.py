"""fpohm_b200 — B200-native geometry core of feature-preserving octree hex meshing.

Python host-side mirror of the C-ABI in include/fpohm.h (ctypes, no torch types cross the boundary).
The directory name carries hyphens, so import it through the repo-root shim:  `import fpohm_b200`.
"""
from .api import *  # noqa: F401,F403
from . import api, procedural  # noqa: F401


def __getattr__(name):
    # torch is plumbing: only pulled in when the multi-GPU helpers are asked for
    if name == "sharding":
        import importlib
        return importlib.import_module(".sharding", __name__)
    raise AttributeError(name)
