"""fpohm_b200 — B200-native geometry core of feature-preserving octree hex meshing.

Python host-side mirror of the C-ABI in include/fpohm.h (ctypes, no torch types cross the boundary).
The directory name carries hyphens, so import it through the repo-root shim:  `import fpohm_b200`.
"""
from .api import *  # noqa: F401,F403
from . import procedural  # noqa: F401
