"""Import shim: the product package lives in `feature-preserving-octree-hex-meshing_b200/` (hyphens are not
importable), so `import fpohm_b200` loads that directory as the package `fpohm_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "feature-preserving-octree-hex-meshing_b200"
_spec = importlib.util.spec_from_file_location("fpohm_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fpohm_b200"] = _mod
_spec.loader.exec_module(_mod)
