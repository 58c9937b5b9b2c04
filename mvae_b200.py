"""Import shim: the package lives in the directory `multimodal-vae_b200/` (a name Python cannot
import directly because of the hyphen); `import mvae_b200` loads it under this name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodal-vae_b200")
_spec = importlib.util.spec_from_file_location(
    "mvae_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mvae_b200"] = _mod
_spec.loader.exec_module(_mod)
