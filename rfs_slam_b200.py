"""Import shim: the package directory is named `rfs-slam_b200/` (not an identifier), so this
module loads it under the importable name `rfs_slam_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "rfs-slam_b200")
_spec = _u.spec_from_file_location("rfs_slam_b200", _os.path.join(_d, "__init__.py"),
                                   submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["rfs_slam_b200"] = _mod
_spec.loader.exec_module(_mod)
