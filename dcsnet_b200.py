"""Import shim: the package directory is `dcs-net_b200/` (not a valid Python identifier), so `import dcsnet_b200`
loads it from there.  All product code lives in `dcs-net_b200/`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dcs-net_b200")
_spec = _u.spec_from_file_location("dcsnet_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["dcsnet_b200"] = _mod
_spec.loader.exec_module(_mod)
