"""Import alias for the package directory ./flair-1_b200/ (a hyphen is not importable).

`import flair1_b200` executes flair-1_b200/__init__.py under this name and registers the directory as
the package search path, so `import flair1_b200.zone_detect.main` etc. work as usual.
"""
import importlib.util as _ilu
import pathlib as _pl
import sys as _sys

_dir = _pl.Path(__file__).resolve().parent / "flair-1_b200"
_spec = _ilu.spec_from_file_location(__name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
