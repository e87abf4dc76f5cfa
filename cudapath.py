"""Import shim: the package directory is named `cs184-final-project-mitsuba0.5_b200` (not a valid Python identifier),
so it is loaded by path and exposed as the module `cudapath` (sub-module `cudapath.scenes`)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cs184-final-project-mitsuba0.5_b200')
_NAME = 'cudapath_b200'

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, '__init__.py'), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
import importlib as _il
scenes = _il.import_module(_NAME + '.scenes')
dist = _il.import_module(_NAME + '.dist')
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith('__')})
