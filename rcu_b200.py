"""Import shim: the package directory is `reliability-challenges-uncertainty_b200/` (a name Python cannot import
directly); this module registers it in sys.modules as `rcu_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reliability-challenges-uncertainty_b200')
_spec = importlib.util.spec_from_file_location('rcu_b200', os.path.join(_DIR, '__init__.py'), submodule_search_locations=[_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules['rcu_b200'] = _module
_spec.loader.exec_module(_module)
