"""Import alias: ``import gnnb200`` loads the package that lives in ``gnn-pretraining_b200/``
(the directory name the build contract fixes is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gnn-pretraining_b200')
_spec = importlib.util.spec_from_file_location('gnnb200', os.path.join(_dir, '__init__.py'),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['gnnb200'] = _mod
_spec.loader.exec_module(_mod)
