"""The C-ABI library loads (no GPU needed) and exports every symbol include/gnnb200.h declares;
the ctypes signature table covers exactly the same set."""
import ctypes
import os
import re

import gnnb200
from gnnb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'gnnb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gnnb200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    assert gnnb200.library_available(), 'build libgnnb200.so first (__graft_entry__.build())'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/gnnb200.h but not exported'


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.gnnb200_version() >= 100
    assert lib.gnnb200_error_string(0) == b'ok'
    assert b'workspace' in lib.gnnb200_error_string(-3)


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from gnnb200 import ops
    with pytest.raises(Exception):
        ops.csr_build(torch.zeros(2, 3, dtype=torch.long), 4, False)
