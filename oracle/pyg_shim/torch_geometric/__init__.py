"""ORACLE (test infrastructure, never shipped): a minimal pure-PyTorch stand-in for the
`torch_geometric` package, exposing exactly the names the reference imports
(SURVEY.md §8c "exact shim surface").  Upstream PyG is absent -> parity UNPINNED."""
__version__ = '0.0.0-oracle-shim'
from . import data, nn, utils, loader, datasets, transforms  # noqa: F401
