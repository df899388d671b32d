"""gnnb200 — B200-native (sm_100a) message-passing hot path of alonbebchuk/GNN-Pretraining.

Import name ``gnnb200`` (see /gnnb200.py); the sources live in ``gnn-pretraining_b200/``:
  csrc/      hand-written CUDA kernels + the C ABI of libgnnb200.so (include/gnnb200.h)
  _lib.py    ctypes binding (fails loudly when the library is missing; no CPU fallback)
  ops.py     torch.library custom ops ``gnnb200::*`` with autograd formulas
  graph.py   cached sorted-CSR view of an edge_index tensor
  nn.py      GINConv / global_{mean,max,add}_pool / Linear drop-ins
  models.py  InputEncoder, GINLayer, GINBackbone, heads, PretrainableGNN, FinetuneGNN
  tasks.py   the six pre-training tasks (compute_loss call surface of the reference)
  partition.py / cuda_graphs.py   node-partitioned multi-GPU execution; CUDA-graph replay of a fixed-batch eval forward
  compat/    a ``torch_geometric`` stand-in so the reference's files run unmodified on these kernels
"""
from . import _lib
from ._lib import Gnnb200Error, LIB_PATH

__all__ = ['Gnnb200Error', 'LIB_PATH', 'library_available']


def library_available() -> bool:
    import os
    return os.path.isfile(LIB_PATH)
