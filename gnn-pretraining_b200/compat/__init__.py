"""``torch_geometric`` stand-in backed by the gnnb200 kernels.

The reference imports exactly these PyG names: nn.{GINConv, global_mean_pool,
global_max_pool}, utils.{to_undirected, batched_negative_sampling, negative_sampling, subgraph,
remove_self_loops}, data.{Data, Batch}, loader.DataLoader, plus import-time-only
datasets.{Planetoid, TUDataset} and transforms.NormalizeFeatures.  ``install()`` puts this package
on ``sys.path`` under the name ``torch_geometric`` so that the reference's own files
(src/models/*.py, src/pretrain/*.py, run_pretrain.py, run_finetune.py) run unmodified with their
message passing, pooling and coalesce on the B200 kernels.  See INTEGRATION.md.
"""
import os
import sys


def install() -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    existing = sys.modules.get('torch_geometric')
    if existing is not None and getattr(existing, '__gnnb200__', False):
        return
    if existing is not None:
        raise RuntimeError('a different torch_geometric is already imported')
    sys.path.insert(0, here)
    import torch_geometric  # noqa: F401
