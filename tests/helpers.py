"""Shared test helpers: deterministic weights that do not depend on module construction order,
and conversion of synthetic graphs into each side's Batch class."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def seeded_state_dict(model: torch.nn.Module, seed: int = 0) -> dict:
    """Fill every floating tensor of model.state_dict() from a generator keyed by (seed, key):
    weights/bias ~ N(0, 0.08), BN weight ~ 1 + N(0, 0.1), running_var in [0.5, 1.5]."""
    out = {}
    for key, ref in model.state_dict().items():
        if not ref.is_floating_point():
            out[key] = ref.clone()
            continue
        h = int.from_bytes(hashlib.sha256(f'{seed}:{key}'.encode()).digest()[:4], 'little')
        g = torch.Generator().manual_seed(h)
        t = torch.randn(ref.shape, generator=g)
        if key.endswith('running_var'):
            t = 0.5 + torch.rand(ref.shape, generator=g)
        elif key.endswith('running_mean'):
            t = 0.1 * t
        elif 'batch_norm' in key and key.endswith('weight') or key.endswith('nn.1.weight'):
            t = 1.0 + 0.1 * t
        elif key.endswith('eps'):
            t = 0.1 * t
        else:
            t = 0.08 * t
        out[key] = t.to(ref.dtype)
    return out


def oracle_batch(graphs):
    from oracle import install_pyg_shim
    install_pyg_shim()
    from torch_geometric.data import Batch, Data
    return Batch.from_data_list([Data(**{k: v.clone() for k, v in g.items()}) for g in graphs])


def product_batch(graphs, device):
    from gnnb200.data import Batch, Data
    b = Batch.from_data_list([Data(**{k: v.clone() for k, v in g.items()}) for g in graphs])
    return b.to(device)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)
