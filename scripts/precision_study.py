"""Which GEMMs decide the end-to-end gradient error?  CPU emulation of kind::tf32 operand truncation around every
Linear of the ORACLE model (forward / dX / dW separately), compared with an fp64 run of the same model.
Usage: python scripts/precision_study.py   (no GPU needed; test infrastructure, imports oracle/)"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import install_pyg_shim  # noqa: E402
from oracle import modules as orc  # noqa: E402
install_pyg_shim()
from torch_geometric.data import Batch, Data  # noqa: E402
import gnnb200  # noqa: E402,F401
from gnnb200 import synthetic  # noqa: E402


def tf32(x):
    if x.dtype != torch.float32:
        return x
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


MODE = {'fwd': False, 'dx': False, 'dw': False}


class EmuLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        xx, ww = (tf32(x), tf32(w)) if MODE['fwd'] else (x, w)
        y = xx @ ww.t()
        return y if b is None else y + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = (tf32(g) @ tf32(w)) if MODE['dx'] else g @ w
        g2, x2 = g.reshape(-1, g.size(-1)), x.reshape(-1, x.size(-1))
        gw = (tf32(g2).t() @ tf32(x2)) if MODE['dw'] else g2.t() @ x2
        return gx, gw, g2.sum(0)


_orig = F.linear


def patched(x, w, b=None):
    return EmuLinear.apply(x, w, b)


def run(model, batch, dtype, train):
    m = model.to(dtype)
    m.train(train)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    b = batch.clone()
    x = b.x.to(dtype).clone().requires_grad_(True)
    b.x = x
    out = m(b)
    out.sum().backward()
    grads = {k: p.grad.double().clone() for k, p in m.named_parameters() if p.grad is not None}
    for p in m.parameters():
        p.grad = None
    return out.detach().double(), x.grad.double(), grads


def fro(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-300))


def main():
    import copy
    ngraphs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    graphs = synthetic.tu_like_graphs('ENZYMES', ngraphs, seed=0)
    batch = Batch.from_data_list([Data(**g) for g in graphs])
    orc.DROPOUT_RATE = 0.0
    for seed in range(3):
        torch.manual_seed(seed)
        base = orc.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
        state = base.state_dict()

        def fresh():
            m = orc.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
            m.load_state_dict(state)
            return m
        for train in (False, True):
            ref = run(fresh(), batch, torch.float64, train)
            rows = []
            torch.nn.functional.linear = _orig
            got = run(fresh(), batch, torch.float32, train)
            rows.append(('fp32 oracle', got))
            torch.nn.functional.linear = patched
            for name, cfg in (('tf32 all', (1, 1, 1)), ('tf32 fwd only', (1, 0, 0)), ('tf32 bwd only (dx+dw)', (0, 1, 1)),
                              ('tf32 dw only', (0, 0, 1)), ('tf32 dx only', (0, 1, 0))):
                MODE['fwd'], MODE['dx'], MODE['dw'] = [bool(c) for c in cfg]
                rows.append((name, run(fresh(), batch, torch.float32, train)))
            torch.nn.functional.linear = _orig
            print(f'seed {seed} train={train}')
            for name, (out, gx, grads) in rows:
                gw = max(fro(grads[k], ref[2][k]) for k in grads if k.endswith('weight') and ref[2][k].norm() > 1e-12)
                print(f'  {name:24s} logits {fro(out, ref[0]):.2e}  dX fro {fro(gx, ref[1]):.2e}  worst dW fro {gw:.2e}')


if __name__ == '__main__':
    main()
