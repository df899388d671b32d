#!/usr/bin/env python
"""Host-side cost of a step with the kernels stubbed out (runs WITHOUT a GPU).

The small-graph configs (BASELINE configs[1..3]) are host-bound on a B200: a step is a few hundred ~4 us kernels and
the Python between them takes longer than they do.  This script replaces every C-ABI call by a no-op (outputs stay
uninitialised), runs the unchanged Python layers (modules -> ops -> autograd -> optimizer) on CPU tensors and reports
the host time per step and its top contributors — the floor a step cannot beat however fast the kernels are, and the
thing to shrink.  Numbers are per host core of THIS machine; compare runs, not absolutes.

  python scripts/host_profile.py [--steps 200] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gnnb200  # noqa: F401,E402
from gnnb200 import data as data_mod, models, ops, synthetic  # noqa: E402


def stub_kernels():
    """Every entry point returns success without touching memory; workspaces are empty."""
    calls = {'n': 0}

    def invoke(name, *args):
        calls['n'] += 1
        return 0

    def call_ws(name, what, device, *args, stream, key=None):
        calls['n'] += 1
        return None

    ops._invoke = invoke
    ops._call_ws = call_ws
    ops.on_device = lambda t: True
    ops._stream = lambda t: 0
    return calls


def c2_step_factory():
    torch.manual_seed(0)
    dev = torch.device('cpu')
    ft = models.FinetuneGNN(dev, 'ENZYMES', 'full_finetune')
    ft.train()
    opt = torch.optim.AdamW(ft.param_groups)
    graphs = synthetic.tu_like_graphs('ENZYMES', 128, seed=42)
    batch = data_mod.Batch.from_data_list([data_mod.Data(**g) for g in graphs])

    def step():
        opt.zero_grad(set_to_none=True)
        logits = ft(batch)
        loss = logits.sum()                      # stand-in for cross_entropy: values are garbage under the stubs
        loss.backward()
        # no opt.step(): on CPU tensors AdamW does real arithmetic (4 ms for 1.3 M parameters) that a GPU run does not
        # spend on the host (there it is ~10 fused multi-tensor launches)
    return step


def c1_step_factory():
    torch.manual_seed(0)
    cora = synthetic.cora_like(42)
    m = torch.nn.ModuleDict({'input_encoder': models.InputEncoder(1433, 256), 'gnn_backbone': models.GINBackbone(3, 256)})
    m.train()
    x, ei = cora['x'], cora['edge_index']

    def step():
        m.zero_grad(set_to_none=True)
        m['gnn_backbone'](m['input_encoder'](x), ei.view_as(ei)).sum().backward()
    return step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--profile', action='store_true')
    args = ap.parse_args()
    torch.set_num_threads(1)
    calls = stub_kernels()
    for name, factory in (('c1 backbone fwd+bwd (L=3)', c1_step_factory), ('c2 fine-tune step', c2_step_factory)):
        step = factory()
        for _ in range(10):
            step()
        calls['n'] = 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
        per = calls['n'] / args.steps
        print(f'{name}: {dt * 1e3:.3f} ms host time per step, {per:.0f} C-ABI calls -> {dt * 1e6 / max(per, 1):.1f} us per call')
        if args.profile:
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(args.steps):
                step()
            pr.disable()
            pstats.Stats(pr).sort_stats('tottime').print_stats(18)


if __name__ == '__main__':
    main()
