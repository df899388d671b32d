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


def c4_step_factory():
    """bench.py's C4 step (sampler draw + pretrain.train_step on four TU-shaped domains, scheme s5)."""
    import importlib.util
    import random
    from gnnb200 import utils
    spec = importlib.util.spec_from_file_location('bench_for_profile', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # losses must stay finite under the stubs; zero-filling every output would dominate the profile, so outputs of one
    # shape share one zero tensor (nothing writes to them here)
    real_empty, cache = torch.empty, {}

    def shared_zeros(*shape, **k):
        key = (shape if not (len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size))) else tuple(shape[0]),
               k.get('dtype'))
        t = cache.get(key)
        if t is None:
            t = cache[key] = real_empty(*shape, **k).zero_()
        return real_empty(0, dtype=t.dtype).set_(t.untyped_storage(), 0, t.shape)     # own version counter
    ops.torch.empty = shared_zeros

    def coalesce_on_host(edge_index, num_nodes):                    # its result width steers host control flow
        key = torch.unique(edge_index[0] * num_nodes + edge_index[1])
        out = torch.zeros_like(edge_index)
        out[0, :key.numel()], out[1, :key.numel()] = key // num_nodes, key % num_nodes
        return out, torch.tensor(key.numel())
    utils.ops.coalesce = coalesce_on_host
    step, sampler = bench.build_c4_step(torch.device('cpu'), rank=0, world=1)
    random.seed(1)
    return lambda: step(sampler.draw())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--only', default=None, choices=[None, 'c1', 'c2', 'c4'])
    args = ap.parse_args()
    torch.set_num_threads(1)
    calls = stub_kernels()
    for name, factory, div in (('c1 backbone fwd+bwd (L=3)', c1_step_factory, 1), ('c2 fine-tune step', c2_step_factory, 1),
                               ('c4 s5 pre-training step (4 domains x 32 graphs)', c4_step_factory, 10)):
        if args.only and not name.startswith(args.only):
            continue
        step = factory()
        steps = max(3, args.steps // div)
        for _ in range(max(2, 10 // div)):
            step()
        calls['n'] = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
        per = calls['n'] / steps
        print(f'{name}: {dt * 1e3:.3f} ms host time per step, {per:.0f} C-ABI calls -> {dt * 1e6 / max(per, 1):.1f} us per call')
        if args.profile:
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(steps):
                step()
            pr.disable()
            pstats.Stats(pr).sort_stats('tottime').print_stats(30)


if __name__ == '__main__':
    main()
