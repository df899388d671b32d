"""CUDA-graph replay of an eval-mode forward on a FIXED batch (the validation pass of the fine-tuning / pre-training loops,
reference src/finetune/finetune.py:328-352 and src/pretrain/pretrain.py:202-258, runs the same batches after every epoch).

A small-graph forward is ~50 launches of ~4 us kernels: launch-bound.  Captured once, it replays as ONE launch.  What is inside
the graph reads the LIVE parameter and BatchNorm buffers, so replays follow the optimizer: the weight split of the compensated
GEMMs (ops.split_weight) and the re-pitched copies of ragged-width operands (ops._tma_rows) are recomputed inside the graph
instead of being served from their caches.  What is fixed at capture time: the batch's tensors (same storage; new feature
values may be copied INTO `x` before a replay) and its structure — the CSR / `ptr` vectors are built before the capture (their
construction reads counts back) and baked into the graph, so `edge_index` / `batch` must not change."""
from typing import Callable

import torch

from . import ops


class CapturedForward:
    """`fn()` -> Tensor (or tuple of tensors), captured after `warmup` eager calls.  `replay()` returns the same output tensors,
    refilled."""

    def __init__(self, fn: Callable[[], object], warmup: int = 2):
        if not torch.cuda.is_available():
            raise ops.L.Gnnb200Error('CUDA graphs need a CUDA device (no CPU fallback)')
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):                       # builds the structure caches, sizes the workspaces
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        ops.BYPASS_OPERAND_CACHES = True                          # the split / re-pitch kernels become part of the graph
        try:
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out = fn()
        finally:
            ops.BYPASS_OPERAND_CACHES = False

    def replay(self):
        self.graph.replay()
        return self.out


def capture_eval(model: torch.nn.Module, *args, warmup: int = 2, **kwargs) -> CapturedForward:
    """Capture `model(*args, **kwargs)` in eval mode (dropout off, BatchNorm on its running statistics)."""
    if model.training:
        raise ops.L.Gnnb200Error('capture_eval: put the model in eval() mode first (training steps draw dropout seeds and '
                                 'update statistics on the host)')
    return CapturedForward(lambda: model(*args, **kwargs), warmup)
