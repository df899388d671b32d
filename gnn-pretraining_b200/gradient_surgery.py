"""Gradient surgery with the reference's call surface (`GradientSurgery(device).apply_gradient_surgery(model,
task_losses, task_names) -> metrics`, reference src/pretrain/gradient_surgery.py:7-27, called from
src/pretrain/pretrain.py:145) — SURVEY.md §8f "next" #1.

Semantics kept from the reference (App. C.2): per-parameter-tensor projection, only against tasks EARLIER in
an (unseeded) `random.shuffle` order, against their ORIGINAL gradients; the averaged result is written only
for parameters present in the first shuffled task's gradient dict, every other parameter keeps whatever
`.grad` the last per-task backward left.  What changes: the T per-task gradients live in one flat [T, P]
buffer and the ~3 host syncs per (task pair, parameter tensor) become one kernel launch plus one read-back
of the conflict counters for the metrics dict.
"""
import random
from typing import Dict, List

import torch
import torch.nn as nn

from . import ops


class GradientSurgery:
    def __init__(self, device: torch.device):
        self.device = device

    def apply_gradient_surgery(self, model: nn.Module, task_losses: Dict[str, torch.Tensor],
                               task_names: List[str]) -> Dict[str, float]:
        if len(task_losses) <= 1:
            return {}
        named = list(model.named_parameters())
        sizes = [p.numel() for _, p in named]
        offsets = [0]
        for n in sizes:
            offsets.append(offsets[-1] + n)
        total = offsets[-1]
        names = list(task_losses.keys())
        T = len(names)
        flat = torch.zeros(T, total, dtype=torch.float32, device=self.device)
        present_host = []
        for t, name in enumerate(names):
            for _, p in named:                                         # = model.zero_grad(set_to_none=True) without its module walk
                p.grad = None
            task_losses[name].backward(retain_graph=True)
            row = flat[t]
            have = []
            pieces, slots = [], []
            for s, (_, p) in enumerate(named):
                if p.grad is not None:
                    have.append(1)
                    pieces.append(p.grad.reshape(-1))
                    slots.append(s)
                else:
                    have.append(0)
            present_host.append(have)
            # copy the present gradients into the flat row (contiguous runs are batched into single copies)
            i = 0
            while i < len(slots):
                j = i
                while j + 1 < len(slots) and slots[j + 1] == slots[j] + 1:
                    j += 1
                seg = torch.cat(pieces[i:j + 1]) if j > i else pieces[i]
                row[offsets[slots[i]]:offsets[slots[j] + 1]].copy_(seg)
                i = j + 1
        order_names = list(task_names)
        random.shuffle(order_names)                                    # gradient_surgery.py:42-43
        order = torch.tensor([names.index(n) for n in order_names], dtype=torch.int32, device=self.device)
        present = torch.tensor(present_host, dtype=torch.uint8, device=self.device)
        seg_off = torch.tensor(offsets, dtype=torch.int64, device=self.device)
        out, has_out, counters = ops.pcgrad(flat, seg_off, present, order)
        # parameters present in the first shuffled task get the averaged projection (views into `out`); the others
        # keep the .grad left by the last backward above (gradient_surgery.py:36-39,60-68)
        first = present_host[names.index(order_names[0])]
        for s, (_, p) in enumerate(named):
            if first[s]:
                p.grad = out[offsets[s]:offsets[s + 1]].view_as(p)
        c = counters.sum(dim=(0, 1)).tolist()                           # the one device->host read
        conflicts, projections = int(c[0]), int(c[1])
        return {'gradient_surgery/total_conflicts': conflicts,
                'gradient_surgery/total_projections': projections,
                'gradient_surgery/conflict_ratio': conflicts / max(projections, 1)}
