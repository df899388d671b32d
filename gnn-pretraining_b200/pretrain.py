"""The caller of the hot path: one iteration of the reference's pre-training loop (`run_training`,
reference src/pretrain/pretrain.py:112-184), its evaluation pass (`run_evaluation`, :187-281), the adaptive loss
balancer (src/pretrain/adaptive_loss_balancer.py:10-60) and the task-specific AdamW groups
(src/pretrain/optimizers.py:18-81) — SURVEY.md §8f "next" #1 (the step around backward: surgery, clip, AdamW).

Same arithmetic, same order, same metric keys and values as the reference; what changes is how often the host waits
for the device.  The reference reads one scalar at a time (`float(t.detach().cpu())` per (domain, task) loss, per
task loss, per domain total, `.item()` per parameter tensor for the gradient norm: ~130 blocking reads per s5 step).
Here every scalar a step logs is appended to a `_ScalarLog` and fetched with ONE stacked device->host copy after the
optimizer step has been queued, so the metrics never stall the kernels of the step itself.

Everything in this file is device-agnostic host logic over the module interfaces (`compute_loss(domain_batches,
generator)`, `apply_gradient_surgery(model, losses, names)`, schedulers); `tests/test_pretrain_step.py` pins it bit
for bit against the UNMODIFIED reference `run_training` / `run_evaluation` on CPU (oracle modules over the PyG shim).
"""
import math
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch
from torch import Tensor

# src/pretrain/pretrain.py:28-55
BATCH_SIZE = 32
EPOCHS = 50
MAX_GRAD_NORM = 0.5
PATIENCE_FRACTION = 0.5
_TU = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']
PRETRAIN_DOMAINS = {'b2': _TU, 'b3': _TU, 'b4': ['ENZYMES'], 's1': _TU, 's2': _TU, 's3': _TU, 's4': _TU, 's5': _TU}
_S4 = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop']
ACTIVE_TASKS = {'b2': ['node_feat_mask'], 'b3': ['node_contrast'], 'b4': _S4, 's1': _S4[:2], 's2': _S4[2:4],
                's3': _S4[:4], 's4': _S4, 's5': _S4 + ['domain_adv']}

# src/pretrain/adaptive_loss_balancer.py:4-6
EPSILON = 1e-8
MIN_TOTAL_LOSS = 1e-6
WARMUP_STEPS = 100

# src/pretrain/optimizers.py:5-15
DEFAULT_LR = 1e-5
DEFAULT_WEIGHT_DECAY = 1e-5
TASK_SPECIFIC_LR = {'link_pred': 5e-7, 'node_feat_mask': 1e-5, 'node_contrast': 1e-5, 'graph_contrast': 1e-5,
                    'graph_prop': 1e-5, 'domain_adv': 5e-6}


class _ScalarLog:
    """Device scalars to report, fetched together: `add` returns a slot, `fetch` does the single device->host copy."""

    def __init__(self):
        self._tensors: List[Tensor] = []
        self._values: Optional[List[float]] = None

    def add(self, t: Tensor) -> int:
        self._tensors.append(t.detach().reshape(()).to(torch.float32))
        return len(self._tensors) - 1

    def fetch(self) -> None:
        if self._tensors:
            by_device: Dict[torch.device, List[int]] = {}
            for i, t in enumerate(self._tensors):
                by_device.setdefault(t.device, []).append(i)
            values = [0.0] * len(self._tensors)
            for idx in by_device.values():                       # one copy per device (normally exactly one)
                got = torch.stack([self._tensors[i] for i in idx]).cpu().tolist()
                for i, v in zip(idx, got):
                    values[i] = v
            self._values = values
        else:
            self._values = []

    def __getitem__(self, slot: int) -> float:
        return self._values[slot]


class AdaptiveLossBalancer:
    """Inverse-magnitude task weights after a warm-up of equal weights (adaptive_loss_balancer.py:10-60).  With one
    task the loss passes through untouched and the step counter does not move (:15-16).  After the warm-up the
    weights need the loss VALUES on the host: one stacked read instead of one per task."""

    def __init__(self):
        self.step_count = 0
        self.current_weights: Dict[str, float] = {}

    def balance_losses(self, task_losses: Dict[str, Tensor], domain_adv_lambda: float) -> Tensor:
        if len(task_losses) == 1:
            return next(iter(task_losses.values()))
        self.step_count += 1
        losses = dict(task_losses)
        if 'domain_adv' in losses:                               # unreachable from run_training (App. C.6); kept for parity
            others = sum(v for k, v in losses.items() if k != 'domain_adv')
            losses['domain_adv'] = torch.clamp(-domain_adv_lambda * losses['domain_adv'], min=-max(others * 0.5, 1.0))
        names = list(losses)
        if self.step_count > WARMUP_STEPS:
            values = torch.stack([losses[n].detach().reshape(()) for n in names]).cpu().tolist()
            magnitude = sum(abs(v) for v in values)
            raw = {n: (1.0 / (abs(v) + EPSILON)) if magnitude > 0 else 1.0 for n, v in zip(names, values)}
            norm = sum(raw.values())
            weights = {n: w / norm for n, w in raw.items()}
        else:
            weights = {n: 1.0 / len(names) for n in names}
        self.current_weights = dict(weights)
        return torch.clamp(torch.stack([weights[n] * losses[n] for n in names]).sum(), min=MIN_TOTAL_LOSS)

    def get_current_weights(self) -> Dict[str, float]:
        return self.current_weights


class TaskSpecificOptimizer:
    """AdamW with one parameter group per active task head (`heads.<task>` in the parameter name) at that task's
    learning rate and a 'default' group for everything else (optimizers.py:18-81).  Same group order, so optimizer
    state dicts are interchangeable with the reference's."""

    def __init__(self, model: torch.nn.Module, active_tasks: List[str]):
        self.model = model
        named = list(model.named_parameters())
        taken = set()
        groups = []
        for task in active_tasks:
            members = [(n, p) for n, p in named if f'heads.{task}' in n]
            taken.update(n for n, _ in members)
            if members:
                groups.append({'params': [p for _, p in members], 'lr': TASK_SPECIFIC_LR[task],
                               'weight_decay': DEFAULT_WEIGHT_DECAY, 'name': task})
        rest = [p for n, p in named if n not in taken]
        if rest:
            groups.append({'params': rest, 'lr': DEFAULT_LR, 'weight_decay': DEFAULT_WEIGHT_DECAY, 'name': 'default'})
        self.param_groups = groups
        self.optimizer = torch.optim.AdamW(self.param_groups)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def step(self) -> None:
        self.optimizer.step()


def _grad_norms(model: torch.nn.Module) -> List[Tensor]:
    grads = [p.grad for p in model.parameters() if p.grad is not None and p.requires_grad]
    if not grads:
        return []
    if grads[0].is_cuda:
        return list(torch._foreach_norm(grads, 2))              # one multi-tensor launch instead of one per tensor
    return [g.norm(2) for g in grads]


def train_step(model, tasks: Dict[str, object], optimizer, domain_batches: Dict[str, object], generator: torch.Generator,
               grl_sched, temperature_scheduler, loss_balancer: AdaptiveLossBalancer, gradient_surgery,
               pretrain_domains: Iterable[str], epoch: int = 0, device: Optional[torch.device] = None,
               allreduce: Optional[Callable[[torch.nn.Module], None]] = None) -> Dict[str, float]:
    """One iteration of `run_training`'s loop body (pretrain.py:112-184): task losses -> loss balancer -> gradient
    surgery over the main tasks (or a plain backward of the balanced loss when surgery does not apply) -> the
    domain-adversarial backward on top -> clip -> optimizer -> schedulers.  Returns the dict the reference hands to
    `wandb.log` (same keys, same values, same insertion order).

    `allreduce(model)` (optional, not in the reference) runs between the backward passes and the clipping: the
    data-parallel hook (`gnnb200.partition.allreduce_gradients` + division by the world size)."""
    if device is not None:
        for name in domain_batches:
            domain_batches[name] = domain_batches[name].to(device)
    log = _ScalarLog()
    domains = list(pretrain_domains)
    per_task: Dict[str, Tensor] = {}
    per_domain_task: Dict[str, Dict[str, Tensor]] = {d: {} for d in domains}
    for task_name, task in tasks.items():
        loss, by_domain = task.compute_loss(domain_batches, generator)
        per_task[task_name] = loss
        for d, part in by_domain.items():
            per_domain_task[d][task_name] = part
    slot_domain_task = {d: {t: log.add(v) for t, v in per_domain_task[d].items()} for d in domains}
    slot_domain = {d: log.add(torch.stack(list(per_domain_task[d].values())).sum()) for d in domains}

    lambda_val = grl_sched()
    main = {k: v for k, v in per_task.items() if k != 'domain_adv'}
    total_loss = loss_balancer.balance_losses(main, lambda_val)

    optimizer.zero_grad(set_to_none=True)
    surgery_metrics = gradient_surgery.apply_gradient_surgery(model, main, list(main))
    if not surgery_metrics:
        # one main task: surgery is a no-op and the balanced loss (= that task's loss) is back-propagated; the task
        # graphs share only leaves, so retain_graph (pretrain.py:147) has nothing to retain
        total_loss.backward()
    if 'domain_adv' in per_task:
        per_task['domain_adv'].backward()
    if allreduce is not None:
        allreduce(model)
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=MAX_GRAD_NORM)
    optimizer.step()
    grl_sched.step()
    temperature_scheduler.step()

    slot_task = {t: log.add(v) for t, v in per_task.items()}
    slot_total = log.add(total_loss)
    norm_slots = [log.add(n) for n in _grad_norms(model)]      # the clipped gradients, as in the reference (:174-180)
    log.fetch()                                                 # the step's only blocking read besides the balancer's

    m: Dict[str, float] = {}
    for d in domains:
        for t, s in slot_domain_task[d].items():
            m[f'train/loss/{d}/{t}'] = log[s]
    for t, s in slot_task.items():
        m[f'train/loss/{t}'] = log[s]
    for d in domains:
        m[f'train/loss/{d}'] = log[slot_domain[d]]
    m['train/loss/total'] = log[slot_total]
    m['train/progress/epoch'] = epoch
    if 'domain_adv' in per_task:
        m['train/domain_adv/lambda'] = grl_sched()             # after grl_sched.step(), like the reference (:165)
        m['train/domain_adv/loss'] = log[slot_task['domain_adv']]
    for t, w in loss_balancer.get_current_weights().items():
        m[f'train/loss_balancer/weight/{t}'] = w
    m.update(surgery_metrics)
    total = 0.0
    for s in norm_slots:
        total += log[s] ** 2
    m['train/gradients/model_grad_norm'] = total ** (1. / 2)
    return m


@torch.no_grad()
def evaluate(model, tasks: Dict[str, object], val_loaders: Dict[str, Iterable], generator: torch.Generator, grl_sched,
             loss_balancer: AdaptiveLossBalancer, device: Optional[torch.device] = None) -> Tuple[Tensor, Dict[str, float]]:
    """`run_evaluation` without its I/O (pretrain.py:202-258): per task the mean over domains of the mean over that
    domain's validation batches, balanced into the total that drives checkpointing / early stopping.  Returns
    (total_loss tensor, val_metrics); the caller compares `total_loss < best`, and `checkpoint_dict` gives the
    reference's on-disk format."""
    model.eval()
    log = _ScalarLog()
    per_task: Dict[str, Tensor] = {}
    slot_domain_task: Dict[str, Dict[str, int]] = {d: {} for d in val_loaders}
    for task_name, task in tasks.items():
        domain_means = []
        for d, loader in val_loaders.items():
            batch_losses = []
            for batch in loader:
                if device is not None:
                    batch = batch.to(device)
                batch_losses.append(task.compute_loss({d: batch}, generator)[0])
            mean = torch.stack(batch_losses).mean()
            domain_means.append(mean)
            slot_domain_task[d][task_name] = log.add(mean)
        per_task[task_name] = torch.stack(domain_means).mean()
    lambda_val = grl_sched()
    main = {k: v for k, v in per_task.items() if k != 'domain_adv'}
    total_loss = loss_balancer.balance_losses(main, lambda_val)
    slot_task = {t: log.add(v) for t, v in per_task.items()}
    slot_total = log.add(total_loss)
    log.fetch()
    m: Dict[str, float] = {}
    for d in slot_domain_task:
        for t, s in slot_domain_task[d].items():
            m[f'val/loss/{d}/{t}'] = log[s]
    for t, s in slot_task.items():
        m[f'val/loss/{t}'] = log[s]
    for d in slot_domain_task:                                   # host mean of the already-fetched floats (:236-239)
        vals = [log[s] for s in slot_domain_task[d].values()]
        m[f'val/loss/{d}'] = float(sum(vals) / len(vals))
    m['val/loss/total'] = log[slot_total]
    if 'domain_adv' in per_task:
        m['val/domain_adv/loss'] = log[slot_task['domain_adv']]
    return total_loss, m


def checkpoint_dict(epoch: int, model: torch.nn.Module, val_metrics: Dict[str, float]) -> Dict[str, object]:
    """The reference's checkpoint layout (pretrain.py:262-266; read back by finetune_model.py:128-146)."""
    return {'epoch': epoch, 'model_state_dict': model.state_dict(), 'val_metrics': val_metrics}


def patience(epochs: int = EPOCHS) -> int:
    """Early-stopping patience in epochs (pretrain.py:31 and the training driver)."""
    return int(math.floor(epochs * PATIENCE_FRACTION))
