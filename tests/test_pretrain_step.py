"""gnnb200.pretrain (the caller of the hot path: one `run_training` iteration, `run_evaluation`, the loss balancer and
the task-specific optimizer) pinned on CPU against the UNMODIFIED reference `src/pretrain/pretrain.py` imported over
the torch_geometric shim: identical metric dicts (keys, order, values) and bit-identical parameters after several
steps.  The step logic is device-agnostic host code, so both sides run the same CPU model/task classes (the reference's
own); the CUDA kernels behind the product's models are covered by tests/test_gpu_*.py.  Container only."""
import random

import pytest
import torch

from helpers import oracle_batch
from oracle.reference_loader import load_reference, reference_available

import gnnb200  # noqa: F401
from gnnb200 import pretrain as prod
from gnnb200 import synthetic

pytestmark = pytest.mark.skipif(not reference_available(), reason='/root/reference not present')

DOMAINS = ['MUTAG', 'ENZYMES']


@pytest.fixture(scope='module')
def ref():
    ns = load_reference()
    import src.pretrain.pretrain as ref_pretrain
    import src.pretrain.adaptive_loss_balancer as ref_balancer
    import src.pretrain.optimizers as ref_optim
    ns.pretrain, ns.balancer, ns.optim = ref_pretrain, ref_balancer, ref_optim
    return ns


def _batches(seed):
    return {d: oracle_batch(synthetic.tu_like_graphs(d, 4, seed=seed + i)) for i, d in enumerate(DOMAINS)}


class _Cfg:
    pretrain_domains = DOMAINS
    exp_name, seed = 'test', 0


def _world(ref, task_names, grl_start_step=0):
    """(model, tasks, optimizer, schedulers, balancer, surgery) built from the reference's own classes."""
    torch.manual_seed(0)
    model = ref.pretrain_model.PretrainableGNN(torch.device('cpu'), DOMAINS, task_names)
    grl = ref.schedulers.GRLScheduler(total_epochs=2, steps_per_epoch=5)
    grl.current_step = grl_start_step
    temp = ref.schedulers.TemperatureScheduler(total_steps=10)
    tasks = ref.pretrain.instantiate_tasks(model, task_names, grl, temp)
    return model, tasks, grl, temp


@pytest.mark.parametrize('task_names', [
    ['node_feat_mask'],                                                         # one task: balanced loss is back-propagated
    ['node_feat_mask', 'graph_prop', 'domain_adv'],                             # surgery over two tasks + adversarial backward
    ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv'],     # scheme s5
])
def test_train_step_equals_reference_run_training(ref, monkeypatch, task_names):
    steps = 3
    logged = []
    # the reference's own LP-decoder backward (index_put with accumulate) is run-to-run non-deterministic on a
    # multi-threaded CPU: one thread makes "the reference" a single well-defined answer
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        _compare_training(ref, monkeypatch, task_names, steps, logged)
    finally:
        torch.set_num_threads(threads)


def _compare_training(ref, monkeypatch, task_names, steps, logged):
    monkeypatch.setattr(ref.pretrain.wandb, 'log', lambda metrics, step=None: logged.append((step, dict(metrics))))
    # ---- reference: run_training over a 3-batch "loader" ----
    model_a, tasks_a, grl_a, temp_a = _world(ref, task_names, grl_start_step=5)
    opt_a = ref.optim.TaskSpecificOptimizer(model_a, task_names)
    bal_a = ref.balancer.AdaptiveLossBalancer()
    bal_a.step_count = 99                                                       # cross the warm-up boundary inside the test
    gen_a = torch.Generator().manual_seed(11)
    random.seed(5)
    torch.manual_seed(21)
    ref.pretrain.run_training(model_a, tasks_a, opt_a, [_batches(100 + s) for s in range(steps)], gen_a, grl_a, temp_a,
                              torch.device('cpu'), 7, [40], _Cfg(), bal_a, ref.gradient_surgery.GradientSurgery(torch.device('cpu')))
    # ---- product step logic over the same classes ----
    model_b, tasks_b, grl_b, temp_b = _world(ref, task_names, grl_start_step=5)
    opt_b = prod.TaskSpecificOptimizer(model_b, task_names)
    bal_b = prod.AdaptiveLossBalancer()
    bal_b.step_count = 99
    gen_b = torch.Generator().manual_seed(11)
    random.seed(5)
    torch.manual_seed(21)
    surgery = ref.gradient_surgery.GradientSurgery(torch.device('cpu'))
    mine = [prod.train_step(model_b, tasks_b, opt_b, _batches(100 + s), gen_b, grl_b, temp_b, bal_b, surgery, DOMAINS,
                            epoch=7, device=torch.device('cpu')) for s in range(steps)]
    assert [s for s, _ in logged] == [41, 42, 43]
    for (_, want), got in zip(logged, mine):
        assert list(got) == list(want)                                          # same keys in the same order
        for k in want:
            assert got[k] == want[k], (k, got[k], want[k])                      # bit-identical floats
    for (k, p), (_, q) in zip(model_a.state_dict().items(), model_b.state_dict().items()):
        assert torch.equal(p, q), k
    assert (grl_a.current_step, temp_a.current_step, bal_a.step_count) == (grl_b.current_step, temp_b.current_step, bal_b.step_count)
    assert bal_a.get_current_weights() == bal_b.get_current_weights()


def test_optimizer_groups_match_reference(ref):
    names = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv']
    model, *_ = _world(ref, names)
    a, b = ref.optim.TaskSpecificOptimizer(model, names), prod.TaskSpecificOptimizer(model, names)
    assert len(a.param_groups) == len(b.param_groups) == 7
    for ga, gb in zip(a.param_groups, b.param_groups):
        assert (ga['name'], ga['lr'], ga['weight_decay']) == (gb['name'], gb['lr'], gb['weight_decay'])
        assert [id(p) for p in ga['params']] == [id(p) for p in gb['params']]
    assert list(a.optimizer.state_dict()['param_groups'][0]) == list(b.optimizer.state_dict()['param_groups'][0])


@pytest.mark.parametrize('step_count', [0, 100, 250])
def test_loss_balancer_equals_reference(ref, step_count):
    g = torch.Generator().manual_seed(step_count)
    losses = {n: (torch.rand((), generator=g) * 10 ** i).requires_grad_(True) for i, n in enumerate(['a', 'b', 'c', 'domain_adv'])}
    a, b = ref.balancer.AdaptiveLossBalancer(), prod.AdaptiveLossBalancer()
    a.step_count = b.step_count = step_count
    for subset in (['a'], ['a', 'b', 'c'], ['a', 'b', 'c', 'domain_adv']):
        la = a.balance_losses({k: losses[k] for k in subset}, 0.3)
        lb = b.balance_losses({k: losses[k] for k in subset}, 0.3)
        assert torch.equal(la, lb) and a.step_count == b.step_count
        assert a.get_current_weights() == b.get_current_weights()
        if len(subset) > 1:
            ga = torch.autograd.grad(la, [losses[k] for k in subset])
            gb = torch.autograd.grad(lb, [losses[k] for k in subset])
            assert all(torch.equal(x, y) for x, y in zip(ga, gb))


def test_evaluate_equals_reference_run_evaluation(ref, monkeypatch, tmp_path):
    names = ['node_feat_mask', 'graph_prop', 'domain_adv']
    logged, saved = [], []
    monkeypatch.setattr(ref.pretrain.wandb, 'log', lambda metrics, step=None: logged.append(dict(metrics)))
    monkeypatch.setattr(ref.pretrain.wandb, 'log_artifact', lambda *a, **k: None)
    monkeypatch.setattr(ref.pretrain.wandb, 'Artifact', lambda **k: type('A', (), {'add_file': lambda self, p: None})())
    monkeypatch.setattr(ref.pretrain, 'OUTPUT_DIR', tmp_path)
    monkeypatch.setattr(ref.pretrain.torch, 'save', lambda obj, path: saved.append(obj))
    model, tasks, grl, temp = _world(ref, names, grl_start_step=6)
    loaders = {d: [oracle_batch(synthetic.tu_like_graphs(d, 3, seed=60 + 10 * i + j)) for j in range(2)] for i, d in enumerate(DOMAINS)}
    bal_a, bal_b = ref.balancer.AdaptiveLossBalancer(), prod.AdaptiveLossBalancer()
    random.seed(1)
    best, since = ref.pretrain.run_evaluation(model, tasks, loaders, torch.Generator().manual_seed(3), grl, torch.device('cpu'),
                                              2, float('inf'), 4, _Cfg(), [9], bal_a)
    random.seed(1)
    total, metrics = prod.evaluate(model, tasks, loaders, torch.Generator().manual_seed(3), grl, bal_b, torch.device('cpu'))
    assert since == 0 and torch.equal(best, total)
    assert list(metrics) == list(logged[0]) and all(metrics[k] == logged[0][k] for k in metrics)
    ck = prod.checkpoint_dict(2, model, metrics)
    assert list(ck) == list(saved[0]) == ['epoch', 'model_state_dict', 'val_metrics']
    assert ck['epoch'] == saved[0]['epoch'] and ck['val_metrics'] == saved[0]['val_metrics']
    assert list(ck['model_state_dict']) == list(saved[0]['model_state_dict'])
    assert prod.patience() == int(ref.pretrain.EPOCHS * ref.pretrain.PATIENCE_FRACTION)
    assert prod.PRETRAIN_DOMAINS == ref.pretrain.PRETRAIN_DOMAINS and prod.ACTIVE_TASKS == ref.pretrain.ACTIVE_TASKS
    assert prod.TASK_SPECIFIC_LR == ref.optim.TASK_SPECIFIC_LR and prod.MAX_GRAD_NORM == ref.pretrain.MAX_GRAD_NORM
