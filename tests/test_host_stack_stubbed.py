"""The whole Python stack of the product (loader -> tasks -> models -> ops -> autograd -> gradient surgery ->
pretrain.train_step, i.e. bench.py's C4 step) executed on CPU with every C-ABI call replaced by a no-op that leaves its
(zero-filled) outputs untouched.  Numbers are meaningless; what this pins without a GPU is the host logic between the
kernels: argument plumbing, shapes, autograd wiring, the order and count of kernel calls, metric keys, scheduler state.
The kernels themselves are covered by tests/test_gpu_*.py."""
import importlib.util
import math
import os
import random

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import ops, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def stubbed(monkeypatch):
    calls = {}

    def invoke(name, *args):
        calls[name] = calls.get(name, 0) + 1
        return 0

    def call_ws(name, what, device, *args, stream, key=None):
        calls[name] = calls.get(name, 0) + 1
        return None

    real_empty = torch.empty

    def zeros_instead_of_empty(*a, **k):          # outputs the stubs never write must still be finite
        return real_empty(*a, **k).zero_()

    def coalesce_on_host(edge_index, num_nodes):  # its result width steers host control flow: answer it for real
        key = torch.unique(edge_index[0] * num_nodes + edge_index[1])
        out = torch.zeros_like(edge_index)
        out[0, :key.numel()], out[1, :key.numel()] = key // num_nodes, key % num_nodes
        return out, torch.tensor(key.numel())

    monkeypatch.setattr(ops, '_invoke', invoke)
    monkeypatch.setattr(ops, '_call_ws', call_ws)
    monkeypatch.setattr(ops, 'on_device', lambda t: True)
    monkeypatch.setattr(ops, '_stream', lambda t: 0)
    monkeypatch.setattr(ops.torch, 'empty', zeros_instead_of_empty)
    monkeypatch.setattr(utils.ops, 'coalesce', coalesce_on_host)
    return calls


def _bench_module():
    spec = importlib.util.spec_from_file_location('bench_for_test', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_c4_step_runs_through_the_whole_host_stack(stubbed):
    bench = _bench_module()
    step, sampler = bench.build_c4_step(torch.device('cpu'), rank=0, world=1)
    grl, temp, balancer = step.schedulers
    before = (grl.current_step, temp.current_step, balancer.step_count)
    random.seed(1)
    metrics = [step(sampler.draw()) for _ in range(2)]
    assert (grl.current_step, temp.current_step, balancer.step_count) == (before[0] + 2, before[1] + 2, before[2] + 2)
    want_keys = {f'train/loss/{d}/{t}' for d in bench.TU_DOMAINS for t in bench.S5_TASKS} | \
                {f'train/loss/{t}' for t in bench.S5_TASKS} | {f'train/loss/{d}' for d in bench.TU_DOMAINS} | \
                {'train/loss/total', 'train/progress/epoch', 'train/domain_adv/lambda', 'train/domain_adv/loss',
                 'train/gradients/model_grad_norm', 'gradient_surgery/total_conflicts', 'gradient_surgery/total_projections',
                 'gradient_surgery/conflict_ratio'} | {f'train/loss_balancer/weight/{t}' for t in bench.S5_TASKS[:-1]}
    for m in metrics:
        assert set(m) == want_keys
        assert all(math.isfinite(float(v)) for v in m.values())
    # every family of kernels was reached, in plausible proportions: per step 4 domains x (NFM 1 + LP 1 + NC 2 + GC 2 +
    # GP 1 + DA 1 = 8 backbone passes + the no-grad encoder pass of NFM) x 5 layers
    per_step = {k: v / 2 for k, v in stubbed.items()}
    # (a GIN layer pass is one composite call, csrc/gin_layer.cu; GNNB200_NATIVE_LAYER=0 makes it ~9 calls)
    layer_fwd = per_step.get('gnnb200_gin_layer_fwd_f32', 0) + per_step.get('gnnb200_aggregate_f32', 0)
    assert layer_fwd >= 4 * 8 * 5                                               # forward gathers alone
    assert per_step.get('gnnb200_gin_layer_bwd_f32', 0) + per_step.get('gnnb200_aggregate_f32', 0) >= 4 * 7 * 5
    for name in ('gnnb200_csr_build_i64', 'gnnb200_gemm_f32', 'gnnb200_bn_act_fwd_f32', 'gnnb200_bn_act_bwd_f32',
                 'gnnb200_segment_pool_fwd_f32', 'gnnb200_segment_pool_bwd_f32', 'gnnb200_rows_gather_f32',
                 'gnnb200_rows_scatter_f32', 'gnnb200_lp_features_f32', 'gnnb200_lp_features_bwd_f32',
                 'gnnb200_ntxent_fwd_f32', 'gnnb200_ntxent_bwd_f32', 'gnnb200_pcgrad_f32'):
        assert per_step.get(name, 0) > 0, name
    assert per_step['gnnb200_pcgrad_f32'] == 1                                  # one surgery kernel per step


def test_finetune_steps_run_through_the_host_stack(stubbed):
    from gnnb200 import finetune, models, synthetic
    from gnnb200.data import Batch, Data
    dev = torch.device('cpu')
    torch.manual_seed(0)
    # graph classification
    model = models.FinetuneGNN(dev, 'ENZYMES', 'full_finetune')
    model.train()
    opt = torch.optim.AdamW(model.param_groups)
    batch = Batch.from_data_list([Data(**g) for g in synthetic.tu_like_graphs('ENZYMES', 8, seed=3)])
    loss, targets, pred, prob = finetune.train_step(model, opt, batch, dev, 'graph_classification', 'ENZYMES')
    assert prob.shape == (8, 6) and pred.shape == targets.shape == (8,) and torch.isfinite(loss)
    # link prediction, training mode: miner (similarity GEMM stubbed -> all ties) + decoder
    d = synthetic.planetoid_like(60, 120, 1433, seed=4)
    lp = models.FinetuneGNN(dev, 'Cora_LP', 'full_finetune')
    lp.train()
    train_edges = d['edge_index'][:, ::2].contiguous()
    data = Data(x=d['x'], edge_index=d['edge_index'])
    loss, targets, pred, prob = finetune.process_batch(lp, (data, train_edges[:, :16], None), dev, 'link_prediction', 'Cora_LP',
                                                       finetune.LinkPredictionHardNegativeMiner(), train_edges)
    assert targets.tolist() == [1] * 16 + [0] * 16 and prob.shape == (32, 2)
    assert stubbed['gnnb200_normalize_rows_f32'] == 1 and stubbed['gnnb200_gemm_f32'] > 1
    # node classification
    nc = models.FinetuneGNN(dev, 'Cora_NC', 'linear_probe')
    nc.train()
    idx = torch.arange(0, 60, 2)
    y = torch.randint(0, 7, (60,))
    loss, targets, pred, prob = finetune.process_batch(nc, (Data(x=d['x'], edge_index=d['edge_index'], y=y), idx, y[idx]), dev,
                                                       'node_classification', 'Cora_NC', None, d['edge_index'])
    assert prob.shape == (30, 7)


# ---- the node-partitioned training step (bench.py's N > 1 workload) over gloo with the kernels stubbed ----------------
def _install_stubs():
    real_empty = torch.empty
    ops._invoke = lambda name, *a: 0
    ops._call_ws = lambda name, what, device, *a, stream, key=None: None
    ops.on_device = lambda t: True
    ops._stream = lambda t: 0
    ops.torch.empty = lambda *a, **k: real_empty(*a, **k).zero_()


def _partition_worker(rank, world, port, halo, out_dir):
    import torch.distributed as dist
    from gnnb200 import models, partition, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        _install_stubs()
        n, e = 301, 2500
        data = synthetic.products_like(n, e, 100, seed=3, locality=0.9, blocks=8)
        runner = partition.PartitionedBackboneStep(models, torch.device('cpu'), 100, 256, 2, n, rank, world, halo=halo)
        before = [p.detach().clone() for p in runner.model.parameters()]
        for _ in range(2):
            loss = runner.step(data['x'], data['edge_index'])
        assert runner.last_halo == {'auto': 'sparse_overlap'}.get(halo, halo)        # 'auto' picks the overlapped sparse exchange here
        assert loss.shape == () and torch.isfinite(loss)
        assert all(p.grad is not None for p in runner.model.parameters())
        assert len(before) == len(list(runner.model.parameters()))
        # every rank ends the step with identical parameters (same all-reduced gradients, same optimizer state)
        flat = torch.cat([p.detach().reshape(-1) for p in runner.model.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert all(torch.equal(both[0], b) for b in both[1:])
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('halo,world', [('dense', 2), ('sparse', 2), ('sparse_overlap', 2), ('auto', 3), ('peer', 2), ('peercopy', 2)])
def test_partitioned_step_host_logic_over_gloo(tmp_path, halo, world):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_partition_worker, args=(world, port, halo, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_bench_secondary_workloads_run_through_the_host_stack(stubbed):
    """bench.py's C1 / C2 / C3 steps (Cora-shaped backbone, ENZYMES fine-tune step, s4 multi-task step on Cora / CiteSeer /
    ENZYMES shaped domains incl. the `random.sample` negative-sampling branch and gradient surgery) with stubbed kernels."""
    bench = _bench_module()
    out = bench.small_graph_steps('gnnb200', torch.device('cpu'), steps=4, warmup=2)
    assert out['c2_shape']['graphs'] == 128 and out['c3_shape']['tasks'] == bench.S4_TASKS
    assert all(out[k] > 0 for k in ('c1_edges_per_s', 'c2_finetune_steps_per_s', 'c3_s4_pretrain_steps_per_s'))
    assert stubbed['gnnb200_pcgrad_f32'] >= 2 and stubbed['gnnb200_ntxent_fwd_f32'] + stubbed.get('gnnb200_ntxent_sim_fwd_f32', 0) > 0


def test_long_rows_plumbing(stubbed, monkeypatch):
    """GNNB200_LONG_ROWS: the CSR build lists the rows above the threshold (host check of the index arithmetic) and the
    aggregation issues the skip-flagged launch followed by the block-per-row launch over exactly those rows."""
    from gnnb200 import _lib as L, graph as graph_mod
    rowptr = torch.tensor([0, 3, 3 + 2000, 3 + 2000 + 1024, 3 + 2000 + 1024 + 1025], dtype=torch.int32)
    monkeypatch.setattr(graph_mod, 'LONG_ROWS', False)
    assert graph_mod.long_rows_of(rowptr, 10 ** 6) is None                            # GNNB200_LONG_ROWS=0
    monkeypatch.setattr(graph_mod, 'LONG_ROWS', True)
    assert graph_mod.long_rows_of(rowptr, 10) is None                                 # small graphs never pay the sync
    ids = graph_mod.long_rows_of(rowptr, 10 ** 6)
    assert ids.tolist() == [1, 3]                                                     # 1024 neighbours is not "long"
    seen = []
    monkeypatch.setattr(ops, '_invoke', lambda name, *a: seen.append((name, a)) or 0)
    x = torch.zeros(4, 256)
    col = torch.zeros(int(rowptr[-1]), dtype=torch.int32)
    ops._aggregate_raw(x, rowptr, col, L.AGG_SUM, x, torch.zeros(1), None, long_rows=ids)
    (n1, a1), (n2, a2) = seen
    assert n1 == 'gnnb200_aggregate_f32' and a1[6] == L.AGG_SUM | L.AGG_SKIP_LONG
    assert n2 == 'gnnb200_aggregate_long_rows_f32' and a2[4] == ids.data_ptr() and a2[5] == 2 and a2[7] == L.AGG_SUM
    assert a2[11] == a1[11]                                                           # same output buffer
    seen.clear()
    ops._aggregate_raw(x, rowptr, col, L.AGG_SUM, x, torch.zeros(1), None, out=torch.zeros(4, 256), long_rows=ids)
    assert seen[0][1][6] == L.AGG_SUM | L.AGG_ACCUMULATE | L.AGG_SKIP_LONG and seen[1][1][7] == L.AGG_SUM | L.AGG_ACCUMULATE
    seen.clear()
    ops._aggregate_raw(torch.zeros(4, 255), rowptr, col, L.AGG_SUM, None, None, None, long_rows=ids)    # no 128-bit layout
    assert [n for n, _ in seen] == ['gnnb200_aggregate_f32'] and seen[0][1][6] == L.AGG_SUM


def test_bench_c4_subprocess_leg_parses_or_reports(monkeypatch):
    """bench.c4_in_subprocess: the child's JSON line becomes the secondary entry; a child that dies is reported, never raised."""
    import json
    import subprocess
    import types
    bench = _bench_module()
    line = {'metric': 'pretrain_steps_per_sec', 'value': 12.5, 'ms_per_step': 80.0, 'gpu_launches': 30000,
            'config': {'graphs_per_sec': 1600.0}, 'e2e': {'value': 11.0}}
    monkeypatch.setattr(subprocess, 'run', lambda *a, **k: types.SimpleNamespace(stdout='noise\n' + json.dumps(line) + '\n', stderr='', returncode=0))
    got = bench.c4_in_subprocess(steps=10)
    assert got == {'steps_per_s': 12.5, 'ms_per_step': 80.0, 'graphs_per_s': 1600.0, 'e2e_steps_per_s': 11.0,
                   'gpu_launches_per_step': 3000.0}
    monkeypatch.setattr(subprocess, 'run', lambda *a, **k: types.SimpleNamespace(stdout='', stderr='Traceback\nRuntimeError: boom', returncode=1))
    assert bench.c4_in_subprocess()['error'] == 'RuntimeError: boom'

    def hang(*a, **k):
        raise subprocess.TimeoutExpired('bench', 1)
    monkeypatch.setattr(subprocess, 'run', hang)
    assert 'TimeoutExpired' in bench.c4_in_subprocess()['error']


def test_graph_and_segment_caches(stubbed):
    """graph_of / segment_ptr_of build once per tensor object and notice in-place edits (`_version`), a different node
    count and a different segment count; the by-source CSR is built lazily, once."""
    from gnnb200 import graph as graph_mod
    ei = torch.randint(0, 50, (2, 300))
    g1 = graph_mod.graph_of(ei, 50)
    assert graph_mod.graph_of(ei, 50) is g1 and stubbed['gnnb200_csr_build_i64'] == 1
    g1.rowptr_t, g1.col_t, g1.rowptr_t                              # noqa: B018 — lazily built, once
    assert stubbed['gnnb200_csr_build_i64'] == 2
    assert graph_mod.graph_of(ei, 60) is not g1                      # other node count
    ei[0, 0] = 7                                                     # in-place edit bumps the tensor's version
    g2 = graph_mod.graph_of(ei, 60)
    assert g2._version == ei._version and stubbed['gnnb200_csr_build_i64'] == 4
    assert graph_mod.graph_of(ei.clone(), 60) is not g2              # another tensor object: its own CSR
    batch = torch.repeat_interleave(torch.arange(4), torch.tensor([3, 1, 0, 5]))
    p1 = graph_mod.segment_ptr_of(batch, 4)
    assert graph_mod.segment_ptr_of(batch, 4) is p1 and stubbed['gnnb200_segment_ptr_i64'] == 1
    assert graph_mod.segment_ptr_of(batch, 6) is not p1 and stubbed['gnnb200_segment_ptr_i64'] == 2
    assert graph_mod.segment_ptr_of(batch).numel() == 4 + 1          # size read from the last element (like PyG)
