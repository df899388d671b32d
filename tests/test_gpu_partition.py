"""The node-partitioned path with the real kernels: world size 1 reproduces the single-device path (forward bit
for bit); world size 2 over NCCL (needs 2 GPUs, skipped otherwise) reproduces the single-device activations
and gradients of the same graph."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnnb200  # noqa: F401
from gnnb200 import models as prod
from gnnb200 import partition, synthetic

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _single(dev, data, n, layers, seed=0):
    torch.manual_seed(seed)
    model = torch.nn.ModuleDict({'input_encoder': prod.InputEncoder(100, 256), 'gnn_backbone': prod.GINBackbone(layers, 256)}).to(dev)
    model.train()
    old = prod.DROPOUT_RATE
    prod.DROPOUT_RATE = 0.0
    model['input_encoder'].dropout.p = 0.0
    try:
        h = model['gnn_backbone'](model['input_encoder'](data['x']), data['edge_index'])
        w = torch.randn(h.shape, generator=torch.Generator().manual_seed(3)).to(dev)
        (h * w).sum().backward()
    finally:
        prod.DROPOUT_RATE = old
    return h.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}, w


def test_world1_partition_equals_single_device():
    dev = torch.device('cuda')
    n, e = 5000, 60000
    data = synthetic.products_like(n, e, 100, seed=1, device=dev)
    torch.manual_seed(0)
    step = partition.PartitionedBackboneStep(prod, dev, 100, 256, 3, n, 0, 1, seed=0)
    torch.manual_seed(5)
    l1 = step.step(data['x'], data['edge_index'])
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({'input_encoder': prod.InputEncoder(100, 256), 'gnn_backbone': prod.GINBackbone(3, 256)}).to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    torch.manual_seed(5)
    opt.zero_grad(set_to_none=True)
    h = model['gnn_backbone'](model['input_encoder'](data['x']), data['edge_index'])
    l2 = h.sum()
    l2.backward()
    opt.step()
    assert torch.equal(l1, l2)                      # same kernels, same order, same dropout seeds in the forward pass
    # The partitioned path keeps one autograd node per kernel, the single-device path one per layer whose transposed
    # gather continues the residual gradient's buffer: dh = (ds + A^T dz) + (1+eps) dz versus ds + (A^T dz + (1+eps) dz).
    # Gradients therefore agree to fp32 re-association, not bitwise.  The weights after the AdamW step are NOT a
    # rounding-level check (Adam's first update is lr * g / (|g| + 1e-8) ~ +-lr, so an element whose near-zero gradient
    # changes sign moves by 2 lr): they are only bounded by that.
    lr = 1e-4
    for (ka, pa), (kb, pb) in zip(step.model.named_parameters(), model.named_parameters()):
        assert ka == kb
        ga, gb, pa, pb = pa.grad, pb.grad, pa.detach(), pb.detach()
        assert (ga is None) == (gb is None), ka
        err, ref = float((ga - gb).norm()), float(gb.norm())
        # measured on B200 (tf32 GEMMs): <= 1.7e-4 for the weights of layers 0/1 and the encoder, 0 for the last layer
        tol = 5e-2 if ka.endswith('eps') else 1e-3          # d(eps) is one cancelling dot product over all rows
        print(f'{ka}: |dg| = {err:.3e}, |g| = {ref:.3e}, max |dw| = {float((pa - pb).abs().max()):.3e}')
        assert err <= tol * ref + 1e-7, (ka, err, ref)
        assert float((pa - pb).abs().max()) <= 2 * lr * 1.01, ka


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n, e, layers = 6001, 80000, 3
        data = synthetic.products_like(n, e, 100, seed=2, device=dev)
        h_ref, g_ref, w = _single(dev, data, n, layers)
        torch.manual_seed(0)
        model = torch.nn.ModuleDict({'input_encoder': prod.InputEncoder(100, 256), 'gnn_backbone': prod.GINBackbone(layers, 256)}).to(dev)
        model.train()
        model['input_encoder'].dropout.p = 0.0
        old = prod.DROPOUT_RATE
        prod.DROPOUT_RATE = 0.0
        graph = partition.PartitionedGraph(data['edge_index'], n, rank, world)
        with partition.partition_scope(graph):
            h = model['gnn_backbone'](model['input_encoder'](data['x'][graph.lo:graph.hi]), graph)
            (h * w[graph.lo:graph.hi]).sum().backward()
        prod.DROPOUT_RATE = old
        partition.allreduce_gradients(model)
        act = float((h.detach() - h_ref[graph.lo:graph.hi]).abs().max() / h_ref.abs().max())
        assert act < 2e-3, act                      # tf32 GEMMs + different BN merge order
        for k, p in model.named_parameters():
            ref = g_ref[k]
            if k.endswith(('linear.bias', 'nn.0.bias', 'nn.3.bias', 'eps')):
                continue        # a bias feeding BatchNorm has a zero gradient; d(eps) is one cancelling sum (tf32 noise dominates)
            fro = float((p.grad - ref).norm() / ref.norm())
            assert fro < 5e-2, (k, fro)
        # running statistics are global, identical on every rank
        rm = model['gnn_backbone'].layers[0].batch_norm.running_mean.clone()
        both = [torch.empty_like(rm) for _ in range(world)]
        dist.all_gather(both, rm)
        assert torch.equal(both[0], both[1])
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_world2_nccl_matches_single_device(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))


def _sparse_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n, e = 6001, 80000
        data = synthetic.products_like(n, e, 100, seed=2, locality=0.9, blocks=16, device=dev)
        x = torch.randn(n, 256, generator=torch.Generator().manual_seed(4)).to(dev)
        eps = torch.tensor([0.25], device=dev)
        dense = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='dense', chunks=1)
        sparse = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='sparse')
        auto = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='auto')
        assert sparse.halo == 'sparse' and sparse.plan.halo_rows < n - sparse.n_local
        assert auto.halo == 'sparse_overlap'                          # a graph with locality: the overlapped sparse exchange
        slabbed = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='dense', chunks=1)
        slabbed.slabs = 4                  # the dense exchange pipelined over column slabs equals the single all-gather bit for bit
        for transposed in (False, True):
            xs = x[dense.lo:dense.hi].contiguous()
            assert torch.equal(dense.aggregate(xs, eps, transposed), slabbed.aggregate(xs, eps, transposed))
        for transposed in (False, True):
            a = dense.aggregate(x[dense.lo:dense.hi].contiguous(), eps, transposed)
            b = sparse.aggregate(x[sparse.lo:sparse.hi].contiguous(), eps, transposed)
            c = auto.aggregate(x[auto.lo:auto.hi].contiguous(), eps, transposed)
            assert torch.equal(a, b)                                  # same edge order per row => same bits
            # local neighbours first, remote neighbours second: the same sum in another association (fp32 rounding)
            assert float((a - c).abs().max()) <= 1e-5 * float(a.abs().max())
            assert torch.equal(c, auto.aggregate(x[auto.lo:auto.hi].contiguous(), eps, transposed))     # deterministic
        # 'sparse_pull': the halo rows pulled from the owners' published buffers instead of sent; same association as
        # 'sparse_overlap', hence the same bits; several passes exercise the two-buffer rotation and the side stream
        pull = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='sparse_pull')
        for it in range(4):
            xi = torch.randn(n, 256, generator=torch.Generator().manual_seed(40 + it)).to(dev)
            for transposed in (False, True):
                want = auto.aggregate(xi[auto.lo:auto.hi].contiguous(), eps, transposed)
                assert torch.equal(pull.aggregate(xi[pull.lo:pull.hi].contiguous(), eps, transposed), want), (it, transposed)
        torch.cuda.synchronize()
        dist.barrier()
        for rows in list(partition.PeerRows._cache.values()):
            rows.close()
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_world2_sparse_halo_equals_dense(tmp_path):
    mp.spawn(_sparse_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))


def _peer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n, e = 6001, 80000
        data = synthetic.products_like(n, e, 100, seed=2, locality=0.9, blocks=16, device=dev)
        eps = torch.tensor([0.25], device=dev)
        dense = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='dense', chunks=1)
        peer = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='peer')
        pull = partition.PartitionedGraph(data['edge_index'], n, rank, world, halo='peercopy')
        # several passes in a row with fresh values: exercises the two-buffer rotation and the per-pass barrier
        for it in range(5):
            x = torch.randn(n, 256, generator=torch.Generator().manual_seed(4 + it)).to(dev)
            for transposed in (False, True):
                a = dense.aggregate(x[dense.lo:dense.hi].contiguous(), eps, transposed)
                b = peer.aggregate(x[peer.lo:peer.hi].contiguous(), eps, transposed)
                c = pull.aggregate(x[pull.lo:pull.hi].contiguous(), eps, transposed)
                assert torch.equal(a, b) and torch.equal(a, c), (it, transposed)   # same edge order per row => same bits
        torch.cuda.synchronize()
        dist.barrier()
        for rows in list(partition.PeerRows._cache.values()):
            rows.close()
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_world2_peer_halo_equals_dense(tmp_path):
    mp.spawn(_peer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))


def _c4_worker(rank, world, port, out_dir):
    """BASELINE configs[3] on NCCL: every rank runs gnnb200.pretrain.train_step on its own graphs with the flat gradient
    all-reduce (mean); replicas start from the same seed, so after 3 steps every parameter and optimizer-visible buffer
    must hold the same BITS on every rank (BatchNorm running statistics stay per replica: DDP semantics)."""
    import importlib.util
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location('bench_for_c4_test', os.path.join(root, 'bench.py'))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        step, sampler = bench.build_c4_step(dev, rank, world)
        start = torch.cat([p.detach().reshape(-1) for p in step.model.parameters()]).clone()
        metrics = [step(sampler.draw()) for _ in range(3)]
        flat = torch.cat([p.detach().reshape(-1) for p in step.model.parameters()])
        assert torch.isfinite(flat).all() and all(torch.isfinite(torch.tensor(float(m['train/loss/total']))) for m in metrics)
        assert not torch.equal(flat, start)                           # the optimizer moved the weights
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        for r in range(1, world):
            assert torch.equal(both[0], both[r]), f'replica {r} diverged from replica 0'
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_world2_c4_data_parallel_replicas_stay_identical(tmp_path):
    mp.spawn(_c4_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(2))
