#!/usr/bin/env python
"""bench.py — headline measurement of the gnnb200 hot path (contract: see the task statement).

Workload (config.workload = "c5_products_backbone"): BASELINE.json configs[4], the largest config
and the one the metric's "% HBM roofline" is quoted on — a synthetic ogbn-products-shaped graph
(N=2,449,029 nodes, E=61,859,140 directed edges in arbitrary COO order, F_in=100, hidden 256),
InputEncoder + 5-layer GIN backbone in TRAIN mode (BatchNorm batch statistics, dropout active),
loss = h.sum(), backward, AdamW step.  One step = CSR+CSC build from edge_index, forward,
backward, optimizer.  metric = aggregated edges/sec = E * L * 2 (fwd + bwd sweeps) / step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S] [--locality P] [--halo dense|sparse|auto|peer|peercopy]
  python bench.py --workload c4 ...      BASELINE configs[3]: data-parallel s5 pre-training step (steps/s, weak scaling)

N > 1 (launched by torch.distributed.run): the same graph node-partitioned into N contiguous
destination ranges with an NCCL all-gather of the layer input per layer ("strong" scaling).
--impl reference times the reference's CPU path (the oracle: reference modules restated over the
PyG shim) on a bounded node/edge sample of the same workload with all host threads.
"""
import argparse
import datetime
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

DTYPE_NAMES = {'f32': 'f32', 'tf32': 'f32+tf32', 'tf32x3': 'f32 (dense transforms: 3xTF32 on tcgen05, fp32-class)',
               'tf32_fwd3': 'f32 storage; forward GEMMs 3xTF32 (fp32-class), backward GEMMs tf32, fp32 accumulate'}
METRIC = 'aggregated_edges_per_sec_fwd_bwd'
UNIT = 'edges/s'
LAYERS = 5
HIDDEN = 256
C5_N, C5_E, C5_F = 2_449_029, 61_859_140, 100


def gnn_precision():
    from gnnb200 import nn as gnn
    return gnn.default_precision()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return {'hbm_gbs': float(p['hbm_gbs']), 'source': 'measured (MEASURED_PEAKS.json, burst copy)'}
    return {'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


def aggregation_bytes(n, e, f):
    """Algorithmic bytes of one aggregation pass: neighbour rows (no-reuse gather
    model) + self-term read + output write + col ids + row pointers."""
    return e * f * 4 + n * f * 4 + n * f * 4 + e * 4 + (n + 1) * 4


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0, period=0.2):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for attr, label in (('nvmlClocksEventReasonHwSlowdown', 'hw_slowdown'),
                            ('nvmlClocksEventReasonHwThermalSlowdown', 'hw_thermal_slowdown'),
                            ('nvmlClocksEventReasonSwThermalSlowdown', 'sw_thermal_slowdown'),
                            ('nvmlClocksEventReasonSwPowerCap', 'sw_power_cap'),
                            ('nvmlClocksThrottleReasonHwSlowdown', 'hw_slowdown'),
                            ('nvmlClocksThrottleReasonHwThermalSlowdown', 'hw_thermal_slowdown'),
                            ('nvmlClocksThrottleReasonSwThermalSlowdown', 'sw_thermal_slowdown'),
                            ('nvmlClocksThrottleReasonSwPowerCap', 'sw_power_cap')):
            if hasattr(nv, attr):
                names[getattr(nv, attr)] = label
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                    getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
                mask = get(self.h)
                for bit, label in names.items():
                    if mask & bit:
                        self.reasons.add(label)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        return {'sm_mhz': statistics.median(self.samples) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_graph(n, e, f, seed, locality, device, skew=0.0):
    from gnnb200 import synthetic
    return synthetic.products_like(n, e, f, seed=seed, locality=locality, device=device, skew=skew)


def build_model(impl_models, device, f_in, seed=0):
    torch.manual_seed(seed)
    enc = impl_models.InputEncoder(f_in, HIDDEN)
    bb = impl_models.GINBackbone(LAYERS, HIDDEN)
    model = torch.nn.ModuleDict({'input_encoder': enc, 'gnn_backbone': bb}).to(device)
    model.train()
    return model


def step_fn(model, opt, x, edge_index):
    """One training step of the hot path.  The sorted CSR / CSC of `edge_index` is cached on the tensor object until it is
    overwritten (gnnb200.graph.graph_of; SURVEY §8d: a one-off cost amortised over every pass on that graph): the
    device-resident leg builds it once (reported as `structure_build_ms`), the end-to-end leg — whose edge list arrives
    from the host every step — rebuilds it every step."""
    ei = edge_index
    opt.zero_grad(set_to_none=True)
    h = model['gnn_backbone'](model['input_encoder'](x), ei)
    loss = h.sum()
    loss.backward()
    opt.step()
    return loss


def partition_selfcheck(prod, dev, rank, world, halo):
    """Parity of the path that is about to be timed, on the hardware it is timed on: a 200 k-node / 5 M-edge graph of the
    same generator, same seeded model on every rank, dropout off.  All ranks run one node-partitioned forward + backward
    (+ the flat gradient all-reduce); rank 0 also runs the single-device path and compares: the aggregation of a fixed
    [n, 256] matrix, forward and transposed (pure gather in edge order: must be bit-identical), the backbone output
    (BatchNorm moments are merged in another order across ranks: fp32 rounding), the loss, and the Frobenius error over
    all parameter gradients.  N = 1 reports
    the single-device loss of the same graph and seed, so the lines of a scaling run can be read side by side."""
    import torch.distributed as dist
    from gnnb200 import partition
    n, e = 200_000, 5_000_000
    data = make_graph(n, e, C5_F, 7, 0.5, dev)
    x, ei = data['x'], data['edge_index']
    w = torch.randn(n, HIDDEN, device=dev, generator=torch.Generator(device=dev).manual_seed(3))    # fixed read-out
    old_p = prod.DROPOUT_RATE
    prod.DROPOUT_RATE = 0.0

    def fresh():
        m = build_model(prod, dev, C5_F, seed=11)
        m['input_encoder'].dropout.p = 0.0
        return m
    out = {'graph': {'nodes': n, 'edges': e}}
    try:
        ref = None
        if rank == 0:
            m1 = fresh()
            from gnnb200 import _lib as L_, ops as ops_
            from gnnb200.graph import graph_of
            h0 = m1['input_encoder'](x)
            h1 = m1['gnn_backbone'](h0, ei)
            loss1 = (h1 * w).sum()
            loss1.backward()
            g1, eps1 = graph_of(ei, n), torch.tensor([0.25], device=dev)
            z1 = torch.cat([ops_._aggregate_raw(w, g1.rowptr, g1.col, L_.AGG_SUM, w, eps1, None),
                            ops_._aggregate_raw(w, g1.rowptr_t, g1.col_t, L_.AGG_SUM, w, eps1, None)], dim=1)
            ref = (z1, h1.detach(), float(loss1.detach()), torch.cat([p.grad.reshape(-1) for p in m1.parameters()]))
            out['loss_single_device'] = ref[2]
        if world > 1:
            m = fresh()
            graph = partition.PartitionedGraph(ei, n, rank, world, halo=halo)
            with partition.partition_scope(graph):
                h0 = m['input_encoder'](x[graph.lo:graph.hi])
                h = m['gnn_backbone'](h0, graph)
                loss = (h * w[graph.lo:graph.hi]).sum()
                loss.backward()
            partition.allreduce_gradients(m)
            w_local, eps1 = w[graph.lo:graph.hi].contiguous(), torch.tensor([0.25], device=dev)
            z = torch.cat([graph.aggregate(w_local, eps1, False), graph.aggregate(w_local, eps1, True)], dim=1)
            total = loss.detach().clone()
            dist.all_reduce(total)

            def everyone(t):                               # [n, F] on every rank from the row shards
                pad = t.new_zeros(graph.per, t.size(1))
                pad[: t.size(0)] = t
                full = t.new_empty(graph.per * world, t.size(1))
                dist.all_gather_into_tensor(full, pad)
                return full[:n]
            z_all, h_all = everyone(z.detach()), everyone(h.detach())
            if rank == 0:
                g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
                out.update({'halo': graph.halo, 'aggregation_bitwise': bool(torch.equal(z_all, ref[0])),
                            'fwd_max_rel': float((h_all - ref[1]).abs().max() / ref[1].abs().max()),
                            'loss_partitioned': float(total), 'loss_rel': abs(float(total) - ref[2]) / abs(ref[2]),
                            'grad_fro': float((g - ref[3]).norm() / ref[3].norm())})
                out['aggregation_max_rel'] = float((z_all - ref[0]).abs().max() / ref[0].abs().max())
                # 'sparse_overlap' sums a row's local neighbours before its remote ones: fp32 rounding, not bit identity
                exact = out['aggregation_bitwise'] or (graph.halo in ('sparse_overlap', 'sparse_pull') and out['aggregation_max_rel'] < 1e-5)
                out['ok'] = bool(exact and out['fwd_max_rel'] < 1e-3 and out['grad_fro'] < 2e-2)
    finally:
        prod.DROPOUT_RATE = old_p
    del data, x, ei, w
    torch.cuda.empty_cache()
    return out


def run_product(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # short collective timeout: a mismatched collective must abort the run, not hold 8 GPUs for ten minutes
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    import gnnb200  # noqa: F401
    from gnnb200 import models as prod, ops
    from gnnb200 import nn as gnn
    if args.precision:
        gnn.set_default_precision(args.precision)

    n = max(1024, int(C5_N * args.scale))
    e = max(4096, int(C5_E * args.scale))
    data = make_graph(n, e, C5_F, 42, args.locality, dev, args.skew)
    x_dev, ei_dev = data['x'], data['edge_index']

    if world > 1:
        from gnnb200 import partition
        runner = partition.PartitionedBackboneStep(prod, dev, C5_F, HIDDEN, LAYERS, n, rank, world, halo=args.halo)
        one_step = lambda x, ei, local=False: runner.step(x, ei, local)  # noqa: E731
    else:
        model = build_model(prod, dev, C5_F)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        one_step = lambda x, ei: step_fn(model, opt, x, ei)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    selfcheck = None if args.no_selfcheck else partition_selfcheck(prod, dev, rank, world, args.halo)

    # ---- device-resident leg -----------------------------------------------------------------
    # the graph structure (CSR + CSC on one GPU; ownership scan, sorts and halo plan when partitioned), timed on its own
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    b0.record()
    if world > 1:
        runner.graph_of(ei_dev)
    else:
        from gnnb200.graph import graph_of as _graph_of
        _g = _graph_of(ei_dev, n)
        _g.rowptr_t                                  # the by-source CSR of the backward pass
    b1.record()
    barrier()
    tb = torch.tensor([b0.elapsed_time(b1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    structure_ms = float(tb)
    loss0 = None
    for i in range(args.warmup):
        l_ = one_step(x_dev, ei_dev)
        if i == 0:                                   # step-0 loss (dropout active: masks depend on the partition)
            l0 = l_.detach().double().clone()
            if world > 1:
                dist.all_reduce(l0)
            loss0 = float(l0)
    barrier()
    ops.reset_counters()
    ops.AGG_TIMER = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            one_step(x_dev, ei_dev)
        stop.record()
        barrier()
    ms = start.elapsed_time(stop)
    main_halo = runner.last_halo if world > 1 else None          # (the generator-(ii) leg below may pick another mode)
    agg_ms = [a.elapsed_time(b) for a, b in ops.AGG_TIMER]
    ops.AGG_TIMER = None
    launches = ops.launch_count()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    ms_per_step = ms / args.steps
    value = e * LAYERS * 2 / (ms_per_step / 1e3)

    # ---- end-to-end leg: host buffers, H2D of the step's inputs and D2H of the loss every step ----
    # Multi-GPU: every rank reads only ITS share from the host — its row shard of x and a 1/N slice of the
    # edge columns; the slices are all-gathered over NVLink inside the timed step to rebuild edge_index.
    if world > 1:
        from gnnb200.partition import shard_bounds
        lo, hi, _ = shard_bounds(n, rank, world)
        per_e = (e + world - 1) // world
        e_lo, e_hi = min(e, rank * per_e), min(e, (rank + 1) * per_e)
        x_host = x_dev[lo:hi].cpu().pin_memory()
        ei_slice = torch.full((2, per_e), -1, dtype=torch.long)
        ei_slice[:, : e_hi - e_lo] = ei_dev[:, e_lo:e_hi].cpu()
        ei_host = ei_slice.pin_memory()
    else:
        x_host = x_dev.cpu().pin_memory()
        ei_host = ei_dev.cpu().pin_memory()
    del x_dev, ei_dev, data
    torch.cuda.empty_cache()

    # Every step's inputs come from pinned host memory and its loss goes back to the host; the copy of
    # step i+1's inputs runs on a side stream while step i computes (double-buffered device inputs).
    # The two sets of device input buffers are allocated ONCE (a fresh 2 GB allocation per step on the side stream cost
    # +32 % on a 16-core box in round 1); a set is overwritten only after the step that read it has finished (`done`).
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty_like(x_host, device=dev), torch.empty_like(ei_host, device=dev)) for _ in range(2)]
    done = [None, None]
    turn = [0]

    def upload():
        b = turn[0] & 1
        turn[0] += 1
        xd, eid = slots[b]
        with torch.cuda.stream(copy_stream):
            if done[b] is not None:
                copy_stream.wait_event(done[b])
            xd.copy_(x_host, non_blocking=True)
            eid.copy_(ei_host, non_blocking=True)
            if world == 1:
                # the new edge list's CSR + CSC are built on the copy stream as well, right behind its upload, so that they
                # overlap the previous step's compute like the copies do (the step then finds them cached on the tensor)
                from gnnb200.graph import graph_of as _graph_of
                _graph_of(eid, n).rowptr_t
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return xd, eid, ready, b

    pending = [None]

    def finish(b):
        done[b] = torch.cuda.Event()
        done[b].record(torch.cuda.current_stream(dev))

    def e2e_step():
        xd, eid, ready, b = pending[0] if pending[0] is not None else upload()
        torch.cuda.current_stream(dev).wait_event(ready)
        if world > 1:
            pending[0] = upload()                  # next step's H2D overlaps this step's compute
            parts = eid.new_empty(world, 2, eid.size(1))
            dist.all_gather_into_tensor(parts.view(world * 2, -1), eid)
            full = parts.permute(1, 0, 2).reshape(2, -1)
            full = full[:, full[0] >= 0]           # drop the padding columns of the last slice
            loss = one_step(xd, full, True)
            finish(b)
        else:
            loss = one_step(xd, eid)               # enqueue this step first: the hub-row listing of the next edge list's
            finish(b)                              # structure build reads one count back, which must not stall the launches
            pending[0] = upload()                  # next step's H2D + structure build overlap this step's compute
        return float(loss.detach().cpu())          # D2H read of the step's result

    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.no_e2e:
        s2.record()
        e2.record()
        barrier()
    else:
        for _ in range(min(args.warmup, 3)):
            e2e_step()
        barrier()
        s2.record()
        for _ in range(args.steps):
            e2e_step()
        e2.record()
        barrier()
        pending[0] = None
    t2 = torch.tensor([max(s2.elapsed_time(e2), 1e-9)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2) / args.steps
    e2e_value = e * LAYERS * 2 / (e2e_ms / 1e3)

    # ---- generator (ii) of SURVEY §8d.5 beside the headline (which runs generator (i), uniform-random edges = worst-case
    #      locality): power-law degrees + 90 % intra-block edges, the same step on the same model, device-resident ----
    gen2 = None
    if not args.no_generator2 and args.locality == 0.0 and args.skew == 0.0:
      try:                                             # a failure here must not lose the headline line
        pending[0] = None
        slots.clear()
        torch.cuda.empty_cache()
        d2 = make_graph(n, e, C5_F, 42, 0.9, dev, 1.8)
        x2, ei2 = d2['x'], d2['edge_index']
        for _ in range(2):
            one_step(x2, ei2)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            one_step(x2, ei2)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        g_ms = float(tg) / args.steps
        gen2 = {'generator': '(ii) power-law degrees (skew 1.8) + 90 % intra-block edges (block = N/64)', 'ms_per_step': g_ms,
                'value': e * LAYERS * 2 / (g_ms / 1e3), 'unit': UNIT,
                'halo': runner.last_halo if world > 1 else None}
        del d2, x2, ei2
      except Exception as exc:                         # noqa: BLE001 — reported in the JSON line
        gen2 = {'error': f'{type(exc).__name__}: {exc}'[:300]}

    out = None
    if rank == 0:
        pk = peaks()
        n_local = (n + world - 1) // world
        e_local = e // world
        agg_bytes = aggregation_bytes(n_local, e_local, HIDDEN)
        # one aggregation PASS = one launch on a single GPU, `chunks` launches (one per halo piece) when partitioned
        passes = args.steps * LAYERS * 2
        agg_avg = (sum(agg_ms) / passes) if agg_ms else None
        achieved = agg_bytes / (agg_avg / 1e3) / 1e9 if agg_avg else None
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.isfile(prof):
            traffic = json.load(open(prof)).get(f'aggregate_scale{args.scale}_loc{args.locality}')
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': DTYPE_NAMES[gnn.default_precision()],
            'data': 'synthetic',
            'config': {'workload': 'c5_products_backbone', 'nodes': n, 'edges': e, 'feat_in': C5_F,
                       'hidden': HIDDEN, 'layers': LAYERS,
                       'mode': 'train fwd+bwd+AdamW; graph structure built once per edge list (every step in the e2e leg)',
                       'edge_locality': args.locality, 'degree_skew': args.skew, 'gemm_precision': gnn.default_precision(),
                       'l2_policy': 'inputs_larger_than_L2 (2.5 GB activations per layer vs 126 MB L2)',
                       'parallelism': 'single' if world == 1 else f'node_partition{world}+' + (
                           {'sparse': 'halo_alltoall_sparse', 'sparse_overlap': 'halo_alltoall_sparse_overlapped_with_local_gather',
                            'sparse_pull': 'halo_rows_pulled_over_nvlink_peer_memory_overlapped_with_local_gather',
                            'peer': 'halo_read_in_gather_over_nvlink_peer_memory',
                            'peercopy': 'halo_allgather_by_copy_engines'}.get(
                               main_halo, 'halo_allgather'))},
            'clocks': clocks.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': x_host.numel() * 4 + ei_host.numel() * 8, 'd2h_bytes_per_step': 4,
                    'note': 'inputs from pinned host memory every step (H2D and the structure build of the new edge list on a side stream, one step ahead), loss read back every step'},
            'gpu_launches': launches,
            'selfcheck': selfcheck, 'loss_step0': loss0, 'structure_build_ms': structure_ms, 'generator_ii': gen2,
            'peak_mem_gb': torch.cuda.max_memory_allocated() / 2**30,
            'roofline': {'bound': 'hbm', 'kernel': 'aggregate_vec_kernel<32,2,SUM,2,6> (fwd and transposed bwd)',
                         'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                         'frac': achieved / pk['hbm_gbs'] if achieved else None, 'traffic': traffic,
                         'peak_source': pk['source'], 'launches_timed': len(agg_ms), 'passes_timed': passes,
                         'avg_launch_ms': agg_avg, 'algorithmic_bytes_per_launch': agg_bytes,
                         'share_of_step': (sum(agg_ms) / ms) if agg_ms else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(args)
        if world == 1 and not args.no_secondary:
            # the small-graph configs are reported beside the headline; a failure there must not lose the headline line
            try:
                sec = small_graph_steps('gnnb200', dev, steps=40, warmup=6)
                if not args.no_cpu_baseline:
                    torch.set_num_threads(os.cpu_count() or 1)
                    cpu = small_graph_steps('oracle', torch.device('cpu'), steps=4, warmup=1)
                    sec['cpu_oracle'] = {k: v for k, v in cpu.items() if k.endswith('_per_s')}
                    sec['cpu_oracle']['cores'] = torch.get_num_threads()
            except Exception as exc:                      # noqa: BLE001 — reported in the JSON line, not swallowed
                sec = {'error': f'{type(exc).__name__}: {exc}'}
            sec['c4_s5_data_parallel_n1'] = c4_in_subprocess()
            out['secondary'] = sec
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def c4_in_subprocess(steps=10, warmup=3, timeout=180):
    """BASELINE configs[3] at one replica (`--workload c4`), in its OWN process: the s5 training iteration is the newest
    code on the device and must not be able to take the headline line down with it.  Returns that run's headline fields."""
    import subprocess
    try:
        torch.cuda.empty_cache()
        run = subprocess.run([sys.executable, os.path.abspath(__file__), '--workload', 'c4', '--steps', str(steps), '--warmup',
                              str(warmup)], capture_output=True, text=True, timeout=timeout,
                             env={k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')})
        for line in reversed(run.stdout.strip().splitlines()):
            if line.startswith('{'):
                js = json.loads(line)
                return {'steps_per_s': js['value'], 'ms_per_step': js['ms_per_step'],
                        'graphs_per_s': js['config'].get('graphs_per_sec'), 'e2e_steps_per_s': js['e2e']['value'],
                        'gpu_launches_per_step': js.get('gpu_launches', 0) / max(steps, 1)}
        return {'error': (run.stderr.strip().splitlines() or ['no output'])[-1][:300], 'returncode': run.returncode}
    except Exception as exc:                              # noqa: BLE001 — reported, not swallowed
        return {'error': f'{type(exc).__name__}: {exc}'[:300]}


# ------------------------------------------------------------------------------------------------
# secondary workloads: BASELINE configs[1] (ENZYMES-shaped fine-tune step) and configs[2] (s4 pre-training step)
# ------------------------------------------------------------------------------------------------
S4_TASKS = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop']
S4_DOMAINS = ['Cora_NC', 'CiteSeer_NC', 'ENZYMES']


def _graph_lists():
    from gnnb200 import synthetic
    return {'c2': synthetic.tu_like_graphs('ENZYMES', 128, seed=42),
            'Cora_NC': [dict(synthetic.cora_like(42), graph_properties=torch.zeros(12))],
            'CiteSeer_NC': [dict(synthetic.citeseer_like(43), graph_properties=torch.zeros(12))],
            'ENZYMES': synthetic.tu_like_graphs('ENZYMES', 32, seed=44)}


def _make_batch(data_mod, graphs, device):
    b = data_mod.Batch.from_data_list([data_mod.Data(**{k: v.clone() for k, v in g.items()}) for g in graphs])
    return b.to(device)


def small_graph_steps(impl, device, steps, warmup):
    """(fine-tune steps/s on C2, s4 pre-training steps/s on C3) for `impl` in {'gnnb200', 'oracle'}.
    The s4 step follows run_training (src/pretrain/pretrain.py:124-153): the 5 task losses of scheme s4 on three
    domains, gradient surgery (one backward per task + the PCGrad projection), gradient clipping, AdamW, temperature
    step; logging / loss balancer (host scalars that never reach the gradients when > 1 task is active) are left out."""
    import random
    if impl == 'gnnb200':
        from gnnb200 import data as data_mod, models, tasks as task_mod
        from gnnb200.gradient_surgery import GradientSurgery
    else:
        from oracle import modules as models
        from oracle import install_pyg_shim
        install_pyg_shim()
        import torch_geometric.data as data_mod
        task_mod = models
        GradientSurgery = models.GradientSurgery
    sync = (lambda: torch.cuda.synchronize()) if device.type == 'cuda' else (lambda: None)
    lists = _graph_lists()
    out = {}
    # ---- C1: Cora-shaped backbone forward+backward (BASELINE configs[0]: 2,708 nodes, 10,556 edges, 1,433 feats, L=3) ----
    from gnnb200 import synthetic as _syn
    cora = _syn.cora_like(42)
    torch.manual_seed(0)
    c1 = torch.nn.ModuleDict({'input_encoder': models.InputEncoder(1433, HIDDEN), 'gnn_backbone': models.GINBackbone(3, HIDDEN)}).to(device)
    c1.train()
    cx, cei = cora['x'].to(device), cora['edge_index'].to(device)

    def c1_step():
        c1.zero_grad(set_to_none=True)
        c1['gnn_backbone'](c1['input_encoder'](cx), cei.view_as(cei)).sum().backward()
    for _ in range(warmup):
        c1_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        c1_step()
    sync()
    dt = (time.perf_counter() - t0) / steps
    out['c1_cora_backbone_fwd_bwd_ms'] = dt * 1e3
    out['c1_edges_per_s'] = 10556 * 3 * 2 / dt
    # ---- C2: graph-classification fine-tune step ----
    torch.manual_seed(0)
    ft = models.FinetuneGNN(device, 'ENZYMES', 'full_finetune')
    ft.train()
    opt = torch.optim.AdamW(ft.param_groups)
    batch = _make_batch(data_mod, lists['c2'], device)

    def ft_step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(ft(batch), batch.y)
        loss.backward()
        opt.step()
        return loss
    for _ in range(warmup):
        ft_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ft_step()
    sync()
    out['c2_finetune_steps_per_s'] = steps / (time.perf_counter() - t0)
    out['c2_shape'] = {'graphs': 128, 'nodes': int(batch.x.size(0)), 'edges': int(batch.edge_index.size(1))}
    # ---- C3: s4 multi-task pre-training step ----
    torch.manual_seed(0)
    pm = models.PretrainableGNN(device, S4_DOMAINS, S4_TASKS)
    pm.train()
    popt = torch.optim.AdamW(pm.parameters(), lr=1e-4)
    temp, grl = task_mod.TemperatureScheduler(1000), task_mod.GRLScheduler(50, 100)
    tasks = task_mod.instantiate_tasks(pm, S4_TASKS, grl, temp)
    batches = {d: _make_batch(data_mod, lists[d], device) for d in S4_DOMAINS}
    gen = torch.Generator().manual_seed(42)
    random.seed(42)

    surgery = GradientSurgery(device)

    def s4_step():
        losses = {name: task.compute_loss(batches, gen)[0] for name, task in tasks.items()}
        popt.zero_grad(set_to_none=True)
        surgery.apply_gradient_surgery(pm, losses, list(losses))
        torch.nn.utils.clip_grad_norm_(pm.parameters(), max_norm=0.5)
        popt.step()
        temp.step()
        return losses
    s4_steps = max(2, steps // 4)
    for _ in range(max(1, warmup // 2)):
        s4_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(s4_steps):
        s4_step()
    sync()
    out['c3_s4_pretrain_steps_per_s'] = s4_steps / (time.perf_counter() - t0)
    out['c3_shape'] = {'domains': S4_DOMAINS, 'tasks': S4_TASKS, 'backbone_passes_per_step': 21}
    return out


# ------------------------------------------------------------------------------------------------
# C4: data-parallel pre-training of scheme s5 (BASELINE configs[3]) — `--workload c4`
# ------------------------------------------------------------------------------------------------
S5_TASKS = S4_TASKS + ['domain_adv']
TU_DOMAINS = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']


def build_c4_step(dev, rank, world):
    """(step(batches) -> metrics, sampler) of the C4 workload on `dev` for this rank (see run_c4)."""
    import random
    import gnnb200  # noqa: F401
    from gnnb200 import data as data_mod, loader, models, partition, pretrain, synthetic, tasks as task_mod
    from gnnb200.gradient_surgery import GradientSurgery
    torch.manual_seed(0)
    pm = models.PretrainableGNN(dev, TU_DOMAINS, S5_TASKS)
    pm.train()
    opt = pretrain.TaskSpecificOptimizer(pm, S5_TASKS)
    temp, grl = task_mod.TemperatureScheduler(1000), task_mod.GRLScheduler(50, 20)
    grl.current_step = 600                                        # past the 40 % ramp start: lambda > 0
    tasks = task_mod.instantiate_tasks(pm, S5_TASKS, grl, temp)
    sets = {d: loader.GraphDataset([data_mod.Data(**g) for g in synthetic.tu_like_graphs(d, 256, seed=42 + rank * 7 + i)],
                                   range(256)) for i, d in enumerate(TU_DOMAINS)}
    gen = torch.Generator().manual_seed(42 + rank)
    random.seed(42 + rank)
    sampler = loader.BalancedMultiDomainSampler(sets, gen, device=dev, batch_size=128)
    balancer, surgery = pretrain.AdaptiveLossBalancer(), GradientSurgery(dev)

    def mean_allreduce(model):
        partition.allreduce_gradients(model)
        for p in model.parameters():
            if p.grad is not None:
                p.grad.div_(world)

    def step(batches):
        return pretrain.train_step(pm, tasks, opt, batches, gen, grl, temp, balancer, surgery, TU_DOMAINS,
                                   allreduce=mean_allreduce if world > 1 else None)
    step.model, step.schedulers = pm, (grl, temp, balancer)
    return step, sampler


def run_c4(args):
    """Every rank trains the s5 multi-task step on its own graphs: 128 per step (32 per TU-shaped domain) drawn by the
    balanced multi-domain sampler from that rank's device-resident dataset (256 graphs per domain, seed 42 + rank).
    One step = gnnb200.pretrain.train_step, i.e. the reference's run_training iteration (src/pretrain/pretrain.py:112-184):
    task losses -> loss balancer -> gradient surgery over the five main tasks -> domain-adversarial backward through the
    GRL on top -> [one flat NCCL all-reduce of the gradients, mean] -> clip -> task-specific AdamW -> schedulers -> metrics.
    BatchNorm statistics stay per replica (DDP semantics).  Weak scaling: value = global steps/s (the same on every N),
    graphs/s = 128 * N * steps/s.  Device leg: the same resident batches every step; e2e leg: a fresh draw per step
    (upload of the packed gather indices) and the metrics dict read back."""
    import random
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    from gnnb200 import ops
    step, sampler = build_c4_step(dev, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(make_batches):
        for _ in range(args.warmup):
            step(make_batches())
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            metrics = step(make_batches())
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / args.steps, metrics

    fixed = sampler.draw()
    ops.reset_counters()
    with ClockSampler(local) as clocks:
        ms, _ = timed(lambda: fixed)
    launches = ops.launch_count() * args.steps // (args.steps + args.warmup)
    e2e_ms, metrics = timed(sampler.draw)
    out = None
    if rank == 0:
        h2d = sum(8 * (b.batch.numel() + b.ptr.numel() + b.x.size(0) + 2 * b.edge_index.size(1) + 13 * b.num_graphs)
                  for b in fixed.values())
        out = {'metric': 'pretrain_steps_per_sec', 'value': 1e3 / ms, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps,
               'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': DTYPE_NAMES[gnn_precision()], 'data': 'synthetic',
               'config': {'workload': 'c4_s5_data_parallel', 'graphs_per_rank': 128, 'global_batch': 128 * world,
                          'domains': TU_DOMAINS, 'tasks': S5_TASKS, 'graphs_per_sec': 128 * world * 1e3 / ms,
                          'step': 'gnnb200.pretrain.train_step (reference run_training iteration incl. balancer, surgery, metrics)',
                          'l2_policy': 'working set < L2 (launch-/host-bound regime; roofline not meaningful)'},
               'clocks': clocks.summary(),
               'e2e': {'value': 1e3 / e2e_ms, 'unit': 'steps/s', 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                       'd2h_bytes_per_step': 4 * sum(isinstance(v, float) for v in metrics.values()),
                       'note': 'fresh sampler draw per step from the device-resident dataset (packed index upload), '
                               'metrics dict read back every step'},
               'gpu_launches': launches}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle = the reference's modules restated over the pure-PyTorch PyG shim)
# ------------------------------------------------------------------------------------------------
def cpu_sample_sizes(args, frac):
    return max(1024, int(C5_N * args.scale * frac)), max(4096, int(C5_E * args.scale * frac))


def cpu_run(args, frac, steps, warmup):
    """The reference's CPU path (oracle port: the reference modules restated over the PyG shim) on a `frac` node/edge
    sample of the workload, all host threads; returns (n, e, [seconds per timed step])."""
    from oracle import modules as orc
    torch.set_num_threads(os.cpu_count() or 1)
    n, e = cpu_sample_sizes(args, frac)
    data = make_graph(n, e, C5_F, 42, args.locality, 'cpu', args.skew)
    model = build_model(orc, torch.device('cpu'), C5_F)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step_fn(model, opt, data['x'], data['edge_index'])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n, e, times


def cpu_budget_fraction(args, probe_frac, probe_sec, steps, budget_s, cap):
    """Largest sample fraction <= cap whose (steps + 1 warm-up) fit `budget_s` seconds (step time is linear in the sample:
    the CPU path is gather / GEMM bound) and whose transient [E, 256] fp32 message tensors (3 alive at the peak) fit half
    of the free host memory."""
    frac = min(cap, probe_frac * budget_s / ((steps + 1) * max(probe_sec, 1e-3)))
    try:
        import psutil
        avail = psutil.virtual_memory().available
        frac = min(frac, 0.5 * avail / (3.0 * C5_E * args.scale * HIDDEN * 4))
    except Exception:                                     # noqa: BLE001 — no psutil: keep the time bound only
        pass
    return max(probe_frac, frac)


def cpu_baseline(args):
    """~20-30 s of CPU work on rank 0: 1 warm-up + 3 timed steps (median) on 1/16 of the workload."""
    frac = args.cpu_sample or 1.0 / 16
    n, e, times = cpu_run(args, frac, steps=3, warmup=1)
    sec = statistics.median(times)
    return {'value': e * LAYERS * 2 / sec, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{n} nodes / {e} edges ({frac:g} of the workload), same model and step, median of 3 timed steps '
                      f'({sec:.2f} s) after 1 warm-up', 'seconds_per_step': sec, 'cpu': cpu_model()}


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return None


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the same step on the host cores.  The reference is pure
    Python over PyTorch-Geometric, which cannot be installed here, so what runs is the oracle port of its
    modules (pinned bit for bit against the unmodified reference files, tests/test_oracle_reference.py).  Every step is a
    bounded SAMPLE of the arm's workload: a quarter of the nodes and edges when that fits ~2.5 minutes and the host memory,
    else the largest fraction that does (probed with one step on 1/16); median over the timed steps."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    steps = max(1, args.steps)
    if args.cpu_sample:
        frac = args.cpu_sample
    else:
        probe = 1.0 / 16
        _, _, t = cpu_run(args, probe, steps=1, warmup=0)
        frac = cpu_budget_fraction(args, probe, t[0], steps, budget_s=150.0, cap=0.25)
    n, e, times = cpu_run(args, frac, steps=steps, warmup=1)
    sec = statistics.median(times)
    value = e * LAYERS * 2 / sec
    n_full, e_full = max(1024, int(C5_N * args.scale)), max(4096, int(C5_E * args.scale))
    sample = (f'{n} nodes / {e} edges per step = {frac:.3g} of the workload ({n_full} nodes / {e_full} edges), same model and '
              f'step; median of {steps} timed steps after 1 warm-up')
    base = {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample,
            'cpu': cpu_model(), 'seconds_per_step': sec}
    return {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        # the arm's own config (the rate is per edge, so a sample of the workload measures the same metric)
        'config': {'workload': 'c5_products_backbone', 'nodes': n_full, 'edges': e_full, 'feat_in': C5_F, 'hidden': HIDDEN,
                   'layers': LAYERS, 'mode': 'train fwd+bwd+AdamW; graph structure built once per edge list (every step in the e2e leg)',
                   'edge_locality': args.locality, 'degree_skew': args.skew,
                   'sampled_step': {'nodes': n, 'edges': e, 'fraction': frac,
                                    'why': 'CPU step on the full graph takes minutes and ~190 GB of transient messages'}},
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


# ------------------------------------------------------------------------------------------------
# C1: BASELINE configs[0], the config the reference runs on CPU — `--workload c1`
# ------------------------------------------------------------------------------------------------
def run_c1(args):
    """Cora-shaped backbone forward+backward (N=2,708, E=10,556 directed, F=1,433, H=256) at L=3 (BASELINE) and L=5 (the
    reference's constant), train mode: gnnb200 on cuda:0 and the oracle port on the host cores, median of `steps` after
    `warmup`.  Launch-bound on the GPU (the whole graph is L2-resident): reported in ms and edges/s, no roofline."""
    import gnnb200  # noqa: F401
    from gnnb200 import models as prod, ops, synthetic
    from oracle import modules as orc
    cora = synthetic.cora_like(42)
    steps, warmup = max(args.steps, 10), max(args.warmup, 3)
    res = {}
    for layers in (3, 5):
        for impl, mods, dev in (('gnnb200', prod, torch.device('cuda', 0)), ('cpu_oracle', orc, torch.device('cpu'))):
            if impl == 'cpu_oracle':
                if args.no_cpu_baseline:
                    continue
                torch.set_num_threads(os.cpu_count() or 1)
            torch.manual_seed(0)
            m = torch.nn.ModuleDict({'input_encoder': mods.InputEncoder(1433, HIDDEN),
                                     'gnn_backbone': mods.GINBackbone(layers, HIDDEN)}).to(dev)
            m.train()
            x_host, ei_host = cora['x'], cora['edge_index']
            if dev.type == 'cuda':
                x_host, ei_host = x_host.pin_memory(), ei_host.pin_memory()
            x, ei = x_host.to(dev), ei_host.to(dev)

            def step(e2e=False):
                xs, eis = (x_host.to(dev, non_blocking=True), ei_host.to(dev, non_blocking=True)) if e2e else (x, ei.view_as(ei))
                m.zero_grad(set_to_none=True)
                loss = m['gnn_backbone'](m['input_encoder'](xs), eis).sum()
                loss.backward()
                return float(loss) if e2e else loss

            def timed(e2e):
                ts = []
                for i in range(warmup + steps):
                    if dev.type == 'cuda':
                        torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    step(e2e)
                    if dev.type == 'cuda':
                        torch.cuda.synchronize()
                    if i >= warmup:
                        ts.append(time.perf_counter() - t0)
                return statistics.median(ts)
            ops.reset_counters()
            sec = timed(False)
            entry = {'ms': sec * 1e3, 'edges_per_s': 10556 * layers * 2 / sec}
            if dev.type == 'cuda':
                entry['gpu_launches_per_step'] = ops.launch_count() / (warmup + steps)
                entry['e2e_ms'] = timed(True) * 1e3
            else:
                entry['cores'] = torch.get_num_threads()
            res[f'L{layers}_{impl}'] = entry
    g = res['L3_gnnb200']
    out = {'metric': METRIC, 'value': g['edges_per_s'], 'unit': UNIT, 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
           'ms_per_step': g['ms'], 'higher_is_better': True, 'scaling': 'replicas only', 'vs_baseline': None,
           'dtype': DTYPE_NAMES[gnn_precision()], 'data': 'synthetic',
           'config': {'workload': 'c1_cora_backbone', 'nodes': 2708, 'edges': 10556, 'feat_in': 1433, 'hidden': HIDDEN,
                      'layers': 3, 'mode': 'train fwd+bwd', 'l2_policy': 'working set < L2 (launch-bound; roofline not meaningful)'},
           'e2e': {'value': 10556 * 3 * 2 / (g['e2e_ms'] / 1e3), 'unit': UNIT, 'ms_per_step': g['e2e_ms'],
                   'h2d_bytes_per_step': 2708 * 1433 * 4 + 2 * 10556 * 8, 'd2h_bytes_per_step': 4},
           'gpu_launches': int(g['gpu_launches_per_step'] * steps), 'all': res}
    if 'L3_cpu_oracle' in res:
        c = res['L3_cpu_oracle']
        out['cpu_baseline'] = {'value': c['edges_per_s'], 'unit': UNIT, 'cores': c['cores'], 'kind': 'port',
                               'sample': f'the whole C1 step, median of {steps}', 'cpu': cpu_model()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='gnnb200', choices=['gnnb200', 'reference'])
    ap.add_argument('--scale', type=float, default=1.0, help='fraction of the C5 graph (debug only; 1.0 = BASELINE config)')
    ap.add_argument('--locality', type=float, default=0.0, help='fraction of intra-block edges (0 = uniform random)')
    ap.add_argument('--skew', type=float, default=0.0,
                    help='> 1: power-law endpoints (1.8 ~ ogbn-products: largest hub 2.8e-4 of all edges); 0 = uniform (default)')
    ap.add_argument('--precision', default=None, choices=[None, 'f32', 'tf32', 'tf32x3', 'tf32_fwd3'])
    ap.add_argument('--halo', default=None, choices=[None, 'dense', 'sparse', 'sparse_overlap', 'sparse_pull', 'auto', 'peer', 'peercopy'],
                    help='N > 1: rows exchanged per layer (default dense = all-gather; see gnnb200/partition.py)')
    ap.add_argument('--cpu-sample', type=float, default=None, dest='cpu_sample',
                    help='fraction of the workload per CPU step (default: 1/16 for cpu_baseline; --impl reference picks up to 1/4 by time and memory)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='profiling runs only: skip the end-to-end leg')
    ap.add_argument('--no-generator2', action='store_true', dest='no_generator2', help='skip the secondary run on generator (ii)')
    ap.add_argument('--no-selfcheck', action='store_true', help='profiling runs only: skip the parity check before the timed region')
    ap.add_argument('--no-secondary', action='store_true', help='skip the small-graph configs (C1 backbone, C2 fine-tune step, C3 s4 step)')
    ap.add_argument('--secondary-steps', type=int, default=40, dest='secondary_steps', help='timed steps per small-graph config with --only-secondary')
    ap.add_argument('--only-secondary', action='store_true', help='time only the small-graph configs and print them')
    ap.add_argument('--workload', default='c5', choices=['c5', 'c4', 'c1'],
                    help='c5 = headline (default); c4 = data-parallel s5 pre-training step; c1 = Cora-shaped backbone, GPU and CPU')
    args = ap.parse_args()
    if args.only_secondary:
        import gnnb200  # noqa: F401
        dev = torch.device('cuda', 0)
        sec = small_graph_steps('gnnb200', dev, steps=args.secondary_steps, warmup=min(6, args.secondary_steps))
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            sec['cpu_oracle'] = small_graph_steps('oracle', torch.device('cpu'), steps=4, warmup=1)
        print(json.dumps(sec), flush=True)
        return
    if args.workload in ('c4', 'c1'):
        out = run_c4(args) if args.workload == 'c4' else run_c1(args)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if args.impl == 'reference':
        out = run_reference(args)
    else:
        out = run_product(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
