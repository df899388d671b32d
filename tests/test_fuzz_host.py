"""Randomised comparison of the host-side planners that every collated batch now goes through (gnnb200.augment two-view
planning, gnnb200.utils symmetrise + batched negative sampling) against the oracle's restatement of the reference, over a few
hundred random batches: graphs with 1-2 nodes, graphs without edges (also at the end of the batch), duplicate edges and self
loops, feature widths below the masking threshold, quotas that put some graphs in the `random.sample` branch.  Everything
must be bit-identical, including how far the CPU generator and Python's `random` stream have advanced."""
import random

import numpy as np
import torch

import gnnb200  # noqa: F401
from gnnb200 import augment, data as pdata, utils
from helpers import oracle_batch
from oracle import install_pyg_shim
from oracle import modules as orc

install_pyg_shim()
from torch_geometric.utils import batched_negative_sampling, to_undirected  # noqa: E402


def test_two_view_augmentation_on_random_batches():
    rng = np.random.default_rng(1)
    for case in range(200):
        feats = int(rng.integers(1, 9))
        graphs = []
        for _ in range(int(rng.integers(1, 9))):
            n = int(rng.integers(1, 16))
            e = int(rng.integers(0, 3 * n + 1)) if rng.random() > 0.25 else 0
            graphs.append({'x': torch.randn(n, feats), 'edge_index': torch.from_numpy(rng.integers(0, n, size=(2, e))).long()})
        seed = int(rng.integers(0, 1 << 30))
        ga, gb = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(seed)
        o1, o2, om1, om2 = orc.GraphAugmentor.create_two_views(oracle_batch(graphs), ga)
        batch = pdata.Batch.from_data_list([pdata.Data(**{k: v.clone() for k, v in g.items()}) for g in graphs])
        if case % 2:                                            # alternate between the host-mirror and the read-back path
            for mirror in ('_ptr_host', '_edge_index_host'):
                batch.__dict__.pop(mirror, None)
        p1, p2, pm1, pm2 = augment.GraphAugmentor.create_two_views(batch, gb)
        for o, p in ((o1, p1), (o2, p2)):
            for k in ('x', 'edge_index', 'batch', 'ptr'):
                assert torch.equal(getattr(o, k), getattr(p, k)), (case, k)
        assert len(om1) == len(pm1) and all(torch.equal(a, b) for a, b in zip(om1 + om2, pm1 + pm2)), case
        assert torch.equal(torch.rand(3, generator=ga), torch.rand(3, generator=gb)), case


def test_negative_sampling_on_random_batches():
    rng = np.random.default_rng(0)
    checked = 0
    for case in range(400):
        sizes = rng.integers(1, 14, size=int(rng.integers(1, 7)))
        if rng.random() < 0.15:
            sizes[rng.integers(0, sizes.size)] = int(rng.integers(40, 90))      # large enough for the random branch
        pieces, start = [], 0
        for n in sizes:
            e = int(rng.integers(0, 3 * n + 1)) if rng.random() > 0.2 else 0
            pieces.append(torch.from_numpy(rng.integers(0, n, size=(2, e))).long() + start)
            start += int(n)
        raw = torch.cat(pieces, dim=1)
        if raw.size(1) == 0:
            continue
        batch = torch.repeat_interleave(torch.arange(sizes.size), torch.from_numpy(sizes))
        und_oracle = to_undirected(raw)
        und_host = utils.to_undirected_host(raw.numpy(), int(raw.max()) + 1)
        assert np.array_equal(und_oracle.numpy(), und_host), case
        quota = int(raw.size(1)) if rng.random() < 0.7 else int(rng.integers(1, 50))
        seed = int(rng.integers(0, 1 << 30))
        random.seed(seed)
        want = batched_negative_sampling(und_oracle, batch, quota)
        state_want = random.getstate()
        random.seed(seed)
        got = utils.batched_negative_sampling_host(und_host, sizes.astype(np.int64), quota)
        got = torch.empty(2, 0, dtype=torch.long) if got is None else torch.from_numpy(got)
        assert torch.equal(want, got) and random.getstate() == state_want, (case, sizes.tolist(), quota)
        checked += 1
    assert checked > 300
