"""Aggregation forward/backward at BASELINE config 5's FULL size on the device, through the checker that
tests/test_full_scale_properties.py validates on CPU.  (Named to sort last: the driver runs the GPU suite with -x.)"""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import synthetic
from test_full_scale_properties import check_aggregation_properties


@pytest.mark.gpu
def test_aggregation_full_c5_scale():
    from gnnb200 import _lib as L, ops
    from gnnb200.graph import Graph
    dev = torch.device('cuda')
    cache = {}

    def graph(ei):
        if cache.get('ei') is not ei:
            cache['ei'], cache['g'] = ei, Graph(ei, synthetic.C5_NODES)
        return cache['g']

    def forward(x, eps, ei):
        g = graph(ei)
        return ops._aggregate_raw(x, g.rowptr, g.col, L.AGG_SUM, x, eps, None)

    def backward(gy, eps, ei):
        g = graph(ei)
        return ops._aggregate_raw(gy, g.rowptr_t, g.col_t, L.AGG_SUM, gy, eps, None)

    check_aggregation_properties(forward, backward, dev, synthetic.C5_NODES, synthetic.C5_EDGES, 256)
