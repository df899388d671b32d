"""gnnb200.cuda_graphs: an eval-mode forward on a fixed batch captured into one CUDA graph replays to the eager result — also
after the weights moved (the weight split and the re-pitched operand copies are recomputed inside the graph) and after new
feature values were copied into the batch's `x`."""
import time

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import cuda_graphs, models as prod, synthetic
from helpers import product_batch, seeded_state_dict

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda')


@pytest.mark.parametrize('domain,count', [('ENZYMES', 128), ('Cora_NC', 1)])
def test_captured_eval_forward_follows_weights_and_features(domain, count):
    if domain == 'ENZYMES':
        graphs = synthetic.tu_like_graphs('ENZYMES', count, seed=3)
    else:
        graphs = [synthetic.cora_like(7)]
    m = prod.FinetuneGNN(DEV, domain, 'full_finetune')
    m.load_state_dict(seeded_state_dict(m, 2))
    m.eval()
    batch = product_batch(graphs, DEV)
    with pytest.raises(Exception):
        m.train()
        cuda_graphs.capture_eval(m, batch)
    m.eval()
    with torch.no_grad():
        eager = m(batch).clone()
    cap = cuda_graphs.capture_eval(m, batch)
    assert torch.equal(cap.replay(), eager)                               # same kernels, same order
    with torch.no_grad():
        assert torch.equal(m(batch), eager)                               # the capture left the eager caches intact
        for p in m.parameters():                                          # "optimizer step": in-place, bumps the versions
            p.add_(0.01 * torch.randn_like(p))
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.add_(0.05)
        moved = m(batch).clone()
    assert not torch.equal(moved, eager)
    assert torch.equal(cap.replay(), moved)                               # the graph reads the live weights / running statistics
    with torch.no_grad():
        batch.x.copy_(batch.x * 0.5 + 0.1)                                # new feature values in the same storage
        fresh = m(batch).clone()
    assert torch.equal(cap.replay(), fresh)
    # what it is for: the launch-bound validation pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.no_grad():
            m(batch)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(20):
        cap.replay()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{domain}: eager {1e3 * (t1 - t0) / 20:.3f} ms, graph replay {1e3 * (t2 - t1) / 20:.3f} ms per forward')   # (reported, not asserted)
