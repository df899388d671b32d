import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch, gnnb200
from gnnb200 import models as prod, synthetic, nn as gnn
from helpers import oracle_batch, product_batch, seeded_state_dict
from oracle import modules as orc
DEV = torch.device('cuda')
d = synthetic.cora_like(seed=42)
for prec in ('f32', 'tf32'):
    gnn.set_default_precision(prec)
    a = orc.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'full_finetune', num_layers=5)
    b = prod.FinetuneGNN(DEV, 'Cora_NC', 'full_finetune', num_layers=5)
    sd = seeded_state_dict(a, 5); a.load_state_dict(sd); b.load_state_dict(sd)
    a.train(); b.train()
    for m in list(a.modules()) + list(b.modules()):
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    prod.DROPOUT_RATE = 0.0; orc.DROPOUT_RATE = 0.0
    ba = oracle_batch([d]); bb = product_batch([d], DEV)
    ha = a.gnn_backbone(a.input_encoder(ba.x), ba.edge_index)
    hb = b.gnn_backbone(b.input_encoder(bb.x), bb.edge_index)
    w = torch.randn(ha.shape, generator=torch.Generator().manual_seed(9))
    (ha * w).sum().backward(); (hb * w.to(DEV)).sum().backward()
    print(prec, 'act rel', float((hb.cpu() - ha).abs().max() / ha.abs().max()))
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for k in pa:
        if pa[k].grad is None: continue
        ga, gb = pa[k].grad.double(), pb[k].grad.double().cpu()
        mx = ga.abs().max().clamp(min=1e-30)
        diff = (ga - gb).abs()
        print(f'  {k:55s} max {float(diff.max()/mx):.2e} fro {float(diff.norm()/ga.norm().clamp(min=1e-30)):.2e} bad(2e-3) {float((diff > 2e-3*mx).double().mean()):.4f} |g|max {float(mx):.2e}')
