"""The reference's OWN files import and construct over gnnb200.compat (the torch_geometric stand-in):
container-only (needs /root/reference); runs in a subprocess because the oracle shim and the stand-in
both answer to the name `torch_geometric`.  No kernels are launched (CPU box): this checks the drop-in
surface — names, signatures, parameter counts and state-dict keys."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
os.environ.setdefault('WANDB_MODE', 'disabled')
import gnnb200
from gnnb200 import compat
compat.install()
sys.path.insert(0, %(ref)r)
import torch
import torch_geometric
assert torch_geometric.__gnnb200__
from src.models.finetune_model import FinetuneGNN
from src.models.pretrain_model import PretrainableGNN
import src.pretrain.tasks as tasks
m = FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 1355915
assert type(m.gnn_backbone.layers[0].gin_conv).__module__.startswith('gnnb200')
assert len(FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'linear_probe').state_dict()) == 84
p = PretrainableGNN(torch.device('cpu'), ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES'],
                    ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv'])
assert sum(q.numel() for q in p.parameters()) == 3702714
assert tasks.global_mean_pool.__module__.startswith('gnnb200')
assert tasks.to_undirected.__module__.startswith('gnnb200')
print('compat ok')
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'src', 'models')), reason='/root/reference not present')
def test_reference_files_import_over_the_stand_in():
    out = subprocess.run([sys.executable, '-c', SCRIPT % {'root': ROOT, 'ref': REF}], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'compat ok' in out.stdout
