"""Known answers the reference itself publishes: trainable-parameter counts from
analysis/results/experiment_results.csv and the state-dict key layout of the shipped checkpoint
header.  Checked for the oracle restatement AND the product modules (CPU construction only)."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import models as prod
from oracle import modules as orc

KAT = {  # (domain, strategy) -> trainable parameters
    ('ENZYMES', 'full_finetune'): 1355915, ('ENZYMES', 'linear_probe'): 33670,
    ('PTC_MR', 'full_finetune'): 1360775, ('PTC_MR', 'linear_probe'): 38530,
    ('Cora_NC', 'full_finetune'): 1691660, ('Cora_NC', 'linear_probe'): 369415,
    ('CiteSeer_NC', 'full_finetune'): 2272523, ('CiteSeer_NC', 'linear_probe'): 950278,
    ('Cora_LP', 'full_finetune'): 1886982, ('Cora_LP', 'linear_probe'): 564737,
    ('CiteSeer_LP', 'full_finetune'): 2468102, ('CiteSeer_LP', 'linear_probe'): 1145857,
}
S5 = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv']


@pytest.mark.parametrize('impl', [orc, prod], ids=['oracle', 'product'])
@pytest.mark.parametrize('key', sorted(KAT))
def test_trainable_parameter_counts(impl, key):
    m = impl.FinetuneGNN(torch.device('cpu'), key[0], key[1])
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == KAT[key]


@pytest.mark.parametrize('impl', [orc, prod], ids=['oracle', 'product'])
def test_pretrain_model_size(impl):
    doms = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']
    assert sum(p.numel() for p in impl.PretrainableGNN(torch.device('cpu'), doms, S5).parameters()) == 3702714
    assert sum(p.numel() for p in impl.PretrainableGNN(torch.device('cpu'), doms, S5[:-1]).parameters()) == 3669302


@pytest.mark.parametrize('impl', [orc, prod], ids=['oracle', 'product'])
def test_state_dict_layout(impl):
    sd = impl.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'linear_probe').state_dict()
    keys = list(sd)
    assert len(keys) == 84
    assert keys[:5] == ['input_encoder.linear.weight', 'input_encoder.linear.bias', 'input_encoder.batch_norm.weight',
                        'input_encoder.batch_norm.bias', 'input_encoder.batch_norm.running_mean']
    layer0 = [k for k in keys if k.startswith('gnn_backbone.layers.0.')]
    assert layer0 == ['gnn_backbone.layers.0.gin_conv.eps'] + \
        [f'gnn_backbone.layers.0.gin_conv.nn.0.{s}' for s in ('weight', 'bias')] + \
        [f'gnn_backbone.layers.0.gin_conv.nn.1.{s}' for s in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked')] + \
        [f'gnn_backbone.layers.0.gin_conv.nn.3.{s}' for s in ('weight', 'bias')] + \
        [f'gnn_backbone.layers.0.batch_norm.{s}' for s in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked')]
    assert keys[-2:] == ['classification_head.mlp.0.weight', 'classification_head.mlp.0.bias']
    assert sd['gnn_backbone.layers.3.gin_conv.nn.0.weight'].shape == (512, 256)
    assert sd['gnn_backbone.layers.3.gin_conv.eps'].shape == (1,)
    # the parameter-free child the checkpoint header lists
    m = impl.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'linear_probe')
    assert hasattr(m.gnn_backbone.layers[0].gin_conv, 'aggr_module')


def test_product_and_oracle_state_dicts_interchange():
    a = orc.PretrainableGNN(torch.device('cpu'), ['MUTAG', 'ENZYMES'], S5)
    b = prod.PretrainableGNN(torch.device('cpu'), ['MUTAG', 'ENZYMES'], S5)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict(), strict=True)


def test_layer_count_is_a_parameter():
    m = prod.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'full_finetune', num_layers=3)
    assert len(m.gnn_backbone.layers) == 3
