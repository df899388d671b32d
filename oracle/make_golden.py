"""ORACLE (test infrastructure): generate tests/golden/*.pt by running the UNMODIFIED reference
modules (imported from /root/reference over the pure-PyTorch torch_geometric shim) on small
seeded inputs.  Run in the dev container only:   python -m oracle.make_golden
The GPU box has no /root/reference; tests there compare the CUDA path with these files."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import GOLDEN, oracle_batch, seeded_state_dict  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

import gnnb200  # noqa: E402,F401  (only the synthetic generators; no kernels involved)
from gnnb200 import synthetic  # noqa: E402


def finetune_case(ref):
    torch.manual_seed(0)
    graphs = synthetic.tu_like_graphs('ENZYMES', 12, seed=7)
    batch = oracle_batch(graphs)
    model = ref.finetune_model.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
    model.load_state_dict(seeded_state_dict(model, 1))
    out = {'graphs': graphs, 'weight_seed': 1}
    model.eval()
    with torch.no_grad():
        out['logits_eval'] = model(batch).clone()
    # Train-mode parity needs identical dropout masks on CPU and CUDA, which the two RNGs cannot give;
    # golden train-mode vectors are therefore taken with p = 0 (BN batch statistics active).
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref.gnn.DROPOUT_RATE = 0.0
    logits = model(batch)
    loss = torch.nn.functional.cross_entropy(logits, batch.y)
    loss.backward()
    ref.gnn.DROPOUT_RATE = 0.2
    out['logits_train'] = logits.detach().clone()
    out['loss_train'] = loss.detach().clone()
    out['grads'] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None and (
        k.endswith('eps') or 'layers.0.gin_conv.nn.0' in k or 'layers.4.batch_norm' in k or 'classification_head' in k)}
    return out


def node_cls_case(ref):
    """Cora-shaped but smaller (N=300, F=1433) node classification forward in eval mode."""
    g = synthetic.planetoid_like(300, 600, 1433, seed=3)
    batch = oracle_batch([g])
    model = ref.finetune_model.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'full_finetune')
    model.load_state_dict(seeded_state_dict(model, 2))
    model.eval()
    with torch.no_grad():
        logits = model(batch, message_passing_edges=batch.edge_index)
    return {'graph': g, 'weight_seed': 2, 'logits_eval': logits.clone()}


def pretrain_case(ref):
    """All six task losses of scheme s5 on two small domains, eval mode (dropout off, BN running
    stats), CPU generator seed 11, random.seed(11)."""
    domains = ['MUTAG', 'ENZYMES']
    tasks = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv']
    graphs = {d: synthetic.tu_like_graphs(d, 6, seed=20 + i) for i, d in enumerate(domains)}
    model = ref.pretrain_model.PretrainableGNN(torch.device('cpu'), domains, tasks)
    model.load_state_dict(seeded_state_dict(model, 3))
    model.eval()
    temp = ref.schedulers.TemperatureScheduler(100)
    grl = ref.schedulers.GRLScheduler(10, 10)
    grl.current_step = 80
    T = ref.tasks
    task_objs = {
        'node_feat_mask': T.NodeFeatureMaskingTask(model), 'link_pred': T.LinkPredictionTask(model),
        'node_contrast': T.NodeContrastiveTask(model, temp), 'graph_contrast': T.GraphContrastiveTask(model, temp),
        'graph_prop': T.GraphPropertyPredictionTask(model), 'domain_adv': T.DomainAdversarialTask(model, grl)}
    out = {'graphs': graphs, 'weight_seed': 3, 'domains': domains, 'losses': {}, 'per_domain': {}, 'grads': {}}
    for name, task in task_objs.items():
        gen = torch.Generator().manual_seed(11)
        random.seed(11)
        batches = {d: oracle_batch(graphs[d]) for d in domains}
        model.zero_grad(set_to_none=True)
        loss, per_dom = task.compute_loss(batches, gen)
        loss.backward()
        out['losses'][name] = loss.detach().clone()
        out['per_domain'][name] = {d: v.detach().clone() for d, v in per_dom.items()}
        gsel = {}
        for k, p in model.named_parameters():
            if p.grad is not None and (k.endswith('layers.0.gin_conv.eps') or k.endswith('layers.4.gin_conv.nn.3.bias')
                                       or k == 'mask_token' or k.startswith('heads.link_pred.predictor.mlp.2')):
                gsel[k] = p.grad.clone()
        out['grads'][name] = gsel
    return out


def main():
    ref = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save(finetune_case(ref), os.path.join(GOLDEN, 'finetune_enzymes.pt'))
    torch.save(node_cls_case(ref), os.path.join(GOLDEN, 'finetune_cora_small.pt'))
    torch.save(pretrain_case(ref), os.path.join(GOLDEN, 'pretrain_s5_small.pt'))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == '__main__':
    main()
