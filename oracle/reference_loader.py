"""ORACLE (test infrastructure): import the UNMODIFIED reference modules from /root/reference
over the pure-PyTorch torch_geometric shim.  Works only where /root/reference exists (the
dev container); the GPU box uses oracle/modules.py + tests/golden/ instead."""
import os
import sys

REFERENCE_ROOT = os.environ.get('GNNB200_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'models', 'gnn.py'))


def load_reference():
    """Returns a namespace with the reference's own modules (gnn, heads, pretrain_model,
    finetune_model, tasks, augmentations, schedulers)."""
    if not reference_available():
        raise RuntimeError(f'{REFERENCE_ROOT} is not present')
    from oracle import install_pyg_shim
    install_pyg_shim()
    os.environ.setdefault('WANDB_MODE', 'disabled')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import types
    import src.models.gnn as gnn
    import src.models.heads as heads
    import src.models.pretrain_model as pretrain_model
    import src.models.finetune_model as finetune_model
    import src.pretrain.tasks as tasks
    import src.pretrain.augmentations as augmentations
    import src.pretrain.schedulers as schedulers
    import src.pretrain.gradient_surgery as gradient_surgery
    return types.SimpleNamespace(gnn=gnn, heads=heads, pretrain_model=pretrain_model,
                                 finetune_model=finetune_model, tasks=tasks,
                                 augmentations=augmentations, schedulers=schedulers,
                                 gradient_surgery=gradient_surgery)
