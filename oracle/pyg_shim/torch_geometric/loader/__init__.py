"""ORACLE shim: torch_geometric.loader.DataLoader (src/data/pretrain_data_loaders.py:9,
src/data/finetune_data_loaders.py:8) = torch DataLoader collating with Batch.from_data_list."""
import torch.utils.data

from torch_geometric.data import Batch, Data


def _collate(items):
    if isinstance(items[0], Data):
        return Batch.from_data_list(items)
    return torch.utils.data.default_collate(items)


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop('collate_fn', None)
        kwargs.pop('follow_batch', None)
        kwargs.pop('exclude_keys', None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=_collate, **kwargs)
