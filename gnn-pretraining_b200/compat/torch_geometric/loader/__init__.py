import torch.utils.data

from gnnb200.data import Batch, Data


def _collate(items):
    return Batch.from_data_list(items) if isinstance(items[0], Data) else torch.utils.data.default_collate(items)


class DataLoader(torch.utils.data.DataLoader):
    """torch DataLoader collating lists of Data into a Batch (reference src/data/*_data_loaders.py)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        for k in ('collate_fn', 'follow_batch', 'exclude_keys'):
            kwargs.pop(k, None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=_collate, **kwargs)
