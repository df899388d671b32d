"""ORACLE (test infrastructure, never shipped): torch_geometric.data.{Data,Batch} restated.

Only the members the reference touches on the hot path are provided (SURVEY.md §8c,
App. A.6): .x .edge_index .batch .ptr .y .num_graphs .graph_properties .num_nodes
.num_edges .num_node_features .clone() .to() .to_data_list(), Batch.from_data_list().
Call sites: src/models/pretrain_model.py:69-74, src/pretrain/tasks.py:106-109,154-155,302,
src/pretrain/augmentations.py:18,31,45,64,91,108, src/data/pretrain_data_loaders.py:41.
Upstream PyG absent -> parity UNPINNED.
"""
import copy
from typing import Any, Dict, List

import torch
from torch import Tensor


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, **kwargs).items():
            if v is not None:
                self._store[k] = v

    # attribute plumbing -------------------------------------------------------------
    def __getattr__(self, key: str) -> Any:
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        if key in ('x', 'edge_index', 'edge_attr', 'y', 'batch', 'ptr'):
            return None
        raise AttributeError(key)

    def __setattr__(self, key: str, value: Any) -> None:
        if key.startswith('_'):
            object.__setattr__(self, key, value)
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def keys(self) -> List[str]:
        return list(self._store.keys())

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)

    # derived sizes --------------------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        if '_num_nodes' in self.__dict__:
            return self.__dict__['_num_nodes']
        if 'x' in self._store:
            return self._store['x'].size(0)
        if 'batch' in self._store:
            return self._store['batch'].size(0)
        ei = self._store.get('edge_index')
        return int(ei.max()) + 1 if ei is not None and ei.numel() > 0 else 0

    @property
    def num_edges(self) -> int:
        ei = self._store.get('edge_index')
        return ei.size(1) if ei is not None else 0

    @property
    def num_node_features(self) -> int:
        x = self._store.get('x')
        if x is None:
            return 0
        return 1 if x.dim() == 1 else x.size(-1)

    num_features = num_node_features

    # copies ---------------------------------------------------------------------------
    def clone(self):
        out = copy.copy(self)
        object.__setattr__(out, '_store', {
            k: (v.clone() if isinstance(v, Tensor) else copy.deepcopy(v)) for k, v in self._store.items()})
        return out

    def to(self, device, *args, **kwargs):
        for k, v in list(self._store.items()):
            if isinstance(v, Tensor):
                self._store[k] = v.to(device, *args, **kwargs)
        return self

    def cpu(self):
        return self.to('cpu')

    def __repr__(self) -> str:
        body = ', '.join(f'{k}={list(v.shape) if isinstance(v, Tensor) else v}' for k, v in self._store.items())
        return f'{type(self).__name__}({body})'


def _is_index_key(key: str) -> bool:
    return 'index' in key or key == 'face'


class Batch(Data):
    """App. A.6 collate: x/y/1-D attrs cat dim 0, *index* attrs cat dim -1 with node offsets,
    batch = graph id per node (sorted int64), ptr = [0, cumsum(n_g)]."""

    @classmethod
    def from_data_list(cls, data_list: List[Data]) -> 'Batch':
        out = cls()
        keys = data_list[0].keys()
        sizes = [d.num_nodes for d in data_list]
        offsets = [0]
        for n in sizes:
            offsets.append(offsets[-1] + n)
        slices: Dict[str, List[int]] = {}
        for k in keys:
            vals = [d._store[k] for d in data_list]
            if isinstance(vals[0], Tensor) and vals[0].dim() > 0:
                dim = -1 if _is_index_key(k) else 0
                if _is_index_key(k):
                    vals = [v + off for v, off in zip(vals, offsets[:-1])]
                cuts = [0]
                for v in vals:
                    cuts.append(cuts[-1] + v.size(dim))
                slices[k] = cuts
                out._store[k] = torch.cat(vals, dim=dim)
            elif isinstance(vals[0], Tensor):
                slices[k] = list(range(len(vals) + 1))
                out._store[k] = torch.stack(vals)
            elif isinstance(vals[0], (int, float)):
                slices[k] = list(range(len(vals) + 1))
                out._store[k] = torch.tensor(vals)
            else:
                out._store[k] = vals
        device = next((v.device for v in out._store.values() if isinstance(v, Tensor)), None)
        out._store['batch'] = torch.repeat_interleave(
            torch.arange(len(data_list), device=device), torch.tensor(sizes, device=device))
        out._store['ptr'] = torch.tensor(offsets, dtype=torch.long, device=device)
        out.__dict__['_slices'] = slices
        out.__dict__['_offsets'] = offsets
        out.__dict__['_num_graphs'] = len(data_list)
        return out

    @property
    def num_graphs(self) -> int:
        if '_num_graphs' in self.__dict__:
            return self.__dict__['_num_graphs']
        if 'ptr' in self._store:
            return self._store['ptr'].numel() - 1
        return int(self._store['batch'].max()) + 1

    @property
    def batch_size(self) -> int:
        return self.num_graphs

    def to_data_list(self) -> List[Data]:
        slices, offsets = self.__dict__['_slices'], self.__dict__['_offsets']
        out = []
        for g in range(self.num_graphs):
            d = Data()
            for k, cuts in slices.items():
                v = self._store[k]
                dim = -1 if _is_index_key(k) else 0
                piece = v.narrow(dim, cuts[g], cuts[g + 1] - cuts[g])
                if _is_index_key(k):
                    piece = piece - offsets[g]
                d._store[k] = piece
            out.append(d)
        return out
