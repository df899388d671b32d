"""Host-side graph containers the hot path consumes (layout of SURVEY.md App. A.6).

These mirror the members of PyG's ``Data`` / ``Batch`` that the reference touches
(src/models/pretrain_model.py:69-74, src/pretrain/tasks.py:106-109,154-155,302,
src/pretrain/augmentations.py:18-109) so that the reference's call sites keep working when
``gnnb200.compat`` stands in for ``torch_geometric``.  Collation itself is host work that
sits *before* the hot path; only the tensors it produces (x, edge_index, batch, ptr) reach
the CUDA kernels.
"""
from typing import Dict, Iterable, List, Optional

import torch
from torch import Tensor

_NODE_OFFSET_KEYS = ('edge_index', 'face')
_STRUCTURE_KEYS = ('edge_index', 'ptr', 'batch')
_HOST_MIRRORS = ('_ptr_host', '_edge_index_host', '_common_rows_host')


def _offset_key(name: str) -> bool:
    return name in _NODE_OFFSET_KEYS or 'index' in name


class Data:
    """One graph: a bag of named tensors with ``x`` [n, F] and ``edge_index`` [2, e] (int64)."""

    def __init__(self, x: Optional[Tensor] = None, edge_index: Optional[Tensor] = None,
                 y: Optional[Tensor] = None, **extra):
        self.__dict__['_t'] = {}
        for name, value in (('x', x), ('edge_index', edge_index), ('y', y), *extra.items()):
            if value is not None:
                self._t[name] = value

    def __getattr__(self, name):
        t = self.__dict__['_t']
        if name in t:
            return t[name]
        if name in ('x', 'edge_index', 'y', 'batch', 'ptr', 'edge_attr'):
            return None
        raise AttributeError(name)

    def _structure_versions(self):
        return tuple((k, self._t[k]._version) for k in _STRUCTURE_KEYS if isinstance(self._t.get(k), Tensor))

    def host_mirror(self, name: str):
        """A host mirror of the structure (`_ptr_host`, `_edge_index_host`, `_common_rows_host`) if it still describes the
        tensors: an in-place edit of edge_index / ptr / batch since the mirror was taken (their `_version` moved) drops all
        mirrors, the same rule gnnb200.graph.graph_of applies to its cached CSR."""
        value = self.__dict__.get(name)
        if value is None:
            return None
        if self.__dict__.get('_mirror_versions') != self._structure_versions():
            for mirror in _HOST_MIRRORS + ('_mirror_versions',):
                self.__dict__.pop(mirror, None)
            return None
        return value

    def __setattr__(self, name, value):
        if name.startswith('_'):
            self.__dict__[name] = value
            if name in _HOST_MIRRORS:
                self.__dict__['_mirror_versions'] = self._structure_versions()
        else:
            if name in _STRUCTURE_KEYS:            # host mirrors of the structure (Batch.from_data_list, gnnb200.loader) are stale now
                for mirror in _HOST_MIRRORS:
                    self.__dict__.pop(mirror, None)
            if value is None:
                self._t.pop(name, None)
            else:
                self._t[name] = value

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        self.__dict__.update(state)

    def keys(self) -> List[str]:
        return list(self._t)

    @property
    def num_nodes(self) -> int:
        if 'x' in self._t:
            return self._t['x'].size(0)
        if 'batch' in self._t:
            return self._t['batch'].numel()
        ei = self._t.get('edge_index')
        return 0 if ei is None or ei.numel() == 0 else int(ei.max()) + 1

    @property
    def num_edges(self) -> int:
        ei = self._t.get('edge_index')
        return 0 if ei is None else ei.size(1)

    @property
    def num_node_features(self) -> int:
        x = self._t.get('x')
        return 0 if x is None else (1 if x.dim() == 1 else x.size(-1))

    def clone(self):
        out = type(self).__new__(type(self))
        out.__dict__.update({k: v for k, v in self.__dict__.items() if k != '_t'})
        out.__dict__['_t'] = {k: (v.clone() if isinstance(v, Tensor) else v) for k, v in self._t.items()}
        if '_mirror_versions' in out.__dict__:
            out.__dict__['_mirror_versions'] = out._structure_versions() if self.host_mirror('_ptr_host') is not None or \
                self.host_mirror('_edge_index_host') is not None else None
        return out

    def to(self, device, non_blocking: bool = False):
        valid = '_mirror_versions' in self.__dict__ and self.__dict__['_mirror_versions'] == self._structure_versions()
        for k, v in self._t.items():
            if isinstance(v, Tensor):
                self._t[k] = v.to(device, non_blocking=non_blocking)
        if '_mirror_versions' in self.__dict__:                    # the copies start their own version counters
            self.__dict__['_mirror_versions'] = self._structure_versions() if valid else None
        return self

    def pin_memory(self):
        for k, v in self._t.items():
            if isinstance(v, Tensor) and not v.is_cuda:
                self._t[k] = v.pin_memory()
        return self

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={tuple(v.shape)}' for k, v in self._t.items() if isinstance(v, Tensor))})"


def host_mirror(obj, name: str):
    """`obj.host_mirror(name)` for gnnb200 containers, plain attribute lookup for anything else."""
    fn = getattr(obj, 'host_mirror', None)
    return fn(name) if fn is not None else getattr(obj, name, None)


class Batch(Data):
    """Disjoint union of graphs: node tensors concatenated, edge indices shifted by the running
    node count, ``batch`` = sorted graph id per node, ``ptr`` = node offsets [B+1]."""

    @classmethod
    def from_data_list(cls, graphs: Iterable[Data]) -> 'Batch':
        graphs = list(graphs)
        counts = [g.num_nodes for g in graphs]
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        out = cls()
        cuts: Dict[str, List[int]] = {}
        for name in graphs[0].keys():
            parts = [g._t[name] for g in graphs]
            if not isinstance(parts[0], Tensor):
                parts = [torch.as_tensor(p) for p in parts]
            if parts[0].dim() == 0:
                parts = [p.view(1) for p in parts]
            axis = parts[0].dim() - 1 if _offset_key(name) else 0
            if _offset_key(name):
                parts = [p + s for p, s in zip(parts, starts)]
            run = [0]
            for p in parts:
                run.append(run[-1] + p.size(axis))
            cuts[name] = run
            out._t[name] = torch.cat(parts, dim=axis)
        dev = out._t['x'].device if 'x' in out._t else None
        out._t['batch'] = torch.repeat_interleave(torch.arange(len(graphs), device=dev),
                                                  torch.tensor(counts, device=dev))
        out._t['ptr'] = torch.tensor(starts, dtype=torch.long, device=dev)
        out._cuts, out._starts, out._n_graphs = cuts, starts, len(graphs)
        # Collation happens on the host: keep what the host-side planners (augmentation, negative sampling, node masking)
        # read — graph boundaries and the edge list — as host mirrors, so that after .to(device) they never have to
        # read the structure back (one sync + one device->host copy per task and domain otherwise).  Assigning a new
        # edge_index / ptr / batch drops them (Data.__setattr__).
        ei = out._t.get('edge_index')
        if ei is not None and not ei.is_cuda and ei.dtype == torch.long:
            out._ptr_host = list(starts)
            out._edge_index_host = ei.numpy()
        return out

    @classmethod
    def from_tensors(cls, x: Tensor, edge_index: Tensor, batch: Optional[Tensor] = None,
                     ptr: Optional[Tensor] = None, **extra) -> 'Batch':
        """Wrap already-collated tensors (single big graph when batch/ptr are omitted)."""
        out = cls(x=x, edge_index=edge_index, **extra)
        n = x.size(0)
        if batch is None:
            batch = torch.zeros(n, dtype=torch.long, device=x.device)
            ptr = torch.tensor([0, n], dtype=torch.long, device=x.device)
        out._t['batch'] = batch
        if ptr is not None:
            out._t['ptr'] = ptr
            out._n_graphs = ptr.numel() - 1
        return out

    @property
    def num_graphs(self) -> int:
        if '_n_graphs' in self.__dict__:
            return self.__dict__['_n_graphs']
        if 'ptr' in self._t:
            return self._t['ptr'].numel() - 1
        return int(self._t['batch'].max()) + 1

    def to_data_list(self) -> List[Data]:
        if '_cuts' not in self.__dict__:
            raise RuntimeError('to_data_list() needs a Batch built by from_data_list()')
        out = []
        for g in range(self.num_graphs):
            d = Data()
            for name, run in self._cuts.items():
                v = self._t[name]
                axis = v.dim() - 1 if _offset_key(name) else 0
                piece = v.narrow(axis, run[g], run[g + 1] - run[g])
                d._t[name] = piece - self._starts[g] if _offset_key(name) else piece
            out.append(d)
        return out
