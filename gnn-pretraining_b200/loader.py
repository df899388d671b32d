"""The feeder of the hot path: mini-batches of small graphs cut from datasets that stay RESIDENT in device memory
(SURVEY.md §8f "next" #2, the batching half; reference src/data/pretrain_data_loaders.py:17-43).

The reference keeps every graph as a separate CPU `Data` object; each step it indexes ~32 of them in Python, collates
them with `Batch.from_data_list` (a dozen small `torch.cat`s per domain) and uploads the result tensor by tensor; later
the tasks read `batch.ptr` / `batch.edge_index` back to the host for the augmentation and masking plans.  A whole TU
dataset is a few MB — it belongs in HBM once.  `ResidentDomain` concatenates a domain's graphs a single time (node
features, LOCAL edge ids, labels, graph properties, each with host-side offset tables) and moves them to the device; a
mini-batch is then ONE small packed index upload plus a handful of `index_select`s on the device.  The batch also
carries host mirrors of what the host-side planners need (`_ptr_host`, `_edge_index_host`), so none of them has to
synchronise with the device to read the structure back.

Batches are identical, tensor for tensor, to `Batch.from_data_list([dataset[i] for i in picks])`, and
`BalancedMultiDomainSampler` draws the same `torch.randint` stream as the reference's sampler (:39), so swapping it in
changes no result (tests/test_loader.py pins both on CPU against the reference's classes over the PyG shim).
Everything here is device-agnostic tensor plumbing in front of the kernels.
"""
import os
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch
from torch import Tensor

from .data import Batch, _offset_key

BATCH_SIZE = 32        # src/data/pretrain_data_loaders.py:13


def _tensors_of(graph) -> Dict[str, Tensor]:
    """Named tensors of a Data-like object (gnnb200.data.Data, the PyG shim's Data, or a plain dict)."""
    if isinstance(graph, dict):
        items = graph.items()
    elif hasattr(graph, 'keys') and callable(graph.keys):
        items = ((k, getattr(graph, k)) for k in graph.keys())
    else:
        items = vars(graph).items()
    return {k: v for k, v in items if isinstance(v, Tensor)}


def _segments(starts: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """concat(arange(s, s + c) for s, c in zip(starts, counts)) without a Python loop."""
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    out_starts = np.cumsum(counts) - counts
    return np.repeat(starts - out_starts, counts) + np.arange(total, dtype=np.int64)


class ResidentDomain:
    """All graphs of one domain, collated once and kept on `device`.  `batch_of(ids)` is the device-side equivalent of
    `Batch.from_data_list([graphs[i] for i in ids])`."""

    def __init__(self, graphs: Sequence, device: Optional[torch.device] = None):
        if len(graphs) == 0:
            raise ValueError('ResidentDomain needs at least one graph')
        per_graph = [_tensors_of(g) for g in graphs]
        self.names = [k for k in per_graph[0] if k not in ('batch', 'ptr')]
        self.num_graphs = len(per_graph)
        self.node_count = np.array([t['x'].size(0) for t in per_graph], dtype=np.int64)
        self.node_start = np.cumsum(self.node_count) - self.node_count
        self.store: Dict[str, Tensor] = {}          # name -> concatenation over all graphs (device)
        self.count: Dict[str, np.ndarray] = {}      # name -> per-graph extent along the concatenation axis (host)
        self.start: Dict[str, np.ndarray] = {}
        self.axis: Dict[str, int] = {}
        for name in self.names:
            parts = [t[name] for t in per_graph]
            if parts[0].dim() == 0:
                parts = [p.view(1) for p in parts]
            axis = parts[0].dim() - 1 if _offset_key(name) else 0
            cnt = np.array([p.size(axis) for p in parts], dtype=np.int64)
            self.axis[name], self.count[name], self.start[name] = axis, cnt, np.cumsum(cnt) - cnt
            self.store[name] = torch.cat(parts, dim=axis)
        ei = self.store.get('edge_index')
        self.edge_index_host = ei.cpu().numpy() if ei is not None else None          # LOCAL ids, host mirror
        self.device = torch.device(device) if device is not None else self.store['x'].device
        self._pin = self.device.type == 'cuda'
        for name in self.names:
            self.store[name] = self.store[name].to(self.device)

    def __len__(self) -> int:
        return self.num_graphs

    def batch_of(self, ids: Sequence[int]) -> Batch:
        ids = np.asarray(list(ids), dtype=np.int64)
        b = ids.size
        n_cnt = self.node_count[ids]
        new_start = np.cumsum(n_cnt) - n_cnt                                          # node offset of each picked graph
        ptr = np.concatenate([np.zeros(1, dtype=np.int64), np.cumsum(n_cnt)])
        batch_vec = np.repeat(np.arange(b, dtype=np.int64), n_cnt)
        # one packed int64 upload: [batch | ptr | per attribute: gather index (| edge offset)]
        pieces, layout, cuts = [batch_vec, ptr], {}, {}
        cursor = batch_vec.size + ptr.size
        for name in self.names:
            cnt = self.count[name][ids]
            idx = _segments(self.start[name][ids], cnt)
            cuts[name] = [0] + np.cumsum(cnt).tolist()
            layout[name] = (cursor, idx.size)
            pieces.append(idx)
            cursor += idx.size
            if _offset_key(name):
                pieces.append(np.repeat(new_start, cnt))
                cursor += idx.size
        packed = torch.from_numpy(np.concatenate(pieces))
        if self._pin:
            packed = packed.pin_memory()
        packed = packed.to(self.device, non_blocking=True)
        tensors = {}
        for name in self.names:
            off, m = layout[name]
            idx = packed[off:off + m]
            piece = self.store[name].index_select(self.axis[name], idx)
            if _offset_key(name):
                piece = piece + packed[off + m:off + 2 * m]
            tensors[name] = piece
        out = Batch.from_tensors(tensors.pop('x'), tensors.pop('edge_index') if 'edge_index' in tensors else None,
                                 packed[:batch_vec.size], packed[batch_vec.size:batch_vec.size + ptr.size], **tensors)
        out._cuts, out._starts, out._n_graphs = cuts, ptr.tolist(), int(b)
        # host mirrors for the host-side planners (augmentation, masking, negative sampling): no device read-back
        out._ptr_host = ptr.tolist()
        if self.edge_index_host is not None:
            e_cnt = self.count['edge_index'][ids]
            e_idx = _segments(self.start['edge_index'][ids], e_cnt)
            out._edge_index_host = self.edge_index_host[:, e_idx] + np.repeat(new_start, e_cnt)
        return out


class GraphDataset:
    """reference pretrain_data_loaders.py:17-27: a split (index list) over a list of graphs."""

    def __init__(self, graphs: Sequence, indices) -> None:
        self.graphs = graphs
        self.indices = [int(i) for i in indices]

    def __len__(self) -> int:
        return len(self.indices)

    def __getitem__(self, idx: int):
        return self.graphs[self.indices[idx]]


class BalancedMultiDomainSampler:
    """reference pretrain_data_loaders.py:30-46: every step draws `BATCH_SIZE // #domains` graphs per domain with
    replacement (`torch.randint` on the shared CPU generator, domains in insertion order) and yields
    {domain: Batch}.  The graphs of each domain's split are made resident on `device` at construction."""

    def __init__(self, domain_datasets: Dict[str, GraphDataset], generator: torch.Generator,
                 device: Optional[torch.device] = None, batch_size: int = BATCH_SIZE) -> None:
        self.domain_datasets = domain_datasets
        self.generator = generator
        self.samples_per_domain = batch_size // len(domain_datasets)
        self.num_steps = max(len(d) for d in domain_datasets.values()) // self.samples_per_domain
        self.resident = {name: ResidentDomain([d[i] for i in range(len(d))], device) for name, d in domain_datasets.items()}

    def __iter__(self) -> Iterator[Dict[str, Batch]]:
        for _ in range(self.num_steps):
            yield self.draw()

    def draw(self) -> Dict[str, Batch]:
        """One step's batches (the body of the reference's __iter__ loop)."""
        out = {}
        for name, dataset in self.domain_datasets.items():
            picks = torch.randint(0, len(dataset), (self.samples_per_domain,), generator=self.generator)
            out[name] = self.resident[name].batch_of(picks.tolist())
        return out

    def __len__(self) -> int:
        return self.num_steps


def sequential_batches(dataset: GraphDataset, batch_size: int = BATCH_SIZE, device: Optional[torch.device] = None
                       ) -> List[Batch]:
    """The validation loader's batches (reference create_val_data_loader, :58-68: PyG DataLoader, no shuffle):
    consecutive slices of the split, cut from the resident copy."""
    resident = ResidentDomain([dataset[i] for i in range(len(dataset))], device)
    return [resident.batch_of(range(lo, min(lo + batch_size, len(dataset)))) for lo in range(0, len(dataset), batch_size)]


# ---- fine-tuning loaders (reference src/data/finetune_data_loaders.py) -----------------------------------------------
# The reference wraps a single graph in a Dataset that returns one (data, node, label) / (data, edge, label) tuple per
# ITEM and rebuilds every batch with `torch.tensor([item[..] for item in batch])` — one Python object and one scalar
# read per node or edge.  Its DataLoaders never shuffle (shuffle=False is the default; the generator is unused), so a
# batch is simply a consecutive slice of the split: the classes below keep the graph and the split resident on the device
# and yield those slices (same tuple layout, dtypes and order).


class NodeBatches:
    """create_node_classification_loader (:81-95): yields (data, node_indices [b] int64, labels [b] int64)."""

    def __init__(self, data, indices, batch_size: int, device: Optional[torch.device] = None) -> None:
        self.data = data.to(device) if device is not None else data
        dev = self.data.x.device
        self.indices = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(dev)
        self.labels = self.data.y[self.indices].to(torch.long)
        self.batch_size = len(self.indices) if batch_size == -1 else int(batch_size)
        self.dataset = self

    def __len__(self) -> int:
        return (len(self.indices) + self.batch_size - 1) // self.batch_size if len(self.indices) else 0

    def __iter__(self):
        for lo in range(0, len(self.indices), self.batch_size):
            yield self.data, self.indices[lo:lo + self.batch_size], self.labels[lo:lo + self.batch_size]


class LinkBatches:
    """create_link_prediction_loader (:98-107) over LinkPredictionDataset (:27-52): training batches are positive edges
    with label 1 (negatives are mined per step, finetune.py:186-196); validation / test batches are positives followed by
    the split's fixed negatives.  `.dataset.train_edges` is what the training loop reads (finetune.py:304, :350)."""

    def __init__(self, data, split_edges: Dict[str, Tensor], split: str, batch_size: int,
                 device: Optional[torch.device] = None) -> None:
        self.data = data.to(device) if device is not None else data
        dev = self.data.x.device
        self.split = split
        self.train_edges = split_edges['train_pos'].to(dev)
        if split == 'train':
            self.edges = self.train_edges
            self.labels = torch.ones(self.edges.size(1), device=dev)
        else:
            pos, neg = split_edges[f'{split}_pos'].to(dev), split_edges[f'{split}_neg'].to(dev)
            self.edges = torch.cat([pos, neg], dim=1)
            self.labels = torch.cat([torch.ones(pos.size(1), device=dev), torch.zeros(neg.size(1), device=dev)])
        self.batch_size = int(batch_size)
        self.dataset = self

    def __len__(self) -> int:
        return (self.edges.size(1) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for lo in range(0, self.edges.size(1), self.batch_size):
            yield self.data, self.edges[:, lo:lo + self.batch_size], self.labels[lo:lo + self.batch_size]


# ---- the reference's on-disk format and loader factories (src/data/data_setup.py:63-72, pretrain_data_loaders.py:49-83,
# finetune_data_loaders.py:68-119) -----------------------------------------------------------------------------------------
# data/processed/<domain>/{data.pt = [Data, ...], splits.pt = {split: indices | edges}, graph_properties.pt = [G, 12]}, written
# with torch.save.  The pickles name `torch_geometric.data.Data`: they load with PyG installed or over gnnb200.compat (the
# package's own torch_geometric stand-in); either kind of Data object is accepted by ResidentDomain.  Same function names,
# argument order and iteration order as the reference; `device` / `root` are additions with defaults.

PROCESSED_DIR = Path(os.environ.get('GNNB200_PROCESSED_DIR', Path('data') / 'processed'))


def _domain_dir(domain_name: str, root) -> Path:
    return Path(root if root is not None else PROCESSED_DIR) / domain_name


def save_processed_data(dataset_name: str, data: List, splits: Dict[str, Tensor], graph_properties: Optional[Tensor] = None,
                        root=None) -> None:
    """data_setup.py:66-72."""
    save_dir = _domain_dir(dataset_name, root)
    os.makedirs(save_dir, exist_ok=True)
    torch.save(data, save_dir / 'data.pt')
    torch.save(splits, save_dir / 'splits.pt')
    if graph_properties is not None:
        torch.save(graph_properties, save_dir / 'graph_properties.pt')


def _load_domain(domain_name: str, root, with_properties_for: Optional[str] = None):
    domain_dir = _domain_dir(domain_name, root)
    graphs = torch.load(domain_dir / 'data.pt', weights_only=False)
    splits = torch.load(domain_dir / 'splits.pt', weights_only=False)
    if with_properties_for is not None:                       # pretrain_data_loaders.py:49-53
        properties = torch.load(domain_dir / 'graph_properties.pt', weights_only=False)
        for idx in splits[with_properties_for]:
            graphs[int(idx)].graph_properties = properties[int(idx)]
    return graphs, splits


def create_train_data_loader(domains: List[str], generator: torch.Generator, device: Optional[torch.device] = None,
                             root=None) -> BalancedMultiDomainSampler:
    """pretrain_data_loaders.py:69-83."""
    sets = {}
    for domain in domains:
        graphs, splits = _load_domain(domain, root, 'train')
        sets[domain] = GraphDataset(graphs, splits['train'])
    return BalancedMultiDomainSampler(sets, generator, device=device)


def create_val_data_loader(domain_name: str, generator: torch.Generator, device: Optional[torch.device] = None, root=None
                           ) -> List[Batch]:
    """pretrain_data_loaders.py:56-66 (the reference's PyG DataLoader does not shuffle: consecutive slices)."""
    graphs, splits = _load_domain(domain_name, root, 'val')
    return sequential_batches(GraphDataset(graphs, splits['val']), BATCH_SIZE, device)


def create_graph_classification_loader(domain_name: str, split: str, batch_size: int, generator: torch.Generator,
                                       device: Optional[torch.device] = None, root=None) -> List[Batch]:
    """finetune_data_loaders.py:68-78."""
    graphs, splits = _load_domain(domain_name, root)
    return sequential_batches(GraphDataset(graphs, splits[split]), batch_size, device)


def create_node_classification_loader(domain_name: str, split: str, batch_size: int, generator: torch.Generator,
                                      device: Optional[torch.device] = None, root=None) -> NodeBatches:
    """finetune_data_loaders.py:81-95."""
    graphs, splits = _load_domain(domain_name, root)
    return NodeBatches(graphs[0], splits[split], batch_size, device)


def create_link_prediction_loader(domain_name: str, split: str, batch_size: int, generator: torch.Generator,
                                  device: Optional[torch.device] = None, root=None) -> LinkBatches:
    """finetune_data_loaders.py:98-107."""
    graphs, splits = _load_domain(domain_name, root)
    return LinkBatches(graphs[0], splits, split, batch_size, device)


def create_finetune_data_loader(domain_name: str, split: str, batch_size: int, generator: torch.Generator,
                                device: Optional[torch.device] = None, root=None
                                ) -> Union[List[Batch], NodeBatches, LinkBatches]:
    """finetune_data_loaders.py:110-119: dispatch on the domain's task type."""
    from .models import TASK_TYPES
    make = {'graph_classification': create_graph_classification_loader, 'node_classification': create_node_classification_loader,
            'link_prediction': create_link_prediction_loader}[TASK_TYPES[domain_name]]
    return make(domain_name, split, batch_size, generator, device, root)
