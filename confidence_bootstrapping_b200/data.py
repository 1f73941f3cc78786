"""Minimal heterogeneous-graph containers with the torch_geometric surface the hot path touches.

The reference feeds `sampling()` / `model(data)` PyG `HeteroData` / `Batch` objects
(utils/sampling.py:78,89-91; datasets/process_mols.py:567-589,448-527).  PyG is not a
dependency of this package: these containers accept the same attribute / key syntax
(`data['ligand'].pos`, `data['ligand', 'ligand'].edge_index`, `data.num_graphs`,
`Batch.from_data_list`, `to_data_list`, `.to(device)`), and every consumer in this
package is duck-typed, so real PyG objects work as well.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, List

import numpy as np
import torch


class _Store:
    """Attribute bag for one node type or one edge type."""

    def __init__(self, key):
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_d", {})

    def __getattr__(self, name):
        d = object.__getattribute__(self, "_d")
        if name in d:
            return d[name]
        if name == "num_nodes":
            for k in ("x", "pos", "batch"):
                if k in d and torch.is_tensor(d[k]):
                    return d[k].shape[0]
            raise AttributeError(name)
        if name == "num_edges":
            if "edge_index" in d:
                return d["edge_index"].shape[1]
            raise AttributeError(name)
        raise AttributeError(f"{object.__getattribute__(self, '_key')!r} store has no attribute {name!r}")

    def __setattr__(self, name, value):
        self._d[name] = value

    def __delattr__(self, name):
        del self._d[name]

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def is_edge(self):
        return isinstance(self._key, tuple)


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_stores", {})
        object.__setattr__(self, "_g", {})

    # -- stores ---------------------------------------------------------------
    def _resolve(self, key):
        if isinstance(key, tuple) and len(key) == 2:
            hits = [k for k in self._stores if isinstance(k, tuple) and k[0] == key[0] and k[2] == key[1]]
            if len(hits) == 1:
                return hits[0]
            if not hits:
                return (key[0], "to", key[1])
            raise KeyError(f"ambiguous edge type {key}")
        return key

    def __getitem__(self, key):
        key = self._resolve(key)
        if key not in self._stores:
            self._stores[key] = _Store(key)
        return self._stores[key]

    @property
    def node_types(self):
        return [k for k in self._stores if not isinstance(k, tuple)]

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    # -- graph-level attributes ----------------------------------------------
    def __getattr__(self, name):
        g = object.__getattribute__(self, "_g")
        if name in g:
            return g[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._g[name] = value

    def __delattr__(self, name):
        g = object.__getattribute__(self, "_g")
        if name not in g:
            raise AttributeError(name)
        del g[name]

    def __contains__(self, name):
        return name in self._g

    @property
    def num_graphs(self):
        return self._g.get("_num_graphs", 1)

    # -- movement -------------------------------------------------------------
    def _apply(self, fn):
        def mv(v):
            if torch.is_tensor(v):
                return fn(v)
            if isinstance(v, dict):
                return {k: mv(u) for k, u in v.items()}
            return v
        for st in self._stores.values():
            for k in list(st.keys()):
                st._d[k] = mv(st._d[k])
        for k in list(self._g.keys()):
            if not k.startswith("_"):
                self._g[k] = mv(self._g[k])
        return self

    def to(self, device, non_blocking=False):
        def mv(t):
            _count_h2d(t, device)
            return t.to(device, non_blocking=non_blocking)
        return self._apply(mv)

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        return copy.deepcopy(self)


def _is_cat_tensor(v):
    return torch.is_tensor(v) and v.dim() >= 1


H2D_BYTES = 0   # bytes moved host -> device by the collate-to-device path and HeteroData.to (bench.py reads it)


def _count_h2d(t, device):
    global H2D_BYTES
    if torch.is_tensor(t) and t.device.type == "cpu" and torch.device(device).type == "cuda":
        H2D_BYTES += t.numel() * t.element_size()


def _replicated(vals):
    """True when every tensor of the list has the same content (the N copies of one complex that
    inference.py / finetune_train.py make with copy.deepcopy share everything but the ligand pose).
    (A thread pool over torch.equal was measured slower than this serial loop: the compare is memory-bound.)"""
    v0 = vals[0]
    if len(vals) < 2 or v0.device.type != "cpu" or v0.numel() == 0:
        return False
    return all(v.shape == v0.shape and v.dtype == v0.dtype and (v.data_ptr() == v0.data_ptr() or torch.equal(v, v0)) for v in vals[1:])


def _cat0(vals, device, rep_out=None):
    """torch.cat(vals, 0), landing on `device` when given.  Replicated host tensors cross the bus once and are
    tiled on the device.  `rep_out` (a list) receives whether the values were identical copies."""
    if device is None:
        if rep_out is not None:
            rep_out.append(False)        # unknown: the host path does not compare
        return torch.cat(vals, 0)
    if _replicated(vals):
        if rep_out is not None:
            rep_out.append(True)
        _count_h2d(vals[0], device)
        d = vals[0].to(device, non_blocking=True)
        return d.repeat((len(vals),) + (1,) * (d.dim() - 1))
    if rep_out is not None:
        rep_out.append(len(vals) < 2)
    out = torch.cat(vals, 0)
    _count_h2d(out, device)
    return out.to(device, non_blocking=True)


class _ReadyEvent:
    """CUDA event recorded when a batch was collated onto the device (its H2D copies are ordered before it).  Lets a consumer
    on ANOTHER stream wait for exactly this batch instead of for everything queued on the collate's stream.  A deep copy of
    the batch is made of later clones, so it does not inherit the event."""

    def __init__(self, ev):
        self.ev = ev

    def __deepcopy__(self, memo):
        return _ReadyEvent(None)


def _static_signature(batch, skip=(("ligand", "pos"),)):
    """(store, attribute, shape, dtype) of every tensor of a collated batch except the ligand pose -- no data is read.
    sampling.reverse_diffusion uses it as the cheap first key of its step-graph cache (a match is then verified element by
    element on the device)."""
    sig = []
    for key, st in batch._stores.items():
        for k, v in st._d.items():
            if torch.is_tensor(v) and (key, k) not in skip:
                sig.append((str(key), k, tuple(v.shape), str(v.dtype)))
    for k, v in batch._g.items():
        if torch.is_tensor(v) and not k.startswith("_"):
            sig.append(("", k, tuple(v.shape), str(v.dtype)))
    return tuple(sorted(sig))


class Batch(HeteroData):
    @classmethod
    def from_data_list(cls, data_list: List[HeteroData], device=None):
        """Collate; with `device` the batch is assembled directly on that device (one transfer per attribute,
        replicated attributes transferred once)."""
        b = cls()
        n = len(data_list)
        first = data_list[0]
        counts: Dict[Any, List[int]] = {}
        for nt in first.node_types:
            counts[nt] = [d[nt].num_nodes for d in data_list]
        offs = {nt: np.concatenate([[0], np.cumsum(c)]) for nt, c in counts.items()}
        slices: Dict[Any, Dict[str, List[int]]] = {}
        rep_flags: Dict[Any, List[bool]] = {}      # per store: were the concatenated attributes identical copies?
        for nt in first.node_types:
            st = b[nt]
            slices[nt] = {}
            rep_flags[nt] = []
            for k in first[nt].keys():
                vals = [d[nt]._d[k] for d in data_list]
                if _is_cat_tensor(vals[0]):
                    st._d[k] = _cat0(vals, device, rep_flags[nt])
                    slices[nt][k] = [0] + list(np.cumsum([v.shape[0] for v in vals]))
                else:
                    st._d[k] = vals
            dev = next((v.device for v in st._d.values() if torch.is_tensor(v)), torch.device("cpu"))
            st._d["batch"] = torch.repeat_interleave(torch.arange(n, device=dev),
                                                     torch.tensor(counts[nt], device=dev))
            st._d["ptr"] = torch.tensor(offs[nt], dtype=torch.long, device=dev)
        for et in first.edge_types:
            st = b[et]
            slices[et] = {}
            rep_flags[et] = []
            for k in first[et].keys():
                vals = [d[et]._d[k] for d in data_list]
                if k == "edge_index":
                    rep_flags[et].append(device is not None and _replicated(vals))
                    if rep_flags[et][-1]:
                        _count_h2d(vals[0], device)
                        shifts = torch.tensor(np.stack([offs[et[0]][:n], offs[et[2]][:n]], 1), dtype=vals[0].dtype).to(device)
                        st._d[k] = (vals[0].to(device).unsqueeze(0) + shifts.unsqueeze(2)).permute(1, 0, 2).reshape(2, -1)
                    else:
                        sh = [torch.tensor([[offs[et[0]][i]], [offs[et[2]][i]]], dtype=v.dtype, device=v.device)
                              for i, v in enumerate(vals)]
                        st._d[k] = torch.cat([v + s for v, s in zip(vals, sh)], 1)
                        if device is not None:
                            _count_h2d(st._d[k], device)
                            st._d[k] = st._d[k].to(device)
                    slices[et][k] = [0] + list(np.cumsum([v.shape[1] for v in vals]))
                elif _is_cat_tensor(vals[0]):
                    st._d[k] = _cat0(vals, device, rep_flags[et])
                    slices[et][k] = [0] + list(np.cumsum([v.shape[0] for v in vals]))
                else:
                    st._d[k] = vals
        for k in first._g.keys():
            if k.startswith("_"):
                continue
            vals = [d._g[k] for d in data_list]
            if _is_cat_tensor(vals[0]):
                b._g[k] = _cat0(vals, device)
            elif torch.is_tensor(vals[0]):
                b._g[k] = torch.stack(vals, 0)
                if device is not None:
                    _count_h2d(b._g[k], device)
                    b._g[k] = b._g[k].to(device)
            else:
                b._g[k] = vals
        # node types whose attributes AND intra-type edges are identical in every graph (the N copies of one complex):
        # models may compute pose-independent quantities of such a type once and tile them
        b._g["_replicated_types"] = sorted(
            nt for nt in first.node_types
            if n > 1 and rep_flags[nt] and all(rep_flags[nt])
            and all(all(rep_flags[et]) and rep_flags[et] for et in first.edge_types if et[0] == nt and et[2] == nt))
        b._g["_num_graphs"] = n
        b._g["_slices"] = slices
        b._g["_offs"] = offs
        # shape-level signature of everything but the ligand pose (sampling.reverse_diffusion: step-graph cache key)
        b._g["_static_sig"] = _static_signature(b) if device is not None else None
        # host-side per-graph counts the models need for buffer caps (node counts are np.diff(_offs[type])): rotatable bonds
        b._g["_n_tor_h"] = None
        if "ligand" in first.node_types and "edge_mask" in first["ligand"].keys():
            ms = [d["ligand"]._d["edge_mask"] for d in data_list]
            if all(torch.is_tensor(m) and m.device.type == "cpu" for m in ms):
                b._g["_n_tor_h"] = [int(m.bool().sum()) for m in ms]
        b._g["_ready"] = _ReadyEvent(None)
        if device is not None and torch.device(device).type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(torch.device(device)))
            b._g["_ready"] = _ReadyEvent(ev)
        return b

    def _rebuild_slices(self):
        """Slices / node offsets re-derived from the `batch` vectors (after an in-place edit such as crop_beyond changed
        the number of nodes or edges per graph).  Edges belong to the graph of their first endpoint."""
        n = self.num_graphs
        counts = {}
        for nt in self.node_types:
            st = self._stores[nt]
            counts[nt] = torch.bincount(st._d["batch"].long().cpu(), minlength=n).numpy() if "batch" in st else np.zeros(n, dtype=np.int64)
        offs = {nt: np.concatenate([[0], np.cumsum(c)]) for nt, c in counts.items()}
        slices = {}
        for key, st in self._stores.items():
            slices[key] = {}
            if isinstance(key, tuple):
                if "edge_index" not in st:
                    continue
                src = st._d["edge_index"][0].long().cpu()
                ec = torch.bincount(torch.bucketize(src, torch.as_tensor(offs[key[0]][1:]), right=True), minlength=n).numpy()
                sl = [0] + list(np.cumsum(ec))
                for k, v in st.items():
                    if k == "edge_index" or (_is_cat_tensor(v) and v.shape[0] == src.shape[0]):
                        slices[key][k] = sl
            else:
                sl = [int(o) for o in offs[key]]
                for k, v in st.items():
                    if k not in ("batch", "ptr") and _is_cat_tensor(v) and v.shape[0] == sl[-1]:
                        slices[key][k] = sl
        self._g["_slices"], self._g["_offs"] = slices, offs
        self._g["_slices_stale"] = False

    def to_data_list(self) -> List[HeteroData]:
        if self._g.get("_slices_stale"):
            self._rebuild_slices()
        n, slices, offs = self.num_graphs, self._g["_slices"], self._g["_offs"]
        out = []
        for i in range(n):
            d = HeteroData()
            for key, st in self._stores.items():
                for k, v in st.items():
                    if k in ("batch", "ptr"):
                        continue
                    sl = slices.get(key, {}).get(k)
                    if sl is not None and torch.is_tensor(v):
                        if k == "edge_index":
                            sh = torch.tensor([[offs[key[0]][i]], [offs[key[2]][i]]], dtype=v.dtype, device=v.device)
                            d[key]._d[k] = v[:, sl[i]: sl[i + 1]] - sh
                        else:
                            d[key]._d[k] = v[sl[i]: sl[i + 1]]
                    elif isinstance(v, list) and len(v) == n:
                        d[key]._d[k] = v[i]
                    # per-step tensors added after collation (node_t, sigma embeddings...) are dropped
            for k, v in self._g.items():
                if k.startswith("_"):
                    continue
                if torch.is_tensor(v) and v.shape[:1] == (n,):
                    d._g[k] = v[i: i + 1]
                elif isinstance(v, list) and len(v) == n:
                    d._g[k] = v[i]
            out.append(d)
        return out


class DataLoader:
    """Sequential, non-shuffling stand-in for torch_geometric.loader.DataLoader (sampling.py:78)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, device=None, **kw):
        assert not shuffle
        self.dataset, self.batch_size, self.device = dataset, batch_size, device

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list([self.dataset[j] for j in range(i, min(i + self.batch_size, len(self.dataset)))],
                                       device=self.device)


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    """Keep edges with both ends in `subset` (bool mask); optionally renumber (utils/utils.py:411-412)."""
    mask = subset if subset.dtype == torch.bool else torch.zeros(num_nodes, dtype=torch.bool).index_fill_(0, subset, True)
    keep = mask[edge_index[0]] & mask[edge_index[1]]
    ei = edge_index[:, keep]
    if relabel_nodes:
        remap = torch.cumsum(mask.long(), 0) - 1
        ei = remap[ei]
    return ei, (edge_attr[keep] if edge_attr is not None else None)
