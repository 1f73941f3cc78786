"""Graph construction on the device: neighbour search (K1) and edge featurisation (K2).

Replaces the `build_*_conv_graph` methods of the reference models
(models/score_model.py:492-664, models/all_atom_score_model.py:515-664).  Edge lists are kept as
CSR over the AGGREGATION node (`edge_index[0]` in the reference, tensor_layers.py:200-206) because
that is the order the fused convolution consumes; `edge_index()` re-materialises the reference's
[2, E] int64 view for tests and callers.

Nothing here synchronises with the host: edge counts stay on the device (`rowptr[-1]`), output
buffers are sized by an upper bound computed from the (host-known) graph sizes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib


@dataclass
class EdgeList:
    """CSR over aggregation nodes.  row[e] = aggregation node, col[e] = neighbour (both local ids)."""
    rowptr: torch.Tensor            # int32 [n_agg + 1]
    row: torch.Tensor               # int32 [cap]
    col: torch.Tensor               # int32 [cap]
    cap: int
    n_agg: int

    @property
    def n_edges_dev(self) -> torch.Tensor:
        return self.rowptr[self.n_agg:]

    def num_edges(self) -> int:
        """Host read (synchronises) -- tests / debugging only."""
        return int(self.rowptr[self.n_agg].item())

    def edge_index(self) -> torch.Tensor:
        e = self.num_edges()
        return torch.stack([self.row[:e].long(), self.col[:e].long()])


def _scratch(dev):
    return torch.empty(4096, dtype=torch.int32, device=dev)


def radius_edges(x, x_ptr, y, y_batch, r, max_neighbors, cap, cutoff=None, exclude_self=False) -> EdgeList:
    """Forward search: for each query y (aggregation node) the candidates x within r, ascending.

    torch_cluster.radius(x, y, r, batch_x, batch_y, max_num_neighbors) semantics with row 0 = y
    (models/score_model.py:568-573,655).  `cutoff` = per-graph scaling of both point sets (:568-570).
    """
    n_y = y.shape[0]
    dev = y.device
    count = torch.empty(max(n_y, 1), dtype=torch.int32, device=dev)
    rowptr = torch.empty(n_y + 1, dtype=torch.int32, device=dev)
    row = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    col = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    search = n_y > 0 and x.shape[0] > 0          # an empty candidate set (e.g. crop_beyond removed every residue): no edges
    if search:
        _lib.radius_count(x, x_ptr, y, y_batch, cutoff, float(r), max_neighbors, exclude_self, count)
    else:
        count.zero_()
    _lib.exclusive_scan(count[:n_y], rowptr, _scratch(dev))
    if search:
        _lib.radius_fill(x, x_ptr, y, y_batch, cutoff, float(r), max_neighbors, exclude_self, rowptr, row, col)
    return EdgeList(rowptr, row, col, cap, n_y)


def radius_edges_transposed(x, x_batch, y, y_ptr, r, cap, cutoff=None, exclude_self=False,
                            kept: Optional[EdgeList] = None) -> EdgeList:
    """Same edge set as `radius_edges`, grouped by the candidate x (aggregation node = x, neighbour = y).

    `kept` (the forward result) is needed only when the forward search may have truncated."""
    n_x = x.shape[0]
    dev = x.device
    count = torch.empty(max(n_x, 1), dtype=torch.int32, device=dev)
    rowptr = torch.empty(n_x + 1, dtype=torch.int32, device=dev)
    row = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    col = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    kr = kept.rowptr if kept is not None else None
    kc = kept.col if kept is not None else None
    search = n_x > 0 and y.shape[0] > 0
    if search:
        _lib.radius_count_t(x, x_batch, y, y_ptr, cutoff, float(r), exclude_self, kr, kc, count)
    else:
        count.zero_()
    _lib.exclusive_scan(count[:n_x], rowptr, _scratch(dev))
    if search:
        _lib.radius_fill_t(x, x_batch, y, y_ptr, cutoff, float(r), exclude_self, kr, kc, rowptr, row, col)
    return EdgeList(rowptr, row, col, cap, n_x)


def static_edges(edge_index: torch.Tensor, n_agg: int):
    """CSR (by edge_index[0]) of a precomputed edge list (receptor / atom / bond graphs).
    Returns (EdgeList, perm) with perm = stable order of the original edges.  One-off per batch."""
    agg = edge_index[0].to(torch.int64)
    perm = torch.sort(agg, stable=True).indices
    row = agg[perm].to(torch.int32).contiguous()
    col = edge_index[1][perm].to(torch.int32).contiguous()
    counts = count_per_bin(agg, n_agg)       # (torch.bincount reads the maximum back to the host: a device sync)
    rowptr = torch.zeros(n_agg + 1, dtype=torch.int32, device=edge_index.device)
    rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    e = int(edge_index.shape[1])
    if e == 0:
        row = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
        col = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
    return EdgeList(rowptr, row, col, e, n_agg), perm


def identity_edges(ptr: torch.Tensor, n_nbr: int) -> EdgeList:
    """Aggregation node g <- every node of graph g (build_center_conv_graph, score_model.py:635-637)."""
    dev = ptr.device
    n_agg = ptr.numel() - 1
    col = torch.arange(n_nbr, dtype=torch.int32, device=dev)
    row = torch.repeat_interleave(torch.arange(n_agg, dtype=torch.int32, device=dev),
                                  (ptr[1:] - ptr[:-1]).long(), output_size=n_nbr).to(torch.int32)
    return EdgeList(ptr.to(torch.int32).contiguous(), row, col, n_nbr, n_agg)


class EdgeEmbedder:
    """Holds the (views of the) parameters of one `*_edge_embedding` Sequential + its GaussianSmearing
    and runs K2.  Feature layout of the first Linear is given by column offsets."""

    def __init__(self, seq, smearing, n_extra, extra_off, sigma_off, sigma_dim, smear_off):
        self.lin1, self.lin2 = seq[0], seq[3]
        self.smearing = smearing
        self.n_extra, self.extra_off = n_extra, extra_off
        self.sigma_off, self.sigma_dim, self.smear_off = sigma_off, sigma_dim, smear_off

    def __call__(self, edges: EdgeList, pos_agg, pos_nbr, agg_graph, sigma_emb, lmax, sh_sign=1.0, extra=None):
        ns = self.lin2.weight.shape[0]
        dev = pos_agg.device
        W1, b1 = self.lin1.weight, self.lin1.bias
        if self.sigma_dim > 0:
            # sigma-embedding block is constant per graph: fold it into a per-graph bias (plumbing GEMM)
            b1g = torch.addmm(b1, sigma_emb, W1[:, self.sigma_off:self.sigma_off + self.sigma_dim].t()).contiguous()
            stride = ns
        else:
            b1g, stride = b1.contiguous(), 0
        S = (lmax + 1) ** 2
        out_attr = torch.empty((max(edges.cap, 1), ns), dtype=torch.float32, device=dev)
        out_sh = torch.empty((max(edges.cap, 1), S), dtype=torch.float32, device=dev)
        a = _lib.EdgeFeatArgs()
        a.row, a.col = _lib.i32(edges.row, "row"), _lib.i32(edges.col, "col")
        a.n_edges_dev = edges.n_edges_dev.data_ptr()
        a.e_cap = edges.cap
        a.pos_agg, a.pos_nbr = _lib.f32(pos_agg, "pos_agg"), _lib.f32(pos_nbr, "pos_nbr")
        a.sh_sign, a.lmax = float(sh_sign), int(lmax)
        a.agg_graph = _lib.i32(agg_graph, "agg_graph")
        a.b1_graph, a.b1_graph_stride = _lib.f32(b1g, "b1_graph"), stride
        a.extra = _lib.f32(extra, "extra", allow_none=True)
        a.n_extra = self.n_extra
        a.smear_offset = _lib.f32(self.smearing.offset, "smear_offset")
        a.smear_coeff = float(self.smearing.coeff)
        a.n_gauss = int(self.smearing.offset.numel())
        a.W1, a.ldw1 = _lib.f32(W1, "W1"), int(W1.shape[1])
        a.extra_off, a.smear_off = self.extra_off, self.smear_off
        a.W2, a.b2, a.ns = _lib.f32(self.lin2.weight, "W2"), _lib.f32(self.lin2.bias, "b2"), ns
        a.out_attr, a.out_sh = out_attr.data_ptr(), out_sh.data_ptr()
        if edges.cap > 0:
            _lib.edge_featurize(a)
        self._keep = (b1g,)  # stream-ordered allocator keeps this safe; reference held for clarity
        return out_attr, out_sh


def host_counts(data, node_types):
    """Per-graph node counts of `node_types` and rotatable-bond counts from the collate's HOST tables (data.Batch:
    `_offs`, `_n_tor_h`) -- no device read, so building a batch's static tables does not wait for whatever is queued on the
    stream.  None when the batch does not carry them or was edited in place (crop_beyond marks the tables stale)."""
    g = getattr(data, "_g", None)
    if not isinstance(g, dict) or g.get("_slices_stale") or g.get("_offs") is None or g.get("_n_tor_h") is None:
        return None
    offs = g["_offs"]
    try:
        counts = {nt: [int(v) for v in np.diff(np.asarray(offs[nt]))] for nt in node_types}
    except KeyError:
        return None
    n = int(g.get("_num_graphs", 0))
    if any(len(c) != n for c in counts.values()) or len(g["_n_tor_h"]) != n:
        return None
    counts["n_tor"] = [int(v) for v in g["_n_tor_h"]]
    return counts


def masked_columns(edge_index: torch.Tensor, mask: torch.Tensor, n_true=None) -> torch.Tensor:
    """edge_index[:, mask]; with the number of set entries known on the host the selection runs without a device read."""
    if n_true is None:
        return edge_index[:, mask]
    idx = torch.nonzero_static(mask, size=int(n_true)).reshape(-1)
    return edge_index.index_select(1, idx)


def count_per_bin(index: torch.Tensor, n_bins: int) -> torch.Tensor:
    """torch.bincount(index, minlength=n_bins) for indices known to lie in [0, n_bins): no host read."""
    out = torch.zeros(n_bins, dtype=torch.int64, device=index.device)
    if index.numel():
        out.index_add_(0, index.to(torch.int64), torch.ones(index.numel(), dtype=torch.int64, device=index.device))
    return out
