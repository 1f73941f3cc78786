"""TensorProductConvLayer on the fused K3 kernel, with the reference's parameter names.

Mirrors models/tensor_layers.py:120-217 (`TensorProductConvLayer`), models/layers.py:8-15 (`FCBlock`)
and e3nn.nn.BatchNorm (parameters `batch_norm.{weight,bias,running_mean,running_var}`), so a
reference state_dict loads with strict=True (FasterTensorProduct has no parameters; e3nn's FCTP
only carries constant buffers, which `TensorProductScoreModel` filters out on load).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .graph import EdgeList, static_edges
from .irreps import (TPProgram, faster_tp_program, fctp_program, get_irrep_seq, irreps_dim, irreps_str,
                     parse_irreps, sh_irreps, transform_plan)

ACTIVATIONS = {"relu": nn.ReLU, "silu": nn.SiLU}
# accumulate kernel of K3: 3 = tcgen05 3xTF32 UMMA with the TRANSPOSED TMEM accumulator (default; layers with more than 240
# f-rows -- the lmax-2 confidence layers -- automatically use 2); 2 = tcgen05 with the row-major accumulator;
# 1 = fp32 FFMA register tiles (kept for shapes outside the UMMA tile limits and for A/B measurements)
ACCUM_MODE = int(os.environ.get("CB200_ACCUM_MODE", "4"))
# bumped whenever a layer drops tensors derived from its weights: captured CUDA graphs hold pointers to those tensors and
# are only replayed under the epoch they were captured in (sampling._graph_cache)
CACHE_EPOCH = 0


_WORKSPACES = {}    # device -> grow-only K3 workspace of the eager launches


def _workspace(n_floats, dev):
    """Accumulator workspace of one K3 call.  Eager launches share ONE grow-only buffer per device: consecutive calls on a
    stream use it one after the other, and the multi-GB cudaMalloc / cudaFree a differently sized request can trigger in the
    caching allocator (the confidence leg's shapes change with every crop_beyond: single 250-380 ms steps among 165 ms ones)
    disappears.  While a CUDA graph is being captured the buffer comes from the graph's memory pool instead, so that the
    graph owns what it points to."""
    if DEBUG_KEEP_WORKSPACE is not None:
        return torch.empty(n_floats, dtype=torch.float32, device=dev)
    if torch.cuda.is_current_stream_capturing():
        # graph pool: 25 % head-room in 256 MB steps, so that the block freed by the previous complex's graph fits the next
        # complex's slightly different request instead of forcing a new multi-GB cudaMalloc (which stalls the device)
        step = 64 << 20
        n_alloc = min(-(-int(n_floats * 1.25) // step) * step, max(WORKSPACE_BYTES // 4, n_floats))
        return torch.empty(max(n_alloc, n_floats), dtype=torch.float32, device=dev)[:n_floats]
    key = (torch.device(dev).index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < n_floats:
        # 2x head-room (within the cap): the filtering leg's request follows the cropped system size, which varies 2x between
        # batches, and a multi-GB cudaMalloc in the middle of a run stalls the device (single 250-380 ms steps)
        cap = max(WORKSPACE_BYTES // 4, n_floats)
        grow = min(2 * n_floats, cap) if buf is None else min(max(n_floats, 2 * buf.numel()), cap)
        _WORKSPACES[key] = buf = None          # release the old block to the allocator before asking for the larger one
        _WORKSPACES[key] = buf = torch.empty(grow, dtype=torch.float32, device=dev)
    return buf[:n_floats]


def _bump_epoch():
    global CACHE_EPOCH
    CACHE_EPOCH += 1

FOLD_E_POST = True            # fold W1e.e_post[graph] into the node projection on the host (tests switch it off to cover the kernel path)
TC_TRANSFORM = os.environ.get("CB200_TC_TRANSFORM", "0") != "0"   # experimental tcgen05 transform kernel (measured 3x slower than the FFMA kernel: DESIGN.md)
DEBUG_KEEP_WORKSPACE = None   # tests may set this to a list to inspect the K3 accumulators
WORKSPACE_BYTES = int(os.environ.get("CB200_WORKSPACE_MB", str(16 << 10))) << 20  # cap on the K3 accumulator workspace (180 GB of HBM3e per GPU); larger layers run in node chunks


def FCBlock(in_dim, hidden_dim, out_dim, layers, dropout, activation="relu"):
    """Linear-act-Dropout-(...)-Linear; Sequential indices 0/3 hold the Linears (layers.py:8-15)."""
    act = ACTIVATIONS[activation]
    assert layers >= 2
    mods = [nn.Linear(in_dim, hidden_dim), act(), nn.Dropout(dropout)]
    for _ in range(layers - 2):
        mods += [nn.Linear(hidden_dim, hidden_dim), act(), nn.Dropout(dropout)]
    mods += [nn.Linear(hidden_dim, out_dim)]
    return nn.Sequential(*mods)


def irrep_to_size(irrep: str) -> int:
    return irreps_dim(irrep)


class EquivariantBatchNorm(nn.Module):
    """Parameter container + eval-mode affine of e3nn.nn.BatchNorm(irreps):
    per irrep block, 0e channels subtract running_mean and add bias; every channel is scaled by
    weight / sqrt(running_var + eps); 0o and l>0 channels have no mean / bias."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1):
        super().__init__()
        self.irreps = parse_irreps(irreps)
        self.eps, self.momentum = eps, momentum
        n_feat = sum(m for m, _, _ in self.irreps)
        n_scalar = sum(m for m, l, p in self.irreps if l == 0 and p == 1)
        self.register_buffer("running_mean", torch.zeros(n_scalar))
        self.register_buffer("running_var", torch.ones(n_feat))
        self.weight = nn.Parameter(torch.ones(n_feat))
        self.bias = nn.Parameter(torch.zeros(n_scalar))
        dims, is_s = [], []
        for m, l, p in self.irreps:
            dims += [2 * l + 1] * m
            is_s += [l == 0 and p == 1] * m
        # static index tables (no boolean masks / repeat_interleave at run time: those synchronise the host)
        self.register_buffer("_scalar_idx", torch.tensor([i for i, f in enumerate(is_s) if f], dtype=torch.long), persistent=False)
        self.register_buffer("_feat_of_chan", torch.tensor([i for i, d in enumerate(dims) for _ in range(d)], dtype=torch.long),
                             persistent=False)
        self._affine_cache = None

    def invalidate_caches(self):
        """Forget the folded (scale, shift).  The cache key (data_ptr, _version) does not see writes through
        `param.data.copy_()` (the reference's ExponentialMovingAverage.copy_to / restore, utils/utils.py:353-392)."""
        self._affine_cache = None
        _bump_epoch()

    def _apply(self, fn, *a, **kw):
        self.invalidate_caches()
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self.invalidate_caches()
        return super()._load_from_state_dict(*a, **kw)

    def affine(self):
        """(scale[d], shift[d]) such that eval-mode BN(x) = x * scale + shift, per feature channel.
        Cached until a parameter or running statistic changes."""
        key = tuple((t.data_ptr(), t._version) for t in (self.weight, self.bias, self.running_mean, self.running_var))
        if self._affine_cache is None or self._affine_cache[0] != key:
            with torch.no_grad():
                scale_c = self.weight * torch.rsqrt(self.running_var + self.eps)
                shift_c = torch.zeros_like(scale_c)
                shift_c.index_copy_(0, self._scalar_idx, self.bias - self.running_mean * scale_c.index_select(0, self._scalar_idx))
                self._affine_cache = (key, scale_c.index_select(0, self._feat_of_chan).contiguous(),
                                      shift_c.index_select(0, self._feat_of_chan).contiguous())
        return self._affine_cache[1], self._affine_cache[2]


@dataclass
class Segment:
    """One edge list handled by one radial MLP (`fc[group]`) inside a layer call."""
    edges: EdgeList
    e_attr: torch.Tensor                 # [cap, ne]
    sh: torch.Tensor                     # [cap, S]
    group: int = 0
    n0: int = 0                          # aggregation-node range [n0, n1) in the output rows
    n1: int = 0
    col_off: int = 0                     # offset of the neighbour node type inside x
    e_post: Optional[torch.Tensor] = None  # [B, ne]
    gate: Optional[object] = None          # dead-output pruning: an EdgeList (nodes without an edge in it are skipped) or a
                                           # uint8 keep mask over the aggregation nodes
    slot: Optional[int] = None           # segments sharing a slot share (n0, n1, group) and one accumulator


def _tf32_round(x):
    """fp32 -> nearest value with a 10-bit mantissa (what tcgen05 kind::tf32 keeps), as fp32."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


class _DeviceProgram:
    def __init__(self, prog: TPProgram, device):
        self.prog = prog
        self.rows = torch.from_numpy(prog.rows.view(np.uint8).copy()).to(device)
        self.terms = torch.from_numpy(prog.terms.view(np.uint8).copy()).to(device)
        self.runs = torch.from_numpy(prog.runs.view(np.uint8).copy()).to(device)
        assert int(prog.runs["mul"].max()) <= 32, "transform kernel stages at most 32 weight rows per row group"


class TensorProductConvLayer(nn.Module):
    def __init__(self, in_irreps, sh_irreps, out_irreps, n_edge_features, residual=True, batch_norm=True, dropout=0.0,
                 hidden_features=None, faster=False, edge_groups=1, tp_weights_layers=2, activation="relu",
                 depthwise=False):
        super().__init__()
        if depthwise:
            raise NotImplementedError("depthwise convolutions are not used by the shipped configurations")
        if tp_weights_layers != 2:
            raise NotImplementedError("the fused kernel implements the 2-layer radial MLP of the shipped configurations")
        if activation != "relu":
            raise NotImplementedError("radial MLP activation must be relu")
        self.in_irreps, self.out_irreps = irreps_str(in_irreps), irreps_str(out_irreps)
        self.sh_irreps = irreps_str(sh_irreps)
        self.residual, self.edge_groups, self.faster = residual, edge_groups, faster
        self.out_size = irreps_dim(out_irreps)
        self.n_edge_features = n_edge_features
        hidden_features = hidden_features or n_edge_features
        self.hidden_features = hidden_features
        if faster:
            assert self.sh_irreps == "1x0e + 1x1o", "sh_irreps don't look like 1st order spherical harmonics"
            self.program = faster_tp_program(self.in_irreps, self.out_irreps)
        else:
            self.program = fctp_program(self.in_irreps, self.sh_irreps, self.out_irreps)
        self.weight_numel = self.program.weight_numel
        if edge_groups == 1:
            self.fc = FCBlock(n_edge_features, hidden_features, self.weight_numel, tp_weights_layers, dropout, activation)
        else:
            self.fc = nn.ModuleList([FCBlock(n_edge_features, hidden_features, self.weight_numel, tp_weights_layers,
                                             dropout, activation) for _ in range(edge_groups)])
        self.batch_norm = EquivariantBatchNorm(out_irreps) if batch_norm else None
        self._dev_prog = None
        self._plan = None          # tensor-core transform plan (irreps.transform_plan) + its device tables
        self._w2t_cache = {}
        self._w2a_cache = {}
        self._param_cache = {}
        self._proj_cache = {}

    # ------------------------------------------------------------------ derived-weight caches
    def invalidate_caches(self):
        """Drop W2a / projection / BatchNorm-affine tensors derived from the parameters.  They are keyed on
        (data_ptr, _version), which misses in-place writes through `param.data` (EMA copy_to / restore, utils/utils.py:353-392):
        `sampling()` calls this on the model at the start of every call, and `.to()` / `load_state_dict` / `train()` do too."""
        self._w2a_cache.clear()
        self._w2t_cache.clear()
        self._proj_cache.clear()
        self._param_cache.clear()
        _bump_epoch()
        if self.batch_norm is not None:
            self.batch_norm.invalidate_caches()

    def _apply(self, fn, *a, **kw):
        self._w2a_cache.clear(); self._w2t_cache.clear(); self._proj_cache.clear(); self._param_cache.clear()
        _bump_epoch()
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self._w2a_cache.clear(); self._w2t_cache.clear(); self._proj_cache.clear(); self._param_cache.clear()
        _bump_epoch()
        return super()._load_from_state_dict(*a, **kw)

    def train(self, mode=True):
        self.invalidate_caches()
        return super().train(mode)

    # ------------------------------------------------------------------ helpers
    def _fc(self, g):
        return self.fc if self.edge_groups == 1 else self.fc[g]

    def _params(self, g):
        """(W1, b1, W2, b2) Parameters of the radial MLP of group g (looked up once: nn.Module attribute access is slow
        and .to()/.cuda()/load_state_dict keep the Parameter objects)."""
        hit = self._param_cache.get(g)
        if hit is None:
            fc = self._fc(g)
            hit = self._param_cache[g] = (fc[0].weight, fc[0].bias, fc[3].weight, fc[3].bias)
        return hit

    def _w2a(self, g):
        """Second Linear of the radial MLP with its bias folded in: [weight_numel, H+4] rows (W2[w], b2[w], 0, 0, 0),
        the layout the transform kernel streams with contiguous bulk copies.  Rebuilt when the parameters change."""
        _, _, W2, b2 = self._params(g)
        key = (W2.data_ptr(), W2._version, b2.data_ptr(), b2._version)
        hit = self._w2a_cache.get(g)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                t = torch.zeros(W2.shape[0], W2.shape[1] + 4, device=W2.device, dtype=torch.float32)
                t[:, :W2.shape[1]] = W2
                t[:, W2.shape[1]] = b2
            hit = (key, t)
            self._w2a_cache[g] = hit
        return hit[1]

    def _tc_plan(self, device):
        """(TransformPlan, chains tensor, blocks tensor, max staging bytes) or None when the program does not fit the kernel."""
        if self._plan is None or (self._plan is not False and self._plan[1].device != device):
            plan = transform_plan(self.program, self.hidden_features)
            if plan is None:
                self._plan = False
            else:
                ha = self.hidden_features + 4
                stage = max(int(c["n_comp"]) * 32 * ha * 4 + 2 * int(c["npad"]) * plan.kp * 4 for c in plan.chains)
                self._plan = (plan, torch.from_numpy(plan.chains.view(np.uint8).copy()).to(device),
                              torch.from_numpy(plan.blocks.view(np.uint8).copy()).to(device), stage)
        return self._plan or None

    def _w2t(self, g, plan):
        """W2a of group g re-laid for the tensor-core transform: per chain [hi | lo] tiles of `npad` weight rows x kp columns in
        the K-major core-matrix layout (element (n, k) of a tile at (n // 8) * (kp // 4 * 32) + (k // 4) * 32 + (n % 8) * 4 + k % 4),
        hi = TF32-rounded weight, lo = TF32-rounded remainder.  Rebuilt when the parameters change (same key as _w2a)."""
        W2a = self._w2a(g)
        key = self._w2a_cache[g][0]
        hit = self._w2t_cache.get(g)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                kp, ha = plan.kp, W2a.shape[1]
                out = torch.zeros(plan.w_floats, dtype=torch.float32, device=W2a.device)
                for npad in sorted({w[2] for w in plan.w_index}):
                    idx = [k for k, w in enumerate(plan.w_index) if w[2] == npad]
                    mul = plan.w_index[idx[0]][1]
                    assert all(plan.w_index[k][1] == mul for k in idx)
                    rows0 = torch.tensor([plan.w_index[k][0] for k in idx], device=W2a.device)
                    offs = torch.tensor([int(plan.chains["w_off"][k]) for k in idx], device=W2a.device)
                    w = torch.zeros((len(idx), npad, kp), dtype=torch.float32, device=W2a.device)
                    w[:, :mul, :ha] = W2a[(rows0[:, None] + torch.arange(mul, device=W2a.device)[None, :])]
                    hi = _tf32_round(w)
                    lo = _tf32_round(w - hi)
                    # (n, k) -> [n // 8][k // 4][n % 8][k % 4]
                    def tile(t):
                        return t.view(len(idx), npad // 8, 8, kp // 4, 4).permute(0, 1, 3, 2, 4).reshape(len(idx), npad * kp)
                    both = torch.cat([tile(hi), tile(lo)], dim=1)          # [chains, 2 * npad * kp]
                    pos = offs[:, None] + torch.arange(2 * npad * kp, device=W2a.device)[None, :]
                    out[pos.reshape(-1)] = both.reshape(-1)
            hit = (key, out)
            self._w2t_cache[g] = hit
        return hit[1]

    def _projection(self, groups, cols):
        """[width, H * len(groups)] = the (offset, width) column block of every group's first Linear, transposed and
        concatenated: the node-level projections of a call are one GEMM with it.  Rebuilt when a weight changes."""
        ws = [self._params(g)[0] for g in groups]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        ck = (groups, cols)
        hit = self._proj_cache.get(ck)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                t = torch.cat([w[:, cols[0]:cols[0] + cols[1]] for w in ws], 0).t().contiguous()
            hit = (key, t)
            self._proj_cache[ck] = hit
        return hit[1]

    def device_program(self, device):
        if self._dev_prog is None or self._dev_prog.rows.device != device:
            self._dev_prog = _DeviceProgram(self.program, device)
        return self._dev_prog

    def _check_mode(self):
        if self.training and (self.batch_norm is not None or any(
                isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules())):
            raise NotImplementedError(
                "cb200 TensorProductConvLayer runs the inference path (eval mode: running-stat BatchNorm, no dropout); "
                "call model.eval() before sampling")

    # ------------------------------------------------------------------ fused path
    def run(self, x, segments: List[Segment], n_out, ns, e_cols, agg_cols=None, nbr_cols=None,
            agg_scalars=None, agg_graph=None, residual=None, raw_sum=False, pre=None):
        """x [N_in, d_in]; returns [n_out, d_out].

        e_cols / agg_cols / nbr_cols: (offset, width) of the edge-embedding, aggregation-side and
        neighbour-side scalar blocks inside the first Linear's input (tensor_layers.py:201-202 applied
        to the concatenations built at score_model.py:290-291,314,367-368,397,439-440).
        agg_scalars: [n_out, ns] table for the aggregation side (defaults to x[:, :ns]).
        raw_sum: return the un-normalised sums over incoming edges (no mean / BatchNorm / residual).
        pre: (sum [period, d_out], deg int32 [period], n0, n1) -- a sample-invariant contribution from an earlier
        raw_sum call, added to aggregation nodes [n0, n1) with period `period` before the mean (cb200.h)."""
        self._check_mode()
        dev = x.device
        P = self.program
        dp = self.device_program(dev)
        H = self.hidden_features
        groups = tuple(sorted({s.group for s in segments}))
        # node-level projections of the first Linear (plain GEMM: plumbing)
        xs = x[:, :ns]
        P_nbr, P_agg = {}, {}
        if nbr_cols is not None:
            pn = xs @ self._projection(groups, tuple(nbr_cols))
            for k, g in enumerate(groups):
                P_nbr[g] = (pn, k * H)
        if agg_cols is not None:
            src = xs if agg_scalars is None else agg_scalars
            pa = src @ self._projection(groups, tuple(agg_cols))
            for k, g in enumerate(groups):
                P_agg[g] = (pa, k * H)
        # per-graph edge-embedding offsets (e_post) are constant per aggregation node: fold W1e.e_post[graph(node)] into the
        # node projection instead of re-deriving it for every (node, slot) inside the kernel.  Only when every segment of
        # the group agrees on (e_post, node range); otherwise the kernel's own e_post path handles it.
        folded = set()
        if FOLD_E_POST and agg_cols is not None and agg_graph is not None:
            by_group = {}
            for s in segments:
                by_group.setdefault(s.group, set()).add((None if s.e_post is None else s.e_post.data_ptr(), s.n0, s.n1))
            graph_l = None
            for s in segments:
                if s.e_post is None or s.group in folded or len(by_group[s.group]) != 1:
                    continue
                if graph_l is None:
                    graph_l = agg_graph.long()
                t, off = P_agg[s.group]
                W1e = self._params(s.group)[0][:, e_cols[0]:e_cols[0] + e_cols[1]]
                t[s.n0:s.n1, off:off + H] += (s.e_post @ W1e.t()).index_select(0, graph_l[s.n0:s.n1])
                folded.add(s.group)
        out = torch.empty((n_out, P.d_out), dtype=torch.float32, device=dev)
        a = _lib.TpConvArgs()
        a.x = _lib.f32(x, "x")
        a.d_in, a.d_out, a.S = x.shape[1], P.d_out, P.sh_dim
        assert x.shape[1] == P.d_in, f"node feature width {x.shape[1]} != {P.d_in}"
        a.ne, a.H, a.n_out = e_cols[1], H, n_out
        a.agg_graph = _lib.i32(agg_graph, "agg_graph", allow_none=True)
        a.rows, a.n_rows = dp.rows.data_ptr(), P.n_rows
        a.terms, a.n_terms = dp.terms.data_ptr(), len(P.terms)
        a.runs, a.n_runs = dp.runs.data_ptr(), len(P.runs)
        a.n_segs = len(segments)
        keep = []
        # slots: explicit ids, else one slot per run of adjacent segments with equal (group, n0, n1, e_post)
        slot_ids, prev = [], None
        for s in segments:
            key = (s.group, s.n0, s.n1, None if s.e_post is None else s.e_post.data_ptr()) if s.slot is None else ("id", s.slot)
            slot_ids.append(slot_ids[-1] + (key != prev) if slot_ids else 0)
            prev = key
        tcp = self._tc_plan(dev) if TC_TRANSFORM else None
        if tcp is not None:
            plan, chains_t, blocks_t, stage_bytes = tcp
            a.chains, a.n_chains = chains_t.data_ptr(), len(plan.chains)
            a.blocks, a.n_blocks = blocks_t.data_ptr(), len(plan.blocks)
            a.kp, a.max_chain_bytes = plan.kp, stage_bytes
        for k, s in enumerate(segments):
            W1, b1 = self._params(s.group)[:2]
            W2a = self._w2a(s.group)
            sg = a.segs[k]
            if tcp is not None:
                sg.W2t = self._w2t(s.group, tcp[0]).data_ptr()
            sg.rowptr, sg.col = _lib.i32(s.edges.rowptr, "rowptr"), _lib.i32(s.edges.col, "col")
            sg.e_attr, sg.sh = _lib.f32(s.e_attr, "e_attr"), _lib.f32(s.sh, "sh")
            assert s.e_attr.shape[1] == e_cols[1] and s.sh.shape[1] == P.sh_dim
            sg.e_post = None if s.group in folded else _lib.f32(s.e_post, "e_post", allow_none=True)
            if torch.is_tensor(s.gate):
                assert s.gate.dtype == torch.uint8 and s.gate.is_cuda and s.gate.is_contiguous() and s.gate.numel() == s.n1 - s.n0, \
                    "gate mask must be a contiguous CUDA uint8 tensor over the segment's aggregation nodes"
                sg.gate_mask = s.gate.data_ptr()
            elif s.gate is not None:
                assert s.gate.n_agg == s.n1 - s.n0, "gate edge list must cover the segment's aggregation nodes"
                sg.gate_rowptr = _lib.i32(s.gate.rowptr, "gate.rowptr")
            if s.group in P_agg:
                t, off = P_agg[s.group]
                sg.P_agg, sg.ldp_agg = t.data_ptr() + 4 * off, t.shape[1]
            if s.group in P_nbr:
                t, off = P_nbr[s.group]
                sg.P_nbr, sg.ldp_nbr = t.data_ptr() + 4 * off, t.shape[1]
            sg.W1e, sg.ldw1 = _lib.f32(W1, "W1") + 4 * e_cols[0], W1.shape[1]
            sg.b1, sg.W2a = _lib.f32(b1, "b1"), _lib.f32(W2a, "W2a")
            sg.n0, sg.n1, sg.col_off, sg.slot = s.n0, s.n1, s.col_off, slot_ids[k]
        if raw_sum:
            a.flags = _lib.CB_TP_RAW_SUM
            assert residual is None and pre is None
        elif self.batch_norm is not None:
            scale, shift = self.batch_norm.affine()
            a.bn_scale, a.bn_shift = scale.data_ptr(), shift.data_ptr()
            keep += [scale, shift]
        if pre is not None:
            p_sum, p_deg, p_n0, p_n1 = pre
            assert p_sum.shape[1] == P.d_out and p_deg.numel() == p_sum.shape[0] and (p_n1 - p_n0) % p_sum.shape[0] == 0
            a.pre_sum, a.pre_deg = _lib.f32(p_sum, "pre_sum"), _lib.i32(p_deg, "pre_deg")
            a.pre_n0, a.pre_n1, a.pre_period = p_n0, p_n1, p_sum.shape[0]
        if residual is not None:
            a.residual, a.d_res, a.ld_res = _lib.f32(residual, "residual"), min(residual.shape[1], P.d_out), residual.shape[1]
        a.out = out.data_ptr()
        # bookkeeping for profilers (bench.py): which edge counters / sizes this launch covers
        a._meta = dict(layer=self, n_in=int(x.shape[0]), n_out=int(n_out), groups=list(groups),
                       edge_counters=[s.edges.n_edges_dev for s in segments],
                       edge_gates=[(s.edges, s.gate) for s in segments])
        # workspace: one R x (H+4) accumulator per (node, slot); process the nodes in chunks if it would be huge
        per_item = P.n_rows * (H + 4)
        a.node_begin, a.node_end = 0, n_out
        np_cols = -(-(H + 1) // 16) * 16
        tc_ok = (P.n_rows <= 384 and -(-P.n_rows // 128) * np_cols + 16 <= 256 and x.shape[1] % 2 == 0
                 and e_cols[1] % 8 == 0 and H <= 120)
        a.accum_mode = ACCUM_MODE if tc_ok else 1
        items = _lib.tp_conv_items(a)
        max_items = max(1, WORKSPACE_BYTES // (4 * per_item))
        n_chunks = max(1, -(-items // max_items))
        step = -(-n_out // n_chunks)
        for c0 in range(0, n_out, step):
            a.node_begin, a.node_end = c0, min(n_out, c0 + step)
            it = _lib.tp_conv_items(a)
            ws = _workspace(max(it, 1) * per_item, dev)
            a.workspace, a.workspace_floats = ws.data_ptr(), ws.numel()
            _lib.tp_conv_forward(a)
            if DEBUG_KEEP_WORKSPACE is not None:
                DEBUG_KEEP_WORKSPACE.append(ws.view(-1, P.n_rows, H + 4))
        return out

    # ------------------------------------------------------------------ reference-style call
    def forward(self, node_attr, edge_index, edge_attr, edge_sh, out_nodes=None, reduce="mean", edge_weight=1.0):
        """Drop-in for tensor_layers.py:195-217 with explicit (already concatenated) edge features.
        edge_attr is a tensor (edge_groups == 1) or a list with one tensor per edge group."""
        assert reduce == "mean"
        if not (isinstance(edge_weight, (int, float)) and float(edge_weight) == 1.0):
            raise NotImplementedError("smooth_edges edge weights are not used by the shipped configurations")
        n_out = int(out_nodes or node_attr.shape[0])
        if edge_index.shape[1] == 0:
            out = torch.zeros((n_out, self.out_size), dtype=node_attr.dtype, device=node_attr.device)
        else:
            attrs = [edge_attr] if self.edge_groups == 1 else list(edge_attr)
            segs, start = [], 0
            for g, ea in enumerate(attrs):
                n = ea.shape[0]
                ei = edge_index[:, start:start + n]
                el, perm = static_edges(ei, n_out)
                segs.append(Segment(el, ea[perm].contiguous(), edge_sh[start:start + n][perm].contiguous(), g, 0, n_out))
                start += n
            out = self.run(node_attr.contiguous(), segs, n_out, 0, (0, self.n_edge_features))
        if self.residual:
            out = out + torch.nn.functional.pad(node_attr[:n_out], (0, out.shape[-1] - node_attr.shape[-1]))
        return out
