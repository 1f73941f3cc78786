"""Coarse-grained (C-alpha) TensorProductScoreModel on the cb200 kernels.

Drop-in for models/score_model.py:44-664: same constructor signature, same parameter / buffer names
(so `load_state_dict(strict=True)` works with reference checkpoints), same `forward(data)` contract:
returns (tr_pred [B,3], rot_pred [B,3], tor_pred [sum R] or empty(0), sidechain_pred None), or
(confidence, atom_confidence) in confidence mode.  `data` is any object with the PyG HeteroData
surface (confidence_bootstrapping_b200.data or torch_geometric).

What runs where: neighbour search, edge featurisation + edge-embedding MLPs, every
TensorProductConvLayer and the SDE step are cb200 CUDA kernels; the remaining tiny dense pieces
(embedding sums, [B,33] heads, table lookups) are torch ops on the same stream.  No host
synchronisation happens inside forward once the per-batch static cache exists and the caller has
recorded the (host-known) diffusion time via `set_time` (diffusion_utils.py).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from . import so3, torus
from .graph import (EdgeEmbedder, EdgeList, count_per_bin, host_counts, identity_edges, masked_columns, radius_edges, radius_edges_transposed,
                    static_edges)
from .irreps import full_tp_low_blocks, get_irrep_seq, irreps_str, sh_irreps
from .synthetic import LIG_FEATURE_DIMS, REC_ATOM_FEATURE_DIMS, REC_RESIDUE_FEATURE_DIMS
from .tensor_layers import Segment, TensorProductConvLayer

lig_feature_dims = LIG_FEATURE_DIMS
rec_residue_feature_dims = REC_RESIDUE_FEATURE_DIMS
rec_atom_feature_dims = REC_ATOM_FEATURE_DIMS


class AtomEncoder(nn.Module):
    """Sum of categorical embeddings (+ Linear over [sum | scalar features]) (score_model.py:18-41)."""

    def __init__(self, emb_dim, feature_dims, sigma_embed_dim, lm_embedding_dim=0):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        self.num_categorical_features = len(feature_dims[0])
        self.additional_features_dim = feature_dims[1] + sigma_embed_dim + lm_embedding_dim
        for dim in feature_dims[0]:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)
        if self.additional_features_dim > 0:
            self.additional_features_embedder = nn.Linear(self.additional_features_dim + emb_dim, emb_dim)

    def categorical(self, x):
        out = 0
        for i in range(self.num_categorical_features):
            out = out + self.atom_embedding_list[i](x[:, i].long())
        return out

    def forward(self, x):
        assert x.shape[1] == self.num_categorical_features + self.additional_features_dim
        emb = self.categorical(x)
        if self.additional_features_dim > 0:
            emb = self.additional_features_embedder(torch.cat([emb, x[:, self.num_categorical_features:]], dim=1))
        return emb


def _segment_sum(x, st):
    """Per-graph sum over the ligand rows as one small GEMM with the (cached) graph-membership matrix: fixed summation
    order and no host synchronisation, so the whole CUDA path stays bit-reproducible (index_add_ / scatter use
    floating-point atomics, torch.segment_reduce reads its offsets on the host)."""
    m = getattr(st, "lig_membership", None)
    if m is None:
        m = (st.lig_batch.unsqueeze(0) == torch.arange(st.B, dtype=st.lig_batch.dtype, device=x.device).unsqueeze(1)).float()
        st.lig_membership = m
    return m @ x


class GaussianSmearing(nn.Module):
    """exp(coeff * (d - offset_k)^2) (score_model.py:667-677); evaluated inside K2."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)


def _edge_mlp(n_in, ns, dropout):
    return nn.Sequential(nn.Linear(n_in, ns), nn.ReLU(), nn.Dropout(dropout), nn.Linear(ns, ns))


def _i32(t):
    return t.to(torch.int32).contiguous()


SHARE_REPLICATED_REC_MESSAGES = True   # tests switch it off to compare with the general path


class _Static:
    """Per-Batch cache of everything that does not depend on the ligand pose or the diffusion time."""


class TensorProductScoreModel(nn.Module):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=None, confidence_mode=False,
                 confidence_dropout=0, confidence_no_batchnorm=False,
                 asyncronous_noise_schedule=False, affinity_prediction=False, parallel=1,
                 parallel_aggregators="mean max min std", num_confidence_outputs=1, atom_num_confidence_outputs=1,
                 fixed_center_conv=False, no_aminoacid_identities=False, include_miscellaneous_atoms=False,
                 differentiate_convolutions=True, tp_weights_layers=2, num_prot_emb_layers=0, reduce_pseudoscalars=False,
                 embed_also_ligand=False, atom_confidence=False, sidechain_pred=False, depthwise_convolution=False):
        super().__init__()
        assert parallel == 1, "not implemented"
        for flag, name in ((smooth_edges, "smooth_edges"), (separate_noise_schedule, "separate_noise_schedule"),
                           (asyncronous_noise_schedule, "asyncronous_noise_schedule"),
                           (include_miscellaneous_atoms, "include_miscellaneous_atoms"), (sidechain_pred, "sidechain_pred"),
                           (depthwise_convolution, "depthwise_convolution"), (use_second_order_repr, "use_second_order_repr"),
                           (odd_parity, "odd_parity"), (affinity_prediction, "affinity_prediction")):
            if flag:
                raise NotImplementedError(f"{name} is outside the hot path of the shipped configurations")
        if lm_embedding_type not in (None, "precomputed"):
            raise NotImplementedError("on-the-fly ESM embeddings are preprocessing (out of scope); pass precomputed ones")
        assert (not no_aminoacid_identities) or (lm_embedding_type is None)
        assert sh_lmax in (1, 2)
        self.t_to_sigma = t_to_sigma
        self.in_lig_edge_features = in_lig_edge_features
        self.sigma_embed_dim = sigma_embed_dim
        self.lig_max_radius, self.rec_max_radius = lig_max_radius, rec_max_radius
        self.cross_max_distance, self.dynamic_max_cross = cross_max_distance, dynamic_max_cross
        self.center_max_distance = center_max_distance
        self.distance_embed_dim, self.cross_distance_embed_dim = distance_embed_dim, cross_distance_embed_dim
        self.sh_lmax = sh_lmax
        self.sh_irreps = irreps_str(sh_irreps(sh_lmax))
        self.ns, self.nv = ns, nv
        self.scale_by_sigma, self.norm_by_sigma = scale_by_sigma, norm_by_sigma
        self.device = device
        self.no_torsion = no_torsion
        self.timestep_emb_func = timestep_emb_func
        self.confidence_mode = confidence_mode
        self.num_conv_layers, self.num_prot_emb_layers = num_conv_layers, num_prot_emb_layers
        self.fixed_center_conv = fixed_center_conv
        self.no_aminoacid_identities = no_aminoacid_identities
        self.differentiate_convolutions = differentiate_convolutions
        self.reduce_pseudoscalars = reduce_pseudoscalars
        self.atom_confidence = atom_confidence
        self.atom_num_confidence_outputs = atom_num_confidence_outputs
        self.lm_embedding_type = lm_embedding_type
        self.embed_also_ligand = embed_also_ligand
        lm_embedding_dim = 1280 if lm_embedding_type == "precomputed" else 0

        self.lig_node_embedding = AtomEncoder(ns, lig_feature_dims, sigma_embed_dim)
        self.lig_edge_embedding = _edge_mlp(in_lig_edge_features + sigma_embed_dim + distance_embed_dim, ns, dropout)
        self.rec_node_embedding = AtomEncoder(ns, rec_residue_feature_dims, 0, lm_embedding_dim)
        self.rec_edge_embedding = _edge_mlp(distance_embed_dim, ns, dropout)
        self.rec_sigma_embedding = _edge_mlp(sigma_embed_dim, ns, dropout)
        self.cross_edge_embedding = _edge_mlp(sigma_embed_dim + cross_distance_embed_dim, ns, dropout)
        self.lig_distance_expansion = GaussianSmearing(0.0, lig_max_radius, distance_embed_dim)
        self.rec_distance_expansion = GaussianSmearing(0.0, rec_max_radius, distance_embed_dim)
        self.cross_distance_expansion = GaussianSmearing(0.0, cross_max_distance, cross_distance_embed_dim)

        irrep_seq = get_irrep_seq(ns, nv, use_second_order_repr, reduce_pseudoscalars)
        faster = sh_lmax == 1 and not use_second_order_repr

        def conv(i, groups):
            return TensorProductConvLayer(
                in_irreps=irrep_seq[min(i, len(irrep_seq) - 1)], sh_irreps=self.sh_irreps,
                out_irreps=irrep_seq[min(i + 1, len(irrep_seq) - 1)], n_edge_features=3 * ns, hidden_features=3 * ns,
                residual=True, batch_norm=batch_norm, dropout=dropout, faster=faster,
                tp_weights_layers=tp_weights_layers, edge_groups=groups, depthwise=depthwise_convolution)

        self.rec_emb_layers = nn.ModuleList([conv(i, 1) for i in range(num_prot_emb_layers)])
        if embed_also_ligand:
            self.lig_emb_layers = nn.ModuleList([conv(i, 1) for i in range(num_prot_emb_layers)])
        last = num_prot_emb_layers + num_conv_layers - 1
        self.conv_layers = nn.ModuleList([
            conv(i, 1 if not differentiate_convolutions else (2 if i == last else 4))
            for i in range(num_prot_emb_layers, num_prot_emb_layers + num_conv_layers)])

        if self.confidence_mode:
            input_size = ns + (nv if reduce_pseudoscalars else ns) if num_conv_layers + num_prot_emb_layers >= 3 else ns

            def head(n_in, n_out):
                bn = (lambda: nn.BatchNorm1d(ns)) if not confidence_no_batchnorm else (lambda: nn.Identity())
                return nn.Sequential(nn.Linear(n_in, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
                                     nn.Linear(ns, ns), bn(), nn.ReLU(), nn.Dropout(confidence_dropout),
                                     nn.Linear(ns, n_out))
            if self.atom_confidence:
                self.atom_confidence_predictor = head(input_size, atom_num_confidence_outputs + ns)
                input_size = ns
            self.confidence_predictor = head(input_size, num_confidence_outputs)
        else:
            self.center_distance_expansion = GaussianSmearing(0.0, center_max_distance, distance_embed_dim)
            self.center_edge_embedding = _edge_mlp(distance_embed_dim + sigma_embed_dim, ns, dropout)
            self.final_conv = TensorProductConvLayer(
                in_irreps=self.conv_layers[-1].out_irreps, sh_irreps=self.sh_irreps, out_irreps="2x1o + 2x1e",
                n_edge_features=2 * ns, residual=False, dropout=dropout, batch_norm=batch_norm)
            self.tr_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
            self.rot_final_layer = nn.Sequential(nn.Linear(1 + sigma_embed_dim, ns), nn.Dropout(dropout), nn.ReLU(), nn.Linear(ns, 1))
            if not no_torsion:
                self._init_torsion_head(ns, sh_lmax, dropout, batch_norm, distance_embed_dim)
        # e3nn modules in reference checkpoints carry constant buffers under `*.tp.*` / `final_tp_tor.*`
        self._register_load_state_dict_pre_hook(self._drop_e3nn_buffers)

    def _init_torsion_head(self, ns, sh_lmax, dropout, batch_norm, distance_embed_dim):
        """final_edge_embedding, final_tp_tor, tor_bond_conv, tor_final_layer (score_model.py:257-280).
        o3.FullTensorProduct(sh_irreps, "2e") has no parameters; only its l<=1 output blocks can reach the
        scalar outputs of tor_bond_conv from l<=1 node features, so only those are evaluated."""
        self.final_edge_embedding = _edge_mlp(distance_embed_dim, ns, dropout)
        tor_sh_irreps, blocks = full_tp_low_blocks(sh_lmax)
        self._tor_blocks = [l_in for l_in, _ in blocks]
        for k, (_, w) in enumerate(blocks):
            self.register_buffer(f"_tor_w{k}", torch.tensor(w, dtype=torch.float32), persistent=False)
        self.tor_bond_conv = TensorProductConvLayer(
            in_irreps=self.conv_layers[-1].out_irreps, sh_irreps=tor_sh_irreps, out_irreps=f"{ns}x0o + {ns}x0e",
            n_edge_features=3 * ns, residual=False, dropout=dropout, batch_norm=batch_norm)
        self.tor_final_layer = nn.Sequential(nn.Linear(2 * ns, ns, bias=False), nn.Tanh(), nn.Dropout(dropout),
                                             nn.Linear(ns, 1, bias=False))

    @staticmethod
    def _drop_e3nn_buffers(state_dict, prefix, *args):
        for k in [k for k in state_dict if ".tp." in k or k.startswith(prefix + "final_tp_tor.")]:
            del state_dict[k]

    # ------------------------------------------------------------------ embedders (views over the parameters)
    def _embedders(self):
        s, d, c = self.sigma_embed_dim, self.distance_embed_dim, self.cross_distance_embed_dim
        e = SimpleNamespace()
        f = self.in_lig_edge_features
        e.lig = EdgeEmbedder(self.lig_edge_embedding, self.lig_distance_expansion, f, 0, f, s, f + s)
        e.rec = EdgeEmbedder(self.rec_edge_embedding, self.rec_distance_expansion, 0, 0, 0, 0, 0)
        e.cross = EdgeEmbedder(self.cross_edge_embedding, self.cross_distance_expansion, 0, 0, 0, s, s)
        if not self.confidence_mode:
            e.center = EdgeEmbedder(self.center_edge_embedding, self.center_distance_expansion, 0, 0, d, s, 0)
            if not self.no_torsion:
                e.final = EdgeEmbedder(self.final_edge_embedding, self.lig_distance_expansion, 0, 0, 0, 0, 0)
        return e

    # ------------------------------------------------------------------ static per-batch cache
    def _static(self, data):
        rec = data["receptor"]
        cache = getattr(rec, "cb200_static", None) if hasattr(rec, "cb200_static") else None
        if cache is not None and cache.get("owner") == id(self):
            return cache["static"]
        st = _Static()
        lig, ll, rr = data["ligand"], data["ligand", "ligand"], data["receptor", "receptor"]
        dev = lig.pos.device
        B = int(data.num_graphs)
        st.B, st.dev = B, dev
        st.lig_batch, st.rec_batch = _i32(lig.batch), _i32(rec.batch)
        st.NL, st.NR = int(lig.pos.shape[0]), int(rec.pos.shape[0])
        nl = count_per_bin(lig.batch, B)
        nr = count_per_bin(rec.batch, B)
        st.lig_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        st.lig_ptr[1:] = torch.cumsum(nl, 0)
        st.rec_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        st.rec_ptr[1:] = torch.cumsum(nr, 0)
        hc = host_counts(data, ("ligand", "receptor"))      # from the collate's host tables: no device read
        if hc is not None:
            nl_h, nr_h = hc["ligand"], hc["receptor"]
        else:
            nl_h, nr_h = nl.tolist(), nr.tolist()         # one host read per batch (sizes for buffer caps)
        st.nl_h, st.nr_h = nl_h, nr_h
        st.cap_cross = int(sum(a * b for a, b in zip(nl_h, nr_h)))
        st.cap_lig = int(sum(a * min(a - 1, 33) for a in nl_h if a > 0)) + 1
        # static topologies
        st.bond_edges, perm = static_edges(ll.edge_index, st.NL)
        st.bond_attr = ll.edge_attr.float()[perm].contiguous()
        st.rec_edges, _ = static_edges(rr.edge_index, st.NR)
        st.rec_pos = rec.pos.float().contiguous()
        st.rec_x = rec.x.float()
        if self.no_aminoacid_identities:
            st.rec_x = st.rec_x * 0
        # rotatable bonds (score_model.py:650-652)
        mask = lig.edge_mask.bool()
        st.tor_bonds = masked_columns(ll.edge_index, mask, None if hc is None else sum(hc["n_tor"])).long()
        st.n_tor = int(st.tor_bonds.shape[1])              # (a host read, once per batch, without the host tables)
        if st.n_tor > 0:
            st.tor_batch = _i32(lig.batch[st.tor_bonds[0]])
            if hc is not None:
                nt_h = hc["n_tor"]
            else:
                nt_h = torch.bincount(lig.batch[st.tor_bonds[0]], minlength=B).tolist()
            st.cap_tor = int(sum(t * min(a, 32) for t, a in zip(nt_h, nl_h)))
        st.lig_cat = self.lig_node_embedding.categorical(lig.x)           # pose/time independent part
        st.center_edges = identity_edges(st.lig_ptr, st.NL)
        st.inv_nl = (1.0 / nl.float().clamp(min=1)).unsqueeze(1)
        # receptor embedding (score_model.py:298-320), cached like the reference caches it on the Batch
        emb = self._embedders()
        zero_sigma = None
        rec_e_attr, rec_sh = emb.rec(st.rec_edges, st.rec_pos, st.rec_pos, st.rec_batch, zero_sigma, self.sh_lmax)
        ns = self.ns
        n_rec_edges = st.rec_edges.cap
        if ("receptor" in getattr(data, "_g", {}).get("_replicated_types", ()) and B > 1 and len(set(nr_h)) == 1
                and n_rec_edges % B == 0 and n_rec_edges > 0):
            # the batch holds B copies of one receptor (flagged by the collate): its pose-independent embedding is computed
            # for the first copy and tiled.  Graph 0 owns the first NR/B nodes and, the static edge list being sorted by
            # aggregation node, the first E/B edges.
            n0, e0 = st.NR // B, n_rec_edges // B
            e_first = EdgeList(st.rec_edges.rowptr[:n0 + 1], st.rec_edges.row[:e0], st.rec_edges.col[:e0], e0, n0)
            x = self.rec_node_embedding(st.rec_x[:n0])
            for layer in self.rec_emb_layers:
                seg = Segment(e_first, rec_e_attr[:e0], rec_sh[:e0], 0, 0, n0)
                x = layer.run(x.contiguous(), [seg], n0, ns, (0, ns), (ns, ns), (2 * ns, ns), residual=x.contiguous())
            x = x.repeat(B, 1)
            # sample-invariant rec->rec messages of the first conv layer are computed on this first copy (forward())
            st.rec_first = SimpleNamespace(n0=n0, e0=e0, edges=e_first,
                                           deg=(st.rec_edges.rowptr[1:n0 + 1] - st.rec_edges.rowptr[:n0]).contiguous())
        else:
            x = self.rec_node_embedding(st.rec_x)
            for layer in self.rec_emb_layers:
                seg = Segment(st.rec_edges, rec_e_attr, rec_sh, 0, 0, st.NR)
                x = layer.run(x.contiguous(), [seg], st.NR, ns, (0, ns), (ns, ns), (2 * ns, ns), residual=x.contiguous())
        st.rec_node_attr, st.rec_e_attr, st.rec_sh = x.contiguous(), rec_e_attr, rec_sh
        rec.cb200_static = {"owner": id(self), "static": st}
        return st

    # ------------------------------------------------------------------ ligand graph + embedding
    def _ligand_embedding(self, st, lig_pos, sigma_emb, emb):
        """build_lig_conv_graph + lig_node/edge_embedding + lig_emb_layers (score_model.py:282-295,492-522).
        Returns (lig_x, lig_segments(group, n1) -> [bond segment, radius segment])."""
        ns, lmax = self.ns, self.sh_lmax
        fwd = radius_edges(lig_pos, st.lig_ptr, lig_pos, st.lig_batch, self.lig_max_radius, 33, st.cap_lig,
                           exclude_self=True)
        rad = radius_edges_transposed(lig_pos, st.lig_batch, lig_pos, st.lig_ptr, self.lig_max_radius,
                                      st.cap_lig, exclude_self=True, kept=fwd)
        bond_attr, bond_sh = emb.lig(st.bond_edges, lig_pos, lig_pos, st.lig_batch, sigma_emb, lmax, extra=st.bond_attr)
        rad_attr, rad_sh = emb.lig(rad, lig_pos, lig_pos, st.lig_batch, sigma_emb, lmax)
        lig_x = self.lig_node_embedding.additional_features_embedder(
            torch.cat([st.lig_cat, sigma_emb[st.lig_batch.long()]], dim=1)).contiguous()
        NL = st.NL

        def lig_segments(group, n1):
            return [Segment(st.bond_edges, bond_attr, bond_sh, group, 0, n1),
                    Segment(rad, rad_attr, rad_sh, group, 0, n1)]

        cols = dict(e_cols=(0, ns), agg_cols=(ns, ns), nbr_cols=(2 * ns, ns))
        if self.embed_also_ligand:
            for layer in self.lig_emb_layers:
                lig_x = layer.run(lig_x, lig_segments(0, NL), NL, ns, residual=lig_x, **cols)
        return lig_x, lig_segments

    # ------------------------------------------------------------------ forward
    def forward(self, data):
        st = self._static(data)
        emb = self._embedders()
        ns, lmax, dev, B = self.ns, self.sh_lmax, st.dev, st.B
        lig_pos = data["ligand"].pos.float().contiguous()
        t = {k: data.complex_t[k].float() for k in ("tr", "rot", "tor")}
        if not self.confidence_mode:
            tr_sigma, rot_sigma, tor_sigma = self.t_to_sigma(t["tr"], t["rot"], t["tor"])
        else:
            tr_sigma, rot_sigma, tor_sigma = t["tr"], t["rot"], t["tor"]
        sigma_emb = self.timestep_emb_func(t["tr"]).float().contiguous()             # [B, sigma_embed_dim]
        rec_sigma_emb = self.rec_sigma_embedding(sigma_emb).contiguous()             # [B, ns]

        lig_x, lig_segments = self._ligand_embedding(st, lig_pos, sigma_emb, emb)
        NL, NR = st.NL, st.NR
        cols = dict(e_cols=(0, ns), agg_cols=(ns, ns), nbr_cols=(2 * ns, ns))

        # ---- cross graph (score_model.py:346-351,564-587); both directions are emitted sorted
        if self.dynamic_max_cross:
            cutoff = (tr_sigma * 3 + 20).float().contiguous()
            r = 1.0
        else:
            cutoff, r = None, float(self.cross_max_distance)
        lr = radius_edges(st.rec_pos, st.rec_ptr, lig_pos, st.lig_batch, r, 10000, st.cap_cross, cutoff=cutoff)
        rl = radius_edges_transposed(st.rec_pos, st.rec_batch, lig_pos, st.lig_ptr, r, st.cap_cross, cutoff=cutoff)
        lr_attr, lr_sh = emb.cross(lr, lig_pos, st.rec_pos, st.lig_batch, sigma_emb, lmax)
        # reversed edges carry SH(-vec) = SH(lig - rec) (score_model.py:359,582): neighbour minus aggregation again
        rl_attr, rl_sh = emb.cross(rl, st.rec_pos, lig_pos, st.rec_batch, sigma_emb, lmax)

        # ---- joint node table [ligand ; receptor] and the conv stack (score_model.py:354-374)
        rec_x = st.rec_node_attr.clone()
        rec_x[:, :ns] += rec_sigma_emb[st.rec_batch.long()]
        if lig_x.shape[1] != rec_x.shape[1]:
            lig_x = torch.nn.functional.pad(lig_x, (0, rec_x.shape[1] - lig_x.shape[1]))
        x = torch.cat([lig_x, rec_x], 0).contiguous()
        node_graph = torch.cat([st.lig_batch, st.rec_batch]).contiguous()
        g = (0, 1, 2, 3) if self.differentiate_convolutions else (0, 0, 0, 0)
        n_layers = len(self.conv_layers)
        gates = self._dead_output_gates(st, rl, n_layers)
        # In the FIRST conv layer the receptor rows of the B copies of one complex are still identical (same receptor, same
        # t), so the rec->rec slot -- 40 % of the layer's edges -- is aggregated and transformed once for the first copy and
        # shared (cb_tp_conv_args.pre_sum).  Needs the collate's "replicated receptor" flag and one diffusion time for the
        # whole batch (what sampling() feeds: utils/sampling.py:110); anything else takes the general path.
        host_t = getattr(data, "complex_t_host", None) if hasattr(data, "complex_t_host") else None
        step = getattr(data, "cb200_step", None) if hasattr(data, "cb200_step") else None      # device-resident step scalars (CUDA graph)
        share_rec = (SHARE_REPLICATED_REC_MESSAGES and getattr(st, "rec_first", None) is not None
                     and (host_t is not None or step is not None) and n_layers > 1)
        for l, layer in enumerate(self.conv_layers):
            if l < n_layers - 1:
                gate = gates.get(l)
                cross = [Segment(lr, lr_attr, lr_sh, g[1], 0, NL, col_off=NL)]
                rec_in = Segment(rl, rl_attr, rl_sh, g[3], NL, NL + NR, col_off=0)
                if l == 0 and share_rec:
                    f = st.rec_first
                    first = Segment(f.edges, st.rec_e_attr[:f.e0], st.rec_sh[:f.e0], g[2], 0, f.n0, col_off=0, e_post=rec_sigma_emb)
                    pre_sum = layer.run(x[NL:NL + f.n0], [first], f.n0, ns, raw_sum=True, **cols)
                    segs = lig_segments(g[0], NL) + cross + [rec_in]
                    x = layer.run(x, segs, NL + NR, ns, agg_graph=node_graph, residual=x, pre=(pre_sum, f.deg, NL, NL + NR), **cols)
                    continue
                segs = lig_segments(g[0], NL) + cross + [
                    Segment(st.rec_edges, st.rec_e_attr, st.rec_sh, g[2], NL, NL + NR, col_off=NL, e_post=rec_sigma_emb, gate=gate),
                    rec_in]
                x = layer.run(x, segs, NL + NR, ns, agg_graph=node_graph, residual=x, **cols)
            else:
                # last layer only updates the ligand rows (the reference normalises the others but never reads them)
                segs = lig_segments(g[0], NL) + [Segment(lr, lr_attr, lr_sh, g[1], 0, NL, col_off=NL)]
                x = layer.run(x, segs, NL, ns, agg_graph=node_graph, residual=x, **cols)
        lig_x = x[:NL].contiguous()

        if self.confidence_mode:
            return self._confidence_head(lig_x, st)

        return self._score_heads(data, st, emb, lig_x, lig_pos, sigma_emb, tr_sigma, rot_sigma, tor_sigma)

    @staticmethod
    def _dead_output_gates(st, rl, n_layers, hops=4):
        """Receptor rows whose rec->rec update nobody reads, per conv layer (exact pruning, no approximation).
        The last layer only updates ligand rows and reads receptor rows through the rec->lig cross edges, so the layer
        before it needs the rec->rec update only for receptors with a cross edge (non-empty row of the flipped list
        `rl`): keep set G0.  One layer earlier the needed rows are G0 plus everything G0 reads through rec->rec edges
        (G1 = G0 u N(G0)), and so on.  Returns {layer index: uint8 keep mask over the receptor nodes}."""
        NR = st.NR
        if n_layers < 2 or NR == 0:
            return {}
        keep = rl.rowptr[1:NR + 1] > rl.rowptr[:NR]
        gates = {n_layers - 2: keep.to(torch.uint8)}
        if st.rec_edges.cap > 0:
            if getattr(st, "rec_row_long", None) is None:
                st.rec_row_long, st.rec_col_long = st.rec_edges.row.long(), st.rec_edges.col.long()
            for h in range(1, hops + 1):
                l = n_layers - 2 - h
                if l < 0:
                    break
                reads = torch.zeros(NR, dtype=torch.int32, device=keep.device)
                reads.index_add_(0, st.rec_col_long, keep.index_select(0, st.rec_row_long).to(torch.int32))   # integer adds: exact
                keep = keep | (reads > 0)
                gates[l] = keep.to(torch.uint8)
        return gates

    def _score_heads(self, data, st, emb, lig_x, lig_pos, sigma_emb, tr_sigma, rot_sigma, tor_sigma):
        ns, lmax, dev, B = self.ns, self.sh_lmax, st.dev, st.B
        # ---- translation / rotation head (score_model.py:394-420)
        center = _segment_sum(lig_pos, st) * st.inv_nl
        graph_ids = torch.arange(B, dtype=torch.int32, device=dev)
        c_attr, c_sh = emb.center(st.center_edges, center.contiguous(), lig_pos, graph_ids, sigma_emb, lmax)
        seg = [Segment(st.center_edges, c_attr, c_sh, 0, 0, B)]
        if self.fixed_center_conv:
            global_pred = self.final_conv.run(lig_x, seg, B, ns, (0, ns), None, (ns, ns))
        else:
            global_pred = self.final_conv.run(lig_x, seg, B, ns, (0, ns), (ns, ns), None,
                                              agg_scalars=lig_x[:B, :ns])
        tr_pred = global_pred[:, :3] + global_pred[:, 6:9]
        rot_pred = global_pred[:, 3:6] + global_pred[:, 9:]
        tr_norm = torch.linalg.vector_norm(tr_pred, dim=1).unsqueeze(1)
        tr_pred = tr_pred / tr_norm * self.tr_final_layer(torch.cat([tr_norm, sigma_emb], dim=1))
        rot_norm = torch.linalg.vector_norm(rot_pred, dim=1).unsqueeze(1)
        rot_pred = rot_pred / rot_norm * self.rot_final_layer(torch.cat([rot_norm, sigma_emb], dim=1))
        host_t = getattr(data, "complex_t_host", None) if hasattr(data, "complex_t_host") else None
        step = getattr(data, "cb200_step", None) if hasattr(data, "cb200_step") else None
        if self.scale_by_sigma:
            tr_pred = tr_pred / tr_sigma.unsqueeze(1)
            so3_norm = step["so3_norm"].expand(B) if step is not None else so3.score_norm_device(rot_sigma, host_t, self.t_to_sigma, dev)
            rot_pred = rot_pred * so3_norm.unsqueeze(1)
        if self.no_torsion or st.n_tor == 0:
            return tr_pred, rot_pred, torch.empty(0, device=dev), None

        # ---- torsion head (score_model.py:432-448)
        tb = st.tor_bonds
        bond_pos = ((lig_pos[tb[0]] + lig_pos[tb[1]]) / 2).contiguous()
        n_tor = st.n_tor
        tor_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)  # unused by the forward search
        te = radius_edges(lig_pos, st.lig_ptr, bond_pos, st.tor_batch, self.lig_max_radius, 32, st.cap_tor)
        t_attr, t_sh = emb.final(te, bond_pos, lig_pos, st.tor_batch, None, lmax)
        bond_vec = lig_pos[tb[1]] - lig_pos[tb[0]]
        u = bond_vec / bond_vec.norm(dim=-1, keepdim=True).clamp(min=1e-12)
        s3 = math.sqrt(3.0)
        X, Y, Z = u[:, 0], u[:, 1], u[:, 2]
        y2 = math.sqrt(5.0) * torch.stack([s3 * X * Z, s3 * X * Y, Y * Y - 0.5 * (X * X + Z * Z), s3 * Y * Z,
                                           (s3 / 2.0) * (Z * Z - X * X)], -1)
        # FullTensorProduct(sh, Y2(bond))'s l<=1 blocks per edge: sqrt(2lo+1) * w3j(l,2,lo)[i,j,k] sh_l[i] Y2[j]
        y2e = y2[te.row.long()]
        tor_sh = torch.cat([torch.einsum("ijk,ei,ej->ek", getattr(self, f"_tor_w{k}"), t_sh[:, l_in * l_in:(l_in + 1) ** 2], y2e)
                            for k, l_in in enumerate(self._tor_blocks)], dim=1).contiguous()
        bond_attr_sum = (lig_x[tb[0], :ns] + lig_x[tb[1], :ns]).contiguous()
        seg = [Segment(te, t_attr, tor_sh, 0, 0, n_tor)]
        tor_pred = self.tor_bond_conv.run(lig_x, seg, n_tor, ns, (0, ns), (2 * ns, ns), (ns, ns), agg_scalars=bond_attr_sum)
        tor_pred = self.tor_final_layer(tor_pred).squeeze(1)
        if self.scale_by_sigma:
            if step is not None:
                tor_pred = tor_pred * torch.sqrt(step["torus_norm"])
            else:
                edge_sigma = tor_sigma[st.tor_batch.long()]
                tor_pred = tor_pred * torch.sqrt(torus.score_norm_device(edge_sigma, host_t, self.t_to_sigma, st.tor_batch, dev))
        return tr_pred, rot_pred, tor_pred, None

    def _confidence_head(self, lig_x, st):
        ns = self.ns
        if self.num_conv_layers + self.num_prot_emb_layers >= 3:
            tail = self.nv if self.reduce_pseudoscalars else ns
            scalar = torch.cat([lig_x[:, :ns], lig_x[:, -tail:]], dim=1)
        else:
            scalar = lig_x[:, :ns]
        if self.atom_confidence:
            scalar = self.atom_confidence_predictor(scalar)
            atom_confidence = scalar[:, :self.atom_num_confidence_outputs]
            scalar = scalar[:, self.atom_num_confidence_outputs:]
        else:
            atom_confidence = torch.zeros((len(lig_x),), device=lig_x.device)
        pooled = _segment_sum(scalar, st) * st.inv_nl
        confidence = self.confidence_predictor(pooled).squeeze(dim=-1)
        return confidence, atom_confidence
