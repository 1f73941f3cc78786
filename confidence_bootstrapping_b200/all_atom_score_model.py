"""All-atom TensorProductScoreModel on the cb200 kernels (the shipped confidence model is this class in
confidence mode).

Drop-in for models/all_atom_score_model.py:21-664: three node types (ligand atoms, receptor residues,
receptor atoms), nine edge groups per conv layer `[lig, lr, la, rec, rl, ra, atom, al, ar]`
(:406-418; the last layer keeps the first three, :427-429), e3nn FullyConnectedTensorProduct with
l<=2 edge harmonics, optional receptor embedding block with four groups (:301-314).  Flipped groups
REUSE the forward edge harmonics (:411-412, :305) -- unlike the CG model, which recomputes SH(-vec).
Parameter names follow the reference so its checkpoints load with strict=True.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from .graph import EdgeEmbedder, count_per_bin, host_counts, masked_columns, radius_edges, radius_edges_transposed, static_edges
from .irreps import get_irrep_seq, irreps_str, sh_irreps
from .score_model import (AtomEncoder, GaussianSmearing, _Static, _edge_mlp, _i32, lig_feature_dims,
                          rec_atom_feature_dims, rec_residue_feature_dims)
from .score_model import TensorProductScoreModel as _CGModel
from .tensor_layers import Segment, TensorProductConvLayer


class TensorProductScoreModel(_CGModel):
    def __init__(self, t_to_sigma, device, timestep_emb_func, in_lig_edge_features=4, sigma_embed_dim=32, sh_lmax=2,
                 ns=16, nv=4, num_conv_layers=2, lig_max_radius=5, rec_max_radius=30, cross_max_distance=250,
                 center_max_distance=30, distance_embed_dim=32, cross_distance_embed_dim=32, no_torsion=False,
                 scale_by_sigma=True, norm_by_sigma=True, use_second_order_repr=False, batch_norm=True,
                 dynamic_max_cross=False, dropout=0.0, smooth_edges=False, odd_parity=False,
                 separate_noise_schedule=False, lm_embedding_type=False, confidence_mode=False,
                 confidence_dropout=0, confidence_no_batchnorm=False,
                 asyncronous_noise_schedule=False, affinity_prediction=False, parallel=1,
                 parallel_aggregators="mean max min std", num_confidence_outputs=1, atom_num_confidence_outputs=1,
                 fixed_center_conv=False, no_aminoacid_identities=False, include_miscellaneous_atoms=False,
                 differentiate_convolutions=True, tp_weights_layers=2, num_prot_emb_layers=0,
                 reduce_pseudoscalars=False, embed_also_ligand=False, atom_confidence=False, sidechain_pred=False,
                 depthwise_convolution=False, crop_beyond=None):
        nn.Module.__init__(self)
        assert not sidechain_pred, "sidechain prediction not implemented/makes sense for all atom model"
        assert not depthwise_convolution, "depthwise convolution not implemented for all atom model"
        assert parallel == 1, "affinity prediction with parallel>1 is outside the shipped configurations"
        for flag, name in ((smooth_edges, "smooth_edges"), (separate_noise_schedule, "separate_noise_schedule"),
                           (asyncronous_noise_schedule, "asyncronous_noise_schedule"),
                           (include_miscellaneous_atoms, "include_miscellaneous_atoms"),
                           (use_second_order_repr, "use_second_order_repr"), (odd_parity, "odd_parity"),
                           (affinity_prediction, "affinity_prediction"), (crop_beyond is not None, "crop_beyond (dead branch, quirk 10)")):
            if flag:
                raise NotImplementedError(f"{name} is outside the hot path of the shipped configurations")
        if lm_embedding_type not in (None, False, "precomputed"):
            raise NotImplementedError("on-the-fly ESM embeddings are preprocessing (out of scope); pass precomputed ones")
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
        self.rec_sigma_embedding = _edge_mlp(sigma_embed_dim, ns, dropout)
        self.rec_node_embedding = AtomEncoder(ns, rec_residue_feature_dims, 0, lm_embedding_dim)
        self.rec_edge_embedding = _edge_mlp(distance_embed_dim, ns, dropout)
        self.atom_node_embedding = AtomEncoder(ns, rec_atom_feature_dims, 0)
        self.atom_edge_embedding = _edge_mlp(distance_embed_dim, ns, dropout)
        self.lr_edge_embedding = _edge_mlp(sigma_embed_dim + cross_distance_embed_dim, ns, dropout)
        self.ar_edge_embedding = _edge_mlp(distance_embed_dim, ns, dropout)
        self.la_edge_embedding = _edge_mlp(sigma_embed_dim + cross_distance_embed_dim, ns, dropout)
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
                tp_weights_layers=tp_weights_layers, edge_groups=groups)

        dc = differentiate_convolutions
        self.rec_emb_layers = nn.ModuleList([conv(i, 4 if dc else 1) for i in range(num_prot_emb_layers)])
        if embed_also_ligand:
            self.lig_emb_layers = nn.ModuleList([conv(i, 1) for i in range(num_prot_emb_layers)])
        last = num_prot_emb_layers + num_conv_layers - 1
        self.conv_layers = nn.ModuleList([conv(i, 1 if not dc else (3 if i == last else 9))
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
        self._register_load_state_dict_pre_hook(self._drop_e3nn_buffers)

    # ------------------------------------------------------------------ embedders
    def _embedders(self):
        s, d = self.sigma_embed_dim, self.distance_embed_dim
        f = self.in_lig_edge_features
        e = SimpleNamespace()
        e.lig = EdgeEmbedder(self.lig_edge_embedding, self.lig_distance_expansion, f, 0, f, s, f + s)
        e.rec = EdgeEmbedder(self.rec_edge_embedding, self.rec_distance_expansion, 0, 0, 0, 0, 0)
        e.atom = EdgeEmbedder(self.atom_edge_embedding, self.lig_distance_expansion, 0, 0, 0, 0, 0)
        e.ar = EdgeEmbedder(self.ar_edge_embedding, self.rec_distance_expansion, 0, 0, 0, 0, 0)
        e.lr = EdgeEmbedder(self.lr_edge_embedding, self.cross_distance_expansion, 0, 0, 0, s, s)
        e.la = EdgeEmbedder(self.la_edge_embedding, self.lig_distance_expansion, 0, 0, 0, s, s)
        if not self.confidence_mode:
            e.center = EdgeEmbedder(self.center_edge_embedding, self.center_distance_expansion, 0, 0, d, s, 0)
            if not self.no_torsion:
                e.final = EdgeEmbedder(self.final_edge_embedding, self.lig_distance_expansion, 0, 0, 0, 0, 0)
        return e

    # ------------------------------------------------------------------ static per-batch cache
    def _static(self, data):
        rec = data["receptor"]
        cache = getattr(rec, "cb200_static_aa", None) if hasattr(rec, "cb200_static_aa") else None
        if cache is not None and cache.get("owner") == id(self):
            return cache["static"]
        st = _Static()
        lig, ll, rr = data["ligand"], data["ligand", "ligand"], data["receptor", "receptor"]
        atom, aa, ar = data["atom"], data["atom", "atom"], data["atom", "receptor"]
        dev = lig.pos.device
        B = int(data.num_graphs)
        st.B, st.dev = B, dev
        st.lig_batch, st.rec_batch, st.atom_batch = _i32(lig.batch), _i32(rec.batch), _i32(atom.batch)
        st.NL, st.NR, st.NA = int(lig.pos.shape[0]), int(rec.pos.shape[0]), int(atom.pos.shape[0])

        def ptr_of(batch):
            n = count_per_bin(batch, B)
            p = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            p[1:] = torch.cumsum(n, 0)
            return p, n

        st.lig_ptr, nl = ptr_of(lig.batch)
        st.rec_ptr, nr = ptr_of(rec.batch)
        st.atom_ptr, na = ptr_of(atom.batch)
        hc = host_counts(data, ("ligand", "receptor", "atom"))            # from the collate's host tables: no device read
        if hc is not None:
            nl_h, nr_h, na_h = hc["ligand"], hc["receptor"], hc["atom"]
        else:
            nl_h, nr_h, na_h = nl.tolist(), nr.tolist(), na.tolist()      # one host read per batch
        st.nl_h = nl_h
        st.cap_cross = int(sum(a * b for a, b in zip(nl_h, nr_h)))
        st.cap_la = int(sum(a * b for a, b in zip(nl_h, na_h)))
        st.cap_lig = int(sum(a * min(a - 1, 33) for a in nl_h if a > 0)) + 1
        st.bond_edges, perm = static_edges(ll.edge_index, st.NL)
        st.bond_attr = ll.edge_attr.float()[perm].contiguous()
        st.rec_pos, st.atom_pos = rec.pos.float().contiguous(), atom.pos.float().contiguous()
        rec_x = rec.x.float()
        if self.no_aminoacid_identities:
            rec_x = rec_x * 0
        mask = lig.edge_mask.bool()
        st.tor_bonds = masked_columns(ll.edge_index, mask, None if hc is None else sum(hc["n_tor"])).long()
        st.n_tor = int(st.tor_bonds.shape[1])
        if st.n_tor > 0:
            st.tor_batch = _i32(lig.batch[st.tor_bonds[0]])
            if hc is not None:
                nt_h = hc["n_tor"]
            else:
                nt_h = torch.bincount(lig.batch[st.tor_bonds[0]], minlength=B).tolist()
            st.cap_tor = int(sum(t * min(a, 32) for t, a in zip(nt_h, nl_h)))
        st.lig_cat = self.lig_node_embedding.categorical(lig.x)
        from .graph import identity_edges
        st.center_edges = identity_edges(st.lig_ptr, st.NL)
        st.inv_nl = (1.0 / nl.float().clamp(min=1)).unsqueeze(1)

        # static receptor-side graphs, each as CSR over its aggregation node (all_atom_score_model.py:297-308)
        emb = self._embedders()
        lmax, ns = self.sh_lmax, self.ns
        st.rec_edges, _ = static_edges(rr.edge_index, st.NR)                       # agg rec  <- rec
        st.atom_edges, _ = static_edges(aa.edge_index, st.NA)                      # agg atom <- atom
        st.ar_edges, _ = static_edges(ar.edge_index, st.NA)                        # agg atom <- rec   (group `ar`)
        st.ra_edges, _ = static_edges(torch.flip(ar.edge_index, dims=[0]), st.NR)  # agg rec  <- atom  (flipped `ar`)
        st.rec_e, st.rec_sh = emb.rec(st.rec_edges, st.rec_pos, st.rec_pos, st.rec_batch, None, lmax)
        st.atom_e, st.atom_sh = emb.atom(st.atom_edges, st.atom_pos, st.atom_pos, st.atom_batch, None, lmax)
        # ar_edge_sh = SH(rec_pos[idx1] - atom_pos[idx0]); the flipped group reuses it (:305, :411-412)
        st.ar_e, st.ar_sh = emb.ar(st.ar_edges, st.atom_pos, st.rec_pos, st.atom_batch, None, lmax, sh_sign=1.0)
        st.ra_e, st.ra_sh = emb.ar(st.ra_edges, st.rec_pos, st.atom_pos, st.rec_batch, None, lmax, sh_sign=-1.0)
        rx = self.rec_node_embedding(rec_x)
        ax = self.atom_node_embedding(atom.x.float())
        x = torch.cat([rx, ax], 0).contiguous()
        NR, NA = st.NR, st.NA
        g = (0, 1, 2, 3) if self.differentiate_convolutions else (0, 0, 0, 0)
        cols = dict(e_cols=(0, ns), agg_cols=(ns, ns), nbr_cols=(2 * ns, ns))
        for layer in self.rec_emb_layers:
            # groups [rec, ar, atom, flipped ar] on the joint table [rec ; atom] (:303-314)
            segs = [Segment(st.rec_edges, st.rec_e, st.rec_sh, g[0], 0, NR, col_off=0),
                    Segment(st.ar_edges, st.ar_e, st.ar_sh, g[1], NR, NR + NA, col_off=0),
                    Segment(st.atom_edges, st.atom_e, st.atom_sh, g[2], NR, NR + NA, col_off=NR),
                    Segment(st.ra_edges, st.ra_e, st.ra_sh, g[3], 0, NR, col_off=NR)]
            x = layer.run(x, segs, NR + NA, ns, residual=x, **cols)
        st.rec_node_attr, st.atom_node_attr = x[:NR].contiguous(), x[NR:].contiguous()
        rec.cb200_static_aa = {"owner": id(self), "static": st}
        return st

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
        sigma_emb = self.timestep_emb_func(t["tr"]).float().contiguous()
        rec_sigma_emb = self.rec_sigma_embedding(sigma_emb).contiguous()
        NL, NR, NA = st.NL, st.NR, st.NA

        lig_x, lig_segments = self._ligand_embedding(st, lig_pos, sigma_emb, emb)
        rec_x = st.rec_node_attr.clone()
        rec_x[:, :ns] += rec_sigma_emb[st.rec_batch.long()]
        atom_x = st.atom_node_attr.clone()
        atom_x[:, :ns] += rec_sigma_emb[st.atom_batch.long()]
        if lig_x.shape[1] != rec_x.shape[1]:
            assert not self.embed_also_ligand
            lig_x = torch.nn.functional.pad(lig_x, (0, rec_x.shape[1] - lig_x.shape[1]))

        # ---- cross graphs (all_atom_score_model.py:587-622): ligand-residue (dynamic cutoff) and ligand-atom (5 A)
        if self.dynamic_max_cross:
            cutoff, r = (tr_sigma * 3 + 20).float().contiguous(), 1.0
        else:
            cutoff, r = None, float(self.cross_max_distance)
        lr = radius_edges(st.rec_pos, st.rec_ptr, lig_pos, st.lig_batch, r, 10000, st.cap_cross, cutoff=cutoff)
        rl = radius_edges_transposed(st.rec_pos, st.rec_batch, lig_pos, st.lig_ptr, r, st.cap_cross, cutoff=cutoff)
        la = radius_edges(st.atom_pos, st.atom_ptr, lig_pos, st.lig_batch, self.lig_max_radius, 10000, st.cap_la)
        al = radius_edges_transposed(st.atom_pos, st.atom_batch, lig_pos, st.lig_ptr, self.lig_max_radius, st.cap_la)
        lr_e, lr_sh = emb.lr(lr, lig_pos, st.rec_pos, st.lig_batch, sigma_emb, lmax)
        rl_e, rl_sh = emb.lr(rl, st.rec_pos, lig_pos, st.rec_batch, sigma_emb, lmax, sh_sign=-1.0)   # reuses SH(rec - lig)
        la_e, la_sh = emb.la(la, lig_pos, st.atom_pos, st.lig_batch, sigma_emb, lmax)
        al_e, al_sh = emb.la(al, st.atom_pos, lig_pos, st.atom_batch, sigma_emb, lmax, sh_sign=-1.0)  # reuses SH(atom - lig)

        x = torch.cat([lig_x, rec_x, atom_x], 0).contiguous()
        node_graph = torch.cat([st.lig_batch, st.rec_batch, st.atom_batch]).contiguous()
        o_rec, o_atom = NL, NL + NR
        g = tuple(range(9)) if self.differentiate_convolutions else (0,) * 9
        cols = dict(e_cols=(0, ns), agg_cols=(ns, ns), nbr_cols=(2 * ns, ns))
        n_layers = len(self.conv_layers)
        for l, layer in enumerate(self.conv_layers):
            lig_part = lig_segments(g[0], NL) + [Segment(lr, lr_e, lr_sh, g[1], 0, NL, col_off=o_rec),
                                                 Segment(la, la_e, la_sh, g[2], 0, NL, col_off=o_atom)]
            if l < n_layers - 1:
                # the last layer reads residue / atom rows only through the cross edges into the ligand: in the layer before
                # it, residues (atoms) without a cross edge (empty row of `rl` / `al`) skip their intra-receptor updates
                g_rec, g_atom = (rl, al) if l == n_layers - 2 else (None, None)
                segs = lig_part + [
                    Segment(st.rec_edges, st.rec_e, st.rec_sh, g[3], o_rec, o_atom, col_off=o_rec, e_post=rec_sigma_emb, gate=g_rec),
                    Segment(rl, rl_e, rl_sh, g[4], o_rec, o_atom, col_off=0),
                    Segment(st.ra_edges, st.ra_e, st.ra_sh, g[5], o_rec, o_atom, col_off=o_atom, e_post=rec_sigma_emb, gate=g_rec),
                    Segment(st.atom_edges, st.atom_e, st.atom_sh, g[6], o_atom, o_atom + NA, col_off=o_atom, e_post=rec_sigma_emb,
                            gate=g_atom),
                    Segment(al, al_e, al_sh, g[7], o_atom, o_atom + NA, col_off=0),
                    Segment(st.ar_edges, st.ar_e, st.ar_sh, g[8], o_atom, o_atom + NA, col_off=o_rec, e_post=rec_sigma_emb, gate=g_atom)]
                x = layer.run(x, segs, NL + NR + NA, ns, agg_graph=node_graph, residual=x, **cols)
            else:
                x = layer.run(x, lig_part, NL, ns, agg_graph=node_graph, residual=x, **cols)
        lig_x = x[:NL].contiguous()
        if self.confidence_mode:
            return self._confidence_head(lig_x, st)
        return self._score_heads(data, st, emb, lig_x, lig_pos, sigma_emb, tr_sigma, rot_sigma, tor_sigma)
