"""CPU restatement of the reference's first-party model code, functional, driven by a state_dict
with the reference's parameter names.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, in the reference formulation (materialised [E, weight_numel] radial-MLP output, gather +
index_add scatter-mean, per-call graph rebuild):
  models/layers.py:8-15                 FCBlock                         -> _fc
  models/tensor_layers.py:39-117        FasterTensorProduct             -> faster_tensor_product
  models/tensor_layers.py:195-217       TensorProductConvLayer.forward  -> tp_conv_layer
  models/score_model.py:18-41           AtomEncoder                     -> _atom_encoder
  models/score_model.py:667-677         GaussianSmearing                -> _smear
  models/score_model.py:282-449,492-664 CG TensorProductScoreModel      -> cg_forward
  models/all_atom_score_model.py:274-507,515-664 all-atom model         -> aa_forward
Pinned against the real reference executed under oracle/shims.py (tests/golden, tests/test_golden_oracle.py).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import o3
from .cluster import radius, radius_graph
from .scatter import scatter, scatter_mean

LIG_FEATURE_DIMS = [119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2]  # datasets/process_mols.py:95-112


# ----------------------------------------------------------------------------- hyper-parameters
def irrep_seq(ns, nv, reduce_pseudoscalars):
    last = nv if reduce_pseudoscalars else ns  # tensor_layers.py:20-26 (use_second_order_repr=False)
    return [f"{ns}x0e", f"{ns}x0e + {nv}x1o", f"{ns}x0e + {nv}x1o + {nv}x1e", f"{ns}x0e + {nv}x1o + {nv}x1e + {last}x0o"]


def hyper_from_args(args, confidence_mode=False):
    """The subset of utils/utils.py:225-283 that the forward pass depends on."""
    has = lambda k: k in args
    lm = any(has(k) and getattr(args, k) is not None for k in (
        "moad_esm_embeddings_path", "pdbbind_esm_embeddings_path", "pdbsidechain_esm_embeddings_path", "esm_embeddings_path"))
    return SimpleNamespace(
        all_atoms=bool(has("all_atoms") and args.all_atoms), ns=args.ns, nv=args.nv, num_conv_layers=args.num_conv_layers,
        sh_lmax=args.sh_lmax if has("sh_lmax") else 2, lig_max_radius=args.max_radius, rec_max_radius=30,
        cross_max_distance=args.cross_max_distance, center_max_distance=30, dynamic_max_cross=args.dynamic_max_cross,
        sigma_embed_dim=args.sigma_embed_dim, distance_embed_dim=args.distance_embed_dim,
        cross_distance_embed_dim=args.cross_distance_embed_dim, no_torsion=args.no_torsion,
        scale_by_sigma=args.scale_by_sigma, batch_norm=not args.no_batch_norm,
        num_prot_emb_layers=args.num_prot_emb_layers if has("num_prot_emb_layers") else 0,
        reduce_pseudoscalars=args.reduce_pseudoscalars if has("reduce_pseudoscalars") else False,
        embed_also_ligand=args.embed_also_ligand if has("embed_also_ligand") else False,
        differentiate_convolutions=not args.no_differentiate_convolutions if has("no_differentiate_convolutions") else True,
        fixed_center_conv=not args.not_fixed_center_conv if has("not_fixed_center_conv") else False,
        atom_confidence=args.atom_confidence_loss_weight > 0.0 if has("atom_confidence_loss_weight") else False,
        confidence_mode=confidence_mode, lm_dim=1280 if lm else 0,
        embedding_scale=args.embedding_scale if has("embedding_type") else 10000,
        in_lig_edge_features=4)


def to_double(sd, batch):
    """fp64 mode of the oracle: (state_dict, batch) with every floating tensor promoted to float64.  The neighbour
    searches keep their fp32 predicate (oracle/cluster.py casts positions to fp32), so the edge sets are the fp32 ones and
    only the arithmetic on them is carried out in double: the resulting outputs are the yardstick that tells fp32
    rounding noise (oracle-fp32 vs oracle-fp64) from a real discrepancy (product vs oracle-fp64)."""
    up = lambda v: v.double() if torch.is_tensor(v) and v.is_floating_point() else v
    sd64 = {k: up(v) for k, v in sd.items()}
    for st in batch._stores.values():
        for k, v in list(st._d.items()):
            st._d[k] = {kk: up(u) for kk, u in v.items()} if isinstance(v, dict) else up(v)
    for k, v in list(batch._g.items()):
        if not k.startswith("_"):
            batch._g[k] = {kk: up(u) for kk, u in v.items()} if isinstance(v, dict) else up(v)
    return sd64, batch


def sinusoidal(t, dim, scale):
    """utils/diffusion_utils.py:99-110 applied to scale*t."""
    half = dim // 2
    dt = t.dtype if torch.is_tensor(t) and t.dtype == torch.float64 else torch.float32      # fp64 mode: see to_double()
    freq = torch.exp(torch.arange(half, dtype=dt) * -(math.log(10000) / (half - 1)))
    ang = (scale * t).to(dt)[:, None] * freq[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


# ----------------------------------------------------------------------------- small pieces
def _lin(sd, name, x):
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"], b)


def _fc(sd, prefix, x):
    """Linear - ReLU - (Dropout) - Linear; Sequential slots 0 and 3."""
    return _lin(sd, prefix + ".3", torch.relu(_lin(sd, prefix + ".0", x)))


def _smear(sd, name, d):
    off = sd[name + ".offset"]
    coeff = -0.5 / (off[1] - off[0]).item() ** 2
    return torch.exp(coeff * torch.pow(d.view(-1, 1) - off.view(1, -1), 2))


def _atom_encoder(sd, prefix, x, n_cat):
    emb = 0
    for i in range(n_cat):
        emb = emb + sd[f"{prefix}.atom_embedding_list.{i}.weight"][x[:, i].long()]
    if x.shape[1] > n_cat:
        emb = _lin(sd, prefix + ".additional_features_embedder", torch.cat([emb, x[:, n_cat:]], dim=1))
    return emb


def _sh(lmax, vec):
    return o3.spherical_harmonics(list(range(lmax + 1)), vec, normalize=True, normalization="component")


def _e3nn_batch_norm(sd, prefix, irreps, x):
    """Eval-mode e3nn BatchNorm."""
    out, ix, im, iv = [], 0, 0, 0
    w, b, rm, rv = (sd[prefix + k] for k in (".weight", ".bias", ".running_mean", ".running_var"))
    for mul, ir in o3.Irreps(irreps):
        d = ir.dim
        f = x[:, ix: ix + mul * d].reshape(-1, mul, d)
        ix += mul * d
        if ir.is_scalar():
            f = f - rm[im: im + mul].reshape(1, mul, 1)
        f = f * ((rv[iv: iv + mul] + 1e-5).pow(-0.5) * w[iv: iv + mul]).reshape(1, mul, 1)
        if ir.is_scalar():
            f = f + b[im: im + mul].reshape(1, mul, 1)
            im += mul
        iv += mul
        out.append(f.reshape(-1, mul * d))
    return torch.cat(out, -1)


def faster_tensor_product(in_irreps, out_irreps, x, sh, w):
    """Closed-form lmax=1 product with the weight layout of FasterTensorProduct."""
    ins, outs = o3.Irreps(in_irreps), o3.Irreps(out_irreps)
    part = {}
    for (mul, ir), sl in zip(ins, ins.slices()):
        v = x[:, sl]
        part[str(ir)] = v.reshape(-1, mul, 3) if ir.l == 1 else v
    mo = {str(ir): mul for mul, ir in outs}
    y0, y1 = sh[:, :1], sh[:, 1:4]
    cross = lambda a: torch.linalg.cross(a, y1[:, None, :].expand_as(a), dim=-1)
    dot = lambda a: (a * y1[:, None, :]).sum(-1)
    mid = {"0e": [], "1o": [], "1e": [], "0o": []}
    if "0e" in part:
        mid["0e"].append(part["0e"] * y0)
        mid["1o"].append(part["0e"][:, :, None] * y1[:, None, :])
    if "1o" in part:
        mid["0e"].append(dot(part["1o"]) / math.sqrt(3))
        mid["1o"].append(part["1o"] * y0[:, :, None])
        mid["1e"].append(cross(part["1o"]) / math.sqrt(2))
    if "1e" in part:
        mid["1o"].append(cross(part["1e"]) / math.sqrt(2))
        mid["1e"].append(part["1e"] * y0[:, :, None])
        mid["0o"].append(dot(part["1e"]) / math.sqrt(3))
    if "0o" in part:
        mid["1e"].append(part["0o"][:, :, None] * y1[:, None, :])
        mid["0o"].append(part["0o"] * y0)
    nin = {k: (part[k].shape[1] if k in part else 0) for k in mid}
    fan = {"0e": nin["0e"] + nin["1o"], "1o": nin["0e"] + nin["1o"] + nin["1e"],
           "1e": nin["1o"] + nin["1e"] + nin["0o"], "0o": nin["1e"] + nin["0o"]}
    res, start = {}, 0
    for k in ("0e", "1o", "1e", "0o"):
        m = mo.get(k, 0)
        if m:
            wk = w[:, start: start + fan[k] * m].reshape(-1, fan[k], m) / math.sqrt(fan[k])
            f = torch.cat(mid[k], dim=1)
            if k[0] == "0":
                res[k] = torch.einsum("zi,zim->zm", f, wk)
            else:
                res[k] = torch.einsum("zic,zim->zmc", f, wk).reshape(f.shape[0], -1)
        start += fan[k] * m
    return torch.cat([res[str(ir)] for _, ir in outs], dim=-1)


def faster_weight_numel(in_irreps, out_irreps):
    nin = {"0e": 0, "1o": 0, "1e": 0, "0o": 0}
    nout = dict(nin)
    for mul, ir in o3.Irreps(in_irreps):
        nin[str(ir)] = mul
    for mul, ir in o3.Irreps(out_irreps):
        nout[str(ir)] = mul
    return ((nin["0e"] + nin["1o"]) * nout["0e"] + (nin["0e"] + nin["1o"] + nin["1e"]) * nout["1o"] +
            (nin["1o"] + nin["1e"] + nin["0o"]) * nout["1e"] + (nin["1e"] + nin["0o"]) * nout["0o"])


def tp_conv_layer(sd, prefix, in_irreps, sh_irreps, out_irreps, faster, edge_groups, residual, batch_norm,
                  node_attr, edge_index, edge_attr, edge_sh, out_nodes=None):
    """TensorProductConvLayer.forward: gather from edge_index[1], mean over edge_index[0]."""
    out_size = o3.Irreps(out_irreps).dim
    if edge_index.shape[1] == 0:
        out = torch.zeros((node_attr.shape[0], out_size), dtype=node_attr.dtype)
    else:
        agg, nbr = edge_index
        if edge_groups == 1:
            w = _fc(sd, prefix + ".fc", edge_attr)
        else:
            w = torch.cat([_fc(sd, f"{prefix}.fc.{g}", edge_attr[g]) for g in range(edge_groups)], dim=0)
        if faster:
            tp = faster_tensor_product(in_irreps, out_irreps, node_attr[nbr], edge_sh, w)
        else:
            tp = o3.FullyConnectedTensorProduct(in_irreps, sh_irreps, out_irreps)(node_attr[nbr], edge_sh, w)
        out = scatter(tp, agg, dim=0, dim_size=out_nodes or node_attr.shape[0], reduce="mean")
        if batch_norm:
            out = _e3nn_batch_norm(sd, prefix + ".batch_norm", out_irreps, out)
    if residual:
        out = out + F.pad(node_attr, (0, out.shape[-1] - node_attr.shape[-1]))
    return out


# ----------------------------------------------------------------------------- CG score model
def _lig_graph(sd, hp, data, sigma_emb_nodes):
    lig, ll = data["ligand"], data["ligand", "ligand"]
    rad = radius_graph(lig.pos, hp.lig_max_radius, lig.batch)
    ei = torch.cat([ll.edge_index, rad], 1).long()
    ea = torch.cat([ll.edge_attr, torch.zeros(rad.shape[-1], hp.in_lig_edge_features)], 0)
    ea = torch.cat([ea, sigma_emb_nodes[ei[0]]], 1)
    node_attr = torch.cat([lig.x, sigma_emb_nodes], 1)
    vec = lig.pos[ei[1]] - lig.pos[ei[0]]
    ea = torch.cat([ea, _smear(sd, "lig_distance_expansion", vec.norm(dim=-1))], 1)
    return node_attr, ei, ea, _sh(hp.sh_lmax, vec)


def _conv_args(hp, i, groups, faster=None):
    seq = irrep_seq(hp.ns, hp.nv, hp.reduce_pseudoscalars)
    return dict(in_irreps=seq[min(i, 3)], sh_irreps=o3.Irreps.spherical_harmonics(hp.sh_lmax), out_irreps=seq[min(i + 1, 3)],
                faster=(hp.sh_lmax == 1) if faster is None else faster, edge_groups=groups, residual=True,
                batch_norm=hp.batch_norm)


def cg_forward(sd, hp, data, t_to_sigma, so3_score_norm, torus_score_norm):
    """models/score_model.py:333-449 (eval mode)."""
    ns, lmax = hp.ns, hp.sh_lmax
    lig, rec = data["ligand"], data["receptor"]
    ll, rr = data["ligand", "ligand"], data["receptor", "receptor"]
    B = data.num_graphs
    ct = data.complex_t
    if not hp.confidence_mode:
        tr_sigma, rot_sigma, tor_sigma = t_to_sigma(ct["tr"], ct["rot"], ct["tor"])
    else:
        tr_sigma, rot_sigma, tor_sigma = ct["tr"], ct["rot"], ct["tor"]
    emb_t = lambda t: sinusoidal(t, hp.sigma_embed_dim, hp.embedding_scale)

    # receptor embedding (:297-326)
    rei = rr.edge_index
    rvec = rec.pos[rei[1].long()] - rec.pos[rei[0].long()]
    rec_edge_attr = _fc(sd, "rec_edge_embedding", _smear(sd, "rec_distance_expansion", rvec.norm(dim=-1)))
    rec_sh = _sh(lmax, rvec)
    rec_x = _atom_encoder(sd, "rec_node_embedding", rec.x, 1)
    for l in range(hp.num_prot_emb_layers):
        ea = torch.cat([rec_edge_attr, rec_x[rei[0], :ns], rec_x[rei[1], :ns]], -1)
        rec_x = tp_conv_layer(sd, f"rec_emb_layers.{l}", node_attr=rec_x, edge_index=rei, edge_attr=ea, edge_sh=rec_sh,
                              **_conv_args(hp, l, 1))
    rec_sigma = _fc(sd, "rec_sigma_embedding", emb_t(ct["tr"]))
    rec_x = rec_x + 0
    rec_x[:, :ns] = rec_x[:, :ns] + rec_sigma[rec.batch]
    rec_edge_attr = rec_edge_attr + rec_sigma[rec.batch[rei[0]]]

    # ligand embedding (:282-295)
    node_sigma = emb_t(lig.node_t["tr"]) if "node_t" in lig else emb_t(ct["tr"])[lig.batch]
    lig_in, lei, lea, lsh = _lig_graph(sd, hp, data, node_sigma)
    lig_x = _atom_encoder(sd, "lig_node_embedding", lig_in, len(LIG_FEATURE_DIMS))
    lig_edge_attr = _fc(sd, "lig_edge_embedding", lea)
    if hp.embed_also_ligand:
        for l in range(hp.num_prot_emb_layers):
            ea = torch.cat([lig_edge_attr, lig_x[lei[0], :ns], lig_x[lei[1], :ns]], -1)
            lig_x = tp_conv_layer(sd, f"lig_emb_layers.{l}", node_attr=lig_x, edge_index=lei, edge_attr=ea, edge_sh=lsh,
                                  **_conv_args(hp, l, 1))

    # cross graph (:346-352, 564-587)
    if hp.dynamic_max_cross:
        cut = (tr_sigma * 3 + 20).unsqueeze(1)
        cei = radius(rec.pos / cut[rec.batch], lig.pos / cut[lig.batch], 1, rec.batch, lig.batch, max_num_neighbors=10000)
    else:
        cei = radius(rec.pos, lig.pos, hp.cross_max_distance, rec.batch, lig.batch, max_num_neighbors=10000)
    cvec = rec.pos[cei[1].long()] - lig.pos[cei[0].long()]
    cea = torch.cat([node_sigma[cei[0].long()], _smear(sd, "cross_distance_expansion", cvec.norm(dim=-1))], 1)
    cea = _fc(sd, "cross_edge_embedding", cea)
    csh, csh_rev = _sh(lmax, cvec), _sh(lmax, -cvec)

    # joint graph + conv stack (:354-374)
    nl = len(lig_x)
    x = torch.cat([lig_x, rec_x], dim=0)
    cei = cei.clone()
    cei[1] = cei[1] + nl
    ei = torch.cat([lei, cei, rei + nl, torch.flip(cei, dims=[0])], dim=1)
    ea_all = torch.cat([lig_edge_attr, cea, rec_edge_attr, cea], dim=0)
    sh_all = torch.cat([lsh, csh, rec_sh, csh_rev], dim=0)
    s1, s2, s3 = lei.shape[1], lei.shape[1] + cei.shape[1], lei.shape[1] + cei.shape[1] + rei.shape[1]
    n_conv = hp.num_conv_layers
    for l in range(n_conv):
        idx = hp.num_prot_emb_layers + l
        if l < n_conv - 1:
            ea = torch.cat([ea_all, x[ei[0], :ns], x[ei[1], :ns]], -1)
            groups = 4 if hp.differentiate_convolutions else 1
            if groups > 1:
                ea = [ea[:s1], ea[s1:s2], ea[s2:s3], ea[s3:]]
            x = tp_conv_layer(sd, f"conv_layers.{l}", node_attr=x, edge_index=ei, edge_attr=ea, edge_sh=sh_all,
                              **_conv_args(hp, idx, groups))
        else:
            ea = torch.cat([ea_all[:s2], x[ei[0, :s2], :ns], x[ei[1, :s2], :ns]], -1)
            groups = 2 if hp.differentiate_convolutions else 1
            if groups > 1:
                ea = [ea[:s1], ea[s1:s2]]
            x = tp_conv_layer(sd, f"conv_layers.{l}", node_attr=x, edge_index=ei[:, :s2], edge_attr=ea, edge_sh=sh_all[:s2],
                              **_conv_args(hp, idx, groups))
    lig_x = x[:nl]
    out_irreps = irrep_seq(hp.ns, hp.nv, hp.reduce_pseudoscalars)[min(hp.num_prot_emb_layers + n_conv, 3)]

    if hp.confidence_mode:
        return _confidence_head(sd, hp, lig_x, lig.batch, B)
    return _score_heads(sd, hp, data, lig_x, node_sigma, out_irreps, emb_t, tr_sigma, rot_sigma, tor_sigma,
                        so3_score_norm, torus_score_norm)


def _score_heads(sd, hp, data, lig_x, node_sigma, out_irreps, emb_t, tr_sigma, rot_sigma, tor_sigma,
                 so3_score_norm, torus_score_norm):
    """Translation / rotation / torsion heads shared by the CG and all-atom score models
    (score_model.py:394-449 == all_atom_score_model.py:457-507)."""
    ns, lmax = hp.ns, hp.sh_lmax
    lig, ll = data["ligand"], data["ligand", "ligand"]
    B, ct = data.num_graphs, data.complex_t
    # translation / rotation head (:394-420)
    cidx = torch.stack([lig.batch, torch.arange(len(lig.batch))])
    center = torch.zeros((B, 3), dtype=lig.pos.dtype).index_add_(0, lig.batch, lig.pos) / torch.bincount(lig.batch, minlength=B).unsqueeze(1)
    cv = lig.pos[cidx[1]] - center[cidx[0]]
    ca = torch.cat([_smear(sd, "center_distance_expansion", cv.norm(dim=-1)), node_sigma[cidx[1]]], 1)
    ca = _fc(sd, "center_edge_embedding", ca)
    ca = torch.cat([ca, lig_x[cidx[1] if hp.fixed_center_conv else cidx[0], :ns]], -1)
    gp = tp_conv_layer(sd, "final_conv", in_irreps=out_irreps, sh_irreps=o3.Irreps.spherical_harmonics(lmax),
                       out_irreps="2x1o + 2x1e", faster=False, edge_groups=1, residual=False, batch_norm=hp.batch_norm,
                       node_attr=lig_x, edge_index=cidx, edge_attr=ca, edge_sh=_sh(lmax, cv), out_nodes=B)
    tr = gp[:, :3] + gp[:, 6:9]
    rot = gp[:, 3:6] + gp[:, 9:]
    gse = emb_t(ct["tr"])
    head = lambda p, v: _lin(sd, p + ".3", torch.relu(_lin(sd, p + ".0", v)))  # Linear, Dropout, ReLU, Linear
    trn = torch.linalg.vector_norm(tr, dim=1).unsqueeze(1)
    tr = tr / trn * head("tr_final_layer", torch.cat([trn, gse], dim=1))
    rotn = torch.linalg.vector_norm(rot, dim=1).unsqueeze(1)
    rot = rot / rotn * head("rot_final_layer", torch.cat([rotn, gse], dim=1))
    if hp.scale_by_sigma:
        tr = tr / tr_sigma.unsqueeze(1)
        rot = rot * so3_score_norm(rot_sigma.cpu()).unsqueeze(1)
    if hp.no_torsion or lig.edge_mask.sum() == 0:
        return tr, rot, torch.empty(0), None

    # torsion head (:432-448, 650-664)
    bonds = ll.edge_index[:, lig.edge_mask].long()
    bpos = (lig.pos[bonds[0]] + lig.pos[bonds[1]]) / 2
    tei = radius(lig.pos, bpos, hp.lig_max_radius, batch_x=lig.batch, batch_y=lig.batch[bonds[0]])
    tv = lig.pos[tei[1]] - bpos[tei[0]]
    tea = _fc(sd, "final_edge_embedding", _smear(sd, "lig_distance_expansion", tv.norm(dim=-1)))
    bvec = lig.pos[bonds[1]] - lig.pos[bonds[0]]
    battr = lig_x[bonds[0]] + lig_x[bonds[1]]
    bsh = o3.spherical_harmonics("2e", bvec, normalize=True, normalization="component")
    ftp = o3.FullTensorProduct(o3.Irreps.spherical_harmonics(lmax), "2e")
    tsh = ftp(_sh(lmax, tv), bsh[tei[0]])
    tea = torch.cat([tea, lig_x[tei[1], :ns], battr[tei[0], :ns]], -1)
    tor = tp_conv_layer(sd, "tor_bond_conv", in_irreps=out_irreps, sh_irreps=ftp.irreps_out,
                        out_irreps=f"{ns}x0o + {ns}x0e", faster=False, edge_groups=1, residual=False,
                        batch_norm=hp.batch_norm, node_attr=lig_x, edge_index=tei, edge_attr=tea, edge_sh=tsh,
                        out_nodes=int(lig.edge_mask.sum()))
    tor = F.linear(torch.tanh(F.linear(tor, sd["tor_final_layer.0.weight"])), sd["tor_final_layer.3.weight"]).squeeze(1)
    if hp.scale_by_sigma:
        edge_sigma = tor_sigma[lig.batch][ll.edge_index[0]][lig.edge_mask]
        tor = tor * torch.sqrt(torch.tensor(torus_score_norm(edge_sigma.cpu().numpy())).to(tor.dtype))
    return tr, rot, tor, None


def _bn1d(sd, p, x):
    return (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + 1e-5) * sd[p + ".weight"] + sd[p + ".bias"]


def _conf_mlp(sd, p, x):
    """Linear, BatchNorm1d, ReLU, Dropout, Linear, BatchNorm1d, ReLU, Dropout, Linear (slots 0,1,4,5,8)."""
    h = torch.relu(_bn1d(sd, p + ".1", _lin(sd, p + ".0", x))) if p + ".1.weight" in sd else torch.relu(_lin(sd, p + ".0", x))
    h = torch.relu(_bn1d(sd, p + ".5", _lin(sd, p + ".4", h))) if p + ".5.weight" in sd else torch.relu(_lin(sd, p + ".4", h))
    return _lin(sd, p + ".8", h)


def _confidence_head(sd, hp, lig_x, lig_batch, B):
    ns = hp.ns
    if hp.num_conv_layers + hp.num_prot_emb_layers >= 3:
        tail = hp.nv if hp.reduce_pseudoscalars else ns
        s = torch.cat([lig_x[:, :ns], lig_x[:, -tail:]], dim=1)
    else:
        s = lig_x[:, :ns]
    if hp.atom_confidence:
        s = _conf_mlp(sd, "atom_confidence_predictor", s)
        atom_conf, s = s[:, :1], s[:, 1:]
    else:
        atom_conf = torch.zeros((len(lig_x),))
    conf = _conf_mlp(sd, "confidence_predictor", scatter_mean(s, lig_batch, dim=0, dim_size=B)).squeeze(dim=-1)
    return conf, atom_conf


# ----------------------------------------------------------------------------- all-atom model
def aa_forward(sd, hp, data, t_to_sigma, so3_score_norm, torus_score_norm):
    """models/all_atom_score_model.py:363-507 (eval mode; the shipped confidence model is this in confidence mode)."""
    ns, lmax = hp.ns, hp.sh_lmax
    lig, rec, atom = data["ligand"], data["receptor"], data["atom"]
    rr, aa, ar = data["receptor", "receptor"], data["atom", "atom"], data["atom", "receptor"]
    B, ct = data.num_graphs, data.complex_t
    if not hp.confidence_mode:
        tr_sigma, rot_sigma, tor_sigma = t_to_sigma(ct["tr"], ct["rot"], ct["tor"])
    else:
        tr_sigma, rot_sigma, tor_sigma = ct["tr"], ct["rot"], ct["tor"]
    emb_t = lambda t: sinusoidal(t, hp.sigma_embed_dim, hp.embedding_scale)
    seq = irrep_seq(hp.ns, hp.nv, hp.reduce_pseudoscalars)

    # receptor / atom embedding (:274-329)
    rei, aei, arei = rr.edge_index.long(), aa.edge_index.long(), ar.edge_index.long()
    rvec = rec.pos[rei[1]] - rec.pos[rei[0]]
    rec_ea = _fc(sd, "rec_edge_embedding", _smear(sd, "rec_distance_expansion", rvec.norm(dim=-1)))
    rec_sh = _sh(lmax, rvec)
    avec = atom.pos[aei[1]] - atom.pos[aei[0]]
    atom_ea = _fc(sd, "atom_edge_embedding", _smear(sd, "lig_distance_expansion", avec.norm(dim=-1)))
    atom_sh = _sh(lmax, avec)
    arvec = rec.pos[arei[1]] - atom.pos[arei[0]]
    ar_ea = _fc(sd, "ar_edge_embedding", _smear(sd, "rec_distance_expansion", arvec.norm(dim=-1)))
    ar_sh = _sh(lmax, arvec)
    rec_x = _atom_encoder(sd, "rec_node_embedding", rec.x, 1)
    atom_x = _atom_encoder(sd, "atom_node_embedding", atom.x, 4)
    nr = len(rec_x)
    if hp.num_prot_emb_layers > 0:
        x = torch.cat([rec_x, atom_x], dim=0)
        are = arei.clone()
        are[0] = are[0] + nr
        ei = torch.cat([rei, are, aei + nr, torch.flip(are, dims=[0])], dim=1)
        ea_all = torch.cat([rec_ea, ar_ea, atom_ea, ar_ea], dim=0)
        sh_all = torch.cat([rec_sh, ar_sh, atom_sh, ar_sh], dim=0)
        s1, s2, s3 = rei.shape[1], rei.shape[1] + are.shape[1], rei.shape[1] + are.shape[1] + aei.shape[1]
        for l in range(hp.num_prot_emb_layers):
            ea = torch.cat([ea_all, x[ei[0], :ns], x[ei[1], :ns]], -1)
            groups = 4 if hp.differentiate_convolutions else 1
            if groups > 1:
                ea = [ea[:s1], ea[s1:s2], ea[s2:s3], ea[s3:]]
            x = tp_conv_layer(sd, f"rec_emb_layers.{l}", node_attr=x, edge_index=ei, edge_attr=ea, edge_sh=sh_all,
                              **_conv_args(hp, l, groups))
        rec_x, atom_x = x[:nr], x[nr:]
    rec_sigma = _fc(sd, "rec_sigma_embedding", emb_t(ct["tr"]))
    rec_x, atom_x = rec_x + 0, atom_x + 0
    rec_x[:, :ns] = rec_x[:, :ns] + rec_sigma[rec.batch]
    atom_x[:, :ns] = atom_x[:, :ns] + rec_sigma[atom.batch]
    rec_ea = rec_ea + rec_sigma[rec.batch[rei[0]]]
    atom_ea = atom_ea + rec_sigma[atom.batch[aei[0]]]
    ar_ea = ar_ea + rec_sigma[atom.batch[arei[0]]]

    # ligand (:345-356)
    node_sigma = emb_t(lig.node_t["tr"]) if "node_t" in lig else emb_t(ct["tr"])[lig.batch]
    lig_in, lei, lea, lsh = _lig_graph(sd, hp, data, node_sigma)
    lig_x = _atom_encoder(sd, "lig_node_embedding", lig_in, len(LIG_FEATURE_DIMS))
    lig_ea = _fc(sd, "lig_edge_embedding", lea)
    if hp.embed_also_ligand:
        for l in range(hp.num_prot_emb_layers):
            ea = torch.cat([lig_ea, lig_x[lei[0], :ns], lig_x[lei[1], :ns]], -1)
            lig_x = tp_conv_layer(sd, f"lig_emb_layers.{l}", node_attr=lig_x, edge_index=lei, edge_attr=ea, edge_sh=lsh,
                                  **_conv_args(hp, l, 1))
    else:
        lig_x = F.pad(lig_x, (0, rec_x.shape[-1] - lig_x.shape[-1]))

    # cross graphs (:587-622)
    if hp.dynamic_max_cross:
        cut = (tr_sigma * 3 + 20).unsqueeze(1)
        lrei = radius(rec.pos / cut[rec.batch], lig.pos / cut[lig.batch], 1, rec.batch, lig.batch, max_num_neighbors=10000)
    else:
        lrei = radius(rec.pos, lig.pos, hp.cross_max_distance, rec.batch, lig.batch, max_num_neighbors=10000)
    lrvec = rec.pos[lrei[1]] - lig.pos[lrei[0]]
    lr_ea = _fc(sd, "lr_edge_embedding", torch.cat([node_sigma[lrei[0]], _smear(sd, "cross_distance_expansion", lrvec.norm(dim=-1))], 1))
    lr_sh = _sh(lmax, lrvec)
    laei = radius(atom.pos, lig.pos, hp.lig_max_radius, atom.batch, lig.batch, max_num_neighbors=10000)
    lavec = atom.pos[laei[1]] - lig.pos[laei[0]]
    la_ea = _fc(sd, "la_edge_embedding", torch.cat([node_sigma[laei[0]], _smear(sd, "lig_distance_expansion", lavec.norm(dim=-1))], 1))
    la_sh = _sh(lmax, lavec)

    # joint graph: nodes [lig ; rec ; atom], nine edge groups (:396-418)
    nl = len(lig_x)
    x = torch.cat([lig_x, rec_x, atom_x], dim=0)
    rei2, aei2 = rei + nl, aei + nl + nr
    lrei2, laei2, arei2 = lrei.clone(), laei.clone(), arei.clone()
    lrei2[1] += nl
    laei2[1] += nl + nr
    arei2[0] += nl + nr
    arei2[1] += nl
    parts = [(lei, lig_ea, lsh), (lrei2, lr_ea, lr_sh), (laei2, la_ea, la_sh), (rei2, rec_ea, rec_sh),
             (torch.flip(lrei2, dims=[0]), lr_ea, lr_sh), (torch.flip(arei2, dims=[0]), ar_ea, ar_sh),
             (aei2, atom_ea, atom_sh), (torch.flip(laei2, dims=[0]), la_ea, la_sh), (arei2, ar_ea, ar_sh)]
    ei = torch.cat([p[0] for p in parts], dim=1)
    ea_all = torch.cat([p[1] for p in parts], dim=0)
    sh_all = torch.cat([p[2] for p in parts], dim=0)
    cuts = np.cumsum([p[0].shape[1] for p in parts]).tolist()
    n_conv = hp.num_conv_layers
    for l in range(n_conv):
        idx = hp.num_prot_emb_layers + l
        if l < n_conv - 1:
            ea = torch.cat([ea_all, x[ei[0], :ns], x[ei[1], :ns]], -1)
            groups = 9 if hp.differentiate_convolutions else 1
            if groups > 1:
                ea = [ea[a:b] for a, b in zip([0] + cuts[:-1], cuts)]
            x = tp_conv_layer(sd, f"conv_layers.{l}", node_attr=x, edge_index=ei, edge_attr=ea, edge_sh=sh_all,
                              **_conv_args(hp, idx, groups))
        else:
            s3 = cuts[2]
            ea = torch.cat([ea_all[:s3], x[ei[0, :s3], :ns], x[ei[1, :s3], :ns]], -1)
            groups = 3 if hp.differentiate_convolutions else 1
            if groups > 1:
                ea = [ea[:cuts[0]], ea[cuts[0]:cuts[1]], ea[cuts[1]:s3]]
            x = tp_conv_layer(sd, f"conv_layers.{l}", node_attr=x, edge_index=ei[:, :s3], edge_attr=ea, edge_sh=sh_all[:s3],
                              **_conv_args(hp, idx, groups))
    lig_x = x[:nl]
    if hp.confidence_mode:
        return _confidence_head(sd, hp, lig_x, lig.batch, B)
    out_irreps = seq[min(hp.num_prot_emb_layers + n_conv, 3)]
    return _score_heads(sd, hp, data, lig_x, node_sigma, out_irreps, emb_t, tr_sigma, rot_sigma, tor_sigma,
                        so3_score_norm, torus_score_norm)
