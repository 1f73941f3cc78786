"""Noise schedules, time embeddings, `set_time` and the batched pose update (utils/diffusion_utils.py).

`modify_conformer_batch` keeps the reference signature but is one cb200 kernel launch (K4) instead of
R sequential bond-rotation micro-kernels plus a batched SVD (diffusion_utils.py:60-78,
torsion.py:75-90, geometry.py:246-276).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib


def t_to_sigma_individual(t, schedule_type, sigma_min, sigma_max, schedule_k=10, schedule_m=0.4):
    if schedule_type == "exponential":
        return sigma_min ** (1 - t) * sigma_max ** t
    if schedule_type == "sigmoid":
        s = lambda v: 1 / (1 + np.e ** (-schedule_k * (v - schedule_m)))
        return (s(t) - s(0)) / (s(1) - s(0)) * (sigma_max - sigma_min) + sigma_min
    raise ValueError(schedule_type)


def t_to_sigma(t_tr, t_rot, t_tor, args):
    """sigma(t) = sigma_min^(1-t) * sigma_max^t per component (diffusion_utils.py:28-32)."""
    return (t_to_sigma_individual(t_tr, "exponential", args.tr_sigma_min, args.tr_sigma_max),
            t_to_sigma_individual(t_rot, "exponential", args.rot_sigma_min, args.rot_sigma_max),
            t_to_sigma_individual(t_tor, "exponential", args.tor_sigma_min, args.tor_sigma_max))


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    assert len(timesteps.shape) == 1
    half = embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -(math.log(max_positions) / (half - 1)))
    ang = timesteps.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1), mode="constant")
    return emb


class GaussianFourierProjection(nn.Module):
    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size // 2) * scale, requires_grad=False)

    def forward(self, x):
        proj = x[:, None] * self.W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    if embedding_type == "sinusoidal":
        return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)
    if embedding_type == "fourier":
        return GaussianFourierProjection(embedding_size=embedding_dim, scale=embedding_scale)
    raise NotImplementedError(embedding_type)


def get_t_schedule(sigma_schedule, inference_steps, inf_sched_alpha=1, inf_sched_beta=1, t_max=1):
    if sigma_schedule == "expbeta":
        from scipy.stats import beta
        lin_max = beta.cdf(t_max, a=inf_sched_alpha, b=inf_sched_beta)
        c = np.linspace(lin_max, 0, inference_steps + 1)[:-1]
        return beta.ppf(c, a=inf_sched_alpha, b=inf_sched_beta)
    raise Exception()


def set_time(complex_graphs, t, t_tr, t_rot, t_tor, batchsize, all_atoms, asyncronous_noise_schedule, device,
             include_miscellaneous_atoms=False, materialize_node_t=False):
    """Records the diffusion time on the batch (diffusion_utils.py:150-179).

    The cb200 models only read `complex_t` ([B] per component) plus the host copy `complex_t_host`,
    which lets them index the so3/torus tables without a device->host round trip; the per-node
    constant tensors of the reference are materialised only on request."""
    assert not asyncronous_noise_schedule, "asyncronous_noise_schedule is outside the shipped configurations"
    ones = torch.ones(batchsize, device=device)
    complex_graphs.complex_t = {"tr": t_tr * ones, "rot": t_rot * ones, "tor": t_tor * ones}
    if all(isinstance(v, (int, float, np.floating, np.integer)) for v in (t_tr, t_rot, t_tor)):
        complex_graphs.complex_t_host = {"tr": float(t_tr), "rot": float(t_rot), "tor": float(t_tor)}
    else:
        complex_graphs.complex_t_host = None
    if materialize_node_t:
        kinds = ["ligand", "receptor"] + (["atom"] if all_atoms else [])
        for k in kinds:
            n = complex_graphs[k].num_nodes
            complex_graphs[k].node_t = {"tr": t_tr * torch.ones(n, device=device), "rot": t_rot * torch.ones(n, device=device),
                                        "tor": t_tor * torch.ones(n, device=device)}


class LigandTopology:
    """Device-side tables of the shared ligand topology of a sampling batch (sampling.py:81,
    diffusion_utils.py:62-64): rotatable bonds (u, v) in edge order and the [R, N] rotation masks."""

    def __init__(self, data, mask_rotate, device, validate=True):
        """`validate=False` skips the device-side orientation check (two host reads): sampling() runs the same check once on
        the host graph instead (`check_rotation_masks`)."""
        B = int(data.num_graphs)
        lig, ll = data["ligand"], data["ligand", "ligand"]
        self.B = B
        self.N = int(lig.num_nodes) // B
        M = int(ll.num_edges) // B
        edge_index, edge_mask = ll.edge_index[:, :M], lig.edge_mask[:M].bool()
        mr = torch.as_tensor(np.asarray(mask_rotate)) if not torch.is_tensor(mask_rotate) else mask_rotate
        n_tor_h = data._g.get("_n_tor_h") if isinstance(getattr(data, "_g", None), dict) and not data._g.get("_slices_stale") else None
        if n_tor_h is not None and len(n_tor_h) == B:
            # rotatable-bond count known on the host (collate): the masked selection needs no device read
            idx = torch.nonzero_static(edge_mask, size=int(n_tor_h[0])).reshape(-1)
            self.bond_uv = edge_index.index_select(1, idx).t().to(torch.int32).contiguous().to(device)
        else:
            self.bond_uv = edge_index.t()[edge_mask].to(torch.int32).contiguous().to(device)
        self.R = int(self.bond_uv.shape[0])
        self.mask_rotate = mr.to(torch.uint8).contiguous().to(device)
        if self.R > 0:
            assert tuple(self.mask_rotate.shape) == (self.R, self.N), "mask_rotate does not match the topology"
            if validate:
                u, v = self.bond_uv[:, 0].long(), self.bond_uv[:, 1].long()
                idx = torch.arange(self.R, device=device)
                # torsion.py:81-82: v must be on the rotating side, u on the fixed side
                assert not bool(self.mask_rotate[idx, u].any()) and bool(self.mask_rotate[idx, v].all())


def check_rotation_masks(graph, mask_rotate):
    """The orientation check of LigandTopology on ONE host graph (torsion.py:81-82: for every rotatable bond (u, v) in edge
    order, v is on the rotating side and u on the fixed side).  No-op for device-resident graphs."""
    ei, em = graph["ligand", "ligand"].edge_index, graph["ligand"].edge_mask
    if not (torch.is_tensor(ei) and ei.device.type == "cpu" and torch.is_tensor(em) and em.device.type == "cpu"):
        return False
    mr = np.asarray(mask_rotate.cpu() if torch.is_tensor(mask_rotate) else mask_rotate).astype(bool)
    uv = ei.t()[em.bool()].numpy()
    if uv.shape[0] == 0:
        return True
    assert mr.shape[0] == uv.shape[0], "mask_rotate does not match the topology"
    r = np.arange(uv.shape[0])
    assert not mr[r, uv[:, 0]].any() and mr[r, uv[:, 1]].all(), "mask_rotate orientation (torsion.py:81-82)"
    return True


def sde_step(pos, topo: LigandTopology, tr_score, rot_score, tor_score, coeffs, z_tr=None, z_rot=None, z_tor=None):
    """In-place pose update: perturbation = c_score * score + c_noise * z, then rigid move, bond
    rotations, Kabsch re-alignment.  coeffs = (c_tr_s, c_tr_n, c_rot_s, c_rot_n, c_tor_s, c_tor_n) as host
    floats, or a float32 CUDA tensor with those six values (read by the kernel: CUDA-graph replays)."""
    a = _lib.SdeStepArgs()
    a.pos = _lib.f32(pos, "pos")
    a.B, a.N, a.R = topo.B, topo.N, topo.R
    assert pos.numel() == topo.B * topo.N * 3
    a.bond_uv = _lib.i32(topo.bond_uv, "bond_uv") if topo.R > 0 else None
    a.mask_rotate = _lib.u8(topo.mask_rotate, "mask_rotate") if topo.R > 0 else None
    a.tr_score, a.rot_score = _lib.f32(tr_score.contiguous(), "tr_score"), _lib.f32(rot_score.contiguous(), "rot_score")
    use_tor = tor_score is not None and topo.R > 0
    if use_tor:
        assert tor_score.numel() == topo.B * topo.R
    a.tor_score = _lib.f32(tor_score.contiguous(), "tor_score") if use_tor else None
    a.z_tr = _lib.f32(z_tr, "z_tr", allow_none=True)
    a.z_rot = _lib.f32(z_rot, "z_rot", allow_none=True)
    a.z_tor = _lib.f32(z_tor, "z_tor", allow_none=True) if use_tor else None
    if torch.is_tensor(coeffs):
        assert coeffs.numel() == 6
        a.coeffs_dev = _lib.f32(coeffs, "coeffs")
    else:
        (a.c_tr_score, a.c_tr_noise, a.c_rot_score, a.c_rot_noise, a.c_tor_score, a.c_tor_noise) = [float(c) for c in coeffs]
    _lib.sde_step(a)
    return pos


def modify_conformer_batch(orig_pos, data, tr_update, rot_update, torsion_updates, mask_rotate):
    """Reference signature (diffusion_utils.py:60-78); returns the new [B*N, 3] positions."""
    topo = LigandTopology(data, mask_rotate, orig_pos.device)
    pos = orig_pos.detach().float().clone().contiguous()
    sde_step(pos, topo, tr_update.float(), rot_update.float(),
             torsion_updates.float() if torsion_updates is not None else None, (1, 0, 1, 0, 1, 0))
    return pos.reshape(-1, 3)
