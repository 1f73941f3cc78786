"""IGSO(3) score-norm lookup (utils/so3.py:90-94) without the minutes-long import-time precomputation.

The 2000-entry `_exp_score_norms` table was produced by running the reference's own utils/so3.py
(oracle/gen_tables.py) and ships as tables/so3_exp_score_norms.npy.  Sampling/score-vector tables
(`sample_vec`, `score_vec`) are training-only and out of scope.
"""
import os

import numpy as np
import torch

MIN_EPS, MAX_EPS, N_EPS = 0.0005, 4, 2000

_exp_score_norms = np.load(os.path.join(os.path.dirname(__file__), "tables", "so3_exp_score_norms.npy"))
_dev_cache = {}


def eps_index(eps):
    """Same arithmetic (numpy, dtype of `eps` preserved) as the reference lookup."""
    eps_idx = (np.log10(eps) - np.log10(MIN_EPS)) / (np.log10(MAX_EPS) - np.log10(MIN_EPS)) * N_EPS
    return np.clip(np.around(eps_idx).astype(int), a_min=0, a_max=N_EPS - 1)


def score_norm(eps):
    """eps: CPU tensor -> float32 tensor of E||score|| (reference signature)."""
    return torch.from_numpy(_exp_score_norms[eps_index(eps.numpy())]).float()


def score_norm_device(rot_sigma, host_t, t_to_sigma, device):
    """[B] device tensor of score norms.  With the host-known diffusion time (recorded by set_time) the
    table index is computed on the host exactly like the CPU reference and no synchronisation happens;
    otherwise this falls back to the reference's own `.cpu()` round trip (score_model.py:420)."""
    if host_t is not None:
        cpu_t = [torch.full((1,), float(host_t[k]), dtype=torch.float32) for k in ("tr", "rot", "tor")]
        sig = t_to_sigma(*cpu_t)[1]
        val = float(score_norm(sig)[0])
        return torch.full((rot_sigma.shape[0],), val, dtype=torch.float32, device=device)
    return score_norm(rot_sigma.detach().cpu()).to(device)
