"""Wrapped-normal (torus) score-norm lookup (utils/torus.py:78-82).

The reference builds `score_norm_` by Monte Carlo with UNSEEDED numpy at import time
(utils/torus.py:71-75), so its values differ from process to process.  This package ships one table,
produced by running the reference's own module under np.random.seed(0) (oracle/gen_tables.py), and
the oracle uses the same one.  `score`, `p`, `sample` are training-only and out of scope.
"""
import os

import numpy as np
import torch

SIGMA_MIN, SIGMA_MAX, SIGMA_N = 3e-3, 2, 5000  # relative to pi

score_norm_ = np.load(os.path.join(os.path.dirname(__file__), "tables", "torus_score_norm_seed0.npy"))


def score_norm(sigma):
    """sigma: numpy array (dtype preserved, like the reference) -> E[score^2] from the table."""
    sigma = np.log(sigma / np.pi)
    sigma = (sigma - np.log(SIGMA_MIN)) / (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)) * SIGMA_N
    sigma = np.round(np.clip(sigma, 0, SIGMA_N)).astype(int)
    return score_norm_[sigma]


def score_norm_device(edge_sigma, host_t, t_to_sigma, tor_batch, device):
    """[n_rotatable] float32 device tensor (score_model.py:444-448); host index when the time is host-known."""
    if host_t is not None:
        cpu_t = [torch.full((1,), float(host_t[k]), dtype=torch.float32) for k in ("tr", "rot", "tor")]
        sig = t_to_sigma(*cpu_t)[2]
        val = float(torch.tensor(score_norm(sig.numpy())).float()[0])
        return torch.full((edge_sigma.shape[0],), val, dtype=torch.float32, device=device)
    return torch.tensor(score_norm(edge_sigma.detach().cpu().numpy())).float().to(device)
