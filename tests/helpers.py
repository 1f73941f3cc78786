"""Shared test helpers: graph (de)serialisation, model construction, BN randomisation, noise injection."""
from __future__ import annotations

import copy
import os
import sys
from argparse import Namespace
from contextlib import contextmanager
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from confidence_bootstrapping_b200.data import Batch, HeteroData  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

_EDGE_TYPES = [("ligand", "lig_bond", "ligand"), ("receptor", "rec_contact", "receptor"),
               ("atom", "atom_contact", "atom"), ("atom", "atom_rec_contact", "receptor")]


def pack_graph(g) -> dict:
    d = {"nodes": {}, "edges": {}, "graph": {}}
    for nt in g.node_types:
        d["nodes"][nt] = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in g[nt].items()
                          if torch.is_tensor(v) or isinstance(v, np.ndarray)}
    for et in g.edge_types:
        d["edges"]["|".join(et)] = {k: v for k, v in g[et].items() if torch.is_tensor(v)}
    d["graph"]["original_center"] = g.original_center
    return d


def unpack_graph(d) -> HeteroData:
    g = HeteroData()
    for nt, attrs in d["nodes"].items():
        for k, v in attrs.items():
            if k in ("mask_rotate", "orig_pos"):
                v = v.numpy()
            setattr(g[nt], k, v)
    for et, attrs in d["edges"].items():
        for k, v in attrs.items():
            setattr(g[tuple(et.split("|"))], k, v)
    g.original_center = d["graph"]["original_center"]
    return g


def small_score_args(**kw) -> Namespace:
    """A tiny CG score configuration (fast on CPU, exercises every code path of the shipped one)."""
    from confidence_bootstrapping_b200.configs import score_model_args
    base = dict(ns=8, nv=2, num_conv_layers=2, num_prot_emb_layers=1, moad_esm_embeddings_path=None)
    base.update(kw)
    return score_model_args(**base)


def randomize_norm_stats(model, seed=0):
    """Make every BatchNorm non-trivial (fresh modules have mean 0 / var 1 / weight 1 / bias 0)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, "running_var") and torch.is_tensor(getattr(m, "running_var", None)):
                dev = m.running_var.device
                m.running_var.copy_((torch.rand(m.running_var.shape, generator=g) + 0.5).to(dev))
                if m.running_mean.numel():
                    m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.2).to(dev))
                if getattr(m, "weight", None) is not None:
                    m.weight.copy_((torch.rand(m.weight.shape, generator=g) + 0.5).to(dev))
                if getattr(m, "bias", None) is not None and m.bias.numel():
                    m.bias.copy_((torch.randn(m.bias.shape, generator=g) * 0.2).to(dev))


class NoiseTape:
    """Deterministic replacement for torch.normal: the k-th call returns the k-th recorded/seeded draw
    (always generated on the CPU, then moved), so CPU oracle, real reference and CUDA product see the
    same noise (sampling.py:126-141 draws with torch.normal on `device`)."""

    def __init__(self, seed=0):
        self.gen = torch.Generator().manual_seed(seed)
        self.calls = 0

    def __call__(self, mean=0, std=1, size=None, device=None, **kw):
        self.calls += 1
        z = torch.randn(tuple(size), generator=self.gen) * std + mean
        return z.to(device) if device is not None else z


@contextmanager
def injected_noise(seed=0):
    tape, orig = NoiseTape(seed), torch.normal
    torch.normal = tape
    try:
        yield tape
    finally:
        torch.normal = orig


def rmsd(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).pow(2).sum(-1).mean().sqrt().item()


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def blockwise_err(a, b, blocks=None):
    """Worst element-wise error with every block judged on its OWN scale (a global max-norm ratio lets
    small-magnitude blocks, e.g. the 1e / 0o channels, be off by far more than the bar and still pass):

        max over blocks, over elements of   |a - b| / max(|b|, rms(b_block))

    `blocks` = list of (start, stop) column ranges of the last dimension (one per irreps block); None = one block.
    Elements below the block's RMS are judged against the RMS (an fp32 sum of O(rms) terms cannot be more accurate
    than eps * rms in absolute terms, whatever its own magnitude)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    if a.dim() == 1:
        a, b = a.unsqueeze(1), b.unsqueeze(1)
    blocks = blocks or [(0, a.shape[-1])]
    worst = 0.0
    for lo, hi in blocks:
        bb, aa = b[..., lo:hi], a[..., lo:hi]
        rms = bb.pow(2).mean().sqrt().clamp(min=1e-30)
        worst = max(worst, ((aa - bb).abs() / torch.maximum(bb.abs(), rms)).max().item())
    return worst


def irreps_blocks(irreps: str):
    """(start, stop) column range of every irreps block of an e3nn-style string such as '32x0e + 6x1o'."""
    out, off = [], 0
    for part in irreps.replace(" ", "").split("+"):
        mul, ir = part.split("x")
        d = int(mul) * (2 * int(ir[:-1]) + 1)
        out.append((off, off + d))
        off += d
    return out


class SlicedNoiseTape(NoiseTape):
    """Noise for an oracle run over the FIRST `rows` graphs of a batch of `full_rows` graphs: every draw is generated
    at the full batch's size (so the stream matches the product run on the full batch) and its first rows returned."""

    def __init__(self, seed, rows, full_rows):
        super().__init__(seed)
        self.rows, self.full_rows = rows, full_rows

    def __call__(self, mean=0, std=1, size=None, device=None, **kw):
        size = tuple(size)
        assert size[0] % self.rows == 0
        per = size[0] // self.rows
        z = super().__call__(mean, std, (per * self.full_rows,) + size[1:], device)
        return z[: size[0]]


@contextmanager
def injected_sliced_noise(seed, rows, full_rows):
    tape, orig = SlicedNoiseTape(seed, rows, full_rows), torch.normal
    torch.normal = tape
    try:
        yield tape
    finally:
        torch.normal = orig


def load_1a0q(all_atoms=True, lm_dim=1280, lm_seed=0):
    """The reference's shipped complex data/1a0q (BASELINE.json config 1) from the committed fixture
    (oracle/make_1a0q_fixture.py).  Residue LM embeddings are seeded N(0,1) (the ESM2 cache is a preprocessing product)."""
    d = torch.load(os.path.join(GOLDEN, "1a0q.pt"), weights_only=False)
    g = unpack_graph(d)
    rec = g["receptor"]
    aa = rec.x.float()
    if lm_dim:
        lm = torch.from_numpy(np.random.default_rng(lm_seed).normal(size=(aa.shape[0], lm_dim)).astype(np.float32))
        aa = torch.cat([aa, lm], 1)
    rec.x = aa
    g["atom"].x = g["atom"].x.float()
    g["ligand"].x = g["ligand"].x.long()
    for et in g.edge_types:
        g[et].edge_index = g[et].edge_index.long()
    if not all_atoms:
        for key in [k for k in list(g._stores) if k == "atom" or (isinstance(k, tuple) and "atom" in (k[0], k[2]))]:
            del g._stores[key]
    g.name = "1a0q"
    return g
