"""Seeded synthetic protein-ligand complexes with the reference's input schema.

Schema follows datasets/process_mols.py (`get_lig_graph` :567-589, `new_extract_receptor_structure`
:448-527) and utils/torsion.py:15-45 (`get_transformation_mask`): the synthetic generator only
replaces the RDKit/BioPython featurisation, which is out of scope.
There are no datasets or checkpoints offline, so bench/tests use these (BASELINE.json configs 2-5).
"""
from __future__ import annotations

import numpy as np
import torch

from .data import HeteroData

# datasets/process_mols.py:95-123 (lengths of the categorical feature vocabularies)
LIG_FEATURE_DIMS = ([119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2], 0)
REC_ATOM_FEATURE_DIMS = ([38, 119, 23, 38], 0)
REC_RESIDUE_FEATURE_DIMS = ([38], 0)


def _knn_radius_edges(pos: np.ndarray, cutoff: float, max_neighbors: int) -> np.ndarray:
    """Row 0 = neighbour, row 1 = centre, grouped by centre (process_mols.py:458-477)."""
    d = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=-1)
    src, dst = [], []
    for i in range(len(pos)):
        nb = [j for j in np.where(d[i] < cutoff)[0] if j != i]
        if len(nb) > max_neighbors:
            nb = list(np.argsort(d[i], kind="stable")[1: max_neighbors + 1])
        if not nb:
            nb = list(np.argsort(d[i], kind="stable")[1:2])
        src += [i] * len(nb)
        dst += nb
    return np.asarray([dst, src], dtype=np.int64)


def _knn_radius_edges_fast(pos: np.ndarray, cutoff: float, max_neighbors: int) -> np.ndarray:
    """Same edge list as `_knn_radius_edges` (verified element for element in tests/test_host_logic.py) from one batched
    KD-tree query instead of the dense N x N distance matrix + Python loop: 1000-residue / 8000-atom receptors in
    milliseconds instead of ten seconds."""
    from scipy.spatial import cKDTree
    n = len(pos)
    k = min(n, max_neighbors + 2)                       # self + (max_neighbors + 1) others
    dist, idx = cKDTree(pos).query(pos, k=k)
    dist, idx = dist.reshape(n, k), idx.reshape(n, k)
    src, dst = [], []
    for i in range(n):
        others = idx[i][idx[i] != i][: k - 1]
        # exact distances as the dense path computes them (the tree's own may differ in the last bit at the cutoff)
        d = np.linalg.norm(pos[i][None, :] - pos[others], axis=-1) if len(others) else np.zeros(0)
        inside = others[d < cutoff]
        if len(inside) > max_neighbors:
            nb = [int(j) for j in others[np.argsort(d, kind="stable")][:max_neighbors]]
        elif len(inside) == 0:
            nb = [int(others[int(np.argmin(d))])] if len(others) else []
        else:
            nb = sorted(int(j) for j in inside)
        src += [i] * len(nb)
        dst += nb
    return np.asarray([dst, src], dtype=np.int64)


def rotatable_bond_masks(n_atoms: int, edges: np.ndarray):
    """`get_transformation_mask` (utils/torsion.py:15-45) without networkx/PyG.

    edges: int [E, 2] directed bond list whose entries come in (u,v),(v,u) pairs.
    Returns (mask_edges bool[E], mask_rotate bool[R, n_atoms]).
    """
    adj = [set() for _ in range(n_atoms)]
    for u, v in edges:
        adj[u].add(v)

    def components(skip):
        seen, comps = [False] * n_atoms, []
        for s in range(n_atoms):
            if seen[s]:
                continue
            stack, comp = [s], []
            seen[s] = True
            while stack:
                a = stack.pop()
                comp.append(a)
                for b in adj[a] | {c for c in range(n_atoms) if a in adj[c]}:
                    if {a, b} == skip or seen[b]:
                        continue
                    seen[b] = True
                    stack.append(b)
            comps.append(comp)
        return comps

    to_rotate = []
    for i in range(0, len(edges), 2):
        assert edges[i, 0] == edges[i + 1, 1]
        comps = components({int(edges[i, 0]), int(edges[i, 1])})
        if len(comps) > 1:
            small = sorted(comps, key=len)[0]
            if len(small) > 1:
                if edges[i, 0] in small:
                    to_rotate += [[], small]
                else:
                    to_rotate += [small, []]
                continue
        to_rotate += [[], []]
    mask_edges = np.asarray([len(l) > 0 for l in to_rotate], dtype=bool)
    mask_rotate = np.zeros((int(mask_edges.sum()), n_atoms), dtype=bool)
    k = 0
    for i, l in enumerate(to_rotate):
        if mask_edges[i]:
            mask_rotate[k][np.asarray(l, dtype=int)] = True
            k += 1
    return mask_edges, mask_rotate


def _ligand(rng, n_atoms: int, ring_prob=0.15):
    """Branched self-avoiding chain with 1.5 A bonds (+ a few ring closures)."""
    pos = np.zeros((n_atoms, 3))
    bonds = []
    for a in range(1, n_atoms):
        for _ in range(200):
            parent = int(rng.integers(max(0, a - 4), a))
            v = rng.normal(size=3)
            cand = pos[parent] + 1.5 * v / np.linalg.norm(v)
            if a == 1 or np.min(np.linalg.norm(pos[:a] - cand, axis=1)) > 1.2:
                break
        pos[a] = cand
        bonds.append((parent, a))
        if a >= 5 and rng.random() < ring_prob:
            d = np.linalg.norm(pos[:a - 1] - cand, axis=1)
            j = int(np.argmin(d))
            if d[j] < 2.6 and (j, a) not in bonds and j != parent:
                bonds.append((j, a))
    row, col, typ = [], [], []
    for (u, v) in bonds:
        t = int(rng.integers(0, 4))
        row += [u, v]
        col += [v, u]
        typ += [t, t]
    edge_index = np.asarray([row, col], dtype=np.int64)
    return pos, edge_index, np.asarray(typ, dtype=np.int64)


def _receptor(rng, n_res: int):
    """C-alpha cloud: ball of protein-like density with >= 3.8 A separation."""
    radius = (3.0 * n_res * 135.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    pts = np.zeros((0, 3))
    while len(pts) < n_res:
        cand = rng.uniform(-radius, radius, size=(4 * n_res, 3))
        cand = cand[np.linalg.norm(cand, axis=1) < radius]
        for c in cand:
            if len(pts) == 0 or np.min(np.linalg.norm(pts - c, axis=1)) >= 3.8:
                pts = np.vstack([pts, c])
                if len(pts) == n_res:
                    break
    return pts


def make_complex(seed: int, n_res: int, n_lig: int, all_atoms: bool = True, lm_dim: int = 1280,
                 atoms_per_res: int = 8, receptor_radius: float = 15.0, c_alpha_max_neighbors: int = 24,
                 atom_radius: float = 5.0, atom_max_neighbors: int = 8, name: str | None = None) -> HeteroData:
    rng = np.random.default_rng(seed)
    g = HeteroData()
    g.name = name or f"syn{seed}_{n_res}x{n_lig}"

    # receptor -----------------------------------------------------------------
    ca = _receptor(rng, n_res)
    aa = rng.integers(0, REC_RESIDUE_FEATURE_DIMS[0][0], size=(n_res, 1)).astype(np.float32)
    feats = [aa]
    if lm_dim:
        feats.append(rng.normal(size=(n_res, lm_dim)).astype(np.float32))
    # pocket = surface-ish point; everything is centred on the receptor centroid like the reference
    center = ca.mean(0, keepdims=True)
    ca = ca - center
    g["receptor"].x = torch.from_numpy(np.concatenate(feats, 1))
    g["receptor"].pos = torch.from_numpy(ca.astype(np.float32))
    g["receptor"].side_chain_vecs = torch.from_numpy(rng.normal(size=(n_res, 11)).astype(np.float32))
    g["receptor", "rec_contact", "receptor"].edge_index = torch.from_numpy(
        _knn_radius_edges_fast(ca, receptor_radius, c_alpha_max_neighbors))
    g.original_center = torch.from_numpy(center.astype(np.float32))

    if all_atoms:
        n_per = rng.integers(max(1, atoms_per_res - 3), atoms_per_res + 4, size=n_res)
        apos, aidx = [], []
        for r in range(n_res):
            off = rng.normal(size=(n_per[r], 3))
            off = off / np.linalg.norm(off, axis=1, keepdims=True) * rng.uniform(0.0, 3.0, size=(n_per[r], 1))
            off[0] = 0.0
            apos.append(ca[r] + off)
            aidx += [r] * int(n_per[r])
        apos = np.concatenate(apos, 0)
        n_atom = len(apos)
        ax = np.stack([rng.integers(0, d, size=n_atom) for d in REC_ATOM_FEATURE_DIMS[0]], 1)
        g["atom"].x = torch.from_numpy(ax.astype(np.float32))
        g["atom"].pos = torch.from_numpy(apos.astype(np.float32))
        g["atom", "atom_contact", "atom"].edge_index = torch.from_numpy(
            _knn_radius_edges_fast(apos, atom_radius, atom_max_neighbors))
        g["atom", "atom_rec_contact", "receptor"].edge_index = torch.from_numpy(
            np.stack([np.arange(n_atom), np.asarray(aidx)]).astype(np.int64))

    # ligand -------------------------------------------------------------------
    lpos, bond_index, bond_type = _ligand(rng, n_lig)
    # place near a random residue (a "pocket"), like a docked pose
    anchor = ca[int(rng.integers(0, n_res))]
    lpos = lpos - lpos.mean(0, keepdims=True) + anchor + rng.normal(size=3) * 2.0
    lx = np.stack([rng.integers(0, d, size=n_lig) for d in LIG_FEATURE_DIMS[0]], 1)
    g["ligand"].x = torch.from_numpy(lx.astype(np.int64))
    g["ligand"].pos = torch.from_numpy(lpos.astype(np.float32))
    g["ligand"].orig_pos = lpos + center
    g["ligand", "lig_bond", "ligand"].edge_index = torch.from_numpy(bond_index)
    g["ligand", "lig_bond", "ligand"].edge_attr = torch.nn.functional.one_hot(
        torch.from_numpy(bond_type), num_classes=4).float()
    mask_edges, mask_rotate = rotatable_bond_masks(n_lig, bond_index.T)
    g["ligand"].edge_mask = torch.from_numpy(mask_edges)
    g["ligand"].mask_rotate = mask_rotate
    return g
