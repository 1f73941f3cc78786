"""Minimal PDB / SDF (V2000) text parser -> HeteroData with the reference's input schema.

Stands in for the RDKit / BioPython / ProDy featurisation of datasets/process_mols.py, which is preprocessing and out of
scope, for the ONE real complex the reference ships: data/1a0q
(BASELINE.json config 1).  What is reproduced from the reference:

  * receptor graph   process_mols.py:448-481 (`new_extract_receptor_structure`, reached from pdbbind.py:396 / moad.py:404):
                     one node per residue at its C-alpha, neighbours within `receptor_radius`, the
                     `c_alpha_max_neighbors` nearest if there are more, the nearest one if there are none;
                     edge_index = [neighbours, centre repeated];
  * atom graph       process_mols.py:492-527: heavy atoms, same radius/kNN rule, atom -> residue edges;
  * ligand graph     process_mols.py:567-589 (`get_lig_graph`): heavy atoms in file order, every bond as the pair
                     (u,v),(v,u), one-hot bond type SINGLE/DOUBLE/TRIPLE/AROMATIC;
  * rotatable bonds  utils/torsion.py:15-45 (`get_transformation_mask`) through synthetic.rotatable_bond_masks;
  * centring         datasets/pdbbind.py:411-422: everything minus the C-alpha centroid, kept as `original_center`.

Categorical ATOM FEATURES that need a chemistry toolkit (chirality, hybridisation, implicit valence, aromaticity flags)
are filled with what the file itself states (element, degree, attached hydrogens, formal charge 0, SDF aromatic bond
type, ring membership from the bond graph); they only index embedding tables, so any in-vocabulary value exercises the
same arithmetic.  Language-model embeddings (1280-d ESM2, a cached preprocessing product) are seeded N(0,1).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .data import HeteroData
from .synthetic import _knn_radius_edges, rotatable_bond_masks

# vocabularies of process_mols.py:56-93 (only their order matters: features are indices into embedding tables)
AMINO_ACIDS = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER",
               "THR", "TRP", "TYR", "VAL", "HIP", "HIE", "TPO", "HID", "LEV", "MEU", "PTR", "GLV", "CYT", "SEP", "HIZ", "CYM",
               "GLM", "ASQ", "TYS", "CYX", "GLZ", "misc"]
ATOM_TYPE_2 = ["C*", "CA", "CB", "CD", "CE", "CG", "CH", "CZ", "N*", "ND", "NE", "NH", "NZ", "O*", "OD", "OE", "OG", "OH", "OX",
               "S*", "SD", "SG", "misc"]
ATOM_TYPE_3 = ["C", "CA", "CB", "CD", "CD1", "CD2", "CE", "CE1", "CE2", "CE3", "CG", "CG1", "CG2", "CH2", "CZ", "CZ2", "CZ3", "N",
               "ND1", "ND2", "NE", "NE1", "NE2", "NH1", "NH2", "NZ", "O", "OD1", "OD2", "OE1", "OE2", "OG", "OG1", "OH", "OXT",
               "SD", "SG", "misc"]
ATOMIC_NUM = {"H": 1, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "P": 15, "S": 16, "CL": 17, "BR": 35, "I": 53, "SE": 34, "FE": 26,
              "ZN": 30, "MG": 12, "CA": 20, "NA": 11, "K": 19, "MN": 25, "CU": 29, "CO": 27, "NI": 28}


def _index(vocab, key):
    """process_mols.py `safe_index`: the last entry ('misc') for anything not listed."""
    try:
        return vocab.index(key)
    except ValueError:
        return len(vocab) - 1


def parse_pdb(path: str):
    """ATOM records -> per residue (name, [(atom name, element, xyz)]) in file order; hydrogens dropped."""
    residues: List[Tuple[str, list]] = []
    last = None
    with open(path) as fh:
        for line in fh:
            if not line.startswith("ATOM"):
                continue
            name, alt, resname = line[12:16].strip(), line[16], line[17:20].strip()
            key = (line[21], line[22:27])
            if alt not in (" ", "A"):
                continue
            elem = line[76:78].strip().upper() or name[:1]
            if elem == "H" or name.startswith("H"):
                continue
            xyz = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
            if key != last:
                residues.append((resname, []))
                last = key
            residues[-1][1].append((name, elem, xyz))
    return [r for r in residues if any(a[0] == "CA" for a in r[1])]      # a residue needs its C-alpha (process_mols.py:492-500)


def parse_sdf(path: str):
    """First molecule of a V2000 SDF -> (elements, xyz [N,3], bonds [(i, j, type)]) with 0-based indices."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    counts = lines[3]
    n_atoms, n_bonds = int(counts[0:3]), int(counts[3:6])
    elems, xyz = [], []
    for line in lines[4:4 + n_atoms]:
        xyz.append((float(line[0:10]), float(line[10:20]), float(line[20:30])))
        elems.append(line[31:34].strip().upper())
    bonds = []
    for line in lines[4 + n_atoms:4 + n_atoms + n_bonds]:
        bonds.append((int(line[0:3]) - 1, int(line[3:6]) - 1, int(line[6:9])))
    return elems, np.asarray(xyz, dtype=np.float64), bonds


def _ring_sizes(n, adj):
    """Per atom: set of sizes (3..8) of simple rings through it (bounded depth-first search on the bond graph)."""
    sizes = [set() for _ in range(n)]
    for s in range(n):
        stack = [(s, [s])]
        while stack:
            a, path = stack.pop()
            for b in adj[a]:
                if b == s and len(path) >= 3:
                    for v in path:
                        sizes[v].add(len(path))
                elif b not in path and b > s and len(path) < 8:
                    stack.append((b, path + [b]))
    return sizes


def ligand_graph(g: HeteroData, elems, xyz, bonds, remove_hs=True):
    """get_lig_graph (process_mols.py:567-589) on the parsed molecule; hydrogens removed like `remove_hs: true`."""
    n = len(elems)
    heavy = [i for i in range(n) if not (remove_hs and elems[i] == "H")]
    new = {old: k for k, old in enumerate(heavy)}
    n_h = [0] * n
    deg = [0] * n
    for i, j, _ in bonds:
        deg[i] += 1
        deg[j] += 1
        if elems[j] == "H":
            n_h[i] += 1
        if elems[i] == "H":
            n_h[j] += 1
    hb = [(new[i], new[j], t) for i, j, t in bonds if i in new and j in new]
    adj = [set() for _ in heavy]
    aromatic = [False] * len(heavy)
    for i, j, t in hb:
        adj[i].add(j)
        adj[j].add(i)
        if t == 4:
            aromatic[i] = aromatic[j] = True
    rings = _ring_sizes(len(heavy), adj)
    feats = []
    for k, old in enumerate(heavy):
        z = ATOMIC_NUM.get(elems[old], 119)
        d = min(deg[old], 11)
        hyb = 1 if aromatic[k] or any(t == 2 for i, j, t in hb if k in (i, j)) else (0 if any(t == 3 for i, j, t in hb if k in (i, j)) else 2)
        feats.append([min(z - 1, 118), 0, d, 5, 0, min(n_h[old], 9), 0, hyb, int(aromatic[k]), min(len(rings[k]), 7),
                      int(3 in rings[k]), int(4 in rings[k]), int(5 in rings[k]), int(6 in rings[k]), int(7 in rings[k]), int(8 in rings[k])])
    row, col, typ = [], [], []
    for i, j, t in hb:
        row += [i, j]
        col += [j, i]
        typ += [t - 1 if 1 <= t <= 4 else 0] * 2
    g["ligand"].x = torch.tensor(feats, dtype=torch.int64)
    g["ligand"].pos = torch.from_numpy(xyz[heavy].astype(np.float32))
    g["ligand", "lig_bond", "ligand"].edge_index = torch.tensor([row, col], dtype=torch.int64)
    g["ligand", "lig_bond", "ligand"].edge_attr = torch.nn.functional.one_hot(torch.tensor(typ), num_classes=4).float()
    mask_edges, mask_rotate = rotatable_bond_masks(len(heavy), np.asarray([row, col]).T)
    g["ligand"].edge_mask = torch.from_numpy(mask_edges)
    g["ligand"].mask_rotate = mask_rotate
    return g


def receptor_graph(g: HeteroData, residues, all_atoms=True, receptor_radius=15.0, c_alpha_max_neighbors=24, atom_radius=5.0,
                   atom_max_neighbors=8, lm_dim=1280, lm_seed=0):
    ca = np.asarray([next(a[2] for a in atoms if a[0] == "CA") for _, atoms in residues], dtype=np.float64)
    aa = np.asarray([[_index(AMINO_ACIDS, name)] for name, _ in residues], dtype=np.float32)
    feats = [aa]
    if lm_dim:
        feats.append(np.random.default_rng(lm_seed).normal(size=(len(residues), lm_dim)).astype(np.float32))
    g["receptor"].x = torch.from_numpy(np.concatenate(feats, 1))
    g["receptor"].pos = torch.from_numpy(ca.astype(np.float32))
    # side_chain_vecs = [chi angles / 360 (5) | N - CA | C - CA] (process_mols.py:450-452); only crop_beyond carries it along
    # (utils/utils.py:414), no model reads it; the chi angles need the side-chain topology tables and are left at zero
    rel = lambda nm: np.asarray([next((a[2] for a in atoms if a[0] == nm), c) for (_, atoms), c in zip(residues, ca)]) - ca
    g["receptor"].side_chain_vecs = torch.from_numpy(np.concatenate([np.zeros((len(ca), 5)), rel("N"), rel("C")], 1).astype(np.float32))
    # rows [neighbour, centre]: what the dataset path builds (pdbbind.py:396 -> process_mols.py:415,448-481)
    g["receptor", "rec_contact", "receptor"].edge_index = torch.from_numpy(_knn_radius_edges(ca, receptor_radius, c_alpha_max_neighbors))
    if all_atoms:
        apos, ax, aidx = [], [], []
        for r, (resname, atoms) in enumerate(residues):
            for name, elem, xyz in atoms:
                apos.append(xyz)
                ax.append([_index(AMINO_ACIDS, resname), min(ATOMIC_NUM.get(elem, 119) - 1, 118), _index(ATOM_TYPE_2, (name + "*")[:2]),
                           _index(ATOM_TYPE_3, name)])
                aidx.append(r)
        apos = np.asarray(apos, dtype=np.float64)
        g["atom"].x = torch.tensor(ax, dtype=torch.float32)
        g["atom"].pos = torch.from_numpy(apos.astype(np.float32))
        g["atom", "atom_contact", "atom"].edge_index = torch.from_numpy(_knn_radius_edges_grid(apos, atom_radius, atom_max_neighbors))
        g["atom", "atom_rec_contact", "receptor"].edge_index = torch.from_numpy(np.stack([np.arange(len(aidx)), np.asarray(aidx)]).astype(np.int64))
    return g


def _knn_radius_edges_grid(pos, cutoff, max_neighbors):
    """Same edge set as synthetic._knn_radius_edges (neighbours within `cutoff`, the nearest `max_neighbors` if more, the
    nearest one if none; rows = [neighbour, centre]) without the dense N x N matrix: KD-tree queries."""
    from scipy.spatial import cKDTree
    tree = cKDTree(pos)
    src, dst = [], []
    for i, p in enumerate(pos):
        nb = [j for j in tree.query_ball_point(p, cutoff) if j != i and np.linalg.norm(pos[j] - p) < cutoff]
        if len(nb) > max_neighbors or not nb:
            k = (max_neighbors if nb else 1) + 1
            d, idx = tree.query(p, k=k)
            nb = [int(j) for j in idx if j != i][: k - 1]
        else:
            nb.sort()
        src += [i] * len(nb)
        dst += nb
    return np.asarray([dst, src], dtype=np.int64)


def load_complex(pdb_path: str, sdf_path: str, all_atoms=True, name=None, **kw) -> HeteroData:
    """PDB + SDF -> one centred complex graph with the keys `sampling()` / the models read."""
    g = HeteroData()
    g.name = name or pdb_path
    receptor_graph(g, parse_pdb(pdb_path), all_atoms=all_atoms, **kw)
    ligand_graph(g, *parse_sdf(sdf_path))
    center = g["receptor"].pos.mean(dim=0, keepdim=True)                       # datasets/pdbbind.py:411-422
    g["ligand"].orig_pos = g["ligand"].pos.numpy().astype(np.float64)
    g["receptor"].pos = g["receptor"].pos - center
    g["ligand"].pos = g["ligand"].pos - center
    if all_atoms:
        g["atom"].pos = g["atom"].pos - center
    g.original_center = center
    return g
