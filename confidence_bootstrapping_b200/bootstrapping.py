"""Confidence Bootstrapping bookkeeping around the sampler: which sampled poses are kept, and the replay buffer they go into.

SURVEY.md section 8f rank 3 -- the host-side stage between `sampling()` and the fine-tuning epoch:
  * `select_confident`  finetune_train.py:223-232 : keep (pose, confidence) pairs whose confidence exceeds the cutoff
                        (first column of a multi-class confidence head, :223-224); the selection runs on the device
                        tensor `sampling()` returned and crosses the bus once;
  * `plain_rmsd`        finetune_train.py:216 : the non-symmetry-corrected RMSD the reference falls back to (the corrected
                        one needs spyrmsd's graph isomorphism, out of scope);
  * `CBBuffer`          bootstrapping/buffer.py:9-116 : the buffer of self-generated complexes -- `add_complexes`
                        (stamps confidence / iteration / t = 0 time tensors, per-couple top-k under
                        confidence + buffer_decay * iteration), `get` (round-robin, or softmax(confidence * temperature)
                        sampling when `fixed_length` is set), `len`.
Same names, argument meaning and policies as the reference; the PyG `Dataset` base class and the hard-coded
`data/BindingMOAD_2020_processed/new_cluster_to_ligands.pkl` lookup are replaced by an explicit `ligand_names`
list (the pickle is still read when no list is given and the file exists).
"""
from __future__ import annotations

import copy
import os
import pickle
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

CLUSTER_FILE = "data/BindingMOAD_2020_processed/new_cluster_to_ligands.pkl"


def select_confident(predictions_list: Sequence, confidences, confidence_cutoff: float, multi_class: bool = False,
                     n_samples: Optional[int] = None) -> List[Tuple[object, object]]:
    """[(graph, confidence)] for the samples with confidence > cutoff (finetune_train.py:223-232)."""
    if confidences is None:
        return []
    conf = confidences
    if torch.is_tensor(conf):
        if multi_class and conf.dim() > 1:
            conf = conf[:, 0]
        keep = torch.nonzero(conf > confidence_cutoff).flatten().cpu().tolist()      # one device->host read
    else:
        conf = np.asarray(conf)
        keep = [int(i) for i in np.nonzero(conf > confidence_cutoff)[0]]
    n = len(predictions_list) if n_samples is None else n_samples
    return [(predictions_list[i], conf[i]) for i in keep if i < n]


def plain_rmsd(ligand_pos, orig_ligand_pos):
    """sqrt(mean_atoms sum_xyz (pos - ref)^2) per sample (finetune_train.py:216); ligand_pos [S, N, 3], ref [N, 3]."""
    p = ligand_pos if torch.is_tensor(ligand_pos) else torch.as_tensor(np.asarray(ligand_pos))
    r = orig_ligand_pos if torch.is_tensor(orig_ligand_pos) else torch.as_tensor(np.asarray(orig_ligand_pos))
    return ((p - r.to(p.device, p.dtype)) ** 2).sum(dim=2).mean(dim=1).sqrt()


def _name_of(graph) -> str:
    n = graph.name
    return n[0] if isinstance(n, (list, tuple)) else n


class CBBuffer:
    def __init__(self, cluster_name=None, root=None, transform=None, multiplicity=1, max_complexes_per_couple=None,
                 fixed_length=None, temperature=1.0, buffer_decay=0.2, reset_buffer=False, ligand_names=None):
        self.root, self.transform = root, transform
        self.multiplicity = multiplicity
        self.complexes = []
        self.iteration = 0
        self.max_complexes_per_couple = max_complexes_per_couple
        self.fixed_length = fixed_length
        self.temperature = temperature
        self.buffer_decay = buffer_decay
        self.reset_buffer = reset_buffer
        if ligand_names is None:
            assert cluster_name is not None
            with open(CLUSTER_FILE, "rb") as f:
                ligand_names = pickle.load(f)[cluster_name]
        self.ligand_names = list(ligand_names)
        self.ligand_cnt = {name: 0 for name in self.ligand_names}

    # -- torch Dataset surface ------------------------------------------------
    def get(self, idx):
        if self.fixed_length is None:
            complex_graph = copy.deepcopy(self.complexes[idx % len(self.complexes)])
        else:
            confidences = np.asarray([float(c.confidence) for c in self.complexes])
            weights = np.exp(confidences * self.temperature)
            weights = weights / np.sum(weights)
            idx = np.random.choice(len(self.complexes), p=weights)
            complex_graph = copy.deepcopy(self.complexes[idx])
        for attr in ("confidence", "iteration"):
            if hasattr(complex_graph, attr):
                delattr(complex_graph, attr)
            for nt in ("receptor", "ligand"):
                if attr in complex_graph[nt]:
                    delattr(complex_graph[nt], attr)
        return complex_graph if self.transform is None else self.transform(complex_graph)

    def len(self):
        return len(self.complexes) * self.multiplicity if self.fixed_length is None else self.fixed_length

    __len__ = len

    def __getitem__(self, idx):
        return self.get(idx)

    def statistics(self):
        return {"complexes": len(self.complexes), "ligand_cnt": dict(self.ligand_cnt)}

    def print_statistics(self):
        print(f"Buffer with {len(self.complexes)} complexes.")
        for ligand, cnt in self.ligand_cnt.items():
            print(f"Ligand: {ligand} Cnt: {cnt}")

    def add_complexes(self, new_complex_list):
        for complex_graph, confidence in new_complex_list:
            complex_graph.confidence = confidence
            complex_graph.iteration = self.iteration
            t = 0
            complex_graph.complex_t = {k: t * torch.ones(1) for k in ("tr", "rot", "tor")}
            for nt in ("ligand", "receptor"):
                n = complex_graph[nt].num_nodes
                complex_graph[nt].node_t = {k: t * torch.ones(n) for k in ("tr", "rot", "tor")}
            self.ligand_cnt[_name_of(complex_graph)] += 1
            complex_graph.cpu()
        self.iteration += 1
        if self.reset_buffer:
            self.complexes = [c for c, _ in new_complex_list]
        else:
            self.complexes.extend([c for c, _ in new_complex_list])
        if self.max_complexes_per_couple is not None:
            c_to_samples = {}
            for s in self.complexes:
                c_to_samples[_name_of(s)[:6]] = []
            for s in self.complexes:      # "the policy is quite arbitrary here" (buffer.py:99)
                c_to_samples[_name_of(s)[:6]].append((float(s.confidence) + self.buffer_decay * s.iteration, s))
            for c in c_to_samples:
                if len(c_to_samples[c]) > self.max_complexes_per_couple:
                    c_to_samples[c] = sorted(c_to_samples[c], key=lambda x: x[0], reverse=True)[:self.max_complexes_per_couple]
            self.complexes = [s for c in c_to_samples for _, s in c_to_samples[c]]
