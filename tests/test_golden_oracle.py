"""The oracle restatements against golden vectors produced by the REAL reference (oracle/make_golden.py)."""
import copy
import os
from argparse import Namespace
from functools import partial

import numpy as np
import pytest
import torch

from helpers import GOLDEN, injected_noise, rmsd, unpack_graph
from oracle import model as om, o3, sampler as osamp
from confidence_bootstrapping_b200 import so3 as pso3, torus as ptorus
from confidence_bootstrapping_b200.data import Batch


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_tables_match_reference_lookups():
    g = load("tables.pt")
    assert torch.equal(pso3.score_norm(g["so3_eps"]), g["so3_score_norm"])
    assert np.array_equal(ptorus.score_norm(g["torus_sigma"].numpy()), g["torus_score_norm"].numpy())


def test_faster_tp_restatement_and_fctp_anchor():
    """(a) oracle FasterTP == reference FasterTensorProduct; (b) e3nn-FCTP restatement == the same numbers
    after re-laying-out the weights -- the in-repo anchor for alpha and the l<=1 Wigner-3j signs."""
    for c in load("faster_tp.pt"):
        y = om.faster_tensor_product(c["in"], c["out"], c["x"], c["sh"], c["w"])
        assert om.faster_weight_numel(c["in"], c["out"]) == c["weight_numel"]
        assert torch.allclose(y, c["y"], atol=2e-6, rtol=1e-6)
        tp = o3.FullyConnectedTensorProduct(c["in"], "1x0e + 1x1o", c["out"])
        assert tp.weight_numel == c["weight_numel"]
        # FasterTP weights: per OUT irrep k a [fan_in, mul_out] block whose rows follow its own concat order;
        # e3nn: per instruction (i_in, i_sh, i_out) a [mul_in, 1, mul_out] block.  Map one onto the other.
        ins, outs = o3.Irreps(c["in"]), o3.Irreps(c["out"])
        order = {"0e": ["0e", "1o"], "1o": ["0e", "1o", "1e"], "1e": ["1o", "1e", "0o"], "0o": ["1e", "0o"]}
        nin = {str(ir): mul for mul, ir in ins}
        start, blocks = 0, {}
        for k in ("0e", "1o", "1e", "0o"):
            mo = {str(ir): mul for mul, ir in outs}.get(k, 0)
            row = 0
            for src in order[k]:
                if src in nin and mo:
                    blocks[(src, k)] = (start + row * mo, nin[src], mo)
                if src in nin:
                    row += nin[src]
            fan = sum(nin.get(s, 0) for s in order[k])
            start += fan * mo
        w_e3 = torch.zeros(c["w"].shape[0], tp.weight_numel)
        for n, (i1, i2, io, _, _) in enumerate(tp.instructions):
            src, dst = str(ins[i1].ir), str(outs[io].ir)
            off, m1, mo = blocks[(src, dst)]
            w_e3[:, tp._woff[n]: tp._woff[n] + m1 * mo] = c["w"][:, off: off + m1 * mo]
        assert torch.allclose(tp(c["x"], c["sh"], w_e3), c["y"], atol=2e-6, rtol=1e-6)


def test_geometry_restatements():
    g = load("geometry.pt")
    assert torch.allclose(osamp.axis_angle_to_matrix(g["axis_angle"]), g["matrix"], atol=1e-7)
    R, t = osamp.kabsch_batch(g["kabsch_A"], g["kabsch_B"])
    assert torch.allclose(R, g["kabsch_R"], atol=1e-5) and torch.allclose(t, g["kabsch_t"], atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(3), atol=1e-5)
    graph = unpack_graph(g["graph"])
    b3 = Batch.from_data_list([copy.deepcopy(graph) for _ in range(3)])
    mr = torch.from_numpy(graph["ligand"].mask_rotate)
    bonds = graph["ligand", "ligand"].edge_index.T[graph["ligand"].edge_mask]
    assert torch.allclose(osamp.twist_batch(g["pos0"].reshape(3, -1, 3), bonds, mr, g["tor"].reshape(3, -1)), g["twisted"], atol=1e-5)
    assert rmsd(osamp.modify_conformer_batch(g["pos0"], b3, g["tr"], g["rot"], g["tor"], mr), g["new_pos"]) < 1e-5
    assert rmsd(osamp.modify_conformer_batch(g["pos0"], b3, g["tr"], g["rot"], None, mr), g["rigid_only"]) < 1e-6


def _small_model_inputs():
    g = load("score_small.pt")
    args = Namespace(**g["args"])
    hp = om.hyper_from_args(args)
    t2s = partial(osamp.t_to_sigma, args=args)
    fwd = lambda b: om.cg_forward(g["state_dict"], hp, b, t2s, pso3.score_norm, ptorus.score_norm)
    return g, args, t2s, fwd


def test_score_model_forward_restatement():
    g, args, t2s, fwd = _small_model_inputs()
    batch = Batch.from_data_list([unpack_graph(x) for x in g["graphs"]])
    osamp.set_time(batch, g["t"], g["t"], g["t"], 2)
    with torch.no_grad():
        tr, rot, tor, _ = fwd(batch)
    for a, b in ((tr, g["tr"]), (rot, g["rot"]), (tor, g["tor"])):
        assert a.shape == b.shape
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)


def test_sampling_restatement():
    g, args, t2s, fwd = _small_model_inputs()
    base = Batch.from_data_list([unpack_graph(g["graphs"][1])])
    data_list = []
    for s in g["sample_start"]:
        d = copy.deepcopy(base)
        d["ligand"].pos = s.clone()
        data_list.append(d)
    sched = osamp.get_t_schedule(g["sample_steps"])
    with injected_noise(seed=g["sample_noise_seed"]):
        out, conf = osamp.sampling(data_list, fwd, g["sample_steps"], sched, sched, sched, t2s, args, batch_size=4)
    assert conf is None
    for d, want in zip(out, g["sample_final"]):
        assert rmsd(d["ligand"].pos, want) < 1e-3   # BASELINE.json: poses within 1e-3 A RMSD


def test_crop_beyond_restatement():
    g = load("crop.pt")
    graph = osamp.crop_beyond(unpack_graph(g["graph"]), g["cutoff"], True)
    want = unpack_graph(g["cropped"])
    for nt in ("receptor", "atom"):
        assert torch.equal(graph[nt].pos, want[nt].pos) and torch.equal(graph[nt].x, want[nt].x)
    for et in (("receptor", "receptor"), ("atom", "atom"), ("atom", "receptor")):
        assert torch.equal(graph[et].edge_index, want[et].edge_index)
    assert 0 < graph["receptor"].pos.shape[0] < unpack_graph(g["graph"])["receptor"].pos.shape[0]


def test_confidence_model_and_filtered_sampling_restatement():
    """All-atom confidence model forward, and sampling -> crop_beyond -> confidence scoring, against the real
    reference's outputs (confidences within 1e-4, BASELINE.json)."""
    g = load("confidence_small.pt")
    gs, sargs, t2s, score_fwd = _small_model_inputs()
    cargs = Namespace(**g["args"])
    hp = om.hyper_from_args(cargs, confidence_mode=True)
    conf_fwd = lambda b: om.aa_forward(g["state_dict"], hp, b, None, pso3.score_norm, ptorus.score_norm)
    batch = Batch.from_data_list([unpack_graph(x) for x in g["graphs"]])
    osamp.set_time(batch, 0, 0, 0, 2, all_atoms=True)
    with torch.no_grad():
        conf, atom_conf = conf_fwd(batch)
    assert torch.allclose(conf, g["confidence"], atol=1e-5) and torch.allclose(atom_conf, g["atom_confidence"], atol=1e-5)
    base = Batch.from_data_list([unpack_graph(g["graphs"][1])])
    data_list = []
    for s in g["sample_start"]:
        d = copy.deepcopy(base)
        d["ligand"].pos = s.clone()
        data_list.append(d)
    filt = copy.deepcopy(data_list)
    sched = g["sample_sched"].numpy()
    with injected_noise(seed=g["sample_noise_seed"]):
        out, sconf = osamp.sampling(data_list, score_fwd, 3, sched, sched, sched, t2s, sargs, batch_size=2,
                                    confidence_forward=conf_fwd, filtering_data_list=filt, filtering_model_args=cargs,
                                    crop_fn=osamp.crop_beyond)
    for d, want in zip(out, g["sample_final"]):
        assert rmsd(d["ligand"].pos, want) < 1e-3
    assert sconf.shape == g["sample_confidence"].shape
    assert torch.allclose(sconf, g["sample_confidence"], atol=1e-4)
