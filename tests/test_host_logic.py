"""CPU-side checks of the product's host logic: TP programs (irreps.py), C-ABI surface, containers."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT
from oracle import model as om, o3
from confidence_bootstrapping_b200 import _lib, irreps
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.synthetic import make_complex, rotatable_bond_masks


def eval_program(P, x, sh, w):
    """Reference evaluator of a TPProgram on the CPU (float64): out = sum_rows f_r * w[w_base+m]."""
    x, sh, w = x.double().numpy(), sh.double().numpy(), w.double().numpy()
    out = np.zeros((x.shape[0], P.d_out))
    for r in P.rows:
        f = np.zeros(x.shape[0])
        for t in P.terms[r["term_begin"]: r["term_end"]]:
            f += t["coef"].astype(np.float64) * x[:, t["x_idx"]] * sh[:, t["sh_idx"]]
        for m in range(r["mul"]):
            out[:, r["out_base"] + m * r["out_step"]] += f * w[:, r["w_base"] + m]
    return torch.from_numpy(out)


SEQ = ["32x0e", "32x0e + 6x1o", "32x0e + 6x1o + 6x1e", "32x0e + 6x1o + 6x1e + 6x0o"]


@pytest.mark.parametrize("i_in,i_out,numel,rows", [(0, 1, 1216, 128), (1, 2, 1480, 170), (2, 3, 1588, 212), (3, 3, 1660, 236)])
def test_faster_program_equals_oracle(i_in, i_out, numel, rows):
    P = irreps.faster_tp_program(SEQ[i_in], SEQ[i_out])
    assert P.weight_numel == numel == om.faster_weight_numel(SEQ[i_in], SEQ[i_out])
    assert P.n_rows == rows
    torch.manual_seed(0)
    x, sh, w = torch.randn(4, P.d_in), torch.randn(4, 4), torch.randn(4, numel)
    want = om.faster_tensor_product(SEQ[i_in], SEQ[i_out], x.double(), sh.double(), w.double())
    assert torch.allclose(eval_program(P, x, sh, w), want, atol=1e-6)


CONF = ["24x0e", "24x0e + 6x1o", "24x0e + 6x1o + 6x1e", "24x0e + 6x1o + 6x1e + 24x0o"]


@pytest.mark.parametrize("in_ir,sh_ir,out_ir,numel", [
    (CONF[0], "1x0e + 1x1o + 1x2e", CONF[1], 720), (CONF[1], "1x0e + 1x1o + 1x2e", CONF[2], 972),
    (CONF[2], "1x0e + 1x1o + 1x2e", CONF[3], 1224), (CONF[3], "1x0e + 1x1o + 1x2e", CONF[3], 1944),
    (SEQ[3], "1x0e + 1x1o", "2x1o + 2x1e", 124), (SEQ[3], "1x1o", "32x0o + 32x0e", 384)])
def test_fctp_program_equals_oracle(in_ir, sh_ir, out_ir, numel):
    P = irreps.fctp_program(in_ir, sh_ir, out_ir)
    tp = o3.FullyConnectedTensorProduct(in_ir, sh_ir, out_ir)
    assert P.weight_numel == numel == tp.weight_numel
    torch.manual_seed(1)
    x, sh, w = torch.randn(3, P.d_in), torch.randn(3, P.sh_dim), torch.randn(3, numel)
    assert torch.allclose(eval_program(P, x, sh, w), tp(x.double(), sh.double(), w.double()), atol=1e-6)
    assert P.n_rows <= 320 and P.out_ptr[-1] == P.n_slots == sum(r["mul"] for r in P.rows)


def test_full_tp_1o_block_matches_oracle():
    dim, off, w = irreps.full_tp_1o_block(1)
    ftp = o3.FullTensorProduct("1x0e + 1x1o", "2e")
    assert dim == ftp.irreps_out.dim == 20 and off == 0
    torch.manual_seed(2)
    sh, y2 = torch.randn(5, 4).double(), torch.randn(5, 5).double()
    want = ftp(sh, y2)[:, :3]
    got = torch.einsum("ijk,ei,ej->ek", torch.from_numpy(w), sh[:, 1:4], y2)
    assert torch.allclose(got, want, atol=1e-12)


def test_wigner_matches_oracle():
    for ls in [(1, 1, 0), (1, 1, 1), (1, 1, 2), (1, 2, 1), (2, 2, 2), (0, 2, 2)]:
        assert np.allclose(irreps.wigner_3j(*ls), o3.wigner_3j(*ls).numpy(), atol=1e-12)


def test_cabi_exports_and_struct_layout():
    """The shared library loads, exports every symbol include/cb200.h declares, and the ctypes mirrors
    have the C sizes (no compute call: this runs without a GPU)."""
    header = open(os.path.join(ROOT, "include", "cb200.h")).read()
    declared = set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", header)) - {"cb_tp_conv_args", "cb_edge_feat_args"}
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in cb200.h but not exported"
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert lib.cb_version() >= 100
    for which, struct in enumerate([_lib.EdgeFeatArgs, _lib.TpSegment, _lib.TpConvArgs, _lib.SdeStepArgs]):
        assert lib.cb_sizeof(which) == ctypes.sizeof(struct)
    assert lib.cb_sizeof(4) == irreps.ROW_DTYPE.itemsize and lib.cb_sizeof(5) == irreps.TERM_DTYPE.itemsize
    assert lib.cb_sizeof(6) == irreps.RUN_DTYPE.itemsize


def test_wrappers_refuse_cpu_tensors():
    """No CPU fallback: handing host tensors to a kernel wrapper is an error, not a slow path."""
    t = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.f32(t, "x")


def test_batch_collate_and_split_roundtrip():
    gs = [make_complex(s, 20 + 5 * s, 6 + s, all_atoms=True, lm_dim=4) for s in range(3)]
    b = Batch.from_data_list(gs)
    assert b.num_graphs == 3
    assert b["ligand"].batch.tolist() == sum([[i] * g["ligand"].num_nodes for i, g in enumerate(gs)], [])
    off = 0
    for i, g in enumerate(gs):
        n = g["receptor"].num_nodes
        e = b["receptor", "receptor"].edge_index
        sel = (b["receptor"].batch[e[0]] == i)
        assert torch.equal(e[:, sel] - off, g["receptor", "receptor"].edge_index)
        off += n
    for g, h in zip(gs, b.to_data_list()):
        assert torch.equal(g["atom", "receptor"].edge_index, h["atom", "receptor"].edge_index)
        assert torch.equal(g["ligand"].pos, h["ligand"].pos)
    # nested collate of 1-graph batches (how the reference's callers build data_list, sampling.py:78)
    nested = Batch.from_data_list([Batch.from_data_list([gs[0]]), Batch.from_data_list([gs[0]])])
    assert nested.num_graphs == 2 and nested["ligand"].num_nodes == 2 * gs[0]["ligand"].num_nodes


def test_rotatable_masks_orientation():
    g = make_complex(4, 30, 18, all_atoms=False, lm_dim=0)
    ei = g["ligand", "ligand"].edge_index.T.numpy()
    me, mr = rotatable_bond_masks(g["ligand"].num_nodes, ei)
    assert mr.shape == (me.sum(), g["ligand"].num_nodes) and me.sum() > 0
    for k, (u, v) in enumerate(ei[me]):
        assert not mr[k, u] and mr[k, v]          # utils/torsion.py:81-82
        assert 1 < mr[k].sum() <= g["ligand"].num_nodes // 2 + 1


def test_collate_to_device_matches_host_collate():
    """Batch.from_data_list(device=...) (replicated attributes transferred once and tiled) == plain collate."""
    import copy
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.synthetic import make_complex
    g = make_complex(3, 40, 12, all_atoms=True, lm_dim=64)
    dl = [copy.deepcopy(g) for _ in range(3)]
    for i, d in enumerate(dl):
        d["ligand"].pos = d["ligand"].pos + float(i)          # the one attribute that differs between the copies
    a, b = Batch.from_data_list(dl), Batch.from_data_list(dl, device="cpu")
    for key, st in a._stores.items():
        for k, v in st.items():
            w = b[key]._d[k]
            if torch.is_tensor(v):
                assert v.dtype == w.dtype and torch.equal(v, w), (key, k)
    assert a.num_graphs == b.num_graphs == 3


def _randomize_position_per_sample(data_list, tr_sigma_max):
    """Literal restatement of utils/sampling.py:15-48 (per-sample loops), the checker for the batched version."""
    from scipy.spatial.transform import Rotation as R
    from confidence_bootstrapping_b200 import sampling as S
    centre_pocket = data_list[0]["receptor"].pos.mean(dim=0)
    for g in data_list:
        upd = np.random.uniform(low=-np.pi, high=np.pi, size=int(g["ligand"].edge_mask.sum()))
        bonds = g["ligand", "ligand"].edge_index.T[g["ligand"].edge_mask]
        g["ligand"].pos = S._twist_numpy(g["ligand"].pos, bonds, S._mask_rotate_of(g), upd)
    for g in data_list:
        c = torch.mean(g["ligand"].pos, dim=0, keepdim=True)
        rot = torch.from_numpy(R.random().as_matrix()).float()
        g["ligand"].pos = (g["ligand"].pos - c) @ rot.T + centre_pocket
        g["ligand"].pos += torch.normal(mean=0, std=tr_sigma_max, size=(1, 3))


def test_randomize_position_batched_matches_per_sample_loop():
    """randomize_position applies the geometry per topology group at once; same random draws, same poses as the
    reference's per-sample loop (also for a list that mixes two ligands)."""
    import copy
    from confidence_bootstrapping_b200.sampling import randomize_position
    from confidence_bootstrapping_b200.synthetic import make_complex
    a, b = make_complex(11, 30, 14, all_atoms=False, lm_dim=0), make_complex(12, 30, 9, all_atoms=False, lm_dim=0)
    base = [copy.deepcopy(a) for _ in range(3)] + [copy.deepcopy(b)] + [copy.deepcopy(a) for _ in range(2)]
    got, want = copy.deepcopy(base), copy.deepcopy(base)
    np.random.seed(4); torch.manual_seed(4)
    randomize_position(got, False, False, 10.0)
    np.random.seed(4); torch.manual_seed(4)
    _randomize_position_per_sample(want, 10.0)
    for x, y in zip(got, want):
        assert x["ligand"].pos.dtype == torch.float32
        assert float((x["ligand"].pos - y["ligand"].pos).abs().max()) < 1e-4


def test_dead_output_gates_are_k_hop_closures():
    """score_model._dead_output_gates: layer n-2 keeps the receptors with a cross edge, layer n-2-k additionally everything
    those read through k rec->rec hops (checked against a brute-force reachability on a random graph)."""
    from types import SimpleNamespace
    from confidence_bootstrapping_b200.graph import EdgeList
    from confidence_bootstrapping_b200.score_model import TensorProductScoreModel
    g = torch.Generator().manual_seed(0)
    NR, E = 60, 150
    row = torch.randint(0, NR, (E,), generator=g)      # aggregation node reads col
    col = torch.randint(0, NR, (E,), generator=g)
    order = torch.argsort(row, stable=True)
    row, col = row[order], col[order]
    rowptr = torch.zeros(NR + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=NR), 0).to(torch.int32)
    rec_edges = EdgeList(rowptr, row.to(torch.int32), col.to(torch.int32), E, NR)
    has_cross = torch.zeros(NR, dtype=torch.bool)
    has_cross[torch.randperm(NR, generator=g)[:6]] = True
    rl_ptr = torch.zeros(NR + 1, dtype=torch.int32)
    rl_ptr[1:] = torch.cumsum(has_cross.int() * 3, 0).to(torch.int32)
    rl = EdgeList(rl_ptr, torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), int(rl_ptr[-1]), NR)
    st = SimpleNamespace(NR=NR, rec_edges=rec_edges, rec_row_long=None, rec_col_long=None)
    gates = TensorProductScoreModel._dead_output_gates(st, rl, n_layers=6, hops=4)
    assert sorted(gates) == [0, 1, 2, 3, 4]
    keep = has_cross.clone()
    for k, layer in enumerate([4, 3, 2, 1, 0]):
        assert torch.equal(gates[layer].bool(), keep), layer
        nxt = keep.clone()
        for r, c in zip(row.tolist(), col.tolist()):
            if keep[r]:
                nxt[c] = True
        keep = nxt
    assert TensorProductScoreModel._dead_output_gates(st, rl, n_layers=1) == {}


def test_collate_flags_replicated_node_types():
    import copy
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.synthetic import make_complex
    a, b = make_complex(3, 40, 12, all_atoms=True, lm_dim=8), make_complex(4, 40, 12, all_atoms=True, lm_dim=8)
    same = [copy.deepcopy(a) for _ in range(3)]
    for i, d in enumerate(same):
        d["ligand"].pos = d["ligand"].pos + float(i)
    flagged = Batch.from_data_list(same, device="cpu")._g["_replicated_types"]
    assert "receptor" in flagged and "atom" in flagged and "ligand" not in flagged
    assert Batch.from_data_list([copy.deepcopy(a), copy.deepcopy(b)], device="cpu")._g["_replicated_types"] == []
    assert Batch.from_data_list(same)._g["_replicated_types"] == []      # the host collate does not compare


def test_to_data_list_after_in_place_edit_rebuilds_slices():
    """crop_beyond edits a batch in place; the collate's slices must not be trusted afterwards."""
    import copy
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.synthetic import make_complex
    gs = [make_complex(s, 20 + 5 * s, 6 + s, all_atoms=True, lm_dim=4) for s in range(3)]
    b = Batch.from_data_list(copy.deepcopy(gs))
    # drop the first two residues of graph 1 by hand (what crop_beyond does, on the host)
    rec = b["receptor"]
    keep = torch.ones(rec.pos.shape[0], dtype=torch.bool)
    o1 = gs[0]["receptor"].num_nodes
    keep[o1:o1 + 2] = False
    remap = torch.cumsum(keep.long(), 0) - 1
    rr, ar, atom, aa = b["receptor", "receptor"], b["atom", "receptor"], b["atom"], b["atom", "atom"]
    ek = keep[rr.edge_index[0]] & keep[rr.edge_index[1]]
    rr.edge_index = remap[rr.edge_index[:, ek]]
    akeep = keep[ar.edge_index[1]]
    amap = torch.cumsum(akeep.long(), 0) - 1
    ar.edge_index = torch.stack([torch.arange(int(akeep.sum())), remap[ar.edge_index[1]][akeep]])
    ek = akeep[aa.edge_index[0]] & akeep[aa.edge_index[1]]
    aa.edge_index = amap[aa.edge_index[:, ek]]
    atom.x, atom.pos, atom.batch = atom.x[akeep], atom.pos[akeep], atom.batch[akeep]
    rec.x, rec.pos, rec.batch = rec.x[keep], rec.pos[keep], rec.batch[keep]
    b._g["_slices_stale"] = True
    out = b.to_data_list()
    assert [d["receptor"].num_nodes for d in out] == [gs[0]["receptor"].num_nodes, gs[1]["receptor"].num_nodes - 2, gs[2]["receptor"].num_nodes]
    assert torch.equal(out[0]["receptor", "receptor"].edge_index, gs[0]["receptor", "receptor"].edge_index)
    assert torch.equal(out[2]["receptor", "receptor"].edge_index, gs[2]["receptor", "receptor"].edge_index)
    assert torch.equal(out[2]["atom", "receptor"].edge_index, gs[2]["atom", "receptor"].edge_index)
    assert torch.equal(out[2]["atom"].pos, gs[2]["atom"].pos)
    assert int(out[1]["receptor", "receptor"].edge_index.max()) < out[1]["receptor"].num_nodes
    assert torch.equal(out[1]["ligand"].pos, gs[1]["ligand"].pos)


def test_derived_weight_caches_are_invalidated():
    """EMA-style `param.data.copy_()` does not bump tensor versions; the folded tensors must still follow."""
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    torch.manual_seed(0)
    layer = TensorProductConvLayer("8x0e + 2x1o", "1x0e + 1x1o", "8x0e + 2x1o", 24, hidden_features=24, faster=True).eval()
    w2a = layer._w2a(0).clone()
    scale, shift = [t.clone() for t in layer.batch_norm.affine()]
    W2 = layer.fc[3].weight
    W2.data.copy_(W2.data * 2)                       # what ExponentialMovingAverage.copy_to does
    layer.batch_norm.weight.data.copy_(layer.batch_norm.weight.data * 3)
    assert torch.equal(layer._w2a(0), w2a)           # the version-keyed cache cannot see it ...
    layer.invalidate_caches()                        # ... which is why sampling() refreshes per call
    assert torch.allclose(layer._w2a(0)[:, :24], 2 * w2a[:, :24])
    assert torch.allclose(layer.batch_norm.affine()[0], 3 * scale)
    # .train()/.eval(), load_state_dict and .to() refresh as well
    W2.data.copy_(W2.data * 0.5)
    layer.train(False)
    assert torch.allclose(layer._w2a(0)[:, :24], w2a[:, :24])
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    sd["fc.3.weight"] = sd["fc.3.weight"] * 4
    layer._w2a(0)
    layer.load_state_dict(sd)
    assert torch.allclose(layer._w2a(0)[:, :24], 4 * w2a[:, :24])


def test_sampling_scopes_eval_mode_and_restores_train_mode():
    """finetune_train.py:177 samples with the score model in train mode; sampling() must not raise."""
    import warnings
    from confidence_bootstrapping_b200 import sampling as smp
    net = torch.nn.Sequential(torch.nn.Linear(2, 2), torch.nn.Dropout(0.5)).train()
    other = torch.nn.Linear(2, 2).eval()
    smp._warned_train_mode = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with smp._inference_mode(net, other, None):
            assert not net.training and not net[1].training and not other.training
    assert net.training and net[1].training and not other.training
    assert any("train mode" in str(x.message) for x in w)


def test_1a0q_fixture_known_answers():
    """BASELINE config 1 input: the parsed data/1a0q complex has the sizes SURVEY.md section 8a derived from the shipped files
    (416 residues, 3183 receptor heavy atoms, 23 ligand heavy atoms, 46 directed bond edges, 11 rotatable bonds, 9962
    rec-rec edges at r = 15 A / 24 neighbours, 242 ligand radius-5 edges with max degree 15)."""
    from helpers import load_1a0q
    from oracle import cluster
    g = load_1a0q()
    assert g["receptor"].num_nodes == 416 and g["receptor"].x.shape == (416, 1281)
    assert g["atom"].num_nodes == 3183 and g["atom"].x.shape[1] == 4
    assert g["ligand"].num_nodes == 23 and g["ligand"].x.shape == (23, 16)
    ll = g["ligand", "ligand"]
    assert ll.edge_index.shape == (2, 46) and ll.edge_attr.shape == (46, 4)
    assert int(g["ligand"].edge_mask.sum()) == 11 and g["ligand"].mask_rotate.shape == (11, 23)
    assert g["receptor", "receptor"].edge_index.shape == (2, 9962)
    assert torch.bincount(g["receptor", "receptor"].edge_index[1]).max() <= 24       # <= 24 neighbours per centre
    assert g["atom", "receptor"].edge_index.shape == (2, 3183)
    # feature indices stay inside the embedding vocabularies (process_mols.py:95-123)
    from confidence_bootstrapping_b200.synthetic import LIG_FEATURE_DIMS, REC_ATOM_FEATURE_DIMS
    assert all(int(g["ligand"].x[:, i].max()) < d for i, d in enumerate(LIG_FEATURE_DIMS[0]))
    assert all(int(g["atom"].x[:, i].max()) < d for i, d in enumerate(REC_ATOM_FEATURE_DIMS[0]))
    assert float(g["receptor"].pos.mean(0).abs().max()) < 1e-3                        # centred on the C-alpha centroid
    b = torch.zeros(23, dtype=torch.long)
    rad = cluster.radius_graph(g["ligand"].pos, 5.0, b)
    assert rad.shape[1] == 242 and int(torch.bincount(rad[1]).max()) == 15
    # torsion.py:81-82 orientation of the rotation masks
    ei = ll.edge_index.T[g["ligand"].edge_mask]
    for k, (u, v) in enumerate(ei.tolist()):
        assert not g["ligand"].mask_rotate[k, u] and g["ligand"].mask_rotate[k, v]


def test_fast_knn_radius_edges_equal_the_dense_loop():
    from confidence_bootstrapping_b200.synthetic import _knn_radius_edges, _knn_radius_edges_fast
    for seed, n, cutoff, mx in [(0, 300, 15.0, 24), (1, 120, 15.0, 24), (2, 900, 5.0, 8), (3, 5, 15.0, 24), (4, 60, 3.0, 8), (5, 2, 1.0, 8)]:
        pos = np.random.default_rng(seed).normal(size=(n, 3)) * (n ** (1 / 3)) * 3
        assert np.array_equal(_knn_radius_edges(pos, cutoff, mx), _knn_radius_edges_fast(pos, cutoff, mx)), (seed, n)


def test_cb_buffer_matches_the_reference_class():
    """bootstrapping/buffer.py:CBBuffer: the cb200 class driven through the same seeded add / get sequence
    as the REAL reference class (golden recorded under oracle/shims.py by oracle/make_buffer_golden.py) holds the same
    complexes with the same confidence / iteration stamps after every round, reports the same length and hands out the same
    samples -- for the plain buffer, the per-couple top-k with decay, softmax sampling at fixed length, and reset mode."""
    import os
    from confidence_bootstrapping_b200.bootstrapping import CBBuffer
    from confidence_bootstrapping_b200.data import HeteroData
    from helpers import GOLDEN
    g = torch.load(os.path.join(GOLDEN, "cb_buffer.pt"), weights_only=False)

    def fake(name, tag):
        c = HeteroData()
        c.name = [name]
        c.tag = tag
        c["ligand"].pos = torch.full((5, 3), float(tag))
        c["receptor"].pos = torch.zeros(7, 3)
        return c

    for case in g["cases"]:
        buf = CBBuffer(cluster_name="c0", ligand_names=g["names"], **case["kwargs"])
        for new, want in zip(case["rounds"], case["trace"]):
            buf.add_complexes([(fake(n, t), torch.tensor(c)) for n, t, c in new])
            held = [(int(c.tag), float(c.confidence), int(c.iteration)) for c in buf.complexes]
            assert held == want["held"], case["kwargs"]
            assert buf.len() == want["len"] and buf.ligand_cnt == want["cnt"]
            np.random.seed(case["get_seed"])
            got = [buf.get(i) for i in range(min(6, buf.len()))]
            assert [int(c.tag) for c in got] == want["got"]
            assert all(not hasattr(c, "confidence") and not hasattr(c, "iteration") for c in got)
        assert float(buf.complexes[0].complex_t["tr"]) == 0.0 and buf.complexes[0]["ligand"].node_t["rot"].shape == (5,)


def test_select_confident_and_plain_rmsd():
    """finetune_train.py:216,223-232: poses above the confidence cutoff (first column of a multi-class head), plain RMSD."""
    from confidence_bootstrapping_b200.bootstrapping import plain_rmsd, select_confident
    preds = [f"pose{i}" for i in range(5)]
    conf = torch.tensor([0.3, -1.0, 2.5, 0.0, 0.31])
    kept = select_confident(preds, conf, 0.3)
    assert [p for p, _ in kept] == ["pose2", "pose4"] and float(kept[0][1]) == 2.5
    multi = torch.stack([conf, conf.flip(0)], 1)
    assert [p for p, _ in select_confident(preds, multi, 0.3, multi_class=True)] == ["pose2", "pose4"]
    assert select_confident(preds, None, 0.0) == []
    pos = torch.zeros(2, 4, 3)
    pos[1, :, 0] = 2.0
    assert torch.allclose(plain_rmsd(pos, torch.zeros(4, 3)), torch.tensor([0.0, 2.0]))


def test_step_graph_cache_key_ignores_the_pose_only():
    """reverse_diffusion re-uses a captured step graph for a batch that equals the captured one up to the ligand pose:
    shape-level signature from the collate (cheap candidate key) + element-wise check against the captured tensors."""
    import copy
    from confidence_bootstrapping_b200 import sampling as smp
    g = make_complex(11, 30, 9, all_atoms=False)

    def batch_of(graph, n, shift):
        dl = [copy.deepcopy(graph) for _ in range(n)]
        for i, d in enumerate(dl):
            d["ligand"].pos = d["ligand"].pos * shift + float(i)
        return Batch.from_data_list(dl, device="cpu"), smp._mask_rotate_of(dl[0])

    b0, mr0 = batch_of(g, 3, 1.0)
    sig = b0._g["_static_sig"]
    assert sig is not None and not any(k == "pos" and "ligand" in key for key, k, _, _ in sig)
    assert Batch.from_data_list([copy.deepcopy(g)])._g.get("_static_sig") is None          # host collate: no key
    entry = dict(static=smp._static_tensors(b0), mask_rotate=copy.deepcopy(mr0))
    assert ("ligand", "pos") not in entry["static"] and ("receptor", "pos") in entry["static"]
    b1, mr1 = batch_of(g, 3, 2.0)                               # other poses, same complex
    assert b1._g["_static_sig"] == sig and smp._same_static(b1, mr1, entry)
    other = make_complex(12, 30, 9, all_atoms=False)            # same sizes, other complex: same shapes, different content
    b2, mr2 = batch_of(other, 3, 1.0)
    if b2._g["_static_sig"] == sig:
        assert not smp._same_static(b2, mr2, entry)
    b3, _ = batch_of(g, 2, 1.0)                                 # other number of copies
    assert b3._g["_static_sig"] != sig
    g4 = copy.deepcopy(g)
    g4["receptor"].pos[0, 0] += 1.0                             # one moved residue
    b4, mr4 = batch_of(g4, 3, 1.0)
    assert b4._g["_static_sig"] == sig and not smp._same_static(b4, mr4, entry)


def test_sampling_refreshes_derived_weights_only_when_the_weights_moved():
    """_inference_mode compares a per-tensor norm signature: unchanged weights keep the folded tensors (and the cached step
    graphs), an EMA-style `.data.copy_()` drops them."""
    from confidence_bootstrapping_b200 import sampling as smp, tensor_layers as tl
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    torch.manual_seed(0)
    layer = TensorProductConvLayer("8x0e + 2x1o", "1x0e + 1x1o", "8x0e + 2x1o", 24, hidden_features=24, faster=True).eval()
    with smp._inference_mode(layer):
        w2a = layer._w2a(0)
    e0 = tl.CACHE_EPOCH
    with smp._inference_mode(layer):
        assert layer._w2a(0) is w2a and tl.CACHE_EPOCH == e0          # nothing moved: same tensor object, same epoch
    layer.fc[3].weight.data.copy_(layer.fc[3].weight.data * 2)
    with smp._inference_mode(layer):
        assert tl.CACHE_EPOCH > e0
        assert torch.allclose(layer._w2a(0)[:, :24], 2 * w2a[:, :24])
    e1 = tl.CACHE_EPOCH
    layer.batch_norm.running_var.data.mul_(4.0)                          # buffers count too
    with smp._inference_mode(layer):
        assert tl.CACHE_EPOCH > e1


def test_filtering_leg_runs_one_batch_behind_and_keeps_the_order():
    """sampling.FilteringLeg: the leg of batch i is executed when batch i+1 is submitted (or at finish), results in batch order."""
    from confidence_bootstrapping_b200.sampling import FilteringLeg
    leg = FilteringLeg(torch.device("cpu"))
    assert not leg.enabled                      # one stream on the CPU: the legs run inline
    calls = []
    for i in range(3):
        leg.submit(lambda p, i=i: (calls.append(i), p * (i + 1))[1], torch.full((2,), 1.0))
        assert calls == list(range(i))          # batch i's leg has not run yet
    out = leg.finish()
    assert calls == [0, 1, 2] and [float(o[0]) for o in out] == [1.0, 2.0, 3.0]
    assert leg.finish() == []


def test_host_count_tables_replace_the_device_reads():
    """The per-batch set-up of the models takes node / rotatable-bond counts from the collate's host tables (graph.host_counts),
    selects masked columns with the host-known count and counts per bin without torch.bincount: same values, no device read."""
    import copy
    from confidence_bootstrapping_b200.diffusion_utils import check_rotation_masks
    from confidence_bootstrapping_b200.graph import count_per_bin, host_counts, masked_columns
    from confidence_bootstrapping_b200.sampling import _mask_rotate_of
    graphs = [make_complex(70 + i, 20 + 7 * i, 8 + 2 * i, all_atoms=True) for i in range(3)]
    b = Batch.from_data_list(copy.deepcopy(graphs), device="cpu")
    hc = host_counts(b, ("ligand", "receptor", "atom"))
    for nt in ("ligand", "receptor", "atom"):
        assert hc[nt] == torch.bincount(b[nt].batch, minlength=3).tolist()
    mask = b["ligand"].edge_mask.bool()
    lig_edge_graph = b["ligand"].batch[b["ligand", "ligand"].edge_index[0]]
    assert hc["n_tor"] == torch.bincount(lig_edge_graph[mask], minlength=3).tolist()
    ei = b["ligand", "ligand"].edge_index
    assert torch.equal(masked_columns(ei, mask, sum(hc["n_tor"])), ei[:, mask]) and torch.equal(masked_columns(ei, mask), ei[:, mask])
    idx = torch.tensor([4, 0, 4, 2])
    assert torch.equal(count_per_bin(idx, 6), torch.bincount(idx, minlength=6))
    assert count_per_bin(idx[:0], 3).tolist() == [0, 0, 0]
    # tables are refused once the batch was edited in place; a host collate carries them too
    b._g["_slices_stale"] = True
    assert host_counts(b, ("ligand",)) is None
    assert host_counts(Batch.from_data_list(copy.deepcopy(graphs)), ("ligand",)) == {"ligand": hc["ligand"], "n_tor": hc["n_tor"]}
    # the host-side orientation check of the rotation masks (torsion.py:81-82)
    g0 = graphs[0]
    mr = _mask_rotate_of(g0)
    assert check_rotation_masks(g0, mr)
    bad = np.asarray(mr).copy()
    if bad.shape[0] > 0:
        bad[0] = ~bad[0]
        with pytest.raises(AssertionError):
            check_rotation_masks(g0, bad)
