"""End-to-end parity of the cb200 score model and sampler against the CPU oracle / reference goldens."""
import copy
import os
from argparse import Namespace
from functools import partial

import numpy as np
import pytest
import torch

from helpers import GOLDEN, injected_noise, randomize_norm_stats, rel_err, rmsd, small_score_args, unpack_graph
from oracle import model as om, sampler as osamp

pytestmark = pytest.mark.gpu


def _build(args, seed=0):
    from confidence_bootstrapping_b200 import so3, torus
    from confidence_bootstrapping_b200.diffusion_utils import t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.utils import get_model
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(seed)
    model = get_model(args, torch.device("cuda"), t_to_sigma=t2s, no_parallel=True)
    randomize_norm_stats(model, seed=seed + 1)
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    hp = om.hyper_from_args(args)
    oracle_fwd = lambda b: om.cg_forward(sd, hp, b, partial(osamp.t_to_sigma, args=args), so3.score_norm, torus.score_norm)
    return model, t2s, oracle_fwd


def test_golden_small_model_forward_and_sampling():
    """Reference-produced golden (real models/score_model.py + utils/sampling.py on CPU): forward outputs
    to 1e-5 relative, final poses to 1e-3 A (BASELINE.json)."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, set_time, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import sampling
    from confidence_bootstrapping_b200.utils import get_model
    g = torch.load(os.path.join(GOLDEN, "score_small.pt"), weights_only=False)
    args = Namespace(**g["args"])
    t2s = partial(t2s_full, args=args)
    dev = torch.device("cuda")
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True)
    model.load_state_dict(g["state_dict"], strict=True)
    model.eval()
    batch = Batch.from_data_list([unpack_graph(x) for x in g["graphs"]]).to(dev)
    set_time(batch, None, g["t"], g["t"], g["t"], 2, False, False, dev)
    with torch.no_grad():
        tr, rot, tor, side = model(batch)
    assert side is None
    for got, want in ((tr, g["tr"]), (rot, g["rot"]), (tor, g["tor"])):
        assert got.shape == want.shape and rel_err(got, want) < 1e-5
    base = Batch.from_data_list([unpack_graph(g["graphs"][1])])
    data_list = []
    for s in g["sample_start"]:
        d = copy.deepcopy(base)
        d["ligand"].pos = s.clone()
        data_list.append(d)
    sched = get_t_schedule("expbeta", g["sample_steps"], 1, 1)
    with injected_noise(seed=g["sample_noise_seed"]):
        out, conf = sampling(data_list=data_list, model=model, inference_steps=g["sample_steps"], tr_schedule=sched,
                             rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s, model_args=args, batch_size=4)
    assert conf is None
    for d, want in zip(out, g["sample_final"]):
        assert d["ligand"].pos.is_cuda
        assert rmsd(d["ligand"].pos, want) < 1e-3


@pytest.mark.parametrize("sizes", [[(60, 14)], [(90, 23), (48, 9), (120, 31)]])
def test_full_size_score_model_vs_oracle(sizes):
    """Shipped hyper-parameters (ns 32, nv 6, 3+5 layers, lmax 1, 1280-d LM features), seeded random weights.
    Single layers are held to 1e-5 relative (test_gpu_kernels.py::test_tp_conv_layer_vs_oracle); the composition
    of 11 TP-conv layers + heads, whose outputs are O(1e-6) differences of O(1) terms at random init, to 1e-4."""
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, oracle_fwd = _build(args)
    graphs = [make_complex(40 + i, nr, nl, all_atoms=False) for i, (nr, nl) in enumerate(sizes)]
    for t in (1.0, 0.35):
        cpu = Batch.from_data_list(copy.deepcopy(graphs))
        osamp.set_time(cpu, t, t, t, len(graphs))
        gpu = Batch.from_data_list(copy.deepcopy(graphs)).to("cuda")
        set_time(gpu, None, t, t, t, len(graphs), False, False, torch.device("cuda"))
        with torch.no_grad():
            want = oracle_fwd(cpu)
            got = model(gpu)
        for a, b, name in zip(got[:3], want[:3], ("tr", "rot", "tor")):
            assert a.shape == b.shape, name
            assert rel_err(a, b) < 1e-4, (name, t, rel_err(a, b))


def test_sampling_vs_oracle_full_size():
    """20 reverse-diffusion steps with identical injected noise: poses within 1e-3 A RMSD of the CPU oracle."""
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, oracle_fwd = _build(args, seed=3)
    g = Batch.from_data_list([make_complex(77, 80, 16, all_atoms=False)])
    np.random.seed(0)
    torch.manual_seed(0)
    dl = [copy.deepcopy(g) for _ in range(4)]
    randomize_position(dl, False, False, args.tr_sigma_max)
    dl_cpu = copy.deepcopy(dl)
    steps = 20
    sched = get_t_schedule("expbeta", steps, 1, 1)
    with injected_noise(seed=9):
        out, _ = sampling(data_list=dl, model=model, inference_steps=steps, tr_schedule=sched, rot_schedule=sched,
                          tor_schedule=sched, device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=4)
    with injected_noise(seed=9):
        ref, _ = osamp.sampling(dl_cpu, oracle_fwd, steps, sched, sched, sched, partial(osamp.t_to_sigma, args=args), args, batch_size=4)
    for a, b in zip(out, ref):
        assert rmsd(a["ligand"].pos, b["ligand"].pos) < 1e-3
    # ragged last batch raises like the reference (sampling.py:126-131)
    with pytest.raises(Exception):
        sampling(data_list=[copy.deepcopy(g) for _ in range(3)], model=model, inference_steps=2, tr_schedule=sched,
                 rot_schedule=sched, tor_schedule=sched, device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=2)


def test_model_refuses_training_mode():
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = small_score_args()
    model, _, _ = _build(args)
    model.train()
    b = Batch.from_data_list([make_complex(1, 30, 8, all_atoms=False, lm_dim=0)]).to("cuda")
    set_time(b, None, 0.5, 0.5, 0.5, 1, False, False, torch.device("cuda"))
    with pytest.raises(NotImplementedError):
        model(b)


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_crop_beyond_bit_exact():
    """Device crop (K1 mask + prefix-sum compaction) vs the reference's host crop: identical tensors/indices,
    single graph (golden from the real utils/utils.py:395-420) and a batch of heterogeneous graphs (oracle)."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import crop_beyond
    g = _load("crop.pt")
    got = crop_beyond(Batch.from_data_list([unpack_graph(g["graph"])]).to("cuda"), g["cutoff"], True)
    want = unpack_graph(g["cropped"])
    for nt in ("receptor", "atom"):
        assert torch.equal(got[nt].pos.cpu(), want[nt].pos) and torch.equal(got[nt].x.cpu(), want[nt].x)
    for et in (("receptor", "receptor"), ("atom", "atom"), ("atom", "receptor")):
        assert torch.equal(got[et].edge_index.cpu(), want[et].edge_index)
    graphs = [make_complex(60 + i, 40 + 9 * i, 8 + 2 * i, all_atoms=True, lm_dim=3) for i in range(3)]
    cropped = [osamp.crop_beyond(copy.deepcopy(x), 10.0, True) for x in graphs]
    want_b = Batch.from_data_list(cropped)
    got_b = crop_beyond(Batch.from_data_list(copy.deepcopy(graphs)).to("cuda"), 10.0, True)
    for nt in ("receptor", "atom"):
        assert torch.equal(got_b[nt].pos.cpu(), want_b[nt].pos) and torch.equal(got_b[nt].batch.cpu(), want_b[nt].batch)
    for et in (("receptor", "receptor"), ("atom", "atom"), ("atom", "receptor")):
        assert torch.equal(got_b[et].edge_index.cpu(), want_b[et].edge_index)


def test_golden_confidence_model_and_filtered_sampling():
    """Reference-produced golden: all-atom confidence model (lmax 2, 9 edge groups) forward, and
    sampling -> crop_beyond -> confidence scoring.  Confidences within 1e-4, poses within 1e-3 A."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import sampling
    from confidence_bootstrapping_b200.utils import get_model
    g, gs = _load("confidence_small.pt"), _load("score_small.pt")
    dev = torch.device("cuda")
    cargs, sargs = Namespace(**g["args"]), Namespace(**gs["args"])
    t2s = partial(t2s_full, args=sargs)
    cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True)
    cmodel.load_state_dict(g["state_dict"], strict=True)
    cmodel.eval()
    smodel = get_model(sargs, dev, t_to_sigma=t2s, no_parallel=True)
    smodel.load_state_dict(gs["state_dict"], strict=True)
    smodel.eval()
    batch = Batch.from_data_list([unpack_graph(x) for x in g["graphs"]]).to(dev)
    set_time(batch, 0, 0, 0, 0, 2, True, False, dev)
    with torch.no_grad():
        conf, atom_conf = cmodel(batch)
    assert torch.allclose(conf.cpu(), g["confidence"], atol=1e-4) and torch.allclose(atom_conf.cpu(), g["atom_confidence"], atol=1e-4)
    base = Batch.from_data_list([unpack_graph(g["graphs"][1])])
    data_list = []
    for s in g["sample_start"]:
        d = copy.deepcopy(base)
        d["ligand"].pos = s.clone()
        data_list.append(d)
    filt = copy.deepcopy(data_list)
    sched = g["sample_sched"].numpy()
    with injected_noise(seed=g["sample_noise_seed"]):
        out, sconf = sampling(data_list=data_list, model=smodel, inference_steps=3, tr_schedule=sched, rot_schedule=sched,
                              tor_schedule=sched, device=dev, t_to_sigma=t2s, model_args=sargs, batch_size=2,
                              confidence_model=cmodel, filtering_data_list=filt, filtering_model_args=cargs)
    for d, want in zip(out, g["sample_final"]):
        assert rmsd(d["ligand"].pos, want) < 1e-3
    assert torch.allclose(sconf.cpu(), g["sample_confidence"], atol=1e-4)


@pytest.mark.parametrize("mode", ["confidence", "score_lmax2"])
def test_all_atom_model_vs_oracle(mode):
    """Shipped confidence hyper-parameters (ns 24, nv 6, 5 layers, lmax 2, 9 groups, 1280-d LM), and an all-atom
    SCORE model with lmax 2 (torsion head through the l<=1 blocks of FullTensorProduct(sh, 2e))."""
    from confidence_bootstrapping_b200 import so3, torus
    from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    dev = torch.device("cuda")
    if mode == "confidence":
        args, cm, t = confidence_model_args(), True, 0.0
    else:
        args, cm, t = score_model_args(all_atoms=True, sh_lmax=2, ns=24, nv=6, num_conv_layers=3, num_prot_emb_layers=1), False, 0.5
    t2s = partial(t2s_full, args=args) if not cm else None
    torch.manual_seed(5)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=cm)
    randomize_norm_stats(model, seed=6)
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    hp = om.hyper_from_args(args, confidence_mode=cm)
    graphs = [make_complex(90 + i, 50 + 20 * i, 11 + 4 * i, all_atoms=True) for i in range(2)]
    cpu = Batch.from_data_list(copy.deepcopy(graphs))
    osamp.set_time(cpu, t, t, t, 2, all_atoms=True)
    gpu = Batch.from_data_list(copy.deepcopy(graphs)).to(dev)
    set_time(gpu, None, t, t, t, 2, True, False, dev)
    with torch.no_grad():
        want = om.aa_forward(sd, hp, cpu, partial(osamp.t_to_sigma, args=args) if not cm else None, so3.score_norm, torus.score_norm)
        got = model(gpu)
    n = 2 if cm else 3
    for a, b in zip(got[:n], want[:n]):
        assert a.shape == b.shape
        if cm:
            assert torch.allclose(a.cpu(), b, atol=1e-4)      # confidences within 1e-4 (BASELINE.json)
        else:
            assert rel_err(a, b) < 1e-4


@pytest.mark.parametrize("far", ["one", "all"])
def test_confidence_model_when_crop_beyond_removes_a_whole_receptor(far):
    """Edge case of the filtering leg (sampling.py:226-233 -> utils.py:395-420): a pose that drifted away from the pocket leaves
    its graph with NO receptor residue / atom after crop_beyond -- for one graph of the batch, or for all of them (empty
    receptor tensors).  The device path must score such batches like the reference arithmetic does (oracle), not fail."""
    from confidence_bootstrapping_b200.configs import confidence_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import crop_beyond, get_model
    dev = torch.device("cuda")
    args = confidence_model_args()
    torch.manual_seed(5)
    model = get_model(args, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True)
    randomize_norm_stats(model, seed=6)
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    hp = om.hyper_from_args(args, confidence_mode=True)
    graphs = [make_complex(95 + i, 45 + 10 * i, 9 + 3 * i, all_atoms=True) for i in range(2)]
    for i, g in enumerate(graphs):
        if far == "all" or i == 1:
            g["ligand"].pos = g["ligand"].pos + 500.0
    gpu = crop_beyond(Batch.from_data_list(copy.deepcopy(graphs)).to(dev), args.crop_beyond, True)
    assert gpu["receptor"].num_nodes == (0 if far == "all" else gpu["receptor"].num_nodes) and (far == "one" or gpu["atom"].num_nodes == 0)
    set_time(gpu, 0, 0, 0, 0, 2, True, False, dev)
    with torch.no_grad():
        conf = model(gpu)[0]
    assert conf.shape[0] == 2 and bool(torch.isfinite(conf).all())
    cpu = Batch.from_data_list([osamp.crop_beyond(copy.deepcopy(g), args.crop_beyond, True) for g in graphs])
    osamp.set_time(cpu, 0, 0, 0, 2, all_atoms=True)
    with torch.no_grad():
        want = om.aa_forward(sd, hp, cpu, None, None, None)[0]
    assert torch.allclose(conf.cpu(), want, atol=1e-4)


@pytest.mark.gpu
def test_e_post_fold_matches_kernel_path():
    """The per-graph edge-embedding offset (rec_sigma_emb) folded into the node projection on the host equals the
    kernel's own e_post path (cb_tp_segment.e_post)."""
    import confidence_bootstrapping_b200.tensor_layers as tl
    from functools import partial
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    from helpers import small_score_args
    args = small_score_args()
    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = get_model(args, dev, t_to_sigma=partial(t_to_sigma, args=args), no_parallel=True).eval()
    dl = [make_complex(5 + i, 60, 14, all_atoms=False, lm_dim=0) for i in range(3)]
    outs = []
    for fold in (True, False):
        tl.FOLD_E_POST = fold
        try:
            batch = Batch.from_data_list(dl).to(dev)
            set_time(batch, None, 0.4, 0.4, 0.4, batch.num_graphs, False, False, dev)
            with torch.no_grad():
                outs.append(model(batch))
        finally:
            tl.FOLD_E_POST = True
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert rel_err(a, b) < 2e-6


@pytest.mark.gpu
def test_sampling_is_bit_reproducible():
    """Same inputs + same injected noise => identical bits: no floating-point atomics anywhere on the path."""
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, _ = _build(args, seed=3)
    g = Batch.from_data_list([make_complex(78, 70, 15, all_atoms=False)])
    np.random.seed(1)
    torch.manual_seed(1)
    dl0 = [copy.deepcopy(g) for _ in range(4)]
    randomize_position(dl0, False, False, args.tr_sigma_max)
    sched = get_t_schedule("expbeta", 6, 1, 1)
    outs = []
    for _ in range(3):
        dl = copy.deepcopy(dl0)
        with injected_noise(seed=5):
            out, _ = sampling(data_list=dl, model=model, inference_steps=6, tr_schedule=sched, rot_schedule=sched,
                              tor_schedule=sched, device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=4)
        outs.append(torch.stack([d["ligand"].pos for d in out]))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.gpu
def test_step_graph_reuse_is_bit_identical_and_follows_the_weights():
    """A later batch of the same complex replays the cached step graph of the first one (no eager step, no capture): same
    bits as a run with the cache off, and a weight change between the calls is honoured (no stale graph)."""
    from confidence_bootstrapping_b200 import sampling as smp
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, _ = _build(args, seed=4)
    g = Batch.from_data_list([make_complex(79, 60, 12, all_atoms=False)])
    sched = get_t_schedule("expbeta", 6, 1, 1)

    def run(seed, cache):
        np.random.seed(seed)
        torch.manual_seed(seed)
        dl = [copy.deepcopy(g) for _ in range(4)]
        randomize_position(dl, False, False, args.tr_sigma_max)
        old = smp.GRAPH_CACHE_SIZE
        smp.GRAPH_CACHE_SIZE = 2 if cache else 0
        try:
            with injected_noise(seed=seed + 100):
                out, _ = sampling(data_list=dl, model=model, inference_steps=6, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched,
                                  device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=4, no_final_step_noise=True)
        finally:
            smp.GRAPH_CACHE_SIZE = old
        return torch.stack([d["ligand"].pos for d in out]).cpu()

    smp._graph_cache.clear()
    ref = [run(s, cache=False) for s in (1, 2, 3)]
    h0 = smp.graph_cache_hits
    got = [run(s, cache=True) for s in (1, 2, 3)]
    assert smp.graph_cache_hits == h0 + 2                      # the first call captures, the next two replay
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    # EMA-style in-place weight change: the cached graph (which points at the old folded weights) must not be replayed
    w = max(model.conv_layers[1].parameters(), key=lambda p: p.numel())          # a second-Linear weight of the radial MLP
    saved = w.data.clone()
    w.data.mul_(1.5)
    try:
        moved_ref = run(2, cache=False)
        h1 = smp.graph_cache_hits
        moved = run(2, cache=True)
        assert smp.graph_cache_hits == h1
        assert torch.equal(moved, moved_ref) and not torch.equal(moved, ref[1])
    finally:
        w.data.copy_(saved)
    smp._graph_cache.clear()


@pytest.mark.gpu
def test_capture_before_the_first_step_is_bit_identical():
    """Once the model is warm, a NEW complex builds its static tables, captures the step graph before step 0 and replays all
    steps (no eager forward): same bits as the eager-first-step path."""
    from confidence_bootstrapping_b200 import sampling as smp
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, _ = _build(args, seed=8)
    sched = get_t_schedule("expbeta", 5, 1, 1)
    gs = [Batch.from_data_list([make_complex(300 + i, 50 + 15 * i, 10 + 3 * i, all_atoms=False)]) for i in range(3)]

    def run_all(first_step_capture):
        old = smp.CAPTURE_FIRST_STEP
        smp.CAPTURE_FIRST_STEP = first_step_capture
        smp._graph_cache.clear()
        outs = []
        try:
            for i, g in enumerate(gs):
                np.random.seed(i)
                torch.manual_seed(i)
                dl = [copy.deepcopy(g) for _ in range(3)]
                randomize_position(dl, False, False, args.tr_sigma_max)
                with injected_noise(seed=40 + i):
                    out, _ = sampling(data_list=dl, model=model, inference_steps=5, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched,
                                      device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=3)
                outs.append(torch.stack([d["ligand"].pos for d in out]).cpu())
        finally:
            smp.CAPTURE_FIRST_STEP = old
        return outs

    want = run_all(False)
    assert getattr(model, "_cb200_warm", None) is not None
    got = run_all(True)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_multi_batch_sampling_with_two_streams_equals_the_single_stream_path():
    """sampling() over several batches of one complex (samples_per_complex > batch_size, what inference.py does) with the
    confidence model: batches 2.. replay the cached step graph and every filtering leg runs one batch behind on the second
    stream.  Poses and confidences must be bit-identical to the plain path (no graph cache, one stream)."""
    from confidence_bootstrapping_b200 import sampling as smp
    from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    dev = torch.device("cuda")
    args, cargs = score_model_args(), confidence_model_args()
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(0)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
    torch.manual_seed(1)
    cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True).eval()
    g = Batch.from_data_list([make_complex(431, 90, 14, all_atoms=True)])
    sched = get_t_schedule("expbeta", 5, 1, 1)

    def run(pipelined):
        np.random.seed(7)
        torch.manual_seed(7)
        dl = [copy.deepcopy(g) for _ in range(9)]
        randomize_position(dl, False, False, args.tr_sigma_max)
        fl = copy.deepcopy(dl)
        saved = (smp.PIPELINE_CONFIDENCE, smp.GRAPH_CACHE_SIZE)
        smp.PIPELINE_CONFIDENCE, smp.GRAPH_CACHE_SIZE = (True, 1) if pipelined else (False, 0)
        smp._graph_cache.clear()
        h0 = smp.graph_cache_hits
        try:
            with injected_noise(seed=77):
                out, conf = sampling(data_list=dl, model=model, inference_steps=5, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched,
                                     device=dev, t_to_sigma=t2s, model_args=args, batch_size=3, confidence_model=cmodel,
                                     filtering_data_list=fl, filtering_model_args=cargs)
        finally:
            smp.PIPELINE_CONFIDENCE, smp.GRAPH_CACHE_SIZE = saved
        return torch.stack([d["ligand"].pos for d in out]).cpu(), conf.cpu(), smp.graph_cache_hits - h0

    pos0, conf0, hits0 = run(False)
    pos1, conf1, hits1 = run(True)
    assert hits0 == 0 and hits1 == 2                    # 3 batches: capture, then two replays of the cached graph
    assert conf1.shape == (9,) and torch.equal(pos0, pos1) and torch.equal(conf0, conf1)
    smp._graph_cache.clear()


@pytest.mark.gpu
def test_sampling_many_equals_the_loop_of_sampling_calls():
    """sampling_many (the callers' per-complex loop as one pipelined call: filtering leg of complex i on the second stream
    while complex i+1 is collated / captured / stepped) returns exactly what separate sampling() calls return."""
    from confidence_bootstrapping_b200 import sampling as smp
    from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling, sampling_many
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    dev = torch.device("cuda")
    args, cargs = score_model_args(), confidence_model_args()
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(0)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
    torch.manual_seed(1)
    cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True).eval()
    sizes = [(120, 12), (60, 9), (200, 20), (60, 9)]
    graphs = [Batch.from_data_list([make_complex(520 + i, nr, nl, all_atoms=True)]) for i, (nr, nl) in enumerate(sizes)]
    sched = get_t_schedule("expbeta", 5, 1, 1)
    kw = dict(model=model, inference_steps=5, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s,
              model_args=args, batch_size=4, confidence_model=cmodel, filtering_model_args=cargs)

    def inputs():
        reqs = []
        for i, g in enumerate(graphs):
            np.random.seed(i)
            torch.manual_seed(i)
            dl = [copy.deepcopy(g) for _ in range(4 if i != 2 else 8)]       # complex 2: two batches
            randomize_position(dl, False, False, args.tr_sigma_max)
            reqs.append((dl, copy.deepcopy(dl)))
        return reqs

    smp._graph_cache.clear()
    with injected_noise(seed=21):
        want = [sampling(data_list=dl, filtering_data_list=fl, **kw) for dl, fl in inputs()]
    smp._graph_cache.clear()
    with injected_noise(seed=21):
        got = sampling_many(inputs(), **kw)
    assert len(got) == len(want)
    for (o1, c1), (o0, c0) in zip(got, want):
        assert torch.equal(torch.stack([d["ligand"].pos for d in o1]), torch.stack([d["ligand"].pos for d in o0]))
        assert torch.equal(c1, c0)
    # bare data lists (no confidence model) work too, and unknown keywords are refused
    with injected_noise(seed=22):
        plain = sampling_many([dl for dl, _ in inputs()], **{**kw, "confidence_model": None})
    assert all(c is None for _, c in plain)
    with pytest.raises(TypeError):
        sampling_many(inputs(), **kw, no_such_flag=1)
    smp._graph_cache.clear()


@pytest.mark.gpu
def test_dead_output_gates_do_not_change_the_scores():
    """The per-layer receptor keep masks (score_model._dead_output_gates) only skip rows nobody reads: the scores are
    bit-identical with and without them."""
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.score_model import TensorProductScoreModel
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, _ = _build(args, seed=4)
    dl = [make_complex(30 + i, 120 + 40 * i, 12 + 3 * i, all_atoms=False) for i in range(3)]
    outs, kept = [], []
    orig = TensorProductScoreModel._dead_output_gates
    for use in (True, False):
        def gates(st, rl, n_layers, hops=4, _use=use):
            g = orig(st, rl, n_layers, hops) if _use else {}
            kept.append({k: float(v.float().mean()) for k, v in g.items()})
            return g
        TensorProductScoreModel._dead_output_gates = staticmethod(gates)
        try:
            batch = Batch.from_data_list(dl).to("cuda")
            set_time(batch, None, 0.05, 0.05, 0.05, batch.num_graphs, False, False, torch.device("cuda"))   # small sigma: short cross cutoff
            with torch.no_grad():
                outs.append(model(batch))
        finally:
            TensorProductScoreModel._dead_output_gates = staticmethod(orig)
    assert kept[0] and min(kept[0].values()) < 1.0, kept       # something was actually pruned
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_replicated_receptor_embedding_is_tiled():
    """A batch collated on the device from N copies of one complex carries the 'receptor is replicated' flag; the score
    model then embeds the receptor once and tiles it.  Same scores as the per-copy embedding."""
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.sampling import randomize_position
    from confidence_bootstrapping_b200.synthetic import make_complex
    args = score_model_args()
    model, t2s, _ = _build(args, seed=5)
    g = Batch.from_data_list([make_complex(41, 350, 18, all_atoms=False)])
    np.random.seed(2); torch.manual_seed(2)
    dl = [copy.deepcopy(g) for _ in range(4)]
    randomize_position(dl, False, False, args.tr_sigma_max)
    outs = []
    for device in ("cuda", None):
        batch = Batch.from_data_list(copy.deepcopy(dl), device=device)
        assert ("receptor" in batch._g["_replicated_types"]) == (device is not None)
        batch = batch.to("cuda")
        set_time(batch, None, 0.3, 0.3, 0.3, batch.num_graphs, False, False, torch.device("cuda"))
        with torch.no_grad():
            outs.append(model(batch))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert rel_err(a, b) < 1e-6


@pytest.mark.gpu
def test_shared_first_layer_receptor_messages_match_the_general_path():
    """Conv layer 0: the rec->rec slot of the S copies of one receptor is aggregated + transformed once for the first copy
    and added through cb_tp_conv_args.pre_sum.  Same scores as the general path up to the order of the fp32 additions
    (the shared sum joins the per-sample sums after the transform instead of inside its register accumulators), and a
    16x smaller layer-0 item count."""
    import confidence_bootstrapping_b200.score_model as sm
    from confidence_bootstrapping_b200 import _lib
    from confidence_bootstrapping_b200.configs import score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.sampling import randomize_position
    from confidence_bootstrapping_b200.synthetic import make_complex
    from helpers import blockwise_err
    args = score_model_args()
    model, t2s, _ = _build(args, seed=7)
    g = Batch.from_data_list([make_complex(43, 300, 22, all_atoms=False)])
    np.random.seed(3); torch.manual_seed(3)
    dl = [copy.deepcopy(g) for _ in range(16)]
    randomize_position(dl, False, False, args.tr_sigma_max)
    outs, items = [], []

    class Count:
        def __init__(self):
            self.n = []

        def __call__(self, a):
            import contextlib
            self.n.append(_lib.tp_conv_items(a))
            return contextlib.nullcontext()

    for share in (True, False):
        sm.SHARE_REPLICATED_REC_MESSAGES = share
        try:
            for t in (1.0, 0.2):
                batch = Batch.from_data_list(copy.deepcopy(dl), device="cuda")
                set_time(batch, None, t, t, t, batch.num_graphs, False, False, torch.device("cuda"))
                c = Count()
                _lib.tp_conv_hook = c
                with torch.no_grad():
                    outs.append(model(batch))
                _lib.tp_conv_hook = None
                items.append(sum(c.n))
        finally:
            sm.SHARE_REPLICATED_REC_MESSAGES = True
            _lib.tp_conv_hook = None
    for k in range(2):
        for a, b in zip(outs[k][:3], outs[k + 2][:3]):
            assert blockwise_err(a, b) < 1e-5      # two fp32 summation orders of the same terms (measured 2e-6), 10x below the forward bar
        assert items[k] < items[k + 2]


def _two_gpu_worker(rank, world, port, q):
    """One rank of the 2-GPU run of test_sharded_sampling_on_two_gpus_matches_one_gpu (spawned process)."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        poses, confs = _sharded_sample(torch.device("cuda", rank))
        q.put((rank, [p.cpu() for p in poses], [c.cpu() for c in confs]))
    finally:
        dist.destroy_process_group()


def _sharded_sample(dev):
    from confidence_bootstrapping_b200 import dist as cbdist
    from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    args, cargs = score_model_args(), confidence_model_args()
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(0)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
    torch.manual_seed(1)
    cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True).eval()
    sizes = [(260, 21), (150, 10), (410, 34), (200, 15), (330, 27)]
    complexes = [Batch.from_data_list([make_complex(800 + i, nr, nl, all_atoms=True)]) for i, (nr, nl) in enumerate(sizes)]
    sched = get_t_schedule("expbeta", 4, 1, 1)

    def sample_fn(c, n):
        i = next(k for k, x in enumerate(complexes) if x is c)
        np.random.seed(i)
        torch.manual_seed(i)
        dl = [copy.deepcopy(c) for _ in range(n)]
        randomize_position(dl, False, False, args.tr_sigma_max)
        fl = copy.deepcopy(dl)
        with injected_noise(seed=50 + i):
            out, conf = sampling(data_list=dl, model=model, inference_steps=4, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched,
                                 device=dev, t_to_sigma=t2s, model_args=args, batch_size=n, confidence_model=cmodel,
                                 filtering_data_list=fl, filtering_model_args=cargs)
        return torch.stack([d["ligand"].pos for d in out]), conf

    return cbdist.sample_complexes(complexes, 3, sample_fn)


@pytest.mark.gpu
def test_sharded_sampling_on_two_gpus_matches_one_gpu():
    """dist.sample_complexes on real GPUs with real payloads -- LPT shards over 2 ranks (NCCL, one process per
    GPU), variable-length all-gather of poses + confidences; every rank ends up with exactly the single-GPU results (the whole
    path is bit-reproducible, so equality is exact)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    single_p, single_c = _sharded_sample(torch.device("cuda", 0))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, poses, confs in res:
        for a, b in zip(poses, single_p):
            assert torch.equal(a, b.cpu())
        for a, b in zip(confs, single_c):
            assert torch.equal(a, b.cpu())
