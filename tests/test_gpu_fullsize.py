"""Parity at the sizes bench.py measures (BASELINE.json configs 2, 3 and 5), not only at toy sizes.

VERDICT r1: "No parity check at any BASELINE size".  These tests run the exact bench batch (400 residues x 40 ligand
atoms x 40 samples: 17 600 nodes, ~0.9 M edges per conv layer, persistent K3 CTAs wrapping > 100 items each, a
multi-GB accumulator workspace) through the CUDA path and compare with the CPU oracle, which processes the same
graphs in slices of 8 (every graph of a batch is independent in eval mode; the slices only bound the oracle's
[E, weight_numel] memory).

Tolerances (written where they are asserted):
  * single TP-conv layer: 1e-5, element-wise per irreps block (helpers.blockwise_err) -- BASELINE.json north_star;
  * whole forward = 11 TP-conv layers + heads composed: 1e-4 element-wise per output (the per-layer 1e-5 does not
    survive an 11-layer composition of fp32 kernels with different summation orders; stated next to each assert);
  * poses after 20 reverse-diffusion steps with identical noise: 1e-3 A RMSD; confidences: 1e-4 absolute.
"""
import copy
from functools import partial

import numpy as np
import pytest
import torch

from helpers import (blockwise_err, injected_noise, injected_sliced_noise, irreps_blocks, randomize_norm_stats, rmsd)
from oracle import model as om, o3, sampler as osamp

pytestmark = pytest.mark.gpu

N_RES, N_LIG, SAMPLES = 400, 40, 40      # bench.py's workload (BASELINE.json configs[1])
FORWARD_TOL = 1e-4                       # whole-forward bar, element-wise (see module docstring)
N_ORACLE = 16                            # graphs of the 40-graph bench batch the CPU oracle re-computes (2 slices of 8)


def _models(seed=0, confidence=False):
    from confidence_bootstrapping_b200 import so3, torus
    from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
    from confidence_bootstrapping_b200.diffusion_utils import t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.utils import get_model
    dev = torch.device("cuda")
    args = score_model_args()
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(seed)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True)
    randomize_norm_stats(model, seed=seed + 1)
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    hp = om.hyper_from_args(args)
    o_t2s = partial(osamp.t_to_sigma, args=args)
    fwd = lambda b: om.cg_forward(sd, hp, b, o_t2s, so3.score_norm, torus.score_norm)
    out = dict(args=args, model=model, t2s=t2s, oracle_fwd=fwd, o_t2s=o_t2s)
    if confidence:
        cargs = confidence_model_args()
        torch.manual_seed(seed + 10)
        cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True)
        randomize_norm_stats(cmodel, seed=seed + 11)
        cmodel.eval()
        csd = {k: v.detach().cpu().clone() for k, v in cmodel.state_dict().items()}
        chp = om.hyper_from_args(cargs, confidence_mode=True)
        out.update(cargs=cargs, cmodel=cmodel,
                   oracle_conf=lambda b: om.aa_forward(csd, chp, b, None, so3.score_norm, torus.score_norm))
    return out


def _bench_data_list(seed, args, all_atoms):
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.sampling import randomize_position
    from confidence_bootstrapping_b200.synthetic import make_complex
    g = Batch.from_data_list([make_complex(seed, N_RES, N_LIG, all_atoms=all_atoms)])
    np.random.seed(seed)
    torch.manual_seed(seed)
    dl = [copy.deepcopy(g) for _ in range(SAMPLES)]
    randomize_position(dl, args.no_torsion, False, args.tr_sigma_max)
    return dl


def _oracle_forward_sliced(fwd, graphs, t, per=8, all_atoms=False):
    from confidence_bootstrapping_b200.data import Batch
    outs = []
    for i in range(0, len(graphs), per):
        part = copy.deepcopy(graphs[i:i + per])
        b = Batch.from_data_list(part)
        osamp.set_time(b, t, t, t, len(part), all_atoms=all_atoms)
        with torch.no_grad():
            outs.append(fwd(b))
    return [torch.cat([o[k] for o in outs]) for k in range(len(outs[0])) if outs[0][k] is not None and torch.is_tensor(outs[0][k])]


@pytest.mark.parametrize("t", [1.0, 0.5, 0.05])
def test_bench_batch_score_forward_vs_oracle(t):
    """ONE score-model forward of the exact bench batch (all-pairs cross graph at t = 1, short cutoff + dead-output gates
    at t = 0.05) against oracle.cg_forward; every output element within FORWARD_TOL of the oracle on its block's scale."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    m = _models(seed=0)
    dl = _bench_data_list(1000, m["args"], all_atoms=False)
    want = _oracle_forward_sliced(m["oracle_fwd"], dl[:N_ORACLE], t)      # the oracle checks the first N_ORACLE of the 40 graphs
    gpu = Batch.from_data_list(copy.deepcopy(dl), device="cuda")
    set_time(gpu, None, t, t, t, SAMPLES, False, False, torch.device("cuda"))
    with torch.no_grad():
        got = m["model"](gpu)
    assert got[0].shape == (SAMPLES, 3) and got[1].shape == (SAMPLES, 3) and got[2].numel() % SAMPLES == 0
    for a, b, name in zip(got[:3], want[:3], ("tr", "rot", "tor")):
        a = a[: b.shape[0]]
        err = blockwise_err(a, b)
        assert err < FORWARD_TOL, (name, t, err)
    assert all(torch.isfinite(x).all() for x in got[:3])


def test_bench_batch_sampling_and_confidence_vs_oracle():
    """The bench step itself: 40 samples x 20 reverse steps + crop_beyond(20 A) + all-atom confidence scoring on the GPU,
    against the oracle run on the first 8 samples with the same noise stream (SlicedNoiseTape): final poses within
    1e-3 A RMSD (BASELINE.json), confidences within 1e-4."""
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import sampling
    m = _models(seed=2, confidence=True)
    args, cargs = m["args"], m["cargs"]
    dl = _bench_data_list(1500, args, all_atoms=True)
    n_ref = 8
    dl_cpu, fl_cpu = copy.deepcopy(dl[:n_ref]), copy.deepcopy(dl[:n_ref])
    fl = copy.deepcopy(dl)
    steps = 20
    sched = get_t_schedule("expbeta", steps, 1, 1)
    dev = torch.device("cuda")
    with injected_noise(seed=21):
        out, conf = sampling(data_list=dl, model=m["model"], inference_steps=steps, tr_schedule=sched, rot_schedule=sched,
                             tor_schedule=sched, device=dev, t_to_sigma=m["t2s"], model_args=args, batch_size=SAMPLES,
                             confidence_model=m["cmodel"], filtering_data_list=fl, filtering_model_args=cargs)
    with injected_sliced_noise(21, n_ref, SAMPLES):
        ref, rconf = osamp.sampling(dl_cpu, m["oracle_fwd"], steps, sched, sched, sched, m["o_t2s"], args, batch_size=n_ref,
                                    confidence_forward=m["oracle_conf"], filtering_data_list=fl_cpu,
                                    filtering_model_args=cargs, crop_fn=osamp.crop_beyond)
    worst = max(rmsd(a["ligand"].pos, b["ligand"].pos) for a, b in zip(out[:n_ref], ref))
    assert worst < 1e-3, worst
    assert conf.shape == (SAMPLES,)
    assert torch.allclose(conf[:n_ref].cpu(), rconf, atol=1e-4), (conf[:n_ref].cpu() - rconf).abs().max()


def test_bench_batch_confidence_forward_vs_oracle():
    """ONE all-atom confidence forward (lmax 2, 9 edge groups) of the cropped bench batch: 40 poses placed in the pocket,
    crop_beyond(20 A) on the device vs per-graph on the host, confidences within 1e-4 (BASELINE.json)."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time
    from confidence_bootstrapping_b200.utils import crop_beyond
    m = _models(seed=4, confidence=True)
    cargs = m["cargs"]
    dl = _bench_data_list(1700, m["args"], all_atoms=True)
    g = torch.Generator().manual_seed(3)
    for d in dl:      # poses near the pocket centre (what the sampler converges to), not N(0, 19 A) away from it
        d["ligand"].pos = d["ligand"].pos - d["ligand"].pos.mean(0, keepdim=True) + torch.randn(1, 3, generator=g) * 2.0
    cropped = [osamp.crop_beyond(copy.deepcopy(d), cargs.crop_beyond, True) for d in dl[:N_ORACLE]]
    want = _oracle_forward_sliced(m["oracle_conf"], cropped, 0.0, per=8, all_atoms=True)
    gpu = crop_beyond(Batch.from_data_list(copy.deepcopy(dl), device="cuda"), cargs.crop_beyond, True)
    assert "receptor" not in gpu._g["_replicated_types"]       # every copy kept a different residue subset
    set_time(gpu, 0, 0, 0, 0, SAMPLES, True, False, torch.device("cuda"))
    with torch.no_grad():
        conf, atom_conf = m["cmodel"](gpu)
    assert conf.shape == (SAMPLES,) and torch.isfinite(conf).all()
    assert torch.allclose(conf[:N_ORACLE].cpu(), want[0], atol=1e-4), (conf[:N_ORACLE].cpu() - want[0]).abs().max()
    assert torch.allclose(atom_conf[: want[1].shape[0]].cpu(), want[1], atol=1e-4)


def test_config3_slice_sharded_sampling_vs_oracle():
    """BASELINE config 3 slice: three heterogeneous complexes (one with >= 900 residues, forced through K3's node-chunked
    workspace path) x 3 samples, 5 reverse steps + confidence, through dist.sample_complexes; poses <= 1e-3 A and
    confidences <= 1e-4 vs the oracle."""
    import confidence_bootstrapping_b200.tensor_layers as tl
    from confidence_bootstrapping_b200 import dist as cbdist
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    m = _models(seed=6, confidence=True)
    args, cargs = m["args"], m["cargs"]
    sizes = [(930, 52), (180, 11), (610, 33)]
    complexes = [Batch.from_data_list([make_complex(300 + i, nr, nl, all_atoms=True)]) for i, (nr, nl) in enumerate(sizes)]
    S, steps = 3, 5
    sched = get_t_schedule("expbeta", steps, 1, 1)
    dev = torch.device("cuda")
    starts = {}

    def start_poses(i):
        np.random.seed(40 + i)
        torch.manual_seed(40 + i)
        dl = [copy.deepcopy(complexes[i]) for _ in range(S)]
        randomize_position(dl, False, False, args.tr_sigma_max)
        return dl

    def sample_fn(c, n):
        i = next(k for k, x in enumerate(complexes) if x is c)
        dl = start_poses(i)
        fl = copy.deepcopy(dl)
        with injected_noise(seed=70 + i):
            out, conf = sampling(data_list=dl, model=m["model"], inference_steps=steps, tr_schedule=sched, rot_schedule=sched,
                                 tor_schedule=sched, device=dev, t_to_sigma=m["t2s"], model_args=args, batch_size=n,
                                 confidence_model=m["cmodel"], filtering_data_list=fl, filtering_model_args=cargs)
        return torch.stack([d["ligand"].pos for d in out]), conf

    old = tl.WORKSPACE_BYTES
    tl.WORKSPACE_BYTES = 192 << 20         # the 930-residue complex needs ~0.5 GB of accumulators: >= 3 node chunks
    try:
        poses, confs = cbdist.sample_complexes(complexes, S, sample_fn)
    finally:
        tl.WORKSPACE_BYTES = old
    for i in range(len(complexes)):
        dl = start_poses(i)
        fl = copy.deepcopy(dl)
        with injected_noise(seed=70 + i):
            ref, rconf = osamp.sampling(dl, m["oracle_fwd"], steps, sched, sched, sched, m["o_t2s"], args, batch_size=S,
                                        confidence_forward=m["oracle_conf"], filtering_data_list=fl,
                                        filtering_model_args=cargs, crop_fn=osamp.crop_beyond)
        assert poses[i].shape == (S, sizes[i][1], 3)
        for s in range(S):
            assert rmsd(poses[i][s], ref[s]["ligand"].pos) < 1e-3, (i, s)
        assert torch.allclose(confs[i].cpu(), rconf, atol=1e-4), (i, (confs[i].cpu() - rconf).abs().max())


def test_config5_slice_all_atom_score_model_vs_oracle():
    """BASELINE config 5 slice: the all-atom SCORE model (all_atom_score_model.py:363-507, lmax 2, 9 edge groups, final_conv +
    torsion head) on a 520-residue receptor (~4 100 atoms) with a 64-atom ligand, 2 samples, t = 0.5 and t = 1.0.
    Hyper-parameters: the confidence YAML's sizes (ns 24, nv 6, 5 conv layers) with confidence_mode=False."""
    from confidence_bootstrapping_b200 import so3, torus
    from confidence_bootstrapping_b200.configs import all_atom_score_model_args
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import randomize_position
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    dev = torch.device("cuda")
    args = all_atom_score_model_args()
    t2s = partial(t2s_full, args=args)
    torch.manual_seed(8)
    model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True)
    randomize_norm_stats(model, seed=9)
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    hp = om.hyper_from_args(args)
    g = Batch.from_data_list([make_complex(500, 520, 64, all_atoms=True)])
    np.random.seed(5)
    torch.manual_seed(5)
    dl = [copy.deepcopy(g) for _ in range(2)]
    randomize_position(dl, False, False, 6.0)       # within reach of the 5 A ligand-atom graph for some atoms
    o_t2s = partial(osamp.t_to_sigma, args=args)
    for t in (0.5, 1.0):
        cpu = Batch.from_data_list(copy.deepcopy(dl))
        osamp.set_time(cpu, t, t, t, 2, all_atoms=True)
        cpu64 = Batch.from_data_list(copy.deepcopy(dl))
        osamp.set_time(cpu64, t, t, t, 2, all_atoms=True)
        sd64, cpu64 = om.to_double(sd, cpu64)
        gpu = Batch.from_data_list(copy.deepcopy(dl), device="cuda")
        set_time(gpu, None, t, t, t, 2, True, False, dev)
        with torch.no_grad():
            want = om.aa_forward(sd, hp, cpu, o_t2s, so3.score_norm, torus.score_norm)
            want64 = om.aa_forward(sd64, hp, cpu64, o_t2s, so3.score_norm, torus.score_norm)      # same edges, fp64 arithmetic
            got = model(gpu)
        for a, b, b64, name in zip(got[:3], want[:3], want64[:3], ("tr", "rot", "tor")):
            assert a.shape == b.shape
            # The heads of a random-init lmax-2 model are differences of O(1) terms: the fp32 oracle itself sits 3e-5 .. 8e-5
            # (element-wise) from the fp64 evaluation of the same graph.  Bar: FORWARD_TOL against the fp64 yardstick, or
            # twice the fp32 oracle's own distance to it where that is larger (i.e. not worse than fp32 rounding noise).
            noise = blockwise_err(b, b64)
            err = blockwise_err(a, b64)
            assert err < max(FORWARD_TOL, 2.0 * noise), (name, t, err, noise)


def test_tor_bond_conv_layer_vs_oracle():
    """The torsion head's convolution (score_model.py:266-274): 74-dim node features (x) the 20-dim output of
    FullTensorProduct(sh, 2e) = 1x1o+1x2o+1x2e+1x3o -> 32x0o + 32x0e, radial MLP 96->96->384, no residual.
    cb200 evaluates only the 1o block of the edge harmonics (the only one that can reach scalars from l <= 1 inputs);
    the oracle gets the full 20 columns.  1e-5 element-wise per irreps block."""
    from confidence_bootstrapping_b200.irreps import full_tp_low_blocks
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    in_ir, out_ir = "32x0e + 6x1o + 6x1e + 6x0o", "32x0o + 32x0e"
    full_sh = "1x1o + 1x2o + 1x2e + 1x3o"
    low_sh, blocks = full_tp_low_blocks(1)
    assert [l for l, _ in blocks] == [1]             # lmax 1: one 1o block survives
    torch.manual_seed(0)
    layer = TensorProductConvLayer(in_ir, low_sh, out_ir, 96, residual=False, batch_norm=True, dropout=0.1, hidden_features=96)
    randomize_norm_stats(layer, seed=1)
    layer.eval()
    n_nodes, n_edges, n_out = 160, 2600, 23          # 23 rotatable bonds aggregating ligand atoms
    x = torch.randn(n_nodes, 74)
    gen = torch.Generator().manual_seed(2)
    ei = torch.stack([torch.randint(0, n_out, (n_edges,), generator=gen), torch.randint(0, n_nodes, (n_edges,), generator=gen)])
    sh20 = torch.randn(n_edges, 20)
    ea = torch.randn(n_edges, 96)
    sd = {"x." + k: v.clone() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        want = om.tp_conv_layer(sd, "x", in_ir, o3.Irreps(full_sh), out_ir, False, 1, False, True, x, ei, ea, sh20, out_nodes=n_out)
        layer = layer.cuda()
        got = layer(x.cuda(), ei.cuda(), ea.cuda(), sh20[:, :3].contiguous().cuda(), out_nodes=n_out)
    assert got.shape == want.shape == (n_out, 64)
    assert blockwise_err(got, want, irreps_blocks(out_ir)) < 1e-5


SEQ3 = "32x0e + 6x1o + 6x1e + 6x0o"
CONF3 = "24x0e + 6x1o + 6x1e + 24x0o"


@pytest.mark.parametrize("in_ir,sh_l,out_ir,faster,groups,nef", [(SEQ3, 1, SEQ3, True, 4, 96), (CONF3, 2, CONF3, False, 9, 72)])
def test_tp_conv_layer_blockwise_at_scale(in_ir, sh_l, out_ir, faster, groups, nef):
    """The two conv-layer shapes that carry the bench (score 74->74 with 4 radial MLPs, confidence 84->84 with 9) on a
    graph big enough that every persistent accumulate CTA wraps over many (node, slot) items and the transform runs
    hundreds of tiles: 6 000 nodes, 120 000 edges, hub nodes with > 1 000 edges.  1e-5 element-wise PER IRREPS BLOCK."""
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    torch.manual_seed(0)
    sh_ir = "1x0e + 1x1o" if sh_l == 1 else "1x0e + 1x1o + 1x2e"
    layer = TensorProductConvLayer(in_ir, sh_ir, out_ir, nef, residual=True, batch_norm=True, dropout=0.1,
                                   hidden_features=nef, faster=faster, edge_groups=groups)
    randomize_norm_stats(layer, seed=1)
    layer.eval()
    n_nodes, n_edges = 6000, 120000
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n_nodes, o3.Irreps(in_ir).dim, generator=gen)
    ei = torch.stack([torch.randint(0, n_nodes, (n_edges,), generator=gen), torch.randint(0, n_nodes, (n_edges,), generator=gen)])
    ei[0, :4000] = torch.randint(0, 3, (4000,), generator=gen)          # three hub nodes
    sh = o3.spherical_harmonics(list(range(sh_l + 1)), torch.randn(n_edges, 3, generator=gen), True, "component")
    ea = torch.randn(n_edges, nef, generator=gen)
    bounds = np.linspace(0, n_edges, groups + 1).astype(int)
    ea_list = [ea[bounds[g]:bounds[g + 1]] for g in range(groups)]
    sd = {"x." + k: v.clone() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        want = om.tp_conv_layer(sd, "x", in_ir, o3.Irreps(sh_ir), out_ir, faster, groups, True, True, x, ei, ea_list, sh, out_nodes=n_nodes)
        layer = layer.cuda()
        got = layer(x.cuda(), ei.cuda(), [e.cuda() for e in ea_list], sh.cuda(), out_nodes=n_nodes)
    assert blockwise_err(got, want, irreps_blocks(out_ir)) < 1e-5


def test_get_model_default_wrapper_and_train_mode_sampling():
    """get_model's default no_parallel=False on CUDA exposes `.module` (finetune_train.py:177 samples
    through model.module) without registering the model as its own child -- .eval() / .state_dict() / .parameters() must
    terminate.  VERDICT r1: sampling() on a model left in train mode (finetune_train.py never calls .eval()) returns
    poses instead of raising; the model is back in train mode afterwards."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma as t2s_full
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from confidence_bootstrapping_b200.synthetic import make_complex
    from confidence_bootstrapping_b200.utils import get_model
    from helpers import small_score_args
    args = small_score_args()
    t2s = partial(t2s_full, args=args)
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = get_model(args, dev, t_to_sigma=t2s)                  # default: no_parallel=False
    assert model.module is model
    assert "module" not in dict(model.named_children())
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params > 0 and len(model.state_dict()) > 0
    model.eval()
    model.train()                                                 # the state finetune_train.py samples in
    g = Batch.from_data_list([make_complex(9, 40, 10, all_atoms=False, lm_dim=0)])
    np.random.seed(0)
    torch.manual_seed(0)
    dl = [copy.deepcopy(g) for _ in range(4)]
    randomize_position(dl, False, False, args.tr_sigma_max)
    ref_dl = copy.deepcopy(dl)
    sched = get_t_schedule("expbeta", 3, 1, 1)
    with injected_noise(seed=3), pytest.warns(UserWarning, match="train mode") if _first_train_warning() else _null():
        out, _ = sampling(data_list=dl, model=model.module, inference_steps=3, tr_schedule=sched, rot_schedule=sched,
                          tor_schedule=sched, device=dev, t_to_sigma=t2s, model_args=args, batch_size=4)
    assert model.training
    model.eval()
    with injected_noise(seed=3):
        ref, _ = sampling(data_list=ref_dl, model=model, inference_steps=3, tr_schedule=sched, rot_schedule=sched,
                          tor_schedule=sched, device=dev, t_to_sigma=t2s, model_args=args, batch_size=4)
    for a, b in zip(out, ref):
        assert torch.isfinite(a["ligand"].pos).all() and torch.equal(a["ligand"].pos, b["ligand"].pos)


def _first_train_warning():
    from confidence_bootstrapping_b200 import sampling as smp
    return not smp._warned_train_mode


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def test_config1_1a0q_sampling_and_confidence_vs_oracle():
    """BASELINE config 1: the reference's shipped complex data/1a0q (416 residues, 23 heavy ligand atoms, 11 torsions) with the
    shipped score / confidence hyper-parameters, 10 samples x 20 steps + crop_beyond(20 A) + confidence scoring, against the CPU
    oracle with the same noise: poses within 1e-3 A RMSD, confidences within 1e-4 (BASELINE.json north_star)."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
    from confidence_bootstrapping_b200.sampling import randomize_position, sampling
    from helpers import load_1a0q
    m = _models(seed=12, confidence=True)
    args, cargs = m["args"], m["cargs"]
    g = Batch.from_data_list([load_1a0q()])
    np.random.seed(11)
    torch.manual_seed(11)
    S, steps = 10, 20
    dl = [copy.deepcopy(g) for _ in range(S)]
    randomize_position(dl, False, False, args.tr_sigma_max)
    dl_cpu, fl, fl_cpu = copy.deepcopy(dl), copy.deepcopy(dl), copy.deepcopy(dl)
    sched = get_t_schedule("expbeta", steps, 1, 1)
    with injected_noise(seed=31):
        out, conf = sampling(data_list=dl, model=m["model"], inference_steps=steps, tr_schedule=sched, rot_schedule=sched,
                             tor_schedule=sched, device=torch.device("cuda"), t_to_sigma=m["t2s"], model_args=args, batch_size=S,
                             confidence_model=m["cmodel"], filtering_data_list=fl, filtering_model_args=cargs)
    with injected_noise(seed=31):
        ref, rconf = osamp.sampling(dl_cpu, m["oracle_fwd"], steps, sched, sched, sched, m["o_t2s"], args, batch_size=S,
                                    confidence_forward=m["oracle_conf"], filtering_data_list=fl_cpu,
                                    filtering_model_args=cargs, crop_fn=osamp.crop_beyond)
    worst = max(rmsd(a["ligand"].pos, b["ligand"].pos) for a, b in zip(out, ref))
    assert worst < 1e-3, worst
    assert torch.allclose(conf.cpu(), rconf, atol=1e-4), (conf.cpu() - rconf).abs().max()
