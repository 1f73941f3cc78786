"""Generate tests/golden/*.pt by running the REAL first-party reference (/root/reference) on CPU under
oracle/shims.py.  Runs only in the build container (the reference tree does not travel to the GPU box).

    cd /tmp/refrun && python /root/repo/oracle/make_golden.py

(the working directory must hold the reference's .so3_*.npy / .p.npy / .score.npy caches, or the
import of utils/so3.py + utils/torus.py recomputes them for ~20 minutes; oracle/gen_tables.py creates them).
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import copy
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import shims  # noqa: E402

np.random.seed(0)  # utils/torus.py draws its Monte-Carlo table with the global numpy RNG at import
shims.install()

import utils.torus as rtorus  # noqa: E402  (first import -> seeded table)
import utils.so3 as rso3  # noqa: E402
from models.tensor_layers import FasterTensorProduct  # noqa: E402
from utils import geometry as rgeo, torsion as rtor  # noqa: E402
from utils.diffusion_utils import get_t_schedule, modify_conformer_batch, set_time  # noqa: E402
from utils.diffusion_utils import t_to_sigma as t_to_sigma_compl  # noqa: E402
from utils.sampling import randomize_position, sampling  # noqa: E402
from utils.utils import crop_beyond, get_model  # noqa: E402

from confidence_bootstrapping_b200 import torus as ptorus  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.synthetic import make_complex  # noqa: E402
from helpers import GOLDEN, injected_noise, pack_graph, randomize_norm_stats, small_score_args  # noqa: E402

os.makedirs(GOLDEN, exist_ok=True)
dev = torch.device("cpu")


def save(name, obj):
    path = os.path.join(GOLDEN, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# 1. tables -------------------------------------------------------------------------------------------
assert np.array_equal(rtorus.score_norm_, ptorus.score_norm_), "shipped torus table != reference run with seed 0"
eps = torch.tensor([0.0005, 0.06, 0.0601, 0.2, 0.431, 1.0, 3.1, 4.0, 7.0], dtype=torch.float32)
sig = np.asarray([0.0314, 0.05, 0.1, 0.5, 1.0, 2.0, 3.14, 0.001, 9.0], dtype=np.float32)
save("tables.pt", {"so3_eps": eps, "so3_score_norm": rso3.score_norm(eps),
                   "torus_sigma": torch.from_numpy(sig), "torus_score_norm": torch.from_numpy(rtorus.score_norm(sig))})

# 2. FasterTensorProduct (the reference's own closed form; anchors the e3nn FCTP semantics at lmax=1) ---
torch.manual_seed(10)
cases = []
seq = ["32x0e", "32x0e + 6x1o", "32x0e + 6x1o + 6x1e", "32x0e + 6x1o + 6x1e + 6x0o"]
for i_in, i_out in ((0, 1), (1, 2), (2, 3), (3, 3)):
    tp = FasterTensorProduct(seq[i_in], "1x0e+1x1o", seq[i_out])
    x = torch.randn(5, tp.in_irreps.dim)
    sh = torch.randn(5, 4)
    w = torch.randn(5, tp.weight_numel)
    cases.append({"in": seq[i_in], "out": seq[i_out], "x": x, "sh": sh, "w": w, "y": tp(x, sh, w),
                  "weight_numel": tp.weight_numel})
save("faster_tp.pt", cases)

# 3. geometry / torsion / pose update -------------------------------------------------------------------
torch.manual_seed(11)
aa = torch.randn(6, 3)
aa[0] *= 1e-8
aa[1] *= 3.0
A, Bm = torch.randn(3, 10, 3), torch.randn(3, 10, 3)
Bm[2] = A[2] * torch.tensor([1.0, 1.0, -1.0])  # forces the reflection branch
R, t = rgeo.rigid_transform_Kabsch_3D_torch_batch(A, Bm)
g = make_complex(5, 40, 16, all_atoms=False, lm_dim=0)
b3 = Batch.from_data_list([copy.deepcopy(g) for _ in range(3)])
mask_rotate = torch.from_numpy(g["ligand"].mask_rotate)
n_tor = int(g["ligand"].edge_mask.sum())
tr_u, rot_u, tor_u = torch.randn(3, 3), torch.randn(3, 3) * 0.5, torch.randn(3 * n_tor)
pos0 = b3["ligand"].pos + torch.randn_like(b3["ligand"].pos) * 0.05
twisted = rtor.modify_conformer_torsion_angles_batch(pos0.reshape(3, -1, 3), g["ligand", "ligand"].edge_index.T[g["ligand"].edge_mask],
                                                     mask_rotate, tor_u.reshape(3, -1))
new_pos = modify_conformer_batch(pos0, b3, tr_u, rot_u, tor_u, mask_rotate)
rigid_only = modify_conformer_batch(pos0, b3, tr_u, rot_u, None, mask_rotate)
save("geometry.pt", {"axis_angle": aa, "matrix": rgeo.axis_angle_to_matrix(aa), "kabsch_A": A, "kabsch_B": Bm,
                     "kabsch_R": R, "kabsch_t": t, "graph": pack_graph(g), "pos0": pos0, "tr": tr_u, "rot": rot_u,
                     "tor": tor_u, "twisted": twisted, "new_pos": new_pos, "rigid_only": rigid_only})

# 4. small CG score model: forward + sampling ---------------------------------------------------------------
args = small_score_args()
t_to_sigma = partial(t_to_sigma_compl, args=args)
torch.manual_seed(12)
model = get_model(args, dev, t_to_sigma=t_to_sigma, no_parallel=True)
randomize_norm_stats(model, seed=3)
model.eval()
graphs = [make_complex(20, 48, 9, all_atoms=False, lm_dim=0), make_complex(21, 64, 17, all_atoms=False, lm_dim=0)]
batch = Batch.from_data_list(copy.deepcopy(graphs))
set_time(batch, None, 0.6, 0.6, 0.6, 2, False, False, dev)
with torch.no_grad():
    tr, rot, tor, _ = model(batch)
fwd = {"args": vars(args), "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
       "graphs": [pack_graph(x) for x in graphs], "t": 0.6, "tr": tr, "rot": rot, "tor": tor}

g1 = Batch.from_data_list([copy.deepcopy(graphs[1])])
np.random.seed(2)
torch.manual_seed(2)
data_list = [copy.deepcopy(g1) for _ in range(4)]
randomize_position(data_list, args.no_torsion, False, args.tr_sigma_max)
start = torch.stack([d["ligand"].pos.clone() for d in data_list])
sched = get_t_schedule("expbeta", 4, 1, 1)
with injected_noise(seed=5):
    out, _ = sampling(data_list=data_list, model=model, inference_steps=4, tr_schedule=sched, rot_schedule=sched,
                      tor_schedule=sched, device=dev, t_to_sigma=t_to_sigma, model_args=args, batch_size=4)
fwd.update({"sample_start": start, "sample_steps": 4, "sample_noise_seed": 5,
            "sample_final": torch.stack([d["ligand"].pos for d in out])})
save("score_small.pt", fwd)

# 5. crop_beyond (utils/utils.py:395-420) on an all-atom complex -------------------------------------------
ga = make_complex(30, 70, 12, all_atoms=True, lm_dim=0)
gc = copy.deepcopy(ga)
crop_beyond(gc, 12.0, True)
save("crop.pt", {"graph": pack_graph(ga), "cutoff": 12.0, "cropped": pack_graph(gc)})
# 6. small all-atom confidence model: forward, and sampling + crop_beyond + confidence scoring -------------
from confidence_bootstrapping_b200.configs import confidence_model_args  # noqa: E402
cargs = confidence_model_args(ns=8, nv=2, num_conv_layers=3, esm_embeddings_path=None, crop_beyond=12.0)
torch.manual_seed(13)
cmodel = get_model(cargs, dev, t_to_sigma=None, no_parallel=True, confidence_mode=True)
randomize_norm_stats(cmodel, seed=4)
cmodel.eval()
agraphs = [make_complex(40, 50, 10, all_atoms=True, lm_dim=0), make_complex(41, 44, 15, all_atoms=True, lm_dim=0)]
cb = Batch.from_data_list(copy.deepcopy(agraphs))
set_time(cb, 0, 0, 0, 0, 2, True, False, dev)
with torch.no_grad():
    conf, atom_conf = cmodel(cb)
g1 = Batch.from_data_list([copy.deepcopy(agraphs[1])])
np.random.seed(3)
torch.manual_seed(3)
data_list = [copy.deepcopy(g1) for _ in range(4)]
randomize_position(data_list, args.no_torsion, False, 3.0)     # small spread so that cropping keeps residues
start = torch.stack([d["ligand"].pos.clone() for d in data_list])
filt = copy.deepcopy(data_list)
with injected_noise(seed=6):
    out, sconf = sampling(data_list=data_list, model=model, inference_steps=3, tr_schedule=sched[:3] * 0.2, rot_schedule=sched[:3] * 0.2,
                          tor_schedule=sched[:3] * 0.2, device=dev, t_to_sigma=t_to_sigma, model_args=args, batch_size=2,
                          confidence_model=cmodel, filtering_data_list=filt, filtering_model_args=cargs)
save("confidence_small.pt", {"args": vars(cargs), "state_dict": {k: v.clone() for k, v in cmodel.state_dict().items()},
                             "graphs": [pack_graph(x) for x in agraphs], "confidence": conf, "atom_confidence": atom_conf,
                             # the score model of this run is the one stored in score_small.pt
                             "sample_start": start, "sample_sched": torch.tensor(sched[:3] * 0.2), "sample_noise_seed": 6,
                             "sample_final": torch.stack([d["ligand"].pos for d in out]), "sample_confidence": sconf})
print("done")
