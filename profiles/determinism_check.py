"""Runs the same 20-step sampling (identical inputs and injected noise) repeatedly on the GPU and on the CPU oracle
and reports run-to-run differences: the CUDA path must be bit-reproducible."""
import copy, os, sys
from functools import partial
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import injected_noise, rmsd
from test_gpu_model import _build
from oracle import sampler as osamp
from confidence_bootstrapping_b200.configs import score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule
from confidence_bootstrapping_b200.sampling import randomize_position, sampling
from confidence_bootstrapping_b200.synthetic import make_complex
args = score_model_args()
model, t2s, oracle_fwd = _build(args, seed=3)
g = Batch.from_data_list([make_complex(77, 80, 16, all_atoms=False)])
np.random.seed(0); torch.manual_seed(0)
dl0 = [copy.deepcopy(g) for _ in range(4)]
randomize_position(dl0, False, False, args.tr_sigma_max)
sched = get_t_schedule("expbeta", 20, 1, 1)
n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 12
outs = []
for i in range(n_gpu):
    dl = copy.deepcopy(dl0)
    with injected_noise(seed=9):
        out, _ = sampling(data_list=dl, model=model, inference_steps=20, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched,
                          device=torch.device("cuda"), t_to_sigma=t2s, model_args=args, batch_size=4)
    outs.append(torch.stack([d["ligand"].pos for d in out]).cpu())
    print("gpu run", i, "max |diff| to run 0:", float((outs[i] - outs[0]).abs().max()))
refs = []
for i in range(2):
    dl = copy.deepcopy(dl0)
    with injected_noise(seed=9):
        ref, _ = osamp.sampling(dl, oracle_fwd, 20, sched, sched, sched, partial(osamp.t_to_sigma, args=args), args, batch_size=4)
    refs.append(torch.stack([d["ligand"].pos for d in ref]))
    print("oracle run", i, "max |diff| to oracle run 0:", float((refs[i] - refs[0]).abs().max()),
          " worst RMSD gpu0 vs oracle:", max(rmsd(a, b) for a, b in zip(outs[0], refs[i])))
