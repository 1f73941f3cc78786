"""A small slice of SURVEY config 3 (complexes of different sizes x 40 samples through sample_complexes / sampling())
as a robustness + throughput check at sizes the bench does not cover (1000-residue receptor => K3 runs in node chunks)."""
import copy, os, sys, time
from functools import partial
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from confidence_bootstrapping_b200 import dist as cbdist
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma
from confidence_bootstrapping_b200.sampling import randomize_position, sampling
from confidence_bootstrapping_b200.synthetic import make_complex
from confidence_bootstrapping_b200.utils import get_model
dev = torch.device("cuda")
args, cargs = score_model_args(), confidence_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
conf = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
sizes = [(150, 10), (1000, 60), (420, 33), (777, 25)]
complexes = [Batch.from_data_list([make_complex(100 + i, nr, nl, all_atoms=True)]) for i, (nr, nl) in enumerate(sizes)]
sched = get_t_schedule("expbeta", 20, 1, 1)
S = 40
def sample_fn(g, n):
    np.random.seed(0); torch.manual_seed(0)
    dl = [copy.deepcopy(g) for _ in range(n)]
    randomize_position(dl, args.no_torsion, False, args.tr_sigma_max)
    fl = copy.deepcopy(dl)
    out, c = sampling(data_list=dl, model=model, inference_steps=20, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev,
                      t_to_sigma=t2s, model_args=args, batch_size=n, confidence_model=conf, filtering_data_list=fl, filtering_model_args=cargs)
    return torch.stack([d["ligand"].pos for d in out]), c
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    poses, confs = cbdist.sample_complexes(complexes, S, sample_fn, device=dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ok = all(torch.isfinite(p).all() and torch.isfinite(c).all() for p, c in zip(poses, confs))
    print(f"rep {rep}: {len(complexes)} complexes x {S} samples in {dt:.2f} s = {len(complexes) * S / dt:.1f} poses/s, finite={bool(ok)}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
