"""cProfile of ONE end-to-end sampling() call with host buffers (collate to device, 20 steps, crop + confidence, D2H):
where does the host time outside the step loop go?   python profiles/e2e_profile.py [n_lines]"""
import cProfile, copy, os, pstats, sys, time
from functools import partial
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma
from confidence_bootstrapping_b200.sampling import sampling
from confidence_bootstrapping_b200.utils import get_model
dev = torch.device("cuda")
args, cargs = score_model_args(), confidence_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
cmodel = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
sched = get_t_schedule("expbeta", 20, 1, 1)
kw = dict(model=model, inference_steps=20, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s,
          model_args=args, batch_size=bench.SAMPLES, confidence_model=cmodel, filtering_model_args=cargs)


def call(seed):
    dl = bench.build_workload(seed, args, bench.SAMPLES)
    fl = copy.deepcopy(dl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, conf = sampling(data_list=dl, filtering_data_list=fl, **kw)
    t1 = time.perf_counter()
    poses = torch.stack([d["ligand"].pos for d in out]).cpu()
    c = conf.cpu()
    torch.cuda.synchronize()
    return t1 - t0, time.perf_counter() - t0


for s in range(3):
    call(s)
print("host-return / total seconds of a call:", [tuple(round(x, 4) for x in call(10 + s)) for s in range(3)])
pr = cProfile.Profile()
pr.enable()
call(99)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 45)
