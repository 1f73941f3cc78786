"""In-process A/B of end-to-end sampling() calls (host buffers in, poses + confidences out; a NEW complex per call, so the
step-graph cache never hits) with sampling.CAPTURE_FIRST_STEP off / on.   python profiles/ab_e2e.py [calls per arm]"""
import copy
import os
import sys
import time
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from confidence_bootstrapping_b200 import sampling as smp  # noqa: E402
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.utils import get_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
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
    out, conf = smp.sampling(data_list=dl, filtering_data_list=fl, **kw)
    torch.stack([d["ligand"].pos for d in out]).cpu()
    conf.cpu()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for s in range(3):
    call(s)
res = {False: [], True: []}
for r in range(n):
    for flag in (False, True):
        smp.CAPTURE_FIRST_STEP = flag
        res[flag].append(call(100 + r))          # the same complexes for both arms
for flag in (False, True):
    v = sorted(res[flag])
    print(f"CAPTURE_FIRST_STEP={flag}: median {1e3 * v[len(v) // 2]:.1f} ms  min {1e3 * v[0]:.1f}  max {1e3 * v[-1]:.1f}   "
          f"({bench.SAMPLES / v[len(v) // 2]:.1f} poses/s end to end)")
