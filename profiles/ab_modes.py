"""A/B of the accumulate kernels inside ONE process on ONE box (box-to-box and host jitter between separate bench.py runs is
+-5 %, larger than the differences being judged): the bench batch (40 samples, 20 reverse-diffusion steps, score model only),
modes alternated, CUDA-event time per pass.   python profiles/ab_modes.py [modes, default 4,3] [rounds] [reps]"""
import copy
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from confidence_bootstrapping_b200 import tensor_layers as tl  # noqa: E402
from confidence_bootstrapping_b200.configs import score_model_args  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.sampling import _mask_rotate_of, reverse_diffusion  # noqa: E402
from confidence_bootstrapping_b200.utils import get_model  # noqa: E402

modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "4,3").split(",")]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
args = score_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
sched = get_t_schedule("expbeta", bench.INF_STEPS, 1, 1)
dl = bench.build_workload(1500, args, bench.SAMPLES)
mask_rotate = _mask_rotate_of(dl[0])
res = {m: [] for m in modes}
with torch.no_grad():
    for r in range(rounds + 1):
        for m in modes:
            tl.ACCUM_MODE = m
            for k in range(reps):
                b = Batch.from_data_list(copy.deepcopy(dl), device=dev)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                reverse_diffusion(b, model, bench.INF_STEPS, sched, sched, sched, dev, t2s, args, mask_rotate)
                e.record()
                torch.cuda.synchronize()
                if r > 0:
                    res[m].append(s.elapsed_time(e))
for m in modes:
    v = sorted(res[m])
    print(f"mode {m}: median {v[len(v) // 2]:.2f} ms  min {v[0]:.2f}  max {v[-1]:.2f}  ({bench.SAMPLES * 1e3 / v[len(v) // 2]:.1f} poses/s, score model only)  n={len(v)}")
