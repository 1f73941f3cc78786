"""cProfile of sampling_many over mixed-size complexes x 8 samples (BASELINE configs[3] slice): where the HOST time per complex
goes once the GPU is no longer the limit.   python profiles/host_profile_many.py [n_complexes] [lines]"""
import cProfile
import copy
import os
import pstats
import sys
import time
from functools import partial

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.sampling import randomize_position, sampling_many  # noqa: E402
from confidence_bootstrapping_b200.synthetic import make_complex  # noqa: E402
from confidence_bootstrapping_b200.utils import get_model  # noqa: E402

n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda")
args, cargs = score_model_args(), confidence_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
cmodel = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
sched = get_t_schedule("expbeta", 20, 1, 1)
rng = np.random.default_rng(2)
sizes = [(int(a), int(b)) for a, b in zip(rng.integers(150, 1001, size=n_c), rng.integers(10, 61, size=n_c))]
kw = dict(model=model, inference_steps=20, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s,
          model_args=args, batch_size=32, confidence_model=cmodel, filtering_model_args=cargs)


def work():
    out = []
    for i, (nr, nl) in enumerate(sizes):
        g = Batch.from_data_list([make_complex(7000 + i, nr, nl, all_atoms=True)])
        np.random.seed(i)
        torch.manual_seed(i)
        dl = [copy.deepcopy(g) for _ in range(8)]
        randomize_position(dl, False, False, args.tr_sigma_max)
        out.append((dl, copy.deepcopy(dl)))
    return out


sampling_many(work()[:3], **kw)
torch.cuda.synchronize()
w = work()
t0 = time.perf_counter()
sampling_many(w, **kw)
torch.cuda.synchronize()
print(f"{n_c} complexes x 8 samples: {1e3 * (time.perf_counter() - t0) / n_c:.1f} ms per complex")
w = work()
pr = cProfile.Profile()
pr.enable()
sampling_many(w, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(int(sys.argv[2]) if len(sys.argv) > 2 else 40)
