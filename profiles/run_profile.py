"""Small driver for ncu: a few score-model forwards on the bench workload (40 samples of the
400-residue / 40-atom complex).  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/run_profile.py 2
  ncu --set full --clock-control none --import-source on -k regex:tp_conv -s 13 -c 3 -o gpurun_out/tp_conv python profiles/run_profile.py 2
"""
import copy
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from confidence_bootstrapping_b200.configs import score_model_args  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.utils import get_model  # noqa: E402

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
t = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
dev = torch.device("cuda")
args = score_model_args()
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=partial(t_to_sigma, args=args), no_parallel=True).eval()
dl = bench.build_workload(1, args, bench.SAMPLES)
batch = Batch.from_data_list(dl, device=dev)   # device collate (flags the replicated receptor) like sampling()
with torch.no_grad():
    for i in range(n_fwd):
        set_time(batch, None, t, t, t, batch.num_graphs, False, False, dev)
        torch.cuda.nvtx.range_push(f"forward{i}")
        out = model(batch)
        torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("ok", [o.shape for o in out[:3]])
