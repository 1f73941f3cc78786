"""Per-phase cycle breakdown of tp_accumulate_tc_kernel.  Needs a library built with
CB200_EXTRA_NVCC_FLAGS=-DCB_PHASE_TIMING python confidence_bootstrapping_b200/build.py --force"""
import ctypes, os, sys
from functools import partial
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from confidence_bootstrapping_b200 import _lib
from confidence_bootstrapping_b200.configs import score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma
from confidence_bootstrapping_b200.utils import get_model
dev = torch.device("cuda")
args = score_model_args()
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=partial(t_to_sigma, args=args), no_parallel=True).eval()
batch = Batch.from_data_list(bench.build_workload(1, args, bench.SAMPLES)).to(dev)
lib = _lib.lib()
buf = (ctypes.c_ulonglong * 24)()
with torch.no_grad():
    for i in range(2):
        set_time(batch, None, 0.5, 0.5, 0.5, batch.num_graphs, False, False, dev)
        if i == 1:
            lib.cb_debug_phases(buf, 1)
        model(batch)
lib.cb_debug_phases(buf, 0)
names = ["iterator+item setup", "top barrier", "wait MMA2(c-1)", "E split+sync+MMA1 issue", "F tiles", "wait MMA1", "H epilogue", "MMA2 issue",
         "item epilogue", "(loop top)", "cp.async wait", "pre-MMA2 barrier"]
for w, label in ((0, "thread 0 (warp 0, issues the MMAs)"), (1, "thread 224 (warp 7)")):
    v = [buf[w * 12 + k] for k in range(12)]
    tot = sum(v)
    print(label, "total Mcycles", tot / 1e6)
    for k in (9, 0, 10, 1, 2, 3, 4, 5, 6, 11, 7, 8):
        print(f"   {names[k]:28s} {100.0 * v[k] / tot:5.1f} %")
