"""Per-phase cycle breakdown of the H role (MMA issuer thread and hidden unit 127) of tp_accumulate_ws_kernel.
Needs CB200_EXTRA_NVCC_FLAGS=-DCB_PHASE_TIMING python confidence_bootstrapping_b200/build.py --force  and CB200_ACCUM_MODE=4."""
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
batch = Batch.from_data_list(bench.build_workload(1, args, bench.SAMPLES), device=dev)
lib = _lib.lib()
buf = (ctypes.c_ulonglong * 24)()
with torch.no_grad():
    for i in range(2):
        set_time(batch, None, 0.5, 0.5, 0.5, batch.num_graphs, False, False, dev)
        if i == 1:
            lib.cb_debug_phases(buf, 1)
        model(batch)
lib.cb_debug_phases(buf, 0)
names = ["wait raw_full", "item setup + E split + fence", "named barrier 1", "hidden MMA issue", "wait hid_bar", "tcgen05.ld", "wait h_free",
         "H~ tile + fences + arrive", "named barrier 2", "issuer: wait f_full", "issuer: wait acc_empty", "issuer: main MMA issue + commits"]
for w, label in ((0, "hidden unit 0 = the MMA issuer"), (1, "hidden unit 127")):
    v = [buf[w * 12 + k] for k in range(12)]
    tot = sum(v)
    print(label, "total Mcycles over all CTAs and K3 launches of one forward:", tot / 1e6)
    for k in range(12):
        print(f"   {names[k]:36s} {100.0 * v[k] / tot:5.1f} %")
