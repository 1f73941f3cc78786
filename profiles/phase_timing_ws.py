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
buf = (ctypes.c_ulonglong * 96)()
with torch.no_grad():
    for i in range(2):
        set_time(batch, None, 0.5, 0.5, 0.5, batch.num_graphs, False, False, dev)
        if i == 1:
            lib.cb_debug_phases(buf, 1)
        model(batch)
lib.cb_debug_phases(buf, 0)
ROLES = [
    ("H warp 8 (hidden units 0-31)", ["wait raw_full(c+1)", "E phase (item setup + E split + fence)", "-", "-", "wait hid_done", "tcgen05.ld", "wait h_free", "H~ tile + fences + arrive"]),
    ("H warp 9 (hidden units 32-63)", ["wait raw_full(c+1)", "E phase (item setup + E split + fence)", "-", "-", "wait hid_done", "tcgen05.ld", "wait h_free", "H~ tile + fences + arrive"]),
    ("F thread 0", ["wait raw_full", "wait f_free", "f-row + fence + arrive", "fsum store + loop"]),
    ("S scheduler", ["wait raw_empty", "iterator + descriptor"]),
    ("G warp 0", ["wait desc_full", "issue cp.async", "wait_group + publish"]),
    ("I issuer", ["polling, nothing ready", "hidden MMA issue + commit", "wait acc_empty", "main MMA issue + commits", "(chunks issued, count)"]),
    ("E warp 16", ["wait acc_full", "TMEM -> workspace", "(items, count)"]),
    ("F thread 224", ["wait raw_full", "wait f_free", "f-row + fence + arrive", "fsum store + loop"]),
]
chunks = buf[5 * 12 + 4]
items = buf[6 * 12 + 2]
print(f"chunks issued (all CTAs, all K3 launches of one forward): {chunks}   items: {items}")
for w, (label, names) in enumerate(ROLES):
    v = [buf[w * 12 + k] for k in range(12)]
    cyc = [x for k, x in enumerate(v) if k < len(names) and "count" not in names[k] and names[k] != "-"]
    tot = sum(cyc)
    print(f"{label}: {tot / 1e6:.1f} Mcycles total, {tot / max(chunks, 1):.0f} cycles per chunk")
    for k, nm in enumerate(names):
        if nm == "-" or "count" in nm:
            continue
        print(f"   {nm:42s} {100.0 * v[k] / max(tot, 1):5.1f} %   {v[k] / max(chunks, 1):7.0f} clk/chunk")
