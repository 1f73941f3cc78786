"""Where do the occasional 60-190 ms long steps of the pipelined resident pass come from?  Replays bench.py's resident pass with
host timestamps per phase and the caching allocator's segment counters per step.   python profiles/stall_probe.py [steps]"""
import copy
import os
import sys
import time
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, set_time, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.sampling import FilteringLeg, _mask_rotate_of, reverse_diffusion  # noqa: E402
from confidence_bootstrapping_b200.utils import crop_beyond, get_model  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
args, cargs = score_model_args(), confidence_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
cmodel = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
sched = get_t_schedule("expbeta", 20, 1, 1)
dl = bench.build_workload(1500, args, bench.SAMPLES)
mr = _mask_rotate_of(dl[0])
pairs = [(Batch.from_data_list(copy.deepcopy(dl), device=dev), Batch.from_data_list(copy.deepcopy(dl), device=dev)) for _ in range(K + 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def leg_of(fb):
    def leg(pos):
        t0 = time.perf_counter()
        fb["ligand"].pos = pos
        c = crop_beyond(fb, cargs.crop_beyond, True)
        t1 = time.perf_counter()
        set_time(c, 0, 0, 0, 0, c.num_graphs, True, False, dev)
        out = cmodel(c)[0]
        t2 = time.perf_counter()
        leg.times = (t1 - t0, t2 - t1, int(c["receptor"].num_nodes), int(c["atom"].num_nodes))
        return out
    return leg


with torch.no_grad():
    reverse_diffusion(copy.deepcopy(pairs[-1][0]), model, 20, sched, sched, sched, dev, t2s, args, mr)
    leg_of(copy.deepcopy(pairs[-1][1]))(pairs[-1][0]["ligand"].pos)
    torch.cuda.synchronize()
    fl = FilteringLeg(dev)
    marks, rows, legs = [], [], []
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for b, fb in pairs[:K]:
        st0 = torch.cuda.memory_stats()
        h0 = time.perf_counter()
        flush.fill_(1)
        pos = reverse_diffusion(b, model, 20, sched, sched, sched, dev, t2s, args, mr)
        h1 = time.perf_counter()
        lg = leg_of(fb)
        legs.append(lg)
        fl.submit(lg, pos)
        h2 = time.perf_counter()
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        marks.append(m)
        st1 = torch.cuda.memory_stats()
        rows.append((h1 - h0, h2 - h1, st1["segment.all.allocated"] - st0["segment.all.allocated"],
                     (st1["reserved_bytes.all.current"] - st0["reserved_bytes.all.current"]) / 2 ** 20))
    fl.finish()
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    torch.cuda.synchronize()
ends = [e0] + marks[:-1] + [e1]
print("step  gpu_ms  host_enqueue_ms  host_leg(prev)_ms  new_segments  reserved_delta_MiB   prev leg: crop_ms model_ms residues atoms")
for i, r in enumerate(rows):
    lt = getattr(legs[i - 1], "times", None) if i > 0 else None
    print(f"{i:3d} {ends[i].elapsed_time(ends[i + 1]):8.1f} {1e3 * r[0]:10.1f} {1e3 * r[1]:14.1f} {r[2]:10d} {r[3]:14.1f}     "
          + (f"{1e3 * lt[0]:8.1f} {1e3 * lt[1]:8.1f} {lt[2]:7d} {lt[3]:7d}" if lt else ""))
