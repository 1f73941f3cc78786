"""Host profile of the FIRST forward on a fresh batch (per-batch static setup) vs a steady-state forward."""
import cProfile, copy, os, pstats, sys, time
from functools import partial
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from confidence_bootstrapping_b200.configs import score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma
from confidence_bootstrapping_b200.utils import get_model
dev = torch.device("cuda")
args = score_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
dl = bench.build_workload(1, args, bench.SAMPLES)
base = Batch.from_data_list(dl).to(dev)
with torch.no_grad():
    for rep in range(3):
        b = copy.deepcopy(base)
        set_time(b, None, 0.5, 0.5, 0.5, b.num_graphs, False, False, dev)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model(b); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        model(b); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f"rep {rep}: first forward host {1e3*(t1-t0):.1f} / device-done {1e3*(t2-t0):.1f} ms; second host {1e3*(t3-t2):.1f} / {1e3*(t4-t2):.1f} ms")
    b = copy.deepcopy(base)
    set_time(b, None, 0.5, 0.5, 0.5, b.num_graphs, False, False, dev)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable(); model(b); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
