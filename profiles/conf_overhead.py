"""Time of the confidence-model leg of one bench step (crop_beyond + all-atom forward on a fresh batch)."""
import copy, os, sys, time
from functools import partial
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from confidence_bootstrapping_b200.configs import confidence_model_args, score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import set_time, t_to_sigma
from confidence_bootstrapping_b200.utils import crop_beyond, get_model
dev = torch.device("cuda")
args = score_model_args()
t2s = partial(t_to_sigma, args=args)
cargs = confidence_model_args()
torch.manual_seed(0)
conf = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
dl = bench.build_workload(1, args, bench.SAMPLES)
base = Batch.from_data_list(dl).to(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
with torch.no_grad():
    for rep in range(4):
        fb = copy.deepcopy(base)
        t0 = T()
        fbc = crop_beyond(fb, cargs.crop_beyond, True)
        t1 = T()
        set_time(fbc, 0, 0, 0, 0, fbc.num_graphs, True, False, dev)
        t2 = time.perf_counter()
        out = conf(fbc)[0]
        t3 = time.perf_counter()
        t4 = T()
        print(f"rep {rep}: crop_beyond {1e3*(t1-t0):.1f} ms | forward host {1e3*(t3-t2):.1f} ms, device done {1e3*(t4-t2):.1f} ms")
    if len(sys.argv) > 1:
        import cProfile, pstats
        fb = crop_beyond(copy.deepcopy(base), cargs.crop_beyond, True)
        set_time(fb, 0, 0, 0, 0, fb.num_graphs, True, False, dev)
        pr = cProfile.Profile(); pr.enable(); conf(fb); torch.cuda.synchronize(); pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
