"""cProfile of the host side of steady-state reverse-diffusion steps (which Python / dispatch costs dominate?)."""
import cProfile, os, pstats, sys
from functools import partial
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from confidence_bootstrapping_b200.configs import score_model_args
from confidence_bootstrapping_b200.data import Batch
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma
from confidence_bootstrapping_b200.sampling import _mask_rotate_of, reverse_diffusion
from confidence_bootstrapping_b200.utils import get_model
dev = torch.device("cuda")
args = score_model_args()
t2s = partial(t_to_sigma, args=args)
torch.manual_seed(0)
model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
dl = bench.build_workload(1, args, bench.SAMPLES)
mr = _mask_rotate_of(dl[0])
sched = get_t_schedule("expbeta", 20, 1, 1)
batch = Batch.from_data_list(dl).to(dev)
with torch.no_grad():
    reverse_diffusion(batch, model, 20, sched, sched, sched, dev, t2s, args, mr)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    reverse_diffusion(batch, model, 20, sched, sched, sched, dev, t2s, args, mr)
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 30)
