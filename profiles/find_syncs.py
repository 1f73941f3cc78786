"""Lists the host<->device synchronisation points of one steady-state reverse-diffusion run
(torch.cuda.set_sync_debug_mode('warn')), grouped by the cb200 source line that triggers them."""
import collections, os, sys, traceback, warnings
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
    reverse_diffusion(batch, model, 3, sched[:3], sched[:3], sched[:3], dev, t2s, args, mr)   # warm caches
    torch.cuda.synchronize()
    hits = collections.Counter()
    def showwarning(message, category, filename, lineno, file=None, line=None):
        if "synchroniz" not in str(message):
            return
        for fr in reversed(traceback.extract_stack()):
            if "confidence_bootstrapping_b200" in fr.filename:
                hits[f"{os.path.basename(fr.filename)}:{fr.lineno} {fr.line}"] += 1
                return
        hits["<outside>"] += 1
    warnings.showwarning = showwarning
    warnings.simplefilter("always")
    torch.cuda.set_sync_debug_mode("warn")
    reverse_diffusion(batch, model, 4, sched[:4], sched[:4], sched[:4], dev, t2s, args, mr)
    torch.cuda.set_sync_debug_mode("default")
for k, v in hits.most_common():
    print(v, k)
print("total syncs in 4 steps:", sum(hits.values()))
