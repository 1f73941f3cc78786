"""BASELINE.json configs[3] and configs[4] at full size (the bench line itself is configs[1]; configs[2] rides in bench.py).

  python profiles/run_configs.py --config 4            # 100 complexes x 8 samples, inference_batch_size 32, score + confidence
  python profiles/run_configs.py --config 5            # all-atom SCORE model, 1000-residue receptor, 64-atom ligand, 128 samples
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         profiles/run_configs.py --config 5            # the same on 8 GPUs

config 4 shards the COMPLEX list over the ranks (dist.partition_lpt); config 5 is one complex, so its 128 SAMPLES are sharded
(dist.partition_samples): every (complex, sample) trajectory is independent (utils/sampling.py:89-233).  One JSON line per run,
wall time = max over ranks of the rank's own time, final gather of poses (+ confidences) included.  Inputs are built before the
timed region (featurisation is out of scope)."""
import argparse
import copy
import json
import os
import sys
import time
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from confidence_bootstrapping_b200 import dist as cbdist  # noqa: E402
from confidence_bootstrapping_b200.configs import all_atom_score_model_args, confidence_model_args, score_model_args  # noqa: E402
from confidence_bootstrapping_b200.data import Batch  # noqa: E402
from confidence_bootstrapping_b200.diffusion_utils import get_t_schedule, t_to_sigma  # noqa: E402
from confidence_bootstrapping_b200.sampling import randomize_position, sampling, sampling_many  # noqa: E402
from confidence_bootstrapping_b200.synthetic import make_complex  # noqa: E402
from confidence_bootstrapping_b200.utils import get_model  # noqa: E402


def main():
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--n", type=int, default=0, help="override the number of complexes (config 4) / samples (config 5)")
    opts = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
        dist.init_process_group("nccl", device_id=dev)
    steps = 20
    sched = get_t_schedule("expbeta", steps, 1, 1)
    torch.manual_seed(0)
    if opts.config == 4:
        args = score_model_args()
        t2s = partial(t_to_sigma, args=args)
        model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
        cargs = confidence_model_args()
        cmodel = get_model(cargs, dev, t_to_sigma=t2s, no_parallel=True, confidence_mode=True).eval()
        n_cplx, S, bs = opts.n or 100, 8, 32
        rng = np.random.default_rng(2)
        sizes = [(int(a), int(b)) for a, b in zip(rng.integers(150, 1001, size=n_cplx), rng.integers(10, 61, size=n_cplx))]
        costs = [cbdist.estimate_cost(nl, nr, S) for nr, nl in sizes]
        mine = cbdist.partition_lpt(costs, world)[rank]
        work = []
        for i in mine:
            g = Batch.from_data_list([make_complex(7000 + i, sizes[i][0], sizes[i][1], all_atoms=True)])
            np.random.seed(i)
            torch.manual_seed(i)
            dl = [copy.deepcopy(g) for _ in range(S)]
            randomize_position(dl, False, False, args.tr_sigma_max)
            work.append((i, dl, copy.deepcopy(dl)))
        kw = dict(model=model, inference_steps=steps, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s,
                  model_args=args, batch_size=bs, confidence_model=cmodel, filtering_model_args=cargs)
        n_units, n_poses = n_cplx, n_cplx * S
        what = (f"configs[3]: {n_cplx} synthetic complexes (N_r ~ U[150,1000], N_l ~ U[10,60], seed 2) x {S} samples x {steps} steps, "
                f"inference_batch_size {bs}, score + confidence models, complexes LPT-sharded")
    else:
        args = all_atom_score_model_args()
        t2s = partial(t_to_sigma, args=args)
        model = get_model(args, dev, t_to_sigma=t2s, no_parallel=True).eval()
        S_total, bs = opts.n or 128, 32
        mine_s = cbdist.partition_samples(S_total, world)[rank]
        g = Batch.from_data_list([make_complex(9000, 1000, 64, all_atoms=True)])
        np.random.seed(5)
        torch.manual_seed(5)
        dl_all = [copy.deepcopy(g) for _ in range(S_total)]
        randomize_position(dl_all, False, False, args.tr_sigma_max)
        work = [(0, [dl_all[k] for k in mine_s], None)]
        mine = [rank]
        kw = dict(model=model, inference_steps=steps, tr_schedule=sched, rot_schedule=sched, tor_schedule=sched, device=dev, t_to_sigma=t2s,
                  model_args=args, batch_size=bs)
        n_units, n_poses = world, S_total
        what = (f"configs[4]: all-atom score model (ns {args.ns}, nv {args.nv}, {args.num_conv_layers} conv layers, lmax {args.sh_lmax}), synthetic "
                f"1000-residue receptor ({g['atom'].num_nodes} atoms), 64-atom ligand, {S_total} samples x {steps} steps, batches of {bs}, "
                f"samples sharded over the ranks")

    def run():
        poses, confs = [], []
        todo = [(dl, fl) for i, dl, fl in work if dl]
        # one pipelined call for the rank's complexes (sampling.sampling_many == the loop of sampling() calls, same results)
        for out, conf in sampling_many(todo, **kw):
            poses.append(torch.stack([d["ligand"].pos for d in out]))
            confs.append(conf)
        return poses, confs

    # warm-up on a tiny slice (module caches, cuBLAS handles, first-touch of the kernels), then the timed pass
    if work and work[0][1]:
        i0, dl0, fl0 = work[0]
        n0 = min(2, len(dl0))
        sampling(data_list=copy.deepcopy(dl0[:n0]), filtering_data_list=copy.deepcopy(fl0[:n0]) if fl0 is not None else None,
                 **{**kw, "inference_steps": 2, "tr_schedule": sched[:2], "rot_schedule": sched[:2], "tor_schedule": sched[:2], "batch_size": n0})
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    poses, confs = run()
    torch.cuda.synchronize()
    busy = time.perf_counter() - t0
    ids = mine if opts.config == 4 else ([rank] if poses else [])
    gp, gc = cbdist.gather_results(ids, poses, confs, n_units, device=dev)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    t = torch.tensor([busy, total, torch.cuda.max_memory_allocated() / 2 ** 30], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        busy_all = [float(x[0]) for x in allt]
        wall = max(float(x[1]) for x in allt)
        finite = all(p is None or bool(torch.isfinite(p).all()) for p in gp)
        print(json.dumps({"config": what, "n_gpus": world, "poses": n_poses, "wall_s": round(wall, 3), "poses_per_s": round(n_poses / wall, 2),
                          "rank_busy_s": [round(b, 3) for b in busy_all],
                          "imbalance_max_over_mean": round(max(busy_all) / (sum(busy_all) / len(busy_all)), 3),
                          "peak_mem_GiB_max": round(max(float(x[2]) for x in allt), 2), "results_finite": finite,
                          "gathered_units": sum(p is not None for p in gp)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
