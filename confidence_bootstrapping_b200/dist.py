"""Multi-GPU sampling: shard the complex list across ranks, gather poses + confidences once at the end.

The reference has no distributed backend: its only multi-GPU construct is PyG
DataParallel for training.  The sampling path is embarrassingly parallel over (complex x sample)
(utils/sampling.py:89-233 never mixes graphs), so the B200 design is one process per GPU
(torch.distributed, NCCL over NVLink), the complex list partitioned longest-processing-time-first,
all samples of a complex kept on one rank (they share the ligand topology, the cached receptor
embedding and the K4 launch), NO collective inside the sampling loop, and a single variable-length
all-gather of final ligand coordinates and confidences.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def estimate_cost(n_lig: int, n_rec: int, n_samples: int, rec_neighbors: int = 24) -> float:
    """Edge-layers per step ~ samples * (cross edges + receptor edges)."""
    return float(n_samples) * (n_lig * n_rec + rec_neighbors * n_rec)


def partition_lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first: heaviest complex to the least loaded rank.  Deterministic
    (ties broken by index), identical on every rank, so no communication is needed to agree on it."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += costs[i]
    for p in parts:
        p.sort()
    return parts


def partition_samples(n_samples: int, world_size: int) -> List[List[int]]:
    """Contiguous, near-equal blocks of the sample indices of ONE complex (BASELINE config 5: a single 1000-residue
    complex x 128 samples on 8 GPUs).  Every (complex, sample) trajectory is independent (utils/sampling.py:89-233), so
    samples shard as freely as complexes do; each rank then runs `sampling()` on its block."""
    base, extra = divmod(n_samples, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_results(local_ids: Sequence[int], poses: Sequence[torch.Tensor], confidences: Sequence[Optional[torch.Tensor]],
                   n_total: int, device=None) -> Tuple[List[Optional[torch.Tensor]], List[Optional[torch.Tensor]]]:
    """All-gather per-complex results.

    local_ids[k] is the global index of the k-th local complex, poses[k] its [S, N_lig, 3] float32 final
    coordinates, confidences[k] its [S] (or [S, k]) confidences (or None).  Every rank returns the full lists
    (index = global complex id).  Payload is KB-MB: one padded all_gather of a flat float buffer plus
    one of the int64 layout table; NCCL on GPUs, gloo for the CPU tests."""
    rank, world = _world()
    if world == 1:
        out_p: List[Optional[torch.Tensor]] = [None] * n_total
        out_c: List[Optional[torch.Tensor]] = [None] * n_total
        for i, p, c in zip(local_ids, poses, confidences):
            out_p[i], out_c[i] = p, c
        return out_p, out_c
    if device is None:
        if len(poses):
            device = poses[0].device
        elif dist.get_backend() == "nccl":     # a rank without complexes still joins the collectives with CUDA tensors
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cpu")
    # layout rows: (global id, S, N, confidence numel (0 = none), confidence trailing width)
    def conf_meta(c):
        if c is None:
            return 0, 0
        return int(c.numel()), (int(c.numel() // max(c.shape[0], 1)) if c.dim() > 1 else 0)
    meta = torch.tensor([[i, p.shape[0], p.shape[1], *conf_meta(c)] for i, p, c in zip(local_ids, poses, confidences)],
                        dtype=torch.int64, device=device).reshape(-1, 5)
    flat = [p.reshape(-1).float() for p in poses] + [c.reshape(-1).float() for c in confidences if c is not None]
    payload = torch.cat(flat) if flat else torch.zeros(0, device=device)
    sizes = torch.tensor([meta.shape[0], payload.numel()], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_meta = max(int(s[0]) for s in all_sizes)
    max_pay = max(int(s[1]) for s in all_sizes)
    meta_pad = torch.zeros((max(max_meta, 1), 5), dtype=torch.int64, device=device)
    meta_pad[: meta.shape[0]] = meta
    pay_pad = torch.zeros(max(max_pay, 1), dtype=torch.float32, device=device)
    pay_pad[: payload.numel()] = payload
    metas = [torch.zeros_like(meta_pad) for _ in range(world)]
    pays = [torch.zeros_like(pay_pad) for _ in range(world)]
    dist.all_gather(metas, meta_pad)
    dist.all_gather(pays, pay_pad)
    out_p = [None] * n_total
    out_c = [None] * n_total
    for r in range(world):
        n_meta = int(all_sizes[r][0])
        rows = metas[r][:n_meta].tolist()
        off = 0
        for (i, S, N, _, _) in rows:
            out_p[i] = pays[r][off: off + S * N * 3].reshape(S, N, 3)
            off += S * N * 3
        for (i, S, N, n_conf, k_conf) in rows:
            if n_conf:      # [S] for a scalar head, [S, k] when rmsd_classification_cutoff is a list (utils/utils.py:262)
                c = pays[r][off: off + n_conf]
                out_c[i] = c.reshape(-1, k_conf) if k_conf else c
                off += n_conf
    return out_p, out_c


def sample_complexes(complexes: Sequence, n_samples: int, sample_fn=None, costs: Optional[Sequence[float]] = None, device=None,
                     sample_many_fn=None):
    """Shard `complexes` over the ranks, run `sample_fn(complex, n_samples) -> (poses [S,N,3], conf [S] or None)`
    on the local shard, and all-gather.  Returns (poses_per_complex, conf_per_complex) on every rank.

    `sample_many_fn(list_of_complexes, n_samples) -> [(poses, conf), ...]` (one entry per complex, in order) replaces the
    per-complex calls by ONE call for the rank's shard -- the hook for `sampling.sampling_many`, which overlaps the filtering leg
    of a complex with the collate / capture / steps of the next one."""
    if (sample_fn is None) == (sample_many_fn is None):
        raise ValueError("sample_complexes: pass exactly one of sample_fn / sample_many_fn")
    rank, world = _world()
    if costs is None:
        costs = [estimate_cost(int(c["ligand"].num_nodes), int(c["receptor"].num_nodes), n_samples) for c in complexes]
    mine = partition_lpt(costs, world)[rank]
    poses, confs = [], []
    if sample_many_fn is not None:
        res = list(sample_many_fn([complexes[i] for i in mine], n_samples)) if mine else []
        if len(res) != len(mine):
            raise RuntimeError(f"sample_many_fn returned {len(res)} results for {len(mine)} complexes")
        for p, c in res:
            poses.append(p)
            confs.append(c)
    else:
        for i in mine:
            p, c = sample_fn(complexes[i], n_samples)
            poses.append(p)
            confs.append(c)
    return gather_results(mine, poses, confs, len(complexes), device=device)
