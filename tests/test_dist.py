"""Host-side logic of the multi-GPU path on CPU: LPT partition, and the variable-length gather with a
world_size-2 gloo group (two processes on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from confidence_bootstrapping_b200 import dist as cbdist


def test_partition_lpt_is_deterministic_and_balanced():
    costs = [5.0, 1.0, 9.0, 3.0, 3.0, 7.0, 2.0]
    parts = cbdist.partition_lpt(costs, 3)
    assert sorted(i for p in parts for i in p) == list(range(7))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert parts == cbdist.partition_lpt(costs, 3)
    assert cbdist.partition_lpt(costs, 1) == [list(range(7))]
    assert cbdist.partition_lpt([], 2) == [[], []]
    assert cbdist.partition_lpt([4.0], 4) == [[0], [], [], []]   # more ranks than complexes: empty shards are fine


def test_partition_samples_covers_every_sample_once():
    for n, w in ((128, 8), (10, 4), (3, 8), (0, 2)):
        parts = cbdist.partition_samples(n, w)
        assert len(parts) == w and sum(parts, []) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _fake_result(i, n_samples):
    n_atoms = 5 + 3 * i
    g = torch.Generator().manual_seed(100 + i)
    pose = torch.randn(n_samples, n_atoms, 3, generator=g)
    if i % 4 == 0:
        return pose, torch.randn(n_samples, generator=g)           # scalar confidence head: [S]
    if i % 4 == 2:
        return pose, torch.randn(n_samples, 3, generator=g)        # multi-class head (rmsd_classification_cutoff list): [S, k]
    return pose, None


def _worker(rank, world, port, n_complexes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        complexes = [{"ligand": type("S", (), {"num_nodes": 5 + 3 * i})(), "receptor": type("S", (), {"num_nodes": 50 + 7 * i})()}
                     for i in range(n_complexes)]
        seen = []

        def sample_fn(c, n):
            i = (c["ligand"].num_nodes - 5) // 3
            seen.append(i)
            return _fake_result(i, n)

        # device=None: a rank that owns no complex (n_complexes=1) must still pick a device the backend accepts
        poses, confs = cbdist.sample_complexes(complexes, 4, sample_fn, device=None)
        # the same through the one-call-per-shard hook (what sampling.sampling_many plugs into)
        poses2, confs2 = cbdist.sample_complexes(complexes, 4, device=None,
                                                 sample_many_fn=lambda cs, n: [_fake_result((c["ligand"].num_nodes - 5) // 3, n) for c in cs])
        ok = all(torch.equal(a, b) for a, b in zip(poses, poses2))
        ok &= all((a is None and b is None) or torch.equal(a, b) for a, b in zip(confs, confs2))
        for i in range(n_complexes):
            p, c = _fake_result(i, 4)
            ok &= torch.equal(poses[i], p)
            ok &= (confs[i] is None) if c is None else torch.equal(confs[i], c)
        q.put((rank, ok, sorted(seen)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_complexes", [5, 1])
def test_sharded_sampling_gathers_identical_results_world2(n_complexes):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_complexes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    seen = sorted(i for _, _, s in res for i in s)
    assert seen == list(range(n_complexes))            # every complex sampled exactly once across the ranks
