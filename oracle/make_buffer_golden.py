"""tests/golden/cb_buffer.pt: the REAL bootstrapping/buffer.py:CBBuffer (imported from /root/reference under oracle/shims.py)
driven through a seeded sequence of add_complexes / get calls; records which complexes it holds after every call and which
it hands out.  Build container only.  TEST INFRASTRUCTURE.

    python oracle/make_buffer_golden.py
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import shims  # noqa: E402

shims.install()
from confidence_bootstrapping_b200.data import HeteroData  # noqa: E402
from helpers import GOLDEN  # noqa: E402

NAMES = ["1abc_LIG_1", "1abc_LIG_2", "2xyz_AAA_1", "3foo_BBB_9"]


def fake_complex(name, tag, n_lig=5, n_rec=7):
    g = HeteroData()
    g.name = [name]
    g.tag = tag
    g["ligand"].pos = torch.full((n_lig, 3), float(tag))
    g["receptor"].pos = torch.zeros(n_rec, 3)
    return g


def script(seed):
    """The same call sequence for the reference class and for the cb200 class (tests/test_host_logic.py)."""
    rng = np.random.default_rng(seed)
    rounds, tag = [], 0
    for it in range(5):
        new = []
        for _ in range(int(rng.integers(2, 7))):
            new.append((NAMES[int(rng.integers(0, len(NAMES)))], tag, float(rng.normal())))
            tag += 1
        rounds.append(new)
    return rounds


def drive(cls, kwargs, rounds, get_seed, **extra):
    buf = cls(cluster_name="c0", **kwargs, **extra)
    trace = []
    for new in rounds:
        buf.add_complexes([(fake_complex(n, t), torch.tensor(c)) for n, t, c in new])
        held = [(int(g.tag), float(g.confidence), int(g.iteration)) for g in buf.complexes]
        np.random.seed(get_seed)
        got = [int(buf.get(i).tag) for i in range(min(6, buf.len()))]
        trace.append({"held": held, "len": int(buf.len()), "got": got, "cnt": dict(buf.ligand_cnt)})
    return trace


def main():
    wd = tempfile.mkdtemp()
    os.makedirs(os.path.join(wd, "data/BindingMOAD_2020_processed"))
    with open(os.path.join(wd, "data/BindingMOAD_2020_processed/new_cluster_to_ligands.pkl"), "wb") as f:
        pickle.dump({"c0": NAMES}, f)
    os.chdir(wd)
    sys.path.insert(0, "/root/reference")
    from bootstrapping.buffer import CBBuffer  # noqa: E402  (the real one)
    cases = [dict(), dict(max_complexes_per_couple=3, buffer_decay=0.2), dict(fixed_length=11, temperature=2.0, max_complexes_per_couple=4),
             dict(reset_buffer=True, multiplicity=3)]
    out = []
    for k, kw in enumerate(cases):
        rounds = script(10 + k)
        out.append({"kwargs": kw, "rounds": rounds, "get_seed": 100 + k, "trace": drive(CBBuffer, kw, rounds, 100 + k)})
    torch.save({"names": NAMES, "cases": out}, os.path.join(GOLDEN, "cb_buffer.pt"))
    print("wrote", os.path.join(GOLDEN, "cb_buffer.pt"), [len(c["trace"]) for c in out])


if __name__ == "__main__":
    main()
