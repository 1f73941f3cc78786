"""Builds tests/golden/1a0q.pt from the reference's only shipped complex (BASELINE.json config 1).  TEST INFRASTRUCTURE.

Run in the build container (reads /root/reference/data/1a0q/*, which does not exist on the GPU box):
    python oracle/make_1a0q_fixture.py
The fixture holds parsed DATA (coordinates, residue / element indices, bonds), not reference source.  LM embeddings are
seeded N(0,1) and regenerated at load time (tests/helpers.load_1a0q) instead of being stored (416 x 1280 floats).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from confidence_bootstrapping_b200.pdbsdf import load_complex  # noqa: E402
from helpers import pack_graph  # noqa: E402

SRC = "/root/reference/data/1a0q"


def main():
    g = load_complex(os.path.join(SRC, "1a0q_protein_processed.pdb"), os.path.join(SRC, "1a0q_ligand.sdf"), all_atoms=True,
                     name="1a0q", lm_dim=0)
    d = pack_graph(g)
    d["nodes"]["receptor"]["x"] = d["nodes"]["receptor"]["x"].to(torch.int16)           # residue-type index only
    d["nodes"]["atom"]["x"] = d["nodes"]["atom"]["x"].to(torch.int16)
    for et in d["edges"]:
        d["edges"][et]["edge_index"] = d["edges"][et]["edge_index"].to(torch.int32)
    d["nodes"]["ligand"]["x"] = d["nodes"]["ligand"]["x"].to(torch.int16)
    out = os.path.join(ROOT, "tests", "golden", "1a0q.pt")
    torch.save(d, out)
    print(out, os.path.getsize(out), "bytes;", g["receptor"].num_nodes, "residues,", g["atom"].num_nodes, "atoms,",
          g["ligand"].num_nodes, "ligand atoms,", g["ligand", "ligand"].num_edges, "bond edges,", int(g["ligand"].edge_mask.sum()),
          "rotatable,", g["receptor", "receptor"].num_edges, "rec edges")


if __name__ == "__main__":
    main()
