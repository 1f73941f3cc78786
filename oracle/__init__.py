"""CPU oracle for the reverse-diffusion pose-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`confidence_bootstrapping_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker / baseline.

Layout
------
o3.py        restatement of the e3nn==0.5.0 ops the reference calls
             (Irreps, spherical_harmonics, wigner_3j, FullyConnectedTensorProduct,
             FullTensorProduct, BatchNorm).                      [third-party recall]
cluster.py   restatement of torch_cluster==1.6.0 radius / radius_graph. [third-party recall]
scatter.py   restatement of torch_scatter==2.0.9 scatter / scatter_mean. [third-party recall]
             (the minimal torch_geometric==2.0.4 surface the hot path touches -- HeteroData, Batch,
             DataLoader, subgraph -- is confidence_bootstrapping_b200/data.py: a container, no arithmetic)
shims.py     installs the modules above (and data.py's containers) under their third-party names so the
             REAL first-party reference (/root/reference, this container only) can
             be imported and executed; used by make_golden.py.
model.py     restatement of the first-party model code (score_model.py,
             all_atom_score_model.py, tensor_layers.py), functional, driven by a
             reference-named state_dict.
sampler.py   restatement of utils/sampling.py + diffusion_utils/torsion/geometry.
make_golden.py  generates tests/golden/* by running the real reference under shims.
make_1a0q_fixture.py  parses the reference's shipped complex data/1a0q into tests/golden/1a0q.pt (BASELINE config 1).
gen_tables.py   runs the reference's own utils/so3.py / utils/torus.py to produce the score-norm tables.

Parity status: the reference has no tests or golden vectors, and its third-party
stack is not installable here.  First-party arithmetic IS pinned: model.py /
sampler.py are checked against the real reference code executed under shims.py
(tests/golden/*).  The third-party semantics in o3/cluster/scatter/pyg are
"parity unpinned" except for the in-repo anchor FasterTensorProduct == FCTP(lmax=1)
(models/tensor_layers.py:39-117) and closed-form/equivariance self-checks.
"""
