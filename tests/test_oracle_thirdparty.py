"""Self-checks of the third-party restatements (oracle/o3.py, cluster.py, scatter.py): closed forms,
orthogonality, equivariance, brute force.  These semantics are [third-party recall] (e3nn 0.5.0,
torch_cluster 1.6.0, torch_scatter 2.0.9 are not installable offline); the in-repo anchor is
FasterTensorProduct == FCTP(lmax=1), checked in test_golden_oracle.py against the reference's own output."""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import cluster, o3, scatter


def _rot(seed):
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    return q * torch.linalg.det(q)


def test_w3j_known_values():
    eps = torch.zeros(3, 3, 3, dtype=torch.float64)
    for i, j, k in [(0, 1, 2), (1, 2, 0), (2, 0, 1)]:
        eps[i, j, k], eps[i, k, j] = 1, -1
    assert torch.allclose(o3.wigner_3j(1, 1, 1), eps / math.sqrt(6), atol=1e-12)
    assert torch.allclose(o3.wigner_3j(1, 1, 0)[:, :, 0], torch.eye(3, dtype=torch.float64) / math.sqrt(3), atol=1e-12)
    for l in (1, 2):
        assert torch.allclose(o3.wigner_3j(0, l, l)[0], torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1), atol=1e-12)
    assert torch.allclose(o3.wigner_3j(1, 2, 1), o3.wigner_3j(1, 1, 2).permute(0, 2, 1), atol=1e-12)
    for (a, b, c) in [(1, 1, 2), (1, 2, 1), (2, 2, 2), (1, 2, 3)]:
        w = o3.wigner_3j(a, b, c)
        assert torch.allclose((w ** 2).sum((0, 1)), torch.full((2 * c + 1,), 1.0 / (2 * c + 1), dtype=torch.float64), atol=1e-12)


def test_sh_closed_forms_and_w3j_link():
    u = torch.randn(7, 3, dtype=torch.float64)
    sh = o3.spherical_harmonics([0, 1, 2], u, normalize=True, normalization="component")
    un = u / u.norm(dim=-1, keepdim=True)
    assert torch.allclose(sh[:, 0], torch.ones(7, dtype=torch.float64))
    assert torch.allclose(sh[:, 1:4], math.sqrt(3) * un)
    assert torch.allclose((sh[:, 4:] ** 2).sum(-1), torch.full((7,), 5.0, dtype=torch.float64))
    y2 = math.sqrt(15 / 2) * torch.einsum("ijk,zi,zj->zk", o3.wigner_3j(1, 1, 2), un, un)
    assert torch.allclose(sh[:, 4:] / math.sqrt(5), y2, atol=1e-12)
    z = o3.spherical_harmonics([1], torch.zeros(1, 3), True, "component")
    assert torch.isfinite(z).all() and z.abs().max() == 0


def test_sh_l2_equivariance():
    R = _rot(1)
    u = torch.randn(9, 3, dtype=torch.float64)
    y, yr = (o3.spherical_harmonics([2], v, True, "component") for v in (u, u @ R.T))
    # D2(R) is the unique matrix with Y2(Rv) = D2 Y2(v): least squares from samples, then check orthogonality
    D = torch.linalg.lstsq(y, yr).solution.T
    assert torch.allclose(D @ D.T, torch.eye(5, dtype=torch.float64), atol=1e-9)
    assert torch.allclose(yr, y @ D.T, atol=1e-9)


@pytest.mark.parametrize("lmax", [1, 2])
def test_fctp_equivariance(lmax):
    """Rotating the inputs rotates the 1o / 1e outputs and leaves the scalars; parity of 1e is even."""
    torch.manual_seed(0)
    irr = "4x0e + 3x1o + 3x1e + 2x0o"
    tp = o3.FullyConnectedTensorProduct(irr, o3.Irreps.spherical_harmonics(lmax), irr)
    x = torch.randn(6, o3.Irreps(irr).dim, dtype=torch.float64)
    v = torch.randn(6, 3, dtype=torch.float64)
    w = torch.randn(6, tp.weight_numel, dtype=torch.float64)
    R = _rot(3)

    def rot_feat(f):
        f = f.clone()
        f[:, 4:13] = (f[:, 4:13].reshape(-1, 3, 3) @ R.T).reshape(-1, 9)
        f[:, 13:22] = (f[:, 13:22].reshape(-1, 3, 3) @ R.T).reshape(-1, 9)
        return f

    sh = lambda vec: o3.spherical_harmonics(list(range(lmax + 1)), vec, True, "component")
    assert torch.allclose(tp(rot_feat(x), sh(v @ R.T), w), rot_feat(tp(x, sh(v), w)), atol=1e-10)
    # improper rotation (inversion): 1o and 0o flip, 1e and 0e do not
    flip = torch.cat([torch.ones(4), -torch.ones(9), torch.ones(9), -torch.ones(2)]).double()
    assert torch.allclose(tp(x * flip, sh(-v), w), tp(x, sh(v), w) * flip, atol=1e-10)


def test_fctp_weight_layout_and_numel():
    tp = o3.FullyConnectedTensorProduct("24x0e + 6x1o + 6x1e + 24x0o", "1x0e + 1x1o + 1x2e", "24x0e + 6x1o + 6x1e + 24x0o")
    assert tp.weight_numel == 1944 and len(tp.instructions) == 12
    tp = o3.FullyConnectedTensorProduct("32x0e + 6x1o + 6x1e + 6x0o", "1x0e + 1x1o", "2x1o + 2x1e")
    assert tp.weight_numel == 124
    ftp = o3.FullTensorProduct("1x0e + 1x1o", "2e")
    assert str(ftp.irreps_out) == "1x1o+1x2o+1x2e+1x3o" and ftp.irreps_out.dim == 20
    assert str(o3.FullTensorProduct("1x0e + 1x1o + 1x2e", "2e").irreps_out) == "1x0e+1x1o+1x1e+1x2o+1x2e+1x2e+1x3o+1x3e+1x4e"
    tor = o3.FullyConnectedTensorProduct("32x0e + 6x1o + 6x1e + 6x0o", ftp.irreps_out, "32x0o + 32x0e")
    assert tor.weight_numel == 384


def test_batchnorm_eval_and_train():
    torch.manual_seed(0)
    bn = o3.BatchNorm("3x0e + 2x1o + 2x0o")
    x = torch.randn(50, 3 + 6 + 2) * 2 + 1
    bn.train()
    y = bn(x)
    assert torch.allclose(y[:, :3].mean(0), torch.zeros(3), atol=1e-5)           # 0e: centred
    assert torch.allclose((y[:, :3] ** 2).mean(0), torch.ones(3), atol=1e-3)
    assert torch.allclose((y[:, 3:9].reshape(50, 2, 3) ** 2).mean((0, 2)), torch.ones(2), atol=1e-3)
    assert y[:, 9:].mean(0).abs().min() > 0.1                                     # 0o: NOT centred
    bn.eval()
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    y = bn(x)
    assert torch.allclose(y[:, 0], (x[:, 0] - rm[0]) / torch.sqrt(rv[0] + 1e-5), atol=1e-6)
    assert torch.allclose(y[:, 9], x[:, 9] / torch.sqrt(rv[5] + 1e-5), atol=1e-6)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 3), st.integers(0, 40), st.integers(0, 25),
       st.sampled_from([0.5, 1.0, 2.5]), st.sampled_from([3, 32, 10000]))
def test_radius_matches_brute_force(seed, n_graphs, nx, ny, r, max_nb):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.rand(int(torch.randint(0, nx + 1, (1,), generator=g)), 3, generator=g) * 3 for _ in range(n_graphs)]
    ys = [torch.rand(int(torch.randint(0, ny + 1, (1,), generator=g)), 3, generator=g) * 3 for _ in range(n_graphs)]
    x, y = torch.cat(xs), torch.cat(ys)
    bx = torch.cat([torch.full((len(a),), i) for i, a in enumerate(xs)]).long()
    by = torch.cat([torch.full((len(a),), i) for i, a in enumerate(ys)]).long()
    e = cluster.radius(x, y, r, bx, by, max_num_neighbors=max_nb)
    want = []
    for q in range(len(y)):
        hits = [c for c in range(len(x)) if bx[c] == by[q] and
                float(((y[q, 0] - x[c, 0]) ** 2 + (y[q, 1] - x[c, 1]) ** 2) + (y[q, 2] - x[c, 2]) ** 2) < np.float32(r) * np.float32(r)]
        want += [(q, c) for c in hits[:max_nb]]
    assert [tuple(p) for p in e.T.tolist()] == want


def test_radius_graph_conventions():
    pos = torch.tensor([[0.0, 0, 0], [1, 0, 0], [3, 0, 0], [0, 0.5, 0]])
    e = cluster.radius_graph(pos, 1.5, torch.zeros(4, dtype=torch.long))
    assert e.tolist() == [[1, 3, 0, 3, 0, 1], [0, 0, 1, 1, 3, 3]]  # row 0 neighbour, row 1 centre, grouped by centre
    # truncation: max_num_neighbors+1 candidates including self, lowest index first
    line = torch.stack([torch.arange(6).float() * 0.1, torch.zeros(6), torch.zeros(6)], 1)
    e = cluster.radius_graph(line, 5.0, None, max_num_neighbors=2)
    assert e[:, e[1] == 5].tolist() == [[0, 1, 2], [5, 5, 5]]   # self not among the first 3 -> 3 neighbours kept
    assert e[:, e[1] == 0].tolist() == [[1, 2], [0, 0]]


def test_scatter_mean():
    src = torch.tensor([[1.0, 2], [3, 4], [5, 6]])
    out = scatter.scatter(src, torch.tensor([2, 0, 2]), dim=0, dim_size=4, reduce="mean")
    assert out.tolist() == [[3, 4], [0, 0], [3, 4], [0, 0]]
