"""Parity of every cb200 kernel (through the C-ABI) against the CPU oracle on seeded inputs.
Integer / index work: bit-exact.  Floating point: tolerance written in each test (BASELINE.json:
per-layer outputs within 1e-5 relative error)."""
import copy
import math

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err, rmsd, unpack_graph
from oracle import cluster, model as om, o3, sampler as osamp

pytestmark = pytest.mark.gpu


def _cloud(seed, sizes_x, sizes_y, scale=4.0):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.rand(n, 3, generator=g) * scale for n in sizes_x]
    ys = [torch.rand(n, 3, generator=g) * scale for n in sizes_y]
    x, y = torch.cat(xs), torch.cat(ys)
    bx = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes_x)])
    by = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes_y)])
    return x, y, bx, by


def _ptr(batch, B):
    p = torch.zeros(B + 1, dtype=torch.int32)
    p[1:] = torch.cumsum(torch.bincount(batch, minlength=B), 0)
    return p


@pytest.mark.parametrize("sizes_x,sizes_y,r,max_nb,use_cutoff", [
    ([40, 0, 75, 33], [7, 5, 0, 12], 1.5, 10000, False),
    ([400, 150], [40, 23], 1.0, 10000, True),
    ([64], [64], 2.0, 32, False),          # truncation: keeps the 32 lowest-index candidates
    ([1], [1], 0.5, 32, False),
    ([1000, 37, 512], [60, 10, 33], 1.0, 10000, True),
])
def test_radius_bit_exact(sizes_x, sizes_y, r, max_nb, use_cutoff):
    from confidence_bootstrapping_b200 import graph
    x, y, bx, by = _cloud(1, sizes_x, sizes_y, scale=4.0 if not use_cutoff else 60.0)
    B = len(sizes_x)
    cut = (torch.rand(B) * 40 + 20) if use_cutoff else None
    if use_cutoff:
        want = cluster.radius(x / cut[bx][:, None], y / cut[by][:, None], r, bx, by, max_num_neighbors=max_nb)
    else:
        want = cluster.radius(x, y, r, bx, by, max_num_neighbors=max_nb)
    d = "cuda"
    cap = int(sum(a * b for a, b in zip(sizes_x, sizes_y)))
    fwd = graph.radius_edges(x.to(d), _ptr(bx, B).to(d), y.to(d), by.int().to(d), r, max_nb, cap,
                             cutoff=cut.to(d) if use_cutoff else None)
    got = fwd.edge_index().cpu()
    assert torch.equal(got, want)                      # sorted by (query, candidate), bit-exact
    assert torch.equal(fwd.rowptr.cpu()[1:].long(), torch.cumsum(torch.bincount(want[0], minlength=len(y)), 0))
    # transposed list = the same edge set sorted by (candidate, query)
    tr = graph.radius_edges_transposed(x.to(d), bx.int().to(d), y.to(d), _ptr(by, B).to(d), r, cap,
                                       cutoff=cut.to(d) if use_cutoff else None, kept=fwd if max_nb < 10000 else None)
    order = torch.argsort(want[1] * (len(y) + 1) + want[0], stable=True)
    assert torch.equal(tr.edge_index().cpu(), torch.stack([want[1][order], want[0][order]]))


@pytest.mark.parametrize("n,r,max_nb", [([23, 40, 9], 5.0, 32), ([60], 3.0, 4), ([2, 1], 5.0, 32)])
def test_radius_graph_bit_exact(n, r, max_nb):
    from confidence_bootstrapping_b200 import graph
    g = torch.Generator().manual_seed(3)
    pos = torch.cat([torch.rand(k, 3, generator=g) * 6 for k in n])
    batch = torch.cat([torch.full((k,), i, dtype=torch.long) for i, k in enumerate(n)])
    want = cluster.radius_graph(pos, r, batch, max_num_neighbors=max_nb)    # row0 neighbour, row1 centre
    d, B = "cuda", len(n)
    cap = int(sum(k * k for k in n))
    fwd = graph.radius_edges(pos.to(d), _ptr(batch, B).to(d), pos.to(d), batch.int().to(d), r, max_nb + 1, cap, exclude_self=True)
    assert torch.equal(fwd.edge_index().cpu(), torch.stack([want[1], want[0]]))       # grouped by centre
    agg = graph.radius_edges_transposed(pos.to(d), batch.int().to(d), pos.to(d), _ptr(batch, B).to(d), r, cap,
                                        exclude_self=True, kept=fwd)
    order = torch.argsort(want[0] * (len(pos) + 1) + want[1], stable=True)              # grouped by edge_index[0]
    assert torch.equal(agg.edge_index().cpu(), want[:, order])


def test_exclusive_scan():
    from confidence_bootstrapping_b200 import _lib
    for n in (0, 1, 5, 4096, 4097, 100_000):
        c = torch.randint(0, 50, (n,), dtype=torch.int32)
        out = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        _lib.exclusive_scan(c.cuda(), out, torch.empty(4096, dtype=torch.int32, device="cuda"))
        want = torch.zeros(n + 1, dtype=torch.int64)
        want[1:] = torch.cumsum(c.long(), 0)
        assert torch.equal(out.cpu().long(), want)


@pytest.mark.parametrize("lmax,n_extra,sigma_first,with_sigma", [(1, 4, True, True), (2, 0, True, True), (1, 0, False, True), (2, 0, True, False)])
def test_edge_featurize(lmax, n_extra, sigma_first, with_sigma):
    """K2 vs GaussianSmearing + spherical_harmonics + 2-layer MLP of the oracle; tolerance 1e-5 relative."""
    from confidence_bootstrapping_b200 import graph
    from confidence_bootstrapping_b200.score_model import GaussianSmearing, _edge_mlp
    torch.manual_seed(0)
    ns, sd_, ng = 24 if lmax == 2 else 32, 32 if with_sigma else 0, 32
    x, y, bx, by = _cloud(5, [50, 70], [12, 9], scale=6.0)
    B = 2
    seq = _edge_mlp(n_extra + sd_ + ng, ns, 0.0)
    sm = GaussianSmearing(0.0, 5.0, ng)
    if sigma_first:
        offs = dict(extra_off=0, sigma_off=n_extra, smear_off=n_extra + sd_)
    else:
        offs = dict(extra_off=0, smear_off=n_extra, sigma_off=n_extra + ng)
    emb = graph.EdgeEmbedder(seq.cuda(), sm.cuda(), n_extra, offs["extra_off"], offs["sigma_off"], sd_, offs["smear_off"])
    edges = graph.radius_edges(x.cuda(), _ptr(bx, B).cuda(), y.cuda(), by.int().cuda(), 2.5, 10000, 50 * 12 + 70 * 9)
    E = edges.num_edges()
    assert E > 50
    sigma = torch.randn(B, sd_) if with_sigma else None
    extra = torch.randn(max(edges.cap, 1), n_extra) if n_extra else None
    attr, sh = emb(edges, y.cuda(), x.cuda(), by.int().cuda(), sigma.cuda() if with_sigma else None, lmax, sh_sign=-1.0,
                   extra=extra.cuda() if n_extra else None)
    ei = edges.edge_index().cpu()
    vec = -(x[ei[1]] - y[ei[0]])
    smear = torch.exp(sm.coeff * (vec.norm(dim=-1)[:, None] - sm.offset.cpu()[None, :]) ** 2)
    parts = {"extra": extra[:E] if n_extra else torch.zeros(E, 0), "sigma": sigma[by[ei[0]]] if with_sigma else torch.zeros(E, 0),
             "smear": smear}
    feat = torch.cat([parts["extra"], parts["sigma"], parts["smear"]] if sigma_first else [parts["extra"], parts["smear"], parts["sigma"]], 1)
    cpu = copy.deepcopy(seq).cpu()
    want_attr = cpu(feat)
    want_sh = o3.spherical_harmonics(list(range(lmax + 1)), vec, True, "component")
    assert rel_err(attr[:E], want_attr) < 1e-5
    assert rel_err(sh[:E], want_sh) < 1e-5


SEQ = ["32x0e", "32x0e + 6x1o", "32x0e + 6x1o + 6x1e", "32x0e + 6x1o + 6x1e + 6x0o"]
CONF = ["24x0e", "24x0e + 6x1o", "24x0e + 6x1o + 6x1e", "24x0e + 6x1o + 6x1e + 24x0o"]


def _random_graph(seed, n_nodes, n_edges, n_out=None):
    g = torch.Generator().manual_seed(seed)
    n_out = n_out or n_nodes
    ei = torch.stack([torch.randint(0, n_out, (n_edges,), generator=g), torch.randint(0, n_nodes, (n_edges,), generator=g)])
    ei[0, : n_edges // 3] = 0            # one heavy aggregation node (hundreds of edges, like a ligand atom)
    return ei


@pytest.mark.parametrize("in_ir,sh_l,out_ir,faster,groups,nef,residual", [
    (SEQ[0], 1, SEQ[1], True, 1, 96, True), (SEQ[1], 1, SEQ[2], True, 1, 96, True), (SEQ[2], 1, SEQ[3], True, 1, 96, True),
    (SEQ[3], 1, SEQ[3], True, 4, 96, True), (SEQ[3], 1, SEQ[3], True, 2, 96, True),
    (CONF[0], 2, CONF[1], False, 9, 72, True), (CONF[2], 2, CONF[3], False, 3, 72, True), (CONF[3], 2, CONF[3], False, 9, 72, True),
    (SEQ[3], 1, "2x1o + 2x1e", False, 1, 64, False),
])
def test_tp_conv_layer_vs_oracle(in_ir, sh_l, out_ir, faster, groups, nef, residual):
    """K3 through the reference-style layer call vs the reference formulation (materialised weights,
    gather, index_add mean, BatchNorm, residual).  Bar: 1e-5 relative (BASELINE.json north_star)."""
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    from helpers import randomize_norm_stats
    torch.manual_seed(0)
    sh_ir = "1x0e + 1x1o" if sh_l == 1 else "1x0e + 1x1o + 1x2e"
    layer = TensorProductConvLayer(in_ir, sh_ir, out_ir, nef, residual=residual, batch_norm=True, dropout=0.1,
                                   hidden_features=nef, faster=faster, edge_groups=groups)
    randomize_norm_stats(layer, seed=1)
    layer.eval()
    n_nodes, n_edges = 120, 2400
    d_in = o3.Irreps(in_ir).dim
    x = torch.randn(n_nodes, d_in)
    n_out = 7 if not residual else n_nodes
    ei = _random_graph(2, n_nodes, n_edges, n_out)
    vec = torch.randn(n_edges, 3)
    sh = o3.spherical_harmonics(list(range(sh_l + 1)), vec, True, "component")
    ea = torch.randn(n_edges, nef)
    bounds = np.linspace(0, n_edges, groups + 1).astype(int)
    ea_list = [ea[bounds[g]:bounds[g + 1]] for g in range(groups)] if groups > 1 else ea
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        want = om.tp_conv_layer({"x." + k: v for k, v in sd.items()}, "x", in_ir, o3.Irreps(sh_ir), out_ir, faster, groups,
                                residual, True, x, ei, ea_list, sh, out_nodes=n_out)
        layer = layer.cuda()
        got = layer(x.cuda(), ei.cuda(), [e.cuda() for e in ea_list] if groups > 1 else ea.cuda(), sh.cuda(), out_nodes=n_out)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-5


def test_tp_conv_empty_and_isolated_nodes():
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    torch.manual_seed(0)
    layer = TensorProductConvLayer(SEQ[3], "1x0e + 1x1o", SEQ[3], 96, faster=True).eval().cuda()
    x = torch.randn(10, 74, device="cuda")
    with torch.no_grad():
        out = layer(x, torch.zeros(2, 0, dtype=torch.long, device="cuda"), torch.zeros(0, 96, device="cuda"), torch.zeros(0, 4, device="cuda"))
        assert torch.equal(out, x)        # zero edges: no BatchNorm, just the residual (tensor_layers.py:197-198)
        ei = torch.tensor([[3, 3], [1, 2]], device="cuda")
        out = layer(x, ei, torch.randn(2, 96, device="cuda"), torch.randn(2, 4, device="cuda"))
        sd = {("x." + k): v.cpu() for k, v in layer.state_dict().items()}
    assert torch.isfinite(out).all()
    # nodes without incoming edges get BatchNorm(0) + residual
    bn0 = om._e3nn_batch_norm(sd, "x.batch_norm", SEQ[3], torch.zeros(1, 74))
    assert torch.allclose(out[0].cpu(), x[0].cpu() + bn0[0], atol=1e-6)


def test_sde_step_vs_reference_golden():
    """K4 vs the reference's modify_conformer_batch outputs (golden) and the oracle on fresh inputs; 1e-4 A."""
    from confidence_bootstrapping_b200.data import Batch
    from confidence_bootstrapping_b200.diffusion_utils import modify_conformer_batch
    import os
    g = torch.load(os.path.join(GOLDEN, "geometry.pt"), weights_only=False)
    graph = unpack_graph(g["graph"])
    b3 = Batch.from_data_list([copy.deepcopy(graph) for _ in range(3)]).to("cuda")
    mr = graph["ligand"].mask_rotate
    new = modify_conformer_batch(g["pos0"].cuda(), b3, g["tr"].cuda(), g["rot"].cuda(), g["tor"].cuda(), mr)
    assert rmsd(new, g["new_pos"]) < 1e-4
    rigid = modify_conformer_batch(g["pos0"].cuda(), b3, g["tr"].cuda(), g["rot"].cuda(), None, mr)
    assert rmsd(rigid, g["rigid_only"]) < 1e-5
    # fresh random updates incl. tiny rotation (series branch of axis_angle_to_quaternion) and a large one
    torch.manual_seed(4)
    B = 37
    bb = Batch.from_data_list([copy.deepcopy(graph) for _ in range(B)])
    R = int(graph["ligand"].edge_mask.sum())
    tr, rot, tor = torch.randn(B, 3), torch.randn(B, 3), torch.randn(B * R) * 2
    rot[0] *= 1e-9
    rot[1] *= 3
    pos0 = bb["ligand"].pos + torch.randn_like(bb["ligand"].pos) * 0.1
    want = osamp.modify_conformer_batch(pos0, bb, tr, rot, tor, torch.from_numpy(mr))
    got = modify_conformer_batch(pos0.cuda(), bb.to("cuda"), tr.cuda(), rot.cuda(), tor.cuda(), mr)
    assert rmsd(got, want) < 1e-4


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("in_ir,sh_l,out_ir,faster,groups,nef", [
    (SEQ[3], 1, SEQ[3], True, 4, 96), (SEQ[0], 1, SEQ[1], True, 1, 96), (CONF[3], 2, CONF[3], False, 9, 72),
    (SEQ[3], 1, "2x1o + 2x1e", False, 1, 64)])
def test_tp_conv_layer_other_accumulate_kernels(in_ir, sh_l, out_ir, faster, groups, nef, mode):
    """K3 with the non-default accumulate kernels: accum_mode 1 = fp32 FFMA register tiles, 2 = tcgen05 3xTF32 with the
    row-major TMEM accumulator (the default, 3, keeps the accumulator transposed and is what every other test exercises;
    layers with more than 240 f-rows fall back to 2 by themselves).  Same 1e-5 relative bar."""
    from confidence_bootstrapping_b200 import tensor_layers
    from confidence_bootstrapping_b200.tensor_layers import TensorProductConvLayer
    from helpers import randomize_norm_stats
    torch.manual_seed(0)
    sh_ir = "1x0e + 1x1o" if sh_l == 1 else "1x0e + 1x1o + 1x2e"
    residual = out_ir != "2x1o + 2x1e"
    layer = TensorProductConvLayer(in_ir, sh_ir, out_ir, nef, residual=residual, batch_norm=True, hidden_features=nef,
                                   faster=faster, edge_groups=groups)
    randomize_norm_stats(layer, seed=1)
    layer.eval()
    n_nodes, n_edges = 150, 3000
    x = torch.randn(n_nodes, o3.Irreps(in_ir).dim)
    n_out = 9 if not residual else n_nodes
    ei = _random_graph(5, n_nodes, n_edges, n_out)
    sh = o3.spherical_harmonics(list(range(sh_l + 1)), torch.randn(n_edges, 3), True, "component")
    ea = torch.randn(n_edges, nef)
    bounds = np.linspace(0, n_edges, groups + 1).astype(int)
    ea_list = [ea[bounds[g]:bounds[g + 1]] for g in range(groups)] if groups > 1 else ea
    sd = {"x." + k: v.clone() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        want = om.tp_conv_layer(sd, "x", in_ir, o3.Irreps(sh_ir), out_ir, faster, groups, residual, True, x, ei, ea_list, sh, out_nodes=n_out)
        layer = layer.cuda()
        old = tensor_layers.ACCUM_MODE
        tensor_layers.ACCUM_MODE = mode
        try:
            got = layer(x.cuda(), ei.cuda(), [e.cuda() for e in ea_list] if groups > 1 else ea.cuda(), sh.cuda(), out_nodes=n_out)
            torch.cuda.synchronize()
        finally:
            tensor_layers.ACCUM_MODE = old
    assert rel_err(got, want) < 1e-5


@pytest.mark.gpu
def test_tp_conv_gate_prunes_dead_outputs():
    """Segment.gate (cb_tp_segment.gate_rowptr): aggregation nodes without an edge in the gate list are skipped
    (output = epilogue of an empty sum); every other node is bit-identical to the ungated call."""
    from confidence_bootstrapping_b200.graph import static_edges
    from confidence_bootstrapping_b200.tensor_layers import Segment, TensorProductConvLayer
    from helpers import randomize_norm_stats
    torch.manual_seed(0)
    layer = TensorProductConvLayer(SEQ[3], "1x0e + 1x1o", SEQ[3], 32, residual=True, batch_norm=True, hidden_features=96,
                                   faster=True, edge_groups=2)
    randomize_norm_stats(layer, seed=1)
    layer = layer.eval().cuda()
    n, e = 100, 1500
    x = torch.randn(n, 74, device="cuda")
    ei_a, ei_b = _random_graph(3, n, e).cuda(), _random_graph(4, n, e // 2).cuda()
    gate_ei = torch.stack([torch.arange(0, n, 3), torch.zeros(len(range(0, n, 3)), dtype=torch.long)]).cuda()   # every 3rd node
    ea, eb = torch.randn(e, 32, device="cuda"), torch.randn(e // 2, 32, device="cuda")
    sha = o3.spherical_harmonics([0, 1], torch.randn(e, 3), True, "component").cuda()
    shb = o3.spherical_harmonics([0, 1], torch.randn(e // 2, 3), True, "component").cuda()
    (la, pa), (lb, pb), (lg, _) = static_edges(ei_a, n), static_edges(ei_b, n), static_edges(gate_ei, n)

    def run(gate):
        segs = [Segment(la, ea[pa].contiguous(), sha[pa].contiguous(), 0, 0, n, gate=gate),
                Segment(lb, eb[pb].contiguous(), shb[pb].contiguous(), 1, 0, n)]
        with torch.no_grad():
            return layer.run(x, segs, n, 0, (0, 32), residual=x)

    full, gated = run(None), run(lg)
    only_b = None
    with torch.no_grad():
        only_b = layer.run(x, [Segment(lb, eb[pb].contiguous(), shb[pb].contiguous(), 1, 0, n)], n, 0, (0, 32), residual=x)
    keep = torch.zeros(n, dtype=torch.bool, device="cuda")
    keep[::3] = True
    assert torch.equal(gated[keep], full[keep])
    # gated-out nodes only see group-1 edges; their mean is over those edges alone
    assert rel_err(gated[~keep], only_b[~keep]) < 1e-5
    assert not torch.equal(gated[~keep], full[~keep])


@pytest.mark.gpu
def test_tp_conv_workspace_chunking_is_exact():
    """Layers whose accumulator workspace exceeds WORKSPACE_BYTES run in node chunks (node_begin/node_end of the
    C-ABI call, workspace tiles counted from node_begin): same bits as the single-chunk call."""
    import confidence_bootstrapping_b200.tensor_layers as tl
    from confidence_bootstrapping_b200.graph import static_edges
    from confidence_bootstrapping_b200.tensor_layers import Segment, TensorProductConvLayer
    from helpers import randomize_norm_stats
    torch.manual_seed(0)
    layer = TensorProductConvLayer(SEQ[3], "1x0e + 1x1o", SEQ[3], 32, residual=True, batch_norm=True, hidden_features=96,
                                   faster=True, edge_groups=2)
    randomize_norm_stats(layer, seed=1)
    layer = layer.eval().cuda()
    n, e = 333, 4000                      # 333 nodes: chunk boundaries that are not multiples of the 32-node tile
    x = torch.randn(n, 74, device="cuda")
    ei_a, ei_b = _random_graph(3, n, e).cuda(), _random_graph(4, n, e // 2, n_out=200).cuda()
    ea, eb = torch.randn(e, 32, device="cuda"), torch.randn(e // 2, 32, device="cuda")
    sha = o3.spherical_harmonics([0, 1], torch.randn(e, 3), True, "component").cuda()
    shb = o3.spherical_harmonics([0, 1], torch.randn(e // 2, 3), True, "component").cuda()
    (la, pa), (lb, pb) = static_edges(ei_a, n), static_edges(ei_b, 200)
    segs = [Segment(la, ea[pa].contiguous(), sha[pa].contiguous(), 0, 0, n),
            Segment(lb, eb[pb].contiguous(), shb[pb].contiguous(), 1, 0, 200)]
    old = tl.WORKSPACE_BYTES
    try:
        with torch.no_grad():
            whole = layer.run(x, segs, n, 0, (0, 32), residual=x)
            tl.WORKSPACE_BYTES = 6 << 20        # 6 MiB: about 65 accumulators per chunk => ~9 chunks
            chunked = layer.run(x, segs, n, 0, (0, 32), residual=x)
    finally:
        tl.WORKSPACE_BYTES = old
    assert torch.equal(whole, chunked)
