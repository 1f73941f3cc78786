"""Layer-level accuracy of the two accumulate paths against the fp32 and fp64 oracle (run on the GPU box)."""
import os, sys, copy, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import randomize_norm_stats, rel_err
from oracle import o3, model as om
import confidence_bootstrapping_b200.tensor_layers as tl
from test_gpu_kernels import SEQ, CONF, _random_graph
cases = [(SEQ[3], 1, SEQ[3], True, 4, 96, True), (SEQ[3], 1, "2x1o + 2x1e", False, 1, 64, False), (CONF[3], 2, CONF[3], False, 9, 72, True), (SEQ[2], 1, SEQ[3], True, 1, 96, True)]
for (in_ir, sh_l, out_ir, faster, groups, nef, residual) in cases:
    torch.manual_seed(0)
    sh_ir = "1x0e + 1x1o" if sh_l == 1 else "1x0e + 1x1o + 1x2e"
    layer = tl.TensorProductConvLayer(in_ir, sh_ir, out_ir, nef, residual=residual, batch_norm=True, dropout=0.1, hidden_features=nef, faster=faster, edge_groups=groups)
    randomize_norm_stats(layer, seed=1); layer.eval()
    n_nodes, n_edges = 120, 2400
    x = torch.randn(n_nodes, o3.Irreps(in_ir).dim)
    n_out = 7 if not residual else n_nodes
    ei = _random_graph(2, n_nodes, n_edges, n_out)
    vec = torch.randn(n_edges, 3)
    sh = o3.spherical_harmonics(list(range(sh_l + 1)), vec, True, "component")
    ea = torch.randn(n_edges, nef)
    bounds = np.linspace(0, n_edges, groups + 1).astype(int)
    ea_list = [ea[bounds[g]:bounds[g + 1]] for g in range(groups)] if groups > 1 else ea
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        want = om.tp_conv_layer({"x." + k: v for k, v in sd.items()}, "x", in_ir, o3.Irreps(sh_ir), out_ir, faster, groups, residual, True, x, ei, ea_list, sh, out_nodes=n_out)
        try:
            dd = lambda t: t.double() if t.is_floating_point() else t
            want64 = om.tp_conv_layer({"x." + k: dd(v) for k, v in sd.items()}, "x", in_ir, o3.Irreps(sh_ir), out_ir, faster, groups, residual, True, dd(x), ei,
                                      [dd(e) for e in ea_list] if groups > 1 else dd(ea), dd(sh), out_nodes=n_out)
        except Exception as e:
            print("fp64 oracle failed:", repr(e)[:200]); want64 = None
        lc = copy.deepcopy(layer).cuda()
        line = f"{out_ir[:14]:14s} g={groups} "
        if want64 is not None: line += f"oracle32-vs-64 {rel_err(want, want64):.2e} | "
        for mode in (1, 2):   # 1 = fp32 FFMA accumulate, 2 = tcgen05 3xTF32 accumulate (default)
            tl.ACCUM_MODE = mode
            got = lc(x.cuda(), ei.cuda(), [e.cuda() for e in ea_list] if groups > 1 else ea.cuda(), sh.cuda(), out_nodes=n_out)
            line += f"mode{mode}: vs32 {rel_err(got, want):.2e}"
            if want64 is not None: line += f" vs64 {rel_err(got, want64):.2e}"
            line += " | "
        print(line)
