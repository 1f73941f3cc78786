"""Irreps bookkeeping and the "TP program" that drives the fused tensor-product kernel (K3).

A tensor product with per-edge weights, as used by TensorProductConvLayer
(models/tensor_layers.py:120-217), is bilinear in (x (x) sh) and linear in the weights.  It is
compiled here, once per layer, into
  rows   one "intermediate" f-row per (path, input multiplicity u, output component k): a short sum of
         coef * x[x_idx] * sh[sh_idx] terms (Clebsch-Gordan coefficients and the path normalisation
         folded into coef), the weight row it multiplies and the output channels it feeds;
  terms  the flat term table;
  out_ptr/out_idx  for every output channel, the partial-sum slots that are added into it.
Two front ends produce the same structure:
  faster_tp_program   closed-form lmax=1 product, weight layout of FasterTensorProduct
                      (tensor_layers.py:58-64,87-93: per output irrep a [fan_in, mul_out] block / sqrt(fan_in))
  fctp_program        e3nn o3.FullyConnectedTensorProduct(shared_weights=False) (tensor_layers.py:185):
                      instructions in nested-loop order, per-instruction [mul1, mul2, mul_out] blocks,
                      alpha = sqrt((2 l_out + 1) / sum_paths mul1*mul2), real Wigner-3j coefficients.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction
from functools import lru_cache
from typing import List, Tuple

import numpy as np

ROW_DTYPE = np.dtype([("term_begin", "<i4"), ("term_end", "<i4"), ("w_base", "<i4"), ("out_base", "<i4"),
                      ("out_step", "<i4"), ("mul", "<i4"), ("p_off", "<i4"), ("pad", "<i4")])
TERM_DTYPE = np.dtype([("x_idx", "<i2"), ("sh_idx", "<i2"), ("coef", "<f4")])
# a run = consecutive f-rows feeding the same output channels (same out_base/out_step/mul) whose weight rows
# are contiguous: row r of the run uses weight rows w_base0 + (r - row_begin) * mul + m
RUN_DTYPE = np.dtype([("row_begin", "<i4"), ("row_end", "<i4"), ("out_base", "<i4"), ("out_step", "<i4"),
                      ("mul", "<i4"), ("w_base0", "<i4"), ("pad0", "<i4"), ("pad1", "<i4")])


def parse_irreps(spec) -> List[Tuple[int, int, int]]:
    """'32x0e + 6x1o' -> [(32, 0, +1), (6, 1, -1)]  (mul, l, parity)."""
    if not isinstance(spec, str):
        out = []
        for it in spec:
            if isinstance(it, str):
                out += parse_irreps(it)
            elif len(it) == 3:
                out.append((int(it[0]), int(it[1]), int(it[2])))
            else:
                mul, ir = it
                if isinstance(ir, str):
                    out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
                else:
                    l, p = ir
                    out.append((int(mul), int(l), int(p)))
        return out
    out = []
    for tok in spec.split("+"):
        tok = tok.strip()
        if not tok:
            continue
        mul, ir = tok.split("x") if "x" in tok else ("1", tok)
        out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
    return out


def irreps_dim(irreps) -> int:
    return sum(m * (2 * l + 1) for m, l, _ in parse_irreps(irreps))


def irreps_str(irreps) -> str:
    return " + ".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in parse_irreps(irreps))


def sh_irreps(lmax: int):
    return [(1, l, (-1) ** l) for l in range(lmax + 1)]


def get_irrep_seq(ns, nv, use_second_order_repr, reduce_pseudoscalars):
    """Feature irreps after 0..3 conv layers (models/tensor_layers.py:12-27)."""
    last = nv if reduce_pseudoscalars else ns
    if use_second_order_repr:
        return [f"{ns}x0e", f"{ns}x0e + {nv}x1o + {nv}x2e", f"{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o",
                f"{ns}x0e + {nv}x1o + {nv}x2e + {nv}x1e + {nv}x2o + {last}x0o"]
    return [f"{ns}x0e", f"{ns}x0e + {nv}x1o", f"{ns}x0e + {nv}x1o + {nv}x1e",
            f"{ns}x0e + {nv}x1o + {nv}x1e + {last}x0o"]


# ----------------------------------------------------------------------------- real Wigner 3j
def _su2_cg(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    pref = Fraction((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
                    f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))
    tot = Fraction(0)
    for v in range(max(-j1 + j2 + m3, -j1 + m1, 0), min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3) + 1):
        tot += Fraction((-1) ** (v + j2 + m2) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
                        f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return math.sqrt(float(pref)) * float(tot)


def _q(l):
    """Change of basis real -> complex spherical harmonics in e3nn's convention."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l - m] = s
        q[l + m, l + m] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + m] = (-1) ** m * s
        q[l + m, l - m] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real Wigner 3j symbol, Frobenius-normalised, e3nn basis (x,y,z order for l=1)."""
    cg = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=np.complex128)
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                cg[l1 + m1, l2 + m2, l3 + m1 + m2] = _su2_cg(l1, m1, l2, m2, l3, m1 + m2)
    w = np.einsum("ij,kl,mn,ikn->jlm", _q(l1), _q(l2), np.conj(_q(l3).T), cg)
    assert np.abs(w.imag).max() < 1e-9
    w = w.real
    return w / np.linalg.norm(w)


# ----------------------------------------------------------------------------- program container
@dataclass
class TPProgram:
    rows: np.ndarray        # ROW_DTYPE [R]
    terms: np.ndarray       # TERM_DTYPE [T]
    out_ptr: np.ndarray     # int32 [d_out + 1]
    out_idx: np.ndarray     # int32 [n_slots]
    weight_numel: int
    d_in: int
    d_out: int
    sh_dim: int
    runs: np.ndarray = None  # RUN_DTYPE [n_runs]

    @property
    def n_rows(self):
        return len(self.rows)

    @property
    def n_slots(self):
        return len(self.out_idx)


class _Builder:
    def __init__(self):
        self.rows, self.terms = [], []

    def row(self, terms, w_base, out_base, out_step, mul):
        tb = len(self.terms)
        self.terms += [(int(x), int(s), float(c)) for (x, s, c) in terms if c != 0.0]
        self.rows.append((tb, len(self.terms), int(w_base), int(out_base), int(out_step), int(mul)))

    def finish(self, weight_numel, d_in, d_out, sh_dim) -> TPProgram:
        rows = np.zeros(len(self.rows), dtype=ROW_DTYPE)
        p = 0
        slots = [[] for _ in range(d_out)]
        for r, (tb, te, wb, ob, os_, mul) in enumerate(self.rows):
            rows[r] = (tb, te, wb, ob, os_, mul, p, 0)
            for m in range(mul):
                slots[ob + m * os_].append(p + m)
            p += mul
        terms = np.zeros(len(self.terms), dtype=TERM_DTYPE)
        for t, v in enumerate(self.terms):
            terms[t] = v
        out_ptr = np.zeros(d_out + 1, dtype=np.int32)
        out_ptr[1:] = np.cumsum([len(s) for s in slots])
        out_idx = np.asarray([q for s in slots for q in s], dtype=np.int32)
        runs = []
        for r, (_, _, wb, ob, os_, mul) in enumerate(self.rows):
            if runs and runs[-1][2:5] == [ob, os_, mul] and wb == runs[-1][5] + (r - runs[-1][0]) * mul:
                runs[-1][1] = r + 1
            else:
                runs.append([r, r + 1, ob, os_, mul, wb])
        run_arr = np.zeros(len(runs), dtype=RUN_DTYPE)
        for k, (rb, re_, ob, os_, mul, wb) in enumerate(runs):
            run_arr[k] = (rb, re_, ob, os_, mul, wb, 0, 0)
        return TPProgram(rows, terms, out_ptr, out_idx, int(weight_numel), int(d_in), int(d_out), int(sh_dim), run_arr)


def _offsets(irreps):
    offs, o = [], 0
    for m, l, _ in irreps:
        offs.append(o)
        o += m * (2 * l + 1)
    return offs, o


# ----------------------------------------------------------------------------- FasterTensorProduct
def faster_tp_program(in_irreps, out_irreps) -> TPProgram:
    """Program + weight layout of the reference FasterTensorProduct (sh = 1x0e + 1x1o)."""
    ins, outs = parse_irreps(in_irreps), parse_irreps(out_irreps)
    key = lambda l, p: f"{l}{'e' if p == 1 else 'o'}"
    in_off, d_in = _offsets(ins)
    out_off, d_out = _offsets(outs)
    im = {"0e": (0, 0), "1o": (0, 0), "1e": (0, 0), "0o": (0, 0)}  # key -> (mul, offset)
    om = dict(im)
    for (m, l, p), o in zip(ins, in_off):
        im[key(l, p)] = (m, o)
    for (m, l, p), o in zip(outs, out_off):
        om[key(l, p)] = (m, o)
    eps = {(0, 1, 2): 1.0, (1, 2, 0): 1.0, (2, 0, 1): 1.0, (0, 2, 1): -1.0, (2, 1, 0): -1.0, (1, 0, 2): -1.0}
    Y0, Y1 = 0, 1  # sh layout: [Y0 | Y1x Y1y Y1z]

    def scalar_times(k, y):       # x_k[i] * sh[y]
        return [[[(im[k][1] + i, y, 1.0)]] for i in range(im[k][0])]

    def scalar_times_vec(k):      # x_k[i] * Y1[c]
        return [[[(im[k][1] + i, Y1 + c, 1.0)] for c in range(3)] for i in range(im[k][0])]

    def vec_times_scalar(k):      # x_k[u][c] * Y0
        return [[[(im[k][1] + 3 * u + c, Y0, 1.0)] for c in range(3)] for u in range(im[k][0])]

    def vec_dot(k):               # (x_k[u] . Y1) / sqrt(3)
        return [[[(im[k][1] + 3 * u + c, Y1 + c, 1.0 / math.sqrt(3)) for c in range(3)]] for u in range(im[k][0])]

    def vec_cross(k):             # (x_k[u] x Y1)[c] / sqrt(2)
        return [[[(im[k][1] + 3 * u + a, Y1 + b, s / math.sqrt(2)) for (a, b, cc), s in eps.items() if cc == c]
                 for c in range(3)] for u in range(im[k][0])]

    # intermediates per output irrep, in the reference's concatenation order (tensor_layers.py:72-85)
    inter = {
        "0e": scalar_times("0e", Y0) + vec_dot("1o"),
        "1o": scalar_times_vec("0e") + vec_times_scalar("1o") + vec_cross("1e"),
        "1e": vec_cross("1o") + vec_times_scalar("1e") + scalar_times_vec("0o"),
        "0o": vec_dot("1e") + scalar_times("0o", Y0),
    }
    fan = {"0e": im["0e"][0] + im["1o"][0], "1o": im["0e"][0] + im["1o"][0] + im["1e"][0],
           "1e": im["1o"][0] + im["1e"][0] + im["0o"][0], "0o": im["1e"][0] + im["0o"][0]}
    b = _Builder()
    start = 0
    for k in ("0e", "1o", "1e", "0o"):  # weight blocks are sliced in this fixed order (:58-63)
        mul_out, o_off = om[k]
        dim = 3 if k[0] == "1" else 1
        if mul_out > 0:
            assert len(inter[k]) == fan[k]
            scale = 1.0 / math.sqrt(fan[k])
            for c in range(dim):          # component-major: consecutive rows share their output channels
                for i, comps in enumerate(inter[k]):
                    b.row([(x, s, v * scale) for (x, s, v) in comps[c]], start + i * mul_out, o_off + c, dim, mul_out)
        start += fan[k] * mul_out
    return b.finish(start, d_in, d_out, 4)


# ----------------------------------------------------------------------------- e3nn FCTP
def fctp_program(in_irreps, sh_irreps_, out_irreps) -> TPProgram:
    """Program + weight layout of e3nn FullyConnectedTensorProduct with per-edge weights ('uvw')."""
    ins, shs, outs = parse_irreps(in_irreps), parse_irreps(sh_irreps_), parse_irreps(out_irreps)
    in_off, d_in = _offsets(ins)
    sh_off, d_sh = _offsets(shs)
    out_off, d_out = _offsets(outs)
    instr = [(i1, i2, io) for i1, (_, l1, p1) in enumerate(ins) for i2, (_, l2, p2) in enumerate(shs)
             for io, (_, lo, po) in enumerate(outs) if abs(l1 - l2) <= lo <= l1 + l2 and p1 * p2 == po]
    fan = [0] * len(outs)
    for (i1, i2, io) in instr:
        fan[io] += ins[i1][0] * shs[i2][0]
    b = _Builder()
    woff = 0
    for (i1, i2, io) in instr:
        m1, l1, _ = ins[i1]
        m2, l2, _ = shs[i2]
        mo, lo, _ = outs[io]
        if m2 != 1:
            raise NotImplementedError("edge harmonics always have multiplicity 1 on this path")
        alpha = math.sqrt((2 * lo + 1) / fan[io])
        w3 = wigner_3j(l1, l2, lo)
        for k in range(2 * lo + 1):       # component-major: consecutive rows share their output channels
            for u in range(m1):
                terms = [(in_off[i1] + u * (2 * l1 + 1) + i, sh_off[i2] + j, alpha * w3[i, j, k])
                         for i in range(2 * l1 + 1) for j in range(2 * l2 + 1) if abs(w3[i, j, k]) > 1e-12]
                b.row(terms, woff + u * mo, out_off[io] + k, 2 * lo + 1, mo)
        woff += m1 * m2 * mo
    return b.finish(woff, d_in, d_out, d_sh)


def full_tp_1o_block(lmax: int):
    """Coefficients of the 1o output block of e3nn FullTensorProduct(sh(lmax), '2e')
    (models/score_model.py:265).  Only this block can reach the 0e/0o outputs of tor_bond_conv from
    l<=1 node features; for lmax=1 it is the first irrep of the sorted output '1x1o+1x2o+1x2e+1x3o'.
    Returns (dim_out_total, offset_of_1o, coef[3 (sh1 i), 5 (Y2 j), 3 (k)]) with the sqrt(2l+1) path weight.
    """
    w = wigner_3j(1, 2, 1) * math.sqrt(3.0)
    # sorted FullTP output irreps by (l, p), odd before even
    outs = []
    for l1 in range(lmax + 1):
        p1 = (-1) ** l1
        for lo in range(abs(l1 - 2), l1 + 2 + 1):
            outs.append((lo, p1))
    order = sorted(range(len(outs)), key=lambda i: (outs[i][0], outs[i][1], i))
    off, o = {}, 0
    for i in order:
        off[i] = o
        o += 2 * outs[i][0] + 1
    first_1o = next(i for i in range(len(outs)) if outs[i] == (1, -1))
    return o, off[first_1o], w


def full_tp_low_blocks(lmax: int):
    """The l<=1 output blocks of e3nn FullTensorProduct(sh(lmax), '2e') in its sorted output order
    ((l, p) ascending, odd before even): the only blocks tor_bond_conv can couple to l<=1 node features.
    Returns (irreps_str, [(l_in, coef[2l_in+1, 5, 2l_out+1])...]) with the sqrt(2 l_out + 1) path weight."""
    blocks = []
    for l1 in range(lmax + 1):
        p1 = (-1) ** l1
        for lo in range(abs(l1 - 2), l1 + 2 + 1):
            if lo <= 1:
                blocks.append((lo, p1, l1))
    blocks.sort(key=lambda b: (b[0], b[1]))
    irreps = " + ".join(f"1x{lo}{'e' if p == 1 else 'o'}" for lo, p, _ in blocks)
    return irreps, [(l1, wigner_3j(l1, 2, lo) * math.sqrt(2 * lo + 1)) for lo, p, l1 in blocks]


# ----------------------------------------------------------------------------- tensor-core transform plan
CHAIN_DTYPE = np.dtype([("row0", "<i4"), ("row1", "<i4"), ("row2", "<i4"), ("n_comp", "<i4"), ("w_off", "<i4"), ("npad", "<i4"),
                        ("acc_col", "<i4"), ("first", "<i4")])
BLOCK_DTYPE = np.dtype([("n_comp", "<i4"), ("mul", "<i4"), ("npad", "<i4"), ("acc_col0", "<i4"), ("n_partials", "<i4"),
                        ("out_step", "<i4"), ("out_base0", "<i4"), ("out_base1", "<i4"), ("out_base2", "<i4"), ("pad0", "<i4"),
                        ("pad1", "<i4"), ("pad2", "<i4")])
MAX_CHAIN_MMAS = 600     # MMAs accumulated into one TMEM partial sum (fp32 accumulation error grows with the chain length)


@dataclass
class TransformPlan:
    """How the tcgen05 transform kernel (tp_transform_tc.cuh) walks a layer's program.

    A *block* = the runs (one per output component) that share their weight rows: for every u of the block the kernel
    stacks the components' accumulator rows (row_c + u) along M (component c, node rank) and multiplies by the `mul` weight rows
    of u -- one MMA chain D[(c, rank)][m] += A[(c, rank)][:] . W[u, m][:].  Chains of a block go to `n_partials` TMEM partial
    sums of at most MAX_CHAIN_MMAS MMAs each, which the epilogue adds in fp32."""
    chains: np.ndarray      # CHAIN_DTYPE
    blocks: np.ndarray      # BLOCK_DTYPE
    kp: int                 # K padded to the MMA's k-step (8)
    w_floats: int           # floats of one group's W tile buffer
    acc_cols: int           # TMEM columns of one accumulator set
    w_index: list           # per chain: (first weight row, mul, npad) for building the W tiles


def transform_plan(prog: TPProgram, H: int):
    """None when the program does not fit the tensor-core transform (more than 3 stacked components, > 32 outputs per row,
    more than 256 TMEM columns per accumulator set)."""
    ha = H + 4
    kp = (ha + 7) // 8 * 8
    mmas_per_chain = (kp // 8) * 3
    upp = max(1, MAX_CHAIN_MMAS // mmas_per_chain)
    runs = prog.runs
    groups = {}
    for r in runs:          # runs that share their weights = the components of one block
        key = (int(r["w_base0"]), int(r["mul"]), int(r["row_end"] - r["row_begin"]), int(r["out_step"]))
        groups.setdefault(key, []).append(r)
    chains, blocks, w_index = [], [], []
    acc_col = 0
    w_off = 0
    for (w_base0, mul, n_u, out_step), rs in groups.items():
        if len(rs) > 3 or mul > 32:
            return None
        npad = (mul + 15) // 16 * 16
        n_part = -(-n_u // upp)
        ob = [int(r["out_base"]) for r in rs] + [0, 0, 0]
        blocks.append((len(rs), mul, npad, acc_col, n_part, out_step, ob[0], ob[1], ob[2], 0, 0, 0))
        rows0 = [int(r["row_begin"]) for r in rs] + [-1, -1, -1]
        for u in range(n_u):
            chains.append((rows0[0] + u, rows0[1] + u if rows0[1] >= 0 else -1, rows0[2] + u if rows0[2] >= 0 else -1, len(rs), w_off, npad,
                           acc_col + (u // upp) * npad, int(u % upp == 0)))
            w_index.append((w_base0 + u * mul, mul, npad))
            w_off += 2 * npad * kp
        acc_col += n_part * npad
    if acc_col > 256:
        return None
    ch = np.zeros(len(chains), dtype=CHAIN_DTYPE)
    for k, c in enumerate(chains):
        ch[k] = c
    bl = np.zeros(len(blocks), dtype=BLOCK_DTYPE)
    for k, b in enumerate(blocks):
        bl[k] = b
    return TransformPlan(ch, bl, kp, w_off, acc_col, w_index)
