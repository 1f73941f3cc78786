"""Restatement of the e3nn==0.5.0 operations the reference hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  e3nn is pinned at 0.5.0 by the
reference (`environment.yml:117`) but is not installable here, so this file
restates its published behaviour [third-party recall].  Call sites it serves:

  o3.Irreps / Irreps.spherical_harmonics  models/score_model.py:72, tensor_layers.py:49-56
  o3.spherical_harmonics                  models/score_model.py:436,519,536,581,582,647,661
  o3.FullyConnectedTensorProduct          models/tensor_layers.py:185
  o3.FullTensorProduct                    models/score_model.py:265
  e3nn.nn.BatchNorm                       models/tensor_layers.py:193

Pinned by in-repo anchors only: FCTP(lmax=1) == reference FasterTensorProduct
(tensor_layers.py:39-117); everything else is checked by closed forms and
equivariance (tests/test_oracle_thirdparty.py).
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache

import torch
from torch import nn


# --------------------------------------------------------------------------- irreps
class Irrep(tuple):
    """(l, p) with p=+1 even / -1 odd; tuple order gives e3nn's sort order (odd before even)."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                return tuple.__new__(cls, (int(s[:-1]), {"e": 1, "o": -1}[s[-1]]))
            l, p = l
        return tuple.__new__(cls, (int(l), int(p)))

    @property
    def l(self):
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self[0] + 1

    def is_scalar(self):
        return self[0] == 0 and self[1] == 1

    def __repr__(self):
        return f"{self[0]}{'e' if self[1] == 1 else 'o'}"

    __str__ = __repr__

    def __mul__(self, other):
        other = Irrep(other)
        return [Irrep(l, self.p * other.p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return tuple.__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self[0] * self[1].dim


class Irreps(tuple):
    def __new__(cls, spec=None):
        if isinstance(spec, Irreps):
            return spec
        items = []
        if spec is None:
            pass
        elif isinstance(spec, str):
            for tok in spec.split("+"):
                tok = tok.strip()
                if not tok:
                    continue
                if "x" in tok:
                    m, ir = tok.split("x")
                    items.append(_MulIr(int(m), Irrep(ir)))
                else:
                    items.append(_MulIr(1, Irrep(tok)))
        else:
            for it in spec:
                if isinstance(it, (str, Irrep)) and not isinstance(it, _MulIr):
                    items.append(_MulIr(1, Irrep(it)))
                else:
                    m, ir = it
                    items.append(_MulIr(m, Irrep(ir)))
        return tuple.__new__(cls, items)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self):
        return sum(mi.mul for mi in self)

    def slices(self):
        out, i = [], 0
        for mi in self:
            out.append(slice(i, i + mi.dim))
            i += mi.dim
        return out

    def sort(self):
        order = sorted((mi.ir, i, mi.mul) for i, mi in enumerate(self))
        inv = tuple(i for _, i, _ in order)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        return Irreps([(mul, ir) for ir, _, mul in order]), tuple(p), inv

    def simplify(self):
        out = []
        for mi in self:
            if out and out[-1][1] == mi.ir:
                out[-1] = (out[-1][0] + mi.mul, mi.ir)
            elif mi.mul > 0:
                out.append((mi.mul, mi.ir))
        return Irreps(out)

    def __contains__(self, ir):
        ir = Irrep(ir)
        return any(mi.ir == ir for mi in self)

    def __repr__(self):
        return "+".join(f"{mi.mul}x{mi.ir}" for mi in self)

    __str__ = __repr__


# --------------------------------------------------------------------------- spherical harmonics
def spherical_harmonics(l, x, normalize, normalization="integral"):
    """Real SH in e3nn's axis convention, l <= 2 (all the reference uses)."""
    if isinstance(l, int):
        ls = [l]
    elif isinstance(l, (list, tuple)) and not isinstance(l, Irreps) and all(isinstance(a, int) for a in l):
        ls = list(l)
    else:
        ls = [mi.ir.l for mi in Irreps(l) for _ in range(mi.mul)]
    if normalize:
        x = x / torch.clamp(torch.linalg.vector_norm(x, dim=-1, keepdim=True), min=1e-12)
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    blocks = []
    for deg in ls:
        if deg == 0:
            b = torch.ones_like(X).unsqueeze(-1)
        elif deg == 1:
            b = torch.stack([X, Y, Z], -1)
        elif deg == 2:
            s3 = math.sqrt(3.0)
            b = torch.stack([s3 * X * Z, s3 * X * Y, Y * Y - 0.5 * (X * X + Z * Z), s3 * Y * Z,
                             (s3 / 2.0) * (Z * Z - X * X)], -1)
        else:
            raise NotImplementedError("oracle SH restated for l<=2 only")
        if normalization == "component":
            b = b * math.sqrt(2 * deg + 1)
        elif normalization == "integral":
            b = b * math.sqrt((2 * deg + 1) / (4 * math.pi))
        elif normalization != "norm":
            raise ValueError(normalization)
        blocks.append(b)
    return torch.cat(blocks, -1)


# --------------------------------------------------------------------------- wigner 3j
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    C = Fraction((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
                 f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))
    S = Fraction(0)
    for v in range(vmin, vmax + 1):
        S += Fraction((-1) ** (v + j2 + m2) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
                      f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return math.sqrt(float(C)) * float(S)


def _real_to_complex(l):
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    r = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = r
        q[l + m, l - abs(m)] = -1j * r
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * r
        q[l + m, l - abs(m)] = 1j * (-1) ** m * r
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _w3j(l1, l2, l3):
    C = torch.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1), dtype=torch.complex128)
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            m3 = m1 + m2
            if abs(m3) <= l3:
                C[l1 + m1, l2 + m2, l3 + m3] = _su2_cg_coeff(l1, m1, l2, m2, l3, m3)
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    W = torch.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, torch.conj(Q3.T), C)
    assert W.imag.abs().max() < 1e-9
    W = W.real
    return W / W.norm()


def wigner_3j(l1, l2, l3, dtype=torch.float64):
    assert abs(l2 - l3) <= l1 <= l2 + l3
    return _w3j(l1, l2, l3).to(dtype).clone()


# --------------------------------------------------------------------------- tensor products
class _TP(nn.Module):
    """Shared machinery: instruction list -> weighted sum of w3j contractions."""

    def _setup(self, in1, in2, out, instructions):
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(in1), Irreps(in2), Irreps(out)
        self.instructions = instructions  # (i1, i2, io, mode, has_weight)
        numel, self._woff = 0, []
        for (i1, i2, io, mode, hw) in instructions:
            self._woff.append(numel)
            if hw:
                assert mode == "uvw"
                numel += self.irreps_in1[i1].mul * self.irreps_in2[i2].mul * self.irreps_out[io].mul
        self.weight_numel = numel
        # component irrep-normalisation, element path-normalisation, all variances / path weights 1
        self._alpha = []
        for (i1, i2, io, mode, hw) in instructions:
            tot = 0
            for (j1, j2, jo, jm, _) in instructions:
                if jo == io:
                    tot += (self.irreps_in1[j1].mul * self.irreps_in2[j2].mul) if jm == "uvw" else 1
            self._alpha.append(math.sqrt(self.irreps_out[io].ir.dim / tot))

    def _run(self, x1, x2, weight):
        lead = x1.shape[:-1]
        x1 = x1.reshape(-1, x1.shape[-1])
        x2 = x2.reshape(-1, x2.shape[-1])
        if weight is not None:
            weight = weight.reshape(-1, weight.shape[-1])
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        out = x1.new_zeros(x1.shape[0], self.irreps_out.dim)
        for n, (i1, i2, io, mode, hw) in enumerate(self.instructions):
            m1, ir1 = self.irreps_in1[i1]
            m2, ir2 = self.irreps_in2[i2]
            mo, iro = self.irreps_out[io]
            a = x1[:, s1[i1]].reshape(-1, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(-1, m2, ir2.dim)
            w3 = wigner_3j(ir1.l, ir2.l, iro.l, dtype=x1.dtype).to(x1.device)
            if mode == "uvw":
                w = weight[:, self._woff[n]: self._woff[n] + m1 * m2 * mo].reshape(-1, m1, m2, mo)
                r = torch.einsum("zuvw,ijk,zui,zvj->zwk", w, w3, a, b)
            elif mode == "uvuv":
                r = torch.einsum("ijk,zui,zvj->zuvk", w3, a, b)
            else:
                raise NotImplementedError(mode)
            out[:, so[io]] += self._alpha[n] * r.reshape(r.shape[0], -1)
        return out.reshape(*lead, -1)


class FullyConnectedTensorProduct(_TP):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=False, **kw):
        super().__init__()
        assert not shared_weights, "reference always passes per-edge weights (tensor_layers.py:185)"
        in1, in2, out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [(i1, i2, io, "uvw", True)
                 for i1, (_, a) in enumerate(in1) for i2, (_, b) in enumerate(in2)
                 for io, (_, c) in enumerate(out) if c in a * b]
        self._setup(in1, in2, out, instr)

    def forward(self, x, y, weight):
        return self._run(x, y, weight)


class FullTensorProduct(_TP):
    def __init__(self, irreps_in1, irreps_in2, **kw):
        super().__init__()
        in1, in2 = Irreps(irreps_in1), Irreps(irreps_in2)
        outs, instr = [], []
        for i1, (m1, a) in enumerate(in1):
            for i2, (m2, b) in enumerate(in2):
                for c in a * b:
                    instr.append((i1, i2, len(outs), "uvuv", False))
                    outs.append((m1 * m2, c))
        out, perm, _ = Irreps(outs).sort()
        instr = [(i1, i2, perm[io], mode, hw) for (i1, i2, io, mode, hw) in instr]
        self._setup(in1, in2, out, instr)

    def forward(self, x, y):
        return self._run(x, y, None)


# --------------------------------------------------------------------------- batch norm
class BatchNorm(nn.Module):
    def __init__(self, irreps, eps=1e-5, momentum=0.1, affine=True, reduce="mean", instance=False,
                 normalization="component"):
        super().__init__()
        assert reduce == "mean" and not instance and normalization == "component"
        self.irreps = Irreps(irreps)
        self.eps, self.momentum, self.affine = eps, momentum, affine
        ns = sum(mi.mul for mi in self.irreps if mi.ir.is_scalar())
        nf = self.irreps.num_irreps
        self.register_buffer("running_mean", torch.zeros(ns))
        self.register_buffer("running_var", torch.ones(nf))
        if affine:
            self.weight = nn.Parameter(torch.ones(nf))
            self.bias = nn.Parameter(torch.zeros(ns))

    def forward(self, x):
        N = x.shape[0]
        out, ix, irm, irv = [], 0, 0, 0
        new_means, new_vars = [], []
        for mul, ir in self.irreps:
            d = ir.dim
            f = x[:, ix: ix + mul * d].reshape(N, mul, d)
            ix += mul * d
            if ir.is_scalar():
                if self.training:
                    mean = f.mean(0).reshape(mul)
                    new_means.append(mean)
                else:
                    mean = self.running_mean[irm: irm + mul]
                f = f - mean.reshape(1, mul, 1)
            if self.training:
                norm = f.pow(2).mean(2).mean(0)
                new_vars.append(norm)
            else:
                norm = self.running_var[irv: irv + mul]
            scale = (norm + self.eps).pow(-0.5)
            if self.affine:
                scale = scale * self.weight[irv: irv + mul]
            f = f * scale.reshape(1, mul, 1)
            if self.affine and ir.is_scalar():
                f = f + self.bias[irm: irm + mul].reshape(1, mul, 1)
            if ir.is_scalar():
                irm += mul
            irv += mul
            out.append(f.reshape(N, mul * d))
        if self.training:
            with torch.no_grad():
                if new_means:
                    m = torch.cat(new_means)
                    self.running_mean.mul_(1 - self.momentum).add_(self.momentum * m.detach())
                v = torch.cat(new_vars)
                self.running_var.mul_(1 - self.momentum).add_(self.momentum * v.detach())
        return torch.cat(out, -1)
