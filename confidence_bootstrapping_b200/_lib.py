"""ctypes binding of libcb200.so (the C-ABI declared in include/cb200.h).

PyTorch is only plumbing here: it owns device memory and the stream; every wrapper checks
dtype / contiguity / device, hands raw pointers to the library and raises RuntimeError with
`cb_last_error()` on failure (the reference's callers catch Exception and halve the batch:
finetune_train.py:187-195).  There is deliberately NO fallback: if the library is missing or the
tensors are not on a CUDA device the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

_i32p = C.c_void_p
_f32p = C.c_void_p


class EdgeFeatArgs(C.Structure):
    _fields_ = [
        ("row", C.c_void_p), ("col", C.c_void_p), ("n_edges_dev", C.c_void_p), ("e_cap", C.c_int32),
        ("pos_agg", C.c_void_p), ("pos_nbr", C.c_void_p), ("sh_sign", C.c_float), ("lmax", C.c_int32),
        ("agg_graph", C.c_void_p), ("b1_graph", C.c_void_p), ("b1_graph_stride", C.c_int32),
        ("extra", C.c_void_p), ("n_extra", C.c_int32),
        ("smear_offset", C.c_void_p), ("smear_coeff", C.c_float), ("n_gauss", C.c_int32),
        ("W1", C.c_void_p), ("ldw1", C.c_int32), ("extra_off", C.c_int32), ("smear_off", C.c_int32),
        ("W2", C.c_void_p), ("b2", C.c_void_p), ("ns", C.c_int32),
        ("out_attr", C.c_void_p), ("out_sh", C.c_void_p),
    ]


class TpSegment(C.Structure):
    _fields_ = [
        ("rowptr", C.c_void_p), ("col", C.c_void_p), ("e_attr", C.c_void_p), ("e_post", C.c_void_p),
        ("sh", C.c_void_p), ("P_agg", C.c_void_p), ("P_nbr", C.c_void_p),
        ("ldp_agg", C.c_int32), ("ldp_nbr", C.c_int32),
        ("W1e", C.c_void_p), ("ldw1", C.c_int32),
        ("b1", C.c_void_p), ("W2a", C.c_void_p),
        ("n0", C.c_int32), ("n1", C.c_int32), ("col_off", C.c_int32), ("slot", C.c_int32),
        ("gate_rowptr", C.c_void_p), ("gate_mask", C.c_void_p), ("W2t", C.c_void_p),
    ]


CB_MAX_SEGS = 12


class TpConvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("d_in", C.c_int32), ("d_out", C.c_int32), ("S", C.c_int32), ("ne", C.c_int32),
        ("H", C.c_int32), ("n_out", C.c_int32), ("agg_graph", C.c_void_p),
        ("rows", C.c_void_p), ("n_rows", C.c_int32), ("terms", C.c_void_p), ("n_terms", C.c_int32),
        ("runs", C.c_void_p), ("n_runs", C.c_int32), ("n_segs", C.c_int32),
        ("segs", TpSegment * CB_MAX_SEGS),
        ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("residual", C.c_void_p),
        ("d_res", C.c_int32), ("ld_res", C.c_int32), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_floats", C.c_int64), ("node_begin", C.c_int32), ("node_end", C.c_int32),
        ("accum_mode", C.c_int32), ("flags", C.c_int32),
        ("pre_sum", C.c_void_p), ("pre_deg", C.c_void_p), ("pre_n0", C.c_int32), ("pre_n1", C.c_int32),
        ("pre_period", C.c_int32), ("pad_", C.c_int32),
        ("chains", C.c_void_p), ("n_chains", C.c_int32), ("blocks", C.c_void_p), ("n_blocks", C.c_int32),
        ("kp", C.c_int32), ("max_chain_bytes", C.c_int32),
    ]


CB_TP_RAW_SUM = 1


class SdeStepArgs(C.Structure):
    _fields_ = [
        ("pos", C.c_void_p), ("B", C.c_int32), ("N", C.c_int32), ("R", C.c_int32),
        ("bond_uv", C.c_void_p), ("mask_rotate", C.c_void_p),
        ("tr_score", C.c_void_p), ("rot_score", C.c_void_p), ("tor_score", C.c_void_p),
        ("z_tr", C.c_void_p), ("z_rot", C.c_void_p), ("z_tor", C.c_void_p),
        ("c_tr_score", C.c_float), ("c_tr_noise", C.c_float), ("c_rot_score", C.c_float),
        ("c_rot_noise", C.c_float), ("c_tor_score", C.c_float), ("c_tor_noise", C.c_float),
        ("coeffs_dev", C.c_void_p),
    ]


EXPORTS = [
    "cb_last_error", "cb_version", "cb_sizeof", "cb_radius_count", "cb_radius_fill", "cb_radius_count_t", "cb_radius_fill_t",
    "cb_exclusive_scan_i32", "cb_edge_featurize", "cb_tp_conv_forward", "cb_tp_conv_items", "cb_sde_step",
]

_lib = None
launch_count = 0  # number of cb200 kernels-launching calls issued (bench.py reports it)


def library_path() -> str:
    return _build.LIB


def lib():
    """Load (never build implicitly on a GPU box: the .so travels with the tree)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the cb200 kernels)")
        l = C.CDLL(path)
        l.cb_last_error.restype = C.c_char_p
        for name in EXPORTS[1:]:
            getattr(l, name).restype = C.c_int
        l.cb_tp_conv_items.restype = C.c_int64
        _lib = l
    return _lib


def _check(code: int, what: str):
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {lib().cb_last_error().decode()}")


def _ptr(t, dtype, name, allow_none=False):
    try:                                            # fast path: the common case is a valid tensor
        if t.dtype is dtype and t.is_cuda and t.is_contiguous():
            return t.data_ptr()
    except AttributeError:
        pass
    if t is None:
        if allow_none:
            return None
        raise RuntimeError(f"{name}: tensor required")
    if not torch.is_tensor(t):
        raise RuntimeError(f"{name}: expected a tensor, got {type(t)}")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the cb200 kernels have no CPU path)")
    raise RuntimeError(f"{name}: expected a contiguous tensor")


def f32(t, name, allow_none=False):
    return _ptr(t, torch.float32, name, allow_none)


def i32(t, name, allow_none=False):
    return _ptr(t, torch.int32, name, allow_none)


def u8(t, name, allow_none=False):
    return _ptr(t, torch.uint8, name, allow_none)


def stream_ptr():
    """Raw handle of torch's current stream on the current device (every cb200 kernel is enqueued on it)."""
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _launched(n=1):
    global launch_count
    launch_count += n


# ------------------------------------------------------------------------------------------- K1
def exclusive_scan(counts: torch.Tensor, out: torch.Tensor, scratch: torch.Tensor):
    """out[0]=0, out[i+1]=sum(counts[:i+1]); out has counts.numel()+1 int32 entries."""
    n = counts.numel()
    if out.numel() < n + 1 or scratch.numel() < 4096:
        raise RuntimeError("exclusive_scan: output/scratch too small")
    _check(lib().cb_exclusive_scan_i32(C.c_void_p(i32(counts, "counts")), C.c_void_p(i32(out, "out")),
                                       C.c_int32(n), C.c_void_p(i32(scratch, "scratch")), stream_ptr()),
           "cb_exclusive_scan_i32")
    _launched(3)


def radius_count(x, x_ptr, y, y_batch, cutoff, r, max_neighbors, exclude_self, count):
    _check(lib().cb_radius_count(C.c_void_p(f32(x, "x")), C.c_void_p(i32(x_ptr, "x_ptr")), C.c_void_p(f32(y, "y")),
                                 C.c_void_p(i32(y_batch, "y_batch")), C.c_void_p(f32(cutoff, "cutoff", True)),
                                 C.c_float(r), C.c_int32(y.shape[0]), C.c_int32(max_neighbors),
                                 C.c_int32(int(exclude_self)), C.c_void_p(i32(count, "count")), stream_ptr()),
           "cb_radius_count")
    _launched()


def radius_fill(x, x_ptr, y, y_batch, cutoff, r, max_neighbors, exclude_self, rowptr, row, col):
    _check(lib().cb_radius_fill(C.c_void_p(f32(x, "x")), C.c_void_p(i32(x_ptr, "x_ptr")), C.c_void_p(f32(y, "y")),
                                C.c_void_p(i32(y_batch, "y_batch")), C.c_void_p(f32(cutoff, "cutoff", True)),
                                C.c_float(r), C.c_int32(y.shape[0]), C.c_int32(max_neighbors),
                                C.c_int32(int(exclude_self)), C.c_void_p(i32(rowptr, "rowptr")),
                                C.c_void_p(i32(row, "row")), C.c_void_p(i32(col, "col")), stream_ptr()),
           "cb_radius_fill")
    _launched()


def radius_count_t(x, x_batch, y, y_ptr, cutoff, r, exclude_self, kept_rowptr, kept_col, count):
    _check(lib().cb_radius_count_t(C.c_void_p(f32(x, "x")), C.c_void_p(i32(x_batch, "x_batch")),
                                   C.c_void_p(f32(y, "y")), C.c_void_p(i32(y_ptr, "y_ptr")),
                                   C.c_void_p(f32(cutoff, "cutoff", True)), C.c_float(r), C.c_int32(x.shape[0]),
                                   C.c_int32(int(exclude_self)), C.c_void_p(i32(kept_rowptr, "kept_rowptr", True)),
                                   C.c_void_p(i32(kept_col, "kept_col", True)), C.c_void_p(i32(count, "count")),
                                   stream_ptr()), "cb_radius_count_t")
    _launched()


def radius_fill_t(x, x_batch, y, y_ptr, cutoff, r, exclude_self, kept_rowptr, kept_col, rowptr_t, row_t, col_t):
    _check(lib().cb_radius_fill_t(C.c_void_p(f32(x, "x")), C.c_void_p(i32(x_batch, "x_batch")),
                                  C.c_void_p(f32(y, "y")), C.c_void_p(i32(y_ptr, "y_ptr")),
                                  C.c_void_p(f32(cutoff, "cutoff", True)), C.c_float(r), C.c_int32(x.shape[0]),
                                  C.c_int32(int(exclude_self)), C.c_void_p(i32(kept_rowptr, "kept_rowptr", True)),
                                  C.c_void_p(i32(kept_col, "kept_col", True)), C.c_void_p(i32(rowptr_t, "rowptr_t")),
                                  C.c_void_p(i32(row_t, "row_t")), C.c_void_p(i32(col_t, "col_t")), stream_ptr()),
           "cb_radius_fill_t")
    _launched()


# ------------------------------------------------------------------------------------------- K2..K4
def edge_featurize(args: EdgeFeatArgs):
    _check(lib().cb_edge_featurize(C.byref(args), stream_ptr()), "cb_edge_featurize")
    _launched()


tp_conv_hook = None  # optional callable(args) -> context manager; bench.py times K3 launches with CUDA events


def tp_conv_items(args: TpConvArgs) -> int:
    return int(lib().cb_tp_conv_items(C.byref(args)))


def tp_conv_forward(args: TpConvArgs):
    if tp_conv_hook is not None:
        with tp_conv_hook(args):
            _check(lib().cb_tp_conv_forward(C.byref(args), stream_ptr()), "cb_tp_conv_forward")
    else:
        _check(lib().cb_tp_conv_forward(C.byref(args), stream_ptr()), "cb_tp_conv_forward")
    _launched(2)


def sde_step(args: SdeStepArgs):
    _check(lib().cb_sde_step(C.byref(args), stream_ptr()), "cb_sde_step")
    _launched()
