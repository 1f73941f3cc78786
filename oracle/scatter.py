"""Restatement of torch_scatter==2.0.9 `scatter` / `scatter_mean` [third-party recall].

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference call sites:
models/tensor_layers.py:206, models/score_model.py:390, models/all_atom_score_model.py:445.
"""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    dim_size = int(dim_size)
    shape = (dim_size,) + tuple(src.shape[1:])
    total = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, src)
    if reduce in ("sum", "add"):
        return total
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(
            0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device)).clamp_(min=1)
        return total / cnt.reshape((dim_size,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, reduce="mean")
