"""Restatement of torch_cluster==1.6.0 `radius` / `radius_graph` [third-party recall].

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference call sites:
models/score_model.py:502,568-573,655; models/all_atom_score_model.py:528,593-598,611-612,657.

Canonical rule: a pair is an edge iff both points share a batch id
and dx*dx + dy*dy + dz*dz < r*r in fp32, summed left to right without FMA contraction;
edges come out sorted by (query index, candidate index); a query with more than
`max_num_neighbors` candidates keeps the lowest-index ones (the torch_cluster CUDA rule).
"""
import torch


def _segments(batch, n):
    if batch is None:
        return [(0, n)]
    if batch.numel() == 0:
        return []
    assert bool((batch[1:] >= batch[:-1]).all()), "batch vector must be sorted"
    nb = int(batch.max()) + 1
    ptr = torch.searchsorted(batch.contiguous(), torch.arange(nb + 1, dtype=batch.dtype))
    return [(int(ptr[b]), int(ptr[b + 1])) for b in range(nb)]


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """-> int64 [2, E]; row 0 indexes y (query), row 1 indexes x (candidate)."""
    x = x.float()
    y = y.float()
    r2 = torch.tensor(float(r), dtype=torch.float32) * torch.tensor(float(r), dtype=torch.float32)
    sx, sy = _segments(batch_x, x.shape[0]), _segments(batch_y, y.shape[0])
    rows, cols = [], []
    for b in range(min(len(sx), len(sy))):
        x0, x1 = sx[b]
        y0, y1 = sy[b]
        if x1 == x0 or y1 == y0:
            continue
        xs, ys = x[x0:x1], y[y0:y1]
        dx = ys[:, None, 0] - xs[None, :, 0]
        dy = ys[:, None, 1] - xs[None, :, 1]
        dz = ys[:, None, 2] - xs[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        hit = d2 < r2
        if max_num_neighbors is not None:
            rank = torch.cumsum(hit.to(torch.int64), dim=1)
            hit = hit & (rank <= max_num_neighbors)
        qi, ci = torch.nonzero(hit, as_tuple=True)  # row-major: sorted by (query, candidate)
        rows.append(qi + y0)
        cols.append(ci + x0)
    if not rows:
        return torch.zeros((2, 0), dtype=torch.int64)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target"):
    """-> int64 [2, E]; row 0 = neighbour, row 1 = centre (grouped by centre)."""
    e = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    if flow == "source_to_target":
        row, col = e[1], e[0]
    else:
        row, col = e[0], e[1]
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col])
