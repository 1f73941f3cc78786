"""CPU restatement of the reference sampler.  TEST INFRASTRUCTURE (see oracle/__init__.py).

  utils/geometry.py:39-86          axis_angle_to_matrix                 -> axis_angle_to_matrix
  utils/geometry.py:246-276        rigid_transform_Kabsch_3D_torch_batch -> kabsch_batch
  utils/torsion.py:75-90           modify_conformer_torsion_angles_batch -> twist_batch
  utils/diffusion_utils.py:60-78   modify_conformer_batch               -> modify_conformer_batch
  utils/diffusion_utils.py:28-32,138-143,150-161  t_to_sigma / get_t_schedule / set_time
  utils/sampling.py:59-274         sampling (default branch + ode / temperature / no_random variants)
Pinned against the real reference executed under oracle/shims.py (tests/test_golden_oracle.py).
"""
from __future__ import annotations

import numpy as np
import torch

from confidence_bootstrapping_b200.data import Batch, DataLoader


def t_to_sigma(t_tr, t_rot, t_tor, args):
    f = lambda t, lo, hi: lo ** (1 - t) * hi ** t
    return (f(t_tr, args.tr_sigma_min, args.tr_sigma_max), f(t_rot, args.rot_sigma_min, args.rot_sigma_max),
            f(t_tor, args.tor_sigma_min, args.tor_sigma_max))


def get_t_schedule(inference_steps, alpha=1, beta_=1, t_max=1):
    from scipy.stats import beta
    c = np.linspace(beta.cdf(t_max, a=alpha, b=beta_), 0, inference_steps + 1)[:-1]
    return beta.ppf(c, a=alpha, b=beta_)


def set_time(batch, t_tr, t_rot, t_tor, b, all_atoms=False):
    kinds = ["ligand", "receptor"] + (["atom"] if all_atoms else [])
    for k in kinds:
        n = batch[k].num_nodes
        batch[k].node_t = {"tr": t_tr * torch.ones(n), "rot": t_rot * torch.ones(n), "tor": t_tor * torch.ones(n)}
    batch.complex_t = {"tr": t_tr * torch.ones(b), "rot": t_rot * torch.ones(b), "tor": t_tor * torch.ones(b)}


def axis_angle_to_matrix(aa):
    ang = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    k = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    q = torch.cat([torch.cos(half), aa * k], dim=-1)
    r, i, j, kk = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - s * (j * j + kk * kk), s * (i * j - kk * r), s * (i * kk + j * r),
                     s * (i * j + kk * r), 1 - s * (i * i + kk * kk), s * (j * kk - i * r),
                     s * (i * kk - j * r), s * (j * kk + i * r), 1 - s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def kabsch_batch(A, B):
    """Rotation R [b,3,3] and translation t [b,3,1] taking A [b,N,3] onto B."""
    A, B = A.permute(0, 2, 1), B.permute(0, 2, 1)
    ca, cb = A.mean(dim=2, keepdim=True), B.mean(dim=2, keepdim=True)
    H = torch.bmm(A - ca, (B - cb).transpose(1, 2))
    U, S, Vt = torch.linalg.svd(H)
    R = torch.bmm(Vt.transpose(1, 2), U.transpose(1, 2))
    flip = torch.diag(torch.tensor([1.0, 1.0, -1.0]))
    Rm = torch.bmm(Vt.transpose(1, 2) @ flip, U.transpose(1, 2))
    R = torch.where(torch.linalg.det(R)[:, None, None] < 0, Rm, R)
    return R, torch.bmm(-R, ca) + cb


def twist_batch(pos, bonds, mask_rotate, updates):
    pos = pos + 0
    for k, e in enumerate(bonds):
        u, v = int(e[0]), int(e[1])
        assert not mask_rotate[k, u] and mask_rotate[k, v]
        axis = pos[:, u] - pos[:, v]
        rot = axis_angle_to_matrix(axis / torch.linalg.norm(axis, dim=-1, keepdims=True) * updates[:, k:k + 1])
        m = mask_rotate[k]
        pos[:, m] = torch.bmm(pos[:, m] - pos[:, v:v + 1], rot.transpose(1, 2)) + pos[:, v:v + 1]
    return pos


def modify_conformer_batch(orig_pos, batch, tr, rot, tor, mask_rotate):
    B = batch.num_graphs
    N, M = batch["ligand"].num_nodes // B, batch["ligand", "ligand"].num_edges // B
    pos = orig_pos.reshape(B, N, 3) + 0
    ei, em = batch["ligand", "ligand"].edge_index[:, :M], batch["ligand"].edge_mask[:M]
    c = pos.mean(dim=1, keepdim=True)
    rigid = torch.bmm(pos - c, axis_angle_to_matrix(rot).permute(0, 2, 1)) + tr.unsqueeze(1) + c
    if tor is None:
        return rigid.reshape(-1, 3)
    flex = twist_batch(rigid, ei.T[em], mask_rotate, tor.reshape(B, -1))
    R, t = kabsch_batch(flex, rigid)
    return (torch.bmm(flex, R.transpose(1, 2)) + t.transpose(1, 2)).reshape(-1, 3)


def sampling(data_list, forward, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_to_sigma_fn, model_args,
             no_random=False, ode=False, batch_size=32, no_final_step_noise=False, temp_sampling=1.0, temp_psi=0.0,
             temp_sigma_data=0.5, confidence_forward=None, filtering_data_list=None, filtering_model_args=None,
             crop_fn=None):
    """`forward(batch) -> (tr, rot, tor, ...)` is the score model (oracle.model.cg_forward bound to a
    state_dict, or the real reference module)."""
    N = len(data_list)
    mr = data_list[0]["ligand"].mask_rotate
    mask_rotate = torch.from_numpy(mr[0] if isinstance(mr, (list, tuple)) else mr)
    ts = [temp_sampling] * 3 if not hasattr(temp_sampling, "__iter__") else list(temp_sampling)
    tp = [temp_psi] * 3 if not hasattr(temp_psi, "__iter__") else list(temp_psi)
    confidence = [] if confidence_forward is not None else None
    f_loader = iter(DataLoader(filtering_data_list, batch_size=batch_size)) if filtering_data_list is not None else None
    all_atoms = "all_atoms" in model_args and model_args.all_atoms
    with torch.no_grad():
        for bid, batch in enumerate(DataLoader(data_list, batch_size=batch_size)):
            b = batch.num_graphs
            for k in range(inference_steps):
                last = k == inference_steps - 1
                t = (tr_schedule[k], rot_schedule[k], tor_schedule[k])
                dt = [s[k] - s[k + 1] if not last else s[k] for s in (tr_schedule, rot_schedule, tor_schedule)]
                sig = t_to_sigma_fn(*t)
                set_time(batch, t[0], t[1], t[2], b, all_atoms)
                score = forward(batch)[:3]
                pert = []
                for c, name in enumerate(("tr", "rot", "tor")):
                    if c == 2 and model_args.no_torsion:
                        pert.append(None)
                        continue
                    smax, smin = getattr(model_args, f"{name}_sigma_max"), getattr(model_args, f"{name}_sigma_min")
                    g = sig[c] * torch.sqrt(torch.tensor(2 * np.log(smax / smin)))
                    if ode:
                        pert.append(0.5 * g ** 2 * dt[c] * score[c])
                        continue
                    shape = (min(batch_size, N), 3) if c < 2 else score[c].shape
                    z = torch.zeros(shape) if no_random or (no_final_step_noise and last) else \
                        torch.normal(mean=0, std=1, size=shape)
                    p = g ** 2 * dt[c] * score[c] + g * np.sqrt(dt[c]) * z
                    if ts[c] != 1.0:
                        sd_ = np.exp(temp_sigma_data * np.log(smax) + (1 - temp_sigma_data) * np.log(smin))
                        lam = (sd_ + sig[c]) / (sd_ + sig[c] / ts[c])
                        p = g ** 2 * dt[c] * (lam + ts[c] * tp[c] / 2) * score[c] + g * np.sqrt(dt[c] * (1 + tp[c])) * z
                    pert.append(p)
                batch["ligand"].pos = modify_conformer_batch(batch["ligand"].pos, batch, pert[0], pert[1], pert[2], mask_rotate)
            n = len(batch["ligand"].pos) // b
            for i in range(b):
                data_list[bid * batch_size + i]["ligand"].pos = batch["ligand"].pos[i * n:n * (i + 1)]
            if confidence_forward is not None:
                if f_loader is not None:
                    fb = next(f_loader)
                    fb["ligand"].pos = batch["ligand"].pos
                    if getattr(filtering_model_args, "crop_beyond", None) is not None:
                        graphs = fb.to_data_list()
                        for gph in graphs:
                            crop_fn(gph, filtering_model_args.crop_beyond, filtering_model_args.all_atoms)
                        fb = Batch.from_data_list(graphs)
                    set_time(fb, 0, 0, 0, b, filtering_model_args.all_atoms)
                    out = confidence_forward(fb)
                else:
                    out = confidence_forward(batch)
                confidence.append(out[0] if isinstance(out, tuple) else out)
    if confidence is not None:
        confidence = torch.nan_to_num(torch.cat(confidence, dim=0), nan=-1000)
    return data_list, confidence


def crop_beyond(graph, cutoff, all_atoms):
    """utils/utils.py:395-420 on a single graph."""
    lp, rp = graph["ligand"].pos, graph["receptor"].pos
    keep = torch.any(torch.sum((lp.unsqueeze(0) - rp.unsqueeze(1)) ** 2, -1) < cutoff ** 2, dim=1)
    remap = torch.cumsum(keep.long(), dim=0) - 1
    rec, rr = graph["receptor"], graph["receptor", "receptor"]
    if all_atoms:
        a2r = graph["atom", "receptor"].edge_index[1]
        akeep = keep[a2r]
        new_a2r = remap[a2r][akeep]
    for k in ("pos", "x", "side_chain_vecs"):
        setattr(rec, k, getattr(rec, k)[keep])
    ei = rr.edge_index
    rr.edge_index = remap[ei[:, keep[ei[0]] & keep[ei[1]]]]
    if all_atoms:
        atom, aa = graph["atom"], graph["atom", "atom"]
        amap = torch.cumsum(akeep.long(), dim=0) - 1
        atom.x, atom.pos = atom.x[akeep], atom.pos[akeep]
        ei = aa.edge_index
        aa.edge_index = amap[ei[:, akeep[ei[0]] & akeep[ei[1]]]]
        graph["atom", "receptor"].edge_index = torch.stack([torch.arange(len(new_a2r)), new_a2r])
    return graph
