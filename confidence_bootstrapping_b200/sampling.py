"""Pose initialisation and the reverse-diffusion sampler (utils/sampling.py) on the cb200 kernels.

`sampling()` keeps the reference signature and contract (sampling.py:59-274): it mutates
`data_list[i]['ligand'].pos` (device tensors), returns `(data_list, confidence)` with NaN -> -1000,
asserts on the flags the reference asserts on, and lets exceptions propagate (callers halve the
batch).  Per step the score-model forward and the pose update are cb200 kernels; the perturbation
arithmetic of sampling.py:119-141 is folded into the K4 launch as scalar coefficients.
"""
from __future__ import annotations

import collections
import contextlib
import copy
import os
import warnings
import weakref

import numpy as np
import torch

from .data import Batch, DataLoader
from .diffusion_utils import LigandTopology, check_rotation_masks, sde_step, set_time
from .utils import crop_beyond


def _is_iterable(x):
    try:
        iter(x)
        return True
    except TypeError:
        return False


def _mask_rotate_of(graph):
    mr = graph["ligand"].mask_rotate
    return mr[0] if isinstance(mr, (list, tuple)) else mr


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, pocket_knowledge=False, pocket_cutoff=7):
    """In-place pose initialisation (sampling.py:15-48): uniform torsions, random rotation about the
    ligand centroid, placement on the pocket centre, N(0, tr_sigma_max) translation.  Host-side numpy / scipy
    with the reference's random draws in the reference's order; the geometry is applied per ligand topology for
    all its samples at once (12 ms instead of 75 ms for 40 samples of a 40-atom ligand)."""
    from scipy.spatial.transform import Rotation as R
    center_pocket = data_list[0]["receptor"].pos.mean(dim=0)
    if pocket_knowledge:
        c = data_list[0]
        d = torch.cdist(c["receptor"].pos, torch.from_numpy(c["ligand"].orig_pos[0]).float() - c.original_center)
        label = torch.any(d < pocket_cutoff, dim=1)
        if torch.any(label):
            center_pocket = c["receptor"].pos[label].mean(dim=0)
        else:
            center_pocket = c["receptor"].pos[torch.argmin(torch.min(d, dim=1)[0])]
    # Random numbers are drawn per sample in the reference's order (all torsion draws first, then per sample
    # R.random() and the translation); the geometry is applied to all samples that share a ligand topology at once
    # instead of sample by sample (the per-sample host loop costs more than the whole sampling).
    if not no_torsion:
        draws = []
        for g in data_list:
            n_tor = int(g["ligand"].edge_mask.sum())
            draws.append(np.random.uniform(low=-np.pi, high=np.pi, size=n_tor))
        for idx in _same_topology_groups(data_list):
            g0 = data_list[idx[0]]
            bonds = g0["ligand", "ligand"].edge_index.T[g0["ligand"].edge_mask]
            if len(idx) == 1:
                g0["ligand"].pos = _twist_numpy(g0["ligand"].pos, bonds, _mask_rotate_of(g0), draws[idx[0]])
                continue
            pos = np.stack([data_list[i]["ligand"].pos.detach().cpu().numpy() for i in idx]).astype(np.float32)
            new = _twist_numpy_batched(pos, bonds.cpu().numpy(), np.asarray(_mask_rotate_of(g0)), np.stack([draws[i] for i in idx]))
            for k, i in enumerate(idx):
                data_list[i]["ligand"].pos = torch.from_numpy(new[k])
    rots, trs = [], []
    for g in data_list:
        rots.append(torch.from_numpy(R.random().as_matrix()).float())
        if not no_random:
            trs.append(torch.normal(mean=0, std=tr_sigma_max, size=(1, 3)))
    for idx in _same_topology_groups(data_list):
        pos = torch.stack([data_list[i]["ligand"].pos for i in idx])                 # [n, atoms, 3]
        rot = torch.stack([rots[i] for i in idx])
        out = torch.bmm(pos - pos.mean(dim=1, keepdim=True), rot.transpose(1, 2)) + center_pocket
        if not no_random:
            out = out + torch.stack([trs[i] for i in idx])
        for k, i in enumerate(idx):
            data_list[i]["ligand"].pos = out[k]


def _same_topology_groups(data_list):
    """Indices of data_list grouped by ligand topology (atom count, bond list, rotation masks), in first-seen order."""
    groups, keys = [], {}
    for i, g in enumerate(data_list):
        lig = g["ligand"]
        mr = np.asarray(_mask_rotate_of(g))
        key = (tuple(lig.pos.shape), g["ligand", "ligand"].edge_index.cpu().numpy().tobytes(),
               lig.edge_mask.cpu().numpy().tobytes(), mr.shape, mr.tobytes())
        if key not in keys:
            keys[key] = len(groups)
            groups.append([])
        groups[keys[key]].append(i)
    return groups


def _twist_numpy_batched(pos, bonds, mask_rotate, updates):
    """modify_conformer_torsion_angles (utils/torsion.py:48-72) for a stack of conformers of ONE topology:
    pos [n, atoms, 3] float32, updates [n, n_tor].  Same arithmetic as the per-sample loop (float64 rotation applied to
    the float32 coordinates, bond after bond); only the loop over samples is vectorised."""
    from scipy.spatial.transform import Rotation as R
    p = pos.copy()
    for k, e in enumerate(bonds):
        u, v = int(e[0]), int(e[1])
        axis = p[:, u] - p[:, v]                                           # float32 [n, 3]
        rotvec = axis * updates[:, k:k + 1] / np.linalg.norm(axis, axis=1, keepdims=True)
        rot = R.from_rotvec(rotvec).as_matrix()                            # float64 [n, 3, 3]
        m = mask_rotate[k]
        moved = np.einsum("nai,nji->naj", p[:, m] - p[:, v:v + 1], rot) + p[:, v:v + 1]
        zero = updates[:, k] == 0                                          # the reference skips zero updates
        p[:, m] = np.where(zero[:, None, None], p[:, m], moved.astype(np.float32))
    return p


def _twist_numpy(pos, bonds, mask_rotate, updates):
    """modify_conformer_torsion_angles (utils/torsion.py:48-72): sequential bond rotations on the host."""
    from scipy.spatial.transform import Rotation as R
    p = pos.detach().cpu().numpy().astype(np.float64).copy() if torch.is_tensor(pos) else np.array(pos, dtype=np.float64)
    p = p.astype(np.float32) if torch.is_tensor(pos) else p
    for k, e in enumerate(bonds.cpu().numpy()):
        if updates[k] == 0:
            continue
        u, v = int(e[0]), int(e[1])
        axis = p[u] - p[v]
        axis = axis * updates[k] / np.linalg.norm(axis)
        rot = R.from_rotvec(axis).as_matrix()
        p[mask_rotate[k]] = (p[mask_rotate[k]] - p[v]) @ rot.T + p[v]
    return torch.from_numpy(p.astype(np.float32))


_warned_train_mode = False


@contextlib.contextmanager
def _inference_mode(*models):
    """Scoped eval() around a sampling call, plus a refresh of the derived-weight caches.

    CONSCIOUS DIVERGENCE: finetune_train.py:177 samples with the score model left in
    train mode (dropout 0.1 active, e3nn BatchNorm on batch statistics) because nothing calls .eval() on that path;
    inference.py:309, dock.py:101 and bootstrapping.py:106 do.  The cb200 kernels implement the inference arithmetic
    (running-stat BatchNorm folded into the K3 epilogue, no dropout), so sampling() runs the models in eval mode and
    restores their previous mode afterwards -- the caller's train_epoch (utils/training.py:185) sets .train() itself.
    A one-time warning says so.  Without this the callers' `except Exception` retry loop (finetune_train.py:187-195)
    would halve the batch five times and skip every complex."""
    global _warned_train_mode
    prev = []
    for m in models:
        if m is None or not isinstance(m, torch.nn.Module):
            continue
        prev.append((m, m.training))
        if m.training:
            if not _warned_train_mode:
                warnings.warn("cb200 sampling(): the model is in train mode (the reference's finetune_train.py samples that way: "
                              "dropout on, batch-statistic BatchNorm); cb200 samples in eval mode and restores train mode afterwards")
                _warned_train_mode = True
            m.eval()
        # folded weights (W2a, projections, BatchNorm affine) must follow in-place `.data` writes, which do not bump tensor
        # versions (ExponentialMovingAverage.copy_to / restore between sampling calls): a per-tensor (L1, L2) signature of
        # every parameter and buffer is compared with the one the caches were built under, and they are dropped when it moved
        sig = _weights_signature(m)
        old = getattr(m, "_cb200_weights_sig", None)
        if old is None or old[0] != sig[0] or not torch.equal(old[1], sig[1]):
            for sub in m.modules():
                inv = getattr(sub, "invalidate_caches", None)
                if inv is not None:
                    inv()
            object.__setattr__(m, "_cb200_weights_sig", sig)
    try:
        yield
    finally:
        for m, was_training in prev:
            if was_training:
                m.train(True)


def _weights_signature(model):
    """(data pointers, 3 float64 checksums of ALL parameters and floating buffers): the flattened weights projected on two fixed
    pseudo-random vectors (sensitive to any element changing, sign flips and permutations included) and their L1 norm --
    two concatenation launches, a few reductions and one small D2H per sampling call."""
    ts = [p.detach() for p in model.parameters()] + [b.detach() for b in model.buffers() if b.is_floating_point()]
    ts = [t for t in ts if t.numel() > 0]
    if not ts:
        return (), torch.zeros(0, dtype=torch.float64)
    ptrs = tuple(t.data_ptr() for t in ts)
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in ts]).double()
    proj = getattr(model, "_cb200_sig_proj", None)
    if proj is None or proj.shape[1] != flat.numel() or proj.device != flat.device:
        g = torch.Generator(device="cpu").manual_seed(0x5EED)
        proj = torch.rand((2, flat.numel()), generator=g, dtype=torch.float64).add_(0.5).to(flat.device)
        object.__setattr__(model, "_cb200_sig_proj", proj)
    sig = torch.cat([proj @ flat, flat.abs().sum().reshape(1)]).cpu()
    return ptrs, sig


USE_CUDA_GRAPH = os.environ.get("CB200_CUDA_GRAPH", "1") != "0"
# Step graphs of recent batches, keyed on (complex fingerprint, model, weights epoch, batch / schedule parameters): a later batch
# of the SAME complex (inference.py samples 40 poses in batches of 10; finetune_train.py re-samples every complex each epoch)
# copies its poses into the captured batch and replays all steps -- no eager first step, no capture.
# At most ONE step graph is alive per process: a miss drops the cached graph BEFORE the new batch's eager step, so its private
# pool (the K3 workspaces: up to 16 GiB per layer for 1000-residue complexes) is free for the new capture.  (Two cached graphs
# drove a mixed-size run -- BASELINE configs[3], N_r up to 1000 -- out of memory: 49 GiB in graph pools + fragmentation.)
GRAPH_CACHE_SIZE = int(os.environ.get("CB200_GRAPH_CACHE", "1"))
# Once a model has run one eager step under its current weights, later batches capture their step graph BEFORE step 0 (static
# tables built by model._static, no eager forward per batch).  Bit-identical (tests); in-process A/B of end-to-end calls on
# new complexes: 194.2 -> 189.5 ms (profiles/ab_e2e.py).
CAPTURE_FIRST_STEP = os.environ.get("CB200_CAPTURE_FIRST_STEP", "1") != "0"
_graph_cache = collections.OrderedDict()
graph_cache_hits = 0
_graph_warned = False
_graph_pools = {}      # device index -> memory-pool handle shared by every step graph captured on that device


_capture_streams = {}


def _capture_stream(device):
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _capture_streams:
        _capture_streams[idx] = torch.cuda.Stream(device=idx)
    return _capture_streams[idx]


def _graph_pool(device):
    """One private memory pool per device for all step graphs: a graph's intermediates (K3 workspaces are GBs) are carved
    from it during capture and return to it when the graph is dropped at the end of the sampling call, so the next call's
    capture re-uses them instead of paying cudaMalloc / cudaFree of several GB per call (measured: +290 ms per call)."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _graph_pools:
        # the allocator drops a pool when the last graph captured into it dies: a tiny keeper graph pins it for the process
        pool = torch.cuda.graph_pool_handle()
        keeper = torch.cuda.CUDAGraph()
        side = _capture_stream(device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            keeper.capture_begin(pool=pool)
            anchor = torch.zeros(8, device=torch.device("cuda", idx))
            keeper.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        _graph_pools[idx] = (pool, keeper, anchor)
    return _graph_pools[idx][0]


_conf_streams = {}
PIPELINE_CONFIDENCE = os.environ.get("CB200_PIPELINE_CONFIDENCE", "1") != "0"


_copy_streams = {}


def _copy_stream(device):
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _copy_streams:
        _copy_streams[idx] = torch.cuda.Stream(device=idx)
    return _copy_streams[idx]


def _conf_stream(device):
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _conf_streams:
        _conf_streams[idx] = torch.cuda.Stream(device=idx)
    return _conf_streams[idx]


class FilteringLeg:
    """The confidence (filtering) leg of batch i runs on a SECOND stream while the reverse-diffusion steps of batch i+1 run
    on the main one.  The leg has host reads (crop_beyond changes the shapes, so its static tables are rebuilt per batch);
    on one stream they would wait for everything already enqueued -- the next batch's 20 graph replays -- and the GPU
    would idle while the host prepares.  On its own stream the leg only waits for the event recorded after ITS batch's last
    step, its kernels fill SMs the replays leave idle, and a host stall during the leg is hidden behind ~150 ms of queued
    replays.  The arithmetic is unchanged (same kernels, same order within the leg).

        leg = FilteringLeg(device)
        for batch: pos = reverse_diffusion(...); leg.submit(fn, pos)     # fn(pos) -> confidence tensor; runs the PREVIOUS one
        confidences = leg.finish()                                       # runs the last one, joins the streams
    """

    def __init__(self, device, enabled=None):
        self.device = device
        self.enabled = (PIPELINE_CONFIDENCE if enabled is None else enabled) and torch.device(device).type == "cuda"
        self.pending, self.results, self._keep = None, [], []

    def _run(self, fn, pos, ev):
        if not self.enabled:
            self.results.append(fn(pos))
            return
        side = _conf_stream(self.device)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            out = fn(pos)
        pos.record_stream(side)
        # whatever the leg read that was allocated on the main stream (the collated filtering batch in fn's closure) must not
        # return to the allocator before the streams are joined
        # (a leg's host reads complete only after every earlier leg on the stream has: two entries are enough)
        self._keep.append(fn)
        del self._keep[:-2]
        self.results.append(out)

    def submit(self, fn, pos):
        ev = None
        if self.enabled:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        prev, self.pending = self.pending, (fn, pos, ev)
        if prev is not None:
            self._run(*prev)

    def finish(self):
        if self.pending is not None:
            self._run(*self.pending)
            self.pending = None
        if self.enabled:
            main, side = torch.cuda.current_stream(self.device), _conf_stream(self.device)
            main.wait_stream(side)
            for r in self.results:
                for t in (r if isinstance(r, (tuple, list)) else (r,)):
                    if torch.is_tensor(t):
                        t.record_stream(main)
        out, self.results, self._keep = self.results, [], []
        return out


def _step_scalars(t_idx, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_to_sigma, model_args, g_const, no_random, ode,
                  no_final_step_noise, temp_sampling, temp_psi, temp_sigma_data, no_torsion):
    """Host arithmetic of one reverse step (sampling.py:94-99,119-167): times, and per component the coefficients of
    perturbation = c_score * score + c_noise * z.  Returns (t(3), coeffs(6), noisy(3 bools))."""
    last = t_idx == inference_steps - 1
    ts = (tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx])
    dts = [s[t_idx] - s[t_idx + 1] if not last else s[t_idx] for s in (tr_schedule, rot_schedule, tor_schedule)]
    sigmas = t_to_sigma(*ts)
    noiseless = no_random or (no_final_step_noise and last)
    coeffs, noisy = [], []
    for k, name in enumerate(("tr", "rot", "tor")):
        if k == 2 and no_torsion:
            coeffs += [0.0, 0.0]
            noisy.append(False)
            continue
        g, dt, temp, psi, sig = sigmas[k] * g_const[name], dts[k], temp_sampling[k], temp_psi[k], sigmas[k]
        smax, smin = getattr(model_args, f"{name}_sigma_max"), getattr(model_args, f"{name}_sigma_min")
        if ode:
            c_s, c_n, z = 0.5 * g ** 2 * dt, 0.0, False
        else:
            z = not noiseless
            c_s, c_n = g ** 2 * dt, g * np.sqrt(dt)
            if temp != 1.0:
                sigma_data = np.exp(temp_sigma_data * np.log(smax) + (1 - temp_sigma_data) * np.log(smin))
                lam = (sigma_data + sig) / (sigma_data + sig / temp)
                c_s, c_n = g ** 2 * dt * (lam + temp * psi / 2), g * np.sqrt(dt * (1 + psi))
        coeffs += [float(c_s), float(c_n)]
        noisy.append(z)
    return ts, coeffs, noisy


def reverse_diffusion(batch, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
                      model_args, mask_rotate, noise_rows=None, no_random=False, ode=False, t_schedule=None,
                      no_final_step_noise=False, temp_sampling=(1.0, 1.0, 1.0), temp_psi=(0.0, 0.0, 0.0),
                      temp_sigma_data=0.5, use_graph=None, validate_topology=True):
    """The hot loop of sampling.py:93-223 on a batch that already lives on the device: `inference_steps` x
    (score-model forward, K4 pose update).  Returns the final [B*N, 3] positions (also left in
    batch['ligand'].pos).  No host synchronisation inside.

    The first step runs eagerly (it builds the per-batch static cache and the derived-weight caches); the second is
    captured into a CUDA graph -- buffers are sized by host-known upper bounds, edge counts stay on the device, and the
    step's scalars (t per component, the six perturbation coefficients, the so3 / torus score norms) are read from a
    small device array -- and every later step is one graph replay after refreshing that array and the noise buffers.
    ~200 launches / 8 ms of host enqueue per step become one launch; the arithmetic is the eager path's, bit for bit."""
    from . import _lib, so3, torus
    b = batch.num_graphs
    N = noise_rows if noise_rows is not None else b
    batch_size = N
    no_torsion = bool(model_args.no_torsion)
    all_atoms = "all_atoms" in model_args and model_args.all_atoms
    asyncronous_noise_schedule = False
    g_const = {k: float(np.sqrt(np.float32(2 * np.log(getattr(model_args, f"{k}_sigma_max") / getattr(model_args, f"{k}_sigma_min")))))
               for k in ("tr", "rot", "tor")}
    if not _is_iterable(temp_sampling):
        temp_sampling = [temp_sampling] * 3
    if not _is_iterable(temp_psi):
        temp_psi = [temp_psi] * 3
    if hasattr(model_args, "crop_beyond") and model_args.crop_beyond is not None:
        raise NotImplementedError("per-step crop_beyond for the score model is not in the shipped score YAML")
    topo = None          # built below unless a cached step graph (which embeds it) is replayed
    pos = batch["ligand"].pos.float().contiguous()
    batch["ligand"].pos = pos
    nb = min(batch_size, N)
    if use_graph is None:
        use_graph = USE_CUDA_GRAPH
    use_graph = bool(use_graph) and pos.is_cuda and inference_steps >= 4 and _lib.tp_conv_hook is None
    scal = lambda k: _step_scalars(k, inference_steps, tr_schedule, rot_schedule, tor_schedule, t_to_sigma, model_args, g_const,
                                   no_random, ode, no_final_step_noise, temp_sampling, temp_psi, temp_sigma_data, no_torsion)

    def draw(noisy, tor_shape):
        """torch.normal in the reference's order (sampling.py:126-141): tr, rot, tor."""
        zs = []
        for k, shape in enumerate(((nb, 3), (nb, 3), tor_shape)):
            if not noisy[k]:
                zs.append(None)
                continue
            z = torch.normal(mean=0, std=1, size=shape, device=device)
            if k < 2 and z.shape[0] != b:
                raise RuntimeError(f"noise batch {z.shape[0]} != graphs in batch {b} (sampling.py:126-131: "
                                   "batch_size must divide the number of samples)")
            zs.append(z.float().contiguous())
        return zs

    graph, vals, z_static, tor_shape, graph_launches, step_table = None, None, None, None, 0, None
    # ---- a step graph captured for an earlier batch of the same complex (same model, weights, batch and schedule)?
    cache_key, reused = None, None
    fp = batch._g.get("_static_sig") if hasattr(batch, "_g") else None
    if use_graph and fp is not None and GRAPH_CACHE_SIZE > 0 and isinstance(model, torch.nn.Module):
        from . import tensor_layers
        wv = (tensor_layers.CACHE_EPOCH, sum(p._version for p in model.parameters()), sum(x._version for x in model.buffers()))
        sch = tuple(tuple(float(v) for v in sc) for sc in (tr_schedule, rot_schedule, tor_schedule))
        cache_key = (fp, id(model), wv, b, nb, int(inference_steps), sch, bool(no_random), bool(ode), bool(no_final_step_noise),
                     tuple(float(v) for v in temp_sampling), tuple(float(v) for v in temp_psi), float(temp_sigma_data), no_torsion,
                     t_schedule is None, torch.device(device).index, tuple(pos.shape))
        for k in [k for k in _graph_cache if k[1] == id(model) and k[2] != wv]:
            del _graph_cache[k]          # captured under other weights: its derived-weight tensors are gone
        reused = _graph_cache.get(cache_key)
        # same shapes is only a candidate: the graph is replayed when the model object is the captured one and every tensor of
        # the batch except the pose (and the rotatable-bond masks the pose update was built from) equals the captured batch's
        if reused is not None and (reused["model"]() is not model or not _same_static(batch, mask_rotate, reused)):
            del _graph_cache[cache_key]
            reused = None
    if reused is None and use_graph and _graph_cache:
        _graph_cache.clear()             # free the previous complex's graph pool before this batch allocates its own
    if reused is None:
        ready = batch._g.get("_ready") if hasattr(batch, "_g") else None
        if getattr(ready, "ev", None) is not None and pos.is_cuda and not validate_topology:
            # the topology tables are uploaded on the copy stream (a pageable H2D copy first synchronises ITS stream: on the
            # main stream the host would wait for the previous batch's replays)
            copy_s, main_s = _copy_stream(device), torch.cuda.current_stream(device)
            copy_s.wait_event(ready.ev)
            with torch.cuda.stream(copy_s):
                topo = LigandTopology(batch, mask_rotate, device, validate=False)
            main_s.wait_stream(copy_s)
            for t in (topo.bond_uv, topo.mask_rotate):
                t.record_stream(main_s)
        else:
            topo = LigandTopology(batch, mask_rotate, device, validate=validate_topology)
    warm_key = None
    if use_graph and isinstance(model, torch.nn.Module):
        from . import tensor_layers
        warm_key = (tensor_layers.CACHE_EPOCH, sum(p._version for p in model.parameters()), sum(x._version for x in model.buffers()))
    if (use_graph and reused is None and warm_key is not None and getattr(model, "_cb200_warm", None) == warm_key
            and hasattr(model, "_static") and CAPTURE_FIRST_STEP):
        # The model has run eagerly under these weights (program tables, folded weights exist): build the batch's static tables
        # (the only host reads of a forward), capture the step BEFORE step 0 and replay all steps -- no eager forward per batch.
        global _graph_warned
        try:
            model._static(batch)
            tor_shape = (0,) if no_torsion else (topo.B * topo.R,)
            graph, vals, z_static, graph_launches = _capture_step(batch, model, pos, topo, b, nb, tor_shape, no_torsion, scal(0)[2], device)
        except Exception as e:
            if not _graph_warned:
                warnings.warn(f"cb200: CUDA-graph capture of the reverse-diffusion step failed ({type(e).__name__}: {e}); "
                              "continuing with eager launches")
                _graph_warned = True
            graph, use_graph = None, False
    if reused is not None:
        global graph_cache_hits
        graph_cache_hits += 1
        _graph_cache.move_to_end(cache_key)
        graph, vals, z_static, tor_shape, graph_launches, step_table = (reused[k] for k in ("graph", "vals", "z_static", "tor_shape",
                                                                                             "graph_launches", "step_table"))
        reused["pos"].copy_(pos)
    for t_idx in range(inference_steps):
        (t_tr, t_rot, t_tor), coeffs, noisy = scal(t_idx)
        if graph is None:
            # ---- eager step (always the first one; every one when graphs are off)
            set_time(batch, t_schedule[t_idx] if t_schedule is not None else None, t_tr, t_rot, t_tor, b, all_atoms,
                     asyncronous_noise_schedule, device)
            if hasattr(batch, "cb200_step"):
                batch.cb200_step = None
            tr_score, rot_score, tor_score = model(batch)[:3]
            tor_shape = tuple(tor_score.shape)
            zs = draw(noisy, tor_shape)
            sde_step(pos, topo, tr_score, rot_score, None if no_torsion else tor_score, coeffs, zs[0], zs[1], zs[2])
            if warm_key is not None:
                object.__setattr__(model, "_cb200_warm", warm_key)
            if use_graph and t_idx == 0:
                try:
                    graph, vals, z_static, graph_launches = _capture_step(batch, model, pos, topo, b, nb, tor_shape, no_torsion,
                                                                          scal(1)[2], device)
                except Exception as e:          # capture is an optimisation: the eager loop is always correct
                    if not _graph_warned:
                        warnings.warn(f"cb200: CUDA-graph capture of the reverse-diffusion step failed ({type(e).__name__}: {e}); "
                                      "continuing with eager launches")
                        _graph_warned = True
                    graph, use_graph = None, False
            continue
        # ---- graph replay: refresh the step's scalars and noise (device-side copies: nothing blocks the host), then one launch
        if step_table is None:
            rows = []
            for k in range(inference_steps):
                (k_tr, k_rot, k_tor), k_coeffs, _ = scal(k)
                sig = t_to_sigma(k_tr, k_rot, k_tor)
                so3_norm = float(so3.score_norm(torch.full((1,), float(sig[1]), dtype=torch.float32))[0])
                torus_norm = float(torch.tensor(torus.score_norm(np.asarray([sig[2]], dtype=np.float32))).float()[0]) if not no_torsion else 1.0
                rows.append([k_tr, k_rot, k_tor, *k_coeffs, so3_norm, torus_norm])
            step_table = torch.tensor(rows, dtype=torch.float32).pin_memory().to(device, non_blocking=True)
        vals.copy_(step_table[t_idx])
        zs = draw(noisy, tor_shape)
        for zb, z in zip(z_static, zs):
            if zb is None and z is not None:
                raise RuntimeError("cb200: a noise term appeared after the step graph was captured without it")
            if zb is not None:
                if z is None:
                    zb.zero_()       # noiseless step (no_final_step_noise): the captured c_noise * z term adds exactly zero
                else:
                    zb.copy_(z)
        graph.replay()
        _lib.launch_count += graph_launches
    if graph is not None:
        if reused is not None:
            pos = reused["pos"].clone()          # the captured buffer belongs to the cache entry
            batch["ligand"].pos = pos
        elif cache_key is not None:
            # keep the graph for later batches of this complex: the entry owns the captured batch (every tensor the graph reads)
            # and the pose buffer the graph updates in place; the caller gets a copy
            _graph_cache[cache_key] = dict(graph=graph, vals=vals, z_static=z_static, tor_shape=tor_shape, graph_launches=graph_launches,
                                           step_table=step_table, pos=pos, topo=topo, keep=_tensor_refs(batch), model=weakref.ref(model),
                                           static=_static_tensors(batch), mask_rotate=copy.deepcopy(mask_rotate))
            while len(_graph_cache) > GRAPH_CACHE_SIZE:
                _graph_cache.popitem(last=False)
            pos = pos.clone()
            batch["ligand"].pos = pos
        if hasattr(batch, "cb200_step"):
            batch.cb200_step = None
        set_time(batch, None, tr_schedule[inference_steps - 1], rot_schedule[inference_steps - 1], tor_schedule[inference_steps - 1], b,
                 all_atoms, asyncronous_noise_schedule, device)
    return pos


def _static_tensors(batch):
    """name -> tensor for the attributes listed in the batch's collate-time signature (data._static_signature): everything a
    collated batch carries except the ligand pose; attributes the model or set_time added later are not part of it."""
    names = {(k, a) for k, a, _, _ in (batch._g.get("_static_sig") or ())}
    out = {}
    for key, st in batch._stores.items():
        for k, v in st._d.items():
            if torch.is_tensor(v) and (str(key), k) in names:
                out[(key, k)] = v
    return out


def _same_static(batch, mask_rotate, entry):
    """Is `batch` the captured batch up to the ligand pose?  One device-side comparison per tensor, one host sync."""
    mr0 = entry["mask_rotate"]
    try:
        if type(mr0) is not type(mask_rotate):
            return False
        if isinstance(mr0, (list, tuple)):
            if len(mr0) != len(mask_rotate) or not all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(mr0, mask_rotate)):
                return False
        elif not np.array_equal(np.asarray(mr0), np.asarray(mask_rotate)):
            return False
    except Exception:
        return False
    pairs = []
    for (key, k), ref in entry["static"].items():
        st = batch._stores.get(key)
        v = st._d.get(k) if st is not None else None
        if not torch.is_tensor(v) or v.shape != ref.shape or v.dtype != ref.dtype or v.device != ref.device:
            return False
        if v.numel() > 0:
            pairs.append((v, ref))
    if not pairs:
        return True
    # The comparison runs on the second stream when the batch carries the event of its collate: the host then waits for that
    # batch only, not for the previous batch's 20 replays still queued on the main stream (it can run a whole batch ahead).
    ready = batch._g.get("_ready") if hasattr(batch, "_g") else None
    ev = getattr(ready, "ev", None)
    if ev is not None and pairs[0][0].is_cuda:
        side = _conf_stream(pairs[0][0].device)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            ok = torch.stack([(v == ref).all() for v, ref in pairs]).all()
            for v, ref in pairs:
                v.record_stream(side)
                ref.record_stream(side)
            return bool(ok)
    return bool(torch.stack([(v == ref).all() for v, ref in pairs]).all())


def _tensor_refs(batch):
    """Every tensor reachable from the batch's stores and graph-level attributes (static caches included): a cached step
    graph reads them by address, so the cache entry keeps them alive whatever the caller does with the batch object."""
    out, seen = [], set()

    def walk(v, depth=0):
        if torch.is_tensor(v):
            out.append(v)
        elif depth < 6 and id(v) not in seen:
            seen.add(id(v))
            if isinstance(v, dict):
                for x in v.values():
                    walk(x, depth + 1)
            elif isinstance(v, (list, tuple)):
                for x in v:
                    walk(x, depth + 1)
            elif hasattr(v, "__dict__") and not isinstance(v, (torch.nn.Module, type)):
                for x in vars(v).values():
                    walk(x, depth + 1)

    for st in batch._stores.values():
        walk(st._d)
    walk(batch._g)
    return out


def _capture_step(batch, model, pos, topo, b, nb, tor_shape, no_torsion, noisy, device):
    """One reverse step (score-model forward + K4) captured into a CUDA graph whose per-step inputs live in static device
    buffers: vals = [t_tr, t_rot, t_tor, c_tr_s, c_tr_n, c_rot_s, c_rot_n, c_tor_s, c_tor_n, so3_norm, torus_norm]."""
    from . import _lib
    vals = torch.zeros(11, dtype=torch.float32, device=device)
    vals[:3] = 0.5          # a valid time while capturing (nothing executes, but shapes / code paths are those of a real step)
    z_static = [torch.zeros((nb, 3), device=device) if noisy[0] else None, torch.zeros((nb, 3), device=device) if noisy[1] else None,
                torch.zeros(tor_shape, device=device) if noisy[2] else None]
    batch.complex_t = {"tr": vals[0:1].expand(b), "rot": vals[1:2].expand(b), "tor": vals[2:3].expand(b)}
    batch.complex_t_host = None
    batch.cb200_step = {"so3_norm": vals[9:10], "torus_norm": vals[10:11]}
    # Manual capture on a side stream instead of the torch.cuda.graph() context: that context synchronises the device,
    # runs the garbage collector and EMPTIES the caching allocator on entry, which costs ~140 ms per sampling call once the
    # eager steps have to cudaMalloc their multi-GB workspaces again.  Here the capture only waits (on the device) for the
    # work already enqueued, and the host records the graph while the GPU is still busy with the eager first step.
    graph = torch.cuda.CUDAGraph()
    l0 = _lib.launch_count
    main = torch.cuda.current_stream()
    side = _capture_stream(device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        graph.capture_begin(pool=_graph_pool(device))
        try:
            tr_score, rot_score, tor_score = model(batch)[:3]
            sde_step(pos, topo, tr_score, rot_score, None if no_torsion else tor_score, vals[3:9], z_static[0], z_static[1], z_static[2])
        finally:
            graph.capture_end()
    main.wait_stream(side)
    return graph, vals, z_static, _lib.launch_count - l0


_SAMPLING_DEFAULTS = dict(no_random=False, ode=False, visualization_list=None, confidence_model=None, filtering_data_list=None,
                          filtering_model_args=None, asyncronous_noise_schedule=False, t_schedule=None, batch_size=32,
                          no_final_step_noise=False, pivot=None, return_full_trajectory=False, temp_sampling=1.0, temp_psi=0.0,
                          temp_sigma_data=0.5, return_features=False)


def _sample_batches(leg, data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args,
                    no_random, ode, visualization_list, confidence_model, filtering_data_list, filtering_model_args,
                    asyncronous_noise_schedule, t_schedule, batch_size, no_final_step_noise, temp_sampling, temp_psi,
                    temp_sigma_data):
    """The batch loop of one sampling() call (sampling.py:78-262): reverse diffusion per batch on the current stream, the
    filtering leg of every batch handed to `leg` (which runs it one submission later).  Returns the number of legs submitted."""
    N = len(data_list)
    # batches are collated straight onto the device: attributes shared by the N copies of a complex cross the bus once
    loader = DataLoader(data_list, batch_size=batch_size, device=device)
    mask_rotate = _mask_rotate_of(data_list[0])
    filtering_loader = None
    if confidence_model is not None and filtering_data_list is not None:
        filtering_loader = iter(DataLoader(filtering_data_list, batch_size=batch_size, device=device))
    if not _is_iterable(temp_sampling):
        temp_sampling = [temp_sampling] * 3
    if not _is_iterable(temp_psi):
        temp_psi = [temp_psi] * 3
    assert len(temp_sampling) == 3 and len(temp_psi) == 3
    n_legs = 0
    # the orientation check of the rotation masks runs once on the host graph (the device-side one needs two host reads)
    checked = check_rotation_masks(data_list[0], mask_rotate)
    # Batches are collated on a COPY stream: pageable H2D copies are ordered behind whatever is queued on their stream, and on
    # the main stream that is the previous batch's (or previous complex's) 20 graph replays -- the host would sit in
    # cudaMemcpy for their whole duration instead of preparing this batch.  The main stream waits for the copy stream's event;
    # every tensor of the batch is marked as used by the main stream for the allocator.
    two_streams = leg.enabled and torch.device(device).type == "cuda"
    it = iter(loader)
    batch_id = -1
    while True:
        if two_streams:
            copy_s, main_s = _copy_stream(device), torch.cuda.current_stream(device)
            with torch.cuda.stream(copy_s):
                batch = next(it, None)
                if batch is not None:
                    batch = batch.to(device)
            if batch is not None:
                main_s.wait_stream(copy_s)
                for t in _tensor_refs(batch):
                    if t.is_cuda:
                        t.record_stream(main_s)
        else:
            batch = next(it, None)
            if batch is not None:
                batch = batch.to(device)
        if batch is None:
            break
        batch_id += 1
        b = batch.num_graphs
        pos = reverse_diffusion(batch, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device,
                                t_to_sigma, model_args, mask_rotate, noise_rows=min(batch_size, N), no_random=no_random,
                                ode=ode, t_schedule=t_schedule, no_final_step_noise=no_final_step_noise,
                                temp_sampling=temp_sampling, temp_psi=temp_psi, temp_sigma_data=temp_sigma_data,
                                validate_topology=not checked)
        n = pos.shape[0] // b
        for i in range(b):
            data_list[batch_id * batch_size + i]["ligand"].pos = pos[i * n:n * (i + 1)]
        if visualization_list is not None:
            for idx, vis in enumerate(visualization_list):
                vis.add((data_list[idx]["ligand"].pos.detach().cpu() + data_list[idx].original_center.detach().cpu()),
                        part=1, order=2)

        if confidence_model is not None:
            # the filtering batch is collated (host compare + H2D, ~7 ms) on the leg's stream AFTER this batch's steps
            # were enqueued: on the main stream its copies would queue behind the 20 replays and block the host
            fb0 = None
            if filtering_loader is not None:
                with (torch.cuda.stream(_conf_stream(device)) if leg.enabled else contextlib.nullcontext()):
                    fb0 = next(filtering_loader)

            def conf_leg(p, fb=fb0, batch=batch, b=b):
                if fb is not None:
                    fb = fb.to(device)           # (already there when the loader collates onto the device)
                    fb["ligand"].pos = p
                    if hasattr(filtering_model_args, "crop_beyond") and filtering_model_args.crop_beyond is not None:
                        fb = crop_beyond(fb, filtering_model_args.crop_beyond, filtering_model_args.all_atoms)
                    set_time(fb, 0, 0, 0, 0, b, filtering_model_args.all_atoms, asyncronous_noise_schedule, device)
                    out = confidence_model(fb)
                else:
                    out = confidence_model(batch)
                return out[0] if type(out) is tuple else out

            leg.submit(conf_leg, pos)        # runs the previous leg while this batch's steps are in flight
            n_legs += 1
    return n_legs


def _check_sampling_flags(N, batch_size, return_features, return_full_trajectory, pivot, asyncronous_noise_schedule, svgd0, svgd1):
    if return_features:
        assert batch_size >= N, "Not implemented yet"
    if svgd0 is not None and svgd1 is not None:
        raise NotImplementedError("SVGD coupling (sampling.py:169-218) is an optional branch outside the hot path")
    assert not (return_full_trajectory or return_features or pivot), "Not implemented yet in new inference version"
    assert not asyncronous_noise_schedule


def _finish_confidence(chunks):
    confidence = torch.cat(chunks, dim=0)
    return torch.nan_to_num(confidence, nan=-1000)


def sampling(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args,
             no_random=False, ode=False, visualization_list=None, confidence_model=None, filtering_data_list=None,
             filtering_model_args=None, asyncronous_noise_schedule=False, t_schedule=None, batch_size=32,
             no_final_step_noise=False, pivot=None, return_full_trajectory=False, temp_sampling=1.0, temp_psi=0.0,
             temp_sigma_data=0.5, return_features=False,
             svgd_weight_log_0=None, svgd_repulsive_weight_log_0=None, svgd_weight_log_1=None,
             svgd_repulsive_weight_log_1=None, svgd_kernel_size_log_0=None, svgd_kernel_size_log_1=None,
             svgd_langevin_weight_log_0=None, svgd_langevin_weight_log_1=None, svgd_rot_log_rel_weight=0.0,
             svgd_tor_log_rel_weight=0.0, svgd_use_x0=False):
    """utils/sampling.py:59 `sampling()`: same signature, same return value `(data_list, confidence)`."""
    _check_sampling_flags(len(data_list), batch_size, return_features, return_full_trajectory, pivot, asyncronous_noise_schedule,
                          svgd_weight_log_0, svgd_weight_log_1)
    # (the filtering leg of a confidence model fed with the SCORE batch shares tensors with the steps: kept on one stream)
    leg = FilteringLeg(device, enabled=None if filtering_data_list is not None else False)
    with torch.no_grad(), _inference_mode(model, confidence_model):
        _sample_batches(leg, data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args,
                        no_random, ode, visualization_list, confidence_model, filtering_data_list, filtering_model_args,
                        asyncronous_noise_schedule, t_schedule, batch_size, no_final_step_noise, temp_sampling, temp_psi,
                        temp_sigma_data)
        done = leg.finish()
    confidence = _finish_confidence(done) if confidence_model is not None else None
    return data_list, confidence


def sampling_many(requests, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma, model_args, **kwargs):
    """The callers' loop `for complex in loader: sampling(data_list, ...)` (inference.py:432-470, finetune_train.py:168-195,
    bootstrapping.py:106-140) as ONE call: `requests` is a sequence of `(data_list, filtering_data_list)` pairs (or bare
    data_lists), the other arguments are sampling()'s and apply to every request.  Returns `[(data_list, confidence), ...]`
    with exactly the values the separate calls return -- but the filtering leg of complex i runs on the second stream while
    complex i+1 is collated, captured and stepped on the main one, so neither the host work of a new complex (collate,
    static tables, graph capture: ~40 ms) nor the shape-dependent host reads of the leg leave the GPU idle."""
    kw = dict(_SAMPLING_DEFAULTS)
    unknown = set(kwargs) - set(kw) - {"filtering_data_list"}
    svgd = {k: kwargs.pop(k) for k in list(kwargs) if k.startswith("svgd_")}
    unknown = {k for k in unknown if not k.startswith("svgd_")}
    if unknown:
        raise TypeError(f"sampling_many() got unexpected keyword arguments {sorted(unknown)}")
    kw.update(kwargs)
    kw.pop("filtering_data_list")
    confidence_model = kw["confidence_model"]
    reqs = [(r, None) if not (isinstance(r, (tuple, list)) and len(r) == 2 and (r[1] is None or isinstance(r[1], (list, tuple)))
                              and isinstance(r[0], (list, tuple))) else tuple(r) for r in requests]
    for dl, _ in reqs:
        _check_sampling_flags(len(dl), kw["batch_size"], kw["return_features"], kw["return_full_trajectory"], kw["pivot"],
                              kw["asyncronous_noise_schedule"], svgd.get("svgd_weight_log_0"), svgd.get("svgd_weight_log_1"))
    pipelined = confidence_model is not None and all(fl is not None for _, fl in reqs)
    leg = FilteringLeg(device, enabled=None if pipelined else False)
    counts = []
    with torch.no_grad(), _inference_mode(model, confidence_model):
        for dl, fl in reqs:
            counts.append(_sample_batches(leg, dl, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
                                          model_args, kw["no_random"], kw["ode"], kw["visualization_list"], confidence_model, fl,
                                          kw["filtering_model_args"], kw["asyncronous_noise_schedule"], kw["t_schedule"],
                                          kw["batch_size"], kw["no_final_step_noise"], kw["temp_sampling"], kw["temp_psi"],
                                          kw["temp_sigma_data"]))
        done = leg.finish()
    out, k = [], 0
    for (dl, _), c in zip(reqs, counts):
        out.append((dl, _finish_confidence(done[k:k + c]) if confidence_model is not None else None))
        k += c
    return out
