"""Model factory and receptor cropping (utils/utils.py:175-288, 395-420)."""
from __future__ import annotations

import torch

from . import _lib
from .diffusion_utils import get_timestep_embedding


def get_model(args, device, t_to_sigma, no_parallel=False, confidence_mode=False, old=False):
    """Same argument -> constructor mapping (with the same defaults for missing keys) as the reference
    factory, utils/utils.py:175-288.  `old` (legacy checkpoints) is out of scope."""
    from .all_atom_score_model import TensorProductScoreModel as AAScoreModel
    from .score_model import TensorProductScoreModel as CGScoreModel
    if old:
        raise NotImplementedError("legacy (old_score_model) checkpoints are outside the shipped configurations")
    has = lambda k: k in args
    timestep_emb_func = get_timestep_embedding(
        embedding_type=args.embedding_type if has("embedding_type") else "sinusoidal",
        embedding_dim=args.sigma_embed_dim,
        embedding_scale=args.embedding_scale if has("embedding_type") else 10000)
    model_class = AAScoreModel if has("all_atoms") and args.all_atoms else CGScoreModel
    lm_embedding_type = None
    if any(has(k) and getattr(args, k) is not None for k in (
            "moad_esm_embeddings_path", "pdbbind_esm_embeddings_path", "pdbsidechain_esm_embeddings_path",
            "esm_embeddings_path")):
        lm_embedding_type = "precomputed"
    if has("esm_embeddings_model") and args.esm_embeddings_model is not None:
        lm_embedding_type = args.esm_embeddings_model
    cut = lambda k: len(getattr(args, k)) + 1 if has(k) and isinstance(getattr(args, k), list) else 1
    model = model_class(
        t_to_sigma=t_to_sigma, device=device, no_torsion=args.no_torsion, timestep_emb_func=timestep_emb_func,
        num_conv_layers=args.num_conv_layers, lig_max_radius=args.max_radius, scale_by_sigma=args.scale_by_sigma,
        sigma_embed_dim=args.sigma_embed_dim, norm_by_sigma=has("norm_by_sigma") and args.norm_by_sigma,
        ns=args.ns, nv=args.nv, distance_embed_dim=args.distance_embed_dim,
        cross_distance_embed_dim=args.cross_distance_embed_dim, batch_norm=not args.no_batch_norm, dropout=args.dropout,
        use_second_order_repr=args.use_second_order_repr, cross_max_distance=args.cross_max_distance,
        dynamic_max_cross=args.dynamic_max_cross, separate_noise_schedule=args.separate_noise_schedule,
        smooth_edges=args.smooth_edges if has("smooth_edges") else False,
        odd_parity=args.odd_parity if has("odd_parity") else False,
        lm_embedding_type=lm_embedding_type, confidence_mode=confidence_mode,
        asyncronous_noise_schedule=args.asyncronous_noise_schedule if has("asyncronous_noise_schedule") else False,
        affinity_prediction=args.affinity_prediction if has("affinity_prediction") else False,
        parallel=args.parallel if has("parallel") else 1,
        num_confidence_outputs=cut("rmsd_classification_cutoff"),
        atom_num_confidence_outputs=cut("atom_rmsd_classification_cutoff"),
        parallel_aggregators=args.parallel_aggregators if has("parallel_aggregators") else "",
        fixed_center_conv=not args.not_fixed_center_conv if has("not_fixed_center_conv") else False,
        no_aminoacid_identities=args.no_aminoacid_identities if has("no_aminoacid_identities") else False,
        include_miscellaneous_atoms=args.include_miscellaneous_atoms if hasattr(args, "include_miscellaneous_atoms") else False,
        sh_lmax=args.sh_lmax if has("sh_lmax") else 2,
        differentiate_convolutions=not args.no_differentiate_convolutions if has("no_differentiate_convolutions") else True,
        tp_weights_layers=args.tp_weights_layers if has("tp_weights_layers") else 2,
        num_prot_emb_layers=args.num_prot_emb_layers if has("num_prot_emb_layers") else 0,
        reduce_pseudoscalars=args.reduce_pseudoscalars if has("reduce_pseudoscalars") else False,
        embed_also_ligand=args.embed_also_ligand if has("embed_also_ligand") else False,
        atom_confidence=args.atom_confidence_loss_weight > 0.0 if has("atom_confidence_loss_weight") else False,
        sidechain_pred=(hasattr(args, "sidechain_loss_weight") and args.sidechain_loss_weight > 0) or
                       (hasattr(args, "backbone_loss_weight") and args.backbone_loss_weight > 0),
        depthwise_convolution=args.depthwise_convolution if hasattr(args, "depthwise_convolution") else False)
    # The reference wraps the model in PyG DataParallel on CUDA (training only; sampling always uses
    # `.module`, finetune_train.py:177).  One process per GPU here: expose `.module` without a wrapper.
    device = torch.device(device)
    if device.type == "cuda" and not no_parallel and not (has("dataset") and args.dataset == "torsional"):
        # plain attribute, NOT a registered child: nn.Module.__setattr__ would make the model its own submodule and every
        # recursive call (.to / .eval / .state_dict / .parameters) would never terminate
        object.__setattr__(model, "module", model)
    model.to(device)
    return model


def _batch_vec(store, n, device):
    return store.batch if "batch" in store else torch.zeros(n, dtype=torch.long, device=device)


def crop_beyond(complex_graph, cutoff, all_atoms):
    """Keep only the residues (and their atoms) with any ligand atom closer than `cutoff`
    (utils/utils.py:395-420), for a single graph or a whole batch, on the device, in place.

    The reference crops graph by graph on the host between `to_data_list` / `from_data_list`
    (sampling.py:245-250).  Here the keep-mask comes from the K1 neighbour kernel (a residue is kept
    iff the transposed radius search finds at least one ligand atom of its own graph: the same fp32
    predicate sum((lig - rec)^2) < cutoff^2), and the compaction / index remapping is done with
    prefix sums on the device."""
    from .graph import radius_edges_transposed  # noqa: F401  (kept for symmetry with the forward search)
    lig, rec = complex_graph["ligand"], complex_graph["receptor"]
    dev = lig.pos.device
    if not lig.pos.is_cuda:
        raise RuntimeError("crop_beyond: expected CUDA tensors (the cb200 kernels have no CPU path)")
    B = int(complex_graph.num_graphs)
    lig_pos, rec_pos = lig.pos.float().contiguous(), rec.pos.float().contiguous()
    lig_batch = _batch_vec(lig, lig_pos.shape[0], dev)
    rec_batch = _batch_vec(rec, rec_pos.shape[0], dev)
    lig_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    lig_ptr[1:] = torch.cumsum(torch.bincount(lig_batch, minlength=B), 0)
    count = torch.empty(rec_pos.shape[0], dtype=torch.int32, device=dev)
    _lib.radius_count_t(rec_pos, rec_batch.to(torch.int32).contiguous(), lig_pos, lig_ptr, None, float(cutoff), False,
                        None, None, count)
    keep = count > 0

    def relabel(mask):
        return torch.cumsum(mask.long(), 0) - 1

    rr = complex_graph["receptor", "receptor"]
    rec_map = relabel(keep)
    if all_atoms:
        atom, aa, ar = complex_graph["atom"], complex_graph["atom", "atom"], complex_graph["atom", "receptor"]
        atom_to_res = ar.edge_index[1]
        atoms_keep = keep[atom_to_res]
        new_atom_res = rec_map[atom_to_res][atoms_keep]
        ar_new = torch.stack([torch.arange(len(new_atom_res), device=dev), new_atom_res])
    for k in ("pos", "x", "side_chain_vecs"):
        if k in rec:
            setattr(rec, k, getattr(rec, k)[keep])
    if "batch" in rec:
        rec.batch = rec_batch[keep]
    ei = rr.edge_index
    ek = keep[ei[0]] & keep[ei[1]]
    rr.edge_index = rec_map[ei[:, ek]]
    if all_atoms:
        atom_map = relabel(atoms_keep)
        atom_batch = _batch_vec(atom, atom.pos.shape[0], dev)
        atom.x, atom.pos = atom.x[atoms_keep], atom.pos[atoms_keep]
        if "batch" in atom:
            atom.batch = atom_batch[atoms_keep]
        ei = aa.edge_index
        ek = atoms_keep[ei[0]] & atoms_keep[ei[1]]
        aa.edge_index = atom_map[ei[:, ek]]
        ar.edge_index = ar_new
    for st in (rec, complex_graph["atom"] if all_atoms else None):
        if st is not None:
            for stale in ("cb200_static", "cb200_static_aa", "ptr"):
                if stale in st:
                    delattr(st, stale)
    # every graph kept a different subset of its receptor: the collate's bookkeeping no longer describes the batch
    g = getattr(complex_graph, "_g", None)
    if isinstance(g, dict):
        if "_replicated_types" in g:
            g["_replicated_types"] = [t for t in g["_replicated_types"] if t not in ("receptor", "atom")]
        if "_slices" in g:
            g["_slices_stale"] = True      # Batch.to_data_list re-derives slices / offsets from the batch vectors
    return complex_graph
