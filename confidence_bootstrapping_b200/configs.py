"""Hyper-parameters of the two shipped (but weight-less) reference checkpoints.

Values transcribed from workdir/pretrained_score/model_parameters.yml and
workdir/pretrained_confidence/model_parameters.yml (only the keys `get_model` / `sampling` read).
Keys ABSENT from the confidence YAML (sh_lmax, num_prot_emb_layers, reduce_pseudoscalars,
embed_also_ligand, embedding_type ...) stay absent here, so the reference's defaults apply
(utils/utils.py:274-283).
"""
from argparse import Namespace

SCORE_MODEL_ARGS = dict(
    all_atoms=False, asyncronous_noise_schedule=False, confidence_dropout=0.0, confidence_no_batchnorm=False,
    cross_distance_embed_dim=32, cross_max_distance=80, distance_embed_dim=32, dropout=0.1, dynamic_max_cross=True,
    embed_also_ligand=True, embedding_scale=1000, embedding_type="sinusoidal", esm_embeddings_model=None,
    include_miscellaneous_atoms=False, max_radius=5.0, no_aminoacid_identities=False, no_batch_norm=False,
    no_differentiate_convolutions=False, no_torsion=False, norm_by_sigma=False, not_fixed_center_conv=False, ns=32,
    num_conv_layers=5, num_prot_emb_layers=3, nv=6, odd_parity=False, receptor_radius=15.0, c_alpha_max_neighbors=24,
    moad_esm_embeddings_path="precomputed", reduce_pseudoscalars=True, rot_sigma_max=3.1, rot_sigma_min=0.06,
    scale_by_sigma=True, separate_noise_schedule=False, sh_lmax=1, sigma_embed_dim=32, smooth_edges=False,
    tor_sigma_max=3.14, tor_sigma_min=0.0314, tp_weights_layers=2, tr_sigma_max=19.0, tr_sigma_min=0.1,
    use_second_order_repr=False, depthwise_convolution=False, sidechain_loss_weight=0, backbone_loss_weight=0,
    inference_steps=20, atom_radius=5, atom_max_neighbors=8)

CONFIDENCE_MODEL_ARGS = dict(
    affinity_prediction=False, all_atoms=True, asyncronous_noise_schedule=False, atom_confidence_loss_weight=0.5,
    atom_max_neighbors=8, atom_radius=5, atom_rmsd_classification_cutoff=2.0, c_alpha_max_neighbors=24,
    confidence_dropout=0.0, confidence_no_batchnorm=False, crop_beyond=20.0, cross_distance_embed_dim=32,
    cross_max_distance=80, distance_embed_dim=32, dropout=0.1, dynamic_max_cross=True, embedding_scale=10000,
    embedding_type="sinusoidal", esm_embeddings_path="precomputed", max_radius=5.0, no_batch_norm=False, no_torsion=False,
    norm_by_sigma=False, ns=24, num_conv_layers=5, nv=6, odd_parity=False, parallel=1,
    parallel_aggregators="mean max min std", receptor_radius=15.0, rmsd_classification_cutoff=2.0,
    scale_by_sigma=True, separate_noise_schedule=False, sigma_embed_dim=32, smooth_edges=False,
    use_second_order_repr=False, inference_steps=20)


def score_model_args(**overrides) -> Namespace:
    return Namespace(**{**SCORE_MODEL_ARGS, **overrides})


def confidence_model_args(**overrides) -> Namespace:
    return Namespace(**{**CONFIDENCE_MODEL_ARGS, **overrides})


def all_atom_score_model_args(**overrides) -> Namespace:
    """BASELINE.json config 5: the all-atom SCORE model.  The reference ships no YAML for it; the
    benchmarked hyper-parameters are the confidence YAML's sizes (ns 24, nv 6, 5 conv layers, lmax 2 by default, no
    receptor-embedding layers, 1280-d precomputed LM features) with the score YAML's diffusion ranges and heads."""
    base = {**SCORE_MODEL_ARGS, "all_atoms": True, "ns": 24, "nv": 6, "num_conv_layers": 5, "num_prot_emb_layers": 0,
            "sh_lmax": 2, "reduce_pseudoscalars": False, "embed_also_ligand": False, "embedding_scale": 10000}
    return Namespace(**{**base, **overrides})
