"""B200-native reverse-diffusion pose-sampling hot path of Confidence Bootstrapping.

Public surface (mirrors the reference modules of the same names):
    utils.get_model, utils.crop_beyond
    sampling.sampling, sampling.randomize_position
    score_model.TensorProductScoreModel, all_atom_score_model.TensorProductScoreModel
    tensor_layers.TensorProductConvLayer
    diffusion_utils.{t_to_sigma, get_t_schedule, set_time, modify_conformer_batch}
"""
__version__ = "0.1.0"
