"""All-atom TensorProductScoreModel (models/all_atom_score_model.py) -- placeholder until implemented."""
from torch import nn


class TensorProductScoreModel(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("all-atom model: next build step")
