"""Parameter holders with the reference's state-dict names and shapes (model/component.py).

ControlKnotLieAlgebra / TransformationLieAlgebra are nn.Embedding tables exactly as upstream
(model/component.py:7-15).  The tone-mapper CRFs are disabled in every shipped config
(optimize_*_crf = False; NeRF.raw2output ignores them, model/nerf.py:127-131): they are kept as
holders of the same four tensors each so that checkpoints round-trip, nothing more.
"""
import torch.nn as nn


class ControlKnotLieAlgebra(nn.Module):
    def __init__(self, knot_num):
        super().__init__()
        self.params = nn.Embedding(knot_num, 6)


class TransformationLieAlgebra(nn.Module):
    def __init__(self, trans_num):
        super().__init__()
        self.params = nn.Embedding(trans_num, 6)


class ColorToneMapper(nn.Module):
    def __init__(self, hidden=0, width=128, input_type="Gray"):
        super().__init__()
        self.mlp_gray = nn.Sequential(nn.Linear(1, width), nn.ReLU(), nn.Linear(width, 1))


class LuminanceToneMapper(nn.Module):
    def __init__(self, hidden=0, width=128, input_type="Gray"):
        super().__init__()
        self.mlp_luminance = nn.Sequential(nn.Linear(1, width), nn.ReLU(), nn.Linear(width, 1))
