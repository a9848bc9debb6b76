"""Import-compatibility stub for the reference's model/cfg_sampler.py.

ClassifierFreeSampleModel is dead code in RoHM (inherited from MDM, imported nowhere; its constructor reads attributes
neither PoseNet nor TrajNet defines, so it cannot be instantiated around them -- SURVEY.md D2).  The class is kept so
``import model.cfg_sampler`` keeps working; it is functionally out of scope.
"""
import torch.nn as nn


class ClassifierFreeSampleModel(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        assert self.model.cond_mask_prob > 0, 'Cannot run a guided diffusion on a model that has not been trained with no conditions'
        self.translation = self.model.translation
        self.njoints = self.model.njoints
        self.data_rep = self.model.data_rep
        self.cond_mode = self.model.cond_mode

    def forward(self, x, timesteps, y=None):
        raise NotImplementedError("classifier-free sampling is not part of RoHM's inference path")
