"""SMPL-X-shaped body model container used when the third-party ``smplx`` package is not installed.

Holds the model tensors as buffers under smplx's own names (v_template, shapedirs, expr_dirs is folded into
shapedirs[..., 10:], posedirs, J_regressor, lbs_weights, parents) and evaluates joints / vertices with the CUDA
kernels of the library (rohm_fk22_* / rohm_lbs_forward).  See DESIGN.md for the provenance of the algorithm
(smplx==0.1.28, not vendored by the reference).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from ._lib import RohmB200Error
from . import synthetic


class BodyOutput:
    def __init__(self, joints=None, vertices=None):
        self.joints = joints
        self.vertices = vertices


class BodyModel(nn.Module):
    NUM_JOINTS = 55
    NUM_BODY_JOINTS = 21

    def __init__(self, tensors):
        super().__init__()
        for name in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
            self.register_buffer(name, tensors[name].to(torch.float32).contiguous())
        self.register_buffer("parents", torch.tensor(tensors["parents"], dtype=torch.long))
        self._handle = None

    @staticmethod
    def create(body_model_path='', device=None, seed=0):
        """Loads ``<path>/smplx/SMPLX_NEUTRAL.npz`` when it exists (official file layout), otherwise builds the
        seeded synthetic SMPL-X-shaped model (tests / benchmarks)."""
        npz = os.path.join(body_model_path or '', 'smplx', 'SMPLX_NEUTRAL.npz')
        if body_model_path and os.path.exists(npz):
            d = np.load(npz, allow_pickle=True)
            shapedirs = np.concatenate([d['shapedirs'][:, :, :10], d['shapedirs'][:, :, 300:310]], axis=-1)
            V = d['v_template'].shape[0]
            tensors = {"v_template": torch.from_numpy(np.asarray(d['v_template'], np.float32)),
                       "shapedirs": torch.from_numpy(np.asarray(shapedirs, np.float32)),
                       "posedirs": torch.from_numpy(np.asarray(d['posedirs'], np.float32).reshape(V * 3, -1).T.copy()),
                       "J_regressor": torch.from_numpy(np.asarray(d['J_regressor'], np.float32)),
                       "lbs_weights": torch.from_numpy(np.asarray(d['weights'], np.float32)),
                       "parents": [int(p) for p in np.asarray(d['kintree_table'][0], np.int64)]}
            tensors["parents"][0] = -1
        else:
            tensors = synthetic.smplx_like_model(seed)
        m = BodyModel(tensors)
        return m.to(device) if device is not None else m

    def as_dict(self):
        return {"v_template": self.v_template, "shapedirs": self.shapedirs, "posedirs": self.posedirs,
                "J_regressor": self.J_regressor, "lbs_weights": self.lbs_weights, "parents": self.parents.tolist()}

    def forward(self, **params):  # filled in with the FK / LBS kernels
        raise RohmB200Error("BodyModel.forward: FK/LBS kernels are not built into this library version")
