"""Shared helpers for the test-suite (not collected by pytest)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: reconstructed parameters within 1e-4 abs (fp32)
TOL = 1e-4


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


class NoiseTape:
    """The seeded CPU noise stream tools/gen_golden.py fed to the reference (same generator calls, same order)."""

    def __init__(self, seed, device="cpu"):
        self.g = torch.Generator().manual_seed(int(seed))
        self.device = device

    def randn(self, *shape, device=None, **kw):
        return torch.randn(*shape, generator=self.g).to(self.device)

    def randn_like(self, x):
        return torch.randn(x.shape, generator=self.g).to(self.device)


def posenet_state_dict(seed=1):
    """Seeded PoseNet weights in the reference's state-dict format, built WITHOUT the reference or the GPU."""
    from rohm_b200 import synthetic
    from rohm_b200.posenet import PoseNet
    ds = synthetic.make_dataset('pose')
    m = PoseNet(dataset=ds, body_feat_dim=294, latent_dim=512, ff_size=1024, num_layers=8, num_heads=4, device=None,
                traj_feat_dim=22)
    return synthetic.synth_state_dict(m, seed), m
