"""Shadows the reference's utils/model_util.py."""
from rohm_b200.diffusion import create_gaussian_diffusion, space_timesteps  # noqa: F401
