"""Shadows the reference's model/heads.py (parameter containers only; see rohm_b200/heads.py)."""
from rohm_b200.heads import *  # noqa: F401,F403
from rohm_b200.heads import (Conv1dBlock, Downsample1d, InputProcess, OutputProcess, PositionalEncoding,  # noqa: F401
                             ResidualTemporalBlock, SinusoidalPosEmb, TimestepEmbedder, Upsample1d, zero_module)
