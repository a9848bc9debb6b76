"""Shadows the reference's model/posenet.py with the B200 implementation."""
from rohm_b200.posenet import PoseNet  # noqa: F401
