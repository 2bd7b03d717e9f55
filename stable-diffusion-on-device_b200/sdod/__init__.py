"""sdod — B200-native drop-in for the hot path of vaenyr/stable-diffusion-on-device.

Import name and exports follow the reference package (sdod/__init__.py:1: ``EfficientGN``).
"""
from .efficient_gn import EfficientGN, efficient_group_norm  # noqa: F401

__version__ = "0.1.0"
