"""Drop-in for reference models/layers/z_buffer_layers.py:11-131 (RasterizePointsXYsBlending)."""
from pixelsynth_b200.models.layers.z_buffer_layers import RasterizePointsXYsBlending  # noqa: F401
