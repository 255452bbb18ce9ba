"""Drop-in for reference models/projection/z_buffer_manipulator.py:11-294 (PtsManipulator)."""
from pixelsynth_b200.models.projection.z_buffer_manipulator import EPS, PtsManipulator, get_splatter  # noqa: F401
