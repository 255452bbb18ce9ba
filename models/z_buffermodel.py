"""Drop-in for reference models/z_buffermodel.py:29 (ZbufferModelPts)."""
from pixelsynth_b200.models.z_buffermodel import *  # noqa: F401,F403
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts  # noqa: F401
