"""Drop-in for reference models/base_model.py:81-103 (BaseModel, evaluation path)."""
from pixelsynth_b200.models.base_model import BaseModel  # noqa: F401
