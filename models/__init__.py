"""Reference import paths (crockwell/pixelsynth `models.*`) re-exporting the sm_100a mirrors in pixelsynth_b200.models,
so `from models.z_buffermodel import ZbufferModelPts`, `from models.base_model import BaseModel` and `python demo.py`
keep working unchanged (BASELINE.json north_star: those entry points stay drop-in)."""
