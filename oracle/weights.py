"""Re-export of the seeded weight factory (it lives in the product: pixelsynth_b200/synthetic.py, because smoke
runs and the bench need random-initialised reference-shaped weights too)."""
from pixelsynth_b200.synthetic import make_state, shapes, synth_image  # noqa: F401
