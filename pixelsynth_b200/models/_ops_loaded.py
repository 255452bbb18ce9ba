from .. import ops  # noqa: F401  importing registers torch.ops.pixelsynth_b200.*
