"""pixelsynth_b200: B200-native (sm_100a) kernels for the PixelSynth novel-view-synthesis inference hot path.

Layout: csrc/ (CUDA kernels + C ABI), _lib.py (ctypes binding), ops.py (torch.ops.pixelsynth_b200.*),
models/ (host-side mirror of the reference's module interface: same class / method names and argument
meaning as crockwell/pixelsynth `models.*`, with the hot callees dispatching to the ops).
"""
__version__ = "0.1.0"
