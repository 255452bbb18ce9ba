"""ctypes binding of libpixelsynth_b200.so (the C ABI declared in include/pixelsynth_b200.h).

There is no CPU fallback: if the shared library is missing the import of any op fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PS_LIB_PATH") or os.path.join(_HERE, "libpixelsynth_b200.so")  # PS_LIB_PATH: developer A/B builds

c_f = ctypes.c_float
c_d = ctypes.c_double
c_i = ctypes.c_int
c_p = ctypes.c_void_p
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); kept in the order of include/pixelsynth_b200.h
SIGNATURES = {
    "ps_abi_version": (c_i, []),
    "ps_error_string": (ctypes.c_char_p, [c_i]),
    "ps_last_error_detail": (ctypes.c_char_p, []),
    "ps_project_pts": (c_i, [c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p]),
    "ps_project_cloud": (c_i, [c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p]),
    "ps_splat_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_d]),
    "ps_splat_points": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_d, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p,
                              c_sz, c_p]),
    "ps_splat_fwd_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_d]),
    "ps_splat_fwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_d, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_p,
                           c_p, c_sz, c_p]),
    "ps_conv_igemm": (c_i, [c_p, c_p]),
    "ps_nchw_to_nhwc_bf16": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_p]),
    "ps_resample": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "ps_instance_norm_stats": (c_i, [c_p, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p]),
    "ps_classifier_input": (c_i, [c_p, ctypes.c_longlong, c_i, c_p, c_p, c_p, c_p]),
    "ps_noise_affine": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_i, c_i, c_p, c_p, c_p]),
    "ps_vq_argmin": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_p, c_p]),
    "ps_embed_codes": (c_i, [c_p, c_i, c_i, c_p, c_i, c_p, c_p]),
    "ps_combine": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "ps_tanh_residual": (c_i, [c_p, c_p, ctypes.c_longlong, c_i, c_p, c_p]),
    "ps_lmconv_glue_host": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "ps_lmconv_tc_cache_bytes": (c_sz, [c_i]),
    "ps_lmconv_levels_host": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_p]),
    "ps_lmconv_tc_run": (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_f, c_p, c_p, c_sz, c_p]),
    "ps_lmconv_tc_set_trace": (None, [c_p]),
    "ps_wedge_poll": (c_i, [c_p]),
    "ps_wedge_log": (c_i, [c_p, c_i]),
    "ps_wedge_reset": (None, []),
    "ps_sm_partition_create": (c_i, [c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_p, c_p]),
    "ps_sm_partition_stream": (c_i, [c_p, c_i, c_i, c_p]),
    "ps_sm_partition_destroy": (c_i, [c_p]),
    "ps_stream_sm_count": (c_i, [c_p]),
    "ps_launch_count": (ctypes.c_longlong, []),
    "ps_launch_count_reset": (None, []),
    "ps_timing_enable": (None, [c_i]),
    "ps_timing_collect": (c_i, [ctypes.c_char_p, ctypes.POINTER(c_d), ctypes.POINTER(c_i)]),
}

_lib = None


class PixelSynthB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PixelSynthB200Error(
                f"{LIB_PATH} is missing: build it with `python -m pixelsynth_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.ps_abi_version() != 1:
            raise PixelSynthB200Error("libpixelsynth_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        L = lib()
        raise PixelSynthB200Error(
            f"{what}: {L.ps_error_string(rc).decode()} ({rc}): {L.ps_last_error_detail().decode()}")


def check_wedge(what="pixelsynth_b200"):
    """Raises if any tensor-core kernel launched so far gave up on a barrier wait (its output is garbage).  Host-memory
    read, no device synchronisation; call it after the synchronisation that hands results to the caller."""
    info = (ctypes.c_uint * 8)()
    if lib().ps_wedge_poll(info):
        raise PixelSynthB200Error(
            f"{what}: a kernel wedged its barrier protocol (block {info[1]}, thread {info[2]}, "
            f"{'progress wait' if info[3] == 0xffffffff else 'mbarrier at shared 0x%x' % info[3]}, value {info[4]}); "
            "results since then are garbage")


def ptr(t):
    """data pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def kernel_time_ms(name=None):
    """(total ms, launches) of the kernels timed since ps_timing_enable(1); name=None sums all and resets."""
    ms, n = c_d(0.0), c_i(0)
    check(lib().ps_timing_collect(None if name is None else name.encode(), ctypes.byref(ms), ctypes.byref(n)),
          "ps_timing_collect")
    return ms.value, n.value
