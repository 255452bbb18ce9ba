"""CPU: libpixelsynth_b200.so loads and exports every symbol include/pixelsynth_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

from util import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pixelsynth_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("ps_splat_fwd", "ps_splat_points", "ps_project_pts", "ps_project_cloud", "ps_abi_version"):
        assert s in syms


def test_library_exports_header_symbols():
    import pixelsynth_b200.build as b

    so = b.build()
    lib = ctypes.CDLL(so)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    lib.ps_abi_version.restype = ctypes.c_int
    assert lib.ps_abi_version() == 1
    lib.ps_error_string.restype = ctypes.c_char_p
    assert lib.ps_error_string(-3) == b"workspace too small"


def test_binding_table_covers_header():
    from pixelsynth_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_invalid_arguments_are_rejected_without_a_gpu():
    from pixelsynth_b200 import _lib

    L = _lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert L.ps_project_pts(None, None, 1, 8, 0.01, None, None, None) == -1
    assert L.ps_splat_points(None, None, 1, 8, 3, 8, 4, 4.0, 1.0, 2, 0, 13, None, None, None, None, None, None, 0,
                             None) == -1
    assert b"argument check failed" in L.ps_last_error_detail()
    assert L.ps_splat_workspace_bytes(1, 65536, 256, 4.0) > 65536 * 4
    # SM partitions: argument errors before any driver call; a NULL handle is a no-op to destroy
    h = ctypes.c_void_p()
    assert L.ps_sm_partition_create(0, 0, 0, None, 0, None, None, None, ctypes.byref(h)) == -1
    assert L.ps_sm_partition_create(0, 24, 1, None, 0, None, None, None, ctypes.byref(h)) == -1   # stream array missing
    assert L.ps_sm_partition_stream(None, 1, 1, ctypes.byref(h)) == -1
    assert L.ps_sm_partition_destroy(None) == 0


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import pixelsynth_b200.ops  # noqa: F401

    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.pixelsynth_b200.splat_points(torch.zeros(1, 4, 3), torch.zeros(1, 3, 4), 8, 4, 2.0, 1.0, 2, 0, 13,
                                               False, False)
