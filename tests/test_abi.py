"""The C-ABI library loads on a CPU box, exports every symbol include/dsdgpu.h declares, the
ctypes binding covers each of them, and without a GPU the path fails loudly (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from dsf2flac_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "dsdgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsdgpu_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 13
    out = subprocess.run(["nm", "-D", "--defined-only", capi.GPU_LIB], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dsdgpu_\w+)", out))
    assert set(names) <= exported, set(names) - exported
    assert set(names) == set(capi.GPU_PROTOTYPES), set(names) ^ set(capi.GPU_PROTOTYPES)
    lib = capi.gpu_lib()
    for n in names:
        assert getattr(lib, n) is not None


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.GPU_LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_product():
    # the product libraries must not link or name anything under oracle/
    for lib in (capi.GPU_LIB, capi.HOST_LIB):
        deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in deps and "dsdref" not in deps
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dsf2flac_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".inc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "libdsdoracle" not in text and "libdsdref" not in text, f


def test_create_argument_errors():
    lib = capi.gpu_lib()
    h = C.c_void_p()
    lut = np.zeros((72, 256))
    p = lut.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.dsdgpu_create(C.byref(h), 0, 0, 72, 4, p, 64) == 1          # nch < 1
    assert lib.dsdgpu_create(C.byref(h), 0, 2, 72, 4, p, 16) == 1          # precision
    assert lib.dsdgpu_create(C.byref(h), 0, 2, 72, 4, None, 64) == 1       # no tables
    assert b"" != lib.dsdgpu_last_error(None)
    lib.dsdgpu_destroy(None)                                               # allowed


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(capi.DsdGpuError, match="no CUDA device"):
        capi.DsdGpu(np.zeros((72, 256)), 2, 4)
    # and the drop-in decimator reports it the reference's way: isValid() false + message
    host = capi.host_lib()
    data = np.zeros((1, 4096), dtype=np.uint8)
    out = np.zeros((200, 1), dtype=np.int32)
    rc = host.dsdhost_run_raw(data.ctypes.data, 1, 4096, 4096 * 8, 0x69, 1, 2822400, 88200, 72, 10, 0, 1.0, 0.0, 0.0, 64,
                              0, out.ctypes.data, None)
    assert rc == 1 and b"GPU path unavailable" in host.dsdhost_last_error()
