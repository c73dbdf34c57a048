"""The drop-in is a drop-in.

oracle/_ref/libdropin_ref*.so hold the REPO's host/dsd_decimator.{h,cpp} compiled against the REFERENCE's own
dsd_sample_reader.h and linked with the reference's unmodified dsf_file_reader.cpp / dsdiff_file_reader.cpp /
dsd_sample_reader.cpp / fstream_plus.cpp -- "readers untouched", the configuration BASELINE's north_star names.
oracle/_ref/libdsdref.so is the same driver source (oracle/app_driver.h = main.cpp:169-191) over the reference's
own decimator.  The PCM of the two must be identical: 12 container cases, dither off and on, and every one of
the five getSamples<T> specialisations (reference dsd_decimator.cpp:153-172).
  CPU:  libdropin_ref_cpu.so (the C ABI answered by the CPU stand-in): compile, link and host logic.
  GPU:  libdropin_ref.so     (the C ABI answered by dsf2flac_b200/libdsdgpu.so): the product path.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.dropin_file_runner import DTYPES, lib_path
from tests.oracle_lib import DSD64, REF_SO, seeded_bytes
from tests.test_containers import CASES, make_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = all(os.path.exists(lib_path(k)) for k in ("ref", "cpu"))
HAVE_GPU_LIB = HAVE and os.path.exists(lib_path("dropin"))
TYPE_NAMES = ["int16", "int32", "int64", "float32", "float64"]


def file_run(which, path, kind, tmp_path, bits, dither, typ=1, read_ahead=0, out_fs=88200):
    scale, amp, clip = filters.app_params(bits, dither=bool(dither))
    dst = str(tmp_path / f"{which}_{typ}_{int(dither)}.npz")
    r = subprocess.run([sys.executable, "-m", "tests.dropin_file_runner", which, path, str(kind), str(out_fs), repr(scale),
                        repr(amp), repr(clip), str(typ), str(read_ahead), dst], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    z = np.load(dst)
    assert int(z["frames"]) >= 0, f"{which}: {int(z['frames'])} {z['error']}"
    return z


def check_file_cases(which, tmp_path, tag, kind, nbytes, nch, kw):
    path, _ = make_file(tmp_path, tag, kind, nbytes, nch, kw)
    for dither in (0, 1):
        ref = file_run("ref", path, kind, tmp_path, 24, dither)
        got = file_run(which, path, kind, tmp_path, 24, dither, read_ahead=(0 if dither else 777))
        assert int(got["frames"]) == int(ref["frames"]) and int(ref["frames"]) > 0
        assert [int(got[k]) for k in ("channels", "fs", "msb", "length")] == [int(ref[k]) for k in ("channels", "fs", "msb", "length")]
        assert np.array_equal(got["pcm"], ref["pcm"])


@pytest.mark.skipif(not HAVE, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES, ids=[c[0] for c in CASES])
def test_cpu_dropin_over_reference_readers(tmp_path, tag, kind, nbytes, nch, kw):
    check_file_cases("cpu", tmp_path, tag, kind, nbytes, nch, kw)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_GPU_LIB, reason="needs oracle/_ref/libdropin_ref.so (built where /root/reference exists)")
@pytest.mark.parametrize("tag,kind,nbytes,nch,kw", CASES, ids=[c[0] for c in CASES])
def test_gpu_dropin_over_reference_readers(tmp_path, tag, kind, nbytes, nch, kw):
    check_file_cases("dropin", tmp_path, tag, kind, nbytes, nch, kw)


def typed_file_check(which, tmp_path):
    """getSamples<T> for all five T through the file path: DSF stereo, 16-bit parameters (so int16 is in range)."""
    path, _ = make_file(tmp_path, "typed", 1, 9000, 2, {})
    for typ in range(5):
        for dither in (0, 1):
            ref = file_run("ref", path, 1, tmp_path, 16, dither, typ)
            got = file_run(which, path, 1, tmp_path, 16, dither, typ, read_ahead=500)
            assert got["pcm"].dtype == DTYPES[typ] and got["pcm"].shape == ref["pcm"].shape and len(ref["pcm"]) > 2000
            assert np.array_equal(got["pcm"], ref["pcm"]), TYPE_NAMES[typ]


@pytest.mark.skipif(not HAVE, reason="needs oracle/_ref")
def test_cpu_dropin_all_sample_types_files(tmp_path):
    typed_file_check("cpu", tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_GPU_LIB, reason="needs oracle/_ref/libdropin_ref.so")
def test_gpu_dropin_all_sample_types_files(tmp_path):
    typed_file_check("dropin", tmp_path)


# ---- raw getSamples<T> calls over an in-memory reader: reference class vs the repo's class, three builds ----------

_RAW_ARGS = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_uint8, C.c_int, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64,
             C.c_int64, C.c_double, C.c_double, C.c_double]


def ref_raw_typed(data, L, msb, fs, out_fs, pre, nframes, chunk, scale, amp, clip, typ):
    lib = C.CDLL(REF_SO)
    f = lib.dsdref_run_raw_typed
    f.restype, f.argtypes = C.c_int, _RAW_ARGS + [C.c_int, C.c_int, C.c_void_p]
    out = np.zeros((nframes, data.shape[0]), dtype=DTYPES[typ])
    assert f(data.ctypes.data, data.shape[0], data.shape[1], L, 0x69, msb, fs, out_fs, pre, nframes, chunk, scale, amp, clip,
             1, typ, out.ctypes.data) == 0
    return out


def dropin_raw_typed(which, data, L, msb, fs, out_fs, pre, nframes, chunk, scale, amp, clip, typ, read_ahead=300):
    lib = C.CDLL(lib_path(which))
    f = lib.dropin_run_raw_typed
    f.restype, f.argtypes = C.c_int, _RAW_ARGS + [C.c_uint32, C.c_int, C.c_void_p]
    out = np.zeros((nframes, data.shape[0]), dtype=DTYPES[typ])
    assert f(data.ctypes.data, data.shape[0], data.shape[1], L, 0x69, msb, fs, out_fs, pre, nframes, chunk, scale, amp, clip,
             read_ahead, typ, out.ctypes.data) == 0
    return out


def host_raw_typed(host, data, L, msb, fs, out_fs, pre, nframes, chunk, scale, amp, clip, typ, read_ahead=300):
    out = np.zeros((nframes, data.shape[0]), dtype=DTYPES[typ])
    rc = host.dsdhost_run_raw_typed(data.ctypes.data, data.shape[0], data.shape[1], L, 0x69, msb, fs, out_fs, pre, nframes,
                                    chunk, scale, amp, clip, 64, read_ahead, typ, out.ctypes.data)
    assert rc == 0, host.dsdhost_last_error()
    return out


RAW_CASES = [(88200, 3, 1, 0), (176400, 2, 0, 1), (352800, 1, 1, 1), (88200, 2, 0, 0)]   # out_fs, nch, msb, dither


def raw_typed_check(run):
    for out_fs, nch, msb, dither in RAW_CASES:
        data = seeded_bytes(nch, 4000, out_fs + nch)
        ratio = DSD64 // out_fs
        nframes = (4000 - 200) * 8 // ratio
        for typ in range(5):
            bits = 16 if typ == 0 else 24
            scale, amp, clip = filters.app_params(bits, dither=bool(dither))
            want = ref_raw_typed(data, 4000 * 8, msb, DSD64, out_fs, 100, nframes, 257, scale, amp, clip, typ)
            got = run(data, 4000 * 8, msb, DSD64, out_fs, 100, nframes, 257, scale, amp, clip, typ)
            assert np.array_equal(got, want), (out_fs, nch, TYPE_NAMES[typ])
            assert np.abs(want.astype(np.float64)).max() > 100


@pytest.mark.skipif(not HAVE, reason="needs oracle/_ref")
def test_cpu_all_sample_types_raw():
    """int16 / int32 / int64 / float32 / float64 (dsd_decimator.cpp:153-172): the repo's class over the reference's reader
    header (stand-in ABI) and the standalone host library (stand-in ABI) against the reference's class."""
    raw_typed_check(lambda *a: dropin_raw_typed("cpu", *a))
    from tests.standin import cpu_host_lib
    host = cpu_host_lib()
    raw_typed_check(lambda *a: host_raw_typed(host, *a))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_SO), reason="needs oracle/_ref")
def test_gpu_all_sample_types_raw():
    host = capi.host_lib()
    raw_typed_check(lambda *a: host_raw_typed(host, *a))
    if HAVE_GPU_LIB:
        raw_typed_check(lambda *a: dropin_raw_typed("dropin", *a))
