"""Host logic without a GPU: the drop-in DsdDecimator (host/dsd_decimator.cpp) linked against a
CPU stand-in of the C ABI (tests/cpp/standin_dsdgpu.cpp) must behave, call for call, like the
reference class: logical position, creep/step, block read-ahead of any size, getSamples call
sizes, dither stream continuity, loop termination and frame counts."""
import ctypes as C
import os

import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.oracle_lib import DSD64, seeded_bytes

CPU_HOST = os.path.join(os.path.dirname(__file__), "cpp", "_build", "libdsdhost_cpu.so")
RATES = {32: 88200, 16: 176400, 8: 352800}


@pytest.fixture(scope="module")
def host():
    return capi.host_lib(CPU_HOST)


def run_app(host, data, L, msb, fs, scale, dither, clip, read_ahead=0, lut_bits=64):
    nch, nbytes = data.shape
    cap = L // 8 + 2048
    out = np.zeros((cap, nch), dtype=np.int32)
    n = host.dsdhost_run_app(data.ctypes.data, nch, nbytes, L, 0x69, int(msb), DSD64, fs, scale, dither, clip, lut_bits,
                             read_ahead, out.ctypes.data, cap)
    assert n >= 0, host.dsdhost_last_error()
    return out[:n]


def run_raw(host, data, L, msb, fs, pre, nframes, chunk, scale, dither, clip, read_ahead=0, want="i32"):
    nch, nbytes = data.shape
    out = np.zeros((nframes, nch), dtype=np.int32 if want == "i32" else np.float64)
    rc = host.dsdhost_run_raw(data.ctypes.data, nch, nbytes, L, 0x69, int(msb), DSD64, fs, pre, nframes, chunk, scale,
                              dither, clip, 64, read_ahead, out.ctypes.data if want == "i32" else None,
                              out.ctypes.data if want != "i32" else None)
    assert rc == 0, host.dsdhost_last_error()
    return out


def test_info_matches_oracle(host, oracle):
    for ratio, fs in RATES.items():
        r, a, b, n = C.c_int32(), C.c_double(), C.c_double(), C.c_int64()
        assert host.dsdhost_info(DSD64, fs, 999 * 8, C.byref(r), C.byref(a), C.byref(b), C.byref(n)) == 1
        info = oracle.info(DSD64, fs, 999 * 8)
        assert (r.value, a.value, b.value, n.value) == (info["ratio"], info["first_valid"], info["last_valid"],
                                                        info["length_pcm"])
    assert host.dsdhost_info(DSD64 * 8, 352800, 0, C.byref(r), C.byref(a), C.byref(b), C.byref(n)) == 0
    assert b"incompatible sample rate" in host.dsdhost_last_error()


@pytest.mark.parametrize("ratio", list(RATES))
@pytest.mark.parametrize("read_ahead", [1, 7, 1024, 5000, 0])
def test_app_loop_any_read_ahead(host, oracle, ratio, read_ahead):
    fs = RATES[ratio]
    data = seeded_bytes(2, 6000, ratio)
    for L in (6000 * 8, 6000 * 8 - 5, 5000 * 8):
        scale, amp, clip = filters.app_params(24, dither=True)
        want = oracle.run_app(data, L, 1, DSD64, fs, scale, amp, clip)
        got = run_app(host, data, L, 1, fs, scale, amp, clip, read_ahead)
        assert np.array_equal(got, want)


def test_call_sizes_and_steps_interleaved(host, oracle):
    # getSamples in odd-sized calls after an odd creep: the dither stream and the position carry on
    data = seeded_bytes(3, 4000, 9)
    scale, amp, clip = filters.app_params(16, dither=True)
    for chunk in (1, 3, 100, 0):
        want = oracle.run_raw(data, 4000 * 8, 0, DSD64, 88200, 75, 700, scale, amp, clip)
        got = run_raw(host, data, 4000 * 8, 0, 88200, 75, 700, chunk, scale, amp, clip, read_ahead=64)
        assert np.array_equal(got, want)
    # float output is the unrounded value
    want = oracle.run_raw(data, 4000 * 8, 0, DSD64, 176400, 0, 300, 2.5, want="f64")
    got = run_raw(host, data, 4000 * 8, 0, 176400, 0, 300, 7, 2.5, 0.0, 0.0, want="f64")
    assert want.tobytes() == got.tobytes()


def test_fp32_tables_within_one_lsb(host, oracle):
    data = seeded_bytes(2, 5000, 5)
    scale, _, clip = filters.app_params(24)
    want = oracle.run_app(data, 5000 * 8, 1, DSD64, 88200, scale, 0.0, clip)
    got = run_app(host, data, 5000 * 8, 1, 88200, scale, 0.0, clip, lut_bits=32)
    assert np.abs(got.astype(np.int64) - want).max() <= 1


def test_bulk_source_equals_stepped_reader(host, oracle):
    nbytes, L = 40_000, 40_000 * 8 - 13
    data = seeded_bytes(3, nbytes, 5)
    scale, amp, clip = filters.app_params(16, dither=True)
    want = oracle.run_app(data, L, 0, DSD64, 176400, scale, amp, clip)
    for fn in (host.dsdhost_run_app, host.dsdhost_run_app_bulk):
        out = np.zeros((len(want) + 16, 3), dtype=np.int32)
        n = fn(data.ctypes.data, 3, nbytes, L, 0x69, 0, DSD64, 176400, scale, amp, clip, 64, 4000, out.ctypes.data, len(out))
        assert n == len(want) and np.array_equal(out[:n], want)


def run_script(host, data, L, fs, bulk, ops, scale_a, scale_b, dither, clip, dsd_fs=DSD64):
    nch, nbytes = data.shape
    flat = (C.c_int64 * (2 * len(ops)))(*[int(x) for pair in ops for x in pair])
    cap = sum(a for w, a in ops if w in (1, 2)) + 8
    out = np.zeros((cap, nch), dtype=np.int32)
    n = host.dsdhost_run_script(data.ctypes.data, nch, nbytes, L, 0x69, 1, dsd_fs, fs, int(bulk), flat, len(ops), scale_a,
                                scale_b, dither, clip, out.ctypes.data, cap)
    assert n >= 0, host.dsdhost_last_error()
    return out[:n]


@pytest.mark.parametrize("ratio", list(RATES))
def test_read_ahead_of_bulk_sources_is_invisible(host, oracle, ratio):
    """With a bulk source the class submits the next block while the caller consumes the current one (the
    stand-in defers that work until the block is claimed).  Whatever the caller does next -- carries on,
    changes the scale, skips bytes with step(), switches table precision, shrinks the read-ahead -- the
    samples are those of the stepped reader path, which other tests hold to the oracle."""
    fs = RATES[ratio]
    data = seeded_bytes(2, 9000, 40 + ratio)
    L = 9000 * 8 - 6
    sa, amp, clip = filters.app_params(24, dither=True)
    sb = sa * 0.5
    fir = filters.pick(DSD64, fs)
    ops = [(4, 300), (0, fir.nlut + 1), (1, 1024), (1, 1), (1, 700),        # carries on across block boundaries
           (2, 500), (2, 64), (1, 10),                                       # scale changes: block in flight is dropped
           (0, 3), (1, 400), (0, 1), (1, 5),                                 # seeks: position no longer matches
           (4, 17), (1, 300), (3, 32), (1, 200), (3, 64), (1, 200),          # read-ahead shrinks; table precision switches
           (4, 5000), (1, 3000)]                                             # runs far past the end of the data
    want = run_script(host, data, L, fs, False, ops, sa, sb, amp, clip)
    got = run_script(host, data, L, fs, True, ops, sa, sb, amp, clip)
    assert np.array_equal(got, want)
    # and the plain part of it against the oracle directly
    first = oracle.run_raw(data, L, 1, DSD64, fs, fir.nlut + 1, 1725, sa, amp, clip)
    assert np.array_equal(got[:1725], first)


def test_gpu_failure_mid_stream_exits_like_the_reference(tmp_path):
    """A GPU error after construction must not turn into silence in the FLAC file: main.cpp checks isValid() once,
    right after construction (main.cpp:232-236).  The drop-in follows the reference's convention for an unrecoverable
    error inside getSamples (dsd_decimator.cpp:182-186): message on stderr, exit(EXIT_FAILURE)."""
    import subprocess
    import sys
    code = (
        "import numpy as np\n"
        "from dsf2flac_b200 import capi, filters\n"
        "from tests.oracle_lib import DSD64, seeded_bytes\n"
        f"host = capi.host_lib({CPU_HOST!r})\n"
        "data = seeded_bytes(2, 6000, 5)\n"
        "out = np.zeros((4096, 2), dtype=np.int32)\n"
        "scale, amp, clip = filters.app_params(24)\n"
        "n = host.dsdhost_run_app(data.ctypes.data, 2, 6000, 48000, 0x69, 1, DSD64, 88200, scale, amp, clip, 64, 200, out.ctypes.data, 4096)\n"
        "print('returned', n)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ok = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert ok.returncode == 0 and "returned 1483" in ok.stdout, ok.stderr
    bad = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, DSDGPU_STANDIN_FAIL_AFTER="3"))
    assert bad.returncode == 1, (bad.returncode, bad.stdout, bad.stderr)
    assert "GPU decimation failed" in bad.stderr and "injected failure" in bad.stderr and "returned" not in bad.stdout


def test_precision_change_after_bulk_source_keeps_the_layout(host, oracle):
    """setLutPrecision() recreates the GPU context; a DSF block-planar bulk source set before must keep its layout on the
    new one (else every channel reads the wrong bytes, silently)."""
    nch, block, nblocks = 2, 256, 40
    data = seeded_bytes(nch, block * nblocks, 99)
    blocks = np.ascontiguousarray(data.reshape(nch, nblocks, block).transpose(1, 0, 2)).reshape(-1)
    L = data.shape[1] * 8
    scale, amp, clip = filters.app_params(24, dither=True)
    n = 2000
    want = oracle.run_app(data, L, 1, DSD64, 88200, scale, amp, clip)[:n]
    out = np.zeros((n, nch), dtype=np.int32)
    got = host.dsdhost_run_precision_after_bulk(blocks.ctypes.data, nch, data.shape[1], block, L, 1, DSD64, 88200, 64, scale, amp, clip,
                                                n, out.ctypes.data)
    assert got == n, host.dsdhost_last_error()
    assert np.array_equal(out, want)
    got = host.dsdhost_run_precision_after_bulk(blocks.ctypes.data, nch, data.shape[1], block, L, 1, DSD64, 88200, 32, scale, amp, clip,
                                                n, out.ctypes.data)
    assert got == n and np.abs(out.astype(np.int64) - want).max() <= 1          # 32-bit tables: within one LSB
