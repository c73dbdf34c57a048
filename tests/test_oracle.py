"""The CPU oracle (oracle/dsd_oracle.c) against (1) the committed golden vectors that
tests/golden/make_golden.py produced from the unmodified reference, and (2) the reference
itself (oracle/_ref) wherever it has been built.  Bit-exact everywhere."""
import hashlib
import os

import numpy as np
import pytest

from dsf2flac_b200 import filters
from tests.oracle_lib import DSD64, seeded_bytes

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
RATES = {32: 88200, 16: 176400, 8: 352800}
CASES = [(r, m) for r in RATES for m in (0, 1)]
S24, C24 = 13295047.713424837, 8388607.0


def same_bits(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def test_rand_matches_glibc_golden(port):
    assert np.array_equal(port.rand(256), GOLD["rand_first_256"])
    # positional start: skipping k draws lands on the same stream
    assert np.array_equal(port.rand(56, skip=200), GOLD["rand_first_256"][200:])


@pytest.mark.parametrize("ratio", list(RATES))
def test_geometry_golden(port, ratio):
    info = port.info(DSD64, RATES[ratio], 12345 * 8)
    got = [info["ratio"], info["nlut"], info["nstep"], info["tzero"], info["length_pcm"]]
    assert got == list(GOLD[f"info_{ratio}"])
    assert [info["first_valid"], info["last_valid"]] == list(GOLD[f"valid_{ratio}"])
    fir = filters.pick(DSD64, RATES[ratio])
    assert (fir.ratio, fir.nlut, fir.nstep, fir.tzero) == tuple(got[:4])


def test_rejected_rates(port):
    # dsd_decimator.cpp:57-68: only /8, /16, /32; DSD512 -> 352.8k is /64
    assert port.info(DSD64 * 8, 352800) is None
    assert port.info(DSD64, 44100) is None
    assert filters.pick(DSD64 * 8, 352800) is None
    assert port.info(DSD64 * 8, 705600)["ratio"] == 32


@pytest.mark.parametrize("ratio,msb", CASES)
def test_tables_golden(port, ratio, msb):
    tag = f"r{ratio}_m{msb}"
    lut = port.lut(DSD64, RATES[ratio], msb)
    assert hashlib.sha256(lut.tobytes()).digest() == GOLD[f"lut_sha_{tag}"].tobytes()
    assert same_bits(lut[[0, lut.shape[0] // 2, lut.shape[0] - 1]], GOLD[f"lut_rows_{tag}"])
    # the host-side builder the GPU path is fed from makes the same tables
    assert same_bits(filters.build_lut(filters.pick(DSD64, RATES[ratio]), msb), lut)


@pytest.mark.parametrize("ratio,msb", CASES)
def test_operator_golden(port, ratio, msb):
    tag = f"r{ratio}_m{msb}"
    data, L, fs = GOLD[f"in_{tag}"], 2048 * 8, RATES[ratio]
    assert np.array_equal(port.run_app(data, L, msb, DSD64, fs, S24, 0.0, C24), GOLD[f"app24_{tag}"])
    assert np.array_equal(port.run_app(data, L, msb, DSD64, fs, 51933.78013056577, 1.0, 32767.0), GOLD[f"app16d_{tag}"])
    assert same_bits(port.run_raw(data, L, msb, DSD64, fs, 0, 96, 1.0, want="f64"), GOLD[f"raw64_{tag}"])
    got = port.run_raw(data[:, :1001], 1001 * 8 - 3, msb, DSD64, fs, 900, 200, 830940.4820890523, 0.0, 524287.0,
                         idle=0x96)
    assert np.array_equal(got, GOLD[f"tail20_{tag}"])


def test_sigma_delta_stream_golden(port):
    for lsb_first in (1, 0):
        got = port.run_app(GOLD[f"sine_in_lsb{lsb_first}"], 1536 * 8, lsb_first, DSD64, 88200, S24, 0.0, C24)
        assert np.array_equal(got, GOLD[f"sine_app24_lsb{lsb_first}"])
    # the sine really is in there: 1 kHz at 0.5 amplitude, 4 dB gain
    x = GOLD["sine_app24_lsb1"][:, 0].astype(np.float64) / 2 ** 23
    assert 0.7 < np.abs(x).max() < 0.85


def test_frame_count_formula(port):
    for ratio, fs in RATES.items():
        fir = filters.pick(DSD64, fs)
        for nbytes in (fir.nlut * 2, 1000, 1003, 4096):
            data = seeded_bytes(1, nbytes, nbytes)
            assert len(port.run_app(data, nbytes * 8, 1, DSD64, fs, 1.0)) == filters.app_frames(nbytes * 8, fir)


# ---- against the reference itself (skipped where oracle/_ref is not built) --------------------

@pytest.mark.parametrize("ratio,msb", CASES)
def test_oracle_equals_reference(port, ref, ratio, msb):
    fs = RATES[ratio]
    assert same_bits(port.lut(DSD64, fs, msb), ref.lut(DSD64, fs, msb))
    data = seeded_bytes(3, 5000, 77 + ratio + msb)
    for L in (5000 * 8, 5000 * 8 - 13, 4000 * 8):
        for bits, dither in ((24, 0.0), (16, 1.0), (20, 1.0)):
            scale, _, clip = filters.app_params(bits)
            a = port.run_app(data, L, msb, DSD64, fs, scale, dither, clip)
            b = ref.run_app(data, L, msb, DSD64, fs, scale, dither, clip)
            assert np.array_equal(a, b)
    # raw operator over the start-up idle region, unclipped fp64 and a hot scale that clips
    a = port.run_raw(data, 5000 * 8, msb, DSD64, fs, 0, 300, 3.0, want="f64")
    b = ref.run_raw(data, 5000 * 8, msb, DSD64, fs, 0, 300, 3.0, want="f64")
    assert same_bits(a, b)
    a = port.run_raw(data, 5000 * 8, msb, DSD64, fs, 10, 300, 8e7, 0.0, 8388607.0)
    b = ref.run_raw(data, 5000 * 8, msb, DSD64, fs, 10, 300, 8e7, 0.0, 8388607.0)
    assert np.array_equal(a, b) and np.abs(a).max() == 8388607


def test_oracle_rand_equals_libc(port, ref):
    assert np.array_equal(port.rand(100000), ref.rand(100000))


def test_one_fma_draw_quotient_sampled(port):
    """dsd_rand.cuh turns a draw into draw / RAND_MAX with one FMA; tools/check_draw_quotient.c compares all 2^31 values
    with the reference's division (0 mismatches).  Here: both ends of the range and a stride through the middle."""
    import ctypes as C
    f = port.lib.dsdoracle_check_draw_quotient
    f.restype, f.argtypes = C.c_longlong, [C.c_longlong] * 3
    assert f(0, 1, 1 << 22) == 0
    assert f((1 << 31) - (1 << 22), 1, 1 << 22) == 0
    assert f(12345, 509, 1 << 22) == 0
