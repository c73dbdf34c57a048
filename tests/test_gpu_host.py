"""The drop-in DsdDecimator (host/dsd_decimator.cpp + libdsdgpu.so) driven the way main.cpp
drives the reference class, byte-by-byte reader underneath, against the oracle's run of the
same loops.  Bit-exact."""
import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.oracle_lib import DSD64, seeded_bytes

pytestmark = pytest.mark.gpu
RATES = {32: 88200, 16: 176400, 8: 352800}


def run_app(data, L, msb, fs, scale, dither, clip, read_ahead=0, lut_bits=64, dsd_fs=DSD64):
    host = capi.host_lib()
    nch, nbytes = data.shape
    cap = L // 8 + 2048
    out = np.zeros((cap, nch), dtype=np.int32)
    n = host.dsdhost_run_app(data.ctypes.data, nch, nbytes, L, 0x69, int(msb), dsd_fs, fs, scale, dither, clip, lut_bits,
                             read_ahead, out.ctypes.data, cap)
    assert n >= 0, host.dsdhost_last_error()
    return out[:n]


@pytest.mark.parametrize("msb", [0, 1])
@pytest.mark.parametrize("ratio", [32, 16, 8])
def test_dropin_matches_oracle_app_loop(oracle, ratio, msb):
    data = seeded_bytes(2, 120000, ratio + msb)
    for L, bits, dither, ra in ((120000 * 8, 24, False, 0), (120000 * 8 - 9, 16, True, 10000), (100000 * 8, 20, True, 1)):
        scale, amp, clip = filters.app_params(bits, dither=dither)
        if ra == 1:
            data_, L_ = np.ascontiguousarray(data[:, :3000]), 3000 * 8       # one GPU round trip per frame: keep it short
        else:
            data_, L_ = data, L
        want = oracle.run_app(data_, L_, msb, DSD64, RATES[ratio], scale, amp, clip)
        got = run_app(data_, L_, msb, RATES[ratio], scale, amp, clip, read_ahead=ra)
        assert np.array_equal(got, want), (L, bits, dither, ra)


def test_dropin_dsd128_six_channels(oracle):
    data = seeded_bytes(6, 60000, 66)
    scale, amp, clip = filters.app_params(24, dither=True)
    want = oracle.run_app(data, 60000 * 8, 0, 2 * DSD64, 176400, scale, amp, clip)
    got = run_app(data, 60000 * 8, 0, 176400, scale, amp, clip, dsd_fs=2 * DSD64)
    assert np.array_equal(got, want)


def test_dropin_fp32_tables(oracle):
    data = seeded_bytes(2, 50000, 12)
    scale, _, clip = filters.app_params(24)
    want = oracle.run_app(data, 50000 * 8, 1, DSD64, 88200, scale, 0.0, clip)
    got = run_app(data, 50000 * 8, 1, 88200, scale, 0.0, clip, lut_bits=32)
    assert np.abs(got.astype(np.int64) - want).max() <= 1


def test_dropin_bulk_source_equals_stepped_reader(oracle):
    """DsdDecimator::setBulkSource: same PCM as draining the reader byte by byte (and as the
    oracle), with dither, for a length that is not a multiple of anything convenient."""
    host = capi.host_lib()
    nbytes, L = 300_000, 300_000 * 8 - 13
    data = seeded_bytes(2, nbytes, 99)
    scale, amp, clip = filters.app_params(24, dither=True)
    want = oracle.run_app(data, L, 1, DSD64, 88200, scale, amp, clip)
    for fn in (host.dsdhost_run_app, host.dsdhost_run_app_bulk):
        out = np.zeros((len(want) + 16, 2), dtype=np.int32)
        n = fn(data.ctypes.data, 2, nbytes, L, 0x69, 1, DSD64, 88200, scale, amp, clip, 64, 50_000, out.ctypes.data, len(out))
        assert n == len(want) and np.array_equal(out[:n], want)


def test_read_ahead_of_bulk_sources_on_gpu(oracle):
    """DsdDecimator over a bulk source keeps the next block in flight on a slot of its own; a scripted caller
    that carries on, changes parameters, seeks and switches table precision gets the stepped reader's samples."""
    from tests.test_host_logic import run_script
    host = capi.host_lib()
    fir = filters.pick(DSD64, 88200)
    data = seeded_bytes(2, 400000, 77)
    L = 400000 * 8 - 6
    sa, amp, clip = filters.app_params(24, dither=True)
    ops = [(4, 20000), (0, fir.nlut + 1), (1, 1024), (1, 1), (1, 30000), (2, 500), (1, 10), (0, 3), (1, 25000), (0, 1), (1, 5),
           (4, 1700), (1, 3000), (3, 32), (1, 2000), (3, 64), (1, 2000), (4, 50000), (1, 30000)]
    want = run_script(host, data, L, 88200, False, ops, sa, sa * 0.5, amp, clip)
    got = run_script(host, data, L, 88200, True, ops, sa, sa * 0.5, amp, clip)
    assert np.array_equal(got, want)
    first = oracle.run_raw(data, L, 1, DSD64, 88200, fir.nlut + 1, 31025, sa, amp, clip)
    assert np.array_equal(got[:31025], first)


@pytest.mark.gpu
def test_gpu_precision_change_after_bulk_source_keeps_the_layout(oracle):
    """The same on the GPU (tests/test_host_logic.py has the stand-in version): 64 -> 32 -> 64 bit tables after a DSF
    block-planar bulk source has been set; the recreated context must address the blocks, not planar rows."""
    from tests.test_host_logic import test_precision_change_after_bulk_source_keeps_the_layout as check
    check(capi.host_lib(), oracle)
