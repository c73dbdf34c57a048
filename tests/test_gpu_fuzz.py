"""Seeded random geometry against the oracle: channel counts 1..8, all three filters, both bit
orders, every kernel that accepts the case, all three source layouts, windows that start before /
end after the data, odd chunk sizes, dither on and off, 64-bit and 32-bit tables.  Small streams
(the oracle finishes each in milliseconds), many of them."""
import numpy as np
import pytest

from dsf2flac_b200 import capi, filters
from tests.oracle_lib import DSD64
from tests.test_gpu_parity import RATES, to_blocks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(48))
def test_random_case_equals_oracle(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    ratio = int(rng.choice([32, 32, 16, 8]))
    fir = filters.FILTERS[ratio]
    msb = int(rng.integers(0, 2))
    nch = int(rng.integers(1, 9))
    nbytes = int(rng.integers(200, 60000))
    data = rng.integers(0, 256, size=(nch, nbytes), dtype=np.uint8)
    first = int(rng.integers(0, max(1, nbytes // 3))) if rng.random() < 0.5 else 0
    layout = int(rng.choice([0, 0, 1, 64, 4096]))
    if layout >= 16:
        first -= first % layout
    count = int(rng.integers(1, nbytes - first + 1))
    p0 = int(rng.integers(-5, max(1, nbytes // 2)))
    nframes = int(rng.integers(0, (nbytes - p0 + 500) // fir.nstep))
    dither = bool(rng.integers(0, 2))
    bits = int(rng.choice([16, 20, 24]))
    scale, amp, clip = filters.app_params(bits, dither=dither)
    precision = 32 if rng.random() < 0.25 else 64
    fill = int(rng.choice([0x69, 0x96, 0x00]))
    kernels = ["auto", "direct", "seg"]
    kernel = str(rng.choice(kernels))
    sample0 = int(rng.integers(0, 10_000_000)) if dither else 0

    masked = np.full_like(data, fill)
    masked[:, first:first + count] = data[:, first:first + count]
    # the oracle's reader delivers `fill` before the stream and after it; its dither stream is positional
    want = oracle.run_raw(masked, 1 << 40, msb, DSD64, RATES[ratio], p0 + 1, nframes, scale, amp, clip, idle=fill,
                          rand_skip=2 * sample0)
    window = data[:, first:first + count]
    if layout == 0:
        src, stride = np.ascontiguousarray(window), count
    elif layout == 1:
        src, stride = np.ascontiguousarray(window.T).reshape(-1), 0
    else:
        src, stride = to_blocks(window, layout), 0
    g = capi.DsdGpu(filters.build_lut(fir, msb), nch, fir.nstep, lut_precision=precision)
    g.set_option("kernel", {"auto": capi.KERNEL_AUTO, "direct": capi.KERNEL_DIRECT, "seg": capi.KERNEL_SEGMENTED}[kernel])
    g.set_option("source_block_bytes", layout)
    g.set_option("chunk_frames", int(rng.integers(1, 5000)))
    got = np.empty((nframes, nch), dtype=np.int32)
    g.decimate_host_raw(src.ctypes.data, stride, first, count, fill, p0, nframes, scale, amp, clip, sample0, capi.OUT_INT32,
                        got.ctypes.data)
    g.close()
    what = dict(ratio=ratio, msb=msb, nch=nch, nbytes=nbytes, first=first, count=count, layout=layout, p0=p0, nframes=nframes,
                dither=dither, bits=bits, precision=precision, kernel=kernel)
    if precision == 64:
        assert np.array_equal(got, want), what
    else:
        assert np.abs(got.astype(np.int64) - want).max(initial=0) <= 1, what
