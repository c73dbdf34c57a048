#!/usr/bin/env python
"""More of tests/test_gpu_fuzz.py than the suite runs every time, plus the same kind of random geometry for the
EXTENSION ratios (ring passes of 72 tables) -- run on a GPU box when the kernels change:
    python tools/fuzz_more.py [first_seed] [count]
Every case must equal the oracle (bit for bit with fp64 tables, within one LSB with 32-bit tables)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsf2flac_b200 import capi, filters  # noqa: E402
from tests.oracle_lib import DSD64, BestOracle  # noqa: E402
from tests import test_gpu_fuzz  # noqa: E402


def extension_case(oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    ratio = int(rng.choice([64, 128, 256]))
    fir = filters.EXTENSION_FILTERS[ratio]
    out_fs = 8 * DSD64 // ratio
    msb = int(rng.integers(0, 2))
    nch = int(rng.integers(1, 7))
    nbytes = int(rng.integers(2000, 120000))
    data = rng.integers(0, 256, size=(nch, nbytes), dtype=np.uint8)
    p0 = int(rng.integers(-5, max(1, nbytes // 2)))
    nframes = int(rng.integers(0, (nbytes - p0 + 500) // fir.nstep))
    dither = bool(rng.integers(0, 2))
    bits = int(rng.choice([16, 20, 24]))
    scale, amp, clip = filters.app_params(bits, dither=dither)
    fill = int(rng.choice([0x69, 0x96, 0x00]))
    kernel = str(rng.choice(["auto", "auto", "seg"]))
    sample0 = int(rng.integers(0, 1_000_000)) if dither and rng.random() < 0.5 else 0
    oracle.set_custom_filter(fir)
    try:
        want = oracle.run_raw(data, 1 << 40, msb, 8 * DSD64, out_fs, p0 + 1, nframes, scale, amp, clip, idle=fill,
                              rand_skip=2 * sample0)
    finally:
        oracle.set_custom_filter(None)
    g = capi.DsdGpu(filters.build_lut(fir, msb), nch, fir.nstep)
    g.set_option("kernel", {"auto": capi.KERNEL_AUTO, "seg": capi.KERNEL_SEGMENTED}[kernel])
    g.set_option("chunk_frames", int(rng.integers(1, 9000)))
    got = np.empty((nframes, nch), dtype=np.int32)
    g.decimate_host_raw(data.ctypes.data, nbytes, 0, nbytes, fill, p0, nframes, scale, amp, clip, sample0, capi.OUT_INT32,
                        got.ctypes.data)
    g.close()
    assert np.array_equal(got, want), dict(seed=seed, ratio=ratio, nch=nch, nbytes=nbytes, p0=p0, nframes=nframes,
                                           dither=dither, kernel=kernel, sample0=sample0)


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    oracle = BestOracle()
    for seed in range(first, first + count):
        test_gpu_fuzz.test_random_case_equals_oracle(oracle, seed)
    print(f"reference filters: seeds {first}..{first + count - 1} equal the oracle")
    for seed in range(first, first + count // 4):
        extension_case(oracle, seed)
    print(f"extension ratios: seeds {first}..{first + count // 4 - 1} equal the oracle")


if __name__ == "__main__":
    main()
