"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference decimator
(oracle/_ref/libdsdref.so, built by oracle/Makefile from /root/reference/src).

Run in the build container only (python tests/golden/make_golden.py): /root/reference does not
exist on the GPU box, the committed vectors are what travels.  The reference ships no golden
vectors of its own, so these pin the oracle to the reference's outputs.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.oracle_lib import DSD64, Oracle, seeded_bytes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
RATES = {32: 88200, 16: 176400, 8: 352800}


def sdm_sine(nbytes, freq, lsb_first, fs=DSD64, amp=0.5):
    """2nd-order sigma-delta sine, numpy restatement of host/dsd_synth.cpp (small sizes only)."""
    n = nbytes * 8
    x = amp * np.sin(2 * np.pi * freq * np.arange(n) / fs)
    s1 = s2 = 0.0
    bits = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        v = 1.0 if s2 >= 0 else -1.0
        s1 += x[i] - v
        s2 += s1 - v
        bits[i] = v > 0
    return np.packbits(bits, bitorder="little" if lsb_first else "big")


def main():
    ref = Oracle(reference=True)
    g = {}
    g["rand_first_256"] = ref.rand(256)
    for ratio, out_fs in RATES.items():
        info = ref.info(DSD64, out_fs, 12345 * 8)
        g[f"info_{ratio}"] = np.array([info["ratio"], info["nlut"], info["nstep"], info["tzero"], info["length_pcm"]],
                                      dtype=np.int64)
        g[f"valid_{ratio}"] = np.array([info["first_valid"], info["last_valid"]])
        for msb in (0, 1):
            tag = f"r{ratio}_m{msb}"
            lut = ref.lut(DSD64, out_fs, msb)
            g[f"lut_sha_{tag}"] = np.frombuffer(hashlib.sha256(lut.tobytes()).digest(), dtype=np.uint8)
            g[f"lut_rows_{tag}"] = lut[[0, lut.shape[0] // 2, lut.shape[0] - 1]]
            data = seeded_bytes(2, 2048, 1000 + ratio + msb)
            g[f"in_{tag}"] = data
            L = 2048 * 8
            # whole-track run, 24 bit, 4 dB, clip on, dither off / on (fresh-process rand stream)
            g[f"app24_{tag}"] = ref.run_app(data, L, msb, DSD64, out_fs, 13295047.713424837, 0.0, 8388607.0)
            g[f"app16d_{tag}"] = ref.run_app(data, L, msb, DSD64, out_fs, 51933.78013056577, 1.0, 32767.0)
            # raw operator straight after rewind: reads the 0x69 pre-fill; unscaled fp64 view
            g[f"raw64_{tag}"] = ref.run_raw(data, L, msb, DSD64, out_fs, 0, 96, 1.0, want="f64")
            # length not a multiple of the ratio, odd idle byte, past-the-end region, 20 bit
            g[f"tail20_{tag}"] = ref.run_raw(data[:, :1001], 1001 * 8 - 3, msb, DSD64, out_fs, 900, 200,
                                             830940.4820890523, 0.0, 524287.0, idle=0x96)
    # a real sigma-delta stream (same bits, both packings must agree: SURVEY 8a)
    for lsb_first in (1, 0):
        sine = np.stack([sdm_sine(1536, 1000.0, lsb_first), sdm_sine(1536, 1500.0, lsb_first)])
        g[f"sine_in_lsb{lsb_first}"] = sine
        g[f"sine_app24_lsb{lsb_first}"] = ref.run_app(sine, 1536 * 8, lsb_first, DSD64, 88200, 13295047.713424837, 0.0,
                                                      8388607.0)
    assert np.array_equal(g["sine_app24_lsb1"], g["sine_app24_lsb0"])
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
