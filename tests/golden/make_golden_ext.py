"""Generates tests/golden/golden_ext_v1.npz: the reference's OWN arithmetic (unmodified
initLookupTable + getSamplesInternal in oracle/_ref/libdsdref.so) running on the extension filters
(decimation by 64, 128, 256: dsf2flac_b200/host/dsd_filter_ext_bits.inc), which oracle/ref_wrap.cpp
installs into a DsdDecimator constructed for ratio 32.  There is no reference parity for these
ratios -- the reference rejects them -- so what these vectors pin is "the reference's operator on
these coefficients".  Build container only (python tests/golden/make_golden_ext.py)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dsf2flac_b200 import filters  # noqa: E402
from tests.oracle_lib import DSD64, Oracle, seeded_bytes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ext_v1.npz")
CASES = {64: (8 * DSD64, 352800), 128: (8 * DSD64, 176400), 256: (8 * DSD64, 88200)}


def main():
    ref = Oracle(reference=True)
    g = {}
    # The reference itself says no to these ratios.  (Asked first: destroying a rejected DsdDecimator after
    # a valid one existed frees an uninitialised table -- the reference's file-scope
    # lookupTableAllocated flag, dsd_decimator.cpp:42,74-83.)
    for ratio, (dsd_fs, out_fs) in CASES.items():
        assert ref.info(dsd_fs, out_fs, 0) is None
    for ratio, (dsd_fs, out_fs) in CASES.items():
        fir = filters.EXTENSION_FILTERS[ratio]
        ref.set_custom_filter(fir)
        info = ref.info(dsd_fs, out_fs, 54321 * 8)
        g[f"info_{ratio}"] = np.array([info["ratio"], info["nlut"], info["nstep"], info["tzero"], info["length_pcm"]],
                                      dtype=np.int64)
        g[f"valid_{ratio}"] = np.array([info["first_valid"], info["last_valid"]])
        for msb in (0, 1):
            tag = f"r{ratio}_m{msb}"
            lut = ref.lut(dsd_fs, out_fs, msb)
            g[f"lut_sha_{tag}"] = np.frombuffer(hashlib.sha256(lut.tobytes()).digest(), dtype=np.uint8)
            nbytes = 40 * ratio
            data = seeded_bytes(2, nbytes, 2000 + ratio + msb)
            g[f"in_{tag}"] = data
            L = nbytes * 8
            g[f"app24_{tag}"] = ref.run_app(data, L, msb, dsd_fs, out_fs, 13295047.713424837, 0.0, 8388607.0)
            g[f"app16d_{tag}"] = ref.run_app(data, L, msb, dsd_fs, out_fs, 51933.78013056577, 1.0, 32767.0)
            g[f"raw64_{tag}"] = ref.run_raw(data, L, msb, dsd_fs, out_fs, 0, 48, 1.0, want="f64")
        ref.set_custom_filter(None)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
