"""Filter geometry and lookup-table construction, host side, for tests and benchmarks.

Mirrors what the drop-in DsdDecimator (host/dsd_decimator.cpp) does before it hands the tables
to dsdgpu_create: the reference's DsdDecimator::DsdDecimator (reference src/dsd_decimator.cpp:44-72)
and initLookupTable (:117-151) on the impulse responses of filters.cpp:52-115, which this repo
carries as binary64 bit patterns in host/dsd_filter_bits.inc.
"""
import os
import re
from dataclasses import dataclass

import numpy as np

_INC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "dsd_filter_bits.inc")
_EXT_INC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "dsd_filter_ext_bits.inc")


@dataclass(frozen=True)
class Fir:
    ratio: int
    ntaps: int
    tzero: int
    coefs: np.ndarray

    @property
    def nlut(self):          # dsd_decimator.cpp:121
        return (self.ntaps + 7) // 8

    @property
    def nstep(self):         # dsd_decimator.cpp:54
        return self.ratio // 8


def _load(path=_INC):
    if not os.path.exists(path):
        return {}
    text = open(path).read()
    out = {}
    for m in re.finditer(r"DSD_FILTER_BEGIN\((\w+), /\*ratio\*/ (\d+), /\*taps\*/ (\d+), /\*tzero\*/ (\d+)\)(.*?)DSD_FILTER_END",
                         text, re.S):
        ratio, ntaps, tzero = int(m.group(2)), int(m.group(3)), int(m.group(4))
        bits = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{16})ULL", m.group(5))], dtype=np.uint64)
        assert len(bits) == ntaps
        out[ratio] = Fir(ratio, ntaps, tzero, bits.view(np.float64))
    return out


FILTERS = _load()
# EXTENSION filters (/64, /128, /256; tools/design_ext_filters.py): no reference counterpart
EXTENSION_FILTERS = _load(_EXT_INC)


def pick(dsd_fs, out_fs, extensions=False):
    """Filter for a rate pair, or None where the reference says "Sorry, incompatible sample rate
    combination" (dsd_decimator.cpp:57-68): the choice is by the integer ratio alone.
    extensions=True also offers the /64, /128 and /256 extension filters."""
    if not out_fs:
        return None
    ratio = int(dsd_fs) // int(out_fs)
    fir = FILTERS.get(ratio)
    if fir is None and extensions:
        fir = EXTENSION_FILTERS.get(ratio)
    return fir


def build_lut(fir, msb_flag):
    """[nlut, 256] float64 tables: entry s of table t is the sum over bit = 0..k-1, in that order,
    of +-coefs[8t+bit]; the sign comes from byte bit 7-bit if msbIsPlayedFirst() else bit `bit`."""
    lut = np.zeros((fir.nlut, 256), dtype=np.float64)
    s = np.arange(256)
    for t in range(fir.nlut):
        k = min(8, fir.ntaps - 8 * t)
        acc = np.zeros(256, dtype=np.float64)
        for bit in range(k):
            on = (s >> (7 - bit)) & 1 if msb_flag else (s >> bit) & 1
            acc = acc + (-1 + 2 * on.astype(np.float64)) * fir.coefs[8 * t + bit]
        lut[t] = acc
    return lut


def app_params(bits, scale_db=4.0, dither=True):
    """scale, dither amplitude and clip as the reference's caller computes them (main.cpp:239-245,
    475-476): userScale = pow(10, dB/20), scale = userScale * 2^(bits-1), clip = 2^(bits-1) - 1."""
    user = 10.0 ** (float(np.float32(scale_db)) / 20)      # the option is a float (cmdline.h)
    scale = user * 2.0 ** (bits - 1)
    return scale, (1.0 if dither else 0.0), 2.0 ** (bits - 1) - 1


def app_frames(length_samples, fir):
    """Frames main.cpp's loops emit for one whole stream."""
    n = length_samples // fir.ratio - fir.nlut // fir.nstep + 1
    return max(n, 0)
