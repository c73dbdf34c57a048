"""TEST INFRASTRUCTURE: ctypes access to the two CPU checkers built by oracle/Makefile.

  Oracle()                oracle/_build/libdsdoracle.so  (C restatement, oracle/dsd_oracle.c)
  Oracle(reference=True)  oracle/_ref/libdsdref.so       (the unmodified reference decimator)

Both expose the same calls.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs use this module; the product never does.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libdsdoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdsdref.so")

_i64, _u32, _dbl, _vp = C.c_int64, C.c_uint32, C.c_double, C.c_void_p


class Oracle:
    def __init__(self, reference=False):
        self.reference = reference
        self.prefix = "dsdref_" if reference else "dsdoracle_"
        self.lib = C.CDLL(REF_SO if reference else ORACLE_SO)
        f = self._fn("info")
        f.restype = C.c_int
        f.argtypes = [_u32, _u32, _i64] + [C.POINTER(C.c_int32)] * 4 + [C.POINTER(_dbl)] * 2 + [C.POINTER(_i64)]
        f = self._fn("lut")
        f.restype = C.c_int
        f.argtypes = [_u32, _u32, C.c_int, _vp, C.c_int]
        f = self._fn("run_raw")
        f.restype = C.c_int
        if reference:   # (..., pre_steps, nframes, chunk_frames, scale, dither, clip, reseed, out_i32, out_f64)
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _u32, _u32, _i64, _i64, _i64, _dbl, _dbl, _dbl,
                          C.c_int, _vp, _vp]
        else:           # (..., pre_steps, nframes, rand_skip, scale, dither, clip, out_i32, out_f64)
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _u32, _u32, _i64, _i64, _i64, _dbl, _dbl, _dbl,
                          _vp, _vp]
        f = self._fn("run_app")
        f.restype = _i64
        if reference:
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _u32, _u32, _dbl, _dbl, _dbl, C.c_int, _vp, _i64]
        else:
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _u32, _u32, _dbl, _dbl, _dbl, _vp, _i64]

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def info(self, dsd_fs, out_fs, length_samples=0):
        r = [C.c_int32() for _ in range(4)]
        d = [_dbl(), _dbl()]
        n = _i64()
        ok = self._fn("info")(dsd_fs, out_fs, length_samples, *[C.byref(x) for x in r], *[C.byref(x) for x in d],
                              C.byref(n))
        if not ok:
            return None
        return dict(ratio=r[0].value, nlut=r[1].value, nstep=r[2].value, tzero=r[3].value, first_valid=d[0].value,
                    last_valid=d[1].value, length_pcm=n.value)

    def lut(self, dsd_fs, out_fs, msb_flag):
        buf = np.zeros((640, 256), dtype=np.float64)
        n = self._fn("lut")(dsd_fs, out_fs, int(msb_flag), buf.ctypes.data, 640)
        return buf[:n].copy() if n else None

    def set_custom_filter(self, fir):
        """EXTENSION filters: run the operator on `fir` (dsf2flac_b200.filters.Fir) for its ratio, which the
        reference itself rejects.  None clears the registration.  Process-global, like the reference's state."""
        f = self._fn("set_custom_filter")
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_int, C.c_int, _vp]
        if fir is None:
            return f(0, 0, 0, None)
        coefs = np.ascontiguousarray(fir.coefs, dtype=np.float64)
        return f(fir.ratio, fir.ntaps, fir.tzero, coefs.ctypes.data)

    def run_raw(self, planar, length_samples, msb_flag, dsd_fs, out_fs, pre_steps, nframes, scale, dither=0.0,
                clip=0.0, idle=0x69, want="i32", chunk_frames=0, rand_skip=0):
        """Raw operator: rewind, pre_steps x step(), nframes frames.  The dither stream is that of
        a fresh process (reference: srand(1) first), optionally `rand_skip` draws in (oracle only)."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        nch, nbytes = planar.shape
        out = np.zeros((nframes, nch), dtype=np.int32 if want == "i32" else np.float64)
        pi = out.ctypes.data if want == "i32" else None
        pf = out.ctypes.data if want != "i32" else None
        if self.reference:
            reseed = 1
            if rand_skip:                       # srand(1) + rand_skip draws, then run without reseeding
                self.rand(rand_skip)
                reseed = 0
            rc = self._fn("run_raw")(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), dsd_fs,
                                     out_fs, pre_steps, nframes, chunk_frames, scale, dither, clip, reseed, pi, pf)
        else:
            rc = self._fn("run_raw")(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), dsd_fs,
                                     out_fs, pre_steps, nframes, rand_skip, scale, dither, clip, pi, pf)
        if rc != 0:
            raise RuntimeError("oracle rejected the rate pair")
        return out

    def run_app(self, planar, length_samples, msb_flag, dsd_fs, out_fs, scale, dither=0.0, clip=0.0, idle=0x69):
        """Whole-track conversion with main.cpp's loop shape.  Returns int32 [frames, nch]."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        nch, nbytes = planar.shape
        cap = length_samples // 8 + 2048
        out = np.zeros((cap, nch), dtype=np.int32)
        if self.reference:
            n = self._fn("run_app")(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), dsd_fs,
                                    out_fs, scale, dither, clip, 1, out.ctypes.data, cap)
        else:
            n = self._fn("run_app")(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), dsd_fs,
                                    out_fs, scale, dither, clip, out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError(f"oracle run_app failed ({n})")
        return out[:n].copy()

    def dop_run_raw(self, planar, length_samples, msb_flag, pre_steps, nframes, idle=0x69, chunk_frames=0):
        """DopPacker::pack_buffer after rewind + pre_steps x reader step().  int32 [nframes, nch]."""
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        nch, nbytes = planar.shape
        out = np.zeros((nframes, nch), dtype=np.int32)
        f = self._fn("dop_run_raw")
        f.restype = C.c_int
        if self.reference:
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _i64, _i64, _i64, _vp]
            f(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), pre_steps, nframes, chunk_frames, out.ctypes.data)
        else:
            f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _i64, _i64, _vp]
            f(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), pre_steps, nframes, out.ctypes.data)
        return out

    def dop_run_app(self, planar, length_samples, msb_flag, idle=0x69):
        """dop_track_helper's loops for --onefile (oracle restatement only).  int32 [frames, nch]."""
        assert not self.reference
        planar = np.ascontiguousarray(planar, dtype=np.uint8)
        nch, nbytes = planar.shape
        cap = length_samples // 16 + 2048
        out = np.zeros((cap, nch), dtype=np.int32)
        f = self.lib.dsdoracle_dop_run_app
        f.restype = _i64
        f.argtypes = [_vp, C.c_int, _i64, _i64, C.c_uint8, C.c_int, _vp, _i64]
        n = f(planar.ctypes.data, nch, nbytes, length_samples, idle, int(msb_flag), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError(f"oracle dop_run_app failed ({n})")
        return out[:n].copy()

    def rand(self, n, skip=0):
        out = np.zeros(n, dtype=np.int32)
        if self.reference:
            out = np.zeros(n + skip, dtype=np.int32)
            self.lib.dsdref_libc_rand.argtypes = [_vp, _i64]
            self.lib.dsdref_libc_rand(out.ctypes.data, n + skip)
            out = out[skip:].copy()
        else:
            self.lib.dsdoracle_rand.argtypes = [_vp, _i64, _i64]
            self.lib.dsdoracle_rand(out.ctypes.data, n, skip)
        return out


class BestOracle:
    """What the GPU parity tests compare with: the UNMODIFIED reference (oracle/_ref) wherever it has been built -- it
    travels to the GPU box as a prebuilt library -- and the C restatement otherwise (tests/test_oracle.py pins the
    two to each other).  Calls only the restatement offers (dop_run_app) go to the restatement."""

    def __init__(self):
        self.port = Oracle()
        self.ref = Oracle(reference=True) if os.path.exists(REF_SO) else None
        self.reference = self.ref is not None

    def __getattr__(self, name):
        return getattr(self.ref if self.ref is not None else self.port, name)

    def info(self, dsd_fs, out_fs, length_samples=0):
        # A rate pair the reference rejects goes to the restatement: the reference's destructor frees an
        # uninitialised lookupTable after a rejected construction once any decimator of the process has set the
        # file-scope `lookupTableAllocated` (reference dsd_decimator.cpp:42,63-68,74-83) -- a latent double free.
        ratio = dsd_fs // out_fs if out_fs else 0
        rejected = ratio not in (8, 16, 32) and ratio != self._custom_ratio
        return (self.port if rejected or self.ref is None else self.ref).info(dsd_fs, out_fs, length_samples)

    _custom_ratio = None

    def set_custom_filter(self, fir):
        self._custom_ratio = fir.ratio if fir is not None else None
        r = self.port.set_custom_filter(fir)
        if self.ref is not None:
            r = self.ref.set_custom_filter(fir)
        return r

    def dop_run_app(self, *a, **k):
        return self.port.dop_run_app(*a, **k)


DSD64 = 2822400


def seeded_bytes(nch, nbytes, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(nch, nbytes), dtype=np.uint8)
