"""TEST INFRASTRUCTURE: runs ONE file conversion in a fresh process and writes the result to an .npz, through either
  ref     oracle/_ref/libdsdref.so          the reference's own reader + the reference's decimator
  dropin  oracle/_ref/libdropin_ref.so      the reference's own reader + the REPO's decimator on the GPU
  cpu     oracle/_ref/libdropin_ref_cpu.so  the same with the CPU stand-in of the C ABI (host logic only)
both driven by oracle/app_driver.h (main.cpp:169-191's loop) with any of the five getSamples<T>.
One conversion per process: the reference's readers keep "allocated" flags in file-scope statics.

usage: python -m tests.dropin_file_runner <ref|dropin|cpu> <path> <kind:1=dsf|2=dff> <out_fs> <scale> <dither> <clip>
                                          <type:0=i16|1=i32|2=i64|3=f32|4=f64> <read_ahead> <out.npz>
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = {"ref": "libdsdref.so", "dropin": "libdropin_ref.so", "cpu": "libdropin_ref_cpu.so"}
DTYPES = [np.int16, np.int32, np.int64, np.float32, np.float64]


def lib_path(which):
    return os.path.join(ROOT, "oracle", "_ref", LIBS[which])


def main():
    which, path, kind, out_fs, scale, dither, clip, typ, read_ahead, dst = sys.argv[1:11]
    lib = C.CDLL(lib_path(which))
    typ = int(typ)
    cap = os.path.getsize(path) + 4096               # frames <= bytes per channel
    flat = np.zeros(cap * 8, dtype=DTYPES[typ])
    info = (C.c_int32 * 4)()
    length = C.c_int64(-1)
    err = C.create_string_buffer(512)
    if which == "ref":
        f = lib.dsdref_file_run_typed
        f.restype = C.c_int64
        f.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p,
                      C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        n = f(path.encode(), int(kind), int(out_fs), float(scale), float(dither), float(clip), 1, typ, flat.ctypes.data,
              cap, info, C.byref(length))
    else:
        f = lib.dropin_file_run_typed
        f.restype = C.c_int64
        f.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_int, C.c_void_p,
                      C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_char_p, C.c_int]
        n = f(path.encode(), int(kind), int(out_fs), float(scale), float(dither), float(clip), int(read_ahead), typ,
              flat.ctypes.data, cap, info, C.byref(length), err, 512)
    nch = max(1, info[0])
    pcm = flat[:max(n, 0) * nch].reshape(-1, nch)
    np.savez(dst, frames=n, channels=info[0], fs=info[1], msb=info[2], length=length.value, pcm=pcm, error=err.value.decode())


if __name__ == "__main__":
    main()
