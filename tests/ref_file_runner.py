"""TEST INFRASTRUCTURE: runs ONE file conversion through the reference's own reader + decimator
(oracle/_ref, dsdref_file_run_app) and writes the result to an .npz.  One conversion per process:
the reference's readers keep "allocated" flags in file-scope statics (dsf_file_reader.cpp:40,
dsdiff_file_reader.cpp:46-49), so a second reader in the same process would crash.

usage: python -m tests.ref_file_runner <path> <kind:1=dsf|2=dff> <out_fs> <scale> <dither> <clip> <out.npz>
       out_fs = 0 runs the DoP path instead (dsdref_file_run_dop: the reference's DopPacker in
       dop_track_helper's loops); scale / dither / clip are then ignored.
"""
import ctypes as C
import sys

import numpy as np

from tests.oracle_lib import REF_SO


def main():
    path, kind, out_fs, scale, dither, clip, dst = sys.argv[1:8]
    lib = C.CDLL(REF_SO)
    lib.dsdref_file_run_app.restype = C.c_int64
    lib.dsdref_file_run_app.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    import os
    cap = os.path.getsize(path) // 1 + 4096          # frames <= bytes per channel
    info = (C.c_int32 * 4)()
    length = C.c_int64(-1)
    out = np.zeros((cap, 8), dtype=np.int32)
    flat = np.zeros(cap * 8, dtype=np.int32)
    if int(out_fs) == 0:
        lib.dsdref_file_run_dop.restype = C.c_int64
        lib.dsdref_file_run_dop.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        n = lib.dsdref_file_run_dop(path.encode(), int(kind), flat.ctypes.data, cap, info, C.byref(length))
    else:
        n = lib.dsdref_file_run_app(path.encode(), int(kind), int(out_fs), float(scale), float(dither), float(clip), 1,
                                    flat.ctypes.data, cap, info, C.byref(length))
    nch = max(1, info[0])
    pcm = flat[:max(n, 0) * nch].reshape(-1, nch)
    np.savez(dst, frames=n, channels=info[0], fs=info[1], msb=info[2], length=length.value, pcm=pcm)


if __name__ == "__main__":
    main()
