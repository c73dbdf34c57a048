"""Small end-to-end cases for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsf2flac_b200 import capi, filters
from tests.oracle_lib import DSD64, Oracle, seeded_bytes

oracle = Oracle()
for nch, nbytes, dither in ((2, 70000, True), (3, 30000, False), (1, 20000, True)):
    fir = filters.pick(DSD64, 88200)
    data = seeded_bytes(nch, nbytes, nch)
    n = filters.app_frames(nbytes * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=dither)
    want = oracle.run_app(data, nbytes * 8, 1, DSD64, 88200, scale, amp, clip)
    with capi.DsdGpu(filters.build_lut(fir, 1), nch, fir.nstep) as g:
        g.set_option("chunk_frames", 9000)
        for kernel in (capi.KERNEL_RING, capi.KERNEL_DIRECT):
            g.set_option("kernel", kernel)
            got = g.decimate_host(data, fir.nlut, n, scale, amp, clip)
            assert np.array_equal(got, want), (nch, kernel)
        g.set_option("kernel", capi.KERNEL_RING)
        for block, src in ((4096, None), (1, np.ascontiguousarray(data.T).reshape(-1))):
            if src is None:
                nb = -(-nbytes // block)
                padded = np.zeros((nch, nb * block), np.uint8); padded[:, :nbytes] = data
                src = np.ascontiguousarray(padded.reshape(nch, nb, block).transpose(1, 0, 2)).reshape(-1)
            g.set_option("source_block_bytes", block)
            got = np.empty((n, nch), np.int32)
            g.decimate_host_raw(src.ctypes.data, 0, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, got.ctypes.data)
            assert np.array_equal(got, want), (nch, block)
print("sanitize_small ok")
