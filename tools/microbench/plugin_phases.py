"""Where the wall time of one drop-in conversion goes (DsdDecimator over a bulk source, 60 s DSD128 stereo):
context creation / destruction, pinned allocations, the refill round trips.  Prints one JSON line."""
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dsf2flac_b200 import capi, filters  # noqa: E402

DSD128 = 2 * 2822400


def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


fir = filters.pick(DSD128, 176400)
lut = filters.build_lut(fir, 1)
res = {}
g0 = capi.DsdGpu(lut, 2, fir.nstep)                       # CUDA context, module load
ctxs = []
res["create_ms"] = t(lambda: ctxs.append(capi.DsdGpu(lut, 2, fir.nstep)))
res["destroy_ms"] = t(lambda: ctxs.pop().close())
bufs = []
res["pinned_alloc_8MB_ms"] = t(lambda: bufs.append(capi.PinnedBuffer((1 << 23,), np.uint8)))
res["pinned_free_8MB_ms"] = t(lambda: bufs.pop().free())
nbytes = DSD128 // 8 * 60
pin_in = capi.PinnedBuffer((2, nbytes), np.uint8)
pin_in.array[:] = np.random.default_rng(1).integers(0, 256, size=(2, nbytes), dtype=np.uint8)
n = filters.app_frames(nbytes * 8, fir)
pin_out = capi.PinnedBuffer((n, 2), np.int32)
scale, amp, clip = filters.app_params(24, dither=False)
g0.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, pin_out.ptr)
res["whole_file_one_call_ms"] = t(lambda: g0.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0,
                                                               capi.OUT_INT32, pin_out.ptr))


def refills():
    f, step = 0, 1 << 20
    while f < n:
        k = min(step, n - f)
        g0.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut + f * fir.nstep, k, scale, amp, clip, 0, capi.OUT_INT32,
                             pin_out.ptr + f * 8)
        f += k


res["whole_file_2p20_frame_calls_ms"] = t(refills)
host = capi.host_lib()
cap = n + 2048
out = capi.PinnedBuffer((cap, 2), np.int32)
res["run_app_bulk_ms"] = t(lambda: host.dsdhost_run_app_bulk(pin_in.ptr, 2, nbytes, nbytes * 8, 0x69, 1, DSD128, 176400, scale, amp, clip,
                                                            64, 0, out.ptr, cap))
res["run_app_bulk_readahead_all_ms"] = t(lambda: host.dsdhost_run_app_bulk(pin_in.ptr, 2, nbytes, nbytes * 8, 0x69, 1, DSD128, 176400, scale,
                                                                          amp, clip, 64, 1 << 24, out.ptr, cap))
print(json.dumps(res))
# the file path: a 60 s DSD128 stereo .dsf on local disk -> dsdhost_convert_file (parse, read into pinned memory, bulk source)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.dsd_files import write_dsf  # noqa: E402
path = "/tmp/plugin_phases.dsf"
write_dsf(path, np.ascontiguousarray(pin_in.array), DSD128)
res2 = {"convert_file_ms": t(lambda: host.dsdhost_convert_file(path.encode(), 0, 176400, scale, amp, clip, 64, 0, out.ptr, cap), reps=4)}
with open(path, "rb") as fh:
    res2["plain_read_of_the_file_ms"] = t(lambda: (fh.seek(0), fh.readinto(memoryview(bytearray(os.path.getsize(path))))), reps=3)
print(json.dumps(res2))
os.remove(path)
