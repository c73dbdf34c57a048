import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dsf2flac_b200 import capi, filters
fs = 2*2822400
fir = filters.pick(fs, 176400)
nbytes = fs//8*60
pin_in = capi.PinnedBuffer((2, nbytes), np.uint8); pin_in.array[:] = np.random.default_rng(1).integers(0,256,size=(2,nbytes),dtype=np.uint8)
n = filters.app_frames(nbytes*8, fir)
pin_out = capi.PinnedBuffer((n,2), np.int32)
scale, amp, clip = filters.app_params(24, dither=False)
g = capi.DsdGpu(filters.build_lut(fir,1), 2, fir.nstep)
for _ in range(3):
    g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, pin_out.ptr)
g.set_option("trace", 1)
g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, pin_out.ptr)
