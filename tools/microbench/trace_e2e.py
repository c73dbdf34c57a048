import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from dsf2flac_b200 import capi, filters
import bench
fir = filters.pick(2*2822400, 176400)
lut = filters.build_lut(fir, 1)
data = bench.synth_stream(2*2822400, 2, 60)
nbytes = data.shape[1]
nframes = filters.app_frames(nbytes*8, fir)
scale, amp, clip = filters.app_params(24, dither=False)
g = capi.DsdGpu(lut, 2, fir.nstep)
pin_in = capi.PinnedBuffer((2, nbytes), np.uint8); pin_in.array[:] = data
pin_out = capi.PinnedBuffer((nframes*2*4,), np.uint8)
for _ in range(3):
    g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, nframes, scale, amp, clip, 0, capi.OUT_INT32, pin_out.ptr)
g.set_option("trace", 1)
g.decimate_host_raw(pin_in.ptr, nbytes, 0, nbytes, 0x69, fir.nlut, nframes, scale, amp, clip, 0, capi.OUT_INT32, pin_out.ptr)
