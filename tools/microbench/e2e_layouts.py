"""End-to-end time (dsdgpu_decimate_host, pinned buffers) of one configs[1]-sized step by source layout: planar\nrows, DSF block-planar, DSDIFF byte-interleaved.  Measured: 2.33 / 2.42 / 2.36 ms -- the link decides, not the layout."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dsf2flac_b200 import capi, filters
fs = 2*2822400; nch=2
fir = filters.pick(fs, 176400)
nbytes = fs//8*60
nb = -(-nbytes//4096)
data = np.random.default_rng(1).integers(0,256,size=(nch,nb*4096),dtype=np.uint8)
n = filters.app_frames(nbytes*8, fir)
scale, amp, clip = filters.app_params(24, dither=False)
res={}
outs={}
for name, block in (("planar",0),("dsf",4096),("dsdiff",1)):
    if block==0: src=data; stride=nb*4096
    elif block==4096: src=np.ascontiguousarray(data.reshape(nch,nb,4096).transpose(1,0,2)).reshape(-1); stride=0
    else: src=np.ascontiguousarray(data[:,:nbytes].T).reshape(-1); stride=0
    pin=capi.PinnedBuffer((src.size,),np.uint8); pin.array[:]=src.reshape(-1)
    out=capi.PinnedBuffer((n,nch),np.int32)
    g=capi.DsdGpu(filters.build_lut(fir,1),nch,fir.nstep); g.set_option("source_block_bytes",block)
    ts=[]
    for k in range(8):
        t0=time.perf_counter()
        g.decimate_host_raw(pin.ptr, stride, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, out.ptr)
        ts.append(time.perf_counter()-t0)
    res[name]=round(sorted(ts[2:])[len(ts[2:])//2]*1e3,3)
    outs[name]=out.array.copy()
res["same"]=bool(np.array_equal(outs["planar"],outs["dsf"]) and np.array_equal(outs["planar"],outs["dsdiff"]))
print(json.dumps(res))
