"""Device-resident cost of the three source layouts (planar rows, DSF block-planar, DSDIFF byte-interleaved) for
DSD256 5.1, 60 s -> 352.8 kHz (BASELINE configs[2]) and DSD128 stereo 60 s: CUDA events around dsdgpu_decimate_device."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dsf2flac_b200 import capi, filters  # noqa: E402

res = {}
for name, fs, out_fs, nch in (("dsd256_6ch", 4 * 2822400, 352800, 6), ("dsd128_2ch", 2 * 2822400, 176400, 2)):
    fir = filters.pick(fs, out_fs)
    nbytes = fs // 8 * 60
    n = filters.app_frames(nbytes * 8, fir)
    scale, amp, clip = filters.app_params(24, dither=False)
    nb = -(-nbytes // 4096)
    planar = torch.randint(0, 256, (nch, nb * 4096), dtype=torch.uint8, device="cuda")      # the last block is padding
    layouts = {"planar": (0, planar, nb * 4096),
               "dsf_4096": (4096, planar.reshape(nch, nb, 4096).permute(1, 0, 2).contiguous().reshape(-1), 0),
               "dsdiff_interleaved": (1, planar[:, :nbytes].t().contiguous().reshape(-1), 0)}
    outs = {}
    for lname, (block, src, stride) in layouts.items():
        g = capi.DsdGpu(filters.build_lut(fir, 1), nch, fir.nstep)
        g.set_option("source_block_bytes", block)
        d_out = torch.empty((n, nch), dtype=torch.int32, device="cuda")
        st = torch.cuda.Stream()
        ts = []
        for k in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.decimate_device(src, stride, 0, nbytes, 0x69, fir.nlut, n, scale, amp, clip, 0, capi.OUT_INT32, d_out, stream=st.cuda_stream)
            e1.record(st)
            st.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[3:])
        outs[lname] = d_out
        res[f"{name}/{lname}"] = {"ms": ts[len(ts) // 2], "tbit_per_s": nbytes * 8 * nch / (ts[len(ts) // 2] * 1e-3) / 1e12}
        g.close()
    res[f"{name}/same_output"] = bool(torch.equal(outs["planar"], outs["dsf_4096"]) and torch.equal(outs["planar"], outs["dsdiff_interleaved"]))
print(json.dumps(res))
