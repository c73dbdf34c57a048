"""Time of the dither generator alone (glibc_rand_fill_kernel through dsdgpu_dither_fill_device) for the
21.2 M output samples of one configs[1] step, CUDA events, median of 20; checks a window against the oracle."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dsf2flac_b200 import capi, filters  # noqa: E402
from tests.oracle_lib import Oracle  # noqa: E402

fir = filters.FILTERS[32]
g = capi.DsdGpu(filters.build_lut(fir, 1), 2, fir.nstep)
n = 21167966
out = torch.empty(n, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
ts = []
for k in range(25):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    g.dither_fill_device(12345, n, out, stream=st.cuda_stream)
    e1.record(st)
    st.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[5:])
r = Oracle().rand(2 * 4000, skip=2 * 12345).astype(np.float64) / 2147483647.0
want = r[0::2] - r[1::2]
ok = bool(np.array_equal(out[:4000].cpu().numpy(), want))
tail = Oracle().rand(2 * 50, skip=2 * (12345 + 1000000)).astype(np.float64) / 2147483647.0
ok = ok and bool(np.array_equal(out[1000000:1000050].cpu().numpy(), tail[0::2] - tail[1::2]))
print(json.dumps({"samples": n, "median_us": ts[len(ts) // 2] * 1e3, "gsamples_per_s": n / (ts[len(ts) // 2] * 1e-3) / 1e9,
                  "write_gbs": n * 8 / (ts[len(ts) // 2] * 1e-3) / 1e9, "matches_oracle": ok}))
