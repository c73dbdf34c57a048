"""Lane-level emulation of the skewed kernel on the CPU (tests/cpp/emu_skew.cpp runs the very
skew_core.h the kernel compiles): every emitted fp64 sum equals the reference-order sum bit for
bit, and every table load of every half-warp is bank-conflict free."""
import json
import os
import subprocess

import pytest

EMU = os.path.join(os.path.dirname(__file__), "cpp", "_build", "emu_skew")


@pytest.mark.parametrize("nthr,nframes,p0,in_first,in_count,seed", [
    (512, 6144, 72, 0, 24700, 1),            # exactly one tile
    (512, 13000, 72, 0, 52100, 2),           # ragged last tile
    (512, 900, -1, 0, 3000, 3),              # starts in the idle pre-fill (p0 - t < 0)
    (512, 7000, 1000, 16, 20000, 4),         # source window starts late and ends early (fill both ends)
    (1024, 6144, 72, 0, 24700, 5),
    (1024, 9000, 75, 3, 30000, 6),           # unaligned in_first
    (512, 1, 72, 0, 100, 7),
])
def test_skew_lanes(nthr, nframes, p0, in_first, in_count, seed):
    r = subprocess.run([EMU, str(nthr), str(nframes), str(p0), str(in_first), str(in_count), str(seed)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout)
    assert rep["checked"] == nframes and rep["mismatches"] == 0
    assert rep["lut_wavefronts"] == rep["lut_halfwarp_loads"]
