"""Lane-level emulation of the ring kernel on the CPU (tests/cpp/emu_ring.cpp runs the very
ring_core.h the kernel compiles, walking each CTA's tiles as the kernel does, buffers overwritten
at the same points): every emitted fp64 sum equals the reference-order sum
(dsd_decimator.cpp:194-198) bit for bit, every (channel, frame) is emitted exactly once, no
shared-memory access leaves the CTA's allocation, every table load of every half-warp is
bank-conflict free, and the window loads are conflict free whenever the frame grid is 4-byte
aligned to the source buffer."""
import json
import os
import subprocess

import pytest

EMU = os.path.join(os.path.dirname(__file__), "cpp", "_build", "emu_ring")


@pytest.mark.parametrize("nwarp,kr,nframes,nch,p0,in_first,in_count,seed,grid", [
    (16, 3, 3072, 2, 79, 0, 13000, 1, 1),        # exactly one pair tile, aligned frame grid
    (16, 3, 20000, 2, 72, 0, 81000, 1, 2),       # the app's p0 = nLUT (unaligned by one byte); 2 CTAs x 3-4 tiles
    (16, 3, 7000, 1, 72, 0, 29000, 2, 1),        # mono: halves hold consecutive frame ranges, ragged end
    (16, 3, 15000, 3, -1, 0, 61000, 3, 3),       # odd channel count; starts in the idle pre-fill
    (16, 1, 9000, 2, 1000, 16, 17000, 4, 2),     # one round per tile; source starts late, ends early
    (8, 6, 9000, 6, 75, 3, 37000, 5, 2),         # 5.1, unaligned in_first, 8-warp geometry
    (16, 3, 1, 2, 72, 0, 100, 7, 4),             # a single frame
    (16, 2, 30000, 2, 75, 0, 121000, 8, 3),      # two rounds per tile, 5 tiles per CTA, aligned
    (76, 1, 20000, 2, 75, 0, 81000, 8, 3),       # nwarp "76" = the 32-bit geometry on 16 warps (80-step periods): two frames per 8-byte entry,
    (76, 1, 20000, 2, 72, 0, 81000, 1, 2),       #   76 pair tables, chains 8 bytes apart, eight frames per lane
    (76, 1, 15000, 3, -1, 0, 61000, 3, 3),
    (76, 1, 9001, 6, 75, 3, 37000, 5, 2),        # odd frame count: the last chain has one frame
    (76, 1, 7001, 1, 72, 0, 29000, 2, 1),
    (76, 1, 1, 2, 72, 0, 100, 7, 4),
    (76, 2, 40000, 2, 75, 0, 161000, 8, 3),
    (768, 1, 20000, 2, 75, 0, 81000, 8, 3),      # "768" = the product's 32-bit instance: 8 warps, 76-step periods (no pad steps,
    (768, 1, 20000, 2, 72, 0, 81000, 1, 2),      #   wrapped lanes read tail copies 20 slots further on)
    (768, 1, 15000, 3, -1, 0, 61000, 3, 3),
    (768, 1, 9001, 6, 75, 3, 37000, 5, 2),
    (768, 1, 7001, 1, 72, 0, 29000, 2, 1),
    (768, 1, 1, 2, 72, 0, 100, 7, 4),
    (30, 2, 30000, 2, 31, 0, 61000, 8, 3),       # nwarp "30" = the divide-by-16 filter (30 tables, 32-step periods)
    (30, 2, 9000, 3, -1, 0, 19000, 3, 2),
    (30, 2, 9000, 1, 30, 2, 18000, 3, 2),
    (12, 2, 30000, 2, 15, 0, 31000, 8, 3),       # nwarp "12" = the divide-by-8 filter (12 tables, 16-step periods)
    (12, 2, 9000, 6, -1, 0, 9500, 3, 2),
    (12, 2, 5000, 1, 12, 1, 5200, 4, 1),
    (12, 16, 70000, 2, 15, 0, 71000, 8, 2),      # the divide-by-8 filter's long tiles (16 rounds per tile), 2-3 tiles per CTA
    (12, 16, 40000, 3, -1, 0, 40500, 3, 1),
    (64, 1, 9000, 2, 75, 0, 73000, 4, 2),        # nwarp "64" = one pass of the divide-by-64 extension: 8-byte frame hop,
    (64, 1, 5000, 3, -1, 0, 41000, 3, 3),        #   six history words, sums start from a plane of initial values
    (64, 1, 3000, 1, 72, 5, 24000, 2, 1),
    (64, 1, 9000, 6, 79, 16, 70000, 5, 2),
    (64, 1, 9000, 2, 72, 0, 73000, 4, 2),        # frame grid off by one byte: window loads two-way
    (128, 1, 5000, 2, 75, 0, 81000, 4, 2),       # nwarp "128" = one pass of the divide-by-128 extension: 16-byte frame hop,
    (128, 1, 3000, 3, -1, 0, 49000, 3, 3),       #   two frames per lane (lanes stay 32 bytes apart), four history words
    (128, 1, 1500, 1, 72, 5, 24000, 2, 1),
    (128, 1, 5000, 6, 79, 16, 80000, 5, 2),
    (256, 1, 2500, 2, 75, 0, 81000, 4, 2),       # nwarp "256" = one pass of the divide-by-256 extension: 32-byte frame hop,
    (256, 1, 1500, 3, -1, 0, 49000, 3, 3),       #   one frame per lane, no history words
    (256, 1, 700, 1, 72, 5, 24000, 2, 1),
    (256, 1, 2500, 6, 79, 16, 80000, 5, 2),
])
def test_ring_lanes(nwarp, kr, nframes, nch, p0, in_first, in_count, seed, grid):
    r = subprocess.run([EMU, *map(str, (nwarp, kr, nframes, nch, p0, in_first, in_count, seed, grid))],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout)
    assert rep["checked"] == nframes * nch and rep["mismatches"] == 0
    assert rep["missing"] == 0 and rep["duplicates"] == 0 and rep["oob"] == 0
    assert rep["lut_wavefronts"] == rep["lut_halfwarp_loads"]
    if (p0 - in_first + 1) % 4 == 0:
        assert rep["word_wavefronts"] == rep["word_loads"]
    else:
        assert rep["word_wavefronts"] <= 2 * rep["word_loads"]
