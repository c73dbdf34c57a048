"""Bounds of the segmented kernel's shared-memory accesses (dsf2flac_b200/csrc/dsd_kernels.cuh: seg_tile_at,
seg_stage_tile, decimate_seg_kernel; csrc/seg_core.h), restated in Python and swept over channel counts,
frame hops, segment lengths, start positions and source origins.  compute-sanitizer is not available on the
GPU pool, so what it would check is checked here by arithmetic: every window word a lane loads lies inside
its channel's row of the tile buffer, every table slot it can address lies inside the image (real taps,
or pad slots that hold -0.0), the 16 lanes of a half-warp sit on 16 distinct bank pairs, and every tile's
samples are covered exactly once."""
import itertools

import pytest

NTHR, GUARD, SLACK, LIMIT = 512, 128, 160, 232448


def row_slots(nstep):
    """SegLayout (96 slots per table row) up to 8 bytes per frame, SegLayoutNarrow (80) from 16 on (dsdgpu_create)."""
    return 80 if nstep >= 16 else 96


def layout(ks, nch, nstep):
    """Two groups of 256 threads per CTA, each with two tile buffers for tiles of 256 * ks output samples."""
    samples = NTHR // 2 * ks
    frames = (samples + nch - 1) // nch + 1
    stride = ((frames - 1) * nstep + SLACK + 15) // 16 * 16
    return samples, stride, GUARD + 256 * row_slots(nstep) * 8 + 4 * nch * stride


def floor16_rel(x):
    return x // 16 * 16            # the kernel rounds toward minus infinity for either sign


@pytest.mark.parametrize("nch,nstep", [(1, 8), (2, 8), (3, 8), (6, 8), (8, 8), (2, 16), (6, 16), (2, 32), (5, 32), (2, 4), (3, 1), (7, 2)])
def test_window_words_stay_inside_the_row(nch, nstep):
    ks = next(k for k in (4, 2, 1) if layout(k, nch, nstep)[2] <= LIMIT)
    samples, stride, smem = layout(ks, nch, nstep)
    ROW = row_slots(nstep)
    LUT_BYTES = GUARD + 256 * ROW * 8
    assert smem <= LIMIT and stride % 16 == 0
    for seg_n, seg_first, p0, in_first, total in itertools.product((1, 12, ROW - 32, ROW - 24, ROW - 15), (0, 72), (-1, 0, 143, 1000003), (0, 5, -7),
                                                                   (1, samples - 1, samples * 3 + 17)):
        nsteps = (seg_n + 15 + 3) // 4 * 4
        ngroups = nsteps // 4
        assert nsteps <= ROW                                   # slots s - j for s < nsteps, j < 16 end below ROW
        ntiles = (total + samples - 1) // samples
        for tile in {0, ntiles - 1, ntiles // 2}:
            i0 = tile * samples
            f_lo = i0 // nch
            pmin = p0 + f_lo * nstep - seg_first
            t0 = in_first + floor16_rel(pmin - nsteps - 16 - in_first)
            assert (t0 - in_first) % 16 == 0 and t0 <= pmin - nsteps - 16
            r0 = i0 - f_lo * nch
            left = total - i0
            for tid in (0, 1, 15, 16, 31, 255):               # thread index within its group
                j = tid & 15
                for k in range(ks):
                    n = tid + k * (NTHR // 2)
                    if n >= left:
                        n = left - 1
                    q = (r0 + n) // nch
                    c = r0 + n - q * nch
                    assert 0 <= c < nch
                    aoff = (pmin + j - t0) + q * nstep            # inside row c
                    phi = (aoff + 1) & 3
                    w_first = aoff - 3 - phi                      # group 0's low word; the word above it is loaded first
                    w_last = w_first - 4 * (ngroups - 1)
                    assert w_first % 4 == 0
                    assert w_last >= 0 and w_first + 8 <= stride, (seg_n, p0, in_first, tile, tid, k)
                    # table slots: step s reads slot s - j of row e: from -15 (guard / previous row's pads) to nsteps - 1
                    lo_addr = GUARD + 0 * ROW * 8 + 8 * (0 - j)
                    hi_addr = GUARD + 255 * ROW * 8 + 8 * (nsteps - 1 - j)
                    assert lo_addr >= 0 and hi_addr + 8 <= LUT_BYTES
    # bank pairs of a half-warp at any step: (s - j) mod 16 for j = 0..15
    for s in range(0, ROW):
        assert len({(s - j) % 16 for j in range(16)}) == 16


@pytest.mark.parametrize("ROW", [96, 80])
def test_pad_slots_cover_every_idle_step(ROW):
    """A lane runs nsteps = roundup4(seg_n + 15) steps; outside its seg_n real taps it must read -0.0: slots
    seg_n .. nsteps - 1 of its own row and slots ROW - 15 .. ROW - 1 of the row before it (or the guard)."""
    for seg_n in range(1, ROW - 15 + 1):
        nsteps = (seg_n + 15 + 3) // 4 * 4
        assert nsteps - 1 <= ROW - 1
        assert seg_n <= ROW - 15          # the previous row's last 15 slots are pads
        for j in range(16):
            real = [s for s in range(nsteps) if 0 <= s - j < seg_n]
            assert len(real) == seg_n and real == list(range(j, j + seg_n))     # taps in order, none skipped
