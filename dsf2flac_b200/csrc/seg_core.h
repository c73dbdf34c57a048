// seg_core.h -- geometry of the "segmented" decimation kernel (dsd_kernels.cuh, decimate_seg_kernel): the
// bank-conflict-free lane-skewed LUT gather for ANY filter geometry, one segment of at most SEG_MAX
// tables per launch, so that filters whose tables do not fit one SM's shared memory (the /64, /128
// and /256 extension filters: 144, 288, 576 tables) run as a chain of passes over a plane of running
// fp64 sums.  Shared with the host so that tests can check the layout arithmetic on the CPU box.
//
// What one pass computes (reference src/dsd_decimator.cpp:194-198, continued across passes): for
// every output sample, sum = (pass 0 ? 0.0 : sum after the previous pass); for t = seg_first ..
// seg_first + seg_n - 1: sum += LUT[t][byte(P - t)], fp64, in this order -- the reference's
// sequential sum over all its tables, cut at segment boundaries, bit for bit.
//
// Layout.  Tables of the segment entry-major, row e = ROW slots of 8 bytes: slots 0..seg_n-1 =
// LUT[seg_first + slot][e], every other slot -0.0 (the identity of fp64 addition).  ROW*8 is a
// multiple of 128, so the bank pair of a load is (slot mod 16) whatever the byte is.
// Schedule.  Lane j (= lane & 15) of a half-warp runs j taps behind lane 0: at step s it is at slot
// s - j.  Slots -15..-1 are the pad slots at the end of the previous row (the guard in front of row
// 0), slots seg_n.. are pads too, so no lane is ever predicated off: every lane runs
// steps 0 .. roundup4(seg_n + 15) - 1 and adds -0.0 outside its seg_n real taps.  The 16 lanes of a
// half-warp sit on 16 consecutive slots = 16 distinct bank pairs: one wavefront per LDS.64.
// Cost of the generality: (seg_n + 15) steps for seg_n taps (72 -> 83 % useful), against the ring
// kernel's 100 % for the reference's three filters.
#pragma once
#include <stdint.h>

namespace dsdgpu {

// ROW = 96 slots per table row: up to 81 tables per pass, 72 by default (144 = 2 x 72).  ROW = 80: up to 65,
// 64 by default -- a 33 KB smaller image, which buys twice the samples per tile where frames are 16 or 32
// bytes apart (decimation by 128 and 256: 5 and 9 passes instead of 4 and 8, each with twice the work per
// barrier and per thread).
template <int ROW_>
struct SegLayoutT {
    static constexpr int NTHR = 512;
    static constexpr int ROW = ROW_;                     // slots per table row
    static constexpr int ROW_BYTES = ROW * 8;
    static constexpr int GUARD_BYTES = 128;              // pads of "row -1"
    static constexpr int LUT_BYTES = GUARD_BYTES + 256 * ROW_BYTES;
    static constexpr int SEG_MAX = ROW - 15;             // tables per pass at most
    static constexpr int SEG_DEFAULT = ROW == 96 ? 72 : ROW - 16;
    static_assert(ROW % 16 == 0, "bank pattern: a row is a whole number of 128-byte lines");
    static constexpr int SLACK = 160;                    // bytes per tile row beyond (frames - 1) * nstep
    static constexpr int SMEM_LIMIT = 232448;            // 227 KB opt-in maximum per CTA (sm_100)

    static int steps(int seg_n) { return (seg_n + 15 + 3) / 4 * 4; }
    // frames a tile of `samples` consecutive output samples (frame-major, channel-minor) can touch
    static long long tile_frames(int samples, int nch) { return (samples + nch - 1) / nch + 1; }
    static long long row_stride(int samples, int nch, int nstep) {
        return ((tile_frames(samples, nch) - 1) * nstep + SLACK + 15) / 16 * 16;
    }
    // two groups of NTHR / 2 threads, each with two tile buffers for tiles of `samples` output samples
    static long long smem_bytes(int samples, int nch, int nstep) {
        return LUT_BYTES + 4 * nch * row_stride(samples, nch, nstep);
    }
};
typedef SegLayoutT<96> SegLayout;
typedef SegLayoutT<80> SegLayoutNarrow;

// Host side of the layout: the shared-memory image of tables [seg_first, seg_first + seg_n).
//   lut   [nlut][256] doubles as the reference builds them (dsd_decimator.cpp:117-151)
template <class L>
inline void seg_build_lut_image(const double* lut, int seg_first, int seg_n, unsigned char* image) {
    double* d = reinterpret_cast<double*>(image);
    for (int i = 0; i < L::LUT_BYTES / 8; ++i) d[i] = -0.0;
    double* rows = reinterpret_cast<double*>(image + L::GUARD_BYTES);
    for (int e = 0; e < 256; ++e)
        for (int t = 0; t < seg_n; ++t) rows[e * L::ROW + t] = lut[(size_t)(seg_first + t) * 256 + e];
}

}  // namespace dsdgpu
