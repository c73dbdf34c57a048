// skew_core.h -- per-lane logic of the bank-conflict-free ("skewed") decimation kernel.
//
// Shared between the CUDA kernel (dsd_kernels.cuh) and a host-side lane emulator
// (tests/cpp/emu_skew.cpp) so the index arithmetic can be proven on the CPU box: a lane never
// talks to another lane, so running this function once per thread id over a byte image of the
// shared memory reproduces the kernel exactly.
//
// What it computes (reference src/dsd_decimator.cpp:194-198): for each output frame a lane
// owns, sum = 0.0; for t = 0..NLUT-1: sum += LUT[t][byte(P - t)], fp64, in this order.
//
// Why it is shaped like this.  A 256-entry fp64 table spans every bank pair 16 times, so 16
// lanes gathering from the SAME table with data-dependent bytes collide ~3.4-way.  Here the
// tables are stored entry-major, LUT_s[byte][t] with a row of TPAD = 80 taps (72 real + 8 of
// -0.0), so the bank pair of an access is (t mod 16) whatever the byte is, and lane j of a
// half-warp runs j taps behind lane 0: at every LDS.64 the 16 lanes sit on 16 consecutive
// taps -> 16 distinct bank pairs -> one conflict-free 128-byte wavefront.  Each lane still
// walks its own output's taps 0,1,2,... in order, so the fp64 sum is the reference's sum bit
// for bit; the pad taps add -0.0, which is the identity for every fp64 value.
//
// Time runs in periods of TPAD steps.  At static step s lane j is at local step s-j: taps of its
// current frame for s >= j, and the last taps (TPAD+s-j) of its previous frame for s < j.  With
// PADS = 8 >= 7 the lanes fall in two classes: class A (j <= PADS) has only pad taps left when a
// period starts, class B (j > PADS) still owes real taps 65..71 in steps 0..6.  Hence two
// capture points per period (step 0 for A, step PADS+1 for B) after which all 32 lanes of the
// warp finish their previous frame together (scale/dither/clip/round/store, non-divergent).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DSD_HD __host__ __device__ __forceinline__
#else
#define DSD_HD inline
#endif

namespace dsdgpu {

template <int NLUT_, int NSTEP_, int NTHR_, int KR_>
struct SkewGeom {
    static constexpr int NLUT = NLUT_;    // tables (8 taps each)
    static constexpr int NSTEP = NSTEP_;  // bytes between successive frames
    static constexpr int NTHR = NTHR_;    // threads per CTA
    static constexpr int KR = KR_;        // frames per thread per tile
    static constexpr int TPAD = (NLUT + 15) / 16 * 16;
    static constexpr int PADS = TPAD - NLUT;
    static constexpr int ROW_BYTES = TPAD * 8;
    static constexpr int GUARD_BYTES = 128;            // -0.0 in front of row 0 (row "-1" pads)
    static constexpr int LUT_ROWS = 257;               // row 256 is only ever read as garbage
    static constexpr int LUT_BYTES = GUARD_BYTES + LUT_ROWS * ROW_BYTES;
    static constexpr int NT = NTHR * KR;               // frames per tile
    static constexpr int HOP = NTHR * NSTEP;           // bytes between a lane's successive frames
    static constexpr int FRONT = 104;                  // tile starts >= FRONT bytes before P_min
    static constexpr int TILE_BYTES = NT * NSTEP + 192;
    static constexpr int SPLIT = PADS + 1;             // class-B capture point
    static constexpr int EARLY_STEPS = 15 - PADS;      // steps in which class B owes real taps
    static constexpr int EARLY_GROUPS = (EARLY_STEPS + 3) / 4;
    static_assert(PADS >= 7, "two capture points need PADS >= 7");
    static_assert(PADS + 1 > 4 * EARLY_GROUPS - 1, "class B must not start inside an early group");
    static_assert((HOP + TPAD) % 4 == 0, "previous frame's window must be word aligned to this one");
    static_assert(LUT_BYTES % 16 == 0 && TILE_BYTES % 16 == 0, "16-byte staging chunks");
};

DSD_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> shift);
#endif
}

template <int K>
DSD_HD uint32_t byte_of(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x4440u | (uint32_t)K);
#else
    return (x >> (8 * K)) & 0xffu;
#endif
}

DSD_HD uint32_t lds_u32(const unsigned char* smem, uint32_t off) {
    return *reinterpret_cast<const uint32_t*>(smem + off);
}
DSD_HD double lds_f64(const unsigned char* smem, uint32_t off) {
    return *reinterpret_cast<const double*>(smem + off);
}
DSD_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

struct NoTrace {
    DSD_HD void lut(int, int, int, uint32_t) const {}
    DSD_HD void word(int, int, int, uint32_t) const {}
};

struct LaneRegs {
    double acc, stash;
    uint32_t lo, x;
};

// Steps [S0, S1) of one period for one lane.  w_early / w_cur: byte offsets (in smem) of the
// aligned word W0 of the window used in the early groups / in the rest of the period.
template <class G, int S0, int S1, class Trace>
DSD_HD void skew_steps(LaneRegs& r, const unsigned char* smem, uint32_t w_early, uint32_t w_cur,
                       uint32_t lut_early, uint32_t lut_cur, uint32_t shift, int tid, int period,
                       const Trace& tr) {
#pragma unroll
    for (int s = S0; s < S1; ++s) {
        const int g = s >> 2, u = s & 3;
        if (u == 0) {
            // bytes of steps 4g..4g+3 are the 4 bytes ending at lane offset A-4g: the aligned
            // words W0-g (low) and W0-g+1 (high), funnel-shifted by the lane's byte phase.
            uint32_t hi;
            if (g == 0) {
                hi = lds_u32(smem, w_early + 4);
                tr.word(tid, period, s, w_early + 4);
                r.lo = lds_u32(smem, w_early);
                tr.word(tid, period, s, w_early);
            } else if (g < G::EARLY_GROUPS) {
                hi = r.lo;
                r.lo = lds_u32(smem, w_early - 4u * (uint32_t)g);
                tr.word(tid, period, s, w_early - 4u * (uint32_t)g);
            } else if (g == G::EARLY_GROUPS) {
                hi = lds_u32(smem, w_cur - 4u * (uint32_t)(g - 1));
                tr.word(tid, period, s, w_cur - 4u * (uint32_t)(g - 1));
                r.lo = lds_u32(smem, w_cur - 4u * (uint32_t)g);
                tr.word(tid, period, s, w_cur - 4u * (uint32_t)g);
            } else {
                hi = r.lo;
                r.lo = lds_u32(smem, w_cur - 4u * (uint32_t)g);
                tr.word(tid, period, s, w_cur - 4u * (uint32_t)g);
            }
            r.x = funnel_r(r.lo, hi, shift);
        }
        uint32_t e;
        if (u == 0) e = byte_of<3>(r.x);
        else if (u == 1) e = byte_of<2>(r.x);
        else if (u == 2) e = byte_of<1>(r.x);
        else e = byte_of<0>(r.x);
        const uint32_t base = (s < G::EARLY_STEPS) ? lut_early : lut_cur;
        const uint32_t addr = e * (uint32_t)G::ROW_BYTES + base + 8u * (uint32_t)s;
        tr.lut(tid, period, s, addr);
        r.acc = add_rn(r.acc, lds_f64(smem, addr));
    }
}

// All KR frames of one thread in one tile.
//   smem      base of the CTA's shared memory; the LUT image sits at offset 0
//   tile_off  byte offset of this tile's byte buffer inside smem (multiple of 16)
//   a_tile    P_min - T0: offset inside the byte buffer of the newest byte of the tile's frame 0
//   emit(k, sum)  called once per owned frame, k = 0..KR-1 (frame = tile_first + k*NTHR + tid)
template <class G, class Emit, class Trace>
DSD_HD void skew_tile_lane(const unsigned char* smem, uint32_t tile_off, int tid, int a_tile,
                           Emit&& emit, const Trace& tr) {
    // skew index: ascending in the lower half-warp, descending in the upper one, so that the
    // 32 lanes' byte-window words (one LDS.32 per 4 taps) span < 32 consecutive words.
    const int j = (tid & 16) ? 15 - (tid & 15) : (tid & 15);
    const bool class_b = j > G::PADS;
    const int a0 = a_tile + tid * G::NSTEP + j;         // lane anchor: offset of "byte at s = 0"
    const int phi = (a0 + 1) & 3;
    const uint32_t shift = 8u * (uint32_t)phi;
    uint32_t w_cur = tile_off + (uint32_t)(a0 - 3 - phi);
    const uint32_t lut_cur = (uint32_t)(G::GUARD_BYTES - 8 * j);
    const uint32_t lut_early = lut_cur + (class_b ? (uint32_t)G::ROW_BYTES : 0u);
    const uint32_t back = (uint32_t)(G::HOP + G::TPAD);  // this window -> previous frame's tail

    LaneRegs r;
    r.acc = 0.0; r.stash = 0.0; r.lo = 0; r.x = 0;
    for (int n = 0; n <= G::KR; ++n) {
        uint32_t w_c = w_cur, w_e = w_cur;
        if (n > 0 && class_b) w_e = w_cur - back;
        if (n == G::KR) { w_c = w_cur - back; w_e = w_c; }   // drain period: stay inside the tile
        if (!class_b) { r.stash = r.acc; r.acc = 0.0; }
        skew_steps<G, 0, G::SPLIT>(r, smem, w_e, w_c, lut_early, lut_cur, shift, tid, n, tr);
        if (class_b) { r.stash = r.acc; r.acc = 0.0; }
        if (n > 0) emit(n - 1, r.stash);
        if (n == G::KR) break;
        skew_steps<G, G::SPLIT, G::TPAD>(r, smem, w_e, w_c, lut_early, lut_cur, shift, tid, n, tr);
        w_cur += (uint32_t)G::HOP;
    }
}

// Where a tile's byte buffer starts in the stream.  f0 = first frame of the tile.
//   t0     stream index of byte 0 of the buffer: a multiple of 16 away from in_first (so the
//          16-byte staging chunks are aligned in the source buffer) and >= FRONT bytes before
//          pmin = p0 + f0*NSTEP, the newest byte of the tile's first frame
//   a_tile pmin - t0
template <class G>
DSD_HD void skew_tile_origin(long long p0, long long in_first, long long f0, long long& t0, int& a_tile) {
    const long long pmin = p0 + f0 * G::NSTEP;
    long long rel = pmin - G::FRONT - in_first;
    rel = (rel >= 0) ? (rel & ~15LL) : -((-rel + 15) & ~15LL);
    t0 = in_first + rel;
    a_tile = (int)(pmin - t0);
}

// Host side of the layout: writes the shared-memory image of the tables.
//   lut   [NLUT][256] doubles as the reference builds them (dsd_decimator.cpp:117-151)
//   image LUT_BYTES bytes
template <class G>
inline void skew_build_lut_image(const double* lut, unsigned char* image) {
    const double negzero = -0.0;
    double* d = reinterpret_cast<double*>(image);
    const int total = G::LUT_BYTES / 8;
    for (int i = 0; i < total; ++i) d[i] = negzero;
    double* rows = reinterpret_cast<double*>(image + G::GUARD_BYTES);
    for (int e = 0; e < 256; ++e)
        for (int t = 0; t < G::NLUT; ++t) rows[e * G::TPAD + t] = lut[t * 256 + e];
}

}  // namespace dsdgpu
