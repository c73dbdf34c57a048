// ring_core.h -- per-lane logic of the "ring" decimation kernel (divide-by-32 filters, 72 tables).
//
// Shared between the CUDA kernel (dsd_kernels.cuh) and the host-side lane emulator
// (tests/cpp/emu_ring.cpp): a lane never talks to another lane inside this function, so running
// it once per thread id over a byte image of the CTA's shared memory reproduces the kernel, and
// the recorded addresses give the exact wavefront count of every shared-memory instruction.
//
// What it computes (reference src/dsd_decimator.cpp:194-198): for each output frame a lane owns,
// sum = 0.0; for t = 0..71: sum += LUT[t][byte(P - t)], fp64, in this order.
//
// Layout.  Tables are stored entry-major, row e = 80 doubles: slots 0..71 hold LUT[t][e], slots
// 72..79 repeat taps 64..71.  A row is 640 B, a multiple of 128 B, so the bank pair of a load
// is (slot mod 16) whatever the byte value is.
//
// Schedule.  Lane j of a half-warp (j = its skew, a permutation of 0..15) runs j taps behind
// lane 0, and every lane walks taps 0,1,...,71,0,1,... with no idle steps: at static step s of a
// period lane j is at tap s-j of its current frame (s >= j) or at tap 72+s-j of its previous
// frame (s < j, "wrapped").  Unwrapped lanes sit on bank pair (s-j) mod 16 -- all distinct.  A
// wrapped lane's natural slot 72+s-j lies 8 bank pairs off that pattern (72 mod 16 = 8), which
// would collide with lane j-8 or j+8; the two lanes of such a pair trade places: lanes j >= 8
// read their taps 64..71 from the duplicate slots 72..79 (slot = tap + 8).  With that every
// LDS.64 of every half-warp covers 16 distinct bank pairs: one 128-byte wavefront, and all 72
// steps of a period are real taps (the emulator asserts both).
//
// Each lane carries KF = 4 consecutive frames (4 independent fp64 chains).  Frame k+1 uses at
// step s the byte frame k used at step s-4, so one 32-bit window word feeds 16 table loads.
// A frame ends at a lane-dependent step (s = j-1), so a lane keeps two sets of sums: `n` for the
// frames started in this period and `p` for those started in the previous one.  In steps 0..15
// each table value x goes to exactly one of them through multipliers instead of selects:
// n = fma(x, 1 or 0, n), p = fma(x, 0 or 1, p) -- fma(x, 1, acc) is acc + x and fma(x, 0, acc) is
// acc, bit for bit (x = -0.0 aside, see dsdgpu_create).  After step 15 every lane's `p` is
// final and the warp emits together; at the end of the period p = n, n = 0.  The pipeline runs
// on across tile boundaries (the previous window then lies in the other tile buffer), so there is
// no drain except once at the very end of a CTA's work.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RING_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define RING_HD inline
#endif

namespace dsdgpu {

// Q32 = false: fp64 tables, the bit-exact path described above.
// Q32 = true : 32-bit tables.  Entries are the same sums rounded to multiples of 2^-qbits (qbits = 30
//              for the reference's filters) and stored as int32; a frame's sum is then an EXACT
//              integer sum -- order independent, so it does not depend on how frames are cut into
//              tiles or shards -- converted to fp64 once per frame.  |error| <= 72 * 2^-31 before
//              scaling (3e-8 of full scale; BASELINE's fp32 variant allows 1e-6).
//
// Frame pairs (Q32 = true, NLUT = 76, NSTEP = 8: the product's 32-bit geometry).  Frame A + 1 of the divide-by-32 filter
//              reads at its tap t + 4 the byte frame A reads at tap t, so ONE 8-byte entry
//              P[tau][e] = (LUT[tau - 4][e], LUT[tau][e])  (tau = 0..75, entries outside 0..71 are 0) serves both:
//              a "chain" = the frames (A, A + 1) at the position of A + 1, 8 bytes from the next chain, 76 pair
//              tables, two int32 sums per chain.  That is the geometry of the divide-by-64 pass below -- LDS.64,
//              80-step periods (4 pad slots of zeros), identity skew, 16-lane groups -- with a lane carrying
//              4 chains = 8 output frames: half the table loads, byte extractions and address computations per
//              frame of the one-entry-per-load layout, for 320 instead of 288 bytes of table reads.
//
// Other geometries whose period is NLUT rounded up to a multiple of 16 steps: the extra slots hold -0.0 (the identity
// of fp64 addition; 0 for integer tables), so a wrapped lane needs no second copy: slot PERIOD + s - j is bank pair
// (s - j) mod 16.  The 30-table filter (2 bytes per frame) runs periods of exactly 30 steps instead (PADLESS below).
//
// Short filters (NLUT = 30, NLUT = 12) keep COPIES = 2 / 4 images of their (small) tables, image a shifted by
// a * 16 / COPIES slots against the bank pattern.  Lane i of a half-warp reads image i % COPIES with skew
// j = i / COPIES, so that its bank pair is (tap + (16 / COPIES) * (i % COPIES)) mod 16 -- 16 distinct values
// with only 16 / COPIES distinct skews: the "early" steps, in which a lane may still be finishing its
// previous frames (two FMAs per table value, two byte windows), shrink from 16 to 8 (of 32) and to 4 (of 12;
// with four images the divide-by-8 period needs no pad slots at all).  The byte buffer of the upper
// half-warp starts 64 bytes further on: its 16 window words then take the banks the lower half leaves free.
//
// NSTEP = 8 (EXTENSION, decimation by 64: 144 tables, run as two passes of 72 over a plane of running
// sums): frames lie 8 bytes apart, so frame k of a lane reads the word stream of frame 0 shifted by
// 2k words -- six history words instead of three -- and a lane's sums start from the previous pass's
// value (Emit::init) instead of 0.  Lanes are 32 bytes apart; the identity skew puts the 16 window
// words of a half-warp in 16 distinct banks (8i + (i >> 2) mod 32), and the byte buffer of the upper
// half starts 16 bytes further on so that its 16 words take the other 16 banks.
//
// NSTEP = 16 / 32 (EXTENSION, decimation by 128 / 256: 288 / 576 tables, four / eight passes of 72): what must not change
// is the 32 bytes between neighbouring lanes, so a lane carries KF = 2 frames 16 bytes apart (frame 1 = the word stream
// shifted by four words) or KF = 1 frame.  Tile bytes, skew and bank pattern are those of the NSTEP = 8 pass; periods are
// short (72 * KF table loads per lane), so these instances keep four tile buffers and stage two tiles ahead.
template <int NWARP_, int KR_, bool Q32_ = false, int NLUT_ = 72, int NSTEP_ = 4>
struct RingGeom {
    static constexpr bool Q32 = Q32_;
    static constexpr int NLUT = NLUT_;                // tables
    static constexpr int NSTEP = NSTEP_;              // bytes between successive frames
    static constexpr bool PAIR = Q32;                 // 32-bit tables: two frames per 8-byte entry (see "Frame pairs" above)
    static constexpr int FPC = PAIR ? 2 : 1;          // output frames per chain (per sum a lane carries)
    // Exact integer sums do not care about the order of their terms: a lane that is j taps behind spends its first j
    // steps on the LAST j taps of the frames it is starting (bytes further back in its own window) instead of finishing
    // the frames of the previous period.  Same table slots, same banks, but one set of sums, nothing to route in the
    // early steps, no window shared between neighbouring tiles and no drain.
    static constexpr bool ROTATE = PAIR;
    static constexpr bool DUP = (NLUT == 72);         // 72-step periods with tail copies (see above)
    static constexpr int COPIES = NLUT == 12 ? 4 : NLUT == 30 ? 2 : 1;   // images of the tables (whatever fits shared memory)
    // 30 tables: a period of exactly 30 steps.  A wrapped lane reads its last taps (23..29) from copies 18 slots further on
    // (slot 48 + s - j: the bank pair it would have in a 32-step pattern), and the window of its previous frames goes on
    // two bytes off the word grid (30 mod 4): one more funnel shift per early group.
    static constexpr bool PADLESS = (NLUT == 30) || (Q32_ && NWARP_ == 8);   // (76 pair tables: the 96-slot rows fit beside the tiles of 8 warps)
    static constexpr int PERIOD = DUP ? 72 : (COPIES == 4 || PADLESS) ? NLUT : (NLUT + 15) / 16 * 16;     // steps per period
    static constexpr int GROUPS = (PERIOD + 3) / 4;   // 4-step groups = window words per period (the last one may be short)
    static constexpr int WRAP_SHIFT = 8 * (PERIOD % 4);   // bits by which the previous window's word stream is off the grid
    // consecutive chains (sums) per lane: four, or as many as keep neighbouring lanes 32 bytes apart (frame hops of 16 / 32 bytes)
    static constexpr int KF = NSTEP_ == 16 ? 2 : NSTEP_ == 32 ? 1 : 4;
    static constexpr int FPL = KF * FPC;              // consecutive output frames per lane
    static constexpr int FBYTES = NSTEP / FPC;        // stream bytes between successive output frames
    static constexpr int P_SHIFT = PAIR ? FBYTES : 0; // a chain sits at the position of the LATER of its two frames
    static constexpr int KR = KR_;                    // rounds (periods) per tile
    static constexpr int NWARP = NWARP_;
    static constexpr int NTHR = 32 * NWARP;
    static constexpr int ELEM = 8;                    // bytes per table entry: a double, or the two int32 of a frame pair
    static constexpr int SKEW_LANES = 16;             // lanes of one conflict-free load (LDS.64: a half-warp)
    static constexpr int NSKEW = SKEW_LANES / COPIES; // distinct skew values
    static constexpr int EARLY_GROUPS = (NSKEW + 3) / 4;  // 4-step groups in which some lane is still wrapped
    static constexpr int EARLY_STEPS = NSKEW;
    static constexpr int ROW = DUP ? 80 : PADLESS ? (PERIOD + NSKEW - 1 + 15) / 16 * 16 : (PERIOD + 15) / 16 * 16;   // slots per table row
    static constexpr int WRAP = PADLESS ? ROW : PERIOD;   // a wrapped lane reads slot WRAP + s - j
    static constexpr int ROW_BYTES = ROW * ELEM;      // a multiple of 128
    static constexpr int GUARD_BYTES = 128;           // keeps every per-lane offset positive
    static constexpr int COPY_STRIDE = 256 * ROW_BYTES + (COPIES > 1 ? 128 : 0);   // image a: + a * (COPY_STRIDE + ELEM * NSKEW)
    static constexpr int LUT_BYTES = GUARD_BYTES + COPIES * COPY_STRIDE;
    // a tile has two halves (lanes 0-15 / 16-31 of every warp), each with its own byte buffer
    static constexpr int LANE_BYTES = KF * NSTEP;     // bytes between the windows of neighbouring lanes
    static constexpr int BLK_PER_ROUND = 16 * NWARP;  // 4-frame blocks per half per round
    static constexpr int FR_PER_ROUND = KF * BLK_PER_ROUND;
    static constexpr int NTP = KR * FR_PER_ROUND;     // chains per half per tile
    static constexpr int NTPF = NTP * FPC;            // output frames per half per tile
    static constexpr int HOP = NSTEP * FR_PER_ROUND;  // bytes between a lane's rounds
    static constexpr int FRONT = 127;                 // buffer starts FRONT..FRONT+15 bytes before pmin
    static constexpr int QH = NSTEP >= 8 ? NSTEP * (KF - 1) / 4 : 3;   // window words of history: frame k reads NSTEP*k bytes further on
    static constexpr int QHA = QH > 0 ? QH : 1;       // (array size)
    static constexpr int HALF_SHIFT = LANE_BYTES == 32 ? 16 : COPIES > 1 ? 64 : 0;   // bank offset of the upper half-warp's buffer
    static constexpr int HALF_BYTES = NSTEP * NTP + 256 + HALF_SHIFT;
    static constexpr int TILE_BYTES = 2 * HALF_BYTES;
    // tile buffers: current, previous (still read), next -- and the one after where a period (KF * 72 table loads per lane)
    // is too short to cover a load from DRAM and a fourth buffer fits
    static constexpr int NBUF = (NSTEP >= 16) ? 4 : 3;
    static constexpr int SMEM_BYTES = LUT_BYTES + NBUF * TILE_BYTES + 2 * NBUF * 8 + 8;   // + full/empty mbarriers, + the table image's
    // where the emitter is called from: spread over the groups after the early ones if there are enough
    static constexpr bool EMIT_SPREAD = GROUPS >= EARLY_GROUPS + 7;   // fetch, four quantise calls, store, prefetch: one group each
    static constexpr int LATE_GROUPS = GROUPS - EARLY_GROUPS;
    static constexpr bool EMIT_SHORT = !EMIT_SPREAD && LATE_GROUPS >= 1;     // quantise calls shared out over the late groups
    static constexpr int EMIT_PER_GROUP = EMIT_SHORT ? (4 + LATE_GROUPS - 1) / LATE_GROUPS : 0;
    static_assert(!Q32 || (NLUT == 76 && NSTEP == 8), "32-bit tables exist for the 72-table filter: 76 pair tables with an 8-byte hop");
    static_assert(NSTEP == 1 || NSTEP == 2 || NSTEP == 4 || NSTEP == 8 || NSTEP == 16 || NSTEP == 32, "frame hop of 8 ... 256 DSD samples");
    static_assert(GROUPS >= EARLY_GROUPS, "every lane must restart within one period");
    static_assert(!PAIR || GROUPS >= EARLY_GROUPS + 15, "two emitter sequences fit the late groups");
    static_assert(!PADLESS || ROW >= NLUT + NSKEW - 1, "tail copies behind the tables");
    static_assert(LUT_BYTES % 16 == 0 && HALF_BYTES % 128 == HALF_SHIFT && ROW_BYTES % 128 == 0 && (WRAP_SHIFT == 0 || (NSTEP == 2 && WRAP_SHIFT == 16)), "staging chunks / bank pattern");
    static_assert(HOP % 16 == 0, "windows of successive rounds share their byte phase");
    static_assert(GUARD_BYTES >= ELEM * (SKEW_LANES - 1), "per-lane table offsets stay positive");
};

RING_HD uint32_t ring_funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> shift);
#endif
}

RING_HD uint32_t ring_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int n = 0; n < 4; ++n) r |= (uint32_t)((v >> (8 * ((sel >> (4 * n)) & 7))) & 0xff) << (8 * n);
    return r;
#endif
}

template <int K>
RING_HD uint32_t ring_byte(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x4440u | (uint32_t)K);
#else
    return (x >> (8 * K)) & 0xffu;
#endif
}

RING_HD uint32_t ring_lds_u32(const unsigned char* smem, uint32_t off) {
    return *reinterpret_cast<const uint32_t*>(smem + off);
}
RING_HD double ring_lds_f64(const unsigned char* smem, uint32_t off) {
    return *reinterpret_cast<const double*>(smem + off);
}
RING_HD double ring_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
RING_HD double ring_fma(double a, double m, double b) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, m, b);
#else
    return std::fma(a, m, b);
#endif
}

struct RingNoTrace {
    RING_HD void lut(int, int, int, int, uint32_t) const {}
    RING_HD void word(int, int, int, uint32_t) const {}
};

// What happens to the finished sums r.p[0..3] of the previous period.  They are final after step
// 15 and stay untouched until the period ends, so the kernel's emitter is called from INSIDE the
// main steps -- quantise(k, sum) before groups 5..8, store() before group 9 -- as straight-line
// code the compiler can schedule into the shadow of the table loads, instead of a block of
// scale/clip/round/store work that every warp of the CTA would hit at the same time.
// Multi-pass filters: prefetch() is called once per period after the stores (the emitter's registers are
// free again), init(k) at the hand-over -- the value the sum of the k-th frame of the NEXT period starts from.
// An emitter that needs data from global memory for quantise() (the dither values) sets LATE: fetch() is
// then called right after the early groups and the quantise / store / prefetch calls move five groups
// further on, so that the loads have a quarter of a period to land.
struct RingNoEmit {
    static constexpr bool LATE = false;
    RING_HD void fetch() {}
    RING_HD void quantise(int, double) {}
    RING_HD void store() {}
    RING_HD void prefetch() {}
    RING_HD double init(int) const { return 0.0; }
    // pair geometry: two sequences (frames 0..3 / 4..7 of the lane)
    RING_HD void fetch(int) {}
    RING_HD void quantise(int, int, double) {}
    RING_HD void store(int) {}
};

// Skew of lane l (0..31) of a warp.
// fp64 tables: within each half-warp the 16 values are a permutation of 0..15 (conflict-free
// LDS.64); across the warp the quarter (j >> 2) is arranged so that the 32 window words of one
// LDS.32 -- lane stride 16 B plus j bytes -- fall in 32 distinct banks when the frame grid is 4-byte
// aligned to the tile.
template <class G>
RING_HD int ring_lane_skew(int l) {
    const int h = (l >> 4) & 1, i = l & 15;
    if (G::COPIES > 1) return i / G::COPIES;            // image i % COPIES (ring_lane_image); see the top of this file
    if (G::LANE_BYTES == 32) return i;                  // lane stride of eight words (see the top of this file)
    if (G::LANE_BYTES == 8) {
        // lane stride of two words: window word of lane i = 2*i + (j >> 2); found by search
        // upper half: j >> 2 = {3,3,3,2,2,2,2,3,0,1,1,1,1,0,0,0}[i], j & 3 = {0,1,2,0,1,2,3,3,0,0,1,2,3,1,2,3}[i]
        const uint32_t q1 = 0x0154eabfu, a1 = 0xe790f924u;              // two bits per lane
        return h ? (int)(4 * ((q1 >> (2 * i)) & 3u) + ((a1 >> (2 * i)) & 3u)) : i;
    }
    // lane stride of four words (and, with two-way conflicts on the window loads, of one word)
    const int q = 2 * (((i >> 2) & 1) ^ h) + ((i >> 3) & 1);
    return (i & 3) + 4 * q;
}

// Which image of the tables lane l reads (COPIES > 1).
template <class G>
RING_HD int ring_lane_image(int l) {
    const int i = l & 15;
    if (G::COPIES == 1) return 0;
    return i % G::COPIES;
}

struct RingI2 { int32_t x, y; };                      // the two sums of a chain of the pair geometry: frames A, A + 1
template <bool Q32> struct RingAcc { typedef double type; };
template <> struct RingAcc<true> { typedef RingI2 type; };
RING_HD void ring_acc_clear(double& a) { a = 0.0; }
RING_HD void ring_acc_clear(RingI2& a) { a.x = 0; a.y = 0; }
RING_HD RingI2 ring_lds_i2(const unsigned char* smem, uint32_t off) {
#if defined(__CUDACC__)
    const int2 v = *reinterpret_cast<const int2*>(smem + off);
    RingI2 r; r.x = v.x; r.y = v.y;
    return r;
#else
    RingI2 r;
    r.x = *reinterpret_cast<const int32_t*>(smem + off);
    r.y = *reinterpret_cast<const int32_t*>(smem + off + 4);
    return r;
#endif
}

// Per-lane state.  It lives across periods AND across tiles: the pipeline never drains between
// tiles, a lane finishes the last frames of one tile during the first steps of the next.
template <class G>
struct RingLane {
    typedef typename RingAcc<G::Q32>::type acc_t;
    acc_t n[G::KF];             // sums of the frames started in the current period
    acc_t p[G::KF];             // sums of the frames started in the previous period (final after EARLY_STEPS)
    uint32_t qn[G::QHA], qo[G::QHA];   // previous window words of the current / the previous frame set
    uint32_t ln, lo;            // last aligned word loaded from either window
    uint32_t vold;              // WRAP_SHIFT: the previous window's last grid-aligned word
    uint32_t wlast;             // smem offset of word L(0) of the previous period's window
    // constants of the lane
    int j;                      // skew
    uint32_t shift;             // 8 * byte phase of the windows
    uint32_t lnew, lwrap, lalt; // table offset at static step 0: current frames / wrapped frames; step to the tail copy
    uint32_t sel[G::EARLY_GROUPS];   // byte selectors mixing the two windows in the early groups
    double unit;                // value of one unit of acc_t (1.0, or 2^-qbits for 32-bit tables)
    // finished sum of output frame k of the lane (k < FPL), exact: |p| < 2^31
    RING_HD double value(int k) const {
        if constexpr (G::PAIR) return (double)((k & 1) ? p[k >> 1].y : p[k >> 1].x) * unit;
        else return (double)p[k] * unit;
    }
};

// a0 = a_tile + LANE_BYTES*(16*warp + (lane & 15)) + j: offset inside the half's buffer of the byte the
// lane's frame 0 reads at its tap 0.  Returns the offset (inside the half's buffer) of L(0), the
// aligned word holding that byte's predecessor window, for round 0.
template <class G>
RING_HD int ring_lane_init(RingLane<G>& r, int tid, int a_tile, double unit = 1.0) {
    const int l = tid & 31, w = tid >> 5;
    const int j = ring_lane_skew<G>(l);
    const int a0 = a_tile + G::LANE_BYTES * (16 * w + (l & 15)) + j;
    const int phi = (a0 + 1) & 3;
    r.j = j;
    r.unit = unit;
    r.shift = 8u * (uint32_t)phi;
    const int image = ring_lane_image<G>(l);
    const int base = G::GUARD_BYTES + image * (G::COPY_STRIDE + G::ELEM * G::NSKEW);
    r.lnew = (uint32_t)(base - G::ELEM * j);
    r.lwrap = (uint32_t)(base + G::ELEM * (G::WRAP - j));
    r.vold = 0u;
    r.lalt = (uint32_t)(G::DUP ? G::ELEM * 8 * (j >> 3) : 0);
#pragma unroll
    for (int g = 0; g < G::EARLY_GROUPS; ++g) {
        uint32_t v = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) v |= (uint32_t)((4 * g + u >= j) ? 3 - u : 7 - u) << (4 * (3 - u));
        r.sel[g] = v;
    }
#pragma unroll
    for (int k = 0; k < G::KF; ++k) { ring_acc_clear(r.n[k]); ring_acc_clear(r.p[k]); }
#pragma unroll
    for (int i = 0; i < G::QHA; ++i) { r.qo[i] = 0u; r.qn[i] = 0u; }
    r.lo = r.ln = 0u;
    r.wlast = 1024u;            // first period: the "previous window" is table bytes (finite garbage, never emitted)
    return a0 - 3 - phi;
}

// The 4-byte windows of the lane's four frames for one group: frame k reads NSTEP*k bytes further
// on than frame 0, i.e. the word stream v = V(g), q[0..2] = V(g-1..g-3) shifted by NSTEP*k bytes.
template <class G>
RING_HD void ring_slot_words(uint32_t v, const uint32_t (&q)[G::QHA], uint32_t (&x)[G::KF]) {
    if constexpr (G::NSTEP >= 4) {
        x[0] = v;
#pragma unroll
        for (int k = 1; k < G::KF; ++k) x[k] = q[G::NSTEP * k / 4 - 1];
    } else if (G::NSTEP == 2) {
        x[0] = v; x[1] = ring_funnel_r(v, q[0], 16u); x[2] = q[0]; x[3] = ring_funnel_r(q[0], q[1], 16u);
    } else {
        x[0] = v; x[1] = ring_funnel_r(v, q[0], 8u); x[2] = ring_funnel_r(v, q[0], 16u); x[3] = ring_funnel_r(v, q[0], 24u);
    }
}

// Steps 4*G0 .. 4*G1-1 of one period.  wnew / wold: smem byte offsets of word L(0) of this
// period's window and of the previous period's window (whose groups 18.. continue here).
template <class G, int G0, int G1, class Emit, class Trace>
RING_HD void ring_groups(RingLane<G>& r, const unsigned char* smem, uint32_t wnew, uint32_t wold, uint32_t& loff,
                         int tid, int period, Emit& emit, const Trace& tr) {
#pragma unroll
    for (int g = G0; g < G1; ++g) {
        if constexpr (G::PAIR) {
            // eight finished frames per lane: two emitter sequences (frames 0..3, then 4..7), each fetch -> four quantise
            // calls -> store, staggered over the late groups (Emit is a RingEmitterPair-like object)
            constexpr int E = G::EARLY_GROUPS;
            constexpr int LEAD = Emit::LATE ? (G::GROUPS >= E + 16 ? 5 : 4) : 1;   // groups between a sequence's fetch (dither loads) and its first quantise
            constexpr int F0 = E + 1, Q0 = F0 + LEAD, F1 = Q0 + 4 - LEAD + 1, Q1 = Q0 + 5;
            static_assert(Q1 + 5 <= G::GROUPS, "both sequences end inside the period");
            if (g == F0) emit.fetch(0);
            if (g >= Q0 && g < Q0 + 4) emit.quantise(0, g - Q0, r.value(g - Q0));
            if (g == Q0 + 4) emit.store(0);
            if (g == F1) emit.fetch(1);
            if (g >= Q1 && g < Q1 + 4) emit.quantise(1, g - Q1, r.value(4 + g - Q1));
            if (g == Q1 + 4) emit.store(1);
        } else if (G::EMIT_SPREAD) {
            constexpr bool late = Emit::LATE && G::GROUPS >= G::EARLY_GROUPS + 12;
            constexpr int q0 = G::EARLY_GROUPS + (late ? 6 : 1);       // group of the first quantise call
            if (late && g == G::EARLY_GROUPS + 1) emit.fetch();
            if (!late && g == q0) emit.fetch();
            if (g >= q0 && g < q0 + G::KF) emit.quantise(g - q0, r.value(g - q0));
            if (g == q0 + 4) emit.store();
            if (g == q0 + 5) emit.prefetch();
        } else if (G::EMIT_SHORT) {
            // few groups after the early ones: the quantise calls shared out in front of them, the store after the last (ring_tail)
            if (g == G::EARLY_GROUPS) emit.fetch();
            if (g >= G::EARLY_GROUPS) {
#pragma unroll
                for (int k = (g - G::EARLY_GROUPS) * G::EMIT_PER_GROUP; k < (g - G::EARLY_GROUPS + 1) * G::EMIT_PER_GROUP && k < G::KF; ++k)
                    emit.quantise(k, r.value(k));
            }
        }
        // window words: v[0] = V(g), v[1..3] = V(g-1..g-3); frame k reads the 4 bytes NSTEP*k further on
        uint32_t x[G::KF];
        {
            const uint32_t a = wnew - 4u * (uint32_t)g;
            const uint32_t w = ring_lds_u32(smem, a);
            tr.word(tid, period, 2 * g, a);
            const uint32_t v = ring_funnel_r(w, r.ln, r.shift);
            r.ln = w;
            ring_slot_words<G>(v, r.qn, x);
#pragma unroll
            for (int i = G::QH - 1; i > 0; --i) r.qn[i] = r.qn[i - 1];
            if (G::QH > 0) r.qn[0] = v;
        }
        if (g < G::EARLY_GROUPS) {
            const uint32_t a = wold - 4u * (uint32_t)(G::GROUPS + g);
            const uint32_t w = ring_lds_u32(smem, a);
            tr.word(tid, period, 2 * g + 1, a);
            uint32_t v = ring_funnel_r(w, r.lo, r.shift);
            r.lo = w;
            if constexpr (G::WRAP_SHIFT != 0) {          // the previous frames' steps PERIOD + 4g ..: off the word grid
                const uint32_t aligned = v;
                v = ring_funnel_r(aligned, r.vold, (uint32_t)G::WRAP_SHIFT);
                r.vold = aligned;
            }
            uint32_t y[G::KF];
            ring_slot_words<G>(v, r.qo, y);
#pragma unroll
            for (int k = 0; k < G::KF; ++k) x[k] = ring_prmt(x[k], y[k], r.sel[g]);
#pragma unroll
            for (int i = G::QH - 1; i > 0; --i) r.qo[i] = r.qo[i - 1];
            if (G::QH > 0) r.qo[0] = v;
        }
        RingI2 pend2[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = 4 * g + u;
            if (s >= G::PERIOD) break;                   // (a period that is not a whole number of groups)
            bool cur = true;
            if (s < G::EARLY_STEPS) {
                // wrapped lanes (s < j) still add to their previous frames; multipliers route each
                // table value into exactly one of the two sums without a select: fma(x, 1, acc) is
                // acc + x and fma(x, 0, acc) is acc, bit for bit (integer tables: x*1 + acc, x*0 + acc)
                if (G::DUP && s + 8 < G::EARLY_STEPS && r.j == s + 8) loff += r.lalt;   // taps 64..71 from this lane's tail copy
                cur = (s >= r.j);
                if (r.j == s) loff = r.lnew;                     // tap 0 of the next frame set
            }
#pragma unroll
            for (int k = 0; k < G::KF; ++k) {
                uint32_t e;
                if (u == 0) e = ring_byte<3>(x[k]);
                else if (u == 1) e = ring_byte<2>(x[k]);
                else if (u == 2) e = ring_byte<1>(x[k]);
                else e = ring_byte<0>(x[k]);
                const uint32_t addr = e * (uint32_t)G::ROW_BYTES + loff + (uint32_t)(G::ELEM * s);
                tr.lut(tid, period, s, k, addr);
                if constexpr (G::PAIR) {
                    const RingI2 v = ring_lds_i2(smem, addr);
                    if ((u & 1) == 0) {
                        pend2[k] = v;            // two entries per three-input add (exact integer sums: any order)
                    } else {
                        r.n[k].x = r.n[k].x + pend2[k].x + v.x;
                        r.n[k].y = r.n[k].y + pend2[k].y + v.y;
                    }
                } else {
                    const double v = ring_lds_f64(smem, addr);
                    if (s < G::EARLY_STEPS) {
                        r.n[k] = ring_fma(v, cur ? 1.0 : 0.0, r.n[k]);
                        r.p[k] = ring_fma(v, cur ? 0.0 : 1.0, r.p[k]);
                    } else {
                        r.n[k] = ring_add(r.n[k], v);
                    }
                }
            }
        }
    }
}

// The first EARLY_STEPS steps of a period whose window word L(0) sits at smem offset wnew.  On
// return r.p[k] holds the finished sums of the frames the lane started in the previous period.
template <class G, class Trace>
RING_HD void ring_head(RingLane<G>& r, const unsigned char* smem, uint32_t wnew, uint32_t& loff, int tid, int period,
                       const Trace& tr) {
    // the QH words ahead of group 0 (frames 1..3 start NSTEP, 2*NSTEP, 3*NSTEP bytes further on)
    uint32_t l[G::QH + 1];                               // l[i] = word L(0) + 1 + i
#pragma unroll
    for (int i = G::QH; i >= 0; --i) {
        l[i] = ring_lds_u32(smem, wnew + 4u * (uint32_t)(i + 1));
        tr.word(tid, period, 100 + G::QH - i, wnew + 4u * (uint32_t)(i + 1));
    }
#pragma unroll
    for (int i = 0; i < G::QH; ++i) r.qn[i] = ring_funnel_r(l[i], l[i + 1], r.shift);
    r.ln = l[0];
    loff = r.lwrap;
    RingNoEmit none;
    if constexpr (G::ROTATE) {
        // the wrapped steps read on in this window, behind its last group: raw words w(g) = word L(0) - g,
        // V(g) = funnel(w(g), w(g - 1)); history V(GROUPS-1 .. GROUPS-QH), last raw word w(GROUPS-1)
        uint32_t w[G::QH + 1];
#pragma unroll
        for (int i = 0; i <= G::QH; ++i) {
            const uint32_t a = wnew - 4u * (uint32_t)(G::GROUPS - 1 - i);
            w[i] = ring_lds_u32(smem, a);
            tr.word(tid, period, 110 + i, a);
        }
#pragma unroll
        for (int i = 0; i < G::QH; ++i) r.qo[i] = ring_funnel_r(w[i], w[i + 1], r.shift);
        r.lo = w[0];
        ring_groups<G, 0, G::EARLY_GROUPS>(r, smem, wnew, wnew, loff, tid, period, none, tr);
        return;
    }
    ring_groups<G, 0, G::EARLY_GROUPS>(r, smem, wnew, r.wlast, loff, tid, period, none, tr);
}

// The remaining steps of the period (with the previous period's sums handed to `emit` on the
// way), then the hand-over to the next one.
template <class G, class Emit, class Trace>
RING_HD void ring_tail(RingLane<G>& r, const unsigned char* smem, uint32_t wnew, uint32_t& loff, int tid, int period,
                       Emit& emit, const Trace& tr) {
    if constexpr (!G::EMIT_SPREAD && !G::EMIT_SHORT) {   // shortest periods: the whole emitter before the remaining steps
        emit.fetch();
#pragma unroll
        for (int k = 0; k < G::KF; ++k) emit.quantise(k, r.value(k));
        emit.store();
        emit.prefetch();
    }
    ring_groups<G, G::EARLY_GROUPS, G::GROUPS>(r, smem, wnew, r.wlast, loff, tid, period, emit, tr);
    if constexpr (G::EMIT_SHORT) {
        emit.store();
        emit.prefetch();
    }
#pragma unroll
    for (int k = 0; k < G::KF; ++k) {
        r.p[k] = r.n[k];
        if constexpr (G::PAIR) ring_acc_clear(r.n[k]);
        else r.n[k] = (typename RingLane<G>::acc_t)emit.init(k);
    }
#pragma unroll
    for (int i = 0; i < G::QH; ++i) r.qo[i] = r.qn[i];   // this window's last words = next period's history
    if constexpr (G::WRAP_SHIFT != 0) {                  // ... in the shifted word stream the wrapped steps read
        r.vold = r.qn[0];
#pragma unroll
        for (int i = 0; i + 1 < G::QH; ++i) r.qo[i] = ring_funnel_r(r.qn[i], r.qn[i + 1], (uint32_t)G::WRAP_SHIFT);
    }
    r.lo = r.ln;
    r.wlast = wnew;
}

// Where the byte buffer of a tile half starts in the stream.  f0 = first frame of the half.
//   t0     stream index of byte 0 of the buffer: a multiple of 16 away from in_first (so the
//          16-byte staging chunks are aligned in the source buffer) and FRONT..FRONT+15 bytes
//          before pmin = p0 + f0*NSTEP, the newest byte of the half's first frame
//   a_tile pmin - t0 (the same for every tile of a launch: NTP*NSTEP is a multiple of 16)
template <class G>
RING_HD void ring_half_origin(long long p0, long long in_first, long long f0, long long& t0, int& a_tile) {
    const long long pmin = p0 + f0 * G::FBYTES + G::P_SHIFT;
    long long rel = pmin - G::FRONT - in_first;
    rel = (rel >= 0) ? (rel & ~15LL) : -((-rel + 15) & ~15LL);
    t0 = in_first + rel;
    a_tile = (int)(pmin - t0);
}

// Tiles of a launch.  Channels are taken two at a time (the halves of a warp then hold the same
// frames of channels c and c+1 and write whole frame pairs); an odd last channel is taken alone,
// its halves holding consecutive frame ranges.  Pair tiles are numbered channel-pair fastest.
struct RingTile { int c[2]; long long f0[2]; };

template <class G>
RING_HD long long ring_tile_count(long long nframes, int nch) {
    const long long nr = (nframes + G::NTPF - 1) / G::NTPF, nr2 = (nframes + 2 * G::NTPF - 1) / (2 * G::NTPF);
    return nr * (nch / 2) + nr2 * (nch & 1);
}

template <class G>
RING_HD RingTile ring_tile_at(long long tile, long long nframes, int nch) {
    const long long nr = (nframes + G::NTPF - 1) / G::NTPF;
    const int npair = nch / 2;
    RingTile t;
    if (tile < nr * npair) {
        // (mono and stereo -- one channel pair at most -- skip the 64-bit division: it runs once per tile and thread, which
        // the 12-step periods of the shortest filter feel)
        const long long rr = npair == 1 ? tile : tile / npair;
        const int u = (int)(tile - rr * npair);
        t.c[0] = 2 * u; t.c[1] = 2 * u + 1;
        t.f0[0] = t.f0[1] = rr * G::NTPF;
    } else {
        const long long rr = tile - nr * npair;
        t.c[0] = t.c[1] = nch - 1;
        t.f0[0] = 2 * rr * G::NTPF; t.f0[1] = t.f0[0] + G::NTPF;
    }
    return t;
}

// Host side of the layout: writes the shared-memory image of the tables.
//   lut   [72][256] doubles as the reference builds them (dsd_decimator.cpp:117-151); for 32-bit
//         tables every entry must already be a multiple of `unit` = 2^-qbits with |entry| / unit < 2^31
template <class G>
inline void ring_build_lut_image(const double* lut, unsigned char* image, double unit = 1.0) {
    for (int i = 0; i < G::LUT_BYTES; ++i) image[i] = 0;
    if constexpr (G::Q32) {
        // lut is the filter's own [72][256]; slot tau of row e = (LUT[tau - 4][e], LUT[tau][e]), zero outside the filter
        for (int e = 0; e < 256; ++e)
            for (int tau = 0; tau < G::NLUT; ++tau) {
                int32_t* at = reinterpret_cast<int32_t*>(image + G::GUARD_BYTES + e * G::ROW_BYTES + tau * G::ELEM);
                at[0] = tau >= G::FBYTES ? (int32_t)(lut[(tau - G::FBYTES) * 256 + e] / unit) : 0;
                at[1] = tau < G::NLUT - G::FBYTES ? (int32_t)(lut[tau * 256 + e] / unit) : 0;
                if (G::PADLESS && tau > G::NLUT - G::NSKEW) {        // tail copies: slot = tau + (ROW - NLUT)
                    int32_t* copy = at + 2 * (G::ROW - G::NLUT);
                    copy[0] = at[0]; copy[1] = at[1];
                }
            }
        return;                                           // (pad slots stay 0)
    }
    for (int i = 0; i < G::LUT_BYTES / 8; ++i) reinterpret_cast<double*>(image)[i] = -0.0;   // pad slots: the identity of fp64 addition
    auto put = [&](int image_no, int e, int slot, double v) {
        unsigned char* at = image + G::GUARD_BYTES + image_no * (G::COPY_STRIDE + G::ELEM * G::NSKEW) + e * G::ROW_BYTES + slot * G::ELEM;
        *reinterpret_cast<double*>(at) = v;
    };
    for (int c = 0; c < G::COPIES; ++c)
        for (int e = 0; e < 256; ++e) {
            for (int t = 0; t < G::NLUT; ++t) put(c, e, t, lut[t * 256 + e]);
            for (int copy = 1; G::DUP && G::NLUT + 8 * copy <= G::ROW; ++copy)      // tail copies: slot = tap + 8*copy
                for (int m = 0; m < 8; ++m) put(c, e, 64 + m + 8 * copy, lut[(64 + m) * 256 + e]);
            if (G::PADLESS)                                                          // tail copies: slot = tap + 18
                for (int t = G::NLUT - G::NSKEW + 1; t < G::NLUT; ++t) put(c, e, t + G::ROW - G::NLUT, lut[t * 256 + e]);
        }
}

}  // namespace dsdgpu
