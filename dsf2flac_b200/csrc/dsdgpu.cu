// dsdgpu.cu -- sm_100a kernels + the C ABI of include/dsdgpu.h.
//
// Hot path replaced: DsdDecimator::getSamplesInternal (reference src/dsd_decimator.cpp:173-222)
// -- the per-byte lookup-table FIR (:194-198) fused with scale / TPDF dither / clip / round
// (:199-216).  Kernels:
//   decimate_ring_kernel    the production kernel for the divide-by-32 filters: bank-conflict-
//                           free lane-skewed LUT gather with no idle taps, 4 frames per lane,
//                           persistent CTAs, LUT image + double-buffered byte tiles in smem,
//                           frame-pair-coalesced int32 stores (ring_core.h)
//   decimate_skew_kernel    its predecessor (80-step periods with 8 pad taps, 1 frame per lane);
//                           kept selectable for A/B measurements (skew_core.h)
//   decimate_direct_kernel  one thread per output sample, tables [t][byte] in smem; any
//                           geometry, used for /16 and /8 filters and as a cross-check
//   glibc_rand_*            positional reconstruction of the reference's dither stream
//                           (unseeded glibc rand(), dsd_decimator.cpp:203-204)
// There is no CPU fallback anywhere in this file.
#include "dsdgpu.h"
#include "skew_core.h"
#include "ring_core.h"
#include "seg_core.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

using namespace dsdgpu;

thread_local char g_create_error[256] = "";

// ---------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------
struct DecimateArgs {
    const uint8_t* in;       // planar [nch][in_stride], or blocks of 2^block_shift bytes per channel
    long long in_stride;
    int block_shift;         // -1: planar rows; >= 4: DSF-style block-planar source (in_stride unused)
    long long in_first;      // stream index of in[c][0]
    long long in_count;      // valid bytes per channel
    long long p0;            // stream index of the newest byte of frame 0
    long long nframes;
    int nch;
    unsigned fill;
    const void* lut;         // kernel specific layout
    const double* dither;    // (r1 - r2) per output sample, [frame*dither_stride + c], or nullptr
    long long dither_stride; // samples per frame of the whole stream (>= nch)
    double scale, dither_amp, clip;
    double unit;             // 32-bit tables: value of one table unit (2^-qbits); 1.0 otherwise
    void* out;
    int out_kind;
    // multi-pass runs (filters longer than one shared-memory image): this launch adds
    // tables [seg_first, seg_first + seg_n) to the running sum of every sample
    int seg_first, seg_n;
    const double* seg_init;  // [frame*nch + c] sums after tables 0..seg_first-1, or nullptr (start from 0.0)
    double* seg_raw;         // where to leave the running sum instead of finishing the sample, or nullptr
    long long plane_stride;  // ring kernel passes keep the running sums planar: [c*plane_stride + frame]
};

// Byte offset from a.in of source byte o (0 <= o < in_count) of channel c.  Planar rows, or the
// DSF data-chunk layout [block 0: ch0 ch1 ..][block 1: ch0 ch1 ..] (dsf_file_reader.cpp:122-149),
// consumed in place instead of being re-packed by a reader.
__device__ __forceinline__ long long src_offset(const DecimateArgs& a, int c, long long o) {
    if (a.block_shift < 0) return (long long)c * a.in_stride + o;
    const long long blk = o >> a.block_shift;
    return (((blk * a.nch + c) << a.block_shift) | (o & ((1LL << a.block_shift) - 1)));
}
__device__ __forceinline__ bool src_aligned16(const DecimateArgs& a) {
    const unsigned long long bits = reinterpret_cast<unsigned long long>(a.in) |
                                    (a.block_shift < 0 ? (unsigned long long)a.in_stride : 0ULL);
    return (bits & 15ULL) == 0;
}

// dsd_decimator.cpp:199-216: scale, dither, clip.  No FMA contraction: the reference is x86-64
// -O3 without FMA, so sum*scale and (r1-r2)*amp are each rounded before the add.
__device__ __forceinline__ double quantise_tail(double sum, double scale, double amp, double dith,
                                                double clip) {
    double v = __dmul_rn(sum, scale);
    if (amp > 0.0) v = __dadd_rn(v, __dmul_rn(dith, amp));
    if (clip > 0.0) {
        if (v > clip) v = clip;
        else if (v < -clip) v = -clip;
    }
    return v;
}

__device__ __forceinline__ void store_sample(const DecimateArgs& a, long long frame, int c, double sum) {
    const long long m = frame * a.nch + c;
    const double dith = (a.dither != nullptr) ? a.dither[frame * a.dither_stride + c] : 0.0;
    const double v = quantise_tail(sum, a.scale, a.dither_amp, dith, a.clip);
    if (a.out_kind == DSDGPU_OUT_INT32) {
        // C round(): half away from zero (dsd_decimator.cpp:213-214)
        reinterpret_cast<int32_t*>(a.out)[m] = __double2int_rz(round(v));
    } else {
        reinterpret_cast<double*>(a.out)[m] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// direct kernel: reference loop order, one thread per (channel, frame)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
decimate_direct_kernel(DecimateArgs a, int nstep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lut_s = reinterpret_cast<double*>(smem_raw);
    const double* lut_g = reinterpret_cast<const double*>(a.lut) + (size_t)a.seg_first * 256;
    for (int i = threadIdx.x; i < a.seg_n * 256; i += blockDim.x) lut_s[i] = lut_g[i];
    __syncthreads();
    const long long total = a.nframes * a.nch;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long f = idx / a.nch;               // idx is the output index: stores coalesce
        const int c = (int)(idx - f * a.nch);
        const long long p = a.p0 + f * nstep - a.seg_first;
        double sum = a.seg_init ? a.seg_init[idx] : 0.0;
        for (int t = 0; t < a.seg_n; ++t) {
            const long long o = p - t - a.in_first;
            const unsigned b = (o >= 0 && o < a.in_count) ? (unsigned)a.in[src_offset(a, c, o)] : a.fill;
            sum = __dadd_rn(sum, lut_s[t * 256 + b]);
        }
        if (a.seg_raw) a.seg_raw[idx] = sum;
        else store_sample(a, f, c, sum);
    }
}

// ---------------------------------------------------------------------------------------------
// skewed kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// mbarrier (shared::cta) helpers for the ring kernel's split-phase tile pipeline
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.shared::cta.b64 t, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrives once every cp.async this thread has issued so far has landed (counts as one of the
// barrier's expected arrivals)
__device__ __forceinline__ void mbar_arrive_after_cp_async(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    for (unsigned spins = 0;; ++spins) {
        unsigned done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();       // a lost arrival must fail the launch, not hang the GPU
    }
}

// One 16-byte chunk of a tile's byte buffer: stream bytes o .. o+15 (relative to in_first) of channel c.
//   interior, 16-byte aligned source   cp.async (LDGSTS), coalesced
//   interior, any other alignment      five aligned words funnel-shifted into four (a caller's odd pointer
//                                      or row pitch costs bandwidth, not 16 byte loads per chunk)
//   touching the ends of the source    byte by byte, with the fill value outside
__device__ __forceinline__ void stage_chunk16(const DecimateArgs& a, int c, long long o, unsigned char* d, bool aligned16) {
    if (o >= 0 && o + 16 <= a.in_count) {
        const uint8_t* p = a.in + src_offset(a, c, o);
        if (aligned16) { cp_async16(d, p); return; }
        const unsigned mis = (unsigned)(reinterpret_cast<unsigned long long>(p) & 3ULL);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p - mis);
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
        const uint32_t w4 = mis ? q[4] : 0u;             // shares its aligned word with byte o+15: never a stray access
        const unsigned sh = 8u * mis;
        *reinterpret_cast<uint4*>(d) = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                                                  __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        return;
    }
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long ob = o + 4 * q + b;
            const unsigned byte = (ob >= 0 && ob < a.in_count) ? (unsigned)a.in[src_offset(a, c, ob)] : a.fill;
            v |= byte << (8 * b);
        }
        w[q] = v;
    }
    *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Tiles are numbered channel-fastest: the CTAs that run side by side write the same frames of
// neighbouring channels, so the 4-byte slots of one 32-byte output sector meet in L2 before the
// sector goes to HBM (frame-major interleaved output, dsd_decimator.cpp:213).
template <class G>
__device__ __forceinline__ void skew_tile_geometry(const DecimateArgs& a, long long tile,
                                                   long long tiles_per_ch, int& c, long long& f0,
                                                   long long& t0, int& a_tile) {
    (void)tiles_per_ch;
    c = (int)(tile % a.nch);
    f0 = (tile / a.nch) * G::NT;
    skew_tile_origin<G>(a.p0, a.in_first, f0, t0, a_tile);
}

// Stage the tile's bytes [t0, t0 + TILE_BYTES) of channel c.  Interior chunks go through
// cp.async (LDGSTS, 16 B, coalesced); chunks that touch the ends of the source are assembled
// byte by byte with the fill value.
template <class G>
__device__ __forceinline__ void skew_stage_tile(const DecimateArgs& a, int c, long long t0,
                                                unsigned char* dst, bool aligned16) {
    const long long o0 = t0 - a.in_first;
    for (int k = threadIdx.x; k < G::TILE_BYTES / 16; k += G::NTHR) {
        const long long o = o0 + 16LL * k;
        unsigned char* d = dst + 16 * k;
        stage_chunk16(a, c, o, d, aligned16);
    }
}

template <class G>
__global__ void __launch_bounds__(G::NTHR, 1) decimate_skew_kernel(DecimateArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [0, LUT_BYTES): table image; then two byte tiles.
    const int tid = threadIdx.x;
    {
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.lut);
        for (int k = tid; k < G::LUT_BYTES / 16; k += G::NTHR) cp_async16(smem + 16 * k, img + 16 * k);
    }
    const long long tiles_per_ch = (a.nframes + G::NT - 1) / G::NT;
    const long long ntiles = tiles_per_ch * a.nch;
    const bool aligned16 = src_aligned16(a);

    long long tile = blockIdx.x;
    int buf = 0;
    int c; long long f0, t0; int a_tile;
    if (tile < ntiles) {
        skew_tile_geometry<G>(a, tile, tiles_per_ch, c, f0, t0, a_tile);
        skew_stage_tile<G>(a, c, t0, smem + G::LUT_BYTES, aligned16);
    }
    cp_async_commit();
    while (tile < ntiles) {
        const long long next = tile + gridDim.x;
        if (next < ntiles) {
            int cn; long long f0n, t0n; int an;
            skew_tile_geometry<G>(a, next, tiles_per_ch, cn, f0n, t0n, an);
            skew_stage_tile<G>(a, cn, t0n, smem + G::LUT_BYTES + (buf ^ 1) * G::TILE_BYTES, aligned16);
        }
        cp_async_commit();
        cp_async_wait<1>();          // everything but the group just committed has landed
        __syncthreads();

        const long long fbase = f0 + tid;
        skew_tile_lane<G>(smem, (uint32_t)(G::LUT_BYTES + buf * G::TILE_BYTES), tid, a_tile,
                          [&](int k, double sum) {
                              const long long f = fbase + (long long)k * G::NTHR;
                              if (f < a.nframes) store_sample(a, f, c, sum);
                          },
                          NoTrace());
        __syncthreads();             // tile `buf` is free again
        tile = next;
        buf ^= 1;
        if (tile < ntiles) skew_tile_geometry<G>(a, tile, tiles_per_ch, c, f0, t0, a_tile);
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// segmented kernel (seg_core.h): any geometry, one segment of the filter's tables per launch
// ---------------------------------------------------------------------------------------------
// Tile = NT = 512 * KS consecutive output samples (frame-major, channel-minor: stores and the reads /
// writes of the running sums coalesce).  Byte buffer of a tile: one row per channel, row c holding
// stream bytes [t0, t0 + row_stride) of channel c.
struct SegTile {
    long long t0;            // stream index of byte 0 of every row
    long long i0;            // first output sample
};

// The CTA works as two groups of 256 threads, each with its own tiles and tile buffers and its own named
// barrier: while one group stages, sets up or stores, the other keeps the shared-memory pipe busy.
constexpr int kSegGroupThreads = SegLayout::NTHR / 2;

template <int KS>
__device__ __forceinline__ SegTile seg_tile_at(const DecimateArgs& a, long long tile, int nstep, int nsteps) {
    SegTile t;
    t.i0 = tile * (long long)(kSegGroupThreads * KS);
    const long long f_lo = t.i0 / a.nch;
    const long long pmin = a.p0 + f_lo * nstep - a.seg_first;        // newest byte of the tile's first frame, this segment's tap 0
    long long rel = pmin - nsteps - 16 - a.in_first;                  // lane 0 reads down to pmin - (nsteps - 1)
    rel = (rel >= 0) ? (rel & ~15LL) : -((-rel + 15) & ~15LL);
    t.t0 = a.in_first + rel;
    return t;
}

__device__ __forceinline__ void seg_stage_tile(const DecimateArgs& a, long long t0, int row_stride, unsigned char* dst,
                                               bool aligned16, int gt) {
    const int chunks_per_row = row_stride / 16;
    const long long o0 = t0 - a.in_first;
    for (int k = gt; k < a.nch * chunks_per_row; k += kSegGroupThreads) {
        const int c = k / chunks_per_row, kk = k - c * chunks_per_row;
        const long long o = o0 + 16LL * kk;
        unsigned char* d = dst + c * row_stride + 16 * kk;
        stage_chunk16(a, c, o, d, aligned16);
    }
}

__device__ __forceinline__ void seg_group_barrier(int group) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + group), "n"(kSegGroupThreads) : "memory");
}

template <int KS, class L>
__global__ void __launch_bounds__(SegLayout::NTHR, 1) decimate_seg_kernel(DecimateArgs a, int nstep, int row_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [0, LUT_BYTES): table image of this segment; then, per group, two byte tiles of nch * row_stride bytes.
    constexpr int NTHR = SegLayout::NTHR, GT = kSegGroupThreads, NT = GT * KS;
    const int tid = threadIdx.x, group = tid / GT, gt = tid - group * GT;
    {
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.lut);
        for (int k = tid; k < L::LUT_BYTES / 16; k += NTHR) cp_async16(smem + 16 * k, img + 16 * k);
    }
    const long long total = a.nframes * a.nch;
    const long long ntiles = (total + NT - 1) / NT;
    const bool aligned16 = src_aligned16(a);
    const int nsteps = (a.seg_n + 15 + 3) / 4 * 4, ngroups = nsteps / 4;
    const int tile_bytes = a.nch * row_stride;
    unsigned char* const bufs = smem + L::LUT_BYTES + group * 2 * tile_bytes;
    const int j = tid & 15;
    const uint32_t lut0 = (uint32_t)(L::GUARD_BYTES - 8 * j);

    const long long tstride = 2LL * gridDim.x;
    long long tile = 2LL * blockIdx.x + group;
    int buf = 0;
    SegTile tl = seg_tile_at<KS>(a, tile, nstep, nsteps);
    if (tile < ntiles) seg_stage_tile(a, tl.t0, row_stride, bufs, aligned16, gt);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();                 // the table image (loaded by all 512 threads) is in place for both groups
    while (tile < ntiles) {
        const long long next = tile + tstride;
        if (next < ntiles) {
            const SegTile tn = seg_tile_at<KS>(a, next, nstep, nsteps);
            seg_stage_tile(a, tn.t0, row_stride, bufs + (buf ^ 1) * tile_bytes, aligned16, gt);
        }
        cp_async_commit();
        cp_async_wait<1>();          // everything but the group just committed has landed
        seg_group_barrier(group);

        const uint32_t tile_off = (uint32_t)(L::LUT_BYTES + (group * 2 + buf) * tile_bytes);
        double acc[KS];
        uint32_t wlo[KS], lo[KS], shift[KS];
        int fl[KS], ch[KS];          // frame (relative to the tile's first) and channel of each sample; fl < 0: past the end
        const long long f_lo = tl.i0 / a.nch;
        const unsigned r0 = (unsigned)(tl.i0 - f_lo * a.nch);
        const long long left = total - tl.i0;                                 // samples from the tile's first to the end
        const int pbase = (int)(a.p0 + f_lo * nstep - a.seg_first - tl.t0) + j;    // byte of step 0 of frame f_lo
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            unsigned n = (unsigned)(gt + k * GT);
            const bool live = (long long)n < left;
            if (!live) n = (unsigned)(left - 1);                               // lanes past the end recompute the last sample
            const unsigned q = (r0 + n) / (unsigned)a.nch;
            const int c = (int)(r0 + n - q * (unsigned)a.nch);
            const int aoff = c * row_stride + pbase + (int)q * nstep;
            const int phi = (aoff + 1) & 3;
            shift[k] = 8u * (uint32_t)phi;
            wlo[k] = tile_off + (uint32_t)(aoff - 3 - phi);
            lo[k] = lds_u32(smem, wlo[k] + 4u);
            acc[k] = a.seg_init ? a.seg_init[tl.i0 + n] : 0.0;
            fl[k] = live ? (int)q : -1;
            ch[k] = c;
        }
        uint32_t lut = lut0;
#pragma unroll 2
        for (int g = 0; g < ngroups; ++g) {
            uint32_t x[KS];
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const uint32_t hi = lo[k];
                lo[k] = lds_u32(smem, wlo[k]);
                wlo[k] -= 4u;
                x[k] = funnel_r(lo[k], hi, shift[k]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    uint32_t e;
                    if (u == 0) e = byte_of<3>(x[k]);
                    else if (u == 1) e = byte_of<2>(x[k]);
                    else if (u == 2) e = byte_of<1>(x[k]);
                    else e = byte_of<0>(x[k]);
                    acc[k] = __dadd_rn(acc[k], lds_f64(smem, e * (uint32_t)L::ROW_BYTES + lut + 8u * (uint32_t)u));
                }
            }
            lut += 32u;
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            if (fl[k] >= 0) {
                if (a.seg_raw) a.seg_raw[tl.i0 + gt + k * GT] = acc[k];
                else store_sample(a, f_lo + fl[k], ch[k], acc[k]);
            }
        }
        seg_group_barrier(group);    // tile `buf` of this group is free again
        tile = next;
        buf ^= 1;
        if (tile < ntiles) tl = seg_tile_at<KS>(a, tile, nstep, nsteps);
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// ring kernel
// ---------------------------------------------------------------------------------------------
// Stage both halves of a tile: bytes [t0_h, t0_h + HALF_BYTES) of channel c_h.
template <class G>
__device__ __forceinline__ void ring_stage_tile(const DecimateArgs& a, const RingTile& tl, unsigned char* dst,
                                                bool aligned16) {
    constexpr int CHUNKS = G::HALF_BYTES / 16;
    for (int k = threadIdx.x; k < 2 * CHUNKS; k += G::NTHR) {
        const int h = k >= CHUNKS ? 1 : 0;
        const int kk = k - h * CHUNKS;
        long long t0; int a_tile;
        ring_half_origin<G>(a.p0, a.in_first, h ? tl.f0[1] : tl.f0[0], t0, a_tile);
        const int c = h ? tl.c[1] : tl.c[0];
        const long long o = t0 - a.in_first + 16LL * kk;
        unsigned char* d = dst + h * G::HALF_BYTES + 16 * kk;
        stage_chunk16(a, c, o, d, aligned16);
    }
}

// Where a lane's four frames of one period go: frames fb..fb+3 of channel c.  Its partner
// (lane ^ 16) holds the same frames of channel c+1 (pair tile, c0 = the even channel) or frames
// NTP further on of the same channel (single tile).
struct RingDest {
    long long fb;
    int c, c0;
    bool pair;
};

// The ring kernel's emitter (protocol: ring_core.h, RingNoEmit): dsd_decimator.cpp:199-216 as
// straight-line code -- scale, dither, clip by selects, C round() as trunc + exact remainder test,
// saturating double->int -- and frame-pair stores.  Called from inside the main steps.
// OUT: 0 = int32 PCM, 1 = fp64 (scaled, dithered, clipped), 2 = the raw running sum into the planar plane
// a.seg_raw (first pass of a two-pass filter).  INIT: sums start from the planar plane a.seg_init (second pass).
template <int OUT, bool DITHER, bool INIT>
struct RingEmitter {
    const DecimateArgs& a;
    RingDest d, nx;              // the frames being finished / the frames the lane starts in the next period
    int h;
    bool out_aligned16;
    int vi[4];
    double vd[4];
    double iv[4];
    double dv[4];
    static constexpr bool LATE = DITHER;

    __device__ __forceinline__ RingEmitter(const DecimateArgs& a_, const RingDest& d_, const RingDest& nx_, int h_, bool al)
        : a(a_), d(d_), nx(nx_), h(h_), out_aligned16(al) {}

    // the dither terms of the four frames being finished, issued well ahead of quantise()
    // Pair tiles read them the way the stores go out: a lane loads the 16-byte pairs {D(f, c0), D(f, c0 + 1)}
    // of the two frames it will store (f = fb + 2h, fb + 2h + 1) and trades the two values its partner
    // (lane ^ 16, the other channel) needs -- two coalesced 16-byte loads and two 64-bit shuffles per lane
    // instead of four 8-byte loads at a stride that touches a different cache line per lane.
    __device__ __forceinline__ void fetch() {
        if (!DITHER) return;
        const bool wide = d.pair && ((a.dither_stride | d.c0) & 1) == 0 &&
                          (reinterpret_cast<unsigned long long>(a.dither) & 15ULL) == 0;    // warp-uniform
        if (wide) {
            const long long f0 = d.fb + 2 * h, f1 = f0 + 1;
            const double2 u = __ldg(reinterpret_cast<const double2*>(a.dither + (f0 < a.nframes ? f0 : 0) * a.dither_stride + d.c0));
            const double2 v = __ldg(reinterpret_cast<const double2*>(a.dither + (f1 < a.nframes ? f1 : 0) * a.dither_stride + d.c0));
            const double r0 = __shfl_xor_sync(0xffffffffu, h ? u.x : u.y, 16);
            const double r1 = __shfl_xor_sync(0xffffffffu, h ? v.x : v.y, 16);
            dv[0] = h ? r0 : u.x; dv[1] = h ? r1 : v.x;
            dv[2] = h ? u.y : r0; dv[3] = h ? v.y : r1;
            return;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long f = d.fb + k < a.nframes ? d.fb + k : 0;     // keep the load in bounds
            dv[k] = __ldg(a.dither + f * a.dither_stride + d.c);
        }
    }

    __device__ __forceinline__ void quantise(int k, double sum) {
        if (OUT == 2) { vd[k] = sum; return; }
        double x = __dmul_rn(sum, a.scale);
        if (DITHER) x = __dadd_rn(x, __dmul_rn(dv[k], a.dither_amp));
        const bool clipping = a.clip > 0.0;
        x = (clipping && x > a.clip) ? a.clip : x;
        x = (clipping && x < -a.clip) ? -a.clip : x;
        if (OUT == 1) {
            vd[k] = x;
        } else {
            // round(): half away from zero.  x - trunc(x) is exact, so this is C's round() bit for bit.
            const double r = trunc(x);
            const double rem = __dsub_rn(x, r);
            const double adj = (rem >= 0.5) ? 1.0 : ((rem <= -0.5) ? -1.0 : 0.0);
            vi[k] = __double2int_rz(__dadd_rn(r, adj));
        }
    }

    // running sums the lane's next four frames start from (planar plane: 32 contiguous bytes per lane)
    __device__ __forceinline__ void prefetch() {
        if (!INIT) return;
        const double* src = a.seg_init + (long long)nx.c * a.plane_stride + nx.fb;
        if (nx.fb + 3 < a.nframes) {
            const double2 u = __ldg(reinterpret_cast<const double2*>(src));
            const double2 v = __ldg(reinterpret_cast<const double2*>(src) + 1);
            iv[0] = u.x; iv[1] = u.y; iv[2] = v.x; iv[3] = v.y;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) iv[k] = nx.fb + k < a.nframes ? __ldg(src + k) : 0.0;
        }
    }
    __device__ __forceinline__ double init(int k) const { return INIT ? iv[k] : 0.0; }

    __device__ __forceinline__ void store() {
        const long long fb = d.fb, nf = a.nframes;
        if (OUT == 2) {
            double* dst = a.seg_raw + (long long)d.c * a.plane_stride + fb;
            if (fb + 3 < nf) {
                reinterpret_cast<double2*>(dst)[0] = make_double2(vd[0], vd[1]);
                reinterpret_cast<double2*>(dst)[1] = make_double2(vd[2], vd[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (fb + k < nf) dst[k] = vd[k];
            }
            return;
        }
        if (OUT == 1) {
            double* out = reinterpret_cast<double*>(a.out);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (fb + k < nf) out[(fb + k) * a.nch + d.c] = vd[k];
            return;
        }
        int32_t* out = reinterpret_cast<int32_t*>(a.out);
        if (d.pair && (a.nch & 1) == 0) {
            // trade halves: lane h=0 ends up with frames fb, fb+1 of both channels, lane h=1 with fb+2, fb+3
            const int s0 = __shfl_xor_sync(0xffffffffu, h ? vi[0] : vi[2], 16);
            const int s1 = __shfl_xor_sync(0xffffffffu, h ? vi[1] : vi[3], 16);
            const long long f = fb + 2 * h;
            const int2 lo2 = h ? make_int2(s0, vi[2]) : make_int2(vi[0], s0);
            const int2 hi2 = h ? make_int2(s1, vi[3]) : make_int2(vi[1], s1);
            if (a.nch == 2 && out_aligned16 && f + 1 < nf) {
                *reinterpret_cast<int4*>(out + f * 2) = make_int4(lo2.x, lo2.y, hi2.x, hi2.y);
            } else {
                if (f < nf) *reinterpret_cast<int2*>(out + f * a.nch + d.c0) = lo2;
                if (f + 1 < nf) *reinterpret_cast<int2*>(out + (f + 1) * a.nch + d.c0) = hi2;
            }
        } else if (a.nch == 1 && out_aligned16 && fb + 3 < nf) {
            *reinterpret_cast<int4*>(out + fb) = make_int4(vi[0], vi[1], vi[2], vi[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (fb + k < nf) out[(fb + k) * a.nch + d.c] = vi[k];
        }
    }
};

// Tile pipeline.  Three tile buffers and two mbarriers per buffer; nothing in the steady state
// makes one warp wait for another:
//   full[b]   expects one arrival per thread; a thread arrives when the cp.async chunks it staged
//             for the tile have landed.  Staging of tile i+1 is issued at the top of tile i, the
//             wait comes a whole tile later.
//   empty[b]  a thread arrives when it has read buffer b for the last time (after the first 16
//             steps of the NEXT tile, which still finish frames of this one); the buffer is
//             refilled another tile later, so that wait does not block either.
template <class G, int OUT, bool DITHER, bool INIT>
__global__ void __launch_bounds__(G::NTHR, 1) decimate_ring_kernel(DecimateArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [0, LUT_BYTES): table image; three tiles of two byte buffers each; six mbarriers.
    constexpr int NBUF = G::NBUF;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::LUT_BYTES + NBUF * G::TILE_BYTES);
    uint64_t* empty = full + NBUF;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) { mbar_init(&full[b], G::NTHR); mbar_init(&empty[b], G::NTHR); }
    }
    __syncthreads();                 // barriers initialised before anybody arrives on them
    {
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.lut);
        for (int k = tid; k < G::LUT_BYTES / 16; k += G::NTHR) cp_async16(smem + 16 * k, img + 16 * k);
    }
    const long long ntiles = ring_tile_count<G>(a.nframes, a.nch);
    const bool aligned16 = src_aligned16(a);
    const bool out_aligned16 = (reinterpret_cast<unsigned long long>(a.out) & 15ULL) == 0;
    int a_tile;
    {
        long long t0;
        ring_half_origin<G>(a.p0, a.in_first, 0, t0, a_tile);   // the same for every half of the launch
    }
    const int l = tid & 31, w = tid >> 5, h = l >> 4;
    RingLane<G> lane;
    const uint32_t w0 = (uint32_t)(G::LUT_BYTES + h * G::HALF_BYTES + ring_lane_init<G>(lane, tid, a_tile, a.unit));
    uint32_t loff = 0;

    long long tile = blockIdx.x;     // grid <= ntiles
    RingTile tl = ring_tile_at<G>(tile, a.nframes, a.nch);
    ring_stage_tile<G>(a, tl, smem + G::LUT_BYTES, aligned16);
    __threadfence_block();           // plain stores of edge chunks are ordered before the arrival
    mbar_arrive_after_cp_async(&full[0]);                       // covers the table image too

    RingDest prev;                   // nothing to emit in the very first period: every store is off
    prev.fb = a.nframes; prev.c = 0; prev.c0 = 0; prev.pair = false;
    if (INIT) {                      // second pass: the sums of the first tile's frames start from the plane
        const int c = h ? tl.c[1] : tl.c[0];
        const long long fb = (h ? tl.f0[1] : tl.f0[0]) + 4LL * (16 * w + (l & 15));
#pragma unroll
        for (int k = 0; k < 4; ++k)
            lane.n[k] = fb + k < a.nframes ? (typename RingLane<G>::acc_t)a.seg_init[(long long)c * a.plane_stride + fb + k] : 0;
    }
    int b = 0, use = 0;              // buffer of the current tile, how often it has been used before
    for (long long i = 0; tile < ntiles; ++i) {
        const long long next = tile + gridDim.x;
        const int bn = b + 1 == NBUF ? 0 : b + 1;
        RingDest after;              // first frames of the next tile (INIT: where the next period's sums start)
        after.fb = a.nframes; after.c = 0; after.c0 = 0; after.pair = false;
        if (next < ntiles) {
            // buffer bn last held tile i-2; its use count is `use` when bn > b, else use + 1, minus one
            if (i + 1 >= NBUF) mbar_wait(&empty[bn], (unsigned)(((i + 1) / NBUF - 1) & 1));
            const RingTile tn = ring_tile_at<G>(next, a.nframes, a.nch);
            ring_stage_tile<G>(a, tn, smem + G::LUT_BYTES + bn * G::TILE_BYTES, aligned16);
            __threadfence_block();
            mbar_arrive_after_cp_async(&full[bn]);
            if (INIT) {
                after.c = h ? tn.c[1] : tn.c[0];
                after.fb = (h ? tn.f0[1] : tn.f0[0]) + 4LL * (16 * w + (l & 15));
            }
        }
        mbar_wait(&full[b], (unsigned)(use & 1));
        RingDest cur;
        cur.c = h ? tl.c[1] : tl.c[0];
        cur.c0 = tl.c[0];
        cur.pair = tl.c[1] != tl.c[0];
        cur.fb = (h ? tl.f0[1] : tl.f0[0]) + 4LL * (16 * w + (l & 15));
#pragma unroll 1
        for (int n = 0; n < G::KR; ++n) {
            const uint32_t wnew = w0 + (uint32_t)(b * G::TILE_BYTES + n * G::HOP);
            ring_head<G>(lane, smem, wnew, loff, tid, n, RingNoTrace());
            if (n == 0 && i > 0) mbar_arrive(&empty[b == 0 ? NBUF - 1 : b - 1]);   // done with the previous tile's bytes
            RingDest nx = cur;           // the frames this lane starts in the next period
            nx.fb += 4LL * G::BLK_PER_ROUND;
            if (n + 1 == G::KR) nx = after;
            RingEmitter<OUT, DITHER, INIT> emit(a, prev, nx, h, out_aligned16);
            ring_tail<G>(lane, smem, wnew, loff, tid, n, emit, RingNoTrace());
            prev = cur;
            cur.fb += 4LL * G::BLK_PER_ROUND;
        }
        tile = next;
        if (bn == 0) ++use;
        b = bn;
        if (tile < ntiles) tl = ring_tile_at<G>(tile, a.nframes, a.nch);
    }
    // the one drain of this CTA: 16 more steps finish the last frames
    ring_head<G>(lane, smem, lane.wlast, loff, tid, 0, RingNoTrace());
    RingEmitter<OUT, DITHER, INIT> emit(a, prev, prev, h, out_aligned16);
    emit.fetch();
#pragma unroll
    for (int k = 0; k < 4; ++k) emit.quantise(k, lane.value(k));
    emit.store();
}

// ---------------------------------------------------------------------------------------------
// DSDIFF sources: the "DSD " chunk is byte-interleaved, in[o*nch + c] (dsdiff_file_reader.cpp:234).
// One pass turns the bytes a run needs into planar rows (out[c*stride + pad + o]) for the kernels
// above; a thread owns one aligned 4-byte word of every row.  Pure HBM traffic: 2 bytes moved per
// source byte, against the 144 bytes of table reads each of them causes later.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
deinterleave_kernel(const uint8_t* __restrict__ in, long long count, int nch, uint8_t* __restrict__ out,
                    long long stride, int pad) {
    const long long words = (count + (pad & 3) + 3) / 4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < words;
         t += (long long)gridDim.x * blockDim.x) {
        const long long o0 = 4 * t - (pad & 3);                  // first source index of this word
        for (int c = 0; c < nch; ++c) {
            uint8_t* dst = out + (long long)c * stride + pad + o0;   // 4-byte aligned by construction
            if (o0 >= 0 && o0 + 4 <= count) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) v |= (uint32_t)in[(o0 + b) * nch + c] << (8 * b);
                *reinterpret_cast<uint32_t*>(dst) = v;
            } else {
                for (int b = 0; b < 4; ++b)
                    if (o0 + b >= 0 && o0 + b < count) dst[b] = in[(o0 + b) * nch + c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DoP packing (reference src/dop_packer.cpp:100-141): output sample (frame i, channel c) =
// marker + (B_c(pm - 1) << 8) + B_c(pm), pm = pm0 + 2*i + 1 the reader's marker after the frame's
// first step(); marker 0x00050000 when (8*pm - 8) % 32 == 0 (C remainder), else 0xFFFA0000; both
// bytes bit-reversed when the reader says msbIsPlayedFirst().  Pure byte traffic: 2 B in, 4 B out.
// One thread per output sample in output order (coalesced 4-byte stores).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dop_pack_kernel(DecimateArgs a, long long pm0, int reverse) {
    const long long total = a.nframes * a.nch;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / a.nch;
        const int c = (int)(idx - i * a.nch);
        const long long pm = pm0 + 2 * i + 1;
        const long long o1 = pm - 1 - a.in_first, o2 = pm - a.in_first;
        unsigned b1 = (o1 >= 0 && o1 < a.in_count) ? (unsigned)a.in[src_offset(a, c, o1)] : a.fill;
        unsigned b2 = (o2 >= 0 && o2 < a.in_count) ? (unsigned)a.in[src_offset(a, c, o2)] : a.fill;
        if (reverse) { b1 = __brev(b1) >> 24; b2 = __brev(b2) >> 24; }
        const int marker = ((8 * pm - 8) % 32 != 0) ? (int)0xFFFA0000 : 0x00050000;
        reinterpret_cast<int32_t*>(a.out)[idx] = marker + (int)(b1 << 8) + (int)b2;
    }
}

// Four consecutive output samples per thread (mono: 4 frames = 8 source bytes; stereo: 2 frames of
// both channels = 4 source bytes per channel), aligned word loads, one 16-byte store.  Requires
// (pm0 - in_first) % 4 == 0 and 4-byte aligned rows; groups that touch the ends of the source fall
// back to the byte-wise expression.
template <int NCH>
__global__ void __launch_bounds__(256)
dop_pack_wide_kernel(DecimateArgs a, long long pm0, int reverse) {
    const long long total = a.nframes * NCH, groups = (total + 3) / 4;
    constexpr int FR = 4 / NCH;                                   // frames per group
    constexpr int NB = 2 * FR;                                    // source bytes per channel per group
    constexpr int NW = NB / 4;
    constexpr int UN = 4;                                         // groups per thread per pass: all loads first
    const long long stride = (long long)gridDim.x * blockDim.x;
    int32_t* __restrict__ out = reinterpret_cast<int32_t*>(a.out);
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < groups; t0 += UN * stride) {
        uint32_t w[UN][NCH][NW];
        bool fast[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long t = t0 + u * stride;
            const long long o = pm0 + 2 * t * FR - a.in_first;   // source index of the group's first byte
            fast[u] = t < groups && o >= 0 && o + NB <= a.in_count && 4 * t + 3 < total;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(a.in + src_offset(a, c, fast[u] ? o : 0));
#pragma unroll
                for (int q = 0; q < NW; ++q) w[u][c][q] = fast[u] ? __ldg(src + q) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long t = t0 + u * stride;
            if (t >= groups) break;
            const long long f0 = t * FR;
            if (fast[u]) {
                int v[4];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int k = 0; k < FR; ++k) {
                        const uint32_t pair = (w[u][c][(2 * k) / 4] >> (8 * ((2 * k) & 3))) & 0xffffu;   // bytes 2k (low), 2k+1 (high)
                        uint32_t b1 = pair & 0xffu, b2 = pair >> 8;
                        if (reverse) { b1 = __brev(b1) >> 24; b2 = __brev(b2) >> 24; }
                        const long long pm = pm0 + 2 * (f0 + k) + 1;
                        const int marker = ((8 * pm - 8) % 32 != 0) ? (int)0xFFFA0000 : 0x00050000;
                        v[k * NCH + c] = marker + (int)(b1 << 8) + (int)b2;
                    }
                }
                reinterpret_cast<int4*>(out)[t] = make_int4(v[0], v[1], v[2], v[3]);
            } else {
                for (int e = 0; e < 4; ++e) {
                    const long long idx = 4 * t + e;
                    if (idx >= total) break;
                    const long long i = idx / NCH;
                    const int c = (int)(idx - i * NCH);
                    const long long pm = pm0 + 2 * i + 1;
                    const long long o1 = pm - 1 - a.in_first, o2 = pm - a.in_first;
                    unsigned b1 = (o1 >= 0 && o1 < a.in_count) ? (unsigned)a.in[src_offset(a, c, o1)] : a.fill;
                    unsigned b2 = (o2 >= 0 && o2 < a.in_count) ? (unsigned)a.in[src_offset(a, c, o2)] : a.fill;
                    if (reverse) { b1 = __brev(b1) >> 24; b2 = __brev(b2) >> 24; }
                    const int marker = ((8 * pm - 8) % 32 != 0) ? (int)0xFFFA0000 : 0x00050000;
                    out[idx] = marker + (int)(b1 << 8) + (int)b2;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// glibc rand() by position.  r[i] = r[i-31] + r[i-3] (mod 2^32) for i >= 34, draw k = r[344+k]>>1.
// With x^31 = x^28 + 1:  r[3+m+i] = sum_k c_k r[3+i+k],  c = x^m mod (x^31 - x^28 - 1).
//
// The stream is cut into pieces of 248 values (124 output samples) per lane, 32 lanes per warp
// (twice that per lane is slower: the kernel is bound by latency and occupancy, not by instruction count).
// The host jumps to the first value of the call: c = x^m as a product of precomputed powers
// (4 bits of m at a time), applied to the seed words -> 61 consecutive words ("segment").
// glibc_rand_fill_kernel: every warp advances that segment to its own start by up to three
// table polynomials (x^(7936*u), x^(7936*128*u), x^(7936*16384*u), u < 128; 31 MACs per lane
// each), every lane then applies its x^(248*lane) and runs the recurrence in registers.  Results
// go through shared memory so that the warp writes whole rows of 31 doubles.
// ---------------------------------------------------------------------------------------------
constexpr int kRandDeg = 31;
constexpr int kRandPowBits = 48;
constexpr int kRandLaneValues = 248;                      // 4 blocks of 62 values = 124 samples
constexpr int kRandLaneSamples = kRandLaneValues / 2;
constexpr int kRandWarpValues = 32 * kRandLaneValues;     // 7936
constexpr int kRandWarpSamples = kRandWarpValues / 2;     // 3968
constexpr int kRandLevel = 128;                           // entries per jump table level
constexpr int kRandBaseWords = 92;                        // seed words r[3..94]
constexpr int kRandFillWarps = 4;                         // warps per CTA of the fill kernel

__host__ __device__ inline void rand_poly_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    uint32_t t[2 * kRandDeg - 1];
    for (int i = 0; i < 2 * kRandDeg - 1; ++i) t[i] = 0;
    for (int i = 0; i < kRandDeg; ++i) {
        const uint32_t ai = a[i];
        if (ai == 0) continue;
        for (int k = 0; k < kRandDeg; ++k) t[i + k] += ai * b[k];
    }
    for (int d = 2 * kRandDeg - 2; d >= kRandDeg; --d) {   // x^d = x^(d-3) + x^(d-31)
        t[d - 3] += t[d];
        t[d - kRandDeg] += t[d];
    }
    for (int i = 0; i < kRandDeg; ++i) out[i] = t[i];
}

// words [31, 61) of a segment from its first 31: w[i] = w[i-31] + w[i-3]; three lanes, ten rounds
__device__ __forceinline__ void rand_extend_segment(uint32_t* seg, int lane) {
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const int i = kRandDeg + 3 * r + lane;
        if (lane < 3) seg[i] = seg[i - kRandDeg] + seg[i - 3];
        __syncwarp();
    }
}

// seg_out[i] = sum_k poly[k] * seg_in[i + k], i < 31 (one lane per word), then extended to 61 words
__device__ __forceinline__ void rand_apply_poly(const uint32_t* __restrict__ poly_g, const uint32_t* seg_in,
                                                uint32_t* seg_out, uint32_t* poly_s, int lane) {
    if (lane < kRandDeg) poly_s[lane] = poly_g[lane];
    __syncwarp();
    if (lane < kRandDeg) {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < kRandDeg; ++k) acc += poly_s[k] * seg_in[lane + k];
        seg_out[lane] = acc;
    }
    __syncwarp();
    rand_extend_segment(seg_out, lane);
}

// The 61-word segment r[n-31 .. n+29] at the call's first value n, computed on the host
// (rand_anchor_segment below: a dozen 31x31 polynomial products) and passed by value.
struct RandSegment { uint32_t w[64]; };

__global__ void __launch_bounds__(32 * kRandFillWarps)
glibc_rand_fill_kernel(const RandSegment segment, const uint32_t* __restrict__ lane_polys,
                       const uint32_t* __restrict__ jump_polys, long long count, double* __restrict__ out) {
    __shared__ uint32_t seg_s[kRandFillWarps][2][64];
    __shared__ uint32_t poly_s[kRandFillWarps][32];
    __shared__ double stage_s[kRandFillWarps][32][33];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const long long wi = (long long)blockIdx.x * kRandFillWarps + wl;
    const long long s_warp = wi * kRandWarpSamples;
    if (s_warp >= count) return;                                   // whole warp
    uint32_t* cur = seg_s[wl][0];
    uint32_t* oth = seg_s[wl][1];
    for (int i = lane; i < 2 * kRandDeg - 1; i += 32) cur[i] = segment.w[i];
    __syncwarp();
    // the call's first value -> this warp's first value
#pragma unroll 1
    for (int level = 2; level >= 0; --level) {
        const int u = (int)((wi >> (7 * level)) & (kRandLevel - 1));
        if (u == 0) continue;
        rand_apply_poly(jump_polys + ((size_t)level * kRandLevel + u) * kRandDeg, cur, oth, poly_s[wl], lane);
        uint32_t* tmp = cur; cur = oth; oth = tmp;
    }
    // this lane's state: x^(kRandLaneValues*lane) applied to the warp's segment
    uint32_t ring[kRandDeg];                                        // ring[i] = r[n - 31 + i]
    {
        uint32_t c[kRandDeg];
#pragma unroll
        for (int k = 0; k < kRandDeg; ++k) { c[k] = lane_polys[k * 32 + lane]; ring[k] = 0; }
#pragma unroll
        for (int d = 0; d < 2 * kRandDeg - 1; ++d) {
            const uint32_t v = cur[d];
#pragma unroll
            for (int i = 0; i < kRandDeg; ++i)
                if (d - i >= 0 && d - i < kRandDeg) ring[i] += c[d - i] * v;
        }
    }
    const double rmax = 2147483647.0, rcp = 1.0 / 2147483647.0;
    const bool whole = s_warp + kRandWarpSamples <= count;         // every sample of this warp is wanted
    double* const row0 = out + s_warp + lane;                       // row r of a block: + r * kRandLaneSamples + blk * 31
#pragma unroll 1
    for (int blk = 0; blk < kRandLaneValues / 62; ++blk) {
        double r1 = 0.0;
#pragma unroll
        for (int k = 0; k < 62; ++k) {
            const int slot = k % kRandDeg;
            const uint32_t v = ring[slot] + ring[(slot + 28) % kRandDeg];   // r[i-31] + r[i-3]
            ring[slot] = v;
            // draw / RAND_MAX, correctly rounded: q0 = a*rcp, q0 + (a - q0*d)*rcp (checked against
            // a/d for all 2^31 possible draws).  The draw becomes a double exactly as 2^52 + draw - 2^52
            // (one add on the fp64 pipe instead of a conversion instruction).
            const double a = __dsub_rn(__hiloint2double(0x43300000, (int)(v >> 1)), 4503599627370496.0);
            const double q0 = __dmul_rn(a, rcp);
            const double q = __fma_rn(__fma_rn(-q0, rmax, a), rcp, q0);
            if (k & 1) stage_s[wl][lane][k >> 1] = __dsub_rn(r1, q);
            else r1 = q;
        }
        __syncwarp();
        // row r of the stage = 31 consecutive samples of lane r
        double* const dst = row0 + blk * 31;
        if (whole) {
            if (lane < 31) {
#pragma unroll
                for (int r = 0; r < 32; ++r) dst[r * kRandLaneSamples] = stage_s[wl][r][lane];
            }
        } else {
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const long long s = s_warp + (long long)r * kRandLaneSamples + blk * 31 + lane;
                if (lane < 31 && s < count) out[s] = stage_s[wl][r][lane];
            }
        }
        __syncwarp();
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct dsdgpu_slot {
    cudaStream_t stream = nullptr;
    uint8_t* d_in = nullptr;
    size_t in_cap = 0;
    uint8_t* d_raw = nullptr;                // interleaved bytes as uploaded (DSDIFF sources)
    size_t raw_cap = 0;
    void* d_out = nullptr;
    size_t out_cap = 0;
    double* d_dither = nullptr;
    size_t dither_cap = 0;
    double* d_partial = nullptr;             // running fp64 sums between the passes of a long filter
    size_t partial_cap = 0;
    bool busy = false;
};

struct dsdgpu_ctx {
    int device = 0, nch = 0, nlut = 0, nstep = 0, lut_precision = 64;
    int sm_count = 0;
    int kernel = DSDGPU_KERNEL_AUTO;
    int threads = 512;
    int dither_stride = 0;                   // 0 = nch
    long long source_block = 0;              // 0: planar rows; 1: byte-interleaved; 2^k >= 16: block-planar
    long long chunk_frames = 1LL << 20;
    double* d_lut_direct = nullptr;          // [nlut][256]
    unsigned char* d_lut_skew = nullptr;     // SkewGeom::LUT_BYTES image (nlut == 72 only)
    unsigned char* d_lut_ring = nullptr;     // RingGeom::LUT_BYTES image (nlut == 72 only): fp64 or 32-bit entries
    unsigned char* d_lut_seg = nullptr;      // nseg x LUT_BYTES images, SEG_DEFAULT tables each (any geometry)
    int nseg = 0;
    bool seg_narrow = false;                 // SegLayoutNarrow (80-slot rows, 64 tables per pass): frame hops of 16 bytes and more
    double lut_unit = 1.0;                   // 32-bit tables: 2^-qbits
    std::vector<uint32_t> rand_pow16;        // host: x^(d * 16^k), d < 16, k < 12, [k][d][31]
    std::vector<uint32_t> rand_base;         // host: seed words r[3..94]
    uint32_t* d_rand_lane = nullptr;         // x^(kRandLaneValues*l), l < 32, stored [k][l]
    uint32_t* d_rand_jump = nullptr;         // x^(kRandWarpValues * 128^level * u), [level][u][k]
    cudaStream_t stream = nullptr;           // for dsdgpu_decimate_device(stream == NULL)
    dsdgpu_slot slots[DSDGPU_SLOTS];
    int pipeline_slots = DSDGPU_SLOTS;       // slots dsdgpu_decimate_host cycles through
    dsdgpu_slot own;                         // scratch for the device entry point
    unsigned long long launches = 0;
    char err[256] = "";
    // "trace" option: per-chunk timeline of dsdgpu_decimate_host (events on the slot streams), printed to stderr
    bool trace = false;
    struct TraceRow { cudaEvent_t e[4]; int slot; long long frames; };
    std::vector<TraceRow> trace_rows;
};

namespace {

using Skew512 = SkewGeom<72, 4, 512, 12>;
using Skew1024 = SkewGeom<72, 4, 1024, 6>;
using Ring16x2 = RingGeom<16, 2>;
using Ring16x2Q = RingGeom<16, 1, true>;   // 32-bit tables: two images (197 KB) leave room for one round per tile
using Ring30 = RingGeom<16, 2, false, 30, 2>;   // divide-by-16 filter: 30 tables, 2 bytes per frame
using Ring12 = RingGeom<16, 2, false, 12, 1>;   // divide-by-8 filter: 12 tables, 1 byte per frame
using Ring64 = RingGeom<16, 1, false, 72, 8>;   // one pass (72 of 144 tables) of the divide-by-64 extension filter

int fail(dsdgpu_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail((ctx), DSDGPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

void free_slot(dsdgpu_slot& s) {
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_raw) cudaFree(s.d_raw);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_dither) cudaFree(s.d_dither);
    if (s.d_partial) cudaFree(s.d_partial);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = dsdgpu_slot();
}

template <class T>
int ensure(dsdgpu_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (*cap >= need) return DSDGPU_OK;
    if (*p) CUDA_TRY(ctx, cudaFree(*p));
    *p = nullptr; *cap = 0;
    size_t want = need + need / 8 + 256;
    CUDA_TRY(ctx, cudaMalloc(reinterpret_cast<void**>(p), want));
    *cap = want;
    return DSDGPU_OK;
}

// glibc's seed-1 initial words r[0..64] (restated from the published TYPE_3 algorithm)
void glibc_initial_words(uint32_t* r, int n) {
    int32_t w[34];
    w[0] = 1;
    for (int i = 1; i < 31; ++i) {
        long long hi = w[i - 1] / 127773, lo = w[i - 1] % 127773;
        long long v = 16807 * lo - 2836 * hi;
        if (v < 0) v += 2147483647;
        w[i] = (int32_t)v;
    }
    for (int i = 31; i < 34; ++i) w[i] = w[i - 31];
    for (int i = 0; i < 34 && i < n; ++i) r[i] = (uint32_t)w[i];
    for (int i = 34; i < n; ++i) r[i] = r[i - 31] + r[i - 3];
}

// x^m mod (x^31 - x^28 - 1) applied to the seed words: the 61 words r[n-31 .. n+29], n = m + 34
void rand_anchor_segment(const dsdgpu_ctx* ctx, unsigned long long first_value, RandSegment& seg) {
    unsigned long long m = first_value - 34ULL;
    uint32_t c[kRandDeg] = {0}, t[kRandDeg];
    c[0] = 1;
    for (int k = 0; k < 12 && m != 0; ++k, m >>= 4) {
        const unsigned d = (unsigned)(m & 15ULL);
        if (!d) continue;
        rand_poly_mul(c, &ctx->rand_pow16[((size_t)k * 16 + d) * kRandDeg], t);
        memcpy(c, t, sizeof(c));
    }
    for (int i = 0; i < 2 * kRandDeg - 1; ++i) {
        uint32_t acc = 0;
        for (int k = 0; k < kRandDeg; ++k) acc += c[k] * ctx->rand_base[i + k];
        seg.w[i] = acc;
    }
}

int launch_dither(dsdgpu_ctx* ctx, dsdgpu_slot& s, unsigned long long first_sample, long long count,
                  double* d_out, cudaStream_t st) {
    (void)s;
    if (count <= 0) return DSDGPU_OK;
    const long long max_samples = (long long)kRandLevel * kRandLevel * kRandLevel * kRandWarpSamples;
    for (long long done = 0; done < count; done += max_samples) {      // one pass unless count > 8.3e9
        const long long n = count - done < max_samples ? count - done : max_samples;
        const unsigned long long first_value = 344ULL + 2ULL * (first_sample + (unsigned long long)done);
        if (first_value - 34ULL >= (1ULL << 48)) return fail(ctx, DSDGPU_ERR_ARG, "dither position beyond 2^47 samples");
        const long long warps = (n + kRandWarpSamples - 1) / kRandWarpSamples;
        RandSegment seg;
        rand_anchor_segment(ctx, first_value, seg);
        glibc_rand_fill_kernel<<<(unsigned)((warps + kRandFillWarps - 1) / kRandFillWarps), 32 * kRandFillWarps, 0, st>>>(
            seg, ctx->d_rand_lane, ctx->d_rand_jump, n, d_out + done);
        ctx->launches += 1;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

template <class G>
int launch_skew(dsdgpu_ctx* ctx, const DecimateArgs& a, cudaStream_t st) {
    const int smem = G::LUT_BYTES + 2 * G::TILE_BYTES;
    static thread_local int configured_device = -1;
    if (configured_device != ctx->device) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(decimate_skew_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_device = ctx->device;
    }
    const long long tiles = ((a.nframes + G::NT - 1) / G::NT) * a.nch;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    decimate_skew_kernel<G><<<grid, G::NTHR, smem, st>>>(a);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

template <class G, int OUT, bool DITHER, bool INIT>
int launch_ring_variant(dsdgpu_ctx* ctx, const DecimateArgs& a, cudaStream_t st) {
    const int smem = G::SMEM_BYTES;
    static thread_local int configured_device = -1;
    if (configured_device != ctx->device) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(decimate_ring_kernel<G, OUT, DITHER, INIT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_device = ctx->device;
    }
    const long long tiles = ring_tile_count<G>(a.nframes, a.nch);
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    decimate_ring_kernel<G, OUT, DITHER, INIT><<<grid, G::NTHR, smem, st>>>(a);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

template <class G, bool INIT = false>
int launch_ring(dsdgpu_ctx* ctx, const DecimateArgs& a, cudaStream_t st) {
    const bool f64 = a.out_kind != DSDGPU_OUT_INT32, dith = a.dither != nullptr;
    if (f64) return dith ? launch_ring_variant<G, 1, true, INIT>(ctx, a, st) : launch_ring_variant<G, 1, false, INIT>(ctx, a, st);
    return dith ? launch_ring_variant<G, 0, true, INIT>(ctx, a, st) : launch_ring_variant<G, 0, false, INIT>(ctx, a, st);
}

// Decimation by 64 (EXTENSION filter, 144 tables): two passes of the ring kernel with 8-byte frame hop.
// Pass 1 adds tables 0..71 and leaves the running sums in a planar fp64 plane; pass 2 starts from them,
// adds tables 72..143 -- the byte of table 72 + t at frame position p is the byte of table t at p - 72 --
// and finishes the samples.  The reference's sequential sum over all tables, cut once, bit for bit.
template <class G>
int launch_ring_two_pass(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, DecimateArgs a, cudaStream_t st) {
    a.plane_stride = (a.nframes + 3) & ~3LL;
    int rc = ensure(ctx, &scratch.d_partial, &scratch.partial_cap, (size_t)a.plane_stride * a.nch * sizeof(double));
    if (rc) return rc;
    DecimateArgs first = a;
    first.lut = ctx->d_lut_ring;
    first.seg_raw = scratch.d_partial;
    first.dither = nullptr;
    rc = launch_ring_variant<G, 2, false, false>(ctx, first, st);
    if (rc) return rc;
    a.lut = ctx->d_lut_ring + G::LUT_BYTES;
    a.p0 -= G::NLUT;
    a.seg_init = scratch.d_partial;
    return launch_ring<G, true>(ctx, a, st);
}

// The running sums of a multi-pass run live in scratch.d_partial, [frame*nch + c].
int ensure_partial(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, const DecimateArgs& a) {
    return ensure(ctx, &scratch.d_partial, &scratch.partial_cap, (size_t)a.nframes * a.nch * sizeof(double));
}

// Direct kernel, in passes of up to 96 tables (what fits its [t][byte] shared-memory layout).
int launch_direct(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, DecimateArgs a, cudaStream_t st) {
    constexpr int kSeg = 96;
    static thread_local int configured_device = -1;
    if (configured_device != ctx->device) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(decimate_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeg * 256 * 8));
        configured_device = ctx->device;
    }
    const long long total = a.nframes * a.nch;
    long long blocks = (total + 1023) / 1024;
    if (blocks > ctx->sm_count) blocks = ctx->sm_count;
    const int npass = (ctx->nlut + kSeg - 1) / kSeg;
    if (npass > 1) { int rc = ensure_partial(ctx, scratch, a); if (rc) return rc; }
    for (int s = 0; s < npass; ++s) {
        a.seg_first = s * kSeg;
        a.seg_n = ctx->nlut - a.seg_first < kSeg ? ctx->nlut - a.seg_first : kSeg;
        a.seg_init = s > 0 ? scratch.d_partial : nullptr;
        a.seg_raw = s + 1 < npass ? scratch.d_partial : nullptr;
        decimate_direct_kernel<<<(int)blocks, 1024, a.seg_n * 256 * (int)sizeof(double), st>>>(a, ctx->nstep);
        ctx->launches += 1;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

// Segmented kernel: one launch per SEG_DEFAULT tables, running sums in between.  KS (samples per
// thread) is the largest of 4, 2, 1 whose two byte tiles fit beside the table image.
template <int KS, class L>
int launch_seg_pass(dsdgpu_ctx* ctx, const DecimateArgs& a, int row_stride, int smem, cudaStream_t st) {
    static thread_local int configured_device = -1;
    if (configured_device != ctx->device) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(decimate_seg_kernel<KS, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_LIMIT));
        configured_device = ctx->device;
    }
    const long long total = a.nframes * a.nch, nt = L::NTHR / 2 * KS;           // tile of one 256-thread group
    const long long tiles = (total + nt - 1) / nt, pairs = (tiles + 1) / 2;
    const int grid = (int)(pairs < ctx->sm_count ? pairs : ctx->sm_count);
    decimate_seg_kernel<KS, L><<<grid, L::NTHR, smem, st>>>(a, ctx->nstep, row_stride);
    ctx->launches += 1;
    return DSDGPU_OK;
}

// samples per thread of the segmented kernel: the largest of 4, 2, 1 that fits shared memory, 0 = none does
template <class L>
int seg_samples_per_thread_of(const dsdgpu_ctx* ctx) {
    int ks = 4;
    while (ks >= 1 && L::smem_bytes(L::NTHR / 2 * ks, ctx->nch, ctx->nstep) > L::SMEM_LIMIT) ks >>= 1;
    return ctx->d_lut_seg ? ks : 0;
}
int seg_samples_per_thread(const dsdgpu_ctx* ctx) {
    return ctx->seg_narrow ? seg_samples_per_thread_of<SegLayoutNarrow>(ctx) : seg_samples_per_thread_of<SegLayout>(ctx);
}

template <class L>
int launch_seg_as(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, DecimateArgs a, cudaStream_t st) {
    if (!ctx->d_lut_seg) return fail(ctx, DSDGPU_ERR_ARG, "segmented kernel: no table images");
    const int ks = seg_samples_per_thread_of<L>(ctx);
    if (ks < 1) return fail(ctx, DSDGPU_ERR_ARG, "segmented kernel: %d channels x %d bytes per frame do not fit shared memory", ctx->nch, ctx->nstep);
    const int samples = L::NTHR / 2 * ks;                                     // per 256-thread group
    const int row_stride = (int)L::row_stride(samples, ctx->nch, ctx->nstep);
    const int smem = (int)L::smem_bytes(samples, ctx->nch, ctx->nstep);
    if (ctx->nseg > 1) { int rc = ensure_partial(ctx, scratch, a); if (rc) return rc; }
    for (int s = 0; s < ctx->nseg; ++s) {
        a.seg_first = s * L::SEG_DEFAULT;
        a.seg_n = ctx->nlut - a.seg_first < L::SEG_DEFAULT ? ctx->nlut - a.seg_first : L::SEG_DEFAULT;
        a.seg_init = s > 0 ? scratch.d_partial : nullptr;
        a.seg_raw = s + 1 < ctx->nseg ? scratch.d_partial : nullptr;
        a.lut = ctx->d_lut_seg + (size_t)s * L::LUT_BYTES;
        int rc = ks == 4 ? launch_seg_pass<4, L>(ctx, a, row_stride, smem, st)
               : ks == 2 ? launch_seg_pass<2, L>(ctx, a, row_stride, smem, st)
                         : launch_seg_pass<1, L>(ctx, a, row_stride, smem, st);
        if (rc) return rc;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

int launch_seg(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, const DecimateArgs& a, cudaStream_t st) {
    return ctx->seg_narrow ? launch_seg_as<SegLayoutNarrow>(ctx, scratch, a, st) : launch_seg_as<SegLayout>(ctx, scratch, a, st);
}

int check_call(dsdgpu_ctx* ctx, const void* in, int64_t in_count, int64_t nframes, int out_kind, const void* out) {
    if (!ctx) return DSDGPU_ERR_ARG;
    if (nframes < 0 || in_count < 0) return fail(ctx, DSDGPU_ERR_ARG, "negative size");
    if (nframes > 0 && !out) return fail(ctx, DSDGPU_ERR_ARG, "null output buffer");
    if (in_count > 0 && !in) return fail(ctx, DSDGPU_ERR_ARG, "null input buffer");
    if (out_kind != DSDGPU_OUT_INT32 && out_kind != DSDGPU_OUT_FLOAT64)
        return fail(ctx, DSDGPU_ERR_ARG, "unknown output kind %d", out_kind);
    return DSDGPU_OK;
}

int block_shift_of(long long block) {
    int sh = -1;
    if (block >= 16) for (sh = 0; (1LL << sh) < block; ++sh) {}
    return sh;
}

// planar rows for a run out of an interleaved device source: out row c = in[o*nch + c], o in [lo, hi)
int launch_deinterleave(dsdgpu_ctx* ctx, const uint8_t* d_raw, long long count, uint8_t* d_planar, long long stride,
                        int pad, cudaStream_t st) {
    if (count <= 0) return DSDGPU_OK;
    const long long words = (count + 3 + 3) / 4;
    long long blocks = (words + 255) / 256;
    if (blocks > 8LL * ctx->sm_count) blocks = 8LL * ctx->sm_count;
    deinterleave_kernel<<<(int)blocks, 256, 0, st>>>(d_raw, count, ctx->nch, d_planar, stride, pad);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

// enqueue dither fill (if any) + the decimation kernel on `st`; block_shift as in DecimateArgs
int enqueue_decimate(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, const uint8_t* d_in, size_t in_stride, int block_shift,
                     int64_t in_first, int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes,
                     double scale, double dither_amp, double clip, uint64_t dither_first_sample,
                     int out_kind, void* d_out, cudaStream_t st) {
    if (nframes == 0) return DSDGPU_OK;
    DecimateArgs a;
    a.in = d_in; a.in_stride = (long long)in_stride; a.block_shift = block_shift;
    a.in_first = in_first; a.in_count = in_count;
    a.p0 = p0; a.nframes = nframes; a.nch = ctx->nch; a.fill = fill;
    a.dither = nullptr; a.dither_stride = ctx->dither_stride > 0 ? ctx->dither_stride : ctx->nch;
    a.scale = scale; a.dither_amp = dither_amp; a.clip = clip;
    a.unit = ctx->lut_unit;
    a.out = d_out; a.out_kind = out_kind;
    a.seg_first = 0; a.seg_n = ctx->nlut; a.seg_init = nullptr; a.seg_raw = nullptr; a.plane_stride = 0;
    if (dither_amp > 0.0) {
        const long long count = (nframes - 1) * a.dither_stride + ctx->nch;
        int rc = ensure(ctx, &scratch.d_dither, &scratch.dither_cap, (size_t)count * sizeof(double));
        if (rc) return rc;
        rc = launch_dither(ctx, scratch, dither_first_sample, count, scratch.d_dither, st);
        if (rc) return rc;
        a.dither = scratch.d_dither;
    }
    const bool skew_ok = (ctx->nlut == 72 && ctx->nstep == 4 && ctx->d_lut_skew != nullptr);
    const bool ring_ok = ctx->d_lut_ring != nullptr;
    int kernel = ctx->kernel;
    if (kernel == DSDGPU_KERNEL_AUTO)
        kernel = ring_ok ? DSDGPU_KERNEL_RING : skew_ok ? DSDGPU_KERNEL_SKEW
               : seg_samples_per_thread(ctx) > 0 ? DSDGPU_KERNEL_SEGMENTED : DSDGPU_KERNEL_DIRECT;
    if (kernel == DSDGPU_KERNEL_RING) {
        if (!ring_ok) return fail(ctx, DSDGPU_ERR_ARG, "ring kernel needs the geometry of one of the reference's filters (72/4, 30/2, 12/1) or of the divide-by-64 extension (144/8), and tables without -0.0 entries");
        a.lut = ctx->d_lut_ring;
        if (ctx->nlut == 144) return launch_ring_two_pass<Ring64>(ctx, scratch, a, st);
        if (ctx->nlut == 30) return launch_ring<Ring30>(ctx, a, st);
        if (ctx->nlut == 12) return launch_ring<Ring12>(ctx, a, st);
        // (a 1024-thread instance of the 32-bit kernel fits in 64 registers but runs 13 % slower)
        return ctx->lut_precision == 32 ? launch_ring<Ring16x2Q>(ctx, a, st) : launch_ring<Ring16x2>(ctx, a, st);
    }
    if (kernel == DSDGPU_KERNEL_SKEW) {
        if (!skew_ok) return fail(ctx, DSDGPU_ERR_ARG, "skew kernel needs nlut=72, nstep=4");
        if (block_shift >= 0) return fail(ctx, DSDGPU_ERR_ARG, "skew kernel reads planar sources only");
        a.lut = ctx->d_lut_skew;
        return ctx->threads == 1024 ? launch_skew<Skew1024>(ctx, a, st) : launch_skew<Skew512>(ctx, a, st);
    }
    if (kernel == DSDGPU_KERNEL_SEGMENTED) return launch_seg(ctx, scratch, a, st);
    a.lut = ctx->d_lut_direct;
    return launch_direct(ctx, scratch, a, st);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int dsdgpu_create(dsdgpu_ctx** out, int device, int nch, int nlut, int nstep, const double* lut,
                  int lut_precision) {
    if (!out) return fail(nullptr, DSDGPU_ERR_ARG, "null ctx pointer");
    *out = nullptr;
    if (nch < 1 || nlut < 1 || nlut > DSDGPU_MAX_TABLES || nstep < 1 || nstep > 64 || !lut)
        return fail(nullptr, DSDGPU_ERR_ARG, "bad geometry nch=%d nlut=%d nstep=%d", nch, nlut, nstep);
    if (lut_precision != 64 && lut_precision != 32)
        return fail(nullptr, DSDGPU_ERR_ARG, "lut_precision must be 64 or 32");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, DSDGPU_ERR_NODEVICE, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev)
        return fail(nullptr, DSDGPU_ERR_NODEVICE, "device %d out of range (%d devices)", device, ndev);
    dsdgpu_ctx* ctx = new (std::nothrow) dsdgpu_ctx();
    if (!ctx) return fail(nullptr, DSDGPU_ERR_ARG, "out of host memory");
    ctx->device = device; ctx->nch = nch; ctx->nlut = nlut; ctx->nstep = nstep;
    ctx->lut_precision = lut_precision;
#define CREATE_TRY(call)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fail(nullptr, DSDGPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));    \
            dsdgpu_destroy(ctx);                                                               \
            return DSDGPU_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int s = 0; s < DSDGPU_SLOTS; ++s) CREATE_TRY(cudaStreamCreateWithFlags(&ctx->slots[s].stream, cudaStreamNonBlocking));

    // Tables.  lut_precision 32 = 32-bit tables: every entry rounded to a multiple of 2^-qbits, qbits
    // the largest value <= 30 for which no sum of nlut entries can leave int32 (30 for the reference's
    // filters, whose largest possible |sum| is 1.33).  Sums of such entries are exact in fp64 too, so
    // every kernel -- integer or fp64 accumulation -- returns the same bits for this variant.
    std::vector<double> tab((size_t)nlut * 256);
    if (lut_precision == 32) {
        double worst = 0.0;
        for (int t = 0; t < nlut; ++t) {
            double m = 0.0;
            for (int e = 0; e < 256; ++e) m = std::fmax(m, std::fabs(lut[(size_t)t * 256 + e]));
            worst += m;
        }
        int qbits = 30;
        while (qbits > 0 && !(worst * std::ldexp(1.0, qbits) + nlut < 2147483647.0)) --qbits;
        if (!(worst * std::ldexp(1.0, qbits) + nlut < 2147483647.0)) {
            fail(nullptr, DSDGPU_ERR_ARG, "table entries too large for the 32-bit variant");
            dsdgpu_destroy(ctx);
            return DSDGPU_ERR_ARG;
        }
        ctx->lut_unit = std::ldexp(1.0, -qbits);
        for (size_t i = 0; i < tab.size(); ++i) tab[i] = std::nearbyint(lut[i] / ctx->lut_unit) * ctx->lut_unit;
    } else {
        for (size_t i = 0; i < tab.size(); ++i) tab[i] = lut[i];
    }
    CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_lut_direct), tab.size() * sizeof(double)));
    CREATE_TRY(cudaMemcpy(ctx->d_lut_direct, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    // The fp64 ring kernels route a value into one of two sums as fma(x, 0 or 1, acc): bit-exact
    // unless x is -0.0 (or anything is non-finite).  No real filter has such a table entry; a table
    // that does simply keeps the older kernels.  (Integer tables have no such case.)
    bool plain = true;
    for (double v : tab) plain = plain && std::isfinite(v) && !(v == 0.0 && std::signbit(v));
    std::vector<unsigned char> rimg;
    if (nlut == 72 && nstep == 4) {
        std::vector<unsigned char> img(Skew512::LUT_BYTES);
        skew_build_lut_image<Skew512>(tab.data(), img.data());
        CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_lut_skew), img.size()));
        CREATE_TRY(cudaMemcpy(ctx->d_lut_skew, img.data(), img.size(), cudaMemcpyHostToDevice));
        if (lut_precision == 32) {
            rimg.resize(Ring16x2Q::LUT_BYTES);
            ring_build_lut_image<Ring16x2Q>(tab.data(), rimg.data(), ctx->lut_unit);
        } else if (plain) {
            rimg.resize(Ring16x2::LUT_BYTES);
            ring_build_lut_image<Ring16x2>(tab.data(), rimg.data());
        }
    } else if (nlut == 30 && nstep == 2 && plain) {
        rimg.resize(Ring30::LUT_BYTES);
        ring_build_lut_image<Ring30>(tab.data(), rimg.data());
    } else if (nlut == 12 && nstep == 1 && plain) {
        rimg.resize(Ring12::LUT_BYTES);
        ring_build_lut_image<Ring12>(tab.data(), rimg.data());
    } else if (nlut == 144 && nstep == 8 && plain && lut_precision == 64) {
        // divide-by-64 extension filter: one image per pass of 72 tables
        rimg.resize(2 * (size_t)Ring64::LUT_BYTES);
        ring_build_lut_image<Ring64>(tab.data(), rimg.data());
        ring_build_lut_image<Ring64>(tab.data() + (size_t)72 * 256, rimg.data() + Ring64::LUT_BYTES);
    }
    if (!rimg.empty()) {
        CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_lut_ring), rimg.size()));
        CREATE_TRY(cudaMemcpy(ctx->d_lut_ring, rimg.data(), rimg.size(), cudaMemcpyHostToDevice));
    }
    {
        // segmented kernel: one image per SEG_DEFAULT tables (fp64 entries; the 32-bit variant's rounded
        // tables sum exactly in fp64 too, so it returns the integer kernel's bits)
        ctx->seg_narrow = nstep >= 16;
        const int seg = ctx->seg_narrow ? SegLayoutNarrow::SEG_DEFAULT : SegLayout::SEG_DEFAULT;
        const size_t img_bytes = ctx->seg_narrow ? SegLayoutNarrow::LUT_BYTES : SegLayout::LUT_BYTES;
        ctx->nseg = (nlut + seg - 1) / seg;
        std::vector<unsigned char> simg((size_t)ctx->nseg * img_bytes);
        for (int sidx = 0; sidx < ctx->nseg; ++sidx) {
            const int first = sidx * seg, count = nlut - first < seg ? nlut - first : seg;
            if (ctx->seg_narrow) seg_build_lut_image<SegLayoutNarrow>(tab.data(), first, count, simg.data() + (size_t)sidx * img_bytes);
            else seg_build_lut_image<SegLayout>(tab.data(), first, count, simg.data() + (size_t)sidx * img_bytes);
        }
        CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_lut_seg), simg.size()));
        CREATE_TRY(cudaMemcpy(ctx->d_lut_seg, simg.data(), simg.size(), cudaMemcpyHostToDevice));
    }
    // dither generator constants
    {
        uint32_t words[3 + kRandBaseWords];
        glibc_initial_words(words, 3 + kRandBaseWords);
        ctx->rand_base.assign(words + 3, words + 3 + kRandBaseWords);
        std::vector<uint32_t> pow2((size_t)kRandPowBits * kRandDeg, 0);
        uint32_t cur[kRandDeg] = {0};
        cur[1] = 1;                               // x^(2^0)
        for (int b = 0; b < kRandPowBits; ++b) {
            memcpy(&pow2[(size_t)b * kRandDeg], cur, sizeof(cur));
            uint32_t sq[kRandDeg];
            rand_poly_mul(cur, cur, sq);
            memcpy(cur, sq, sizeof(cur));
        }
        // x^e from the squarings
        auto power = [&](unsigned long long e, uint32_t* out) {
            uint32_t acc[kRandDeg] = {0}, tmp[kRandDeg];
            acc[0] = 1;
            for (int b = 0; b < kRandPowBits && (e >> b) != 0; ++b)
                if ((e >> b) & 1ULL) { rand_poly_mul(acc, &pow2[(size_t)b * kRandDeg], tmp); memcpy(acc, tmp, sizeof(acc)); }
            memcpy(out, acc, sizeof(acc));
        };
        std::vector<uint32_t> lane_polys((size_t)kRandDeg * 32), jump((size_t)3 * kRandLevel * kRandDeg);
        uint32_t step[kRandDeg], acc[kRandDeg], tmp[kRandDeg];
        ctx->rand_pow16.assign((size_t)12 * 16 * kRandDeg, 0);          // x^(d * 16^k) for the host-side anchor
        for (int k = 0; k < 12; ++k) {
            memset(acc, 0, sizeof(acc)); acc[0] = 1;
            for (int d = 0; d < 16; ++d) {
                memcpy(&ctx->rand_pow16[((size_t)k * 16 + d) * kRandDeg], acc, sizeof(acc));
                rand_poly_mul(acc, &pow2[(size_t)4 * k * kRandDeg], tmp); memcpy(acc, tmp, sizeof(acc));
            }
        }
        power(kRandLaneValues, step);
        memset(acc, 0, sizeof(acc)); acc[0] = 1;
        for (int l = 0; l < 32; ++l) {
            for (int k = 0; k < kRandDeg; ++k) lane_polys[(size_t)k * 32 + l] = acc[k];
            rand_poly_mul(acc, step, tmp); memcpy(acc, tmp, sizeof(acc));
        }
        unsigned long long stride = kRandWarpValues;
        for (int level = 0; level < 3; ++level, stride *= kRandLevel) {
            power(stride, step);
            memset(acc, 0, sizeof(acc)); acc[0] = 1;
            for (int u = 0; u < kRandLevel; ++u) {
                memcpy(&jump[((size_t)level * kRandLevel + u) * kRandDeg], acc, sizeof(acc));
                rand_poly_mul(acc, step, tmp); memcpy(acc, tmp, sizeof(acc));
            }
        }
        CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_rand_lane), lane_polys.size() * sizeof(uint32_t)));
        CREATE_TRY(cudaMemcpy(ctx->d_rand_lane, lane_polys.data(), lane_polys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->d_rand_jump), jump.size() * sizeof(uint32_t)));
        CREATE_TRY(cudaMemcpy(ctx->d_rand_jump, jump.data(), jump.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
#undef CREATE_TRY
    *out = ctx;
    return DSDGPU_OK;
}

void dsdgpu_destroy(dsdgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int s = 0; s < DSDGPU_SLOTS; ++s) free_slot(ctx->slots[s]);
    free_slot(ctx->own);
    if (ctx->d_lut_direct) cudaFree(ctx->d_lut_direct);
    if (ctx->d_lut_skew) cudaFree(ctx->d_lut_skew);
    if (ctx->d_lut_ring) cudaFree(ctx->d_lut_ring);
    if (ctx->d_lut_seg) cudaFree(ctx->d_lut_seg);
    if (ctx->d_rand_lane) cudaFree(ctx->d_rand_lane);
    if (ctx->d_rand_jump) cudaFree(ctx->d_rand_jump);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dsdgpu_last_error(const dsdgpu_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int dsdgpu_set_option(dsdgpu_ctx* ctx, const char* key, long long value) {
    if (!ctx || !key) return DSDGPU_ERR_ARG;
    if (!strcmp(key, "kernel")) {
        if (value < DSDGPU_KERNEL_AUTO || value > DSDGPU_KERNEL_SEGMENTED) return fail(ctx, DSDGPU_ERR_ARG, "bad kernel %lld", value);
        ctx->kernel = (int)value;
    } else if (!strcmp(key, "threads")) {
        if (value != 512 && value != 1024) return fail(ctx, DSDGPU_ERR_ARG, "threads must be 512 or 1024");
        ctx->threads = (int)value;
    } else if (!strcmp(key, "dither_stride")) {
        if (value != 0 && value < ctx->nch) return fail(ctx, DSDGPU_ERR_ARG, "dither_stride must be 0 or >= nch");
        ctx->dither_stride = (int)value;
    } else if (!strcmp(key, "source_block_bytes")) {
        const bool pow2 = value > 0 && (value & (value - 1)) == 0;
        if (!(value == 0 || value == 1 || (pow2 && value >= 16 && value <= (1LL << 30))))
            return fail(ctx, DSDGPU_ERR_ARG, "source_block_bytes must be 0 (planar), 1 (interleaved) or a power of two >= 16");
        ctx->source_block = value;
    } else if (!strcmp(key, "slots")) {
        if (value < 1 || value > DSDGPU_SLOTS) return fail(ctx, DSDGPU_ERR_ARG, "slots must be 1..%d", DSDGPU_SLOTS);
        ctx->pipeline_slots = (int)value;
    } else if (!strcmp(key, "trace")) {
        ctx->trace = value != 0;
    } else if (!strcmp(key, "chunk_frames")) {
        if (value < 1) return fail(ctx, DSDGPU_ERR_ARG, "chunk_frames must be >= 1");
        ctx->chunk_frames = value;
    } else {
        return fail(ctx, DSDGPU_ERR_ARG, "unknown option '%s'", key);
    }
    return DSDGPU_OK;
}

unsigned long long dsdgpu_kernel_launches(const dsdgpu_ctx* ctx) { return ctx ? ctx->launches : 0ULL; }

void* dsdgpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void dsdgpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int dsdgpu_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return DSDGPU_ERR_ARG;
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? DSDGPU_OK : DSDGPU_ERR_CUDA;
}
int dsdgpu_host_unregister(void* p) {
    if (!p) return DSDGPU_ERR_ARG;
    return cudaHostUnregister(p) == cudaSuccess ? DSDGPU_OK : DSDGPU_ERR_CUDA;
}

// the window of stream bytes a run of frames can touch, clipped to what the caller has: [lo, hi)
static void run_window(const dsdgpu_ctx* ctx, int64_t in_first, int64_t in_count, int64_t p0, int64_t nframes,
                       int64_t& lo, int64_t& hi) {
    lo = p0 - (ctx->nlut - 1);
    hi = p0 + (nframes - 1) * ctx->nstep + 1;
    if (lo < in_first) lo = in_first;
    if (hi > in_first + in_count) hi = in_first + in_count;
    if (hi < lo) hi = lo;
}

// Interleaved device bytes d_raw[(j - raw_first)*nch + c] for j in [lo, hi) -> planar rows in
// scratch.d_in, placed like dsdgpu_submit places host bytes (frame grid 16-byte aligned; pad bytes
// that stand for positions before the source are set to `fill`).  Returns the planar view.
static int planar_from_interleaved(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, const uint8_t* d_raw, int64_t raw_first,
                                   int64_t lo, int64_t hi, int64_t p0, uint8_t fill, cudaStream_t st,
                                   size_t& dstride, int64_t& first, int64_t& count) {
    const int64_t n = hi - lo;
    const int64_t pad = ((lo - p0 - 1) % 16 + 16) % 16;
    dstride = ((size_t)(n + pad) + 255) & ~(size_t)255;
    int rc = ensure(ctx, &scratch.d_in, &scratch.in_cap, dstride * ctx->nch + 256);
    if (rc) return rc;
    if (pad > 0 && p0 - (ctx->nlut - 1) < lo)
        CUDA_TRY(ctx, cudaMemset2DAsync(scratch.d_in, dstride, fill, (size_t)pad, (size_t)ctx->nch, st));
    rc = launch_deinterleave(ctx, d_raw + (size_t)(lo - raw_first) * ctx->nch, n, scratch.d_in, (long long)dstride, (int)pad, st);
    first = lo - pad;
    count = n + pad;
    return rc;
}

int dsdgpu_decimate_device(dsdgpu_ctx* ctx, const uint8_t* d_in, size_t in_stride, int64_t in_first,
                           int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes, double scale,
                           double dither_amp, double clip, uint64_t dither_first_sample, int out_kind,
                           void* d_out, void* cuda_stream) {
    int rc = check_call(ctx, d_in, in_count, nframes, out_kind, d_out);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    if (ctx->source_block == 1 && nframes > 0) {
        int64_t lo, hi, first, count;
        size_t dstride;
        run_window(ctx, in_first, in_count, p0, nframes, lo, hi);
        rc = planar_from_interleaved(ctx, ctx->own, d_in, in_first, lo, hi, p0, fill, st, dstride, first, count);
        if (rc) return rc;
        rc = enqueue_decimate(ctx, ctx->own, ctx->own.d_in, dstride, -1, first, count, fill, p0, nframes, scale,
                              dither_amp, clip, dither_first_sample, out_kind, d_out, st);
    } else {
        rc = enqueue_decimate(ctx, ctx->own, d_in, in_stride, block_shift_of(ctx->source_block), in_first, in_count, fill,
                              p0, nframes, scale, dither_amp, clip, dither_first_sample, out_kind, d_out, st);
    }
    if (rc) return rc;
    if (!cuda_stream) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return DSDGPU_OK;
}

// What a kernel reads after a host window has been brought to the device.
struct SourceView {
    const uint8_t* d = nullptr;
    size_t stride = 0;
    int block_shift = -1;
    int64_t first = 0, count = 0;
};

// Upload stream bytes [lo, hi) (already clipped to the caller's window) of every channel in the
// context's source layout on slot `s`.  `oldest` = the oldest stream index the run reads, `grid` =
// a stream index the device buffer should be 16-byte aligned to (planar / interleaved sources).
static int stage_source(dsdgpu_ctx* ctx, dsdgpu_slot& s, const uint8_t* in, size_t in_stride, int64_t in_first,
                        int64_t in_count, uint8_t fill, int64_t lo, int64_t hi, int64_t oldest, int64_t grid,
                        SourceView& v) {
    const int64_t n = hi - lo;
    int rc;
    if (ctx->source_block >= 16) {
        // DSF-style source: whole blocks of every channel are one contiguous range of the host buffer
        const int64_t B = ctx->source_block;
        if (in_first % B != 0) return fail(ctx, DSDGPU_ERR_ARG, "block sources must start on a block boundary");
        const int64_t b0 = (lo - in_first) / B, b1 = n > 0 ? (hi - in_first + B - 1) / B : b0;
        const size_t bytes = (size_t)(b1 - b0) * (size_t)B * ctx->nch;
        rc = ensure(ctx, &s.d_in, &s.in_cap, bytes + 256);
        if (rc) return rc;
        if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(s.d_in, in + (size_t)b0 * B * ctx->nch, bytes, cudaMemcpyHostToDevice, s.stream));
        int64_t cnt = in_count - b0 * B;
        if (cnt > (b1 - b0) * B) cnt = (b1 - b0) * B;
        v.d = s.d_in; v.stride = 0; v.block_shift = block_shift_of(B); v.first = in_first + b0 * B; v.count = cnt;
        return DSDGPU_OK;
    }
    // Planar rows on the device, placed so that (grid - first) is a multiple of 16, `first` being the
    // stream index of the first byte of the device buffer: with grid = p0 + 1 the ring kernel's window
    // loads are bank-conflict free (ring_core.h).  The pad bytes in front count as part of the source
    // (d_in stays 16-byte aligned): either they are older than anything the run reads, or -- when the
    // run starts before the caller's first byte -- they are set to the fill byte they stand for.
    const int64_t pad = ((lo - grid) % 16 + 16) % 16;
    const size_t dstride = ((size_t)(n + pad) + 255) & ~(size_t)255;
    rc = ensure(ctx, &s.d_in, &s.in_cap, dstride * ctx->nch + 256);
    if (rc) return rc;
    if (pad > 0 && oldest < lo)
        CUDA_TRY(ctx, cudaMemset2DAsync(s.d_in, dstride, fill, (size_t)pad, (size_t)ctx->nch, s.stream));
    if (ctx->source_block == 1) {
        // DSDIFF-style source: upload the interleaved range as it is, split it on the device
        const size_t bytes = (size_t)n * ctx->nch;
        rc = ensure(ctx, &s.d_raw, &s.raw_cap, bytes + 256);
        if (rc) return rc;
        if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(s.d_raw, in + (size_t)(lo - in_first) * ctx->nch, bytes, cudaMemcpyHostToDevice, s.stream));
        rc = launch_deinterleave(ctx, s.d_raw, n, s.d_in, (long long)dstride, (int)pad, s.stream);
        if (rc) return rc;
    } else if (n > 0) {
        CUDA_TRY(ctx, cudaMemcpy2DAsync(s.d_in + pad, dstride ? dstride : 256, in + (lo - in_first), in_stride, (size_t)n,
                                        (size_t)ctx->nch, cudaMemcpyHostToDevice, s.stream));
    }
    v.d = s.d_in; v.stride = dstride; v.block_shift = -1; v.first = lo - pad; v.count = n + pad;
    return DSDGPU_OK;
}

int dsdgpu_submit(dsdgpu_ctx* ctx, int slot, const uint8_t* in, size_t in_stride, int64_t in_first,
                  int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes, double scale,
                  double dither_amp, double clip, uint64_t dither_first_sample, int out_kind, void* out) {
    int rc = check_call(ctx, in, in_count, nframes, out_kind, out);
    if (rc) return rc;
    if (slot < 0 || slot >= DSDGPU_SLOTS) return fail(ctx, DSDGPU_ERR_ARG, "slot must be 0..%d", DSDGPU_SLOTS - 1);
    dsdgpu_slot& s = ctx->slots[slot];
    if (s.busy) return fail(ctx, DSDGPU_ERR_STATE, "slot %d is still in flight", slot);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (nframes == 0) { s.busy = true; return DSDGPU_OK; }
    const size_t elem = out_kind == DSDGPU_OUT_INT32 ? sizeof(int32_t) : sizeof(double);
    const size_t out_bytes = (size_t)nframes * ctx->nch * elem;
    rc = ensure(ctx, reinterpret_cast<unsigned char**>(&s.d_out), &s.out_cap, out_bytes);
    if (rc) return rc;
    // bytes this run of frames can touch, clipped to what the caller has
    int64_t lo, hi;
    run_window(ctx, in_first, in_count, p0, nframes, lo, hi);
    SourceView v;
    dsdgpu_ctx::TraceRow tr;
    if (ctx->trace) {
        for (auto& e : tr.e) CUDA_TRY(ctx, cudaEventCreate(&e));
        tr.slot = slot; tr.frames = nframes;
        CUDA_TRY(ctx, cudaEventRecord(tr.e[0], s.stream));
    }
    rc = stage_source(ctx, s, in, in_stride, in_first, in_count, fill, lo, hi, p0 - (ctx->nlut - 1), p0 + 1, v);
    if (rc) return rc;
    if (ctx->trace) CUDA_TRY(ctx, cudaEventRecord(tr.e[1], s.stream));
    rc = enqueue_decimate(ctx, s, v.d, v.stride, v.block_shift, v.first, v.count, fill, p0, nframes, scale, dither_amp,
                          clip, dither_first_sample, out_kind, s.d_out, s.stream);
    if (rc) return rc;
    if (ctx->trace) CUDA_TRY(ctx, cudaEventRecord(tr.e[2], s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
    if (ctx->trace) { CUDA_TRY(ctx, cudaEventRecord(tr.e[3], s.stream)); ctx->trace_rows.push_back(tr); }
    s.busy = true;
    return DSDGPU_OK;
}

// prints and clears the timeline collected since the last call (times in ms from the first chunk's submission point)
static void dump_trace(dsdgpu_ctx* ctx) {
    if (ctx->trace_rows.empty()) return;
    cudaEvent_t base = ctx->trace_rows[0].e[0];
    fprintf(stderr, "# dsdgpu trace: chunk slot frames | upload starts, upload done, kernel done, download done [ms]\n");
    for (size_t k = 0; k < ctx->trace_rows.size(); ++k) {
        float t[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], base, ctx->trace_rows[k].e[i]);
        fprintf(stderr, "%3zu %d %8lld | %7.3f %7.3f %7.3f %7.3f\n", k, ctx->trace_rows[k].slot, ctx->trace_rows[k].frames, t[0], t[1], t[2], t[3]);
    }
    for (auto& r : ctx->trace_rows)
        for (auto& e : r.e) cudaEventDestroy(e);
    ctx->trace_rows.clear();
}

static int launch_dop(dsdgpu_ctx* ctx, const SourceView& v, uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits,
                      int32_t* d_out, cudaStream_t st) {
    if (nframes == 0) return DSDGPU_OK;
    DecimateArgs a = {};
    a.in = v.d; a.in_stride = (long long)v.stride; a.block_shift = v.block_shift;
    a.in_first = v.first; a.in_count = v.count; a.nframes = nframes; a.nch = ctx->nch; a.fill = fill; a.out = d_out;
    const long long total = nframes * ctx->nch;
    // mono / stereo with everything 4-byte aligned: four output samples per thread, word loads, 16-byte stores
    const bool wide = (ctx->nch == 1 || ctx->nch == 2) && ((pm0 - v.first) & 3) == 0 &&
                      ((reinterpret_cast<unsigned long long>(v.d) | (v.block_shift < 0 ? (unsigned long long)v.stride : 0ULL)) & 3ULL) == 0 &&
                      (reinterpret_cast<unsigned long long>(d_out) & 15ULL) == 0;
    if (wide) {
        const long long groups = (total + 3) / 4;
        long long blocks = (groups + 4 * 256 - 1) / (4 * 256);       // four groups per thread per pass
        if (blocks > 16LL * ctx->sm_count) blocks = 16LL * ctx->sm_count;
        if (blocks < 1) blocks = 1;
        if (ctx->nch == 1) dop_pack_wide_kernel<1><<<(int)blocks, 256, 0, st>>>(a, pm0, reverse_bits);
        else dop_pack_wide_kernel<2><<<(int)blocks, 256, 0, st>>>(a, pm0, reverse_bits);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
        return DSDGPU_OK;
    }
    long long blocks = (total + 255) / 256;
    if (blocks > 32LL * ctx->sm_count) blocks = 32LL * ctx->sm_count;
    dop_pack_kernel<<<(int)blocks, 256, 0, st>>>(a, pm0, reverse_bits);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    return DSDGPU_OK;
}

int dsdgpu_dop_pack_device(dsdgpu_ctx* ctx, const uint8_t* d_in, size_t in_stride, int64_t in_first, int64_t in_count,
                           uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits, int32_t* d_out,
                           void* cuda_stream) {
    int rc = check_call(ctx, d_in, in_count, nframes, DSDGPU_OUT_INT32, d_out);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    SourceView v;
    if (ctx->source_block == 1) {
        int64_t lo = pm0, hi = pm0 + 2 * nframes + 1, first, count;
        if (lo < in_first) lo = in_first;
        if (hi > in_first + in_count) hi = in_first + in_count;
        if (hi < lo) hi = lo;
        size_t dstride;
        rc = planar_from_interleaved(ctx, ctx->own, d_in, in_first, lo, hi, lo - 1, fill, st, dstride, first, count);
        if (rc) return rc;
        v.d = ctx->own.d_in; v.stride = dstride; v.first = first; v.count = count;
    } else {
        v.d = d_in; v.stride = in_stride; v.block_shift = block_shift_of(ctx->source_block); v.first = in_first; v.count = in_count;
    }
    rc = launch_dop(ctx, v, fill, pm0, nframes, reverse_bits, d_out, st);
    if (rc) return rc;
    if (!cuda_stream) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return DSDGPU_OK;
}

int dsdgpu_dop_pack_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first, int64_t in_count,
                         uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits, int32_t* out) {
    int rc = check_call(ctx, in, in_count, nframes, DSDGPU_OUT_INT32, out);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const long long chunk = ctx->chunk_frames;
    const int nslots = ctx->pipeline_slots;
    bool inflight[DSDGPU_SLOTS] = {false};
    long long k = 0;
    for (int64_t f = 0; f < nframes && !rc; f += chunk, ++k) {
        const int slot = (int)(k % nslots);
        dsdgpu_slot& s = ctx->slots[slot];
        if (inflight[slot]) { rc = dsdgpu_wait(ctx, slot); inflight[slot] = false; if (rc) break; }
        if (s.busy) { rc = fail(ctx, DSDGPU_ERR_STATE, "slot %d is still in flight", slot); break; }
        const int64_t n = nframes - f < chunk ? nframes - f : chunk;
        const int64_t pm = pm0 + 2 * f;
        int64_t lo = pm, hi = pm + 2 * n + 1;                       // frames f .. f+n-1 read bytes pm .. pm + 2n
        if (lo < in_first) lo = in_first;
        if (hi > in_first + in_count) hi = in_first + in_count;
        if (hi < lo) hi = lo;
        const size_t out_bytes = (size_t)n * ctx->nch * sizeof(int32_t);
        rc = ensure(ctx, reinterpret_cast<unsigned char**>(&s.d_out), &s.out_cap, out_bytes);
        if (rc) break;
        SourceView v;
        rc = stage_source(ctx, s, in, in_stride, in_first, in_count, fill, lo, hi, pm, lo, v);
        if (rc) break;
        rc = launch_dop(ctx, v, fill, pm, n, reverse_bits, static_cast<int32_t*>(s.d_out), s.stream);
        if (rc) break;
        if (cudaMemcpyAsync(out + (size_t)f * ctx->nch, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess) {
            rc = fail(ctx, DSDGPU_ERR_CUDA, "cudaMemcpyAsync failed");
            break;
        }
        s.busy = true;
        inflight[slot] = true;
    }
    for (int sidx = 0; sidx < DSDGPU_SLOTS; ++sidx)
        if (inflight[sidx]) { int r2 = dsdgpu_wait(ctx, sidx); if (!rc) rc = r2; }
    return rc;
}

int dsdgpu_wait(dsdgpu_ctx* ctx, int slot) {
    if (!ctx) return DSDGPU_ERR_ARG;
    if (slot < 0 || slot >= DSDGPU_SLOTS) return fail(ctx, DSDGPU_ERR_ARG, "slot must be 0..%d", DSDGPU_SLOTS - 1);
    dsdgpu_slot& s = ctx->slots[slot];
    if (!s.busy) return fail(ctx, DSDGPU_ERR_STATE, "slot %d has nothing in flight", slot);
    s.busy = false;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(s.stream));
    return DSDGPU_OK;
}

int dsdgpu_decimate_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first,
                         int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes, double scale,
                         double dither_amp, double clip, uint64_t dither_first_sample, int out_kind,
                         void* out) {
    int rc = check_call(ctx, in, in_count, nframes, out_kind, out);
    if (rc) return rc;
    const size_t elem = out_kind == DSDGPU_OUT_INT32 ? sizeof(int32_t) : sizeof(double);
    const long long chunk = ctx->chunk_frames;
    long long k = 0;
    const int nslots = ctx->pipeline_slots;
    bool inflight[DSDGPU_SLOTS] = {false};
    for (int64_t f = 0; f < nframes; f += chunk, ++k) {
        const int slot = (int)(k % nslots);
        if (inflight[slot]) { rc = dsdgpu_wait(ctx, slot); inflight[slot] = false; if (rc) break; }
        const int64_t n = nframes - f < chunk ? nframes - f : chunk;
        rc = dsdgpu_submit(ctx, slot, in, in_stride, in_first, in_count, fill, p0 + f * ctx->nstep, n, scale,
                           dither_amp, clip, dither_first_sample + (uint64_t)f * (uint64_t)(ctx->dither_stride > 0 ? ctx->dither_stride : ctx->nch), out_kind,
                           static_cast<unsigned char*>(out) + (size_t)f * ctx->nch * elem);
        if (rc) break;
        inflight[slot] = true;
    }
    for (int s = 0; s < DSDGPU_SLOTS; ++s)
        if (inflight[s]) { int r2 = dsdgpu_wait(ctx, s); if (!rc) rc = r2; }
    if (ctx->trace) dump_trace(ctx);
    return rc;
}

int dsdgpu_dither_fill_device(dsdgpu_ctx* ctx, uint64_t first_sample, int64_t count, double* d_out, void* cuda_stream) {
    if (!ctx) return DSDGPU_ERR_ARG;
    if (count < 0 || (count > 0 && !d_out)) return fail(ctx, DSDGPU_ERR_ARG, "bad dither request");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    int rc = launch_dither(ctx, ctx->own, first_sample, count, d_out, st);
    if (rc) return rc;
    if (!cuda_stream) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return DSDGPU_OK;
}

}  // extern "C"
