// dsd_kernels.cuh -- the sm_100a kernels of the decimation path (included once, by dsdgpu.cu).
//
//   decimate_ring_kernel    production kernel (ring_core.h): divide-by-32 / -16 / -8 filters, 32-bit tables, and the
//                           two 72-table passes of the divide-by-64 extension filter
//   decimate_seg_kernel     any geometry, <= 72 (64) tables per pass over a plane of running sums (seg_core.h)
//   decimate_direct_kernel  one thread per output sample, reference loop order: the cross-check partner
//   deinterleave_kernel     DSDIFF byte-interleaved source -> planar rows
//   dop_pack_*_kernel       DoP packing (reference dop_packer.cpp:100-141)
// Reference path replaced: DsdDecimator::getSamplesInternal (reference src/dsd_decimator.cpp:173-222).
#pragma once
#include "dsdgpu.h"
#include "ring_core.h"
#include "seg_core.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

using namespace dsdgpu;

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
    double clip_hi;          // clip, or +infinity when clip <= 0 (no clipping): min / max instead of compare-and-select
    int clip_lo_i, clip_hi_i;   // the same bounds after rounding: -round(clip), round(clip) saturated to int32; INT_MIN, INT_MAX without clipping
    double unit;             // 32-bit tables: value of one table unit (2^-qbits); 1.0 otherwise
    void* out;
    int out_kind;
    int pcm_bytes;           // DSDGPU_OUT_PCM24 / _PCM16: 3 / 2 bytes per sample (else 0)
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

// sample m of a packed PCM output: the low `bytes` bytes of the rounded integer, little endian
__device__ __forceinline__ void store_pcm(unsigned char* out, long long m, int q, int bytes) {
    unsigned char* p = out + m * bytes;
    p[0] = (unsigned char)q;
    p[1] = (unsigned char)(q >> 8);
    if (bytes == 3) p[2] = (unsigned char)(q >> 16);
}

__device__ __forceinline__ void store_sample(const DecimateArgs& a, long long frame, int c, double sum) {
    const long long m = frame * a.nch + c;
    const double dith = (a.dither != nullptr) ? a.dither[frame * a.dither_stride + c] : 0.0;
    const double v = quantise_tail(sum, a.scale, a.dither_amp, dith, a.clip);
    if (a.out_kind == DSDGPU_OUT_FLOAT64) {
        reinterpret_cast<double*>(a.out)[m] = v;
        return;
    }
    // C round(): half away from zero (dsd_decimator.cpp:213-214)
    const int q = __double2int_rz(round(v));
    if (a.out_kind == DSDGPU_OUT_INT32) reinterpret_cast<int32_t*>(a.out)[m] = q;
    else store_pcm(reinterpret_cast<unsigned char*>(a.out), m, q, a.pcm_bytes);
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
// staging helpers: cp.async, mbarriers
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
    // running sums this thread's samples start from (later passes): loaded one tile ahead of their use
    double init_next[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        const long long i = tl.i0 + gt + k * GT;
        init_next[k] = (a.seg_init && tile < ntiles) ? a.seg_init[i < total ? i : total - 1] : 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();                 // the table image (loaded by all 512 threads) is in place for both groups
    while (tile < ntiles) {
        const long long next = tile + tstride;
        double acc[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) acc[k] = init_next[k];
        if (next < ntiles) {
            const SegTile tn = seg_tile_at<KS>(a, next, nstep, nsteps);
            seg_stage_tile(a, tn.t0, row_stride, bufs + (buf ^ 1) * tile_bytes, aligned16, gt);
            if (a.seg_init) {
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const long long i = tn.i0 + gt + k * GT;
                    init_next[k] = a.seg_init[i < total ? i : total - 1];
                }
            }
        }
        cp_async_commit();
        cp_async_wait<1>();          // everything but the group just committed has landed
        seg_group_barrier(group);

        const uint32_t tile_off = (uint32_t)(L::LUT_BYTES + (group * 2 + buf) * tile_bytes);
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
            lo[k] = ring_lds_u32(smem, wlo[k] + 4u);
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
                lo[k] = ring_lds_u32(smem, wlo[k]);
                wlo[k] -= 4u;
                x[k] = ring_funnel_r(lo[k], hi, shift[k]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    uint32_t e;
                    if (u == 0) e = ring_byte<3>(x[k]);
                    else if (u == 1) e = ring_byte<2>(x[k]);
                    else if (u == 2) e = ring_byte<1>(x[k]);
                    else e = ring_byte<0>(x[k]);
                    acc[k] = __dadd_rn(acc[k], ring_lds_f64(smem, e * (uint32_t)L::ROW_BYTES + lut + 8u * (uint32_t)u));
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
// Bulk copy (TMA, cp.async.bulk: SASS UBLKCP) of `bytes` contiguous bytes into shared memory, completion counted in bytes
// on an mbarrier of this CTA; source, destination and size multiples of 16.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one arrival of this thread + `bytes` more transaction bytes to wait for
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Stage both halves of a tile: bytes [t0_h, t0_h + HALF_BYTES) of channel c_h, and arrive on the tile's `full` barrier.
// A tile whose halves lie inside a planar, 16-byte aligned source -- every tile but those at the ends of the stream, for
// every buffer the host entry points build -- is ONE thread's work: two bulk copies (TMA) that report to the barrier in
// bytes; the other 511 threads just arrive.  (What a tile costs a thread besides its table loads -- the staging loop,
// its 64-bit geometry -- is per period for the instances with one round per tile.)  Other tiles: every thread copies its
// 16-byte chunks (cp.async; edges byte by byte with the fill value) and arrives when they have landed.
// rel0: source offset (from a.in_first) of byte 0 of a half buffer whose first frame is frame 0 -- a half that starts at
// frame f0 lies f0 * FBYTES further on (whole 16-byte chunks: ring_half_origin).
template <class G>
__device__ __forceinline__ void ring_stage_tile(const DecimateArgs& a, const RingTile& tl, unsigned char* dst,
                                                bool aligned16, uint64_t* full_bar, long long rel0) {
    constexpr int CHUNKS = G::HALF_BYTES / 16;
    if (aligned16 && a.block_shift < 0) {
        const long long o0 = rel0 + tl.f0[0] * G::FBYTES, o1 = rel0 + tl.f0[1] * G::FBYTES;
        if (o0 >= 0 && o1 >= 0 && o0 + G::HALF_BYTES <= a.in_count && o1 + G::HALF_BYTES <= a.in_count) {   // CTA-uniform
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(full_bar, 2u * G::HALF_BYTES);
                bulk_copy_g2s(dst, a.in + (long long)tl.c[0] * a.in_stride + o0, G::HALF_BYTES, full_bar);
                bulk_copy_g2s(dst + G::HALF_BYTES, a.in + (long long)tl.c[1] * a.in_stride + o1, G::HALF_BYTES, full_bar);
            } else {
                mbar_arrive(full_bar);
            }
            return;
        }
    }
    for (int k = threadIdx.x; k < 2 * CHUNKS; k += G::NTHR) {
        const int h = k >= CHUNKS ? 1 : 0;
        const int kk = k - h * CHUNKS;
        const int c = h ? tl.c[1] : tl.c[0];
        const long long o = rel0 + (h ? tl.f0[1] : tl.f0[0]) * G::FBYTES + 16LL * kk;
        unsigned char* d = dst + h * G::HALF_BYTES + 16 * kk;
        stage_chunk16(a, c, o, d, aligned16);
    }
    __threadfence_block();           // plain stores of edge chunks are ordered before the arrival
    mbar_arrive_after_cp_async(full_bar);
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
// a.seg_raw (first pass of a two-pass filter), 3 = packed PCM of a.pcm_bytes (3 or 2) bytes per sample.
// INIT: sums start from the planar plane a.seg_init (second pass).
template <int OUT, bool DITHER, bool INIT, int KF = 4>
struct RingEmitter {
    const DecimateArgs& a;
    RingDest d, nx;              // the frames being finished / the frames the lane starts in the next period
    int h;
    unsigned out_low_bits4;                // a.out & 3
    bool out_aligned16, out_aligned8;      // of a.out: the 16- and 8-byte stores below need them (a caller's device pointer may be a slice)
    int vi[KF];
    double vd[KF];
    double iv[KF];
    double dv[KF];
    double2 du, dw;              // dither terms as loaded (pair tiles), traded with the partner lane in settle()
    bool wide;
    static constexpr bool LATE = DITHER;

    __device__ __forceinline__ RingEmitter(const DecimateArgs& a_, const RingDest& d_, const RingDest& nx_, int h_, unsigned out_low_bits)
        : a(a_), d(d_), nx(nx_), h(h_), out_low_bits4(out_low_bits & 3u), out_aligned16((out_low_bits & 15u) == 0),
          out_aligned8((out_low_bits & 7u) == 0) {}

    // the dither terms of the four frames being finished, issued well ahead of quantise()
    // Pair tiles read them the way the stores go out: a lane loads the 16-byte pairs {D(f, c0), D(f, c0 + 1)}
    // of the two frames it will store (f = fb + 2h, fb + 2h + 1) and trades the two values its partner
    // (lane ^ 16, the other channel) needs -- two coalesced 16-byte loads and two 64-bit shuffles per lane
    // instead of four 8-byte loads at a stride that touches a different cache line per lane.
    __device__ __forceinline__ void fetch() {
        if (!DITHER) return;
        // The plane comes from DRAM (it is larger than L2 for a whole file) and the loads below are issued only a
        // quarter period before their use: ask L2 for the lines of the frames this lane STARTS next period -- they are
        // emitted two periods from now -- so that the loads of that period find them there.  One hint per lane; the 32
        // hints of a warp name 8 to 16 distinct lines.
        if (nx.fb < a.nframes) {
            const double* ahead = a.dither + nx.fb * a.dither_stride + nx.c;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ahead));
        }
        wide = KF == 4 && d.pair && ((a.dither_stride | d.c0) & 1) == 0 &&
               (reinterpret_cast<unsigned long long>(a.dither) & 15ULL) == 0;    // warp-uniform
        if (wide) {
            // loads only: the trade with the partner lane waits until the values are used (settle), so that the warp does
            // not sit on the load latency here
            const long long f0 = d.fb + 2 * h, f1 = f0 + 1;
            du = __ldg(reinterpret_cast<const double2*>(a.dither + (f0 < a.nframes ? f0 : 0) * a.dither_stride + d.c0));
            dw = __ldg(reinterpret_cast<const double2*>(a.dither + (f1 < a.nframes ? f1 : 0) * a.dither_stride + d.c0));
            return;
        }
#pragma unroll
        for (int k = 0; k < KF; ++k) {
            const long long f = d.fb + k < a.nframes ? d.fb + k : 0;     // keep the load in bounds
            dv[k] = __ldg(a.dither + f * a.dither_stride + d.c);
        }
    }
    // called once, right before the first quantise of the period
    __device__ __forceinline__ void settle() {
        if (!DITHER || KF != 4 || !wide) return;
        const double r0 = __shfl_xor_sync(0xffffffffu, h ? du.x : du.y, 16);
        const double r1 = __shfl_xor_sync(0xffffffffu, h ? dw.x : dw.y, 16);
        dv[0] = h ? r0 : du.x; dv[1 % KF] = h ? r1 : dw.x;
        dv[2 % KF] = h ? du.y : r0; dv[3 % KF] = h ? dw.y : r1;
    }

    __device__ __forceinline__ void quantise(int k, double sum) {
        if (OUT == 2) { vd[k] = sum; return; }
        if (k == 0) settle();
        double x = __dmul_rn(sum, a.scale);
        if (DITHER) x = __dadd_rn(x, __dmul_rn(dv[k], a.dither_amp));
        if (OUT == 1) {
            // clip (dsd_decimator.cpp:207-212): a.clip_hi is the clip amplitude, or +infinity when the caller passed none
            vd[k] = fmin(fmax(x, -a.clip_hi), a.clip_hi);
        } else {
            // C round() -- half away from zero (dsd_decimator.cpp:213-214) -- as trunc(x + copysign(0.5 - 2^-54, x)): the
            // largest double below one half pushes every tie and nothing but the ties over the next integer.  Checked
            // against round() around every integer and every tie below 2^25, around all powers of two, and on 4e8
            // random doubles up to 2^52 (tools/check_round.c); the direct and segmented kernels keep libdevice's round(),
            // so every parity test that compares kernels compares the two.
            // The clip (dsd_decimator.cpp:207-212) comes AFTER the rounding, on the integer: round() is monotone, so
            // round(min(max(x, -c), c)) = min(max(round(x), round(-c)), round(c)) for every x and c -- two integer
            // min / max instead of two fp64 compare-and-select sequences per sample.
            const int q = __double2int_rz(__dadd_rn(x, copysign(0.49999999999999994, x)));
            vi[k] = min(max(q, a.clip_lo_i), a.clip_hi_i);
        }
    }

    // running sums the lane's next four frames start from (planar plane: 32 contiguous bytes per lane)
    __device__ __forceinline__ void prefetch() {
        if (!INIT) return;
        const double* src = a.seg_init + (long long)nx.c * a.plane_stride + nx.fb;
        if (KF == 4 && nx.fb + 3 < a.nframes) {
            const double2 u = __ldg(reinterpret_cast<const double2*>(src));
            const double2 v = __ldg(reinterpret_cast<const double2*>(src) + 1);
            iv[0] = u.x; iv[1 % KF] = u.y; iv[2 % KF] = v.x; iv[3 % KF] = v.y;
        } else if (KF == 2 && nx.fb + 1 < a.nframes) {
            const double2 u = __ldg(reinterpret_cast<const double2*>(src));
            iv[0] = u.x; iv[1 % KF] = u.y;
        } else {
#pragma unroll
            for (int k = 0; k < KF; ++k) iv[k] = nx.fb + k < a.nframes ? __ldg(src + k) : 0.0;
        }
    }
    __device__ __forceinline__ double init(int k) const { return INIT ? iv[k] : 0.0; }

    __device__ __forceinline__ void store() {
        const long long fb = d.fb, nf = a.nframes;
        if (OUT == 2) {
            double* dst = a.seg_raw + (long long)d.c * a.plane_stride + fb;
            if (KF == 4 && fb + 3 < nf) {
                reinterpret_cast<double2*>(dst)[0] = make_double2(vd[0], vd[1 % KF]);
                reinterpret_cast<double2*>(dst)[1] = make_double2(vd[2 % KF], vd[3 % KF]);
            } else if (KF == 2 && fb + 1 < nf) {
                reinterpret_cast<double2*>(dst)[0] = make_double2(vd[0], vd[1 % KF]);
            } else {
#pragma unroll
                for (int k = 0; k < KF; ++k)
                    if (fb + k < nf) dst[k] = vd[k];
            }
            return;
        }
        if (OUT == 1) {
            double* out = reinterpret_cast<double*>(a.out);
#pragma unroll
            for (int k = 0; k < KF; ++k)
                if (fb + k < nf) out[(fb + k) * a.nch + d.c] = vd[k];
            return;
        }
        if (OUT == 3) { store_packed(); return; }
        int32_t* out = reinterpret_cast<int32_t*>(a.out);
        if (KF == 4 && d.pair && (a.nch & 1) == 0) {
            // trade halves: lane h=0 ends up with frames fb, fb+1 of both channels, lane h=1 with fb+2, fb+3
            const int s0 = __shfl_xor_sync(0xffffffffu, h ? vi[0] : vi[2 % KF], 16);
            const int s1 = __shfl_xor_sync(0xffffffffu, h ? vi[1 % KF] : vi[3 % KF], 16);
            const long long f = fb + 2 * h;
            const int2 lo2 = h ? make_int2(s0, vi[2 % KF]) : make_int2(vi[0], s0);
            const int2 hi2 = h ? make_int2(s1, vi[3 % KF]) : make_int2(vi[1 % KF], s1);
            if (a.nch == 2 && out_aligned16 && f + 1 < nf) {
                *reinterpret_cast<int4*>(out + f * 2) = make_int4(lo2.x, lo2.y, hi2.x, hi2.y);
            } else if (out_aligned8) {
                if (f < nf) *reinterpret_cast<int2*>(out + f * a.nch + d.c0) = lo2;
                if (f + 1 < nf) *reinterpret_cast<int2*>(out + (f + 1) * a.nch + d.c0) = hi2;
            } else {                // output base only 4-byte aligned: scalar stores
                if (f < nf) { out[f * a.nch + d.c0] = lo2.x; out[f * a.nch + d.c0 + 1] = lo2.y; }
                if (f + 1 < nf) { out[(f + 1) * a.nch + d.c0] = hi2.x; out[(f + 1) * a.nch + d.c0 + 1] = hi2.y; }
            }
        } else if (KF == 4 && a.nch == 1 && out_aligned16 && fb + 3 < nf) {
            *reinterpret_cast<int4*>(out + fb) = make_int4(vi[0], vi[1 % KF], vi[2 % KF], vi[3 % KF]);
        } else {
#pragma unroll
            for (int k = 0; k < KF; ++k)
                if (fb + k < nf) out[(fb + k) * a.nch + d.c] = vi[k];
        }
    }

    // Packed PCM (3 or 2 bytes per sample).  Stereo: after the same trade of halves a lane owns two whole frames =
    // 12 (8) contiguous bytes, 4-byte aligned because fb is a multiple of 4; a warp instruction covers 768 (512)
    // contiguous bytes.  Every other shape goes out byte by byte (the output side is 12x looser than the table loads).
    __device__ __forceinline__ void store_packed() {
        const long long fb = d.fb, nf = a.nframes;
        unsigned char* out = reinterpret_cast<unsigned char*>(a.out);
        const int B = a.pcm_bytes;
        const bool aligned4 = out_aligned8 || (out_low_bits4 == 0);
        if (KF == 4 && d.pair && (a.nch & 1) == 0) {
            const int s0 = __shfl_xor_sync(0xffffffffu, h ? vi[0] : vi[2 % KF], 16);
            const int s1 = __shfl_xor_sync(0xffffffffu, h ? vi[1 % KF] : vi[3 % KF], 16);
            const long long f = fb + 2 * h;
            const int2 lo2 = h ? make_int2(s0, vi[2 % KF]) : make_int2(vi[0], s0);
            const int2 hi2 = h ? make_int2(s1, vi[3 % KF]) : make_int2(vi[1 % KF], s1);
            if (a.nch == 2 && aligned4 && f + 1 < nf) {
                if (B == 3) {
                    const uint32_t q0 = (uint32_t)lo2.x & 0xffffffu, q1 = (uint32_t)lo2.y & 0xffffffu;
                    const uint32_t q2 = (uint32_t)hi2.x & 0xffffffu, q3 = (uint32_t)hi2.y & 0xffffffu;
                    uint32_t* p = reinterpret_cast<uint32_t*>(out + f * 6);
                    p[0] = q0 | (q1 << 24);
                    p[1] = (q1 >> 8) | (q2 << 16);
                    p[2] = (q2 >> 16) | (q3 << 8);
                } else {
                    uint32_t* p = reinterpret_cast<uint32_t*>(out + f * 4);
                    p[0] = ((uint32_t)lo2.x & 0xffffu) | ((uint32_t)lo2.y << 16);
                    p[1] = ((uint32_t)hi2.x & 0xffffu) | ((uint32_t)hi2.y << 16);
                }
            } else {
                if (f < nf) { store_pcm(out, f * a.nch + d.c0, lo2.x, B); store_pcm(out, f * a.nch + d.c0 + 1, lo2.y, B); }
                if (f + 1 < nf) { store_pcm(out, (f + 1) * a.nch + d.c0, hi2.x, B); store_pcm(out, (f + 1) * a.nch + d.c0 + 1, hi2.y, B); }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KF; ++k)
                if (fb + k < nf) store_pcm(out, (fb + k) * a.nch + d.c, vi[k], B);
        }
    }
};

// The pair geometry (ring_core.h, "Frame pairs") finishes eight frames per lane and period: two emitters, the second four
// frames further on, driven as two staggered sequences.
template <int OUT, bool DITHER>
struct RingEmitterPair {
    RingEmitter<OUT, DITHER, false, 4> e0, e1;
    static constexpr bool LATE = DITHER;
    static __device__ __forceinline__ RingDest further(RingDest d) { d.fb += 4; return d; }
    __device__ __forceinline__ RingEmitterPair(const DecimateArgs& a, const RingDest& d, const RingDest& nx, int h, unsigned out_low_bits)
        : e0(a, d, nx, h, out_low_bits), e1(a, further(d), further(nx), h, out_low_bits) {}
    __device__ __forceinline__ void fetch(int q) { if (q) e1.fetch(); else e0.fetch(); }
    __device__ __forceinline__ void quantise(int q, int k, double sum) { if (q) e1.quantise(k, sum); else e0.quantise(k, sum); }
    __device__ __forceinline__ void store(int q) { if (q) e1.store(); else e0.store(); }
    __device__ __forceinline__ void prefetch() {}
    __device__ __forceinline__ double init(int) const { return 0.0; }
};
template <class G, int OUT, bool DITHER, bool INIT> struct RingEmitterOf { typedef RingEmitter<OUT, DITHER, INIT, G::KF> type; };
template <int NW, int KR, int NL, int OUT, bool DITHER, bool INIT>
struct RingEmitterOf<RingGeom<NW, KR, true, NL, 8>, OUT, DITHER, INIT> { typedef RingEmitterPair<OUT, DITHER> type; };

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
    uint64_t* lut_bar = empty + NBUF;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) { mbar_init(&full[b], G::NTHR); mbar_init(&empty[b], G::NTHR); }
        mbar_init(lut_bar, 1);
    }
    __syncthreads();                 // barriers initialised before anybody arrives on them
    if (tid == 0) {
        // the table image: a handful of bulk copies (TMA) by one thread, counted in bytes on a barrier of their own
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.lut);
        constexpr int PART = 32768;
        mbar_arrive_expect_tx(lut_bar, (unsigned)G::LUT_BYTES);
#pragma unroll 1
        for (int o = 0; o < G::LUT_BYTES; o += PART)
            bulk_copy_g2s(smem + o, img + o, (unsigned)(G::LUT_BYTES - o < PART ? G::LUT_BYTES - o : PART), lut_bar);
    }
    const long long ntiles = ring_tile_count<G>(a.nframes, a.nch);
    const bool aligned16 = src_aligned16(a);
    const unsigned out_low_bits = (unsigned)(reinterpret_cast<unsigned long long>(a.out) & 15ULL);
    int a_tile;
    long long rel0;                  // source offset of byte 0 of the half buffer of frame 0 (ring_stage_tile)
    {
        long long t0;
        ring_half_origin<G>(a.p0, a.in_first, 0, t0, a_tile);   // a_tile is the same for every half of the launch
        rel0 = t0 - a.in_first;
    }
    const int l = tid & 31, w = tid >> 5, h = l >> 4;
    RingLane<G> lane;
    const uint32_t w0 = (uint32_t)(G::LUT_BYTES + h * G::HALF_BYTES + ring_lane_init<G>(lane, tid, a_tile, a.unit));
    uint32_t loff = 0;

    // Tiles are staged LEAD = NBUF - 2 tiles ahead (one, or two where a period is too short to cover a load from DRAM):
    // tq[0] is the current tile, tq[d] the tile d further on in this CTA's sequence.
    constexpr int LEAD = NBUF - 2;
    long long tile = blockIdx.x;     // grid <= ntiles
    RingTile tq[LEAD + 1];
    cp_async_wait<0>();              // the table image of this thread (the tile's barrier may be a bulk-copy one)
#pragma unroll
    for (int d = 0; d < LEAD; ++d) {
        const long long t = tile + (long long)d * gridDim.x;
        tq[d] = ring_tile_at<G>(t < ntiles ? t : tile, a.nframes, a.nch);
        if (t < ntiles) ring_stage_tile<G>(a, tq[d], smem + G::LUT_BYTES + d * G::TILE_BYTES, aligned16, &full[d], rel0);
    }
    tq[LEAD] = tq[0];
    const RingTile& tl = tq[0];

    RingDest prev;                   // nothing to emit in the very first period: every store is off
    prev.fb = a.nframes; prev.c = 0; prev.c0 = 0; prev.pair = false;
    if constexpr (INIT && !G::PAIR) {   // second pass: the sums of the first tile's frames start from the plane
        const int c = h ? tl.c[1] : tl.c[0];
        const long long fb = (h ? tl.f0[1] : tl.f0[0]) + (long long)G::FPL * (16 * w + (l & 15));
#pragma unroll
        for (int k = 0; k < G::KF; ++k)
            lane.n[k] = fb + k < a.nframes ? (typename RingLane<G>::acc_t)a.seg_init[(long long)c * a.plane_stride + fb + k] : 0;
    }
    typedef typename RingEmitterOf<G, OUT, DITHER, INIT>::type Emitter;
    constexpr long long FPL = G::FPL;   // output frames per lane and period
    mbar_wait(lut_bar, 0u);          // the table image has landed
    int b = 0;                       // buffer of the current tile (the i-th of this CTA): i % NBUF, used i / NBUF times before
    for (int i = 0; tile < ntiles; ++i) {
        const long long next = tile + gridDim.x, ahead = tile + (long long)LEAD * gridDim.x;
        const int bn = b + 1 == NBUF ? 0 : b + 1;
        RingDest after;              // first frames of the next tile (INIT: where the next period's sums start)
        after.fb = a.nframes; after.c = 0; after.c0 = 0; after.pair = false;
        if (ahead < ntiles) {
            // the (i + LEAD)-th tile goes to the buffer that held the (i - 2)-th
            const int bs = b + LEAD >= NBUF ? b + LEAD - NBUF : b + LEAD;
            if (i + LEAD >= NBUF) mbar_wait(&empty[bs], (unsigned)(((i + LEAD) / NBUF - 1) & 1));
            tq[LEAD] = ring_tile_at<G>(ahead, a.nframes, a.nch);
            ring_stage_tile<G>(a, tq[LEAD], smem + G::LUT_BYTES + bs * G::TILE_BYTES, aligned16, &full[bs], rel0);
        }
        if (next < ntiles && (INIT || DITHER)) {
            const RingTile& tn = tq[1];
            after.c = h ? tn.c[1] : tn.c[0];
            after.fb = (h ? tn.f0[1] : tn.f0[0]) + FPL * (16 * w + (l & 15));
        }
        mbar_wait(&full[b], (unsigned)((i / NBUF) & 1));
        RingDest cur;
        cur.c = h ? tl.c[1] : tl.c[0];
        cur.c0 = tl.c[0];
        cur.pair = tl.c[1] != tl.c[0];
        cur.fb = (h ? tl.f0[1] : tl.f0[0]) + FPL * (16 * w + (l & 15));
#pragma unroll 1
        for (int n = 0; n < G::KR; ++n) {
            const uint32_t wnew = w0 + (uint32_t)(b * G::TILE_BYTES + n * G::HOP);
            ring_head<G>(lane, smem, wnew, loff, tid, n, RingNoTrace());
            if (n == 0 && i > 0) mbar_arrive(&empty[b == 0 ? NBUF - 1 : b - 1]);   // done with the previous tile's bytes
            RingDest nx = cur;           // the frames this lane starts in the next period
            nx.fb += FPL * G::BLK_PER_ROUND;
            if (n + 1 == G::KR) nx = after;
            Emitter emit(a, prev, nx, h, out_low_bits);
            ring_tail<G>(lane, smem, wnew, loff, tid, n, emit, RingNoTrace());
            prev = cur;
            cur.fb += FPL * G::BLK_PER_ROUND;
        }
        tile = next;
        b = bn;
#pragma unroll
        for (int d = 0; d < LEAD; ++d) tq[d] = tq[d + 1];
    }
    // the one drain of this CTA: 16 more steps finish the last frames (ROTATE: they are finished, only the emitter is left)
    if constexpr (!G::ROTATE) ring_head<G>(lane, smem, lane.wlast, loff, tid, 0, RingNoTrace());
    Emitter emit(a, prev, prev, h, out_low_bits);
    if constexpr (G::PAIR) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            emit.fetch(q);
#pragma unroll
            for (int k = 0; k < 4; ++k) emit.quantise(q, k, lane.value(4 * q + k));
            emit.store(q);
        }
    } else {
        emit.fetch();
#pragma unroll
        for (int k = 0; k < G::KF; ++k) emit.quantise(k, lane.value(k));
        emit.store();
    }
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
    // a warp takes UN * 32 consecutive groups per pass (UN x 128 contiguous source bytes per channel, UN x 512 contiguous
    // output bytes), group u * 32 + lane in its u-th set of loads / stores; the warps of the grid take consecutive such
    // tiles (one contiguous share per warp instead: 5.15 against 5.38 TB/s)
    constexpr long long stride = 32;
    const long long warps = (long long)gridDim.x * blockDim.x / 32;
    const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int32_t* __restrict__ out = reinterpret_cast<int32_t*>(a.out);
    for (long long t0 = wg * (UN * 32) + (threadIdx.x & 31); t0 - (threadIdx.x & 31) < groups; t0 += warps * (UN * 32)) {
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

}  // namespace
