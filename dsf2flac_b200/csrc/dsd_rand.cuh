// dsd_rand.cuh -- the reference's dither stream by position (included once, by dsdgpu.cu): glibc's unseeded
// rand() (TYPE_3 additive feedback generator) as the reference draws it, two values per output sample
// (reference src/dsd_decimator.cpp:203-204), regenerated for any sample range on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

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

// draw / RAND_MAX, correctly rounded, from the 31-bit draw a:  x = a * 2^-31 is exact (the draw becomes the low
// mantissa bits of 2^21 + x, one subtraction away), and a / (2^31 - 1) = x * (1 + 2^-31 + 2^-62 + ...) rounds to
// fma(x, 2^-31 + 2^-62, x) for EVERY one of the 2^31 possible draws (checked exhaustively against a / 2147483647.0 on
// the CPU, tools/check_draw_quotient.c; tests/test_oracle.py samples it).  Two fp64 operations per draw.
__device__ __forceinline__ double rand_draw_quotient(uint32_t r) {
    const double x = __dsub_rn(__hiloint2double(0x41400000, (int)(r >> 1)), 2097152.0);
    return __fma_rn(x, 4.6566128752457969e-10 /* 2^-31 + 2^-62 */, x);
}

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
            const double q = rand_draw_quotient(v);
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
