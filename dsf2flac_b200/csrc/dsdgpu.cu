// dsdgpu.cu -- the C ABI of include/dsdgpu.h: context, table images, launches, the pipelined host path.
//
// Hot path replaced: DsdDecimator::getSamplesInternal (reference src/dsd_decimator.cpp:173-222)
// -- the per-byte lookup-table FIR (:194-198) fused with scale / TPDF dither / clip / round
// (:199-216).  The kernels live in dsd_kernels.cuh (lane logic: ring_core.h, seg_core.h),
// the positional reconstruction of the reference's dither stream in dsd_rand.cuh.
// There is no CPU fallback anywhere in this library.
#include "dsdgpu.h"
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

}  // namespace

#include "dsd_kernels.cuh"
#include "dsd_rand.cuh"


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
    int dither_stride = 0;                   // 0 = nch
    long long source_block = 0;              // 0: planar rows; 1: byte-interleaved; 2^k >= 16: block-planar
    long long chunk_frames = 1LL << 20;
    double* d_lut_direct = nullptr;          // [nlut][256]
    unsigned char* d_lut_ring = nullptr;     // RingGeom::LUT_BYTES image (nlut == 72 only): fp64 or 32-bit entries
    unsigned char* d_lut_seg = nullptr;      // nseg x LUT_BYTES images, SEG_DEFAULT tables each (any geometry)
    int nseg = 0;
    bool seg_narrow = false;                 // SegLayoutNarrow (80-slot rows, 64 tables per pass): frame hops of 16 bytes and more
    double lut_unit = 1.0;                   // 32-bit tables: 2^-qbits
    std::vector<uint32_t> rand_pow16;        // host: x^(d * 16^k), d < 16, k < 12, [k][d][31]
    std::vector<uint32_t> rand_base;         // host: seed words r[3..94]
    uint32_t* d_rand_lane = nullptr;         // x^(kRandLaneValues*l), l < 32, stored [k][l]
    uint32_t* d_rand_jump = nullptr;         // x^(kRandWarpValues * 128^level * u), [level][u][k]
    // Dither plane cache ("dither_cache" option).  The reference draws from the UNSEEDED process-global rand()
    // (dsd_decimator.cpp:203-204), and dsf2flac converts one file per process: the (r1 - r2) term of output sample m
    // is a constant of the reference, the same for every file -- like the filter tables.  A context therefore keeps
    // the prefix [0, dither_len) of that plane once generated; calls whose samples lie inside the limit read it in
    // place and only ever generate what is not there yet.  Grow-only; a region is never rewritten once valid.
    long long dither_cache_limit = 1LL << 27;   // samples (8 bytes each); 0 = off: every call generates its own plane
    double* d_dither_cache = nullptr;
    long long dither_cap = 0, dither_len = 0;   // allocated / valid samples
    cudaEvent_t dither_ready = nullptr;         // recorded after the last generation; readers on other streams wait for it
    std::vector<double*> dither_retired;        // outgrown planes: possibly still read by launches in flight, freed at destroy
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

using Ring16x2 = RingGeom<16, 2>;
using Ring16x1 = RingGeom<16, 1>;               // the same kernel with one round per tile: half-size tiles for short launches
static_assert(Ring16x1::LUT_BYTES == Ring16x2::LUT_BYTES, "both read the same table image");
using Ring16x2Q = RingGeom<8, 1, true, 76, 8>;   // 32-bit tables: two frames per 8-byte entry (76 pair tables, chains 8 bytes apart)
using Ring30 = RingGeom<16, 2, false, 30, 2>;   // divide-by-16 filter: 30 tables, 2 bytes per frame
using Ring12 = RingGeom<16, 2, false, 12, 1>;   // divide-by-8 filter: 12 tables, 1 byte per frame
using Ring12L = RingGeom<16, 16, false, 12, 1>; // the same with 16 rounds per tile: what a tile costs besides its table loads (geometry,
                                                // staging, barriers: once per tile and thread) weighs on 12-step periods -- 0.58 -> 0.70 of roofline
static_assert(Ring12L::LUT_BYTES == Ring12::LUT_BYTES, "both read the same table image");
using Ring64 = RingGeom<16, 1, false, 72, 8>;   // one pass (72 of 144 tables) of the divide-by-64 extension filter
using Ring128 = RingGeom<16, 1, false, 72, 16>; // ... of the divide-by-128 filter (288 tables, four passes): two frames per lane, 16 bytes apart
using Ring256 = RingGeom<16, 1, false, 72, 32>; // ... of the divide-by-256 filter (576 tables, eight passes): one frame per lane
static_assert(Ring64::LUT_BYTES == Ring128::LUT_BYTES && Ring64::LUT_BYTES == Ring256::LUT_BYTES, "one image layout for all passes");

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

constexpr int DSDGPU_ERR_NOMEM_SOFT = -1000;   // internal: the dither plane could not grow (the caller falls back to a plane of its own)

// Make samples [0, end) of the cached dither plane valid for work enqueued on `st` afterwards.
int dither_cache_cover(dsdgpu_ctx* ctx, dsdgpu_slot& s, long long end, cudaStream_t st) {
    if (!ctx->dither_ready) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->dither_ready, cudaEventDisableTiming));
    if (ctx->dither_len > 0) CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->dither_ready, 0));   // earlier generations, on whatever stream
    if (end <= ctx->dither_len) return DSDGPU_OK;
    if (end > ctx->dither_cap) {
        long long cap = 1LL << 22;
        while (cap < end) cap <<= 1;
        if (cap > ctx->dither_cache_limit) cap = ctx->dither_cache_limit;
        double* fresh = nullptr;
        if (cudaMalloc(reinterpret_cast<void**>(&fresh), (size_t)cap * sizeof(double)) != cudaSuccess) {
            (void)cudaGetLastError();               // no room for the plane: this call and the later ones make their own
            ctx->dither_cache_limit = ctx->dither_len;
            return DSDGPU_ERR_NOMEM_SOFT;
        }
        if (ctx->dither_len > 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(fresh, ctx->d_dither_cache, (size_t)ctx->dither_len * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (ctx->d_dither_cache) ctx->dither_retired.push_back(ctx->d_dither_cache);
        ctx->d_dither_cache = fresh;
        ctx->dither_cap = cap;
    }
    // whole generator warps: the new part starts where the valid part ends
    int rc = launch_dither(ctx, s, (unsigned long long)ctx->dither_len, end - ctx->dither_len, ctx->d_dither_cache + ctx->dither_len, st);
    if (rc) return rc;
    ctx->dither_len = end;
    CUDA_TRY(ctx, cudaEventRecord(ctx->dither_ready, st));
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
    const bool f64 = a.out_kind == DSDGPU_OUT_FLOAT64, dith = a.dither != nullptr;
    if (f64) return dith ? launch_ring_variant<G, 1, true, INIT>(ctx, a, st) : launch_ring_variant<G, 1, false, INIT>(ctx, a, st);
    if (a.pcm_bytes) return dith ? launch_ring_variant<G, 3, true, INIT>(ctx, a, st) : launch_ring_variant<G, 3, false, INIT>(ctx, a, st);
    return dith ? launch_ring_variant<G, 0, true, INIT>(ctx, a, st) : launch_ring_variant<G, 0, false, INIT>(ctx, a, st);
}

// Decimation by 64 / 128 / 256 (EXTENSION filters, 144 / 288 / 576 tables): passes of the ring kernel over 72 tables each,
// frames 8 / 16 / 32 bytes apart.  Pass 0 adds tables 0..71 and leaves the running sums in a planar fp64 plane; pass s
// starts from them and adds tables 72s .. 72s+71 -- the byte of table 72s + t at frame position p is the byte of table t
// at p - 72s -- in place (a lane reads the sums of its next frames a period before it writes those of its last ones,
// and every sample belongs to one lane); the last pass finishes the samples.  The reference's sequential sum over all
// tables, cut at table boundaries, bit for bit.
template <class G>
int launch_ring_passes(dsdgpu_ctx* ctx, dsdgpu_slot& scratch, DecimateArgs a, cudaStream_t st) {
    const int npass = ctx->nlut / G::NLUT;
    a.plane_stride = (a.nframes + 3) & ~3LL;
    int rc = ensure(ctx, &scratch.d_partial, &scratch.partial_cap, (size_t)a.plane_stride * a.nch * sizeof(double));
    if (rc) return rc;
    DecimateArgs pass = a;
    pass.seg_raw = scratch.d_partial;
    pass.dither = nullptr;
    for (int s = 0; s + 1 < npass; ++s) {
        pass.lut = ctx->d_lut_ring + (size_t)s * G::LUT_BYTES;
        pass.p0 = a.p0 - (long long)s * G::NLUT;
        pass.seg_init = s > 0 ? scratch.d_partial : nullptr;
        rc = s > 0 ? launch_ring_variant<G, 2, false, true>(ctx, pass, st) : launch_ring_variant<G, 2, false, false>(ctx, pass, st);
        if (rc) return rc;
    }
    a.lut = ctx->d_lut_ring + (size_t)(npass - 1) * G::LUT_BYTES;
    a.p0 -= (long long)(npass - 1) * G::NLUT;
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
    if (out_kind != DSDGPU_OUT_INT32 && out_kind != DSDGPU_OUT_FLOAT64 && out_kind != DSDGPU_OUT_PCM24 && out_kind != DSDGPU_OUT_PCM16)
        return fail(ctx, DSDGPU_ERR_ARG, "unknown output kind %d", out_kind);
    return DSDGPU_OK;
}

// bytes per output sample; packed PCM only holds what the clip amplitude guarantees to fit
size_t out_elem_bytes(int out_kind) {
    return out_kind == DSDGPU_OUT_INT32 ? sizeof(int32_t) : out_kind == DSDGPU_OUT_FLOAT64 ? sizeof(double)
         : out_kind == DSDGPU_OUT_PCM24 ? 3 : 2;
}
int check_packed(dsdgpu_ctx* ctx, int out_kind, double clip) {
    if (out_kind == DSDGPU_OUT_PCM24 && !(clip > 0.0 && clip <= 8388607.0))
        return fail(ctx, DSDGPU_ERR_ARG, "DSDGPU_OUT_PCM24 needs 0 < clip <= 2^23-1 (every sample must fit 24 bits)");
    if (out_kind == DSDGPU_OUT_PCM16 && !(clip > 0.0 && clip <= 32767.0))
        return fail(ctx, DSDGPU_ERR_ARG, "DSDGPU_OUT_PCM16 needs 0 < clip <= 2^15-1 (every sample must fit 16 bits)");
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
    a.clip_hi = clip > 0.0 ? clip : HUGE_VAL;
    {   // round(clip) as C round() does it, saturated like the kernels' double -> int conversion
        const double rc = std::round(clip);
        a.clip_hi_i = !(clip > 0.0) || rc >= 2147483647.0 ? INT32_MAX : (int)rc;
        a.clip_lo_i = !(clip > 0.0) || rc > 2147483647.0 ? INT32_MIN : -(int)rc;
    }
    a.unit = ctx->lut_unit;
    a.out = d_out; a.out_kind = out_kind;
    a.pcm_bytes = out_kind == DSDGPU_OUT_PCM24 ? 3 : out_kind == DSDGPU_OUT_PCM16 ? 2 : 0;
    { int rc = check_packed(ctx, out_kind, clip); if (rc) return rc; }
    a.seg_first = 0; a.seg_n = ctx->nlut; a.seg_init = nullptr; a.seg_raw = nullptr; a.plane_stride = 0;
    if (dither_amp > 0.0) {
        const long long count = (nframes - 1) * a.dither_stride + ctx->nch;
        // the cached plane serves calls that start inside (or right at the end of) what is already valid -- a file converted
        // from its first sample, whole or chunk after chunk; a call far into the stream (a time shard) makes its own plane
        if (dither_first_sample <= (uint64_t)ctx->dither_len &&
            dither_first_sample + (uint64_t)count <= (uint64_t)ctx->dither_cache_limit) {
            int rc = dither_cache_cover(ctx, scratch, (long long)dither_first_sample + count, st);
            if (rc && rc != DSDGPU_ERR_NOMEM_SOFT) return rc;
            if (rc == DSDGPU_OK) a.dither = ctx->d_dither_cache + dither_first_sample;
        }
        if (a.dither == nullptr) {
            int rc = ensure(ctx, &scratch.d_dither, &scratch.dither_cap, (size_t)count * sizeof(double));
            if (rc) return rc;
            rc = launch_dither(ctx, scratch, dither_first_sample, count, scratch.d_dither, st);
            if (rc) return rc;
            a.dither = scratch.d_dither;
        }
    }
    const bool ring_ok = ctx->d_lut_ring != nullptr;
    int kernel = ctx->kernel;
    if (kernel == DSDGPU_KERNEL_AUTO)
        kernel = ring_ok ? DSDGPU_KERNEL_RING
               : seg_samples_per_thread(ctx) > 0 ? DSDGPU_KERNEL_SEGMENTED : DSDGPU_KERNEL_DIRECT;
    if (kernel == DSDGPU_KERNEL_RING) {
        if (!ring_ok) return fail(ctx, DSDGPU_ERR_ARG, "ring kernel needs the geometry of one of the reference's filters (72/4, 30/2, 12/1) or of the extension filters (144/8, 288/16, 576/32), and tables without -0.0 entries");
        a.lut = ctx->d_lut_ring;
        if (ctx->nlut == 144) return launch_ring_passes<Ring64>(ctx, scratch, a, st);
        if (ctx->nlut == 288) return launch_ring_passes<Ring128>(ctx, scratch, a, st);
        if (ctx->nlut == 576) return launch_ring_passes<Ring256>(ctx, scratch, a, st);
        if (ctx->nlut == 30) return launch_ring<Ring30>(ctx, a, st);
        if (ctx->nlut == 12)      // long tiles where there are enough of them to keep every SM busy, short ones for short launches
            return ring_tile_count<Ring12L>(a.nframes, a.nch) >= 6LL * ctx->sm_count ? launch_ring<Ring12L>(ctx, a, st) : launch_ring<Ring12>(ctx, a, st);
        // (a 1024-thread instance of the 32-bit kernel fits in 64 registers but runs 13 % slower)
        if (ctx->lut_precision == 32) return launch_ring<Ring16x2Q>(ctx, a, st);
        // Tiles of 2 x 2048 frames run 1 % faster on a long stream; a short launch (a 60 s DSD64 file, a chunk of
        // the pipelined host path) leaves the CTAs with an uneven handful of them: tiles of 2 x 1024 frames then
        // (measured: configs[0] +2 %, configs[1] -1 %)
        if (ring_tile_count<Ring16x2>(a.nframes, a.nch) < 24LL * ctx->sm_count) return launch_ring<Ring16x1>(ctx, a, st);
        return launch_ring<Ring16x2>(ctx, a, st);
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
    } else if (((nlut == 144 && nstep == 8) || (nlut == 288 && nstep == 16) || (nlut == 576 && nstep == 32)) && plain && lut_precision == 64) {
        // divide-by-64 / -128 / -256 extension filters: one image per pass of 72 tables (the same layout for all three)
        const int npass = nlut / 72;
        rimg.resize((size_t)npass * Ring64::LUT_BYTES);
        for (int sidx = 0; sidx < npass; ++sidx)
            ring_build_lut_image<Ring64>(tab.data() + (size_t)sidx * 72 * 256, rimg.data() + (size_t)sidx * Ring64::LUT_BYTES);
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
    if (ctx->d_lut_ring) cudaFree(ctx->d_lut_ring);
    if (ctx->d_lut_seg) cudaFree(ctx->d_lut_seg);
    if (ctx->d_rand_lane) cudaFree(ctx->d_rand_lane);
    if (ctx->d_rand_jump) cudaFree(ctx->d_rand_jump);
    if (ctx->d_dither_cache) cudaFree(ctx->d_dither_cache);
    for (double* p : ctx->dither_retired) cudaFree(p);
    if (ctx->dither_ready) cudaEventDestroy(ctx->dither_ready);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dsdgpu_last_error(const dsdgpu_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int dsdgpu_set_option(dsdgpu_ctx* ctx, const char* key, long long value) {
    if (!ctx || !key) return DSDGPU_ERR_ARG;
    if (!strcmp(key, "kernel")) {
        if (value != DSDGPU_KERNEL_AUTO && value != DSDGPU_KERNEL_DIRECT && value != DSDGPU_KERNEL_RING && value != DSDGPU_KERNEL_SEGMENTED)
            return fail(ctx, DSDGPU_ERR_ARG, "bad kernel %lld", value);
        ctx->kernel = (int)value;
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
    } else if (!strcmp(key, "dither_cache")) {
        if (value < 0) return fail(ctx, DSDGPU_ERR_ARG, "dither_cache must be >= 0 (samples)");
        if (value < ctx->dither_len) {              // shrinking below what is valid: drop the plane (launches in flight keep reading it)
            if (ctx->d_dither_cache) ctx->dither_retired.push_back(ctx->d_dither_cache);
            ctx->d_dither_cache = nullptr; ctx->dither_cap = 0; ctx->dither_len = 0;
        }
        ctx->dither_cache_limit = value;
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
    rc = check_packed(ctx, out_kind, clip);
    if (rc) return rc;
    const size_t elem = out_elem_bytes(out_kind);
    const size_t out_bytes = (size_t)nframes * ctx->nch * elem;
    rc = ensure(ctx, reinterpret_cast<unsigned char**>(&s.d_out), &s.out_cap, out_bytes + 16);
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
        if (blocks > 64LL * ctx->sm_count) blocks = 64LL * ctx->sm_count;   // (8 / 16 / 32 / 64 per SM: 5.29 / 5.38 / 5.43 / 5.48 TB/s on 600 s)
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
    const size_t elem = out_elem_bytes(out_kind);
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
