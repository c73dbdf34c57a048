// TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the C ABI of include/dsdgpu.h.
//
// Linked with dsf2flac_b200/host/dsd_decimator.cpp into tests/cpp/_build/libdsdhost_cpu.so so
// that the HOST logic of the drop-in DsdDecimator (logical position, reader draining, block
// cache, dither-stream bookkeeping, loop termination) can be tested where there is no GPU.
// It is never built into, linked with or loaded by the product (dsf2flac_b200/*.so).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "dsdgpu.h"

struct dsdgpu_ctx {
    int nch, nlut, nstep;
    std::vector<double> lut;
    unsigned long long launches = 0;
    long long chunk_frames = 1 << 20;
    long long dither_stride = 0;
    long long source_block = 0;        // include/dsdgpu.h "source_block_bytes"
    char err[128] = "";
    // dsdgpu_submit is DEFERRED here: the arguments are kept and the work is done in dsdgpu_wait, so a host
    // that touches its buffers between submit and wait (which the real, asynchronous library forbids) fails
    struct Pending {
        bool busy = false;
        const uint8_t* in; size_t in_stride; int64_t in_first, in_count; uint8_t fill; int64_t p0, nframes;
        double scale, dither_amp, clip; uint64_t dither_first_sample; int out_kind; void* out;
    } pending[DSDGPU_SLOTS];
};

namespace {
// glibc rand() draws by position: srand(1), skip, draw (tests are small, so the skip is cheap)
std::mutex g_rand_lock;      // libc's generator is process-global; dsdgpu_run_batch drives this stand-in from several threads
void dither_terms(uint64_t first_sample, int64_t count, std::vector<double>& out) {
    std::lock_guard<std::mutex> guard(g_rand_lock);
    out.resize((size_t)count);
    srand(1);
    for (uint64_t k = 0; k < 2 * first_sample; ++k) (void)rand();
    for (int64_t i = 0; i < count; ++i) {
        const double r1 = (double)rand() / (double)RAND_MAX;
        const double r2 = (double)rand() / (double)RAND_MAX;
        out[(size_t)i] = r1 - r2;
    }
}
}  // namespace

extern "C" {

int dsdgpu_create(dsdgpu_ctx** ctx, int, int nch, int nlut, int nstep, const double* lut, int lut_precision) {
    dsdgpu_ctx* c = new dsdgpu_ctx();
    c->nch = nch; c->nlut = nlut; c->nstep = nstep;
    c->lut.assign(lut, lut + (size_t)nlut * 256);
    if (lut_precision == 32) {   // 32-bit tables: multiples of 2^-30 (include/dsdgpu.h; the reference's filters need no smaller exponent)
        const double unit = std::ldexp(1.0, -30);
        for (auto& v : c->lut) v = std::nearbyint(v / unit) * unit;
    }
    *ctx = c;
    return DSDGPU_OK;
}
void dsdgpu_destroy(dsdgpu_ctx* ctx) { delete ctx; }
const char* dsdgpu_last_error(const dsdgpu_ctx* ctx) { return ctx ? ctx->err : "stand-in"; }
int dsdgpu_set_option(dsdgpu_ctx* ctx, const char* key, long long value) {
    if (!strcmp(key, "chunk_frames")) ctx->chunk_frames = value;
    if (!strcmp(key, "dither_stride")) ctx->dither_stride = value;
    if (!strcmp(key, "source_block_bytes")) ctx->source_block = value;
    return DSDGPU_OK;
}
unsigned long long dsdgpu_kernel_launches(const dsdgpu_ctx* ctx) { return ctx->launches; }
void* dsdgpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void dsdgpu_host_free(void* p) { free(p); }
int dsdgpu_host_register(void*, size_t) { return DSDGPU_OK; }
int dsdgpu_host_unregister(void*) { return DSDGPU_OK; }

int dsdgpu_decimate_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first,
                         int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes, double scale,
                         double dither_amp, double clip, uint64_t dither_first_sample, int out_kind,
                         void* out) {
    // failure injection for the "fail loudly" test: DSDGPU_STANDIN_FAIL_AFTER=n makes every call after the n-th fail
    if (const char* env = getenv("DSDGPU_STANDIN_FAIL_AFTER")) {
        if ((long long)ctx->launches >= atoll(env)) {
            snprintf(ctx->err, sizeof ctx->err, "injected failure (stand-in)");
            return DSDGPU_ERR_CUDA;
        }
    }
    std::vector<double> dith;
    const long long stride = ctx->dither_stride > 0 ? ctx->dither_stride : ctx->nch;
    if (dither_amp > 0 && nframes > 0) dither_terms(dither_first_sample, (nframes - 1) * stride + ctx->nch, dith);
    for (int64_t i = 0; i < nframes; ++i)
        for (int c = 0; c < ctx->nch; ++c) {
            const int64_t p = p0 + i * ctx->nstep;
            double sum = 0.0;
            for (int t = 0; t < ctx->nlut; ++t) {
                const int64_t o = p - t - in_first;
                const long long B = ctx->source_block;
                const size_t at = B == 0 ? (size_t)c * in_stride + (size_t)o
                                : B == 1 ? (size_t)o * ctx->nch + c
                                         : (size_t)(((o / B) * ctx->nch + c) * B + o % B);
                const unsigned b = (o >= 0 && o < in_count) ? in[at] : fill;
                sum += ctx->lut[(size_t)t * 256 + b];
            }
            sum = sum * scale;
            if (dither_amp > 0) sum = sum + dith[(size_t)(i * stride + c)] * dither_amp;
            if (clip > 0) { if (sum > clip) sum = clip; else if (sum < -clip) sum = -clip; }
            const int64_t m = i * ctx->nch + c;
            if (out_kind == DSDGPU_OUT_FLOAT64) { static_cast<double*>(out)[m] = sum; continue; }
            const int32_t q = (int32_t)round(sum);
            if (out_kind == DSDGPU_OUT_INT32) static_cast<int32_t*>(out)[m] = q;
            else if (out_kind == DSDGPU_OUT_PCM16) static_cast<int16_t*>(out)[m] = (int16_t)q;
            else { uint8_t* p = static_cast<uint8_t*>(out) + 3 * m; p[0] = (uint8_t)q; p[1] = (uint8_t)(q >> 8); p[2] = (uint8_t)(q >> 16); }
        }
    ctx->launches += 1;
    return DSDGPU_OK;
}

int dsdgpu_dop_pack_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first, int64_t in_count,
                         uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits, int32_t* out) {
    auto rev = [](unsigned b) { unsigned r = 0; for (int k = 0; k < 8; ++k) r |= ((b >> k) & 1u) << (7 - k); return r; };
    auto byte_at = [&](int c, int64_t j) -> unsigned {
        const int64_t o = j - in_first;
        if (o < 0 || o >= in_count) return fill;
        const long long B = ctx->source_block;
        const size_t at = B == 0 ? (size_t)c * in_stride + (size_t)o : B == 1 ? (size_t)o * ctx->nch + c
                                 : (size_t)(((o / B) * ctx->nch + c) * B + o % B);
        return in[at];
    };
    for (int64_t i = 0; i < nframes; ++i) {
        const int64_t pm = pm0 + 2 * i + 1;
        for (int c = 0; c < ctx->nch; ++c) {
            unsigned b1 = byte_at(c, pm - 1), b2 = byte_at(c, pm);
            if (reverse_bits) { b1 = rev(b1); b2 = rev(b2); }
            out[i * ctx->nch + c] = (((8 * pm - 8) % 32 != 0) ? (int32_t)0xFFFA0000 : 0x00050000) + (int32_t)(b1 << 8) + (int32_t)b2;
        }
    }
    ctx->launches += 1;
    return DSDGPU_OK;
}
int dsdgpu_dop_pack_device(dsdgpu_ctx* ctx, const uint8_t*, size_t, int64_t, int64_t, uint8_t, int64_t, int64_t, int, int32_t*,
                           void*) {
    snprintf(ctx->err, sizeof ctx->err, "stand-in has no device");
    return DSDGPU_ERR_NODEVICE;
}
int dsdgpu_decimate_device(dsdgpu_ctx* ctx, const uint8_t*, size_t, int64_t, int64_t, uint8_t, int64_t,
                           int64_t, double, double, double, uint64_t, int, void*, void*) {
    snprintf(ctx->err, sizeof ctx->err, "stand-in has no device");
    return DSDGPU_ERR_NODEVICE;
}
int dsdgpu_submit(dsdgpu_ctx* ctx, int slot, const uint8_t* in, size_t in_stride, int64_t in_first, int64_t in_count,
                  uint8_t fill, int64_t p0, int64_t nframes, double scale, double dither_amp, double clip,
                  uint64_t dither_first_sample, int out_kind, void* out) {
    if (slot < 0 || slot >= DSDGPU_SLOTS) return DSDGPU_ERR_ARG;
    dsdgpu_ctx::Pending& p = ctx->pending[slot];
    if (p.busy) { snprintf(ctx->err, sizeof ctx->err, "slot %d is still in flight", slot); return DSDGPU_ERR_STATE; }
    p = {true, in, in_stride, in_first, in_count, fill, p0, nframes, scale, dither_amp, clip, dither_first_sample, out_kind, out};
    return DSDGPU_OK;
}
int dsdgpu_wait(dsdgpu_ctx* ctx, int slot) {
    if (slot < 0 || slot >= DSDGPU_SLOTS) return DSDGPU_ERR_ARG;
    dsdgpu_ctx::Pending& p = ctx->pending[slot];
    if (!p.busy) { snprintf(ctx->err, sizeof ctx->err, "slot %d has nothing in flight", slot); return DSDGPU_ERR_STATE; }
    p.busy = false;
    return dsdgpu_decimate_host(ctx, p.in, p.in_stride, p.in_first, p.in_count, p.fill, p.p0, p.nframes, p.scale, p.dither_amp,
                                p.clip, p.dither_first_sample, p.out_kind, p.out);
}
int dsdgpu_dither_fill_device(dsdgpu_ctx*, uint64_t, int64_t, double*, void*) { return DSDGPU_ERR_NODEVICE; }

}  // extern "C"
