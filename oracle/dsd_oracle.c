/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the DSD->PCM decimation hot path.
 *
 * A plain-C restatement of the reference algorithm, written from SURVEY.md Appendix A and
 * checked line by line against the reference sources cited below.  It is what the GPU path is
 * compared with in tests/, in __graft_entry__.smoke() and in bench.py's cpu_baseline leg; it
 * is never imported, linked or executed by the product path (dsf2flac_b200/csrc, host/).
 *
 * Parity pinning: the reference ships no tests, fixtures or golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference ITSELF: the
 * unmodified dsd_decimator.cpp / filters.cpp / dsd_sample_reader.cpp compiled into
 * oracle/_ref/libdsdref.so (oracle/Makefile, oracle/ref_wrap.cpp).  tests/test_oracle.py
 * requires bit-equality between the two on every function below (tables, positions, int32 and
 * fp64 outputs, dither draws) and tests/golden/ holds vectors generated from oracle/_ref by
 * tests/golden/make_golden.py, so the pin also holds where /root/reference is absent.
 *
 * Build: oracle/Makefile (gcc -O3 -ffp-contract=off; no -ffast-math, no -march=native: the
 * reference is built with plain -O3, src/Makefile.am:2, and x86-64 baseline has no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- filter impulse responses (numeric spec of filters.cpp:52-115) ----------------------- */
#define DSD_FILTER_BEGIN(tag, ratio, taps, tzero) \
    static const int tag##_taps = taps, tag##_tzero = tzero; \
    static const uint64_t tag##_bits[taps] = {
#define DSD_FILTER_END(tag) };
#include "dsd_filter_bits.inc"

typedef struct {
    int ratio, ntaps, tzero, nlut, nstep;
    const uint64_t* bits;
} fir_spec;

/* EXTENSION filters (ratios the reference rejects).  A test registers one impulse
 * response; the operator below then runs on it exactly as it runs on the reference's three.      */
static struct { int ratio, ntaps, tzero; uint64_t* bits; } g_custom = {0, 0, 0, NULL};

int dsdoracle_set_custom_filter(int ratio, int ntaps, int tzero, const double* coefs) {
    free(g_custom.bits);
    g_custom.bits = NULL; g_custom.ratio = 0;
    if (ratio <= 0 || ntaps <= 0 || !coefs) return 0;
    g_custom.bits = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)ntaps);
    memcpy(g_custom.bits, coefs, sizeof(uint64_t) * (size_t)ntaps);
    g_custom.ratio = ratio; g_custom.ntaps = ntaps; g_custom.tzero = tzero;
    return 1;
}

/* dsd_decimator.cpp:52-68: ratio = dsdFs / outFs (integer divide), filter chosen by ratio only */
static int pick_filter(uint32_t dsd_fs, uint32_t out_fs, fir_spec* f) {
    if (out_fs == 0) return 0;
    uint32_t ratio = dsd_fs / out_fs;
    if (g_custom.ratio > 0 && ratio == (uint32_t)g_custom.ratio) {
        f->ntaps = g_custom.ntaps; f->tzero = g_custom.tzero; f->bits = g_custom.bits;
    } else
    if (ratio == 8) { f->ntaps = div8_taps; f->tzero = div8_tzero; f->bits = div8_bits; }
    else if (ratio == 16) { f->ntaps = div16_taps; f->tzero = div16_tzero; f->bits = div16_bits; }
    else if (ratio == 32) { f->ntaps = div32_taps; f->tzero = div32_tzero; f->bits = div32_bits; }
    else return 0;
    f->ratio = (int)ratio;
    f->nstep = (int)(ratio / 8);          /* dsd_decimator.cpp:54 */
    f->nlut = (f->ntaps + 7) / 8;         /* dsd_decimator.cpp:121 */
    return 1;
}

static double coef(const fir_spec* f, int i) {
    double d;
    memcpy(&d, &f->bits[i], 8);
    return d;
}

/* dsd_decimator.cpp:117-151: one 256-entry table per 8 taps, +-coef summed in bit order 0..k-1;
 * coefficient (8t+bit) pairs with byte bit (7-bit) when msbIsPlayedFirst(), else with bit `bit`. */
static void build_lut(const fir_spec* f, int msb_flag, double* lut) {
    for (int t = 0; t < f->nlut; ++t) {
        int k = f->ntaps - t * 8;
        if (k > 8) k = 8;
        for (int s = 0; s < 256; ++s) {
            double acc = 0.0;
            for (int bit = 0; bit < k; ++bit) {
                int on = msb_flag ? ((s >> (7 - bit)) & 1) : ((s >> bit) & 1);
                double val = -1 + 2 * (double)on;
                acc += val * coef(f, t * 8 + bit);
            }
            lut[t * 256 + s] = acc;
        }
    }
}

/* ---- glibc rand(), TYPE_3 additive feedback generator, as the reference draws dither from --
 * dsd_decimator.cpp:203-204 calls rand() twice per output sample and never srand(), so the
 * stream is that of seed 1.  r[i] = 16807*r[i-1] mod (2^31-1) for i=1..30 (Schrage form),
 * r[31..33] = r[0..2], r[i] = r[i-31] + r[i-3] mod 2^32; draw k is r[344+k] >> 1.            */
typedef struct {
    uint32_t ring[34];
    int pos;          /* index (mod 34) of the next word to produce */
} glibc_rand;

static void glibc_rand_seed(glibc_rand* g, uint32_t seed) {
    int32_t r[34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 0; i < 34; ++i) g->ring[i] = (uint32_t)r[i];
    g->pos = 0;                     /* next produced index is 34, stored at slot 34 % 34 = 0 */
    for (int i = 34; i < 344; ++i) {
        uint32_t v = g->ring[(g->pos + 34 - 31) % 34] + g->ring[(g->pos + 34 - 3) % 34];
        g->ring[g->pos] = v;
        g->pos = (g->pos + 1) % 34;
    }
}

static int32_t glibc_rand_next(glibc_rand* g) {
    uint32_t v = g->ring[(g->pos + 34 - 31) % 34] + g->ring[(g->pos + 34 - 3) % 34];
    g->ring[g->pos] = v;
    g->pos = (g->pos + 1) % 34;
    return (int32_t)(v >> 1);
}

#define GLIBC_RAND_MAX 2147483647

/* ---- the reader model (dsd_sample_reader.h:100-106,122; dsf_file_reader.cpp:83-103) -------
 * Byte j of channel c as the decimator sees it: the ring is pre-filled with the idle byte
 * (j < 0); step number j pushes data only if samplesAvailable() held BEFORE the marker moved,
 * i.e. (j-1)*8 < length, and the source actually has a byte there.                            */
typedef struct {
    const uint8_t* planar;
    int nch;
    int64_t nbytes, length;
    uint8_t idle;
} byte_source;

static inline uint8_t source_byte(const byte_source* s, int c, int64_t j) {
    if (j < 0 || j >= s->nbytes) return s->idle;
    if ((j - 1) * 8 >= s->length) return s->idle;
    return s->planar[(int64_t)c * s->nbytes + j];
}

/* ---- the operator (dsd_decimator.cpp:173-222) -------------------------------------------- */
typedef struct {
    fir_spec fir;
    double* lut;
    byte_source src;
    int64_t p;                      /* reader posMarker: index of the newest byte pushed */
    glibc_rand rng;
} oracle_dec;

static int dec_open(oracle_dec* d, const uint8_t* planar, int nch, int64_t nbytes, int64_t length,
                    uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs) {
    if (!pick_filter(dsd_fs, out_fs, &d->fir)) return 0;
    d->lut = (double*)malloc(sizeof(double) * 256 * (size_t)d->fir.nlut);
    build_lut(&d->fir, msb_flag, d->lut);
    d->src.planar = planar; d->src.nch = nch; d->src.nbytes = nbytes; d->src.length = length;
    d->src.idle = idle;
    d->p = -1;                      /* dsf_file_reader.cpp:117 */
    glibc_rand_seed(&d->rng, 1);
    return 1;
}

static void dec_close(oracle_dec* d) { free(d->lut); d->lut = NULL; }

/* dsd_decimator.cpp:90-93 with dsd_sample_reader.h:89 */
static double dec_position(const oracle_dec* d) {
    int64_t dsd_pos = d->p * 8;
    return (double)(dsd_pos - (int64_t)d->fir.tzero) / (uint32_t)d->fir.ratio;
}
static int64_t dec_length(const oracle_dec* d) { return d->src.length / (uint32_t)d->fir.ratio; }
static double dec_first_valid(const oracle_dec* d) {
    return (double)(uint32_t)d->fir.nlut / (uint32_t)d->fir.nstep
         - (double)(uint32_t)d->fir.tzero / (uint32_t)d->fir.ratio;
}
static double dec_last_valid(const oracle_dec* d) {
    return (double)dec_length(d) - (double)(uint32_t)d->fir.tzero / (uint32_t)d->fir.ratio;
}

/* One getSamples call of `nframes` frames.  Exactly one of out_i32 / out_f64 is written
 * (int types round half away from zero via round(), float types are a plain cast).          */
static void dec_get_samples(oracle_dec* d, int64_t nframes, double scale, double dither,
                            double clip_amp, int32_t* out_i32, double* out_f64) {
    const int nch = d->src.nch, nlut = d->fir.nlut;
    const int clip = clip_amp > 0;
    for (int64_t i = 0; i < nframes; ++i) {
        for (int c = 0; c < nch; ++c) {
            double sum = 0.0;
            for (int t = 0; t < nlut; ++t)                    /* sequential, t ascending */
                sum += d->lut[t * 256 + source_byte(&d->src, c, d->p - t)];
            sum = sum * scale;
            if (dither > 0) {
                double rand1 = ((double)glibc_rand_next(&d->rng)) / ((double)GLIBC_RAND_MAX);
                double rand2 = ((double)glibc_rand_next(&d->rng)) / ((double)GLIBC_RAND_MAX);
                sum = sum + (rand1 - rand2) * dither;
            }
            if (clip) {
                if (sum > clip_amp) sum = clip_amp;
                else if (sum < -clip_amp) sum = -clip_amp;
            }
            if (out_i32) out_i32[i * nch + c] = (int32_t)round(sum);
            else out_f64[i * nch + c] = sum;
        }
        d->p += d->fir.nstep;                                 /* nStep x reader->step() */
    }
}

/* ---- C ABI (mirrors oracle/ref_wrap.cpp one to one) -------------------------------------- */

int dsdoracle_info(uint32_t dsd_fs, uint32_t out_fs, int64_t length_samples, int32_t* ratio,
                   int32_t* nlut, int32_t* nstep, int32_t* tzero, double* first_valid,
                   double* last_valid, int64_t* length_pcm) {
    oracle_dec d;
    if (!dec_open(&d, NULL, 1, 0, length_samples, 0x69, 1, dsd_fs, out_fs)) return 0;
    *ratio = d.fir.ratio; *nlut = d.fir.nlut; *nstep = d.fir.nstep; *tzero = d.fir.tzero;
    *first_valid = dec_first_valid(&d);
    *last_valid = dec_last_valid(&d);
    *length_pcm = dec_length(&d);
    dec_close(&d);
    return 1;
}

int dsdoracle_lut(uint32_t dsd_fs, uint32_t out_fs, int msb_flag, double* lut, int cap_tables) {
    fir_spec f;
    if (!pick_filter(dsd_fs, out_fs, &f)) return 0;
    double* full = (double*)malloc(sizeof(double) * 256 * (size_t)f.nlut);
    build_lut(&f, msb_flag, full);
    int n = f.nlut < cap_tables ? f.nlut : cap_tables;
    memcpy(lut, full, sizeof(double) * 256 * (size_t)n);
    free(full);
    return f.nlut;
}

/* rand_skip: dither draws already consumed before this call (lets a test start mid-stream). */
int dsdoracle_run_raw(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                      uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs,
                      int64_t pre_steps, int64_t nframes, int64_t rand_skip, double scale,
                      double dither, double clip, int32_t* out_i32, double* out_f64) {
    oracle_dec d;
    if (!dec_open(&d, planar, nch, nbytes, length_samples, idle, msb_flag, dsd_fs, out_fs)) return 1;
    if (out_i32) {
        d.p = -1 + pre_steps;
        glibc_rand_seed(&d.rng, 1);
        for (int64_t k = 0; k < rand_skip; ++k) (void)glibc_rand_next(&d.rng);
        dec_get_samples(&d, nframes, scale, dither, clip, out_i32, NULL);
    }
    if (out_f64) {
        d.p = -1 + pre_steps;
        glibc_rand_seed(&d.rng, 1);
        for (int64_t k = 0; k < rand_skip; ++k) (void)glibc_rand_next(&d.rng);
        dec_get_samples(&d, nframes, scale, dither, clip, NULL, out_f64);
    }
    dec_close(&d);
    return 0;
}

/* main.cpp:169-191 with the --onefile bounds of main.cpp:249-258. */
int64_t dsdoracle_run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                          uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs,
                          double scale, double dither, double clip, int32_t* out,
                          int64_t out_cap_frames) {
    oracle_dec d;
    if (!dec_open(&d, planar, nch, nbytes, length_samples, idle, msb_flag, dsd_fs, out_fs)) return -1;
    double start_pos = dec_first_valid(&d), end_pos = dec_last_valid(&d);
    if (start_pos > (double)(dec_length(&d) - 1)) start_pos = (double)(dec_length(&d) - 1);
    if (end_pos > (double)dec_length(&d)) end_pos = (double)dec_length(&d);
    const int block = 1024;
    int64_t done = 0;
    while (dec_position(&d) < start_pos) d.p += 1;            /* creep: step() only */
    while (dec_position(&d) <= end_pos - block) {
        if (done + block > out_cap_frames) { dec_close(&d); return -2; }
        dec_get_samples(&d, block, scale, dither, clip, out + done * nch, NULL);
        done += block;
    }
    while (dec_position(&d) <= end_pos) {
        if (done + 1 > out_cap_frames) { dec_close(&d); return -2; }
        dec_get_samples(&d, 1, scale, dither, clip, out + done * nch, NULL);
        done += 1;
    }
    dec_close(&d);
    return done;
}

/* ---- DoP packing (dop_packer.cpp:100-141) -------------------------------------------------
 * pack_buffer, per frame: reader->step(); then for every channel the marker is 0x00050000 if
 * (getPosition() - 8) % 32 == 0 else 0xFFFA0000 (C remainder), plus ring[1] << 8 (the byte before the
 * newest) plus ring[0] (the newest), both bit-reversed when msbIsPlayedFirst(); reader->step() again.
 * With pm = the reader's marker after the first step, getPosition() = 8*pm.  Starting right after
 * rewind() (pre_steps = 0) every marker comes out 0xFFFA0000; main.cpp's creep loop (:362-364) takes
 * one step() first, after which they alternate 0x05 / 0xFA as DoP requires.                       */
static uint8_t reverse_bits(uint8_t b) {
    b = (uint8_t)((b & 0xF0) >> 4 | (b & 0x0F) << 4);
    b = (uint8_t)((b & 0xCC) >> 2 | (b & 0x33) << 2);
    b = (uint8_t)((b & 0xAA) >> 1 | (b & 0x55) << 1);
    return b;
}

int dsdoracle_dop_run_raw(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                          uint8_t idle, int msb_flag, int64_t pre_steps, int64_t nframes, int32_t* out) {
    byte_source src = {planar, nch, nbytes, length_samples, idle};
    int64_t pm = -1 + pre_steps;
    for (int64_t i = 0; i < nframes; ++i) {
        pm += 1;
        for (int c = 0; c < nch; ++c) {
            int32_t v = ((8 * pm - 8) % 32 != 0) ? (int32_t)0xFFFA0000 : 0x00050000;
            uint8_t b1 = source_byte(&src, c, pm - 1), b2 = source_byte(&src, c, pm);
            if (msb_flag) { b1 = reverse_bits(b1); b2 = reverse_bits(b2); }
            out[i * nch + c] = v + ((int32_t)b1 << 8) + (int32_t)b2;
        }
        pm += 1;
    }
    return 0;
}

/* dop_track_helper's loops for the --onefile case (main.cpp:300-306, 362-383): creep to position 0,
 * blocks of 1024 frames while position <= end - 1024*16, then single frames while position <= end. */
int64_t dsdoracle_dop_run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                              uint8_t idle, int msb_flag, int32_t* out, int64_t out_cap_frames) {
    int64_t start = 0, end = length_samples;
    if (start > length_samples - 1) start = length_samples - 1;
    int64_t pm = -1, done = 0;
    while (8 * pm < start) pm += 1;
    while (8 * pm <= end - 1024 * 16) {
        if (done + 1024 > out_cap_frames) return -2;
        dsdoracle_dop_run_raw(planar, nch, nbytes, length_samples, idle, msb_flag, pm + 1, 1024, out + done * nch);
        done += 1024; pm += 2048;
    }
    while (8 * pm <= end) {
        if (done + 1 > out_cap_frames) return -2;
        dsdoracle_dop_run_raw(planar, nch, nbytes, length_samples, idle, msb_flag, pm + 1, 1, out + done * nch);
        done += 1; pm += 2;
    }
    return done;
}

void dsdoracle_rand(int32_t* out, int64_t n, int64_t skip) {
    glibc_rand g;
    glibc_rand_seed(&g, 1);
    for (int64_t k = 0; k < skip; ++k) (void)glibc_rand_next(&g);
    for (int64_t i = 0; i < n; ++i) out[i] = glibc_rand_next(&g);
}

/* The GPU dither generator's draw / RAND_MAX: fma(x, 2^-31 + 2^-62, x), x = a * 2^-31, against the quotient the
 * reference computes (dsd_decimator.cpp:203-204) for a = first, first + stride, ... (count values).  Returns the
 * number of mismatches (tools/check_draw_quotient.c runs all 2^31 values: 0). */
long long dsdoracle_check_draw_quotient(long long first, long long stride, long long count) {
    const double c = ldexp(1.0, -31) + ldexp(1.0, -62);
    long long bad = 0;
    for (long long k = 0; k < count; ++k) {
        const long long a = first + k * stride;
        if (a < 0 || a > 2147483647LL) continue;
        const double x = ldexp((double)a, -31);
        if (fma(x, c, x) != (double)a / 2147483647.0) ++bad;
    }
    return bad;
}
