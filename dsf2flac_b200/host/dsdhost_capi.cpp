// dsdhost_capi.cpp -- C-callable drivers around the drop-in DsdDecimator, shaped like its one
// caller in the reference (pcm_track_helper, reference src/main.cpp:111-210), so tests/ and
// bench.py can run "what dsf2flac would run" through ctypes: an in-memory DsdSampleReader feeds
// the decimator byte by byte exactly as the DSF/DSDIFF readers do, the decimator drains it in
// blocks and runs the FIR on the GPU through include/dsdgpu.h.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "container_sample_reader.h"
#include "dsd_container.h"
#include "dsd_decimator.h"
#include "dsdgpu.h"
#include "memory_sample_reader.h"

namespace {
thread_local std::string g_error;
}

namespace {
template <typename T>
void raw_calls(DsdDecimator& dec, int nch, int64_t nframes, int64_t chunk_frames, double scale, double dither, double clip, void* out) {
    T* dst = static_cast<T*>(out);
    for (int64_t f = 0; f < nframes; f += chunk_frames) {
        const int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
        dec.getSamples(dst + f * nch, (dsf2flac_uint32)(n * nch), scale, dither, clip);
    }
}
}  // namespace

extern "C" {

const char* dsdhost_last_error() { return g_error.c_str(); }

// DsdDecimator::setExtensionFilters: decimation by 64, 128, 256 (the reference rejects these).
void dsdhost_set_extension_filters(int on) { DsdDecimator::setExtensionFilters(on != 0); }

// Geometry as the decimator reports it (reference dsd_decimator.cpp:44-72,85-101).
// Returns 1 if the rate pair is accepted, 0 if not ("Sorry, incompatible ..."); a missing GPU
// does not matter here: the getters do not depend on it.
int dsdhost_info(uint32_t dsd_fs, uint32_t out_fs, int64_t length_samples, int32_t* ratio,
                 double* first_valid, double* last_valid, int64_t* length_pcm) {
    uint8_t dummy = 0;
    MemorySampleReader rd(&dummy, 1, 0, length_samples, dsd_fs, true);
    DsdDecimator dec(&rd, out_fs);
    g_error = dec.getErrorMsg();
    if (g_error.rfind("Sorry", 0) == 0) return 0;            // the reference's rejection, not a missing GPU
    *ratio = (int32_t)dec.getDecimationRatio();
    *first_valid = dec.getFirstValidSample();
    *last_valid = dec.getLastValidSample();
    *length_pcm = dec.getLength();
    return 1;
}

// rewind, `pre_steps` x step(), then getSamples in calls of `chunk_frames` frames (0 = one call).
// Exactly one of out_i32 / out_f64 is filled ([nframes][nch]).  Returns 0, or 1 with the
// decimator's error message in dsdhost_last_error().
int dsdhost_run_raw(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                    uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int64_t pre_steps,
                    int64_t nframes, int64_t chunk_frames, double scale, double dither, double clip,
                    int lut_bits, uint32_t read_ahead, int32_t* out_i32, double* out_f64) {
    MemorySampleReader rd(planar, (uint32_t)nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return 1; }
    if (lut_bits != 64 && !dec.setLutPrecision(lut_bits)) { g_error = dec.getErrorMsg(); return 1; }
    if (read_ahead) dec.setReadAheadFrames(read_ahead);
    for (int64_t i = 0; i < pre_steps; ++i) dec.step();
    if (chunk_frames <= 0) chunk_frames = nframes;
    for (int64_t f = 0; f < nframes; f += chunk_frames) {
        const int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
        if (out_i32) dec.getSamples(out_i32 + f * nch, (dsf2flac_uint32)(n * nch), scale, dither, clip);
        else dec.getSamples(out_f64 + f * nch, (dsf2flac_uint32)(n * nch), scale, dither, clip);
    }
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return 1; }
    return 0;
}

// The same through any of the five getSamples<T> specialisations (reference dsd_decimator.cpp:153-172).
// type: 0 int16, 1 int32, 2 int64, 3 float32, 4 float64; out is [nframes][nch] of that type.

int dsdhost_run_raw_typed(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                          int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int64_t pre_steps, int64_t nframes,
                          int64_t chunk_frames, double scale, double dither, double clip, int lut_bits, uint32_t read_ahead,
                          int type, void* out) {
    MemorySampleReader rd(planar, (uint32_t)nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return 1; }
    if (lut_bits != 64 && !dec.setLutPrecision(lut_bits)) { g_error = dec.getErrorMsg(); return 1; }
    if (read_ahead) dec.setReadAheadFrames(read_ahead);
    for (int64_t i = 0; i < pre_steps; ++i) dec.step();
    if (chunk_frames <= 0) chunk_frames = nframes;
    switch (type) {
    case 0: raw_calls<dsf2flac_int16>(dec, nch, nframes, chunk_frames, scale, dither, clip, out); break;
    case 1: raw_calls<dsf2flac_int32>(dec, nch, nframes, chunk_frames, scale, dither, clip, out); break;
    case 2: raw_calls<dsf2flac_int64>(dec, nch, nframes, chunk_frames, scale, dither, clip, out); break;
    case 3: raw_calls<dsf2flac_float32>(dec, nch, nframes, chunk_frames, scale, dither, clip, out); break;
    case 4: raw_calls<dsf2flac_float64>(dec, nch, nframes, chunk_frames, scale, dither, clip, out); break;
    default: g_error = "unknown sample type"; return 1;
    }
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return 1; }
    return 0;
}

// One whole track, --onefile bounds (reference main.cpp:249-258) and the loop of
// main.cpp:169-191: creep with step(), blocks of 1024 frames, then single frames.
// Returns frames written, -1 on error, -2 if `out` is too small.
static volatile uint32_t g_consumed = 0;   // keeps the consumer of the out == NULL mode alive

static int64_t run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                       uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs, double scale,
                       double dither, double clip, int lut_bits, uint32_t read_ahead, int32_t* out,
                       int64_t out_cap_frames, bool bulk) {
    MemorySampleReader rd(planar, (uint32_t)nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    if (lut_bits != 64 && !dec.setLutPrecision(lut_bits)) { g_error = dec.getErrorMsg(); return -1; }
    if (bulk) {
        // the reader would deliver bytes 0 .. ceil(L/8) (as far as they exist), idle bytes afterwards
        int64_t delivered = (length_samples + 7) / 8 + 1;
        if (delivered > nbytes) delivered = nbytes;
        if (!dec.setBulkSource(planar, 0, delivered, nbytes)) { g_error = dec.getErrorMsg(); return -1; }
    }
    if (read_ahead) dec.setReadAheadFrames(read_ahead);
    double startPos = dec.getFirstValidSample();
    double endPos = dec.getLastValidSample();
    if (startPos > dec.getLength() - 1) startPos = dec.getLength() - 1;
    if (endPos > dec.getLength()) endPos = dec.getLength();
    const int block = 1024;
    while (dec.getPosition() < startPos) dec.step();
    int64_t done = 0;
    std::vector<int32_t> buf((size_t)nch * block);
    // out == NULL: the caller only wants the loop run the way main.cpp runs it -- one block buffer, reused, handed to a
    // consumer that reads it (main.cpp:173-178: the FLAC encoder) -- not a second copy of the whole PCM stream
    uint32_t consumed = 0;
    while (dec.getPosition() <= endPos - block) {
        dec.getSamples(buf.data(), (dsf2flac_uint32)(nch * block), scale, dither, clip);
        if (out) {
            if (done + block > out_cap_frames) return -2;
            std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch * block);
        } else {
            for (int32_t v : buf) consumed += (uint32_t)v;
        }
        done += block;
    }
    while (dec.getPosition() <= endPos) {
        dec.getSamples(buf.data(), (dsf2flac_uint32)nch, scale, dither, clip);
        if (out) {
            if (done + 1 > out_cap_frames) return -2;
            std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch);
        } else {
            for (int c = 0; c < nch; ++c) consumed += (uint32_t)buf[(size_t)c];
        }
        done += 1;
    }
    g_consumed = consumed;
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    return done;
}

int64_t dsdhost_run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                        uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs, double scale,
                        double dither, double clip, int lut_bits, uint32_t read_ahead, int32_t* out,
                        int64_t out_cap_frames) {
    return run_app(planar, nch, nbytes, length_samples, idle, msb_flag, dsd_fs, out_fs, scale, dither, clip, lut_bits,
                   read_ahead, out, out_cap_frames, false);
}

// The same loop with the bytes handed over as a bulk source (DsdDecimator::setBulkSource) instead
// of being stepped out of the reader one byte per channel at a time.
int64_t dsdhost_run_app_bulk(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                             uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs, double scale,
                             double dither, double clip, int lut_bits, uint32_t read_ahead, int32_t* out,
                             int64_t out_cap_frames) {
    return run_app(planar, nch, nbytes, length_samples, idle, msb_flag, dsd_fs, out_fs, scale, dither, clip, lut_bits,
                   read_ahead, out, out_cap_frames, true);
}

// A scripted caller, for tests of the block cache and the read-ahead: ops[2k] = what, ops[2k+1] = argument.
//   0: step() x arg          1: getSamples(arg frames, scale_a)     2: getSamples(arg frames, scale_b)
//   3: setLutPrecision(arg)  4: setReadAheadFrames(arg)
// int32 output frames are appended to `out`.  Returns frames written, -1 on error, -2 if `out` is too small.
int64_t dsdhost_run_script(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                           int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int bulk, const int64_t* ops, int nops,
                           double scale_a, double scale_b, double dither, double clip, int32_t* out,
                           int64_t out_cap_frames) {
    MemorySampleReader rd(planar, (uint32_t)nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    if (bulk) {
        int64_t delivered = (length_samples + 7) / 8 + 1;
        if (delivered > nbytes) delivered = nbytes;
        if (!dec.setBulkSource(planar, 0, delivered, nbytes)) { g_error = dec.getErrorMsg(); return -1; }
    }
    int64_t done = 0;
    for (int k = 0; k < nops; ++k) {
        const int64_t what = ops[2 * k], arg = ops[2 * k + 1];
        if (what == 0) {
            for (int64_t i = 0; i < arg; ++i) dec.step();
        } else if (what == 1 || what == 2) {
            if (done + arg > out_cap_frames) return -2;
            dec.getSamples(out + done * nch, (dsf2flac_uint32)(arg * nch), what == 1 ? scale_a : scale_b, dither, clip);
            done += arg;
        } else if (what == 3) {
            if (!dec.setLutPrecision((int)arg)) { g_error = dec.getErrorMsg(); return -1; }
        } else if (what == 4) {
            dec.setReadAheadFrames((dsf2flac_uint32)arg);
        }
        if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    }
    return done;
}

// setLutPrecision() AFTER setBulkSource() on a block-planar (DSF-style) source: the context is recreated and must keep
// the source layout (DsdDecimator::openGpu).  `blocks` holds the stream in blocks of block_bytes per channel, channel after
// channel; precision goes 64 -> 32 -> 64 before the first sample is asked for.  int32 frames to `out`; returns frames or -1.
int64_t dsdhost_run_precision_after_bulk(const uint8_t* blocks, int nch, int64_t nbytes, int64_t block_bytes, int64_t length_samples,
                                         int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int final_bits, double scale, double dither,
                                         double clip, int64_t nframes, int32_t* out) {
    const uint8_t dummy = 0;
    MemorySampleReader rd(&dummy, (uint32_t)nch, 0, length_samples, dsd_fs, msb_flag != 0, 0x69);     // geometry only
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    if (!dec.setBulkSource(blocks, block_bytes, nbytes)) { g_error = dec.getErrorMsg(); return -1; }
    if (!dec.setLutPrecision(32) || !dec.setLutPrecision(64) || (final_bits != 64 && !dec.setLutPrecision(final_bits))) {
        g_error = dec.getErrorMsg();
        return -1;
    }
    while (dec.getPosition() < dec.getFirstValidSample()) dec.step();
    dec.getSamples(out, (dsf2flac_uint32)(nframes * nch), scale, dither, clip);
    if (!dec.isValid()) { g_error = dec.getErrorMsg(); return -1; }
    return nframes;
}

// Header facts of a DSF / DSDIFF file as the reference's readers report them, plus what the bulk
// path derives: info = {channels, sampling frequency, msbIsPlayedFirst, kind (1 DSF, 2 DSDIFF),
// source_block_bytes}; *length_samples = getLength(); *delivered_bytes = stream bytes per channel
// the reference reader hands out before idle bytes.  Needs no GPU.  Returns 1, or 0 + error text.
int dsdhost_file_info(const char* path, int kind, int32_t* info, int64_t* length_samples, int64_t* delivered_bytes) {
    DsdContainer c;
    if (!dsd_container_load(path, kind, c)) { g_error = c.error; return 0; }
    info[0] = (int32_t)c.channels; info[1] = (int32_t)c.samplingFreq; info[2] = c.msbIsPlayedFirst ? 1 : 0;
    info[3] = (int32_t)c.kind; info[4] = (int32_t)c.blockBytes;
    *length_samples = c.lengthSamples;
    *delivered_bytes = c.deliveredBytes;
    return 1;
}

// The stream bytes of channel `ch` the container path will feed the GPU, [0, delivered) -> planar
// (for CPU-side checks of the container model).  Returns bytes written or -1.
int64_t dsdhost_file_channel_bytes(const char* path, int kind, int ch, uint8_t* out, int64_t cap) {
    DsdContainer c;
    if (!dsd_container_load(path, kind, c)) { g_error = c.error; return -1; }
    if (ch < 0 || ch >= (int)c.channels) return -1;
    const int64_t n = c.deliveredBytes < cap ? c.deliveredBytes : cap, nch = c.channels, B = c.blockBytes;
    for (int64_t k = 0; k < n; ++k)
        out[k] = B == 1 ? c.payload[(size_t)(k * nch + ch)] : c.payload[(size_t)(((k / B) * nch + ch) * B + k % B)];
    return n;
}

// One whole file, what `dsf2flac -i file` computes before FLAC (reference main.cpp:232-258 with
// onefile, :169-191), with the file's sample bytes handed to the GPU in their own layout instead of
// being stepped through a reader.  Returns frames written, -1 on error, -2 if `out` is too small.
int64_t dsdhost_convert_file(const char* path, int kind, uint32_t out_fs, double scale, double dither, double clip,
                             int lut_bits, uint32_t read_ahead, int32_t* out, int64_t out_cap_frames) {
    DsdContainer c;
    if (!dsd_container_load(path, kind, c)) { g_error = c.error; return -1; }
    int64_t done = -1;
    {
        ContainerSampleReader rd(c);
        DsdDecimator dec(&rd, out_fs);
        bool ok = dec.isValid();
        if (ok && lut_bits != 64) ok = dec.setLutPrecision(lut_bits);
        if (ok) ok = dec.setBulkSource(c.payload.data(), c.blockBytes, c.deliveredBytes);
        if (!ok) {
            g_error = dec.getErrorMsg();
        } else {
            if (read_ahead) dec.setReadAheadFrames(read_ahead);
            double startPos = dec.getFirstValidSample();
            double endPos = dec.getLastValidSample();
            if (startPos > dec.getLength() - 1) startPos = dec.getLength() - 1;
            if (endPos > dec.getLength()) endPos = dec.getLength();
            const int block = 1024;
            const int nch = (int)c.channels;
            while (dec.getPosition() < startPos) dec.step();
            done = 0;
            while (dec.getPosition() <= endPos - block) {
                if (done + block > out_cap_frames) { done = -2; break; }
                dec.getSamples(out + done * nch, (dsf2flac_uint32)(nch * block), scale, dither, clip);
                done += block;
            }
            while (done >= 0 && dec.getPosition() <= endPos) {
                if (done + 1 > out_cap_frames) { done = -2; break; }
                dec.getSamples(out + done * nch, (dsf2flac_uint32)nch, scale, dither, clip);
                done += 1;
            }
            if (!dec.isValid()) { g_error = dec.getErrorMsg(); done = -1; }
        }
    }
    return done;
}

// ---- DoP (reference dop_packer.cpp + dop_track_helper, main.cpp:294-403) -------------------------
// The frames dop_track_helper emits for a whole file (--onefile: start 0, end getLength(),
// main.cpp:423-424): creep with step() to position >= start, blocks of 1024 frames while position <=
// end - 1024*16, single frames while position <= end; a frame advances the reader by two bytes.
// Returns the number of frames; *pm0 = the reader's marker when the first frame is packed.
static int64_t dop_app_frames(int64_t length_samples, int64_t* pm0) {
    int64_t start = 0, end = length_samples;
    if (start > length_samples - 1) start = length_samples - 1;
    int64_t pm = -1, n = 0;
    while (8 * pm < start) pm += 1;
    *pm0 = pm;
    if (8 * pm <= end - 1024 * 16) {
        const int64_t blocks = (end - 1024 * 16 - 8 * pm) / (16 * 1024) + 1;
        n += 1024 * blocks; pm += 2048 * blocks;
    }
    if (8 * pm <= end) n += (end - 8 * pm) / 16 + 1;
    return n;
}

static int64_t dop_pack_all(const uint8_t* data, size_t stride, int64_t block_bytes, int64_t delivered, int nch,
                            int64_t length_samples, uint8_t idle, int msb_flag, int32_t* out, int64_t out_cap_frames) {
    int64_t pm0 = -1;
    const int64_t n = dop_app_frames(length_samples, &pm0);
    if (n > out_cap_frames) return -2;
    const double dummy[256] = {0.0};                       // the packer uses no tables
    dsdgpu_ctx* gpu = nullptr;
    int device = 0;
    if (const char* env = std::getenv("DSDGPU_DEVICE")) device = std::atoi(env);
    if (dsdgpu_create(&gpu, device, nch, 1, 1, dummy, 64) != DSDGPU_OK) { g_error = dsdgpu_last_error(nullptr); return -1; }
    int rc = dsdgpu_set_option(gpu, "source_block_bytes", block_bytes);
    if (rc == DSDGPU_OK) rc = dsdgpu_dop_pack_host(gpu, data, stride, 0, delivered, idle, pm0, n, msb_flag, out);
    if (rc != DSDGPU_OK) g_error = dsdgpu_last_error(gpu);
    dsdgpu_destroy(gpu);
    return rc == DSDGPU_OK ? n : -1;
}

// Planar bytes in memory -> DoP frames of a whole stream.
int64_t dsdhost_dop_run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                            int msb_flag, int32_t* out, int64_t out_cap_frames) {
    int64_t delivered = (length_samples + 7) / 8 + 1;        // the reader hands out bytes 0 .. ceil(L/8), as far as they exist
    if (delivered > nbytes) delivered = nbytes;
    return dop_pack_all(planar, (size_t)nbytes, 0, delivered, nch, length_samples, idle, msb_flag, out, out_cap_frames);
}

// One DSF / DSDIFF file -> DoP frames, what `dsf2flac -d` computes before FLAC.
int64_t dsdhost_dop_file(const char* path, int kind, int32_t* out, int64_t out_cap_frames) {
    DsdContainer c;
    if (!dsd_container_load(path, kind, c)) { g_error = c.error; return -1; }
    const int64_t n = dop_pack_all(c.payload.data(), 0, c.blockBytes, c.deliveredBytes, (int)c.channels, c.lengthSamples,
                                   c.idle, c.msbIsPlayedFirst ? 1 : 0, out, out_cap_frames);
    return n;
}

}  // extern "C"
