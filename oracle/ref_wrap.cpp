// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// C-ABI wrapper around the UNMODIFIED reference decimator and container readers.  oracle/Makefile
// compiles /root/reference/src/dsd_decimator.cpp (+ filters.cpp through its #include),
// dsd_sample_reader.cpp, dsf_file_reader.cpp, dsdiff_file_reader.cpp, fstream_plus.cpp and
// libdstdec/*.c where they lie, against oracle/ref_shim, together with this file, into
// oracle/_ref/libdsdref.so.  This file adds only (1) an in-memory DsdSampleReader whose
// step() follows dsf_file_reader.cpp:83-103 and (2) drivers shaped like main.cpp:169-191.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// The lookup table is a private member (dsd_decimator.h:139); the unit tests want to compare
// it entry by entry, so open the class up for this translation unit only.
#define private public
#include "dsd_decimator.h"
#undef private
#include "dop_packer.h"
#include "dsdiff_file_reader.h"
#include "dsf_file_reader.h"

#include "app_driver.h"
#include "ref_memory_reader.h"

namespace {

// EXTENSION filters.  The reference rejects every ratio but 8, 16 and 32
// (dsd_decimator.cpp:57-68).  To run ITS arithmetic -- initLookupTable and getSamplesInternal, both
// unmodified -- on a longer impulse response, the decimator is constructed for ratio 32 (valid) and
// then given the registered filter through its own private initLookupTable.
struct CustomFilter { int ratio = 0, ntaps = 0, tzero = 0; std::vector<double> coefs; } g_custom;

class RefDecimator : public DsdDecimator {
public:
    RefDecimator(DsdSampleReader* rd, uint32_t out_fs)
        : DsdDecimator(rd, custom_for(rd, out_fs) ? rd->getSamplingFreq() / 32 : out_fs) {
        if (!custom_for(rd, out_fs)) return;
        for (dsf2flac_uint32 n = 0; n < nLookupTable; n++) delete[] lookupTable[n];
        delete[] lookupTable;
        outputSampleRate = out_fs;
        ratio = (dsf2flac_uint32)g_custom.ratio;
        nStep = ratio / 8;
        initLookupTable(g_custom.ntaps, g_custom.coefs.data(), g_custom.tzero);
        if (nLookupTable > rd->getBufferLength()) rd->setBufferLength(nLookupTable);
    }
private:
    static bool custom_for(DsdSampleReader* rd, uint32_t out_fs) {
        return g_custom.ratio > 0 && out_fs != 0 && rd->getSamplingFreq() / out_fs == (uint32_t)g_custom.ratio;
    }
};

}  // namespace

extern "C" {

int dsdref_set_custom_filter(int ratio, int ntaps, int tzero, const double* coefs) {
    g_custom = CustomFilter();
    if (ratio <= 0 || ntaps <= 0 || !coefs) return 0;
    g_custom.ratio = ratio; g_custom.ntaps = ntaps; g_custom.tzero = tzero;
    g_custom.coefs.assign(coefs, coefs + ntaps);
    return 1;
}

// Geometry the reference derives for (dsd_fs, out_fs).  Returns 0 if the pair is rejected.
int dsdref_info(uint32_t dsd_fs, uint32_t out_fs, int64_t length_samples, int32_t* ratio,
                int32_t* nlut, int32_t* nstep, int32_t* tzero, double* first_valid,
                double* last_valid, int64_t* length_pcm) {
    uint8_t dummy = 0;
    MemoryReader rd(&dummy, 1, 0, length_samples, dsd_fs, true, 0x69);
    RefDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return 0;
    *ratio = (int32_t)dec.getDecimationRatio();
    *nlut = (int32_t)dec.nLookupTable;
    *nstep = (int32_t)dec.nStep;
    *tzero = (int32_t)dec.tzero;
    *first_valid = dec.getFirstValidSample();
    *last_valid = dec.getLastValidSample();
    *length_pcm = dec.getLength();
    return 1;
}

// Copies the reference's lookup table, lut[t*256 + s].  Returns nLookupTable (0 = invalid).
int dsdref_lut(uint32_t dsd_fs, uint32_t out_fs, int msb_flag, double* lut, int cap_tables) {
    uint8_t dummy = 0;
    MemoryReader rd(&dummy, 1, 0, 0, dsd_fs, msb_flag != 0, 0x69);
    RefDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return 0;
    int n = (int)dec.nLookupTable;
    for (int t = 0; t < n && t < cap_tables; ++t)
        for (int s = 0; s < 256; ++s) lut[t * 256 + s] = (double)dec.lookupTable[t][s];
    return n;
}

// Raw operator: rewind, `pre_steps` x step(), then getSamples in calls of `chunk_frames` frames
// (0 = one call).  reseed != 0 -> srand(1) first, i.e. the dither stream of a fresh process.
// out_i32 and/or out_f64 (either may be NULL) are [nframes][nch].  Returns 0 on success.
int dsdref_run_raw(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                   uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs,
                   int64_t pre_steps, int64_t nframes, int64_t chunk_frames, double scale,
                   double dither, double clip, int reseed, int32_t* out_i32, double* out_f64) {
    MemoryReader rd(planar, nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    RefDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return 1;
    if (reseed) srand(1);
    for (int64_t i = 0; i < pre_steps; ++i) dec.step();
    if (chunk_frames <= 0) chunk_frames = nframes;
    // out_f64 is produced from a second, identically driven decimator so both views see the
    // same dither draws only when dither == 0; with dither on ask for one view at a time.
    if (out_i32) {
        for (int64_t f = 0; f < nframes; f += chunk_frames) {
            int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
            dec.getSamples(out_i32 + f * nch, (dsf2flac_uint32)(n * nch), scale, dither, clip);
        }
    }
    if (out_f64) {
        MemoryReader rd2(planar, nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
        RefDecimator dec2(&rd2, out_fs);
        if (reseed) srand(1);
        for (int64_t i = 0; i < pre_steps; ++i) dec2.step();
        for (int64_t f = 0; f < nframes; f += chunk_frames) {
            int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
            dec2.getSamples(out_f64 + f * nch, (dsf2flac_uint32)(n * nch), scale, dither, clip);
        }
    }
    return 0;
}

// Whole-track conversion with the loop shape of pcm_track_helper (main.cpp:169-191) for the
// --onefile case: startPos = getFirstValidSample(), endPos = getLastValidSample()
// (main.cpp:249-258).  Returns frames written, or -1 (invalid) / -2 (out too small).
int64_t dsdref_run_app(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples,
                       uint8_t idle, int msb_flag, uint32_t dsd_fs, uint32_t out_fs,
                       double scale, double dither, double clip, int reseed, int32_t* out,
                       int64_t out_cap_frames) {
    MemoryReader rd(planar, nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    RefDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return -1;
    if (reseed) srand(1);
    double startPos = dec.getFirstValidSample();
    double endPos = dec.getLastValidSample();
    if (startPos > dec.getLength() - 1) startPos = dec.getLength() - 1;
    if (endPos > dec.getLength()) endPos = dec.getLength();
    const int block = 1024;
    while (dec.getPosition() < startPos) dec.step();
    int64_t done = 0;
    std::vector<int32_t> buf((size_t)nch * block);
    while (dec.getPosition() <= endPos - block) {
        dec.getSamples(buf.data(), (dsf2flac_uint32)(nch * block), scale, dither, clip);
        if (done + block > out_cap_frames) return -2;
        std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch * block);
        done += block;
    }
    while (dec.getPosition() <= endPos) {
        dec.getSamples(buf.data(), (dsf2flac_uint32)nch, scale, dither, clip);
        if (done + 1 > out_cap_frames) return -2;
        std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch);
        done += 1;
    }
    return done;
}

// ---- file level: the reference's own container readers in front of its decimator ------------
// kind 1 = .dsf (DsfFileReader, dsf_file_reader.cpp), 2 = .dff (DsdiffFileReader,
// dsdiff_file_reader.cpp).  info[0..3] = channels, sampling frequency, msbIsPlayedFirst, valid;
// *length = getLength() (DSD samples per channel).  Returns 1 if the reader accepted the file.
// The readers are constructed in zero-filled memory: DsdSampleReader's constructor leaves posMarker
// unset (dsd_sample_reader.cpp:40-44) and DsdiffFileReader::rewind() reads it before assigning it
// (dsdiff_file_reader.cpp:152-154 -> readNextBlock :167), so on recycled heap memory the first
// 1024-byte block can come out as idle bytes.  Zeroed storage is what a fresh process's first
// allocations look like and gives the behaviour the code intends (posMarker = 0 there).
static DsdSampleReader* open_reader(const char* path, int kind) {
    if (kind == 1) return new (calloc(1, sizeof(DsfFileReader))) DsfFileReader((char*)path);
    if (kind == 2) return new (calloc(1, sizeof(DsdiffFileReader))) DsdiffFileReader((char*)path);
    return nullptr;
}
static void close_reader(DsdSampleReader* rd) {
    if (!rd) return;
    rd->~DsdSampleReader();
    free(rd);
}

int dsdref_file_info(const char* path, int kind, int32_t* info, int64_t* length) {
    DsdSampleReader* rd = open_reader(path, kind);
    if (!rd) return 0;
    const int ok = rd->isValid() ? 1 : 0;
    if (ok) {
        info[0] = (int32_t)rd->getNumChannels();
        info[1] = (int32_t)rd->getSamplingFreq();
        info[2] = rd->msbIsPlayedFirst() ? 1 : 0;
        info[3] = 1;
        *length = rd->getLength();
    }
    close_reader(rd);
    return ok;
}

// What `dsf2flac -i file` computes before FLAC (main.cpp:232-258 with onefile, :169-191).
// Returns frames written, -1 reader invalid, -3 decimator invalid, -2 out too small.
// NOTE: dsf_file_reader.cpp:40 keeps `blockBufferAllocated` in a file-scope static that is never
// reset, so a process can open ONE .dsf file; callers run each conversion in a fresh process and get
// the reader's geometry from the same call (info[0..2] = channels, sampling frequency,
// msbIsPlayedFirst; *length = getLength(); either may be NULL).
int64_t dsdref_file_run_app(const char* path, int kind, uint32_t out_fs, double scale, double dither,
                            double clip, int reseed, int32_t* out, int64_t out_cap_frames,
                            int32_t* info, int64_t* length) {
    DsdSampleReader* rd = open_reader(path, kind);
    if (!rd || !rd->isValid()) { close_reader(rd); return -1; }
    if (info) {
        info[0] = (int32_t)rd->getNumChannels();
        info[1] = (int32_t)rd->getSamplingFreq();
        info[2] = rd->msbIsPlayedFirst() ? 1 : 0;
    }
    if (length) *length = rd->getLength();
    int64_t done = 0;
    {
        const int nch = (int)rd->getNumChannels();
        DsdDecimator dec(rd, out_fs);
        if (!dec.isValid()) { close_reader(rd); return -3; }
        if (reseed) srand(1);
        double startPos = dec.getFirstValidSample();
        double endPos = dec.getLastValidSample();
        if (startPos > dec.getLength() - 1) startPos = dec.getLength() - 1;
        if (endPos > dec.getLength()) endPos = dec.getLength();
        const int block = 1024;
        while (dec.getPosition() < startPos) dec.step();
        std::vector<int32_t> buf((size_t)nch * block);
        while (dec.getPosition() <= endPos - block) {
            dec.getSamples(buf.data(), (dsf2flac_uint32)(nch * block), scale, dither, clip);
            if (done + block > out_cap_frames) { done = -2; break; }
            std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch * block);
            done += block;
        }
        while (done >= 0 && dec.getPosition() <= endPos) {
            dec.getSamples(buf.data(), (dsf2flac_uint32)nch, scale, dither, clip);
            if (done + 1 > out_cap_frames) { done = -2; break; }
            std::memcpy(out + done * nch, buf.data(), sizeof(int32_t) * nch);
            done += 1;
        }
    }
    close_reader(rd);
    return done;
}

// ---- DoP: the reference's DopPacker (dop_packer.cpp) ------------------------------------------
// rewind, `pre_steps` x reader step(), then pack_buffer in calls of `chunk_frames` frames (0 = one call).
int dsdref_dop_run_raw(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                       int msb_flag, int64_t pre_steps, int64_t nframes, int64_t chunk_frames, int32_t* out) {
    MemoryReader rd(planar, nch, nbytes, length_samples, 2822400, msb_flag != 0, idle);
    DopPacker dopp(&rd);
    for (int64_t i = 0; i < pre_steps; ++i) rd.step();
    if (chunk_frames <= 0) chunk_frames = nframes;
    for (int64_t f = 0; f < nframes; f += chunk_frames) {
        const int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
        dopp.pack_buffer(out + f * nch, (dsf2flac_uint32)(n * nch));
    }
    return 0;
}

// What `dsf2flac -d -i file` computes before FLAC: dop_track_helper's loops (main.cpp:300-306,
// 362-383) with the --onefile bounds (:423-424), over the reference's own file reader.
// Returns frames written, -1 reader invalid, -2 out too small.  One call per process (see above).
int64_t dsdref_file_run_dop(const char* path, int kind, int32_t* out, int64_t out_cap_frames,
                            int32_t* info, int64_t* length) {
    DsdSampleReader* dsr = open_reader(path, kind);
    if (!dsr || !dsr->isValid()) { close_reader(dsr); return -1; }
    if (info) {
        info[0] = (int32_t)dsr->getNumChannels();
        info[1] = (int32_t)dsr->getSamplingFreq();
        info[2] = dsr->msbIsPlayedFirst() ? 1 : 0;
    }
    if (length) *length = dsr->getLength();
    int64_t done = 0;
    {
        const int nch = (int)dsr->getNumChannels();
        dsf2flac_int64 startPos = 0, endPos = dsr->getLength();
        if (startPos > dsr->getLength() - 1) startPos = dsr->getLength() - 1;
        if (endPos > dsr->getLength()) endPos = dsr->getLength();
        DopPacker dopp(dsr);
        while (dsr->getPosition() < startPos) dsr->step();
        const int block = 1024;
        while (dsr->getPosition() <= endPos - block * 16) {
            if (done + block > out_cap_frames) { done = -2; break; }
            dopp.pack_buffer(out + done * nch, (dsf2flac_uint32)(nch * block));
            done += block;
        }
        while (done >= 0 && dsr->getPosition() <= endPos) {
            if (done + 1 > out_cap_frames) { done = -2; break; }
            dopp.pack_buffer(out + done * nch, (dsf2flac_uint32)nch);
            done += 1;
        }
    }
    close_reader(dsr);
    return done;
}

// ---- all five getSamples<T> specialisations (dsd_decimator.cpp:153-172), driven by oracle/app_driver.h ----
// type: 0 int16, 1 int32, 2 int64, 3 float32, 4 float64; out is [nframes][nch] of that type.
int dsdref_run_raw_typed(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                         int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int64_t pre_steps, int64_t nframes,
                         int64_t chunk_frames, double scale, double dither, double clip, int reseed, int type, void* out) {
    MemoryReader rd(planar, nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    RefDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return 1;
    if (reseed) srand(1);
    return app_driver::raw_loop_typed(dec, nch, type, pre_steps, nframes, chunk_frames, scale, dither, clip, out);
}

// One file through the reference's own reader + decimator with main.cpp's loop, any sample type.
// Returns frames written, -1 reader invalid, -3 decimator invalid, -2 out too small.  One call per process.
int64_t dsdref_file_run_typed(const char* path, int kind, uint32_t out_fs, double scale, double dither, double clip,
                              int reseed, int type, void* out, int64_t out_cap_frames, int32_t* info, int64_t* length) {
    DsdSampleReader* rd = open_reader(path, kind);
    if (!rd || !rd->isValid()) { close_reader(rd); return -1; }
    if (info) {
        info[0] = (int32_t)rd->getNumChannels();
        info[1] = (int32_t)rd->getSamplingFreq();
        info[2] = rd->msbIsPlayedFirst() ? 1 : 0;
    }
    if (length) *length = rd->getLength();
    int64_t done;
    {
        DsdDecimator dec(rd, out_fs);
        if (!dec.isValid()) { close_reader(rd); return -3; }
        if (reseed) srand(1);
        done = app_driver::app_loop_typed(dec, (int)rd->getNumChannels(), type, scale, dither, clip, out, out_cap_frames);
    }
    close_reader(rd);
    return done;
}


// libc rand() as the reference sees it: srand(1) then n draws (pins the restated generator).
void dsdref_libc_rand(int32_t* out, int64_t n) {
    srand(1);
    for (int64_t i = 0; i < n; ++i) out[i] = rand();
}

}  // extern "C"
