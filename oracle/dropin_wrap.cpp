// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// "The drop-in is a drop-in": this file is compiled together with
//   * the REPO's  dsf2flac_b200/host/dsd_decimator.{h,cpp}   (the class under test), which here see
//   * the REFERENCE's own  src/dsd_sample_reader.h           (through oracle/ref_shim for Boost / id3lib), and
//   * the REFERENCE's unmodified  dsd_sample_reader.cpp, dsf_file_reader.cpp, dsdiff_file_reader.cpp,
//     fstream_plus.cpp and libdstdec/*.c, compiled where they lie under /root/reference/src
// -- i.e. exactly what a dsf2flac maintainer gets after swapping the two decimator files (INTEGRATION.md
// section 1), minus main.cpp's FLAC encoder.  The caller side is oracle/app_driver.h, the same source that
// drives the all-reference build (oracle/ref_wrap.cpp), so the two libraries differ in nothing but the class.
//   _ref/libdropin_ref.so      linked with dsf2flac_b200/libdsdgpu.so          (GPU tests)
//   _ref/libdropin_ref_cpu.so  linked with tests/cpp/standin_dsdgpu.cpp        (CPU tests of the host logic)
#include <cstdint>
#include <cstdlib>
#include <new>

#include "dsd_decimator.h"          // the repo's class: -I dsf2flac_b200/host comes before -I $(REF)/src
#include "dsdiff_file_reader.h"     // the reference's readers
#include "dsf_file_reader.h"

#include "app_driver.h"
#include "ref_memory_reader.h"

namespace {

// zero-filled storage: see oracle/ref_wrap.cpp (DsdiffFileReader::rewind() reads posMarker before assigning it)
DsdSampleReader* open_reader(const char* path, int kind) {
    if (kind == 1) return new (calloc(1, sizeof(DsfFileReader))) DsfFileReader((char*)path);
    if (kind == 2) return new (calloc(1, sizeof(DsdiffFileReader))) DsdiffFileReader((char*)path);
    return nullptr;
}
void close_reader(DsdSampleReader* rd) {
    if (!rd) return;
    rd->~DsdSampleReader();
    free(rd);
}

}  // namespace

extern "C" {

// The reference's file reader, untouched, stepped by the repo's decimator; main.cpp's loop; any getSamples<T>.
// type: 0 int16, 1 int32, 2 int64, 3 float32, 4 float64.  read_ahead: DsdDecimator::setReadAheadFrames (0 = default).
// Returns frames written, -1 reader invalid, -3 decimator invalid, -2 out too small.  One call per process
// (the reference's readers keep "allocated" flags in file-scope statics).
int64_t dropin_file_run_typed(const char* path, int kind, uint32_t out_fs, double scale, double dither, double clip,
                              uint32_t read_ahead, int type, void* out, int64_t out_cap_frames, int32_t* info,
                              int64_t* length, char* err, int err_cap) {
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
        if (!dec.isValid()) {
            if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", dec.getErrorMsg().c_str());
            close_reader(rd);
            return -3;
        }
        if (read_ahead) dec.setReadAheadFrames(read_ahead);
        done = app_driver::app_loop_typed(dec, (int)rd->getNumChannels(), type, scale, dither, clip, out, out_cap_frames);
    }
    close_reader(rd);
    return done;
}

// The same class over an in-memory subclass of the reference's DsdSampleReader: raw getSamples<T> calls.
int dropin_run_raw_typed(const uint8_t* planar, int nch, int64_t nbytes, int64_t length_samples, uint8_t idle,
                         int msb_flag, uint32_t dsd_fs, uint32_t out_fs, int64_t pre_steps, int64_t nframes,
                         int64_t chunk_frames, double scale, double dither, double clip, uint32_t read_ahead, int type,
                         void* out) {
    MemoryReader rd(planar, nch, nbytes, length_samples, dsd_fs, msb_flag != 0, idle);
    DsdDecimator dec(&rd, out_fs);
    if (!dec.isValid()) return 1;
    if (read_ahead) dec.setReadAheadFrames(read_ahead);
    return app_driver::raw_loop_typed(dec, nch, type, pre_steps, nframes, chunk_frames, scale, dither, clip, out);
}

}  // extern "C"
