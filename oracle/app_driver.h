// TEST INFRASTRUCTURE ONLY -- never linked into the product path.
//
// The caller side of the drop-in boundary, written once and compiled twice: against the REFERENCE's
// DsdDecimator (oracle/ref_wrap.cpp -> _ref/libdsdref.so) and against the REPO's DsdDecimator built over
// the reference's own dsd_sample_reader.h and file readers (oracle/dropin_wrap.cpp ->
// _ref/libdropin_ref*.so).  Whatever class named DsdDecimator the including translation unit sees is
// driven exactly as pcm_track_helper drives it (reference src/main.cpp:169-191, --onefile bounds
// main.cpp:249-258), through any of the five getSamples<T> specialisations
// (reference src/dsd_decimator.cpp:153-172).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace app_driver {

// type codes shared with tests/: 0 int16, 1 int32, 2 int64, 3 float32, 4 float64
inline int sample_size(int type) { return type == 0 ? 2 : type == 1 ? 4 : type == 2 ? 8 : type == 3 ? 4 : 8; }

// creep to the first valid sample, blocks of 1024 frames, then single frames.
// Returns frames written, -2 if `out` is too small.
template <typename T, typename Dec>
int64_t app_loop(Dec& dec, int nch, double scale, double dither, double clip, T* out, int64_t cap_frames) {
    double startPos = dec.getFirstValidSample();
    double endPos = dec.getLastValidSample();
    if (startPos > dec.getLength() - 1) startPos = dec.getLength() - 1;
    if (endPos > dec.getLength()) endPos = dec.getLength();
    const int block = 1024;
    while (dec.getPosition() < startPos) dec.step();
    int64_t done = 0;
    std::vector<T> buf((size_t)nch * block);
    while (dec.getPosition() <= endPos - block) {
        dec.getSamples(buf.data(), (uint32_t)(nch * block), scale, dither, clip);
        if (done + block > cap_frames) return -2;
        std::memcpy(out + done * nch, buf.data(), sizeof(T) * nch * block);
        done += block;
    }
    while (dec.getPosition() <= endPos) {
        dec.getSamples(buf.data(), (uint32_t)nch, scale, dither, clip);
        if (done + 1 > cap_frames) return -2;
        std::memcpy(out + done * nch, buf.data(), sizeof(T) * nch);
        done += 1;
    }
    return done;
}

template <typename Dec>
int64_t app_loop_typed(Dec& dec, int nch, int type, double scale, double dither, double clip, void* out, int64_t cap_frames) {
    switch (type) {
    case 0: return app_loop(dec, nch, scale, dither, clip, static_cast<int16_t*>(out), cap_frames);
    case 1: return app_loop(dec, nch, scale, dither, clip, static_cast<int32_t*>(out), cap_frames);
    case 2: return app_loop(dec, nch, scale, dither, clip, static_cast<int64_t*>(out), cap_frames);
    case 3: return app_loop(dec, nch, scale, dither, clip, static_cast<float*>(out), cap_frames);
    case 4: return app_loop(dec, nch, scale, dither, clip, static_cast<double*>(out), cap_frames);
    }
    return -4;
}

// rewind (done by the caller), `pre_steps` x step(), then getSamples<T> in calls of `chunk_frames` frames
template <typename T, typename Dec>
void raw_loop(Dec& dec, int nch, int64_t pre_steps, int64_t nframes, int64_t chunk_frames, double scale, double dither,
              double clip, T* out) {
    for (int64_t i = 0; i < pre_steps; ++i) dec.step();
    if (chunk_frames <= 0) chunk_frames = nframes;
    for (int64_t f = 0; f < nframes; f += chunk_frames) {
        const int64_t n = nframes - f < chunk_frames ? nframes - f : chunk_frames;
        dec.getSamples(out + f * nch, (uint32_t)(n * nch), scale, dither, clip);
    }
}

template <typename Dec>
int raw_loop_typed(Dec& dec, int nch, int type, int64_t pre_steps, int64_t nframes, int64_t chunk_frames, double scale,
                   double dither, double clip, void* out) {
    switch (type) {
    case 0: raw_loop(dec, nch, pre_steps, nframes, chunk_frames, scale, dither, clip, static_cast<int16_t*>(out)); return 0;
    case 1: raw_loop(dec, nch, pre_steps, nframes, chunk_frames, scale, dither, clip, static_cast<int32_t*>(out)); return 0;
    case 2: raw_loop(dec, nch, pre_steps, nframes, chunk_frames, scale, dither, clip, static_cast<int64_t*>(out)); return 0;
    case 3: raw_loop(dec, nch, pre_steps, nframes, chunk_frames, scale, dither, clip, static_cast<float*>(out)); return 0;
    case 4: raw_loop(dec, nch, pre_steps, nframes, chunk_frames, scale, dither, clip, static_cast<double*>(out)); return 0;
    }
    return 4;
}

}  // namespace app_driver
