// dsd_decimator.cpp -- host side of the B200 drop-in DsdDecimator.
//
// Reference behaviour reproduced here (reference src/dsd_decimator.cpp):
//   :44-72    ratio = dsdFs/outFs, nStep = ratio/8, filter by ratio, "Sorry, incompatible ..."
//   :85-101   getLength / getPosition / getFirstValidSample / getLastValidSample
//   :117-151  initLookupTable -> buildTables()
//   :153-172  getSamples<int16|int32|int64|float32|float64>
//   :182-186  bufferLen not a multiple of the channel count -> message on stderr + exit
//   :191-221  the FIR + scale/dither/clip/round loop -> GPU, via include/dsdgpu.h
// What differs by design: the reader is drained ahead of the consumer (a block of frames per
// GPU round trip), so this class keeps its own LOGICAL position: step() and getSamples() move
// `pos` exactly as reader->step() moved the reference's posMarker, and every position query is
// answered from `pos`, never from how far the reader has been drained.
#include "dsd_decimator.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include "dsdgpu.h"
#include "pinned_pool.h"

namespace {

struct FirSpec { int ratio, ntaps, tzero; const dsf2flac_uint64* bits; };

#define DSD_FILTER_BEGIN(tag, ratio_, taps_, tzero_) \
    const int tag##_ratio = ratio_, tag##_taps = taps_, tag##_tzero = tzero_; \
    const dsf2flac_uint64 tag##_bits[taps_] = {
#define DSD_FILTER_END(tag) };
#include "dsd_filter_bits.inc"
// EXTENSION (no reference counterpart): /64, /128, /256 -- the /32 response stretched to faster DSD
// clocks (tools/design_ext_filters.py).  Off unless asked for: the reference rejects these ratios.
#include "dsd_filter_ext_bits.inc"

const FirSpec kFilters[3] = {
    {div8_ratio, div8_taps, div8_tzero, div8_bits},
    {div16_ratio, div16_taps, div16_tzero, div16_bits},
    {div32_ratio, div32_taps, div32_tzero, div32_bits},
};
const FirSpec kExtensionFilters[3] = {
    {div64_ratio, div64_taps, div64_tzero, div64_bits},
    {div128_ratio, div128_taps, div128_tzero, div128_bits},
    {div256_ratio, div256_taps, div256_tzero, div256_bits},
};

int g_extensionFilters = -1;        // -1: ask the environment (DSDGPU_EXTENSION_FILTERS=1), 0 off, 1 on

bool extensionFiltersEnabled() {
    if (g_extensionFilters >= 0) return g_extensionFilters != 0;
    const char* env = std::getenv("DSDGPU_EXTENSION_FILTERS");
    return env && std::atoi(env) != 0;
}

const FirSpec* findFilter(dsf2flac_uint32 ratio) {
    for (const FirSpec& f : kFilters)
        if ((dsf2flac_uint32)f.ratio == ratio) return &f;
    if (extensionFiltersEnabled())
        for (const FirSpec& f : kExtensionFilters)
            if ((dsf2flac_uint32)f.ratio == ratio) return &f;
    return nullptr;
}

inline double tapValue(const FirSpec* f, int i) {
    double d;
    std::memcpy(&d, &f->bits[i], sizeof d);
    return d;
}

const dsf2flac_uint32 kDefaultReadAhead = 1u << 20;

PinnedPool& g_pinned = pinnedPool();

// Creating a context costs about 3 ms (streams, table images, dither constants, their uploads) -- more than the GPU
// needs for a minute of DSD128.  A process that converts several streams (tracks of one file, a batch of files)
// gets the context of a finished decimator back when geometry, precision and tables are the same.
struct ContextPool {
    static const int kSlots = 4;
    struct Entry { dsdgpu_ctx* ctx = nullptr; int device = 0, nch = 0, nlut = 0, nstep = 0, bits = 0; std::vector<double> tables; };
    std::mutex lock;
    Entry slot[kSlots];
    dsdgpu_ctx* take(int device, int nch, int nlut, int nstep, int bits, const std::vector<double>& tables) {
        std::lock_guard<std::mutex> g(lock);
        for (Entry& e : slot)
            if (e.ctx && e.device == device && e.nch == nch && e.nlut == nlut && e.nstep == nstep && e.bits == bits &&
                e.tables.size() == tables.size() && std::memcmp(e.tables.data(), tables.data(), tables.size() * sizeof(double)) == 0) {
                dsdgpu_ctx* c = e.ctx;
                e.ctx = nullptr;
                return c;
            }
        return nullptr;
    }
    void give(dsdgpu_ctx* ctx, int device, int nch, int nlut, int nstep, int bits, const std::vector<double>& tables) {
        if (!ctx) return;
        dsdgpu_ctx* evicted = ctx;
        {
            std::lock_guard<std::mutex> g(lock);
            for (Entry& e : slot)
                if (!e.ctx) {
                    e.ctx = ctx; e.device = device; e.nch = nch; e.nlut = nlut; e.nstep = nstep; e.bits = bits; e.tables = tables;
                    evicted = nullptr;
                    break;
                }
            if (evicted) {                                   // full: the oldest entry makes room
                evicted = slot[0].ctx;
                for (int i = 0; i + 1 < kSlots; ++i) slot[i] = std::move(slot[i + 1]);
                Entry& e = slot[kSlots - 1];
                e.ctx = ctx; e.device = device; e.nch = nch; e.nlut = nlut; e.nstep = nstep; e.bits = bits; e.tables = tables;
            }
        }
        if (evicted) dsdgpu_destroy(evicted);
    }
    // no destructor, like PinnedPool: the CUDA runtime may be gone at static-destruction time
};
ContextPool g_contexts;

int deviceFromEnv() {
    const char* env = std::getenv("DSDGPU_DEVICE");
    return env ? std::atoi(env) : 0;
}

}  // namespace

DsdDecimator::DsdDecimator(DsdSampleReader *r, dsf2flac_uint32 rate)
    : reader(r), outputSampleRate(rate), nLookupTable(0), tzero(0), ratio(0), nStep(0), valid(true),
      gpu(nullptr), gpuDevice(0), lutBits(64), pos(-1), harvested(0), samplesOut(0), bytes(nullptr), bytesStride(0),
      bytesCap(0), bytesFirst(0), bytesCount(0), block(nullptr), blockCap(0), blockKind(0), blockPos(0),
      blockFrames(0), blockSample0(0), blockScale(0), blockDither(0), blockClip(0),
      readAhead(kDefaultReadAhead), packedTransfer(true), bulk(nullptr), bulkBlock(0), bulkCount(0), bulkStride(0),
      ahead(nullptr), aheadCap(0), aheadBusy(false), aheadKind(0), aheadPos(0), aheadFrames(0), aheadSample0(0),
      aheadScale(0), aheadDither(0), aheadClip(0)
{
    ratio = rate ? r->getSamplingFreq() / rate : 0;
    nStep = ratio / 8;
    const FirSpec* f = findFilter(ratio);
    if (!f) {
        valid = false;
        errorMsg = "Sorry, incompatible sample rate combination";
        return;
    }
    tzero = (dsf2flac_uint32)f->tzero;
    nLookupTable = (dsf2flac_uint32)((f->ntaps + 7) / 8);
    buildTables(r->msbIsPlayedFirst() ? 1 : 0);
    if (nLookupTable > reader->getBufferLength())
        reader->setBufferLength(nLookupTable);
    // logical marker = the reader's marker at hand-over (-1 right after rewind())
    pos = reader->getPosition() / 8;
    harvested = pos + 1;
    bytesFirst = harvested;
    openGpu(lutBits);
}

DsdDecimator::~DsdDecimator()
{
    dropAhead();
    releaseGpu();
    if (bytes) dsdgpu_host_free(bytes);
    g_pinned.give(block, blockCap);
    g_pinned.give(ahead, aheadCap);
}

// One 256-entry table per 8 taps; entries are the signed tap sums in bit order 0..k-1, tap
// 8t+bit pairing with byte bit 7-bit when the reader says msbIsPlayedFirst, else with bit `bit`.
void DsdDecimator::buildTables(int msbFlag)
{
    const FirSpec* f = findFilter(ratio);
    tables.assign((size_t)nLookupTable * 256, 0.0);
    for (dsf2flac_uint32 t = 0; t < nLookupTable; ++t) {
        int k = f->ntaps - (int)t * 8;
        if (k > 8) k = 8;
        for (int seq = 0; seq < 256; ++seq) {
            double acc = 0.0;
            for (int bit = 0; bit < k; ++bit) {
                const int on = msbFlag ? (seq >> (7 - bit)) & 1 : (seq >> bit) & 1;
                const double val = -1 + 2 * (double)on;
                acc += val * tapValue(f, (int)t * 8 + bit);
            }
            tables[(size_t)t * 256 + seq] = acc;
        }
    }
}

namespace {
const int kAheadSlot = DSDGPU_SLOTS - 1;    // the synchronous runs cycle through the slots below it
size_t outElemBytes(int outKind) {
    return outKind == DSDGPU_OUT_INT32 ? sizeof(dsf2flac_int32) : outKind == DSDGPU_OUT_FLOAT64 ? sizeof(double)
         : outKind == DSDGPU_OUT_PCM24 ? 3 : 2;
}
}  // namespace

bool DsdDecimator::openGpu(int bits)
{
    dropAhead();
    releaseGpu();
    gpuDevice = deviceFromEnv();
    gpu = g_contexts.take(gpuDevice, (int)reader->getNumChannels(), (int)nLookupTable, (int)nStep, bits, tables);
    if (gpu) {
        // a recycled context: back to the defaults this class does not set itself
        dsdgpu_set_option(gpu, "source_block_bytes", 0);
        dsdgpu_set_option(gpu, "dither_stride", 0);
    } else {
        const int rc = dsdgpu_create(&gpu, gpuDevice, (int)reader->getNumChannels(), (int)nLookupTable, (int)nStep,
                                     tables.data(), bits);
        if (rc != DSDGPU_OK) {
            fail(std::string("DsdDecimator: GPU path unavailable: ") + dsdgpu_last_error(nullptr));
            gpu = nullptr;
            return false;
        }
    }
    dsdgpu_set_option(gpu, "slots", DSDGPU_SLOTS - 1);
    // a bulk source set before this (re)creation keeps its layout on the new context
    if (bulk && dsdgpu_set_option(gpu, "source_block_bytes", bulkBlock) != DSDGPU_OK) {
        fail(std::string("DsdDecimator: ") + dsdgpu_last_error(gpu));
        dsdgpu_destroy(gpu);
        gpu = nullptr;
        return false;
    }
    lutBits = bits;
    blockFrames = 0;
    return true;
}

// The context goes back to the pool (nothing of this decimator is in flight on it any more).
void DsdDecimator::releaseGpu()
{
    if (!gpu) return;
    g_contexts.give(gpu, gpuDevice, (int)reader->getNumChannels(), (int)nLookupTable, (int)nStep, lutBits, tables);
    gpu = nullptr;
}

// A block in flight is waited for (its buffers must not be freed or reused under it) and forgotten.
void DsdDecimator::dropAhead()
{
    if (aheadBusy && gpu) dsdgpu_wait(gpu, kAheadSlot);
    aheadBusy = false;
}

// Submit the block that follows the current one, asynchronously, with the current block's parameters.
void DsdDecimator::prefetch(int outKind)
{
    if (!bulk || !gpu || aheadBusy || blockFrames <= 0) return;
    const dsf2flac_uint32 nch = reader->getNumChannels();
    const dsf2flac_int64 nextPos = blockPos + blockFrames * (dsf2flac_int64)nStep;
    const dsf2flac_int64 lastDataByte = reader->getLength() / 8 + 1;
    dsf2flac_int64 frames = (lastDataByte + (dsf2flac_int64)nLookupTable - nextPos) / (dsf2flac_int64)nStep + 1;
    if (frames > (dsf2flac_int64)readAhead) frames = readAhead;
    if (frames < 1) return;                               // nothing but idle bytes from there on
    const size_t need = (size_t)frames * nch * outElemBytes(outKind);
    if (need > aheadCap) {
        g_pinned.give(ahead, aheadCap);
        ahead = g_pinned.take(need, aheadCap);
        if (!ahead) return;                               // no read-ahead then; the next refill runs synchronously
    }
    const dsf2flac_uint64 sample0 = blockSample0 + (dsf2flac_uint64)blockFrames * nch;
    if (dsdgpu_submit(gpu, kAheadSlot, bulk, (size_t)bulkStride, 0, bulkCount, reader->getIdleSample(), nextPos, frames,
                      blockScale, blockDither, blockClip, sample0, outKind, ahead) != DSDGPU_OK)
        return;
    aheadBusy = true; aheadKind = outKind; aheadPos = nextPos; aheadFrames = frames; aheadSample0 = sample0;
    aheadScale = blockScale; aheadDither = blockDither; aheadClip = blockClip;
}

void DsdDecimator::fail(const std::string& why)
{
    valid = false;
    errorMsg = why;
}

bool DsdDecimator::isValid() { return valid; }
std::string DsdDecimator::getErrorMsg() { return errorMsg; }
dsf2flac_uint32 DsdDecimator::getOutputSampleRate() { return outputSampleRate; }

dsf2flac_int64 DsdDecimator::getLength() { return reader->getLength() / ratio; }

dsf2flac_float64 DsdDecimator::getPosition()
{
    const dsf2flac_int64 dsdPos = pos * 8;
    return (dsf2flac_float64)(dsdPos - tzero) / ratio;
}

dsf2flac_float64 DsdDecimator::getFirstValidSample()
{
    return (dsf2flac_float64)nLookupTable / nStep - (dsf2flac_float64)tzero / ratio;
}

dsf2flac_float64 DsdDecimator::getLastValidSample()
{
    return (dsf2flac_float64)getLength() - (dsf2flac_float64)tzero / ratio;
}

void DsdDecimator::step() { pos += 1; }

void DsdDecimator::setExtensionFilters(bool on) { g_extensionFilters = on ? 1 : 0; }

void DsdDecimator::setReadAheadFrames(dsf2flac_uint32 frames) { readAhead = frames ? frames : 1; }

void DsdDecimator::setPackedTransfer(bool on) { packedTransfer = on; }

bool DsdDecimator::setLutPrecision(int bits)
{
    if (bits != 64 && bits != 32) return false;
    if (!valid && !gpu) return false;
    return openGpu(bits);
}

bool DsdDecimator::setBulkSource(const dsf2flac_uint8* data, dsf2flac_int64 blockBytes, dsf2flac_int64 deliveredBytes,
                                 dsf2flac_int64 planarStride)
{
    if (!gpu || !data || deliveredBytes < 0) return false;
    if (dsdgpu_set_option(gpu, "source_block_bytes", blockBytes) != DSDGPU_OK) {
        fail(std::string("DsdDecimator: ") + dsdgpu_last_error(gpu));
        return false;
    }
    dropAhead();
    bulk = data; bulkBlock = blockBytes; bulkCount = deliveredBytes; bulkStride = planarStride;
    blockFrames = 0;
    return true;
}

// Pull bytes out of the reader until stream byte `hi` has been seen, keeping [lo, hi] (and
// whatever later bytes were already drained) contiguous in `bytes`.  The reader can only move
// forward one byte per step(), so anything before `lo` that was not drained yet is stepped over.
void DsdDecimator::harvest(dsf2flac_int64 lo, dsf2flac_int64 hi)
{
    const dsf2flac_uint32 nch = reader->getNumChannels();
    if (lo < bytesFirst) lo = bytesFirst;              // older bytes are gone (never needed again)
    dsf2flac_int64 last = bytesFirst + bytesCount;      // one past the newest stored byte
    if (hi < last - 1) hi = last - 1;
    const size_t need = (size_t)(hi - lo + 1);
    if (need > bytesStride) {
        // grow: new pinned block, carry the still-needed tail over
        const size_t stride = (need + need / 4 + 4095) & ~(size_t)4095;
        dsf2flac_uint8* fresh = static_cast<dsf2flac_uint8*>(dsdgpu_host_alloc(stride * nch));
        if (!fresh) { fail("DsdDecimator: pinned host allocation failed"); return; }
        const dsf2flac_int64 keepFrom = lo < last ? lo : last;
        const size_t keep = (size_t)(last - keepFrom);
        for (dsf2flac_uint32 c = 0; c < nch; ++c)
            if (keep) std::memcpy(fresh + c * stride, bytes + c * bytesStride + (keepFrom - bytesFirst), keep);
        if (bytes) dsdgpu_host_free(bytes);
        bytes = fresh; bytesStride = stride; bytesCap = stride * nch;
        bytesFirst = keepFrom; bytesCount = (dsf2flac_int64)keep;
    } else if ((size_t)(hi - bytesFirst + 1) > bytesStride) {
        // slide the still-needed tail to the front
        const dsf2flac_int64 keepFrom = lo < last ? lo : last;
        const size_t keep = (size_t)(last - keepFrom);
        for (dsf2flac_uint32 c = 0; c < nch; ++c)
            if (keep) std::memmove(bytes + c * bytesStride, bytes + c * bytesStride + (keepFrom - bytesFirst), keep);
        bytesFirst = keepFrom; bytesCount = (dsf2flac_int64)keep;
    }
    auto* rings = reader->getBuffer();
    while (harvested <= hi) {
        reader->step();                                  // pushes byte number `harvested` of every channel
        if (harvested >= lo) {
            if (bytesCount == 0) bytesFirst = harvested;
            const size_t k = (size_t)(harvested - bytesFirst);
            for (dsf2flac_uint32 c = 0; c < nch; ++c) bytes[c * bytesStride + k] = rings[c][0];
            bytesCount = (dsf2flac_int64)k + 1;
        }
        ++harvested;
    }
}

bool DsdDecimator::refill(dsf2flac_uint32 wantFrames, dsf2flac_float64 scale, dsf2flac_float64 dither,
                          dsf2flac_float64 clip, int outKind)
{
    if (!gpu) return false;
    const dsf2flac_uint32 nch = reader->getNumChannels();
    if (aheadBusy) {
        // the block submitted while the caller was busy with the last one: take it if it is what is asked for
        const int rc = dsdgpu_wait(gpu, kAheadSlot);
        aheadBusy = false;
        if (rc == DSDGPU_OK && aheadPos == pos && aheadKind == outKind && aheadScale == scale && aheadDither == dither &&
            aheadClip == clip && aheadSample0 == samplesOut && aheadFrames >= 1) {
            std::swap(block, ahead);
            std::swap(blockCap, aheadCap);
            blockKind = outKind; blockPos = pos; blockFrames = aheadFrames; blockSample0 = samplesOut;
            blockScale = scale; blockDither = dither; blockClip = clip;
            prefetch(outKind);
            return true;
        }
    }
    // how far ahead is worth computing: past the end of the data every byte is the idle byte
    dsf2flac_int64 frames = readAhead;
    const dsf2flac_int64 lastDataByte = reader->getLength() / 8 + 1;
    dsf2flac_int64 useful = (lastDataByte + (dsf2flac_int64)nLookupTable - pos) / (dsf2flac_int64)nStep + 1;
    if (useful < 1) useful = 1;
    if (frames > useful) frames = useful;
    if (frames < (dsf2flac_int64)wantFrames) frames = wantFrames;

    if (!bulk) {
        const dsf2flac_int64 lo = pos - ((dsf2flac_int64)nLookupTable - 1);
        const dsf2flac_int64 hi = pos + (frames - 1) * (dsf2flac_int64)nStep;
        harvest(lo, hi);
        if (!valid) return false;
    }

    const size_t needOut = (size_t)frames * nch * outElemBytes(outKind);
    if (needOut > blockCap) {
        g_pinned.give(block, blockCap);
        block = g_pinned.take(needOut, blockCap);
        if (!block) { fail("DsdDecimator: pinned host allocation failed"); return false; }
    }
    const int rc = bulk
        ? dsdgpu_decimate_host(gpu, bulk, (size_t)bulkStride, 0, bulkCount, reader->getIdleSample(), pos, frames, scale,
                               dither, clip, samplesOut, outKind, block)
        : dsdgpu_decimate_host(gpu, bytes, bytesStride, bytesFirst, bytesCount, reader->getIdleSample(), pos, frames,
                               scale, dither, clip, samplesOut, outKind, block);
    if (rc != DSDGPU_OK) {
        fail(std::string("DsdDecimator: GPU decimation failed: ") + dsdgpu_last_error(gpu));
        blockFrames = 0;
        return false;
    }
    blockKind = outKind; blockPos = pos; blockFrames = frames; blockSample0 = samplesOut;
    blockScale = scale; blockDither = dither; blockClip = clip;
    prefetch(outKind);
    return true;
}

template <typename T, typename Conv>
void DsdDecimator::serve(T* buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 dither,
                         dsf2flac_float64 clip, int outKind, Conv conv)
{
    const dsf2flac_uint32 nch = reader->getNumChannels();
    const ldiv_t d = ldiv(bufferLen, nch);
    if (d.rem) {
        fputs("Buffer length is not a multiple of getNumChannels()", stderr);
        exit(EXIT_FAILURE);
    }
    dsf2flac_uint32 frames = (dsf2flac_uint32)d.quot, done = 0;
    while (done < frames) {
        // is the current position inside the cached block, on its frame grid, same parameters,
        // and at the same place in the dither stream?
        dsf2flac_int64 idx = -1;
        if (blockFrames > 0 && blockKind == outKind && blockScale == scale && blockDither == dither &&
            blockClip == clip && pos >= blockPos && (pos - blockPos) % nStep == 0) {
            idx = (pos - blockPos) / nStep;
            if (idx >= blockFrames || blockSample0 + (dsf2flac_uint64)idx * nch != samplesOut) idx = -1;
        }
        if (idx < 0) {
            if (!refill(frames - done, scale, dither, clip, outKind)) {
                // The GPU path failed and there is deliberately no CPU fallback.  main.cpp checks isValid() once,
                // after construction (main.cpp:232-236); handing back made-up samples here would end in a FLAC
                // file with silence in it and exit code 0.  The reference's convention for an unrecoverable error
                // inside this function is message + exit (dsd_decimator.cpp:182-186): do the same.
                fputs(errorMsg.empty() ? "DsdDecimator: GPU decimation failed" : errorMsg.c_str(), stderr);
                fputs("\n", stderr);
                exit(EXIT_FAILURE);
            }
            idx = 0;
        }
        dsf2flac_int64 n = blockFrames - idx;
        if (n > (dsf2flac_int64)(frames - done)) n = frames - done;
        conv(buffer + (size_t)done * nch, block, (size_t)idx * nch, (size_t)n * nch);
        done += (dsf2flac_uint32)n;
        pos += n * (dsf2flac_int64)nStep;
        samplesOut += (dsf2flac_uint64)n * nch;
    }
}

namespace {
template <typename T>
struct RoundFromDouble {          // static_cast<T>(round(sum)), reference dsd_decimator.cpp:213-214
    void operator()(T* dst, const void* blk, size_t first, size_t count) const {
        const double* src = static_cast<const double*>(blk) + first;
        for (size_t i = 0; i < count; ++i) dst[i] = static_cast<T>(std::round(src[i]));
    }
};
template <typename T>
struct CastFromDouble {           // static_cast<T>(sum), reference dsd_decimator.cpp:215-216
    void operator()(T* dst, const void* blk, size_t first, size_t count) const {
        const double* src = static_cast<const double*>(blk) + first;
        for (size_t i = 0; i < count; ++i) dst[i] = static_cast<T>(src[i]);
    }
};
struct CopyInt32 {                // rounded on the GPU already
    void operator()(dsf2flac_int32* dst, const void* blk, size_t first, size_t count) const {
        std::memcpy(dst, static_cast<const dsf2flac_int32*>(blk) + first, count * sizeof(dsf2flac_int32));
    }
};

// The block came over PCIe as packed 24-bit / 16-bit samples (include/dsdgpu.h, DSDGPU_OUT_PCM24 / _PCM16): the copy
// into the caller's FLAC__int32 buffer (main.cpp:173-177) is the widening.  serve() copies every sample once anyway.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("ssse3"))) void widen24_ssse3(dsf2flac_int32* dst, const dsf2flac_uint8* src, size_t count) {
    // 12 source bytes -> 4 samples: each sample's 3 bytes go to the top of a 32-bit lane, an arithmetic shift by 8
    // brings them down sign-extended.  Loads are 16 bytes wide, so stop 4 bytes short of the end.
    const __m128i pick = _mm_setr_epi8(-1, 0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11);
    size_t i = 0;
    for (; i + 4 <= count && 3 * i + 16 <= 3 * count; i += 4) {
        const __m128i raw = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 3 * i));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_srai_epi32(_mm_shuffle_epi8(raw, pick), 8));
    }
    for (; i < count; ++i) {
        const dsf2flac_uint8* p = src + 3 * i;
        dst[i] = (dsf2flac_int32)((dsf2flac_uint32)p[0] << 8 | (dsf2flac_uint32)p[1] << 16 | (dsf2flac_uint32)p[2] << 24) >> 8;
    }
}
#endif
void widen24(dsf2flac_int32* dst, const dsf2flac_uint8* src, size_t count) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool fast = __builtin_cpu_supports("ssse3");
    if (fast) { widen24_ssse3(dst, src, count); return; }
#endif
    for (size_t i = 0; i < count; ++i) {
        const dsf2flac_uint8* p = src + 3 * i;
        dst[i] = (dsf2flac_int32)((dsf2flac_uint32)p[0] << 8 | (dsf2flac_uint32)p[1] << 16 | (dsf2flac_uint32)p[2] << 24) >> 8;
    }
}
struct WidenPcm24 {
    void operator()(dsf2flac_int32* dst, const void* blk, size_t first, size_t count) const {
        const dsf2flac_uint8* src = static_cast<const dsf2flac_uint8*>(blk) + 3 * first;
        widen24(dst, src, count);
#if defined(__x86_64__) && defined(__GNUC__)
        // The next call reads the bytes that follow, cold: page-locked memory the copy engine has just written.  Ask
        // for them now -- the caller is about to spend its time in the encoder (a prefetch never faults, so running
        // past the end of the block is harmless).
        for (size_t off = 3 * count; off < 6 * count; off += 64) __builtin_prefetch(src + off, 0, 3);
#endif
    }
};
struct WidenPcm16 {
    void operator()(dsf2flac_int32* dst, const void* blk, size_t first, size_t count) const {
        const dsf2flac_int16* src = static_cast<const dsf2flac_int16*>(blk) + first;
        for (size_t i = 0; i < count; ++i) dst[i] = src[i];           // vectorised by the compiler (pmovsxwd / unpack + shift)
    }
};
}  // namespace

template<> void DsdDecimator::getSamples(dsf2flac_int16 *buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 tpdfDitherPeakAmplitude, dsf2flac_float64 clipAmplitude)
{
    serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_FLOAT64, RoundFromDouble<dsf2flac_int16>());
}
template<> void DsdDecimator::getSamples(dsf2flac_int32 *buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 tpdfDitherPeakAmplitude, dsf2flac_float64 clipAmplitude)
{
    // main.cpp clips to 2^(bits-1)-1 (:245): the samples then fit 16 / 24 bits and cross PCIe that narrow
    if (packedTransfer && clipAmplitude > 0 && clipAmplitude <= 32767.0)
        serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_PCM16, WidenPcm16());
    else if (packedTransfer && clipAmplitude > 0 && clipAmplitude <= 8388607.0)
        serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_PCM24, WidenPcm24());
    else
        serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_INT32, CopyInt32());
}
template<> void DsdDecimator::getSamples(dsf2flac_int64 *buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 tpdfDitherPeakAmplitude, dsf2flac_float64 clipAmplitude)
{
    serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_FLOAT64, RoundFromDouble<dsf2flac_int64>());
}
template<> void DsdDecimator::getSamples(dsf2flac_float32 *buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 tpdfDitherPeakAmplitude, dsf2flac_float64 clipAmplitude)
{
    serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_FLOAT64, CastFromDouble<dsf2flac_float32>());
}
template<> void DsdDecimator::getSamples(dsf2flac_float64 *buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 tpdfDitherPeakAmplitude, dsf2flac_float64 clipAmplitude)
{
    serve(buffer, bufferLen, scale, tpdfDitherPeakAmplitude, clipAmplitude, DSDGPU_OUT_FLOAT64, CastFromDouble<dsf2flac_float64>());
}
