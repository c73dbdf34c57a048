// dsd_decimator.h -- B200 drop-in for the reference's DsdDecimator.
//
// The public section is the reference's operator surface (reference src/dsd_decimator.h:62-127):
// same class name, same member names, same argument meaning, so main.cpp, the DSF/DSDIFF
// readers, the DoP packer, tag conversion and the FLAC encode loop compile and link unchanged
// against this header + dsd_decimator.cpp + libdsdgpu.so.  Everything below
// `private:` is new: the FIR no longer runs in getSamples() on the CPU, it runs on the GPU
// through the C ABI of include/dsdgpu.h, a large block of frames at a time.
#ifndef DSDDECIMATOR_H
#define DSDDECIMATOR_H

#include <string>
#include <vector>

#include "dsd_sample_reader.h"

struct dsdgpu_ctx;

class DsdDecimator
{
public:
    // reference dsd_decimator.h:72 -- reader must be valid; rate must divide the DSD rate by 8, 16 or 32
    DsdDecimator(DsdSampleReader *reader, dsf2flac_uint32 outputSampleRate);
    virtual ~DsdDecimator();

    bool isValid();
    std::string getErrorMsg();

    dsf2flac_uint32 getOutputSampleRate();
    dsf2flac_uint32 getDecimationRatio() { return ratio; }
    dsf2flac_int64 getLength();                                   // PCM samples
    dsf2flac_uint32 getNumChannels() { return reader->getNumChannels(); }
    dsf2flac_float64 getPosition();                               // PCM samples
    dsf2flac_float64 getPositionInSeconds() { return getPosition() / outputSampleRate; }
    dsf2flac_float64 getPositionAsPercent() { return getPosition() / getLength() * 100; }
    dsf2flac_float64 getFirstValidSample();
    dsf2flac_float64 getLastValidSample();
    void step();                                                  // forward by 8 DSD samples

    // bufferLen = frames * getNumChannels(), interleaved; int16/int32/int64 round half away from
    // zero, float32/float64 are not rounded (reference dsd_decimator.h:96-121, .cpp:153-172)
    template <typename sampleType> void getSamples(
            sampleType *buffer,
            dsf2flac_uint32 bufferLen,
            dsf2flac_float64 scale,
            dsf2flac_float64 tpdfDitherPeakAmplitude = 0,
            dsf2flac_float64 clipAmplitude = 0);

    // --- additions (not in the reference) ---------------------------------------------------
    // Decimation by 64, 128 and 256 (DSD512 -> 352.8k, ... -> 88.2k): the reference answers "Sorry,
    // incompatible sample rate combination" for these (dsd_decimator.cpp:64-68) and so does this
    // class unless the extension filters are switched on here (or DSDGPU_EXTENSION_FILTERS=1 in the
    // environment) before construction.  Same operator, longer impulse responses.
    static void setExtensionFilters(bool on);
    // Frames decimated per GPU round trip; larger = fewer, bigger transfers. Default 1<<20.
    void setReadAheadFrames(dsf2flac_uint32 frames);
    // getSamples<int32> with a clip amplitude of at most 2^23-1 / 2^15-1 (what main.cpp passes for 20/24-bit and 16-bit
    // output, :245) fetches its blocks from the GPU as packed 24-bit / 16-bit samples and widens them while copying
    // into the caller's buffer -- same integers, 25 % / 50 % fewer bytes over PCIe.  On by default; false = int32.
    void setPackedTransfer(bool on);
    // 64 (default, bit-exact) or 32 (32-bit fixed-point tables, exact integer sums: within +-1 LSB at
    // 24 bit, about twice as fast). Takes effect immediately.
    bool setLutPrecision(int bits);
    // Bulk source: the stream bytes of all channels in the container's own layout, so that nothing
    // is pulled through reader->step() any more.  `data` holds stream bytes 0 ..
    // deliveredBytes-1 of every channel: blockBytes = 0 planar rows of `planarStride` bytes,
    // blockBytes = 2^k >= 16 DSF block-planar, blockBytes = 1 DSDIFF byte-interleaved; every other
    // stream index reads as the reader's idle byte.  Ideally pinned memory.  The reader still
    // supplies the geometry (channels, rates, length, bit order).  Call before the first getSamples().
    bool setBulkSource(const dsf2flac_uint8* data, dsf2flac_int64 blockBytes, dsf2flac_int64 deliveredBytes,
                       dsf2flac_int64 planarStride = 0);

private:
    void buildTables(int msbFlag);
    bool openGpu(int lutBits);
    void releaseGpu();
    void fail(const std::string& why);
    // make stream bytes [lo, hi] of every channel available in `bytes`
    void harvest(dsf2flac_int64 lo, dsf2flac_int64 hi);
    // run the GPU over `frames` frames starting at logical position pos into `block`
    bool refill(dsf2flac_uint32 wantFrames, dsf2flac_float64 scale, dsf2flac_float64 dither,
                dsf2flac_float64 clip, int outKind);
    // copy `frames` frames at the current position out of `block`, refilling as needed
    template <typename T, typename Conv>
    void serve(T* buffer, dsf2flac_uint32 bufferLen, dsf2flac_float64 scale, dsf2flac_float64 dither,
               dsf2flac_float64 clip, int outKind, Conv conv);

    DsdSampleReader *reader;
    dsf2flac_uint32 outputSampleRate;
    dsf2flac_uint32 nLookupTable;
    dsf2flac_uint32 tzero;
    dsf2flac_uint32 ratio;
    dsf2flac_uint32 nStep;
    bool valid;
    std::string errorMsg;

    std::vector<double> tables;          // [nLookupTable][256], built as the reference builds them
    dsdgpu_ctx* gpu;
    int gpuDevice;
    int lutBits;

    dsf2flac_int64 pos;                  // logical reader marker: newest byte the consumer has reached
    dsf2flac_int64 harvested;            // reader->step() calls made so far (bytes 0..harvested-1 seen)
    dsf2flac_uint64 samplesOut;          // output samples handed out (dither stream position)

    dsf2flac_uint8* bytes;               // pinned, planar [nch][bytesStride]; bytes[c][k] = stream byte bytesFirst+k
    size_t bytesStride, bytesCap;
    dsf2flac_int64 bytesFirst, bytesCount;

    void* block;                         // pinned output of the last GPU run
    size_t blockCap;
    int blockKind;                       // DSDGPU_OUT_*
    dsf2flac_int64 blockPos;             // logical position of frame 0 of the block
    dsf2flac_int64 blockFrames;
    dsf2flac_uint64 blockSample0;        // samplesOut at frame 0
    dsf2flac_float64 blockScale, blockDither, blockClip;
    dsf2flac_uint32 readAhead;
    bool packedTransfer;

    const dsf2flac_uint8* bulk;          // bulk source (nullptr: drain the reader)
    dsf2flac_int64 bulkBlock, bulkCount, bulkStride;

    // Bulk sources only: while the caller consumes `block` (FLAC-encodes it), the block after it is already
    // on its way -- submitted on a slot of its own into `ahead` with the parameters of the last call.
    // refill() takes it if position, parameters and dither position still match, else drops it.
    void prefetch(int outKind);
    void dropAhead();
    void* ahead;
    size_t aheadCap;
    bool aheadBusy;
    int aheadKind;
    dsf2flac_int64 aheadPos, aheadFrames;
    dsf2flac_uint64 aheadSample0;
    dsf2flac_float64 aheadScale, aheadDither, aheadClip;
};

#endif // DSDDECIMATOR_H
