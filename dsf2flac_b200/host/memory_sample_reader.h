// In-memory DsdSampleReader used by the tests, the bench and the smoke run: planar packed DSD
// bytes [nch][nbytes].  step() has the shape of the reference's DsfFileReader::step()
// (dsf_file_reader.cpp:83-103): availability is tested before the marker moves, data is pushed
// while available (and present), the idle byte otherwise.
#ifndef MEMORY_SAMPLE_READER_H
#define MEMORY_SAMPLE_READER_H

#include "dsd_sample_reader.h"

class MemorySampleReader : public DsdSampleReader {
public:
    MemorySampleReader(const dsf2flac_uint8* planar, dsf2flac_uint32 nch, dsf2flac_int64 nbytes,
                       dsf2flac_int64 lengthSamples, dsf2flac_uint32 fs, bool msbFlag,
                       dsf2flac_uint8 idle = 0x69)
        : planar_(planar), nch_(nch), nbytes_(nbytes), length_(lengthSamples), fs_(fs), msb_(msbFlag), idle_(idle) {
        allocateBuffer();
        rewind();
    }
    dsf2flac_uint32 getSamplingFreq() override { return fs_; }
    dsf2flac_uint32 getNumChannels() override { return nch_; }
    dsf2flac_int64 getLength() override { return length_; }
    bool msbIsPlayedFirst() override { return msb_; }
    dsf2flac_uint8 getIdleSample() override { return idle_; }
    void rewind() override { posMarker = -1; clearBuffer(); }
    bool step() override {
        const bool have = samplesAvailable();
        const dsf2flac_int64 k = posMarker + 1;
        for (dsf2flac_uint32 c = 0; c < nch_; ++c) {
            dsf2flac_uint8 b = idle_;
            if (have && k >= 0 && k < nbytes_) b = planar_[(dsf2flac_int64)c * nbytes_ + k];
            circularBuffers[c].push_front(b);
        }
        ++posMarker;
        return have;
    }
private:
    const dsf2flac_uint8* planar_;
    dsf2flac_uint32 nch_;
    dsf2flac_int64 nbytes_, length_;
    dsf2flac_uint32 fs_;
    bool msb_;
    dsf2flac_uint8 idle_;
};

#endif
