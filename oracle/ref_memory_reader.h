// TEST INFRASTRUCTURE ONLY.  In-memory subclass of the REFERENCE's DsdSampleReader (reference
// src/dsd_sample_reader.h:51-145), shared by oracle/ref_wrap.cpp and oracle/dropin_wrap.cpp.
// step() has the shape of DsfFileReader::step() (reference src/dsf_file_reader.cpp:83-103): availability
// is tested BEFORE the marker moves, so the byte at index L/8 is still fetched when L is a multiple of 8.
#pragma once
#include <cstdint>

#include "dsd_sample_reader.h"

class MemoryReader : public DsdSampleReader {
public:
    MemoryReader(const uint8_t* planar, int nch, int64_t nbytes, int64_t lengthSamples,
                 uint32_t fs, bool msbFirstFlag, uint8_t idle)
        : data_(planar), nch_(nch), nbytes_(nbytes), len_(lengthSamples), fs_(fs),
          msb_(msbFirstFlag), idle_(idle) {
        samplesPerChar = 8;
        valid = true;
        posMarker = -1;
        allocateBuffer();
        rewind();
    }
    dsf2flac_uint32 getSamplingFreq() { return fs_; }
    dsf2flac_uint32 getNumChannels() { return (dsf2flac_uint32)nch_; }
    dsf2flac_int64 getLength() { return len_; }
    bool msbIsPlayedFirst() { return msb_; }
    dsf2flac_uint8 getIdleSample() { return idle_; }
    void rewind() { posMarker = -1; clearBuffer(); }
    bool step() {
        bool ok = samplesAvailable();
        const int64_t k = posMarker + 1;
        for (int c = 0; c < nch_; ++c) {
            uint8_t b = idle_;
            if (ok && k >= 0 && k < nbytes_) b = data_[(int64_t)c * nbytes_ + k];
            circularBuffers[c].push_front(b);
        }
        posMarker++;
        return ok;
    }
private:
    const uint8_t* data_;
    int nch_;
    int64_t nbytes_, len_;
    uint32_t fs_;
    bool msb_;
    uint8_t idle_;
};
