// DsdSampleReader over a loaded DsdContainer: the geometry the reference's DsfFileReader /
// DsdiffFileReader report (dsf_file_reader.h:66-75, dsdiff_file_reader.h:117-127) and, for callers
// that still step, the same bytes one at a time.  With DsdDecimator::setBulkSource() nobody steps.
#ifndef CONTAINER_SAMPLE_READER_H
#define CONTAINER_SAMPLE_READER_H

#include "dsd_container.h"
#include "dsd_sample_reader.h"

class ContainerSampleReader : public DsdSampleReader {
public:
    explicit ContainerSampleReader(const DsdContainer& c) : c_(c) {
        valid = c.valid;
        errorMsg = c.error;
        if (valid) { allocateBuffer(); rewind(); }
    }
    dsf2flac_uint32 getSamplingFreq() override { return c_.samplingFreq; }
    dsf2flac_uint32 getNumChannels() override { return c_.channels; }
    dsf2flac_int64 getLength() override { return c_.lengthSamples; }
    bool msbIsPlayedFirst() override { return c_.msbIsPlayedFirst; }
    dsf2flac_uint8 getIdleSample() override { return c_.idle; }
    void rewind() override { posMarker = -1; clearBuffer(); }
    bool step() override {
        const bool have = samplesAvailable();
        const dsf2flac_int64 k = posMarker + 1, nch = c_.channels, B = c_.blockBytes;
        for (dsf2flac_int64 ch = 0; ch < nch; ++ch) {
            dsf2flac_uint8 b = c_.idle;
            if (have && k >= 0 && k < c_.deliveredBytes)
                b = B == 1 ? c_.payload[(size_t)(k * nch + ch)] : c_.payload[(size_t)(((k / B) * nch + ch) * B + k % B)];
            circularBuffers[ch].push_front(b);
        }
        ++posMarker;
        return have;
    }
private:
    const DsdContainer& c_;
};

#endif
