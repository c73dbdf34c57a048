// dsd_sample_reader.h -- stand-alone mirror of the reference's reader interface
// (reference src/dsd_sample_reader.h:51-145), for building and testing the drop-in decimator
// in this repo, where Boost and id3lib are absent.  Inside dsf2flac itself the reference's own
// header is used instead: dsd_decimator.{h,cpp} only touch the members declared in both.
//
// Contract kept (reference file:line):
//   * one byte ring per channel, index 0 = newest byte, filled by push_front (h:100-106)
//   * rings start out full of the idle byte 0x69                       (h:122, cpp:119-130)
//   * getPosition() = posMarker * samplesPerChar, in DSD samples      (h:89)
//   * setBufferLength() re-sizes the rings, refills them and rewinds   (cpp:63-77)
#ifndef DSD_SAMPLE_READER_MIRROR_H
#define DSD_SAMPLE_READER_MIRROR_H

#include <cstddef>
#include <string>
#include <vector>

#include "dsf2flac_types.h"

// Newest-first byte ring with the three operations the hot path needs.
class DsdByteRing {
public:
    explicit DsdByteRing(std::size_t capacity = 0) { set_capacity(capacity); }
    void set_capacity(std::size_t capacity) { store_.assign(capacity, 0); newest_ = 0; }
    std::size_t capacity() const { return store_.size(); }
    void push_front(dsf2flac_uint8 v) {
        if (store_.empty()) return;
        newest_ = (newest_ == 0 ? store_.size() : newest_) - 1;
        store_[newest_] = v;
    }
    dsf2flac_uint8 operator[](std::size_t age) const {
        std::size_t i = newest_ + age;
        if (i >= store_.size()) i -= store_.size();
        return store_[i];
    }
private:
    std::vector<dsf2flac_uint8> store_;
    std::size_t newest_ = 0;
};

static const dsf2flac_uint32 defaultBufferLength = 5000;

class DsdSampleReader {
public:
    DsdSampleReader() : circularBuffers(nullptr), posMarker(-1), samplesPerChar(8), valid(true),
                        bufferLength(defaultBufferLength), buffersReady(false) {}
    virtual ~DsdSampleReader() { delete[] circularBuffers; }

    bool isValid() { return valid; }
    std::string getErrorMsg() { return errorMsg; }

    virtual dsf2flac_uint32 getSamplingFreq() = 0;
    virtual dsf2flac_uint32 getNumChannels() = 0;
    virtual dsf2flac_int64 getLength() = 0;                       // DSD samples per channel
    dsf2flac_float64 getLengthInSeconds() { return getLength() / (dsf2flac_float64)getSamplingFreq(); }
    virtual dsf2flac_uint32 getNumTracks() { return 1; }
    virtual dsf2flac_uint64 getTrackStart(dsf2flac_uint32) { return 0; }
    virtual dsf2flac_uint64 getTrackEnd(dsf2flac_uint32) { return getLength(); }

    virtual bool step() = 0;                                      // advance 8 DSD samples
    virtual bool samplesAvailable() { return getPosition() < getLength(); }
    dsf2flac_int64 getPosition() { return posMarker * samplesPerChar; }
    dsf2flac_float64 getPositionInSeconds() { return getPosition() / (dsf2flac_float64)getSamplingFreq(); }
    dsf2flac_float64 getPositionAsPercent() { return 100 * getPosition() / (dsf2flac_float64)getLength(); }
    virtual void rewind() = 0;

    DsdByteRing* getBuffer() { return circularBuffers; }
    dsf2flac_uint32 getBufferLength() { return bufferLength; }
    bool setBufferLength(dsf2flac_uint32 n) {
        if (n < 1) { errorMsg = "dsdSampleReader::setBufferLength:buffer length must be >0"; return false; }
        bufferLength = n;
        if (!buffersReady) allocateBuffer();
        for (dsf2flac_uint32 c = 0; c < getNumChannels(); ++c) circularBuffers[c].set_capacity(bufferLength);
        clearBuffer();
        rewind();
        return true;
    }
    virtual bool msbIsPlayedFirst() = 0;
    virtual dsf2flac_uint8 getIdleSample() { return 0x69; }

protected:
    void allocateBuffer() {
        if (buffersReady) return;
        circularBuffers = new DsdByteRing[getNumChannels()];
        for (dsf2flac_uint32 c = 0; c < getNumChannels(); ++c) circularBuffers[c].set_capacity(bufferLength);
        buffersReady = true;
        clearBuffer();
    }
    void clearBuffer() {
        if (!buffersReady) { allocateBuffer(); return; }
        const dsf2flac_uint8 idle = getIdleSample();
        for (dsf2flac_uint32 c = 0; c < getNumChannels(); ++c)
            for (dsf2flac_uint32 k = 0; k < bufferLength; ++k) circularBuffers[c].push_front(idle);
    }

    DsdByteRing* circularBuffers;
    dsf2flac_int64 posMarker;        // implementors bump this in step()
    dsf2flac_uint32 samplesPerChar;  // 8
    bool valid;
    std::string errorMsg;

private:
    dsf2flac_uint32 bufferLength;
    bool buffersReady;
};

#endif
