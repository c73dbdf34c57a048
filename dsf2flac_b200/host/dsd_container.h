// dsd_container.h -- the two containers dsf2flac reads, as far as the decimation path needs them.
//
// The reference pulls DSD bytes through DsfFileReader / DsdiffFileReader one byte per channel per
// virtual step() (dsf_file_reader.cpp:83-103, dsdiff_file_reader.cpp:223-243) and re-packs them
// into per-channel rings.  Here the payload of the file is handed to the GPU where it lies:
// parse the headers, say where the sample bytes are and how they are laid out, and say exactly
// which stream bytes the reference's reader would have delivered (everything else is the idle
// byte 0x69) -- including what it delivers around the end of the data, which the reference
// decides by side effects of its block reads (see dsd_container.cpp).
#ifndef DSD_CONTAINER_H
#define DSD_CONTAINER_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct DsdContainer {
    enum Kind { NONE = 0, DSF = 1, DSDIFF = 2 };
    Kind kind = NONE;
    bool valid = false;
    std::string error;

    uint32_t channels = 0;
    uint32_t samplingFreq = 0;        // DSD samples per second
    int64_t lengthSamples = 0;        // reader->getLength(): DSD samples per channel
    bool msbIsPlayedFirst = false;    // reader->msbIsPlayedFirst(): DSF true, DSDIFF false
    uint8_t idle = 0x69;              // reader->getIdleSample()

    int64_t dataOffset = 0;           // file offset of the first sample byte
    int64_t blockBytes = 0;           // dsdgpu "source_block_bytes": DSF block size, 1 for DSDIFF
    int64_t deliveredBytes = 0;       // stream bytes per channel the reference reader hands out before
                                      // it turns to idle bytes (the operator's in_count)
    // the sample bytes in their file layout, ready for dsdgpu_decimate_host: payload.size() >=
    // what deliveredBytes needs; for DSF it may end with one block the file does not contain (the
    // reference's stale-buffer read past the last block, reproduced).  Read from the file straight into
    // page-locked memory (dsdgpu_host_alloc): one copy between the page cache and the upload, no staging vector.
    struct Payload {
        Payload() = default;
        Payload(const Payload&) = delete;
        Payload& operator=(const Payload&) = delete;
        Payload(Payload&& o) noexcept : ptr_(o.ptr_), size_(o.size_), cap_(o.cap_) { o.ptr_ = nullptr; o.size_ = o.cap_ = 0; }
        Payload& operator=(Payload&& o) noexcept { release(); ptr_ = o.ptr_; size_ = o.size_; cap_ = o.cap_; o.ptr_ = nullptr; o.size_ = o.cap_ = 0; return *this; }
        ~Payload() { release(); }
        bool allocate(size_t bytes);          // page-locked (recycled through pinned_pool.h), NOT cleared; false if it fails
        void release();
        uint8_t* data() { return ptr_; }
        const uint8_t* data() const { return ptr_; }
        size_t size() const { return size_; }
        bool empty() const { return size_ == 0; }
        uint8_t operator[](size_t i) const { return ptr_[i]; }
    private:
        uint8_t* ptr_ = nullptr;
        size_t size_ = 0, cap_ = 0;
    } payload;

    DsdContainer() = default;
    DsdContainer(DsdContainer&&) = default;
    DsdContainer& operator=(DsdContainer&&) = default;
};

// Parses `path` (DSF if it starts with "DSD ", DSDIFF if it starts with "FRM8") and loads the sample
// bytes.  kind_hint: 0 = by magic, 1 = DSF, 2 = DSDIFF.  Returns c.valid.
bool dsd_container_load(const char* path, int kind_hint, DsdContainer& c);

#endif  // DSD_CONTAINER_H
