// dsd_container.cpp -- DSF / DSDIFF headers and "what would the reference's reader deliver".
//
// Header layouts follow the reference parsers: DSF dsf_file_reader.cpp:163-288 (little endian:
// "DSD " 28 B, "fmt " 52 B, "data" 12 B + channel blocks), DSDIFF dsdiff_file_reader.cpp:288-383,
// 470-545, 1021-1039, 1142-1167 (big endian IFF: FRM8{"DSD ", FVER, PROP{"SND ", FS, CHNL, CMPR},
// "DSD " data}; a chunk occupies size + 12 bytes, + 1 if the size is odd).
//
// The delivered-byte rules restate side effects of the reference's block reads.  Its stream
// wrapper reports only badbit (fstream_plus.cpp:84-92), so a read that runs into the end of the
// file "succeeds" but leaves eofbit set, and samplesAvailable() (dsf_file_reader.h:71,
// dsdiff_file_reader.h:122) is false from the next step on: the byte pushed in the step that
// triggered the read is the last one delivered.  tests/test_containers.py pins every case below
// against the reference's own readers (oracle/_ref).
#include "dsd_container.h"

#include <cstdio>
#include <cstring>
#include <sys/types.h>

#include "dsdgpu.h"
#include "pinned_pool.h"

namespace {

uint64_t le64(const uint8_t* p) { uint64_t v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | p[i]; return v; }
uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
uint64_t be64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v = (v << 8) | p[i]; return v; }
uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | (uint32_t)p[3]; }
uint32_t be16(const uint8_t* p) { return (uint32_t)p[0] << 8 | (uint32_t)p[1]; }
bool is(const uint8_t* p, const char* id) { return std::memcmp(p, id, 4) == 0; }

bool fail(DsdContainer& c, const std::string& why) {
    c.valid = false;
    c.error = why;
    return false;
}

// The file, read where the parsers ask: a few header-sized pieces, then the sample bytes in one go.
struct File {
    FILE* f = nullptr;
    uint64_t size = 0;
    ~File() { if (f) std::fclose(f); }
    bool open(const char* path) {
        f = std::fopen(path, "rb");
        if (!f) return false;
        if (fseeko(f, 0, SEEK_END) != 0) return false;
        const off_t n = ftello(f);
        size = n > 0 ? (uint64_t)n : 0;
        return true;
    }
    bool has(uint64_t off, uint64_t n) const { return off <= size && n <= size - off; }
    // n bytes at `off` (n <= 64: a chunk header and its first fields)
    bool peek(uint64_t off, size_t n, uint8_t* dst) const {
        if (!has(off, n)) return false;
        return fseeko(f, (off_t)off, SEEK_SET) == 0 && std::fread(dst, 1, n, f) == n;
    }
    bool read(uint64_t off, size_t n, uint8_t* dst) const {
        if (n == 0) return true;
        return has(off, n) && fseeko(f, (off_t)off, SEEK_SET) == 0 && std::fread(dst, 1, n, f) == n;
    }
};

bool load_dsf(const File& f, DsdContainer& c) {
    uint8_t h[64];
    uint64_t at = 0;
    if (!f.peek(at, 28, h) || !is(h, "DSD ")) return fail(c, "dsfFileReader::readHeaders:DSD ident error");
    at += le64(h + 4);
    if (!f.peek(at, 52, h) || !is(h, "fmt ")) return fail(c, "dsfFileReader::readHeaders:file ident error");
    const uint64_t fmtSz = le64(h + 4);
    c.channels = le32(h + 24);
    c.samplingFreq = le32(h + 28);
    const uint32_t bits = le32(h + 32);
    c.lengthSamples = (int64_t)le64(h + 36);
    const uint32_t block = le32(h + 44);
    at += fmtSz;
    if (!f.peek(at, 12, h) || !is(h, "data")) return fail(c, "dsfFileReader::readHeaders:file ident error");
    c.dataOffset = (int64_t)at + 12;
    // header fields of an untrusted file: a sample count that is negative as int64, or so large that rounding it up
    // to bytes would overflow, would otherwise turn into a negative / wrapped payload size below
    if (c.lengthSamples < 0 || c.lengthSamples > (int64_t)1 << 62) return fail(c, "dsfFileReader::readHeaders:implausible sample count");
    if (bits != 1) return fail(c, "Sorry, only one bit data is supported");          // dsf_file_reader.cpp:56-61
    if (c.channels < 1 || c.channels > 64) return fail(c, "unsupported channel count");
    if (block < 16 || (block & (block - 1)) != 0) return fail(c, "unsupported DSF block size (GPU path needs a power of two >= 16)");
    c.msbIsPlayedFirst = true;                                                         // dsf_file_reader.h:73
    c.blockBytes = block;

    // Stream indices 0 .. ceil(L/8) are asked for (availability is tested before the marker moves,
    // dsf_file_reader.cpp:87,101).  Whole blocks come from the file.  If the byte after the data
    // would need a block the file does not (fully) hold, the reference's read comes back short:
    // the channel buffers keep the previous block except where trailing file bytes landed, that
    // one byte is delivered, then idle.
    const int64_t B = block, nch = c.channels;
    const int64_t want = (c.lengthSamples + 7) / 8 + 1;
    const int64_t fileData = (int64_t)f.size - c.dataOffset;
    const int64_t fullBlocks = fileData > 0 ? fileData / (nch * B) : 0;
    const int64_t needBlocks = (want + B - 1) / B;
    const bool synth = needBlocks > fullBlocks;
    const int64_t haveBlocks = synth ? fullBlocks + 1 : needBlocks;
    if (!c.payload.allocate((size_t)(haveBlocks * nch * B))) return fail(c, "out of (page-locked) memory for the sample data");
    const int64_t copyBlocks = synth ? fullBlocks : needBlocks;
    if (copyBlocks > 0 && !f.read((uint64_t)c.dataOffset, (size_t)(copyBlocks * nch * B), c.payload.data()))
        return fail(c, "dsfFileReader::readNextBlock:File read error");
    if (synth) {
        uint8_t* last = c.payload.data() + fullBlocks * nch * B;
        if (fullBlocks > 0) std::memcpy(last, last - nch * B, (size_t)(nch * B));       // stale buffers
        else std::memset(last, 0, (size_t)(nch * B));                                    // the reader's buffers start out zeroed
        const int64_t rest = fileData > 0 ? fileData - fullBlocks * nch * B : 0;        // short read fills from channel 0 on
        if (rest > 0 && !f.read((uint64_t)(c.dataOffset + fullBlocks * nch * B), (size_t)rest, last))
            return fail(c, "dsfFileReader::readNextBlock:File read error");
        c.deliveredBytes = want < fullBlocks * B + 1 ? want : fullBlocks * B + 1;
    } else {
        c.deliveredBytes = want;
    }
    c.valid = true;
    return true;
}

bool load_dsdiff(const File& f, DsdContainer& c) {
    uint8_t h[32];
    // top level: the first FRM8 chunk (dsdiff_file_reader.cpp:288-312)
    uint64_t at = 0, frmSz = 0;
    for (;;) {
        if (!f.peek(at, 12, h)) return fail(c, "dsdiffFileReader::readChunkHeader:File read error");
        uint64_t sz = be64(h + 4);
        if (sz > (uint64_t)1 << 62) return fail(c, "dsdiffFileReader::readChunkHeader:implausible chunk size");   // size + 13 must not wrap (a chunk of 0 bytes would never advance)
        sz += (sz & 1) ? 13 : 12;
        if (is(h, "FRM8")) { frmSz = sz; break; }
        at += sz;
    }
    const uint64_t frmStart = at;
    if (!f.peek(at, 16, h) || !is(h + 12, "DSD ")) return fail(c, "dsdiffFileReader::readChunk_FRM8:chunk ident error");
    bool fver = false, prop = false, dsd = false, dst = false;
    bool fs = false, chnl = false, cmpr = false;
    char compression[5] = "    ";
    uint64_t dataChunkLen = 0;
    for (uint64_t sub = at + 16; sub < frmStart + frmSz;) {
        if (!f.peek(sub, 12, h)) return fail(c, "dsdiffFileReader::readChunkHeader:File read error");
        uint64_t sz = be64(h + 4);
        if (sz > (uint64_t)1 << 62) return fail(c, "dsdiffFileReader::readChunkHeader:implausible chunk size");
        sz += (sz & 1) ? 13 : 12;
        if (!fver && is(h, "FVER")) {
            fver = f.has(sub + 12, 4);
        } else if (!prop && is(h, "PROP")) {
            uint8_t q[16];
            if (!f.peek(sub + 12, 4, q) || !is(q, "SND ")) return fail(c, "dsdiffFileReader::readChunk_PROP:PROP chunk format error");
            for (uint64_t p = sub + 16; p < sub + sz;) {
                if (!f.peek(p, 12, q)) return fail(c, "dsdiffFileReader::readChunkHeader:File read error");
                uint64_t psz = be64(q + 4);
                if (psz > (uint64_t)1 << 62) return fail(c, "dsdiffFileReader::readChunkHeader:implausible chunk size");
                psz += (psz & 1) ? 13 : 12;
                uint8_t v[4];
                if (!fs && is(q, "FS  ") && f.peek(p + 12, 4, v)) { c.samplingFreq = be32(v); fs = true; }
                else if (!chnl && is(q, "CHNL") && f.peek(p + 12, 2, v)) { c.channels = be16(v); chnl = true; }
                else if (!cmpr && is(q, "CMPR") && f.peek(p + 12, 4, v)) { std::memcpy(compression, v, 4); cmpr = true; }
                p += psz;
            }
            if (!fs) return fail(c, "dsdiffFileReader::readChunk_PROP: no valid FS chunk");
            if (!chnl) return fail(c, "dsdiffFileReader::readChunk_PROP: no valid CHNL chunk");
            if (!cmpr) return fail(c, "dsdiffFileReader::readChunk_PROP: no valid CMPR chunk");
            prop = true;
        } else if (!dsd && !dst && is(h, "DSD ")) {
            c.dataOffset = (int64_t)sub + 12;
            dataChunkLen = sz;
            dsd = true;
        } else if (!dsd && !dst && is(h, "DST ")) {
            dst = true;
        }
        sub += sz;
    }
    if (!fver) return fail(c, "dsdiffFileReader::readChunk_FRM8:No valid FVER chunk found");
    if (!prop) return fail(c, "dsdiffFileReader::readChunk_FRM8:No valid PROP chunk found");
    if (dst || std::memcmp(compression, "DST ", 4) == 0)
        return fail(c, "DST-compressed DSDIFF is decoded on the CPU by the reference (libdstdec); the GPU path reads plain DSD only");
    if (!dsd) return fail(c, "dsdiffFileReader::readChunk_FRM8:No valid DSD or DST chunk found");
    if (std::memcmp(compression, "DSD ", 4) != 0) return fail(c, "unknown DSDIFF compression type");
    if (c.channels < 1 || c.channels > 64) return fail(c, "unsupported channel count");
    const int64_t nch = c.channels;
    // dsdiff_file_reader.cpp:1037 divides the PADDED chunk length (readChunkHeader has already added the pad byte of an
    // odd-sized chunk, :1160-1163), so an odd "DSD " chunk counts its pad byte as a sample byte; reproduced, not fixed
    const int64_t N = (int64_t)((dataChunkLen - 12) / (uint64_t)nch);
    c.lengthSamples = N * 8;
    c.msbIsPlayedFirst = false;                                                        // dsdiff_file_reader.h:121
    c.blockBytes = 1;

    // The reader refills a 1024-byte-per-channel buffer at stream indices k = 1024 m.  For m >= 1
    // it asks for min(N - k + 1, 1024) bytes per channel -- its position is still one byte behind
    // when it computes what is left (dsdiff_file_reader.cpp:167-181, 227-241) -- i.e. one byte per
    // channel MORE than the chunk holds once it reaches the last buffer.  If the file goes on (a
    // chunk follows), those bytes are that chunk's; if the file ends there, the read comes back
    // short, eofbit sticks, and only the byte of that very step is still delivered.
    const int64_t total = (int64_t)f.size - c.dataOffset;                              // bytes from the first sample to EOF
    int64_t delivered = N + 1;
    const int64_t first = (N < 1024 ? N : 1024) * nch;                                  // rewind(): position 0, nothing "behind"
    if (first > total) {
        delivered = 0;
    } else {
        for (int64_t k = 1024; k <= N; k += 1024) {
            const int64_t ask = (N - k + 1 < 1024 ? N - k + 1 : 1024) * nch, have = total - k * nch;
            if (ask > have) { delivered = have >= nch ? k + 1 : k; break; }
        }
    }
    if (delivered > total / nch) delivered = total / nch;
    if (delivered < 0) delivered = 0;
    c.deliveredBytes = delivered;
    if (!c.payload.allocate((size_t)(delivered * nch))) return fail(c, "out of (page-locked) memory for the sample data");
    if (!f.read((uint64_t)c.dataOffset, (size_t)(delivered * nch), c.payload.data()))
        return fail(c, "dsdiffFileReader::readNextBlock:File read error");
    c.valid = true;
    return true;
}

}  // namespace

bool DsdContainer::Payload::allocate(size_t bytes) {
    release();
    if (bytes == 0) return true;
    ptr_ = static_cast<uint8_t*>(pinnedPool().take(bytes, cap_));
    if (!ptr_) { cap_ = 0; return false; }
    size_ = bytes;
    return true;
}

void DsdContainer::Payload::release() {
    if (ptr_) pinnedPool().give(ptr_, cap_);
    ptr_ = nullptr;
    size_ = cap_ = 0;
}

bool dsd_container_load(const char* path, int kind_hint, DsdContainer& c) {
    c = DsdContainer();
    File f;
    if (!f.open(path)) return fail(c, "could not open file");
    if (f.size < 12) return fail(c, "file too short");
    int kind = kind_hint;
    uint8_t magic[4];
    if (!f.peek(0, 4, magic)) return fail(c, "file too short");
    if (kind == 0) kind = is(magic, "DSD ") ? 1 : is(magic, "FRM8") ? 2 : 0;
    if (kind == 1) { c.kind = DsdContainer::DSF; return load_dsf(f, c); }
    if (kind == 2) { c.kind = DsdContainer::DSDIFF; return load_dsdiff(f, c); }
    return fail(c, "not a DSF or DSDIFF file");
}
