// TEST INFRASTRUCTURE ONLY. Stand-in for id3lib's <id3/tag.h>.  The reference's reader interface
// names ID3_Tag as a return type (dsd_sample_reader.h:117-119) and the DSF / DSDIFF file readers
// parse or synthesise tags (dsf_file_reader.cpp:303-345, dsdiff_file_reader.cpp:258-285, 588-630);
// none of it touches the sample path, so every tag operation here is a no-op and IsV2Tag says
// "no tag".  The real header also drags in the libc/libstdc++ headers the reference relies on
// implicitly.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#define ID3_TAGHEADERSIZE 10
enum ID3_FrameID { ID3FID_LEADARTIST, ID3FID_ALBUM, ID3FID_TITLE, ID3FID_TRACKNUM };
enum ID3_FieldID { ID3FN_TEXT };
class ID3_Field {
public:
    size_t Set(const char*) { return 0; }
    size_t Set(const unsigned char*, size_t) { return 0; }
};
class ID3_Frame {
public:
    ID3_Frame() {}
    ID3_Frame(ID3_FrameID) {}
    ID3_Field& Field(ID3_FieldID) { return f; }
private:
    ID3_Field f;
};
class ID3_Tag {
public:
    ID3_Tag() {}
    ID3_Tag(const void*) {}
    static int IsV2Tag(const unsigned char*) { return 0; }
    size_t Parse(const unsigned char*, size_t) { return 0; }
    void AttachFrame(ID3_Frame* fr) { delete fr; }
    void AddFrame(const ID3_Frame*) {}
    void AddFrame(const ID3_Frame&) {}
};
