// TEST INFRASTRUCTURE ONLY. Stand-in for id3lib's <id3/tag.h>: the reference's reader interface
// only names ID3_Tag as a return type (dsd_sample_reader.h:117-119); nothing on the hot path uses
// it.  The real header drags in the libc/libstdc++ headers the reference relies on implicitly.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
class ID3_Tag {
public:
    ID3_Tag() {}
    ID3_Tag(const void*) {}
};
