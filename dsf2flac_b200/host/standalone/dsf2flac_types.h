// Fixed-width aliases under the names the reference uses (its src/dsf2flac_types.h), so the
// drop-in decimator compiles unchanged against either header.
#ifndef DSF2FLAC_TYPES_H
#define DSF2FLAC_TYPES_H
#include <cstdint>
using dsf2flac_int8 = char;
using dsf2flac_int16 = std::int16_t;
using dsf2flac_int32 = std::int32_t;
using dsf2flac_int64 = std::int64_t;
using dsf2flac_uint8 = std::uint8_t;
using dsf2flac_uint16 = std::uint16_t;
using dsf2flac_uint32 = std::uint32_t;
using dsf2flac_uint64 = std::uint64_t;
using dsf2flac_float32 = float;
using dsf2flac_float64 = double;
#endif
