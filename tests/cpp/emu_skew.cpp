// Host-side lane emulator for the skewed kernel (dsf2flac_b200/csrc/skew_core.h).
//
// Runs skew_tile_lane() once per thread id over a byte image of the CTA's shared memory (table
// image + one staged byte tile), for every tile of a small stream, and
//   1. compares every emitted fp64 sum, bit for bit, with the straightforward
//      sum_{t=0..NLUT-1} lut[t][B(p - t)] in reference order (dsd_decimator.cpp:194-198);
//   2. replays the recorded shared-memory addresses warp by warp and counts wavefronts:
//      LDS.64 is served per half-warp (16 lanes x 8 B), LDS.32 per warp (32 lanes x 4 B); a
//      wavefront can serve one 4-byte word per bank (32 banks), equal addresses broadcast.
// Prints one JSON line; exit code 0 iff all sums match and every table load is conflict free.
//
// usage: emu_skew <nthr:512|1024> <nframes> <p0> <in_first> <in_count> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "skew_core.h"

using namespace dsdgpu;

struct Access { int tid, period, step; uint32_t addr; };

struct Recorder {
    std::vector<Access>* lut_log;
    std::vector<Access>* word_log;
    void lut(int tid, int period, int s, uint32_t a) const { if (lut_log) lut_log->push_back({tid, period, s, a}); }
    void word(int tid, int period, int s, uint32_t a) const { if (word_log) word_log->push_back({tid, period, s, a}); }
};

// wavefronts needed by one group of lanes: max over banks of distinct words requested
static int wavefronts(const std::vector<uint32_t>& addrs, int bytes_per_lane) {
    std::map<int, std::vector<uint32_t>> per_bank;
    for (uint32_t a : addrs)
        for (int k = 0; k < bytes_per_lane / 4; ++k) {
            uint32_t w = a / 4 + k;
            auto& v = per_bank[w % 32];
            bool seen = false;
            for (uint32_t x : v) seen |= (x == w);
            if (!seen) v.push_back(w);
        }
    size_t worst = 0;
    for (auto& kv : per_bank) worst = kv.second.size() > worst ? kv.second.size() : worst;
    return (int)worst;
}

template <class G>
static int run(long long nframes, long long p0, long long in_first, long long in_count, unsigned seed) {
    // random tables (any finite doubles do) and random bytes
    srand(seed);
    std::vector<double> lut((size_t)G::NLUT * 256);
    for (auto& v : lut) v = (rand() / (double)RAND_MAX - 0.5) * 0.35;
    std::vector<unsigned char> src((size_t)(in_count > 0 ? in_count : 1));
    for (auto& b : src) b = (unsigned char)(rand() & 0xff);
    const unsigned char fill = 0x69;
    auto stream_byte = [&](long long j) -> unsigned char {
        long long o = j - in_first;
        return (o >= 0 && o < in_count) ? src[(size_t)o] : fill;
    };

    std::vector<unsigned char> smem((size_t)G::LUT_BYTES + G::TILE_BYTES);
    skew_build_lut_image<G>(lut.data(), smem.data());

    long long mismatches = 0, checked = 0;
    long long lut_instr = 0, lut_wavefronts = 0, lut_ideal = 0;
    long long word_instr = 0, word_wavefronts = 0;
    const long long ntiles = (nframes + G::NT - 1) / G::NT;
    for (long long tile = 0; tile < ntiles; ++tile) {
        const long long f0 = tile * G::NT;
        long long t0; int a_tile;
        skew_tile_origin<G>(p0, in_first, f0, t0, a_tile);
        if ((t0 - in_first) % 16 != 0 || a_tile < G::FRONT) { fprintf(stderr, "bad tile origin\n"); return 2; }
        for (int k = 0; k < G::TILE_BYTES; ++k) smem[(size_t)G::LUT_BYTES + k] = stream_byte(t0 + k);

        std::vector<Access> lut_log, word_log;
        const bool trace = (tile == 0 || tile == ntiles - 1);
        Recorder rec{trace ? &lut_log : nullptr, trace ? &word_log : nullptr};
        for (int tid = 0; tid < G::NTHR; ++tid) {
            skew_tile_lane<G>(smem.data(), (uint32_t)G::LUT_BYTES, tid, a_tile,
                              [&](int k, double sum) {
                                  const long long f = f0 + (long long)k * G::NTHR + tid;
                                  if (f >= nframes) return;
                                  const long long p = p0 + f * G::NSTEP;
                                  double ref = 0.0;
                                  for (int t = 0; t < G::NLUT; ++t) ref += lut[(size_t)t * 256 + stream_byte(p - t)];
                                  ++checked;
                                  if (memcmp(&ref, &sum, 8) != 0) {
                                      if (mismatches < 5) fprintf(stderr, "mismatch f=%lld tid=%d k=%d got %.17g want %.17g\n", f, tid, k, sum, ref);
                                      ++mismatches;
                                  }
                              },
                              rec);
        }
        if (trace) {
            // group by (warp, period, step[, issue order]) -> lanes issue the same instruction
            std::map<long long, std::vector<uint32_t>> groups;
            for (auto& a : lut_log) {
                long long key = (((long long)(a.tid / 16) * 64 + a.period) * 128 + a.step);
                groups[key].push_back(a.addr);
            }
            for (auto& kv : groups) {
                ++lut_instr;
                lut_wavefronts += wavefronts(kv.second, 8);
                lut_ideal += 1;
            }
            std::map<long long, std::vector<std::vector<uint32_t>>> wgroups;  // two loads can share a step
            std::map<long long, int> seen_count;
            std::map<long long, std::vector<uint32_t>> wg;
            std::map<long long, int> ordinal;
            for (auto& a : word_log) {
                long long lane_key = (((long long)a.tid * 64 + a.period) * 128 + a.step);
                int ord = ordinal[lane_key]++;
                long long key = ((((long long)(a.tid / 32) * 64 + a.period) * 128 + a.step) * 4 + ord);
                wg[key].push_back(a.addr);
            }
            for (auto& kv : wg) {
                ++word_instr;
                word_wavefronts += wavefronts(kv.second, 4);
            }
        }
    }
    printf("{\"nthr\": %d, \"kr\": %d, \"frames\": %lld, \"checked\": %lld, \"mismatches\": %lld, "
           "\"lut_halfwarp_loads\": %lld, \"lut_wavefronts\": %lld, \"word_loads\": %lld, \"word_wavefronts\": %lld}\n",
           G::NTHR, G::KR, nframes, checked, mismatches, lut_instr, lut_wavefronts, word_instr, word_wavefronts);
    if (checked != nframes) return 3;
    if (mismatches) return 1;
    if (lut_wavefronts != lut_ideal) return 4;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: emu_skew nthr nframes p0 in_first in_count seed\n"); return 64; }
    const int nthr = atoi(argv[1]);
    const long long nframes = atoll(argv[2]), p0 = atoll(argv[3]), in_first = atoll(argv[4]), in_count = atoll(argv[5]);
    const unsigned seed = (unsigned)atoi(argv[6]);
    if (nthr == 512) return run<SkewGeom<72, 4, 512, 12>>(nframes, p0, in_first, in_count, seed);
    if (nthr == 1024) return run<SkewGeom<72, 4, 1024, 6>>(nframes, p0, in_first, in_count, seed);
    fprintf(stderr, "nthr must be 512 or 1024\n");
    return 64;
}
