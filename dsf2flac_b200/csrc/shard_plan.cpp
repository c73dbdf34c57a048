// shard_plan.cpp -- dsdgpu_plan_shards: cut a batch of DSD streams into per-GPU shards
// (file / channel / time segment).  Host arithmetic only; see include/dsdgpu.h.
#include <cstdint>
#include <vector>

#include "dsdgpu.h"

namespace {

struct Lane {              // one channel of one job on the line of work
    int32_t job, ch;
    int64_t frames;
    long double cost;      // per frame
    long double start;     // position of its frame 0 on the line
};

// frame index inside `l` of line position x, rounded to the alignment and clamped
int64_t frame_at(const Lane& l, long double x, int64_t align) {
    if (x <= l.start) return 0;
    long double f = (x - l.start) / l.cost;
    if (f >= (long double)l.frames) return l.frames;
    int64_t k = (int64_t)(f / align + 0.5L) * align;
    return k > l.frames ? l.frames : k;
}

}  // namespace

extern "C" int64_t dsdgpu_plan_shards(const int64_t* job_frames, const int32_t* job_nch, const int32_t* job_cost,
                                      int njobs, int rank, int world, int64_t align_frames, dsdgpu_shard* out,
                                      int64_t cap) {
    if (njobs < 0 || world < 1 || rank < 0 || rank >= world || align_frames < 1) return -1;
    if (njobs > 0 && (!job_frames || !job_nch)) return -1;
    std::vector<Lane> lanes;
    long double total = 0;
    for (int k = 0; k < njobs; ++k) {
        if (job_frames[k] < 0 || job_nch[k] < 1) return -1;
        const long double cost = job_cost ? (long double)(job_cost[k] > 0 ? job_cost[k] : 1) : 1.0L;
        for (int c = 0; c < job_nch[k]; ++c) {
            lanes.push_back({k, c, job_frames[k], cost, total});
            total += cost * (long double)job_frames[k];
        }
    }
    // cut points are functions of (lane, boundary index) only, so neighbouring ranks agree on them
    const long double lo = total * rank / world, hi = total * (rank + 1) / world;
    std::vector<dsdgpu_shard> mine;
    for (const Lane& l : lanes) {
        const int64_t a = rank == 0 ? 0 : frame_at(l, lo, align_frames);
        const int64_t b = rank == world - 1 ? l.frames : frame_at(l, hi, align_frames);
        if (b <= a) continue;
        if (!mine.empty()) {
            dsdgpu_shard& p = mine.back();
            if (p.job == l.job && p.ch_first + p.ch_count == l.ch && p.frame_first == a && p.frame_count == b - a) {
                p.ch_count += 1;
                continue;
            }
        }
        mine.push_back({l.job, l.ch, 1, 0, a, b - a});
    }
    for (size_t i = 0; i < mine.size() && (int64_t)i < cap; ++i) out[i] = mine[i];
    return (int64_t)mine.size();
}
