// batch_run.cpp -- dsdgpu_run_batch: a batch of DSD streams on the GPUs of one process, from C.
//
// The reference converts one file at a time on one thread; its only batch notion is the serial shell loop of
// scripts/batch_dsf2flac.sh:16-19.  Here a caller describes every stream of the batch as a dsdgpu_job (what
// DsdDecimator::getSamples would be asked for the whole stream, reference src/dsd_decimator.cpp:173-222 as driven by
// main.cpp:169-191) and gets all of them converted on `ndevices` GPUs, without Python and without a process per GPU:
//   * dsdgpu_plan_shards cuts the batch into world * ndevices equal-cost pieces (by file, by channel, by time segment
//     with an (nlut-1)-byte input halo); this process runs pieces rank*ndevices .. rank*ndevices + ndevices-1,
//   * one host thread per device; each thread turns its shards into chunks of the context's "chunk_frames" and keeps up
//     to DSDGPU_SLOTS of them in flight (upload, kernel and download of neighbouring chunks overlap ACROSS shards and
//     files -- a batch of short files never drains the pipeline between files),
//   * contexts (tables on the device) are created once per (tables, geometry, channel count, precision, layout).
// Shards are independent (a sample depends on nlut bytes of its own channel, dsd_decimator.cpp:193-198): there is no
// data-path collective.  Only the public entry points of include/dsdgpu.h are used, so the same file runs against the
// CPU stand-in of the ABI in the host-logic tests.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dsdgpu.h"

namespace {

size_t elem_bytes(int out_kind) {
    return out_kind == DSDGPU_OUT_INT32 ? 4 : out_kind == DSDGPU_OUT_FLOAT64 ? 8 : out_kind == DSDGPU_OUT_PCM24 ? 3 : 2;
}

struct CtxKey {
    int device, nlut, nstep, nch, precision; long long source_block;
    std::vector<double> tables;                      // compared by content: a caller's pointer says nothing across calls
    bool same(int dev, const dsdgpu_job& job, int k) const {
        return device == dev && nlut == job.nlut && nstep == job.nstep && nch == k && precision == job.lut_precision &&
               source_block == job.source_block_bytes &&
               std::memcmp(tables.data(), job.lut, tables.size() * sizeof(double)) == 0;
    }
};

// Contexts outlive a call: creating one costs 3-5 ms (streams, table images, dither constants), a batch call on a warm
// process should not pay that per call.  A bounded, process-wide pool; a context is held by one worker at a time.
struct PoolEntry { CtxKey key; dsdgpu_ctx* ctx; bool in_use; };
std::mutex g_pool_lock;
std::vector<PoolEntry>& pool() {
    static std::vector<PoolEntry>* p = new std::vector<PoolEntry>();   // never destroyed: the CUDA runtime may go first
    return *p;
}
const size_t kPoolMax = 24;

struct Worker {
    int device = 0;
    // contexts this worker holds, with what identifies them WITHIN one call (the caller's tables pointer stands for its content)
    struct Held { const double* lut; int nlut, nstep, k, precision; long long source_block; dsdgpu_ctx* ctx; };
    std::vector<Held> held;
    std::string error;

    ~Worker() {
        std::lock_guard<std::mutex> g(g_pool_lock);
        for (auto& h : held)
            for (PoolEntry& e : pool())
                if (e.ctx == h.ctx) e.in_use = false;
    }

    dsdgpu_ctx* ctx_for(const dsdgpu_job& job, int k) {
        for (const Held& h : held)
            if (h.lut == job.lut && h.nlut == job.nlut && h.nstep == job.nstep && h.k == k && h.precision == job.lut_precision &&
                h.source_block == job.source_block_bytes)
                return h.ctx;
        dsdgpu_ctx* ctx = nullptr;
        {
            std::lock_guard<std::mutex> g(g_pool_lock);
            for (PoolEntry& e : pool())
                if (!e.in_use && e.key.same(device, job, k)) { e.in_use = true; ctx = e.ctx; break; }
        }
        if (!ctx) {
            if (dsdgpu_create(&ctx, device, k, job.nlut, job.nstep, job.lut, job.lut_precision) != DSDGPU_OK) {
                error = std::string("dsdgpu_create: ") + dsdgpu_last_error(nullptr);
                return nullptr;
            }
            if (dsdgpu_set_option(ctx, "source_block_bytes", job.source_block_bytes) != DSDGPU_OK) {
                error = std::string("source_block_bytes: ") + dsdgpu_last_error(ctx);
                dsdgpu_destroy(ctx);
                return nullptr;
            }
            CtxKey key{device, job.nlut, job.nstep, k, job.lut_precision, job.source_block_bytes,
                       std::vector<double>(job.lut, job.lut + (size_t)job.nlut * 256)};
            dsdgpu_ctx* evicted = nullptr;
            {
                std::lock_guard<std::mutex> g(g_pool_lock);
                if (pool().size() >= kPoolMax)
                    for (size_t i = 0; i < pool().size(); ++i)
                        if (!pool()[i].in_use) { evicted = pool()[i].ctx; pool().erase(pool().begin() + (long)i); break; }
                pool().push_back(PoolEntry{std::move(key), ctx, true});
            }
            if (evicted) dsdgpu_destroy(evicted);
        }
        held.push_back(Held{job.lut, job.nlut, job.nstep, k, job.lut_precision, job.source_block_bytes, ctx});
        return ctx;
    }
};

// a channel-subset shard of an interleaved output: the chunk lands dense in `tmp`, then goes to its columns
struct Scatter { void* tmp; unsigned char* dst; int64_t frames; int k, nch, c0; size_t elem; };

struct InFlight { dsdgpu_ctx* ctx; int slot; bool scatter; Scatter sc; };

void scatter_columns(const Scatter& s) {
    const unsigned char* src = static_cast<const unsigned char*>(s.tmp);
    const size_t row = (size_t)s.k * s.elem;
    for (int64_t f = 0; f < s.frames; ++f)
        std::memcpy(s.dst + ((size_t)f * s.nch + s.c0) * s.elem, src + (size_t)f * row, row);
}

void run_piece(Worker& w, const dsdgpu_job* jobs, int njobs, int piece, int pieces, int64_t align_frames) {
    std::vector<int64_t> frames(njobs);
    std::vector<int32_t> nch(njobs), cost(njobs);
    for (int j = 0; j < njobs; ++j) { frames[j] = jobs[j].nframes; nch[j] = jobs[j].nch; cost[j] = jobs[j].nlut; }
    int64_t cap = 0;
    for (int j = 0; j < njobs; ++j) cap += jobs[j].nch;
    std::vector<dsdgpu_shard> shards((size_t)cap + 4);
    const int64_t n = dsdgpu_plan_shards(frames.data(), nch.data(), cost.data(), njobs, piece, pieces, align_frames, shards.data(),
                                         (int64_t)shards.size());
    if (n < 0 || n > (int64_t)shards.size()) { w.error = "dsdgpu_plan_shards failed"; return; }

    const long long chunk = 1LL << 20;
    std::deque<InFlight> flying;                              // oldest first; one entry per busy slot
    std::vector<std::pair<dsdgpu_ctx*, int>> next_slot;       // round-robin slot per context
    std::vector<void*> temps;                                 // pinned scratch of channel-subset shards
    auto retire = [&]() -> bool {
        InFlight f = flying.front();
        flying.pop_front();
        if (dsdgpu_wait(f.ctx, f.slot) != DSDGPU_OK) { w.error = std::string("dsdgpu_wait: ") + dsdgpu_last_error(f.ctx); return false; }
        if (f.scatter) scatter_columns(f.sc);
        return true;
    };
    for (int64_t s = 0; s < n && w.error.empty(); ++s) {
        const dsdgpu_shard& sh = shards[(size_t)s];
        const dsdgpu_job& job = jobs[sh.job];
        const bool subset = sh.ch_count != job.nch;
        if (subset && job.source_block_bytes != 0) {
            w.error = "a channel-subset shard needs planar source rows (source_block_bytes = 0)";
            break;
        }
        dsdgpu_ctx* ctx = w.ctx_for(job, sh.ch_count);
        if (!ctx) break;
        dsdgpu_set_option(ctx, "dither_stride", job.nch);
        const size_t elem = elem_bytes(job.out_kind);
        const uint8_t* in = job.in + (job.source_block_bytes == 0 ? (size_t)sh.ch_first * job.in_stride : 0);
        int* slot_of = nullptr;
        for (auto& p : next_slot)
            if (p.first == ctx) slot_of = &p.second;
        if (!slot_of) { next_slot.emplace_back(ctx, 0); slot_of = &next_slot.back().second; }
        void* tmp = nullptr;
        if (subset) {
            tmp = dsdgpu_host_alloc((size_t)sh.frame_count * sh.ch_count * elem);
            if (!tmp) { w.error = "out of page-locked memory"; break; }
            temps.push_back(tmp);
        }
        for (int64_t f = 0; f < sh.frame_count; f += chunk) {
            const int64_t cnt = sh.frame_count - f < chunk ? sh.frame_count - f : chunk;
            const int64_t f0 = sh.frame_first + f;
            // the slot this context uses next must be free: retire in-flight chunks, oldest first, until it is
            const int slot = *slot_of;
            *slot_of = (slot + 1) % DSDGPU_SLOTS;
            for (;;) {
                bool busy = false;
                for (const InFlight& x : flying) busy = busy || (x.ctx == ctx && x.slot == slot);
                if (!busy && flying.size() < (size_t)DSDGPU_SLOTS) break;
                if (!retire()) break;
            }
            if (!w.error.empty()) break;
            unsigned char* dense = static_cast<unsigned char*>(job.out) + (size_t)f0 * job.nch * elem;
            void* out = subset ? static_cast<void*>(static_cast<unsigned char*>(tmp) + (size_t)f * sh.ch_count * elem) : static_cast<void*>(dense);
            const int rc = dsdgpu_submit(ctx, slot, in, job.in_stride, job.in_first, job.in_count, job.fill, job.p0 + f0 * job.nstep, cnt,
                                         job.scale, job.dither_amp, job.clip,
                                         job.dither_first_sample + (uint64_t)f0 * (uint64_t)job.nch + (uint64_t)sh.ch_first, job.out_kind, out);
            if (rc != DSDGPU_OK) { w.error = std::string("dsdgpu_submit: ") + dsdgpu_last_error(ctx); break; }
            InFlight fl{ctx, slot, subset, Scatter{out, dense, cnt, sh.ch_count, job.nch, sh.ch_first, elem}};
            flying.push_back(fl);
        }
    }
    while (!flying.empty()) {
        if (!retire() && flying.empty()) break;
    }
    for (void* t : temps) dsdgpu_host_free(t);
}

}  // namespace

extern "C" int dsdgpu_run_batch(const dsdgpu_job* jobs, int njobs, const int* devices, int ndevices, int rank, int world,
                                int64_t align_frames, char* err, size_t err_cap) {
    auto fail = [&](const std::string& why, int code) {
        if (err && err_cap) std::snprintf(err, err_cap, "%s", why.c_str());
        return code;
    };
    if (err && err_cap) err[0] = 0;
    if (njobs < 0 || (njobs > 0 && !jobs) || ndevices < 1 || !devices || world < 1 || rank < 0 || rank >= world || align_frames < 1)
        return fail("dsdgpu_run_batch: bad arguments", DSDGPU_ERR_ARG);
    for (int j = 0; j < njobs; ++j) {
        const dsdgpu_job& job = jobs[j];
        if (job.nch < 1 || job.nlut < 1 || job.nstep < 1 || !job.lut || job.nframes < 0 || (job.nframes > 0 && !job.out) ||
            (job.in_count > 0 && !job.in))
            return fail("dsdgpu_run_batch: job " + std::to_string(j) + " is incomplete", DSDGPU_ERR_ARG);
    }
    std::vector<Worker> workers((size_t)ndevices);
    std::vector<std::thread> threads;
    const int pieces = world * ndevices;
    for (int d = 0; d < ndevices; ++d) {
        workers[(size_t)d].device = devices[d];
        threads.emplace_back([&, d]() { run_piece(workers[(size_t)d], jobs, njobs, rank * ndevices + d, pieces, align_frames); });
    }
    for (auto& t : threads) t.join();
    for (int d = 0; d < ndevices; ++d)
        if (!workers[(size_t)d].error.empty())
            return fail("device " + std::to_string(devices[d]) + ": " + workers[(size_t)d].error, DSDGPU_ERR_CUDA);
    return DSDGPU_OK;
}
