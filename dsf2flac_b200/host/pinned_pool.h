// pinned_pool.h -- page-locked staging blocks, recycled across the decimators and containers of a process.
//
// Page-locking memory costs about 0.45 ms per MB (measured: 3.5 ms for the 8 MB PCM block of a stereo stream, 38 ms
// for the 85 MB payload of a one-minute DSD128 file), more than the GPU needs for the conversion itself.  A process
// that converts several streams (tracks of one file, a batch) therefore keeps the blocks of finished conversions.
#ifndef DSD_PINNED_POOL_H
#define DSD_PINNED_POOL_H

#include <cstddef>
#include <mutex>

#include "dsdgpu.h"

struct PinnedPool {
    static const int kSlots = 8;
    std::mutex lock;
    void* ptr[kSlots] = {nullptr};
    size_t cap[kSlots] = {0};
    // a block of at least `need` bytes (and not absurdly larger); got = its capacity
    void* take(size_t need, size_t& got) {
        {
            std::lock_guard<std::mutex> g(lock);
            int best = -1;
            for (int i = 0; i < kSlots; ++i)
                if (ptr[i] && cap[i] >= need && cap[i] <= 2 * need + (1u << 20) && (best < 0 || cap[i] < cap[best])) best = i;
            if (best >= 0) { void* p = ptr[best]; got = cap[best]; ptr[best] = nullptr; cap[best] = 0; return p; }
        }
        void* p = dsdgpu_host_alloc(need);
        got = p ? need : 0;
        return p;
    }
    void give(void* p, size_t n) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(lock);
            for (int i = 0; i < kSlots; ++i)
                if (!ptr[i]) { ptr[i] = p; cap[i] = n; return; }
            // full: the smallest block makes room for a larger one
            int small = 0;
            for (int i = 1; i < kSlots; ++i)
                if (cap[i] < cap[small]) small = i;
            if (cap[small] < n) { void* q = ptr[small]; ptr[small] = p; cap[small] = n; p = q; }
        }
        dsdgpu_host_free(p);
    }
    // no destructor: at static-destruction time the CUDA runtime may already be gone; the process's pinned
    // pages go back to the system with it
};

inline PinnedPool& pinnedPool() {
    static PinnedPool* pool = new PinnedPool();     // never destroyed, see above
    return *pool;
}

#endif
