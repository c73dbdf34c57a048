// pcie_lab.cu -- which shape of the upload -> kernel -> download pipeline gets closest to the duplex copy floor?
// One configs[1] step (84.7 MB each way) in chunks; the kernel is a stand-in of `spin` iterations per 16 bytes.
// Variants (best of 6, wall clock around enqueue + sync):
//   slots S      the product's shape: S streams, each upload / kernel / download in order, host waits to reuse a stream
//   chain3       three streams (up, compute, down) chained by timing-disabled events
//   chain3x2     two upload and two download streams (alternating chunks), one compute stream, events
//   updown2      upload stream -> event -> one stream carrying kernel AND download
//   memops       three streams chained by cuStreamWriteValue32 / cuStreamWaitValue32 instead of events
//   persist      uploads + write-value flags on one stream; ONE persistent kernel polls the flags, works chunk by chunk
//                into a device buffer and counts finished blocks per chunk; the download stream waits on the counters
// Build: nvcc -O2 -arch=sm_100a -o pcie_lab pcie_lab.cu -lcuda ; run: pcie_lab <spin> <chunk MB>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("%s: error %d\n", #x, (int)r_); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void touch(const uint4* in, uint4* out, size_t n, int spin) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
        out[i] = v;
    }
}

__global__ void persistent(const uint4* in, uint4* out, size_t total16, size_t chunk16, int spin,
                           const volatile unsigned* ready, unsigned* done, unsigned epoch) {
    const size_t nchunks = (total16 + chunk16 - 1) / chunk16;
    for (size_t c = 0; c < nchunks; ++c) {
        if (threadIdx.x == 0) {
            unsigned long long t0 = clock64();
            while (ready[c] != epoch) { if (clock64() - t0 > 4000000000ULL) __trap(); }
        }
        __syncthreads();
        const size_t lo = c * chunk16, hi = lo + chunk16 < total16 ? lo + chunk16 : total16;
        for (size_t i = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = __ldcg(in + i);
            for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
            out[i] = v;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(done + c, 1u);
    }
}

// the same, storing straight into pinned host memory; keep = 4: every 16-byte piece, keep = 3: three of four (the
// volume of packed 24-bit output)
__global__ void persistent_host(const uint4* in, uint4* host_out, size_t total16, size_t chunk16, int spin,
                                const volatile unsigned* ready, unsigned epoch, int keep) {
    const size_t nchunks = (total16 + chunk16 - 1) / chunk16;
    for (size_t c = 0; c < nchunks; ++c) {
        if (threadIdx.x == 0) {
            unsigned long long t0 = clock64();
            while (ready[c] != epoch) { if (clock64() - t0 > 4000000000ULL) __trap(); }
        }
        __syncthreads();
        const size_t lo = c * chunk16, hi = lo + chunk16 < total16 ? lo + chunk16 : total16;
        for (size_t i = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = __ldcg(in + i);
            for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
            if ((i & 3) < (size_t)keep) host_out[i - (keep == 3 ? i / 4 : 0)] = v;
        }
    }
}

int main(int argc, char** argv) {
    const size_t total = 84672000;
    const int spin = argc > 1 ? atoi(argv[1]) : 3000;
    const size_t chunk = (size_t)(argc > 2 ? atof(argv[2]) : 8.0) * (1 << 20);
    const int nchunks = (int)((total + chunk - 1) / chunk);
    unsigned char *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    CK(cudaMalloc(&d_in, total)); CK(cudaMalloc(&d_out, total));
    unsigned *d_ready, *d_done;
    CK(cudaMalloc(&d_ready, 256 * 4)); CK(cudaMalloc(&d_done, 256 * 4));
    CK(cudaMemset(d_ready, 0, 256 * 4)); CK(cudaMemset(d_done, 0, 256 * 4));
    const int NS = 8;
    cudaStream_t st[NS];
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(2 * 256);
    for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // time of the stand-in kernel alone
    {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        touch<<<592, 256, 0, st[0]>>>((const uint4*)d_in, (uint4*)d_out, chunk / 16, spin);
        CK(cudaEventRecord(a, st[0]));
        for (int i = 0; i < 10; ++i) touch<<<592, 256, 0, st[0]>>>((const uint4*)d_in, (uint4*)d_out, chunk / 16, spin);
        CK(cudaEventRecord(b, st[0])); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf("# spin %d, chunk %.0f MB x %d: stand-in kernel %.1f us per chunk, %.3f ms per step\n", spin, chunk / 1048576.0, nchunks, ms * 100, ms / 10 * nchunks);
    }
    auto span = [&](int k, size_t& o, size_t& n) { o = (size_t)k * chunk; n = total - o < chunk ? total - o : chunk; };
    auto bench = [&](const char* name, auto&& body) {
        double best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            body(rep);
            CK(cudaDeviceSynchronize());
            const double t = now() - t0;
            if (t < best) best = t;
        }
        printf("%-28s %.3f ms\n", name, best * 1e3);
        fflush(stdout);
    };
    bench("copies only, two streams", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[1]));
        }
    });
    for (int S : {2, 3, 4, 6, 8}) {
        char name[64]; snprintf(name, sizeof name, "slots %d", S);
        bench(name, [&](int) {
            for (int k = 0; k < nchunks; ++k) {
                size_t o, n; span(k, o, n);
                cudaStream_t s = st[k % S];
                if (k >= S) CK(cudaStreamSynchronize(s));
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, s));
                touch<<<592, 256, 0, s>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, s));
            }
        });
        snprintf(name, sizeof name, "slots %d, no host waits", S);
        bench(name, [&](int) {
            for (int k = 0; k < nchunks; ++k) {
                size_t o, n; span(k, o, n);
                cudaStream_t s = st[k % S];
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, s));
                touch<<<592, 256, 0, s>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, s));
            }
        });
    }
    bench("chain3 (events)", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            CK(cudaEventRecord(ev[2 * k], st[0]));
            CK(cudaStreamWaitEvent(st[1], ev[2 * k], 0));
            touch<<<592, 256, 0, st[1]>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CK(cudaEventRecord(ev[2 * k + 1], st[1]));
            CK(cudaStreamWaitEvent(st[2], ev[2 * k + 1], 0));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[2]));
        }
    });
    bench("chain3, uploads first", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            CK(cudaEventRecord(ev[2 * k], st[0]));
        }
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaStreamWaitEvent(st[1], ev[2 * k], 0));
            touch<<<592, 256, 0, st[1]>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CK(cudaEventRecord(ev[2 * k + 1], st[1]));
            CK(cudaStreamWaitEvent(st[2], ev[2 * k + 1], 0));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[2]));
        }
    });
    bench("chain3x2 (2 up, 2 down)", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            cudaStream_t up = st[k & 1], down = st[2 + (k & 1)], mid = st[4];
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, up));
            CK(cudaEventRecord(ev[2 * k], up));
            CK(cudaStreamWaitEvent(mid, ev[2 * k], 0));
            touch<<<592, 256, 0, mid>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CK(cudaEventRecord(ev[2 * k + 1], mid));
            CK(cudaStreamWaitEvent(down, ev[2 * k + 1], 0));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, down));
        }
    });
    bench("updown2 (kernel+down share)", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            CK(cudaEventRecord(ev[2 * k], st[0]));
            CK(cudaStreamWaitEvent(st[1], ev[2 * k], 0));
            touch<<<592, 256, 0, st[1]>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[1]));
        }
    });
    bench("upkernel2 (up+kernel share)", [&](int) {
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            touch<<<592, 256, 0, st[0]>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CK(cudaEventRecord(ev[2 * k], st[0]));
            CK(cudaStreamWaitEvent(st[1], ev[2 * k], 0));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[1]));
        }
    });
    unsigned epoch = 0;
    bench("memops (write/wait value)", [&](int) {
        ++epoch;
        for (int k = 0; k < nchunks; ++k) {
            size_t o, n; span(k, o, n);
            CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
            CU(cuStreamWriteValue32(st[0], (CUdeviceptr)(d_ready + k), epoch, 0));
            CU(cuStreamWaitValue32(st[1], (CUdeviceptr)(d_ready + k), epoch, CU_STREAM_WAIT_VALUE_EQ));
            touch<<<592, 256, 0, st[1]>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), n / 16, spin);
            CU(cuStreamWriteValue32(st[1], (CUdeviceptr)(d_done + k), epoch, 0));
            CU(cuStreamWaitValue32(st[2], (CUdeviceptr)(d_done + k), epoch, CU_STREAM_WAIT_VALUE_EQ));
            CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[2]));
        }
    });
    for (int blocks : {148, 592}) {
        char name[64]; snprintf(name, sizeof name, "persist, %d blocks", blocks);
        bench(name, [&](int) {
            ++epoch;
            CK(cudaMemsetAsync(d_done, 0, 256 * 4, st[1]));
            persistent<<<blocks, 256, 0, st[1]>>>((const uint4*)d_in, (uint4*)d_out, total / 16, chunk / 16, spin, d_ready, d_done, epoch);
            for (int k = 0; k < nchunks; ++k) {
                size_t o, n; span(k, o, n);
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
                CU(cuStreamWriteValue32(st[0], (CUdeviceptr)(d_ready + k), epoch, 0));
                CU(cuStreamWaitValue32(st[2], (CUdeviceptr)(d_done + k), (unsigned)blocks, CU_STREAM_WAIT_VALUE_GEQ));
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, st[2]));
            }
        });
    }
    for (int keep : {4, 3})
        for (int blocks : {148, 296, 592}) {
            char name[64]; snprintf(name, sizeof name, "persist->host %s, %d blk", keep == 4 ? "int32" : "pcm24", blocks);
            bench(name, [&](int) {
                ++epoch;
                persistent_host<<<blocks, 256, 0, st[1]>>>((const uint4*)d_in, (uint4*)h_out, total / 16, chunk / 16, spin, d_ready, epoch, keep);
                for (int k = 0; k < nchunks; ++k) {
                    size_t o, n; span(k, o, n);
                    CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, st[0]));
                    CU(cuStreamWriteValue32(st[0], (CUdeviceptr)(d_ready + k), epoch, 0));
                }
            });
        }
    // tapered schedules: few big pieces in the middle (fewer hand-offs), small ones at both ends (short fill and drain)
    {
        const double MB = 1048576.0;
        std::vector<std::vector<double>> plans = {
            {2, 6, 24, 24, 16.75, 6, 2}, {1, 3, 8, 16, 24, 16, 8, 3, 1.75}, {4, 12, 24, 24, 12.75, 4}, {2, 6, 12, 12, 12, 12, 12, 6, 4.75, 2},
            {1, 2, 4, 8, 16, 20, 16, 8, 3.75, 2}, {0.5, 1.5, 4, 8, 16, 21.25, 16, 8, 4, 1.5}};
        for (auto& plan : plans) {
            std::vector<size_t> off, len;
            size_t o = 0;
            for (size_t i = 0; i < plan.size() && o < total; ++i) {
                size_t n = (size_t)(plan[i] * MB) & ~(size_t)255;
                if (i + 1 == plan.size() || o + n > total) n = total - o;
                off.push_back(o); len.push_back(n); o += n;
            }
            if (o < total) { off.push_back(o); len.push_back(total - o); }
            char name[128]; int w = snprintf(name, sizeof name, "taper");
            for (double p : plan) w += snprintf(name + w, sizeof name - w, " %g", p);
            bench(name, [&](int) {
                for (size_t k = 0; k < off.size(); ++k) {
                    CK(cudaMemcpyAsync(d_in + off[k], h_in + off[k], len[k], cudaMemcpyHostToDevice, st[0]));
                    CK(cudaEventRecord(ev[2 * k], st[0]));
                    CK(cudaStreamWaitEvent(st[1], ev[2 * k], 0));
                    // kernels in 8 MB sub-pieces so that the stand-in time scales like the real kernel's
                    for (size_t s = 0; s < len[k]; s += chunk) {
                        const size_t m = len[k] - s < chunk ? len[k] - s : chunk;
                        touch<<<592, 256, 0, st[1]>>>((const uint4*)(d_in + off[k] + s), (uint4*)(d_out + off[k] + s), m / 16, spin);
                    }
                    CK(cudaEventRecord(ev[2 * k + 1], st[1]));
                    CK(cudaStreamWaitEvent(st[2], ev[2 * k + 1], 0));
                    CK(cudaMemcpyAsync(h_out + off[k], d_out + off[k], len[k], cudaMemcpyDeviceToHost, st[2]));
                }
            });
        }
    }
    return 0;
}
