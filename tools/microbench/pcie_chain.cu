// Which link of the upload -> kernel -> download chain costs time?  One configs[1] step in 8 MB chunks, three
// streams chained by events (structure B of pcie_pipeline.cu), with parts of the chain switched off.
// Build: nvcc -O2 -arch=sm_100a -o pcie_chain pcie_chain.cu ; run: pcie_chain <spin>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void touch(const uint4* in, uint4* out, size_t n, int spin) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
        out[i] = v;
    }
}
// the stand-in kernel, started without any stream dependency on its upload: block 0 .. n wait for a flag the
// upload stream copies behind the chunk's data (epoch value, so stale flags of earlier runs do not count)
__global__ void touch_flag(const uint4* in, uint4* out, size_t n, int spin, const volatile unsigned* flag, unsigned epoch) {
    if (threadIdx.x == 0) {
        unsigned tries = 0;
        while (*flag != epoch) { if (++tries > (1u << 26)) __trap(); }
    }
    __syncthreads();
    __threadfence();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldcg(in + i);
        for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
        out[i] = v;
    }
}
int main(int argc, char** argv) {
    const size_t total = 84672000, chunk = 1 << 23;
    const int spin = argc > 1 ? atoi(argv[1]) : 6000;
    unsigned char *h_in, *h_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    const int R = 16;
    unsigned char *d_in[R], *d_out[R];
    for (int r = 0; r < R; ++r) { CK(cudaMalloc(&d_in[r], chunk)); CK(cudaMalloc(&d_out[r], chunk)); }
    cudaStream_t up, mid, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&mid, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> e_up(32), e_k(32);
    for (auto& e : e_up) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : e_k) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const char* names[] = {"upload only", "upload -> kernel (chained)", "upload, kernel (not chained)", "kernel -> download (chained)",
                           "upload -> kernel -> download", "upload -> kernel -> download, kernel on 1/4 of the SMs' worth of blocks",
                           "upload -> download (no kernel)", "upload, kernel, download: three free-running streams, no events",
                           "upload free-running; kernel -> download chained",
                           "upload + flag copy; kernel polls the flag (no event upload -> kernel); kernel -> download chained",
                           "upload -> kernel -> download chained, kernel works on unrelated buffers",
                           "upload -> kernel -> download chained, kernel is a pure spin (no memory traffic)"};
    unsigned char *d_x, *d_y; CK(cudaMalloc(&d_x, chunk)); CK(cudaMalloc(&d_y, chunk));
    unsigned* d_flags; CK(cudaMalloc(&d_flags, 64 * sizeof(unsigned))); CK(cudaMemset(d_flags, 0, 64 * sizeof(unsigned)));
    unsigned* h_epoch; CK(cudaMallocHost(&h_epoch, 4));
    unsigned epoch = 0;
    for (int mode = 0; mode < 12; ++mode) {
        double best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            int k = 0;
            *h_epoch = ++epoch;
            for (size_t o = 0; o < total; o += chunk, ++k) {
                const size_t n = total - o < chunk ? total - o : chunk;
                const bool upl = mode != 3, ker = mode != 0 && mode != 6, dwn = mode == 3 || mode >= 4;
                if (mode == 10 || mode == 11) {
                    CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up)); CK(cudaEventRecord(e_up[k], up));
                    CK(cudaStreamWaitEvent(mid, e_up[k], 0));
                    if (mode == 10) touch<<<592, 256, 0, mid>>>((const uint4*)d_x, (uint4*)d_y, n / 16, spin);
                    else touch<<<592, 256, 0, mid>>>((const uint4*)d_x, (uint4*)d_y, 592 * 256, spin * 14);
                    CK(cudaEventRecord(e_k[k], mid)); CK(cudaStreamWaitEvent(down, e_k[k], 0));
                    CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
                    continue;
                }
                if (mode == 9) {
                    CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up));
                    CK(cudaMemcpyAsync(d_flags + k, h_epoch, 4, cudaMemcpyHostToDevice, up));
                    touch_flag<<<592, 256, 0, mid>>>((const uint4*)d_in[k], (uint4*)d_out[k], n / 16, spin, d_flags + k, epoch);
                    CK(cudaEventRecord(e_k[k], mid)); CK(cudaStreamWaitEvent(down, e_k[k], 0));
                    CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
                    continue;
                }
                if (mode == 7 || mode == 8) {
                    CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up));
                    touch<<<592, 256, 0, mid>>>((const uint4*)d_in[k], (uint4*)d_out[k], n / 16, spin);
                    if (mode == 8) { CK(cudaEventRecord(e_k[k], mid)); CK(cudaStreamWaitEvent(down, e_k[k], 0)); }
                    CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
                    continue;
                }
                if (upl) { CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up)); CK(cudaEventRecord(e_up[k], up)); }
                if (ker) {
                    if (upl && mode != 2) CK(cudaStreamWaitEvent(mid, e_up[k], 0));
                    touch<<<mode == 5 ? 37 : 592, 256, 0, mid>>>((const uint4*)d_in[k], (uint4*)d_out[k], n / 16, mode == 5 ? spin / 16 : spin);
                    CK(cudaEventRecord(e_k[k], mid));
                }
                if (dwn) {
                    if (ker) CK(cudaStreamWaitEvent(down, e_k[k], 0)); else CK(cudaStreamWaitEvent(down, e_up[k], 0));
                    CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
                }
            }
            CK(cudaDeviceSynchronize());
            const double t = now() - t0;
            if (t < best) best = t;
        }
        printf("%-75s %.3f ms\n", names[mode], best * 1e3);
    }
    return 0;
}
