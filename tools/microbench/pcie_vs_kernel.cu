// Does a running kernel slow PCIe copies down?  Both-direction copies of one configs[1] step (8 MB chunks, all
// queued up front) alone, beside a compute-only kernel that fills every SM, and beside a kernel that streams HBM.
// Build: nvcc -O2 -arch=sm_100a -o pcie_vs_kernel pcie_vs_kernel.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void spin_kernel(unsigned* out, int iters) {
    unsigned v = threadIdx.x;
    for (int k = 0; k < iters; ++k) v = v * 1664525u + 1013904223u;
    if (v == 12345u) out[0] = v;
}
__global__ void stream_kernel(const uint4* in, uint4* out, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    const size_t total = 84672000, chunk = 1 << 23;
    unsigned char *h_in, *h_out, *d_in, *d_out, *d_a, *d_b;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    CK(cudaMalloc(&d_in, total)); CK(cudaMalloc(&d_out, total));
    const size_t big = (size_t)1 << 30;
    CK(cudaMalloc(&d_a, big)); CK(cudaMalloc(&d_b, big));
    cudaStream_t up, down, k;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&k, cudaStreamNonBlocking));
    for (int mode = 0; mode < 5; ++mode) {
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaDeviceSynchronize());
            if (mode == 1) spin_kernel<<<148 * 2, 1024, 0, k>>>((unsigned*)d_a, 4000000);              // a few ms of pure ALU work
            if (mode == 2) stream_kernel<<<148 * 8, 256, 0, k>>>((const uint4*)d_a, (uint4*)d_b, big / 16, 8);   // a few ms of HBM traffic
            if (mode == 3) for (int q = 0; q < 60; ++q) spin_kernel<<<148 * 2, 1024, 0, k>>>((unsigned*)d_a, 30000);   // 60 short kernels back to back
            if (mode == 4) for (int q = 0; q < 60; ++q) stream_kernel<<<148 * 8, 256, 0, k>>>((const uint4*)d_a, (uint4*)d_b, ((size_t)1 << 27) / 16, 1);
            const double t0 = now();
            for (size_t o = 0; o < total; o += chunk) {
                const size_t n = total - o < chunk ? total - o : chunk;
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, up));
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, down));
            }
            CK(cudaStreamSynchronize(up)); CK(cudaStreamSynchronize(down));
            const double t = now() - t0;
            if (t < best) best = t;
            CK(cudaDeviceSynchronize());
        }
        printf("%s: %.3f ms for 84.7 MB each way\n", mode == 0 ? "copies alone" : mode == 1 ? "beside an ALU-bound kernel" : mode == 2 ? "beside an HBM-streaming kernel" : mode == 3 ? "beside 60 short ALU kernels" : "beside 60 short HBM kernels", best * 1e3);
    }
    return 0;
}
