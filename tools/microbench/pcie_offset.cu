// Both-direction copies of one configs[1] step in 8 MB chunks, all queued up front, the download stream held back
// by a spin kernel of varying length first: does it matter whether chunk boundaries of the two directions coincide?
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
int main() {
    const size_t total = 84672000, chunk = 1 << 23;
    unsigned char *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    CK(cudaMalloc(&d_in, total)); CK(cudaMalloc(&d_out, total));
    cudaStream_t up, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    for (int iters : {0, 10000, 20000, 40000, 80000}) {
        double best = 1e9, kbest = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaDeviceSynchronize());
            double k0 = now();
            if (iters) { spin_kernel<<<1, 32, 0, down>>>((unsigned*)d_out, iters); CK(cudaStreamSynchronize(down)); }
            double kt = now() - k0;
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            if (iters) spin_kernel<<<1, 32, 0, down>>>((unsigned*)d_out, iters);
            for (size_t o = 0; o < total; o += chunk) {
                const size_t n = total - o < chunk ? total - o : chunk;
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, up));
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, down));
            }
            CK(cudaDeviceSynchronize());
            const double t = now() - t0;
            if (t < best) { best = t; kbest = kt; }
        }
        printf("download stream held back %.0f us: %.3f ms\n", kbest * 1e6, best * 1e3);
    }
    return 0;
}
