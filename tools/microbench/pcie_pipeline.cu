// Upload -> kernel -> download pipelines over the PCIe link, one configs[1] step (84.7 MB each way) in chunks:
//   A  what dsdgpu_decimate_host does: S slot streams, each chunk's three operations on its slot's stream, the
//      host waits for a slot before it reuses it
//   B  three streams (uploads / kernels / downloads) chained by events, a ring of R device buffers, everything
//      enqueued up front
// The "kernel" is a stand-in that touches the chunk (read in, write out) -- about what the decimation kernel moves.
// Build: nvcc -O2 -arch=sm_100a -o pcie_pipeline pcie_pipeline.cu
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

int main(int argc, char** argv) {
    const size_t total = 84672000;
    const int spin = argc > 1 ? atoi(argv[1]) : 600;
    const int blocks = argc > 2 ? atoi(argv[2]) : 148 * 4;
    const int flat = argc > 3 ? atoi(argv[3]) : 0;           // 1: uploads are plain 1-D copies instead of 2-row pitched ones   // grid of the stand-in kernel (fewer blocks: leaves SMs idle)     // tunes the stand-in kernel to ~40-80 us per 8 MB chunk
    unsigned char *h_in, *h_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    const int R = 8;
    const size_t cap = (size_t)1 << 25;
    unsigned char *d_in[R], *d_out[R];
    for (int r = 0; r < R; ++r) { CK(cudaMalloc(&d_in[r], cap)); CK(cudaMalloc(&d_out[r], cap)); }
    cudaStream_t st[R], up, mid, down;
    for (int r = 0; r < R; ++r) CK(cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&mid, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> e_up(64), e_k(64), e_dn(64);
    for (auto& e : e_up) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : e_k) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : e_dn) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {   // kernel alone on one chunk
        CK(cudaDeviceSynchronize());
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        touch<<<148 * 4, 256>>>((const uint4*)d_in[0], (uint4*)d_out[0], ((size_t)1 << 23) / 16, spin);
        CK(cudaEventRecord(a)); touch<<<148 * 4, 256>>>((const uint4*)d_in[0], (uint4*)d_out[0], ((size_t)1 << 23) / 16, spin); CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize()); float ms; CK(cudaEventElapsedTime(&ms, a, b)); printf("stand-in kernel on 8 MB: %.1f us\n", ms * 1e3);
    }
    for (size_t chunk : {(size_t)1 << 22, (size_t)1 << 23, (size_t)1 << 24}) {
        for (int S : {2, 4, 8}) {
            double bestA = 1e9, bestB = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                CK(cudaDeviceSynchronize());
                double t0 = now();
                int k = 0;
                for (size_t o = 0; o < total; o += chunk, ++k) {
                    const size_t n = total - o < chunk ? total - o : chunk;
                    const int s = k % S;
                    if (k >= S) CK(cudaStreamSynchronize(st[s]));
                    if (flat) CK(cudaMemcpyAsync(d_in[s], h_in + o, n, cudaMemcpyHostToDevice, st[s])); else CK(cudaMemcpy2DAsync(d_in[s], cap / 2, h_in + o / 2, total / 2, n / 2, 2, cudaMemcpyHostToDevice, st[s]));
                    if (spin >= 0) touch<<<blocks, 256, 0, st[s]>>>((const uint4*)d_in[s], (uint4*)d_out[s], n / 16, spin);
                    CK(cudaMemcpyAsync(h_out + o, d_out[s], n, cudaMemcpyDeviceToHost, st[s]));
                }
                CK(cudaDeviceSynchronize());
                double t = now() - t0; if (t < bestA) bestA = t;
                // B
                CK(cudaDeviceSynchronize());
                t0 = now(); k = 0;
                for (size_t o = 0; o < total; o += chunk, ++k) {
                    const size_t n = total - o < chunk ? total - o : chunk;
                    const int s = k % S;
                    if (k >= S) CK(cudaStreamWaitEvent(up, e_k[k - S], 0));          // d_in[s] free once kernel k-S has read it
                    if (flat) CK(cudaMemcpyAsync(d_in[s], h_in + o, n, cudaMemcpyHostToDevice, up)); else CK(cudaMemcpy2DAsync(d_in[s], cap / 2, h_in + o / 2, total / 2, n / 2, 2, cudaMemcpyHostToDevice, up));
                    CK(cudaEventRecord(e_up[k], up));
                    CK(cudaStreamWaitEvent(mid, e_up[k], 0));
                    if (k >= S) CK(cudaStreamWaitEvent(mid, e_dn[k - S], 0));         // d_out[s] free once download k-S is done
                    if (spin >= 0) touch<<<blocks, 256, 0, mid>>>((const uint4*)d_in[s], (uint4*)d_out[s], n / 16, spin);
                    CK(cudaEventRecord(e_k[k], mid));
                    CK(cudaStreamWaitEvent(down, e_k[k], 0));
                    CK(cudaMemcpyAsync(h_out + o, d_out[s], n, cudaMemcpyDeviceToHost, down));
                    CK(cudaEventRecord(e_dn[k], down));
                }
                CK(cudaDeviceSynchronize());
                t = now() - t0; if (t < bestB) bestB = t;
            }
            if (S == 8 && chunk == ((size_t)1 << 23)) {      // timeline of structure B
                std::vector<cudaEvent_t> tu(32), tk(32), td(32);
                for (auto& x : tu) CK(cudaEventCreate(&x));
                for (auto& x : tk) CK(cudaEventCreate(&x));
                for (auto& x : td) CK(cudaEventCreate(&x));
                cudaEvent_t base; CK(cudaEventCreate(&base));
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(base, up));
                int k = 0;
                for (size_t o = 0; o < total; o += chunk, ++k) {
                    const size_t n = total - o < chunk ? total - o : chunk;
                    const int s = k % S;
                    if (k >= S) CK(cudaStreamWaitEvent(up, tk[k - S], 0));
                    if (flat) CK(cudaMemcpyAsync(d_in[s], h_in + o, n, cudaMemcpyHostToDevice, up)); else CK(cudaMemcpy2DAsync(d_in[s], cap / 2, h_in + o / 2, total / 2, n / 2, 2, cudaMemcpyHostToDevice, up));
                    CK(cudaEventRecord(tu[k], up));
                    CK(cudaStreamWaitEvent(mid, tu[k], 0));
                    if (k >= S) CK(cudaStreamWaitEvent(mid, td[k - S], 0));
                    if (spin >= 0) touch<<<blocks, 256, 0, mid>>>((const uint4*)d_in[s], (uint4*)d_out[s], n / 16, spin);
                    CK(cudaEventRecord(tk[k], mid));
                    CK(cudaStreamWaitEvent(down, tk[k], 0));
                    CK(cudaMemcpyAsync(h_out + o, d_out[s], n, cudaMemcpyDeviceToHost, down));
                    CK(cudaEventRecord(td[k], down));
                }
                CK(cudaDeviceSynchronize());
                for (int i = 0; i < k; ++i) {
                    float a, b, c;
                    CK(cudaEventElapsedTime(&a, base, tu[i])); CK(cudaEventElapsedTime(&b, base, tk[i])); CK(cudaEventElapsedTime(&c, base, td[i]));
                    printf("  B chunk %2d: upload done %.3f  kernel done %.3f  download done %.3f\n", i, a, b, c);
                }
            }
            printf("chunk %9zu buffers %d: A (slot streams, host waits) %.3f ms | B (3 streams, events) %.3f ms\n", chunk, S, bestA * 1e3, bestB * 1e3);
        }
    }
    return 0;
}
