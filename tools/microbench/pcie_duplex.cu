// What the PCIe link gives for the end-to-end path's copy pattern: N uploads and N downloads of `chunk` bytes
// each, queued on two streams so that both directions are busy, for 1-D copies and for the pitched 2-D copy
// (2 rows) dsdgpu_submit uses for planar stereo bytes.  Build: nvcc -O2 -o pcie_duplex pcie_duplex.cu
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    const size_t total = 84672000;                      // one configs[1] step, each way
    unsigned char *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    CK(cudaMalloc(&d_in, total + (1 << 20))); CK(cudaMalloc(&d_out, total));
    cudaStream_t up, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    for (size_t chunk : {(size_t)1 << 21, (size_t)1 << 22, (size_t)1 << 23, (size_t)1 << 24, (size_t)1 << 25, total}) {
        for (int mode = 0; mode < 4; ++mode) {          // 0: up only 1-D, 1: down only, 2: both 1-D, 3: both, upload pitched 2 rows
            double best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaDeviceSynchronize());
                const double t0 = now();
                for (size_t o = 0; o < total; o += chunk) {
                    const size_t n = total - o < chunk ? total - o : chunk;
                    if (mode == 0 || mode == 2) CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, up));
                    if (mode == 3) CK(cudaMemcpy2DAsync(d_in + o / 2, total / 2 + 256, h_in + o / 2, total / 2, n / 2, 2, cudaMemcpyHostToDevice, up));
                    if (mode != 0) CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, down));
                }
                CK(cudaDeviceSynchronize());
                const double t = now() - t0;
                if (t < best) best = t;
            }
            printf("chunk %9zu mode %d: %.3f ms (%.1f GB/s per direction)\n", chunk, mode, best * 1e3, total / best / 1e9);
        }
    }
    return 0;
}
