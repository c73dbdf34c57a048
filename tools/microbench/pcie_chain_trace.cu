// Timeline of the chained upload -> kernel -> download pipeline (three streams, events), 8 MB chunks:
// start and end of every operation.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)
__global__ void touch(const uint4* in, uint4* out, size_t n, int spin) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        for (int k = 0; k < spin; ++k) v.x = v.x * 1664525u + 1013904223u;
        out[i] = v;
    }
}
int main(int argc, char** argv) {
    const size_t total = 84672000, chunk = 1 << 23;
    const int spin = argc > 1 ? atoi(argv[1]) : 6000;
    const int chained = argc > 2 ? atoi(argv[2]) : 1;
    unsigned char *h_in, *h_out;
    CK(cudaMallocHost(&h_in, total)); CK(cudaMallocHost(&h_out, total));
    const int R = 16;
    unsigned char *d_in[R], *d_out[R];
    for (int r = 0; r < R; ++r) { CK(cudaMalloc(&d_in[r], chunk)); CK(cudaMalloc(&d_out[r], chunk)); }
    cudaStream_t up, mid, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&mid, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(6 * 16);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    cudaEvent_t base; CK(cudaEventCreate(&base));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(base, up));
        int k = 0;
        for (size_t o = 0; o < total; o += chunk, ++k) {
            const size_t n = total - o < chunk ? total - o : chunk;
            cudaEvent_t* e = &ev[6 * k];
            CK(cudaEventRecord(e[0], up));
            CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up));
            CK(cudaEventRecord(e[1], up));
            if (chained) CK(cudaStreamWaitEvent(mid, e[1], 0));
            CK(cudaEventRecord(e[2], mid));
            touch<<<592, 256, 0, mid>>>((const uint4*)d_in[k], (uint4*)d_out[k], n / 16, spin);
            CK(cudaEventRecord(e[3], mid));
            CK(cudaStreamWaitEvent(down, e[3], 0));
            CK(cudaEventRecord(e[4], down));
            CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
            CK(cudaEventRecord(e[5], down));
        }
        CK(cudaDeviceSynchronize());
        if (rep == 2)
            for (int i = 0; i < k; ++i) {
                float t[6];
                for (int j = 0; j < 6; ++j) CK(cudaEventElapsedTime(&t[j], base, ev[6 * i + j]));
                printf("chunk %2d: upload %.3f-%.3f  kernel %.3f-%.3f  download %.3f-%.3f\n", i, t[0], t[1], t[2], t[3], t[4], t[5]);
            }
    }
    return 0;
}
