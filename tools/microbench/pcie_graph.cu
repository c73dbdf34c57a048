// The chained upload -> kernel -> download pipeline (8 MB chunks, three streams, events) captured once into a
// CUDA graph and replayed: does the graph's dependency handling overlap better than stream-ordered work?
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
    cudaEvent_t fork, join1, join2;
    CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&join1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
    cudaGraph_t graph; cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(up, cudaStreamCaptureModeGlobal));
    CK(cudaEventRecord(fork, up)); CK(cudaStreamWaitEvent(mid, fork, 0)); CK(cudaStreamWaitEvent(down, fork, 0));
    int k = 0;
    for (size_t o = 0; o < total; o += chunk, ++k) {
        const size_t n = total - o < chunk ? total - o : chunk;
        CK(cudaMemcpyAsync(d_in[k], h_in + o, n, cudaMemcpyHostToDevice, up)); CK(cudaEventRecord(e_up[k], up));
        CK(cudaStreamWaitEvent(mid, e_up[k], 0));
        touch<<<592, 256, 0, mid>>>((const uint4*)d_in[k], (uint4*)d_out[k], n / 16, spin);
        CK(cudaEventRecord(e_k[k], mid)); CK(cudaStreamWaitEvent(down, e_k[k], 0));
        CK(cudaMemcpyAsync(h_out + o, d_out[k], n, cudaMemcpyDeviceToHost, down));
    }
    CK(cudaEventRecord(join1, mid)); CK(cudaEventRecord(join2, down));
    CK(cudaStreamWaitEvent(up, join1, 0)); CK(cudaStreamWaitEvent(up, join2, 0));
    CK(cudaStreamEndCapture(up, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    double best = 1e9;
    for (int rep = 0; rep < 8; ++rep) {
        CK(cudaDeviceSynchronize());
        const double t0 = now();
        CK(cudaGraphLaunch(exec, up));
        CK(cudaStreamSynchronize(up));
        const double t = now() - t0;
        if (t < best) best = t;
    }
    printf("graph replay of the chained pipeline: %.3f ms\n", best * 1e3);
    return 0;
}
