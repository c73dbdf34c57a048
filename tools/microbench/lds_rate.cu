// Microbenchmark: what wavefront rate does the shared-memory pipe of one SM sustain for
// conflict-free LDS.64 streams?  (Practical ceiling under the 128 B/clk/SM roofline used in
// bench.py / DESIGN.md.)  Prints SM cycles per warp-level LDS.64 (ideal 2.0 = two 128-B wavefronts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_rate lds_rate.cu && ./lds_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROW = 640;            // bytes per table row (80 doubles)
constexpr int LUT_BYTES = 256 * ROW;
constexpr int ITER = 2048;

// MODE 0: fixed conflict-free addresses, integer sink.   MODE 1: + one DADD per load (4 chains).
// MODE 2: data-dependent row (byte from a register stream) + IMAD address + DADD, like the ring kernel.
template <int MODE, int CHAINS, int SHFL = 0>
__global__ void lds_kernel(unsigned long long* cycles, double* sink, uint32_t seed) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < LUT_BYTES / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, j = lane & 15;
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u;
    double acc[CHAINS];
    unsigned long long isink = 0;
    uint32_t sh = x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = 0.0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            if (MODE == 2 && (s & 3) == 0) x = x * 1664525u + 1013904223u;
            if (SHFL && (s % SHFL) == 0) sh = __shfl_up_sync(0xffffffffu, sh, 3) + x;     // SHFL-th of the steps carries one warp shuffle
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                uint32_t addr;
                if (MODE == 2) {
                    const uint32_t e = __byte_perm(x, 0u, 0x4440u | ((s + c) & 3));
                    addr = e * ROW + 8u * (uint32_t)((s - j) & 15) + 128u * (uint32_t)c;
                } else {
                    addr = (uint32_t)(c * 4 + (it & 63)) * ROW + 8u * (uint32_t)((s - j) & 15);
                }
                const double v = *reinterpret_cast<const double*>(smem + addr);
                if (MODE == 0) isink += (unsigned long long)__double_as_longlong(v);
                else acc[c] = __dadd_rn(acc[c], v);
            }
        }
    }
    const unsigned long long t1 = clock64();
    double tot = (double)isink + (double)sh;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) tot += acc[c];
    if (tot == 1.2345) sink[0] = tot;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int CHAINS, int SHFL = 0>
void run(const char* name, int threads) {
    unsigned long long* d_cyc; double* d_sink;
    cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 8);
    cudaFuncSetAttribute(lds_kernel<MODE, CHAINS, SHFL>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_BYTES + 1024);
    for (int rep = 0; rep < 2; ++rep) lds_kernel<MODE, CHAINS, SHFL><<<148, threads, LUT_BYTES + 1024>>>(d_cyc, d_sink, 7);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
    double worst = 0, sum = 0;
    for (auto c : h) { worst = c > worst ? c : worst; sum += c; }
    const double loads_per_sm = (double)ITER * 16 * CHAINS * (threads / 32);
    printf("%-44s threads=%4d  cycles/LDS.64 per SM: mean %.3f worst %.3f  -> %.1f%% of 128 B/clk  (%s)\n", name, threads,
           sum / 148 / loads_per_sm, worst / loads_per_sm, 100.0 * 2.0 / (sum / 148 / loads_per_sm), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
    run<0, 4>("fixed rows, integer sink, 4 loads/step", 512);
    run<0, 4>("fixed rows, integer sink, 4 loads/step", 1024);
    run<0, 8>("fixed rows, integer sink, 8 loads/step", 512);
    run<1, 4>("fixed rows, 4 fp64 add chains", 512);
    run<1, 4>("fixed rows, 4 fp64 add chains", 1024);
    run<2, 4>("data-dependent rows + IMAD + 4 fp64 chains", 512);
    run<2, 4>("data-dependent rows + IMAD + 4 fp64 chains", 1024);
    run<2, 1>("data-dependent rows + IMAD + 1 fp64 chain", 512);
    run<2, 1>("data-dependent rows + IMAD + 1 fp64 chain", 1024);
    run<2, 2>("data-dependent rows + IMAD + 2 fp64 chains", 512);
    run<2, 2>("data-dependent rows + IMAD + 2 fp64 chains", 1024);
    // do warp shuffles take wavefronts of the same pipe?  4 loads per step; one SHFL every SHFL-th step
    run<2, 4, 4>("data-dependent + 1 SHFL per 16 LDS.64", 512);
    run<2, 4, 2>("data-dependent + 1 SHFL per 8 LDS.64", 512);
    run<2, 4, 1>("data-dependent + 1 SHFL per 4 LDS.64", 512);
    return 0;
}
