// Microbenchmark: which part of the ring kernel's period costs what?  Runs the very
// ring_head()/ring_tail() of csrc/ring_core.h on a static shared-memory image (random bytes,
// finite table) and reports SM cycles per warp-level LDS.64 for: the 56 main steps alone, the 16
// early steps alone, a whole period, and a whole period plus a quantise+store epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../dsf2flac_b200/csrc -o ring_parts ring_parts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ring_core.h"

using namespace dsdgpu;
using G = RingGeom<16, 3>;
constexpr int SMEM = G::LUT_BYTES + 2 * G::TILE_BYTES;
constexpr int ITER = 300;

struct SinkEmit {                 // a cheap stand-in for the kernel's emitter
    int sink = 0;
    __device__ void quantise(int, double sum) {
        double v = __dmul_rn(sum, 13295047.7);
        if (v > 8388607.0) v = 8388607.0; else if (v < -8388607.0) v = -8388607.0;
        const double r = trunc(v);
        sink += __double2int_rz(r + ((v - r >= 0.5) ? 1.0 : ((v - r <= -0.5) ? -1.0 : 0.0)));
    }
    __device__ void store() { sink ^= __shfl_xor_sync(0xffffffffu, sink, 16); }
};

template <int MODE>
__global__ void __launch_bounds__(G::NTHR, 1) parts_kernel(unsigned long long* cycles, int* out, uint32_t seed) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < G::LUT_BYTES / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-3 * (i % 977);
    uint32_t x = seed + threadIdx.x * 2654435761u;
    for (int i = threadIdx.x; i < 2 * G::TILE_BYTES / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t*>(smem + G::LUT_BYTES)[i] = x;
    }
    __syncthreads();
    const int tid = threadIdx.x, l = tid & 31, h = l >> 4;
    RingLane lane;
    const uint32_t w0 = (uint32_t)(G::LUT_BYTES + h * G::HALF_BYTES + ring_lane_init<G>(lane, tid, G::FRONT));
    lane.wlast = w0 + G::HOP;
    uint32_t loff = lane.lnew;
    int sink = 0;
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const uint32_t wnew = w0 + (uint32_t)((it % 3) * G::HOP);
        if (MODE != 0) ring_head<G>(lane, smem, wnew, loff, tid, 0, RingNoTrace());
        if (MODE == 3) { SinkEmit e; ring_tail<G>(lane, smem, wnew, loff, tid, 0, e, RingNoTrace()); sink += e.sink; }
        else if (MODE != 1) { RingNoEmit e; ring_tail<G>(lane, smem, wnew, loff, tid, 0, e, RingNoTrace()); sink += (int)lane.p[it & 3]; }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { lane.p[k] = lane.n[k]; lane.n[k] = 0.0; }
            lane.wlast = wnew;
        }
    }
    const unsigned long long t1 = clock64();
    double tot = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) tot += lane.n[k] + lane.p[k];
    if (tot == 1.2345 || sink == 12345) out[tid] = (int)tot;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int lds64_per_iter) {
    unsigned long long* d_cyc; int* d_out;
    cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_out, 4096);
    cudaFuncSetAttribute(parts_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    for (int rep = 0; rep < 2; ++rep) parts_kernel<MODE><<<148, G::NTHR, SMEM>>>(d_cyc, d_out, 7);
    cudaDeviceSynchronize();
    unsigned long long hc[148];
    cudaMemcpy(hc, d_cyc, sizeof hc, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (auto c : hc) sum += c;
    const double loads = (double)ITER * lds64_per_iter * G::NWARP;
    const double cyc = sum / 148 / loads;
    printf("%-46s cycles per LDS.64: %.3f  (table-load pipe time share %.1f%%)  %s\n", name, cyc, 100.0 * 2.0 / cyc,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_cyc); cudaFree(d_out);
}

int main() {
    run<0>("main steps 16..71 only (224 LDS.64 / warp)", 224);
    run<1>("early steps 0..15 only (64 LDS.64 / warp)", 64);
    run<2>("whole period (288 LDS.64 / warp)", 288);
    run<3>("whole period + quantise/shuffle epilogue", 288);
    return 0;
}
