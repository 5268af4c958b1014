// Microbenchmark of the integer-pipe roofline that binds the XOR+POPC compare ("Both peaks must be
// measured by microbenchmarks" - MEASURED_PEAKS.json only holds HBM and bf16 tensor numbers).
//
// Every thread keeps 8 query words in registers and streams synthetic database words generated in registers
// (one IMAD per word, fma pipe), accumulating popc(q ^ r): per 32-bit compare exactly one LOP3, one POPC and one
// IADD - the instruction mix of the compare itself with no memory traffic, no counters and no epilogue.  The rate
// it reaches is the ceiling any popc-based Hamming kernel can approach on this chip.
#include "common.cuh"

namespace cmh {

constexpr int PEAK_Q = 8;

__global__ void __launch_bounds__(256) popc_peak_kernel(int iters, uint32_t seed, uint32_t* __restrict__ sink) {
    uint32_t q[PEAK_Q];
#pragma unroll
    for (int i = 0; i < PEAK_Q; ++i) q[i] = seed * (2 * i + 1) + threadIdx.x * 0x9E3779B9u + blockIdx.x;
    uint32_t acc[PEAK_Q];
#pragma unroll
    for (int i = 0; i < PEAK_Q; ++i) acc[i] = 0;
    uint32_t r = seed ^ (blockIdx.x * 256u + threadIdx.x);
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        r = r * 1664525u + 1013904223u;
#pragma unroll
        for (int i = 0; i < PEAK_Q; ++i) acc[i] += __popc(q[i] ^ r);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < PEAK_Q; ++i) s += acc[i];
    if (s == 0xffffffffu) sink[0] = s;  // never true in practice; keeps the loop alive
}

}  // namespace cmh

using namespace cmh;

// Runs the microbenchmark on the current device / `stream` and returns the best of `reps` timings as
// 32-bit XOR+POPC compares per second.  Synchronises the stream.
extern "C" int cmh_measure_popc_peak(int iters, int reps, double* popc32_per_s, void* stream) {
    CMH_REQUIRE(iters > 0 && reps > 0 && popc32_per_s, CMH_ERR_ARG, "cmh_measure_popc_peak: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* sink = nullptr;
    CMH_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CMH_CUDA(cudaEventCreate(&e0));
    CMH_CUDA(cudaEventCreate(&e1));
    const int grid = sm_count() * 8, block = 256;
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {  // first launch is a warm-up
        CMH_CUDA(cudaEventRecord(e0, st));
        popc_peak_kernel<<<grid, block, 0, st>>>(iters, 12345u + r, sink);
        CMH_LAUNCH_CHECK("popc_peak_kernel");
        CMH_CUDA(cudaEventRecord(e1, st));
        CMH_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CMH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)grid * block * (double)iters * PEAK_Q / (ms * 1e-3);
        if (r > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *popc32_per_s = best;
    return CMH_OK;
}
