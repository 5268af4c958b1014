// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_X32(PACK)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32" PACK ".b32 "                                            \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                    \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"    \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),        \
                   "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),    \
                   "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), \
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                         \
                 : "r"(addr) : "memory")
#define LD_X16()                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                   \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"             \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),        \
                   "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),    \
                   "=r"(v[14]), "=r"(v[15])                                                                   \
                 : "r"(addr) : "memory")
#define LD_X8()                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"             \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
                 : "r"(addr) : "memory")

// mode 0: x32 ; 1: x32 pack16 ; 2: x16 ; 3: x8 ; 4: tcgen05.st x32
template <int MODE, int INFLIGHT>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = r;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INFLIGHT; ++j) {
            const uint32_t addr = base + ((it * INFLIGHT + j) * 32 + (warp >> 2) * 64) % 448;
            if (MODE == 0) LD_X32("");
            if (MODE == 1) LD_X32(".pack::16b");
            if (MODE == 2) LD_X16();
            if (MODE == 3) LD_X8();
            if (MODE == 4)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                             "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
                             "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(addr), "r"(acc) : "memory");
            if (INFLIGHT > 1 && MODE != 4) acc += v[0] ^ v[7];   // consume lightly (after the wait below would be cleaner)
        }
        if (MODE == 4) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += v[0] + v[5];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE, int INFLIGHT>
void run(const char* name, int warps, int cols_per_ld) {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 20000;
    k<MODE, INFLIGHT><<<148, warps * 32>>>(iters, out, sink);
    cudaDeviceSynchronize();
    k<MODE, INFLIGHT><<<148, warps * 32>>>(iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double clk = (double)h[0];
    double lds = (double)iters * INFLIGHT * warps;
    printf("%-14s warps=%2d inflight=%d : %7.1f clk per ld per SM-warp-slot, %6.1f clk/ld/SMSP, %7.1f col*lanes/clk/SM  (%s)\n", name, warps,
           INFLIGHT, clk / (iters * INFLIGHT), clk / (lds / 4), lds * 32.0 * cols_per_ld / clk, cudaGetErrorString(e));
    cudaFree(out); cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0, 1>("x32", w, 32); run<0, 2>("x32", w, 32); run<0, 4>("x32", w, 32);
        run<1, 1>("x32.pack16", w, 64); run<1, 2>("x32.pack16", w, 64); run<1, 4>("x32.pack16", w, 64);
        run<2, 1>("x16", w, 16); run<2, 4>("x16", w, 16);
        run<3, 4>("x8", w, 8);
        run<4, 1>("st.x32", w, 32); run<4, 4>("st.x32", w, 32);
    }
    return 0;
}
