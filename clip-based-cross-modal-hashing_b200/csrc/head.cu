// f1 - the DCHMT hash head fused with binarise + pack + scatter.
//
// Reference: model/DCHMT.py:16-26 - after fc(512 -> 128) + relu, `bits` SEPARATE nn.Linear(128, 2) layers each followed
// by a softmax; train/base.py:150-158 (make_hash_code_DCHMT) stacks the `bits` [n, 2] outputs, takes the argmax and
// maps class 0 to -1; train/base.py:176-177 scatters the float codes by dataset index.  That is bits tiny GEMMs, bits
// softmaxes, a stack, a permute, an argmax and a float [N, bits] buffer - to produce one BIT per (row, layer):
//     bit = argmax(softmax([l0, l1])) = [l1 > l0]        (softmax is monotonic; a tie is class 0 = -1)
// Here the bits x 2 weight rows form ONE [hidden, 2 * bits] matrix held in shared memory; a warp computes both logits of
// 32 bits at a time for two batch rows (lane = bit), one vote turns the 32 comparisons into a packed word, and the word
// is stored at the row's dataset index.  Logits never reach memory.
//
// Rounding: logits are accumulated in float32 over `hidden` in index order; the reference's cuBLAS / softmax path rounds
// differently, so bits whose two logits agree to ~1e-6 relative may differ (and the reference itself turns l1 - l0 <
// 2^-25 into a tie through exp rounding).  tests/test_gpu_parity.py::test_hash_head_fused checks equality everywhere
// else.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace cmh {

constexpr int HEAD_WARPS = 8;
constexpr int HEAD_ROWS = 2;      // batch rows per warp and pass over the weights

template <typename T> __device__ __forceinline__ float head_f(T v) { return (float)v; }
template <> __device__ __forceinline__ float head_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float head_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// smem: Wt [hidden][2 * bits] float (column 2 j + c = class c of bit j), bias [2 * bits], then per warp HEAD_ROWS x hidden floats
template <typename T>
__global__ void __launch_bounds__(HEAD_WARPS * 32) hash_head_pack_kernel(const T* __restrict__ x, int64_t n, int hidden, int64_t ld,
                                                                         int relu, const float* __restrict__ W /*[bits][2][hidden]*/,
                                                                         const float* __restrict__ bias /*[bits][2] or NULL*/,
                                                                         int bits, const long long* __restrict__ index, int64_t n_out,
                                                                         int w32_out, uint32_t* __restrict__ sign32,
                                                                         uint32_t* __restrict__ valid32,
                                                                         unsigned long long* counters) {
    extern __shared__ __align__(16) float smem_head[];
    const int cols = 2 * bits;
    float* Wt = smem_head;                         // [hidden][cols]
    float* sb = Wt + (size_t)hidden * cols;        // [cols]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* e = sb + cols + (size_t)warp * HEAD_ROWS * hidden;
    for (int i = threadIdx.x; i < hidden * cols; i += blockDim.x) {
        const int col = i / hidden, h = i - col * hidden;      // W is [col][h] with col = 2 j + c: coalesced read, transposed store
        Wt[(size_t)h * cols + col] = W[i];
    }
    for (int i = threadIdx.x; i < cols; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int n_seg = (bits + 31) >> 5;
    unsigned long long n_bad = 0;
    for (int64_t r0 = ((int64_t)blockIdx.x * HEAD_WARPS + warp) * HEAD_ROWS; r0 < n; r0 += (int64_t)gridDim.x * HEAD_WARPS * HEAD_ROWS) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u < HEAD_ROWS; ++u)
            for (int h = lane; h < hidden; h += 32) {
                float v = r0 + u < n ? head_f<T>(x[(r0 + u) * ld + h]) : 0.f;
                e[u * hidden + h] = relu ? fmaxf(v, 0.f) : v;
            }
        __syncwarp();
        int64_t dst[HEAD_ROWS];
#pragma unroll
        for (int u = 0; u < HEAD_ROWS; ++u) {
            dst[u] = r0 + u < n ? (index ? (int64_t)index[r0 + u] : r0 + u) : -1;
            if (r0 + u < n && (dst[u] < 0 || dst[u] >= n_out)) { n_bad += lane == 0; dst[u] = -1; }
        }
        for (int seg = 0; seg < n_seg; ++seg) {
            const int j = seg * 32 + lane;                     // this lane's bit
            const bool live = j < bits;
            const float2* wcol = reinterpret_cast<const float2*>(Wt) + (live ? j : 0);   // (class 0, class 1) of bit j, row stride bits
            float a0[HEAD_ROWS], a1[HEAD_ROWS];
#pragma unroll
            for (int u = 0; u < HEAD_ROWS; ++u) { a0[u] = 0.f; a1[u] = 0.f; }
#pragma unroll 4
            for (int h = 0; h < hidden; ++h) {
                const float2 w = wcol[(size_t)h * bits];
#pragma unroll
                for (int u = 0; u < HEAD_ROWS; ++u) {
                    const float ev = e[u * hidden + h];
                    a0[u] = fmaf(w.x, ev, a0[u]);
                    a1[u] = fmaf(w.y, ev, a1[u]);
                }
            }
            const float b0 = sb[2 * (live ? j : 0)], b1 = sb[2 * (live ? j : 0) + 1];
#pragma unroll
            for (int u = 0; u < HEAD_ROWS; ++u) {
                const uint32_t sw = __ballot_sync(0xffffffffu, live && (a1[u] + b1) > (a0[u] + b0));
                const uint32_t vw = __ballot_sync(0xffffffffu, live);
                if (lane == 0 && dst[u] >= 0) {
                    sign32[dst[u] * w32_out + seg] = sw;
                    if (valid32) valid32[dst[u] * w32_out + seg] = vw;
                }
            }
        }
        if (lane == 0 && (n_seg & 1))                          // the unused high half of the last 64-bit word
#pragma unroll
            for (int u = 0; u < HEAD_ROWS; ++u)
                if (dst[u] >= 0) {
                    sign32[dst[u] * w32_out + n_seg] = 0u;
                    if (valid32) valid32[dst[u] * w32_out + n_seg] = 0u;
                }
    }
    if (lane == 0 && n_bad && counters) atomicAdd(&counters[1], n_bad);
}

}  // namespace cmh

using namespace cmh;

extern "C" int cmh_hash_head_pack(const void* x, int dtype, int64_t n, int hidden, int64_t ld, int relu, const float* weight,
                                  const float* bias, int bits, const int64_t* index, int64_t n_out, uint64_t* sign_out,
                                  uint64_t* valid_out, unsigned long long* counters, void* stream) {
    CMH_REQUIRE(n >= 0 && hidden >= 1 && ld >= hidden && bits >= 1 && n_out >= 0, CMH_ERR_ARG,
                "cmh_hash_head_pack: bad shape n=%lld hidden=%d ld=%lld bits=%d", (long long)n, hidden, (long long)ld, bits);
    CMH_REQUIRE(bits <= CMH_MAX_BITS, CMH_ERR_UNSUPPORTED, "cmh_hash_head_pack: bits=%d > %d", bits, CMH_MAX_BITS);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(x && weight && sign_out, CMH_ERR_ARG, "cmh_hash_head_pack: NULL pointer");
    const size_t smem = ((size_t)hidden * 2 * bits + 2 * bits + (size_t)HEAD_WARPS * HEAD_ROWS * hidden) * sizeof(float);
    CMH_REQUIRE(smem <= 227 * 1024, CMH_ERR_UNSUPPORTED,
                "cmh_hash_head_pack: hidden=%d x bits=%d needs %zu bytes of shared memory (max 227 KB; 128 x 128 fits)", hidden, bits,
                smem);
    cudaStream_t st = (cudaStream_t)stream;
    const int w32 = 2 * ((bits + 63) / 64);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, HEAD_WARPS * HEAD_ROWS), (int64_t)sm_count()));
    uint32_t* s32 = reinterpret_cast<uint32_t*>(sign_out);
    uint32_t* v32 = reinterpret_cast<uint32_t*>(valid_out);
    const long long* idx = reinterpret_cast<const long long*>(index);
#define CMH_HEAD_CASE(DT, T)                                                                                              \
    case DT:                                                                                                              \
        CMH_CUDA(cudaFuncSetAttribute(hash_head_pack_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hash_head_pack_kernel<T><<<grid, HEAD_WARPS * 32, smem, st>>>((const T*)x, n, hidden, ld, relu, weight, bias, bits, idx, \
                                                                      n_out, w32, s32, v32, counters);                   \
        break;
    switch (dtype) {
        CMH_HEAD_CASE(CMH_F32, float)
        CMH_HEAD_CASE(CMH_F16, __half)
        CMH_HEAD_CASE(CMH_BF16, __nv_bfloat16)
        default:
            set_error("cmh_hash_head_pack: unsupported dtype %d (float32 / float16 / bfloat16 activations)", dtype);
            return CMH_ERR_UNSUPPORTED;
    }
#undef CMH_HEAD_CASE
    CMH_LAUNCH_CHECK("hash_head_pack_kernel");
    return CMH_OK;
}
