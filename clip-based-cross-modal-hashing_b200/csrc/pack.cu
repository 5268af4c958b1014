// K1 - sign + bit-pack.  Replaces the float32 code buffers that train/base.py:141-146 fills with torch.sign
// output (and the float multi-hot labels of dataset/base.py:89-94) by packed 64-bit words:
//   sign plane  bit = (x > 0)      valid plane  bit = (x != 0)      label mask  bit = (L != 0)
// HBM-bound: 4*n*bits bytes in, n*ceil(bits/64)*8 (x2 with the valid plane) out.
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace cmh {

template <typename T>
struct Cls {
    __device__ static __forceinline__ void get(T v, bool& pos, bool& zero, bool& unit) {
        float f = (float)v;
        pos = f > 0.f; zero = f == 0.f; unit = fabsf(f) == 1.f;
    }
};
template <>
struct Cls<double> {
    __device__ static __forceinline__ void get(double v, bool& pos, bool& zero, bool& unit) {
        pos = v > 0.0; zero = v == 0.0; unit = fabs(v) == 1.0;
    }
};
template <>
struct Cls<long long> {
    __device__ static __forceinline__ void get(long long v, bool& pos, bool& zero, bool& unit) {
        pos = v > 0; zero = v == 0; unit = (v == 1 || v == -1);
    }
};
template <>
struct Cls<__half> {
    __device__ static __forceinline__ void get(__half v, bool& pos, bool& zero, bool& unit) {
        Cls<float>::get(__half2float(v), pos, zero, unit);
    }
};
template <>
struct Cls<__nv_bfloat16> {
    __device__ static __forceinline__ void get(__nv_bfloat16 v, bool& pos, bool& zero, bool& unit) {
        Cls<float>::get(__bfloat162float(v), pos, zero, unit);
    }
};

__device__ __forceinline__ void warp_count_flush(unsigned long long a, unsigned long long b,
                                                 unsigned long long* counters) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0 && counters != nullptr) {
        if (a) atomicAdd(&counters[0], a);
        if (b) atomicAdd(&counters[1], b);
    }
}

// ---- fast path: float32, contiguous rows, bits % 32 == 0 -------------------------------------------------------
// The matrix is a flat stream of floats: 32 consecutive floats are one 32-bit output word, and a warp turns them into
// that word with ONE vote per plane (lane = element).  PACK_UNROLL independent, fully coalesced 128-byte loads per lane
// are in flight before the first is used; lane u keeps word u and lanes 0..7 store 32 contiguous bytes per plane.
// ~0.3 warp instructions per float (r02a capture of the float4 / shuffle version this replaces: 0.9, ALU pipe 75 %
// busy at 54 % of the HBM rate - the per-element 64-bit counters and a 64-bit division per stored word).
constexpr int PACK_UNROLL = 16;

template <bool SAME_STRIDE>   // words per row in == words per row out (bits % 64 == 0): the flat word index is the output index
__global__ void __launch_bounds__(256) pack_codes_f32_fast(const float* __restrict__ x, int64_t n_elems,
                                                           int w32_per_row_in,   // bits / 32
                                                           int w32_per_row_out,  // words * 2
                                                           uint32_t* __restrict__ sign32,
                                                           uint32_t* __restrict__ valid32,
                                                           unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_words = n_elems >> 5;            // n_elems is a multiple of 32 here
    uint32_t n_zero = 0, n_odd = 0;                  // per thread: < 2^32 elements each
    for (int64_t w0 = warp * PACK_UNROLL; w0 < n_words; w0 += n_warps * PACK_UNROLL) {
        float v[PACK_UNROLL];
#pragma unroll
        for (int u = 0; u < PACK_UNROLL; ++u)
            v[u] = w0 + u < n_words ? __ldcs(x + ((w0 + u) << 5) + lane) : 1.f;
        uint32_t sw = 0, vw = 0xffffffffu;
        // Fast trip: what `torch.sign` / the argmax heads emit is +-1 almost everywhere, so only the sign plane is voted
        // (compare + vote + select per element) while one predicate per lane remembers whether everything it saw was
        // +-1; a single vote afterwards sends the 16 words through the full classification (zeros, other values) only
        // when something else turned up.  (r02p capture of the version that voted both planes: ALU pipe 89 % busy.)
        bool plain = true;
#pragma unroll
        for (int u = 0; u < PACK_UNROLL; ++u) {
            const uint32_t s = __ballot_sync(0xffffffffu, v[u] > 0.f);
            plain = plain && fabsf(v[u]) == 1.f;
            if (lane == u) sw = s;
        }
        if (!__all_sync(0xffffffffu, plain)) {
#pragma unroll
            for (int u = 0; u < PACK_UNROLL; ++u) {
                const bool nz = v[u] != 0.f;
                const uint32_t z = __ballot_sync(0xffffffffu, nz);
                n_odd += (nz && fabsf(v[u]) != 1.f) ? 1u : 0u;
                if (lane == u) vw = z;
            }
        }
        if (lane < PACK_UNROLL && w0 + lane < n_words) {
            n_zero += (uint32_t)__popc(~vw);
            const int64_t w_flat = w0 + lane;
            int64_t o = w_flat;
            bool last_odd = false;
            if (!SAME_STRIDE) {
                const int64_t row = w_flat / w32_per_row_in;
                const int w = (int)(w_flat - row * w32_per_row_in);
                o = row * w32_per_row_out + w;
                last_odd = w == w32_per_row_in - 1 && (w32_per_row_in & 1);
            }
            sign32[o] = sw;
            if (valid32) valid32[o] = vw;
            if (last_odd) {                          // clear the unused high half-word
                sign32[o + 1] = 0u;
                if (valid32) valid32[o + 1] = 0u;
            }
        }
    }
    warp_count_flush((unsigned long long)n_zero, (unsigned long long)n_odd, counters);
}

// ---- fast path for labels: float32, contiguous rows of <= 32 columns (21 / 24 labels: NUS-WIDE, MIRFlickr) -----------
// 32 rows = 32 * ncols floats = ncols flat 32-bit words (one vote each, lane w keeps word w); lane j then cuts row j -
// bits [j * ncols, (j + 1) * ncols) of the flat stream - out of two neighbouring words with a funnel shift and stores
// its 64-bit mask (256 contiguous bytes per warp).
__global__ void __launch_bounds__(256) pack_labels_f32_fast(const float* __restrict__ x, int64_t n, int ncols,
                                                            uint64_t* __restrict__ out, unsigned long long* neg_counter) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_elems = n * ncols;
    const uint32_t row_mask = ncols >= 32 ? 0xffffffffu : ((1u << ncols) - 1u);
    uint32_t n_neg = 0;
    for (int64_t r0 = warp * 32; r0 < n; r0 += n_warps * 32) {
        const int64_t e0 = r0 * ncols;
        uint32_t mine = 0;
        for (int w = 0; w < ncols; w += 16) {         // up to 16 independent 128-byte loads per lane in flight
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int64_t e = e0 + (int64_t)(w + u) * 32 + lane;
                v[u] = (w + u < ncols && e < n_elems) ? __ldcs(x + e) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const uint32_t word = __ballot_sync(0xffffffffu, v[u] != 0.f);
                n_neg += v[u] < 0.f ? 1u : 0u;
                if (lane == w + u) mine = word;
            }
        }
        const int bit = lane * ncols, w0 = bit >> 5, sh = bit & 31;
        const uint32_t lo = __shfl_sync(0xffffffffu, mine, w0);
        const uint32_t hi = __shfl_sync(0xffffffffu, mine, min(w0 + 1, 31));
        const uint32_t mask = __funnelshift_r(lo, hi, sh) & row_mask;
        if (r0 + lane < n) out[r0 + lane] = (uint64_t)mask;
    }
    warp_count_flush((unsigned long long)n_neg, 0ull, neg_counter ? neg_counter : nullptr);
}

// ---- generic path: any dtype / leading dimension / width; one warp per row, lane = column, ballot packs --------
constexpr int PACK_ROWS = 4;  // rows in flight per warp

template <typename T, bool LABELS>
__global__ void __launch_bounds__(256) pack_rows_generic(const T* __restrict__ x, int64_t n, int ncols, int64_t ld,
                                                         int w32_out, uint32_t* __restrict__ sign32,
                                                         uint32_t* __restrict__ valid32,
                                                         unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n_seg = (ncols + 31) >> 5;
    unsigned long long c0 = 0, c1 = 0;
    for (int64_t r0 = warp * PACK_ROWS; r0 < n; r0 += n_warps * PACK_ROWS) {
        for (int seg = 0; seg < w32_out; ++seg) {
            const int col = seg * 32 + lane;
            T v[PACK_ROWS];
#pragma unroll
            for (int u = 0; u < PACK_ROWS; ++u) {
                const int64_t r = r0 + u;
                v[u] = (seg < n_seg && col < ncols && r < n) ? x[r * ld + col] : T(0.f);
            }
#pragma unroll
            for (int u = 0; u < PACK_ROWS; ++u) {
                const int64_t r = r0 + u;
                bool pos, zero, unit;
                Cls<T>::get(v[u], pos, zero, unit);
                const bool live = seg < n_seg && col < ncols && r < n;
                uint32_t sw, vw;
                if (LABELS) {
                    sw = __ballot_sync(0xffffffffu, live && !zero);
                    vw = 0;
                    c0 += live && !pos && !zero;  // negative label entry
                } else {
                    sw = __ballot_sync(0xffffffffu, live && pos);
                    vw = __ballot_sync(0xffffffffu, live && !zero);
                    c0 += live && zero;
                    c1 += live && !zero && !unit;
                }
                if (lane == 0 && r < n) {
                    sign32[r * w32_out + seg] = sw;
                    if (!LABELS && valid32) valid32[r * w32_out + seg] = vw;
                }
            }
        }
    }
    warp_count_flush(c0, c1, counters);
}

// ---- binarise at the source: sign / argmax + pack + scatter by dataset index ---------------------------------------
// What get_code does per batch (train/base.py:141-146: torch.sign of the encoder output, rows stored at the loader's
// `index`) and its DCHMT variant (train/base.py:150-158: argmax over [n, bits, 2] logits, class 0 -> -1), written
// straight into the packed planes: the float [N, bits] buffers never exist.  One warp per row, lane = column.
//   MODE 0: value x      -> sign bit x > 0, valid bit x != 0 (torch.sign(0) == 0), zeros counted
//   MODE 1: logits (a,b) -> sign bit b > a (argmax, first maximum wins: a tie is class 0 = -1), valid bit 1
template <typename T, int MODE>
__global__ void __launch_bounds__(256) pack_scatter_kernel(const T* __restrict__ x, int64_t n, int bits, int64_t ld,
                                                           const long long* __restrict__ index, int64_t n_out,
                                                           int w32_out, uint32_t* __restrict__ sign32,
                                                           uint32_t* __restrict__ valid32,
                                                           unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_zero = 0, n_bad = 0;
    for (int64_t r = warp; r < n; r += n_warps) {
        const int64_t dst = index ? (int64_t)index[r] : r;
        const bool ok = dst >= 0 && dst < n_out;
        if (!ok) { n_bad += lane == 0; continue; }
        for (int seg = 0; seg < w32_out; ++seg) {
            const int col = seg * 32 + lane;
            const bool live = col < bits;
            bool pos = false, zero = false;
            if (live) {
                if (MODE == 0) {
                    bool unit;
                    Cls<T>::get(x[r * ld + col], pos, zero, unit);
                } else {
                    const float a = (float)x[r * ld + 2 * (int64_t)col], b = (float)x[r * ld + 2 * (int64_t)col + 1];
                    pos = b > a;
                }
            }
            const uint32_t sw = __ballot_sync(0xffffffffu, live && pos);
            const uint32_t vw = __ballot_sync(0xffffffffu, live && !zero);
            n_zero += live && zero;
            if (lane == 0) {
                sign32[dst * w32_out + seg] = sw;
                if (valid32) valid32[dst * w32_out + seg] = vw;
            }
        }
    }
    warp_count_flush(n_zero, n_bad, counters);
}

// ---- unpack: packed planes -> float32 {-1, 0, +1} codes (the .mat export of train/base.py:307-349 stores float arrays) -----
__global__ void __launch_bounds__(256) unpack_codes_kernel(const uint32_t* __restrict__ sign32,
                                                           const uint32_t* __restrict__ valid32, int64_t n, int bits,
                                                           int w32, float* __restrict__ out, int64_t ld) {
    const int64_t total = n * (int64_t)bits;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / bits;
        const int c = (int)(i - r * bits);
        const uint32_t sw = sign32[r * w32 + (c >> 5)];
        const uint32_t vw = valid32 ? valid32[r * w32 + (c >> 5)] : 0xffffffffu;
        const bool v = (vw >> (c & 31)) & 1u, p = (sw >> (c & 31)) & 1u;
        out[r * ld + c] = v ? (p ? 1.f : -1.f) : 0.f;
    }
}

__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) synth_codes_kernel(uint64_t base, int64_t row0, int64_t n, int words,
                                                          uint64_t tail_mask, uint64_t* __restrict__ out) {
    const int64_t total = n * words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words;
        const int w = (int)(i - r * words);
        uint64_t v = splitmix64_dev(base + (uint64_t)((row0 + r) * words + w));
        if (w == words - 1) v &= tail_mask;
        out[i] = v;
    }
}

template <bool LABELS>
static int launch_generic(const void* x, int dtype, int64_t n, int ncols, int64_t ld, int w32_out, uint32_t* s32,
                          uint32_t* v32, unsigned long long* counters, cudaStream_t st) {
    const int block = 256;
    const int64_t warps_needed = ceil_div(n, PACK_ROWS);
    int grid = (int)std::min<int64_t>(ceil_div(warps_needed * 32, block), (int64_t)sm_count() * 16);
    if (grid < 1) grid = 1;
#define CMH_PACK_CASE(DT, T)                                                                                     \
    case DT:                                                                                                     \
        pack_rows_generic<T, LABELS><<<grid, block, 0, st>>>((const T*)x, n, ncols, ld, w32_out, s32, v32, counters); \
        break;
    switch (dtype) {
        CMH_PACK_CASE(CMH_F32, float)
        CMH_PACK_CASE(CMH_F16, __half)
        CMH_PACK_CASE(CMH_BF16, __nv_bfloat16)
        CMH_PACK_CASE(CMH_F64, double)
        CMH_PACK_CASE(CMH_I8, signed char)
        CMH_PACK_CASE(CMH_I32, int)
        CMH_PACK_CASE(CMH_I64, long long)
        CMH_PACK_CASE(CMH_U8, unsigned char)
        default:
            set_error("unsupported dtype %d", dtype);
            return CMH_ERR_UNSUPPORTED;
    }
#undef CMH_PACK_CASE
    CMH_LAUNCH_CHECK("pack_rows_generic");
    return CMH_OK;
}

}  // namespace cmh

using namespace cmh;

extern "C" int cmh_pack_codes(const void* x, int dtype, int64_t n, int bits, int64_t ld, uint64_t* sign_out,
                              uint64_t* valid_out, unsigned long long* counters, void* stream) {
    CMH_REQUIRE(n >= 0 && bits > 0 && ld >= bits, CMH_ERR_ARG, "cmh_pack_codes: bad shape n=%lld bits=%d ld=%lld",
                (long long)n, bits, (long long)ld);
    CMH_REQUIRE(bits <= CMH_MAX_BITS, CMH_ERR_UNSUPPORTED, "cmh_pack_codes: bits=%d > %d", bits, CMH_MAX_BITS);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(x && sign_out, CMH_ERR_ARG, "cmh_pack_codes: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int words = (bits + 63) / 64;
    const bool fast = dtype == CMH_F32 && ld == bits && (bits % 32) == 0 && (((uintptr_t)x) & 3) == 0;
    if (fast) {
        const int64_t n_elems = n * bits;
        const int block = 256;
        // 8 resident CTAs per SM (2048 threads): 8 x 128-byte loads per lane keep ~64 KB per SM in flight
        int grid = (int)std::min<int64_t>(ceil_div(n_elems / 32, (int64_t)(block / 32) * PACK_UNROLL), (int64_t)sm_count() * 6);
        if (grid < 1) grid = 1;
        if (bits % 64 == 0)
            pack_codes_f32_fast<true><<<grid, block, 0, st>>>((const float*)x, n_elems, bits / 32, words * 2,
                                                              (uint32_t*)sign_out, (uint32_t*)valid_out, counters);
        else
            pack_codes_f32_fast<false><<<grid, block, 0, st>>>((const float*)x, n_elems, bits / 32, words * 2,
                                                               (uint32_t*)sign_out, (uint32_t*)valid_out, counters);
        CMH_LAUNCH_CHECK("pack_codes_f32_fast");
        return CMH_OK;
    }
    return launch_generic<false>(x, dtype, n, bits, ld, words * 2, (uint32_t*)sign_out, (uint32_t*)valid_out,
                                 counters, st);
}

extern "C" int cmh_pack_scatter(const void* x, int dtype, int64_t n, int bits, int64_t ld, int mode, const int64_t* index,
                                int64_t n_out, uint64_t* sign_out, uint64_t* valid_out, unsigned long long* counters,
                                void* stream) {
    CMH_REQUIRE(n >= 0 && bits > 0 && n_out >= 0 && (mode == 0 || mode == 1) && ld >= (int64_t)bits * (mode ? 2 : 1),
                CMH_ERR_ARG, "cmh_pack_scatter: bad shape n=%lld bits=%d ld=%lld mode=%d", (long long)n, bits, (long long)ld, mode);
    CMH_REQUIRE(bits <= CMH_MAX_BITS, CMH_ERR_UNSUPPORTED, "cmh_pack_scatter: bits=%d > %d", bits, CMH_MAX_BITS);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(x && sign_out, CMH_ERR_ARG, "cmh_pack_scatter: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int w32 = 2 * ((bits + 63) / 64);
    const int block = 256;
    int grid = (int)std::min<int64_t>(ceil_div(n * 32, block), (int64_t)sm_count() * 16);
    if (grid < 1) grid = 1;
    uint32_t* s32 = reinterpret_cast<uint32_t*>(sign_out);
    uint32_t* v32 = reinterpret_cast<uint32_t*>(valid_out);
    const long long* idx = reinterpret_cast<const long long*>(index);
#define CMH_SCATTER_CASE(DT, T)                                                                                          \
    case DT:                                                                                                             \
        if (mode == 0)                                                                                                   \
            pack_scatter_kernel<T, 0><<<grid, block, 0, st>>>((const T*)x, n, bits, ld, idx, n_out, w32, s32, v32, counters); \
        else                                                                                                             \
            pack_scatter_kernel<T, 1><<<grid, block, 0, st>>>((const T*)x, n, bits, ld, idx, n_out, w32, s32, v32, counters); \
        break;
    switch (dtype) {
        CMH_SCATTER_CASE(CMH_F32, float)
        CMH_SCATTER_CASE(CMH_F16, __half)
        CMH_SCATTER_CASE(CMH_BF16, __nv_bfloat16)
        CMH_SCATTER_CASE(CMH_F64, double)
        default:
            set_error("cmh_pack_scatter: unsupported dtype %d (floating-point activations only)", dtype);
            return CMH_ERR_UNSUPPORTED;
    }
#undef CMH_SCATTER_CASE
    CMH_LAUNCH_CHECK("pack_scatter_kernel");
    return CMH_OK;
}

extern "C" int cmh_unpack_codes(const uint64_t* sign, const uint64_t* valid, int64_t n, int bits, float* out, int64_t ld,
                                void* stream) {
    CMH_REQUIRE(n >= 0 && bits > 0 && bits <= CMH_MAX_BITS && ld >= bits, CMH_ERR_ARG,
                "cmh_unpack_codes: bad shape n=%lld bits=%d ld=%lld", (long long)n, bits, (long long)ld);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(sign && out, CMH_ERR_ARG, "cmh_unpack_codes: NULL pointer");
    const int64_t total = n * (int64_t)bits;
    const int grid = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 16);
    unpack_codes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(sign),
                                                              reinterpret_cast<const uint32_t*>(valid), n, bits,
                                                              2 * ((bits + 63) / 64), out, ld);
    CMH_LAUNCH_CHECK("unpack_codes_kernel");
    return CMH_OK;
}

extern "C" int cmh_pack_labels(const void* L, int dtype, int64_t n, int nlab, int64_t ld, uint64_t* out,
                               unsigned long long* neg_counter, void* stream) {
    CMH_REQUIRE(n >= 0 && nlab > 0 && ld >= nlab, CMH_ERR_ARG, "cmh_pack_labels: bad shape n=%lld nlab=%d ld=%lld",
                (long long)n, nlab, (long long)ld);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(L && out, CMH_ERR_ARG, "cmh_pack_labels: NULL pointer");
    const int lwords = (nlab + 63) / 64;
    if (dtype == CMH_F32 && ld == nlab && nlab <= 32 && (((uintptr_t)L) & 3) == 0) {
        const int block = 256;
        int grid = (int)std::min<int64_t>(ceil_div(ceil_div(n, 32), block / 32), (int64_t)sm_count() * 8);
        if (grid < 1) grid = 1;
        pack_labels_f32_fast<<<grid, block, 0, (cudaStream_t)stream>>>((const float*)L, n, nlab, out, neg_counter);
        CMH_LAUNCH_CHECK("pack_labels_f32_fast");
        return CMH_OK;
    }
    return launch_generic<true>(L, dtype, n, nlab, ld, lwords * 2, (uint32_t*)out, nullptr, neg_counter,
                                (cudaStream_t)stream);
}

extern "C" int cmh_synth_codes(uint64_t seed, int64_t row0, int64_t n, int bits, uint64_t* out, void* stream) {
    CMH_REQUIRE(n >= 0 && bits > 0 && row0 >= 0, CMH_ERR_ARG, "cmh_synth_codes: bad arguments");
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(out, CMH_ERR_ARG, "cmh_synth_codes: NULL pointer");
    const int words = (bits + 63) / 64;
    const int tail = bits - 64 * (words - 1);
    const uint64_t tail_mask = tail >= 64 ? ~0ull : ((1ull << tail) - 1ull);
    const int block = 256;
    int grid = (int)std::min<int64_t>(ceil_div(n * words, block), (int64_t)sm_count() * 16);
    synth_codes_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(seed * 0x100000001B3ull, row0, n, words, tail_mask, out);
    CMH_LAUNCH_CHECK("synth_codes_kernel");
    return CMH_OK;
}
