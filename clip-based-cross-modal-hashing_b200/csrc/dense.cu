// Dense blocks: calc_hammingDist (utils/calc_utils.py:8-13) and calc_neighbor (:4-5,:42-45) on packed rows.
// Both are bound by the float32 output they must materialise (4 B per pair to HBM), not by the compare.
#include <algorithm>

#include "common.cuh"

namespace cmh {

constexpr int DENSE_QROWS = 8;    // query rows per CTA (their words are broadcast from shared memory)
constexpr int DENSE_THREADS = 256;

// out[i][j] = 0.5 * (bits - dot),  dot = popc(vq & vd) - 2 * popc((q ^ d) & vq & vd)
template <bool TERN>
__global__ void __launch_bounds__(DENSE_THREADS) hamming_dense_kernel(
    const uint64_t* __restrict__ qs, const uint64_t* __restrict__ qv, int64_t nq, const uint64_t* __restrict__ ds,
    const uint64_t* __restrict__ dv, int64_t nd, int words, int bits, float* __restrict__ out, int64_t ld_out) {
    extern __shared__ uint64_t sq[];  // [DENSE_QROWS][words] sign, then [DENSE_QROWS][words] valid
    const int64_t q0 = (int64_t)blockIdx.y * DENSE_QROWS;
    const int nrows = (int)min((int64_t)DENSE_QROWS, nq - q0);
    for (int i = threadIdx.x; i < nrows * words; i += blockDim.x) {
        sq[i] = qs[q0 * words + i];
        if (TERN) sq[DENSE_QROWS * words + i] = qv ? qv[q0 * words + i] : ~0ull;
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nd) return;
    int acc[DENSE_QROWS];
#pragma unroll
    for (int r = 0; r < DENSE_QROWS; ++r) acc[r] = 0;
    for (int w = 0; w < words; ++w) {
        const uint64_t dw = ds[j * words + w];
        uint64_t dvw = ~0ull;
        if (TERN) {
            // padding bits are 0 in every sign plane and in every explicit valid plane; an implicit (NULL)
            // valid plane is all-ones only over real bit positions
            const int live = bits - 64 * w;
            const uint64_t full = live >= 64 ? ~0ull : ((1ull << live) - 1ull);
            dvw = dv ? dv[j * words + w] : full;
        }
#pragma unroll
        for (int r = 0; r < DENSE_QROWS; ++r) {
            if (r < nrows) {
                const uint64_t x = sq[r * words + w] ^ dw;
                if (TERN) {
                    const uint64_t both = sq[DENSE_QROWS * words + r * words + w] & dvw;
                    acc[r] += 2 * __popcll(x & both) - __popcll(both);  // = -dot contribution
                } else {
                    acc[r] += __popcll(x);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < DENSE_QROWS; ++r) {
        if (r < nrows) {
            const float v = TERN ? 0.5f * (float)(bits + acc[r]) : (float)acc[r];
            __stcs(out + (q0 + r) * ld_out + j, v);
        }
    }
}

__global__ void __launch_bounds__(DENSE_THREADS) neighbor_dense_kernel(const uint64_t* __restrict__ la, int64_t na,
                                                                      const uint64_t* __restrict__ lb, int64_t nb,
                                                                      int lwords, float* __restrict__ out,
                                                                      int64_t ld_out) {
    extern __shared__ uint64_t sq[];
    const int64_t q0 = (int64_t)blockIdx.y * DENSE_QROWS;
    const int nrows = (int)min((int64_t)DENSE_QROWS, na - q0);
    for (int i = threadIdx.x; i < nrows * lwords; i += blockDim.x) sq[i] = la[q0 * lwords + i];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    uint64_t hit[DENSE_QROWS];
#pragma unroll
    for (int r = 0; r < DENSE_QROWS; ++r) hit[r] = 0;
    for (int w = 0; w < lwords; ++w) {
        const uint64_t bw = lb[j * lwords + w];
#pragma unroll
        for (int r = 0; r < DENSE_QROWS; ++r)
            if (r < nrows) hit[r] |= sq[r * lwords + w] & bw;
    }
#pragma unroll
    for (int r = 0; r < DENSE_QROWS; ++r)
        if (r < nrows) __stcs(out + (q0 + r) * ld_out + j, hit[r] ? 1.0f : 0.0f);
}

}  // namespace cmh

using namespace cmh;

extern "C" int cmh_hamming_dense(const cmh_codeset* q, const cmh_codeset* d, int bits, float* out, int64_t ld_out,
                                 void* stream) {
    CMH_REQUIRE(q && d, CMH_ERR_ARG, "cmh_hamming_dense: NULL codeset");
    CMH_REQUIRE(bits > 0 && bits <= CMH_MAX_BITS, CMH_ERR_UNSUPPORTED, "cmh_hamming_dense: bits=%d", bits);
    CMH_REQUIRE(q->n >= 0 && d->n >= 0 && ld_out >= d->n, CMH_ERR_ARG, "cmh_hamming_dense: bad sizes");
    if (q->n == 0 || d->n == 0) return CMH_OK;
    CMH_REQUIRE(q->sign && d->sign && out, CMH_ERR_ARG, "cmh_hamming_dense: NULL pointer");
    const int words = (bits + 63) / 64;
    const bool tern = q->valid || d->valid;
    dim3 grid((unsigned)ceil_div(d->n, DENSE_THREADS), (unsigned)ceil_div(q->n, DENSE_QROWS));
    CMH_REQUIRE(grid.y <= 65535, CMH_ERR_UNSUPPORTED, "cmh_hamming_dense: too many query rows per call (%lld)",
                (long long)q->n);
    const size_t smem = (size_t)DENSE_QROWS * words * 8 * (tern ? 2 : 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (tern)
        hamming_dense_kernel<true><<<grid, DENSE_THREADS, smem, st>>>(q->sign, q->valid, q->n, d->sign, d->valid, d->n,
                                                                     words, bits, out, ld_out);
    else
        hamming_dense_kernel<false><<<grid, DENSE_THREADS, smem, st>>>(q->sign, nullptr, q->n, d->sign, nullptr, d->n,
                                                                      words, bits, out, ld_out);
    CMH_LAUNCH_CHECK("hamming_dense_kernel");
    return CMH_OK;
}

extern "C" int cmh_neighbor_dense(const uint64_t* la, int64_t na, const uint64_t* lb, int64_t nb_rows, int lwords,
                                  float* out, int64_t ld_out, void* stream) {
    CMH_REQUIRE(na >= 0 && nb_rows >= 0 && lwords > 0 && ld_out >= nb_rows, CMH_ERR_ARG, "cmh_neighbor_dense: bad sizes");
    if (na == 0 || nb_rows == 0) return CMH_OK;
    CMH_REQUIRE(la && lb && out, CMH_ERR_ARG, "cmh_neighbor_dense: NULL pointer");
    dim3 grid((unsigned)ceil_div(nb_rows, DENSE_THREADS), (unsigned)ceil_div(na, DENSE_QROWS));
    CMH_REQUIRE(grid.y <= 65535, CMH_ERR_UNSUPPORTED, "cmh_neighbor_dense: too many rows per call (%lld)", (long long)na);
    neighbor_dense_kernel<<<grid, DENSE_THREADS, (size_t)DENSE_QROWS * lwords * 8, (cudaStream_t)stream>>>(
        la, na, lb, nb_rows, lwords, out, ld_out);
    CMH_LAUNCH_CHECK("neighbor_dense_kernel");
    return CMH_OK;
}
