// Host side of the ranking-by-counting path: launch plan, workspace carve-up, the small scan / reduce / finalise
// kernels between the two heavy passes, and the extern "C" entry points declared in include/cmh_b200.h.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "eval_common.cuh"

namespace cmh {

// launchers implemented in eval_tile.cu / eval_warp.cu
size_t tile_smem_bytes(const EvalArgs& a, bool tern, int kind);
int launch_hist_tile(const EvalArgs& a, bool tern, uint32_t* chunk_hist, cudaStream_t st);
int launch_rank_tile(const EvalArgs& a, bool tern, const uint2* base, const uint32_t* total, const TopnList& topn,
                     double* ap_part, uint32_t* hits_part, cudaStream_t st);
int launch_select_tile(const EvalArgs& a, bool tern, const uint2* base, const int32_t* thr, uint64_t* keys,
                       cudaStream_t st);
int warp_queries_per_cta(const EvalArgs& a, bool tern);
size_t warp_smem_bytes(const EvalArgs& a, bool tern, int wpc);
int launch_hist_warp(const EvalArgs& a, bool tern, int wpc, uint32_t* chunk_hist, cudaStream_t st);
int launch_rank_warp(const EvalArgs& a, bool tern, int wpc, const uint2* base, const uint32_t* total,
                     const TopnList& topn, double* ap_part, uint32_t* hits_part, cudaStream_t st);
int launch_select_warp(const EvalArgs& a, bool tern, int wpc, const uint2* base, const int32_t* thr, uint64_t* keys,
                       cudaStream_t st);
// eval_lane.cu: warp-per-query, lane-per-row kernels for binary codes up to 128 bits (layout W, 8 queries per CTA)
bool lane_supported(const EvalArgs& a, bool tern);
size_t lane_smem_bytes(const EvalArgs& a, int kind);
int launch_hist_lane(const EvalArgs& a, uint32_t* chunk_hist, cudaStream_t st);
int launch_rank_lane(const EvalArgs& a, const uint2* base, const uint32_t* total, const TopnList& topn, double* ap_part,
                     uint32_t* hits_part, cudaStream_t st);
constexpr int LANE_Q_TILE = 8;

constexpr size_t MAX_DYN_SMEM = 227 * 1024;

constexpr int MAX_CHUNK_ROWS = 65520;  // pass-1 counters are 16 bit; multiple of 16 keeps bulk copies aligned
constexpr int MIN_CHUNK_ROWS = 1024;   // amortises the per-CTA counter prologue / epilogue

static EvalArgs geometry_args(int bits, int nlab, bool ternary) {
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.bits = bits;
    a.nb = ternary ? 2 * bits + 1 : bits + 1;
    a.cw = (bits + 31) / 32;
    a.cw_stride = 2 * ((bits + 63) / 64);
    a.lw = (nlab + 31) / 32;
    a.lw_stride = 2 * ((nlab + 63) / 64);
    a.ntopn = 0;
    a.kq = a.kd = 1;
    return a;
}

Workspace carve_workspace(const cmh_plan& p, void* ws) {
    Workspace w;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) {
        unsigned char* r = base ? base + off : nullptr;
        off += (bytes + 255) & ~(uint64_t)255;
        return r;
    };
    const uint64_t cells = (uint64_t)p.n_chunks * p.nb * p.nq_pad;
    w.chunk_hist = reinterpret_cast<uint32_t*>(take(cells * 4));
    w.base = reinterpret_cast<uint2*>(take(cells * 8));
    w.shard_all = reinterpret_cast<uint32_t*>(take((uint64_t)p.nq * p.nb * 4));
    w.shard_rel = reinterpret_cast<uint32_t*>(take((uint64_t)p.nq * p.nb * 4));
    w.total = reinterpret_cast<uint32_t*>(take((uint64_t)p.nq_pad * 4));
    w.thr = reinterpret_cast<int32_t*>(take((uint64_t)p.nq_pad * 4));
    w.thr_quota = reinterpret_cast<uint32_t*>(take((uint64_t)p.nq_pad * 4));
    w.ap_part = reinterpret_cast<double*>(take((uint64_t)p.n_chunks * p.nq_pad * 8));
    w.hits_part = reinterpret_cast<uint32_t*>(take((uint64_t)p.n_chunks * p.max_topn * p.nq_pad * 4));
    w.bytes = off;
    return w;
}

static int fill_args(const cmh_plan& p, const cmh_codeset* q, const cmh_codeset* d, bool need_labels, EvalArgs* out) {
    CMH_REQUIRE(q && d, CMH_ERR_ARG, "NULL codeset");
    CMH_REQUIRE(q->n == p.nq && d->n == p.nd, CMH_ERR_ARG, "codeset sizes (%lld, %lld) do not match the plan (%lld, %lld)",
                (long long)q->n, (long long)d->n, (long long)p.nq, (long long)p.nd);
    CMH_REQUIRE(q->sign && d->sign, CMH_ERR_ARG, "NULL sign plane");
    if (p.ternary) CMH_REQUIRE(q->valid && d->valid, CMH_ERR_ARG, "ternary plan needs both valid planes");
    if (need_labels) CMH_REQUIRE(p.lwords > 0 && q->labels && d->labels, CMH_ERR_ARG, "labels required");
    EvalArgs a = geometry_args(p.bits, p.nlab, p.ternary != 0);
    a.qs = reinterpret_cast<const uint32_t*>(q->sign);
    a.qv = reinterpret_cast<const uint32_t*>(q->valid);
    a.ql = reinterpret_cast<const uint32_t*>(q->labels);
    a.ds = reinterpret_cast<const uint32_t*>(d->sign);
    a.dv = reinterpret_cast<const uint32_t*>(d->valid);
    a.dl = reinterpret_cast<const uint32_t*>(d->labels);
    if (!(q->labels && d->labels)) { a.lw = 0; a.lw_stride = 0; a.ql = a.dl = nullptr; }
    a.nq = p.nq; a.nd = p.nd; a.nq_pad = p.nq_pad;
    a.chunk_rows = p.chunk_rows; a.n_chunks = p.n_chunks;
    a.kq = p.kq > 0 ? p.kq : 1; a.kd = p.kd > 0 ? p.kd : 1; a.ap_mode = p.ap_mode;
    const uintptr_t align = (uintptr_t)a.ds | (uintptr_t)a.dv | (uintptr_t)a.dl;
    a.bulk_ok = (align & 15) == 0;
    *out = a;
    return CMH_OK;
}

// ---- small kernels -----------------------------------------------------------------------------------------------

// shard totals [nq][nb] from the per-chunk histograms
template <bool LAYOUT_W>
__global__ void __launch_bounds__(256) reduce_hist_kernel(const EvalArgs a, const uint32_t* __restrict__ chunk_hist,
                                                          uint32_t* __restrict__ s_all, uint32_t* __restrict__ s_rel,
                                                          uint32_t* __restrict__ u_all, uint32_t* __restrict__ u_rel) {
    int64_t q; int b;
    if (LAYOUT_W) { b = blockIdx.x * blockDim.x + threadIdx.x; q = blockIdx.y; if (b >= a.nb) return; }
    else { q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b = blockIdx.y; }
    if (q >= a.nq) return;
    uint32_t sa = 0, sr = 0;
    for (int c = 0; c < a.n_chunks; ++c) {
        const uint32_t h = chunk_hist[LAYOUT_W ? hist_index_W(a, c, b, q) : hist_index_T(a, c, b, q)];
        sa += h & 0xffffu; sr += h >> 16;
    }
    const int64_t o = q * a.nb + b;
    s_all[o] = sa; s_rel[o] = sr;
    if (u_all) u_all[o] = sa;
    if (u_rel) u_rel[o] = sr;
}

struct ScanArgs {
    const uint32_t *g_all, *g_rel;   // [nq][nb] totals over all shards
    const uint32_t *l_all, *l_rel;   // [nq][nb] totals over lower shards (or NULL)
    int64_t k;                       // < 0: all
    int K;                           // > 0: also derive the top-K threshold bucket
    uint2* base; uint32_t* total; int64_t* n_rel; int32_t* thr;
};

// tile layout: block = 32 queries (lane) x 8 warps striding over buckets
__global__ void __launch_bounds__(256) scan_tile_kernel(const EvalArgs a, const uint32_t* __restrict__ chunk_hist,
                                                        const ScanArgs sa) {
    extern __shared__ uint32_t sm[];  // [nb][32] all, [nb][32] rel
    uint32_t* sA = sm; uint32_t* sR = sm + a.nb * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 32 + lane;
    const bool live = q < a.nq;
    for (int b = warp; b < a.nb; b += 8) {
        sA[b * 32 + lane] = live ? sa.g_all[q * a.nb + b] : 0u;
        sR[b * 32 + lane] = (live && sa.g_rel) ? sa.g_rel[q * a.nb + b] : 0u;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t ra = 0, rr = 0; int thr = -1;
        for (int b = 0; b < a.nb; ++b) {
            const uint32_t ca = sA[b * 32 + lane], cr = sR[b * 32 + lane];
            sA[b * 32 + lane] = ra; sR[b * 32 + lane] = rr;
            ra += ca; rr += cr;
            if (thr < 0 && sa.K > 0 && ra >= (uint32_t)sa.K) thr = b;
        }
        if (thr < 0) thr = a.nb - 1;
        if (q < a.nq_pad) {
            const int64_t kk = (sa.k < 0 || sa.k > (int64_t)rr) ? (int64_t)rr : sa.k;
            sa.total[q] = live ? (uint32_t)kk : 0u;
            if (sa.thr) sa.thr[q] = live ? thr : -1;
            if (live && sa.n_rel) sa.n_rel[q] = (int64_t)rr;
        }
    }
    __syncthreads();
    if (q >= a.nq_pad) return;
    for (int b = warp; b < a.nb; b += 8) {
        uint32_t ra = sA[b * 32 + lane] + ((live && sa.l_all) ? sa.l_all[q * a.nb + b] : 0u);
        uint32_t rr = sR[b * 32 + lane] + ((live && sa.l_rel) ? sa.l_rel[q * a.nb + b] : 0u);
        int c = 0;
        for (; c + 4 <= a.n_chunks; c += 4) {
            uint32_t h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) h[u] = chunk_hist[hist_index_T(a, c + u, b, q)];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sa.base[hist_index_T(a, c + u, b, q)] = make_uint2(ra, rr);
                ra += h[u] & 0xffffu; rr += h[u] >> 16;
            }
        }
        for (; c < a.n_chunks; ++c) {
            const uint32_t h = chunk_hist[hist_index_T(a, c, b, q)];
            sa.base[hist_index_T(a, c, b, q)] = make_uint2(ra, rr);
            ra += h & 0xffffu; rr += h >> 16;
        }
    }
}

// warp layout: one warp per query, lanes over buckets with a running carry
__global__ void __launch_bounds__(256) scan_warp_kernel(const EvalArgs a, const uint32_t* __restrict__ chunk_hist,
                                                        const ScanArgs sa) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= a.nq_pad) return;
    const bool live = q < a.nq;
    uint32_t carry_a = 0, carry_r = 0; int thr = -1;
    for (int b0 = 0; b0 < a.nb; b0 += 32) {
        const int b = b0 + lane;
        const bool in = b < a.nb;
        const uint32_t ca = (live && in) ? sa.g_all[q * a.nb + b] : 0u;
        const uint32_t cr = (live && in && sa.g_rel) ? sa.g_rel[q * a.nb + b] : 0u;
        uint32_t ia = ca, ir = cr;  // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
            if (lane >= o) { ia += ta; ir += tr; }
        }
        if (sa.K > 0 && thr < 0) {
            const uint32_t hit = __ballot_sync(0xffffffffu, in && carry_a + ia >= (uint32_t)sa.K);
            if (hit) thr = b0 + __ffs(hit) - 1;
        }
        if (in) {
            uint32_t ra = carry_a + ia - ca + ((live && sa.l_all) ? sa.l_all[q * a.nb + b] : 0u);
            uint32_t rr = carry_r + ir - cr + ((live && sa.l_rel) ? sa.l_rel[q * a.nb + b] : 0u);
            for (int c = 0; c < a.n_chunks; ++c) {
                const uint32_t h = chunk_hist[hist_index_W(a, c, b, q)];
                sa.base[hist_index_W(a, c, b, q)] = make_uint2(ra, rr);
                ra += h & 0xffffu; rr += h >> 16;
            }
        }
        carry_a += __shfl_sync(0xffffffffu, ia, 31);
        carry_r += __shfl_sync(0xffffffffu, ir, 31);
    }
    if (lane == 0) {
        if (thr < 0) thr = a.nb - 1;
        const int64_t kk = (sa.k < 0 || sa.k > (int64_t)carry_r) ? (int64_t)carry_r : sa.k;
        sa.total[q] = live ? (uint32_t)kk : 0u;
        if (sa.thr) sa.thr[q] = live ? thr : -1;
        if (live && sa.n_rel) sa.n_rel[q] = (int64_t)carry_r;
    }
}

// sum the per-chunk partial results in a fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_parts_kernel(const EvalArgs a, const double* __restrict__ ap_part,
                                                           const uint32_t* __restrict__ hits_part, const TopnList topn,
                                                           double* __restrict__ ap_sum, uint32_t* __restrict__ hits) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nq) return;
    double s = 0.0;
    for (int c = 0; c < a.n_chunks; ++c) s += ap_part[(int64_t)c * a.nq_pad + q];
    ap_sum[q] = s;
    uint32_t cum = 0;  // bins hold first crossings: hits@N_i = sum of the bins of all cutoffs <= N_i
    for (int i = 0; i < a.ntopn; ++i) {
        for (int c = 0; c < a.n_chunks; ++c) cum += hits_part[((int64_t)c * a.ntopn + i) * a.nq_pad + q];
        hits[q * a.ntopn + topn.perm[i]] = cum;
    }
}

__global__ void __launch_bounds__(256) fill_u64_kernel(uint64_t* p, int64_t n, uint64_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ double block_sum_1024(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

// AP = ap_sum / min(k, n_rel); mAP = sum AP / nq (label-free queries stay in the divisor, calc_utils.py:28-29,38)
__global__ void __launch_bounds__(1024) finalize_map_kernel(const double* __restrict__ ap_sum,
                                                            const int64_t* __restrict__ n_rel, int64_t nq, int64_t k,
                                                            double* __restrict__ ap, float* __restrict__ map) {
    __shared__ double sh[32];
    double local = 0.0;
    for (int64_t q = threadIdx.x; q < nq; q += blockDim.x) {
        const int64_t nr = n_rel[q];
        const int64_t tot = k < 0 ? nr : (k < nr ? k : nr);
        const double v = tot > 0 ? ap_sum[q] / (double)tot : 0.0;
        if (ap) ap[q] = v;
        local += v;
    }
    const double s = block_sum_1024(local, sh);
    if (threadIdx.x == 0) map[0] = nq > 0 ? (float)(s / (double)nq) : 0.f;
}

// textbook AP@k (train/DPSIH/_utils.py:22-29): AP = ap_sum / #relevant rows within the first k (0 when there is none)
__global__ void __launch_bounds__(1024) finalize_map_hits_kernel(const double* __restrict__ ap_sum,
                                                                 const uint32_t* __restrict__ hits, int ntopn, int64_t nq,
                                                                 double* __restrict__ ap, float* __restrict__ map) {
    __shared__ double sh[32];
    double local = 0.0;
    for (int64_t q = threadIdx.x; q < nq; q += blockDim.x) {
        const uint32_t h = hits[q * ntopn];
        const double v = h > 0 ? ap_sum[q] / (double)h : 0.0;
        if (ap) ap[q] = v;
        local += v;
    }
    const double s = block_sum_1024(local, sh);
    if (threadIdx.x == 0) map[0] = nq > 0 ? (float)(s / (double)nq) : 0.f;
}

__global__ void __launch_bounds__(1024) finalize_topn_kernel(const uint32_t* __restrict__ hits,
                                                             const int64_t* __restrict__ n_rel, int64_t nq, int ntopn,
                                                             const TopnList topn_unsorted, int64_t nd_total,
                                                             float* __restrict__ prec) {
    __shared__ double sh[32];
    const int i = blockIdx.x;
    double local = 0.0;  // integer-valued: exact below 2^53
    for (int64_t q = threadIdx.x; q < nq; q += blockDim.x)
        if (n_rel[q] > 0) local += (double)hits[q * ntopn + i];
    const double s = block_sum_1024(local, sh);
    if (threadIdx.x == 0) {
        const int64_t ni = (int64_t)topn_unsorted.n[i];
        const double n = (double)(ni < nd_total ? ni : nd_total);
        prec[i] = (nq > 0 && n > 0) ? (float)(s / n / (double)nq) : 0.f;
    }
}

// PR curve, stage A: one warp per 32 queries (lane = query) walks the radii; warp-reduced partial sums per radius.
__global__ void __launch_bounds__(256) pr_partial_kernel(const uint32_t* __restrict__ h_all,
                                                         const uint32_t* __restrict__ h_rel, int64_t nq, int bits,
                                                         int ternary, double* __restrict__ part /*[nw][bits+1][3]*/) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t q = w * 32 + lane;
    if (w * 32 >= nq) return;
    const int nb = ternary ? 2 * bits + 1 : bits + 1;
    const bool live = q < nq;
    double n_rel = 0.0;
    if (live)
        for (int b = 0; b < nb; ++b) n_rel += (double)h_rel[q * nb + b];
    double ca = 0.0, cr = 0.0;
    int b = 0;
    for (int r = 0; r <= bits; ++r) {
        const int b_end = ternary ? 2 * r : r;  // buckets with dist <= r
        for (; b <= b_end; ++b)
            if (live) { ca += (double)h_all[q * nb + b]; cr += (double)h_rel[q * nb + b]; }
        double p = 0.0, rc = 0.0, sup = 0.0;
        if (live && n_rel > 0.0) {
            p = cr / (ca > 0.0 ? ca : 0.1);
            rc = cr / n_rel;
            sup = p > 0.0 ? 1.0 : 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p += __shfl_xor_sync(0xffffffffu, p, o);
            rc += __shfl_xor_sync(0xffffffffu, rc, o);
            sup += __shfl_xor_sync(0xffffffffu, sup, o);
        }
        if (lane == 0) {
            double* o = part + (w * (bits + 1) + r) * 3;
            o[0] = p; o[1] = rc; o[2] = sup;
        }
    }
}

__global__ void __launch_bounds__(256) pr_final_kernel(const double* __restrict__ part, int64_t n_warps, int bits,
                                                       float* __restrict__ P, float* __restrict__ R) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > bits) return;
    double p = 0.0, rc = 0.0, sup = 0.0;
    for (int64_t w = 0; w < n_warps; ++w) {
        const double* o = part + (w * (bits + 1) + r) * 3;
        p += o[0]; rc += o[1]; sup += o[2];
    }
    if (sup == 0.0) sup = 0.1;
    P[r] = (float)(p / sup);
    R[r] = (float)(rc / sup);
}

// ---- top-K merge: rank of every candidate among all lists by binary search; unique keys -> unique ranks ----------
// The lists of a query are staged in shared memory when they fit (in_smem); beyond that (n_lists * K * 8 bytes above the
// budget, e.g. 8 shards x K = 4096) they are searched where they are - L2-resident, slower, but without a size limit.
__global__ void __launch_bounds__(256) topk_merge_kernel(const uint64_t* __restrict__ keys_in, int n_lists, int64_t nq,
                                                         int K, int in_smem, uint64_t* __restrict__ keys_out) {
    extern __shared__ uint64_t sk[];  // [n_lists][K] for this query
    const int64_t q = blockIdx.x;
    if (in_smem) {
        for (int i = threadIdx.x; i < n_lists * K; i += blockDim.x) {
            const int g = i / K, j = i - g * K;
            sk[i] = keys_in[((int64_t)g * nq + q) * K + j];
        }
        __syncthreads();
    }
    auto list_of = [&](int g) { return in_smem ? sk + g * K : keys_in + ((int64_t)g * nq + q) * K; };
    for (int i = threadIdx.x; i < n_lists * K; i += blockDim.x) {
        const int g = i / K, j = i - g * K;
        const uint64_t key = list_of(g)[j];
        if (key == ~0ull) continue;  // padding
        int rank = j;
        for (int o = 0; o < n_lists && rank < K; ++o) {
            if (o == g) continue;
            const uint64_t* lst = list_of(o);
            int lo = 0, hi = K;  // first position with lst[pos] >= key  (keys are unique across lists)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lst[mid] < key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < K) keys_out[q * K + rank] = key;
    }
}

}  // namespace cmh

using namespace cmh;

// =====================================================================================================================
// extern "C"
// =====================================================================================================================
// Which design ranks an (nq, nd, nb) problem when the caller does not say: thread-per-query tile kernels (0), generic
// warp-per-query kernels (1: any bucket count, ternary codes), lane kernels (2: binary codes <= 128 bits).
static int auto_design(int64_t nq, int64_t nd, int nb, int nlab, bool tile_ok, bool lane_ok) {
    (void)nq; (void)nd;
    // Measured at the BASELINE.json shapes (profiles/r02c_map_designs.txt; hist + rank, ms): 64-bit codes tile 0.55 + 1.59
    // / lane 0.48 + 1.39 (NUS-WIDE), 128-bit tile 2.29 + 6.71 / lane 0.91 + 2.46 (MS-COCO) - the tile design's per-thread
    // counter columns leave one to three CTAs per SM there; 16 / 32-bit codes tile 0.31 + 0.68 / lane 0.59 + 0.89 (few
    // buckets: the lanes of a step collide on the same counters).  Top-K (no labels) keeps the tile select kernel.
    if (lane_ok && nlab > 0 && nb >= 65) return 2;
    return tile_ok ? 0 : 1;
}

static int make_plan(int64_t nq, int64_t nd, int bits, int nlab, int ternary, int max_topn, int design, cmh_plan* plan,
                     int kq = 1, int kd = 1, int ap_mode = 0) {
    CMH_REQUIRE(plan, CMH_ERR_ARG, "cmh_eval_plan: NULL plan");
    CMH_REQUIRE(nq >= 0 && nd >= 0 && nlab >= 0 && max_topn >= 0 && max_topn <= CMH_MAX_TOPN, CMH_ERR_ARG,
                "cmh_eval_plan: bad sizes nq=%lld nd=%lld nlab=%d max_topn=%d", (long long)nq, (long long)nd, nlab, max_topn);
    CMH_REQUIRE(bits > 0 && bits <= CMH_MAX_BITS, CMH_ERR_UNSUPPORTED, "cmh_eval_plan: bits=%d outside (0, %d]", bits,
                CMH_MAX_BITS);
    CMH_REQUIRE(nd < (1ll << 32), CMH_ERR_UNSUPPORTED, "cmh_eval_plan: nd=%lld needs 32-bit ranks", (long long)nd);
    memset(plan, 0, sizeof(*plan));
    plan->bits = bits; plan->words = (bits + 63) / 64; plan->nlab = nlab; plan->lwords = (nlab + 63) / 64;
    plan->ternary = ternary ? 1 : 0; plan->nb = ternary ? 2 * bits + 1 : bits + 1;
    plan->max_topn = max_topn; plan->nq = nq; plan->nd = nd;
    plan->kq = kq; plan->kd = kd; plan->ap_mode = ap_mode;
    EvalArgs a = geometry_args(bits, nlab, ternary != 0);
    a.ntopn = max_topn;
    a.kq = kq; a.kd = kd;
    if (kq > 1 || kd > 1 || ap_mode) {                        // set-valued codes / textbook AP@k: the generic warp kernels
        CMH_REQUIRE(!ternary, CMH_ERR_UNSUPPORTED, "cmh_eval_plan_sets: binary codes only");
        CMH_REQUIRE(design < 0 || design == 1, CMH_ERR_UNSUPPORTED, "cmh_eval_plan_sets: the warp design walks set-valued codes");
        design = 1;
    }
    const bool tile_ok = a.cw <= MAX_CW && a.lw <= MAX_LW && plan->nb <= TILE_MAX_NB &&
                         tile_smem_bytes(a, ternary != 0, 1) <= MAX_DYN_SMEM;
    const bool lane_ok = lane_supported(a, ternary != 0);
    if (design < 0) {
        // CMH_EVAL_DESIGN=0/1/2 forces a design where it applies (measurement aid)
        static const int forced = [] { const char* e = getenv("CMH_EVAL_DESIGN"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1; }();
        if (forced == 2 && lane_ok) design = 2;
        else if (forced == 1) design = 1;
        else if (forced == 0 && tile_ok) design = 0;
        else design = auto_design(nq, nd, plan->nb, nlab, tile_ok, lane_ok);
    }
    CMH_REQUIRE(design != 0 || tile_ok, CMH_ERR_UNSUPPORTED, "cmh_eval_plan: tile design cannot hold %d buckets", plan->nb);
    CMH_REQUIRE(design != 2 || lane_ok, CMH_ERR_UNSUPPORTED, "cmh_eval_plan: lane design needs binary codes of <= 128 bits");
    plan->design = design;
    size_t smem; int max_cta_by_threads; int waves = 2;
    if (design == 0) {
        plan->q_tile = QT; smem = tile_smem_bytes(a, ternary != 0, 1); max_cta_by_threads = 2048 / QT;
    } else if (design == 2) {
        plan->q_tile = LANE_Q_TILE; smem = lane_smem_bytes(a, 1); max_cta_by_threads = 2048 / (LANE_Q_TILE * 32); waves = 4;
    } else {
        plan->q_tile = warp_queries_per_cta(a, ternary != 0);
        smem = warp_smem_bytes(a, ternary != 0, plan->q_tile); max_cta_by_threads = 2048 / (plan->q_tile * 32);
    }
    CMH_REQUIRE(smem <= MAX_DYN_SMEM, CMH_ERR_UNSUPPORTED, "cmh_eval_plan: %zu bytes of shared memory needed", smem);
    plan->nq_pad = round_up(std::max<int64_t>(nq, 1), plan->q_tile);
    plan->n_qtiles = (int32_t)(plan->nq_pad / plan->q_tile);
    const int ctas_per_sm = std::max(1, std::min<int>((int)(MAX_DYN_SMEM / smem), max_cta_by_threads));
    const int64_t target_ctas = (int64_t)sm_count() * ctas_per_sm * waves;  // full waves of resident CTAs
    const int64_t want = std::max<int64_t>(1, ceil_div(target_ctas, plan->n_qtiles));
    int64_t rows = round_up(std::max<int64_t>(1, ceil_div(std::max<int64_t>(nd, 1), want)), 16);
    rows = std::max<int64_t>(rows, MIN_CHUNK_ROWS);
    rows = std::min<int64_t>(rows, MAX_CHUNK_ROWS);
    plan->chunk_rows = (int32_t)rows;
    plan->n_chunks = (int32_t)std::max<int64_t>(1, ceil_div(nd, rows));
    CMH_REQUIRE(plan->n_chunks <= 65535, CMH_ERR_UNSUPPORTED, "cmh_eval_plan: %d chunks exceed the grid limit", plan->n_chunks);
    plan->workspace_bytes = carve_workspace(*plan, nullptr).bytes;
    return CMH_OK;
}

extern "C" int cmh_eval_plan(int64_t nq, int64_t nd, int bits, int nlab, int ternary, int max_topn, cmh_plan* plan) {
    return make_plan(nq, nd, bits, nlab, ternary, max_topn, -1, plan);
}
extern "C" int cmh_eval_plan_sets(int64_t nq, int64_t nd, int bits, int nlab, int kq, int kd, int ap_mode, cmh_plan* plan) {
    CMH_REQUIRE(kq >= 1 && kd >= 1 && kq <= 64 && kd <= 64 && (ap_mode == 0 || ap_mode == 1), CMH_ERR_ARG,
                "cmh_eval_plan_sets: kq=%d kd=%d ap_mode=%d", kq, kd, ap_mode);
    return make_plan(nq, nd, bits, nlab, 0, 1, -1, plan, kq, kd, ap_mode);
}
extern "C" int cmh_eval_plan_design(int64_t nq, int64_t nd, int bits, int nlab, int ternary, int max_topn, int design,
                                    cmh_plan* plan) {
    CMH_REQUIRE(design >= -1 && design <= 2, CMH_ERR_ARG, "cmh_eval_plan_design: design=%d", design);
    return make_plan(nq, nd, bits, nlab, ternary, max_topn, design, plan);
}

extern "C" int cmh_eval_hist(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d, uint32_t* hist_all,
                             uint32_t* hist_rel, void* workspace, void* stream) {
    CMH_REQUIRE(plan && workspace, CMH_ERR_ARG, "cmh_eval_hist: NULL plan / workspace");
    if (plan->nq == 0) return CMH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace w = carve_workspace(*plan, workspace);
    if (plan->nd == 0) {  // empty shard: all-zero histograms
        CMH_CUDA(cudaMemsetAsync(w.shard_all, 0, (size_t)plan->nq * plan->nb * 4, st));
        CMH_CUDA(cudaMemsetAsync(w.shard_rel, 0, (size_t)plan->nq * plan->nb * 4, st));
        CMH_CUDA(cudaMemsetAsync(w.chunk_hist, 0, (size_t)plan->n_chunks * plan->nb * plan->nq_pad * 4, st));
        if (hist_all) CMH_CUDA(cudaMemsetAsync(hist_all, 0, (size_t)plan->nq * plan->nb * 4, st));
        if (hist_rel) CMH_CUDA(cudaMemsetAsync(hist_rel, 0, (size_t)plan->nq * plan->nb * 4, st));
        return CMH_OK;
    }
    EvalArgs a;
    int rc = fill_args(*plan, q, d, false, &a);
    if (rc) return rc;
    if (plan->design == 0) {
        if ((rc = launch_hist_tile(a, plan->ternary != 0, w.chunk_hist, st))) return rc;
        const dim3 grid((unsigned)ceil_div(a.nq, 256), (unsigned)a.nb);
        reduce_hist_kernel<false><<<grid, 256, 0, st>>>(a, w.chunk_hist, w.shard_all, w.shard_rel, hist_all, hist_rel);
    } else {
        if (plan->design == 2) rc = launch_hist_lane(a, w.chunk_hist, st);
        else rc = launch_hist_warp(a, plan->ternary != 0, plan->q_tile, w.chunk_hist, st);
        if (rc) return rc;
        // grid.y (= query) is capped at 65535: walk the queries in slabs.  Shifting every [q][b]-major pointer by
        // q0 * nb moves the query origin (the chunk stride nq_pad * nb is unaffected).
        for (int64_t q0 = 0; q0 < a.nq; q0 += 65535) {
            EvalArgs s = a;
            s.nq = std::min<int64_t>(65535, a.nq - q0);
            const dim3 g((unsigned)ceil_div(a.nb, 256), (unsigned)s.nq);
            reduce_hist_kernel<true><<<g, 256, 0, st>>>(s, w.chunk_hist + q0 * a.nb, w.shard_all + q0 * a.nb,
                                                        w.shard_rel + q0 * a.nb, hist_all ? hist_all + q0 * a.nb : nullptr,
                                                        hist_rel ? hist_rel + q0 * a.nb : nullptr);
            CMH_LAUNCH_CHECK("reduce_hist_kernel");
        }
        return CMH_OK;
    }
    CMH_LAUNCH_CHECK("reduce_hist_kernel");
    return CMH_OK;
}

static int run_scan(const cmh_plan& plan, const EvalArgs& a, const Workspace& w, const ScanArgs& sa, cudaStream_t st) {
    if (plan.design == 0) {
        const size_t smem = (size_t)a.nb * 32 * 4 * 2;
        CMH_CUDA(cudaFuncSetAttribute(scan_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        scan_tile_kernel<<<(unsigned)(a.nq_pad / 32), 256, smem, st>>>(a, w.chunk_hist, sa);
    } else {
        scan_warp_kernel<<<(unsigned)ceil_div(a.nq_pad * 32, 256), 256, 0, st>>>(a, w.chunk_hist, sa);
    }
    CMH_LAUNCH_CHECK("scan kernel");
    return CMH_OK;
}

extern "C" int cmh_eval_rank(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d, int64_t k,
                             const uint32_t* lower_all, const uint32_t* lower_rel, const uint32_t* global_all,
                             const uint32_t* global_rel, const int64_t* topn, int ntopn, uint32_t* hits, double* ap_sum,
                             int64_t* n_rel, void* workspace, void* stream) {
    CMH_REQUIRE(plan && workspace, CMH_ERR_ARG, "cmh_eval_rank: NULL plan / workspace");
    CMH_REQUIRE(ntopn >= 0 && ntopn <= plan->max_topn, CMH_ERR_ARG, "cmh_eval_rank: ntopn=%d exceeds the plan's max_topn=%d",
                ntopn, plan->max_topn);
    CMH_REQUIRE(ntopn == 0 || (topn && hits), CMH_ERR_ARG, "cmh_eval_rank: topn / hits NULL");
    CMH_REQUIRE((global_all == nullptr) == (global_rel == nullptr) && (lower_all == nullptr) == (lower_rel == nullptr),
                CMH_ERR_ARG, "cmh_eval_rank: shard histograms must come in (all, rel) pairs");
    CMH_REQUIRE(ap_sum && n_rel, CMH_ERR_ARG, "cmh_eval_rank: NULL outputs");
    if (plan->nq == 0) return CMH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace w = carve_workspace(*plan, workspace);
    EvalArgs a;
    int rc;
    if (plan->nd == 0) {
        // nothing to rank on this shard; n_rel still has to reflect the other shards
        CMH_CUDA(cudaMemsetAsync(ap_sum, 0, (size_t)plan->nq * 8, st));
        if (ntopn) CMH_CUDA(cudaMemsetAsync(hits, 0, (size_t)plan->nq * ntopn * 4, st));
        a = geometry_args(plan->bits, plan->nlab, plan->ternary != 0);
        a.nq = plan->nq; a.nd = 0; a.nq_pad = plan->nq_pad; a.chunk_rows = plan->chunk_rows; a.n_chunks = 0;
    } else if ((rc = fill_args(*plan, q, d, true, &a))) {
        return rc;
    }
    TopnList tl;
    int order[CMH_MAX_TOPN];
    for (int i = 0; i < ntopn; ++i) {
        CMH_REQUIRE(topn[i] >= 1, CMH_ERR_ARG, "cmh_eval_rank: topn[%d]=%lld must be >= 1", i, (long long)topn[i]);
        order[i] = i;
    }
    std::stable_sort(order, order + ntopn, [&](int x, int y) { return topn[x] < topn[y]; });
    for (int i = 0; i < CMH_MAX_TOPN; ++i) {
        tl.n[i] = i < ntopn ? (uint32_t)std::min<int64_t>(topn[order[i]], 0xfffffffell) : 0xffffffffu;
        tl.perm[i] = i < ntopn ? order[i] : 0;
    }
    a.ntopn = ntopn;
    a.nmax = ntopn ? tl.n[ntopn - 1] : 0u;
    a.big_ranks = (global_all != nullptr || plan->nd >= (1ll << 23)) ? 1 : 0;   // ranks may exceed 2^23: no float bit tricks
    a.ap_mode = plan->ap_mode;
    a.kcut = (uint32_t)((k < 0 || k > 0xfffffffell) ? 0xfffffffeu : k);
    CMH_REQUIRE(plan->ap_mode == 0 || plan->design == 1, CMH_ERR_UNSUPPORTED, "cmh_eval_rank: ap_mode 1 needs a cmh_eval_plan_sets plan");
    CMH_REQUIRE(plan->ap_mode == 0 || (ntopn == 1 && topn[0] == (k < 0 ? topn[0] : k)), CMH_ERR_ARG,
                "cmh_eval_rank: ap_mode 1 takes topn = {k}");

    ScanArgs sa;
    sa.g_all = global_all ? global_all : w.shard_all;
    sa.g_rel = global_rel ? global_rel : w.shard_rel;
    sa.l_all = lower_all; sa.l_rel = lower_rel;
    sa.k = k; sa.K = 0;
    sa.base = w.base; sa.total = w.total; sa.n_rel = n_rel; sa.thr = nullptr;
    if ((rc = run_scan(*plan, a, w, sa, st))) return rc;
    if (plan->nd > 0) {
        if (plan->design == 0)
            rc = launch_rank_tile(a, plan->ternary != 0, w.base, w.total, tl, w.ap_part, w.hits_part, st);
        else if (plan->design == 2)
            rc = launch_rank_lane(a, w.base, w.total, tl, w.ap_part, w.hits_part, st);
        else
            rc = launch_rank_warp(a, plan->ternary != 0, plan->q_tile, w.base, w.total, tl, w.ap_part, w.hits_part, st);
        if (rc) return rc;
        reduce_parts_kernel<<<(unsigned)ceil_div(a.nq, 256), 256, 0, st>>>(a, w.ap_part, w.hits_part, tl, ap_sum, hits);
        CMH_LAUNCH_CHECK("reduce_parts_kernel");
    }
    return CMH_OK;
}

extern "C" int cmh_finalize_map(const double* ap_sum, const int64_t* n_rel, int64_t nq, int64_t k, double* ap,
                                float* map, void* stream) {
    CMH_REQUIRE(nq >= 0 && map, CMH_ERR_ARG, "cmh_finalize_map: bad arguments");
    CMH_REQUIRE(nq == 0 || (ap_sum && n_rel), CMH_ERR_ARG, "cmh_finalize_map: NULL inputs");
    finalize_map_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ap_sum, n_rel, nq, k, ap, map);
    CMH_LAUNCH_CHECK("finalize_map_kernel");
    return CMH_OK;
}

extern "C" int cmh_finalize_map_hits(const double* ap_sum, const uint32_t* hits, int ntopn, int64_t nq, double* ap, float* map,
                                     void* stream) {
    CMH_REQUIRE(nq >= 0 && ntopn >= 1 && map, CMH_ERR_ARG, "cmh_finalize_map_hits: bad arguments");
    CMH_REQUIRE(nq == 0 || (ap_sum && hits), CMH_ERR_ARG, "cmh_finalize_map_hits: NULL inputs");
    finalize_map_hits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ap_sum, hits, ntopn, nq, ap, map);
    CMH_LAUNCH_CHECK("finalize_map_hits_kernel");
    return CMH_OK;
}

extern "C" int cmh_finalize_topn(const uint32_t* hits, const int64_t* n_rel, int64_t nq, const int64_t* topn, int ntopn,
                                 int64_t nd_total, float* prec, void* stream) {
    CMH_REQUIRE(nq >= 0 && ntopn >= 0 && ntopn <= CMH_MAX_TOPN, CMH_ERR_ARG, "cmh_finalize_topn: bad sizes");
    if (ntopn == 0) return CMH_OK;
    CMH_REQUIRE(topn && prec && (nq == 0 || (hits && n_rel)), CMH_ERR_ARG, "cmh_finalize_topn: NULL pointer");
    TopnList tl;
    for (int i = 0; i < CMH_MAX_TOPN; ++i) {
        tl.n[i] = i < ntopn ? (uint32_t)std::min<int64_t>(topn[i], 0xfffffffell) : 0u;
        tl.perm[i] = i;
    }
    finalize_topn_kernel<<<ntopn, 1024, 0, (cudaStream_t)stream>>>(hits, n_rel, nq, ntopn, tl, nd_total, prec);
    CMH_LAUNCH_CHECK("finalize_topn_kernel");
    return CMH_OK;
}

extern "C" uint64_t cmh_finalize_pr_workspace_bytes(int64_t nq, int bits) {
    return (uint64_t)std::max<int64_t>(1, ceil_div(nq, 32)) * (uint64_t)(bits + 1) * 3 * 8;
}

extern "C" int cmh_finalize_pr(const uint32_t* hist_all, const uint32_t* hist_rel, int64_t nq, int bits, int ternary,
                               float* P, float* R, void* workspace, void* stream) {
    CMH_REQUIRE(nq >= 0 && bits > 0 && bits <= CMH_MAX_BITS && P && R && workspace, CMH_ERR_ARG,
                "cmh_finalize_pr: bad arguments");
    CMH_REQUIRE(nq == 0 || (hist_all && hist_rel), CMH_ERR_ARG, "cmh_finalize_pr: NULL histograms");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_warps = ceil_div(nq, 32);
    if (n_warps > 0) {
        pr_partial_kernel<<<(unsigned)ceil_div(n_warps * 32, 256), 256, 0, st>>>(hist_all, hist_rel, nq, bits, ternary,
                                                                                 (double*)workspace);
        CMH_LAUNCH_CHECK("pr_partial_kernel");
    }
    pr_final_kernel<<<(unsigned)ceil_div(bits + 1, 256), 256, 0, st>>>((const double*)workspace, n_warps, bits, P, R);
    CMH_LAUNCH_CHECK("pr_final_kernel");
    return CMH_OK;
}

extern "C" int cmh_map_k(const cmh_codeset* q, const cmh_codeset* d, int bits, int nlab, int64_t k, double* ap,
                         float* map, void* workspace, uint64_t workspace_bytes, void* stream) {
    CMH_REQUIRE(q && d && map, CMH_ERR_ARG, "cmh_map_k: NULL argument");
    const int ternary = (q->valid || d->valid) ? 1 : 0;
    cmh_plan plan;
    int rc = cmh_eval_plan(q->n, d->n, bits, nlab, ternary, 0, &plan);
    if (rc) return rc;
    const uint64_t extra = ((uint64_t)q->n * 16 + 511) & ~255ull;  // ap_sum + n_rel behind the plan's workspace
    CMH_REQUIRE(workspace && workspace_bytes >= plan.workspace_bytes + extra, CMH_ERR_WORKSPACE,
                "cmh_map_k: workspace of %llu bytes, %llu needed", (unsigned long long)workspace_bytes,
                (unsigned long long)(plan.workspace_bytes + extra));
    double* ap_sum = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + plan.workspace_bytes);
    int64_t* n_rel = reinterpret_cast<int64_t*>(ap_sum + q->n);
    if ((rc = cmh_eval_hist(&plan, q, d, nullptr, nullptr, workspace, stream))) return rc;
    if ((rc = cmh_eval_rank(&plan, q, d, k, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, ap_sum, n_rel,
                            workspace, stream)))
        return rc;
    return cmh_finalize_map(ap_sum, n_rel, q->n, k, ap, map, stream);
}

extern "C" uint64_t cmh_map_k_workspace_bytes(int64_t nq, int64_t nd, int bits, int nlab, int ternary) {
    cmh_plan plan;
    if (cmh_eval_plan(nq, nd, bits, nlab, ternary, 0, &plan)) return 0;
    return plan.workspace_bytes + (((uint64_t)nq * 16 + 511) & ~255ull);
}

extern "C" int cmh_topk(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d, int K, int64_t index_base,
                        uint64_t* keys, void* workspace, void* stream) {
    CMH_REQUIRE(plan && workspace && keys, CMH_ERR_ARG, "cmh_topk: NULL argument");
    CMH_REQUIRE(K >= 1 && index_base >= 0 && index_base + plan->nd <= (1ll << 32), CMH_ERR_ARG,
                "cmh_topk: K=%d index_base=%lld", K, (long long)index_base);
    if (plan->nq == 0) return CMH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // rows that do not exist are reported as UINT64_MAX; rows that do are overwritten at their exact rank
    fill_u64_kernel<<<(unsigned)std::min<int64_t>(ceil_div(plan->nq * (int64_t)K, 256), (int64_t)sm_count() * 16), 256, 0, st>>>(
        keys, plan->nq * (int64_t)K, ~0ull);
    CMH_LAUNCH_CHECK("fill_u64_kernel");
    if (plan->nd == 0) return CMH_OK;
    int rc;
    cmh_codeset qn = *q, dn = *d;
    qn.labels = nullptr; dn.labels = nullptr;
    if ((rc = cmh_eval_hist(plan, &qn, &dn, nullptr, nullptr, workspace, stream))) return rc;
    const Workspace w = carve_workspace(*plan, workspace);
    EvalArgs a;
    if ((rc = fill_args(*plan, &qn, &dn, false, &a))) return rc;
    a.K = K; a.index_base = index_base;
    ScanArgs sa;
    sa.g_all = w.shard_all; sa.g_rel = nullptr; sa.l_all = nullptr; sa.l_rel = nullptr;
    sa.k = -1; sa.K = K; sa.base = w.base; sa.total = w.total; sa.n_rel = nullptr; sa.thr = w.thr;
    if ((rc = run_scan(*plan, a, w, sa, st))) return rc;
    if (plan->design == 0) return launch_select_tile(a, plan->ternary != 0, w.base, w.thr, keys, st);
    return launch_select_warp(a, plan->ternary != 0, plan->q_tile, w.base, w.thr, keys, st);
}

extern "C" int cmh_topk_merge(const uint64_t* keys_in, int n_lists, int64_t nq, int K, uint64_t* keys_out, void* stream) {
    CMH_REQUIRE(n_lists >= 1 && nq >= 0 && K >= 1, CMH_ERR_ARG, "cmh_topk_merge: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(keys_in && keys_out, CMH_ERR_ARG, "cmh_topk_merge: NULL pointer");
    const size_t want = (size_t)n_lists * K * 8;
    const int in_smem = want <= 200 * 1024 ? 1 : 0;
    const size_t smem = in_smem ? want : 0;
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_topk_merge: too many queries per call");
    cudaStream_t st = (cudaStream_t)stream;
    fill_u64_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nq * (int64_t)K, 256), (int64_t)sm_count() * 16), 256, 0, st>>>(
        keys_out, nq * (int64_t)K, ~0ull);
    CMH_LAUNCH_CHECK("fill_u64_kernel");
    if (in_smem) CMH_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<(unsigned)nq, 256, smem, st>>>(keys_in, n_lists, nq, K, in_smem, keys_out);
    CMH_LAUNCH_CHECK("topk_merge_kernel");
    return CMH_OK;
}
