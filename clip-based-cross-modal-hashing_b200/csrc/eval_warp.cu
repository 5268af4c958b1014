// Warp-per-query ("warp") kernels of the ranking-by-counting path - the generic design for any bucket count
// (TwDH long codes 256...2048 bits, train/TwDH/hash_train.py:206-217, and ternary codes of any length), where the
// per-thread counter columns of eval_tile.cu no longer fit in shared memory.
//
// One CTA = WPC queries (one per warp) x one database chunk; a lane owns one database row of the current 32-row
// group, so counters are per WARP (nb entries) instead of per thread.  Index order inside a bucket is recovered with
// __match_any_sync: lanes holding the same bucket form a group, a lane's position in the group is the popcount of
// the group mask below it, and the group's last lane advances the counter.  Rows are staged in shared memory with
// an odd 32-bit row stride so that "lane = row" reads are conflict-free.
#include "eval_common.cuh"

namespace cmh {

constexpr int W_ROWS = 64;  // database rows per shared-memory stage (two 32-row groups)

struct WarpGeom {
    int wpc;          // warps (= queries) per CTA
    int cs, ls;       // padded (odd) 32-bit row strides of the staged codes / labels
    size_t smem;
};

static inline int odd(int x) { return x | 1; }

// wpc_fixed > 0: use that many warps per CTA (the plan's q_tile) instead of the largest that fits
WarpGeom warp_geom(const EvalArgs& a, bool tern, int kind /*0 hist, 1 rank, 2 select*/, int wpc_fixed = 0) {
    WarpGeom g;
    g.cs = odd(a.cw * (a.kd > 0 ? a.kd : 1));      // an item = kd sub-codes of cw words (set-valued codes; 1 otherwise)
    g.ls = a.lw ? odd(a.lw) : 1;
    const size_t per_warp_cnt = (size_t)(a.nb + 1) * (kind == 1 ? 8 : 4);
    const size_t per_warp_q = (size_t)a.cw * (a.kq > 0 ? a.kq : 1) * 4 * (tern ? 2 : 1) + (size_t)a.lw * 4 + (size_t)CMH_MAX_TOPN * 4;
    const size_t stage = (size_t)W_ROWS * (g.cs * (tern ? 2 : 1) + g.ls) * 4;
    int wpc = (int)((200 * 1024 - stage) / (per_warp_cnt + per_warp_q));
    g.wpc = wpc_fixed > 0 ? wpc_fixed : (wpc < 1 ? 1 : (wpc > 8 ? 8 : wpc));
    g.smem = stage + (size_t)g.wpc * (per_warp_cnt + per_warp_q) + 64;
    return g;
}

struct WarpSmem {
    uint32_t *codes, *valid, *labels;  // staged rows
    uint32_t *qs, *qv, *ql;            // this warp's query words
    uint32_t* hits;                    // this warp's precision@N bins
    unsigned char* cnt;                // this warp's counters
};

template <bool TERN>
__device__ __forceinline__ WarpSmem carve_warp_smem(unsigned char* raw, const EvalArgs& a, int cs, int ls, int wpc,
                                                    int cnt_elem_bytes) {
    WarpSmem s;
    const int warp = threadIdx.x >> 5;
    uint32_t* p = reinterpret_cast<uint32_t*>(raw);
    s.codes = p; p += W_ROWS * cs;
    s.valid = p; if (TERN) p += W_ROWS * cs;
    s.labels = p; p += W_ROWS * ls;
    const int qw = a.cw * a.kq;                      // query words per item
    s.qs = p + warp * qw; p += wpc * qw;
    s.qv = p + warp * qw; if (TERN) p += wpc * qw;
    s.ql = p + warp * a.lw; p += wpc * a.lw;
    s.hits = p + warp * CMH_MAX_TOPN; p += wpc * CMH_MAX_TOPN;
    uintptr_t c = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
    s.cnt = reinterpret_cast<unsigned char*>(c) + (size_t)warp * (a.nb + 1) * cnt_elem_bytes;
    return s;
}

template <bool TERN>
__device__ __forceinline__ void stage_rows(const EvalArgs& a, const WarpSmem& s, int cs, int ls, int64_t row0, int rows) {
    const int iw = a.cw * a.kd;                      // used words per item; in memory an item is kd x cw_stride words
    for (int i = threadIdx.x; i < rows * iw; i += blockDim.x) {
        const int r = i / iw, w = i - r * iw;
        const int64_t src = (row0 + r) * (int64_t)(a.kd * a.cw_stride) + (w / a.cw) * a.cw_stride + (w % a.cw);
        s.codes[r * cs + w] = a.ds[src];
        if (TERN) s.valid[r * cs + w] = a.dv[src];
    }
    for (int i = threadIdx.x; i < rows * a.lw; i += blockDim.x) {
        const int r = i / a.lw, w = i - r * a.lw;
        s.labels[r * ls + w] = a.dl[(row0 + r) * a.lw_stride + w];
    }
}

template <bool TERN>
__device__ __forceinline__ void load_query(const EvalArgs& a, const WarpSmem& s, int64_t q) {
    const int lane = threadIdx.x & 31;
    const bool live = q < a.nq;
    for (int w = lane; w < a.cw * a.kq; w += 32) {
        const int64_t src = q * (int64_t)(a.kq * a.cw_stride) + (w / a.cw) * a.cw_stride + (w % a.cw);
        s.qs[w] = live ? a.qs[src] : 0u;
        if (TERN) s.qv[w] = live ? a.qv[src] : 0u;
    }
    for (int w = lane; w < a.lw; w += 32) s.ql[w] = live ? a.ql[q * a.lw_stride + w] : 0u;
}

template <bool TERN>
__device__ __forceinline__ int row_bucket(const EvalArgs& a, const WarpSmem& s, int cs, int r) {
    const uint32_t* rc = s.codes + r * cs;
    if (!TERN) {
        if (a.kq * a.kd > 1) {
            // set-valued codes (train/DPSIH/_utils.py:15-21): similarity = max over the kq x kd pairs of sub-codes, i.e.
            // the distance of two items is the MINIMUM of the pairwise sub-code distances
            int best = a.bits;
            for (int x = 0; x < a.kq; ++x)
                for (int y = 0; y < a.kd; ++y) {
                    int d = 0;
                    for (int w = 0; w < a.cw; ++w) d += __popc(s.qs[x * a.cw + w] ^ rc[y * a.cw + w]);
                    best = min(best, d);
                }
            return best;
        }
        int d = 0;
        for (int w = 0; w < a.cw; ++w) d += __popc(s.qs[w] ^ rc[w]);
        return d;
    }
    const uint32_t* rv = s.valid + r * cs;
    int d = a.bits;
    for (int w = 0; w < a.cw; ++w) {
        const uint32_t both = s.qv[w] & rv[w];
        d += 2 * __popc((s.qs[w] ^ rc[w]) & both) - __popc(both);
    }
    return d;
}

__device__ __forceinline__ bool row_relevant(const EvalArgs& a, const WarpSmem& s, int ls, int r) {
    uint32_t acc = 0;
    const uint32_t* rl = s.labels + r * ls;
    for (int w = 0; w < a.lw; ++w) acc |= s.ql[w] & rl[w];
    return acc != 0u;
}

// ---- pass 1 ------------------------------------------------------------------------------------------------------
template <bool TERN>
__global__ void __launch_bounds__(256) hist_warp_kernel(const EvalArgs a, int cs, int ls, int wpc,
                                                        uint32_t* __restrict__ chunk_hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const WarpSmem s = carve_warp_smem<TERN>(smem_raw, a, cs, ls, wpc, 4);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(s.cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * wpc + warp;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);

    load_query<TERN>(a, s, q);
    for (int b = lane; b <= a.nb; b += 32) cnt[b] = 0u;
    for (int r0 = 0; r0 < c_rows; r0 += W_ROWS) {
        const int rows = min(W_ROWS, c_rows - r0);
        __syncthreads();
        stage_rows<TERN>(a, s, cs, ls, c_begin + r0, rows);
        __syncthreads();
        for (int g = 0; g < rows; g += 32) {
            const int r = g + lane;
            if (r < rows) {
                const int d = row_bucket<TERN>(a, s, cs, r);
                atomicAdd(&cnt[d], row_relevant(a, s, ls, r) ? 0x10001u : 1u);
            }
        }
    }
    __syncwarp();
    if (q < a.nq_pad)
        for (int b = lane; b < a.nb; b += 32) chunk_hist[hist_index_W(a, chunk, b, q)] = cnt[b];
}

// ---- pass 2 ------------------------------------------------------------------------------------------------------
template <bool TERN>
__global__ void __launch_bounds__(256) rank_warp_kernel(const EvalArgs a, int cs, int ls, int wpc,
                                                        const uint2* __restrict__ base,
                                                        const uint32_t* __restrict__ total_arr, const TopnList topn,
                                                        double* __restrict__ ap_part, uint32_t* __restrict__ hits_part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const WarpSmem s = carve_warp_smem<TERN>(smem_raw, a, cs, ls, wpc, 8);
    uint2* cnt = reinterpret_cast<uint2*>(s.cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * wpc + warp;
    const bool qlive = q < a.nq_pad;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const uint32_t lt = lanemask_lt();

    load_query<TERN>(a, s, q);
    const uint32_t total = qlive ? total_arr[q] : 0u;
    for (int b = lane; b <= a.nb; b += 32)
        cnt[b] = (qlive && b < a.nb) ? base[hist_index_W(a, chunk, b, q)] : make_uint2(0u, 0u);
    for (int i = lane; i < a.ntopn; i += 32) s.hits[i] = 0u;
    double acc = 0.0;
    for (int r0 = 0; r0 < c_rows; r0 += W_ROWS) {
        const int rows = min(W_ROWS, c_rows - r0);
        __syncthreads();
        stage_rows<TERN>(a, s, cs, ls, c_begin + r0, rows);
        __syncthreads();
        for (int g = 0; g < rows; g += 32) {
            const int r = g + lane;
            const bool in = r < rows;
            const int d = in ? row_bucket<TERN>(a, s, cs, r) : a.nb;  // nb = parking bucket of idle lanes
            const bool rel = in && row_relevant(a, s, ls, r);
            const uint32_t grp = __match_any_sync(0xffffffffu, d);
            const uint32_t relmask = __ballot_sync(0xffffffffu, rel);
            const uint2 c = cnt[d];
            const uint32_t rank = c.x + __popc(grp & lt) + 1u;
            const uint32_t rr = c.y + __popc(grp & relmask & lt) + 1u;
            if (rel) {
                if (a.ap_mode ? rank <= a.kcut : rr <= total) acc += (double)__fdiv_rn((float)rr, (float)rank);
                if (rank <= a.nmax) {
                    int i = 0;
                    while (rank > topn.n[i]) ++i;
                    atomicAdd(&s.hits[i], 1u);
                }
            }
            __syncwarp();
            if ((grp >> lane) == 1u)  // last lane of the group advances the bucket's counters
                cnt[d] = make_uint2(c.x + __popc(grp), c.y + __popc(grp & relmask));
            __syncwarp();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (qlive) {
        if (lane == 0) ap_part[(int64_t)chunk * a.nq_pad + q] = acc;
        for (int i = lane; i < a.ntopn; i += 32)
            hits_part[((int64_t)chunk * a.ntopn + i) * a.nq_pad + q] = s.hits[i];
    }
}

// ---- top-K select --------------------------------------------------------------------------------------------------
template <bool TERN>
__global__ void __launch_bounds__(256) select_warp_kernel(const EvalArgs a, int cs, int ls, int wpc,
                                                          const uint2* __restrict__ base,
                                                          const int32_t* __restrict__ thr_arr,
                                                          uint64_t* __restrict__ keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const WarpSmem s = carve_warp_smem<TERN>(smem_raw, a, cs, ls, wpc, 4);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(s.cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * wpc + warp;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const uint32_t lt = lanemask_lt();

    load_query<TERN>(a, s, q);
    const int thr = q < a.nq ? thr_arr[q] : -1;
    for (int b = lane; b <= a.nb; b += 32) cnt[b] = (b <= thr && b < a.nb) ? base[hist_index_W(a, chunk, b, q)].x : 0u;
    uint64_t* __restrict__ out = keys + q * (int64_t)a.K;
    for (int r0 = 0; r0 < c_rows; r0 += W_ROWS) {
        const int rows = min(W_ROWS, c_rows - r0);
        __syncthreads();
        stage_rows<TERN>(a, s, cs, ls, c_begin + r0, rows);
        __syncthreads();
        for (int g = 0; g < rows; g += 32) {
            const int r = g + lane;
            int d = r < rows ? row_bucket<TERN>(a, s, cs, r) : a.nb;
            const bool take = d <= thr;
            if (__any_sync(0xffffffffu, take)) {
                if (!take) d = a.nb;
                const uint32_t grp = __match_any_sync(0xffffffffu, d);
                const uint32_t pos = cnt[d] + __popc(grp & lt);
                if (take && pos < (uint32_t)a.K)
                    out[pos] = ((uint64_t)(TERN ? d : 2 * d) << 32) | (uint64_t)(a.index_base + c_begin + r0 + r);
                __syncwarp();
                if ((grp >> lane) == 1u) cnt[d] += __popc(grp);
                __syncwarp();
            }
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------------------------------
int warp_queries_per_cta(const EvalArgs& a, bool tern) { return warp_geom(a, tern, 1).wpc; }
size_t warp_smem_bytes(const EvalArgs& a, bool tern, int wpc) { return warp_geom(a, tern, 1, wpc).smem; }

template <typename Kern>
static int prep(Kern kern, size_t smem) {
    CMH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return CMH_OK;
}

int launch_hist_warp(const EvalArgs& a, bool tern, int wpc, uint32_t* chunk_hist, cudaStream_t st) {
    const WarpGeom g = warp_geom(a, tern, 0, wpc);
    const dim3 grid((unsigned)ceil_div(a.nq_pad, wpc), (unsigned)a.n_chunks);
    int rc;
    if (tern) {
        if ((rc = prep(hist_warp_kernel<true>, g.smem))) return rc;
        hist_warp_kernel<true><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, chunk_hist);
    } else {
        if ((rc = prep(hist_warp_kernel<false>, g.smem))) return rc;
        hist_warp_kernel<false><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, chunk_hist);
    }
    CMH_LAUNCH_CHECK("hist_warp_kernel");
    return CMH_OK;
}

int launch_rank_warp(const EvalArgs& a, bool tern, int wpc, const uint2* base, const uint32_t* total,
                     const TopnList& tl, double* ap_part, uint32_t* hits_part, cudaStream_t st) {
    const WarpGeom g = warp_geom(a, tern, 1, wpc);
    const dim3 grid((unsigned)ceil_div(a.nq_pad, wpc), (unsigned)a.n_chunks);
    int rc;
    if (tern) {
        if ((rc = prep(rank_warp_kernel<true>, g.smem))) return rc;
        rank_warp_kernel<true><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, base, total, tl, ap_part, hits_part);
    } else {
        if ((rc = prep(rank_warp_kernel<false>, g.smem))) return rc;
        rank_warp_kernel<false><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, base, total, tl, ap_part, hits_part);
    }
    CMH_LAUNCH_CHECK("rank_warp_kernel");
    return CMH_OK;
}

int launch_select_warp(const EvalArgs& a, bool tern, int wpc, const uint2* base, const int32_t* thr, uint64_t* keys,
                       cudaStream_t st) {
    const WarpGeom g = warp_geom(a, tern, 2, wpc);
    const dim3 grid((unsigned)ceil_div(a.nq_pad, wpc), (unsigned)a.n_chunks);
    int rc;
    if (tern) {
        if ((rc = prep(select_warp_kernel<true>, g.smem))) return rc;
        select_warp_kernel<true><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, base, thr, keys);
    } else {
        if ((rc = prep(select_warp_kernel<false>, g.smem))) return rc;
        select_warp_kernel<false><<<grid, wpc * 32, g.smem, st>>>(a, g.cs, g.ls, wpc, base, thr, keys);
    }
    CMH_LAUNCH_CHECK("select_warp_kernel");
    return CMH_OK;
}

}  // namespace cmh
