// Thread-per-query ("tile") kernels of the ranking-by-counting path - the fast design for nb <= TILE_MAX_NB
// buckets (binary codes up to 199 bits, ternary up to 99), i.e. every BASELINE.json config.
//
// One CTA = QT (128) queries x one database chunk.  A thread owns ONE query: its code / label words live in
// registers, and its per-bucket counters live in a private shared-memory column cnt[bucket][tid] (bank = tid % 32,
// so never a conflict and never an atomic).  Database rows stream through shared memory in TILE_ROWS-row stages
// filled by the TMA engine (cp.async.bulk + mbarrier, SASS UBLKCP) while the previous stage is consumed; every
// thread reads the same row (a shared-memory broadcast).  Because a thread walks its chunk in ascending row order,
// "#{i < j : d_i == d_j}" is simply the running counter - the stable tie order of the reference's forced-stable
// torch.sort (utils/calc_utils.py:31) costs nothing extra.
//
//   hist_tile_kernel    pass 1: counts per (query, chunk, bucket), all + relevant, packed 16|16        (:26-27,:30)
//   rank_tile_kernel    pass 2: counters start at the exclusive-scan bases; every relevant row adds
//                       relrank / rank to the query's AP sum and to the precision@N first-crossing bins  (:31-37)
//   select_tile_kernel  top-K: rows at or below the query's threshold bucket are written at their exact rank
#include "eval_common.cuh"

namespace cmh {

constexpr int LW_NONE = -1;

struct TileSmem {
    uint64_t* bar;          // [2] full barriers
    uint32_t* codes[2];
    uint32_t* valid[2];
    uint32_t* labels[2];
    unsigned char* rest;    // counters etc.
};

__host__ __device__ inline size_t tile_stage_bytes(int cw_stride, int lw_stride, bool tern) {
    size_t b = (size_t)TILE_ROWS * cw_stride * 4 * (tern ? 2 : 1) + (size_t)TILE_ROWS * lw_stride * 4;
    return (b + 127) & ~(size_t)127;
}

template <bool TERN>
__device__ __forceinline__ TileSmem carve_tile_smem(unsigned char* raw, const EvalArgs& a) {
    TileSmem s;
    s.bar = reinterpret_cast<uint64_t*>(raw);
    unsigned char* p = raw + 128;
    const size_t code_b = (size_t)TILE_ROWS * a.cw_stride * 4;
    const size_t stage_b = tile_stage_bytes(a.cw_stride, a.lw_stride, TERN);
#pragma unroll
    for (int st = 0; st < 2; ++st) {
        unsigned char* base = p + st * stage_b;
        s.codes[st] = reinterpret_cast<uint32_t*>(base);
        s.valid[st] = reinterpret_cast<uint32_t*>(base + code_b);
        s.labels[st] = reinterpret_cast<uint32_t*>(base + code_b * (TERN ? 2 : 1));
    }
    s.rest = p + 2 * stage_b;
    return s;
}

// Fill one stage with database rows [row0, row0 + rows).  Full stages go through the bulk-copy engine; the ragged
// last stage of a chunk is copied by the CTA (row counts that are not a multiple of 16 bytes cannot be bulk-copied),
// and so is every stage of a shard whose planes are not 16-byte aligned (a row-range view starting at an odd row).
// Must be called by all threads of the CTA.
template <bool TERN>
__device__ __forceinline__ void load_stage(const EvalArgs& a, const TileSmem& s, int st, int64_t row0, int rows) {
    const uint32_t code_b = (uint32_t)rows * a.cw_stride * 4;
    const uint32_t lab_b = (uint32_t)rows * a.lw_stride * 4;
    if (rows == TILE_ROWS && a.bulk_ok) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s.bar[st], code_b * (TERN ? 2u : 1u) + lab_b);
            bulk_g2s(s.codes[st], a.ds + row0 * a.cw_stride, code_b, &s.bar[st]);
            if (TERN) bulk_g2s(s.valid[st], a.dv + row0 * a.cw_stride, code_b, &s.bar[st]);
            if (lab_b) bulk_g2s(s.labels[st], a.dl + row0 * a.lw_stride, lab_b, &s.bar[st]);
        }
    } else {
        const int nc = rows * a.cw_stride, nl = rows * a.lw_stride;
        for (int i = threadIdx.x; i < nc; i += QT) {
            s.codes[st][i] = a.ds[row0 * a.cw_stride + i];
            if (TERN) s.valid[st][i] = a.dv[row0 * a.cw_stride + i];
        }
        for (int i = threadIdx.x; i < nl; i += QT) s.labels[st][i] = a.dl[row0 * a.lw_stride + i];
        __syncthreads();
        if (threadIdx.x == 0) mbar_arrive(&s.bar[st]);
    }
}

// ---- per-thread query state ------------------------------------------------------------------------------------
template <int CW, int LW, bool TERN>
struct Query {
    static constexpr int NCW = CW > 0 ? CW : MAX_CW;
    static constexpr int NLW = LW > 0 ? LW : (LW == 0 ? MAX_LW : 1);
    uint32_t s[NCW];
    uint32_t v[TERN ? NCW : 1];
    uint32_t l[NLW];

    __device__ __forceinline__ void load(const EvalArgs& a, int64_t q) {
        const bool live = q < a.nq;
#pragma unroll
        for (int w = 0; w < NCW; ++w) {
            const bool in = live && w < a.cw;
            s[w] = in ? a.qs[q * a.cw_stride + w] : 0u;
            if (TERN) v[w] = in ? a.qv[q * a.cw_stride + w] : 0u;
        }
        if (LW != LW_NONE) {
#pragma unroll
            for (int w = 0; w < NLW; ++w) l[w] = (live && w < a.lw) ? a.ql[q * a.lw_stride + w] : 0u;
        }
    }

    // bucket of database row `rc` (binary: Hamming distance; ternary: bits - dot = 2 * dist)
    __device__ __forceinline__ int bucket(const uint32_t* __restrict__ rc, const uint32_t* __restrict__ rv,
                                          const EvalArgs& a) const {
        if (!TERN) {
            if (CW == 1) return __popc(s[0] ^ rc[0]);
            if (CW == 2) {
                const uint2 r = *reinterpret_cast<const uint2*>(rc);
                return __popc(s[0] ^ r.x) + __popc(s[1] ^ r.y);
            }
            if (CW == 4) {
                const uint4 r = *reinterpret_cast<const uint4*>(rc);
                return __popc(s[0] ^ r.x) + __popc(s[1] ^ r.y) + __popc(s[2] ^ r.z) + __popc(s[3] ^ r.w);
            }
            int d = 0;
#pragma unroll
            for (int w = 0; w < NCW; ++w)
                if (w < a.cw) d += __popc(s[w] ^ rc[w]);
            return d;
        } else {
            int d = a.bits;
#pragma unroll
            for (int w = 0; w < NCW; ++w)
                if (w < a.cw) {
                    const uint32_t both = v[w] & rv[w];
                    d += 2 * __popc((s[w] ^ rc[w]) & both) - __popc(both);
                }
            return d;
        }
    }

    __device__ __forceinline__ bool relevant(const uint32_t* __restrict__ rl, const EvalArgs& a) const {
        if (LW == LW_NONE) return false;
        if (LW == 1) return (l[0] & rl[0]) != 0u;
        if (LW == 4) {
            const uint4 r = *reinterpret_cast<const uint4*>(rl);
            return ((l[0] & r.x) | (l[1] & r.y) | (l[2] & r.z) | (l[3] & r.w)) != 0u;
        }
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < NLW; ++w)
            if (w < a.lw) acc |= l[w] & rl[w];
        return acc != 0u;
    }
};

// =================================================================================================================
// pass 1
// =================================================================================================================
template <int CW, int LW, bool TERN>
__global__ void __launch_bounds__(QT) hist_tile_kernel(const EvalArgs a, uint32_t* __restrict__ chunk_hist) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem s = carve_tile_smem<TERN>(smem_raw, a);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(s.rest);  // [nb][QT]
    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * QT + tid;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const int n_tiles = (c_rows + TILE_ROWS - 1) / TILE_ROWS;

    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        mbar_fence_init();
    }
    Query<CW, LW, TERN> qu;
    qu.load(a, q);
    for (int b = 0; b < a.nb; ++b) cnt[b * QT + tid] = 0u;
    __syncthreads();
    load_stage<TERN>(a, s, 0, c_begin, min(TILE_ROWS, c_rows));
    if (n_tiles > 1) load_stage<TERN>(a, s, 1, c_begin + TILE_ROWS, min(TILE_ROWS, c_rows - TILE_ROWS));

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        mbar_wait(&s.bar[st], (t >> 1) & 1);
        const int rows = min(TILE_ROWS, c_rows - t * TILE_ROWS);
        const uint32_t* __restrict__ tc = s.codes[st];
        const uint32_t* __restrict__ tv = s.valid[st];
        const uint32_t* __restrict__ tl = s.labels[st];
#pragma unroll 4
        for (int j = 0; j < rows; ++j) {
            const int d = qu.bucket(tc + j * a.cw_stride, tv + j * a.cw_stride, a);
            const bool rel = qu.relevant(tl + j * a.lw_stride, a);
            cnt[d * QT + tid] += rel ? 0x10001u : 1u;
        }
        __syncthreads();  // everyone is done with stage st
        if (t + 2 < n_tiles)
            load_stage<TERN>(a, s, st, c_begin + (int64_t)(t + 2) * TILE_ROWS, min(TILE_ROWS, c_rows - (t + 2) * TILE_ROWS));
    }
    for (int b = 0; b < a.nb; ++b) chunk_hist[hist_index_T(a, chunk, b, q)] = cnt[b * QT + tid];
}

// =================================================================================================================
// pass 2 - average precision + precision@N
// =================================================================================================================
template <int CW, int LW, bool TERN>
__global__ void __launch_bounds__(QT) rank_tile_kernel(const EvalArgs a, const uint2* __restrict__ base,
                                                       const uint32_t* __restrict__ total_arr,
                                                       const TopnList topn,
                                                       double* __restrict__ ap_part,
                                                       uint32_t* __restrict__ hits_part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem s = carve_tile_smem<TERN>(smem_raw, a);
    uint2* cnt = reinterpret_cast<uint2*>(s.rest);                                 // [nb][QT]
    uint32_t* hits = reinterpret_cast<uint32_t*>(s.rest + (size_t)a.nb * QT * 8);  // [ntopn][QT]
    uint32_t* s_topn = hits + (size_t)a.ntopn * QT;                                // [ntopn]
    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * QT + tid;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const int n_tiles = (c_rows + TILE_ROWS - 1) / TILE_ROWS;

    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        mbar_fence_init();
    }
    Query<CW, LW, TERN> qu;
    qu.load(a, q);
    const uint32_t total = total_arr[q];
    for (int b = 0; b < a.nb; ++b) cnt[b * QT + tid] = base[hist_index_T(a, chunk, b, q)];
    for (int i = 0; i < a.ntopn; ++i) hits[i * QT + tid] = 0u;
    if (tid < CMH_MAX_TOPN) s_topn[tid] = topn.n[tid];
    __syncthreads();
    load_stage<TERN>(a, s, 0, c_begin, min(TILE_ROWS, c_rows));
    if (n_tiles > 1) load_stage<TERN>(a, s, 1, c_begin + TILE_ROWS, min(TILE_ROWS, c_rows - TILE_ROWS));

    const uint32_t nmax = a.nmax;
    double acc = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        mbar_wait(&s.bar[st], (t >> 1) & 1);
        const int rows = min(TILE_ROWS, c_rows - t * TILE_ROWS);
        const uint32_t* __restrict__ tc = s.codes[st];
        const uint32_t* __restrict__ tv = s.valid[st];
        const uint32_t* __restrict__ tl = s.labels[st];
        // float32 partial sums over 4 rows (each term <= 1, so the rounding of the partial sum stays below
        // 4 * 2^-24 relative), flushed into the float64 accumulator: keeps F2F.F64 off the per-row path.
        for (int j0 = 0; j0 < rows; j0 += 4) {
            float acc4 = 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < rows) {
                    const int d = qu.bucket(tc + j * a.cw_stride, tv + j * a.cw_stride, a);
                    const bool rel = qu.relevant(tl + j * a.lw_stride, a);
                    uint2* p = cnt + d * QT + tid;
                    uint2 c = *p;
                    c.x += 1u;                      // rank of this row
                    c.y += rel ? 1u : 0u;           // its rank among the relevant rows
                    // count / tindex (:35), predicated instead of branched: the 128 queries of a CTA disagree on
                    // relevance at almost every row, so a branch is paid in full anyway.  x * rcp(y): <= 1.5 ulp per
                    // term, i.e. <= 2e-7 on an AP in [0, 1] (the bar is 1e-6)
                    const float term = __fdividef((float)c.y, (float)c.x);
                    acc4 += (rel && c.y <= total) ? term : 0.f;
                    if (rel && c.x <= nmax) {       // precision@N, rare
                        int i = 0;
                        while (c.x > s_topn[i]) ++i;
                        hits[i * QT + tid] += 1u;
                    }
                    *p = c;
                }
            }
            acc += (double)acc4;
        }
        __syncthreads();
        if (t + 2 < n_tiles)
            load_stage<TERN>(a, s, st, c_begin + (int64_t)(t + 2) * TILE_ROWS, min(TILE_ROWS, c_rows - (t + 2) * TILE_ROWS));
    }
    ap_part[(int64_t)chunk * a.nq_pad + q] = acc;
    for (int i = 0; i < a.ntopn; ++i)
        hits_part[((int64_t)chunk * a.ntopn + i) * a.nq_pad + q] = hits[i * QT + tid];
}

// =================================================================================================================
// top-K select: exact rank of every row at or below the threshold bucket
// =================================================================================================================
template <int CW, bool TERN>
__global__ void __launch_bounds__(QT) select_tile_kernel(const EvalArgs a, const uint2* __restrict__ base,
                                                         const int32_t* __restrict__ thr_arr,
                                                         uint64_t* __restrict__ keys) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem s = carve_tile_smem<TERN>(smem_raw, a);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(s.rest);  // [nb][QT] rank of the next row of the bucket
    const int tid = threadIdx.x;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * QT + tid;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const int n_tiles = (c_rows + TILE_ROWS - 1) / TILE_ROWS;

    if (tid == 0) {
        mbar_init(&s.bar[0], 1);
        mbar_init(&s.bar[1], 1);
        mbar_fence_init();
    }
    Query<CW, LW_NONE, TERN> qu;
    qu.load(a, q);
    const int thr = q < a.nq ? thr_arr[q] : -1;
    for (int b = 0; b < a.nb; ++b) cnt[b * QT + tid] = b <= thr ? base[hist_index_T(a, chunk, b, q)].x : 0u;
    __syncthreads();
    load_stage<TERN>(a, s, 0, c_begin, min(TILE_ROWS, c_rows));
    if (n_tiles > 1) load_stage<TERN>(a, s, 1, c_begin + TILE_ROWS, min(TILE_ROWS, c_rows - TILE_ROWS));

    const uint32_t K = (uint32_t)a.K;
    uint64_t* __restrict__ out = keys + q * (int64_t)a.K;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        mbar_wait(&s.bar[st], (t >> 1) & 1);
        const int rows = min(TILE_ROWS, c_rows - t * TILE_ROWS);
        const uint32_t* __restrict__ tc = s.codes[st];
        const uint32_t* __restrict__ tv = s.valid[st];
        const int64_t row_base = a.index_base + c_begin + (int64_t)t * TILE_ROWS;
#pragma unroll 4
        for (int j = 0; j < rows; ++j) {
            const int d = qu.bucket(tc + j * a.cw_stride, tv + j * a.cw_stride, a);
            if (d <= thr) {
                const uint32_t pos = cnt[d * QT + tid]++;
                if (pos < K) out[pos] = ((uint64_t)(TERN ? d : 2 * d) << 32) | (uint64_t)(row_base + j);
            }
        }
        __syncthreads();
        if (t + 2 < n_tiles)
            load_stage<TERN>(a, s, st, c_begin + (int64_t)(t + 2) * TILE_ROWS, min(TILE_ROWS, c_rows - (t + 2) * TILE_ROWS));
    }
}

// =================================================================================================================
// host-side launchers
// =================================================================================================================
size_t tile_smem_bytes(const EvalArgs& a, bool tern, int kind /*0 hist, 1 rank, 2 select*/) {
    size_t b = 128 + 2 * tile_stage_bytes(a.cw_stride, a.lw_stride, tern);
    if (kind == 0) b += (size_t)a.nb * QT * 4;
    if (kind == 1) b += (size_t)a.nb * QT * 8 + (size_t)a.ntopn * QT * 4 + (size_t)CMH_MAX_TOPN * 4;
    if (kind == 2) b += (size_t)a.nb * QT * 4;
    return b;
}

template <typename Kern>
static int prep(Kern kern, size_t smem) {
    CMH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return CMH_OK;
}

#define CMH_TILE_DISPATCH_CW(CWV, BODY)  \
    if (a.cw == CWV) {                   \
        constexpr int CW_ = CWV;         \
        BODY                             \
    } else

int launch_hist_tile(const EvalArgs& a, bool tern, uint32_t* chunk_hist, cudaStream_t st) {
    const dim3 grid((unsigned)(a.nq_pad / QT), (unsigned)a.n_chunks);
    const size_t smem = tile_smem_bytes(a, tern, 0);
#define CMH_GO(CW_, LW_, T_)                                                         \
    do {                                                                             \
        auto k = hist_tile_kernel<CW_, LW_, T_>;                                     \
        int rc = prep(k, smem);                                                      \
        if (rc) return rc;                                                           \
        k<<<grid, QT, smem, st>>>(a, chunk_hist);                                    \
    } while (0)
    if (tern) {
        if (a.lw == 0) CMH_GO(0, LW_NONE, true); else CMH_GO(0, 0, true);
    } else {
        const int cwsel = (a.cw == 1 || a.cw == 2 || a.cw == 4) ? a.cw : 0;
        const int lwsel = a.lw == 0 ? LW_NONE : (a.lw == 1 ? 1 : ((a.lw == 3 || a.lw == 4) && a.lw_stride == 4 ? 4 : 0));
#define CMH_ROW(CW_)                                                 \
    switch (lwsel) {                                                 \
        case LW_NONE: CMH_GO(CW_, LW_NONE, false); break;            \
        case 1: CMH_GO(CW_, 1, false); break;                        \
        case 4: CMH_GO(CW_, 4, false); break;                        \
        default: CMH_GO(CW_, 0, false); break;                       \
    }
        switch (cwsel) {
            case 1: CMH_ROW(1) break;
            case 2: CMH_ROW(2) break;
            case 4: CMH_ROW(4) break;
            default: CMH_ROW(0) break;
        }
#undef CMH_ROW
    }
#undef CMH_GO
    CMH_LAUNCH_CHECK("hist_tile_kernel");
    return CMH_OK;
}

int launch_rank_tile(const EvalArgs& a, bool tern, const uint2* base, const uint32_t* total, const TopnList& topn_sorted,
                     double* ap_part, uint32_t* hits_part, cudaStream_t st) {
    const dim3 grid((unsigned)(a.nq_pad / QT), (unsigned)a.n_chunks);
    const size_t smem = tile_smem_bytes(a, tern, 1);
#define CMH_GO(CW_, LW_, T_)                                                         \
    do {                                                                             \
        auto k = rank_tile_kernel<CW_, LW_, T_>;                                     \
        int rc = prep(k, smem);                                                      \
        if (rc) return rc;                                                           \
        k<<<grid, QT, smem, st>>>(a, base, total, topn_sorted, ap_part, hits_part);  \
    } while (0)
    if (tern) {
        CMH_GO(0, 0, true);
    } else {
        const int cwsel = (a.cw == 1 || a.cw == 2 || a.cw == 4) ? a.cw : 0;
        const int lwsel = a.lw == 1 ? 1 : ((a.lw == 3 || a.lw == 4) && a.lw_stride == 4 ? 4 : 0);
#define CMH_ROW(CW_)                                                 \
    switch (lwsel) {                                                 \
        case 1: CMH_GO(CW_, 1, false); break;                        \
        case 4: CMH_GO(CW_, 4, false); break;                        \
        default: CMH_GO(CW_, 0, false); break;                       \
    }
        switch (cwsel) {
            case 1: CMH_ROW(1) break;
            case 2: CMH_ROW(2) break;
            case 4: CMH_ROW(4) break;
            default: CMH_ROW(0) break;
        }
#undef CMH_ROW
    }
#undef CMH_GO
    CMH_LAUNCH_CHECK("rank_tile_kernel");
    return CMH_OK;
}

int launch_select_tile(const EvalArgs& a, bool tern, const uint2* base, const int32_t* thr, uint64_t* keys,
                       cudaStream_t st) {
    const dim3 grid((unsigned)(a.nq_pad / QT), (unsigned)a.n_chunks);
    const size_t smem = tile_smem_bytes(a, tern, 2);
#define CMH_GO(CW_, T_)                                          \
    do {                                                         \
        auto k = select_tile_kernel<CW_, T_>;                    \
        int rc = prep(k, smem);                                  \
        if (rc) return rc;                                       \
        k<<<grid, QT, smem, st>>>(a, base, thr, keys);           \
    } while (0)
    if (tern) {
        CMH_GO(0, true);
    } else {
        switch (a.cw) {
            case 1: CMH_GO(1, false); break;
            case 2: CMH_GO(2, false); break;
            case 4: CMH_GO(4, false); break;
            default: CMH_GO(0, false); break;
        }
    }
#undef CMH_GO
    CMH_LAUNCH_CHECK("select_tile_kernel");
    return CMH_OK;
}

}  // namespace cmh
