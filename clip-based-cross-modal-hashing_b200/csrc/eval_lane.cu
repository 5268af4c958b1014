// Warp-per-query, lane-per-row ("lane") kernels of the ranking-by-counting path for binary codes of up to 128 bits -
// the design for the reference's own calls (calc_map_k_matrix, utils/calc_utils.py:16-39) at the BASELINE.json shapes.
//
// Why a third design.  The thread-per-query kernels (eval_tile.cu) keep one counter column per THREAD: nb buckets x 8 B
// x 128 threads - 132 KB of shared memory per CTA for 128-bit codes, i.e. one 4-warp CTA per SM (ncu, profiles/r02a:
// rank_tile_kernel at the MS-COCO shape issues on 16 % of the cycles, 6 % of the warp slots are occupied).  Here a WARP
// owns a query and its 32 lanes own 32 consecutive database rows, so the counters are per warp (nb x 8 B = 1 KB) and
// the SM is full of warps.  The price is that "how many earlier rows of my bucket" is no longer a private running
// counter: within the 32 rows of a step it comes from __match_any_sync (lanes holding the same bucket) + a popcount of
// the peers below the lane, across steps from the warp's shared-memory counter, advanced by the last lane of each
// group.  Index order - the stable tie order of the reference's forced-stable torch.sort (:31) - is lane order.
//
//   hist_lane_kernel   pass 1: per (query, chunk) bucket counts, all + relevant packed 16|16 (shared-memory atomics,
//                      order does not matter)                                                        (:26-27,:30)
//   rank_lane_kernel   pass 2: ranks from the exclusive-scan bases; every relevant row adds relrank / rank (:31-37)
//
// Database rows stream through shared memory in LANE_ROWS-row stages filled by the bulk-copy engine (cp.async.bulk +
// mbarrier) while the previous stage is consumed; the 8 warps (= 8 queries) of a CTA share the stage.  Layout of the
// per-chunk histograms / bases: "W" ([chunk][query][bucket], eval_common.cuh), shared with the generic warp kernels.
#include <cstdlib>

#include "eval_common.cuh"

namespace cmh {

constexpr int LANE_WARPS = 8;     // queries (consumer warps) per CTA; one more warp feeds the stages
constexpr int LANE_THREADS = (LANE_WARPS + 1) * 32;
constexpr int LANE_ROWS = 512;    // database rows per stage

__host__ __device__ inline size_t lane_stage_bytes(int cws, int lws) { return (size_t)LANE_ROWS * (cws + lws) * 4; }
__host__ __device__ inline int lane_nbp(int nb) { return (nb + 1 + 3) & ~3; }   // + the parking bucket of idle lanes

constexpr int LANE_MODE_DEFAULT = 3;   // peers by shared-memory OR, four sub-histograms (see lane_mode())

size_t lane_smem_bytes(const EvalArgs& a, int kind /*0 hist, 1 rank*/) {
    // pass 1: up to four uint32 sub-histograms per warp; pass 2: 16-byte counter entries + the precision@N bins
    return 128 + 2 * lane_stage_bytes(a.cw_stride, a.lw_stride) + (size_t)LANE_WARPS * lane_nbp(a.nb) * 16 +
           (kind ? (size_t)LANE_WARPS * CMH_MAX_TOPN * 4 : 0);
}

bool lane_supported(const EvalArgs& a, bool tern) {
    return !tern && (a.cw_stride == 2 || a.cw_stride == 4) && (a.lw_stride == 0 || a.lw_stride == 2 || a.lw_stride == 4) &&
           a.nb <= 257;
}

// smem: [bars: 128 B][stage 0: codes, labels][stage 1: codes, labels][counters ...]; stage pointers are computed, not
// stored in an array (an array indexed by the stage parity would live in local memory)
struct LaneSmem {
    unsigned char* raw;
    uint32_t code_b, stage_b;
    __device__ __forceinline__ uint64_t* bar(int st) const { return reinterpret_cast<uint64_t*>(raw) + st; }         // stage full
    __device__ __forceinline__ uint64_t* empty(int st) const { return reinterpret_cast<uint64_t*>(raw) + 2 + st; }  // stage consumed
    __device__ __forceinline__ uint32_t* codes(int st) const { return reinterpret_cast<uint32_t*>(raw + 128 + (size_t)st * stage_b); }
    __device__ __forceinline__ uint32_t* labels(int st) const {
        return reinterpret_cast<uint32_t*>(raw + 128 + (size_t)st * stage_b + code_b);
    }
    __device__ __forceinline__ unsigned char* rest() const { return raw + 128 + 2 * (size_t)stage_b; }
};

__device__ __forceinline__ LaneSmem carve_lane_smem(unsigned char* raw, int cws, int lws) {
    LaneSmem s;
    s.raw = raw;
    s.code_b = (uint32_t)LANE_ROWS * cws * 4;
    s.stage_b = (uint32_t)lane_stage_bytes(cws, lws);
    return s;
}

// The producer warp: fills stage t % 2 with the rows of tile t as soon as the eight consumer warps have released it -
// whole stages through the bulk-copy engine (one lane), the ragged last stage of a chunk (and every stage of a shard
// view that is not 16-byte aligned) with the warp's own loads.  No CTA-wide barrier in the loop: a consumer warp moves
// on to the next stage as soon as that stage is full, whatever the other queries are doing (the r02p captures had 2-3
// of ~20 stall cycles per issue on __syncthreads).
template <int CWS, int LWS>
__device__ __forceinline__ void lane_producer(const EvalArgs& a, const LaneSmem& s, int64_t c_begin, int c_rows, int n_tiles) {
    const int lane = threadIdx.x & 31;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        if (t >= 2) mbar_wait(s.empty(st), ((t >> 1) - 1) & 1);          // the consumers are done with tile t - 2
        const int64_t row0 = c_begin + (int64_t)t * LANE_ROWS;
        const int rows = min(LANE_ROWS, c_rows - t * LANE_ROWS);
        const uint32_t code_b = (uint32_t)rows * CWS * 4, lab_b = (uint32_t)rows * LWS * 4;
        if (rows == LANE_ROWS && a.bulk_ok) {
            if (lane == 0) {
                mbar_expect_tx(s.bar(st), code_b + lab_b);
                bulk_g2s(s.codes(st), a.ds + row0 * CWS, code_b, s.bar(st));
                if (LWS) bulk_g2s(s.labels(st), a.dl + row0 * LWS, lab_b, s.bar(st));
            }
        } else {
            for (int i = lane; i < rows * CWS; i += 32) s.codes(st)[i] = a.ds[row0 * CWS + i];
            if (LWS)
                for (int i = lane; i < rows * LWS; i += 32) s.labels(st)[i] = a.dl[row0 * LWS + i];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(s.bar(st));                       // release: the stores above are visible to the waiters
        }
        __syncwarp();
    }
}

// ---- shared memory by 32-bit address (no generic-pointer window arithmetic in the loops: the r02d capture showed 14 % of
// the stall samples on the S2UR / ULEA pair that rebuilds the shared window base every iteration) ----------------------
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_shared(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_shared(uint32_t a, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// the query of a warp: the same words in every lane
template <int CWS, int LWS>
struct LaneQuery {
    uint32_t s[CWS];
    uint32_t l[LWS ? LWS : 1];

    __device__ __forceinline__ void load(const EvalArgs& a, int64_t q) {
        const bool live = q < a.nq;
#pragma unroll
        for (int w = 0; w < CWS; ++w) s[w] = live ? a.qs[q * CWS + w] : 0u;
        if (LWS) {
#pragma unroll
            for (int w = 0; w < LWS; ++w) l[w] = live ? a.ql[q * LWS + w] : 0u;
        }
    }
    // Hamming distance to the staged row at shared address `rc` (padding words are zero on both sides)
    __device__ __forceinline__ int bucket(uint32_t rc) const {
        if (CWS == 2) {
            const uint2 r = lds64(rc);
            return __popc(s[0] ^ r.x) + __popc(s[1] ^ r.y);
        }
        const uint4 r = lds128(rc);
        return __popc(s[0] ^ r.x) + __popc(s[1] ^ r.y) + __popc(s[2 % CWS] ^ r.z) + __popc(s[3 % CWS] ^ r.w);
    }
    __device__ __forceinline__ bool relevant(uint32_t rl) const {
        constexpr int N = LWS ? LWS : 1;
        if (LWS == 0) return false;
        if (LWS == 2) {
            const uint2 r = lds64(rl);
            return ((l[0] & r.x) | (l[1 % N] & r.y)) != 0u;
        }
        const uint4 r = lds128(rl);
        return ((l[0] & r.x) | (l[1 % N] & r.y) | (l[2 % N] & r.z) | (l[3 % N] & r.w)) != 0u;
    }
};

// =================================================================================================================
// pass 1
// =================================================================================================================
// COPIES sub-histograms per warp (lane % COPIES picks one): the lanes of a step collide on the few buckets around the
// mean distance, and a collision costs a shared-memory wavefront each (r02d: 4.6 extra wavefronts per step with one copy)
template <int CWS, int LWS, int COPIES>
__global__ void __launch_bounds__(LANE_THREADS) hist_lane_kernel(const EvalArgs a, uint32_t* __restrict__ chunk_hist) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LaneSmem s = carve_lane_smem(smem_raw, CWS, LWS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbp = lane_nbp(a.nb);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(s.rest()) + warp * (COPIES * nbp);
    const uint32_t cnt_u = smem_u32(cnt) + (uint32_t)(lane % COPIES) * (uint32_t)nbp * 4u;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * LANE_WARPS + warp;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const int n_tiles = (c_rows + LANE_ROWS - 1) / LANE_ROWS;

    if (threadIdx.x == 0) {
        mbar_init(s.bar(0), 1);
        mbar_init(s.bar(1), 1);
        mbar_init(s.empty(0), LANE_WARPS);
        mbar_init(s.empty(1), LANE_WARPS);
        mbar_fence_init();
    }
    if (warp == LANE_WARPS) {                         // (uniform per warp)
        __syncthreads();
        lane_producer<CWS, LWS>(a, s, c_begin, c_rows, n_tiles);
        return;
    }
    LaneQuery<CWS, LWS> qu;
    qu.load(a, q);
    for (int b = lane; b < COPIES * nbp; b += 32) cnt[b] = 0u;
    __syncthreads();

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        mbar_wait(s.bar(st), (t >> 1) & 1);
        const int rows = min(LANE_ROWS, c_rows - t * LANE_ROWS);
        const uint32_t tc = smem_u32(s.codes(st)) + (uint32_t)lane * (CWS * 4u);
        const uint32_t tl = smem_u32(s.labels(st)) + (uint32_t)lane * (LWS * 4u);
        const int full = rows & ~31;
#pragma unroll 4
        for (int g = 0; g < full; g += 32) {
            const int d = qu.bucket(tc + (uint32_t)g * (CWS * 4u));
            red_add_shared(cnt_u + 4u * (uint32_t)d, qu.relevant(tl + (uint32_t)g * (LWS * 4u)) ? 0x10001u : 1u);   // order is irrelevant here
        }
        if (full + lane < rows) {
            const int d = qu.bucket(tc + (uint32_t)full * (CWS * 4u));
            red_add_shared(cnt_u + 4u * (uint32_t)d, qu.relevant(tl + (uint32_t)full * (LWS * 4u)) ? 0x10001u : 1u);
        }
        __syncwarp();                                 // every lane has read its rows
        if (lane == 0) mbar_arrive(s.empty(st));      // this warp is done with stage st
    }
    __syncwarp();
    if (q < a.nq_pad)
        for (int b = lane; b < a.nb; b += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int c = 0; c < COPIES; ++c) v += cnt[c * nbp + b];     // (all | rel << 16): no carry between the halves, counts <= 65520
            chunk_hist[hist_index_W(a, chunk, b, q)] = v;
        }
}

// =================================================================================================================
// pass 2 - average precision + precision@N
// =================================================================================================================
// integer -> float without the XU pipe (POPC, I2F and MUFU.RCP share it): exact for x < 2^23
template <bool BIG>
__device__ __forceinline__ float lane_u2f(uint32_t x) {
    if (BIG) return (float)x;
    return __uint_as_float(0x4B000000u | x) - 8388608.0f;
}
// 1 / y for y >= 1: one MUFU.RCP, without the denormal rescaling __fdividef / __frcp_rn carry around
__device__ __forceinline__ float lane_rcp(float y) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
}

// A counter entry is 16 bytes: (rows ranked so far in the bucket, relevant rows among them, peer mask of the even lanes,
// peer mask of the odd lanes).  PEERS = 0: the lanes that share a bucket in a step are found with __match_any_sync;
// PEERS = 1: every lane ORs its bit into the entry of its bucket (a shared-memory reduction; even and odd lanes use
// different words, i.e. different banks) and reads the mask back with the counters - MATCH.ANY carries ~half of the stall
// samples of the r02d capture.  Either way the last lane of a group writes the group's final ranks and clears the masks.
template <int CWS, int LWS, bool BIG, int PEERS>
__global__ void __launch_bounds__(LANE_THREADS) rank_lane_kernel(const EvalArgs a, const uint2* __restrict__ base,
                                                                    const uint32_t* __restrict__ total_arr, const TopnList topn,
                                                                    double* __restrict__ ap_part,
                                                                    uint32_t* __restrict__ hits_part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LaneSmem s = carve_lane_smem(smem_raw, CWS, LWS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbp = lane_nbp(a.nb);
    uint4* cnt = reinterpret_cast<uint4*>(s.rest()) + warp * nbp;
    const uint32_t cnt_u = smem_u32(cnt);
    uint32_t* hits = reinterpret_cast<uint32_t*>(s.rest() + (size_t)LANE_WARPS * nbp * 16) + warp * CMH_MAX_TOPN;
    const int chunk = blockIdx.y;
    const int64_t q = (int64_t)blockIdx.x * LANE_WARPS + warp;
    const bool qlive = q < a.nq_pad;
    const int64_t c_begin = (int64_t)chunk * a.chunk_rows;
    const int c_rows = (int)min((int64_t)a.chunk_rows, a.nd - c_begin);
    const int n_tiles = (c_rows + LANE_ROWS - 1) / LANE_ROWS;
    const uint32_t lt = lanemask_lt();
    const uint32_t my_bit = 1u << lane, my_mask_off = 8u + 4u * (uint32_t)(lane & 1);

    if (threadIdx.x == 0) {
        mbar_init(s.bar(0), 1);
        mbar_init(s.bar(1), 1);
        mbar_init(s.empty(0), LANE_WARPS);
        mbar_init(s.empty(1), LANE_WARPS);
        mbar_fence_init();
    }
    if (warp == LANE_WARPS) {                         // (uniform per warp)
        __syncthreads();
        lane_producer<CWS, LWS>(a, s, c_begin, c_rows, n_tiles);
        return;
    }
    LaneQuery<CWS, LWS> qu;
    qu.load(a, q);
    const uint32_t total = qlive ? total_arr[q] : 0u;
    for (int b = lane; b < nbp; b += 32) {
        const uint2 v = (qlive && b < a.nb) ? base[hist_index_W(a, chunk, b, q)] : make_uint2(0u, 0u);
        cnt[b] = make_uint4(v.x, v.y, 0u, 0u);
    }
    for (int i = lane; i < a.ntopn; i += 32) hits[i] = 0u;
    __syncthreads();

    const uint32_t nmax = a.nmax;
    double acc = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t & 1;
        mbar_wait(s.bar(st), (t >> 1) & 1);
        const int rows = min(LANE_ROWS, c_rows - t * LANE_ROWS);
        const uint32_t tc = smem_u32(s.codes(st)) + (uint32_t)lane * (CWS * 4u);
        const uint32_t tl = smem_u32(s.labels(st)) + (uint32_t)lane * (LWS * 4u);
        float acc_t = 0.f;      // <= 16 terms <= 1 per lane and stage: float32 rounding stays below 16 * 2^-24 relative
        // one step = 32 rows, lane = row.  `in`: false only for the idle lanes of a chunk's ragged last step.
        auto step = [&](int g, bool in) {
            const int d = in ? qu.bucket(tc + (uint32_t)g * (CWS * 4u)) : a.nb;      // idle lanes park in bucket nb
            const bool rel = in && qu.relevant(tl + (uint32_t)g * (LWS * 4u));
            const uint32_t ent = cnt_u + 16u * (uint32_t)d;
            uint32_t grp;
            uint4 c;
            if (PEERS == 0) {
                grp = __match_any_sync(0xffffffffu, d);              // the lanes (rows) of this step in my bucket
                c = lds128(ent);
            } else {
                red_or_shared(ent + my_mask_off, my_bit);
                __syncwarp();                                        // every lane of the step has announced itself
                c = lds128(ent);
                grp = c.z | c.w;
            }
            const uint32_t relmask = __ballot_sync(0xffffffffu, rel);
            const uint32_t below = grp & lt;                         // ... of lower index
            const uint32_t rank = c.x + (uint32_t)__popc(below) + 1u;
            const uint32_t rr = c.y + (uint32_t)__popc(below & relmask) + (rel ? 1u : 0u);
            // count / tindex (:35): x * rcp(y), <= 1.5 ulp per term, i.e. <= 2e-7 on an AP in [0, 1] (the bar is 1e-6)
            const float term = lane_u2f<BIG>(rr) * lane_rcp(lane_u2f<BIG>(rank));
            acc_t += (rel && rr <= total) ? term : 0.f;
            if (rel && rank <= nmax) {                               // precision@N, rare
                int i = 0;
                while (rank > topn.n[i]) ++i;
                atomicAdd(&hits[i], 1u);
            }
            __syncwarp();                                            // every lane has read its entry
            if ((grp >> lane) == 1u) sts128(ent, make_uint4(rank, rr, 0u, 0u));   // the group's last row leaves its ranks behind
            __syncwarp();
        };
        const int full = rows & ~31;
#pragma unroll 2
        for (int g = 0; g < full; g += 32) step(g, true);
        if (full < rows) step(full, full + lane < rows);
        acc += (double)acc_t;
        __syncwarp();                                 // every lane has read its rows
        if (lane == 0) mbar_arrive(s.empty(st));      // this warp is done with stage st
    }
    // fixed-order reduction over the lanes: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (qlive) {
        if (lane == 0) ap_part[(int64_t)chunk * a.nq_pad + q] = acc;
        __syncwarp();
        for (int i = lane; i < a.ntopn; i += 32) hits_part[((int64_t)chunk * a.ntopn + i) * a.nq_pad + q] = hits[i];
    }
}

// =================================================================================================================
// launchers
// =================================================================================================================
template <typename Kern>
static int prep_lane(Kern kern, size_t smem) {
    CMH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return CMH_OK;
}

// CMH_LANE_MODE (measurement aid): bit 0 = peers through shared-memory OR instead of MATCH.ANY (pass 2), bit 1 = four
// sub-histograms per warp instead of one (pass 1)
static int lane_mode() {
    static const int mode = [] { const char* e = getenv("CMH_LANE_MODE"); return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : LANE_MODE_DEFAULT; }();
    return mode;
}

int launch_hist_lane(const EvalArgs& a, uint32_t* chunk_hist, cudaStream_t st) {
    const dim3 grid((unsigned)(a.nq_pad / LANE_WARPS), (unsigned)a.n_chunks);
    const size_t smem = lane_smem_bytes(a, 0);
    const bool four = (lane_mode() & 2) != 0;
#define CMH_GO(CWS_, LWS_)                                                       \
    do {                                                                         \
        if (four) {                                                              \
            auto k = hist_lane_kernel<CWS_, LWS_, 4>;                            \
            int rc = prep_lane(k, smem);                                         \
            if (rc) return rc;                                                   \
            k<<<grid, LANE_THREADS, smem, st>>>(a, chunk_hist);               \
        } else {                                                                 \
            auto k = hist_lane_kernel<CWS_, LWS_, 1>;                            \
            int rc = prep_lane(k, smem);                                         \
            if (rc) return rc;                                                   \
            k<<<grid, LANE_THREADS, smem, st>>>(a, chunk_hist);               \
        }                                                                        \
    } while (0)
    if (a.cw_stride == 2) {
        if (a.lw_stride == 0) CMH_GO(2, 0); else if (a.lw_stride == 2) CMH_GO(2, 2); else CMH_GO(2, 4);
    } else {
        if (a.lw_stride == 0) CMH_GO(4, 0); else if (a.lw_stride == 2) CMH_GO(4, 2); else CMH_GO(4, 4);
    }
#undef CMH_GO
    CMH_LAUNCH_CHECK("hist_lane_kernel");
    return CMH_OK;
}

int launch_rank_lane(const EvalArgs& a, const uint2* base, const uint32_t* total, const TopnList& tl, double* ap_part,
                     uint32_t* hits_part, cudaStream_t st) {
    const dim3 grid((unsigned)(a.nq_pad / LANE_WARPS), (unsigned)a.n_chunks);
    const size_t smem = lane_smem_bytes(a, 1);
    const bool big = a.big_ranks != 0;
    const bool peers_or = (lane_mode() & 1) != 0;
#define CMH_GO(CWS_, LWS_, BIG_, P_)                                                    \
    do {                                                                                \
        auto k = rank_lane_kernel<CWS_, LWS_, BIG_, P_>;                                \
        int rc = prep_lane(k, smem);                                                    \
        if (rc) return rc;                                                              \
        k<<<grid, LANE_THREADS, smem, st>>>(a, base, total, tl, ap_part, hits_part); \
    } while (0)
#define CMH_PICK(CWS_, LWS_)                                                            \
    do {                                                                                \
        if (big) { if (peers_or) CMH_GO(CWS_, LWS_, true, 1); else CMH_GO(CWS_, LWS_, true, 0); }   \
        else     { if (peers_or) CMH_GO(CWS_, LWS_, false, 1); else CMH_GO(CWS_, LWS_, false, 0); } \
    } while (0)
    if (a.cw_stride == 2) { if (a.lw_stride == 2) CMH_PICK(2, 2); else CMH_PICK(2, 4); }
    else                  { if (a.lw_stride == 2) CMH_PICK(4, 2); else CMH_PICK(4, 4); }
#undef CMH_PICK
#undef CMH_GO
    CMH_LAUNCH_CHECK("rank_lane_kernel");
    return CMH_OK;
}

}  // namespace cmh
