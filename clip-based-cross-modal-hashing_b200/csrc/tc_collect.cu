// K3 - the tensor-core variant of the query x database Hamming compare (north star (2)):
// the reference's identity  dist = 0.5 * (bits - qB . rB^T)  (utils/calc_utils.py:12) evaluated as a +-1 int8 GEMM
// on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), with the top-K candidate filter
// fused as the epilogue so that the Q x D distance matrix never exists.
//
// One CTA = TC_T x 128 queries (the A operand, expanded once to +-1 int8 in shared memory) x one database chunk.
// Warp roles (warp-specialised, mbarrier pipelines, no __syncthreads in the main loop):
//   warp 0        TMEM allocation; one elected lane issues tcgen05.mma + tcgen05.commit
//   warps 1-4     producers: read packed 64-bit code words (8 B per row, coalesced), expand every bit to a +-1 byte
//                 with a 16-entry nibble LUT and store 128x... core matrices (no-swizzle K-major UMMA layout) into a
//                 TC_STAGES-deep ring of B tiles (256 database rows each)
//   warps 5-12    epilogue: thread = query (TMEM lane), tcgen05.ld 32 columns at a time, VIMNMX3 max-tree, compare
//                 with the query's threshold; a rare hit appends key (2*dist << 32 | global row) to the query's
//                 candidate list in global memory
// TMEM: 2 accumulator buffers x 256 columns (the whole 512-column TMEM, one CTA per SM).
//
// Exactness: thresholds only have to be upper bounds of the K-th distance (cmh_topk_threshold derives them from a
// sample histogram); cmh_topk_finalize sorts the candidates by key - (distance, index), all keys distinct - which IS
// the stable ranking, and flags any query whose candidate list is short or overflowed for the exact two-pass path.
#include <algorithm>

#include "common.cuh"

namespace cmh {

constexpr int TC_M = 128;
constexpr int TC_N = 256;
constexpr int TC_PROD_WARPS = 4;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (1 + TC_PROD_WARPS + TC_EPI_WARPS) * 32;
constexpr int TC_MAX_T = 4;

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// all tcgen05.mma issued so far by this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1):
//   16-byte units: element (row r, 16-byte K chunk c) lives at  (r % 8) + (r / 8) * SBO + c * LBO
// i.e. a core matrix is 8 rows x 16 B = 128 contiguous bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, signed int8 A and B, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 16 bits -> 16 bytes of +-1 (bit set -> +1 = 0x01, clear -> -1 = 0xFF) through the nibble LUT
__device__ __forceinline__ uint4 expand16(uint32_t bits16, const uint32_t* __restrict__ lut) {
    uint4 o;
    o.x = lut[bits16 & 15u];
    o.y = lut[(bits16 >> 4) & 15u];
    o.z = lut[(bits16 >> 8) & 15u];
    o.w = lut[(bits16 >> 12) & 15u];
    return o;
}

constexpr int TC_SLOTS = 8;        // staged hits per epilogue LANE (private slots: the hit path has no atomics)
constexpr int TC_FLUSH_AT = 4;     // a warp flushes once any of its lanes holds this many staged hits
constexpr int TC_REFRESH = 128;    // tile iterations between threshold refreshes

// per-query bookkeeping shared by every CTA working on the query (global memory, zeroed by cmh_tc_collect)
struct TcAux {
    uint32_t cum_le[4];   // candidates found so far with dist <= thr0 - j   (j = 0..3)
    uint32_t force_fail;  // a staging buffer overflowed: a candidate may have been dropped
    uint32_t pad[3];
};

struct TcArgs {
    const uint64_t* q;      // [nq][words]
    const uint64_t* d;      // [nd][words]
    const int32_t* thr;     // [nq] initial threshold bucket thr0 (Hamming distance): rows with dist <= thr qualify
    uint64_t* cand;         // [nq][cap]
    uint32_t* cnt;          // [nq] candidates found (may exceed cap)
    TcAux* aux;             // [nq]
    int64_t nq, nd, index_base;
    int chunk_rows;         // database rows per CTA (multiple of TC_N)
    int cap, bits;
    int K;                  // > 0: tighten thresholds while scanning (once K rows at dist <= thr0 - j are known)
};

__device__ __forceinline__ int max8(const int* v) {
    int m = __vimax3_s32(v[0], v[1], v[2]);
    int n = __vimax3_s32(v[3], v[4], v[5]);
    return __vimax3_s32(m, n, max(v[6], v[7]));
}

// Out of line on purpose: one copy of the append code keeps the epilogue's instruction footprint small (an inlined,
// fully unrolled hit path was ~160 KB of SASS and every rare hit paid a chain of instruction-cache misses).
__device__ __noinline__ int tc_stage_hit(uint64_t* my_key, unsigned char* my_t, int n_staged, uint64_t key, int t,
                                         uint32_t* force_fail) {
    if (n_staged < TC_SLOTS) {
        my_key[n_staged * 32] = key;
        my_t[n_staged * 32] = (unsigned char)t;
        return n_staged + 1;
    }
    *force_fail = 1u;
    return n_staged;
}

// smem: [A: T tiles][B: STAGES tiles][lut 64 B][barriers][tmem slot][staging][thresholds]
template <int WORDS, int T, int TC_STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_collect_kernel(const TcArgs a) {
    constexpr int KBYTES = WORDS * 64;           // int8 elements (= bytes) per row
    constexpr int KSTEPS = KBYTES / 32;          // tcgen05.mma kind::i8 has K = 32
    constexpr int CHUNKS = KBYTES / 16;          // 16-byte K chunks per row
    constexpr uint32_t A_TILE = TC_M * KBYTES;
    constexpr uint32_t B_TILE = TC_N * KBYTES;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + T * A_TILE;
    uint32_t* lut = reinterpret_cast<uint32_t*>(sB + TC_STAGES * B_TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(lut + 16);
    uint64_t* b_full = bars;                     // [STAGES] producers -> MMA
    uint64_t* b_empty = bars + TC_STAGES;        // [STAGES] MMA -> producers
    uint64_t* t_full = bars + 2 * TC_STAGES;     // [2] MMA -> epilogue
    uint64_t* t_empty = bars + 2 * TC_STAGES + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    // staged hits: [EPI_WARPS][SLOTS][32 lanes] (slot-major, lane fastest: conflict-free), 8-byte aligned
    uint64_t* stg_key = reinterpret_cast<uint64_t*>(tmem_slot + 2);
    unsigned char* stg_t = reinterpret_cast<unsigned char*>(stg_key + TC_EPI_WARPS * TC_SLOTS * 32);
    int* s_thr = reinterpret_cast<int*>(stg_t + TC_EPI_WARPS * TC_SLOTS * 32);    // [T][128] current dot thresholds

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * (T * TC_M);
    const int64_t c_begin = (int64_t)blockIdx.y * a.chunk_rows;
    const int64_t c_end = min(a.nd, c_begin + a.chunk_rows);
    const int n_tiles = (int)((c_end - c_begin + TC_N - 1) / TC_N);

    // ---- prologue -----------------------------------------------------------------------------------------------
    if (tid < 16) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= (((tid >> i) & 1) ? 0x01u : 0xFFu) << (8 * i);
        lut[tid] = w;
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&b_full[s], TC_PROD_WARPS * 32);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&t_full[b], 1);
            mbar_init(&t_empty[b], TC_EPI_WARPS * 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // A operand: T x 128 query rows, expanded by everyone (rows beyond nq are all -1; their threshold never fires)
    for (int i = tid; i < T * TC_M * WORDS * 4; i += TC_THREADS) {
        const int c = i % (WORDS * 4);           // 16-byte chunk of the row
        const int r = (i / (WORDS * 4)) % TC_M;
        const int t = i / (WORDS * 4 * TC_M);
        const int64_t q = q0 + t * TC_M + r;
        const uint64_t word = q < a.nq ? a.q[q * WORDS + (c >> 2)] : 0ull;
        const uint32_t b16 = (uint32_t)(word >> (16 * (c & 3))) & 0xffffu;
        *reinterpret_cast<uint4*>(sA + t * A_TILE + c * (TC_M * 16) + r * 16) = expand16(b16, lut);
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_M, TC_N);
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            int it = 0;
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TC_STAGES;
                mbar_wait(&b_full[s], (i / TC_STAGES) & 1);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < T; ++t, ++it) {
                    const int buf = it & 1;
                    mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                        const uint64_t ad = umma_desc(a_addr + t * A_TILE + k * 2 * (TC_M * 16), TC_M * 16, 128);
                        const uint64_t bd = umma_desc(b_addr + s * B_TILE + k * 2 * (TC_N * 16), TC_N * 16, 128);
                        umma_i8(tmem_base + buf * TC_N, ad, bd, idesc, k > 0 ? 1u : 0u);
                    }
                    umma_commit(&t_full[buf]);
                }
                umma_commit(&b_empty[s]);
            }
        }
    } else if (warp <= TC_PROD_WARPS) {
        // ================= producers: packed bits -> +-1 int8 core matrices =================
        const int pt = tid - 32;                 // 0..127
        uint64_t w[2][WORDS];
        auto fetch = [&](int i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t row = c_begin + (int64_t)i * TC_N + h * 128 + pt;
#pragma unroll
                for (int x = 0; x < WORDS; ++x) w[h][x] = row < c_end ? __ldg(a.d + row * WORDS + x) : 0ull;
            }
        };
        if (n_tiles > 0) fetch(0);
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % TC_STAGES;
            uint64_t cur[2][WORDS];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int x = 0; x < WORDS; ++x) cur[h][x] = w[h][x];
            if (i + 1 < n_tiles) fetch(i + 1);   // next tile's words are in flight while this one is expanded
            mbar_wait(&b_empty[s], ((i / TC_STAGES) & 1) ^ 1);
            unsigned char* dst = sB + s * B_TILE;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = h * 128 + pt;
                const bool live = c_begin + (int64_t)i * TC_N + r < c_end;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    const uint32_t b16 = (uint32_t)(cur[h][c >> 2] >> (16 * (c & 3))) & 0xffffu;
                    const uint4 e = live ? expand16(b16, lut) : make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(dst + c * (TC_N * 16) + r * 16) = e;
                }
            }
            fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&b_full[s]);
        }
    } else {
        // ================= epilogue: threshold filter on the dot products =================
        const int ew = warp - (1 + TC_PROD_WARPS);       // 0..7
        const int quarter = warp & 3;                     // TMEM lanes this warp may touch: 32 * (warp % 4)
        const int col_half = ew >> 2;                     // columns [0,128) or [128,256)
        const int qrow = quarter * 32 + lane;             // query row inside the 128-row tile
        uint64_t* my_key = stg_key + ew * (TC_SLOTS * 32) + lane;      // slot s at my_key[s * 32]
        unsigned char* my_t = stg_t + ew * (TC_SLOTS * 32) + lane;
        int n_staged = 0;                                              // this lane's staged hits
        // dist <= thr  <=>  dot = bits - 2 dist >= bits - 2 thr ; padding queries never fire.  Both column halves
        // own the same (t, qrow) entries and write identical values.
        for (int t = 0; t < T; ++t) {
            const int64_t q = q0 + t * TC_M + qrow;
            s_thr[t * TC_M + qrow] = q < a.nq ? a.bits - 2 * a.thr[q] : 0x7fffffff;
        }
        __syncwarp();
        // staged hits -> global candidate lists.  Every lane drains its own slots, so the 32 returning atomics of a
        // trip are in flight together and their latency is paid once per flush instead of once per hit.
        auto flush = [&](int at_least) {
            if (__any_sync(0xffffffffu, n_staged >= at_least)) {
                for (int e = 0; e < n_staged; ++e) {
                    const uint64_t key = my_key[e * 32];
                    const int64_t q = q0 + (int)my_t[e * 32] * TC_M + qrow;
                    const uint32_t pos = atomicAdd(&a.cnt[q], 1u);
                    if (pos < (uint32_t)a.cap) a.cand[q * a.cap + pos] = key;
                    if (a.K > 0) {
                        const int slack = a.thr[q] - (int)((uint32_t)(key >> 32) >> 1);   // thr0 - dist >= 0
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (slack >= j) atomicAdd(&a.aux[q].cum_le[j], 1u);
                    }
                }
                n_staged = 0;
            }
        };
        // 32 dot products of this lane's query: block maxima (ILP 4), one compare; the hit path is rare
        auto scan32 = [&](const int (&x)[32], int thr, int t, int64_t row0) {
            int mb[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) mb[b] = max8(x + 8 * b);
            const int m = __vimax3_s32(mb[0], mb[1], max(mb[2], mb[3]));
            if (m >= thr) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (mb[b] >= thr) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int val = x[8 * b + j];
                            const int64_t row = row0 + 8 * b + j;
                            if (val >= thr && row < c_end)
                                n_staged = tc_stage_hit(my_key, my_t, n_staged,
                                                        ((uint64_t)(uint32_t)(a.bits - val) << 32) |
                                                            (uint64_t)(a.index_base + row),
                                                        t, &a.aux[q0 + t * TC_M + qrow].force_fail);
                        }
                    }
                }
            }
        };
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + col_half * 128;
        int it = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int64_t tile_row0 = c_begin + (int64_t)i * TC_N + col_half * 128;
#pragma unroll 1
            for (int t = 0; t < T; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&t_full[buf], (it >> 1) & 1);
                tc_fence_after();
                const int thr = s_thr[t * TC_M + qrow];
                const uint32_t taddr = lane_addr + buf * TC_N;
                int v0[32], v1[32];
                tmem_ld32(taddr, v0);
#pragma unroll 1
                for (int gp = 0; gp < 2; ++gp) {      // rolled: two static scan sites in the whole epilogue
                    tmem_ld_wait();
                    tmem_ld32(taddr + (2 * gp + 1) * 32, v1);          // in flight while v0 is scanned
                    scan32(v0, thr, t, tile_row0 + (2 * gp) * 32);
                    tmem_ld_wait();
                    if (gp == 0) tmem_ld32(taddr + 64, v0);
                    scan32(v1, thr, t, tile_row0 + (2 * gp + 1) * 32);
                }
                tc_fence_before();
                mbar_arrive(&t_empty[buf]);
                flush(TC_FLUSH_AT);
                if (a.K > 0 && (it & (TC_REFRESH - 1)) == TC_REFRESH - 1) {
                    // tighten: once K rows at dist <= thr0 - j are known, nothing beyond that bucket can be in the top K
                    for (int u = 0; u < T; ++u) {
                        const int64_t q = q0 + u * TC_M + qrow;
                        if (q < a.nq) {
                            const uint4 c = __ldcg(reinterpret_cast<const uint4*>(a.aux[q].cum_le));
                            const uint32_t K = (uint32_t)a.K;
                            const int j = c.w >= K ? 3 : (c.z >= K ? 2 : (c.y >= K ? 1 : 0));
                            s_thr[u * TC_M + qrow] = a.bits - 2 * (a.thr[q] - j);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        flush(1);
    }
    // ---- teardown -----------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- threshold from a sample histogram -------------------------------------------------------------------------------
// hist: [nq][nb] rows of a SAMPLE of n_sample database rows per bucket (bucket = Hamming distance).  Chooses the
// smallest bucket whose cumulative sample count reaches  m = K f + 6 sqrt(K f) + 8  (f = n_sample / nd; Poisson
// margin), or exactly K when the sample is the whole shard.  Any choice is SAFE - a threshold that turns out too low
// is caught by cmh_topk_finalize (fewer than K candidates) - this only sets how often that happens.
__global__ void __launch_bounds__(256) topk_threshold_kernel(const uint32_t* __restrict__ hist, int64_t nq, int nb,
                                                             double need, int32_t* __restrict__ thr) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double cum = 0.0;
    int t = nb - 1;
    for (int b = 0; b < nb; ++b) {
        cum += (double)hist[q * nb + b];
        if (cum >= need) { t = b; break; }
    }
    thr[q] = t;
}

// ---- finalize: exact K-th bucket from the candidates' own histogram, compact, sort, emit -------------------------------
// One CTA per query.  Every database row at or below the query's final threshold was collected, and that threshold
// is an upper bound of the K-th distance, so the smallest bucket T whose cumulative candidate count reaches K is the
// true K-th distance; only candidates with dist <= T (between K and a few K of them) are sorted - by key, i.e. by
// (distance, global index), which is the stable ranking.
constexpr int FIN_MAX = 4096;
constexpr int FIN_BINS = 129;   // bits <= 128 on the tensor path

__global__ void __launch_bounds__(512) topk_finalize_kernel(const uint64_t* __restrict__ cand,
                                                            const uint32_t* __restrict__ cnt,
                                                            const TcAux* __restrict__ aux, int cap, int K, int64_t nd,
                                                            uint64_t* __restrict__ keys,
                                                            uint32_t* __restrict__ fail_flags,
                                                            uint32_t* __restrict__ fail_count) {
    __shared__ uint64_t sk[FIN_MAX];
    __shared__ uint32_t hist[FIN_BINS];
    __shared__ int s_T, s_keep;
    __shared__ uint32_t s_n;
    const int64_t q = blockIdx.x;
    const uint32_t n_found = cnt[q];
    const int64_t need = nd < (int64_t)K ? nd : (int64_t)K;
    bool fail = n_found > (uint32_t)cap || (int64_t)n_found < need || aux[q].force_fail != 0u;
    const int n = (int)n_found;
    const uint64_t* __restrict__ mine = cand + q * cap;
    if (!fail) {
        for (int i = threadIdx.x; i < FIN_BINS; i += blockDim.x) hist[i] = 0u;
        if (threadIdx.x == 0) s_n = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[(uint32_t)(mine[i] >> 33)], 1u);
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t cum = 0;
            int T = -1;
            for (int b = 0; b < FIN_BINS && cum < need; ++b) { cum += hist[b]; T = b; }
            s_T = T;
            s_keep = cum > FIN_MAX ? -1 : (int)cum;
        }
        __syncthreads();
        fail = s_keep < 0;
    }
    if (threadIdx.x == 0) {
        fail_flags[q] = fail ? 1u : 0u;
        if (fail) atomicAdd(fail_count, 1u);
    }
    if (fail) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) keys[q * K + i] = ~0ull;
        return;
    }
    const int T = s_T, keep = s_keep;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = mine[i];
        if ((int)(uint32_t)(key >> 33) <= T) sk[atomicAdd(&s_n, 1u)] = key;
    }
    int p2 = 1;
    while (p2 < keep) p2 <<= 1;
    __syncthreads();
    for (int i = keep + threadIdx.x; i < p2; i += blockDim.x) sk[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t x = sk[i], y = sk[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sk[i] = y; sk[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x) keys[q * K + i] = i < keep ? sk[i] : ~0ull;
}

}  // namespace cmh

using namespace cmh;

// B-ring depth: 6 stages for 64-bit codes (128 KB of shared memory per CTA, so exactly one CTA - which owns the whole
// TMEM - is resident per SM), 4 for 128-bit codes (192 KB).
static int tc_stages(int words) { return words == 1 ? 6 : 4; }
static size_t tc_smem_bytes(int words, int T) {
    const int st = tc_stages(words);
    return (size_t)T * TC_M * words * 64 + (size_t)st * TC_N * words * 64 + 64 + (2 * st + 4) * 8 + 8 +
           (size_t)TC_EPI_WARPS * TC_SLOTS * 32 * 9 + (size_t)T * TC_M * 4 + 16;
}

extern "C" int cmh_tc_supported(int bits, int ternary) {
    return (!ternary && (bits == 64 || bits == 128)) ? 1 : 0;
}

extern "C" int cmh_tc_collect(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                              int64_t index_base, const int32_t* thr, int K, int cap, uint64_t* cand, uint32_t* cnt,
                              uint32_t* aux, void* stream) {
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_collect: bits=%d (64 or 128, +-1 codes only)", bits);
    CMH_REQUIRE(nq >= 0 && nd >= 0 && cap >= 1 && K >= 0 && index_base >= 0 && index_base + nd <= (1ll << 32), CMH_ERR_ARG,
                "cmh_tc_collect: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(q_sign && thr && cand && cnt && aux, CMH_ERR_ARG, "cmh_tc_collect: NULL pointer");
    static_assert(sizeof(TcAux) == 32, "cmh_tc_collect: aux is uint32 [nq][8]");
    CMH_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nq * 4, st));
    CMH_CUDA(cudaMemsetAsync(aux, 0, (size_t)nq * sizeof(TcAux), st));
    if (nd == 0) return CMH_OK;
    CMH_REQUIRE(d_sign, CMH_ERR_ARG, "cmh_tc_collect: NULL database");
    const int words = bits / 64;
    const int T = TC_MAX_T;
    const int64_t n_qgroups = ceil_div(nq, (int64_t)T * TC_M);
    // enough chunks for ~4 CTAs per SM over the launch, chunk = whole tiles
    int64_t n_chunks = std::max<int64_t>(1, ceil_div((int64_t)sm_count() * 4, n_qgroups));
    int64_t chunk_rows = round_up(ceil_div(nd, n_chunks), TC_N);
    chunk_rows = std::max<int64_t>(chunk_rows, 16 * TC_N);
    n_chunks = ceil_div(nd, chunk_rows);
    CMH_REQUIRE(n_chunks <= 65535 && n_qgroups <= 0x7fffffffll && chunk_rows <= 0x7fffffff, CMH_ERR_UNSUPPORTED,
                "cmh_tc_collect: launch geometry out of range");
    TcArgs a;
    a.q = q_sign; a.d = d_sign; a.thr = thr; a.cand = cand; a.cnt = cnt; a.aux = reinterpret_cast<TcAux*>(aux);
    a.nq = nq; a.nd = nd; a.index_base = index_base; a.chunk_rows = (int)chunk_rows; a.cap = cap; a.bits = bits;
    a.K = K;
    const size_t smem = tc_smem_bytes(words, T);
    const dim3 grid((unsigned)n_qgroups, (unsigned)n_chunks);
    if (words == 1) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, TC_MAX_T, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, TC_MAX_T, 6><<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<2, TC_MAX_T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<2, TC_MAX_T, 4><<<grid, TC_THREADS, smem, st>>>(a);
    }
    CMH_LAUNCH_CHECK("tc_collect_kernel");
    return CMH_OK;
}

extern "C" int cmh_topk_threshold(const uint32_t* hist, int64_t nq, int nb, int64_t n_sample, int64_t nd, int K,
                                  int32_t* thr, void* stream) {
    CMH_REQUIRE(nq >= 0 && nb >= 1 && n_sample >= 0 && nd >= n_sample && K >= 1, CMH_ERR_ARG, "cmh_topk_threshold: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(hist && thr, CMH_ERR_ARG, "cmh_topk_threshold: NULL pointer");
    double need;
    if (n_sample >= nd) {
        need = (double)std::min<int64_t>(K, nd);
    } else {
        const double kf = (double)K * (double)n_sample / (double)nd;
        need = std::min((double)n_sample, kf + 6.0 * std::sqrt(kf) + 8.0);
    }
    topk_threshold_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(hist, nq, nb, need, thr);
    CMH_LAUNCH_CHECK("topk_threshold_kernel");
    return CMH_OK;
}

extern "C" int cmh_topk_finalize(const uint64_t* cand, const uint32_t* cnt, const uint32_t* aux, int64_t nq, int cap,
                                 int K, int64_t nd, uint64_t* keys, uint32_t* fail_flags, uint32_t* fail_count,
                                 void* stream) {
    CMH_REQUIRE(nq >= 0 && cap >= 1 && K >= 1 && nd >= 0, CMH_ERR_ARG, "cmh_topk_finalize: bad sizes");
    CMH_REQUIRE(K <= FIN_MAX, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: K=%d > %d", K, FIN_MAX);
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(cand && cnt && aux && keys && fail_flags && fail_count, CMH_ERR_ARG, "cmh_topk_finalize: NULL pointer");
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: too many queries per call");
    cudaStream_t st = (cudaStream_t)stream;
    CMH_CUDA(cudaMemsetAsync(fail_count, 0, 4, st));
    topk_finalize_kernel<<<(unsigned)nq, 512, 0, st>>>(cand, cnt, reinterpret_cast<const TcAux*>(aux), cap, K, nd, keys,
                                                       fail_flags, fail_count);
    CMH_LAUNCH_CHECK("topk_finalize_kernel");
    return CMH_OK;
}
