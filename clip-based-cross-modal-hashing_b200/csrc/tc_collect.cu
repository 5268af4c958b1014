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

// Pipeline trace (debug builds only, -DCMH_TC_TRACE): CTA (0,0) records clock64() at the hand-offs of its first
// TC_TRACE_ITERS iterations into the buffer registered with cmh_tc_set_trace: [role][iteration][event].
#ifdef CMH_TC_TRACE
#define TC_TRACE_ITERS 96
#define TC_TRACE_EVENTS 8
#define TC_TRACE(role, iter, ev)                                                                         \
    do {                                                                                                 \
        if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (iter) < TC_TRACE_ITERS)                    \
            a.trace[((role) * TC_TRACE_ITERS + (iter)) * TC_TRACE_EVENTS + (ev)] = clock64();            \
    } while (0)
#else
#define TC_TRACE(role, iter, ev) do {} while (0)
#endif

namespace cmh {

constexpr int TC_M = 128;       // queries per MMA (TMEM lanes)
constexpr int TC_N = 256;       // database rows per shared-memory stage (two MMAs of TC_NM columns)
constexpr int TC_NM = 128;      // database rows per MMA = TMEM columns per accumulator buffer
constexpr int TC_BUFS = 4;      // accumulator buffers (4 x 128 columns = the whole TMEM)
constexpr int TC_MMA_WARPS = 4;   // one MMA issuer (elected lane) per accumulator buffer
constexpr int TC_PROD_WARPS = 4;
constexpr int TC_EPI_WARPS = 16;  // 4 groups of 4 warps; group g drains buffer g
constexpr int TC_GROUPS = TC_EPI_WARPS / 4;
constexpr int TC_THREADS = (TC_MMA_WARPS + TC_PROD_WARPS + TC_EPI_WARPS) * 32;
constexpr int TC_RING = 8;        // packed database tiles in flight (bulk copies)
constexpr int TC_MAX_CHUNKS = 1024;  // candidate segments per query (cmh_topk_finalize walks them)
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
// 64 consecutive columns, two per register: column 2i in the low half of v[i], column 2i+1 in the high half (the
// int32 accumulators are small - |dot| <= bits <= 128 - so their low 16 bits are the exact int16 value)
__device__ __forceinline__ void tmem_ld64p(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
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

// 4 bits -> 4 bytes of +-1 (bit set -> +1 = 0x01, clear -> -1 = 0xFF), ALU only: the multiply spreads bit j to bit 8j
// (the four shifted copies of a nibble do not overlap, so the product is their OR), and 0xFF - 0xFE * {0,1} per byte
// never borrows.
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
    const uint32_t x = (nib * 0x00204081u) & 0x01010101u;
    return 0xFFFFFFFFu - x * 0xFEu;
}
__device__ __forceinline__ uint4 expand16(uint32_t bits16) {
    uint4 o;
    o.x = expand4(bits16 & 15u);
    o.y = expand4((bits16 >> 4) & 15u);
    o.z = expand4((bits16 >> 8) & 15u);
    o.w = expand4(bits16 >> 12);
    return o;
}

constexpr int TC_REFRESH = 32;     // tiles (per epilogue group) between threshold refreshes

// per-query bookkeeping shared by every CTA working on the query (global memory, zeroed by cmh_tc_collect)
struct TcAux {
    uint32_t h[4];        // candidates seen so far at dist == thr0 - j (j = 0, 1, 2) and at dist <= thr0 - 3 (j = 3)
    uint32_t force_fail;  // a candidate segment overflowed: entries were dropped
    uint32_t pad[3];
};

struct TcArgs {
    const uint64_t* q;      // [nq][words]
    const uint64_t* d;      // [nd][words]
    const int32_t* thr;     // [nq] initial threshold bucket thr0 (Hamming distance): rows with dist <= thr qualify
    uint64_t* cand;         // [nq][n_chunks][seg_cap]  one private candidate segment per (query, database chunk)
    uint32_t* cnt;          // [n_chunks][nq]           candidates the chunk's CTA found (may exceed seg_cap)
    TcAux* aux;             // [nq]
    int64_t nq, nd, index_base;
    int chunk_rows;         // database rows per CTA (multiple of TC_N)
    int n_chunks, seg_cap, bits;
    int K;                  // > 0: tighten thresholds while scanning (once K rows at dist <= thr0 - j are known)
    long long* trace;       // CMH_TC_TRACE builds only
    int probe;              // measurement aid (cmh_tc_probe): 1 = no tcgen05.mma, 2 = no TMEM drain, 4 = drain without scan
};

// per-half maximum of 8 registers of packed int16 pairs (VIMNMX3.S16x2)
__device__ __forceinline__ uint32_t max8p(const uint32_t* v) {
    const uint32_t m = __vimax3_s16x2(v[0], v[1], v[2]);
    const uint32_t n = __vimax3_s16x2(v[3], v[4], v[5]);
    return __vimax3_s16x2(m, n, __vmaxs2(v[6], v[7]));
}

// The hit path.  Out of line on purpose: one copy of the append keeps the epilogue's instruction footprint small (an
// inlined, fully unrolled hit path was ~160 KB of SASS and every rare hit paid a chain of instruction-cache misses).
// The slot comes from a shared-memory counter (the segment is private to this CTA), the key goes straight to global
// memory and the tightening statistics are a fire-and-forget RED: nothing on this path waits for global memory.
__device__ __noinline__ void tc_append(uint64_t* seg, uint32_t* pos_ctr, uint32_t seg_cap, uint32_t* h, int slack,
                                       uint32_t key_hi, uint32_t key_lo) {
    const uint32_t pos = atomicAdd(pos_ctr, 1u);
    if (pos < seg_cap) seg[pos] = ((uint64_t)key_hi << 32) | key_lo;
    if (h != nullptr) atomicAdd(h + min(slack, 3), 1u);
}

// smem: [A: T tiles][B: STAGES tiles][packed ring][barriers][tmem slot][thr][thr0][pos]
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
    uint64_t* ring = reinterpret_cast<uint64_t*>(sB + TC_STAGES * B_TILE);   // [TC_RING][256 rows][WORDS] packed words
    uint64_t* bars = ring + TC_RING * TC_N * WORDS;
    uint64_t* b_full = bars;                     // [STAGES] producers -> MMA
    uint64_t* b_empty = bars + TC_STAGES;        // [STAGES] MMA -> producers
    uint64_t* t_full = bars + 2 * TC_STAGES;               // [BUFS] MMA -> epilogue group
    uint64_t* t_empty = bars + 2 * TC_STAGES + TC_BUFS;    // [BUFS] epilogue group -> MMA
    uint64_t* r_full = bars + 2 * TC_STAGES + 2 * TC_BUFS;               // [RING] bulk copy -> producers
    uint64_t* r_empty = bars + 2 * TC_STAGES + 2 * TC_BUFS + TC_RING;    // [RING] producers -> bulk copy issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_BUFS + 2 * TC_RING);
    int* s_thr = reinterpret_cast<int*>(tmem_slot + 2);    // [T][128] current dot-product thresholds
    int* s_thr0 = s_thr + T * TC_M;                        // [T][128] dot-product threshold of thr0
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_thr0 + T * TC_M);   // [T][128] entries appended by this CTA

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * (T * TC_M);
    const int64_t c_begin = (int64_t)blockIdx.y * a.chunk_rows;
    const int64_t c_end = min(a.nd, c_begin + a.chunk_rows);
    const int n_tiles = (int)((c_end - c_begin + TC_N - 1) / TC_N);

    // ---- prologue -----------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&b_full[s], TC_PROD_WARPS * 32);
            mbar_init(&b_empty[s], TC_MMA_WARPS);
        }
        for (int b = 0; b < TC_BUFS; ++b) {
            mbar_init(&t_full[b], 1);
            mbar_init(&t_empty[b], 128);         // the 4 warps that drain one accumulator tile
        }
        for (int r = 0; r < TC_RING; ++r) {
            mbar_init(&r_full[r], 1);
            mbar_init(&r_empty[r], TC_PROD_WARPS * 32);
        }
        mbar_fence_init();
    }
    // dist <= thr  <=>  dot = bits - 2 dist >= bits - 2 thr ; padding queries never fire (int16 max)
    for (int e = tid; e < T * TC_M; e += TC_THREADS) {
        const int64_t q = q0 + e;
        const int v = q < a.nq ? a.bits - 2 * a.thr[q] : 0x7fff;
        s_thr[e] = v;
        s_thr0[e] = v;
        s_pos[e] = 0u;
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
        *reinterpret_cast<uint4*>(sA + t * A_TILE + c * (TC_M * 16) + r * 16) = expand16(b16);
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_MMA_WARPS) {
        // ================= MMA issuers =================
        // Issuer w owns accumulator buffer w: tiles w, w + 4, ...  One thread doing wait -> tcgen05.mma -> commit for
        // every tile spends ~300 cycles per tile on instruction latency alone (measured), more than the 128 cycles
        // the tensor pipe needs for it; four issuers overlap that latency.
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_M, TC_NM);
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            const bool no_mma = (a.probe & 1) != 0;
            const int n_iters = n_tiles * T * 2;     // iteration = (stage i, query tile t, row half h)
            const uint32_t d_tmem = tmem_base + warp * TC_NM;
            int round = 0;
#pragma unroll 1
            for (int it = warp; it < n_iters; it += TC_MMA_WARPS, ++round) {
                const int i = it / (2 * T), rem = it - i * (2 * T);
                const int t = rem >> 1, h = rem & 1;
                const int s = i % TC_STAGES;
                TC_TRACE(0, it, 3);
                if (rem < TC_MMA_WARPS) {            // this issuer's first tile of the stage
                    mbar_wait(&b_full[s], (i / TC_STAGES) & 1);
                    tc_fence_after();
                }
                TC_TRACE(0, it, 0);
                mbar_wait(&t_empty[warp], (round & 1) ^ 1);
                tc_fence_after();
                TC_TRACE(0, it, 1);
                if (!no_mma) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                        const uint64_t ad = umma_desc(a_addr + t * A_TILE + k * 2 * (TC_M * 16), TC_M * 16, 128);
                        const uint64_t bd = umma_desc(b_addr + s * B_TILE + k * 2 * (TC_N * 16) + h * (TC_NM * 16),
                                                      TC_N * 16, 128);
                        umma_i8(d_tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
                    }
                }
                umma_commit(&t_full[warp]);
                if (rem >= 2 * T - TC_MMA_WARPS) umma_commit(&b_empty[s]);   // ... and its last one
                TC_TRACE(0, it, 2);
            }
        }
    } else if (warp < TC_MMA_WARPS + TC_PROD_WARPS) {
        // ================= producers: packed bits -> +-1 int8 core matrices =================
        // The packed words of a tile (256 rows, 2-4 KB) are staged in a TC_RING-deep shared-memory ring by the bulk-copy
        // engine (cp.async.bulk, issued by one thread, completion on an mbarrier).  The producer threads must not have
        // global loads of their own in flight: fence.proxy.async is a MEMBAR that waits for them, which bounded the
        // kernel by one tile per global-memory round trip.  Tiles the engine cannot take (a ragged last tile, a shard
        // view that is not 16-byte aligned) are read with plain loads.
        const int pt = tid - TC_MMA_WARPS * 32;  // 0..127
        const bool aligned = (reinterpret_cast<uintptr_t>(a.d) & 15) == 0;
        auto bulk_tile = [&](int i) { return aligned && c_begin + (int64_t)(i + 1) * TC_N <= c_end; };
        auto issue = [&](int i) {                // pt == 0
            if (i >= n_tiles) return;
            const int slot = i % TC_RING;
            if (bulk_tile(i)) {
                mbar_expect_tx(&r_full[slot], TC_N * WORDS * 8);
                bulk_g2s(ring + slot * (TC_N * WORDS), a.d + (c_begin + (int64_t)i * TC_N) * WORDS, TC_N * WORDS * 8,
                         &r_full[slot]);
            } else {
                mbar_arrive(&r_full[slot]);
            }
        };
        if (pt == 0)
            for (int j = 0; j < TC_RING; ++j) issue(j);
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % TC_STAGES, slot = i % TC_RING;
            uint64_t cur[2][WORDS];
            if (pt == 0) TC_TRACE(1, i, 0);
            mbar_wait(&r_full[slot], (i / TC_RING) & 1);
            if (pt == 0) TC_TRACE(1, i, 1);
            if (bulk_tile(i)) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int x = 0; x < WORDS; ++x) cur[h][x] = ring[slot * (TC_N * WORDS) + (h * 128 + pt) * WORDS + x];
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t row = c_begin + (int64_t)i * TC_N + h * 128 + pt;
#pragma unroll
                    for (int x = 0; x < WORDS; ++x) cur[h][x] = row < c_end ? __ldg(a.d + row * WORDS + x) : 0ull;
                }
            }
            mbar_arrive(&r_empty[slot]);
            if (pt == 0) {
                mbar_wait(&r_empty[slot], (i / TC_RING) & 1);
                issue(i + TC_RING);              // refill the slot everyone has just read
            }
            if (pt == 0) TC_TRACE(1, i, 2);
            mbar_wait(&b_empty[s], ((i / TC_STAGES) & 1) ^ 1);
            if (pt == 0) TC_TRACE(1, i, 3);
            unsigned char* dst = sB + s * B_TILE;
            // rows past the end of the chunk are expanded like any other (as all -1): the hit path checks the row
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = h * 128 + pt;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    const uint32_t b16 = (uint32_t)(cur[h][c >> 2] >> (16 * (c & 3))) & 0xffffu;
                    *reinterpret_cast<uint4*>(dst + c * (TC_N * 16) + r * 16) = expand16(b16);
                }
            }
            if (pt == 0) TC_TRACE(1, i, 4);
            fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
            if (pt == 0) TC_TRACE(1, i, 5);
            mbar_arrive(&b_full[s]);
            if (pt == 0) TC_TRACE(1, i, 6);
        }
    } else {
        // ================= epilogue: threshold filter on the dot products =================
        const int ew = warp - (TC_MMA_WARPS + TC_PROD_WARPS);   // 0 .. TC_EPI_WARPS-1
        const int grp = ew >> 2;                          // group of 4 warps (one per TMEM lane quarter)
        const int quarter = warp & 3;                     // TMEM lanes this warp may touch: 32 * (warp % 4)
        const int qrow = quarter * 32 + lane;             // query row inside the 128-row tile
        const bool no_drain = (a.probe & 2) != 0, no_scan = (a.probe & 4) != 0;
        // Rare: at least one of the 64 dot products in x reaches the threshold.  Two-level search (4 block maxima,
        // then the 8 registers of a block), both halves of a register by a rolled loop: 32 call sites in all.
        auto collect64 = [&](const uint32_t (&x)[32], const uint32_t (&mb)[4], int thr, uint32_t thr2m, int t,
                             int64_t row0) {
            const int e = t * TC_M + qrow;
            const int64_t q = q0 + e;
            uint64_t* seg = a.cand + ((uint64_t)q * a.n_chunks + blockIdx.y) * (uint64_t)a.seg_cap;
            uint32_t* h = a.K > 0 ? a.aux[q].h : nullptr;
            const int dthr0 = s_thr0[e];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (__vmaxs2(mb[b], thr2m) == thr2m) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t w = x[8 * b + j];
                    if (__vmaxs2(w, thr2m) == thr2m) continue;
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        const int val = hh ? ((int)w >> 16) : ((int)(w << 16) >> 16);
                        const int64_t row = row0 + 16 * b + 2 * j + hh;
                        if (val >= thr && row < c_end)
                            tc_append(seg, &s_pos[e], (uint32_t)a.seg_cap, h, (val - dthr0) >> 1,
                                      (uint32_t)(a.bits - val), (uint32_t)(a.index_base + row));
                    }
                }
            }
        };
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int n_iters = n_tiles * T * 2;
        int round = 0;
#pragma unroll 1
        for (int it = grp; it < n_iters; it += TC_GROUPS, ++round) {
            const int i = it / (2 * T), rem = it - i * (2 * T);
            const int t = rem >> 1, h = rem & 1;
            const int buf = it & (TC_BUFS - 1);
            const int64_t row0 = c_begin + (int64_t)i * TC_N + h * TC_NM;
            if (qrow == 0) TC_TRACE(2 + grp, round, 0);
            mbar_wait(&t_full[buf], (it / TC_BUFS) & 1);
            tc_fence_after();
            if (qrow == 0) TC_TRACE(2 + grp, round, 1);
            const int thr = s_thr[t * TC_M + qrow];
            const uint32_t thr2m = (uint32_t)((thr - 1) & 0xffff) * 0x10001u;   // "any half > thr - 1"
            const uint32_t taddr = lane_taddr + buf * TC_NM;
            if (no_drain) {
                tc_fence_before();
                mbar_arrive(&t_empty[buf]);
                if (qrow == 0) TC_TRACE(2 + grp, round, 4);
                continue;
            }
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {       // rolled: one scan site (and one hit path) in the kernel
                uint32_t v[32];
                tmem_ld64p(taddr + half * 64, v);
                tmem_ld_wait();
                if (qrow == 0) TC_TRACE(2 + grp, round, 2 + half);
                if (half == 1) {
                    tc_fence_before();
                    mbar_arrive(&t_empty[buf]);          // the values are in registers: the tile can be overwritten
                }
                if (no_scan) {
                    uint32_t o = 0;
#pragma unroll
                    for (int r = 0; r < 32; ++r) o |= v[r];
                    if (o == 0xdeadbeefu) s_pos[0] = o;  // keep the loads alive
                    continue;
                }
                uint32_t mb[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) mb[b] = max8p(v + 8 * b);
                const uint32_t m = __vimax3_s16x2(mb[0], mb[1], __vmaxs2(mb[2], mb[3]));
                if (__vmaxs2(m, thr2m) != thr2m) collect64(v, mb, thr, thr2m, t, row0 + half * 64);
            }
            if (qrow == 0) TC_TRACE(2 + grp, round, 4);
            if (a.K > 0 && (round & (TC_REFRESH - 1)) == TC_REFRESH - 1) {
                // tighten: once K rows at dist <= thr0 - j are known, nothing beyond that bucket can be in the top K
                for (int u = 0; u < T; ++u) {
                    const int e = u * TC_M + qrow;
                    const int64_t q = q0 + e;
                    if (q < a.nq) {
                        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(a.aux[q].h));
                        const uint32_t K = (uint32_t)a.K;
                        const uint32_t c3 = c.w, c2 = c3 + c.z, c1 = c2 + c.y;
                        const int j = c3 >= K ? 3 : (c2 >= K ? 2 : (c1 >= K ? 1 : 0));
                        s_thr[e] = s_thr0[e] + 2 * j;   // the four groups store the same or a newer (tighter) value
                    }
                }
                __syncwarp();
            }
        }
    }
    // ---- teardown -----------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    for (int e = tid; e < T * TC_M; e += TC_THREADS) {
        const int64_t q = q0 + e;
        if (q < a.nq) {
            const uint32_t n = s_pos[e];
            a.cnt[(int64_t)blockIdx.y * a.nq + q] = n;
            if (n > (uint32_t)a.seg_cap) a.aux[q].force_fail = 1u;
        }
    }
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
// (distance, global index), which is the stable ranking.  A warp walks one (query, chunk) segment at a time.
constexpr int FIN_MAX = 4096;
constexpr int FIN_BINS = 129;   // bits <= 128 on the tensor path
constexpr int FIN_THREADS = 512;

__global__ void __launch_bounds__(FIN_THREADS) topk_finalize_kernel(const uint64_t* __restrict__ cand,
                                                                    const uint32_t* __restrict__ cnt,
                                                                    const TcAux* __restrict__ aux, int64_t nq,
                                                                    int n_chunks, int seg_cap, int K, int64_t nd,
                                                                    uint64_t* __restrict__ keys,
                                                                    uint32_t* __restrict__ fail_flags,
                                                                    uint32_t* __restrict__ fail_count) {
    __shared__ uint64_t sk[FIN_MAX];
    __shared__ uint32_t hist[FIN_BINS];
    __shared__ int s_T, s_keep;
    __shared__ uint32_t s_n, s_total, s_over;
    const int64_t q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = FIN_THREADS / 32;
    const int64_t need = nd < (int64_t)K ? nd : (int64_t)K;
    const uint64_t* __restrict__ mine = cand + (uint64_t)q * n_chunks * (uint64_t)seg_cap;
    for (int i = threadIdx.x; i < FIN_BINS; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x == 0) { s_n = 0u; s_total = 0u; s_over = 0u; }
    __syncthreads();
    for (int c = warp; c < n_chunks; c += NW) {
        uint32_t n = cnt[(int64_t)c * nq + q];
        if (lane == 0) {
            atomicAdd(&s_total, n);
            if (n > (uint32_t)seg_cap) s_over = 1u;
        }
        n = min(n, (uint32_t)seg_cap);
        const uint64_t* seg = mine + (uint64_t)c * seg_cap;
        for (uint32_t i = lane; i < n; i += 32) atomicAdd(&hist[(uint32_t)(seg[i] >> 33)], 1u);
    }
    __syncthreads();
    bool fail = s_over != 0u || (int64_t)s_total < need || aux[q].force_fail != 0u;
    if (!fail) {
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
    for (int c = warp; c < n_chunks; c += NW) {
        const uint32_t n = cnt[(int64_t)c * nq + q];     // <= seg_cap (checked above)
        const uint64_t* seg = mine + (uint64_t)c * seg_cap;
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t key = seg[i];
            if ((int)(uint32_t)(key >> 33) <= T) sk[atomicAdd(&s_n, 1u)] = key;
        }
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

// B-ring depth: 6 stages for 64-bit codes (>114 KB of shared memory per CTA, so exactly one CTA - which owns the whole
// TMEM - is resident per SM), 3 for 128-bit codes (64 KB of A + 96 KB of B).
static int tc_stages(int words) { return words == 1 ? 6 : 3; }
static size_t tc_smem_bytes(int words, int T) {
    const int st = tc_stages(words);
    return (size_t)T * TC_M * words * 64 + (size_t)st * TC_N * words * 64 + 64 + (2 * st + 2 * TC_BUFS + 2 * TC_RING) * 8 + 8 +
           (size_t)3 * T * TC_M * 4 + (size_t)TC_RING * 2 * words * 128 * 8 + 16;
}

extern "C" int cmh_tc_supported(int bits, int ternary) {
    return (!ternary && (bits == 64 || bits == 128)) ? 1 : 0;
}

// launch geometry: query groups of TC_MAX_T x 128 rows, database chunks of whole tiles, ~4 CTAs per SM over the launch
static void tc_geometry(int64_t nq, int64_t nd, int64_t* n_qgroups, int64_t* n_chunks, int64_t* chunk_rows) {
    *n_qgroups = std::max<int64_t>(1, ceil_div(nq, (int64_t)TC_MAX_T * TC_M));
    int64_t want = std::max<int64_t>(1, ceil_div((int64_t)sm_count() * 4, *n_qgroups));
    want = std::min<int64_t>(want, TC_MAX_CHUNKS);
    int64_t rows = round_up(std::max<int64_t>(1, ceil_div(nd, want)), TC_N);
    rows = std::max<int64_t>(rows, 16 * TC_N);
    *chunk_rows = rows;
    *n_chunks = std::max<int64_t>(1, ceil_div(nd, rows));
}

static long long* g_tc_trace = nullptr;
#ifdef CMH_TC_TRACE
extern "C" int cmh_tc_set_trace(void* device_buffer) { g_tc_trace = (long long*)device_buffer; return CMH_OK; }
#endif

extern "C" int cmh_tc_plan(int64_t nq, int64_t nd, int bits, int* n_chunks) {
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_plan: bits=%d (64 or 128, +-1 codes only)", bits);
    CMH_REQUIRE(nq >= 0 && nd >= 0 && n_chunks, CMH_ERR_ARG, "cmh_tc_plan: bad arguments");
    int64_t g, c, r;
    tc_geometry(nq, nd, &g, &c, &r);
    *n_chunks = (int)c;
    return CMH_OK;
}

static int tc_collect_impl(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                           int64_t index_base, const int32_t* thr, int K, int n_chunks_in, int seg_cap, uint64_t* cand,
                           uint32_t* cnt, uint32_t* aux, int probe, void* stream) {
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_collect: bits=%d (64 or 128, +-1 codes only)", bits);
    CMH_REQUIRE(nq >= 0 && nd >= 0 && seg_cap >= 1 && K >= 0 && index_base >= 0 && index_base + nd <= (1ll << 32),
                CMH_ERR_ARG, "cmh_tc_collect: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (nq == 0) return CMH_OK;
    int64_t n_qgroups, n_chunks, chunk_rows;
    tc_geometry(nq, nd, &n_qgroups, &n_chunks, &chunk_rows);
    CMH_REQUIRE(n_chunks_in == (int)n_chunks, CMH_ERR_ARG, "cmh_tc_collect: n_chunks=%d, cmh_tc_plan says %d",
                n_chunks_in, (int)n_chunks);
    CMH_REQUIRE(q_sign && thr && cand && cnt && aux, CMH_ERR_ARG, "cmh_tc_collect: NULL pointer");
    static_assert(sizeof(TcAux) == 32, "cmh_tc_collect: aux is uint32 [nq][8]");
    CMH_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_chunks * nq * 4, st));
    CMH_CUDA(cudaMemsetAsync(aux, 0, (size_t)nq * sizeof(TcAux), st));
    if (nd == 0) return CMH_OK;
    CMH_REQUIRE(d_sign, CMH_ERR_ARG, "cmh_tc_collect: NULL database");
    CMH_REQUIRE(n_qgroups <= 0x7fffffffll && chunk_rows <= 0x7fffffff, CMH_ERR_UNSUPPORTED,
                "cmh_tc_collect: launch geometry out of range");
    const int words = bits / 64;
    const int T = TC_MAX_T;
    TcArgs a;
    a.q = q_sign; a.d = d_sign; a.thr = thr; a.cand = cand; a.cnt = cnt; a.aux = reinterpret_cast<TcAux*>(aux);
    a.nq = nq; a.nd = nd; a.index_base = index_base; a.chunk_rows = (int)chunk_rows; a.n_chunks = (int)n_chunks;
    a.seg_cap = seg_cap; a.bits = bits; a.K = K; a.probe = probe; a.trace = g_tc_trace;
    const size_t smem = tc_smem_bytes(words, T);
    const dim3 grid((unsigned)n_qgroups, (unsigned)n_chunks);
    if (words == 1) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, TC_MAX_T, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, TC_MAX_T, 6><<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<2, TC_MAX_T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<2, TC_MAX_T, 3><<<grid, TC_THREADS, smem, st>>>(a);
    }
    CMH_LAUNCH_CHECK("tc_collect_kernel");
    return CMH_OK;
}

extern "C" int cmh_tc_collect(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                              int64_t index_base, const int32_t* thr, int K, int n_chunks, int seg_cap, uint64_t* cand,
                              uint32_t* cnt, uint32_t* aux, void* stream) {
    return tc_collect_impl(q_sign, nq, d_sign, nd, bits, index_base, thr, K, n_chunks, seg_cap, cand, cnt, aux, 0, stream);
}

extern "C" int cmh_tc_probe(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                            const int32_t* thr, int n_chunks, int seg_cap, uint64_t* cand, uint32_t* cnt, uint32_t* aux,
                            int probe, void* stream) {
    CMH_REQUIRE(probe >= 0 && probe < 16, CMH_ERR_ARG, "cmh_tc_probe: probe=%d", probe);
    return tc_collect_impl(q_sign, nq, d_sign, nd, bits, 0, thr, 0, n_chunks, seg_cap, cand, cnt, aux, probe, stream);
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

extern "C" int cmh_topk_finalize(const uint64_t* cand, const uint32_t* cnt, const uint32_t* aux, int64_t nq,
                                 int n_chunks, int seg_cap, int K, int64_t nd, uint64_t* keys, uint32_t* fail_flags,
                                 uint32_t* fail_count, void* stream) {
    CMH_REQUIRE(nq >= 0 && n_chunks >= 1 && seg_cap >= 1 && K >= 1 && nd >= 0, CMH_ERR_ARG, "cmh_topk_finalize: bad sizes");
    CMH_REQUIRE(K <= FIN_MAX, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: K=%d > %d", K, FIN_MAX);
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(cand && cnt && aux && keys && fail_flags && fail_count, CMH_ERR_ARG, "cmh_topk_finalize: NULL pointer");
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: too many queries per call");
    cudaStream_t st = (cudaStream_t)stream;
    CMH_CUDA(cudaMemsetAsync(fail_count, 0, 4, st));
    topk_finalize_kernel<<<(unsigned)nq, FIN_THREADS, 0, st>>>(cand, cnt, reinterpret_cast<const TcAux*>(aux), nq, n_chunks,
                                                              seg_cap, K, nd, keys, fail_flags, fail_count);
    CMH_LAUNCH_CHECK("topk_finalize_kernel");
    return CMH_OK;
}
