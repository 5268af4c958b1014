// K3 - the tensor-core variant of the query x database Hamming compare (north star (2)):
// the reference's identity  dist = 0.5 * (bits - qB . rB^T)  (utils/calc_utils.py:12) evaluated as a +-1 int8 GEMM
// on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), with the top-K candidate filter
// fused as the epilogue so that the Q x D distance matrix never exists.
//
// One CTA = T x 128 queries x one contiguous database chunk, walked in tiles of 256 rows.
//
// Two database rows per accumulator.  The pipeline is bound by the round trip of an accumulator buffer (fill, drain
// through tcgen05.ld, hand back) over the 512 TMEM columns, so every 32-bit accumulator carries TWO dot products and
// the per-query threshold, all produced by the tensor core:
//     acc[q][j] = bias_q + <q, d_j> + B * <q, d_{j+128}>            (j = 0..127 inside a 256-row tile)
// from 2 * KSTEPS + 1 K-steps into the same TMEM columns: A = +-1 bytes against rows 0..127 expanded to +-1, then
// A' = +-S bytes against rows 128..255 expanded to +-S (S * S = B), then one K-step of per-query digits against a
// constant row of weights whose product is  bias_q = B - (B + 1) * T_q  (T_q = the dot product a row needs to qualify).
// With e1 = dot1 - T and e2 = dot2 - T + 1,  acc = e1 + B * e2  and, F = log2 B,
//     bit F-1  of acc clear  <=>  e1 >= 0  <=>  row j qualifies
//     bit 2F-1 of acc clear  <=>  e2 - [e1 < 0] >= 0  <=>  row j+128 qualifies  (e2 is odd: dot products have the parity
//                                                                              of the code length, so e2 >= 0 <=> e2 >= 1)
// so the filter is an AND-reduction of the registers and one mask test per slice.  64-bit codes: F = 8, S = 16, acc
// fits 16 bits and tcgen05.ld packs two columns per register (4 rows per register); 128-bit codes: F = 10, S = 32.
// The fields cannot wrap (|e| < 2^(F-1) for 1 <= T <= bits).
// The fields of a flagged accumulator decode to the exact distances, which are compared with the query's CURRENT
// threshold - it tightens while the scan proceeds - before the row is appended.
//
// Warp roles (warp-specialised, mbarrier pipelines, no __syncthreads in the main loop):
//   warps 0-3     MMA issuers, one per accumulator buffer (an elected lane: wait, 2 * KSTEPS + 1 tcgen05.mma, commit)
//   warps 4-7     producers: packed code words arrive through a bulk-copy (cp.async.bulk) ring; every bit is expanded
//                 to a +-1 / +-S byte with two integer multiplies per nibble and stored as 8x16 B core matrices
//                 (no-swizzle K-major UMMA layout) into a STAGES-deep ring of B tiles
//   warps 8-23    epilogue: 4 groups of 4 warps, group g drains accumulator buffer g; thread = one query (a TMEM
//                 lane) for the whole CTA, so its threshold, candidate count and segment pointer live in registers;
//                 slices with a flagged row are parked in a per-warp shared-memory queue and worked off - by the whole
//                 warp, one entry at a time - while the warp waits for its next tile
//   warps 24-27   (experimental variant WK only) hit workers: take the parked slices off the queues instead
// TMEM: 4 accumulator buffers x 128 columns (the whole 512-column TMEM, one CTA per SM).
//
// Exactness: thresholds only have to be upper bounds of the K-th distance (cmh_topk_threshold derives them from a
// sample histogram, cmh_tc_cand_hist + cmh_tc_choose refine them from the pilot launches, cmh_tc_choose_prefix /
// cmh_tc_choose_seen tighten them - exactly - from the candidates of the rows already scanned); cmh_topk_finalize
// orders the candidates by key - (distance, index), all keys distinct - which IS the stable ranking, and flags any
// query whose candidate list is short or overflowed for the exact two-pass path.
#include <algorithm>
#include <cmath>

#include "tc_internal.cuh"

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
constexpr int TC_N = 256;       // database rows per tile (shared-memory stage): 2 rows per accumulator column
constexpr int TC_NM = 128;      // accumulator columns per tile = N of one tcgen05.mma
constexpr int TC_BUFS = 4;      // accumulator buffers (4 x 128 columns = the whole TMEM)
constexpr int TC_MMA_WARPS = 4;   // one MMA issuer (elected lane) per accumulator buffer
constexpr int TC_PROD_WARPS = 4;
constexpr int TC_EPI_WARPS = 16;  // 4 groups of 4 warps; group g drains buffer g
constexpr int TC_THREADS = (TC_MMA_WARPS + TC_PROD_WARPS + TC_EPI_WARPS) * 32;
// registers per thread: launched with 80 (768 threads); the issuer and producer warpgroups hand theirs to the 16 epilogue
// warps, which keep two slices of accumulators in flight:  32 * 128 + 32 * 128 + 104 * 512 = 80 * 768
constexpr int TC_REGS_MMA = 32, TC_REGS_PROD = 32, TC_REGS_EPI = 104;
static_assert(TC_REGS_MMA * TC_MMA_WARPS * 32 + TC_REGS_PROD * TC_PROD_WARPS * 32 + TC_REGS_EPI * TC_EPI_WARPS * 32 <=
              80 * TC_THREADS, "register budget");
// Experimental variant (template flag WK, CMH_TC_WORKERS=1): four more warps, one HIT WORKER per epilogue group, take
// the parked slices off the draining warps' queues.  896 threads are launched with 72 registers; issuers 32, producers
// 32, epilogue 96, workers 56:  32 * 256 + 96 * 512 + 56 * 128 = 72 * 896.
constexpr int TC_WORK_WARPS = 4;
constexpr int TC_THREADS_WK = TC_THREADS + TC_WORK_WARPS * 32;
constexpr int TC_REGS_EPI_WK = 96, TC_REGS_WORK = 56;
static_assert(TC_REGS_MMA * TC_MMA_WARPS * 32 + TC_REGS_PROD * TC_PROD_WARPS * 32 + TC_REGS_EPI_WK * TC_EPI_WARPS * 32 +
              TC_REGS_WORK * TC_WORK_WARPS * 32 <= 72 * TC_THREADS_WK, "register budget (worker variant)");
constexpr int TC_WK_WORDS = 3 * TC_BUFS * TC_M + 3 * TC_EPI_WARPS;   // per-query position / threshold / initial threshold
                                                                     // of every group; tail, head, done of every queue
constexpr int TC_RING = 8;        // packed database tiles in flight (bulk copies)
constexpr int TC_MAX_CHUNKS = 1024;  // candidate segments per query (cmh_topk_finalize walks them)
constexpr int TC_REFRESH = 16;    // tiles (per epilogue group) between threshold refreshes
constexpr int TC_PARK = 4;        // flagged lanes of a warp parked per trip (one __syncwarp)
constexpr int TC_BACKLOG = 8;     // parked slices an epilogue warp can hold (power of two)
constexpr int TC_PARK_WORDS = 36; // one parked slice: 32 registers, then owner lane, first row (relative to the chunk), pad
constexpr int TC_BIAS_SLOTS = 12; // K slots of the bias step that carry weight 127 (the 13th carries weight 1)

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
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
// register reallocation between the role warpgroups (all 4 warps of an aligned warpgroup execute it)
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
// one lane of a converged warp (warp-uniform code around it lets the operands live in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 64 consecutive columns, two per register: column 2i in the low half of v[i], column 2i+1 in the high half (the
// accumulators of 64-bit codes fit 16 bits, so their low halves are the exact values)
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

// 4 bits -> 4 bytes of +-S (bit set -> +S, clear -> -S), ALU only: the multiply spreads bit j to bit 8j (the four
// shifted copies of a nibble do not overlap, so the product is their OR), and (256 - S) - (256 - 2S) * {0,1} per byte
// never borrows.  S = 1: 0xFF / 0x01;  S = 32: 0xE0 / 0x20.
template <int S>
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
    const uint32_t x = (nib * 0x00204081u) & 0x01010101u;
    return (uint32_t)(0x100 - S) * 0x01010101u - x * (uint32_t)(0x100 - 2 * S);
}
template <int S>
__device__ __forceinline__ uint4 expand16(uint32_t bits16) {
    uint4 o;
    o.x = expand4<S>(bits16 & 15u);
    o.y = expand4<S>((bits16 >> 4) & 15u);
    o.z = expand4<S>((bits16 >> 8) & 15u);
    o.w = expand4<S>(bits16 >> 12);
    return o;
}

// per-query bookkeeping shared by every CTA working on the query (global memory, zeroed by cmh_tc_collect)
struct TcAux {
    uint32_t h[4];        // candidates of THIS launch so far at dist == thr0 - j (j = 0, 1, 2) and at dist <= thr0 - 3 (j = 3)
    uint32_t pad[4];      // (an overflowed segment shows in its count: cnt > seg_cap)
};

struct TcArgs {
    const uint64_t* q;      // [nq][words]
    const uint64_t* d;      // [nd][words]
    const int32_t* thr;     // [nq] initial threshold bucket thr0 (Hamming distance): rows with dist <= thr qualify
    uint64_t* cand;         // [nq][n_segs][seg_cap]  one private candidate segment per (query, chunk, draining thread)
    uint32_t* cnt;          // [n_segs][nq]           candidates found per segment (may exceed seg_cap)
    TcAux* aux;             // [nq]
    int64_t nq, nd, index_base;
    int chunk_rows;         // database rows per CTA (multiple of TC_N)
    int n_segs, seg_base, seg_cap, bits;   // segments per query in all, first segment of this launch
    int K;                  // > 0: tighten thresholds while scanning (once K rows at dist <= thr0 - j are known)
    long long* trace;       // CMH_TC_TRACE builds only
    int probe;              // measurement aid (cmh_tc_probe): 1 no tcgen05.mma, 2 no TMEM drain, 4 drain without scan,
                            // 8 hits decoded but not stored, 16 parked hits dropped, 32 flagged slices not parked,
                            // 64 parked slices also worked off in the rounds that parked them
};

// smem: [A: T tiles x {+-1, +-S, bias digits}][B: STAGES tiles][bias weights][packed ring][barriers][tmem slot][park]
// KBITS: the code length the operands are expanded to - 32 (codes of up to 32 bits: the low half of the packed word, one
// K-step per field), 64 or 128
template <int WORDS, int T, int TC_STAGES, bool WK = false, int KBITS = WORDS * 64>
__global__ void __launch_bounds__(WK ? TC_THREADS_WK : TC_THREADS, 1) tc_collect_kernel(const TcArgs a) {
    constexpr int NTHREADS = WK ? TC_THREADS_WK : TC_THREADS;
    constexpr int KBYTES = KBITS;                // int8 elements (= bytes) per row
    static_assert(KBITS == WORDS * 64 || (WORDS == 1 && KBITS == 32), "operand width");
    constexpr int KSTEPS = KBYTES / 32;          // tcgen05.mma kind::i8 has K = 32
    constexpr int CHUNKS = KBYTES / 16;          // 16-byte K chunks per row
    constexpr int FIELD = WORDS == 1 ? 8 : 10;   // bits per packed dot-product field
    constexpr int SCALE = 1 << (FIELD / 2);      // SCALE * SCALE = 1 << FIELD
    constexpr bool PACKED = WORDS == 1;          // accumulators fit 16 bits: two columns per register
    constexpr uint32_t FLAG_LO = 1u << (FIELD - 1), FLAG_HI = 1u << (2 * FIELD - 1);
    constexpr uint32_t FLAGS = PACKED ? (FLAG_LO | FLAG_HI) * 0x10001u : (FLAG_LO | FLAG_HI);
    constexpr uint32_t A_TILE = TC_M * KBYTES;   // one scale of one query tile
    constexpr uint32_t A_BIAS = TC_M * 32;       // the bias K-step of one query tile
    constexpr uint32_t A_ALL = 2 * A_TILE + A_BIAS;
    constexpr uint32_t B_TILE = TC_N * KBYTES;
    constexpr uint32_t B_BIAS = TC_NM * 32;
    static_assert(TC_BUFS % T == 0, "a group of epilogue warps serves one query tile");
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                    // [T][A_ALL]
    unsigned char* sB = smem + T * A_ALL;
    unsigned char* sW = sB + TC_STAGES * B_TILE; // bias weights: 128 identical rows of 32 bytes
    uint64_t* ring = reinterpret_cast<uint64_t*>(sW + B_BIAS);   // [TC_RING][256 rows][WORDS] packed words
    uint64_t* bars = ring + TC_RING * TC_N * WORDS;
    uint64_t* b_full = bars;                     // [STAGES] producers -> MMA
    uint64_t* b_empty = bars + TC_STAGES;        // [STAGES] MMA -> producers
    uint64_t* t_full = bars + 2 * TC_STAGES;               // [BUFS] MMA -> epilogue group
    uint64_t* t_empty = bars + 2 * TC_STAGES + TC_BUFS;    // [BUFS] epilogue group -> MMA
    uint64_t* r_full = bars + 2 * TC_STAGES + 2 * TC_BUFS;               // [RING] bulk copy -> producers
    uint64_t* r_empty = bars + 2 * TC_STAGES + 2 * TC_BUFS + TC_RING;    // [RING] producers -> bulk copy issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_BUFS + 2 * TC_RING);
    uint32_t* scratch = tmem_slot + 4;           // [EPI_WARPS][TC_BACKLOG][TC_PARK_WORDS] parked slices of the hit path
    // worker variant: [BUFS][128] positions, [BUFS][128] thresholds, [BUFS][128] initial thresholds, then per queue
    // (= epilogue warp) tail (written by its producer), head (by the worker), done
    uint32_t* wk_pos = scratch + TC_EPI_WARPS * TC_BACKLOG * TC_PARK_WORDS;
    int* wk_thr = reinterpret_cast<int*>(wk_pos + TC_BUFS * TC_M);
    int* wk_thr0 = wk_thr + TC_BUFS * TC_M;
    volatile uint32_t* wk_tail = reinterpret_cast<volatile uint32_t*>(wk_thr0 + TC_BUFS * TC_M);
    volatile uint32_t* wk_head = wk_tail + TC_EPI_WARPS;
    volatile uint32_t* wk_done = wk_head + TC_EPI_WARPS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * (T * TC_M);
    const int64_t c_begin = (int64_t)blockIdx.y * a.chunk_rows;
    const int64_t c_end = min(a.nd, c_begin + a.chunk_rows);
    const int n_tiles = (int)((c_end - c_begin + TC_N - 1) / TC_N);
    const int n_iters = n_tiles * T;             // iteration it = (tile i, query tile t) -> buffer / issuer / group it % 4
    // the flag arithmetic needs 1 <= T_q; a larger threshold is clamped (the query then comes out short and takes the
    // exact path)
    const int thr_max = (a.bits - 1) >> 1;

    // ---- prologue -----------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&b_full[s], TC_PROD_WARPS * 32);
            mbar_init(&b_empty[s], T);           // one commit per query tile
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
        if (WK)
            for (int i = 0; i < 3 * TC_EPI_WARPS; ++i) wk_tail[i] = 0u;
    }
    // A operand: T x 128 query rows at both scales, expanded by everyone (rows beyond nq are all -1 / -S; their
    // threshold never fires)
    for (int i = tid; i < T * TC_M * CHUNKS; i += NTHREADS) {
        const int c = i % CHUNKS;                // 16-byte chunk of the row
        const int r = (i / CHUNKS) % TC_M;
        const int t = i / (CHUNKS * TC_M);
        const int64_t q = q0 + t * TC_M + r;
        const uint64_t word = q < a.nq ? a.q[q * WORDS + (c >> 2)] : 0ull;
        const uint32_t b16 = (uint32_t)(word >> (16 * (c & 3))) & 0xffffu;
        unsigned char* dst = sA + t * A_ALL + c * (TC_M * 16) + r * 16;
        *reinterpret_cast<uint4*>(dst) = expand16<1>(b16);
        *reinterpret_cast<uint4*>(dst + A_TILE) = expand16<SCALE>(b16);
    }
    // bias K-step: digits of  bias_q = B - (B + 1) * T_q  against the weights (127 x TC_BIAS_SLOTS, 1, 0 ...)
    for (int i = tid; i < T * TC_M; i += NTHREADS) {
        const int r = i % TC_M, t = i / TC_M;
        const int64_t q = q0 + i;
        // padding queries (all -1) get the tightest valid threshold; whatever they flag is dropped by the hit path
        const int Tq = q < a.nq ? a.bits - 2 * max(0, min(a.thr[q], thr_max)) : a.bits;
        int rest = (1 << FIELD) - ((1 << FIELD) + 1) * Tq;                             // < 0
        int c1 = rest / 127;
        const int c0 = rest - c1 * 127;
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < TC_BIAS_SLOTS; ++k) {
            const int dgt = max(c1, -127);       // c1 <= 0
            c1 -= dgt;
            w[k >> 2] |= (uint32_t)(dgt & 0xff) << (8 * (k & 3));
        }
        w[TC_BIAS_SLOTS >> 2] |= (uint32_t)(c0 & 0xff) << (8 * (TC_BIAS_SLOTS & 3));
        unsigned char* dst = sA + t * A_ALL + 2 * A_TILE + r * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + TC_M * 16) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    for (int i = tid; i < TC_NM * 2; i += NTHREADS) {          // weights: row i % 128, 16-byte chunk i / 128
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < TC_BIAS_SLOTS; ++k) w[k >> 2] |= 127u << (8 * (k & 3));
        w[TC_BIAS_SLOTS >> 2] |= 1u << (8 * (TC_BIAS_SLOTS & 3));
        const int c = i / TC_NM;
        *reinterpret_cast<uint4*>(sW + c * (TC_NM * 16) + (i % TC_NM) * 16) =
            make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_MMA_WARPS) {
        // ================= MMA issuers =================
        // Issuer w owns accumulator buffer w: iterations w, w + 4, ...  One thread doing wait -> tcgen05.mma -> commit
        // for every tile spends ~300 cycles per tile on instruction latency alone (measured), more than the tensor
        // pipe needs for it; four issuers overlap that latency.
        reg_dec<TC_REGS_MMA>();
        {
            // The whole warp runs the loop and one elected lane issues: with warp-uniform control flow and a provably
            // uniform warp index the descriptors live in uniform registers and tcgen05.mma takes them as they are
            // (under `if (lane == 0)` every operand went through an ELECT / R2UR loop: ~90 instructions per tile on the
            // pipeline's critical path).
            const int w = __shfl_sync(0xffffffffu, warp, 0);
            constexpr uint32_t idesc = umma_idesc_i8(TC_M, TC_NM);
            const uint32_t b_addr = smem_u32(sB);
            const bool no_mma = (a.probe & 1) != 0;
            const uint32_t d_tmem = tmem_base + w * TC_NM;
            const uint32_t a_addr = smem_u32(sA) + (w % T) * A_ALL;              // this issuer's query tile
            const uint64_t bias_a = umma_desc(a_addr + 2 * A_TILE, TC_M * 16, 128);
            const uint64_t bias_b = umma_desc(smem_u32(sW), TC_NM * 16, 128);
            int round = 0;
            int s = (w / T) % TC_STAGES;                     // stage of this issuer's current iteration
            uint32_t ph = ((w / T) / TC_STAGES) & 1;
#pragma unroll 1
            for (int it = w; it < n_iters; it += TC_MMA_WARPS, ++round) {
                TC_TRACE(0, it, 3);
                mbar_wait(&b_full[s], ph);
                TC_TRACE(0, it, 0);
                mbar_wait(&t_empty[w], (round & 1) ^ 1);
                tc_fence_after();
                TC_TRACE(0, it, 1);
                if (elect_one()) {
                    if (!no_mma) {
                        umma_i8(d_tmem, bias_a, bias_b, idesc, 0u);  // acc = bias_q
#pragma unroll
                        for (int f = 0; f < 2; ++f)          // field 0: rows 0..127 (+-1); field 1: rows 128..255 (+-S)
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_i8(d_tmem, umma_desc(a_addr + f * A_TILE + k * 2 * (TC_M * 16), TC_M * 16, 128),
                                        umma_desc(b_addr + s * B_TILE + k * 2 * (TC_N * 16) + f * (TC_NM * 16), TC_N * 16, 128),
                                        idesc, 1u);
                    }
                    umma_commit(&t_full[w]);
                    umma_commit(&b_empty[s]);
                }
                __syncwarp();
                TC_TRACE(0, it, 2);
                s += TC_MMA_WARPS / T;                       // next iteration of this issuer: 4 / T tiles further
                if (s >= TC_STAGES) { s -= TC_STAGES; ph ^= 1; }
            }
        }
    } else if (warp < TC_MMA_WARPS + TC_PROD_WARPS) {
        // ================= producers: packed bits -> int8 core matrices =================
        // The packed words of a tile (256 rows, 2-4 KB) are staged in a TC_RING-deep shared-memory ring by the bulk-copy
        // engine (cp.async.bulk, issued by one thread, completion on an mbarrier), so the producer threads have no
        // global loads of their own in flight when they reach fence.proxy.async (a MEMBAR).  Tiles the engine cannot
        // take (a ragged last tile, a shard view that is not 16-byte aligned) are read with plain loads.
        reg_dec<TC_REGS_PROD>();
        const int pt = tid - TC_MMA_WARPS * 32;  // 0..127: rows pt (scale 1) and 128 + pt (scale S) of every tile
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
            // (the slot is handed back BELOW, after the words have been used: an arrive right behind the loads can be
            // performed while they are still queued - measured with dense hits: the refill for tile i + 8 then lands first
            // and rows of this warp come out as the rows eight tiles further on)
            if (pt == 0) TC_TRACE(1, i, 2);
            mbar_wait(&b_empty[s], ((i / TC_STAGES) & 1) ^ 1);
            if (pt == 0) TC_TRACE(1, i, 3);
            unsigned char* dst = sB + s * B_TILE + pt * 16;
            // rows past the end of the chunk are expanded like any other (as all -1): the hit path checks the row
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const uint32_t lo16 = (uint32_t)(cur[0][c >> 2] >> (16 * (c & 3))) & 0xffffu;
                const uint32_t hi16 = (uint32_t)(cur[1][c >> 2] >> (16 * (c & 3))) & 0xffffu;
                *reinterpret_cast<uint4*>(dst + c * (TC_N * 16)) = expand16<1>(lo16);
                *reinterpret_cast<uint4*>(dst + c * (TC_N * 16) + 128 * 16) = expand16<SCALE>(hi16);
            }
            if (pt == 0) TC_TRACE(1, i, 4);
            fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
            if (pt == 0) TC_TRACE(1, i, 5);
            mbar_arrive(&b_full[s]);
            // the ring slot goes back only now: the loaded words have been consumed by the stores above and the fence has
            // drained this thread's memory operations, so no read of the slot can still be in flight when it is refilled
            mbar_arrive(&r_empty[slot]);
            if (pt == 0) {
                mbar_wait(&r_empty[slot], (i / TC_RING) & 1);
                issue(i + TC_RING);              // refill the slot everyone has read
            }
            if (pt == 0) TC_TRACE(1, i, 6);
        }
    } else if (WK && warp >= TC_MMA_WARPS + TC_PROD_WARPS + TC_EPI_WARPS) {
        // ================= hit workers (experimental variant): one warp per epilogue group =================
        // The draining warps of group g only park flagged slices; this warp takes them off their four queues and owns
        // the group's per-query state (position in the candidate segment, current threshold) in shared memory - nobody
        // else touches it, so there are no atomics, and nothing the draining warps do depends on how long this takes.
        reg_dec<TC_REGS_WORK>();
        const int g = warp - (TC_MMA_WARPS + TC_PROD_WARPS + TC_EPI_WARPS);
        const int t = g % T;
        const int seg_id = a.seg_base + blockIdx.y * (TC_BUFS / T) + g / T;
        const bool hq_on = a.K > 0, drop = (a.probe & 16) != 0, store = !(a.probe & 8);
        const uint32_t row_base = (uint32_t)a.index_base + (uint32_t)c_begin;
        const uint32_t chunk_len = (uint32_t)(c_end - c_begin);
        uint32_t* pos_s = wk_pos + g * TC_M;
        int* thr_s = wk_thr + g * TC_M;
        int* thr0_s = wk_thr0 + g * TC_M;
        for (int i = lane; i < TC_M; i += 32) {
            const int64_t q = q0 + t * TC_M + i;
            const int t0 = q < a.nq ? min(a.thr[q], thr_max) : -1;
            pos_s[i] = 0u;
            thr_s[i] = t0;
            thr0_s[i] = t0;
        }
        __syncwarp();
        uint32_t heads[4] = {0u, 0u, 0u, 0u};             // warp-uniform
        for (;;) {
            bool progressed = false;
            int n_done = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                 // queue k = the group's warp that drains TMEM lanes 32k..32k+31
                const int ew = g * 4 + k;
                uint32_t dn = 0u, tail = 0u;
                if (lane == 0) {
                    dn = wk_done[ew];                     // done first: a tail read after it is final
                    __threadfence_block();
                    tail = wk_tail[ew];
                }
                dn = __shfl_sync(0xffffffffu, dn, 0);
                tail = __shfl_sync(0xffffffffu, tail, 0);
                __threadfence_block();                    // the entries below the tail are visible
                while (heads[k] != tail) {
                    const uint32_t* mine = scratch + (ew * TC_BACKLOG + (heads[k] & (TC_BACKLOG - 1))) * TC_PARK_WORDS;
                    const uint32_t xr = mine[lane];       // lane r looks at register r of the slice
                    const uint2 hdr = *reinterpret_cast<const uint2*>(mine + 32);   // owner lane, first row (rel.)
                    const int qrow_o = k * 32 + (int)hdr.x;
                    const int thr_o = thr_s[qrow_o], thr0_o = thr0_s[qrow_o];
                    const uint32_t pos_in = pos_s[qrow_o];
                    uint32_t pos_o = pos_in;
                    const int64_t q_o = q0 + t * TC_M + qrow_o;
                    const int dot_thr0_o = a.bits - 2 * max(0, thr0_o);
                    uint32_t fl = (q_o < a.nq && !drop) ? (~xr & FLAGS) : 0u;
                    uint64_t* const seg_o = a.cand + ((uint64_t)(q_o < a.nq ? q_o : 0) * a.n_segs + seg_id) * (uint64_t)a.seg_cap;
                    uint32_t* const hq_o = a.aux[q_o < a.nq ? q_o : 0].h;
                    do {                                  // one trip unless a register holds two flagged rows
                        bool pass = false;
                        int dist = 0;
                        uint32_t rel = 0;
                        if (fl) {
                            const int bit = 31 - __clz(fl);
                            fl &= ~(1u << bit);
                            const int col = PACKED ? 2 * lane + (bit >> 4) : lane;
                            const int f = PACKED ? ((bit >> 3) & 1) : (bit >= FIELD ? 1 : 0);
                            const int val = PACKED ? ((bit >> 4) ? (int)xr >> 16 : (int)(xr << 16) >> 16) : (int)xr;
                            const int e1 = (int)((uint32_t)val << (32 - FIELD)) >> (32 - FIELD);
                            const int dot = f ? ((val - e1) >> FIELD) + dot_thr0_o - 1 : e1 + dot_thr0_o;
                            dist = (a.bits - dot) >> 1;
                            rel = hdr.y + (uint32_t)(col + f * TC_NM);
                            pass = dist <= thr_o && rel < chunk_len;
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, pass);
                        if (pass && store) {
                            const uint32_t p = pos_o + __popc(m & lanemask_lt());
                            if (p < (uint32_t)a.seg_cap) seg_o[p] = ((uint64_t)(uint32_t)(2 * dist) << 32) | (row_base + rel);
                            if (hq_on) atomicAdd(hq_o + min(thr0_o - dist, 3), 1u);
                        }
                        pos_o += __popc(m);
                    } while (__any_sync(0xffffffffu, fl != 0u));
                    if (lane == 0) {
                        pos_s[qrow_o] = pos_o;
                        if (hq_on && ((pos_in ^ pos_o) >> 4)) {
                            // tighten (every 16 candidates of a query): once K rows at dist <= thr0 - j are known,
                            // nothing beyond that bucket can be in the top K
                            const uint4 c4 = __ldcg(reinterpret_cast<const uint4*>(hq_o));
                            const uint32_t K = (uint32_t)a.K;
                            const uint32_t c3 = c4.w, c2 = c3 + c4.z, c1 = c2 + c4.y;
                            thr_s[qrow_o] = thr0_o - (c3 >= K ? 3 : (c2 >= K ? 2 : (c1 >= K ? 1 : 0)));
                        }
                    }
                    __syncwarp();                         // everyone has read the entry; the state is written
                    ++heads[k];
                    if (lane == 0) {
                        __threadfence_block();
                        wk_head[ew] = heads[k];           // the producer may reuse the entry
                    }
                    progressed = true;
                }
                n_done += dn ? 1 : 0;
            }
            if (n_done == 4) break;                       // all four producers had finished before their tails were read
            if (!progressed) __nanosleep(200);
        }
        for (int i = lane; i < TC_M; i += 32) {
            const int64_t q = q0 + t * TC_M + i;
            if (q < a.nq) {
                a.cnt[(int64_t)seg_id * a.nq + q] = pos_s[i];
            }
        }
    } else {
        // ================= epilogue: flag filter on the packed dot products =================
        reg_inc<(WK ? TC_REGS_EPI_WK : TC_REGS_EPI)>();
        const int ew = warp - (TC_MMA_WARPS + TC_PROD_WARPS);   // 0 .. 15
        const int grp = ew >> 2;                          // accumulator buffer this warp drains
        const int quarter = warp & 3;                     // TMEM lanes this warp may touch: 32 * (warp % 4)
        const int qrow = quarter * 32 + lane;             // query row inside the 128-row tile
        const int t = grp % T;                            // ... of query tile t, for the whole CTA
        const bool no_drain = (a.probe & 2) != 0, no_scan = (a.probe & 4) != 0;
        const int64_t q = q0 + t * TC_M + qrow;
        const bool live = q < a.nq;
        // this thread's query: rows with dist <= thr qualify; thr0 is what the bias K-step (the flags) was built from
        const int thr0 = live ? min(a.thr[q], thr_max) : -1;
        int thr = thr0;
        const int dot_thr0 = a.bits - 2 * max(0, thr0);  // T0: the dot product behind the bias K-step of this query
        // with T < 4 a query is drained by 4 / T threads of this CTA (one per group): each owns a segment of its own
        const int seg_id = a.seg_base + blockIdx.y * (TC_BUFS / T) + grp / T;
        uint32_t* hq = (a.K > 0 && live) ? a.aux[q].h : nullptr;
        uint32_t pos = 0;
        // what the hit path needs of ANOTHER lane's query is derived from this lane's by pointer arithmetic (adjacent lanes
        // are adjacent queries): nothing is recomputed from the launch parameters while a parked slice is worked off
        uint64_t* const seg_mine = a.cand + ((uint64_t)q * a.n_segs + seg_id) * (uint64_t)a.seg_cap;   // not dereferenced unless live
        uint32_t* const hq_mine = a.aux[q].h;
        const int seg_stride = a.n_segs * a.seg_cap;     // keys between the segments of adjacent queries (< 2^31, host-checked)
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        const uint32_t row_base = (uint32_t)a.index_base + (uint32_t)c_begin;   // global index of the chunk's first row
        const uint32_t chunk_len = (uint32_t)(c_end - c_begin);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + grp * TC_NM;
        // The hit path, taken out of the pipeline's round trip.  A slice that holds a flagged row is PARKED: each flagged
        // lane copies its 32 registers into an entry of the warp's shared-memory queue (TC_BACKLOG entries) and the warp
        // goes on draining.  The queue is worked off while the warp would otherwise wait for its next accumulator tile,
        // one entry at a time and with a look at the tile's barrier in between: a buffer that is handed back late stalls
        // its MMA issuer, and the four buffers of TMEM are all the slack the pipeline has.  An entry is worked off by the
        // WHOLE warp - lane r decodes register r of the parked slice, the owner's threshold and position travel by
        // shuffle, the candidate segment is private to (query, chunk, group) so the position is a register counter, the
        // key goes straight to global memory and the tightening statistics are a fire-and-forget RED.  (A single lane
        // walking its own slice took ~4x as long: a dependent chain of ~100 instructions at one warp's issue latency.)
        uint32_t* park = scratch + ew * (TC_BACKLOG * TC_PARK_WORDS);
        int n_parked = 0, head = 0;                      // warp-uniform
        bool parked_now = false;                         // ... this round's scan parked a slice
        const bool hq_on = a.K > 0, drop = (a.probe & 16) != 0, store = !(a.probe & 8);
        auto work_off = [&]() {                          // the oldest parked slice, by the whole warp
            const uint32_t* mine = park + head * TC_PARK_WORDS;
            const uint32_t xr = mine[lane];              // lane r looks at register r of the slice
            const uint2 hdr = *reinterpret_cast<const uint2*>(mine + 32);    // owner lane, first row (relative to the chunk)
            const int owner = (int)hdr.x;
            const int thr_o = __shfl_sync(0xffffffffu, thr, owner);
            const int thr0_o = __shfl_sync(0xffffffffu, thr0, owner);
            uint32_t pos_o = __shfl_sync(0xffffffffu, pos, owner);
            const int dot_thr0_o = a.bits - 2 * max(0, thr0_o);
            uint32_t fl = (((live_mask >> owner) & 1u) && !drop) ? (~xr & FLAGS) : 0u;
            uint64_t* const seg_o = seg_mine + (int64_t)(owner - lane) * seg_stride;
            do {                                         // one trip unless a register holds two flagged rows
                bool pass = false;
                int dist = 0;
                uint32_t rel = 0;
                if (fl) {
                    const int bit = 31 - __clz(fl);
                    fl &= ~(1u << bit);
                    // packed: bit 7 / 15 = column 2r, field 0 / 1; bit 23 / 31 = column 2r + 1, field 0 / 1
                    const int col = PACKED ? 2 * lane + (bit >> 4) : lane;
                    const int f = PACKED ? ((bit >> 3) & 1) : (bit >= FIELD ? 1 : 0);
                    // acc = e1 + B * e2 with e1 = dot1 - T0, e2 = dot2 - T0 + 1 (no wrap): decode the flagged field
                    const int val = PACKED ? ((bit >> 4) ? (int)xr >> 16 : (int)(xr << 16) >> 16) : (int)xr;
                    const int e1 = (int)((uint32_t)val << (32 - FIELD)) >> (32 - FIELD);
                    const int dot = f ? ((val - e1) >> FIELD) + dot_thr0_o - 1 : e1 + dot_thr0_o;
                    dist = (a.bits - dot) >> 1;
                    rel = hdr.y + (uint32_t)(col + f * TC_NM);
                    pass = dist <= thr_o && rel < chunk_len;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, pass);
                if (pass && store) {
                    const uint32_t p = pos_o + __popc(m & lanemask_lt());
                    if (p < (uint32_t)a.seg_cap) seg_o[p] = ((uint64_t)(uint32_t)(2 * dist) << 32) | (row_base + rel);
                    if (hq_on) atomicAdd(hq_mine + (owner - lane) * 8 + min(thr0_o - dist, 3), 1u);
                }
                pos_o += __popc(m);
            } while (__any_sync(0xffffffffu, fl != 0u));
            if (lane == owner) pos = pos_o;
            __syncwarp();                                // the entry may be overwritten now
            head = (head + 1) & (TC_BACKLOG - 1);
            --n_parked;
        };
        auto park_slice = [&](const uint32_t (&v)[32], bool flagged, int64_t row0) {
            uint32_t pend = __ballot_sync(0xffffffffu, flagged);
            if (a.probe & 32) pend = 0;
            while (pend) {                               // warp-uniform
                if (n_parked == TC_BACKLOG) work_off();  // backlog full: this one is paid for on the spot
                const int room = TC_BACKLOG - n_parked;
                uint32_t rest = pend;                    // the lowest flagged lanes there is room for, TC_PARK at most
#pragma unroll
                for (int k = 0; k < TC_PARK; ++k)
                    if (k < room) rest &= rest - 1;
                const uint32_t take = pend & ~rest;
                if ((take >> lane) & 1u) {               // 8 x 16-byte stores: parking must stay cheap, it is on the clock
                    uint32_t* mine = park + ((head + n_parked + __popc(take & lanemask_lt())) & (TC_BACKLOG - 1)) * TC_PARK_WORDS;
#pragma unroll
                    for (int r = 0; r < 32; r += 4)
                        *reinterpret_cast<uint4*>(mine + r) = make_uint4(v[r], v[r + 1], v[r + 2], v[r + 3]);
                    *reinterpret_cast<uint2*>(mine + 32) = make_uint2((uint32_t)lane, (uint32_t)(row0 - c_begin));
                }
                __syncwarp();
                n_parked += __popc(take);
                pend = rest;
            }
        };
        // worker variant: the queue is drained by the group's hit worker; this warp only appends (and waits when it is full)
        uint32_t tail_l = 0u, head_seen = 0u;            // warp-uniform
        auto park_slice_wk = [&](const uint32_t (&v)[32], bool flagged, int64_t row0) {
            uint32_t pend = __ballot_sync(0xffffffffu, flagged);
            if (a.probe & 32) pend = 0;
            while (pend) {                               // warp-uniform
                while (tail_l - head_seen >= (uint32_t)TC_BACKLOG) {
                    uint32_t h = 0u;
                    if (lane == 0) h = wk_head[ew];
                    head_seen = __shfl_sync(0xffffffffu, h, 0);
                    if (tail_l - head_seen >= (uint32_t)TC_BACKLOG) __nanosleep(64);
                }
                const int room = TC_BACKLOG - (int)(tail_l - head_seen);
                uint32_t rest = pend;
#pragma unroll
                for (int k = 0; k < TC_PARK; ++k)
                    if (k < room) rest &= rest - 1;
                const uint32_t take = pend & ~rest;
                if ((take >> lane) & 1u) {
                    uint32_t* mine = park + ((tail_l + __popc(take & lanemask_lt())) & (TC_BACKLOG - 1)) * TC_PARK_WORDS;
#pragma unroll
                    for (int r = 0; r < 32; r += 4)
                        *reinterpret_cast<uint4*>(mine + r) = make_uint4(v[r], v[r + 1], v[r + 2], v[r + 3]);
                    *reinterpret_cast<uint2*>(mine + 32) = make_uint2((uint32_t)lane, (uint32_t)(row0 - c_begin));
                }
                __syncwarp();
                tail_l += (uint32_t)__popc(take);
                if (lane == 0) {
                    __threadfence_block();               // the entries before the tail
                    wk_tail[ew] = tail_l;
                }
                pend = rest;
            }
        };
        // one slice (32 registers: 64 or 128 rows): AND-reduce, one mask test, one vote; the hit path is rare
        // two slices (32 registers each: 64 or 128 rows): AND-reduce, one mask test per slice, ONE vote for both; the
        // hit path is rare
        auto scan2 = [&](const uint32_t (&v0)[32], int64_t r0, const uint32_t (&v1)[32], int64_t r1) {
            if (no_scan) {
                uint32_t o = 0;
#pragma unroll
                for (int r = 0; r < 32; ++r) o |= v0[r] | v1[r];
                if (o == 0xdeadbeefu) a.cnt[0] = o;      // keep the loads alive
                return;
            }
            uint32_t ab0[4], ab1[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                ab0[b] = (v0[8 * b] & v0[8 * b + 1] & v0[8 * b + 2]) & (v0[8 * b + 3] & v0[8 * b + 4] & v0[8 * b + 5]) &
                         (v0[8 * b + 6] & v0[8 * b + 7]);
                ab1[b] = (v1[8 * b] & v1[8 * b + 1] & v1[8 * b + 2]) & (v1[8 * b + 3] & v1[8 * b + 4] & v1[8 * b + 5]) &
                         (v1[8 * b + 6] & v1[8 * b + 7]);
            }
            const bool f0 = ((ab0[0] & ab0[1] & ab0[2] & ab0[3]) & FLAGS) != FLAGS;
            const bool f1 = ((ab1[0] & ab1[1] & ab1[2] & ab1[3]) & FLAGS) != FLAGS;
            if (__any_sync(0xffffffffu, f0 || f1)) {
                if (WK) {
                    park_slice_wk(v0, f0, r0);
                    park_slice_wk(v1, f1, r1);
                } else {
                    park_slice(v0, f0, r0);
                    park_slice(v1, f1, r1);
                    parked_now = true;
                }
            }
        };
        auto release = [&]() {                   // the values are in registers: the buffer goes back to its issuer
            tc_fence_before();
            mbar_arrive(&t_empty[grp]);
        };
        int round = 0;
#pragma unroll 1
        for (int it = grp; it < n_iters; it += TC_BUFS, ++round) {
            const int i = it / T;
            const int64_t row0 = c_begin + (int64_t)i * TC_N;
            if (qrow == 0) TC_TRACE(2 + grp, round, 0);
            // The backlog of parked hits is worked off in the time this warp would spend waiting for its tile - but not
            // in a round that has just parked: scan + parking + one work-off is about as long as the wait, and a warp
            // that is late for its tile holds the whole group's buffer back (the tensor pipe has no slack to catch up).
            const bool busy_round = parked_now && !(a.probe & 64);
            parked_now = false;
            if (!WK && !busy_round)
                while (n_parked > 0 && !__any_sync(0xffffffffu, mbar_test(&t_full[grp], round & 1))) work_off();
            mbar_wait(&t_full[grp], round & 1);
            tc_fence_after();
            if (qrow == 0) TC_TRACE(2 + grp, round, 1);
            if (no_drain) {
                release();
                continue;
            }
            // two slices (register buffers) per trip to TMEM
            uint32_t va[32], vb[32];
            if (PACKED) {
                tmem_ld64p(taddr, va);
                tmem_ld64p(taddr + 64, vb);
                tmem_ld_wait();
                if (qrow == 0) TC_TRACE(2 + grp, round, 2);
                release();                       // before anything is looked at: a hit never holds the buffer
                scan2(va, row0, vb, row0 + 64);
            } else {
                tmem_ld32(taddr, va);
                tmem_ld32(taddr + 32, vb);
                tmem_ld_wait();
                scan2(va, row0, vb, row0 + 32);
                tmem_ld32(taddr + 64, va);
                tmem_ld32(taddr + 96, vb);
                tmem_ld_wait();
                if (qrow == 0) TC_TRACE(2 + grp, round, 2);
                release();
                scan2(va, row0 + 64, vb, row0 + 96);
            }
            if (!WK && hq != nullptr && (round & (TC_REFRESH - 1)) == TC_REFRESH - 1) {
                // tighten: once K rows at dist <= thr0 - j are known, nothing beyond that bucket can be in the top K.
                // Here, after the buffer went back, the L2 round trip of the counters costs the pipeline nothing.
                const uint4 c4 = __ldcg(reinterpret_cast<const uint4*>(hq));
                const uint32_t K = (uint32_t)a.K;
                const uint32_t c3 = c4.w, c2 = c3 + c4.z, c1 = c2 + c4.y;
                const int j = c3 >= K ? 3 : (c2 >= K ? 2 : (c1 >= K ? 1 : 0));
                thr = thr0 - j;
            }
            if (qrow == 0) TC_TRACE(2 + grp, round, 4);
        }
        if (WK) {
            __syncwarp();
            if (lane == 0) {                             // this queue gets no more entries (the tail is final)
                __threadfence_block();
                wk_done[ew] = 1u;
            }
        } else {
            while (n_parked > 0) work_off();
            if (live) {
                a.cnt[(int64_t)seg_id * a.nq + q] = pos;
            }
        }
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

constexpr int FIN_MAX = 4096;
constexpr uint64_t TC_KEY_FAIL = ~0ull - 1;   // first key of a failed per-shard list (no real key: dist field 2^31 - 1)
constexpr int FIN_BINS = 129;   // bits <= 128 on the tensor path
constexpr int FIN_THREADS = 512;

// ---- thresholds from the candidates of a pilot launch ------------------------------------------------------------------
// The pilot launch scanned n_seen of the nd rows with thresholds thr_in and kept EVERY row at or below them (no
// tightening), so per query the candidates' histogram is an exact sample of the distance distribution below thr_in.
// Step 1 (tc_cand_hist_kernel): that histogram, uint32 [nq][nb] (bucket = Hamming distance), and whether a segment
// overflowed.  Sharded databases all-reduce the histograms here.  Step 2 (tc_choose_kernel): thr_out = the smallest
// bucket whose cumulative count reaches `need`, capped by thr_in; unchanged when a segment overflowed.  As with
// cmh_topk_threshold any outcome is safe: a threshold that turns out too low is caught after the merge.
// A query's candidates are spread over many short segments (one per chunk, launch and draining group: a few hundred,
// ~10 entries each).  Walking them one segment after the other is a chain of dependent L2/HBM round trips (count, then
// entries); instead a block of segments is INDEXED: every thread loads the counts of SEG_IT segments at once, a block
// scan turns them into offsets in shared memory, and the candidates are then addressed by their flat position (a binary
// search in the offsets) - all loads of a pass are in flight together.
constexpr int SEG_IT = 4;         // segments per thread and block of segments

// offsets of the segments [c0, c0 + THREADS * SEG_IT) (those beyond c1 are empty) into s_off[THREADS * SEG_IT + 1];
// returns the number of (clamped) entries.  raw / over accumulate per thread.
template <int THREADS>
__device__ __forceinline__ uint32_t seg_block_offsets(const uint32_t* __restrict__ cnt, int64_t nq, int64_t q, int c0, int c1,
                                                      uint32_t seg_cap, uint32_t* s_off, uint32_t* s_warp, uint32_t& raw,
                                                      uint32_t& over) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[SEG_IT], sum = 0;
#pragma unroll
    for (int k = 0; k < SEG_IT; ++k) {
        const int c = c0 + tid * SEG_IT + k;
        uint32_t n = c < c1 ? cnt[(int64_t)c * nq + q] : 0u;
        raw += n;
        if (n > seg_cap) { over = 1u; n = seg_cap; }
        v[k] = n;
        sum += n;
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    __syncthreads();                             // the previous block's walk is over: s_off / s_warp may be rewritten
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t ex = inc - sum;
    for (int w = 0; w < warp; ++w) ex += s_warp[w];
#pragma unroll
    for (int k = 0; k < SEG_IT; ++k) {
        s_off[tid * SEG_IT + k] = ex;
        ex += v[k];
    }
    if (tid == THREADS - 1) s_off[THREADS * SEG_IT] = ex;
    __syncthreads();
    return s_off[THREADS * SEG_IT];
}

// the key at flat position i (< s_off[n_off]) of the indexed block: segment = the last one whose offset is <= i
__device__ __forceinline__ uint64_t seg_flat_load(const uint64_t* __restrict__ block_base, const uint32_t* s_off, int n_off,
                                                  uint32_t seg_cap, uint32_t i) {
    int lo = 0, hi = n_off;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid;
    }
    return block_base[(uint64_t)lo * seg_cap + (i - s_off[lo])];
}

constexpr int CH_THREADS = 256;
// the threshold rule shared by tc_choose_kernel and the fused form below: the smallest bucket b <= thr_in whose
// cumulative count reaches `need` gives min(thr_in, b + offset); thr_in when there is none
struct ChooseRule {
    double need;          // < 0: no fused choose
    int offset;
    const int32_t* thr_in;
    int32_t* thr_out;
};
__global__ void __launch_bounds__(CH_THREADS) tc_cand_hist_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                                                  int64_t nq, int seg_lo, int seg_hi, int seg_total, int seg_cap,
                                                                  int nb, uint32_t* __restrict__ hist_out,
                                                                  uint32_t* __restrict__ overflow, const ChooseRule rule) {
    constexpr int BLOCK = CH_THREADS * SEG_IT;
    __shared__ uint32_t hist[FIN_BINS];
    __shared__ uint32_t s_off[BLOCK + 1], s_warp[CH_THREADS / 32];
    __shared__ uint32_t s_over;
    const int64_t q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < FIN_BINS; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x == 0) s_over = 0u;
    const uint64_t* __restrict__ mine = cand + (uint64_t)q * seg_total * (uint64_t)seg_cap;
    uint32_t raw = 0, over = 0;
    for (int c0 = seg_lo; c0 < seg_hi; c0 += BLOCK) {
        const uint32_t total = seg_block_offsets<CH_THREADS>(cnt, nq, q, c0, seg_hi, (uint32_t)seg_cap, s_off, s_warp, raw, over);
        const uint64_t* base = mine + (uint64_t)c0 * seg_cap;
        for (uint32_t i0 = threadIdx.x - lane; i0 < total; i0 += CH_THREADS) {   // warp-uniform trip count
            const uint32_t i = i0 + lane;
            const uint32_t bkt = i < total ? min((uint32_t)(seg_flat_load(base, s_off, BLOCK, (uint32_t)seg_cap, i) >> 33),
                                                 (uint32_t)(FIN_BINS - 1)) : (uint32_t)FIN_BINS;
            const uint32_t same = __match_any_sync(0xffffffffu, bkt);
            if (bkt < (uint32_t)FIN_BINS && lane == __ffs(same) - 1) atomicAdd(&hist[bkt], (uint32_t)__popc(same));
        }
    }
    if (over) s_over = 1u;
    __syncthreads();
    if (hist_out != nullptr) {
        for (int i = threadIdx.x; i < nb; i += blockDim.x) hist_out[q * nb + i] = i < FIN_BINS ? hist[i] : 0u;
        if (threadIdx.x == 0) overflow[q] = s_over;
    }
    if (rule.need >= 0.0 && threadIdx.x == 0) {      // one GPU: nothing to exchange between the histogram and the rule
        const int t_in = rule.thr_in[q];
        int t = t_in;
        if (s_over == 0u) {
            double cum = 0.0;
            for (int b = 0; b <= min(t_in, min(nb, FIN_BINS) - 1); ++b) {
                cum += (double)hist[b];
                if (cum >= rule.need) { t = min(t_in, b + rule.offset); break; }
            }
        }
        rule.thr_out[q] = t;
    }
}

// sharded, contiguous ranges: every[r][q][b] all-gathered candidate histograms -> lower = sum over ranks <= rank
// (rows of lower index than what this shard still has to scan), seen = sum over all ranks
__global__ void __launch_bounds__(256) tc_sum_ranks_kernel(const uint32_t* __restrict__ every, int world, int rank, int64_t n,
                                                           uint32_t* __restrict__ lower, uint32_t* __restrict__ seen) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, all = 0;
    for (int r = 0; r < world; ++r) {
        const uint32_t v = every[(int64_t)r * n + i];
        all += v;
        if (r <= rank) lo += v;
    }
    lower[i] = lo;
    seen[i] = all;
}

__global__ void __launch_bounds__(256) tc_choose_kernel(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ overflow,
                                                        int64_t nq, int nb, double need, int offset,
                                                        const int32_t* __restrict__ thr_in, int32_t* __restrict__ thr_out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int t_in = thr_in[q];
    int t = t_in;
    if (overflow == nullptr || overflow[q] == 0u) {
        double cum = 0.0;
        for (int b = 0; b <= min(t_in, nb - 1); ++b) {
            cum += (double)hist[q * nb + b];
            if (cum >= need) { t = min(t_in, b + offset); break; }
        }
    }
    thr_out[q] = t;
}

// ---- verify: the K-th key of a (merged) result must come from complete buckets ------------------------------------------
__global__ void __launch_bounds__(256) topk_verify_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ thr_limit,
                                                          int64_t nq, int K, int64_t need, uint32_t* __restrict__ fail_flags,
                                                          uint32_t* __restrict__ fail_count) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    bool fail = fail_flags[q] != 0u;
    if (!fail && need > 0) {
        const uint64_t last = keys[q * K + (need - 1)];
        fail = last == ~0ull || (int)(uint32_t)(last >> 33) > thr_limit[q];
    }
    fail_flags[q] = fail ? 1u : 0u;
    if (fail) atomicAdd(fail_count, 1u);
}

// ---- finalize: exact K-th bucket from the candidates' own histogram, compact, sort, emit -------------------------------
// One CTA per query.  Every database row at or below the query's final threshold was collected, and that threshold
// is an upper bound of the K-th distance, so the smallest bucket T whose cumulative candidate count reaches K is the
// true K-th distance; only candidates with dist <= T (between K and a few K of them) are sorted - by key, i.e. by
// (distance, global index), which is the stable ranking.  A warp walks one (query, chunk) segment at a time.

// THREADS x MAXK: 512 x 4096 for a whole database; 256 x 2048 for the short per-shard lists of a sharded search (twice the
// resident CTAs, half the threads per barrier: the kernel is a chain of dependent small steps per query)
template <int THREADS, int MAXK>
__global__ void __launch_bounds__(THREADS) topk_finalize_kernel(const uint64_t* __restrict__ cand,
                                                                    const uint32_t* __restrict__ cnt,
                                                                    const int32_t* __restrict__ thr_limit, int64_t nq,
                                                                    int n_chunks, int seg_cap, int K, int64_t nd, int partial,
                                                                    int width, int order_slack, uint64_t* __restrict__ keys,
                                                                    uint32_t* __restrict__ fail_flags,
                                                                    uint32_t* __restrict__ fail_count) {
    constexpr int BLOCK = THREADS * SEG_IT;
    __shared__ uint64_t sk[MAXK];
    __shared__ uint32_t hist[FIN_BINS];
    __shared__ uint32_t s_off[BLOCK + 1], s_warp[THREADS / 32];
    __shared__ uint32_t s_wc[THREADS / 32][8], s_base[8];
    __shared__ int s_T, s_keep;
    __shared__ uint32_t s_total, s_over;
    const int64_t q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t need = nd < (int64_t)K ? nd : (int64_t)K;
    const uint64_t* __restrict__ mine = cand + (uint64_t)q * n_chunks * (uint64_t)seg_cap;
    for (int i = threadIdx.x; i < FIN_BINS; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x == 0) { s_total = 0u; s_over = 0u; }
    // pass 1: the candidates' histogram (they sit in 3-4 buckets: one add per bucket per warp)
    uint32_t raw = 0, over = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += BLOCK) {
        const uint32_t total = seg_block_offsets<THREADS>(cnt, nq, q, c0, n_chunks, (uint32_t)seg_cap, s_off, s_warp, raw, over);
        const uint64_t* base = mine + (uint64_t)c0 * seg_cap;
        for (uint32_t i0 = threadIdx.x - lane; i0 < total; i0 += THREADS) {  // warp-uniform trip count
            const uint32_t i = i0 + lane;
            const uint32_t bkt = i < total ? min((uint32_t)(seg_flat_load(base, s_off, BLOCK, (uint32_t)seg_cap, i) >> 33),
                                                 (uint32_t)(FIN_BINS - 1)) : (uint32_t)FIN_BINS;
            const uint32_t same = __match_any_sync(0xffffffffu, bkt);
            if (bkt < (uint32_t)FIN_BINS && lane == __ffs(same) - 1) atomicAdd(&hist[bkt], (uint32_t)__popc(same));
        }
    }
    raw = __reduce_add_sync(0xffffffffu, raw);
    if (lane == 0 && raw) atomicAdd(&s_total, raw);
    if (over) s_over = 1u;
    __syncthreads();
    // partial (one shard of several): emit what there is, up to `width`; the K-th key is judged after the merge
    const int64_t want = partial ? min(min(need, (int64_t)width), (int64_t)s_total) : need;
    bool fail = s_over != 0u || (int64_t)s_total < want;
    if (!fail) {
        if (threadIdx.x == 0) {
            int64_t cum = 0;
            int T = -1;
            for (int b = 0; b < FIN_BINS && cum < want; ++b) { cum += hist[b]; T = b; }
            s_T = T;
            // buckets above thr_limit may be incomplete (launches with different thresholds): the K-th distance must
            // not come from there
            // More candidates at or below the K-th bucket than the sort buffer holds (short codes: the K-th bucket of a
            // 16-bit code holds thousands of rows and the index decides).  The buckets below T always fit (fewer than
            // `want` rows); of bucket T only the first rows in STORED order are kept - exact when that order is the index
            // order up to a bounded displacement: order_slack = 256 says a query's candidates are appended tile by tile
            // (one segment per chunk: operand widths up to 64 bits), so the `want` lowest indices of the bucket are among
            // its first want + 255 stored entries; -1 = no such bound (two interleaved segments per chunk at 128 bits,
            // caller-built segments): the query is then flagged as before.
            const bool fits = cum <= MAXK || (order_slack >= 0 && want + order_slack <= (int64_t)MAXK);
            s_keep = (!fits || (thr_limit != nullptr && T > thr_limit[q])) ? -1 : (int)(cum < (int64_t)MAXK ? cum : (int64_t)MAXK);
        }
        __syncthreads();
        fail = s_keep < 0;
    }
    if (threadIdx.x == 0) {
        fail_flags[q] = fail ? 1u : 0u;
        if (fail) atomicAdd(fail_count, 1u);
    }
    if (fail) {
        // a failed shard list starts with the marker: the merge flags the query whatever the other shards hold
        for (int i = threadIdx.x; i < width; i += blockDim.x) keys[q * width + i] = (partial && i == 0) ? TC_KEY_FAIL : ~0ull;
        return;
    }
    // pass 2: place the candidates at or below the K-th bucket, bucket by bucket and - inside a bucket - in the order
    // they are stored.  That order is ascending in the row index (segments follow the chunks and launches, a segment
    // is appended to tile after tile) except inside one 256-row tile and between the two groups that share a query
    // tile of 128-bit codes, so the result is usually sorted already: it is checked, and only an unsorted one goes
    // through the bitonic network (which was 90 % of this kernel's instructions).
    // Classes: slot s = T - bucket for the 7 highest buckets, slot 7 = everything below (rarely populated).
    const int T = s_T, keep = s_keep;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) {
        const int b_lo = threadIdx.x == 7 ? 0 : T - (int)threadIdx.x;     // first bucket of the class
        uint32_t below = 0;
        for (int b = 0; b < b_lo && b < FIN_BINS; ++b) below += hist[b];
        s_base[threadIdx.x] = b_lo < 0 ? 0u : below;
    }
    for (int c0 = 0; c0 < n_chunks; c0 += BLOCK) {
        // (with a single block of segments - the usual case - the offsets of pass 1 are still in place)
        const uint32_t total = n_chunks <= BLOCK ? s_off[BLOCK]
                                                 : seg_block_offsets<THREADS>(cnt, nq, q, c0, n_chunks, (uint32_t)seg_cap, s_off, s_warp, raw, over);
        const uint64_t* base = mine + (uint64_t)c0 * seg_cap;
        for (uint32_t i0 = 0; i0 < total; i0 += THREADS) {          // block-uniform trip count
            const uint32_t i = i0 + threadIdx.x;
            const uint64_t key = i < total ? seg_flat_load(base, s_off, BLOCK, (uint32_t)seg_cap, i) : ~0ull;
            const int bkt = (int)(uint32_t)(key >> 33);
            const int cls = (i < total && bkt <= T) ? min(T - bkt, 7) : 8;
            if (threadIdx.x < (THREADS / 32) * 8) s_wc[threadIdx.x >> 3][threadIdx.x & 7] = 0u;
            __syncthreads();
            const uint32_t same = __match_any_sync(0xffffffffu, cls);
            if (cls < 8 && lane == __ffs(same) - 1) s_wc[warp][cls] = (uint32_t)__popc(same);
            __syncthreads();
            if (cls < 8) {
                uint32_t pos = s_base[cls] + (uint32_t)__popc(same & lanemask_lt());
                for (int w = 0; w < warp; ++w) pos += s_wc[w][cls];
                if (pos < (uint32_t)MAXK) sk[pos] = key;
            }
            __syncthreads();
            if (threadIdx.x < 8) {
                uint32_t tot = 0;
#pragma unroll
                for (int w = 0; w < THREADS / 32; ++w) tot += s_wc[w][threadIdx.x];
                s_base[threadIdx.x] += tot;
            }
            __syncthreads();                     // s_wc is zeroed again by the next trip
        }
    }
    __syncthreads();
    bool unsorted = false;
    for (int i = threadIdx.x; i + 1 < keep; i += blockDim.x) unsorted |= sk[i] > sk[i + 1];
    if (__syncthreads_or(unsorted)) {
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
    }
    for (int i = threadIdx.x; i < width; i += blockDim.x) keys[q * width + i] = i < keep ? sk[i] : ~0ull;
}

// ---- sharded: merge of the per-shard lists of one query slice + verification, one CTA per query -----------------------
// lists: [n_lists][nq][W] ascending keys (pads UINT64_MAX; a list that starts with TC_KEY_FAIL comes from a shard whose
// candidate segments overflowed).  Every key's rank among all lists by binary search (keys are unique across shards).
// A list that arrives FULL (W < K keys, none a pad) may have been cut: what it lost lies above its last key, so the
// result is exact unless that key ranks before the K-th of the merge.  The K-th key must also come from a complete
// bucket (<= thr_limit).  Lists live in shared memory when they fit, else they are searched where they are.
__global__ void __launch_bounds__(256) topk_merge_verify_kernel(const uint64_t* __restrict__ lists, int n_lists, int64_t nq /* rows per list block */,
                                                                int W, int K, int64_t need, const int32_t* __restrict__ thr_limit,
                                                                int in_smem, uint64_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ fail_flags) {
    extern __shared__ uint64_t sk[];  // [n_lists][W] for this query (in_smem)
    __shared__ uint32_t s_fail;
    __shared__ uint64_t s_kth;
    const int64_t q = blockIdx.x;
    if (threadIdx.x == 0) { s_fail = 0u; s_kth = ~0ull; }
    for (int i = threadIdx.x; i < K; i += blockDim.x) keys_out[q * K + i] = ~0ull;
    if (in_smem)
        for (int i = threadIdx.x; i < n_lists * W; i += blockDim.x) {
            const int g = i / W, j = i - g * W;
            sk[i] = lists[((int64_t)g * nq + q) * W + j];
        }
    __syncthreads();
    auto list_of = [&](int g) { return in_smem ? sk + g * W : lists + ((int64_t)g * nq + q) * W; };
    for (int i = threadIdx.x; i < n_lists * W; i += blockDim.x) {
        const int g = i / W, j = i - g * W;
        const uint64_t key = list_of(g)[j];
        if (key == TC_KEY_FAIL) { s_fail = 1u; continue; }
        if (key == ~0ull) continue;  // padding
        int rank = j;
        for (int o = 0; o < n_lists && rank < K; ++o) {
            if (o == g) continue;
            const uint64_t* lst = list_of(o);
            int lo = 0, hi = W;      // first position with lst[pos] >= key
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lst[mid] < key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < K) keys_out[q * K + rank] = key;
        if (need > 0 && rank == (int)need - 1) s_kth = key;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool fail = s_fail != 0u;
        const uint64_t kth = s_kth;
        if (need > 0 && (kth == ~0ull || (thr_limit != nullptr && (int)(uint32_t)(kth >> 33) > thr_limit[q]))) fail = true;
        if (W < K)
            for (int g = 0; g < n_lists && !fail; ++g) {
                const uint64_t last = list_of(g)[W - 1];
                if (last != ~0ull && last != TC_KEY_FAIL && last < kth) fail = true;   // the cut may have cost a key
            }
        fail_flags[q] = fail ? 1u : 0u;
    }
}

// fail_count = number of flagged queries (flags identical on every rank after the all-gather); failed queries' keys -> pads
__global__ void __launch_bounds__(256) tc_count_flags_kernel(const uint32_t* __restrict__ flags, int64_t n, uint32_t* __restrict__ count) {
    uint32_t c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += flags[i] != 0u ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

}  // namespace cmh

using namespace cmh;

// Per code length: query tiles per CTA and B-ring depth.  64-bit codes: 4 x 128 queries (80 KB of A: two scales + bias),
// 6 stages of 16 KB; 128-bit codes: 2 x 128 queries (64 KB), 3 stages of 32 KB.  Either way > 114 KB of shared memory,
// so exactly one CTA - which owns the whole TMEM - is resident per SM.
static int tc_T(int words) { return words == 1 ? 4 : 2; }
static int tc_stages(int words) { return words == 1 ? 6 : 3; }
// CMH_TC_WORKERS=0 / 1 (or cmh_tc_set_workers) forces the draining-warp / hit-worker variant for every launch
// (measurement aid; 64-bit codes); unset / -1: chosen per launch (tc_collect_launch)
static int g_tc_workers_mode = [] { const char* e = getenv("CMH_TC_WORKERS"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1; }();
static int tc_workers_mode() { return g_tc_workers_mode; }
extern "C" int cmh_tc_set_workers(int mode) {
    CMH_REQUIRE(mode >= -1 && mode <= 1, CMH_ERR_ARG, "cmh_tc_set_workers: mode=%d", mode);
    g_tc_workers_mode = mode;
    return CMH_OK;
}
static size_t tc_smem_bytes(int words, int kbytes) {
    const int st = tc_stages(words), T = tc_T(words);
    return (size_t)T * (2 * TC_M * kbytes + TC_M * 32) + (size_t)st * TC_N * kbytes + TC_NM * 32 +
           (size_t)TC_RING * TC_N * words * 8 + (2 * st + 2 * TC_BUFS + 2 * TC_RING) * 8 + 16 +
           (size_t)TC_EPI_WARPS * TC_BACKLOG * TC_PARK_WORDS * 4;
}

extern "C" int cmh_tc_supported(int bits, int ternary) {
    return (!ternary && bits >= 1 && bits <= 128) ? 1 : 0;
}

// launch geometry: query groups of T x 128 rows, database chunks of whole tiles, at most 4 CTAs per SM over the launch
static void tc_geometry(int64_t nq, int64_t nd, int words, int64_t* n_qgroups, int64_t* n_chunks, int64_t* chunk_rows) {
    *n_qgroups = std::max<int64_t>(1, ceil_div(nq, (int64_t)tc_T(words) * TC_M));
    // CTAs = query groups x chunks: as many as fit 4 full waves of one CTA per SM (rounding UP would add a fifth,
    // nearly empty wave: +25 % time)
    static const int waves = [] { const char* e = getenv("CMH_TC_WAVES"); return (e && e[0] >= '1' && e[0] <= '8') ? e[0] - '0' : 4; }();
    int64_t want = std::max<int64_t>(1, ((int64_t)sm_count() * waves) / *n_qgroups);
    want = std::min<int64_t>(want, TC_MAX_CHUNKS / (TC_BUFS / tc_T(words)));
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
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_plan: bits=%d (1..128, +-1 codes only)", bits);
    CMH_REQUIRE(nq >= 0 && nd >= 0 && n_chunks, CMH_ERR_ARG, "cmh_tc_plan: bad arguments");
    bits = tc_eff_bits(bits);
    int64_t g, c, r;
    tc_geometry(nq, nd, tc_words(bits), &g, &c, &r);
    *n_chunks = (int)c * (TC_BUFS / tc_T(tc_words(bits)));     // candidate segments per query of one launch
    return CMH_OK;
}

namespace cmh {
int tc_collect_launch(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits, int64_t index_base,
                      const int32_t* thr, int K, int seg_base, int seg_total, int seg_cap, uint64_t* cand, uint32_t* cnt,
                      uint32_t* aux, int probe, bool skip_cnt_zero, cudaStream_t st) {
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_collect: bits=%d (1..128, +-1 codes only)", bits);
    bits = tc_eff_bits(bits);
    CMH_REQUIRE(nq >= 0 && nd >= 0 && seg_cap >= 1 && K >= 0 && index_base >= 0 && index_base + nd <= (1ll << 32),
                CMH_ERR_ARG, "cmh_tc_collect: bad sizes");
    if (nq == 0) return CMH_OK;
    int64_t n_qgroups, n_chunks, chunk_rows;
    const int words = tc_words(bits);
    tc_geometry(nq, nd, words, &n_qgroups, &n_chunks, &chunk_rows);
    const int n_segs = (int)n_chunks * (TC_BUFS / tc_T(words));
    CMH_REQUIRE(seg_base >= 0 && seg_base + n_segs <= seg_total, CMH_ERR_ARG,
                "cmh_tc_collect: segments [%d, %d) do not fit seg_total=%d (see cmh_tc_plan)", seg_base, seg_base + n_segs,
                seg_total);
    CMH_REQUIRE((int64_t)seg_total * seg_cap < (1ll << 31), CMH_ERR_UNSUPPORTED,
                "cmh_tc_collect: seg_total * seg_cap = %lld keys per query (must stay below 2^31)", (long long)seg_total * seg_cap);
    CMH_REQUIRE(q_sign && thr && cand && cnt && aux, CMH_ERR_ARG, "cmh_tc_collect: NULL pointer");
    static_assert(sizeof(TcAux) == 32, "cmh_tc_collect: aux is uint32 [nq][8]");
    // every (segment, query) count of the launch is written by the kernel; only a launch without rows needs the memset
    if (!skip_cnt_zero || nd == 0) CMH_CUDA(cudaMemsetAsync(cnt + (size_t)seg_base * nq, 0, (size_t)n_segs * nq * 4, st));
    if (nd == 0) return CMH_OK;
    if (K > 0) CMH_CUDA(cudaMemsetAsync(aux, 0, (size_t)nq * sizeof(TcAux), st));   // the tightening counters of THIS launch
    CMH_REQUIRE(d_sign, CMH_ERR_ARG, "cmh_tc_collect: NULL database");
    CMH_REQUIRE(n_qgroups <= 0x7fffffffll && chunk_rows <= 0x7fffffff, CMH_ERR_UNSUPPORTED,
                "cmh_tc_collect: launch geometry out of range");
    TcArgs a;
    a.q = q_sign; a.d = d_sign; a.thr = thr; a.cand = cand; a.cnt = cnt; a.aux = reinterpret_cast<TcAux*>(aux);
    a.nq = nq; a.nd = nd; a.index_base = index_base; a.chunk_rows = (int)chunk_rows; a.n_segs = seg_total;
    a.seg_base = seg_base; a.seg_cap = seg_cap; a.bits = bits; a.K = K; a.probe = probe; a.trace = g_tc_trace;
    // Which variant drains the hits (profiles/r02a: 8192 x 100M, 64-bit): the main launches (K > 0: thresholds close to the
    // K-th distance, few hits) run 6 % faster with the hit workers - the draining warps only park - while the pilot launches
    // (K = 0: loose thresholds, 20-30x the candidates) saturate the one worker per group and are faster with the warps
    // working their own queues off.  128-bit codes: the worker state does not fit next to 2 x 64 KB of operands.
    const bool workers = tc_workers_mode() == 1 || (tc_workers_mode() < 0 && words == 1 && K > 0 && !(probe & 128));
    const size_t smem = tc_smem_bytes(words, bits) + (workers ? (size_t)TC_WK_WORDS * 4 : 0);
    const dim3 grid((unsigned)n_qgroups, (unsigned)n_chunks);
    if (workers && bits == 32) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, 4, 6, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, 4, 6, true, 32><<<grid, TC_THREADS_WK, smem, st>>>(a);
    } else if (bits == 32) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, 4, 6, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, 4, 6, false, 32><<<grid, TC_THREADS, smem, st>>>(a);
    } else if (workers && words == 1) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, 4, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, 4, 6, true><<<grid, TC_THREADS_WK, smem, st>>>(a);
    } else if (words == 1) {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<1, 4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<1, 4, 6><<<grid, TC_THREADS, smem, st>>>(a);
    } else {
        CMH_CUDA(cudaFuncSetAttribute(tc_collect_kernel<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_collect_kernel<2, 2, 3><<<grid, TC_THREADS, smem, st>>>(a);
    }
    CMH_LAUNCH_CHECK("tc_collect_kernel");
    return CMH_OK;
}
}  // namespace cmh

extern "C" int cmh_tc_collect(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                              int64_t index_base, const int32_t* thr, int K, int seg_base, int seg_total, int seg_cap,
                              uint64_t* cand, uint32_t* cnt, uint32_t* aux, void* stream) {
    return tc_collect_launch(q_sign, nq, d_sign, nd, bits, index_base, thr, K, seg_base, seg_total, seg_cap, cand, cnt, aux,
                             0, false, (cudaStream_t)stream);
}

extern "C" int cmh_tc_probe(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                            const int32_t* thr, int seg_total, int seg_cap, uint64_t* cand, uint32_t* cnt, uint32_t* aux,
                            int probe, void* stream) {
    CMH_REQUIRE(probe >= 0 && probe < 256, CMH_ERR_ARG, "cmh_tc_probe: probe=%d", probe);
    // bit 8 (256 would collide): probes time the kernel of the MAIN launches (hit workers) unless bit 7 asks for the pilot's
    return tc_collect_launch(q_sign, nq, d_sign, nd, bits, 0, thr, (probe & 128) ? 0 : 1000000000, 0, seg_total, seg_cap, cand, cnt,
                             aux, probe, false, (cudaStream_t)stream);
}

namespace cmh {
// histogram of the candidates in segments [seg_lo, seg_hi) and / or - `need` >= 0 - the threshold rule applied to it in
// the same launch (hist / overflow may then be NULL)
int tc_cand_hist_rule(const uint64_t* cand, const uint32_t* cnt, int64_t nq, int seg_lo, int seg_hi, int seg_total, int seg_cap,
                      int nb, uint32_t* hist, uint32_t* overflow, double need, int offset, const int32_t* thr_in,
                      int32_t* thr_out, cudaStream_t st) {
    if (nq == 0) return CMH_OK;
    const ChooseRule rule{need, offset, thr_in, thr_out};
    tc_cand_hist_kernel<<<(unsigned)nq, CH_THREADS, 0, st>>>(cand, cnt, nq, seg_lo, seg_hi, seg_total, seg_cap, nb, hist, overflow, rule);
    CMH_LAUNCH_CHECK("tc_cand_hist_kernel");
    return CMH_OK;
}
int tc_choose_rule(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, double need, int offset,
                   const int32_t* thr_in, int32_t* thr_out, cudaStream_t st) {
    if (nq == 0) return CMH_OK;
    tc_choose_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(hist, overflow, nq, nb, need, offset, thr_in, thr_out);
    CMH_LAUNCH_CHECK("tc_choose_kernel");
    return CMH_OK;
}
double tc_refine_need(int64_t n_seen, int64_t nd, int K, double sigma) {
    const double kf = (double)K * (double)n_seen / (double)std::max<int64_t>(nd, 1);
    return n_seen >= nd ? (double)std::min<int64_t>(K, nd) : kf + sigma * std::sqrt(kf) + 4.0;
}
int tc_sum_ranks(const uint32_t* every, int world, int rank, int64_t n, uint32_t* lower, uint32_t* seen, cudaStream_t st) {
    if (n == 0) return CMH_OK;
    tc_sum_ranks_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(every, world, rank, n, lower, seen);
    CMH_LAUNCH_CHECK("tc_sum_ranks_kernel");
    return CMH_OK;
}
int tc_count_flags(const uint32_t* flags, int64_t n, uint32_t* count, cudaStream_t st) {
    CMH_CUDA(cudaMemsetAsync(count, 0, 4, st));
    if (n == 0) return CMH_OK;
    tc_count_flags_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 64), 256, 0, st>>>(flags, n, count);
    CMH_LAUNCH_CHECK("tc_count_flags_kernel");
    return CMH_OK;
}
constexpr size_t TC_MERGE_SMEM_MAX = 200 * 1024;
int tc_merge_verify(const uint64_t* lists, int n_lists, int64_t nq_lists, int64_t nq, int W, int K, int64_t nd_total,
                    const int32_t* thr_limit, uint64_t* keys_out, uint32_t* fail_flags, cudaStream_t st) {
    if (nq == 0) return CMH_OK;
    const size_t smem = (size_t)n_lists * W * 8;
    const int in_smem = smem <= TC_MERGE_SMEM_MAX ? 1 : 0;
    if (in_smem) CMH_CUDA(cudaFuncSetAttribute(topk_merge_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_verify_kernel<<<(unsigned)nq, 256, in_smem ? smem : 0, st>>>(lists, n_lists, nq_lists, W, K, std::min<int64_t>(K, nd_total),
                                                                          thr_limit, in_smem, keys_out, fail_flags);
    CMH_LAUNCH_CHECK("topk_merge_verify_kernel");
    return CMH_OK;
}
int tc_finalize(const uint64_t* cand, const uint32_t* cnt, const int32_t* thr_limit, int64_t nq, int n_chunks, int seg_cap, int K,
                int64_t nd, int partial, int width, int order_slack, uint64_t* keys, uint32_t* fail_flags, uint32_t* fail_count,
                cudaStream_t st) {
    if (nq == 0) return CMH_OK;
    CMH_CUDA(cudaMemsetAsync(fail_count, 0, 4, st));
    // the short variant holds at most 2048 candidates at or below the list's last bucket: for lists of a few hundred keys
    // (6+ shards at K = 1000); a query that holds more is flagged and redone exactly, so the choice only costs time
    if (partial && width <= 512)
        topk_finalize_kernel<256, 2048><<<(unsigned)nq, 256, 0, st>>>(cand, cnt, nullptr, nq, n_chunks, seg_cap, K, nd, partial, width,
                                                                     order_slack, keys, fail_flags, fail_count);
    else
        topk_finalize_kernel<FIN_THREADS, FIN_MAX><<<(unsigned)nq, FIN_THREADS, 0, st>>>(cand, cnt, partial ? nullptr : thr_limit, nq, n_chunks,
                                                                                        seg_cap, K, nd, partial, width, order_slack, keys,
                                                                                        fail_flags, fail_count);
    CMH_LAUNCH_CHECK("topk_finalize_kernel");
    return CMH_OK;
}
int tc_geometry_segs(int64_t nq, int64_t nd, int bits) {
    bits = tc_eff_bits(bits);
    int64_t g, c, r;
    tc_geometry(nq, nd, tc_words(bits), &g, &c, &r);
    return (int)c * (TC_BUFS / tc_T(tc_words(bits)));
}
}  // namespace cmh

extern "C" int cmh_tc_cand_hist(const uint64_t* cand, const uint32_t* cnt, int64_t nq, int seg_lo, int seg_hi, int seg_total,
                                int seg_cap, int nb, uint32_t* hist, uint32_t* overflow, void* stream) {
    CMH_REQUIRE(nq >= 0 && 0 <= seg_lo && seg_lo <= seg_hi && seg_hi <= seg_total && seg_cap >= 1 && nb >= 1, CMH_ERR_ARG,
                "cmh_tc_cand_hist: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(cand && cnt && hist && overflow, CMH_ERR_ARG, "cmh_tc_cand_hist: NULL pointer");
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_tc_cand_hist: too many queries per call");
    return tc_cand_hist_rule(cand, cnt, nq, seg_lo, seg_hi, seg_total, seg_cap, nb, hist, overflow, -1.0, 0, nullptr, nullptr,
                             (cudaStream_t)stream);
}

extern "C" int cmh_tc_choose(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int64_t n_seen, int64_t nd,
                             int K, double sigma, const int32_t* thr_in, int32_t* thr_out, void* stream) {
    CMH_REQUIRE(nq >= 0 && nb >= 1 && n_seen >= 0 && nd >= n_seen && K >= 1 && sigma >= 0.0, CMH_ERR_ARG,
                "cmh_tc_choose: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(hist && thr_in && thr_out, CMH_ERR_ARG, "cmh_tc_choose: NULL pointer");
    const double need = tc_refine_need(n_seen, nd, K, sigma);
    tc_choose_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(hist, overflow, nq, nb, need, 0, thr_in, thr_out);
    CMH_LAUNCH_CHECK("tc_choose_kernel");
    return CMH_OK;
}

extern "C" int cmh_tc_choose_prefix(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int K,
                                    const int32_t* thr_in, int32_t* thr_out, void* stream) {
    CMH_REQUIRE(nq >= 0 && nb >= 1 && K >= 1, CMH_ERR_ARG, "cmh_tc_choose_prefix: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(hist && thr_in && thr_out, CMH_ERR_ARG, "cmh_tc_choose_prefix: NULL pointer");
    tc_choose_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(hist, overflow, nq, nb, (double)K, -1, thr_in,
                                                                                   thr_out);
    CMH_LAUNCH_CHECK("tc_choose_kernel");
    return CMH_OK;
}

extern "C" int cmh_tc_choose_seen(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int K,
                                  const int32_t* thr_in, int32_t* thr_out, void* stream) {
    CMH_REQUIRE(nq >= 0 && nb >= 1 && K >= 1, CMH_ERR_ARG, "cmh_tc_choose_seen: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(hist && thr_in && thr_out, CMH_ERR_ARG, "cmh_tc_choose_seen: NULL pointer");
    tc_choose_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(hist, overflow, nq, nb, (double)K, 0, thr_in,
                                                                                   thr_out);
    CMH_LAUNCH_CHECK("tc_choose_kernel");
    return CMH_OK;
}

extern "C" int cmh_topk_verify(const uint64_t* keys, const int32_t* thr_limit, int64_t nq, int K, int64_t nd,
                               uint32_t* fail_flags, uint32_t* fail_count, void* stream) {
    CMH_REQUIRE(nq >= 0 && K >= 1 && nd >= 0, CMH_ERR_ARG, "cmh_topk_verify: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(keys && thr_limit && fail_flags && fail_count, CMH_ERR_ARG, "cmh_topk_verify: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CMH_CUDA(cudaMemsetAsync(fail_count, 0, 4, st));
    topk_verify_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(keys, thr_limit, nq, K, std::min<int64_t>(K, nd), fail_flags,
                                                                   fail_count);
    CMH_LAUNCH_CHECK("topk_verify_kernel");
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

extern "C" int cmh_topk_finalize(const uint64_t* cand, const uint32_t* cnt, const int32_t* thr_limit, int64_t nq, int n_chunks,
                                 int seg_cap, int K, int64_t nd, int partial, int width, uint64_t* keys, uint32_t* fail_flags,
                                 uint32_t* fail_count, void* stream) {
    CMH_REQUIRE(nq >= 0 && n_chunks >= 1 && seg_cap >= 1 && K >= 1 && nd >= 0, CMH_ERR_ARG, "cmh_topk_finalize: bad sizes");
    CMH_REQUIRE(K <= FIN_MAX, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: K=%d > %d", K, FIN_MAX);
    CMH_REQUIRE(partial ? (width >= 1 && width <= K) : (width == K || width == 0), CMH_ERR_ARG,
                "cmh_topk_finalize: width=%d (K for a whole database, 1..K for a shard)", width);
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(cand && cnt && keys && fail_flags && fail_count, CMH_ERR_ARG, "cmh_topk_finalize: NULL pointer");
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_topk_finalize: too many queries per call");
    return tc_finalize(cand, cnt, thr_limit, nq, n_chunks, seg_cap, K, nd, partial, partial ? width : K, -1, keys, fail_flags,
                       fail_count, (cudaStream_t)stream);
}

extern "C" int cmh_topk_merge_verify(const uint64_t* lists, int n_lists, int64_t nq, int width, int K, int64_t nd_total,
                                     const int32_t* thr_limit, uint64_t* keys, uint32_t* fail_flags, void* stream) {
    CMH_REQUIRE(n_lists >= 1 && nq >= 0 && width >= 1 && K >= 1 && width <= K && nd_total >= 0, CMH_ERR_ARG,
                "cmh_topk_merge_verify: bad sizes");
    if (nq == 0) return CMH_OK;
    CMH_REQUIRE(lists && keys && fail_flags, CMH_ERR_ARG, "cmh_topk_merge_verify: NULL pointer");
    CMH_REQUIRE(nq <= 0x7fffffffll, CMH_ERR_UNSUPPORTED, "cmh_topk_merge_verify: too many queries per call");
    return tc_merge_verify(lists, n_lists, nq, nq, width, K, nd_total, thr_limit, keys, fail_flags, (cudaStream_t)stream);
}
