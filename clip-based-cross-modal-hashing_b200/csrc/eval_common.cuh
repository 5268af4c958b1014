// Shared pieces of the ranking-by-counting kernels (eval_tile.cu, eval_warp.cu, eval_host.cu).
//
// Restatement of utils/calc_utils.py:25-37 that never sorts and never materialises the Q x D distance matrix:
//   rank_j    = #{i : d_i < d_j} + #{i < j : d_i == d_j} + 1      (torch.sort :31 forced stable)
//   relrank_j = same over relevant rows only                        (gnd[ind], nonzero :33,36)
//   AP_q      = (1/total) * sum_{j relevant, relrank_j <= total} relrank_j / rank_j,  total = min(k, n_rel)  (:34-37)
// Pass 1 counts rows per (query, database chunk, bucket); an exclusive scan turns the counts into the rank of the
// first row of every (chunk, bucket); pass 2 walks each chunk in index order and hands out ranks.
#pragma once

#include "common.cuh"

namespace cmh {

constexpr int QT = 128;        // queries per CTA in the thread-per-query (tile) design
constexpr int TILE_ROWS = 256; // database rows per shared-memory stage
constexpr int MAX_CW = 8;      // 32-bit code words the tile design keeps in registers (bits <= 256)
constexpr int MAX_LW = 16;     // 32-bit label words kept in registers (<= 512 labels)
constexpr int TILE_MAX_NB = 200;   // buckets the tile design can keep per thread (8 B each, 128 threads)
constexpr int WARP_MAX_NB = 2 * CMH_MAX_BITS + 1;

// Workspace carve-up (all offsets 256-byte aligned).  Layout "T": [chunk][bucket][query] (query fastest, so a
// CTA's 128 threads read their private counter columns coalesced).  Layout "W": [chunk][query][bucket].
struct Workspace {
    uint32_t* chunk_hist;   // packed (all | rel << 16) per (chunk, bucket, query)
    uint2* base;            // (all, rel) rows ranked before the first row of (chunk, bucket) for the query
    uint32_t* shard_all;    // [nq][nb] this shard's totals
    uint32_t* shard_rel;    // [nq][nb]
    uint32_t* total;        // [nq_pad] min(k, n_rel)
    int32_t* thr;           // [nq_pad] top-K threshold bucket
    uint32_t* thr_quota;    // [nq_pad] rows of the threshold bucket that still fit in the top K
    double* ap_part;        // [n_chunks][nq_pad]
    uint32_t* hits_part;    // [n_chunks][max_topn][nq_pad]  (first-crossing counts)
    uint64_t bytes;
};

Workspace carve_workspace(const cmh_plan& p, void* ws);

struct EvalArgs {
    const uint32_t *qs, *qv, *ql;   // query planes viewed as 32-bit words
    const uint32_t *ds, *dv, *dl;   // database planes
    int64_t nq, nd, nq_pad;
    int cw_stride, lw_stride;       // 32-bit words per row in memory (2 * words, 2 * lwords)
    int cw, lw;                     // 32-bit words that can be non-zero
    int bits, nb;
    int chunk_rows, n_chunks;
    int bulk_ok;                    // database planes are 16-byte aligned: stages may use the bulk-copy engine
    // pass 2
    int64_t index_base;
    int ntopn, K;
    uint32_t nmax;                  // largest precision@N cutoff (0 = none)
    int big_ranks;                  // ranks may reach 2^23 and beyond (a long shard, or bases from other shards)
    int kq, kd;                     // set-valued codes: sub-codes per query / database item (1 = plain; warp kernels only)
    int ap_mode;                    // 1: textbook AP@k - a relevant row counts when its RANK is within kcut
    uint32_t kcut;
};

// precision@N cutoffs, passed to kernels by value
struct TopnList {
    uint32_t n[CMH_MAX_TOPN];    // ascending cutoffs (unused slots = UINT32_MAX)
    int32_t perm[CMH_MAX_TOPN];  // position of each sorted cutoff in the caller's list
};

__device__ __forceinline__ int64_t hist_index_T(const EvalArgs& a, int chunk, int b, int64_t q) {
    return ((int64_t)chunk * a.nb + b) * a.nq_pad + q;
}
__device__ __forceinline__ int64_t hist_index_W(const EvalArgs& a, int chunk, int b, int64_t q) {
    return ((int64_t)chunk * a.nq_pad + q) * a.nb + b;
}

}  // namespace cmh
