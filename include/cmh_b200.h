/* cmh_b200 - C ABI of the B200-native retrieval-evaluation path (libcmh_b200.so).
 *
 * Drop-in boundary for ONE hot path of QinLab-WFU/CLIP-based-Cross-Modal-Hashing: the retrieval evaluation in
 * utils/calc_utils.py that train/base.py:259-262,299-302 and train/TwDH/hash_train.py:221-224 run on the
 * sign()-binarised image / text hash codes.  The reference is pure Python on torch CPU ops and has no FFI of
 * its own; the entry points below are what a ctypes binding of that path binds (INTEGRATION.md shows the stub),
 * one group per reference function:
 *
 *   reference site (file:line)                                   entry points
 *   ------------------------------------------------------------ -------------------------------------------
 *   torch.sign + float buffers, train/base.py:141-146            cmh_pack_codes
 *   float multi-hot labels, dataset/base.py:89-94                cmh_pack_labels
 *   calc_hammingDist, utils/calc_utils.py:8-13                   cmh_hamming_dense
 *   calc_neighbor, utils/calc_utils.py:4-5,42-45                 cmh_neighbor_dense
 *   calc_map_k_matrix, utils/calc_utils.py:16-39                 cmh_eval_plan / cmh_eval_hist / cmh_eval_rank /
 *     gnd :26-27, sort+gather :31-33, AP :34-38                  cmh_finalize_map  (one call each: cmh_map_k)
 *   p_topK / pr_curve (north star; not in the reference)         cmh_eval_rank (topn hits) + cmh_finalize_topn,
 *                                                                cmh_finalize_pr
 *   stable top-K ranking (north star config 4)                   cmh_topk, cmh_topk_merge
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer on the current device, 8-byte aligned (16-byte aligned
 *     database planes let the ranking kernels stage rows with the bulk-copy engine; others take a slower path);
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only enqueue work;
 *     nothing synchronises unless stated.
 *   - every function returns 0 on success, a negative CMH_ERR_* for argument errors, or a positive
 *     cudaError_t.  cmh_last_error() returns a thread-local message for the last non-zero return.
 *   - packed codes: uint64 [n][words], words = ceil(bits/64); bit (c % 64) of word (c / 64) is column c;
 *     padding bits are 0.  "sign" plane: x > 0.  "valid" plane: x != 0 (NULL = every entry is +-1).
 *   - packed labels: uint64 [n][lwords], lwords = ceil(nlab/64), bit = (L != 0).
 *   - buckets: with both valid planes NULL (binary mode) bucket b = Hamming distance, nb = bits + 1; otherwise
 *     (ternary mode) bucket b = 2 * dist = bits - <q, r>, nb = 2 * bits + 1.   top-K keys always carry 2*dist.
 *   - in ternary mode the ranking entry points need BOTH valid planes (cmh_pack_codes always can write one).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef CMH_B200_H
#define CMH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMH_ABI_VERSION 2

#define CMH_OK 0
#define CMH_ERR_ARG (-1)         /* NULL / negative size / inconsistent arguments            */
#define CMH_ERR_UNSUPPORTED (-2) /* bits > CMH_MAX_BITS, unknown dtype ...                    */
#define CMH_ERR_WORKSPACE (-3)   /* workspace too small for the plan                          */
#define CMH_ERR_DEVICE (-4)      /* not an sm_100 device / no device                          */

#define CMH_MAX_BITS 4096
#define CMH_MAX_TOPN 64

enum cmh_dtype { CMH_F32 = 0, CMH_F16 = 1, CMH_BF16 = 2, CMH_F64 = 3, CMH_I8 = 4, CMH_I32 = 5, CMH_I64 = 6, CMH_U8 = 7 };

/* One side of a comparison (queries or database shard), all device pointers. */
typedef struct cmh_codeset {
    const uint64_t* sign;   /* [n][words]                                  */
    const uint64_t* valid;  /* [n][words] or NULL (all entries are +-1)    */
    const uint64_t* labels; /* [n][lwords] or NULL (no relevance needed)   */
    int64_t n;
} cmh_codeset;

/* Launch geometry chosen by cmh_eval_plan; pass the same plan to hist and rank. */
typedef struct cmh_plan {
    int32_t bits, words, nlab, lwords, ternary;
    int32_t max_topn;      /* precision@N cutoffs the workspace has room for                       */
    int32_t nb;            /* buckets (bits+1 or 2*bits+1)                                         */
    int32_t design;        /* 0 = thread-per-query tile kernels, 1 = warp-per-query generic kernels */
    int32_t q_tile;        /* queries per CTA                                                      */
    int32_t n_qtiles;
    int32_t chunk_rows;    /* database rows per CTA column (multiple of 16, <= 65520)              */
    int32_t n_chunks;
    int64_t nq, nd, nq_pad;
    uint64_t workspace_bytes;
    int32_t kq, kd;        /* sub-codes per query item / database item (1 = plain codes); cmh_eval_plan_sets */
    int32_t ap_mode;       /* 0: calc_map_k_matrix's AP; 1: textbook AP@k (cmh_eval_rank: hits within k, ranks <= k) */
    int32_t reserved;
} cmh_plan;

int cmh_abi_version(void);
const char* cmh_last_error(void);
/* sm count / compute capability / total memory of the current device */
int cmh_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem);
/* kernels launched by this library in this process so far (bench.py reports the delta as "gpu_launches") */
unsigned long long cmh_launch_count(void);
/* Integer-pipe roofline of the XOR+POPC compare, measured live: a register-only kernel with the compare's
 * instruction mix (one LOP3 + one POPC + one IADD per 32-bit word).  Best of `reps` launches of `iters`
 * iterations, in 32-bit compares per second.  Synchronises `stream`. */
int cmh_measure_popc_peak(int iters, int reps, double* popc32_per_s, void* stream);

/* ---- K1: sign + bit-pack ---------------------------------------------------------------------------------- */
/* x: device [n][ld] of `dtype` (first `bits` columns used).  sign_out / valid_out: device [n][words]
 * (valid_out may be NULL).  counters: device uint64[2], INCREMENTED by (#entries == 0, #entries not in
 * {-1, 0, +1}); may be NULL. */
int cmh_pack_codes(const void* x, int dtype, int64_t n, int bits, int64_t ld,
                   uint64_t* sign_out, uint64_t* valid_out, unsigned long long* counters, void* stream);
/* Inverse of cmh_pack_codes: float32 [n][bits] (leading dimension ld) of {-1, 0, +1} from the packed planes (valid may
 * be NULL: every entry is +-1).  The reference's .mat export (train/base.py:307-349) stores the codes as float arrays. */
int cmh_unpack_codes(const uint64_t* sign, const uint64_t* valid, int64_t n, int bits, float* out, int64_t ld, void* stream);
/* Binarise at the source (train/base.py:130-158, get_code / make_hash_code_DCHMT): row i of a batch of encoder
 * outputs is binarised, packed and written at row index[i] (index == NULL: row i) of the packed planes
 * sign_out / valid_out (device uint64 [n_out][ceil(bits/64)]; valid_out may be NULL) - the float [N, bits] code
 * buffers of the reference never exist.
 *   mode 0: x is [n][bits] (leading dimension ld) of activations: sign bit x > 0, valid bit x != 0 (torch.sign(0) == 0)
 *   mode 1: x is [n][bits][2] logits (ld >= 2*bits): sign bit = argmax is class 1 (a tie is class 0 = -1), valid bit 1
 * counters (device uint64 [2], may be NULL) += (#exact zeros, #rows whose index is outside [0, n_out): not written). */
int cmh_pack_scatter(const void* x, int dtype, int64_t n, int bits, int64_t ld, int mode, const int64_t* index,
                     int64_t n_out, uint64_t* sign_out, uint64_t* valid_out, unsigned long long* counters, void* stream);
/* The DCHMT hash head fused with binarise + pack + scatter (model/DCHMT.py:16-26 + train/base.py:150-158,176-177): `bits`
 * separate Linear(hidden, 2) + softmax + argmax (class 0 -> -1) as ONE [hidden, 2 * bits] product whose epilogue emits
 * packed bits, bit j = [logit(j, 1) > logit(j, 0)] (a tie is class 0), written at row index[i] (NULL: row i) of the
 * packed planes.  x: device [n][ld] activations of `dtype` (F32 / F16 / BF16; the first `hidden` columns; relu != 0
 * applies max(x, 0) first, model/DCHMT.py:22); weight: device float [bits][2][hidden] (the Linear weights stacked); bias:
 * device float [bits][2] or NULL; valid_out (may be NULL) gets all real bits set; counters[1] += rows whose index is
 * outside [0, n_out).  Logits are accumulated in float32 and never stored. */
int cmh_hash_head_pack(const void* x, int dtype, int64_t n, int hidden, int64_t ld, int relu, const float* weight,
                       const float* bias, int bits, const int64_t* index, int64_t n_out, uint64_t* sign_out,
                       uint64_t* valid_out, unsigned long long* counters, void* stream);
/* L: device [n][ld] multi-hot (non-negative).  out: device [n][lwords].  neg_counter: device uint64[1],
 * incremented by #entries < 0 (the reference's `dot > 0` predicate is only a set intersection for L >= 0). */
int cmh_pack_labels(const void* L, int dtype, int64_t n, int nlab, int64_t ld,
                    uint64_t* out, unsigned long long* neg_counter, void* stream);
/* counter-based synthetic packed codes (bench / tests): word w of global row r =
 * splitmix64(seed * 0x100000001B3 + r * words + w), tail bits cleared.  out: device [n][words]. */
int cmh_synth_codes(uint64_t seed, int64_t row0, int64_t n, int bits, uint64_t* out, void* stream);

/* ---- a2 / a4: dense blocks -------------------------------------------------------------------------------- */
/* out[i][j] = 0.5 * (bits - <q_i, d_j>)  float32, device [q->n][ld_out] */
int cmh_hamming_dense(const cmh_codeset* q, const cmh_codeset* d, int bits, float* out, int64_t ld_out, void* stream);
/* out[i][j] = 1.0f if label rows i, j intersect else 0.0f */
int cmh_neighbor_dense(const uint64_t* la, int64_t na, const uint64_t* lb, int64_t nb_rows, int lwords,
                       float* out, int64_t ld_out, void* stream);

/* ---- a3 / a5 / a6: ranking by counting --------------------------------------------------------------------- */
/* Fills *plan (host struct) for nq queries against an nd-row database shard on the current device.
 * nlab: number of labels (0 = no relevance, e.g. top-K); ternary: 1 when either side has a valid plane;
 * max_topn: largest ntopn a later cmh_eval_rank may pass.  plan->workspace_bytes is the device scratch needed. */
int cmh_eval_plan(int64_t nq, int64_t nd, int bits, int nlab, int ternary, int max_topn, cmh_plan* plan);
/* Same with the kernel design forced (0 tile, 1 warp, -1 automatic) - used by the tests to cover both designs
 * on small inputs. */
int cmh_eval_plan_design(int64_t nq, int64_t nd, int bits, int nlab, int ternary, int max_topn, int design,
                         cmh_plan* plan);

/* Set-valued codes (train/DPSIH/_utils.py:4-30: K embeddings per item, similarity = max over the K x K pairs, i.e.
 * distance = MIN over the pairs of sub-code Hamming distances): a query item is kq consecutive packed sub-codes
 * ([nq][kq][words]), a database item kd of them; buckets are that minimum distance (nb = bits + 1).  Binary codes only;
 * the generic warp kernels walk the passes.  ap_mode 1 makes cmh_eval_rank accumulate the TEXTBOOK AP@k that function
 * computes - relevant rows ranked within the first k, relrank / rank - with the number of such rows in hits[q][0] (the
 * caller passes topn = {k}); cmh_finalize_map_hits divides. */
int cmh_eval_plan_sets(int64_t nq, int64_t nd, int bits, int nlab, int kq, int kd, int ap_mode, cmh_plan* plan);
/* ap[q] = ap_sum[q] / hits[q * ntopn] (0 when there is no hit), *map = float(sum_q ap[q] / nq)   (DPSIH _utils.py:22-29) */
int cmh_finalize_map_hits(const double* ap_sum, const uint32_t* hits, int ntopn, int64_t nq, double* ap, float* map, void* stream);

/* Pass 1.  Per-(query, chunk) bucket histograms into `workspace` and their sum over this shard into
 * hist_all / hist_rel: device uint32 [nq][nb] (hist_rel may be NULL when q->labels is NULL). */
int cmh_eval_hist(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d,
                  uint32_t* hist_all, uint32_t* hist_rel, void* workspace, void* stream);

/* Pass 2.  Must follow cmh_eval_hist with the same plan / inputs / workspace.
 *   k            reference `k` (<0 = None = all)                                 calc_utils.py:23-24,34
 *   lower_*      device uint32 [nq][nb]: rows of LOWER-indexed shards per bucket, NULL on one GPU
 *   global_*     device uint32 [nq][nb]: rows of ALL shards per bucket,          NULL on one GPU
 *   topn/ntopn   HOST int64 list for precision@N (ntopn <= CMH_MAX_TOPN), hits: device uint32 [nq][ntopn]
 *   ap_sum       device double [nq]: sum over this shard's relevant rows with relrank <= min(k, n_rel) of
 *                relrank / rank   (AP = ap_sum / min(k, n_rel) once summed over shards)
 *   n_rel        device int64 [nq]: relevant rows over all shards                 calc_utils.py:27 */
int cmh_eval_rank(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d, int64_t k,
                  const uint32_t* lower_all, const uint32_t* lower_rel,
                  const uint32_t* global_all, const uint32_t* global_rel,
                  const int64_t* topn, int ntopn, uint32_t* hits,
                  double* ap_sum, int64_t* n_rel, void* workspace, void* stream);

/* ap[q] = ap_sum[q] / min(k, n_rel[q]) (0 when n_rel == 0), *map = float(sum_q ap[q] / nq).
 * ap (device double [nq]) may be NULL; map: device float[1].                    calc_utils.py:37-38 */
int cmh_finalize_map(const double* ap_sum, const int64_t* n_rel, int64_t nq, int64_t k,
                     double* ap, float* map, void* stream);
/* prec[i] = (1/nq) * sum over queries with n_rel > 0 of hits[q][i] / min(topn[i], nd_total).  prec: device float [ntopn] */
int cmh_finalize_topn(const uint32_t* hits, const int64_t* n_rel, int64_t nq, const int64_t* topn, int ntopn,
                      int64_t nd_total, float* prec, void* stream);
/* Hamming-radius PR curve from (global) bucket histograms.  P, R: device float [bits+1];
 * workspace: device scratch of cmh_finalize_pr_workspace_bytes(nq, bits). */
uint64_t cmh_finalize_pr_workspace_bytes(int64_t nq, int bits);
int cmh_finalize_pr(const uint32_t* hist_all, const uint32_t* hist_rel, int64_t nq, int bits, int ternary,
                    float* P, float* R, void* workspace, void* stream);

/* One-call single-GPU calc_map_k_matrix (plan + hist + rank + finalize).
 * workspace >= cmh_map_k_workspace_bytes(...). */
uint64_t cmh_map_k_workspace_bytes(int64_t nq, int64_t nd, int bits, int nlab, int ternary);
int cmh_map_k(const cmh_codeset* q, const cmh_codeset* d, int bits, int nlab, int64_t k,
              double* ap /* device [nq] or NULL */, float* map /* device [1] */,
              void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- top-K retrieval ---------------------------------------------------------------------------------------- */
/* keys: device uint64 [nq][K], ascending (2*dist << 32) | (index_base + row); padded with UINT64_MAX when
 * nd < K.  Uses the same plan / workspace as the eval passes (lwords = 0). */
int cmh_topk(const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d, int K, int64_t index_base,
             uint64_t* keys, void* workspace, void* stream);
/* keys_in: device uint64 [n_lists][nq][K] (each row ascending) -> keys_out: device [nq][K] smallest. */
int cmh_topk_merge(const uint64_t* keys_in, int n_lists, int64_t nq, int K, uint64_t* keys_out, void* stream);

/* ---- top-K retrieval on the tensor cores (tcgen05 / TMEM) ---------------------------------------------------- */
/* The +-1 int8 GEMM form of calc_hammingDist (utils/calc_utils.py:12: dist = 0.5 * (bits - qB . rB^T)) with the
 * candidate filter fused as the epilogue.  Supported: +-1 codes (no valid plane) of 1..128 bits.  A code runs at the
 * width of its packed words (64 or 128): padding bits are zero on both sides, i.e. equal, and add nothing to a distance -
 * distances, thresholds and keys are those of the real code length (cmh_tc_search.bits holds the width the search runs at). */
int cmh_tc_supported(int bits, int ternary);
/* Launch geometry of cmh_tc_collect for an (nq, nd) problem on the current device: the database is cut into
 * contiguous chunks, one CTA per (group of 512 / 256 queries, chunk); *n_chunks = the number of private candidate
 * segments every query owns in ONE launch (one per chunk for 64-bit codes, two for 128-bit codes). */
int cmh_tc_plan(int64_t nq, int64_t nd, int bits, int* n_chunks);
/* thr: device int32 [nq] - per-query upper bound of the K-th Hamming distance; every database row with
 * dist <= thr[q] (or <= a tighter bound derived while scanning when K > 0: once K rows of THIS launch at
 * dist <= thr[q] - j are known, j <= 3) is appended, in no particular order, to one of the query's segments
 * cand[q][seg_base + s][0..seg_cap) (device uint64 [nq][seg_total][seg_cap]; s < cmh_tc_plan's count) as key
 * (2*dist << 32) | (index_base + row).  cnt (device uint32 [seg_total][nq]) counts the rows found per segment and may
 * exceed seg_cap; aux: device uint32 [nq][8] bookkeeping; the launch's own segments of cnt and aux are written by the
 * call.  Several launches (row ranges of one database, e.g. a pilot range and the rest) fill disjoint segment
 * ranges of the same arrays.  K = 0 keeps the thresholds fixed.  Thresholds above (bits-1)/2 are clamped. */
int cmh_tc_collect(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits,
                   int64_t index_base, const int32_t* thr, int K, int seg_base, int seg_total, int seg_cap,
                   uint64_t* cand, uint32_t* cnt, uint32_t* aux, void* stream);
/* Measurement aid for the roofline of cmh_tc_collect: the same kernel and launch with parts of the pipeline
 * disabled, to time the ceilings in situ.  probe bit 0: no tcgen05.mma is issued (operand expansion + TMEM drain +
 * filter only); bit 1: accumulators are not drained (operand expansion + MMA only = tensor-pipe ceiling);
 * bit 2: accumulators are drained but not scanned; bits 3-5 switch off stages of the hit path (candidates decoded but
 * not stored / parked slices dropped / flagged slices not parked).  Outputs are meaningless. */
int cmh_tc_probe(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits, const int32_t* thr,
                 int seg_total, int seg_cap, uint64_t* cand, uint32_t* cnt, uint32_t* aux, int probe, void* stream);
/* Measurement aid: which variant of the kernel drains the hits - 0 the draining warps work their own queues off, 1 hit
 * workers (64-bit codes), -1 (default) chosen per launch: workers for the main launches of 64-bit codes. */
int cmh_tc_set_workers(int mode);
/* Thresholds from a pilot launch.  Step 1: hist (device uint32 [nq][nb], bucket = Hamming distance) of the candidates
 * in segments [seg_lo, seg_hi) - every row at or below thr_in[q] among the rows a K = 0 launch scanned - and
 * overflow[q] (device uint32 [nq]) = 1 when one of those segments lost entries.  A sharded database all-reduces the
 * histograms (sum) and the flags (max) between the two steps. */
int cmh_tc_cand_hist(const uint64_t* cand, const uint32_t* cnt, int64_t nq, int seg_lo, int seg_hi, int seg_total,
                     int seg_cap, int nb, uint32_t* hist, uint32_t* overflow, void* stream);
/* Step 2: thr_out[q] = the smallest bucket whose cumulative count reaches K*f + sigma*sqrt(K*f) + 4 (f = n_seen / nd,
 * n_seen = rows behind the histogram), never above thr_in[q]; thr_in[q] when overflow[q] (overflow may be NULL). */
int cmh_tc_choose(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int64_t n_seen, int64_t nd, int K,
                  double sigma, const int32_t* thr_in, int32_t* thr_out, void* stream);
/* The prefix rule (exact, no statistics): hist holds the candidates of rows that ALL precede the rows still to be
 * scanned in index order.  If K of them lie at dist <= b, a later row at dist >= b ranks after all K (larger or equal
 * distance, larger index - the stable tie order), so the rest of the database only has to be searched below b:
 * thr_out[q] = min(thr_in[q], b - 1) for the smallest such b; thr_in[q] when there is none (or overflow[q]). */
int cmh_tc_choose_prefix(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int K, const int32_t* thr_in,
                         int32_t* thr_out, void* stream);
/* The same bound without the index order (a database sharded over several GPUs): hist holds candidates of rows already
 * scanned ANYWHERE in the database (the all-gathered sum of the shards' prefix histograms).  K of them at dist <= b put
 * the K-th result at dist <= b, so no row above b is needed - rows AT b still are, their index may be lower:
 * thr_out[q] = min(thr_in[q], b).  A shard combines both: cmh_tc_choose_prefix on the sum over the shards of lower rank
 * plus its own prefix, cmh_tc_choose_seen on the sum over all shards. */
int cmh_tc_choose_seen(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, int K, const int32_t* thr_in,
                       int32_t* thr_out, void* stream);
/* thr[q] from a histogram (cmh_eval_hist, binary mode, nb = bits + 1) over a SAMPLE of n_sample rows of an nd-row
 * shard: smallest bucket whose cumulative sample count reaches K*f + 6*sqrt(K*f) + 8 (f = n_sample / nd), exactly
 * min(K, nd) when n_sample == nd. */
int cmh_topk_threshold(const uint32_t* hist, int64_t nq, int nb, int64_t n_sample, int64_t nd, int K,
                       int32_t* thr, void* stream);
/* Per query: K-th distance from the candidates' own histogram over all n_chunks (= seg_total) segments, sort of the
 * candidates at or below it, smallest keys out (UINT64_MAX pads; K <= 4096).  thr_limit (device int32 [nq] or
 * NULL): the smallest initial threshold any launch used for the query - buckets above it are incomplete.
 * partial = 0 (width = K): the segments cover the whole nd-row database; keys: device [nq][K]; fail_flags[q] (device
 * uint32 [nq]) = 1 and *fail_count (device uint32) incremented when a candidate segment overflowed, the query holds fewer
 * than min(K, nd) candidates, more than 4096 at or below the K-th distance, or its K-th distance lies above thr_limit:
 * those queries must be re-run through cmh_topk (exact path).
 * partial = 1: the segments cover one shard; the `width` (<= K) smallest keys the query holds are emitted (keys: device
 * [nq][width]); only overflow fails - the list then starts with the marker UINT64_MAX - 1 - and the merged result is
 * judged by cmh_topk_merge_verify. */
int cmh_topk_finalize(const uint64_t* cand, const uint32_t* cnt, const int32_t* thr_limit, int64_t nq, int n_chunks,
                      int seg_cap, int K, int64_t nd, int partial, int width, uint64_t* keys, uint32_t* fail_flags,
                      uint32_t* fail_count, void* stream);
/* Merge of per-shard lists (cmh_topk_finalize, partial = 1) + verdict, per query: lists: device uint64
 * [n_lists][nq][width]; keys: device [nq][K] smallest; fail_flags[q] = 1 when a list carries the overflow marker, the
 * min(K, nd_total)-th key is missing or lies above thr_limit[q] (an incomplete bucket; thr_limit may be NULL), or a list
 * that arrived full (width < K) ends below the K-th key (the cut may have cost a row). */
int cmh_topk_merge_verify(const uint64_t* lists, int n_lists, int64_t nq, int width, int K, int64_t nd_total,
                          const int32_t* thr_limit, uint64_t* keys, uint32_t* fail_flags, void* stream);
/* After the merge of per-shard results: fail_flags[q] |= the min(K, nd)-th key is a pad or its distance lies above
 * thr_limit[q] (an incomplete bucket); *fail_count = number of flagged queries. */
int cmh_topk_verify(const uint64_t* keys, const int32_t* thr_limit, int64_t nq, int K, int64_t nd, uint32_t* fail_flags,
                    uint32_t* fail_count, void* stream);

/* ---- multi-GPU transport --------------------------------------------------------------------------------------- */
/* The exchange steps of the sharded paths (one exchange per metric) go through this table.  Either the
 * built-in NCCL transport (cmh_comm_create / cmh_comm_create_rank; libnccl.so.2 is bound at run time with dlopen, the
 * copy the process has already loaded - e.g. torch's - is preferred) or a caller-supplied one: fill the struct with
 * your own functions (they enqueue on `stream` and return 0 / a positive error code).  All ranks make the same calls
 * in the same order.  With the NCCL transport one host thread drives one device. */
typedef struct cmh_comm {
    void* ctx;
    int32_t rank, world;
    /* in-place element-wise reduction of `count` uint32 over the ranks; op 0 = sum, 1 = max */
    int (*all_reduce_u32)(void* ctx, uint32_t* buf, int64_t count, int op, void* stream);
    /* recv (world * bytes) = the ranks' send buffers (bytes each), in rank order */
    int (*all_gather)(void* ctx, const void* send, void* recv, int64_t bytes, void* stream);
    /* block s (bytes each) of send goes to rank s; block s of recv came from rank s */
    int (*all_to_all)(void* ctx, const void* send, void* recv, int64_t bytes, void* stream);
} cmh_comm;

#define CMH_COMM_ID_BYTES 128
/* One process driving `ndev` devices (ncclCommInitAll): out[i] is the transport of devs[i] (devs == NULL: 0..ndev-1). */
int cmh_comm_create(int ndev, const int* devs, cmh_comm** out);
/* One process per GPU: rank 0 calls cmh_comm_unique_id, ships the 128 bytes to the other ranks by any means (a file,
 * MPI, torch.distributed.broadcast), and every rank calls cmh_comm_create_rank on its current device. */
int cmh_comm_unique_id(void* id128);
int cmh_comm_create_rank(const void* id128, int world, int rank, cmh_comm** out);
/* Measurement aid: ONE GPU plays rank `rank` of `world` statistically identical shards (sums are world x the local
 * value, gathered / exchanged blocks are copies of the local block), so that the per-shard GPU work of an N-GPU search
 * can be timed on one GPU.  Results are not a ranking of any database. */
int cmh_comm_create_loopback(int world, int rank, cmh_comm** out);
int cmh_comm_destroy(cmh_comm* comm);

/* Exact (popc) global top-K of a database sharded over the ranks of `comm`: local cmh_topk, all-gather, K-way merge.
 * plan / workspace as for cmh_topk; gathered: device uint64 [world][nq][K] scratch; keys: device [nq][K], identical on
 * every rank.  comm == NULL (or world 1): plain cmh_topk. */
int cmh_topk_sharded(const cmh_comm* comm, const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d_shard, int K,
                     int64_t index_base, uint64_t* gathered, uint64_t* keys, void* workspace, void* stream);
/* calc_map_k_matrix (utils/calc_utils.py:16-39) - plus precision@N and the PR curve - with the database sharded by
 * contiguous row ranges over the ranks of `comm` (rank r holds the rows after those of ranks < r): pass 1 per shard,
 * all-gather of the shard histograms, pass 2 with the global bucket bases + the rows of lower shards, all-gather of the
 * partial AP sums (added in rank order: the result is bit-identical on every rank and for every world size up to the
 * order of that one sum), all-reduce of the precision@N hit counts.
 *   ternary     1 when ANY shard or the queries hold exact zeros (then every side needs its valid plane)
 *   topn/ntopn  HOST list of precision@N cut-offs (may be NULL / 0); prec: device float [ntopn]
 *   ap          device double [nq] or NULL;  map: device float [1];  n_rel: device int64 [nq] or NULL
 *   P, R        device float [bits + 1] or NULL (both): Hamming-radius PR curve over the whole database
 * workspace >= cmh_map_k_sharded_workspace_bytes(world, ...).  comm == NULL: one shard. */
uint64_t cmh_map_k_sharded_workspace_bytes(int world, int64_t nq, int64_t nd_shard, int bits, int nlab, int ternary, int ntopn);
int cmh_map_k_sharded(const cmh_comm* comm, const cmh_codeset* q, const cmh_codeset* d_shard, int bits, int nlab,
                      int ternary, int64_t k, int64_t nd_total, const int64_t* topn, int ntopn, double* ap, float* map,
                      int64_t* n_rel, float* prec, float* P, float* R, void* workspace, uint64_t workspace_bytes,
                      void* stream);

/* ---- the tensor-core top-K search as ONE call ------------------------------------------------------------------- */
/* cmh_topk_tc owns the whole launch chain of the benchmarked path (north star config 4): sample histogram ->
 * thresholds -> pilot launches + refinement -> main launches with the exact prefix rule at every cut -> finalize ->
 * (sharded: all-to-all by query slice, merge + verify).  cmh_tc_search_plan fixes the geometry once per
 * (queries per call, shard) - it is a collective call when comm spans several ranks - and says how much device
 * scratch a search needs. */
#define CMH_TC_MAX_STRIPES 8
#define CMH_TC_MAX_STAGES 4
#define CMH_TC_MAX_CUTS 8
#define CMH_TC_MAX_SPANS 32
#define CMH_TC_MAX_READY 16
#define CMH_TC_PHASES 8

typedef struct cmh_tc_opts {
    int32_t n_pilot;                        /* -1: automatic stages; else cumulative local row counts in pilot_rows     */
    int64_t pilot_rows[CMH_TC_MAX_STAGES];
    int32_t prefix;                         /* apply the exact prefix rule (default 1)                                   */
    int32_t n_prefix;                       /* -1: automatic fractions; else prefix_frac[0..n_prefix) of the shard's rows */
    double prefix_frac[CMH_TC_MAX_CUTS];
    int64_t prefix_min_rows;                /* rows per shard below which the rule is not worth its launches (-1: default) */
    int32_t tighten;                        /* tighten thresholds inside the main launches (default 1)                   */
    int32_t cap;                            /* candidate slots per query and launch (0: default 16384)                   */
    int32_t seg_cap;                        /* slots per candidate segment (0: cap / segments of the widest launch)      */
    int32_t exact_thresholds;               /* 1: thresholds from a full popc histogram of the shard (no sample; world 1)  */
    int32_t gather;                         /* sharded: 1 = all-gather the merged query slices so every rank holds all keys */
    int32_t n_ready;                        /* rows that arrive with run-time events (a database still being uploaded)   */
    int64_t ready_rows[CMH_TC_MAX_READY];
    double sigma;                           /* margin of the refined thresholds (0: default 5)                           */
} cmh_tc_opts;

typedef struct cmh_tc_search {
    int32_t bits, K, world, rank;
    int64_t nq, nd, nd_total, n_sample, n_sample_all;
    int32_t n_stripes;
    int64_t stripe_row[CMH_TC_MAX_STRIPES], stripe_index[CMH_TC_MAX_STRIPES];
    int32_t n_stages;
    int64_t stage_rows[CMH_TC_MAX_STAGES], stage_rows_all[CMH_TC_MAX_STAGES];
    int32_t n_prefix_cuts;
    int64_t prefix_cut[CMH_TC_MAX_CUTS];
    int32_t lockstep;
    int32_t n_spans;
    int64_t span_lo[CMH_TC_MAX_SPANS], span_hi[CMH_TC_MAX_SPANS], span_index[CMH_TC_MAX_SPANS];
    int32_t span_seg_base[CMH_TC_MAX_SPANS], span_n_segs[CMH_TC_MAX_SPANS];
    int32_t seg_total, seg_cap;
    int32_t n_thr;                          /* threshold vectors a search writes (int32 [nq] each, at off_thr)           */
    int32_t thr_limit_slot, thr_final_slot; /* the statistical bound / the last (prefix-tightened) thresholds            */
    int64_t per_rank;                       /* queries merged by each rank (ceil(nq / world))                            */
    int32_t exch_width;                     /* keys per (query, shard) in the exchange                                   */
    cmh_tc_opts opts;
    cmh_plan sample_plan;                   /* popc histogram of the queries against the sample                          */
    uint64_t off_cand, off_cnt, off_aux, off_thr, off_hist, off_sample_hist, off_part, off_recv, off_flags, off_eval,
             off_gather;
    uint64_t workspace_bytes;
} cmh_tc_search;

void cmh_tc_default_opts(cmh_tc_opts* opts);
/* the automatic pilot stages of a shard of nd rows in a database of nd_total rows over `world` shards: cumulative local
 * row counts into rows[CMH_TC_MAX_STAGES]; returns their number (0: database too small for a pilot) */
int cmh_tc_pilot_stages(int64_t nd, int64_t nd_total, int world, int64_t* rows);
/* sizeof of {cmh_codeset, cmh_plan, cmh_comm, cmh_tc_opts, cmh_tc_search} and CMH_ABI_VERSION into sizes[0..n): lets a
 * binding check that its struct mirrors match the library.  Returns the number of values available. */
int cmh_struct_sizes(int32_t* sizes, int n);
/* stripes: (local_row, global_index) pairs - the local rows from stripe_row[j] up to the next stripe are the database
 * rows stripe_index[j], stripe_index[j] + 1, ... (one contiguous shard: n_stripes = 1, {0, index_base}).  With a comm
 * of several ranks and several stripes the stripes must be LOCKSTEP stripes (stripe j of every shard below stripe j+1
 * of every shard in global index); the prefix rule is then applied at the stripe boundaries.  n_sample: rows of the
 * sample this shard passes to cmh_topk_tc (ignored with opts.exact_thresholds). */
int cmh_tc_search_plan(const cmh_comm* comm, int64_t nq, int64_t nd, int64_t nd_total, int bits, int K, int n_stripes,
                       const int64_t* stripe_row, const int64_t* stripe_index, int64_t n_sample, const cmh_tc_opts* opts,
                       cmh_tc_search* plan);
/* optional per-phase device timing of a search (CUDA events owned by the handle) */
typedef struct cmh_tc_timing cmh_tc_timing;
int cmh_tc_timing_create(cmh_tc_timing** out);
int cmh_tc_timing_destroy(cmh_tc_timing* t);
/* after the search has completed: phase_ms[CMH_TC_PHASES] = thresholds, pilot, main, finalize, exchange (then zeros),
 * *collect_ms = sum over the tc_collect launches, *n_collect = their number.  Synchronises the last event. */
int cmh_tc_timing_read(cmh_tc_timing* t, float* phase_ms, float* collect_ms, int* n_collect);
/* the duration of each of the search's tc_collect launches, in launch order, into ms[0..n) (zeros past the last) */
int cmh_tc_timing_launches(cmh_tc_timing* t, float* ms, int n);

/* q_sign: device [nq][words]; d_sign: this shard's rows; sample_sign: device [n_sample][words] (any subset of the
 * shard's rows; NULL with opts.exact_thresholds); ready_events: opts.n_ready cudaEvent_t (rows below ready_rows[i] are
 * valid once event i has completed), else NULL.
 * keys: device uint64.  world 1: [nq][K].  Sharded, gather 0: [per_rank][K] - the rows of queries
 * rank * per_rank ... (-1 pads past nq); gather 1: [per_rank * world][K].
 * fail_flags: device uint32 [per_rank * world] (identical on every rank), *fail_count (device uint32) their number:
 * those queries must be redone through the exact path (cmh_topk / cmh_topk_sharded); their keys are pads.
 * Everything is enqueued on `stream`; nothing synchronises. */
int cmh_topk_tc(const cmh_tc_search* plan, const cmh_comm* comm, const uint64_t* q_sign, const uint64_t* d_sign,
                const uint64_t* sample_sign, void* const* ready_events, uint64_t* keys, uint32_t* fail_flags,
                uint32_t* fail_count, void* workspace, cmh_tc_timing* timing, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMH_B200_H */
