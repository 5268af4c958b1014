// Internal interfaces between tc_collect.cu (kernels of the tensor-core top-K) and tc_search.cu (the one-call
// orchestration, cmh_topk_tc) - not part of the C ABI.
#pragma once

#include "common.cuh"

namespace cmh {

// The width a code length runs at on the tensor path: the width of its packed words.  Padding bits are zero on both
// sides by contract (cmh_pack_codes, cmh_synth_codes), i.e. equal, so they add nothing to a Hamming distance: a 16-bit
// code IS a 64-bit code whose upper 48 bits agree everywhere, and every distance, threshold and key is the same number.
// Codes of up to 32 bits use the low half of their word only: one K-step per field instead of two (K = 32 of tcgen05.mma
// kind::i8 is exactly such a code), half the expansion work for the producers.
inline int tc_eff_bits(int bits) { return bits <= 32 ? 32 : (bits <= 64 ? 64 : 128); }
inline int tc_words(int eff_bits) { return (eff_bits + 63) / 64; }

// candidate segments one cmh_tc_collect launch over nd rows fills per query (cmh_tc_plan)
int tc_geometry_segs(int64_t nq, int64_t nd, int bits);
// the launch itself; skip_zero: cnt / aux of the launch's segments have been zeroed by the caller (one memset per search)
int tc_collect_launch(const uint64_t* q_sign, int64_t nq, const uint64_t* d_sign, int64_t nd, int bits, int64_t index_base,
                      const int32_t* thr, int K, int seg_base, int seg_total, int seg_cap, uint64_t* cand, uint32_t* cnt,
                      uint32_t* aux, int probe, bool skip_cnt_zero, cudaStream_t st);
// histogram of the candidates in segments [seg_lo, seg_hi) (hist / overflow may be NULL) and - need >= 0 - the threshold
// rule applied to it in the same launch: thr_out = min(thr_in, b + offset) for the smallest bucket b <= thr_in whose
// cumulative count reaches `need` (thr_in when none, or when a segment overflowed)
int tc_cand_hist_rule(const uint64_t* cand, const uint32_t* cnt, int64_t nq, int seg_lo, int seg_hi, int seg_total, int seg_cap,
                      int nb, uint32_t* hist, uint32_t* overflow, double need, int offset, const int32_t* thr_in,
                      int32_t* thr_out, cudaStream_t st);
int tc_choose_rule(const uint32_t* hist, const uint32_t* overflow, int64_t nq, int nb, double need, int offset,
                   const int32_t* thr_in, int32_t* thr_out, cudaStream_t st);
double tc_refine_need(int64_t n_seen, int64_t nd, int K, double sigma);
int tc_sum_ranks(const uint32_t* every, int world, int rank, int64_t n, uint32_t* lower, uint32_t* seen, cudaStream_t st);
int tc_count_flags(const uint32_t* flags, int64_t n, uint32_t* count, cudaStream_t st);
// lists [n_lists][nq_lists][W]; the first nq queries of every block are merged
int tc_merge_verify(const uint64_t* lists, int n_lists, int64_t nq_lists, int64_t nq, int W, int K, int64_t nd_total,
                    const int32_t* thr_limit, uint64_t* keys_out, uint32_t* fail_flags, cudaStream_t st);
// order_slack: >= 0 when a query's candidates are stored in index order up to a displacement of that many rows (256: one
// segment per chunk, appended tile by tile); -1 when nothing is known about the stored order
int tc_finalize(const uint64_t* cand, const uint32_t* cnt, const int32_t* thr_limit, int64_t nq, int n_chunks, int seg_cap, int K,
                int64_t nd, int partial, int width, int order_slack, uint64_t* keys, uint32_t* fail_flags, uint32_t* fail_count,
                cudaStream_t st);

}  // namespace cmh
