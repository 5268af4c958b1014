// cmh_topk_tc - the tensor-core top-K search as ONE call (include/cmh_b200.h): the launch chain of the benchmarked
// path (north star config 4: stable top-K of utils/calc_utils.py:30-31 over a 100M-row database), owned by the
// library so that any caller of the C ABI reaches it - and so that the chain of ~30 small launches between the scans is
// issued in microseconds instead of through an interpreter.
//
//   sample histogram (popc) -> [all-reduce] -> thresholds
//   pilot launches (keep EVERY row at or below the threshold) -> candidate histogram -> [all-reduce] -> refined thresholds
//   main launches, cut at the prefix-rule rows: candidate histogram of everything scanned so far -> [all-reduce /
//   all-gather] -> exact tightening (cmh_tc_choose_prefix / _seen)
//   finalize (per query: K-th bucket, ordered emit)
//   sharded: all-to-all of the per-shard lists by query slice -> merge + verify -> all-gather of the verdicts
//
// Exactness never depends on a threshold: a too-low one shows up as a short candidate list or as a K-th key from an
// incomplete bucket and the query is flagged for the exact path.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>

#include "tc_internal.cuh"

using namespace cmh;

namespace {

constexpr int64_t PILOT_MIN_ROWS = 8000000;        // databases at least this long get pilot launches over their first rows
constexpr int64_t PILOT_FRACTION = 64;             // ... the first 1/64 of the rows (measured optimum)
constexpr int64_t PILOT_EARLY = 512;               // shards of at least PILOT_EARLY_MIN_ROWS refine once more, after 1/512:
constexpr int64_t PILOT_EARLY_MIN_ROWS = 64000000; // the sample's thresholds are loose there (K f < 1 sample rows)
constexpr int64_t PREFIX_MIN_ROWS = 4000000;
const double PREFIX_FRACTIONS[4] = {0.3, 0.5, 0.7, 0.85};     // one GPU (swept: 43.1 ms against 50.5 without)
const double PREFIX_FRACTIONS_SHARDED[2] = {0.3, 0.6};        // contiguous shards: each cut is an all-gather
constexpr int DEFAULT_CAP = 16384, MIN_SEG = 64, MAX_K = 4096;

uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

}  // namespace

struct cmh_tc_timing {
    cudaEvent_t phase[CMH_TC_PHASES + 1];
    cudaEvent_t collect[2 * CMH_TC_MAX_SPANS];
    int n_phase, n_collect;
    int phase_kind[CMH_TC_PHASES + 1];
};

extern "C" void cmh_tc_default_opts(cmh_tc_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->n_pilot = -1;
    o->prefix = 1;
    o->n_prefix = -1;
    o->prefix_min_rows = -1;
    o->tighten = 1;
    o->gather = 1;
}

// Cumulative local row counts (multiples of the 256-row tile) after which the thresholds are refined.  The NUMBER of
// stages depends on the whole database and the number of shards only - every shard takes part in every refinement.
extern "C" int cmh_tc_pilot_stages(int64_t nd, int64_t nd_total, int world, int64_t* rows) {
    int n = 0;
    if (nd < 0 || nd_total < PILOT_MIN_ROWS || !rows) return 0;
    if (nd_total / std::max(1, world) >= PILOT_EARLY_MIN_ROWS) rows[n++] = (nd / PILOT_EARLY) / 256 * 256;
    rows[n++] = (nd / PILOT_FRACTION) / 256 * 256;
    return n;
}

extern "C" int cmh_struct_sizes(int32_t* sizes, int n) {
    const int32_t v[6] = {(int32_t)sizeof(cmh_codeset), (int32_t)sizeof(cmh_plan), (int32_t)sizeof(cmh_comm),
                          (int32_t)sizeof(cmh_tc_opts), (int32_t)sizeof(cmh_tc_search), CMH_ABI_VERSION};
    for (int i = 0; i < n && i < 6; ++i) sizes[i] = v[i];
    return 6;
}

extern "C" int cmh_tc_search_plan(const cmh_comm* comm, int64_t nq, int64_t nd, int64_t nd_total, int bits, int K, int n_stripes,
                                  const int64_t* stripe_row, const int64_t* stripe_index, int64_t n_sample,
                                  const cmh_tc_opts* opts_in, cmh_tc_search* p) {
    CMH_REQUIRE(p, CMH_ERR_ARG, "cmh_tc_search_plan: NULL plan");
    CMH_REQUIRE(cmh_tc_supported(bits, 0), CMH_ERR_UNSUPPORTED, "cmh_tc_search_plan: bits=%d (1..128, +-1 codes only)", bits);
    bits = tc_eff_bits(bits);       // the search runs at 32 / 64 / 128 bits (padding bits agree: same distances)
    CMH_REQUIRE(nq >= 1 && nd >= 0 && nd_total >= nd && K >= 1 && K <= MAX_K, CMH_ERR_ARG,
                "cmh_tc_search_plan: bad sizes nq=%lld nd=%lld nd_total=%lld K=%d", (long long)nq, (long long)nd,
                (long long)nd_total, K);
    CMH_REQUIRE(n_stripes >= 1 && n_stripes <= CMH_TC_MAX_STRIPES && stripe_row && stripe_index, CMH_ERR_ARG,
                "cmh_tc_search_plan: 1..%d stripes", CMH_TC_MAX_STRIPES);
    cmh_tc_opts o;
    if (opts_in) o = *opts_in; else cmh_tc_default_opts(&o);
    CMH_REQUIRE(o.n_pilot <= CMH_TC_MAX_STAGES && o.n_prefix <= CMH_TC_MAX_CUTS && o.n_ready >= 0 && o.n_ready <= CMH_TC_MAX_READY,
                CMH_ERR_ARG, "cmh_tc_search_plan: too many stages / cuts / ready ranges");
    memset(p, 0, sizeof(*p));
    const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
    CMH_REQUIRE(world >= 1 && rank >= 0 && rank < world, CMH_ERR_ARG, "cmh_tc_search_plan: bad comm");
    CMH_REQUIRE(!(o.exact_thresholds && world > 1), CMH_ERR_ARG, "cmh_tc_search_plan: exact_thresholds is a one-GPU mode");
    p->bits = bits; p->K = K; p->world = world; p->rank = rank;
    p->nq = nq; p->nd = nd; p->nd_total = nd_total;
    p->n_sample = o.exact_thresholds ? nd : n_sample;
    CMH_REQUIRE(p->n_sample >= 0 && p->n_sample <= nd, CMH_ERR_ARG, "cmh_tc_search_plan: n_sample=%lld", (long long)n_sample);
    p->n_stripes = n_stripes;
    for (int j = 0; j < n_stripes; ++j) {
        p->stripe_row[j] = stripe_row[j];
        p->stripe_index[j] = stripe_index[j];
        CMH_REQUIRE(stripe_row[j] >= (j ? stripe_row[j - 1] : 0) && stripe_row[j] <= nd && stripe_index[j] >= 0, CMH_ERR_ARG,
                    "cmh_tc_search_plan: stripes must start at local row 0 and ascend within the shard");
    }
    CMH_REQUIRE(stripe_row[0] == 0, CMH_ERR_ARG, "cmh_tc_search_plan: the first stripe starts at local row 0");
    // ---- refinement stages: cumulative local row counts, the same NUMBER on every shard -------------------------------
    std::vector<int64_t> stages;
    if (o.n_pilot < 0) {
        int64_t rows[CMH_TC_MAX_STAGES];
        const int n = cmh_tc_pilot_stages(nd, nd_total, world, rows);
        stages.assign(rows, rows + n);
    } else {
        for (int i = 0; i < o.n_pilot; ++i)
            if (o.pilot_rows[i] > 0) stages.push_back(o.pilot_rows[i]);
    }
    if (o.exact_thresholds) stages.clear();          // exact thresholds need no refinement
    for (auto& e : stages) e = std::min(e, nd);
    p->n_stages = (int)stages.size();
    for (int i = 0; i < p->n_stages; ++i) p->stage_rows[i] = stages[(size_t)i];
    const int64_t last_stage = stages.empty() ? 0 : stages.back();
    // ---- the prefix rule: after these rows the candidates so far bound what later rows can still contribute ---------
    bool ascending = true;
    for (int j = 1; j < n_stripes; ++j)
        ascending = ascending && stripe_index[j] >= stripe_index[j - 1] + (stripe_row[j] - stripe_row[j - 1]);
    p->lockstep = (world > 1 && n_stripes > 1) ? 1 : 0;
    std::vector<int64_t> pcuts;
    const int64_t min_rows = o.prefix_min_rows >= 0 ? o.prefix_min_rows : PREFIX_MIN_ROWS;
    if (p->lockstep) {
        if (o.prefix)
            for (int j = 1; j < n_stripes; ++j) pcuts.push_back(stripe_row[j]);
    } else if (o.prefix && ascending && (nd_total + world - 1) / world >= min_rows) {
        if (o.n_prefix >= 0) {
            for (int i = 0; i < o.n_prefix; ++i) pcuts.push_back(std::min(nd, (int64_t)((double)nd * o.prefix_frac[i]) / 256 * 256));
        } else if (world == 1) {
            for (double f : PREFIX_FRACTIONS) pcuts.push_back(std::min(nd, (int64_t)((double)nd * f) / 256 * 256));
        } else {
            for (double f : PREFIX_FRACTIONS_SHARDED) pcuts.push_back(std::min(nd, (int64_t)((double)nd * f) / 256 * 256));
        }
    }
    CMH_REQUIRE((int)pcuts.size() <= CMH_TC_MAX_CUTS, CMH_ERR_ARG, "cmh_tc_search_plan: too many prefix cuts");
    if (p->lockstep && o.prefix)
        // a prefix exchange is never made before the last pilot stage has closed (the order of the collectives must be
        // the same on every shard); a stripe boundary inside the pilot rows would apply the strict rule (b - 1) after
        // other shards have already scanned into the next stripe - exclude it up front
        for (int64_t c : pcuts)
            CMH_REQUIRE(c >= last_stage, CMH_ERR_ARG,
                        "cmh_tc_search_plan: lockstep stripe boundary at local row %lld lies inside the pilot rows (%lld)",
                        (long long)c, (long long)last_stage);
    p->n_prefix_cuts = (int)pcuts.size();
    for (int i = 0; i < p->n_prefix_cuts; ++i) p->prefix_cut[i] = pcuts[(size_t)i];
    // ---- spans: the launches, cut at stages, prefix cuts, upload boundaries and stripe boundaries -------------------
    std::set<int64_t> cuts = {0, nd};
    for (int64_t e : stages) if (e > 0 && e < nd) cuts.insert(e);
    for (int64_t e : pcuts) if (e > last_stage && e < nd) cuts.insert(e);
    for (int i = 0; i < o.n_ready; ++i) if (o.ready_rows[i] > 0 && o.ready_rows[i] < nd) cuts.insert(o.ready_rows[i]);
    for (int j = 0; j < n_stripes; ++j) if (stripe_row[j] > 0 && stripe_row[j] < nd) cuts.insert(stripe_row[j]);
    std::vector<int64_t> cv(cuts.begin(), cuts.end());
    CMH_REQUIRE((int)cv.size() - 1 <= CMH_TC_MAX_SPANS, CMH_ERR_ARG, "cmh_tc_search_plan: %d launches (max %d)", (int)cv.size() - 1,
                CMH_TC_MAX_SPANS);
    int n_spans = 0, seg_total = 0, widest = 1;
    for (size_t i = 0; i + 1 < cv.size(); ++i) {
        if (cv[i + 1] <= cv[i]) continue;
        const int64_t lo = cv[i], hi = cv[i + 1];
        int j = 0;
        while (j + 1 < n_stripes && stripe_row[j + 1] <= lo) ++j;     // spans never straddle a stripe boundary
        p->span_lo[n_spans] = lo; p->span_hi[n_spans] = hi;
        p->span_index[n_spans] = stripe_index[j] + (lo - stripe_row[j]);
        const int segs = tc_geometry_segs(nq, hi - lo, bits);
        p->span_seg_base[n_spans] = seg_total; p->span_n_segs[n_spans] = segs;
        seg_total += segs; widest = std::max(widest, segs);
        ++n_spans;
    }
    if (n_spans == 0) {                                               // an empty shard still takes part in every exchange
        const int segs = tc_geometry_segs(nq, 0, bits);
        p->span_lo[0] = p->span_hi[0] = 0; p->span_index[0] = stripe_index[0];
        p->span_seg_base[0] = 0; p->span_n_segs[0] = segs;
        seg_total = segs; widest = segs; n_spans = 1;
    }
    p->n_spans = n_spans;
    p->seg_total = seg_total;
    p->seg_cap = o.seg_cap > 0 ? o.seg_cap : std::max(MIN_SEG, (o.cap > 0 ? o.cap : DEFAULT_CAP) / widest);
    CMH_REQUIRE((int64_t)p->seg_total * p->seg_cap < (1ll << 31), CMH_ERR_UNSUPPORTED,
                "cmh_tc_search_plan: %lld candidate slots per query", (long long)p->seg_total * p->seg_cap);
    // thresholds: slot 0 from the sample, one per closed stage, one per prefix exchange
    p->n_thr = 1 + p->n_stages + p->n_prefix_cuts;
    p->thr_limit_slot = p->n_stages;                                  // the statistical bound: after the last refinement
    p->thr_final_slot = p->n_thr - 1;
    // ---- the exchange: rank r merges query slice r; a shard sends the `exch_width` smallest keys it holds per query ---
    p->per_rank = (nq + world - 1) / world;
    // Lockstep stripes spread every stretch of the index order over all shards, so a shard's share of the K results stays
    // near K / world even when a whole tie bucket is decided by the row index: twice the mean + 64, in 64s.  Contiguous
    // shards send full lists (the lowest-index shard may own every winner of a tie bucket).  Any width is exact - a list
    // that arrives full and ends below the merged K-th key flags the query (cmh_topk_merge_verify).
    p->exch_width = (world > 1 && p->lockstep) ? (int)std::min<int64_t>(K, ((int64_t)2 * K / world + 64 + 63) / 64 * 64) : K;
    o.gather = o.gather ? 1 : 0;
    p->opts = o;
    // ---- global counts behind the statistics (sample rows, rows per stage over all shards) --------------------------
    p->n_sample_all = p->n_sample;
    for (int i = 0; i < p->n_stages; ++i) p->stage_rows_all[i] = p->stage_rows[i];
    if (world > 1) {
        uint32_t host[2 * (1 + CMH_TC_MAX_STAGES)], *dev = nullptr;
        host[0] = (uint32_t)(p->n_sample & 0xffffffff); host[1] = (uint32_t)(p->n_sample >> 32);
        for (int i = 0; i < p->n_stages; ++i) {
            host[2 + 2 * i] = (uint32_t)(p->stage_rows[i] & 0xffffffff);
            host[3 + 2 * i] = (uint32_t)(p->stage_rows[i] >> 32);
        }
        const int n = 2 * (1 + p->n_stages);
        // a few words of device memory per device, allocated once for the life of the process (cudaMalloc / cudaFree per
        // plan would serialise the device every time an index is rebuilt)
        static uint32_t* scratch[64] = {nullptr};
        int device = 0;
        CMH_CUDA(cudaGetDevice(&device));
        CMH_REQUIRE(device >= 0 && device < 64, CMH_ERR_UNSUPPORTED, "cmh_tc_search_plan: device %d", device);
        if (!scratch[device]) CMH_CUDA(cudaMalloc(&scratch[device], 256));
        dev = scratch[device];
        cudaError_t e = cudaMemcpy(dev, host, (size_t)n * 4, cudaMemcpyHostToDevice);
        int rc = e == cudaSuccess ? comm->all_reduce_u32(comm->ctx, dev, n, 0, nullptr) : cuda_fail(e, "cudaMemcpy");
        if (!rc) {
            e = cudaStreamSynchronize(nullptr);
            if (e == cudaSuccess) e = cudaMemcpy(host, dev, (size_t)n * 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "cmh_tc_search_plan: count exchange");
        }
        if (rc) return rc;
        // low and high halves were summed separately (shards hold < 2^32 rows each and there are few of them)
        p->n_sample_all = (int64_t)host[0] + ((int64_t)host[1] << 32);
        for (int i = 0; i < p->n_stages; ++i) p->stage_rows_all[i] = (int64_t)host[2 + 2 * i] + ((int64_t)host[3 + 2 * i] << 32);
    }
    // ---- workspace ---------------------------------------------------------------------------------------------------
    // candidate histograms are exchanged up to the highest bucket a threshold can name: the kernel clamps thresholds to
    // (bits - 1) / 2 (a query beyond that takes the exact path), so nothing above it is ever consulted
    const int nb = std::min(bits + 1, (bits - 1) / 2 + 2);
    const int64_t nq_all = p->per_rank * world;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { const uint64_t r = off; off += align256(bytes); return r; };
    p->off_cand = take((uint64_t)nq * p->seg_total * p->seg_cap * 8);
    p->off_cnt = take((uint64_t)p->seg_total * nq * 4);
    p->off_aux = take((uint64_t)nq * 32);
    p->off_thr = take((uint64_t)p->n_thr * nq * 4);
    // candidate histogram [nq][nb] followed by the overflow flags [nq] (one all-reduce), twice (lower / seen) + gathered
    p->off_hist = take((uint64_t)nq * (nb + 1) * 4 * 2);
    p->off_gather = (world > 1 && !p->lockstep && p->n_prefix_cuts) ? take((uint64_t)world * nq * (nb + 1) * 4) : off;
    p->off_sample_hist = take((uint64_t)nq * (bits + 1) * 4);
    p->off_part = world > 1 ? take((uint64_t)nq_all * p->exch_width * 8) : off;
    p->off_recv = world > 1 ? take((uint64_t)nq_all * p->exch_width * 8) : off;
    p->off_flags = take((uint64_t)nq_all * 4 * 2);
    if (p->n_sample > 0) {
        int rc = cmh_eval_plan(nq, p->n_sample, bits, 0, 0, 0, &p->sample_plan);
        if (rc) return rc;
        p->off_eval = take(p->sample_plan.workspace_bytes);
    } else {
        p->off_eval = off;
    }
    p->workspace_bytes = off;
    return CMH_OK;
}

extern "C" int cmh_tc_timing_create(cmh_tc_timing** out) {
    CMH_REQUIRE(out, CMH_ERR_ARG, "cmh_tc_timing_create: NULL");
    cmh_tc_timing* t = new cmh_tc_timing;
    memset(t, 0, sizeof(*t));
    for (auto& e : t->phase) CMH_CUDA(cudaEventCreate(&e));
    for (auto& e : t->collect) CMH_CUDA(cudaEventCreate(&e));
    *out = t;
    return CMH_OK;
}

extern "C" int cmh_tc_timing_destroy(cmh_tc_timing* t) {
    if (!t) return CMH_OK;
    for (auto& e : t->phase) cudaEventDestroy(e);
    for (auto& e : t->collect) cudaEventDestroy(e);
    delete t;
    return CMH_OK;
}

extern "C" int cmh_tc_timing_read(cmh_tc_timing* t, float* phase_ms, float* collect_ms, int* n_collect) {
    CMH_REQUIRE(t, CMH_ERR_ARG, "cmh_tc_timing_read: NULL");
    if (phase_ms) for (int i = 0; i < CMH_TC_PHASES; ++i) phase_ms[i] = 0.f;
    if (collect_ms) *collect_ms = 0.f;
    if (n_collect) *n_collect = t->n_collect / 2;
    if (t->n_phase == 0) return CMH_OK;
    CMH_CUDA(cudaEventSynchronize(t->phase[t->n_phase - 1]));
    for (int i = 1; i < t->n_phase; ++i) {
        float ms = 0.f;
        CMH_CUDA(cudaEventElapsedTime(&ms, t->phase[i - 1], t->phase[i]));
        if (phase_ms) phase_ms[t->phase_kind[i]] += ms;
    }
    for (int i = 0; i + 1 < t->n_collect; i += 2) {
        float ms = 0.f;
        CMH_CUDA(cudaEventElapsedTime(&ms, t->collect[i], t->collect[i + 1]));
        if (collect_ms) *collect_ms += ms;
    }
    return CMH_OK;
}

extern "C" int cmh_tc_timing_launches(cmh_tc_timing* t, float* ms, int n) {
    CMH_REQUIRE(t && (ms || n == 0), CMH_ERR_ARG, "cmh_tc_timing_launches: NULL");
    int k = 0;
    for (int i = 0; i + 1 < t->n_collect && k < n; i += 2, ++k) {
        CMH_CUDA(cudaEventSynchronize(t->collect[i + 1]));
        CMH_CUDA(cudaEventElapsedTime(&ms[k], t->collect[i], t->collect[i + 1]));
    }
    for (int i = k; i < n; ++i) ms[i] = 0.f;
    return CMH_OK;
}

namespace {

enum Phase { PH_THRESHOLDS = 0, PH_PILOT = 1, PH_MAIN = 2, PH_FINALIZE = 3, PH_EXCHANGE = 4 };

struct Search {
    const cmh_tc_search& p;
    const cmh_comm* comm;
    unsigned char* ws;
    cudaStream_t st;
    cmh_tc_timing* tm;
    int nb;
    int si = 0, pj = 0, thr_cur = 0;     // next stage / prefix exchange to close; slot of the current thresholds

    int32_t* thr(int slot) const { return reinterpret_cast<int32_t*>(ws + p.off_thr) + (int64_t)slot * p.nq; }
    uint64_t* cand() const { return reinterpret_cast<uint64_t*>(ws + p.off_cand); }
    uint32_t* cnt() const { return reinterpret_cast<uint32_t*>(ws + p.off_cnt); }
    uint32_t* hist(int which) const { return reinterpret_cast<uint32_t*>(ws + p.off_hist) + (int64_t)which * p.nq * (nb + 1); }

    int mark(int kind) {
        if (tm && tm->n_phase <= CMH_TC_PHASES) {
            tm->phase_kind[tm->n_phase] = kind;
            CMH_CUDA(cudaEventRecord(tm->phase[tm->n_phase++], st));
        }
        return CMH_OK;
    }
    int mark_collect() {
        if (tm && tm->n_collect < 2 * CMH_TC_MAX_SPANS) CMH_CUDA(cudaEventRecord(tm->collect[tm->n_collect++], st));
        return CMH_OK;
    }

    // the candidates of the rows scanned so far (kept at thresholds >= the current ones): exact counts of every bucket at
    // or below the current threshold, all-reduced over the shards, decide the thresholds of the rows still to come
    int refine(int stage, int seg_hi) {
        const double need = tc_refine_need(p.stage_rows_all[stage], p.nd_total, p.K, p.opts.sigma > 0 ? p.opts.sigma : 5.0);
        const int out = thr_cur + 1;
        int rc;
        if (p.world == 1) {
            if (seg_hi > 0) {
                rc = tc_cand_hist_rule(cand(), cnt(), p.nq, 0, seg_hi, p.seg_total, p.seg_cap, nb, nullptr, nullptr, need, 0,
                                       thr(thr_cur), thr(out), st);
            } else {
                CMH_CUDA(cudaMemsetAsync(hist(0), 0, (size_t)p.nq * (nb + 1) * 4, st));
                rc = tc_choose_rule(hist(0), nullptr, p.nq, nb, need, 0, thr(thr_cur), thr(out), st);
            }
            if (rc) return rc;
        } else {
            uint32_t* h = hist(0);
            uint32_t* over = h + (int64_t)p.nq * nb;
            if (seg_hi > 0) {
                if ((rc = tc_cand_hist_rule(cand(), cnt(), p.nq, 0, seg_hi, p.seg_total, p.seg_cap, nb, h, over, -1.0, 0, nullptr,
                                            nullptr, st)))
                    return rc;
            } else {
                CMH_CUDA(cudaMemsetAsync(h, 0, (size_t)p.nq * (nb + 1) * 4, st));
            }
            if ((rc = comm->all_reduce_u32(comm->ctx, h, p.nq * (nb + 1), 0, st))) return rc;   // flags summed: non-zero = overflow
            if ((rc = tc_choose_rule(h, over, p.nq, nb, need, 0, thr(thr_cur), thr(out), st))) return rc;
        }
        thr_cur = out;
        return CMH_OK;
    }

    // The prefix rule (exact, no statistics).  K candidates at dist <= b among rows of LOWER index than what is still to be
    // scanned: later rows only matter below b.  K candidates at dist <= b ANYWHERE among the rows scanned so far: later
    // rows only matter at or below b.
    int prefix(int seg_hi) {
        const int out = thr_cur + 1;
        const double need = (double)p.K;
        int rc;
        if (p.world == 1) {
            if (seg_hi > 0) {
                rc = tc_cand_hist_rule(cand(), cnt(), p.nq, 0, seg_hi, p.seg_total, p.seg_cap, nb, nullptr, nullptr, need, -1,
                                       thr(thr_cur), thr(out), st);
            } else {
                CMH_CUDA(cudaMemcpyAsync(thr(out), thr(thr_cur), (size_t)p.nq * 4, cudaMemcpyDeviceToDevice, st));
                rc = CMH_OK;
            }
            if (rc) return rc;
            thr_cur = out;
            return CMH_OK;
        }
        uint32_t* h = hist(0);
        if (seg_hi > 0) {
            if ((rc = tc_cand_hist_rule(cand(), cnt(), p.nq, 0, seg_hi, p.seg_total, p.seg_cap, nb, h, h + (int64_t)p.nq * nb, -1.0,
                                        0, nullptr, nullptr, st)))
                return rc;
        } else {
            CMH_CUDA(cudaMemsetAsync(h, 0, (size_t)p.nq * (nb + 1) * 4, st));
        }
        if (p.lockstep) {
            // lockstep stripes: whatever any shard has scanned lies below whatever any shard has still to scan
            if ((rc = comm->all_reduce_u32(comm->ctx, h, p.nq * (nb + 1), 0, st))) return rc;
            if ((rc = tc_choose_rule(h, h + (int64_t)p.nq * nb, p.nq, nb, need, -1, thr(thr_cur), thr(out), st))) return rc;
        } else {
            // contiguous shards: the prefixes of lower-ranked shards + this one are "lower index" (b - 1), everything seen
            // anywhere bounds the K-th distance (b)
            uint32_t* every = reinterpret_cast<uint32_t*>(ws + p.off_gather);
            const int64_t n = p.nq * (nb + 1);
            if ((rc = comm->all_gather(comm->ctx, h, every, n * 4, st))) return rc;
            uint32_t* lower = hist(0);
            uint32_t* seen = hist(1);
            if ((rc = tc_sum_ranks(every, p.world, p.rank, n, lower, seen, st))) return rc;
            // an overflowed segment anywhere invalidates the counts of that query on every shard: seen's flag column
            if ((rc = tc_choose_rule(lower, seen + (int64_t)p.nq * nb, p.nq, nb, need, -1, thr(thr_cur), thr(out), st))) return rc;
            if ((rc = tc_choose_rule(seen, seen + (int64_t)p.nq * nb, p.nq, nb, need, 0, thr(out), thr(out), st))) return rc;
        }
        thr_cur = out;
        return CMH_OK;
    }

    // exchanges due once the rows below `hi` have been scanned: the pilot stages, then - never before the last stage, so
    // that the order is the same on every shard - the prefix rule.  A shard without rows at a cut still takes part.
    int close(int64_t hi, int seg_hi) {
        int rc;
        while (si < p.n_stages && p.stage_rows[si] <= hi) {
            if (p.stage_rows_all[si] > 0) {
                if ((rc = refine(si, seg_hi))) return rc;
            } else {
                CMH_CUDA(cudaMemcpyAsync(thr(thr_cur + 1), thr(thr_cur), (size_t)p.nq * 4, cudaMemcpyDeviceToDevice, st));
                ++thr_cur;
            }
            ++si;
            if ((rc = mark(PH_PILOT))) return rc;
        }
        while (si == p.n_stages && pj < p.n_prefix_cuts && p.prefix_cut[pj] <= hi) {
            if ((rc = prefix(seg_hi))) return rc;
            ++pj;
        }
        return CMH_OK;
    }
};

}  // namespace

extern "C" int cmh_topk_tc(const cmh_tc_search* plan, const cmh_comm* comm, const uint64_t* q_sign, const uint64_t* d_sign,
                           const uint64_t* sample_sign, void* const* ready_events, uint64_t* keys, uint32_t* fail_flags,
                           uint32_t* fail_count, void* workspace, cmh_tc_timing* timing, void* stream) {
    CMH_REQUIRE(plan && q_sign && keys && fail_flags && fail_count && workspace, CMH_ERR_ARG, "cmh_topk_tc: NULL argument");
    const cmh_tc_search& p = *plan;
    const int world = comm ? comm->world : 1;
    CMH_REQUIRE(world == p.world && (comm ? comm->rank : 0) == p.rank, CMH_ERR_ARG, "cmh_topk_tc: comm does not match the plan");
    CMH_REQUIRE(p.nd == 0 || d_sign, CMH_ERR_ARG, "cmh_topk_tc: NULL database");
    CMH_REQUIRE(p.opts.n_ready == 0 || ready_events, CMH_ERR_ARG, "cmh_topk_tc: the plan expects %d ready events", p.opts.n_ready);
    CMH_REQUIRE(p.opts.exact_thresholds || p.n_sample == 0 || sample_sign, CMH_ERR_ARG, "cmh_topk_tc: NULL sample");
    cudaStream_t st = (cudaStream_t)stream;
    Search s{p, comm, reinterpret_cast<unsigned char*>(workspace), st, timing, std::min(p.bits + 1, (p.bits - 1) / 2 + 2)};
    const int nb = p.bits + 1, K = p.K, words = tc_words(p.bits);      // (s.nb: buckets of the exchanged candidate histograms)
    const int64_t nq = p.nq, nq_all = p.per_rank * world;
    int rc;
    if (timing) { timing->n_phase = 0; timing->n_collect = 0; }
    if ((rc = s.mark(PH_THRESHOLDS))) return rc;
    // ---- thresholds from the sample ---------------------------------------------------------------------------------
    uint32_t* h_sample = reinterpret_cast<uint32_t*>(s.ws + p.off_sample_hist);
    if (p.n_sample > 0) {
        cmh_codeset qs = {q_sign, nullptr, nullptr, nq};
        cmh_codeset ss = {p.opts.exact_thresholds ? d_sign : sample_sign, nullptr, nullptr, p.n_sample};
        if ((rc = cmh_eval_hist(&p.sample_plan, &qs, &ss, h_sample, nullptr, s.ws + p.off_eval, st))) return rc;
    } else {
        CMH_CUDA(cudaMemsetAsync(h_sample, 0, (size_t)nq * nb * 4, st));
    }
    if (world > 1 && (rc = comm->all_reduce_u32(comm->ctx, h_sample, nq * nb, 0, st))) return rc;
    if ((rc = cmh_topk_threshold(h_sample, nq, nb, p.n_sample_all, p.nd_total, K, s.thr(0), st))) return rc;
    if ((rc = s.mark(PH_THRESHOLDS))) return rc;
    // ---- the scan -----------------------------------------------------------------------------------------------------
    const int64_t last_stage = p.n_stages ? p.stage_rows[p.n_stages - 1] : 0;
    bool launched = false;
    if ((rc = s.close(0, 0))) return rc;
    for (int i = 0; i < p.n_spans; ++i) {
        const int64_t lo = p.span_lo[i], hi = p.span_hi[i];
        if (hi <= lo) continue;
        const bool in_pilot = hi <= last_stage;      // pilot spans keep EVERY row at or below the threshold (K = 0)
        for (int r = 0; r < p.opts.n_ready; ++r)     // a database still being uploaded: wait only for the rows this launch reads
            if (p.opts.ready_rows[r] >= hi) {
                CMH_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ready_events[r], 0));
                break;
            }
        if ((rc = s.mark_collect())) return rc;
        // tightening (main spans) uses this launch's own counts: K rows found locally are K rows found globally
        if ((rc = tc_collect_launch(q_sign, nq, d_sign + lo * words, hi - lo, p.bits, p.span_index[i], s.thr(s.thr_cur),
                                    (in_pilot || !p.opts.tighten) ? 0 : K, p.span_seg_base[i], p.seg_total, p.seg_cap, s.cand(),
                                    s.cnt(), reinterpret_cast<uint32_t*>(s.ws + p.off_aux), 0, true, st)))
            return rc;
        if ((rc = s.mark_collect())) return rc;
        launched = true;
        if (hi < p.nd && (rc = s.close(hi, p.span_seg_base[i] + p.span_n_segs[i]))) return rc;
    }
    if ((rc = s.close(p.nd, launched ? p.seg_total : 0))) return rc;
    if (!launched) CMH_CUDA(cudaMemsetAsync(s.cnt(), 0, (size_t)p.seg_total * nq * 4, st));
    // thresholds that were never written (an exchange that did not come due) do not exist: every slot up to thr_cur is
    const int32_t* thr_limit = s.thr(std::min(p.thr_limit_slot, s.thr_cur));
    if ((rc = s.mark(PH_MAIN))) return rc;
    // ---- finalize -------------------------------------------------------------------------------------------------------
    if (world == 1) {
        if ((rc = tc_finalize(s.cand(), s.cnt(), thr_limit, nq, p.seg_total, p.seg_cap, K, p.nd_total, 0, K, p.bits <= 64 ? 256 : -1, keys, fail_flags,
                              fail_count, st)))
            return rc;
        if ((rc = s.mark(PH_FINALIZE))) return rc;
        return s.mark(PH_EXCHANGE);
    }
    const int W = p.exch_width;
    uint64_t* part = reinterpret_cast<uint64_t*>(s.ws + p.off_part);       // [nq_all][W]: the slices in rank order
    uint64_t* recv = reinterpret_cast<uint64_t*>(s.ws + p.off_recv);       // [world][per_rank][W]
    uint32_t* flags_slice = reinterpret_cast<uint32_t*>(s.ws + p.off_flags);
    uint32_t* part_flags = flags_slice + nq_all;                           // finalize's own verdicts (the marker carries them)
    if (nq_all > nq) CMH_CUDA(cudaMemsetAsync(part + nq * W, 0xff, (size_t)(nq_all - nq) * W * 8, st));
    if ((rc = tc_finalize(s.cand(), s.cnt(), nullptr, nq, p.seg_total, p.seg_cap, K, p.nd_total, 1, W, p.bits <= 64 ? 256 : -1, part, part_flags, fail_count,
                          st)))
        return rc;
    if ((rc = s.mark(PH_FINALIZE))) return rc;
    // ---- exchange: slice r of every shard's lists goes to rank r, which merges and verifies it --------------------------
    if ((rc = comm->all_to_all(comm->ctx, part, recv, p.per_rank * W * 8, st))) return rc;
    const int64_t q0 = (int64_t)p.rank * p.per_rank;
    const int64_t mine = std::max<int64_t>(0, std::min<int64_t>(p.per_rank, nq - q0));     // live queries of this slice
    uint64_t* out_slice = p.opts.gather ? keys + q0 * K : keys;
    // (lists of padding queries are all pads: they would "fail" as short - only the live ones are merged)
    if ((rc = tc_merge_verify(recv, world, p.per_rank, mine, W, K, p.nd_total, thr_limit + q0, out_slice, flags_slice, st))) return rc;
    if (mine < p.per_rank) {
        CMH_CUDA(cudaMemsetAsync(flags_slice + mine, 0, (size_t)(p.per_rank - mine) * 4, st));
        CMH_CUDA(cudaMemsetAsync(out_slice + mine * K, 0xff, (size_t)(p.per_rank - mine) * K * 8, st));
    }
    if ((rc = comm->all_gather(comm->ctx, flags_slice, fail_flags, p.per_rank * 4, st))) return rc;
    if ((rc = tc_count_flags(fail_flags, nq_all, fail_count, st))) return rc;
    if (p.opts.gather && (rc = comm->all_gather(comm->ctx, out_slice, keys, p.per_rank * K * 8, st))) return rc;
    return s.mark(PH_EXCHANGE);
}
