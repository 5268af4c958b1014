// The sharded forms of the counting path behind the C ABI: cmh_topk_sharded (exact popc top-K of a
// database split over the ranks of a cmh_comm) and cmh_map_k_sharded (calc_map_k_matrix, utils/calc_utils.py:16-39,
// plus precision@N and the PR curve).  One exchange step per metric; ranks are integers, so N shards == 1 shard bit for
// bit, and the one floating-point sum over shards is taken in rank order on every rank.
#include <algorithm>

#include "common.cuh"

namespace cmh {

// gathered[r][2][nq*nb] shard histograms (all, relevant) -> rows of LOWER shards and of ALL shards per (query, bucket)
__global__ void __launch_bounds__(256) shard_hist_sums_kernel(const uint32_t* __restrict__ gathered, int world, int rank, int64_t n,
                                                              uint32_t* __restrict__ lower_all, uint32_t* __restrict__ lower_rel,
                                                              uint32_t* __restrict__ glob_all, uint32_t* __restrict__ glob_rel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t la = 0, lr = 0, ga = 0, gr = 0;
    for (int r = 0; r < world; ++r) {
        const uint32_t a = gathered[((int64_t)r * 2) * n + i], b = gathered[((int64_t)r * 2 + 1) * n + i];
        ga += a; gr += b;
        if (r < rank) { la += a; lr += b; }
    }
    lower_all[i] = la; lower_rel[i] = lr; glob_all[i] = ga; glob_rel[i] = gr;
}

// ap_sum[q] = sum over ranks, in rank order, of the shards' partial sums
__global__ void __launch_bounds__(256) shard_ap_sum_kernel(const double* __restrict__ parts, int world, int64_t nq,
                                                           double* __restrict__ ap_sum) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += parts[(int64_t)r * nq + q];
    ap_sum[q] = s;
}

static uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

struct MapWs {
    uint64_t plan, pair, gathered, sums, ap_sum, ap_parts, n_rel, hits, pr, bytes;
};

static MapWs map_ws(int world, const cmh_plan& plan, int64_t nq, int bits, int ntopn) {
    MapWs w;
    uint64_t off = 0;
    auto take = [&](uint64_t b) { const uint64_t r = off; off += align256(b); return r; };
    const uint64_t cells = (uint64_t)nq * plan.nb;
    w.plan = take(plan.workspace_bytes);
    w.pair = take(cells * 4 * 2);
    w.gathered = take(cells * 4 * 2 * (uint64_t)world);
    w.sums = take(cells * 4 * 4);
    w.ap_sum = take((uint64_t)nq * 8);
    w.ap_parts = take((uint64_t)nq * 8 * (uint64_t)world);
    w.n_rel = take((uint64_t)nq * 8);
    w.hits = take((uint64_t)nq * std::max(1, ntopn) * 4);
    w.pr = take(cmh_finalize_pr_workspace_bytes(nq, bits));
    w.bytes = off;
    return w;
}

}  // namespace cmh

using namespace cmh;

extern "C" int cmh_topk_sharded(const cmh_comm* comm, const cmh_plan* plan, const cmh_codeset* q, const cmh_codeset* d_shard,
                                int K, int64_t index_base, uint64_t* gathered, uint64_t* keys, void* workspace, void* stream) {
    CMH_REQUIRE(plan && keys, CMH_ERR_ARG, "cmh_topk_sharded: NULL argument");
    const int world = comm ? comm->world : 1;
    if (world == 1) return cmh_topk(plan, q, d_shard, K, index_base, keys, workspace, stream);
    CMH_REQUIRE(gathered, CMH_ERR_ARG, "cmh_topk_sharded: NULL gather buffer");
    const int64_t nq = plan->nq;
    uint64_t* mine = gathered + (int64_t)comm->rank * nq * K;         // in place: rank r's block of the gathered lists
    int rc;
    if ((rc = cmh_topk(plan, q, d_shard, K, index_base, mine, workspace, stream))) return rc;
    if ((rc = comm->all_gather(comm->ctx, mine, gathered, nq * (int64_t)K * 8, stream))) return rc;
    return cmh_topk_merge(gathered, world, nq, K, keys, stream);
}

extern "C" uint64_t cmh_map_k_sharded_workspace_bytes(int world, int64_t nq, int64_t nd_shard, int bits, int nlab, int ternary,
                                                      int ntopn) {
    cmh_plan plan;
    if (world < 1 || cmh_eval_plan(nq, nd_shard, bits, nlab, ternary, ntopn, &plan)) return 0;
    return map_ws(world, plan, nq, bits, ntopn).bytes;
}

extern "C" int cmh_map_k_sharded(const cmh_comm* comm, const cmh_codeset* q, const cmh_codeset* d_shard, int bits, int nlab,
                                 int ternary, int64_t k, int64_t nd_total, const int64_t* topn, int ntopn, double* ap, float* map,
                                 int64_t* n_rel_out, float* prec, float* P, float* R, void* workspace, uint64_t workspace_bytes,
                                 void* stream) {
    CMH_REQUIRE(q && d_shard && map && workspace, CMH_ERR_ARG, "cmh_map_k_sharded: NULL argument");
    CMH_REQUIRE(ntopn >= 0 && ntopn <= CMH_MAX_TOPN && (ntopn == 0 || (topn && prec)), CMH_ERR_ARG, "cmh_map_k_sharded: topn / prec");
    CMH_REQUIRE((P == nullptr) == (R == nullptr), CMH_ERR_ARG, "cmh_map_k_sharded: P and R come together");
    const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
    cmh_plan plan;
    int rc = cmh_eval_plan(q->n, d_shard->n, bits, nlab, ternary, ntopn, &plan);
    if (rc) return rc;
    const int64_t nq = q->n;
    const MapWs w = map_ws(world, plan, nq, bits, ntopn);
    CMH_REQUIRE(workspace_bytes >= w.bytes, CMH_ERR_WORKSPACE, "cmh_map_k_sharded: workspace of %llu bytes, %llu needed",
                (unsigned long long)workspace_bytes, (unsigned long long)w.bytes);
    if (nq == 0) return cmh_finalize_map(nullptr, nullptr, 0, k, ap, map, stream);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
    const int64_t cells = nq * plan.nb;
    uint32_t* pair = reinterpret_cast<uint32_t*>(base + w.pair);      // [2][nq][nb]: this shard's histograms
    uint32_t* sums = reinterpret_cast<uint32_t*>(base + w.sums);      // lower_all, lower_rel, glob_all, glob_rel
    double* ap_sum = reinterpret_cast<double*>(base + w.ap_sum);
    int64_t* n_rel = reinterpret_cast<int64_t*>(base + w.n_rel);
    uint32_t* hits = reinterpret_cast<uint32_t*>(base + w.hits);
    if ((rc = cmh_eval_hist(&plan, q, d_shard, pair, pair + cells, base + w.plan, stream))) return rc;
    const uint32_t *glob_all = pair, *glob_rel = pair + cells;
    if (world > 1) {
        uint32_t* gathered = reinterpret_cast<uint32_t*>(base + w.gathered);
        if ((rc = comm->all_gather(comm->ctx, pair, gathered, cells * 4 * 2, stream))) return rc;
        shard_hist_sums_kernel<<<(unsigned)ceil_div(cells, 256), 256, 0, st>>>(gathered, world, rank, cells, sums, sums + cells,
                                                                              sums + 2 * cells, sums + 3 * cells);
        CMH_LAUNCH_CHECK("shard_hist_sums_kernel");
        glob_all = sums + 2 * cells; glob_rel = sums + 3 * cells;
        rc = cmh_eval_rank(&plan, q, d_shard, k, sums, sums + cells, glob_all, glob_rel, topn, ntopn, hits, ap_sum, n_rel,
                           base + w.plan, stream);
    } else {
        rc = cmh_eval_rank(&plan, q, d_shard, k, nullptr, nullptr, nullptr, nullptr, topn, ntopn, hits, ap_sum, n_rel,
                           base + w.plan, stream);
    }
    if (rc) return rc;
    if (world > 1) {
        double* parts = reinterpret_cast<double*>(base + w.ap_parts);
        CMH_CUDA(cudaMemcpyAsync(parts + (int64_t)rank * nq, ap_sum, (size_t)nq * 8, cudaMemcpyDeviceToDevice, st));
        if ((rc = comm->all_gather(comm->ctx, parts + (int64_t)rank * nq, parts, nq * 8, stream))) return rc;
        shard_ap_sum_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(parts, world, nq, ap_sum);
        CMH_LAUNCH_CHECK("shard_ap_sum_kernel");
        if (ntopn && (rc = comm->all_reduce_u32(comm->ctx, hits, nq * ntopn, 0, stream))) return rc;
    }
    if ((rc = cmh_finalize_map(ap_sum, n_rel, nq, k, ap, map, stream))) return rc;
    if (n_rel_out) CMH_CUDA(cudaMemcpyAsync(n_rel_out, n_rel, (size_t)nq * 8, cudaMemcpyDeviceToDevice, st));
    if (ntopn && (rc = cmh_finalize_topn(hits, n_rel, nq, topn, ntopn, nd_total, prec, stream))) return rc;
    if (P) rc = cmh_finalize_pr(glob_all, glob_rel, nq, bits, ternary, P, R, base + w.pr, stream);
    return rc;
}
