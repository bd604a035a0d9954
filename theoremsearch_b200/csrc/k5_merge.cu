// K5 — merge of sorted top-k candidate lists (per-CTA lists inside one GPU, per-shard lists
// after the cross-GPU all-gather).
//
// The reference has no counterpart (it has no sharding); semantically this is the
// tail of `np.argsort(-scores)[:k]` (test_app.py:77) / `ORDER BY ... LIMIT k`
// (streamlit_app.py:281-282) applied to the union of the partial results.
//
// One CTA per query. Input lists are each sorted descending, so a warp folding a list into its
// register-resident WarpTopK stops at the first element that does not beat its current k-th
// key; 8 warps fold disjoint subsets of the lists, then warp 0 folds the 8 warp lists.
// Latency-bound (a few microseconds); the payload is nlists*k*8 bytes per query.
#include <algorithm>

#include "merge_device.cuh"

namespace ts {

template <int KPL>
__global__ void __launch_bounds__(256) merge_topk_kernel(const MergeParams p) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    uint64_t* lists = reinterpret_cast<uint64_t*>(merge_smem);  // [8][KPL*32]
    const int nwork = p.qcount ? *p.qcount : p.nq;
    for (int qi = blockIdx.x; qi < nwork; qi += gridDim.x) {
        const int qo = p.qlist ? p.qlist[qi] : qi;  // output row
        merge_lists<KPL>(p, qi, qo, lists, 8);
    }
}

int launch_merge_strided(const uint64_t* keys, int nlists, int nq, int k, int64_t stride_list,
                         int64_t stride_query, const int64_t* list_base, const int64_t* id_map,
                         uint64_t* out_keys, float* out_scores, int64_t* out_ids, cudaStream_t s,
                         const int* qlist, const int* qcount, int64_t out_stride) {
    TS_REQUIRE(k >= 1 && k <= TS_MAX_K, TS_ERR_BAD_ARG, "merge: k=%d out of range [1, %d]", k, TS_MAX_K);
    TS_REQUIRE(nlists >= 1 && nq >= 0, TS_ERR_BAD_ARG, "merge: nlists=%d nq=%d", nlists, nq);
    if (nq == 0) return TS_OK;
    MergeParams p;
    p.keys = keys;
    p.nlists = nlists;
    p.nq = nq;
    p.k = k;
    p.stride_list = stride_list;
    p.stride_query = stride_query;
    p.list_base = list_base;
    p.id_map = id_map;
    p.out_keys = out_keys;
    p.out_scores = out_scores;
    p.out_ids = out_ids;
    p.out_stride = out_stride ? out_stride : k;
    p.qlist = qlist;
    p.qcount = qcount;
    const int grid = qcount ? std::min(nq, 64) : nq;
    if (k <= 32) {
        merge_topk_kernel<1><<<grid, 256, 8 * 32 * 8, s>>>(p);
    } else if (k <= 128) {
        merge_topk_kernel<4><<<grid, 256, 8 * 128 * 8, s>>>(p);
    } else if (k <= 256) {
        merge_topk_kernel<8><<<grid, 256, 8 * 256 * 8, s>>>(p);
    } else {
        TS_CHECK_CUDA(cudaFuncSetAttribute(merge_topk_kernel<32>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8));
        merge_topk_kernel<32><<<grid, 256, 8 * 1024 * 8, s>>>(p);
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

int launch_merge(const uint64_t* keys, int nlists, int nq, int k, bool query_major,
                 const int64_t* list_base, const int64_t* id_map, uint64_t* out_keys,
                 float* out_scores, int64_t* out_ids, cudaStream_t s) {
    if (query_major)  // [nq][nlists][k]
        return launch_merge_strided(keys, nlists, nq, k, k, (int64_t)nlists * k, list_base, id_map, out_keys,
                                    out_scores, out_ids, s);
    // [nlists][nq][k]
    return launch_merge_strided(keys, nlists, nq, k, (int64_t)nq * k, k, list_base, id_map, out_keys, out_scores,
                                out_ids, s);
}

}  // namespace ts
