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
    p.done_flag = nullptr;
    p.done_value = 0;
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

// ---- exchange kernel of the sharded exact search (ts_search_sharded, two-kernel form) ----------------------
// One CTA per query. Launched with programmatic stream serialisation right behind the scan kernel of the
// same query: it becomes resident while the scan is still streaming (256 threads and 2-64 KB of shared
// memory fit beside the scan's CTA), immediately lets ITS successor — the scan of the next query — launch,
// and then sleeps in griddepcontrol.wait until the scan has completed and its per-CTA lists are visible.
// The next scan therefore starts filling SMs the moment this query's scan CTAs retire, while this CTA
// merges, pushes the keys to the peers over NVLink, waits for theirs and writes the result: neither the
// launch gap nor the cross-GPU wait sits on the scan stream's critical path.
template <int KPL>
__global__ void __launch_bounds__(256) xchg_finish_kernel(const MergeParams local, const MergeParams fin,
                                                           const XchgDev x) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    uint64_t* lists = reinterpret_cast<uint64_t*>(merge_smem);  // [8][KPL*32]
    griddep_launch_dependents();
    griddep_wait();
    const int qi = blockIdx.x;
    exchange_and_merge<KPL>(x, fin, local, qi, qi, local.k, lists, 8);
}

int launch_xchg_finish(const uint64_t* part_keys, int nparts, int nq, int k, const XchgDev& x, const int64_t* id_map,
                       float* out_scores, int64_t* out_ids, cudaStream_t s, uint32_t* done_flag, uint32_t done_value) {
    TS_REQUIRE(k >= 1 && k <= TS_MAX_K && nq >= 1, TS_ERR_BAD_ARG, "xchg_finish: nq=%d k=%d", nq, k);
    MergeParams local;
    memset(&local, 0, sizeof(local));
    local.keys = part_keys;
    local.nlists = nparts;
    local.nq = nq;
    local.k = k;
    local.stride_list = k;
    local.stride_query = (int64_t)nparts * k;
    MergeParams fin;
    memset(&fin, 0, sizeof(fin));
    fin.nq = nq;
    fin.k = k;
    fin.id_map = id_map;
    fin.out_scores = out_scores;
    fin.out_ids = out_ids;
    fin.out_stride = k;
    fin.done_flag = done_flag;
    fin.done_value = done_value;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nq);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (k <= 32) {
        cfg.dynamicSmemBytes = 8 * 32 * 8;
        TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, xchg_finish_kernel<1>, local, fin, x));
    } else if (k <= 128) {
        cfg.dynamicSmemBytes = 8 * 128 * 8;
        TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, xchg_finish_kernel<4>, local, fin, x));
    } else if (k <= 256) {
        cfg.dynamicSmemBytes = 8 * 256 * 8;
        TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, xchg_finish_kernel<8>, local, fin, x));
    } else {
        cfg.dynamicSmemBytes = 8 * 1024 * 8;
        TS_CHECK_CUDA(cudaFuncSetAttribute(xchg_finish_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           8 * 1024 * 8));
        TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, xchg_finish_kernel<32>, local, fin, x));
    }
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // namespace ts
