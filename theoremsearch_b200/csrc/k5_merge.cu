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

#include "ts_common.cuh"

namespace ts {

struct MergeParams {
    const uint64_t* keys;
    int nlists, nq, k;
    int64_t stride_list, stride_query;  // element (l, q, i) at keys[l*stride_list + q*stride_query + i]
    const int64_t* list_base;           // [nlists] row offset added to each list's rows, or null
    const int64_t* id_map;              // row -> caller id, or null
    uint64_t* out_keys;                 // [nq, out_stride] or null
    float* out_scores;                  // [nq, k] or null
    int64_t* out_ids;                   // [nq, k] or null
    int64_t out_stride;                 // elements between consecutive queries in out_keys
    const int* qlist;                   // fix-up mode: work item w reads lists of item w, writes query qlist[w]
    const int* qcount;                  // ... for w < *qcount
};

__device__ __forceinline__ uint64_t rebase_key(uint64_t key, int64_t base) {
    if (key == 0ull || base == 0) return key;
    const uint32_t row = key_row(key) + (uint32_t)base;
    return (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - row);
}

template <int KPL>
__global__ void __launch_bounds__(256) merge_topk_kernel(const MergeParams p) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    uint64_t(*lists)[KPL * 32] = reinterpret_cast<uint64_t(*)[KPL * 32]>(merge_smem);  // [8][KPL*32]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = p.k;
    const int nwork = p.qcount ? *p.qcount : p.nq;
    for (int qi = blockIdx.x; qi < nwork; qi += gridDim.x) {
    const int qo = p.qlist ? p.qlist[qi] : qi;   // output row

    WarpTopK<KPL> list;
    list.clear();
    // Lists are read 32 keys at a time (one coalesced load per chunk); the first chunk of the
    // next list is requested before the current one is folded in.
    auto chunk_ptr = [&](int l) {
        return p.keys + (int64_t)l * p.stride_list + (int64_t)qi * p.stride_query;
    };
    uint64_t next = 0ull;
    if (warp < p.nlists && lane < k) next = chunk_ptr(warp)[lane];
    for (int l = warp; l < p.nlists; l += 8) {
        const uint64_t* src = chunk_ptr(l);
        const int64_t base = p.list_base ? p.list_base[l] : 0;
        uint64_t cur = next;
        if (l + 8 < p.nlists && lane < k) next = chunk_ptr(l + 8)[lane];
        uint64_t thr = list.at(k - 1);
        bool done = false;
        for (int i0 = 0; i0 < k && !done; i0 += 32) {
            if (i0 > 0) cur = (i0 + lane < k) ? src[i0 + lane] : 0ull;
            const int n = (k - i0 < 32) ? (k - i0) : 32;
            for (int i = 0; i < n; ++i) {
                const uint64_t x = rebase_key(__shfl_sync(0xFFFFFFFFu, cur, i), base);
                if (x <= thr) {
                    done = true;
                    break;
                }
                list.insert(x, lane);
                thr = list.at(k - 1);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KPL; ++j) lists[warp][j * 32 + lane] = list.key[j];
    __syncthreads();
    if (warp == 0) {
    for (int w = 1; w < 8; ++w) merge_sorted_into<KPL>(list, lists[w], k, k, lane);
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const int pos = j * 32 + lane;
        if (pos >= k) continue;
        const uint64_t key = list.key[j];
        const size_t o = (size_t)qo * k + pos;
        if (p.out_keys) p.out_keys[(size_t)qo * p.out_stride + pos] = key;
        if (p.out_scores) p.out_scores[o] = key ? key_score(key) : -INFINITY;
        if (p.out_ids) {
            int64_t id = -1;
            if (key) {
                const uint32_t row = key_row(key);
                id = p.id_map ? p.id_map[row] : (int64_t)row;
            }
            p.out_ids[o] = id;
        }
    }
    }
    __syncthreads();
    }  // work items
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
