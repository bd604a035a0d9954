// Device-side merge of descending candidate lists — shared by K5 (stand-alone merge kernel) and
// K2 (whose last CTA to finish merges the grid's per-CTA lists in-kernel).
#pragma once

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
    // host-buffer latency path (single query): the outputs above point into pinned, device-mapped HOST memory and,
    // once they are written, *done_flag (also host-mapped) receives done_value — the host polls the flag instead
    // of paying two D2H copies and a stream synchronise. nullptr: no signal.
    uint32_t* done_flag;
    uint32_t done_value;
};

__device__ __forceinline__ void merge_signal_done(const MergeParams& p) {
    if (p.done_flag != nullptr && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(p.done_flag) = p.done_value;
    }
}

// Descending bitonic sort of P keys (power of two, >= 64) in shared memory by every thread of the CTA.
// A warp owns 64-key windows: all compare-exchange levels with stride <= 32 run in registers (two keys per
// lane, shuffles), so only strides >= 64 cost a shared-memory pass and a barrier — 15 barriers instead of 55
// for 1024 keys.
__device__ __forceinline__ void cta_bitonic_sort_desc(uint64_t* buf, int P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    auto reg_block = [&](int size_first, int size_last) {
        for (int wb = warp * 64; wb < P; wb += nw * 64) {
            uint64_t a0 = buf[wb + lane], a1 = buf[wb + 32 + lane];
            for (int size = size_first; size <= size_last; size <<= 1) {
                const bool desc0 = ((wb + lane) & size) == 0;
                const bool desc1 = ((wb + 32 + lane) & size) == 0;
                for (int st = (size >> 1) > 32 ? 32 : (size >> 1); st >= 1; st >>= 1) {
                    if (st == 32) {
                        if ((a0 < a1) == desc0) {
                            const uint64_t t = a0;
                            a0 = a1;
                            a1 = t;
                        }
                    } else {
                        const bool lower = (lane & st) != 0;
                        const uint64_t o0 = __shfl_xor_sync(0xFFFFFFFFu, a0, st);
                        const uint64_t o1 = __shfl_xor_sync(0xFFFFFFFFu, a1, st);
                        const bool t0 = (desc0 != lower) ? (o0 > a0) : (o0 < a0);
                        const bool t1 = (desc1 != lower) ? (o1 > a1) : (o1 < a1);
                        a0 = t0 ? o0 : a0;
                        a1 = t1 ? o1 : a1;
                    }
                }
            }
            buf[wb + lane] = a0;
            buf[wb + 32 + lane] = a1;
        }
        __syncthreads();
    };
    reg_block(2, 64 < P ? 64 : P);
    for (int size = 128; size <= P; size <<= 1) {
        for (int st = size >> 1; st >= 64; st >>= 1) {
            for (int i = threadIdx.x; i < P / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (st - 1));
                const int hi = lo + st;
                const bool desc = (lo & size) == 0;
                const uint64_t a = buf[lo], b = buf[hi];
                if ((a < b) == desc) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
            __syncthreads();
        }
        reg_block(size, size);
    }
}

__device__ __forceinline__ uint64_t rebase_key(uint64_t key, int64_t base) {
    if (key == 0ull || base == 0) return key;
    const uint32_t row = key_row(key) + (uint32_t)base;
    return (key & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - row);
}

// Merge work item `qi` (its p.nlists lists) and write output row `qo`. Called by ALL threads of a
// CTA with `nwarps` warps; `lists` is smem scratch [nwarps][KPL*32]. Lists are read with ld.global.cg
// (L2) so a CTA may merge keys other CTAs of the same grid wrote just before (after a fence).
// Write output position `pos` of query row `qo` (key 0 = empty slot).
__device__ __forceinline__ void merge_emit(const MergeParams& p, int qo, int pos, uint64_t key) {
    const size_t o = (size_t)qo * p.k + pos;
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

// k <= 32, 16..256 lists (the 148 per-CTA lists of a scan): the pruned merge. The k-th largest list HEAD T is a lower
// bound on the k-th best key overall (the k largest heads are k distinct keys >= T), so only keys >= T can be in
// the result — typically ~2k of the nlists * k keys, held by ~k lists. Two L2 round trips (heads; the qualifying
// lists) and two rank-by-counting passes in shared memory replace ~nlists/nwarps dependent register merges per
// warp plus a serial fold of the warps' lists: ~3 us instead of ~9 us for 148 lists. Returns false (nothing
// written) when more keys survive than the 256-entry buffer holds — the caller then runs the general merge.
constexpr int MERGE_PRUNE_MAX = 256;
__device__ __forceinline__ bool merge_lists_pruned(const MergeParams& p, int qi, int qo, uint64_t* surv /* [>= 256] */) {
    __shared__ uint64_t s_heads[MERGE_PRUNE_MAX];
    __shared__ unsigned long long s_T;
    __shared__ int s_cnt;
    const int k = p.k, nl = p.nlists;
    auto list_ptr = [&](int l) { return p.keys + (int64_t)l * p.stride_list + (int64_t)qi * p.stride_query; };
    uint64_t head = 0ull;
    int64_t base = 0;
    const int l = threadIdx.x;                       // thread l owns list l (blockDim.x >= 256 >= nl)
    if (l < nl) {
        base = p.list_base ? p.list_base[l] : 0;
        head = rebase_key(__ldcg(list_ptr(l)), base);
    }
    if (l < MERGE_PRUNE_MAX) s_heads[l] = head;
    if (threadIdx.x == 0) {
        s_T = 0ull;                                  // fewer than k non-empty lists: every key survives
        s_cnt = 0;
    }
    __syncthreads();
    if (head != 0ull) {                              // rank of this head among the heads (unique keys)
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < nl; ++j) rank += (s_heads[j] > head) ? 1 : 0;
        if (rank == k - 1) s_T = head;
    }
    __syncthreads();
    const uint64_t T = s_T;
    if (head != 0ull && head >= T) {                 // a qualifying list: fetch it whole (independent loads), keep keys >= T
        const uint64_t* src = list_ptr(l);
        uint64_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (j < k) ? __ldcg(src + j) : 0ull;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint64_t key = rebase_key(v[j], base);
            if (key != 0ull && key >= T) {
                const int pos = atomicAdd(&s_cnt, 1);
                if (pos < MERGE_PRUNE_MAX) surv[pos] = key;
            }
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt > MERGE_PRUNE_MAX) return false;         // uniform: the buffer overflowed, nothing has been written
    if ((int)threadIdx.x < cnt) {
        const uint64_t key = surv[threadIdx.x];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) rank += (surv[j] > key) ? 1 : 0;
        if (rank < k) merge_emit(p, qo, rank, key);
    }
    for (int pos = cnt + (int)threadIdx.x; pos < k; pos += blockDim.x) merge_emit(p, qo, pos, 0ull);   // padding
    __syncthreads();
    merge_signal_done(p);
    return true;
}

template <int KPL>
__device__ __forceinline__ void merge_lists(const MergeParams& p, int qi, int qo, uint64_t* lists, int nwarps) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = p.k;
    if constexpr (KPL == 1) {
        if (p.nlists >= 16 && p.nlists <= MERGE_PRUNE_MAX && blockDim.x >= MERGE_PRUNE_MAX && nwarps * 32 >= MERGE_PRUNE_MAX) {
            if (merge_lists_pruned(p, qi, qo, lists)) return;
        }
    }
    WarpTopK<KPL> list;
    list.clear();
    auto list_ptr = [&](int l) { return p.keys + (int64_t)l * p.stride_list + (int64_t)qi * p.stride_query; };
    if constexpr (KPL == 1) {
        // k <= 32: a list is one key per lane. Fetch this warp's lists eight at a time (eight independent L2
        // loads in flight instead of a dependent chain of ~1 us round trips), then fold them in.
        for (int l0 = warp; l0 < p.nlists; l0 += 8 * nwarps) {
            uint64_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int l = l0 + u * nwarps;
                v[u] = (l < p.nlists && lane < k) ? __ldcg(list_ptr(l) + lane) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int l = l0 + u * nwarps;
                if (l >= p.nlists) break;
                uint64_t cur[1] = {rebase_key(v[u], p.list_base ? p.list_base[l] : 0)};
                const uint64_t head = __shfl_sync(0xFFFFFFFFu, cur[0], 0);
                if (head > list.kth(k)) list.merge_desc(cur, lane);
            }
        }
    } else if constexpr (KPL <= 8) {
        // Every list is fetched whole into registers in the WarpTopK layout (coalesced 256-byte loads; the
        // next list is requested before the current one is merged) and folded in with the bitonic merge.
        auto fetch = [&](int l, uint64_t (&b)[KPL]) {
            const uint64_t* src = list_ptr(l);
#pragma unroll
            for (int j = 0; j < KPL; ++j) b[j] = (j * 32 + lane < k) ? __ldcg(src + j * 32 + lane) : 0ull;
        };
        uint64_t nxt[KPL];
        if (warp < p.nlists) fetch(warp, nxt);
        for (int l = warp; l < p.nlists; l += nwarps) {
            uint64_t cur[KPL];
#pragma unroll
            for (int j = 0; j < KPL; ++j) cur[j] = nxt[j];
            if (l + nwarps < p.nlists) fetch(l + nwarps, nxt);
            const int64_t base = p.list_base ? p.list_base[l] : 0;
#pragma unroll
            for (int j = 0; j < KPL; ++j) cur[j] = rebase_key(cur[j], base);
            const uint64_t head = __shfl_sync(0xFFFFFFFFu, cur[0], 0);
            if (head > list.kth(k)) list.merge_desc(cur, lane);
        }
    } else {
        // k > 256: lists are read 32 keys at a time and inserted element-wise, stopping at the first key of a
        // list that cannot enter (see merge_sorted_into)
        for (int l = warp; l < p.nlists; l += nwarps) {
            const uint64_t* src = list_ptr(l);
            const int64_t base = p.list_base ? p.list_base[l] : 0;
            uint64_t thr = list.kth(k);
            bool done = false;
            for (int i0 = 0; i0 < k && !done; i0 += 32) {
                const uint64_t cur = (i0 + lane < k) ? __ldcg(src + i0 + lane) : 0ull;
                const int n = (k - i0 < 32) ? (k - i0) : 32;
                for (int i = 0; i < n; ++i) {
                    const uint64_t x = rebase_key(__shfl_sync(0xFFFFFFFFu, cur, i), base);
                    if (x <= thr) {
                        done = true;
                        break;
                    }
                    list.insert(x, lane);
                    thr = list.kth(k);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KPL; ++j) lists[(size_t)warp * (KPL * 32) + j * 32 + lane] = list.key[j];
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < nwarps; ++w) merge_sorted_into<KPL>(list, lists + (size_t)w * (KPL * 32), k, k, lane);
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
    merge_signal_done(p);
}

// ---------------------------------------------------------------------------------------------- peer exchange
__device__ __forceinline__ void griddep_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns2() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The cross-GPU step of the sharded exact search, run by ALL threads of one CTA for query `qi`:
//   1) merge this shard's per-CTA lists (`local`, work item `wi`) into the shard's top-k (local rows), parked
//      in this rank's own slot of its own receive area;
//   2) rebase the keys to global rows and store them into EVERY rank's receive area (plain stores through the
//      CUDA-IPC mappings: NVLink for the peers), system-scope fence, then raise the (rank, query) sequence flag
//      in every area with st.release.sys;
//   3) wait (ld.acquire.sys, bounded by x.timeout_ns) until all `world` flags of this query carry x.seq;
//   4) merge the world lists and write the final (scores, ids) — or, if a peer did not arrive in time, raise the
//      sticky error flag (host-mapped memory: the host sees it without synchronising) and write a poisoned
//      result (-inf, -1) so that a stale or partial merge can never pass for an answer.
// Slots and flags are double-buffered by sequence parity: a rank finishes query s only after every rank has
// pushed s, so nobody can be more than one query ahead of a reader. Exchanges of one rank run strictly in
// sequence: step 0 waits until *x.done_seq == seq - 1 (the exchange kernels of consecutive searches are chained
// by programmatic dependent launch, which orders their starts, not their completions) and the last step
// publishes *x.done_seq = seq, which also releases the scan kernel's ring slot (k2_scan_impl.cuh).
// The queries of one search finish in separate CTAs: the last of them publishes the search as done.
__device__ __forceinline__ void xchg_publish_done(const XchgDev& x) {
    if (threadIdx.x == 0) {
        __threadfence();
        if (x.nq == 1 || atomicAdd(x.done_count, 1u) == (unsigned)x.nq - 1u) {
            if (x.nq > 1) *x.done_count = 0u;
            __threadfence();
            st_release_gpu_u32(x.done_seq, x.seq);
        }
    }
}

template <int KPL>
__device__ __forceinline__ void exchange_and_merge(const XchgDev& x, const MergeParams& fin, MergeParams local, int wi,
                                                   int qi, int k, uint64_t* lists, int nwarps) {
    __shared__ int s_timed_out;
    const uint32_t par = x.seq & 1u;
    const size_t slot_sz = (size_t)x.max_k;
    const size_t my_slot = (((size_t)par * x.world + x.rank) * x.max_nq + qi) * slot_sz;
    local.out_keys = x.my_slots + my_slot - (size_t)qi * slot_sz;   // merge_lists adds qi * out_stride
    local.out_stride = (int64_t)slot_sz;
    local.out_scores = nullptr;
    local.out_ids = nullptr;
    local.id_map = nullptr;
    if (threadIdx.x == 0) {
        s_timed_out = 0;
        const unsigned long long t0 = globaltimer_ns2();
        while ((int32_t)(ld_acquire_gpu_u32(x.done_seq) - (x.seq - 1u)) < 0) {
            if (globaltimer_ns2() - t0 > x.timeout_ns) {
                s_timed_out = 1;
                *reinterpret_cast<volatile int*>(x.error) = 1;
                break;
            }
        }
    }
    merge_lists<KPL>(local, wi, qi, lists, nwarps);
    __threadfence();
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = rebase_key(__ldcg(x.my_slots + my_slot + i), x.base);
        for (int g = 0; g < x.world; ++g) x.peer_slots[g][my_slot + i] = key;
    }
    __threadfence_system();
    __syncthreads();
    const size_t flag_idx = ((size_t)par * x.world + x.rank) * x.max_nq + qi;
    if ((int)threadIdx.x < x.world) {
        if (!x.debug_no_flag) st_release_sys_u32(x.peer_flags[threadIdx.x] + flag_idx, x.seq);
        const uint32_t* f = x.my_flags + ((size_t)par * x.world + threadIdx.x) * x.max_nq + qi;
        const unsigned long long t0 = globaltimer_ns2();
        while (ld_acquire_sys_u32(f) != x.seq) {
            if (globaltimer_ns2() - t0 > x.timeout_ns) {
                s_timed_out = 1;
                *reinterpret_cast<volatile int*>(x.error) = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (s_timed_out) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            if (fin.out_keys) fin.out_keys[(size_t)qi * fin.out_stride + i] = 0ull;
            if (fin.out_scores) fin.out_scores[(size_t)qi * k + i] = -INFINITY;
            if (fin.out_ids) fin.out_ids[(size_t)qi * k + i] = -1;
        }
        __syncthreads();
        merge_signal_done(fin);
        xchg_publish_done(x);
        return;
    }
    MergeParams w = fin;
    w.keys = x.my_slots + (size_t)par * x.world * x.max_nq * slot_sz;
    w.nlists = x.world;
    w.k = k;
    w.stride_list = (int64_t)x.max_nq * slot_sz;
    w.stride_query = (int64_t)slot_sz;
    w.list_base = nullptr;     // already global rows
    merge_lists<KPL>(w, qi, qi, lists, nwarps);
    xchg_publish_done(x);
}

}  // namespace ts
