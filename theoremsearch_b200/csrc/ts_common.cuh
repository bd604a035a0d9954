// Shared device/host helpers for libtheoremsearch (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>

#include "../../include/theoremsearch.h"

namespace ts {

// ---------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define TS_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ts::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                          __FILE__, __LINE__);                                           \
            return (_e == cudaErrorMemoryAllocation) ? TS_ERR_OOM : TS_ERR_CUDA;         \
        }                                                                                \
    } while (0)

#define TS_REQUIRE(cond, code, ...)       \
    do {                                  \
        if (!(cond)) {                    \
            ts::set_error(__VA_ARGS__);   \
            return (code);                \
        }                                 \
    } while (0)

// Every kernel launch goes through this so ts_kernel_launches() is an honest count.
#define TS_LAUNCH_CHECK()                                   \
    do {                                                    \
        ts::g_launches.fetch_add(1, std::memory_order_relaxed); \
        TS_CHECK_CUDA(cudaGetLastError());                  \
    } while (0)

inline int sm_count(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

// selects a CUDA device for the lifetime of the object, restores the previous one
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            ok = false;
            return;
        }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------- tunables
struct Tunables {
    int scan_ctas_per_sm = 1;   // persistent CTAs per SM for K2
    int scan_warps = 8;         // consumer warps per CTA
    int scan_stages = 2;        // TMA ring depth per warp (2 measured best on B200: sweep in profiles/)
    int scan_tile_rows = 0;     // rows per TMA bulk copy; 0 = default (~8 KB tiles)
    int batch_min_nq = 2;       // nq >= this goes to the batched tcgen05 path (bf16 corpus): one pass at ~3.5 ms beats nq scans of 2.8 ms
    int batch_cap = 3072;       // K3 candidate slots per query per chunk
    int batch_first_chunk = 1024;  // rows of the first K3 chunk (every row passes thr = -inf)
    int batch_growth = 3;       // next chunk = growth x rows already seen
    int batch_dense = 1;        // small corpora (nq * N * 4 B <= 1 GiB): dense score matrix + select instead of chunked filtering
    int batch_a_policy = 2;     // K3 corpus-tile L2 policy: 0 evict_first, 1 evict_normal, 2 evict_last. With evict_first the 16 n-blocks
                                // of a corpus tile re-fetched it from DRAM (ncu: 15.7 GB read for a 7.4 GB chunk, 26 GB with CTA pairs) and
                                // DRAM traffic costs power under the cap: 71.9 -> 70.8 ms (single CTAs), 76.0 -> 70.2 ms (pairs)
    int batch_tf32 = 1;         // fp32-stored corpora: batches take the TF32 GEMM (0 = one K2 pass per query, as in round 1)
    int batch_cta_pair = 1;     // large batches use the cta_group::2 kernel (CTA pairs: 256-row tiles, each CTA stages half of the
                                // query block). Round 2, 4096 x 10M x 1024, 10 iterations, power-capped: pairs 70.2 ms vs single CTAs
                                // 70.8 ms with the corpus tiles kept in L2 (profiles/k3_sweep_r2b.txt); ncu on the large chunk: 33.1 vs 35.4 ms
    int batch_pair_min_nq = 512;  // batches at least this large use CTA pairs
    int ivf_warps = 0;          // K4b warps per CTA; 0 = auto (latency mode 16, throughput mode 8)
    int ivf_tile_rows = 0;      // K4b rows per TMA tile (4 or 8, e4m3 D<=1024 lists only); 0 = auto
    int ivf_parts = 0;          // K4b CTAs per query; 0 = auto (2*SMs/nq clamped to [1, SMs])
    int ivf_timeline = 0;       // 1 = K4b CTAs record globaltimer stamps per phase (ts_debug_ivf_timeline)
    int ivf_group_min_nq = 16;  // batches at least this large take the list-major scan (K4d); 0 = never (sweep: profiles/sweep_ivf_batch_r1.txt)
    int ivf_group_min_lists = 0;  // K4d needs at least this many lists (and query-list pairs); 0 = 4 x SM count
    int ivf_fuse_rescore = 1;   // small IVF batches: the list scan's last CTA re-scores its candidates (no separate K4c launch)
    int ivf_select_warp = 1;    // K4d candidate selection: 1 = warp-per-query register select, 0 = CTA-per-query smem select
    int scan_timeline = 0;      // 1 = K2 CTAs record %globaltimer stamps per phase (ts_debug_scan_timeline)
    int store_no_vmm = 0;       // 1 = plain cudaMalloc row stores (copy on grow) instead of virtual-memory mapping
    int xchg_debug_no_flag = 0; // test hook: the sharded exchange does not raise its own flag, so its wait times out
    int ivf_group_mma = 3;      // K4d scoring: 3 / 4 = tcgen05 kind::f8f6f4 (e4m3 rows straight into the tensor cores, two-term
                                // e4m3 queries) with 16 / 8 queries per group; 1 / 2 = legacy mma.sync f16 variant with 8 / 16;
                                // 0 = packed HFMA2 on the CUDA cores (4 queries per group)
};
Tunables& tunables();

// ---------------------------------------------------------------------------- keys
// key = orderable_u32(score) << 32 | (0xFFFFFFFF - row). Unsigned max == (score desc, row asc).
// 0 is reserved for "empty".
__host__ __device__ __forceinline__ uint32_t orderable_from_float(float f) {
#ifdef __CUDA_ARCH__
    if (f != f) return 1u;
    f = f + 0.0f;  // -0.0 -> +0.0
    uint32_t u = __float_as_uint(f);
#else
    if (f != f) return 1u;
    f = f + 0.0f;
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    uint32_t hi = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return hi ? hi : 1u;
}
__host__ __device__ __forceinline__ float float_from_orderable(uint32_t hi) {
    uint32_t u = (hi & 0x80000000u) ? (hi & 0x7FFFFFFFu) : ~hi;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float score, uint32_t row) {
    return ((uint64_t)orderable_from_float(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) {
    return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
    return float_from_orderable((uint32_t)(key >> 32));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------- warp top-k list
// A descending list of KPL*32 keys held across one warp: position p = j*32 + lane lives in
// slot j of lane `lane`. Slot values are unique keys or 0 (empty, sorts last).
__host__ __device__ constexpr int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

template <int KPL>
struct WarpTopK {
    uint64_t key[KPL];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int j = 0; j < KPL; ++j) key[j] = 0ull;
    }

    // A key that nothing in the top k can be below, broadcast to every lane: the k-th key for
    // single-register lists, the LAST (32*KPL-th) key otherwise. The weaker bound keeps the register
    // index static — `key[(k-1) >> 5]` with a run-time k makes nvcc index the array dynamically, which
    // moves the whole list to local memory — and only admits a few more candidates (K ln(n/K) instead of
    // k ln(n/k) over n rows).
    __device__ __forceinline__ uint64_t kth(int k) const {
        if constexpr (KPL == 1) return __shfl_sync(0xFFFFFFFFu, key[0], k - 1);
        else return __shfl_sync(0xFFFFFFFFu, key[KPL - 1], 31);
    }

    // Insert x (warp-uniform, x != any present key). Entries below x shift down one
    // position; the last one falls off.
    __device__ __forceinline__ void insert(uint64_t x, int lane) {
        uint64_t carry = ~0ull;  // "previous position" of position 0: +inf, never < x
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            uint64_t up = __shfl_up_sync(0xFFFFFFFFu, key[j], 1);
            uint64_t last = __shfl_sync(0xFFFFFFFFu, key[j], 31);
            uint64_t prev = (lane == 0) ? carry : up;
            uint64_t mine = key[j];
            key[j] = (mine > x) ? mine : ((prev > x) ? x : prev);
            carry = last;
        }
    }

    // Merge another descending list held in the same register layout (b[j] = B[j*32 + lane], 0-padded)
    // and keep the best 32*KPL of the union, sorted descending. Bitonic top-K merge: position p takes
    // max(A[p], B[K-1-p]) — a bitonic sequence holding the K largest — then log2(K) compare-exchange
    // stages (register-to-register for strides >= 32 positions, shuffles below). ~100 instructions for
    // K = 128 where element-wise insertion costs ~60 cycles per entering key.
    __device__ __forceinline__ void merge_desc(const uint64_t (&b)[KPL], int lane) {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const uint64_t rb = __shfl_sync(0xFFFFFFFFu, b[KPL - 1 - j], 31 - lane);
            key[j] = key[j] > rb ? key[j] : rb;
        }
        // (loops run over log2 of the stride with unit steps so that nvcc fully unrolls them and every
        // register index is a compile-time constant — a shift-stepped loop is left rolled and drags the
        // whole list into local memory)
#pragma unroll
        for (int ls = ilog2_c(KPL) - 1; ls >= 0; --ls) {
            const int s = 1 << ls;
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                if ((j & s) == 0) {
                    const uint64_t x = key[j], y = key[j | s];
                    key[j] = x > y ? x : y;
                    key[j | s] = x > y ? y : x;
                }
            }
        }
#pragma unroll
        for (int ls = 4; ls >= 0; --ls) {
            const int s = 1 << ls;
            const bool lower = (lane & s) != 0;   // the half of each pair that keeps the smaller key
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, key[j], s);
                const bool take = lower ? (o < key[j]) : (o > key[j]);
                key[j] = take ? o : key[j];
            }
        }
    }
};

// Fold the descending-sorted list src[0..n) (n <= 32*KPL, unique keys, 0 = empty) into `list`.
// All lanes call with the same args. Skipped outright when src[0] cannot enter the top k.
template <int KPL>
__device__ __forceinline__ void merge_sorted_into(WarpTopK<KPL>& list, const uint64_t* src, int n,
                                                  int k, int lane) {
    if (n <= 0 || src[0] <= list.kth(k)) return;   // sorted: nothing after src[0] can enter either
    if constexpr (KPL <= 8) {
        uint64_t b[KPL];
#pragma unroll
        for (int j = 0; j < KPL; ++j) b[j] = (j * 32 + lane < n) ? src[j * 32 + lane] : 0ull;
        list.merge_desc(b, lane);
    } else {
        // k > 256 (rare): element-wise insertion — the 1024-key bitonic network unrolls to tens of thousands
        // of instructions per call site and dominated the build time
        uint64_t thr = list.kth(k);
        for (int i = 0; i < n; ++i) {
            const uint64_t x = src[i];
            if (x <= thr) break;
            list.insert(x, lane);
            thr = list.kth(k);
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---------------------------------------------------------------------------- mbarrier / TMA (1-D bulk)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (TMA engine, no tensor map): bytes % 16 == 0, both 16B aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src,
                                                 uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif  // __CUDACC__

}  // namespace ts

// ---------------------------------------------------------------------------- handles
namespace ts {
struct RowStore;
}
struct ts_index {
    int device = 0;
    int dim = 0;          // logical embedding dimension
    int dim_pad = 0;      // stored elements per row (multiple of 8 -> 16-byte rows)
    int dtype = TS_BF16;  // storage dtype of `data`
    int64_t capacity = 0;
    int64_t size = 0;
    void* data = nullptr;     // [capacity, dim_pad] of dtype (== row_store_ptr(store))
    ts::RowStore* store = nullptr;   // owns `data`: a reserved address range, physical memory mapped as the index grows
    ts::RowStore* ids_store = nullptr;   // owns `ids`
    int64_t auto_next = 0;           // id of the next row added WITHOUT an id: one past the largest row position ever
                                     // occupied (a SERIAL column: positions shrink on delete, this never does)
    ts::RowStore* pos_store = nullptr;   // owns `pos_of_row`
    int64_t* ids = nullptr;   // [capacity] caller ids; valid only when has_ids
    bool has_ids = false;
    float* max_norm2 = nullptr;  // device scalar: max squared L2 norm over stored (quantised) rows

    // IVF-Flat state (K4)
    int nlist = 0;
    float* centroids = nullptr;         // [nlist, dim_pad] fp32 unit vectors
    __nv_bfloat16* centroids_bf16 = nullptr;  // same, bf16, what the coarse scan reads
    int64_t* list_offsets = nullptr;    // [nlist + 1] into the permuted row order
    uint32_t* list_rows = nullptr;      // [size] original row of each permuted position
    void* list_data = nullptr;          // [size, dim_pad] permuted rows in list_dtype
    float* list_scales = nullptr;       // [size] per-row scale (fp8 only)
    int list_dtype = TS_BF16;
    uint32_t list_row_bytes = 0;        // bytes per row of list_data (multiple of 16)
    float* centroid_max_norm2 = nullptr;  // device scalar: max ||centroid_bf16||^2
    bool ivf_built = false;
    // Incremental IVF state (pgvector's ivfflat accepts inserts after the build; so does this):
    // the list arrays hold `built_n` positions of packed main lists followed by `ovf_n` positions of OVERFLOW
    // lists — rows added or replaced since the build, filed under the same centroids and packed the same way.
    // Overflow list l is virtual list nlist + l: list_offsets has 2*nlist + 1 entries and simply continues, so
    // the scan kernels see one more list per probe and need no other change. A replaced row's old position is
    // tombstoned (list_rows[pos] = TS_DEAD_ROW) and its new content filed in the overflow.
    int64_t list_cap = 0;            // rows the list arrays (list_rows / list_data / list_scales) can hold
    int64_t built_n = 0;
    int64_t ovf_n = 0;
    int64_t ivf_moved = 0;           // deletes since the build: list entries were renamed in place (rows no longer ascend)
    int64_t ivf_dead = 0;            // tombstoned positions among the main lists
    uint32_t* pos_of_row = nullptr;  // [capacity] list position of every corpus row
    uint32_t* ovf_set = nullptr;     // [ovf_cap] corpus rows filed in the overflow lists, ascending
    int64_t ovf_cap = 0;
    // id -> row (host side; upserts arrive from the host writer, ec2/generate_embeddings/__main__.py:84-101)
    std::unordered_map<int64_t, int64_t>* id_map_host = nullptr;
    bool id_map_valid = false;

    size_t elem_bytes() const { return dtype == TS_F32 ? 4 : 2; }
    size_t row_bytes() const { return (size_t)dim_pad * elem_bytes(); }
};

// Peer exchange of the sharded exact search: every rank owns a receive area other ranks store into over
// NVLink (CUDA IPC mappings). Layout of one area: slots[2 parities][world][max_nq][max_k] packed keys and
// flags[2][world][max_nq] sequence numbers.
struct ts_xchg {
    int device = 0, world = 1, rank = 0, max_nq = 0, max_k = 0;
    uint64_t* slots = nullptr;        // this rank's receive area (cudaMalloc, exported by IPC handle)
    uint32_t* flags = nullptr;        // same allocation, after the slots
    void* base = nullptr;             // the allocation
    size_t bytes = 0;
    void* peer_base[16] = {nullptr};  // mapped base of every rank's area (own entry = base)
    uint64_t** d_peer_slots = nullptr;   // device array [world]
    uint32_t** d_peer_flags = nullptr;   // device array [world]
    uint32_t* d_tickets = nullptr;    // [max_nq] last-CTA tickets: zero at creation, left zero by every kernel (no per-call memset)
    int* h_error = nullptr;           // sticky flag in pinned, device-mapped host memory: 1 = a peer did not arrive
                                      // within the time-out; the host reads it without synchronising
    int* d_error = nullptr;           // the device-side address of h_error
    uint32_t* d_done = nullptr;       // [2]: sequence number of the last finished exchange, per-search query counter
    uint64_t* part_ring = nullptr;    // [4][max_nq][ring_nparts][max_k] per-CTA lists of the last four searches
    int ring_nparts = 0;
    uint32_t seq = 0;                 // searches issued so far (identical on every rank)
    unsigned long long timeout_ns = 10000000000ull;   // 10 s: covers lazy module loads and per-rank host skew
    bool connected = false;
    // host-buffer path (ts_search_sharded_host): stream, pinned staging, device buffers, workspace
    cudaStream_t stream = nullptr;
    float* h_queries = nullptr;
    float* h_scores = nullptr;
    int64_t* h_ids = nullptr;
    float* d_queries = nullptr;
    float* d_scores = nullptr;
    int64_t* d_ids = nullptr;
    void* workspace = nullptr;
    size_t workspace_bytes = 0;
    int host_dim = 0;
    float* m_scores = nullptr;        // device addresses of the mapped h_scores / h_ids / h_done (see ts_ctx)
    int64_t* m_ids = nullptr;
    uint32_t* h_done = nullptr;
    uint32_t* m_done = nullptr;
    uint32_t done_seq = 0;
};

struct ts_ctx {
    ts_index* index = nullptr;
    int max_nq = 0, max_k = 0;
    cudaStream_t stream = nullptr;
    float* h_queries = nullptr;   // pinned [max_nq, dim]
    float* h_scores = nullptr;    // pinned [max_nq, max_k]
    int64_t* h_ids = nullptr;     // pinned [max_nq, max_k]
    float* d_queries = nullptr;
    float* d_scores = nullptr;
    int64_t* d_ids = nullptr;
    void* workspace = nullptr;
    size_t workspace_bytes = 0;
    void* ivf_workspace = nullptr;      // grown on demand by ts_ivf_search_host (outside the timed path after warm-up)
    size_t ivf_workspace_bytes = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = -1.f;
    // single-query latency path: the final merge writes (scores, ids) straight into pinned, device-mapped host
    // memory and raises a mapped flag the host polls (no D2H copies, no stream synchronise)
    float* m_scores = nullptr;        // device address of h_scores
    int64_t* m_ids = nullptr;         // device address of h_ids
    uint32_t* h_done = nullptr;       // mapped flag (host address) and its device address
    uint32_t* m_done = nullptr;
    uint32_t done_seq = 0;
};

#define TS_DEAD_ROW 0xFFFFFFFFu

// internal kernel entry points (one per .cu)
namespace ts {
// growable row store (vmm_store.cu): virtual-memory mapped, grows in place
struct RowStore;
int row_store_create(RowStore** out, int device, size_t bytes, size_t row_bytes);
void row_store_destroy(RowStore* st);
void* row_store_ptr(const RowStore* st);
bool row_store_is_vmm(const RowStore* st);
int row_store_reserve(RowStore* st, size_t bytes, size_t used);
// the side tables that grow with the index (caller ids, IVF row positions) live in row stores of their own
template <typename T>
inline int side_table_create(RowStore** st, T** table, int device, int64_t capacity) {
    int rc = row_store_create(st, device, (size_t)(capacity > 0 ? capacity : 1) * sizeof(T), sizeof(T));
    if (rc == 0) *table = static_cast<T*>(row_store_ptr(*st));
    return rc;
}
template <typename T>
inline int side_table_reserve(RowStore* st, T** table, int64_t used, int64_t capacity) {
    int rc = row_store_reserve(st, (size_t)capacity * sizeof(T), (size_t)used * sizeof(T));
    if (rc == 0) *table = static_cast<T*>(row_store_ptr(st));
    return rc;
}
template <typename T>
inline void side_table_destroy(RowStore** st, T** table) {
    row_store_destroy(*st);
    *st = nullptr;
    *table = nullptr;
}
// IVF upkeep after rows were replaced in place (device list of corpus rows) and/or appended ([app_first,
// app_first + app_n)): tombstones, overflow lists, automatic re-pack (k4_ivf.cu). No-op unless lists are built.
int ivf_apply_mutation(ts_index* ix, const uint32_t* replaced_rows, int64_t n_replaced, int64_t app_first,
                       int64_t app_n, cudaStream_t s);
// after ts_index_delete compacted the rows: remap[old row] = new row or TS_DEAD_ROW (device, one entry per old row)
int ivf_apply_delete(ts_index* ix, const uint32_t* remap, int64_t n_deleted, cudaStream_t s);
void ivf_free_all(ts_index* ix);
// grow the row capacity (data, ids, pos_of_row); synchronises the device
int index_reserve(ts_index* ix, int64_t capacity);
int index_make_room(ts_index* ix, int64_t extra);
// exact search over `ix` (K2 or K3 by batch size); exactly one of out_keys / (out_scores, out_ids) is used
int search_impl(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int normalize_queries,
                const uint32_t* allow_mask, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                void* workspace, size_t workspace_bytes, cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1, uint32_t* done_flag = nullptr,
                uint32_t done_value = 0, float* q_out = nullptr);
int launch_normalize_cast(const void* src, int src_dtype, int64_t n, int dim, int dim_pad,
                          int normalize, void* dst, int dst_dtype, cudaStream_t s,
                          float* max_norm2 = nullptr, const int64_t* dst_rows = nullptr);
// K4d: list-major batched IVF scan (k4_ivf_grouped.cu)
bool ivf_grouped_supported(const ts_index* ix, int kc);
size_t ivf_grouped_workspace_bytes(const ts_index* ix, int nq, int nprobe);
int launch_ivf_grouped(const ts_index* ix, const uint64_t* probes, const float* q32, int nq, int nprobe_real, int vmul, int kc,
                       const uint32_t* allow_mask, void* workspace, uint64_t* cand, const uint32_t** flag_out,
                       cudaStream_t s);
int launch_max_norm2(const void* rows, int dtype, int64_t n, int dim_pad, float* out, cudaStream_t s);
int launch_dequant_rows(const void* src, int src_dtype, int64_t n, int dim, int dim_pad, float* dst,
                        cudaStream_t s);
int launch_prepare_queries(const void* q, int q_dtype, int nq, int dim, int dim_pad, int normalize,
                           float* out_f32, cudaStream_t s);
// K2: per-CTA candidate lists part_keys[nq][nparts][k]
int scan_nparts(const ts_index* ix);
int debug_scan_timeline(uint64_t* out_host, int launches_back, int n_ctas);
// Optional fused stages of K2: in-kernel query normalisation and in-kernel final merge.
// device view of a ts_xchg for one search (world == 0: no exchange)
struct XchgDev {
    uint64_t* const* peer_slots;   // [world] every rank's slot area (mapped)
    uint32_t* const* peer_flags;   // [world] every rank's flag area (mapped)
    uint64_t* my_slots;
    uint32_t* my_flags;
    int* error;                    // sticky time-out flag in host-mapped memory
    uint32_t* done_seq;            // sequence number of this rank's last finished exchange
    uint32_t* done_count;          // queries of the current search finished so far (nq > 1)
    int world, rank, max_nq, max_k, nq;
    uint32_t seq;
    int64_t base;                  // global row of this shard's row 0
    unsigned long long timeout_ns; // bound on the wait for the peers' flags
    int debug_no_flag;             // test hook ("xchg.debug_no_flag"): do not raise the own flag -> the wait times out
};
struct ScanFused {
    const void* q_raw;      // caller's queries [nq, dim] (nullptr: read prepared fp32 queries instead)
    int q_dtype, q_normalize;
    uint32_t* tickets;      // [nq] zeroed counters (nullptr: no fused merge, only part_keys are written)
    const int64_t* id_map;
    uint64_t* out_keys;     // any of the three may be nullptr
    float* out_scores;
    int64_t* out_ids;
    XchgDev xchg;           // world > 0: exchange the shard results with the peers inside the kernel
    int pdl;                // ScanParams::pdl
    const uint32_t* ring_gate;   // ScanParams::ring_gate / ring_need
    uint32_t ring_need;
    uint32_t* done_flag = nullptr;   // MergeParams::done_flag / done_value (host-mapped completion flag)
    uint32_t done_value = 0;
    float* q_out = nullptr;          // ScanParams::q_out
};
int launch_scan_topk(const ts_index* ix, const void* data, int data_dtype, int64_t n_rows,
                     const float* queries_f32, int nq, int k, const uint32_t* allow_mask,
                     uint64_t* part_keys, int nparts, cudaStream_t s, cudaEvent_t ev0,
                     cudaEvent_t ev1, const int* qlist = nullptr, const int* qcount = nullptr,
                     const ScanFused* fused = nullptr);
// exchange kernel of the two-kernel sharded search (k5_merge.cu): merges the scan's per-CTA lists, exchanges the
// shard's keys with the peers, merges the world lists. Launched with programmatic stream serialisation.
int launch_xchg_finish(const uint64_t* part_keys, int nparts, int nq, int k, const XchgDev& x, const int64_t* id_map,
                       float* out_scores, int64_t* out_ids, cudaStream_t s, uint32_t* done_flag = nullptr,
                       uint32_t done_value = 0);
// host-side wait for a device-written completion flag in mapped memory (falls back to a stream synchronise)
int wait_done_flag(const volatile uint32_t* flag, uint32_t value, cudaStream_t s, const char* what);
// K5: lists[nlists][nq][k] (list-major) or [nq][nlists][k] (query-major) -> top-k
int launch_merge(const uint64_t* keys, int nlists, int nq, int k, bool query_major,
                 const int64_t* list_base, const int64_t* id_map, uint64_t* out_keys,
                 float* out_scores, int64_t* out_ids, cudaStream_t s);
int launch_merge_strided(const uint64_t* keys, int nlists, int nq, int k, int64_t stride_list,
                         int64_t stride_query, const int64_t* list_base, const int64_t* id_map,
                         uint64_t* out_keys, float* out_scores, int64_t* out_ids, cudaStream_t s,
                         const int* qlist = nullptr, const int* qcount = nullptr, int64_t out_stride = 0);
// K3: batched tcgen05 GEMM + fused top-k
size_t batched_workspace_bytes(const ts_index* ix, int nq, int k);
int debug_last_batched_fixups();
int launch_batched_search(const ts_index* ix, const void* queries, int q_dtype, int nq, int k, int normalize,
                          const uint32_t* allow_mask, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                          void* workspace, size_t workspace_bytes, cudaStream_t s, cudaEvent_t ev0,
                          cudaEvent_t ev1);
}  // namespace ts
