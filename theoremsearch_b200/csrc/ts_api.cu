// C ABI of libtheoremsearch.so — see include/theoremsearch.h for the contract and the
// reference call sites each entry point replaces.
#include "ts_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

namespace ts {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

Tunables& tunables() {
    static Tunables t;
    return t;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Poll a completion flag the device writes into mapped host memory. The result bytes were written (and fenced at
// system scope) before the flag, so once the flag shows `value` the outputs beside it are final. After ~200 ms of
// polling — a kernel that faulted never raises the flag — fall back to a stream synchronise, which reports the error.
int wait_done_flag(const volatile uint32_t* flag, uint32_t value, cudaStream_t s, const char* what) {
    for (uint64_t spins = 0; *flag != value; ++spins) {
#if defined(__x86_64__) || defined(_M_X64)
        __builtin_ia32_pause();
#endif
        if ((spins & 0xFFFFu) == 0xFFFFu) {
            const cudaError_t q = cudaStreamQuery(s);
            if (q == cudaSuccess) break;                       // the stream has drained: the flag write is done too
            if (q != cudaErrorNotReady) {
                set_error("%s: kernel failed: %s", what, cudaGetErrorString(q));
                return TS_ERR_CUDA;
            }
        }
    }
    return TS_OK;
}

// Workspace layout for exact search: [prepared queries fp32 | per-CTA candidate keys]
struct SearchWs {
    float* q_f32;
    uint64_t* part_keys;
    uint32_t* tickets;
    int nparts;
    size_t bytes;
};
static SearchWs carve_ws(const ts_index* ix, int nq, int k, void* base) {
    SearchWs w;
    w.nparts = scan_nparts(ix);
    size_t off = 0;
    w.q_f32 = (float*)((char*)base + off);
    off += align_up((size_t)nq * ix->dim_pad * sizeof(float), 256);
    w.part_keys = (uint64_t*)((char*)base + off);
    off += align_up((size_t)nq * w.nparts * k * sizeof(uint64_t), 256);
    w.tickets = (uint32_t*)((char*)base + off);
    off += align_up((size_t)nq * sizeof(uint32_t), 256);
    w.bytes = off;
    return w;
}

// nq >= batch.min_nq goes to K3 (tcgen05 GEMM: kind::f16 on bf16 rows, kind::tf32 on fp32 rows); a single query to K2.
static bool use_batched(const ts_index* ix, int nq) {
    if (ix->dtype == TS_F32 && !tunables().batch_tf32) return false;
    return (ix->dtype == TS_BF16 || ix->dtype == TS_F32) && nq >= tunables().batch_min_nq && ix->size > 0;
}

int search_impl(ts_index* ix, const void* queries, int q_dtype, int nq, int k,
                       int normalize_queries, const uint32_t* allow_mask, uint64_t* out_keys,
                       float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                       cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1, uint32_t* done_flag, uint32_t done_value,
                       float* q_out) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "search: index is NULL");
    TS_REQUIRE(nq >= 0, TS_ERR_BAD_ARG, "search: nq=%d", nq);
    TS_REQUIRE(k >= 1 && k <= TS_MAX_K, TS_ERR_BAD_ARG, "search: k=%d out of range [1, %d]", k, TS_MAX_K);
    TS_REQUIRE(q_dtype == TS_F32 || q_dtype == TS_BF16 || q_dtype == TS_F16, TS_ERR_BAD_ARG,
               "search: query dtype %d", q_dtype);
    if (nq == 0) return TS_OK;
    TS_REQUIRE(queries != nullptr, TS_ERR_BAD_ARG, "search: queries is NULL");
    TS_REQUIRE(workspace != nullptr, TS_ERR_BAD_ARG, "search: workspace is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "search: cannot select CUDA device %d", ix->device);
    if (use_batched(ix, nq)) {
        if (q_out != nullptr) {   // the caller wants the prepared queries too (IVF): K3 keeps its own copy internally
            int prc = launch_prepare_queries(queries, q_dtype, nq, ix->dim, ix->dim_pad, normalize_queries, q_out, s);
            if (prc) return prc;
        }
        return launch_batched_search(ix, queries, q_dtype, nq, k, normalize_queries, allow_mask, out_keys,
                                     out_scores, out_ids, workspace, workspace_bytes, s, ev0, ev1);
    }
    SearchWs w = carve_ws(ix, nq, k, workspace);
    TS_REQUIRE(workspace_bytes >= w.bytes, TS_ERR_CAPACITY, "search: workspace %zu < %zu bytes",
               workspace_bytes, w.bytes);
    // One kernel does everything: each warp normalises the query itself, the scan keeps per-warp
    // top-k lists, and the last CTA to finish merges the 148 per-CTA lists and maps rows to ids.
    // The only other stream operation is zeroing the nq ticket counters.
    TS_CHECK_CUDA(cudaMemsetAsync(w.tickets, 0, (size_t)nq * sizeof(uint32_t), s));
    ScanFused f;
    f.q_raw = queries;
    f.q_dtype = q_dtype;
    f.q_normalize = normalize_queries;
    f.tickets = w.tickets;
    f.id_map = ix->has_ids ? ix->ids : nullptr;
    f.out_keys = out_keys;
    f.out_scores = out_scores;
    f.out_ids = out_ids;
    memset(&f.xchg, 0, sizeof(f.xchg));
    f.pdl = 0;
    f.ring_gate = nullptr;
    f.ring_need = 0;
    f.done_flag = done_flag;     // only the single-query host path passes one (nq == 1: one merging CTA)
    f.done_value = done_value;
    f.q_out = q_out;
    return launch_scan_topk(ix, ix->data, ix->dtype, ix->size, w.q_f32, nq, k, allow_mask, w.part_keys, w.nparts,
                            s, ev0, ev1, nullptr, nullptr, &f);
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_abi_version(void) { return 2; }

const char* ts_last_error(void) { return t_error.c_str(); }

uint64_t ts_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

uint64_t ts_pack_key(float score, uint32_t row) { return pack_key(score, row); }

void ts_unpack_key(uint64_t key, float* score, uint32_t* row) {
    if (score) *score = key ? key_score(key) : -INFINITY;
    if (row) *row = key_row(key);
}

// ------------------------------------------------------------------------------------ index
int ts_index_create(ts_index** out, int device, int dim, int dtype, int64_t capacity) {
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "index_create: out is NULL");
    *out = nullptr;
    TS_REQUIRE(dim >= 1 && dim <= TS_MAX_DIM, TS_ERR_BAD_ARG, "index_create: dim=%d out of range [1, %d]",
               dim, TS_MAX_DIM);
    TS_REQUIRE(dtype == TS_BF16 || dtype == TS_F32, TS_ERR_BAD_ARG,
               "index_create: storage dtype must be TS_BF16 or TS_F32 (got %d)", dtype);
    TS_REQUIRE(capacity >= 0 && capacity < (int64_t)0xFFFFFFFFll, TS_ERR_BAD_ARG,
               "index_create: capacity=%lld", (long long)capacity);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("index_create: no CUDA device (%s); libtheoremsearch has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return TS_ERR_CUDA;
    }
    TS_REQUIRE(device >= 0 && device < ndev, TS_ERR_BAD_ARG, "index_create: device %d of %d", device, ndev);
    DeviceGuard g(device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_create: cannot select CUDA device %d", device);
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    TS_REQUIRE(major == 10, TS_ERR_UNSUPPORTED,
               "index_create: device %d is sm_%d0; this library is built for sm_100a only", device, major);
    ts_index* ix = new ts_index();
    ix->device = device;
    ix->dim = dim;
    ix->dim_pad = (dim + 7) / 8 * 8;
    ix->dtype = dtype;
    ix->capacity = capacity;
    size_t bytes = (size_t)std::max<int64_t>(capacity, 1) * ix->row_bytes();
    {
        int src = row_store_create(&ix->store, device, bytes, ix->row_bytes());
        if (src != TS_OK) {
            delete ix;
            return src;
        }
        ix->data = row_store_ptr(ix->store);
    }
    e = cudaMalloc(&ix->max_norm2, sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(ix->max_norm2, 0, sizeof(float));
    if (e != cudaSuccess) {
        set_error("index_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        row_store_destroy(ix->store);
        delete ix;
        cudaGetLastError();
        return TS_ERR_OOM;
    }
    *out = ix;
    return TS_OK;
}

void ts_index_destroy(ts_index* ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    row_store_destroy(ix->store);
    cudaFree(ix->max_norm2);
    side_table_destroy(&ix->ids_store, &ix->ids);
    cudaFree(ix->centroids);
    cudaFree(ix->centroids_bf16);
    cudaFree(ix->list_offsets);
    cudaFree(ix->list_rows);
    cudaFree(ix->list_data);
    cudaFree(ix->list_scales);
    cudaFree(ix->centroid_max_norm2);
    side_table_destroy(&ix->pos_store, &ix->pos_of_row);
    cudaFree(ix->ovf_set);
    delete ix->id_map_host;
    delete ix;
}

__global__ void iota_ids_kernel(int64_t* ids, int64_t first, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ids[i] = first + i;
}

int ts_index_add(ts_index* ix, const void* rows, int src_dtype, int64_t n, int normalize,
                 const int64_t* ids, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_add: index is NULL");
    TS_REQUIRE(n >= 0, TS_ERR_BAD_ARG, "index_add: n=%lld", (long long)n);
    if (n == 0) return TS_OK;
    TS_REQUIRE(rows != nullptr, TS_ERR_BAD_ARG, "index_add: rows is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_add: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (ix->size + n > ix->capacity) {   // the pgvector table this replaces has no fixed size (rds_schema.sql:50-53)
        int grc = index_make_room(ix, n);
        if (grc) return grc;
    }
    if (ids != nullptr && !ix->has_ids) {
        // first explicit ids: materialise the id table, rows added so far keep id = row
        int irc = side_table_create(&ix->ids_store, &ix->ids, ix->device, ix->capacity);
        if (irc) return irc;
        if (ix->size > 0) {
            iota_ids_kernel<<<256, 256, 0, s>>>(ix->ids, 0, ix->size);
            TS_LAUNCH_CHECK();
        }
        ix->has_ids = true;
    }
    if (ix->has_ids) {
        if (ids != nullptr) {
            TS_CHECK_CUDA(cudaMemcpyAsync(ix->ids + ix->size, ids, (size_t)n * sizeof(int64_t),
                                          cudaMemcpyDeviceToDevice, s));
        } else {
            // no ids given: the rows are numbered on from the largest position ever used (after a delete the row
            // positions shrink; ids are not handed out twice)
            const int64_t auto_first = std::max(ix->auto_next, ix->size);
            iota_ids_kernel<<<256, 256, 0, s>>>(ix->ids + ix->size, auto_first, n);
            TS_LAUNCH_CHECK();
            ix->auto_next = auto_first + n;
        }
    }
    void* dst = (char*)ix->data + (size_t)ix->size * ix->row_bytes();
    int rc = launch_normalize_cast(rows, src_dtype, n, ix->dim, ix->dim_pad, normalize, dst, ix->dtype, s,
                                   ix->max_norm2);
    if (rc) return rc;
    const int64_t first = ix->size;
    ix->size += n;
    ix->auto_next = std::max(ix->auto_next, ix->size);
    ix->id_map_valid = false;
    // built IVF lists stay valid: the new rows are filed in overflow lists (re-packed once they reach a tenth of the corpus)
    return ivf_apply_mutation(ix, nullptr, 0, first, n, s);
}

int ts_index_add_host(ts_index* ix, const void* rows, int src_dtype, int64_t n, int normalize,
                      const int64_t* ids) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_add_host: index is NULL");
    TS_REQUIRE(n >= 0, TS_ERR_BAD_ARG, "index_add_host: n=%lld", (long long)n);
    if (n == 0) return TS_OK;
    TS_REQUIRE(rows != nullptr, TS_ERR_BAD_ARG, "index_add_host: rows is NULL");
    TS_REQUIRE(src_dtype == TS_F32 || src_dtype == TS_BF16 || src_dtype == TS_F16, TS_ERR_BAD_ARG,
               "index_add_host: source dtype %d", src_dtype);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_add_host: cannot select CUDA device %d", ix->device);
    const size_t src_row = (size_t)ix->dim * (src_dtype == TS_F32 ? 4 : 2);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)((64u << 20) / src_row)));
    void* d_rows = nullptr;
    int64_t* d_ids = nullptr;
    TS_CHECK_CUDA(cudaMalloc(&d_rows, (size_t)chunk * src_row));
    if (ids) {
        cudaError_t e = cudaMalloc(&d_ids, (size_t)chunk * sizeof(int64_t));
        if (e != cudaSuccess) {
            cudaFree(d_rows);
            set_error("index_add_host: cudaMalloc ids failed: %s", cudaGetErrorString(e));
            return TS_ERR_OOM;
        }
    }
    int rc = TS_OK;
    for (int64_t pos = 0; pos < n && rc == TS_OK; pos += chunk) {
        const int64_t m = std::min(chunk, n - pos);
        cudaError_t e = cudaMemcpy(d_rows, (const char*)rows + (size_t)pos * src_row, (size_t)m * src_row,
                                   cudaMemcpyHostToDevice);
        if (e == cudaSuccess && ids)
            e = cudaMemcpy(d_ids, ids + pos, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("index_add_host: H2D copy failed: %s", cudaGetErrorString(e));
            rc = TS_ERR_CUDA;
            break;
        }
        rc = ts_index_add(ix, d_rows, src_dtype, m, normalize, ids ? d_ids : nullptr, nullptr);
        if (rc == TS_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) {
            set_error("index_add_host: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = TS_ERR_CUDA;
        }
    }
    cudaFree(d_rows);
    cudaFree(d_ids);
    return rc;
}

int64_t ts_index_size(const ts_index* ix) { return ix ? ix->size : -1; }
int64_t ts_index_capacity(const ts_index* ix) { return ix ? ix->capacity : -1; }
int ts_index_dim(const ts_index* ix) { return ix ? ix->dim : -1; }
int ts_index_dtype(const ts_index* ix) { return ix ? ix->dtype : -1; }
int ts_index_device(const ts_index* ix) { return ix ? ix->device : -1; }
const void* ts_index_data(const ts_index* ix) { return ix ? ix->data : nullptr; }
size_t ts_index_row_bytes(const ts_index* ix) { return ix ? ix->row_bytes() : 0; }
int ts_index_grows_in_place(const ts_index* ix) { return ix ? (row_store_is_vmm(ix->store) ? 1 : 0) : -1; }

int ts_index_get_rows(const ts_index* ix, int64_t first, int64_t n, float* out, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_get_rows: index is NULL");
    TS_REQUIRE(first >= 0 && n >= 0 && first + n <= ix->size, TS_ERR_BAD_ARG,
               "index_get_rows: [%lld, %lld) outside [0, %lld)", (long long)first, (long long)(first + n),
               (long long)ix->size);
    if (n == 0) return TS_OK;
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "index_get_rows: out is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_get_rows: cannot select CUDA device %d", ix->device);
    return launch_dequant_rows((const char*)ix->data + (size_t)first * ix->row_bytes(), ix->dtype, n, ix->dim,
                               ix->dim_pad, out, (cudaStream_t)stream);
}

int ts_ivf_search_host(ts_ctx* c, const float* queries, int nq, int k, int nprobe, int rescore_k, int normalize_queries,
                       const uint32_t* allow_mask, float* out_scores, int64_t* out_ids) {
    TS_REQUIRE(c != nullptr, TS_ERR_BAD_ARG, "ivf_search_host: ctx is NULL");
    TS_REQUIRE(nq >= 0 && nq <= c->max_nq, TS_ERR_CAPACITY, "ivf_search_host: nq=%d exceeds ctx max_nq=%d", nq, c->max_nq);
    TS_REQUIRE(k >= 1 && k <= c->max_k, TS_ERR_CAPACITY, "ivf_search_host: k=%d exceeds ctx max_k=%d", k, c->max_k);
    if (nq == 0) return TS_OK;
    TS_REQUIRE(queries && out_scores && out_ids, TS_ERR_BAD_ARG, "ivf_search_host: NULL buffer");
    ts_index* ix = c->index;
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ivf_search_host: cannot select CUDA device %d", ix->device);
    const size_t need = ts_ivf_workspace_bytes(ix, nq, k, nprobe, rescore_k);
    TS_REQUIRE(need > 0, TS_ERR_STATE, "ivf_search_host: the index has no IVF lists (ts_ivf_train + ts_ivf_build)");
    if (need > c->ivf_workspace_bytes) {   // first call with this shape only
        TS_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->ivf_workspace);
        c->ivf_workspace = nullptr;
        c->ivf_workspace_bytes = 0;
        TS_CHECK_CUDA(cudaMalloc(&c->ivf_workspace, need));
        c->ivf_workspace_bytes = need;
    }
    const size_t qb = (size_t)nq * ix->dim * sizeof(float);
    memcpy(c->h_queries, queries, qb);
    TS_CHECK_CUDA(cudaMemcpyAsync(c->d_queries, c->h_queries, qb, cudaMemcpyHostToDevice, c->stream));
    int rc = ts_ivf_search(ix, c->d_queries, TS_F32, nq, k, nprobe, rescore_k, normalize_queries, allow_mask, c->d_scores,
                           c->d_ids, c->ivf_workspace, c->ivf_workspace_bytes, c->stream);
    if (rc) return rc;
    const size_t n = (size_t)nq * k;
    TS_CHECK_CUDA(cudaMemcpyAsync(c->h_scores, c->d_scores, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TS_CHECK_CUDA(cudaMemcpyAsync(c->h_ids, c->d_ids, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    TS_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(out_scores, c->h_scores, n * sizeof(float));
    memcpy(out_ids, c->h_ids, n * sizeof(int64_t));
    return TS_OK;
}

// ------------------------------------------------------------------------------------ raw access (save / load)
int ts_index_has_ids(const ts_index* ix) { return ix ? (ix->has_ids ? 1 : 0) : -1; }

int ts_index_read_raw_host(const ts_index* ix, int64_t first, int64_t n, void* rows_out, int64_t* ids_out) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_read_raw_host: index is NULL");
    TS_REQUIRE(first >= 0 && n >= 0 && first + n <= ix->size, TS_ERR_BAD_ARG,
               "index_read_raw_host: [%lld, %lld) outside [0, %lld)", (long long)first, (long long)(first + n),
               (long long)ix->size);
    if (n == 0) return TS_OK;
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_read_raw_host: cannot select CUDA device %d", ix->device);
    if (rows_out)
        TS_CHECK_CUDA(cudaMemcpy(rows_out, (const char*)ix->data + (size_t)first * ix->row_bytes(),
                                 (size_t)n * ix->row_bytes(), cudaMemcpyDeviceToHost));
    if (ids_out) {
        TS_REQUIRE(ix->has_ids, TS_ERR_STATE, "index_read_raw_host: the index has no id table (ids are row positions)");
        TS_CHECK_CUDA(cudaMemcpy(ids_out, ix->ids + first, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    return TS_OK;
}

int ts_index_append_raw_host(ts_index* ix, const void* rows, int64_t n, const int64_t* ids) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_append_raw_host: index is NULL");
    TS_REQUIRE(n >= 0, TS_ERR_BAD_ARG, "index_append_raw_host: n=%lld", (long long)n);
    if (n == 0) return TS_OK;
    TS_REQUIRE(rows != nullptr, TS_ERR_BAD_ARG, "index_append_raw_host: rows is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_append_raw_host: cannot select CUDA device %d", ix->device);
    if (ix->size + n > ix->capacity) {
        int grc = index_make_room(ix, n);
        if (grc) return grc;
    }
    if (ids != nullptr && !ix->has_ids) {
        int irc = side_table_create(&ix->ids_store, &ix->ids, ix->device, ix->capacity);
        if (irc) return irc;
        if (ix->size > 0) {
            iota_ids_kernel<<<256, 256>>>(ix->ids, 0, ix->size);
            TS_LAUNCH_CHECK();
        }
        ix->has_ids = true;
    }
    char* dst = (char*)ix->data + (size_t)ix->size * ix->row_bytes();
    TS_CHECK_CUDA(cudaMemcpy(dst, rows, (size_t)n * ix->row_bytes(), cudaMemcpyHostToDevice));
    if (ix->has_ids) {
        if (ids != nullptr) {
            TS_CHECK_CUDA(cudaMemcpy(ix->ids + ix->size, ids, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
        } else {
            const int64_t auto_first = std::max(ix->auto_next, ix->size);
            iota_ids_kernel<<<256, 256>>>(ix->ids + ix->size, auto_first, n);
            TS_LAUNCH_CHECK();
            ix->auto_next = auto_first + n;
        }
    }
    int rc = launch_max_norm2(dst, ix->dtype, n, ix->dim_pad, ix->max_norm2, nullptr);
    if (rc) return rc;
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    const int64_t first = ix->size;
    ix->size += n;
    ix->auto_next = std::max(ix->auto_next, ix->size);
    if (ids != nullptr)          // a reloaded index keeps numbering id-less rows past its largest stored id
        for (int64_t i = 0; i < n; ++i) ix->auto_next = std::max(ix->auto_next, ids[i] + 1);
    ix->id_map_valid = false;
    return ivf_apply_mutation(ix, nullptr, 0, first, n, nullptr);
}

int ts_ivf_list_dtype(const ts_index* ix) { return (ix && ix->ivf_built) ? ix->list_dtype : -1; }

// ------------------------------------------------------------------------------------ search
size_t ts_workspace_bytes(const ts_index* ix, int nq, int k) {
    if (!ix || nq < 0 || k < 1) return 0;
    const size_t scan = carve_ws(ix, std::max(nq, 1), k, nullptr).bytes;
    const size_t batched = batched_workspace_bytes(ix, std::max(nq, 1), k);
    return std::max(scan, batched);
}

int ts_search(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int normalize_queries,
              const uint32_t* allow_mask, float* out_scores, int64_t* out_ids, void* workspace,
              size_t workspace_bytes, void* stream) {
    TS_REQUIRE(nq == 0 || (out_scores != nullptr && out_ids != nullptr), TS_ERR_BAD_ARG,
               "search: output pointers are NULL");
    return search_impl(ix, queries, q_dtype, nq, k, normalize_queries, allow_mask, nullptr, out_scores,
                       out_ids, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr);
}

int ts_search_keys(ts_index* ix, const void* queries, int q_dtype, int nq, int k, int normalize_queries,
                   const uint32_t* allow_mask, uint64_t* out_keys, void* workspace, size_t workspace_bytes,
                   void* stream) {
    TS_REQUIRE(nq == 0 || out_keys != nullptr, TS_ERR_BAD_ARG, "search_keys: out_keys is NULL");
    return search_impl(ix, queries, q_dtype, nq, k, normalize_queries, allow_mask, out_keys, nullptr, nullptr,
                       workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr);
}

int ts_merge_topk(const uint64_t* keys, int nshards, int nq, int k, const int64_t* shard_base,
                  const int64_t* id_map, float* out_scores, int64_t* out_ids, void* stream) {
    TS_REQUIRE(keys != nullptr || nq == 0, TS_ERR_BAD_ARG, "merge_topk: keys is NULL");
    TS_REQUIRE(nq == 0 || (out_scores != nullptr && out_ids != nullptr), TS_ERR_BAD_ARG,
               "merge_topk: output pointers are NULL");
    return launch_merge(keys, nshards, nq, k, /*query_major=*/false, shard_base, id_map, nullptr, out_scores,
                        out_ids, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ ctx
int ts_ctx_create(ts_ctx** out, ts_index* ix, int max_nq, int max_k) {
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "ctx_create: out is NULL");
    *out = nullptr;
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "ctx_create: index is NULL");
    TS_REQUIRE(max_nq >= 1 && max_k >= 1 && max_k <= TS_MAX_K, TS_ERR_BAD_ARG, "ctx_create: max_nq=%d max_k=%d",
               max_nq, max_k);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "ctx_create: cannot select CUDA device %d", ix->device);
    ts_ctx* c = new ts_ctx();
    c->index = ix;
    c->max_nq = max_nq;
    c->max_k = max_k;
    c->workspace_bytes = ts_workspace_bytes(ix, max_nq, max_k);
    const size_t qb = (size_t)max_nq * ix->dim * sizeof(float);
    const size_t sb = (size_t)max_nq * max_k * sizeof(float);
    const size_t ib = (size_t)max_nq * max_k * sizeof(int64_t);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_queries, qb);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_scores, sb, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_ids, ib, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_done, sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *c->h_done = 0u;
        e = cudaHostGetDevicePointer((void**)&c->m_scores, c->h_scores, 0);
    }
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&c->m_ids, c->h_ids, 0);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&c->m_done, c->h_done, 0);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_queries, qb);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_scores, sb);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ids, ib);
    if (e == cudaSuccess) e = cudaMalloc(&c->workspace, c->workspace_bytes);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) {
        set_error("ctx_create: allocation failed: %s", cudaGetErrorString(e));
        ts_ctx_destroy(c);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? TS_ERR_OOM : TS_ERR_CUDA;
    }
    *out = c;
    return TS_OK;
}

void ts_ctx_destroy(ts_ctx* c) {
    if (!c) return;
    DeviceGuard g(c->index ? c->index->device : 0);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFreeHost(c->h_queries);
    cudaFreeHost(c->h_scores);
    cudaFreeHost(c->h_ids);
    cudaFreeHost(c->h_done);
    cudaFree(c->d_queries);
    cudaFree(c->d_scores);
    cudaFree(c->d_ids);
    cudaFree(c->workspace);
    cudaFree(c->ivf_workspace);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ts_ctx_set_timing(ts_ctx* c, int enabled) {
    TS_REQUIRE(c != nullptr, TS_ERR_BAD_ARG, "ctx_set_timing: ctx is NULL");
    c->timing = enabled != 0;
    c->last_ms = -1.f;
    return TS_OK;
}

float ts_ctx_last_kernel_ms(const ts_ctx* c) { return c ? c->last_ms : -1.f; }

int ts_search_host(ts_ctx* c, const float* queries, int nq, int k, int normalize_queries,
                   const uint32_t* allow_mask, float* out_scores, int64_t* out_ids) {
    TS_REQUIRE(c != nullptr, TS_ERR_BAD_ARG, "search_host: ctx is NULL");
    TS_REQUIRE(nq >= 0 && nq <= c->max_nq, TS_ERR_CAPACITY, "search_host: nq=%d exceeds ctx max_nq=%d", nq,
               c->max_nq);
    TS_REQUIRE(k >= 1 && k <= c->max_k, TS_ERR_CAPACITY, "search_host: k=%d exceeds ctx max_k=%d", k, c->max_k);
    if (nq == 0) return TS_OK;
    TS_REQUIRE(queries && out_scores && out_ids, TS_ERR_BAD_ARG, "search_host: NULL buffer");
    ts_index* ix = c->index;
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "search_host: cannot select CUDA device %d", ix->device);
    // the batched path's workspace depends on the index SIZE (dense small-corpus path): rows added since the ctx
    // was created may need more than it was sized for
    const size_t need = ts_workspace_bytes(ix, nq, k);
    if (need > c->workspace_bytes) {
        TS_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->workspace);
        c->workspace = nullptr;
        c->workspace_bytes = 0;
        TS_CHECK_CUDA(cudaMalloc(&c->workspace, need));
        c->workspace_bytes = need;
    }
    const size_t qb = (size_t)nq * ix->dim * sizeof(float);
    memcpy(c->h_queries, queries, qb);
    TS_CHECK_CUDA(cudaMemcpyAsync(c->d_queries, c->h_queries, qb, cudaMemcpyHostToDevice, c->stream));
    if (nq == 1 && !c->timing && !use_batched(ix, nq)) {
        // latency path: the kernel's last CTA writes the k results into mapped host memory and raises a flag
        const uint32_t seq = ++c->done_seq;
        int rc1 = search_impl(ix, c->d_queries, TS_F32, 1, k, normalize_queries, allow_mask, nullptr, c->m_scores, c->m_ids,
                              c->workspace, c->workspace_bytes, c->stream, nullptr, nullptr, c->m_done, seq);
        if (rc1) return rc1;
        rc1 = wait_done_flag(c->h_done, seq, c->stream, "search_host");
        if (rc1) return rc1;
        memcpy(out_scores, c->h_scores, (size_t)k * sizeof(float));
        memcpy(out_ids, c->h_ids, (size_t)k * sizeof(int64_t));
        return TS_OK;
    }
    int rc = search_impl(ix, c->d_queries, TS_F32, nq, k, normalize_queries, allow_mask, nullptr, c->d_scores,
                         c->d_ids, c->workspace, c->workspace_bytes, c->stream, c->timing ? c->ev0 : nullptr,
                         c->timing ? c->ev1 : nullptr, nullptr, 0);
    if (rc) return rc;
    const size_t n = (size_t)nq * k;
    TS_CHECK_CUDA(cudaMemcpyAsync(c->h_scores, c->d_scores, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TS_CHECK_CUDA(cudaMemcpyAsync(c->h_ids, c->d_ids, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    TS_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(out_scores, c->h_scores, n * sizeof(float));
    memcpy(out_ids, c->h_ids, n * sizeof(int64_t));
    if (c->timing) TS_CHECK_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    return TS_OK;
}

// ------------------------------------------------------------------------------------ peer exchange
static size_t xchg_slot_count(const ts_xchg* x) { return (size_t)2 * x->world * x->max_nq * x->max_k; }
static size_t xchg_flag_count(const ts_xchg* x) { return (size_t)2 * x->world * x->max_nq; }

int ts_xchg_create(ts_xchg** out, int device, int world, int rank, int max_nq, int max_k) {
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "xchg_create: out is NULL");
    *out = nullptr;
    TS_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, TS_ERR_BAD_ARG, "xchg_create: world=%d rank=%d",
               world, rank);
    TS_REQUIRE(max_nq >= 1 && max_k >= 1 && max_k <= TS_MAX_K, TS_ERR_BAD_ARG, "xchg_create: max_nq=%d max_k=%d", max_nq,
               max_k);
    DeviceGuard g(device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "xchg_create: cannot select CUDA device %d", device);
    ts_xchg* x = new ts_xchg();
    x->device = device;
    x->world = world;
    x->rank = rank;
    x->max_nq = max_nq;
    x->max_k = max_k;
    if (const char* env = getenv("TS_XCHG_TIMEOUT_MS")) {
        const long long ms = atoll(env);
        if (ms > 0) x->timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    x->bytes = xchg_slot_count(x) * sizeof(uint64_t) + xchg_flag_count(x) * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&x->base, x->bytes);
    if (e == cudaSuccess) e = cudaMemset(x->base, 0, x->bytes);
    if (e == cudaSuccess) e = cudaMalloc(&x->d_peer_slots, 16 * sizeof(void*));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_peer_flags, 16 * sizeof(void*));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_tickets, (size_t)max_nq * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(x->d_tickets, 0, (size_t)max_nq * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&x->d_done, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(x->d_done, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&x->h_error, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *x->h_error = 0;
        e = cudaHostGetDevicePointer((void**)&x->d_error, x->h_error, 0);
    }
    if (e != cudaSuccess) {
        set_error("xchg_create: allocation failed: %s", cudaGetErrorString(e));
        ts_xchg_destroy(x);
        cudaGetLastError();
        return TS_ERR_OOM;
    }
    x->slots = (uint64_t*)x->base;
    x->flags = (uint32_t*)((char*)x->base + xchg_slot_count(x) * sizeof(uint64_t));
    x->peer_base[rank] = x->base;
    *out = x;
    return TS_OK;
}

void ts_xchg_destroy(ts_xchg* x) {
    if (!x) return;
    DeviceGuard g(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    cudaFree(x->base);
    cudaFree(x->d_peer_slots);
    cudaFree(x->d_peer_flags);
    cudaFree(x->d_tickets);
    cudaFree(x->d_done);
    cudaFree(x->part_ring);
    cudaFreeHost(x->h_error);
    cudaFreeHost(x->h_queries);
    cudaFreeHost(x->h_scores);
    cudaFreeHost(x->h_ids);
    cudaFreeHost(x->h_done);
    cudaFree(x->d_queries);
    cudaFree(x->d_scores);
    cudaFree(x->d_ids);
    cudaFree(x->workspace);
    if (x->stream) cudaStreamDestroy(x->stream);
    delete x;
}

int ts_xchg_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int ts_xchg_handle(const ts_xchg* x, void* out_handle) {
    TS_REQUIRE(x != nullptr && out_handle != nullptr, TS_ERR_BAD_ARG, "xchg_handle: NULL argument");
    DeviceGuard g(x->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "xchg_handle: cannot select CUDA device %d", x->device);
    cudaIpcMemHandle_t h;
    TS_CHECK_CUDA(cudaIpcGetMemHandle(&h, x->base));
    memcpy(out_handle, &h, sizeof(h));
    return TS_OK;
}

int ts_xchg_connect(ts_xchg* x, const void* all_handles) {
    TS_REQUIRE(x != nullptr && (all_handles != nullptr || x->world == 1), TS_ERR_BAD_ARG, "xchg_connect: NULL argument");
    TS_REQUIRE(!x->connected, TS_ERR_STATE, "xchg_connect: the peers are already mapped");
    DeviceGuard g(x->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "xchg_connect: cannot select CUDA device %d", x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + (size_t)r * sizeof(h), sizeof(h));
        TS_CHECK_CUDA(cudaIpcOpenMemHandle(&x->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    uint64_t* hs[16] = {nullptr};
    uint32_t* hf[16] = {nullptr};
    for (int r = 0; r < x->world; ++r) {
        hs[r] = (uint64_t*)x->peer_base[r];
        hf[r] = (uint32_t*)((char*)x->peer_base[r] + xchg_slot_count(x) * sizeof(uint64_t));
    }
    TS_CHECK_CUDA(cudaMemcpy(x->d_peer_slots, hs, sizeof(hs), cudaMemcpyHostToDevice));
    TS_CHECK_CUDA(cudaMemcpy(x->d_peer_flags, hf, sizeof(hf), cudaMemcpyHostToDevice));
    x->connected = true;
    return TS_OK;
}

int ts_xchg_error(const ts_xchg* x) {
    if (!x || !x->h_error) return -1;
    return *reinterpret_cast<volatile int*>(x->h_error);   // host-mapped: no device synchronisation
}

int ts_xchg_set_timeout_ms(ts_xchg* x, int64_t ms) {
    TS_REQUIRE(x != nullptr && ms > 0, TS_ERR_BAD_ARG, "xchg_set_timeout_ms: ms=%lld", (long long)ms);
    x->timeout_ns = (unsigned long long)ms * 1000000ull;
    return TS_OK;
}

int ts_xchg_reset(ts_xchg* x) {
    TS_REQUIRE(x != nullptr, TS_ERR_BAD_ARG, "xchg_reset: NULL handle");
    DeviceGuard g(x->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "xchg_reset: cannot select CUDA device %d", x->device);
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    TS_CHECK_CUDA(cudaMemset(x->base, 0, x->bytes));
    TS_CHECK_CUDA(cudaMemset(x->d_tickets, 0, (size_t)x->max_nq * sizeof(uint32_t)));
    TS_CHECK_CUDA(cudaMemset(x->d_done, 0, 2 * sizeof(uint32_t)));
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    *reinterpret_cast<volatile int*>(x->h_error) = 0;
    x->seq = 0;
    return TS_OK;
}

uint32_t ts_xchg_seq(const ts_xchg* x) { return x ? x->seq : 0u; }

}  // extern "C"

static int search_sharded_impl(ts_index* ix, ts_xchg* x, const void* queries, int q_dtype, int nq, int k, int normalize_queries,
                               const uint32_t* allow_mask, int64_t shard_base, const int64_t* id_map, float* out_scores,
                               int64_t* out_ids, void* workspace, size_t workspace_bytes, int flags, void* stream,
                               uint32_t* done_flag, uint32_t done_value) {
    TS_REQUIRE(ix != nullptr && x != nullptr, TS_ERR_BAD_ARG, "search_sharded: NULL handle");
    TS_REQUIRE(x->connected, TS_ERR_STATE, "search_sharded: ts_xchg_connect has not been called");
    TS_REQUIRE(x->device == ix->device, TS_ERR_BAD_ARG, "search_sharded: index on device %d, exchange on %d", ix->device,
               x->device);
    TS_REQUIRE(*reinterpret_cast<volatile int*>(x->h_error) == 0, TS_ERR_STATE,
               "search_sharded: an earlier search timed out waiting for a peer (its result was poisoned with -inf / -1); "
               "bring every rank to a barrier and call ts_xchg_reset before searching again");
    TS_REQUIRE(nq >= 1 && nq <= x->max_nq, TS_ERR_CAPACITY, "search_sharded: nq=%d outside [1, %d]", nq, x->max_nq);
    TS_REQUIRE(k >= 1 && k <= x->max_k, TS_ERR_CAPACITY, "search_sharded: k=%d outside [1, %d]", k, x->max_k);
    TS_REQUIRE(!use_batched(ix, nq), TS_ERR_UNSUPPORTED,
               "search_sharded: nq=%d goes to the batched path, whose results are exchanged with an all-gather", nq);
    TS_REQUIRE(q_dtype == TS_F32 || q_dtype == TS_BF16 || q_dtype == TS_F16, TS_ERR_BAD_ARG, "search_sharded: query dtype %d",
               q_dtype);
    TS_REQUIRE(queries && out_scores && out_ids && workspace, TS_ERR_BAD_ARG, "search_sharded: NULL buffer");
    TS_REQUIRE(shard_base >= 0 && shard_base + ix->size < (int64_t)0xFFFFFFFFll, TS_ERR_UNSUPPORTED,
               "search_sharded: global rows must fit 32 bits");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "search_sharded: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    SearchWs w = carve_ws(ix, nq, k, workspace);
    TS_REQUIRE(workspace_bytes >= w.bytes, TS_ERR_CAPACITY, "search_sharded: workspace %zu < %zu bytes", workspace_bytes,
               w.bytes);
    if (id_map == nullptr && x->world == 1 && shard_base == 0 && ix->has_ids) id_map = ix->ids;   // a world of one: the index's own ids
    XchgDev xd;
    xd.peer_slots = x->d_peer_slots;
    xd.peer_flags = x->d_peer_flags;
    xd.my_slots = x->slots;
    xd.my_flags = x->flags;
    xd.error = x->d_error;
    xd.done_seq = x->d_done;
    xd.done_count = x->d_done + 1;
    xd.nq = nq;
    xd.world = x->world;
    xd.rank = x->rank;
    xd.max_nq = x->max_nq;
    xd.max_k = x->max_k;
    xd.seq = x->seq + 1u;      // committed (x->seq) only once the search is enqueued: a failed call must not leave this
                               // rank one sequence number ahead of its peers
    xd.base = shard_base;
    xd.timeout_ns = x->timeout_ns;
    xd.debug_no_flag = tunables().xchg_debug_no_flag;
    ScanFused f;
    f.q_raw = queries;
    f.q_dtype = q_dtype;
    f.q_normalize = normalize_queries;
    f.id_map = id_map;
    f.out_keys = nullptr;
    if (flags & TS_SHARDED_ONE_KERNEL) {
        // the exchange context owns self-cleaning tickets: the whole search is ONE stream operation (the kernel)
        f.tickets = x->d_tickets;
        f.out_scores = out_scores;
        f.out_ids = out_ids;
        f.xchg = xd;
        f.pdl = 0;
        f.ring_gate = nullptr;
        f.ring_need = 0;
        f.done_flag = done_flag;
        f.done_value = done_value;
        int rc1 = launch_scan_topk(ix, ix->data, ix->dtype, ix->size, w.q_f32, nq, k, allow_mask, w.part_keys, w.nparts, s,
                                   nullptr, nullptr, nullptr, nullptr, &f);
        if (rc1 == TS_OK) x->seq = xd.seq;
        return rc1;
    }
    // two kernels chained by programmatic dependent launch: the scan writes its per-CTA lists, the exchange
    // kernel (resident early, asleep until the scan completes) merges, exchanges and writes the result while
    // the NEXT search's scan already streams the corpus
    // The per-CTA lists go into a ring of four buffers owned by the exchange (search seq uses slot seq % 4), so a
    // scan never has to wait for the previous search's exchange kernel before it can write its lists.
    if (x->part_ring == nullptr || x->ring_nparts != w.nparts) {
        TS_CHECK_CUDA(cudaDeviceSynchronize());
        cudaFree(x->part_ring);
        x->part_ring = nullptr;
        TS_CHECK_CUDA(cudaMalloc(&x->part_ring, (size_t)4 * x->max_nq * w.nparts * x->max_k * sizeof(uint64_t)));
        x->ring_nparts = w.nparts;
    }
    uint64_t* part = x->part_ring + (size_t)(xd.seq & 3u) * x->max_nq * w.nparts * x->max_k;
    f.tickets = nullptr;
    f.out_scores = nullptr;
    f.out_ids = nullptr;
    memset(&f.xchg, 0, sizeof(f.xchg));
    f.pdl = 1 | ((flags & TS_SHARDED_INDEPENDENT) ? 0 : 2);
    f.ring_gate = x->d_done;
    f.ring_need = xd.seq - 4u;
    int rc = launch_scan_topk(ix, ix->data, ix->dtype, ix->size, w.q_f32, nq, k, allow_mask, part, w.nparts, s,
                              nullptr, nullptr, nullptr, nullptr, &f);
    if (rc) return rc;
    x->seq = xd.seq;
    return launch_xchg_finish(part, w.nparts, nq, k, xd, id_map, out_scores, out_ids, s, done_flag, done_value);
}

extern "C" {

int ts_search_sharded(ts_index* ix, ts_xchg* x, const void* queries, int q_dtype, int nq, int k, int normalize_queries,
                      const uint32_t* allow_mask, int64_t shard_base, const int64_t* id_map, float* out_scores,
                      int64_t* out_ids, void* workspace, size_t workspace_bytes, int flags, void* stream) {
    return search_sharded_impl(ix, x, queries, q_dtype, nq, k, normalize_queries, allow_mask, shard_base, id_map, out_scores,
                               out_ids, workspace, workspace_bytes, flags, stream, nullptr, 0);
}

int ts_search_sharded_host(ts_index* ix, ts_xchg* x, const float* queries, int nq, int k, int normalize_queries,
                           const uint32_t* allow_mask, int64_t shard_base, const int64_t* id_map, float* out_scores,
                           int64_t* out_ids) {
    TS_REQUIRE(ix != nullptr && x != nullptr, TS_ERR_BAD_ARG, "search_sharded_host: NULL handle");
    TS_REQUIRE(nq >= 1 && nq <= x->max_nq, TS_ERR_CAPACITY, "search_sharded_host: nq=%d outside [1, %d]", nq, x->max_nq);
    TS_REQUIRE(k >= 1 && k <= x->max_k, TS_ERR_CAPACITY, "search_sharded_host: k=%d outside [1, %d]", k, x->max_k);
    TS_REQUIRE(queries && out_scores && out_ids, TS_ERR_BAD_ARG, "search_sharded_host: NULL buffer");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "search_sharded_host: cannot select CUDA device %d", ix->device);
    if (x->host_dim != ix->dim) {   // first call: stream, pinned staging, device buffers
        TS_REQUIRE(x->host_dim == 0, TS_ERR_BAD_ARG, "search_sharded_host: the exchange was set up for dim %d", x->host_dim);
        const size_t qb = (size_t)x->max_nq * ix->dim * sizeof(float);
        const size_t sb = (size_t)x->max_nq * x->max_k * sizeof(float);
        const size_t ib = (size_t)x->max_nq * x->max_k * sizeof(int64_t);
        x->workspace_bytes = carve_ws(ix, x->max_nq, x->max_k, nullptr).bytes;
        cudaError_t e = x->stream ? cudaSuccess : cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking);
        auto host = [&](void** p, size_t bytes, unsigned fl) {   // an earlier attempt may have got part of the way
            if (e == cudaSuccess && *p == nullptr) e = cudaHostAlloc(p, bytes, fl);
        };
        auto dev = [&](void** p, size_t bytes) {
            if (e == cudaSuccess && *p == nullptr) e = cudaMalloc(p, bytes);
        };
        host((void**)&x->h_queries, qb, cudaHostAllocDefault);
        host((void**)&x->h_scores, sb, cudaHostAllocMapped);
        host((void**)&x->h_ids, ib, cudaHostAllocMapped);
        host((void**)&x->h_done, sizeof(uint32_t), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&x->m_scores, x->h_scores, 0);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&x->m_ids, x->h_ids, 0);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&x->m_done, x->h_done, 0);
        dev((void**)&x->d_queries, qb);
        dev((void**)&x->d_scores, sb);
        dev((void**)&x->d_ids, ib);
        dev((void**)&x->workspace, x->workspace_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("search_sharded_host: staging allocation failed: %s", cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? TS_ERR_OOM : TS_ERR_CUDA;
        }
        *x->h_done = 0u;
        x->done_seq = 0u;
        x->host_dim = ix->dim;
    }
    const size_t qb = (size_t)nq * ix->dim * sizeof(float);
    memcpy(x->h_queries, queries, qb);
    TS_CHECK_CUDA(cudaMemcpyAsync(x->d_queries, x->h_queries, qb, cudaMemcpyHostToDevice, x->stream));
    const size_t n = (size_t)nq * k;
    if (nq == 1) {
        // latency path: the exchange kernel writes the k results into mapped host memory and raises a flag
        const uint32_t seq = ++x->done_seq;
        int rc1 = search_sharded_impl(ix, x, x->d_queries, TS_F32, 1, k, normalize_queries, allow_mask, shard_base, id_map,
                                      x->m_scores, x->m_ids, x->workspace, x->workspace_bytes, 0, x->stream, x->m_done, seq);
        if (rc1) return rc1;
        rc1 = wait_done_flag(x->h_done, seq, x->stream, "search_sharded_host");
        if (rc1) return rc1;
        memcpy(out_scores, x->h_scores, n * sizeof(float));
        memcpy(out_ids, x->h_ids, n * sizeof(int64_t));
        TS_REQUIRE(*reinterpret_cast<volatile int*>(x->h_error) == 0, TS_ERR_STATE,
                   "search_sharded_host: timed out waiting for a peer's keys; the result was poisoned (-inf / -1)");
        return TS_OK;
    }
    int rc = ts_search_sharded(ix, x, x->d_queries, TS_F32, nq, k, normalize_queries, allow_mask, shard_base, id_map,
                               x->d_scores, x->d_ids, x->workspace, x->workspace_bytes, 0, x->stream);
    if (rc) return rc;
    TS_CHECK_CUDA(cudaMemcpyAsync(x->h_scores, x->d_scores, n * sizeof(float), cudaMemcpyDeviceToHost, x->stream));
    TS_CHECK_CUDA(cudaMemcpyAsync(x->h_ids, x->d_ids, n * sizeof(int64_t), cudaMemcpyDeviceToHost, x->stream));
    TS_CHECK_CUDA(cudaStreamSynchronize(x->stream));
    memcpy(out_scores, x->h_scores, n * sizeof(float));
    memcpy(out_ids, x->h_ids, n * sizeof(int64_t));
    TS_REQUIRE(*reinterpret_cast<volatile int*>(x->h_error) == 0, TS_ERR_STATE,
               "search_sharded_host: timed out waiting for a peer's keys; the result was poisoned (-inf / -1)");
    return TS_OK;
}

// ------------------------------------------------------------------------------------ tunables
static int* tunable_slot(const char* name) {
    Tunables& t = tunables();
    if (!name) return nullptr;
    if (!strcmp(name, "scan.ctas_per_sm")) return &t.scan_ctas_per_sm;
    if (!strcmp(name, "scan.warps")) return &t.scan_warps;
    if (!strcmp(name, "scan.stages")) return &t.scan_stages;
    if (!strcmp(name, "scan.tile_rows")) return &t.scan_tile_rows;
    if (!strcmp(name, "batch.min_nq")) return &t.batch_min_nq;
    if (!strcmp(name, "batch.cap")) return &t.batch_cap;
    if (!strcmp(name, "batch.first_chunk")) return &t.batch_first_chunk;
    if (!strcmp(name, "batch.growth")) return &t.batch_growth;
    if (!strcmp(name, "batch.dense")) return &t.batch_dense;
    if (!strcmp(name, "batch.tf32")) return &t.batch_tf32;
    if (!strcmp(name, "batch.a_policy")) return &t.batch_a_policy;
    if (!strcmp(name, "batch.cta_pair")) return &t.batch_cta_pair;
    if (!strcmp(name, "batch.pair_min_nq")) return &t.batch_pair_min_nq;
    if (!strcmp(name, "ivf.warps")) return &t.ivf_warps;
    if (!strcmp(name, "ivf.tile_rows")) return &t.ivf_tile_rows;
    if (!strcmp(name, "ivf.parts")) return &t.ivf_parts;
    if (!strcmp(name, "ivf.timeline")) return &t.ivf_timeline;
    if (!strcmp(name, "ivf.group_min_nq")) return &t.ivf_group_min_nq;
    if (!strcmp(name, "ivf.group_mma")) return &t.ivf_group_mma;
    if (!strcmp(name, "ivf.select_warp")) return &t.ivf_select_warp;
    if (!strcmp(name, "ivf.fuse_rescore")) return &t.ivf_fuse_rescore;
    if (!strcmp(name, "xchg.debug_no_flag")) return &t.xchg_debug_no_flag;
    if (!strcmp(name, "scan.timeline")) return &t.scan_timeline;
    if (!strcmp(name, "store.no_vmm")) return &t.store_no_vmm;
    if (!strcmp(name, "ivf.group_min_lists")) return &t.ivf_group_min_lists;
    return nullptr;
}
int ts_debug_last_batched_fixups(void) { return debug_last_batched_fixups(); }
int ts_debug_scan_timeline(uint64_t* out_host, int launches_back, int n_ctas) {
    return debug_scan_timeline(out_host, launches_back, n_ctas);
}

int ts_set_tunable(const char* name, int value) {
    int* s = tunable_slot(name);
    TS_REQUIRE(s != nullptr, TS_ERR_BAD_ARG, "set_tunable: unknown tunable '%s'", name ? name : "(null)");
    *s = value;
    return TS_OK;
}
int ts_get_tunable(const char* name, int* value) {
    int* s = tunable_slot(name);
    TS_REQUIRE(s != nullptr && value != nullptr, TS_ERR_BAD_ARG, "get_tunable: unknown tunable '%s'",
               name ? name : "(null)");
    *value = *s;
    return TS_OK;
}

}  // extern "C"
