// Index mutation: growable capacity and upsert-by-id.
//
// Reference writer: ec2/generate_embeddings/__main__.py:84-101 —
//     INSERT INTO theorem_embedding_qwen (slogan_id, embedding) VALUES ...
//     ON CONFLICT (slogan_id) DO UPDATE SET embedding = EXCLUDED.embedding
// i.e. a row whose id is already stored is REPLACED IN PLACE, any other row is appended; the table has no fixed
// capacity. Here: the id -> row map lives on the host (upserts arrive from the host writer in batches), the rows
// are normalised / quantised by K1 straight into their (scattered) row slots, the row store grows geometrically,
// and built IVF lists are kept valid incrementally (tombstone + overflow lists, k4_ivf.cu:ivf_apply_mutation).
#include "ts_common.cuh"

#include <algorithm>
#include <vector>

namespace ts {

__global__ void iota_ids2_kernel(int64_t* ids, int64_t first, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        ids[i] = first + i;
}
__global__ void scatter_ids_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ dst_rows, int64_t n,
                                   int64_t* __restrict__ table) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (dst_rows[i] >= 0) table[dst_rows[i]] = ids[i];
}

// The row store and the side tables (ids, list positions) grow IN PLACE: more physical memory is mapped behind the
// data in each one's reserved address range (vmm_store.cu) — nothing is copied, no second allocation is needed.
int index_reserve(ts_index* ix, int64_t capacity) {
    if (capacity <= ix->capacity) return TS_OK;
    TS_REQUIRE(capacity < (int64_t)0xFFFFFFFFll, TS_ERR_UNSUPPORTED, "index_reserve: capacity %lld exceeds 2^32-2 rows",
               (long long)capacity);
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    int rc = row_store_reserve(ix->store, (size_t)capacity * ix->row_bytes(), (size_t)ix->size * ix->row_bytes());
    if (rc) return rc;
    ix->data = row_store_ptr(ix->store);
    if (ix->has_ids && (rc = side_table_reserve(ix->ids_store, &ix->ids, ix->size, capacity))) return rc;
    if (ix->pos_of_row != nullptr && (rc = side_table_reserve(ix->pos_store, &ix->pos_of_row, ix->size, capacity))) return rc;
    ix->capacity = capacity;
    return TS_OK;
}

// room for `extra` more rows: geometric growth (x1.5), falling back to the exact size when memory is short
int index_make_room(ts_index* ix, int64_t extra) {
    const int64_t need = ix->size + extra;
    if (need <= ix->capacity) return TS_OK;
    const int64_t want = std::min<int64_t>(std::max<int64_t>(need, ix->capacity + ix->capacity / 2), (int64_t)0xFFFFFFFEll);
    if (want > need && index_reserve(ix, want) == TS_OK) return TS_OK;
    return index_reserve(ix, need);
}

static int ensure_id_table(ts_index* ix, cudaStream_t s) {
    if (ix->has_ids) return TS_OK;
    int rc = side_table_create(&ix->ids_store, &ix->ids, ix->device, ix->capacity);
    if (rc) return rc;
    if (ix->size > 0) {
        iota_ids2_kernel<<<256, 256, 0, s>>>(ix->ids, 0, ix->size);
        TS_LAUNCH_CHECK();
    }
    ix->has_ids = true;
    return TS_OK;
}

static int ensure_host_id_map(ts_index* ix) {
    if (ix->id_map_host == nullptr) ix->id_map_host = new std::unordered_map<int64_t, int64_t>();
    if (ix->id_map_valid) return TS_OK;
    auto& m = *ix->id_map_host;
    m.clear();
    m.reserve((size_t)ix->size * 2 + 16);
    if (ix->has_ids) {
        std::vector<int64_t> h((size_t)ix->size);
        if (ix->size > 0)
            TS_CHECK_CUDA(cudaMemcpy(h.data(), ix->ids, (size_t)ix->size * sizeof(int64_t), cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r < ix->size; ++r) m[h[(size_t)r]] = r;   // a duplicated id resolves to its LAST row
    } else {
        for (int64_t r = 0; r < ix->size; ++r) m[r] = r;
    }
    ix->id_map_valid = true;
    return TS_OK;
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_index_reserve(ts_index* ix, int64_t capacity) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_reserve: index is NULL");
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_reserve: cannot select CUDA device %d", ix->device);
    return index_reserve(ix, capacity);
}

int ts_index_upsert(ts_index* ix, const void* rows, int src_dtype, int64_t n, int normalize, const int64_t* ids_host,
                    int64_t* n_replaced_out, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_upsert: index is NULL");
    TS_REQUIRE(n >= 0, TS_ERR_BAD_ARG, "index_upsert: n=%lld", (long long)n);
    if (n_replaced_out) *n_replaced_out = 0;
    if (n == 0) return TS_OK;
    TS_REQUIRE(rows != nullptr && ids_host != nullptr, TS_ERR_BAD_ARG, "index_upsert: rows / ids is NULL (ids are what an upsert keys on)");
    TS_REQUIRE(src_dtype == TS_F32 || src_dtype == TS_BF16 || src_dtype == TS_F16, TS_ERR_BAD_ARG,
               "index_upsert: source dtype %d", src_dtype);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_upsert: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_host_id_map(ix);
    if (rc) return rc;
    auto& map = *ix->id_map_host;
    // resolve every source row to its destination: an existing row (replace) or the next free one (append). An id
    // that occurs twice in the batch keeps its LAST occurrence, as a row-by-row upsert would; the earlier ones
    // are skipped (-1) so that no two source rows race for one destination.
    std::vector<int64_t> dst((size_t)n);
    std::vector<uint32_t> replaced;
    std::unordered_map<int64_t, int64_t> last_writer;   // destination row -> source index, for in-batch duplicates
    int64_t n_new = 0;
    bool identity = !ix->has_ids;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t id = ids_host[i];
        auto it = map.find(id);
        int64_t row;
        if (it != map.end()) {
            row = it->second;
            if (row < ix->size) replaced.push_back((uint32_t)row);
        } else {
            row = ix->size + n_new++;
            map.emplace(id, row);
            if (id != row) identity = false;
        }
        auto lw = last_writer.find(row);
        if (lw != last_writer.end()) {
            dst[(size_t)lw->second] = -1;
            lw->second = i;
        } else {
            last_writer.emplace(row, i);
        }
        dst[(size_t)i] = row;
    }
    std::sort(replaced.begin(), replaced.end());
    replaced.erase(std::unique(replaced.begin(), replaced.end()), replaced.end());
    auto rollback = [&]() { ix->id_map_valid = false; };   // the map was edited optimistically
    if ((rc = index_make_room(ix, n_new))) {
        rollback();
        return rc;
    }
    if (!identity && (rc = ensure_id_table(ix, s))) {
        rollback();
        return rc;
    }
    int64_t *d_dst = nullptr, *d_ids = nullptr;
    uint32_t* d_replaced = nullptr;
    cudaError_t e = cudaMalloc(&d_dst, (size_t)n * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dst, dst.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && ix->has_ids) {
        e = cudaMalloc(&d_ids, (size_t)n * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ids, ids_host, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess && !replaced.empty()) {
        e = cudaMalloc(&d_replaced, replaced.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_replaced, replaced.data(), replaced.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
    }
    rc = TS_OK;
    if (e != cudaSuccess) {
        set_error("index_upsert: staging failed: %s", cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? TS_ERR_OOM : TS_ERR_CUDA;
    }
    if (rc == TS_OK && ix->has_ids) {
        scatter_ids_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, s>>>(d_ids, d_dst, n, ix->ids);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = TS_ERR_CUDA;
    }
    if (rc == TS_OK)
        rc = launch_normalize_cast(rows, src_dtype, n, ix->dim, ix->dim_pad, normalize, ix->data, ix->dtype, s, ix->max_norm2,
                                   d_dst);
    const int64_t old_size = ix->size;
    if (rc == TS_OK) {
        ix->size += n_new;
        rc = ivf_apply_mutation(ix, d_replaced, (int64_t)replaced.size(), old_size, n_new, s);
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == TS_OK) {
        set_error("index_upsert: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = TS_ERR_CUDA;
    }
    cudaFree(d_dst);
    cudaFree(d_ids);
    cudaFree(d_replaced);
    if (rc != TS_OK) {
        rollback();
        return rc;
    }
    if (n_replaced_out) *n_replaced_out = (int64_t)replaced.size();
    return TS_OK;
}

int ts_index_upsert_host(ts_index* ix, const void* rows, int src_dtype, int64_t n, int normalize, const int64_t* ids_host,
                         int64_t* n_replaced_out) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_upsert_host: index is NULL");
    TS_REQUIRE(n >= 0, TS_ERR_BAD_ARG, "index_upsert_host: n=%lld", (long long)n);
    if (n_replaced_out) *n_replaced_out = 0;
    if (n == 0) return TS_OK;
    TS_REQUIRE(rows != nullptr && ids_host != nullptr, TS_ERR_BAD_ARG, "index_upsert_host: rows / ids is NULL");
    TS_REQUIRE(src_dtype == TS_F32 || src_dtype == TS_BF16 || src_dtype == TS_F16, TS_ERR_BAD_ARG,
               "index_upsert_host: source dtype %d", src_dtype);
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_upsert_host: cannot select CUDA device %d", ix->device);
    const size_t src_row = (size_t)ix->dim * (src_dtype == TS_F32 ? 4 : 2);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)((64u << 20) / src_row)));
    void* d_rows = nullptr;
    TS_CHECK_CUDA(cudaMalloc(&d_rows, (size_t)chunk * src_row));
    int rc = TS_OK;
    int64_t replaced_total = 0;
    for (int64_t pos = 0; pos < n && rc == TS_OK; pos += chunk) {   // chunks are upserted in order: later ids win
        const int64_t m = std::min(chunk, n - pos);
        cudaError_t e = cudaMemcpy(d_rows, (const char*)rows + (size_t)pos * src_row, (size_t)m * src_row,
                                   cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("index_upsert_host: H2D copy failed: %s", cudaGetErrorString(e));
            rc = TS_ERR_CUDA;
            break;
        }
        int64_t r = 0;
        rc = ts_index_upsert(ix, d_rows, src_dtype, m, normalize, ids_host + pos, &r, nullptr);
        replaced_total += r;
    }
    cudaFree(d_rows);
    if (n_replaced_out) *n_replaced_out = replaced_total;
    return rc;
}

}  // extern "C"
