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

// delete by id: the last rows move into the freed slots (one warp per row, 16-byte chunks), their ids with them
__global__ void relocate_rows_kernel(uint8_t* __restrict__ data, uint32_t row_bytes, const int64_t* __restrict__ from,
                                     const int64_t* __restrict__ to, int64_t n, int64_t* __restrict__ ids) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (int64_t)gridDim.x * 8) {
        const uint8_t* src = data + (size_t)from[i] * row_bytes;
        uint8_t* dst = data + (size_t)to[i] * row_bytes;
        for (uint32_t off = lane * 16u; off < row_bytes; off += 512u)
            *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(src + off);
        if (lane == 0) ids[to[i]] = ids[from[i]];
    }
}
__global__ void identity_remap_kernel(uint32_t* remap, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        remap[i] = (uint32_t)i;
}
__global__ void scatter_remap_kernel(uint32_t* __restrict__ remap, const int64_t* __restrict__ rows,
                                     const int64_t* __restrict__ now, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        remap[rows[i]] = now ? (uint32_t)now[i] : TS_DEAD_ROW;
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
        ix->auto_next = std::max(ix->auto_next, ix->size);
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

// DELETE: the reference re-parses a paper with `DELETE FROM theorem WHERE paper_id = ANY(%s)`
// (ec2/parse_arxiv_papers/__main__.py:271-274); theorem_slogan and theorem_embedding_qwen reference it
// ON DELETE CASCADE (rds_schema.sql:35,46,51), so the embedding rows of those slogans vanish from the corpus table.
int ts_index_delete(ts_index* ix, const int64_t* ids_host, int64_t n, int64_t* n_deleted_out, int64_t* moved_from,
                    int64_t* moved_to, int64_t* n_moved_out, void* stream) {
    TS_REQUIRE(ix != nullptr, TS_ERR_BAD_ARG, "index_delete: index is NULL");
    TS_REQUIRE(n >= 0 && (n == 0 || ids_host != nullptr), TS_ERR_BAD_ARG, "index_delete: n=%lld / ids", (long long)n);
    if (n_deleted_out) *n_deleted_out = 0;
    if (n_moved_out) *n_moved_out = 0;
    if (n == 0 || ix->size == 0) return TS_OK;
    DeviceGuard g(ix->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "index_delete: cannot select CUDA device %d", ix->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_host_id_map(ix);
    if (rc) return rc;
    auto& map = *ix->id_map_host;
    std::vector<int64_t> dead;                       // stored rows to delete; an id that is not stored matches nothing
    for (int64_t i = 0; i < n; ++i) {
        auto it = map.find(ids_host[i]);
        if (it == map.end()) continue;               // unknown id, or a repeat of one already taken
        dead.push_back(it->second);
        map.erase(it);
    }
    if (dead.empty()) return TS_OK;
    auto rollback = [&]() { ix->id_map_valid = false; };   // the map was edited optimistically
    std::sort(dead.begin(), dead.end());
    const int64_t d = (int64_t)dead.size(), old_size = ix->size, new_size = old_size - d;
    // rows >= new_size that survive move into the deleted slots < new_size (same count on both sides)
    std::vector<int64_t> from, to;
    {
        size_t tail = std::lower_bound(dead.begin(), dead.end(), new_size) - dead.begin();   // dead[tail..) lie in the tail
        size_t hole = 0;
        for (int64_t r = new_size; r < old_size; ++r) {
            if (tail < dead.size() && dead[tail] == r) {
                ++tail;
                continue;
            }
            from.push_back(r);
            to.push_back(dead[hole++]);
        }
    }
    const int64_t m = (int64_t)from.size();
    if ((rc = ensure_id_table(ix, s))) {             // positions stop being ids once rows move
        rollback();
        return rc;
    }
    std::vector<int64_t> tail_ids((size_t)(old_size - new_size));
    cudaError_t e = cudaMemcpyAsync(tail_ids.data(), ix->ids + new_size, tail_ids.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    int64_t *d_from = nullptr, *d_to = nullptr, *d_dead = nullptr;
    uint32_t* d_remap = nullptr;
    if (e == cudaSuccess && m > 0) {
        e = cudaMalloc(&d_from, (size_t)m * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMalloc(&d_to, (size_t)m * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_from, from.data(), (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_to, to.data(), (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess && ix->ivf_built) {
        // the lists name rows by position: old position -> new position / deleted, consumed by ivf_apply_delete
        e = cudaMalloc(&d_remap, (size_t)old_size * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&d_dead, (size_t)d * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dead, dead.data(), (size_t)d * sizeof(int64_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            identity_remap_kernel<<<1024, 256, 0, s>>>(d_remap, old_size);
            scatter_remap_kernel<<<(int)std::min<int64_t>((d + 255) / 256, 1024), 256, 0, s>>>(d_remap, d_dead, nullptr, d);
            if (m > 0) scatter_remap_kernel<<<(int)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, s>>>(d_remap, d_from, d_to, m);
            g_launches.fetch_add(m > 0 ? 3 : 2, std::memory_order_relaxed);
            e = cudaGetLastError();
        }
    }
    rc = TS_OK;
    if (e != cudaSuccess) {
        set_error("index_delete: staging failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        rc = e == cudaErrorMemoryAllocation ? TS_ERR_OOM : TS_ERR_CUDA;
    }
    if (rc == TS_OK && m > 0) {
        relocate_rows_kernel<<<(int)std::min<int64_t>((m + 7) / 8, 148 * 16), 256, 0, s>>>(
            (uint8_t*)ix->data, (uint32_t)ix->row_bytes(), d_from, d_to, m, ix->ids);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = TS_ERR_CUDA;
    }
    if (rc == TS_OK) {
        ix->size = new_size;
        for (int64_t i = 0; i < m; ++i) map[tail_ids[(size_t)(from[(size_t)i] - new_size)]] = to[(size_t)i];
        rc = ivf_apply_delete(ix, d_remap, d, s);
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == TS_OK) {
        set_error("index_delete: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = TS_ERR_CUDA;
    }
    cudaFree(d_from);
    cudaFree(d_to);
    cudaFree(d_dead);
    cudaFree(d_remap);
    if (rc != TS_OK) {
        rollback();
        return rc;
    }
    if (n_deleted_out) *n_deleted_out = d;
    if (n_moved_out) *n_moved_out = m;
    if (moved_from && moved_to)
        for (int64_t i = 0; i < m; ++i) {
            moved_from[i] = from[(size_t)i];
            moved_to[i] = to[(size_t)i];
        }
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
