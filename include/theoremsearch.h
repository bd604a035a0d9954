/*
 * theoremsearch.h — C ABI of libtheoremsearch.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for TheoremSearch's retrieval hot path: "score a query
 * embedding (or a batch) against the corpus of theorem/slogan embeddings, return the top-k
 * theorem ids".  The reference has no FFI of its own; every entry point below names the
 * reference expression it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - Plain C: pointers, sizes, ints.  No torch / C++ types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device
 *     work is enqueued on it; no entry point synchronises the device unless it says so.
 *   - Unless a name ends in `_host`, data pointers are DEVICE pointers owned by the caller.
 *   - Return value: 0 = TS_OK, negative = error class; text via ts_last_error() (thread-local).
 *   - The library owns only what hangs off a ts_index / ts_ctx handle.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     TS_ERR_CUDA.
 *
 * Result order everywhere: score descending, ties broken by LOWER row/id first
 * (BASELINE.json north_star); entries past the number of eligible rows have score = -inf,
 * id = -1.
 */
#ifndef THEOREMSEARCH_H_
#define THEOREMSEARCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TS_API
#else
#define TS_API __attribute__((visibility("default")))
#endif

/* ---- enums ------------------------------------------------------------------------- */

enum ts_status {
    TS_OK = 0,
    TS_ERR_BAD_ARG = -1,      /* null pointer, bad dtype, k out of range, dim mismatch ...   */
    TS_ERR_CUDA = -2,         /* a CUDA runtime call failed (message has the cudaError text) */
    TS_ERR_OOM = -3,          /* device or pinned-host allocation failed                     */
    TS_ERR_UNSUPPORTED = -4,  /* valid request this build has no kernel for                  */
    TS_ERR_CAPACITY = -5,     /* index full / workspace too small                            */
    TS_ERR_STATE = -6         /* call order (e.g. ivf_search before ivf_build)               */
};

enum ts_dtype {
    TS_F32 = 0,       /* IEEE binary32                                   */
    TS_BF16 = 1,      /* bfloat16 (corpus storage for exact search)      */
    TS_FP8_E4M3 = 2,  /* e4m3 + one fp32 scale per row (IVF list storage) */
    TS_F16 = 3        /* IEEE binary16 (accepted as a source dtype only) */
};

#define TS_MAX_K 1024   /* largest k any search entry point accepts            */
#define TS_MAX_DIM 2048 /* largest embedding dimension (reference uses 768/1024) */
#define TS_IVF_MAX_CANDIDATES 256 /* largest k / rescore_k of the IVF entry points */

typedef struct ts_index ts_index; /* the corpus: quantised rows (+ ids, + IVF lists) on one GPU */
typedef struct ts_ctx ts_ctx;     /* per-caller search context: workspace, pinned staging, stream */

/* ---- library ----------------------------------------------------------------------- */

/* Version of this ABI (bumped on any signature change). */
TS_API int ts_abi_version(void);

/* Last error message of the calling thread ("" if none). Never NULL. */
TS_API const char* ts_last_error(void);

/* Number of CUDA kernels this library has launched since load (all threads). bench.py reads
 * it before/after the timed region to report "gpu_launches". */
TS_API uint64_t ts_kernel_launches(void);

/* ---- index: the corpus ---------------------------------------------------------------
 * Replaces the reference's corpus containers: the `embeddings_db` tensor
 * (test_app.py:129-130, app_showcase_model.py:52 `torch.load('corpus_embeddings.pt')`) and
 * the pgvector table `theorem_embedding_qwen(slogan_id BIGINT, embedding vector(1024))`
 * (rds_schema.sql:50-53). */

/* Allocate an empty index with room for `capacity` rows of `dim` elements stored as `dtype`
 * (TS_BF16 or TS_F32) on CUDA device `device`; the room grows on demand (see "mutation"). */
TS_API int ts_index_create(ts_index** out, int device, int dim, int dtype, int64_t capacity);

/* Free the index and everything it owns. NULL is a no-op. */
TS_API void ts_index_destroy(ts_index* index);

/* Append `n` rows (device pointer, row-major [n, dim], `src_dtype` in {F32, BF16, F16}).
 * normalize != 0: each row becomes x / max(||x||_2, 1e-12) before quantisation — the
 * `F.normalize` inside sentence_transformers.util.cos_sim (test_app.py:76) and
 * `model.encode(..., normalize_embeddings=True)` (ec2/generate_embeddings/embeddings.py:27,35).
 * `ids` (device int64[n]) are the caller's row keys (slogan_id / theorem_id,
 * rds_schema.sql:22,34); NULL = the row's position in the index — after a ts_index_delete, one past
 * the largest position ever occupied, like a SERIAL column: ids are not handed out twice.  Kernel K1. */
TS_API int ts_index_add(ts_index* index, const void* rows, int src_dtype, int64_t n,
                        int normalize, const int64_t* ids, void* stream);

/* Same, rows/ids in HOST memory; staged through pinned buffers in chunks. Synchronises. */
TS_API int ts_index_add_host(ts_index* index, const void* rows, int src_dtype, int64_t n,
                             int normalize, const int64_t* ids);

/* ---- mutation ------------------------------------------------------------------------
 * The table this index replaces has no fixed size and is written with
 *     INSERT ... ON CONFLICT (slogan_id) DO UPDATE SET embedding = EXCLUDED.embedding
 * (ec2/generate_embeddings/__main__.py:84-101). `capacity` is therefore only the initial reservation:
 * ts_index_add / ts_index_upsert grow the row store geometrically when it is full (the old and the new
 * allocation are resident together while the rows are copied across, device to device).
 * Built IVF lists stay valid across add / upsert, as pgvector's ivfflat accepts inserts after its build:
 * new and replaced rows are filed under the existing centroids in per-list OVERFLOW segments (scanned with
 * their list), a replaced row's old list entry is tombstoned, and once tombstones + overflow exceed a tenth
 * of the corpus the lists are re-packed from scratch.
 * add / upsert / reserve / train / build are exclusive with searches on the same index. */

/* 1 if the row store is virtual-memory mapped (a reserved address range into which physical memory is mapped as the
 * index grows: growth copies no row and needs no second allocation, ts_index_data() never changes), 0 if it is a
 * plain allocation that is copied on growth (driver entry points unavailable, or tunable "store.no_vmm"). */
TS_API int ts_index_grows_in_place(const ts_index* index);

/* Make room for at least `capacity` rows (never shrinks). Synchronises the device. */
TS_API int ts_index_reserve(ts_index* index, int64_t capacity);

/* Upsert n rows (DEVICE, row-major [n, dim] of src_dtype) keyed by ids (HOST int64[n]): a row whose id is
 * already stored is replaced in place, any other is appended. An id that occurs more than once in the batch
 * keeps its last occurrence (what row-by-row execution would leave). An index whose rows were added without
 * ids treats the row position as the id. *n_replaced (may be NULL) receives the number of distinct stored
 * rows that were replaced. Synchronises `stream` before returning. */
TS_API int ts_index_upsert(ts_index* index, const void* rows, int src_dtype, int64_t n, int normalize,
                           const int64_t* ids_host, int64_t* n_replaced, void* stream);
/* Same with the rows in HOST memory (staged in chunks, in order). */
TS_API int ts_index_upsert_host(ts_index* index, const void* rows, int src_dtype, int64_t n, int normalize,
                                const int64_t* ids_host, int64_t* n_replaced);

/* DELETE by id. Replaces the cascade of `DELETE FROM theorem WHERE paper_id = ANY(%s)`
 * (ec2/parse_arxiv_papers/__main__.py:271-274; theorem_slogan and theorem_embedding_qwen reference theorem
 * ON DELETE CASCADE, rds_schema.sql:35,46,51) on the corpus table. ids: HOST int64[n]; ids that are not stored
 * (or repeated) match nothing. The store stays dense: the LAST rows move into the freed slots, so ts_index_size
 * drops by the number deleted and row positions (the bit positions of an allow mask, the order of exact score
 * ties) change for the moved rows — moved_from / moved_to (HOST int64[n] each, or NULL) receive the
 * n_moved_out <= n relocations (old row -> new row) for callers that keep row-aligned side tables. Built IVF
 * lists stay valid (entries of deleted rows are tombstoned, entries of moved rows renamed; re-packed past 10 %
 * churn; deleting the last row drops the lists, the centroids stay). Enqueues on `stream` and synchronises it
 * before returning. */
TS_API int ts_index_delete(ts_index* index, const int64_t* ids, int64_t n, int64_t* n_deleted_out,
                           int64_t* moved_from, int64_t* moved_to, int64_t* n_moved_out, void* stream);

TS_API int64_t ts_index_size(const ts_index* index);     /* rows added so far */
TS_API int64_t ts_index_capacity(const ts_index* index);
TS_API int ts_index_dim(const ts_index* index);
TS_API int ts_index_dtype(const ts_index* index);
TS_API int ts_index_device(const ts_index* index);

/* Copy stored rows [first, first+n) back, dequantised to fp32, into device buffer
 * out[n, dim]. This is what parity tests feed the oracle ("same inputs"). */
TS_API int ts_index_get_rows(const ts_index* index, int64_t first, int64_t n, float* out,
                             void* stream);

/* Raw device pointer / byte size of the stored corpus (for save/load and diagnostics). */
TS_API const void* ts_index_data(const ts_index* index);
TS_API size_t ts_index_row_bytes(const ts_index* index);

/* ---- raw access: the stored (normalised, quantised, row-padded) bytes, for saving an index and
 * loading it back without re-quantising.  Replaces the reference's on-disk corpus container
 * `torch.save(corpus_embeddings, 'corpus_embeddings.pt')` / `torch.load` (app_create_embeddings.py:85-89,
 * app_showcase_model.py:52) for the quantised form. HOST buffers; both calls synchronise. */
TS_API int ts_index_has_ids(const ts_index* index);
/* rows_out: n * ts_index_row_bytes() bytes, ids_out: int64[n]; either may be NULL. */
TS_API int ts_index_read_raw_host(const ts_index* index, int64_t first, int64_t n, void* rows_out,
                                  int64_t* ids_out);
/* Append n rows that are ALREADY in the index's storage format (as ts_index_read_raw_host returned
 * them); ids: int64[n] or NULL. */
TS_API int ts_index_append_raw_host(ts_index* index, const void* rows, int64_t n, const int64_t* ids);

/* ---- exact search --------------------------------------------------------------------
 * Replaces `util.cos_sim(q, corpus)[0]` + `np.argsort(-scores)[:k]` (test_app.py:75-77),
 * `torch.topk(scores, k=min(200, N), sorted=True)` (app_showcase_model.py:92-96),
 * `util.cos_sim(q_emb, s_emb)` + `np.argsort(-sim_matrix, axis=1)` (compare_embeddings.py:61,105)
 * and pgvector's `ORDER BY e.embedding <#> q LIMIT k` (streamlit_app.py:275-282). */

/* Bytes of device workspace ts_search / ts_search_keys need for this (nq, k). */
TS_API size_t ts_workspace_bytes(const ts_index* index, int nq, int k);

/* Exact top-k. queries: device [nq, dim] of q_dtype (F32 or BF16); normalize_queries != 0
 * L2-normalises them first (cos_sim semantics), 0 takes them as-is (pgvector `<#>` semantics).
 * allow_mask: optional device bitmask, bit r of word r/32 set = row r is eligible (the SQL WHERE
 * of streamlit_app.py:175-243 applied BEFORE ORDER BY/LIMIT); NULL = all rows.
 * out_scores: device float[nq, k]; out_ids: device int64[nq, k] (caller ids, or rows if none).
 * Dispatch: nq == 1 -> K2 bandwidth-bound scan; nq >= 2 -> K3 tcgen05 GEMM (kind::f16 on a bf16 index,
 * kind::tf32 on an fp32 index — the reference's own storage precision, rds_schema.sql:50-53), candidates
 * re-scored and certified so that results are bit-identical to K2's. */
TS_API int ts_search(ts_index* index, const void* queries, int q_dtype, int nq, int k,
                     int normalize_queries, const uint32_t* allow_mask, float* out_scores,
                     int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);

/* Same scan, but stops before the id mapping and returns this GPU's candidates as packed
 * 64-bit keys out_keys[nq, k] (see ts_pack_key): the payload of the cross-GPU all-gather. */
TS_API int ts_search_keys(ts_index* index, const void* queries, int q_dtype, int nq, int k,
                          int normalize_queries, const uint32_t* allow_mask, uint64_t* out_keys,
                          void* workspace, size_t workspace_bytes, void* stream);

/* K5: merge gathered per-shard candidates. keys: device [nshards, nq, k] packed keys, each
 * [k] slice sorted descending (as ts_search_keys writes them). shard_base: device
 * int64[nshards], global row of each shard's row 0 (NULL = zeros). id_map: optional device
 * int64 table indexed by global row (NULL = ids are global rows). Writes the global top-k. */
TS_API int ts_merge_topk(const uint64_t* keys, int nshards, int nq, int k,
                         const int64_t* shard_base, const int64_t* id_map, float* out_scores,
                         int64_t* out_ids, void* stream);

/* ---- sharded exact search with the cross-GPU exchange done by the GPUs themselves ---------------
 * The reference has no sharding; BASELINE.json specifies it: GPU g holds rows
 * [g*N/G, (g+1)*N/G), every GPU scans its shard, the k (score,row) candidates per GPU are exchanged
 * and merged.  ts_search_keys + an NCCL all-gather + ts_merge_topk is the portable form; below the
 * exchange is device-initiated: the shard's k packed keys (rebased to global rows) are stored straight
 * into every peer's receive area over NVLink (CUDA-IPC mapped peer memory), a sequence flag is raised,
 * the peers' flags are awaited (bounded) and the G lists merged — no collective call.
 *   default (two kernels chained by programmatic dependent launch): the scan kernel writes its per-CTA
 *     lists; the exchange kernel — resident early, asleep until the scan completes — merges, exchanges
 *     and writes the result while the scan of the NEXT search already streams the corpus, so a stream
 *     of searches runs at the local scan rate and no launch gap or peer wait sits on its critical path;
 *   TS_SHARDED_ONE_KERNEL: the last CTA of the scan kernel does the exchange itself (one launch per
 *     search, the peers' arrival is awaited inside the scan kernel). */
typedef struct ts_xchg ts_xchg;
enum ts_sharded_flags {
    /* Promise: queries, allow_mask and the corpus are NOT written by the kernel enqueued immediately before
     * this call on `stream` (e.g. the queries were uploaded earlier). Lets the scan start streaming the
     * corpus before that kernel — normally the previous search's exchange kernel — has completed. Without
     * it the scan waits for the preceding kernel first (always correct, loses the overlap). */
    TS_SHARDED_INDEPENDENT = 1,
    TS_SHARDED_ONE_KERNEL = 2
};
/* One per rank (one process per GPU). max_nq / max_k bound the searches that will use it. */
TS_API int ts_xchg_create(ts_xchg** out, int device, int world, int rank, int max_nq, int max_k);
TS_API void ts_xchg_destroy(ts_xchg* xchg);
/* Size of / this rank's IPC handle of its receive area (HOST buffer of ts_xchg_handle_bytes()). */
TS_API int ts_xchg_handle_bytes(void);
TS_API int ts_xchg_handle(const ts_xchg* xchg, void* out_handle);
/* all_handles: HOST buffer of world handles in rank order (gathered by the caller with any host-side
 * collective). Maps every peer's area. world == 1 needs no handles (NULL). */
TS_API int ts_xchg_connect(ts_xchg* xchg, const void* all_handles);
/* Sticky time-out flag: 1 once a search on this rank waited longer than the time-out for a peer's keys
 * (that search's result was written as score = -inf, id = -1 in every position, never a partial merge),
 * 0 otherwise. Lives in host-mapped memory: reading it does NOT synchronise the device, so it is checked
 * at the start of every ts_search_sharded call (TS_ERR_STATE once set). */
TS_API int ts_xchg_error(const ts_xchg* xchg);
/* Time-out of the wait for the peers (default 10 s, or the TS_XCHG_TIMEOUT_MS environment variable). */
TS_API int ts_xchg_set_timeout_ms(ts_xchg* xchg, int64_t ms);
/* Recovery after a time-out or after the ranks' call sequences diverged (one rank raised between calls):
 * synchronises the device, clears this rank's receive area, the sequence counter and the error flag.
 * Collective by contract: every rank calls it between two host-side barriers (nobody may be searching). */
TS_API int ts_xchg_reset(ts_xchg* xchg);
/* Searches issued on this exchange so far (identical on every rank while they are in step). */
TS_API uint32_t ts_xchg_seq(const ts_xchg* xchg);
/* Exact top-k over the GLOBAL corpus: every rank calls it with the same queries, nq, k in the same
 * order. shard_base = global row of this shard's row 0; id_map = optional device table global row ->
 * caller id. Outputs on every rank. flags: ts_sharded_flags. Only the single-query scan path (nq below
 * the batched threshold); larger batches return TS_ERR_UNSUPPORTED (use ts_search_keys + all-gather +
 * ts_merge_topk). */
TS_API int ts_search_sharded(ts_index* index, ts_xchg* xchg, const void* queries, int q_dtype, int nq,
                             int k, int normalize_queries, const uint32_t* allow_mask,
                             int64_t shard_base, const int64_t* id_map, float* out_scores,
                             int64_t* out_ids, void* workspace, size_t workspace_bytes, int flags,
                             void* stream);
/* The host-buffer (end-to-end) form: queries HOST [nq, dim] fp32, outputs HOST [nq, k]. Pinned staging,
 * H2D copy, scan + exchange, D2H copy and the stream synchronise happen inside the call, on a stream and
 * buffers owned by the exchange handle (created on first use). Returns TS_ERR_STATE, with the outputs
 * poisoned, if a peer did not arrive in time. */
TS_API int ts_search_sharded_host(ts_index* index, ts_xchg* xchg, const float* queries, int nq, int k,
                                  int normalize_queries, const uint32_t* allow_mask, int64_t shard_base,
                                  const int64_t* id_map, float* out_scores, int64_t* out_ids);

/* Packed candidate key: high 32 bits = order-preserving image of the fp32 score, low 32 bits
 * = 0xFFFFFFFF - row, so unsigned `max` == "higher score, then lower row". 0 = empty slot. */
TS_API uint64_t ts_pack_key(float score, uint32_t row);
TS_API void ts_unpack_key(uint64_t key, float* score, uint32_t* row);

/* ---- search context: the host-buffer (end-to-end) path --------------------------------
 * What a Streamlit callback actually does (streamlit_app.py:173,284-286): one host query
 * vector in, k rows out.  A ctx owns a stream, pinned staging and workspace sized for
 * (max_nq, max_k); one ctx per host thread makes concurrent searches on one index safe. */
TS_API int ts_ctx_create(ts_ctx** out, ts_index* index, int max_nq, int max_k);
TS_API void ts_ctx_destroy(ts_ctx* ctx);

/* queries: HOST [nq, dim] fp32. out_scores / out_ids: HOST [nq, k]. allow_mask: DEVICE or NULL.
 * H2D copy, search, D2H copy and a stream synchronise all happen inside the call. */
TS_API int ts_search_host(ts_ctx* ctx, const float* queries, int nq, int k, int normalize_queries,
                          const uint32_t* allow_mask, float* out_scores, int64_t* out_ids);

/* Device time (ms, CUDA events on the ctx stream) of the dominant kernel of the last
 * ts_search_host call when timing is enabled with ts_ctx_set_timing(ctx, 1); -1 otherwise. */
TS_API int ts_ctx_set_timing(ts_ctx* ctx, int enabled);
TS_API float ts_ctx_last_kernel_ms(const ts_ctx* ctx);

/* ---- IVF-Flat (pgvector `ivfflat` equivalent; the reference never creates the index,
 * rds_schema.sql has no CREATE INDEX, so production runs the exact scan of
 * streamlit_app.py:275-282 — BASELINE.json config 5 asks for the ANN variant) ----------
 * Semantics restated from pgvector's ivfflat: k-means centroids over a sample, every row
 * filed under its nearest centroid (inner product on unit vectors), a query ranks only the
 * rows of its `nprobe` nearest lists.  The corpus must be stored as TS_BF16. */

/* Spherical k-means (`iters` Lloyd iterations from `nlist` distinct seeded sample rows)
 * -> nlist unit centroids.  sample: device fp32 [n_sample, dim] (normalised internally), or
 * NULL = train on every (size/n_sample)-th STORED row in place (n_sample <= 0: all rows).
 * Kernel K4a (tcgen05 GEMM + fused arg-max) does the assignment step. Synchronises. */
TS_API int ts_ivf_train(ts_index* index, const float* sample, int64_t n_sample, int nlist,
                        int iters, uint64_t seed, void* stream);
/* Install centroids computed elsewhere (device fp32 [nlist, dim], taken as given) — how the
 * ranks of a sharded index share one coarse quantiser (train on rank 0, broadcast). */
TS_API int ts_ivf_set_centroids(ts_index* index, const float* centroids, int nlist, void* stream);
/* Copy the centroids to a device buffer float[nlist, dim]. */
TS_API int ts_ivf_get_centroids(const ts_index* index, float* out, void* stream);
/* Assign every stored row to its nearest centroid (ties -> lower list) and build the
 * inverted lists: rows permuted into (list, ascending row) order, stored as `list_dtype`:
 * TS_BF16 (verbatim) or TS_FP8_E4M3 (e4m3 + one fp32 scale per row, scale = max|x|/448).
 * Synchronises. */
TS_API int ts_ivf_build(ts_index* index, int list_dtype, void* stream);
TS_API size_t ts_ivf_workspace_bytes(const ts_index* index, int nq, int k, int nprobe,
                                     int rescore_k);
/* ANN search: exact top-nprobe centroids per query (K2/K3 over the centroid table), scan of
 * those lists keeping max(k, rescore_k) candidates by list-precision score (K4b), exact
 * re-score of the candidates against the stored bf16 rows with the fp32 query (K4c; returned
 * scores are bit-identical to ts_search's for the same rows), top-k out.  Same output
 * conventions as ts_search. nprobe is clamped to min(nlist, TS_MAX_K). allow_mask: optional device
 * bitmask over corpus rows, applied INSIDE the list scan (pgvector filters after its ivfflat scan
 * and can return fewer than k rows; here the k best ELIGIBLE rows of the probed lists come back). */
TS_API int ts_ivf_search(ts_index* index, const void* queries, int q_dtype, int nq, int k,
                         int nprobe, int rescore_k, int normalize_queries,
                         const uint32_t* allow_mask, float* out_scores, int64_t* out_ids,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Same, returning packed keys out_keys[nq, k] (score, local row): the all-gather payload of
 * the sharded IVF path, merged by ts_merge_topk. */
TS_API int ts_ivf_search_keys(ts_index* index, const void* queries, int q_dtype, int nq, int k,
                              int nprobe, int rescore_k, int normalize_queries,
                              const uint32_t* allow_mask, uint64_t* out_keys, void* workspace,
                              size_t workspace_bytes, void* stream);
/* The host-buffer form of ts_ivf_search on a ts_ctx (see ts_search_host): queries HOST [nq, dim] fp32,
 * outputs HOST [nq, k]; H2D, the four IVF stages, D2H and the synchronise happen inside the call. The ctx
 * grows its IVF workspace the first time a shape needs it. */
TS_API int ts_ivf_search_host(ts_ctx* ctx, const float* queries, int nq, int k, int nprobe, int rescore_k,
                              int normalize_queries, const uint32_t* allow_mask, float* out_scores,
                              int64_t* out_ids);
TS_API int ts_ivf_nlist(const ts_index* index);
/* Rows filed in overflow lists / tombstoned list positions since the last build or re-pack (either may be NULL). */
TS_API int ts_ivf_pending(const ts_index* index, int64_t* overflow_rows, int64_t* dead_positions);
/* Re-pack the lists now (same centroids, every stored row re-filed): clears overflow and tombstones. No-op on
 * pristine lists. Synchronises. */
TS_API int ts_ivf_repack(ts_index* index, void* stream);
/* Storage dtype of the built lists (TS_BF16 / TS_FP8_E4M3), -1 if the lists are not built. */
TS_API int ts_ivf_list_dtype(const ts_index* index);
/* Copy list sizes (int64[nlist]) to a device buffer — for balance diagnostics. */
TS_API int ts_ivf_list_sizes(const ts_index* index, int64_t* out, void* stream);
/* Copy the list layout to device buffers: offsets int64[nlist+1] (list l occupies positions
 * [offsets[l], offsets[l+1])) and rows int64[size] (corpus row at each position). Either may
 * be NULL. What parity tests feed the oracle. */
TS_API int ts_ivf_get_lists(const ts_index* index, int64_t* offsets_out, int64_t* rows_out,
                            void* stream);
/* List rows at positions [first, first+n) dequantised to fp32 into device out[n, dim]. */
TS_API int ts_ivf_get_list_data(const ts_index* index, int64_t first, int64_t n, float* out,
                                void* stream);

/* ---- evaluation metrics from the batched top-k, on the device (K6) -------------------- */
/* Replaces compare_embeddings.py:55-92 `evaluate_retrieval`'s six metrics — precision_at_k :95, hit_at_k :120,
 * mrr_at_k :143, ndcg_at_k :216, err_at_k :257, q_measure_at_k :315 — each of which argsorts the full [Q, N]
 * similarity matrix on the host. Here they are computed from the [Q, k] ids a batched ts_search left in device
 * memory; six doubles come back. */
typedef struct ts_eval ts_eval;   /* the relevance judgements (the reference's `qrels`) resident on one GPU */

/* qrels as CSR over HOST arrays: query q judges docs[offsets[q] .. offsets[q+1]) with relevances rels[...], in the
 * order the reference's {doc: relevance} dict would iterate (the "correct" document of a query is the FIRST one of
 * relevance exactly 1, compare_embeddings.py:111). A query may judge nothing. max_k (1..64): the largest k nDCG
 * will be asked for when some query judges more than max_k documents. TS_ERR_BAD_ARG on a negative doc, a
 * non-finite relevance or a doc judged twice by one query. */
TS_API int ts_eval_create(ts_eval** out, int device, int nq, const int64_t* offsets, const int64_t* docs,
                          const double* rels, int max_k);
TS_API void ts_eval_destroy(ts_eval* ev);
TS_API int ts_eval_num_queries(const ts_eval* ev);
/* Largest relevance judged anywhere (0 if nothing is judged): the default `max_rel` of ERR / Q-measure. */
TS_API double ts_eval_max_relevance(const ts_eval* ev);
/* First query that has no document of relevance exactly 1 (precision / hit / MRR are undefined for it: the
 * reference's `next(...)` raises StopIteration there), or -1. */
TS_API int64_t ts_eval_first_query_without_correct_doc(const ts_eval* ev);

/* ranked: DEVICE int64 [nq, stride], the first `width` entries of a row are doc ids best first, -1 = padding
 * (skipped; ranks count valid entries). k6: HOST int[6], the cut of precision, hit, MRR, nDCG, ERR, Q-measure in
 * that order (<= 0: no cut). gain_exp: nDCG gain 2^rel - 1 (1) or rel (0). max_rel: normaliser 2^max_rel of the
 * ERR / Q-measure gains, < 0 = the table's largest relevance. out_means: DEVICE double[6], the metrics averaged
 * over the nq queries, same order; out_per_query: DEVICE double[nq, 6] or NULL. Enqueues two kernels on `stream`;
 * does not synchronise. */
TS_API int ts_eval_rankings(ts_eval* ev, const int64_t* ranked, int64_t stride, int width, const int* k6,
                            int gain_exp, double max_rel, double* out_means, double* out_per_query, void* stream);

/* ---- tuning / diagnostics (not part of the drop-in surface) -------------------------- */

/* Set an internal tunable by name (e.g. "scan.ctas_per_sm", "scan.stages"); returns
 * TS_ERR_BAD_ARG for unknown names. Used by the bench sweeps only. */
TS_API int ts_set_tunable(const char* name, int value);
TS_API int ts_get_tunable(const char* name, int* value);

/* With tunable "ivf.timeline" = 1 the list-scan CTAs of query 0 record %globaltimer (ns) at 8 phase
 * boundaries; this copies [n_ctas][12] stamps to HOST memory. Synchronises the device. */
TS_API int ts_debug_ivf_timeline(uint64_t* out_host, int n_ctas);

/* With tunable "scan.timeline" = 1 every K2 launch records, per CTA, %globaltimer (ns) at: 0 entry, 1 query
 * ready, 2 first tile landed, 3 warp 0 done, 4 all warps done, 5 CTA list written, 6 (last CTA) final merge
 * done; slot 7 = SM id. Copies the [n_ctas][8] stamps of the launch `launches_back` launches ago (0 = latest,
 * ring of 8) to HOST memory. Synchronises the device. */
TS_API int ts_debug_scan_timeline(uint64_t* out_host, int launches_back, int n_ctas);

/* How many queries of the most recent batched (K3) search failed the exactness certificate or
 * overflowed their candidate buffer and were therefore re-scanned by the exact K2 path.
 * Synchronises the device; -1 if no batched search has run. Valid until the caller frees or
 * reuses that search's workspace. */
TS_API int ts_debug_last_batched_fixups(void);

#ifdef __cplusplus
}
#endif
#endif /* THEOREMSEARCH_H_ */
