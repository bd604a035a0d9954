// K3 — batched exact search: tcgen05/TMEM bf16 tile GEMM with the top-k fused into the epilogue.
//
// Replaces the reference's batched path
//     sim_matrix = util.cos_sim(q_emb, s_emb).cpu().numpy()        (compare_embeddings.py:61)
//     ranked = np.argsort(-sim_matrix, axis=1)                     (compare_embeddings.py:105,...)
// which materialises the full Q x N score matrix on the host and fully sorts every row (six
// times). Here the Q x N scores exist only as 128 x 256 fp32 accumulator tiles in TMEM.
//
// GEMM:   D[128 corpus rows, 256 queries] = A[128, K] * B[256, K]^T   (both K-major, bf16, fp32 acc)
//   * A = corpus tile, B = query tile; TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages
//     64-element k-blocks of both through a 4-deep smem ring guarded by full/empty mbarriers;
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128 N=256 K=16) into one of
//     two 256-column TMEM accumulators, tcgen05.commit releases smem slots / publishes the tile;
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
//
// Fused top-k epilogue: thread <-> corpus row (TMEM lane), columns <-> queries. Each thread reads
// 32 columns at a time with tcgen05.ld and compares them with the per-query running thresholds
// (k-th best score so far) held in smem; only survivors — rare once thresholds are warm — are
// appended (one atomicAdd + one 8-byte store) to that query's candidate buffer in HBM.
//
// Thresholds are refreshed between launches: the corpus is processed in geometrically growing row
// chunks; after each chunk K3b (`compact_candidates_kernel`, one CTA per query, bitonic sort in
// smem) folds the appended candidates into the query's sorted top-k and publishes the new
// threshold. With chunk sizes growing 4x the expected number of survivors per query per chunk is
// ~3k, independent of N. If a buffer ever overflows (adversarial row order) the query is flagged
// and K3c re-scans it exactly with the K2 path, so the result is exact in all cases.
//
// Algorithmic FLOPs: 2 * Q * N * D per batch.
#include <algorithm>

#include "merge_device.cuh"
#include "scan_device.cuh"
#include "umma_device.cuh"

namespace ts {

namespace k3 {
using namespace umma;            // BM = 128 corpus rows, BN = 256 queries, BK = 64, UMMA_K = 16
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int STG_CAP = 256;           // survivors an epilogue warp stages in smem before flushing to HBM
constexpr size_t SMEM_TILES = (size_t)STAGES * (A_BYTES + B_BYTES);
constexpr size_t SMEM_STAGING = 4 * (size_t)STG_CAP * (sizeof(uint64_t) + sizeof(uint32_t));
constexpr size_t SMEM_BYTES =
    SMEM_TILES + ACC_STAGES * BN * sizeof(float) + 18 * sizeof(uint64_t) + SMEM_STAGING + 1024;
}  // namespace k3

struct BatchParams {
    int64_t row_begin, row_end;   // corpus rows of this chunk
    int nq, k, cap;
    int num_k_blocks;
    int num_m_blocks, num_n_blocks;
    int a_policy;                 // L2 policy of the corpus-tile loads: 0 = evict_first, 1 = evict_normal, 2 = evict_last
    int bn;                       // queries per n-block actually used (multiple of 32, <= BN): UMMA N, TMA box rows
    const float* thr;             // [num_n_blocks * bn] running k-th best score per query (-inf initially)
    uint32_t* count;              // [nq] candidates appended in this chunk
    uint64_t* cand;               // [nq][k + cap]: [0,k) sorted best so far, [k, k+cap) appended
    const uint32_t* mask;         // allow bitmask or nullptr
    float* dense;                 // small tables: write EVERY score to dense[query][row] instead of filtering
    int64_t dense_stride;         // elements between consecutive queries in `dense`
};

// ---------------------------------------------------------------------------------- K3 kernel
// PAIR = false: one CTA per tile, tcgen05 cta_group::1, tile = 128 corpus rows x bn queries.
// PAIR = true : a cluster of two CTAs (one TPC) per tile, cta_group::2, tile = 256 corpus rows x bn
//               queries. Each CTA stages its own 128 corpus rows and HALF of the query block; the
//               tensor cores of both SMs read both halves, so operand traffic from L2 per SM drops
//               from (128 + bn) to (128 + bn/2) rows per k-block. CTA rank 0 (leader) issues the MMAs;
//               completion is multicast to both CTAs' mbarriers; each CTA runs the epilogue on the 128
//               accumulator rows that live in its own TMEM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
    // bit 24 of a shared::cluster address selects the odd CTA of the pair: clearing it makes the
    // transaction bytes land on the LEADER's barrier from either CTA
    const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {   // arrive on this barrier in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}

// TF32 = true: both operands are fp32 rows (an fp32-stored corpus — the reference's own precision, pg
//               `vector(1024)`, rds_schema.sql:50-53 — and the fp32 queries themselves), multiplied as TF32
//               (tcgen05 kind::tf32, K = 8 per instruction). A k-block is still one 128-byte swizzle span per row
//               (32 fp32 instead of 64 bf16), so the smem ring, the descriptors and the expect-tx bytes are unchanged.
template <bool PAIR, bool TF32>
__global__ void __launch_bounds__(k3::THREADS, 1)
batched_gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const BatchParams p) {
    using namespace k3;
    static_assert(!(PAIR && TF32), "the CTA-pair kernel is bf16 only");
    constexpr int KB_ELEMS = TF32 ? BK / 2 : BK;           // elements per k-block (one 128-byte span per row)
    constexpr int NST = PAIR ? 6 : STAGES;                 // smem ring depth
    constexpr int BSLOT = PAIR ? B_BYTES / 2 : B_BYTES;    // bytes reserved per stage for this CTA's part of B
    constexpr int TILE_ROWS = PAIR ? 2 * BM : BM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)NST * A_BYTES;
    static_assert((size_t)NST * (A_BYTES + BSLOT) == SMEM_TILES, "ring must fill the tile area exactly");
    float* thr_s = reinterpret_cast<float*>(smem + SMEM_TILES);                   // [ACC_STAGES][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_s + ACC_STAGES * BN);
    uint64_t* full = bars;                    // [NST]   (PAIR: only the leader's are used)
    uint64_t* empty = bars + 6;               // [NST]
    uint64_t* tmem_full = bars + 12;          // [ACC_STAGES]
    uint64_t* tmem_empty = bars + 14;         // [ACC_STAGES] (PAIR: only the leader's are used)
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 16);
    uint64_t* stg_keys_all = bars + 18;                                             // [4][STG_CAP]
    uint32_t* stg_q_all = reinterpret_cast<uint32_t*>(stg_keys_all + 4 * STG_CAP);  // [4][STG_CAP]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // tile-scheduling unit (CTA or CTA pair)
    const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;

    if constexpr (PAIR) cluster_sync_all();   // both CTAs resident before the paired TMEM allocation
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], PAIR ? 2 * EPI_THREADS : EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {  // TMEM owner: all 512 columns (1 CTA per SM by launch bounds + smem)
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                         "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                         "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == 0) {
        // ===================== TMA producer (every CTA loads its own rows / its half of B) =====================
        if (lane == 0) {
            // corpus tile: streamed once per n-block; the other n-blocks of the same rows read it within microseconds
            const uint64_t pol_a = p.a_policy == 0 ? l2_policy_evict_first() : (p.a_policy == 1 ? l2_policy_evict_normal() : l2_policy_evict_last());
            const uint64_t pol_b = l2_policy_evict_last();   // query block: re-read by every corpus tile
            const uint32_t b_rows = PAIR ? (uint32_t)p.bn / 2 : (uint32_t)p.bn;
            int stage = 0;
            uint32_t phase = 0;
            for (int t = unit; t < num_tiles; t += nunits) {
                const int m_blk = t / p.num_n_blocks, n_blk = t % p.num_n_blocks;
                const int row0 = (int)(p.row_begin + (int64_t)m_blk * TILE_ROWS + (int64_t)cta_rank * BM);
                const int b0 = n_blk * p.bn + (int)(cta_rank * b_rows);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait_wd(&empty[stage], phase ^ 1u);
                    if constexpr (PAIR) {
                        if (leader) mbar_expect_tx(&full[stage], 2u * (A_BYTES + b_rows * (BK * 2)));
                        tma_load_2d_pair(smem_a + (size_t)stage * A_BYTES, &tmap_a, kb * KB_ELEMS, row0, &full[stage], pol_a);
                        tma_load_2d_pair(smem_b + (size_t)stage * BSLOT, &tmap_b, kb * KB_ELEMS, b0, &full[stage], pol_b);
                    } else {
                        mbar_expect_tx(&full[stage], A_BYTES + b_rows * (BK * 2));
                        tma_load_2d(smem_a + (size_t)stage * A_BYTES, &tmap_a, kb * KB_ELEMS, row0, &full[stage], pol_a);
                        tma_load_2d(smem_b + (size_t)stage * BSLOT, &tmap_b, kb * KB_ELEMS, b0, &full[stage], pol_b);
                    }
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread; in PAIR mode only in the leader CTA) =====================
        if (lane == 0 && leader) {
            // instruction descriptor: D=f32 at [4,6), A / B format at [7,10) / [10,13) (1 = bf16, 2 = tf32), both
            // K-major, N>>3 at [17,23), M>>4 at [24,29)
            constexpr uint32_t FMT = TF32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.bn >> 3) << 17) |
                                   ((uint32_t)(TILE_ROWS >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = unit; t < num_tiles; t += nunits, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                mbar_wait_wd(&tmem_empty[acc], acc_phase ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait_wd(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint64_t da = umma_smem_desc(smem_a + (size_t)stage * A_BYTES);
                    const uint64_t db = umma_smem_desc(smem_b + (size_t)stage * BSLOT);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // +32 bytes per UMMA_K step inside the 128-byte swizzle span (address field is >>4)
                        if constexpr (PAIR)
                            umma_bf16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else if constexpr (TF32)
                            umma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else
                            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    if constexpr (PAIR) {
                        tcgen05_commit_pair(&empty[stage]);  // smem slot reusable in BOTH CTAs once these MMAs retire
                        if (kb == p.num_k_blocks - 1) tcgen05_commit_pair(&tmem_full[acc]);
                    } else {
                        tcgen05_commit(&empty[stage]);
                        if (kb == p.num_k_blocks - 1) tcgen05_commit(&tmem_full[acc]);
                    }
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: threshold filter + append =====================
        // Survivors are staged per warp in smem with ballot/popc slot assignment (no atomics, no
        // divergence-serialised round trips) and flushed to the per-query HBM buffers 32 at a
        // time, so the global atomicAdd latencies overlap instead of adding up.
        const int ep_tid = threadIdx.x - 64;              // 0..127
        const int ew = warp - 2;                          // epilogue warp 0..3
        const int lane_base = 32 * (warp & 3);            // TMEM lanes this warp may touch
        const int row_in_tile = (int)cta_rank * BM + lane_base + lane;
        const size_t cand_stride = (size_t)p.k + p.cap;
        uint64_t* stg_keys = stg_keys_all + ew * STG_CAP;
        uint32_t* stg_q = stg_q_all + ew * STG_CAP;
        const uint32_t lt_mask = (1u << lane) - 1u;
        int staged = 0;                                   // warp-uniform
        auto flush = [&]() {
            __syncwarp();
            for (int e = lane; e < staged; e += 32) {
                const uint32_t q = stg_q[e];
                const uint32_t pos = atomicAdd(p.count + q, 1u);
                if (pos < (uint32_t)p.cap) p.cand[(size_t)q * cand_stride + p.k + pos] = stg_keys[e];
            }
            __syncwarp();
            staged = 0;
        };
        int it = 0;
        for (int t = unit; t < num_tiles; t += nunits, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const int m_blk = t / p.num_n_blocks, n_blk = t % p.num_n_blocks;
            const int q0 = n_blk * p.bn;
            float* thr_t = thr_s + acc * BN;
            if (ep_tid < p.bn) thr_t[ep_tid] = p.thr[q0 + ep_tid];
            if (ep_tid + EPI_THREADS < p.bn) thr_t[ep_tid + EPI_THREADS] = p.thr[q0 + ep_tid + EPI_THREADS];
            asm volatile("bar.sync 1, 128;" ::: "memory");

            const int64_t row = p.row_begin + (int64_t)m_blk * TILE_ROWS + row_in_tile;
            bool row_ok = row < p.row_end;
            if (row_ok && p.mask != nullptr) row_ok = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;

            mbar_wait_wd(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c = 0; c < p.bn / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                if (p.dense != nullptr) {
                    // small corpus (e.g. an IVF centroid table): no threshold exists yet that would filter anything,
                    // so the tile goes to a dense score matrix — for a fixed query the warp's 32 rows are 32
                    // consecutive floats (one coalesced 128-byte store per column) — and a select kernel follows
                    if (row < p.row_end) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int q = q0 + c * 32 + j;
                            if (q < p.nq) p.dense[(int64_t)q * p.dense_stride + row] = row_ok ? __uint_as_float(v[j]) : -INFINITY;
                        }
                    }
                    continue;
                }
                const float4* th4 = reinterpret_cast<const float4*>(thr_t + c * 32);
                bool any = false;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 th = th4[j];
                    any |= (__uint_as_float(v[4 * j + 0]) > th.x) | (__uint_as_float(v[4 * j + 1]) > th.y) |
                           (__uint_as_float(v[4 * j + 2]) > th.z) | (__uint_as_float(v[4 * j + 3]) > th.w);
                }
                if (__any_sync(0xFFFFFFFFu, any && row_ok)) {
                    // rare path. Per-lane 32-bit pass mask (straight-line), OR-reduced over the warp to the
                    // set of columns that have a survivor anywhere; each such column is re-read from TMEM
                    // (one register per lane, no dynamic register indexing) and compacted with a ballot.
                    unsigned pm = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        pm |= (__uint_as_float(v[j]) > thr_t[c * 32 + j]) ? (1u << j) : 0u;   // padded queries: thr = +inf
                    if (!row_ok) pm = 0;
                    unsigned cols = __reduce_or_sync(0xFFFFFFFFu, pm);
                    while (cols) {
                        const int j = __ffs(cols) - 1;
                        cols &= cols - 1;
                        const float sc = __uint_as_float(tmem_ld_32x32b_x1(taddr0 + (uint32_t)(c * 32 + j)));
                        tmem_ld_wait();
                        const bool pass = (pm >> j) & 1u;
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
                        if (pass) {
                            const int slot = staged + __popc(m & lt_mask);
                            stg_keys[slot] = pack_key(sc, (uint32_t)row);
                            stg_q[slot] = (uint32_t)(q0 + c * 32 + j);
                        }
                        staged += __popc(m);
                        if (staged > STG_CAP - 32) flush();
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back (to the leader's MMA thread) before any HBM work
            tcgen05_fence_before();
            if (PAIR && !leader)
                mbar_arrive_remote(&tmem_empty[acc], 0);
            else
                mbar_arrive(&tmem_empty[acc]);
            if (staged >= STG_CAP / 2) flush();
        }
        flush();
    }

    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------- K3b compaction
// One CTA per query: sort [best k | appended candidates] descending (bitonic, smem), keep k,
// publish the new threshold, reset the append counter, flag overflow.
__global__ void __launch_bounds__(256) compact_candidates_kernel(uint64_t* cand, uint32_t* count, float* thr,
                                                                 uint32_t* overflow, int k, int cap) {
    extern __shared__ uint64_t sort_buf[];
    const int q = blockIdx.x;
    const size_t stride = (size_t)k + cap;
    uint64_t* mine = cand + (size_t)q * stride;
    uint32_t cnt = count[q];
    if (cnt > (uint32_t)cap) {
        if (threadIdx.x == 0) atomicOr(overflow + q, 1u);
        cnt = cap;
    }
    const int n = k + (int)cnt;
    if (cnt == 0) {  // nothing appended in this chunk: list and threshold stand
        return;
    }
    int P = 64;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) sort_buf[i] = (i < n) ? mine[i] : 0ull;
    __syncthreads();
    cta_bitonic_sort_desc(sort_buf, P);   // register-blocked: strides <= 32 in shuffles, barriers only above
    for (int i = threadIdx.x; i < k; i += blockDim.x) mine[i] = sort_buf[i];
    if (threadIdx.x == 0) {
        const uint64_t kth = sort_buf[k - 1];
        thr[q] = kth ? key_score(kth) : -INFINITY;
        count[q] = 0;
    }
}

__global__ void init_batch_state_kernel(uint64_t* cand, uint32_t* count, float* thr, uint32_t* overflow, int nq,
                                        int nq_pad, int k, int cap) {
    const size_t stride = (size_t)k + cap;
    const size_t total = (size_t)nq * k;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t q = i / k, j = i % k;
        cand[q * stride + j] = 0ull;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq_pad; i += gridDim.x * blockDim.x) {
        thr[i] = (i < nq) ? -INFINITY : INFINITY;   // padded queries can never pass the filter
        if (i < nq) {
            count[i] = 0;
            overflow[i] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------- query rounding
// q32 (normalised fp32, what K2 dots with) -> bf16 copy for the tensor cores, plus
// qerr[q], a bound (per unit of ||c||) on how far a row's exact-path score can sit above its GEMM score:
//   ||q32 - bf16(q32)||_2        Cauchy-Schwarz: |<q32,c> - <q16,c>| <= that * ||c||
// + 2 * dim_pad * 2^-23 * ||q16||  fp32 accumulation of dim_pad products, in whatever order the tensor core uses and
//                                allowing truncation instead of rounding (error <= n * 2^-23 * sum|q_i c_i| <=
//                                n * 2^-23 * ||q|| ||c||), once for the excluded row's GEMM score and once for the
//                                threshold score it was compared with.
__global__ void __launch_bounds__(256) round_queries_kernel(const float* __restrict__ q32, int nq, int dim_pad,
                                                            __nv_bfloat16* __restrict__ q16, float* __restrict__ qerr) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    float e2 = 0.f, n2 = 0.f;
    for (int i = lane; i < dim_pad; i += 32) {
        const float v = q32[(size_t)q * dim_pad + i];
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        q16[(size_t)q * dim_pad + i] = b;
        const float bf = __bfloat162float(b);
        const float d = v - bf;
        e2 = fmaf(d, d, e2);
        n2 = fmaf(bf, bf, n2);
    }
    e2 = warp_sum(e2);
    n2 = warp_sum(n2);
    if (lane == 0)
        qerr[q] = sqrtf(e2) * 1.0001f + 2.0f * (float)dim_pad * 1.1920929e-7f * sqrtf(n2) * 1.0001f + 1e-30f;
}

// fp32 corpus (TF32 GEMM): the queries go to the tensor core as they are; the tensor core drops the low 13
// mantissa bits of BOTH operands (relative error < 2^-10 each), so |<q,c> - tf32 score| <= (2^-9 + 2^-20) ||q|| ||c||,
// plus the same accumulation allowance as above.
__global__ void __launch_bounds__(256) tf32_query_error_kernel(const float* __restrict__ q32, int nq, int dim_pad,
                                                               float* __restrict__ qerr) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    float n2 = 0.f;
    for (int i = lane; i < dim_pad; i += 32) {
        const float v = q32[(size_t)q * dim_pad + i];
        n2 = fmaf(v, v, n2);
    }
    n2 = warp_sum(n2);
    if (lane == 0)
        qerr[q] = sqrtf(n2) * 1.0001f * (1.9550e-3f + 2.0f * (float)dim_pad * 1.1920929e-7f) + 1e-30f;
}

// ---------------------------------------------------------------------------------- K3c rescore + certify
// One CTA per query. The GEMM ranked candidates with bf16-ROUNDED queries; the exact path (K2)
// scores with the fp32 query. Re-score the kp best candidates with the fp32 query using K2's exact
// summation order (so both paths return bit-identical scores), sort, keep k, and CERTIFY the
// result: every row the GEMM left out has bf16-query score <= b (the kp-th kept one), hence
// fp32-query score <= b + qerr*max||row||; if the k-th rescored score beats that bound the top-k is
// provably the exact one, otherwise (or if a candidate buffer overflowed) the query is flagged and
// re-scanned by K2.
template <int ELEM>   // bytes per stored corpus element: 2 = bf16 rows, 4 = fp32 rows
__global__ void __launch_bounds__(256) rescore_certify_kernel(uint64_t* cand, size_t cand_stride, int kp, int k,
                                                              const float* __restrict__ q32,
                                                              const uint8_t* __restrict__ corpus, uint32_t row_bytes,
                                                              int dim_pad, const float* __restrict__ qerr,
                                                              const float* __restrict__ max_norm2,
                                                              const uint32_t* __restrict__ overflow,
                                                              int* __restrict__ flags) {
    extern __shared__ uint64_t rs_buf[];   // [P] rescored keys
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* mine = cand + (size_t)q * cand_stride;
    int P = 64;
    while (P < kp) P <<= 1;
    const uint64_t worst_kept = mine[kp - 1];   // 0 when fewer than kp rows were eligible at all
    const float* qv = q32 + (size_t)q * dim_pad;
    for (int j = warp; j < P; j += 8) {
        uint64_t key = (j < kp) ? mine[j] : 0ull;
        if (key != 0ull) {
            const uint32_t row = key_row(key);
            const uint8_t* r = corpus + (size_t)row * row_bytes;
            float acc = 0.f;
            for (uint32_t off = (uint32_t)lane * 16u; off < row_bytes; off += 512u) {   // K2's chunk order
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(r + off));
                if constexpr (ELEM == 4) {
                    const float4 qa = __ldg(reinterpret_cast<const float4*>(qv + off / 4));
                    acc = fmaf(__uint_as_float(v.x), qa.x, acc);
                    acc = fmaf(__uint_as_float(v.y), qa.y, acc);
                    acc = fmaf(__uint_as_float(v.z), qa.z, acc);
                    acc = fmaf(__uint_as_float(v.w), qa.w, acc);
                    continue;
                }
                const float4 qa = __ldg(reinterpret_cast<const float4*>(qv + off / 2));
                const float4 qb = __ldg(reinterpret_cast<const float4*>(qv + off / 2 + 4));
                acc = fmaf(__uint_as_float(v.x << 16), qa.x, acc);
                acc = fmaf(__uint_as_float(v.x & 0xFFFF0000u), qa.y, acc);
                acc = fmaf(__uint_as_float(v.y << 16), qa.z, acc);
                acc = fmaf(__uint_as_float(v.y & 0xFFFF0000u), qa.w, acc);
                acc = fmaf(__uint_as_float(v.z << 16), qb.x, acc);
                acc = fmaf(__uint_as_float(v.z & 0xFFFF0000u), qb.y, acc);
                acc = fmaf(__uint_as_float(v.w << 16), qb.z, acc);
                acc = fmaf(__uint_as_float(v.w & 0xFFFF0000u), qb.w, acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);   // == K2's reduce tree
            key = pack_key(acc, row);
        }
        if (lane == 0) rs_buf[j] = key;
    }
    __syncthreads();
    cta_bitonic_sort_desc(rs_buf, P);
    for (int i = threadIdx.x; i < k; i += blockDim.x) mine[i] = rs_buf[i];
    if (threadIdx.x == 0) {
        bool ok = overflow[q] == 0;
        if (ok && worst_kept != 0ull) {
            const uint64_t kth = rs_buf[k - 1];
            const float bound = key_score(worst_kept) + qerr[q] * sqrtf(*max_norm2) + 2e-6f;   // + the exact path's own rounding
            ok = kth != 0ull && key_score(kth) > bound;   // NaN bound -> not certified -> re-scan
        }
        flags[q] = ok ? 0 : 1;
    }
}

// Compact the flagged queries into a work list for the K2 fix-up pass (single CTA, ballot scan).
__global__ void __launch_bounds__(1024) build_fix_list_kernel(const int* __restrict__ flags, int nq,
                                                              int* __restrict__ list, int* __restrict__ count) {
    __shared__ int base;
    __shared__ int warp_tot[32];
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < nq; start += blockDim.x) {
        const int q = start + threadIdx.x;
        const bool f = q < nq && flags[q] != 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (f) list[off + __popc(m & ((1u << lane) - 1u))] = q;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// ---------------------------------------------------------------------------------- dense select (small corpora)
// One CTA per query: the best kp keys of its dense score row. The CTA walks the row 1024 scores per step; keys above
// the running threshold are appended to a shared buffer (warp-aggregated slots) which is compacted with a cooperative
// sort (keep kp, raise the threshold) when the next step could overflow it. Output: cand[q][0..kp) sorted descending.
// Warp-per-query form for kp <= 256 (the IVF coarse step: top-96 of 16384 centroid scores per query): the row is
// walked with float4 loads against a float threshold, candidates go through a register-resident WarpSelect — no
// shared memory, no barriers, no CTA sorts. 4096 queries x 16384 scores: 0.97 ms with the CTA-per-query kernel below.
template <int KPL>
__global__ void __launch_bounds__(256, 4) dense_select_warp_kernel(const float* __restrict__ dense, int64_t stride, int64_t n_rows,
                                                                int nq, int kp, uint64_t* __restrict__ cand, size_t cand_stride,
                                                                uint32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    if (lane == 0) overflow[q] = 0u;
    WarpSelect<KPL> sel;
    sel.init();
    const float* row = dense + (int64_t)q * stride;
    for (int64_t c0 = 0; c0 < n_rows; c0 += (int64_t)1 << 30)     // (rows <= 2^18 here; the loop keeps `len` an int)
        warp_select_run<KPL>(sel, row + c0, (int)std::min<int64_t>(n_rows - c0, (int64_t)1 << 30), (uint32_t)c0, kp, lane);
    sel.flush(kp, lane);
    uint64_t* mine = cand + (size_t)q * cand_stride;
#pragma unroll
    for (int j = 0; j < KPL; ++j)
        if (j * 32 + lane < kp) mine[j * 32 + lane] = sel.best.key[j];
}

constexpr int DSEL_CAP = 3072;
__global__ void __launch_bounds__(256) dense_select_kernel(const float* __restrict__ dense, int64_t stride, int64_t n_rows,
                                                           int kp, uint64_t* __restrict__ cand, size_t cand_stride,
                                                           uint32_t* __restrict__ overflow) {
    __shared__ __align__(16) uint64_t buf[4096];
    __shared__ int s_cnt;
    __shared__ unsigned long long s_thr;
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float* row = dense + (int64_t)q * stride;
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_thr = 0ull;
        overflow[q] = 0u;
    }
    __syncthreads();
    auto compact = [&]() {
        const int n = s_cnt;
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = n + threadIdx.x; i < P; i += blockDim.x) buf[i] = 0ull;
        __syncthreads();
        cta_bitonic_sort_desc(buf, P);
        if (threadIdx.x == 0) {
            s_cnt = n < kp ? n : kp;
            s_thr = n >= kp ? buf[kp - 1] : 0ull;
        }
        __syncthreads();
    };
    for (int64_t r0 = 0; r0 < n_rows; r0 += 1024) {
        const unsigned long long thr = s_thr;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + u * 256 + threadIdx.x;   // coalesced 4-byte loads (the row stride need not be aligned)
            const float f = (r < n_rows) ? __ldg(row + r) : -INFINITY;
            const uint64_t key = f > -INFINITY ? pack_key(f, (uint32_t)r) : 0ull;
            const bool pass = key > thr;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (pass) buf[base + __popc(m & lt_mask)] = key;
            }
        }
        __syncthreads();
        if (s_cnt > DSEL_CAP - 1024) compact();
    }
    compact();
    uint64_t* mine = cand + (size_t)q * cand_stride;
    for (int i = threadIdx.x; i < kp; i += blockDim.x) mine[i] = (i < s_cnt) ? buf[i] : 0ull;
}

// Small corpora take the dense path: the select kernel walks a query's whole score row with ONE CTA, so the
// row must be short (2^18 scores = 1 MB, ~10 us) whatever nq is — at 10M rows and a handful of queries the
// matrix would fit the byte budget but the select would take 20 ms (measured) against 3.3 ms chunked.
constexpr int64_t DENSE_MAX_ROWS = (int64_t)1 << 18;
static bool dense_eligible(const ts_index* ix, int nq) {
    return ix->size > 0 && ix->size <= DENSE_MAX_ROWS && (size_t)nq * (size_t)ix->size * 4 <= ((size_t)1 << 30);
}
static size_t dense_stride_of(const ts_index* ix) { return ((size_t)ix->size + 31) / 32 * 32; }

// ---------------------------------------------------------------------------------- host side
// candidates kept per query by the GEMM stage: k plus a margin that makes the exactness
// certificate succeed (the gap between the k-th and kp-th score must exceed the bf16 query
// rounding bound ~1.1e-3; doubling k gives ~5e-3 on 10M random unit rows).
static int batched_kp(int k) { return std::max(2 * k, k + 64); }
// candidate slots per query per chunk: the tunable, raised so that it holds >= 3 chunks' worth of
// expected survivors for large k, within the 8192-key compaction sort.
static int batched_cap(int kp) { return std::min(std::max(tunables().batch_cap, 3 * kp), 8192 - kp); }
// chunk growth: expected survivors per query per chunk ~ growth * kp must stay well below cap.
static int batched_growth(int kp, int cap) { return std::max(1, std::min(tunables().batch_growth, cap / (2 * kp))); }

// diagnostics: where the last batched search left its fix-up count (device memory in the caller's workspace)
static const int* g_last_fix_count = nullptr;
static int g_last_fix_device = 0;
int debug_last_batched_fixups() {
    if (!g_last_fix_count) return -1;
    int prev = 0, v = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g_last_fix_device);
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(&v, g_last_fix_count, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        v = -1;
    cudaSetDevice(prev);
    return v;
}

size_t batched_workspace_bytes(const ts_index* ix, int nq, int k) {
    const int nq_pad = (nq + k3::BN - 1) / k3::BN * k3::BN + k3::BN;
    const int kp = batched_kp(k);
    const int nparts = scan_nparts(ix);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t b = 0;
    b += al((size_t)nq * ix->dim_pad * 4);                   // fp32 normalised queries
    b += al((size_t)nq * ix->dim_pad * 2);                   // bf16 queries
    b += al((size_t)nq_pad * 4) * 4;                         // thr, count, overflow, qerr
    b += al((size_t)nq * 4) * 2 + 256;                       // flags, fix list, fix count
    b += al((size_t)nq * ((size_t)kp + batched_cap(kp)) * 8);  // candidates
    b += al((size_t)nq * nparts * k * 8);                    // K2 fix-up partial lists (worst case: every query)
    if (dense_eligible(ix, nq)) b += al((size_t)nq * dense_stride_of(ix) * 4);   // dense scores of a small corpus
    return b;
}

int launch_batched_search(const ts_index* ix, const void* queries, int q_dtype, int nq, int k, int normalize,
                          const uint32_t* allow_mask, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                          void* workspace, size_t workspace_bytes, cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1) {
    using namespace k3;
    TS_REQUIRE(ix->dtype == TS_BF16 || ix->dtype == TS_F32, TS_ERR_UNSUPPORTED, "batched: corpus must be stored as bf16 or fp32");
    const bool f32 = ix->dtype == TS_F32;     // fp32 rows: TF32 GEMM on the rows and the fp32 queries themselves
    const Tunables& t = tunables();
    const int k_out = k;          // what the caller asked for
    k = batched_kp(k_out);        // what the GEMM stage keeps per query (kp)
    const int cap = batched_cap(k);
    const int growth = batched_growth(k, cap);
    const int nparts = scan_nparts(ix);
    TS_REQUIRE(workspace_bytes >= batched_workspace_bytes(ix, nq, k_out), TS_ERR_CAPACITY,
               "batched: workspace %zu < %zu bytes", workspace_bytes, batched_workspace_bytes(ix, nq, k_out));
    const int nq_pad = (nq + BN - 1) / BN * BN + BN;
    // n-blocks: as few as possible, equal width, width a multiple of 32 (UMMA N % 16, epilogue reads 32 columns)
    const int num_n_blocks = (nq + BN - 1) / BN;
    const int bn = ((nq + num_n_blocks - 1) / num_n_blocks + 31) / 32 * 32;
    char* w = (char*)workspace;
    auto take = [&](size_t bytes) {
        char* p = w;
        w += (bytes + 255) / 256 * 256;
        return p;
    };
    float* q32 = (float*)take((size_t)nq * ix->dim_pad * 4);
    __nv_bfloat16* q16 = (__nv_bfloat16*)take((size_t)nq * ix->dim_pad * 2);
    float* thr = (float*)take((size_t)nq_pad * 4);
    uint32_t* count = (uint32_t*)take((size_t)nq_pad * 4);
    uint32_t* overflow = (uint32_t*)take((size_t)nq_pad * 4);
    float* qerr = (float*)take((size_t)nq_pad * 4);
    int* flags = (int*)take((size_t)nq * 4);
    int* fix_list = (int*)take((size_t)nq * 4);
    int* fix_count = (int*)take(256);
    uint64_t* cand = (uint64_t*)take((size_t)nq * ((size_t)k + cap) * 8);
    uint64_t* fix_parts = (uint64_t*)take((size_t)nq * nparts * k_out * 8);
    const bool dense = t.batch_dense != 0 && dense_eligible(ix, nq);
    float* dscores = dense ? (float*)take((size_t)nq * dense_stride_of(ix) * 4) : nullptr;

    int rc = launch_prepare_queries(queries, q_dtype, nq, ix->dim, ix->dim_pad, normalize, q32, s);
    if (rc) return rc;
    if (f32) tf32_query_error_kernel<<<(nq + 7) / 8, 256, 0, s>>>(q32, nq, ix->dim_pad, qerr);
    else round_queries_kernel<<<(nq + 7) / 8, 256, 0, s>>>(q32, nq, ix->dim_pad, q16, qerr);
    TS_LAUNCH_CHECK();
    init_batch_state_kernel<<<256, 256, 0, s>>>(cand, count, thr, overflow, nq, nq_pad, k, cap);
    TS_LAUNCH_CHECK();

    CUtensorMap tmap_a, tmap_b;
    rc = make_tmap_rows(&tmap_a, ix->data, (uint64_t)ix->size, (uint64_t)ix->dim_pad, ix->row_bytes(), BM, f32);
    if (rc) return rc;
    // CTA pairs pay off once the batch is tensor-bound; small batches (HBM-bound) keep single CTAs, which
    // spread the corpus stream over all 148 SMs' TMA queues.
    const bool pair = t.batch_cta_pair != 0 && nq >= t.batch_pair_min_nq && !dense && !f32;   // the dense pass uses single CTAs
    if (f32)
        rc = make_tmap_rows(&tmap_b, q32, (uint64_t)nq, (uint64_t)ix->dim_pad, (uint64_t)ix->dim_pad * 4, bn, true);
    else
        rc = make_tmap_rows(&tmap_b, q16, (uint64_t)nq, (uint64_t)ix->dim_pad, (uint64_t)ix->dim_pad * 2,
                            pair ? bn / 2 : bn, false);
    if (rc) return rc;

    static bool attr_set[64] = {false};   // function attributes are per device
    const int dslot = ix->device & 63;
    if (!attr_set[dslot]) {
        TS_CHECK_CUDA(cudaFuncSetAttribute(batched_gemm_topk_kernel<false, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        TS_CHECK_CUDA(cudaFuncSetAttribute(batched_gemm_topk_kernel<true, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        TS_CHECK_CUDA(cudaFuncSetAttribute(batched_gemm_topk_kernel<false, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        TS_CHECK_CUDA(cudaFuncSetAttribute(compact_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           64 * 1024));
        attr_set[dslot] = true;
    }
    int P = 64;
    while (P < k + cap) P <<= 1;
    const size_t sort_smem = (size_t)P * 8;
    TS_REQUIRE(sort_smem <= 64 * 1024, TS_ERR_UNSUPPORTED, "batched: kp + cap = %d too large for the compaction sort", k + cap);

    BatchParams p;
    p.nq = nq;
    p.k = k;
    p.cap = cap;
    p.num_k_blocks = f32 ? (ix->dim_pad + BK / 2 - 1) / (BK / 2) : (ix->dim_pad + BK - 1) / BK;
    p.num_n_blocks = num_n_blocks;
    p.bn = bn;
    p.thr = thr;
    p.count = count;
    p.cand = cand;
    p.mask = allow_mask;
    p.dense = nullptr;
    p.dense_stride = 0;
    p.a_policy = num_n_blocks > 1 ? t.batch_a_policy : 0;   // one n-block: every corpus tile is read exactly once
    const int sms = sm_count(ix->device);

    // chunk schedule: first chunk fills the buffers (every row passes thr = -inf), then chunks
    // grow by `batch_growth` x the rows already seen, so expected survivors per query stay ~growth*k.
    int64_t first = std::min<int64_t>((int64_t)(cap / 2) / (2 * BM) * (2 * BM),
                                      (int64_t)t.batch_first_chunk / (2 * BM) * (2 * BM));
    if (first < 2 * BM) first = 2 * BM;
    int64_t pos = 0;
    if (ev0) TS_CHECK_CUDA(cudaEventRecord(ev0, s));
    if (dense) {
        // Small corpus: one GEMM pass writes every score, a select kernel keeps the kp best per query. The
        // chunked threshold filter below has nothing to filter with on its first chunk (threshold -inf: every
        // score is appended and sorted), which dominates when the whole corpus is a few chunks.
        p.dense = dscores;
        p.dense_stride = (int64_t)dense_stride_of(ix);
        p.row_begin = 0;
        p.row_end = ix->size;
        p.num_m_blocks = (int)((ix->size + BM - 1) / BM);
        const int tiles = p.num_m_blocks * p.num_n_blocks;
        if (f32) batched_gemm_topk_kernel<false, true><<<tiles < sms ? tiles : sms, THREADS, SMEM_BYTES, s>>>(tmap_a, tmap_b, p);
        else batched_gemm_topk_kernel<false, false><<<tiles < sms ? tiles : sms, THREADS, SMEM_BYTES, s>>>(tmap_a, tmap_b, p);
        TS_LAUNCH_CHECK();
        // one WARP per query pays off once there are enough queries to fill the SMs with warps (a lone warp walks its
        // row serially: 73 queries x 100k scores took 0.9 ms that way, 0.45 ms with a CTA per query)
        const bool warp_select = nq >= 8 * sms;
        if (!warp_select)
            dense_select_kernel<<<nq, 256, 0, s>>>(dscores, p.dense_stride, ix->size, k, cand, (size_t)k + cap, overflow);
        else if (k <= 32)
            dense_select_warp_kernel<1><<<(nq + 7) / 8, 256, 0, s>>>(dscores, p.dense_stride, ix->size, nq, k, cand, (size_t)k + cap, overflow);
        else if (k <= 128)
            dense_select_warp_kernel<4><<<(nq + 7) / 8, 256, 0, s>>>(dscores, p.dense_stride, ix->size, nq, k, cand, (size_t)k + cap, overflow);
        else if (k <= 256)
            dense_select_warp_kernel<8><<<(nq + 7) / 8, 256, 0, s>>>(dscores, p.dense_stride, ix->size, nq, k, cand, (size_t)k + cap, overflow);
        else
            dense_select_kernel<<<nq, 256, 0, s>>>(dscores, p.dense_stride, ix->size, k, cand, (size_t)k + cap, overflow);
        TS_LAUNCH_CHECK();
        pos = ix->size;
    }
    while (pos < ix->size) {
        int64_t chunk = (pos == 0) ? first : pos * (int64_t)growth;
        chunk = (chunk + 2 * BM - 1) / (2 * BM) * (2 * BM);
        const int64_t end = std::min<int64_t>(ix->size, pos + chunk);
        p.row_begin = pos;
        p.row_end = end;
        const int tile_rows = pair ? 2 * BM : BM;
        p.num_m_blocks = (int)((end - pos + tile_rows - 1) / tile_rows);
        const int tiles = p.num_m_blocks * p.num_n_blocks;
        if (pair) {
            const int pairs = std::min(tiles, sms / 2);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * pairs);
            cfg.blockDim = dim3(THREADS);
            cfg.dynamicSmemBytes = SMEM_BYTES;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, batched_gemm_topk_kernel<true, false>, tmap_a, tmap_b, p));
        } else {
            const int grid = tiles < sms ? tiles : sms;
            if (f32) batched_gemm_topk_kernel<false, true><<<grid, THREADS, SMEM_BYTES, s>>>(tmap_a, tmap_b, p);
            else batched_gemm_topk_kernel<false, false><<<grid, THREADS, SMEM_BYTES, s>>>(tmap_a, tmap_b, p);
        }
        TS_LAUNCH_CHECK();
        compact_candidates_kernel<<<nq, 256, sort_smem, s>>>(cand, count, thr, overflow, k, cap);
        TS_LAUNCH_CHECK();
        pos = end;
    }
    if (ev1) TS_CHECK_CUDA(cudaEventRecord(ev1, s));
    // cand[q][0..kp): best kp by bf16-query score. Re-score with the fp32 query, keep k_out, certify.
    const size_t cstride = (size_t)k + cap;
    int PR = 64;
    while (PR < k) PR <<= 1;
    if (f32)
        rescore_certify_kernel<4><<<nq, 256, (size_t)PR * 8, s>>>(cand, cstride, k, k_out, q32, (const uint8_t*)ix->data,
                                                                  (uint32_t)ix->row_bytes(), ix->dim_pad, qerr,
                                                                  ix->max_norm2, overflow, flags);
    else
        rescore_certify_kernel<2><<<nq, 256, (size_t)PR * 8, s>>>(cand, cstride, k, k_out, q32, (const uint8_t*)ix->data,
                                                                  (uint32_t)ix->row_bytes(), ix->dim_pad, qerr,
                                                                  ix->max_norm2, overflow, flags);
    TS_LAUNCH_CHECK();
    // Fix-up: queries that could not be certified (or overflowed) are re-scanned exactly by K2.
    // The work list lives on the device; with nothing flagged these launches exit immediately.
    build_fix_list_kernel<<<1, 1024, 0, s>>>(flags, nq, fix_list, fix_count);
    TS_LAUNCH_CHECK();
    g_last_fix_count = fix_count;
    g_last_fix_device = ix->device;
    rc = launch_scan_topk(ix, ix->data, ix->dtype, ix->size, q32, nq, k_out, allow_mask, fix_parts, nparts, s, nullptr,
                          nullptr, fix_list, fix_count);
    if (rc) return rc;
    rc = launch_merge_strided(fix_parts, nparts, nq, k_out, /*stride_list=*/k_out, /*stride_query=*/(int64_t)nparts * k_out,
                              nullptr, nullptr, cand, nullptr, nullptr, s, fix_list, fix_count, (int64_t)cstride);
    if (rc) return rc;
    // cand[q][0..k_out) now holds each query's exact sorted top-k keys: emit scores / ids (or keys)
    return launch_merge_strided(cand, 1, nq, k_out, /*stride_list=*/0, /*stride_query=*/(int64_t)cstride, nullptr,
                                ix->has_ids ? ix->ids : nullptr, out_keys, out_scores, out_ids, s);
}

}  // namespace ts
