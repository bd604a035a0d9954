// K2 — single-query exact scan with the top-k fused into the scan (HBM-bandwidth bound).
//
// Replaces, per query, the reference's
//     cosine_scores = util.cos_sim(query_embedding, embeddings_db)[0]      (test_app.py:76)
//     top = np.argsort(-cosine_scores.cpu())[:5]                           (test_app.py:77)
//     torch.topk(cosine_scores, k=min(200, N), sorted=True)                (app_showcase_model.py:96)
//     ORDER BY e.embedding <#> q ASC LIMIT k                               (streamlit_app.py:281-282)
// which re-normalise the corpus, materialise all N scores and fully sort them.
//
// Design (B200):
//   * persistent grid, ctas_per_sm CTAs per SM, W warps each; every warp runs its OWN
//     TMA pipeline: `stages` shared-memory slots, one mbarrier per slot; lane 0 issues
//     cp.async.bulk (1-D TMA, whole rows are contiguous so no tensor map is needed) for the
//     warp's next tile as soon as the warp has finished reading a slot. No block-wide
//     barrier inside the scan loop.
//   * a tile is R whole rows (~8 KB). Lanes read the slot with conflict-free 128-bit LDS
//     (lane l takes bytes [16l, 16l+16) of each 512-byte chunk of a row), convert bf16->fp32
//     with one shift/mask per element and FMA against the query held in registers as fp32.
//   * the R per-row partial sums are reduced with a transposing butterfly (R-1 + log2(32/R)
//     shuffles per tile instead of 5R).
//   * top-k: each warp keeps a sorted list of 64-bit keys (score,row) in registers
//     (WarpTopK); a candidate is inserted only if it beats the warp's current k-th key, which
//     after warm-up is rare, so the steady state costs one compare + one ballot per tile.
//     Scores never go to HBM: each CTA writes only its k best keys; K5 merges the CTA lists.
//
// Algorithmic bytes: N * row_bytes per query (+ D*4 query + nparts*k*8 candidates).
#pragma once

#include <algorithm>

#include "merge_device.cuh"
#include "scan_device.cuh"

namespace ts {

struct ScanParams {
    const uint8_t* data;      // [n_rows, row_bytes]
    int64_t n_rows;
    uint32_t row_bytes;       // dim_pad * elem size, multiple of 16
    int dim_pad;
    const float* queries;     // [nq, dim_pad] fp32, already normalised / zero padded
    int k;
    const uint32_t* mask;     // allow bitmask or nullptr
    uint64_t* part_keys;      // [work item, nparts, k]
    int stages;
    int nq;                   // work items when qcount == nullptr (work item w scans query w)
    const int* qlist;         // fix-up mode: work item w scans query qlist[w] ...
    const int* qcount;        // ... for w < *qcount (device-side count; 0 = every CTA exits at once)
    // fused query preparation: when q_raw != nullptr the kernel normalises the caller's query itself
    // (bit-identical to K1's arithmetic) instead of reading a prepared fp32 copy from `queries`
    const void* q_raw;        // [nq, dim] of q_dtype
    int q_dtype, q_normalize, dim;
    // fused final merge: when tickets != nullptr the last CTA of a work item to finish merges the
    // grid's per-CTA lists and writes the final result (fin.keys / nlists / strides are filled in-kernel)
    uint32_t* tickets;        // [work items], zero on entry, left zero on exit
    MergeParams fin;
    // fused cross-GPU exchange (sharded search): the last CTA stores this shard's k keys into every peer's
    // receive area over NVLink, waits for the peers' keys and merges the world lists — no NCCL call, no
    // second kernel
    XchgDev xchg;
    // programmatic dependent launch (the sharded stream, ts_search_sharded): bit 0 = this grid was launched
    // with programmatic stream serialisation and lets its successor launch at once; bit 1 = queries / mask /
    // corpus may have been written by the kernel that precedes this one on the stream: wait for it before
    // touching anything (the safe default; without it the scan of query n+1 streams the corpus while query
    // n's exchange kernel is still merging). part_keys then points into a ring of 4 per-search buffers; before
    // a CTA writes its list it checks that the exchange kernel which last read this ring slot (search
    // ring_need) has finished: *ring_gate is the sequence number of the last finished exchange.
    int pdl;
    const uint32_t* ring_gate;
    uint32_t ring_need;
    // diagnostics ("scan.timeline"): [gridDim.x][8] %globaltimer stamps of this launch, or nullptr
    unsigned long long* timeline;
    // optional: the prepared (normalised, zero padded) fp32 queries are also written here [nq, dim_pad] — the IVF
    // coarse scan hands them to the list scan and the re-score, saving a separate preparation launch
    float* q_out;
};

__device__ __forceinline__ void scan_stamp(const ScanParams& p, int slot) {
    if (p.timeline != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ float load_query_elem(const void* base, int dtype, size_t idx) {
    if (dtype == TS_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
    if (dtype == TS_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
    return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}

template <int ELEM, int NCHUNK, int KPL, int R>
__global__ void __launch_bounds__(512, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int CN = Chunk<ELEM>::N;       // elements per 16-byte chunk
    constexpr int GROUP = 32 / R;            // lanes that end up holding the same row's score
    extern __shared__ __align__(128) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int stages = p.stages;
    const uint32_t tile_bytes = R * p.row_bytes;

    uint8_t* my_slots = smem + (size_t)warp * stages * tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)W * stages * tile_bytes);
    uint64_t* my_bars = bars + warp * stages;

    scan_stamp(p, 0);                                  // 0: CTA entry
    if (p.timeline != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.timeline[(size_t)blockIdx.x * 8 + 7] = smid;
    }
    if (p.pdl & 1) griddep_launch_dependents();
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&my_bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (p.pdl & 2) griddep_wait();

    const int64_t num_tiles = (p.n_rows + R - 1) / R;
    const int64_t gw = (int64_t)blockIdx.x * W + warp;
    const int64_t tw = (int64_t)gridDim.x * W;
    const uint64_t policy = l2_policy_evict_first();
    const int k = p.k;
    const int my_row = row_of_lane<R>(lane);
    const bool leader = (lane & (GROUP - 1)) == 0;

    auto issue = [&](int64_t tile, int s) {
        const int64_t row0 = tile * R;
        const int64_t rows = (p.n_rows - row0 < R) ? (p.n_rows - row0) : R;
        const uint32_t bytes = (uint32_t)rows * p.row_bytes;
        mbar_expect_tx(&my_bars[s], bytes);
        tma_load_1d_hint(my_slots + (size_t)s * tile_bytes, p.data + (size_t)row0 * p.row_bytes, bytes,
                         &my_bars[s], policy);
    };

    // ring position persists across work items (mbarrier phases cannot be rewound)
    int s = 0;
    uint32_t parity = 0;
    const int nwork = p.qcount ? *p.qcount : p.nq;
    for (int wi = blockIdx.y; wi < nwork; wi += gridDim.y) {
    const int qi = p.qlist ? p.qlist[wi] : wi;

    if (lane == 0) {   // get the corpus stream going before anything else
        int ss = s;
        for (int i = 0; i < stages; ++i) {
            const int64_t t = gw + (int64_t)i * tw;
            if (t < num_tiles) issue(t, ss);
            if (++ss == stages) ss = 0;
        }
    }

    // query slice of this lane, fp32 in registers
    float q[NCHUNK * CN];
    if (p.q_raw == nullptr) {
        const float* qv = p.queries + (size_t)qi * p.dim_pad;
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const int e0 = (j * 32 + lane) * CN;
#pragma unroll
            for (int i = 0; i < CN; ++i) q[j * CN + i] = (e0 + i < p.dim_pad) ? __ldg(qv + e0 + i) : 0.f;
        }
    } else {
        // K1's arithmetic, replicated per warp (4 KB from L2): ||x|| accumulated in fp64 with lane l
        // taking the pairs (2l, 2l+1) + 64j in ascending order, then the butterfly; fp32 division by
        // max((float)sqrt, 1e-12).
        const size_t qoff = (size_t)qi * p.dim;
        float den = 1.0f;
        if (p.q_dtype == TS_F32) {
            // fp32 queries (the common case): every load of a phase is independent and issued back to back —
            // one L2 round trip per phase. (A rolled loop of dependent-latency loads cost 10-13 us here, during
            // which the warp's two TMA slots sat full: 3 % of a 1.25M-row shard's scan.)
            const float* qv = reinterpret_cast<const float*>(p.q_raw) + qoff;
            // both views of the query are requested at once (ONE L2 round trip): the pair view K1's norm is summed
            // in, and the chunk view the scan multiplies with
            constexpr int NP = NCHUNK * CN / 2;            // pairs per lane: (2l, 2l+1) + 64j
            const bool vec2 = ((p.dim & 1) == 0) && ((reinterpret_cast<uintptr_t>(qv) & 7u) == 0);
            const bool vec4 = ((p.dim & 3) == 0) && ((reinterpret_cast<uintptr_t>(qv) & 15u) == 0);
            float2 v[NP];
            if (p.q_normalize) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = 2 * lane + 64 * j;
                    v[j] = make_float2(0.f, 0.f);
                    if (vec2) {
                        if (i < p.dim) v[j] = __ldg(reinterpret_cast<const float2*>(qv + i));
                    } else {
                        if (i < p.dim) v[j].x = __ldg(qv + i);
                        if (i + 1 < p.dim) v[j].y = __ldg(qv + i + 1);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) {
                const int e0 = (j * 32 + lane) * CN;
#pragma unroll
                for (int i = 0; i < CN; i += 4) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (vec4) {
                        if (e0 + i < p.dim) f = __ldg(reinterpret_cast<const float4*>(qv + e0 + i));
                    } else {
                        if (e0 + i < p.dim) f.x = __ldg(qv + e0 + i);
                        if (e0 + i + 1 < p.dim) f.y = __ldg(qv + e0 + i + 1);
                        if (e0 + i + 2 < p.dim) f.z = __ldg(qv + e0 + i + 2);
                        if (e0 + i + 3 < p.dim) f.w = __ldg(qv + e0 + i + 3);
                    }
                    q[j * CN + i] = f.x;
                    q[j * CN + i + 1] = f.y;
                    q[j * CN + i + 2] = f.z;
                    q[j * CN + i + 3] = f.w;
                }
            }
            if (p.q_normalize) {
                double ss = 0.0;
#pragma unroll
                for (int j = 0; j < NP; ++j) {                 // zeros past dim leave ss unchanged (K1 does the same)
                    const double a = (double)v[j].x, b = (double)v[j].y;
                    ss = fma(a, a, ss);
                    ss = fma(b, b, ss);
                }
                ss = warp_sum(ss);
                den = fmaxf((float)sqrt(ss), 1e-12f);
#pragma unroll
                for (int i = 0; i < NCHUNK * CN; ++i) q[i] = __fdiv_rn(q[i], den);
            }
        } else {
        if (p.q_normalize) {
            double ss = 0.0;
            for (int i = 2 * lane; i < p.dim; i += 64) {   // K1's order: lane l owns pairs (2l, 2l+1) + 64j
                const double a = (double)load_query_elem(p.q_raw, p.q_dtype, qoff + i);
                ss = fma(a, a, ss);
                if (i + 1 < p.dim) {
                    const double b = (double)load_query_elem(p.q_raw, p.q_dtype, qoff + i + 1);
                    ss = fma(b, b, ss);
                }
            }
            ss = warp_sum(ss);
            den = fmaxf((float)sqrt(ss), 1e-12f);
        }
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const int e0 = (j * 32 + lane) * CN;
#pragma unroll
            for (int i = 0; i < CN; ++i) {
                float v = (e0 + i < p.dim) ? load_query_elem(p.q_raw, p.q_dtype, qoff + e0 + i) : 0.f;
                if (p.q_normalize) v = __fdiv_rn(v, den);
                q[j * CN + i] = v;
            }
        }
        }
    }

    if (p.q_out != nullptr && blockIdx.x == 0 && warp == 0) {
        float* qo = p.q_out + (size_t)qi * p.dim_pad;
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const int e0 = (j * 32 + lane) * CN;
#pragma unroll
            for (int i = 0; i < CN; ++i)
                if (e0 + i < p.dim_pad) qo[e0 + i] = q[j * CN + i];
        }
    }
    WarpTopK<KPL> list;
    list.clear();
    uint64_t thr = 0ull;  // current k-th key of this warp's list (0 while it has < k entries)
    scan_stamp(p, 1);                                  // 1: query in registers, scan loop starts
    bool first_tile = true;

    for (int64_t tile = gw; tile < num_tiles; tile += tw) {
        const int64_t row0 = tile * R;
        const int64_t row = row0 + my_row;
        const bool in_range = row < p.n_rows;

        // fetch the allow bit early so its latency hides behind the barrier wait + FMAs
        bool allowed = in_range && leader;
        if (p.mask != nullptr && allowed) allowed = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;

        mbar_wait(&my_bars[s], parity);
        if (first_tile) {
            scan_stamp(p, 2);                          // 2: warp 0's first tile has landed
            first_tile = false;
        }

        const uint8_t* slot = my_slots + (size_t)s * tile_bytes;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            const uint32_t off = (uint32_t)(j * 32 + lane) * 16u;
            if (off < p.row_bytes) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    // rows past the end of a short last tile hold stale smem; masked below
                    const uint4 v = *reinterpret_cast<const uint4*>(slot + (size_t)r * p.row_bytes + off);
                    acc[r] = Chunk<ELEM>::dot(v, &q[j * CN], acc[r]);
                }
            }
        }
        __syncwarp();  // every lane is done reading the slot
        if (lane == 0) {
            const int64_t nt = tile + (int64_t)stages * tw;
            if (nt < num_tiles) issue(nt, s);
        }

        transpose_reduce<R>(acc, lane);
        const uint64_t key = allowed ? pack_key(acc[0], (uint32_t)row) : 0ull;
        unsigned m = __ballot_sync(0xFFFFFFFFu, key > thr);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t x = __shfl_sync(0xFFFFFFFFu, key, src);
            if (x > thr) {
                list.insert(x, lane);
                thr = list.kth(k);
            }
        }
        if (++s == stages) {
            s = 0;
            parity ^= 1u;
        }
    }

    // ---- CTA merge: warps park their lists in smem (the pipeline has drained: every issued
    // copy was waited on), warp 0 folds them into its own and writes the CTA's k best.
    scan_stamp(p, 3);                                  // 3: warp 0 finished its tiles
    __syncthreads();
    scan_stamp(p, 4);                                  // 4: every warp of the CTA finished
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem);  // [W][KPL*32]
#pragma unroll
    for (int j = 0; j < KPL; ++j) lists[(size_t)warp * (KPL * 32) + j * 32 + lane] = list.key[j];
    __syncthreads();
    if constexpr (KPL == 1) {
        // k <= 32: rank by counting. Every thread holds one key of its warp's sorted list; its position in the CTA's
        // merged list is the number of larger keys among the W * k parked ones (smem broadcast reads) — 80 compares
        // for 8 warps x top-10 instead of seven dependent register merges by one warp (4 us -> under 1 us).
        if (p.ring_gate != nullptr && threadIdx.x == 0) {   // (never waits in practice: the slot's reader is four searches back)
            while ((int32_t)(ld_acquire_gpu_u32(p.ring_gate) - p.ring_need) < 0) __nanosleep(200);
        }
        const uint64_t mine = (lane < k) ? list.key[0] : 0ull;
        const int live = __syncthreads_count(mine != 0ull);       // also orders thread 0's gate wait before the stores
        uint64_t* out = p.part_keys + ((size_t)wi * gridDim.x + blockIdx.x) * k;
        if (mine != 0ull) {
            int rank = 0;
            for (int w = 0; w < W; ++w) {
                const uint64_t* lw = lists + (size_t)w * 32;
#pragma unroll 4
                for (int j = 0; j < k; ++j) rank += (lw[j] > mine) ? 1 : 0;
            }
            if (rank < k) out[rank] = mine;
        }
        for (int pos = live + (int)threadIdx.x; pos < k; pos += blockDim.x) out[pos] = 0ull;   // fewer than k rows seen
    } else if (warp == 0) {
        for (int w = 1; w < W; ++w) merge_sorted_into<KPL>(list, lists + (size_t)w * (KPL * 32), k, k, lane);
        if (p.ring_gate != nullptr) {    // (never waits in practice: the slot's reader is four searches back)
            while ((int32_t)(ld_acquire_gpu_u32(p.ring_gate) - p.ring_need) < 0) __nanosleep(200);
        }
        uint64_t* out = p.part_keys + ((size_t)wi * gridDim.x + blockIdx.x) * k;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            const int pos = j * 32 + lane;
            if (pos < k) out[pos] = list.key[j];
        }
    }
    __syncthreads();  // the list area aliases the TMA slots of the next work item
    scan_stamp(p, 5);                                  // 5: CTA list written
    if (p.tickets != nullptr) {
        // ---- fused final merge: the last CTA of this work item to arrive folds all gridDim.x lists.
        // One gpu-scope fence on each side of the ticket, by thread 0 only: the barriers order the other
        // threads' stores before it / loads after it (fences are cumulative).
        __shared__ int s_is_last;
        if (threadIdx.x == 0) {
            __threadfence();
            const uint32_t t = atomicAdd(p.tickets + wi, 1u);
            s_is_last = (t == gridDim.x - 1);
            if (s_is_last) {
                p.tickets[wi] = 0u;   // self-cleaning: the next launch finds zeros
                __threadfence();
            }
        }
        __syncthreads();
        if (s_is_last) {
            MergeParams mp = p.fin;
            mp.keys = p.part_keys;
            mp.nlists = gridDim.x;
            mp.k = k;
            mp.stride_list = k;
            mp.stride_query = (int64_t)gridDim.x * k;
            if (p.xchg.world == 0) {
                merge_lists<KPL>(mp, wi, qi, lists, W);
            } else {
                // ---- fused exchange (one-kernel form): this CTA also pushes the shard's keys to the peers,
                // waits for theirs and merges the world lists
                mp.done_flag = nullptr;   // the shard-local merge is an intermediate step: only the world merge signals
                exchange_and_merge<KPL>(p.xchg, p.fin, mp, wi, qi, k, lists, W);
            }
            scan_stamp(p, 6);                          // 6: (last CTA only) final merge / exchange done
        }
    }
    }  // work items
}

// ------------------------------------------------------------------------------------ host side
struct ScanConfig {
    int grid, warps, stages;
    size_t smem;
};

template <int ELEM, int NCHUNK, int R>
static ScanConfig scan_config(const ts_index* ix, int kpl) {
    const Tunables& t = tunables();
    const size_t tile_bytes = (size_t)R * ix->dim_pad * ELEM;
    int stages = t.scan_stages < 2 ? 2 : t.scan_stages;
    int ctas = t.scan_ctas_per_sm < 1 ? 1 : t.scan_ctas_per_sm;
    int warps = t.scan_warps < 1 ? 1 : (t.scan_warps > 16 ? 16 : t.scan_warps);
    const size_t budget = (size_t)(220 * 1024) / ctas - 1024;
    while (warps > 1 && (size_t)warps * stages * tile_bytes + 8 * warps * stages > budget) --warps;
    while (stages > 2 && (size_t)warps * stages * tile_bytes + 8 * warps * stages > budget) --stages;
    size_t smem = (size_t)warps * stages * tile_bytes + 8 * (size_t)warps * stages;
    const size_t list_bytes = (size_t)warps * kpl * 32 * 8;
    if (smem < list_bytes) smem = list_bytes;
    ScanConfig c;
    c.grid = sm_count(ix->device) * ctas;
    c.warps = warps;
    c.stages = stages;
    c.smem = smem;
    return c;
}

template <int ELEM, int NCHUNK, int KPL, int R>
static int launch_r(const ts_index* ix, const ScanParams& p0, int nq, int nparts, cudaStream_t s,
                    cudaEvent_t ev0, cudaEvent_t ev1) {
    ScanConfig c = scan_config<ELEM, NCHUNK, R>(ix, KPL);
    ScanParams p = p0;
    p.stages = c.stages;
    auto kern = scan_topk_kernel<ELEM, NCHUNK, KPL, R>;
    TS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    if (ev0) TS_CHECK_CUDA(cudaEventRecord(ev0, s));
    // Small tables (e.g. the IVF centroid table): fewer CTAs, so that every warp still gets a few tiles and
    // the fused final merge folds fewer lists. Only when the merge is fused (it reads gridDim.x lists);
    // the unfused callers size their merge for nparts lists.
    int grid = nparts;
    if (p.tickets != nullptr) {
        const int64_t tiles = (p.n_rows + R - 1) / R;
        grid = (int)std::min<int64_t>(nparts, std::max<int64_t>(1, (tiles + c.warps * 4 - 1) / (c.warps * 4)));
    }
    if (p.pdl & 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid, p.qcount ? std::min(nq, 8) : nq);
        cfg.blockDim = dim3(c.warps * 32);
        cfg.dynamicSmemBytes = c.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    } else {
        kern<<<dim3(grid, p.qcount ? std::min(nq, 8) : nq), c.warps * 32, c.smem, s>>>(p);
    }
    TS_LAUNCH_CHECK();
    if (ev1) TS_CHECK_CUDA(cudaEventRecord(ev1, s));
    return TS_OK;
}

// rows per TMA tile: the default aims at ~8 KB; "scan.tile_rows" (2/4/8/16) overrides it for the
// common k <= 32 kernels (the large-k instances keep the default to bound compile time).
template <int ELEM, int NCHUNK, int KPL>
static int launch_one(const ts_index* ix, const ScanParams& p, int nq, int nparts, cudaStream_t s,
                      cudaEvent_t ev0, cudaEvent_t ev1) {
    constexpr int RD = RowsPerTile<NCHUNK>::value;
    if constexpr (KPL == 1) {
        switch (tunables().scan_tile_rows) {
            case 2: return launch_r<ELEM, NCHUNK, KPL, 2>(ix, p, nq, nparts, s, ev0, ev1);
            case 4: return launch_r<ELEM, NCHUNK, KPL, 4>(ix, p, nq, nparts, s, ev0, ev1);
            case 8: return launch_r<ELEM, NCHUNK, KPL, 8>(ix, p, nq, nparts, s, ev0, ev1);
            case 16: return launch_r<ELEM, NCHUNK, KPL, 16>(ix, p, nq, nparts, s, ev0, ev1);
            default: break;
        }
    }
    return launch_r<ELEM, NCHUNK, KPL, RD>(ix, p, nq, nparts, s, ev0, ev1);
}

template <int ELEM, int NCHUNK>
static int launch_k(const ts_index* ix, const ScanParams& p, int nq, int nparts, cudaStream_t s,
                    cudaEvent_t ev0, cudaEvent_t ev1) {
    if (p.k <= 32) return launch_one<ELEM, NCHUNK, 1>(ix, p, nq, nparts, s, ev0, ev1);
    if (p.k <= 128) return launch_one<ELEM, NCHUNK, 4>(ix, p, nq, nparts, s, ev0, ev1);
    if (p.k <= 256) return launch_one<ELEM, NCHUNK, 8>(ix, p, nq, nparts, s, ev0, ev1);
    return launch_one<ELEM, NCHUNK, 32>(ix, p, nq, nparts, s, ev0, ev1);
}

}  // namespace ts
