// K4d — list-major batched IVF scan: every probed list is read ONCE per group of QB queries.
//
// The per-query list scan (K4b) reads a list once for every query that probes it; with a batch of 4096
// queries x 32 probes over 16384 lists every list is wanted by ~8 queries, so the batch reads the probed part
// of the corpus 8 times over and is HBM-bound at that inflated volume. Here the (query, list) pairs are
// inverted into a per-list query table and a CTA takes one (list, group of <= QB queries) work item and streams
// the list's e4m3 rows once. Two scoring variants:
//   * tensor cores (default): legacy mma.sync m16n8k16, fp16 queries in smem, e4m3 rows converted in registers,
//     fp32 accumulation, QB = 8 (or 16) — back at the HBM bound of the reduced volume (see the comment above
//     ivf_grouped_mma_kernel);
//   * CUDA cores: rows converted to half2 once, QB = 4 queries held in registers, packed HFMA2 — bound by the
//     FP16 FMA pipe (kept as `ivf.group_mma = 0`).
// Scores go to a dense fp32 buffer laid out per (query, probe) pair — 4 bytes written per row-query against
// 1 KB of row read per QB queries — and a selection kernel picks each query's best k' (threshold filter into a
// shared buffer + cooperative CTA sorts), which the exact re-score (K4c) consumes.
//
//   G1 ivf_invert_count_kernel   histogram of probes per list
//   G2 ivf_invert_scan_kernel    exclusive scans: table slots, work items (ceil(cnt/QB)), score-buffer bases
//   G3 ivf_invert_fill_kernel    per-list query table + per-pair score offsets; ivf_item_table_kernel: item -> list
//   G4 ivf_grouped_mma_kernel / ivf_grouped_scan_kernel   the scan
//   G5 ivf_select_kernel         per-query top-k' over its pairs' score runs
//
// If the score buffer the caller's workspace provides is too small for this batch (heavily skewed lists) G2
// raises a flag, G4/G5 exit at once and the per-query kernel (K4b) runs instead — decided on the device.
#include <algorithm>

#include "merge_device.cuh"
#include "scan_device.cuh"
#include "umma_device.cuh"

namespace ts {

namespace g4 {
constexpr int QB = 4;   // queries per work item (their slices live in registers: QB * 16 half2 per lane at D = 1024)
constexpr int R = 8;    // rows per TMA tile
}  // namespace g4

__global__ void ivf_invert_count_kernel(const uint64_t* __restrict__ probes, int n_pairs, uint32_t* __restrict__ cnt) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pairs; e += gridDim.x * blockDim.x) {
        const uint64_t key = probes[e];
        if (key != 0ull) atomicAdd(cnt + key_row(key), 1u);
    }
}

// One CTA. slot_start[l] = sum cnt[<l], item_start[l] = sum ceil(cnt/QB)[<l], base[l] = sum (cnt*len)[<l].
// totals[0] = number of work items, totals[1] = 1 if the score buffer is too small (fall back to K4b).
__global__ void __launch_bounds__(1024) ivf_invert_scan_kernel(const uint32_t* __restrict__ cnt,
                                                               const int64_t* __restrict__ list_offsets, int nlist,
                                                               uint32_t* __restrict__ slot_start,
                                                               uint32_t* __restrict__ item_start,
                                                               unsigned long long* __restrict__ base,
                                                               uint32_t* __restrict__ totals,
                                                               unsigned long long score_cap, int qb) {
    __shared__ unsigned long long w_a[32], w_b[32], w_c[32];
    __shared__ unsigned long long carry[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) carry[threadIdx.x] = 0ull;
    __syncthreads();
    for (int l0 = 0; l0 < nlist; l0 += 1024) {
        const int l = l0 + threadIdx.x;
        unsigned long long a = 0, b = 0, c = 0;
        if (l < nlist) {
            const unsigned long long n = cnt[l];
            a = n;
            b = (n + qb - 1) / qb;
            c = n * (unsigned long long)(((list_offsets[l + 1] - list_offsets[l]) + 3) & ~3ll);   // runs padded to 16 B
        }
        unsigned long long ia = a, ib = b, ic = c;   // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long ua = __shfl_up_sync(0xFFFFFFFFu, ia, o);
            const unsigned long long ub = __shfl_up_sync(0xFFFFFFFFu, ib, o);
            const unsigned long long uc = __shfl_up_sync(0xFFFFFFFFu, ic, o);
            if (lane >= o) {
                ia += ua;
                ib += ub;
                ic += uc;
            }
        }
        if (lane == 31) {
            w_a[warp] = ia;
            w_b[warp] = ib;
            w_c[warp] = ic;
        }
        __syncthreads();
        if (warp == 0) {   // exclusive scan of the 32 warp totals
            unsigned long long ta = w_a[lane], tb = w_b[lane], tc = w_c[lane];
            const unsigned long long oa = ta, ob = tb, oc = tc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ua = __shfl_up_sync(0xFFFFFFFFu, ta, o);
                const unsigned long long ub = __shfl_up_sync(0xFFFFFFFFu, tb, o);
                const unsigned long long uc = __shfl_up_sync(0xFFFFFFFFu, tc, o);
                if (lane >= o) {
                    ta += ua;
                    tb += ub;
                    tc += uc;
                }
            }
            w_a[lane] = ta - oa;
            w_b[lane] = tb - ob;
            w_c[lane] = tc - oc;
        }
        __syncthreads();
        const unsigned long long pa = carry[0] + w_a[warp] + ia - a;
        const unsigned long long pb = carry[1] + w_b[warp] + ib - b;
        const unsigned long long pc = carry[2] + w_c[warp] + ic - c;
        if (l < nlist) {
            slot_start[l] = (uint32_t)pa;
            item_start[l] = (uint32_t)pb;
            base[l] = pc;
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry[0] = pa + a;
            carry[1] = pb + b;
            carry[2] = pc + c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        slot_start[nlist] = (uint32_t)carry[0];
        item_start[nlist] = (uint32_t)carry[1];
        base[nlist] = carry[2];
        totals[0] = (uint32_t)carry[1];
        totals[1] = carry[2] > score_cap ? 1u : 0u;
        totals[2] = 0u;   // work-item counter of the persistent scan kernel (dynamic scheduling)
    }
}

__global__ void ivf_invert_fill_kernel(const uint64_t* __restrict__ probes, int n_pairs,
                                       const uint32_t* __restrict__ slot_start,
                                       const unsigned long long* __restrict__ base,
                                       const int64_t* __restrict__ list_offsets, uint32_t* __restrict__ cursor,
                                       uint32_t* __restrict__ inv, unsigned long long* __restrict__ pair_off,
                                       uint32_t* __restrict__ pair_len, uint32_t* __restrict__ pair_pos0) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pairs; e += gridDim.x * blockDim.x) {
        const uint64_t key = probes[e];
        if (key == 0ull) {
            pair_len[e] = 0u;
            continue;
        }
        const uint32_t l = key_row(key);
        const int64_t start = list_offsets[l];
        const uint32_t len = (uint32_t)(list_offsets[l + 1] - start);
        const uint32_t i = atomicAdd(cursor + l, 1u);
        inv[slot_start[l] + i] = (uint32_t)e;
        pair_off[e] = base[l] + (unsigned long long)i * ((len + 3u) & ~3u);
        pair_len[e] = len;
        pair_pos0[e] = (uint32_t)start;
    }
}

// item_list[item] = the list a work item belongs to (so a CTA needs one load, not a binary search)
__global__ void ivf_item_table_kernel(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ item_start, int nlist,
                                      uint32_t* __restrict__ item_list, int qb) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const uint32_t n = (cnt[l] + qb - 1) / qb;
    for (uint32_t g = 0; g < n; ++g) item_list[item_start[l] + g] = (uint32_t)l;
}

struct GroupedParams {
    const uint8_t* list_data;
    uint32_t row_bytes;
    int dim_pad;
    const float* scales;
    const int64_t* list_offsets;
    const float* queries;          // [nq, dim_pad] fp32 normalised
    int nprobe, nlist;
    const uint32_t* mask;          // allow bitmask over corpus rows or nullptr
    const uint32_t* list_rows;
    const uint32_t* cnt;
    const uint32_t* slot_start;
    const uint32_t* item_start;
    const uint32_t* item_list;     // work item -> list
    const uint32_t* inv;           // table slot -> pair index (query * nprobe + probe rank)
    const unsigned long long* pair_off;
    const uint32_t* totals;
    float* scores;
    int stages;
};

template <int NCHUNK>
__global__ void __launch_bounds__(384, 1) ivf_grouped_scan_kernel(const GroupedParams p) {
    using namespace g4;
    constexpr int NH = NCHUNK * 8;   // half2 per lane per row
    constexpr int GROUP = 32 / R;
    extern __shared__ __align__(128) uint8_t smem[];
    if (p.totals[1] != 0u) return;
    const uint32_t item = blockIdx.x;
    if (item >= p.totals[0]) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int W = blockDim.x >> 5;
    const int stages = p.stages;
    const uint32_t tile_bytes = R * p.row_bytes;
    uint8_t* my_slots = smem + (size_t)warp * stages * tile_bytes;
    uint64_t* my_bars = reinterpret_cast<uint64_t*>(smem + (size_t)W * stages * tile_bytes) + warp * stages;
    if (lane == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&my_bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int l = (int)p.item_list[item];
    const int g = (int)(item - p.item_start[l]);
    const int c = (int)p.cnt[l];
    const int nqg = (c - g * QB < QB) ? (c - g * QB) : QB;
    const uint32_t s0 = p.slot_start[l] + (uint32_t)g * QB;
    const int64_t start = p.list_offsets[l];
    const int len = (int)(p.list_offsets[l + 1] - start);

    const int T = (len + R - 1) / R;
    const int t0 = (int)((int64_t)T * warp / W), t1 = (int)((int64_t)T * (warp + 1) / W);
    const uint64_t policy = l2_policy_evict_first();
    const int my_row = row_of_lane<R>(lane);
    const bool leader = (lane & (GROUP - 1)) == 0;
    auto issue = [&](int t, int s) {
        const int left = len - t * R;
        const uint32_t bytes = (uint32_t)(left < R ? left : R) * p.row_bytes;
        mbar_expect_tx(&my_bars[s], bytes);
        tma_load_1d_hint(my_slots + (size_t)s * tile_bytes, p.list_data + (size_t)(start + (int64_t)t * R) * p.row_bytes,
                         bytes, &my_bars[s], policy);
    };
    if (lane == 0)   // get the list streaming before anything else
        for (int st = 0; st < stages && t0 + st < t1; ++st) issue(t0 + st, st);

    // query slices (half2) and output bases of the group
    __half2 qh[QB][NH];
    unsigned long long obase[QB];
#pragma unroll
    for (int qq = 0; qq < QB; ++qq) {
        obase[qq] = 0ull;
        if (qq < nqg) {
            const uint32_t e = p.inv[s0 + qq];
            obase[qq] = p.pair_off[e];
            const float* qv = p.queries + (size_t)(e / (uint32_t)p.nprobe) * p.dim_pad;
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) {
                const int e0 = (j * 32 + lane) * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e0 + i < p.dim_pad) f = __ldg(reinterpret_cast<const float4*>(qv + e0 + i));
                    qh[qq][j * 8 + i / 2] = __floats2half2_rn(f.x, f.y);
                    qh[qq][j * 8 + i / 2 + 1] = __floats2half2_rn(f.z, f.w);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NH; ++i) qh[qq][i] = __float2half2_rn(0.f);
        }
    }

    int s = 0;
    uint32_t parity = 0;
    for (int t = t0; t < t1; ++t) {
        const int left = len - t * R;
        const int row_in_list = t * R + my_row;
        const int64_t pos = start + row_in_list;
        bool mine = leader && my_row < left;
        bool allowed = true;
        if (p.mask != nullptr && mine) {
            const uint32_t row = __ldg(p.list_rows + pos);
            allowed = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
        }
        float scale = 1.0f;
        if (mine) scale = __ldg(p.scales + pos);

        mbar_wait(&my_bars[s], parity);
        const uint8_t* slot = my_slots + (size_t)s * tile_bytes;
        float acc[QB][R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // the row's e4m3 values as half2, converted once for all QB queries
            __half2 e[NH];
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) {
                const uint32_t off = (uint32_t)(j * 32 + lane) * 16u;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (off < p.row_bytes) v = *reinterpret_cast<const uint4*>(slot + (size_t)r * p.row_bytes + off);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2_raw a = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] & 0xFFFFu), __NV_E4M3);
                    const __half2_raw b = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] >> 16), __NV_E4M3);
                    e[j * 8 + 2 * i] = *reinterpret_cast<const __half2*>(&a);
                    e[j * 8 + 2 * i + 1] = *reinterpret_cast<const __half2*>(&b);
                }
            }
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
                // 16 products per fp16 accumulator at D = 1024 (|e4m3| <= 448, |q| <= 1: far from overflow);
                // the rounding this adds is well below the e4m3 noise, and candidates are re-scored exactly
                __half2 a0 = __float2half2_rn(0.f), a1 = __float2half2_rn(0.f);
#pragma unroll
                for (int i = 0; i < NH; i += 2) {
                    a0 = __hfma2(e[i], qh[qq][i], a0);
                    a1 = __hfma2(e[i + 1], qh[qq][i + 1], a1);
                }
                const float2 f = __half22float2(__hadd2(a0, a1));
                acc[qq][r] = f.x + f.y;
            }
        }
        __syncwarp();
        if (lane == 0 && t + stages < t1) issue(t + stages, s);

#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
            transpose_reduce<R>(acc[qq], lane);
            if (mine && qq < nqg) p.scores[obase[qq] + (unsigned long long)row_in_list] = allowed ? acc[qq][0] * scale : -INFINITY;
        }
        if (++s == stages) {
            s = 0;
            parity ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------------- G4 on the tensor cores
// Same work item, 8 queries per group, scored with legacy mma.sync m16n8k16 (f16 x f16 -> f32): the CUDA-core
// variant above is bound by the FP16 FMA pipe (16 HFMA2 per row-query per lane); here 16 rows x 8 queries x 16
// dims cost one HMMA + 4 F2FP + 3 LDS, i.e. ~4 instructions per row-query per lane, accumulation is fp32 and the
// queries stay fp16 (no query quantisation) — the scan goes back to being HBM-bound, at 1/8 of K4b's volume.
//   * A (rows): a 16-row tile is copied row by row (one 1-D TMA copy each, one mbarrier per tile) into smem with
//     a row stride of row_bytes + 16, so the fragment loads — lane (g, t) reads the 4 e4m3 bytes at k0 + 4t of
//     rows g and g + 8 — hit 32 distinct banks. The 4 bytes become a0/a2 (row g) and a1/a3 (row g + 8): the
//     fragment's logical k pairs (2t, 2t+1) and (2t+8, 2t+9) are mapped onto the PHYSICAL bytes 4t..4t+3, a
//     permutation of k inside each 16-block that B uses too, so the dot products are unchanged.
//   * B (queries): fp16 in smem [8][D + 8]; lane (g, t) reads query g's 4 halves at k0 + 4t as one 8-byte load.
//   * C: c0/c1 = (row g, queries 2t, 2t+1), c2/c3 = (row g + 8, same) -> x row scale -> the dense score buffer.
namespace g4m {
constexpr int R = 16;
constexpr int STAGES = 2;
// NT = n-tiles of 8 queries per group. 2 stages x 16 padded rows per warp (33 KB) + the queries (16.5 KB per
// n-tile) must fit 227 KB: 6 warps with 8 queries, 5 with 16.
constexpr int warps_for(int nt) { return nt == 1 ? 6 : 5; }
}  // namespace g4m

__device__ __forceinline__ void mma_m16n8k16_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// four e4m3 bytes of one 32-bit word -> two packed f16x2 registers (bytes 0,1 and bytes 2,3): two F2FP, no repacking
__device__ __forceinline__ void e4m3x4_to_f16x2x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    asm("{\n\t.reg .b16 l, h;\n\t"
        "mov.b32 {l, h}, %2;\n\t"
        "cvt.rn.f16x2.e4m3x2 %0, l;\n\t"
        "cvt.rn.f16x2.e4m3x2 %1, h;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(w));
}

template <int NT>
__global__ void __launch_bounds__(g4m::warps_for(NT) * 32, 1) ivf_grouped_mma_kernel(const GroupedParams p) {
    using namespace g4m;
    constexpr int QB = 8 * NT;
    constexpr int WARPS = warps_for(NT);
    extern __shared__ __align__(128) uint8_t smem[];
    if (p.totals[1] != 0u) return;
    const uint32_t item = blockIdx.x;
    if (item >= p.totals[0]) return;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t row_stride = p.row_bytes + 16u;              // padded: fragment loads are bank-conflict free
    const uint32_t tile_bytes = R * row_stride;
    const uint32_t q_stride = (uint32_t)p.row_bytes + 16u;      // halves per query row in smem: word stride = 8 mod 32,
                                                                // so a half-warp's 8-byte B loads cover all 32 banks once
    uint8_t* my_slots = smem + (size_t)warp * STAGES * tile_bytes;
    uint64_t* my_bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * tile_bytes) + warp * STAGES;
    __half* qs = reinterpret_cast<__half*>(smem + (size_t)WARPS * STAGES * tile_bytes + 8 * WARPS * STAGES);
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&my_bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int l = (int)p.item_list[item];
    const int grp = (int)(item - p.item_start[l]);
    const int c = (int)p.cnt[l];
    const int nqg = (c - grp * QB < QB) ? (c - grp * QB) : QB;
    const uint32_t s0 = p.slot_start[l] + (uint32_t)grp * QB;
    const int64_t start = p.list_offsets[l];
    const int len = (int)(p.list_offsets[l + 1] - start);
    const int T = (len + R - 1) / R;
    const int t0 = (int)((int64_t)T * warp / WARPS), t1 = (int)((int64_t)T * (warp + 1) / WARPS);
    const uint64_t policy = l2_policy_evict_first();
    auto issue = [&](int tile, int s) {   // whole warp: lane 0 arms the barrier, lanes 0..rows-1 copy one row each
        const int left = len - tile * R;
        const int rows = left < R ? left : R;
        if (lane == 0) mbar_expect_tx(&my_bars[s], (uint32_t)rows * p.row_bytes);
        __syncwarp();
        if (lane < rows)
            tma_load_1d_hint(my_slots + (size_t)s * tile_bytes + (size_t)lane * row_stride,
                             p.list_data + (size_t)(start + (int64_t)tile * R + lane) * p.row_bytes, p.row_bytes,
                             &my_bars[s], policy);
    };
    for (int s = 0; s < STAGES && t0 + s < t1; ++s) issue(t0 + s, s);

    // the group's queries as fp16 in smem (zeros for missing queries and past dim_pad)
    unsigned long long obase[NT][2];   // score runs of the queries this lane's accumulators belong to
    for (int i = threadIdx.x; i < QB * (int)p.row_bytes / 4; i += blockDim.x) {   // 4 dims per thread per step
        const int qq = (4 * i) / (int)p.row_bytes, d = (4 * i) % (int)p.row_bytes;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qq < nqg && d < p.dim_pad) {   // dim_pad is a multiple of 8: the four dims are inside or outside together
            const uint32_t e = p.inv[s0 + qq];
            v = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)(e / (uint32_t)p.nprobe) * p.dim_pad + d));
        }
        __half2* dst = reinterpret_cast<__half2*>(qs + (size_t)qq * q_stride + d);
        dst[0] = __floats2half2_rn(v.x, v.y);
        dst[1] = __floats2half2_rn(v.z, v.w);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int qq = nt * 8 + 2 * t + h;
            obase[nt][h] = qq < nqg ? p.pair_off[p.inv[s0 + qq]] : 0ull;
        }
    __syncthreads();

    const int ksteps = (int)p.row_bytes / 16;
    const __half* qrow = qs + (size_t)g * q_stride + 4 * t;
    int s = 0;
    uint32_t parity = 0;
    for (int tile = t0; tile < t1; ++tile) {
        const int left = len - tile * R;
        const int r_lo = tile * R + g, r_hi = r_lo + 8;        // rows in the list this lane's accumulators cover
        const bool ok_lo = g < left, ok_hi = g + 8 < left;
        float sc_lo = 0.f, sc_hi = 0.f;
        bool al_lo = true, al_hi = true;
        if (t == 0) {   // one lane per row fetches the row's scale / allow bit, shared below by shuffle
            if (ok_lo) sc_lo = __ldg(p.scales + start + r_lo);
            if (ok_hi) sc_hi = __ldg(p.scales + start + r_hi);
            if (p.mask != nullptr) {
                if (ok_lo) {
                    const uint32_t row = __ldg(p.list_rows + start + r_lo);
                    al_lo = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
                }
                if (ok_hi) {
                    const uint32_t row = __ldg(p.list_rows + start + r_hi);
                    al_hi = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
                }
            }
        }
        mbar_wait(&my_bars[s], parity);
        const uint8_t* slot = my_slots + (size_t)s * tile_bytes;
        const uint8_t* a_lo = slot + (size_t)g * row_stride + 4 * t;
        const uint8_t* a_hi = a_lo + 8 * (size_t)row_stride;
        // four independent accumulator chains (k-steps 4i, 4i+1, 4i+2, 4i+3): an HMMA depends on its accumulator,
        // a single chain would serialise the 64 k-steps on the instruction's latency
        float acc4[4][NT][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc4[u][nt][i] = 0.f;
        auto kstep = [&](int ks, float (&c)[NT][4]) {
            const uint32_t w_lo = *reinterpret_cast<const uint32_t*>(a_lo + ks * 16);
            const uint32_t w_hi = *reinterpret_cast<const uint32_t*>(a_hi + ks * 16);
            uint32_t a[4];   // a0/a2: row g, bytes (0,1)/(2,3); a1/a3: row g + 8 — converted once for all n-tiles
            e4m3x4_to_f16x2x2(w_lo, a[0], a[2]);
            e4m3x4_to_f16x2x2(w_hi, a[1], a[3]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint2 bq = *reinterpret_cast<const uint2*>(qrow + (size_t)nt * 8 * q_stride + ks * 16);
                const uint32_t b[2] = {bq.x, bq.y};
                mma_m16n8k16_f16(c[nt], a, b);
            }
        };
        const int ks_full = ksteps & ~3;
        for (int ks0 = 0; ks0 < ks_full; ks0 += 4) {   // branch-free body: the fragment loads issue back to back
#pragma unroll
            for (int u = 0; u < 4; ++u) kstep(ks0 + u, acc4[u]);
        }
        for (int ks = ks_full; ks < ksteps; ++ks) kstep(ks, acc4[0]);
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = (acc4[0][nt][i] + acc4[1][nt][i]) + (acc4[2][nt][i] + acc4[3][nt][i]);
        __syncwarp();
        if (tile + STAGES < t1) issue(tile + STAGES, s);

        sc_lo = __shfl_sync(0xFFFFFFFFu, sc_lo, lane & ~3);
        sc_hi = __shfl_sync(0xFFFFFFFFu, sc_hi, lane & ~3);
        al_lo = __shfl_sync(0xFFFFFFFFu, (int)al_lo, lane & ~3) != 0;
        al_hi = __shfl_sync(0xFFFFFFFFu, (int)al_hi, lane & ~3) != 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (nt * 8 + 2 * t + h < nqg) {
                    if (ok_lo) p.scores[obase[nt][h] + (unsigned long long)r_lo] = al_lo ? acc[nt][h] * sc_lo : -INFINITY;
                    if (ok_hi) p.scores[obase[nt][h] + (unsigned long long)r_hi] = al_hi ? acc[nt][2 + h] * sc_hi : -INFINITY;
                }
            }
        if (++s == STAGES) {
            s = 0;
            parity ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------------- G4 on tcgen05 (the default)
// Blackwell's tensor cores multiply e4m3 natively (tcgen05.mma kind::f8f6f4, fp32 accumulation in TMEM), so the list
// rows go from HBM to the MMA untouched: no conversion instructions, no register staging.
//   * A (rows): 128-row x 128-byte k-blocks of the list by 2-D TMA (a u8 tensor map over list_data, SWIZZLE_128B)
//     into an 8-stage smem ring — one full 128-row tile of D = 1024 in flight per SM.
//   * B (queries): a group of QB <= 16 queries. Queries are quantised ONCE per batch to TWO e4m3 terms with one
//     fp32 scale per query — q / s = hi + lo / 16 — and both terms are MMA columns (N = 2 * QB), so the query side
//     carries ~8 mantissa bits while the row side stays single e4m3; the candidate ranking is as good as with fp16
//     queries. Two loader warps copy the group's 2 * QB byte rows into smem in the 128-byte-swizzled K-major layout
//     the UMMA descriptor expects (double-buffered: the next item's queries load while this item streams).
//   * D: 128 rows x 2 * QB fp32 columns in TMEM, double-buffered; four epilogue warps read it back (thread = row),
//     fold hi + lo / 16, apply query scale x row scale (and the allow mask / tombstones) and store to the dense
//     score buffer — for a fixed query a warp's 32 rows are 32 consecutive floats, one coalesced store.
//   * persistent CTAs (one per SM) claim work items from a global counter (lists differ in length by 30x: a fixed
//     stride left the SMs 14 % idle at the tail — ncu: smsp cycles active 86 % of elapsed): the TMA thread claims
//     the next item one item ahead and publishes it through a small smem ring that the other roles follow;
//     every role derives the item's geometry itself, the pipeline never drains between items.
// Algorithmic bytes per item = list rows x row_bytes, read once per QB = 16 queries (K4b: once per query).
namespace g4u {
constexpr int NST = 8;                 // A ring stages (16 KB each)
constexpr int A_BYTES = 128 * 128;
constexpr int THREADS = 256;           // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue, warps 6-7 query loaders
constexpr int LOADERS = 64;
constexpr int KB_MAX = 8;              // k-blocks per row: row_bytes <= 1024
constexpr int SR = 8;                  // scheduler ring entries (items claimed ahead of their consumers)
constexpr int SCHED_CONSUMERS = 7;     // MMA thread + 4 epilogue warps + 2 loader warps
constexpr uint32_t NO_ITEM = 0xFFFFFFFFu;
}  // namespace g4u

__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// q / s = hi + lo / 16 with hi, lo in e4m3 and s = max|q| / 448: out[q][0][*] = hi bytes, out[q][1][*] = lo bytes
// (row_bytes each, zero padded), qscale[q] = s. One warp per query.
__global__ void __launch_bounds__(256) quantize_queries_e4m3x2_kernel(const float* __restrict__ q32, int nq, int dim_pad,
                                                                      uint32_t row_bytes, uint8_t* __restrict__ out,
                                                                      float* __restrict__ qscale) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float* src = q32 + (size_t)q * dim_pad;
    float amax = 0.f;
    for (int i = lane; i < dim_pad; i += 32) amax = fmaxf(amax, fabsf(src[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
    const float scale = amax > 0.f ? __fdiv_rn(amax, 448.0f) : 1.0f;
    if (lane == 0) qscale[q] = scale;
    uint8_t* hi = out + (size_t)q * 2 * row_bytes;
    uint8_t* lo = hi + row_bytes;
    for (int i = lane; i < (int)row_bytes; i += 32) {
        uint8_t h = 0, l = 0;
        if (i < dim_pad) {
            const float x = __fdiv_rn(src[i], scale);
            h = (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
            const __half_raw hr = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)h, __NV_E4M3);
            const float back = __half2float(*reinterpret_cast<const __half*>(&hr));
            l = (uint8_t)__nv_cvt_float_to_fp8((x - back) * 16.0f, __NV_SATFINITE, __NV_E4M3);
        }
        hi[i] = h;
        lo[i] = l;
    }
}

struct UmmaItem {
    int nqg;          // queries in this group
    uint32_t s0;      // first table slot of the group
    int64_t start;    // first list position
    int len, tiles;   // rows, 128-row tiles
};
__device__ __forceinline__ UmmaItem umma_item(const GroupedParams& p, uint32_t item, int qb) {
    UmmaItem it;
    const int l = (int)p.item_list[item];
    const int grp = (int)(item - p.item_start[l]);
    const int c = (int)p.cnt[l];
    it.nqg = (c - grp * qb < qb) ? (c - grp * qb) : qb;
    it.s0 = p.slot_start[l] + (uint32_t)grp * qb;
    it.start = p.list_offsets[l];
    it.len = (int)(p.list_offsets[l + 1] - it.start);
    it.tiles = (it.len + 127) / 128;
    return it;
}

template <int QB>
__global__ void __launch_bounds__(g4u::THREADS, 1)
ivf_grouped_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const GroupedParams p, const uint8_t* __restrict__ q8,
                        const float* __restrict__ qscale, int has_dead) {
    using namespace g4u;
    constexpr int N = 2 * QB;                       // MMA columns: hi and lo term of every query
    constexpr int B_KB_BYTES = N * 128;             // one k-block of the query operand
    constexpr int TMEM_COLS = (2 * N < 32) ? 32 : 2 * N;
    extern __shared__ uint8_t smem_raw[];
    if (p.totals[1] != 0u) return;                  // score buffer too small for this batch: K4b does the work
    const uint32_t n_items = p.totals[0];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KB = (int)((p.row_bytes + 127) / 128);
    uint8_t* smem_a = smem;                                            // [NST][128 rows][128 B]
    uint8_t* smem_b = smem + (size_t)NST * A_BYTES;                    // [2][KB_MAX][N rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)2 * KB_MAX * B_KB_BYTES);
    uint64_t* full = bars;                 // [NST]
    uint64_t* empty = bars + NST;          // [NST]
    uint64_t* tmem_full = bars + 2 * NST;  // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint64_t* b_full = tmem_empty + 2;     // [2]
    uint64_t* b_empty = b_full + 2;        // [2]
    uint64_t* sched_full = b_empty + 2;    // [SR]
    uint64_t* sched_empty = sched_full + SR;   // [SR]
    uint32_t* sched_item = reinterpret_cast<uint32_t*>(sched_empty + SR);   // [SR]
    uint32_t* tmem_ptr_s = sched_item + SR;
    uint32_t* item_counter = const_cast<uint32_t*>(p.totals) + 2;
    // the i-th item of this CTA, as published by the TMA thread (NO_ITEM = no more work)
    auto next_item = [&](uint32_t i) -> uint32_t {
        const int slot = (int)(i % SR);
        mbar_wait_wd(&sched_full[slot], (i / SR) & 1u);
        return *reinterpret_cast<volatile uint32_t*>(&sched_item[slot]);
    };
    auto release_item = [&](uint32_t i) { mbar_arrive(&sched_empty[i % SR]); };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 128);
            mbar_init(&b_full[s], LOADERS);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < SR; ++s) {
            mbar_init(&sched_full[s], 1);
            mbar_init(&sched_empty[s], SCHED_CONSUMERS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            int stage = 0;
            uint32_t phase = 0;
            // Items are claimed two ahead and PUBLISHED one ahead of the copies: the loader warps need an item's id a
            // whole item early (metadata + 32 KB of query bytes + the swizzled smem writes must be done before the
            // item's first MMA), and the claim's atomic round trip hides behind the current item's copies.
            auto claim = [&]() -> uint32_t {
                const uint32_t c = atomicAdd(item_counter, 1u);
                return c < n_items ? c : NO_ITEM;
            };
            auto publish = [&](uint32_t i, uint32_t item) {
                const int slot = (int)(i % SR);
                mbar_wait_wd(&sched_empty[slot], ((i / SR) & 1u) ^ 1u);
                sched_item[slot] = item;
                mbar_arrive(&sched_full[slot]);          // (mbarrier arrive has release semantics: the store above is visible)
            };
            uint32_t c_a = claim(), c_b = claim();
            publish(0, c_a);
            for (uint32_t i = 0;; ++i) {
                const uint32_t item = c_a;
                if (item == NO_ITEM) break;              // its sentinel has been published already
                publish(i + 1, c_b);
                const uint32_t c_c = c_b == NO_ITEM ? NO_ITEM : claim();
                c_a = c_b;
                c_b = c_c;
                const UmmaItem it = umma_item(p, item, QB);
                for (int tile = 0; tile < it.tiles; ++tile) {
                    const int row0 = (int)(it.start + (int64_t)tile * 128);
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait_wd(&empty[stage], phase ^ 1u);
                        mbar_expect_tx(&full[stage], A_BYTES);
                        tma_load_2d(smem_a + (size_t)stage * A_BYTES, &tmap_a, kb * 128, row0, &full[stage], pol);
                        if (++stage == NST) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = f32 at [4,6); A / B format at [7,10) / [10,13): 0 = e4m3; K-major; N>>3 at
            // [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t tcount = 0;
            for (uint32_t icount = 0;; ++icount) {
                const uint32_t item = next_item(icount);
                if (item == NO_ITEM) break;
                const UmmaItem it = umma_item(p, item, QB);
                release_item(icount);
                const int bbuf = (int)(icount & 1u);
                mbar_wait_wd(&b_full[bbuf], (icount >> 1) & 1u);
                tcgen05_fence_after();
                const uint8_t* bq = smem_b + (size_t)bbuf * KB_MAX * B_KB_BYTES;
                for (int tile = 0; tile < it.tiles; ++tile, ++tcount) {
                    const int acc = (int)(tcount & 1u);
                    mbar_wait_wd(&tmem_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait_wd(&full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = umma_smem_desc(smem_a + (size_t)stage * A_BYTES);
                        const uint64_t db = umma_smem_desc(bq + (size_t)kb * B_KB_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)   // K = 32 e4m3 = 32 bytes per instruction: +2 in the >>4 address field
                            umma_f8(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        tcgen05_commit(&empty[stage]);
                        if (kb == KB - 1) tcgen05_commit(&tmem_full[acc]);
                        if (++stage == NST) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
                tcgen05_commit(&b_empty[bbuf]);   // the item's MMAs have read the queries: the buffer may be refilled
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue: TMEM -> scaled scores -> dense score buffer =====================
        const int lane_base = 32 * (warp & 3);            // TMEM lanes this warp may touch
        uint32_t tcount = 0;
        for (uint32_t icount = 0;; ++icount) {
            const uint32_t item = next_item(icount);
            if (item == NO_ITEM) break;
            const UmmaItem it = umma_item(p, item, QB);
            __syncwarp();
            if (lane == 0) release_item(icount);
            unsigned long long obase[QB];
            float qs[QB];
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
                obase[qq] = 0ull;
                qs[qq] = 0.f;
                if (qq < it.nqg) {
                    const uint32_t e = p.inv[it.s0 + qq];
                    obase[qq] = p.pair_off[e];
                    qs[qq] = __ldg(qscale + e / (uint32_t)p.nprobe);
                }
            }
            for (int tile = 0; tile < it.tiles; ++tile, ++tcount) {
                const int acc = (int)(tcount & 1u);
                const int r = tile * 128 + lane_base + lane;           // row inside the list
                const bool valid = r < it.len;
                float rscale = 0.f;
                bool allowed = valid;
                if (valid) {
                    rscale = __ldg(p.scales + it.start + r);
                    if (p.mask != nullptr || has_dead) {
                        const uint32_t row = __ldg(p.list_rows + it.start + r);
                        allowed = row != TS_DEAD_ROW;
                        if (allowed && p.mask != nullptr) allowed = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
                    }
                }
                mbar_wait_wd(&tmem_full[acc], (tcount >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(acc * N);
                uint32_t v[N];
                if constexpr (N == 32) {
                    tmem_ld_32x32b_x32(taddr, v);
                } else {
                    tmem_ld_32x32b_x16(taddr, v);
                }
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(&tmem_empty[acc]);                         // accumulator drained: the next tile may start
                if (valid) {
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq) {
                        if (qq < it.nqg) {
                            const float sc = fmaf(__uint_as_float(v[QB + qq]), 0.0625f, __uint_as_float(v[qq])) * (qs[qq] * rscale);
                            p.scores[obase[qq] + (unsigned long long)r] = allowed ? sc : -INFINITY;
                        }
                    }
                }
            }
        }
    } else {
        // ===================== query loaders: the group's 2 * QB byte rows -> swizzled K-major smem =====================
        const int lt = threadIdx.x - 192;                 // 0..63
        const int chunks_per_row = KB * 8;                // 16-byte chunks
        for (uint32_t icount = 0;; ++icount) {
            const uint32_t item = next_item(icount);
            if (item == NO_ITEM) break;
            const UmmaItem it = umma_item(p, item, QB);
            __syncwarp();
            if (lane == 0) release_item(icount);
            const int bbuf = (int)(icount & 1u);
            mbar_wait_wd(&b_empty[bbuf], ((icount >> 1) & 1u) ^ 1u);
            uint8_t* bq = smem_b + (size_t)bbuf * KB_MAX * B_KB_BYTES;
            for (int c = lt; c < N * chunks_per_row; c += LOADERS) {
                const int n = c / chunks_per_row, ch = c % chunks_per_row;    // operand row, 16-byte chunk in the row
                const int qq = n < QB ? n : n - QB;
                uint4 val = make_uint4(0u, 0u, 0u, 0u);
                if (qq < it.nqg && (uint32_t)ch * 16u < p.row_bytes) {
                    const uint32_t e = p.inv[it.s0 + qq];
                    const size_t qi = (size_t)(e / (uint32_t)p.nprobe);
                    val = __ldg(reinterpret_cast<const uint4*>(q8 + (qi * 2 + (n < QB ? 0 : 1)) * p.row_bytes + (size_t)ch * 16));
                }
                const int kb = ch >> 3, ci = ch & 7;
                // SWIZZLE_128B, K-major: 8-row groups 1024 B apart, rows 128 B apart, chunk index XOR (row mod 8)
                uint8_t* dst = bq + (size_t)kb * B_KB_BYTES + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 +
                               (size_t)((ci ^ (n & 7)) * 16);
                *reinterpret_cast<uint4*>(dst) = val;
            }
            fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core's reads
            mbar_arrive(&b_full[bbuf]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Per query: best k keys over the score runs of its probed lists. The whole CTA walks the runs 1024 rows at a
// time (one 16-byte load per thread, the next step's load issued before the current one is consumed); keys that
// beat the running threshold are appended to one shared buffer (warp-aggregated slot allocation); when the buffer
// could overflow in the next step the CTA sorts it cooperatively, keeps the best k and raises the threshold to the
// k-th key. After the first compaction only ~k ln(n / 1024) more keys ever pass, so a query costs two CTA sorts.
// Warp-per-query form (the default): the query's score runs are walked by one warp with float4 loads against a float
// threshold, candidates go through a register-resident WarpSelect (warp_select_run) — no shared memory, no barriers, no
// CTA sorts; 4096 queries run as 4096 independent warps. 0.93 ms -> see profiles/launches_ivf_batch_r2*.csv.
template <int KPL>
__global__ void __launch_bounds__(256, 4) ivf_select_warp_kernel(const float* __restrict__ scores,
                                                              const unsigned long long* __restrict__ pair_off,
                                                              const uint32_t* __restrict__ pair_len,
                                                              const uint32_t* __restrict__ pair_pos0, int nq, int nprobe, int k,
                                                              const uint32_t* __restrict__ totals, uint64_t* __restrict__ cand) {
    if (totals[1] != 0u) return;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpSelect<KPL> sel;
    sel.init();
    for (int j = 0; j < nprobe; ++j) {
        const size_t e = (size_t)q * nprobe + j;
        const int len = (int)pair_len[e];
        if (len == 0) continue;
        warp_select_run<KPL>(sel, scores + pair_off[e], len, pair_pos0[e], k, lane);
    }
    sel.flush(k, lane);
#pragma unroll
    for (int j = 0; j < KPL; ++j)
        if (j * 32 + lane < k) cand[(size_t)q * k + j * 32 + lane] = sel.best.key[j];
}

constexpr int SEL_CAP = 3072;   // keys the buffer holds (24 KB); a step appends at most 1024
__global__ void __launch_bounds__(256) ivf_select_kernel(const float* __restrict__ scores,
                                                         const unsigned long long* __restrict__ pair_off,
                                                         const uint32_t* __restrict__ pair_len,
                                                         const uint32_t* __restrict__ pair_pos0, int nprobe, int k,
                                                         const uint32_t* __restrict__ totals,
                                                         uint64_t* __restrict__ cand) {
    __shared__ __align__(16) uint64_t buf[4096];   // sort space: SEL_CAP keys padded to a power of two
    __shared__ int s_cnt;
    __shared__ unsigned long long s_thr;
    if (totals[1] != 0u) return;
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_thr = 0ull;
    }
    __syncthreads();
    auto compact = [&]() {   // all threads: sort buf[0, s_cnt), keep k, publish the threshold
        const int n = s_cnt;
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = n + threadIdx.x; i < P; i += blockDim.x) buf[i] = 0ull;
        __syncthreads();
        cta_bitonic_sort_desc(buf, P);
        if (threadIdx.x == 0) {
            s_cnt = n < k ? n : k;
            s_thr = n >= k ? buf[k - 1] : 0ull;
        }
        __syncthreads();
    };
    for (int j = 0; j < nprobe; ++j) {
        const size_t e = (size_t)q * nprobe + j;
        const int len = (int)pair_len[e];
        if (len == 0) continue;
        const float* run = scores + pair_off[e];
        const uint32_t pos0 = pair_pos0[e];
        const int r_me = 4 * (int)threadIdx.x;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r_me < len) nxt = __ldg(reinterpret_cast<const float4*>(run + r_me));
        for (int r0 = 0; r0 < len; r0 += 1024) {
            const float4 cur = nxt;
            const int r = r0 + r_me;
            if (r + 1024 < len) nxt = __ldg(reinterpret_cast<const float4*>(run + r + 1024));
            const unsigned long long thr = s_thr;
            const float f[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // masked rows carry -inf; the padded tail of a run is unwritten memory and is cut by r + i < len
                const bool live = (r + i < len) && f[i] > -INFINITY;
                const uint64_t key = live ? pack_key(f[i], pos0 + (uint32_t)(r + i)) : 0ull;
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
            if (s_cnt > SEL_CAP - 1024) compact();   // uniform: every thread reads the same counter after the barrier
        }
    }
    compact();
    for (int i = threadIdx.x; i < k; i += blockDim.x) cand[(size_t)q * k + i] = (i < s_cnt) ? buf[i] : 0ull;
}

// ------------------------------------------------------------------------------------ host side
static inline size_t g_al(size_t v) { return (v + 255) / 256 * 256; }

size_t ivf_grouped_score_cap(const ts_index* ix, int nq, int nprobe) {
    const double avg = (double)ix->size / (double)std::max(ix->nlist, 1);
    double want = 2.0 * (double)nq * (double)nprobe * (avg + 1.0) + 1048576.0;   // 2x the balanced-list volume
    const double max_floats = 8.0 * 1024 * 1024 * 1024 / 4.0;                    // never more than 8 GiB
    if (want > max_floats) want = max_floats;
    return (size_t)want;
}

// Sized for the mutated index as well: overflow lists double the list count (virtual lists nlist .. 2 nlist - 1) and
// the probes per query; the score buffer follows the real probe count (overflow lists hold < 10 % of the rows).
size_t ivf_grouped_workspace_bytes(const ts_index* ix, int nq, int nprobe) {
    const size_t np = (size_t)nq * nprobe * 2, nl = (size_t)ix->nlist * 2;
    size_t b = 0;
    b += g_al(nl * 4) * 2;                 // cnt, cursor
    b += g_al((nl + 1) * 4) * 2;           // slot_start, item_start
    b += g_al((np / g4::QB + std::min(nl, np) + 1) * 4);   // item_list (sized for the smaller group width)
    b += g_al((nl + 1) * 8);               // base
    b += 256;                              // totals
    b += g_al(np * 4) * 3;                 // inv, pair_len, pair_pos0
    b += g_al(np * 8);                     // pair_off
    b += g_al((size_t)nq * 2 * ix->list_row_bytes) + g_al((size_t)nq * 4);   // two-term e4m3 queries + their scales
    b += g_al(ivf_grouped_score_cap(ix, nq, nprobe) * 4);
    return b;
}

bool ivf_grouped_supported(const ts_index* ix, int kc) {
    return ix->list_dtype == TS_FP8_E4M3 && ix->list_row_bytes <= 1024 && kc <= 256;
}

// Enqueues G1..G5. cand[nq][kc] receives the candidates unless the device-side capacity flag trips; `flag_out`
// points at that flag (1 = the caller's K4b launch must do the work instead).
// `vmul` = 2 when the index carries overflow lists: `probes` then holds 2 * nprobe keys per query (every probed list and
// its overflow list) and lists are numbered 0 .. 2 nlist - 1.
int launch_ivf_grouped(const ts_index* ix, const uint64_t* probes, const float* q32, int nq, int nprobe_real, int vmul, int kc,
                       const uint32_t* allow_mask, void* workspace, uint64_t* cand, const uint32_t** flag_out,
                       cudaStream_t s) {
    const int nprobe = nprobe_real * vmul;                   // probes per query as the kernels see them
    const int nlist_v = ix->nlist * vmul;
    const size_t np_max = (size_t)nq * nprobe_real * 2, nl_max = (size_t)ix->nlist * 2;   // the carve-out is the same for both
    const size_t np = (size_t)nq * nprobe, nl = (size_t)nlist_v;
    TS_REQUIRE(np_max < ((size_t)1 << 31), TS_ERR_UNSUPPORTED, "ivf grouped scan: nq * nprobe = %zu too large", np_max);
    char* w = (char*)workspace;
    auto take = [&](size_t bytes) {
        char* p = w;
        w += g_al(bytes);
        return p;
    };
    uint32_t* cnt = (uint32_t*)take(nl_max * 4);
    uint32_t* cursor = (uint32_t*)take(nl_max * 4);
    uint32_t* slot_start = (uint32_t*)take((nl_max + 1) * 4);
    uint32_t* item_start = (uint32_t*)take((nl_max + 1) * 4);
    const size_t max_items_alloc = np_max / g4::QB + std::min(nl_max, np_max) + 1;
    const size_t max_items = np / g4::QB + std::min(nl, np) + 1;   // upper bound on sum ceil(cnt/QB)
    uint32_t* item_list = (uint32_t*)take(max_items_alloc * 4);
    unsigned long long* base = (unsigned long long*)take((nl_max + 1) * 8);
    uint32_t* totals = (uint32_t*)take(256);
    uint32_t* inv = (uint32_t*)take(np_max * 4);
    uint32_t* pair_len = (uint32_t*)take(np_max * 4);
    uint32_t* pair_pos0 = (uint32_t*)take(np_max * 4);
    unsigned long long* pair_off = (unsigned long long*)take(np_max * 8);
    uint8_t* q8 = (uint8_t*)take((size_t)nq * 2 * ix->list_row_bytes);
    float* qscale = (float*)take((size_t)nq * 4);
    const size_t cap = ivf_grouped_score_cap(ix, nq, nprobe_real);
    float* scores = (float*)take(cap * 4);
    *flag_out = totals + 1;

    TS_CHECK_CUDA(cudaMemsetAsync(cnt, 0, g_al(nl_max * 4) * 2, s));   // cnt and cursor are adjacent
    const int pb = (int)std::min<size_t>((np + 255) / 256, 148 * 8);
    ivf_invert_count_kernel<<<pb, 256, 0, s>>>(probes, (int)np, cnt);
    TS_LAUNCH_CHECK();
    // ivf.group_mma: 3 / 4 = tcgen05 kind::f8f6f4 with 16 / 8 queries per group (3 is the default), 1 / 2 = legacy
    // mma.sync with 8 / 16, 0 = CUDA cores with 4
    const int mode = tunables().ivf_group_mma;
    const bool use_umma = mode >= 3;
    const bool use_mma = mode == 1 || mode == 2;
    const int qb = use_umma ? (mode == 3 ? 16 : 8) : use_mma ? (mode == 2 ? 16 : 8) : g4::QB;
    ivf_invert_scan_kernel<<<1, 1024, 0, s>>>(cnt, ix->list_offsets, nlist_v, slot_start, item_start, base, totals,
                                              (unsigned long long)cap, qb);
    TS_LAUNCH_CHECK();
    ivf_invert_fill_kernel<<<pb, 256, 0, s>>>(probes, (int)np, slot_start, base, ix->list_offsets, cursor, inv, pair_off,
                                              pair_len, pair_pos0);
    TS_LAUNCH_CHECK();
    ivf_item_table_kernel<<<(nlist_v + 255) / 256, 256, 0, s>>>(cnt, item_start, nlist_v, item_list, qb);
    TS_LAUNCH_CHECK();

    GroupedParams p;
    p.list_data = (const uint8_t*)ix->list_data;
    p.row_bytes = ix->list_row_bytes;
    p.dim_pad = ix->dim_pad;
    p.scales = ix->list_scales;
    p.list_offsets = ix->list_offsets;
    p.queries = q32;
    p.nprobe = nprobe;
    p.nlist = nlist_v;
    p.mask = allow_mask;
    p.list_rows = ix->list_rows;
    p.cnt = cnt;
    p.slot_start = slot_start;
    p.item_start = item_start;
    p.item_list = item_list;
    p.inv = inv;
    p.pair_off = pair_off;
    p.totals = totals;
    p.scores = scores;
    p.stages = 2;
    if (use_umma) {
        quantize_queries_e4m3x2_kernel<<<(nq + 7) / 8, 256, 0, s>>>(q32, nq, ix->dim_pad, p.row_bytes, q8, qscale);
        TS_LAUNCH_CHECK();
        CUtensorMap tmap;
        {
            auto enc = get_encode_fn();
            TS_REQUIRE(enc != nullptr, TS_ERR_CUDA, "ivf grouped scan: cuTensorMapEncodeTiled entry point unavailable");
            cuuint64_t dims[2] = {(cuuint64_t)p.row_bytes, (cuuint64_t)std::max<int64_t>(ix->built_n + ix->ovf_n, 1)};
            cuuint64_t strides[1] = {(cuuint64_t)p.row_bytes};
            cuuint32_t box[2] = {128, 128};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>((const void*)p.list_data), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            TS_REQUIRE(r == CUDA_SUCCESS, TS_ERR_CUDA, "ivf grouped scan: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        }
        const int n_cols = 2 * qb;
        const size_t smem = (size_t)g4u::NST * g4u::A_BYTES + (size_t)2 * g4u::KB_MAX * n_cols * 128 + 64 * sizeof(uint64_t) + 1024;
        const unsigned grid = (unsigned)std::min<size_t>(max_items, (size_t)sm_count(ix->device));
        const int has_dead = ix->ivf_dead > 0 ? 1 : 0;
        if (qb == 16) {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_umma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_umma_kernel<16><<<grid, g4u::THREADS, smem, s>>>(tmap, p, q8, qscale, has_dead);
        } else {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_umma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_umma_kernel<8><<<grid, g4u::THREADS, smem, s>>>(tmap, p, q8, qscale, has_dead);
        }
    } else if (use_mma) {
        const int nt = qb / 8;
        const int warps = g4m::warps_for(nt);
        const size_t tile_bytes = (size_t)g4m::R * (p.row_bytes + 16);
        const size_t smem = (size_t)warps * g4m::STAGES * tile_bytes + 8 * warps * g4m::STAGES +
                            (size_t)qb * (p.row_bytes + 16) * sizeof(__half);
        if (nt == 1) {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_mma_kernel<1><<<(unsigned)max_items, warps * 32, smem, s>>>(p);
        } else {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_mma_kernel<2><<<(unsigned)max_items, warps * 32, smem, s>>>(p);
        }
    } else {
        const int warps = 12;   // 152 registers x 384 threads fill the register file: 3 warps per scheduler
        const size_t tile_bytes = (size_t)g4::R * p.row_bytes;
        const size_t smem = (size_t)warps * p.stages * tile_bytes + 8 * (size_t)warps * p.stages;
        if (p.row_bytes <= 512) {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_scan_kernel<1><<<(unsigned)max_items, warps * 32, smem, s>>>(p);
        } else {
            TS_CHECK_CUDA(cudaFuncSetAttribute(ivf_grouped_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ivf_grouped_scan_kernel<2><<<(unsigned)max_items, warps * 32, smem, s>>>(p);
        }
    }
    TS_LAUNCH_CHECK();

    if (tunables().ivf_select_warp == 0)
        ivf_select_kernel<<<nq, 256, 0, s>>>(scores, pair_off, pair_len, pair_pos0, nprobe, kc, totals, cand);
    else if (kc <= 32)
        ivf_select_warp_kernel<1><<<(nq + 7) / 8, 256, 0, s>>>(scores, pair_off, pair_len, pair_pos0, nq, nprobe, kc, totals, cand);
    else if (kc <= 128)
        ivf_select_warp_kernel<4><<<(nq + 7) / 8, 256, 0, s>>>(scores, pair_off, pair_len, pair_pos0, nq, nprobe, kc, totals, cand);
    else
        ivf_select_warp_kernel<8><<<(nq + 7) / 8, 256, 0, s>>>(scores, pair_off, pair_len, pair_pos0, nq, nprobe, kc, totals, cand);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // namespace ts
