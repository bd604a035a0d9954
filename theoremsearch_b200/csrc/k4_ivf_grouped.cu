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

// Per query: best k keys over the score runs of its probed lists. The whole CTA walks the runs 1024 rows at a
// time (one 16-byte load per thread, the next step's load issued before the current one is consumed); keys that
// beat the running threshold are appended to one shared buffer (warp-aggregated slot allocation); when the buffer
// could overflow in the next step the CTA sorts it cooperatively, keeps the best k and raises the threshold to the
// k-th key. After the first compaction only ~k ln(n / 1024) more keys ever pass, so a query costs two CTA sorts.
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

size_t ivf_grouped_workspace_bytes(const ts_index* ix, int nq, int nprobe) {
    const size_t np = (size_t)nq * nprobe, nl = (size_t)ix->nlist;
    size_t b = 0;
    b += g_al(nl * 4) * 2;                 // cnt, cursor
    b += g_al((nl + 1) * 4) * 2;           // slot_start, item_start
    b += g_al((np / g4::QB + std::min(nl, np) + 1) * 4);   // item_list (sized for the smaller group width)
    b += g_al((nl + 1) * 8);               // base
    b += 256;                              // totals
    b += g_al(np * 4) * 3;                 // inv, pair_len, pair_pos0
    b += g_al(np * 8);                     // pair_off
    b += g_al(ivf_grouped_score_cap(ix, nq, nprobe) * 4);
    return b;
}

bool ivf_grouped_supported(const ts_index* ix, int kc) {
    return ix->list_dtype == TS_FP8_E4M3 && ix->list_row_bytes <= 1024 && kc <= 256;
}

// Enqueues G1..G5. cand[nq][kc] receives the candidates unless the device-side capacity flag trips; `flag_out`
// points at that flag (1 = the caller's K4b launch must do the work instead).
int launch_ivf_grouped(const ts_index* ix, const uint64_t* probes, const float* q32, int nq, int nprobe, int kc,
                       const uint32_t* allow_mask, void* workspace, uint64_t* cand, const uint32_t** flag_out,
                       cudaStream_t s) {
    const size_t np = (size_t)nq * nprobe, nl = (size_t)ix->nlist;
    TS_REQUIRE(np < ((size_t)1 << 31), TS_ERR_UNSUPPORTED, "ivf grouped scan: nq * nprobe = %zu too large", np);
    char* w = (char*)workspace;
    auto take = [&](size_t bytes) {
        char* p = w;
        w += g_al(bytes);
        return p;
    };
    uint32_t* cnt = (uint32_t*)take(nl * 4);
    uint32_t* cursor = (uint32_t*)take(nl * 4);
    uint32_t* slot_start = (uint32_t*)take((nl + 1) * 4);
    uint32_t* item_start = (uint32_t*)take((nl + 1) * 4);
    const size_t max_items = np / g4::QB + std::min(nl, np) + 1;   // upper bound on sum ceil(cnt/QB)
    uint32_t* item_list = (uint32_t*)take(max_items * 4);
    unsigned long long* base = (unsigned long long*)take((nl + 1) * 8);
    uint32_t* totals = (uint32_t*)take(256);
    uint32_t* inv = (uint32_t*)take(np * 4);
    uint32_t* pair_len = (uint32_t*)take(np * 4);
    uint32_t* pair_pos0 = (uint32_t*)take(np * 4);
    unsigned long long* pair_off = (unsigned long long*)take(np * 8);
    const size_t cap = ivf_grouped_score_cap(ix, nq, nprobe);
    float* scores = (float*)take(cap * 4);
    *flag_out = totals + 1;

    TS_CHECK_CUDA(cudaMemsetAsync(cnt, 0, g_al(nl * 4) * 2, s));   // cnt and cursor are adjacent
    const int pb = (int)std::min<size_t>((np + 255) / 256, 148 * 8);
    ivf_invert_count_kernel<<<pb, 256, 0, s>>>(probes, (int)np, cnt);
    TS_LAUNCH_CHECK();
    const bool use_mma = tunables().ivf_group_mma != 0;
    const int qb = use_mma ? (tunables().ivf_group_mma >= 2 ? 16 : 8) : g4::QB;
    ivf_invert_scan_kernel<<<1, 1024, 0, s>>>(cnt, ix->list_offsets, ix->nlist, slot_start, item_start, base, totals,
                                              (unsigned long long)cap, qb);
    TS_LAUNCH_CHECK();
    ivf_invert_fill_kernel<<<pb, 256, 0, s>>>(probes, (int)np, slot_start, base, ix->list_offsets, cursor, inv, pair_off,
                                              pair_len, pair_pos0);
    TS_LAUNCH_CHECK();
    ivf_item_table_kernel<<<(ix->nlist + 255) / 256, 256, 0, s>>>(cnt, item_start, ix->nlist, item_list, qb);
    TS_LAUNCH_CHECK();

    GroupedParams p;
    p.list_data = (const uint8_t*)ix->list_data;
    p.row_bytes = ix->list_row_bytes;
    p.dim_pad = ix->dim_pad;
    p.scales = ix->list_scales;
    p.list_offsets = ix->list_offsets;
    p.queries = q32;
    p.nprobe = nprobe;
    p.nlist = ix->nlist;
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
    if (use_mma) {
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

    ivf_select_kernel<<<nq, 256, 0, s>>>(scores, pair_off, pair_len, pair_pos0, nprobe, kc, totals, cand);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // namespace ts
