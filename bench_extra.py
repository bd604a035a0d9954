#!/usr/bin/env python
"""Secondary measurements (not the driver's contract — that is bench.py): the other BASELINE.json
configs and tuning sweeps.  Each sub-command prints one JSON line per measurement.

  python bench_extra.py batched   [--rows 10000000 --nq 4096 --k 100]     configs[2]: K3 tcgen05 GEMM + top-k
  python bench_extra.py small-batch [--rows 10000000]                     nq = 1..4096 latency curve
  python bench_extra.py sweep-scan [--rows 10000000]                      K2 tunables sweep
  python bench_extra.py shard-stream [--rows 1250000 --iters 400]          per-query cost of every sharded search form at the shard size (1 GPU)
  python bench_extra.py k1                                                K1 normalise+quantise throughput
  python bench_extra.py torch-compare                                     library comparator: torch.topk(corpus @ q) vs K2
  python bench_extra.py cfg0                                              configs[0] (73 queries x 100k x 1024 fp32, top-10):
                                                                          host-buffer calls beside the reference expression on the CPU
  python bench_extra.py fp8-scan  [--rows 10000000]                       exhaustive e4m3 scan + exact re-score
  python bench_extra.py ivf [--rows 40000000 --nlist 16384 --nprobe 32 --rescore 100 --data clustered|gaussian]
                                                                          configs[4] shape on one GPU: IVF-Flat fp8 + rescore
  python bench_extra.py ivf-q1 / ivf-q1-sweep                             minimal runs for ncu / K4b tunables + phase timeline
  torchrun --nproc-per-node 8 bench_extra.py sharded [--rows 100000000 --nlist 16384]
                                                                          configs[3]+[4] at full size (exact + IVF, sharded)
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import theoremsearch_b200 as ts  # noqa: E402
from theoremsearch_b200 import synthetic  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0,
                                                          "bf16_tflops_sustained": 1400.0}


def build(rows, dim, dev):
    index = ts.TheoremIndex(dim, rows, dtype="bf16", device=dev)
    synthetic.fill_index(index, 0, rows, seed=0)
    torch.cuda.synchronize()
    return index


_CLOCKS = None     # bench.ClockSampler of this process; every timed region samples SM clocks / throttle reasons


class _Sampling:
    """`with _Sampling():` samples clocks for the enclosed timed region (no-op before main() set the sampler up)."""

    def __enter__(self):
        if _CLOCKS is not None:
            _CLOCKS.__enter__()

    def __exit__(self, *a):
        if _CLOCKS is not None:
            _CLOCKS.__exit__(*a)


_json_dumps = json.dumps


def _dumps_with_clocks(obj, *a, **k):
    """Every result line carries the clocks of the most recent timed region (a number without them is unusable)."""
    if isinstance(obj, dict) and "bench" in obj and "clocks" not in obj and _CLOCKS is not None:
        obj = dict(obj)
        obj["clocks"] = _CLOCKS.summary()
    return _json_dumps(obj, *a, **k)


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with _Sampling():
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cmd_batched(a):
    dev = torch.device("cuda", 0)
    index = build(a.rows, a.dim, dev)
    q = synthetic.make_queries(a.nq, a.dim, dev)
    for name, val in (a.tunable or []):
        ts.set_tunable(name, int(val))
    from bench import ClockSampler
    l0 = ts.kernel_launches()
    with ClockSampler(0) as clocks:
        ms = timed(lambda: index.search(q, a.k), a.warmup, a.iters)
    launches = (ts.kernel_launches() - l0) // (a.warmup + a.iters)
    fix = ts.last_batched_fixups()
    flops = 2.0 * a.nq * a.rows * a.dim
    pk = peaks()
    tf = flops / (ms * 1e-3) / 1e12
    print(json.dumps({"bench": "batched", "rows": a.rows, "dim": a.dim, "nq": a.nq, "k": a.k, "ms_per_batch": ms,
                      "queries_per_s": a.nq / (ms * 1e-3), "tflops": tf,
                      "frac_of_measured_bf16_burst": tf / pk["bf16_tflops"],
                      "frac_of_measured_bf16_sustained": tf / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                      "launches_per_batch": launches, "tunables": a.tunable, "fixups": fix, "iters": a.iters,
                      "clocks": clocks.summary()}))


def cmd_sweep_scan(a):
    dev = torch.device("cuda", 0)
    index = build(a.rows, a.dim, dev)
    q = synthetic.make_queries(64, a.dim, dev)
    pk = peaks()
    nbytes = a.rows * a.dim * 2
    grid = list(itertools.product((1, 2), (6, 8, 12, 16), (2, 3), (2, 4, 8)))
    iters = 40
    if a.fine:   # the near-optimal region, more iterations, interleaved twice to expose drift
        grid = [(1, 12, 2, 2), (1, 10, 2, 2), (1, 14, 2, 2), (1, 16, 2, 2), (1, 6, 2, 4), (1, 7, 2, 4), (1, 8, 2, 4),
                (2, 6, 2, 2), (2, 8, 2, 2), (2, 3, 2, 4), (2, 4, 2, 4), (1, 3, 2, 8), (1, 4, 2, 8)] * 2
        iters = 150
    for ctas, warps, stages, rows in grid:
        if ctas * warps > 16 or ctas * warps * stages * rows * a.dim * 2 > 200 * 1024:
            continue
        ts.set_tunable("scan.tile_rows", rows)
        ts.set_tunable("scan.ctas_per_sm", ctas)
        ts.set_tunable("scan.warps", warps)
        ts.set_tunable("scan.stages", stages)
        it = iter(itertools.cycle(range(64)))
        try:
            ms = timed(lambda: index.search(q[next(it)], a.k), 5, iters)
        except ts.TheoremSearchError as e:
            print(json.dumps({"bench": "sweep-scan", "ctas": ctas, "warps": warps, "stages": stages, "rows": rows,
                              "error": str(e)}))
            continue
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"bench": "sweep-scan", "ctas": ctas, "warps": warps, "stages": stages, "rows": rows, "ms": ms,
                          "gbs_incl_merge": gbs, "frac_measured": gbs / pk["hbm_gbs"]}))


def cmd_small_batch(a):
    dev = torch.device("cuda", 0)
    index = build(a.rows, a.dim, dev)
    q = synthetic.make_queries(4096, a.dim, dev)
    for nq in (1, 2, 4, 8, 16, 64, 256, 1024, 4096):
        ms = timed(lambda: index.search(q[:nq], a.k), 3, 10 if nq < 1024 else 3)
        print(json.dumps({"bench": "small-batch", "nq": nq, "k": a.k, "ms": ms, "queries_per_s": nq / (ms * 1e-3),
                          "corpus_gbs": a.rows * a.dim * 2 / (ms * 1e-3) / 1e9}))


def cmd_cfg0(a):
    """configs[0] at the reference's own shape. GPU side: host fp32 queries in, host results out through
    ``ts_search_host`` (copies inside the timed region, wall clock). CPU side (oracle = the checker, used here as
    the reported baseline only): ``util.cos_sim`` + ``argsort`` per query (test_app.py:76-77), the batched form
    (compare_embeddings.py:61,105) and the pgvector-shaped form (rows normalised at write time, dot + top-k)."""
    import time

    import numpy as np

    from oracle import oracle
    n, d, nq, k = 100_000, a.dim, 73, 10
    rows = oracle.synthetic_rows(0, n, d, seed=0) * 3.0
    queries = oracle.synthetic_queries(nq, d) * 0.25
    out = {"bench": "cfg0", "rows": n, "dim": d, "nq": nq, "k": k, "host_threads": torch.get_num_threads(),
           "host_cores": os.cpu_count()}

    t0 = time.perf_counter()
    index32 = ts.build_index(rows, dtype="f32", normalize=True)
    torch.cuda.synchronize()
    out["gpu_build_f32_ms_incl_h2d"] = (time.perf_counter() - t0) * 1e3
    index16 = ts.build_index(rows, dtype="bf16", normalize=True)
    torch.cuda.synchronize()

    def wall(fn, warmup, iters):
        for _ in range(warmup):
            fn()
        lat = []
        for _ in range(iters):
            t = time.perf_counter()
            fn()
            lat.append(time.perf_counter() - t)
        return np.array(lat)

    for name, index in (("f32_rows", index32), ("bf16_rows", index16)):
        it = itertools.cycle(range(nq))
        with _Sampling():
            lat = wall(lambda: index.search_host(queries[next(it)], k), 10, 3 * nq)
        out[f"gpu_{name}_single_p50_ms"] = float(np.median(lat) * 1e3)
        out[f"gpu_{name}_single_qps"] = float(1.0 / lat.mean())
        with _Sampling():
            lat = wall(lambda: index.search_host(queries, k), 3, 200)
        out[f"gpu_{name}_batch73_ms"] = float(np.median(lat) * 1e3)
        out[f"gpu_{name}_batch73_qps"] = float(nq / np.median(lat))

    rows_t, q_t = torch.from_numpy(rows), torch.from_numpy(queries)
    lat = wall(lambda i=itertools.cycle(range(nq)): oracle.reference_single_query_verbatim(q_t[next(i)], rows_t, k), 1, 12)
    out["cpu_reference_single_p50_ms"] = float(np.median(lat) * 1e3)
    out["cpu_reference_single_qps"] = float(1.0 / lat.mean())
    lat = wall(lambda: oracle.batched_ranking(q_t, rows_t, k), 0, 2)
    out["cpu_reference_batch73_ms"] = float(np.median(lat) * 1e3)
    stored = oracle.normalize(rows_t)
    qn = oracle.normalize(q_t)

    def pg_shaped(i=itertools.cycle(range(nq))):
        sc = torch.mv(stored, qn[next(i)])
        return torch.topk(sc, k)
    lat = wall(pg_shaped, 2, 40)
    out["cpu_prenormalised_dot_topk_p50_ms"] = float(np.median(lat) * 1e3)
    out["speedup_single_query_vs_reference"] = out["cpu_reference_single_p50_ms"] / out["gpu_f32_rows_single_p50_ms"]
    out["speedup_batch73_vs_reference"] = out["cpu_reference_batch73_ms"] / out["gpu_f32_rows_batch73_ms"]
    print(json.dumps(out))


def cmd_torch_compare(a):
    """Optional GPU comparator of SURVEY §8(d): the same single-query top-10 done with library calls on the same
    box — ``torch.topk(corpus_bf16 @ q_bf16, 10)`` (cuBLAS GEMV writing N scores + a separate top-k kernel) —
    beside K2 on an index of the same shape. Library arithmetic is bf16 x bf16 (the query is rounded), so it is
    a speed comparator only, not a parity one."""
    dev = torch.device("cuda", 0)
    n, d = a.rows, a.dim
    x = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for lo in range(0, n, 1 << 20):
        hi = min(n, lo + (1 << 20))
        blk = torch.randn((hi - lo, d), generator=g, device=dev)
        x[lo:hi] = torch.nn.functional.normalize(blk, dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn((64, d), generator=g, device=dev), dim=1)
    qb = q.to(torch.bfloat16)
    it = itertools.cycle(range(64))
    ms_mv = timed(lambda: torch.mv(x, qb[next(it)]), 5, 40)
    ms_lib = timed(lambda: torch.topk(torch.mv(x, qb[next(it)]), a.k), 5, 40)
    index = ts.TheoremIndex(d, n, dtype="bf16", device=dev)
    for lo in range(0, n, 1 << 20):
        index.add(x[lo:min(n, lo + (1 << 20))], normalize=False)
    ms_k2 = timed(lambda: index.search(q[next(it)], a.k, normalize=False), 5, 40)
    s_lib, i_lib = torch.topk(torch.mv(x, qb[0]).float(), a.k)
    s_k2, i_k2 = index.search(q[0], a.k, normalize=False)
    nbytes = n * d * 2
    print(json.dumps({"bench": "torch-compare", "rows": n, "dim": d, "k": a.k,
                      "torch_mv_ms": ms_mv, "torch_mv_topk_ms": ms_lib, "k2_scan_topk_ms": ms_k2,
                      "torch_gbs": nbytes / (ms_lib * 1e-3) / 1e9, "k2_gbs": nbytes / (ms_k2 * 1e-3) / 1e9,
                      "speedup_vs_torch": ms_lib / ms_k2,
                      "top10_overlap_with_bf16_query_library_result": len(set(i_lib.tolist()) & set(i_k2[0].tolist()))}))


def timed_graph(fn, warmup, iters):
    """Device time per call with the launch sequence replayed from a CUDA graph (no host launch cost)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(warmup):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with _Sampling():
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def recall_at_k(got: torch.Tensor, exact: torch.Tensor) -> float:
    hits = (got.unsqueeze(2) == exact.unsqueeze(1)).any(dim=2).sum().item()
    return hits / exact.numel()


def cmd_ivf(a):
    import time
    dev = torch.device("cuda", 0)
    for name, val in (a.tunable or []):
        ts.set_tunable(name, int(val))
    pk = peaks()
    index = ts.TheoremIndex(a.dim, a.rows, dtype="bf16", device=dev)
    if a.data == "clustered":
        centers = synthetic.fill_index_clustered(index, a.rows, a.centers or a.nlist, a.sigma, seed=0)
        q_all = synthetic.make_clustered_queries(max(a.nq, a.nq_recall), centers, a.sigma)
    elif a.data == "hier":
        model = synthetic.HierarchicalCorpus(a.dim, n_leaves=10 * a.nlist, device=dev, seed=0)
        model.fill(index, a.rows)
        q_all = model.queries(max(a.nq, a.nq_recall), seed=3_000_000)
    else:
        synthetic.fill_index(index, 0, a.rows, seed=0)
        q_all = synthetic.make_queries(max(a.nq, a.nq_recall), a.dim, dev)
    torch.cuda.synchronize()
    t0 = time.time()
    index.ivf_train(a.nlist, n_sample=a.train_sample, iters=a.train_iters, seed=0)
    torch.cuda.synchronize()
    t1 = time.time()
    index.ivf_build(a.list_dtype)
    torch.cuda.synchronize()
    t2 = time.time()
    sizes = index.ivf_list_sizes().to(torch.float64)
    out = {"bench": "ivf", "rows": a.rows, "dim": a.dim, "data": a.data, "nlist": a.nlist, "nprobe": a.nprobe,
           "rescore_k": a.rescore, "k": a.k, "list_dtype": a.list_dtype, "train_sample": a.train_sample,
           "train_iters": a.train_iters, "train_s": t1 - t0, "build_s": t2 - t1,
           "list_rows_min_mean_max": [sizes.min().item(), sizes.mean().item(), sizes.max().item()],
           "empty_lists": int((sizes == 0).sum().item())}
    # recall@k against the exact path on the same index
    qr = q_all[:a.nq_recall]
    _, exact_ids = index.search(qr, a.k)
    _, ivf_ids = index.ivf_search(qr, a.k, nprobe=a.nprobe, rescore_k=a.rescore)
    out["recall_at_k"] = recall_at_k(ivf_ids, exact_ids)
    out["recall_queries"] = a.nq_recall
    for np_ in a.recall_sweep:
        _, ids2 = index.ivf_search(qr, a.k, nprobe=np_, rescore_k=a.rescore)
        out[f"recall_nprobe_{np_}"] = recall_at_k(ids2, exact_ids)
    # bytes the list scan must read per query: rows of the probed lists x list row bytes (+ 4 B scale)
    cent = index.ivf_centroids()
    qn = torch.nn.functional.normalize(qr, dim=1)
    probe = (qn @ cent.T).topk(min(a.nprobe, a.nlist), dim=1).indices
    rows_probed = sizes[probe].sum(dim=1)
    row_bytes = a.dim * (1 if a.list_dtype == "fp8" else 2) + (4 if a.list_dtype == "fp8" else 0)
    out["rows_probed_mean"] = rows_probed.mean().item()
    out["frac_of_corpus_probed"] = rows_probed.mean().item() / a.rows
    scan_bytes = rows_probed.mean().item() * row_bytes
    coarse_bytes = a.nlist * a.dim * 2
    out["algorithmic_bytes_per_query"] = {"list_scan": scan_bytes, "coarse": coarse_bytes,
                                          "rescore": a.rescore * a.dim * 2}
    # single-query latency: eager (host launch cost included) and graph-replayed (device only)
    q1 = [q_all[i:i + 1].contiguous() for i in range(64)]
    it = iter(itertools.cycle(range(64)))
    ms_eager = timed(lambda: index.ivf_search(q1[next(it)], a.k, nprobe=a.nprobe, rescore_k=a.rescore), 10, 200)
    qs = q_all[:1].clone()
    ms_graph = timed_graph(lambda: index.ivf_search(qs, a.k, nprobe=a.nprobe, rescore_k=a.rescore), 10, 200)
    out["single_query_ms_eager"] = ms_eager
    out["single_query_ms_graph"] = ms_graph
    out["single_query_qps_graph"] = 1e3 / ms_graph
    out["single_query_gbs_all_stages"] = (scan_bytes + coarse_bytes) / (ms_graph * 1e-3) / 1e9
    ms_exact = timed(lambda: index.search(q1[next(it)], a.k), 3, 20)
    out["exact_single_query_ms"] = ms_exact
    # batched throughput
    qb = q_all[:a.nq].contiguous()
    ms_b = timed(lambda: index.ivf_search(qb, a.k, nprobe=a.nprobe, rescore_k=a.rescore), 2, 5)
    out["batch_nq"] = a.nq
    out["batch_ms"] = ms_b
    out["batch_qps"] = a.nq / (ms_b * 1e-3)
    # what the list-major scan must READ: every (list, group of <= QB probing queries) streams the list once
    # (nq x per-query bytes would count a list once per query although it is read once per group)
    cent_ix = ts.build_index(index.ivf_centroids(), dtype="bf16", normalize=False, device=dev)
    _, probes_b = cent_ix.search(qb, min(a.nprobe, a.nlist))
    cnt = torch.bincount(probes_b.reshape(-1), minlength=a.nlist)
    cent_ix.close()
    lrow = ((a.dim + 15) // 16) * 16 + 4 if a.list_dtype == "fp8" else a.dim * 2
    sizes_i = index.ivf_list_sizes()
    out["batch_bytes_if_read_per_query"] = int((cnt * sizes_i).sum().item()) * lrow
    out["batch_by_scoring_backend"] = {}
    for mode in (a.mma_modes or [ts.get_tunable("ivf.group_mma")]):
        qbw = {0: 4, 1: 8, 2: 16, 3: 16, 4: 8}[mode]
        ts.set_tunable("ivf.group_mma", mode)
        index._ws = {}
        ms_m = timed(lambda: index.ivf_search(qb, a.k, nprobe=a.nprobe, rescore_k=a.rescore), 2, 8)
        unique = int((((cnt + qbw - 1) // qbw) * sizes_i).sum().item()) * lrow
        _, ids_m = index.ivf_search(qr, a.k, nprobe=a.nprobe, rescore_k=a.rescore)
        out["batch_by_scoring_backend"][str(mode)] = {
            "queries_per_group": qbw, "batch_ms": ms_m, "qps": a.nq / (ms_m * 1e-3), "unique_list_bytes": unique,
            "whole_batch_frac_of_measured_hbm": unique / (ms_m * 1e-3) / 1e9 / pk["hbm_gbs"],
            "recall_at_k": recall_at_k(ids_m, exact_ids), "clocks": _CLOCKS.summary() if _CLOCKS else None}
    out["tunables"] = a.tunable
    print(json.dumps(out), flush=True)


def cmd_ivf_q1(a):
    """Minimal single-query IVF run for the ncu launch list: build, then `iters` single queries."""
    dev = torch.device("cuda", 0)
    index = ts.TheoremIndex(a.dim, a.rows, dtype="bf16", device=dev)
    centers = synthetic.fill_index_clustered(index, a.rows, a.centers or a.nlist, a.sigma, seed=0)
    q_all = synthetic.make_clustered_queries(64, centers, a.sigma)
    index.ivf_train(a.nlist, n_sample=a.train_sample, iters=2, seed=0)
    index.ivf_build(a.list_dtype)
    torch.cuda.synchronize()
    if a.profile_nq > 1:
        q_all = synthetic.make_clustered_queries(a.profile_nq, centers, a.sigma)
    for i in range(a.iters):
        qq = q_all[i:i + 1] if a.profile_nq == 1 else q_all
        index.ivf_search(qq, a.k, nprobe=a.nprobe, rescore_k=a.rescore)
    torch.cuda.synchronize()
    print(json.dumps({"bench": "ivf-q1", "profile_nq": a.profile_nq, "launches": ts.kernel_launches()}))


def cmd_ivf_q1_sweep(a):
    """Single-query IVF latency (graph-replayed) over the K4b tunables, plus the phase timeline of one query."""
    import ctypes as C
    import numpy as np
    from theoremsearch_b200 import _lib
    dev = torch.device("cuda", 0)
    index = ts.TheoremIndex(a.dim, a.rows, dtype="bf16", device=dev)
    centers = synthetic.fill_index_clustered(index, a.rows, a.centers or a.nlist, a.sigma, seed=0)
    q_all = synthetic.make_clustered_queries(64, centers, a.sigma)
    index.ivf_train(a.nlist, n_sample=a.train_sample, iters=3, seed=0)
    index.ivf_build(a.list_dtype)
    torch.cuda.synchronize()
    qs = q_all[:1].clone()
    fn = lambda: index.ivf_search(qs, a.k, nprobe=a.nprobe, rescore_k=a.rescore)
    for warps, rows, parts in [(0, 0, 0), (8, 8, 0), (16, 4, 0), (8, 4, 0), (12, 4, 0), (16, 4, 74), (16, 4, 296), (8, 8, 74)]:
        ts.set_tunable("ivf.warps", warps)
        ts.set_tunable("ivf.tile_rows", rows)
        ts.set_tunable("ivf.parts", parts)
        index._ws = {}
        ms = timed_graph(fn, 10, 200)
        print(json.dumps({"bench": "ivf-q1-sweep", "rows": a.rows, "nlist": a.nlist, "warps": warps, "tile_rows": rows,
                          "parts": parts, "q1_ms_graph": ms}))
    for name in ("ivf.warps", "ivf.tile_rows", "ivf.parts"):
        ts.set_tunable(name, 0)
    index._ws = {}
    ts.set_tunable("ivf.timeline", 1)
    for _ in range(3):
        fn()
    buf = np.zeros((148, 12), dtype=np.uint64)
    _lib.check(_lib.lib.ts_debug_ivf_timeline(buf.ctypes.data, 148))
    ts.set_tunable("ivf.timeline", 0)
    t = buf.astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    last = int(np.argmax(t[:, 7]))
    t[t == 0] = t0
    print(json.dumps({"bench": "ivf-q1-timeline", "unit": "us since first CTA start",
                      "phases": ["start", "table", "prologue", "scan_done", "gathered", "cta_sorted", "final_begin", "final_end",
                                 "heads_sorted", "survivors_compacted", "survivors_sorted"],
                      "median_cta": [float(np.median(rel[:, i])) for i in range(6)],
                      "max_cta": [float(rel[:, i].max()) for i in range(6)],
                      "last_cta": [float(x) for x in rel[last][:11]]}))


def cmd_k1(a):
    """K1 (normalise + quantise) throughput: fp32 rows in HBM -> bf16 index rows. Bytes = 4*D read + 2*D written."""
    dev = torch.device("cuda", 0)
    pk = peaks()
    blk = 1 << 20
    x = torch.randn((blk, a.dim), dtype=torch.float32, device=dev)
    reps = 8
    index = ts.TheoremIndex(a.dim, blk * (reps + 2), dtype="bf16", device=dev)
    from theoremsearch_b200._lib import lib, check
    from theoremsearch_b200.index import _stream_ptr
    def add():
        check(lib.ts_index_add(index.handle, x.data_ptr(), 0, blk, 1, None, _stream_ptr(dev)))   # reads `index` at call time
    add(); add()
    torch.cuda.synchronize()
    # ~100 launches (each writes the next 1 Mi rows; the index is emptied between rounds of 8 by re-creating it)
    total_ms, launches = 0.0, 0
    with _Sampling():
        for _round in range(12):
            index.close()
            index = ts.TheoremIndex(a.dim, blk * (reps + 2), dtype="bf16", device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                add()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            launches += reps
    ms = total_ms / launches
    gbs = blk * a.dim * 6 / (ms * 1e-3) / 1e9
    print(json.dumps({"bench": "k1", "rows_per_launch": blk, "dim": a.dim, "ms_per_launch": ms, "gbs": gbs,
                      "frac_of_measured_hbm": gbs / pk["hbm_gbs"], "rows_per_s": blk / (ms * 1e-3)}))


def cmd_fp8_scan(a):
    """BASELINE north_star: single-query scan over an fp8-e4m3 corpus, re-scored in fp32 — half the bytes."""
    dev = torch.device("cuda", 0)
    pk = peaks()
    index = build(a.rows, a.dim, dev)
    index.build_fp8_shadow()
    q_all = synthetic.make_queries(max(256, a.nq_recall), a.dim, dev)
    qr = q_all[:a.nq_recall]
    _, exact_ids = index.search(qr, a.k)
    out = {"bench": "fp8-scan", "rows": a.rows, "dim": a.dim, "k": a.k, "data": "gaussian (hardest case: no structure)"}
    for rk in (32, 64, 128, 256):
        ids = torch.cat([index.search_fp8(qr[i:i + 64], a.k, rescore_k=rk)[1] for i in range(0, a.nq_recall, 64)])
        out[f"recall_at_{a.k}_rescore_{rk}"] = recall_at_k(ids, exact_ids)
    qs = q_all[:1].clone()
    ms = timed_graph(lambda: index.search_fp8(qs, a.k, rescore_k=a.rescore), 10, 100)
    ms_exact = timed_graph(lambda: index.search(qs, a.k), 10, 100)
    nbytes = a.rows * (a.dim + 4)
    out.update({"rescore_k": a.rescore, "fp8_q1_ms": ms, "fp8_q1_qps": 1e3 / ms, "fp8_scan_gbs": nbytes / (ms * 1e-3) / 1e9,
                "fp8_frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "bf16_exact_q1_ms": ms_exact,
                "speedup_vs_bf16_exact": ms_exact / ms})
    print(json.dumps(out), flush=True)


def cmd_sharded(a):
    """configs[3] + configs[4] at their named size under torchrun: rows row-sharded over WORLD_SIZE GPUs.
    Exact Q=1 / Q=batch, then IVF-Flat (one shared coarse quantiser) with recall vs the sharded exact path."""
    import time
    import torch.distributed as dist
    from theoremsearch_b200.sharded import ShardedIndex, shard_bounds
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "WARN").upper() in ("WARN", "VERSION"):   # (the launcher's default) NCCL prints its version banner on STDOUT at these levels
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(a.rows, world)[rank]
    index = ts.TheoremIndex(a.dim, hi - lo, dtype="bf16", device=dev)
    t0 = time.time()
    if a.data == "clustered":
        centers = synthetic.fill_index_clustered(index, hi - lo, a.centers or a.nlist, a.sigma, seed=0, first_row=lo)
        q_all = synthetic.make_clustered_queries(max(a.nq, a.nq_recall), centers, a.sigma)
    elif a.data == "hier":
        model = synthetic.HierarchicalCorpus(a.dim, n_leaves=10 * a.nlist, device=dev, seed=0)
        model.fill(index, hi - lo, first_row=lo)
        q_all = model.queries(max(a.nq, a.nq_recall), seed=3_000_000)
    else:
        synthetic.fill_index(index, lo, hi - lo, seed=0)
        q_all = synthetic.make_queries(max(a.nq, a.nq_recall), a.dim, dev)
    torch.cuda.synchronize()
    fill_s = time.time() - t0
    sh = ShardedIndex(index, a.rows)
    peer = False
    try:
        sh.enable_peer_exchange(max_nq=1, max_k=32)
        peer = True
    except ts.TheoremSearchError:
        pass

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_max(fn, warmup, iters):
        for _ in range(warmup):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with _Sampling():
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            sync()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pk = peaks()
    out = {"bench": "sharded", "rows": a.rows, "dim": a.dim, "n_gpus": world, "rows_per_gpu": hi - lo, "data": a.data,
           "fill_s": fill_s}
    q1 = [q_all[i:i + 1].contiguous() for i in range(64)]
    it = iter(itertools.cycle(range(64)))
    ms1 = timed_max(lambda: sh.search(q1[next(it)], a.k, independent=peer), 10, 200)
    out["exact_q1_form"] = ("device-initiated exchange, scan + exchange kernel (PDL), independent stream" if peer
                            else "NCCL all-gather + merge kernel")
    out["exact_q1_clocks"] = _CLOCKS.summary() if _CLOCKS else None
    if peer:
        out["exact_q1_dependent_ms"] = timed_max(lambda: sh.search(q1[next(it)], a.k), 10, 100)
        out["exact_q1_one_kernel_ms"] = timed_max(lambda: sh.search(q1[next(it)], a.k, one_kernel=True), 10, 100)
        saved, sh._xchg = sh._xchg, None
        out["exact_q1_nccl_ms"] = timed_max(lambda: sh.search(q1[next(it)], a.k), 10, 100)
        sh._xchg = saved
    out["exact_q1_ms"] = ms1
    out["exact_q1_qps"] = 1e3 / ms1
    out["exact_q1_aggregate_gbs"] = a.rows * a.dim * 2 / (ms1 * 1e-3) / 1e9
    out["exact_q1_frac_of_measured_hbm_per_gpu"] = out["exact_q1_aggregate_gbs"] / world / pk["hbm_gbs"]
    ms1_local = timed_max(lambda: index.search_keys(q1[next(it)], a.k), 5, 50)
    out["exact_q1_local_scan_ms"] = ms1_local
    qb = q_all[:a.nq].contiguous()
    msb = timed_max(lambda: sh.search(qb, a.kb), 1, 2)
    out["exact_batch"] = {"nq": a.nq, "k": a.kb, "ms": msb, "qps": a.nq / (msb * 1e-3),
                          "aggregate_tflops": 2.0 * a.nq * a.rows * a.dim / (msb * 1e-3) / 1e12}
    if a.nlist > 0:
        sync()
        t0 = time.time()
        sh.ivf_train_build(a.nlist, n_sample=a.train_sample, iters=a.train_iters, seed=0, list_dtype=a.list_dtype)
        sync()
        out["ivf_train_build_s"] = time.time() - t0
        qr = q_all[:a.nq_recall].contiguous()
        _, exact_ids = sh.search(qr, a.k)
        _, ivf_ids = sh.ivf_search(qr, a.k, nprobe=a.nprobe, rescore_k=a.rescore)
        ivf = {"nlist": a.nlist, "nprobe": a.nprobe, "rescore_k": a.rescore, "list_dtype": a.list_dtype,
               "recall_at_k": recall_at_k(ivf_ids, exact_ids), "recall_queries": a.nq_recall}
        for np_ in a.recall_sweep:
            _, ids2 = sh.ivf_search(qr, a.k, nprobe=np_, rescore_k=a.rescore)
            ivf[f"recall_nprobe_{np_}"] = recall_at_k(ids2, exact_ids)
        sizes = index.ivf_list_sizes().to(torch.float64)
        ivf["local_list_rows_min_mean_max"] = [sizes.min().item(), sizes.mean().item(), sizes.max().item()]
        ms = timed_max(lambda: sh.ivf_search(q1[next(it)], a.k, nprobe=a.nprobe, rescore_k=a.rescore), 10, 100)
        ivf["q1_ms"] = ms
        ivf["q1_qps"] = 1e3 / ms
        msb = timed_max(lambda: sh.ivf_search(qb, a.k, nprobe=a.nprobe, rescore_k=a.rescore), 1, 3)
        ivf["batch"] = {"nq": a.nq, "ms": msb, "qps": a.nq / (msb * 1e-3)}
        out["ivf"] = ivf
    if rank == 0:
        print(json.dumps(out), flush=True)
    sh.close()
    index.close()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)


def cmd_shard_stream(a):
    """One GPU holding ONE shard of the N-way sharded corpus (default 1.25M rows = 10M / 8), world = 1 exchange:
    the per-step cost of every search form at the shard size, without a second GPU. Isolates what the
    programmatic-dependent-launch chain buys (launch gap, merge tail) from cross-GPU skew."""
    from theoremsearch_b200.sharded import ShardedIndex
    dev = torch.device("cuda", 0)
    index = build(a.rows, a.dim, dev)
    sh = ShardedIndex(index, a.rows).enable_peer_exchange(max_nq=1, max_k=32)
    q = synthetic.make_queries(256, a.dim, dev)
    torch.cuda.synchronize()
    it = itertools.count()
    forms = {
        "plain ts_search (one kernel, last-CTA merge)": lambda: index.search(q[next(it) % 256:][:1], a.k),
        "sharded one-kernel": lambda: sh.search(q[next(it) % 256:][:1], a.k, one_kernel=True),
        "sharded scan+exchange (PDL), dependent": lambda: sh.search(q[next(it) % 256:][:1], a.k),
        "sharded scan+exchange (PDL), independent stream": lambda: sh.search(q[next(it) % 256:][:1], a.k, independent=True),
    }
    bytes_ = a.rows * a.dim * 2
    out = {"bench": "shard-stream", "rows": a.rows, "dim": a.dim, "k": a.k, "iters": a.iters}
    for name, fn in forms.items():
        ms = timed(fn, 20, a.iters)
        out[name] = {"ms_per_query": ms, "gbs": bytes_ / (ms * 1e-3) / 1e9}
    for i in range(3):
        index.search_host(q[i].cpu().numpy(), a.k, timing=True)
    km = []
    for i in range(50):
        index.search_host(q[i].cpu().numpy(), a.k, timing=True)
        km.append(index.last_kernel_ms)
    out["scan_kernel_alone_ms (events around the kernel)"] = sum(km) / len(km)
    print(json.dumps(out), flush=True)
    sh.close()


def cmd_scan_timeline(a):
    """Per-CTA phase stamps of consecutive K2 launches at the shard size (tunable scan.timeline): where a
    query's time goes besides streaming — start-up, the spread of the CTAs' finishing times, the merge tail —
    and how far the next launch's CTAs overlap the previous launch's tail, for every search form."""
    import ctypes as C
    import numpy as np
    from theoremsearch_b200 import _lib
    from theoremsearch_b200.sharded import ShardedIndex
    dev = torch.device("cuda", 0)
    index = build(a.rows, a.dim, dev)
    sh = ShardedIndex(index, a.rows).enable_peer_exchange(max_nq=1, max_k=32)
    q = synthetic.make_queries(64, a.dim, dev)
    torch.cuda.synchronize()
    forms = {"plain": lambda i: index.search(q[i:i + 1], a.k),
             "one_kernel": lambda i: sh.search(q[i:i + 1], a.k, one_kernel=True),
             "pdl_dependent": lambda i: sh.search(q[i:i + 1], a.k),
             "pdl_independent": lambda i: sh.search(q[i:i + 1], a.k, independent=True)}
    n = 148
    out = {"bench": "scan-timeline", "rows": a.rows, "stamps": ["entry", "query_ready", "first_tile", "warp0_done",
                                                                "cta_done", "list_written", "final_merge_done"]}
    for name, fn in forms.items():
        for i in range(20):
            fn(i)
        torch.cuda.synchronize()
        _lib.set_tunable("scan.timeline", 1)
        for i in range(20, 36):
            fn(i)
        torch.cuda.synchronize()
        _lib.set_tunable("scan.timeline", 0)
        tl = []
        for back in range(4):
            buf = np.zeros((n, 8), dtype=np.uint64)
            _lib.check(_lib.lib.ts_debug_scan_timeline(buf.ctypes.data, back, n))
            tl.append(buf.astype(np.int64))
        tl = tl[::-1]                       # oldest first: launches L-3 .. L
        t0 = int(tl[1][:, 0].min())         # reference: first CTA entry of launch L-2
        rep = {}
        for li, lab in ((1, "launch_n"), (2, "launch_n_plus_1")):
            d = tl[li][:, :7] - t0
            rep[lab] = {st: {"min": int(d[:, j].min()), "p50": int(np.median(d[:, j])), "p95": int(np.percentile(d[:, j], 95)),
                             "max": int(d[:, j].max())} for j, st in enumerate(out["stamps"][:6])}
            last = tl[li][:, 6]
            rep[lab]["final_merge_done"] = int(last[last > 0].max() - t0) if (last > t0).any() else None
        rep["period_ns (entry min of n+1 minus entry min of n)"] = int(tl[2][:, 0].min() - tl[1][:, 0].min())
        rep["cta_busy_ns p50 (entry -> list_written)"] = int(np.median(tl[1][:, 5] - tl[1][:, 0]))
        rep["cta_stream_ns p50 (first_tile -> cta_done)"] = int(np.median(tl[1][:, 4] - tl[1][:, 2]))
        rep["spread_cta_done_ns (max - min)"] = int(tl[1][:, 4].max() - tl[1][:, 4].min())
        rep["spread_entry_ns (max - min)"] = int(tl[1][:, 0].max() - tl[1][:, 0].min())
        # per-SM gap between the end of launch n's CTA and the entry of launch n+1's CTA on the same SM
        sm_n = {int(r[7]): int(r[5]) for r in tl[1]}
        gaps = [int(r[0]) - sm_n[int(r[7])] for r in tl[2] if int(r[7]) in sm_n]
        rep["same_sm_gap_ns (n+1 entry - n list_written)"] = {"p50": int(np.median(gaps)), "min": int(min(gaps)), "max": int(max(gaps))}
        out[name] = rep
    print(json.dumps(out), flush=True)
    sh.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["batched", "sweep-scan", "small-batch", "ivf", "sharded", "ivf-q1", "ivf-q1-sweep", "fp8-scan", "k1", "cfg0", "torch-compare", "shard-stream", "scan-timeline"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=4096)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tunable", nargs=2, action="append", metavar=("NAME", "VALUE"))
    ap.add_argument("--fine", action="store_true")
    ap.add_argument("--kb", type=int, default=100, help="k of the batched exact search in `sharded`")
    ap.add_argument("--nlist", type=int, default=16384)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--rescore", type=int, default=100)
    ap.add_argument("--list-dtype", default="fp8", choices=["fp8", "bf16"])
    ap.add_argument("--data", default="clustered", choices=["clustered", "gaussian", "hier"])
    ap.add_argument("--mma-modes", type=int, nargs="*", default=[],
                    help="ivf: time the batch with each K4d scoring back end (ivf.group_mma value)")
    ap.add_argument("--centers", type=int, default=0)
    ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--train-sample", type=int, default=2_000_000)
    ap.add_argument("--train-iters", type=int, default=10)
    ap.add_argument("--nq-recall", type=int, default=1000)
    ap.add_argument("--recall-sweep", type=int, nargs="*", default=[])
    ap.add_argument("--profile-nq", type=int, default=1, help="ivf-q1: queries per call (1 = latency mode)")
    a = ap.parse_args()
    global _CLOCKS
    from bench import ClockSampler
    _CLOCKS = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    json.dumps = _dumps_with_clocks
    if a.k is None:
        a.k = 100 if a.cmd == "batched" else 10
    {"batched": cmd_batched, "sweep-scan": cmd_sweep_scan, "small-batch": cmd_small_batch, "ivf": cmd_ivf, "sharded": cmd_sharded, "ivf-q1": cmd_ivf_q1, "ivf-q1-sweep": cmd_ivf_q1_sweep, "fp8-scan": cmd_fp8_scan, "k1": cmd_k1, "cfg0": cmd_cfg0, "torch-compare": cmd_torch_compare, "shard-stream": cmd_shard_stream, "scan-timeline": cmd_scan_timeline}[a.cmd](a)


if __name__ == "__main__":
    main()
