#!/usr/bin/env python
"""Secondary measurements (not the driver's contract — that is bench.py): the other BASELINE.json
configs and tuning sweeps.  Each sub-command prints one JSON line per measurement.

  python bench_extra.py batched   [--rows 10000000 --nq 4096 --k 100]     configs[2]: K3 tcgen05 GEMM + top-k
  python bench_extra.py sweep-scan [--rows 10000000]                      K2 tunables sweep
  python bench_extra.py small-batch [--rows 10000000]                     nq = 1..256 latency curve
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


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["batched", "sweep-scan", "small-batch"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=4096)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tunable", nargs=2, action="append", metavar=("NAME", "VALUE"))
    ap.add_argument("--fine", action="store_true")
    a = ap.parse_args()
    if a.k is None:
        a.k = 100 if a.cmd == "batched" else 10
    {"batched": cmd_batched, "sweep-scan": cmd_sweep_scan, "small-batch": cmd_small_batch}[a.cmd](a)


if __name__ == "__main__":
    main()
