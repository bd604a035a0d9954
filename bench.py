#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

Metric: queries/sec (exact top-10) of the single-query scan over a 10M x 1024 bf16 corpus
(configs[1]); a "step" is ONE query = one full pass over the corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* N = 1: the whole corpus on one B200.  N > 1 (under torchrun): the SAME corpus row-sharded
  over N GPUs (strong scaling); per query every rank scans its shard and the GPUs exchange their k
  packed keys themselves over NVLink (stores into CUDA-IPC mapped peer memory + sequence flags,
  `ts_search_sharded`): a scan kernel and an exchange kernel chained by programmatic dependent launch,
  so the next query's scan streams the corpus while this query's keys are exchanged and merged.
  `--exchange one-kernel` puts the exchange into the scan kernel's last CTA, `--exchange nccl` selects
  the all-gather + K5 merge form.
* `value`   : device-timed (CUDA events, max over ranks), queries already resident in HBM.
* `e2e`     : wall-clock through the public host-buffer API (`TheoremIndex.search_host` /
              `ShardedIndex.search_host` -> `ts_search_host` / `ts_search_sharded_host`): per step a
              4 KB pinned H2D query copy and a k*(4+8) B D2H result, synchronised.
* `roofline`: N*D*2 algorithmic bytes / average duration of the K2 scan kernel, measured live with
              CUDA events recorded around that kernel on its launch stream (ts_ctx timing).
* `parity_check`: an untimed block — planted rows come back first, the host-buffer path equals the device
              path, and at N > 1 the three exchange forms agree bit for bit on every rank.
* `batched` / `ivf` (N = 1): BASELINE.json configs[2] (4096 queries x top-100 over the same corpus, K3) and
              configs[4] (IVF-Flat fp8, nlist 16384, recall@10 vs exact and q/s per nprobe) with clocks;
              `batched.evaluation`: the six compare_embeddings metrics from that batch's ids on the GPU (K6).
* `cpu_baseline` / `--impl reference`: the reference's CPU path (util.cos_sim + argsort,
  test_app.py:76-77) restated in oracle/oracle.py, on the host's cores. The reference arm runs it on the
  FULL 10M x 1024 fp32 corpus (41 GB; `--reference-sample-rows R` selects a row sample instead and says so
  in config.workload); the in-line `cpu_baseline` of our arm is a bounded 1M-row sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "queries/sec (single-query exact top-10, 10M x 1024 bf16 corpus)"
UNIT = "queries/s"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while a timed region runs. Construct it BEFORE the
    barrier that precedes the region (nvmlInit takes milliseconds and differs per rank); `with` only starts
    and stops the sampling thread."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() \
                else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
        except Exception as e:  # NVML missing: report that, never fake numbers
            self.nv = None
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.01)      # NVML queries take driver locks: sampling harder than this perturbs latency-bound loops

    def __enter__(self):
        self.samples, self.reasons = [], set()
        self._stop.clear()
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()
            self._t = None

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU arm
def host_rows(rows: int, dim: int, threads: int) -> np.ndarray:
    """`rows` x `dim` i.i.d. N(0,1) fp32 rows in host memory, filled by `threads` numpy generators in
    parallel (the CPU arm times arithmetic on them; their values need not equal the device corpus)."""
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty((rows, dim), dtype=np.float32)
    blk = 1 << 17

    def fill(c):
        lo = c * blk
        np.random.Generator(np.random.SFC64(c)).standard_normal(out=out[lo:min(rows, lo + blk)], dtype=np.float32)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(fill, range((rows + blk - 1) // blk)))
    return out


def mem_available_bytes() -> int:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def cpu_reference_arm(corpus_fp32: np.ndarray, queries: np.ndarray, k: int, steps: int, warmup: int):
    """The reference's own CPU path, restated (oracle.reference_single_query_verbatim == test_app.py:75-77:
    cos_sim re-normalises the corpus on every call, then a full argsort).  One step = one query over
    `corpus_fp32`; exactly `warmup` untimed and `steps` timed queries."""
    from oracle import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    db = torch.from_numpy(corpus_fp32)
    times = []
    for i in range(warmup + steps):
        q = torch.from_numpy(queries[i % len(queries)])
        t0 = time.perf_counter()
        oracle.reference_single_query_verbatim(q, db, k)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return {"s_per_step": sum(times) / len(times), "p50_s": statistics.median(times)}


def cpu_prenormalised_arm(corpus_fp32: np.ndarray, queries: np.ndarray, k: int, reps: int):
    """pgvector-shaped variant: corpus normalised once at write time, dot + top-k per query.
    Normalises `corpus_fp32` IN PLACE (call it after cpu_reference_arm)."""
    dbn = torch.from_numpy(corpus_fp32)
    for lo in range(0, dbn.shape[0], 1 << 20):          # "write time": normalise in place, once, untimed
        blk = dbn[lo:lo + (1 << 20)]
        blk.div_(blk.norm(dim=1, keepdim=True).clamp_min_(1e-12))
    t0 = time.perf_counter()
    for i in range(reps):
        s = torch.mv(dbn, torch.from_numpy(queries[i % len(queries)]))
        torch.topk(s, k)
    return (time.perf_counter() - t0) / reps


def run_reference(args, config) -> int:
    cores = os.cpu_count() or 1
    from oracle import oracle
    rows = args.rows
    sampled = args.reference_sample_rows > 0
    need = int(args.rows * args.dim * 4 * 2.3)      # corpus + the normalised copy cos_sim makes per query
    avail = mem_available_bytes()
    note = None
    if sampled:
        rows = min(args.reference_sample_rows, args.rows)
    elif avail and avail < need:
        rows = max(100_000, int(avail * 0.35 / (args.dim * 4)))
        sampled = True
        note = f"host has {avail / 1e9:.0f} GB available, full corpus needs {need / 1e9:.0f} GB"
    corpus = host_rows(rows, args.dim, cores)
    qs = oracle.normalize_f64(oracle.synthetic_queries(32, args.dim))
    r = cpu_reference_arm(corpus, qs, args.k, args.steps, args.warmup)
    pg = cpu_prenormalised_arm(corpus, qs, args.k, max(3, min(args.steps, 10)))
    scale = args.rows / rows
    value = 1.0 / (r["s_per_step"] * scale)
    if sampled:
        config = dict(config)
        config["workload"] += (f" — REFERENCE ARM ON A ROW SAMPLE: {rows} of {args.rows} rows, q/s scaled by rows "
                               f"({scale:.2f}x; the path is linear in N)" + (f" [{note}]" if note else ""))
    sample = (f"{args.steps} queries (+{args.warmup} warm-up) x {rows} of {args.rows} fp32 rows in host memory: "
              f"util.cos_sim (re-normalises the corpus per query) + full argsort as test_app.py:76-77, torch CPU, "
              f"{cores} threads" + ("; q/s scaled by rows" if sampled else "; nothing extrapolated"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["s_per_step"] * (scale if sampled else 1.0),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config, "extrapolated": bool(sampled), "rows_timed": rows,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "p50_ms": r["p50_s"] * 1e3 * (scale if sampled else 1.0),
                         "prenormalised_dot_topk_value": 1.0 / (pg * scale)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ extras (N = 1)
def batched_section(ts, index, args, dev, clocks, peaks):
    """BASELINE.json configs[2]: 4096 queries x top-100 over the resident corpus through K3 (tcgen05 GEMM +
    fused top-k), timed with CUDA events; 8 sampled queries must equal the single-query scan bit for bit."""
    from theoremsearch_b200 import synthetic
    nq, k = args.batch_nq, args.batch_k
    q = synthetic.make_queries(nq, args.dim, dev, seed=2_000_000)
    index.search(q, k)                                             # warm-up (workspace, thresholds path)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        e0.record()
        for _ in range(reps):
            s, ids = index.search(q, k)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fixups = ts.last_batched_fixups()
    pick = [0, 1, nq // 3, nq // 2, nq - 2, nq - 1, 1234 % nq, 4000 % nq]
    same = True
    for j in pick:
        s1, i1 = index.search(q[j:j + 1], k)
        same &= bool(torch.equal(s1[0], s[j]) and torch.equal(i1[0], ids[j]))
    flops = 2.0 * nq * len(index) * args.dim
    tf = flops / (ms * 1e-3) / 1e12
    # the consumer of the reference's batched path (compare_embeddings.py:69-92): six metrics from the ids where K3
    # left them (K6). Query j's correct document is planted at rank j % 10 + 1, so Hit@5 / MRR@5 are known exactly.
    from theoremsearch_b200 import metrics
    host = ids[:, :10].cpu().numpy()
    qrels = {j: {int(host[j, j % 10]): 1.0} for j in range(nq)}
    table = metrics.JudgedTable(qrels, nq, dev, max_k=10)
    cuts = {"precision": 1, "hit": 5, "mrr": 5, "ndcg": 5, "err": 5, "q_measure": 5}
    got = table.evaluate(ids, cuts)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        table.evaluate(ids, cuts)
    e1.record()
    torch.cuda.synchronize()
    eval_ms = e0.elapsed_time(e1) / 10
    table.close()
    ranks = np.arange(nq) % 10 + 1
    want_hit, want_mrr = float(np.mean(ranks <= 5)), float(np.mean(np.where(ranks <= 5, 1.0 / ranks, 0.0)))
    evaluation = {"ms_incl_d2h_of_6_doubles": eval_ms, "metrics": got,
                  "planted_ranks_recovered": bool(abs(got["hit"] - want_hit) < 1e-12 and abs(got["mrr"] - want_mrr) < 1e-12)}
    return {"workload": f"configs[2]: {nq} queries x top-{k} over {len(index)}x{args.dim} bf16 (K3)",
            "evaluation": evaluation,
            "ms": ms, "queries_per_s": nq / (ms * 1e-3), "tflops": tf,
            "frac_burst": tf / peaks["bf16_tflops"],
            "frac_sustained": tf / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
            "uncertified_queries_rescanned": fixups, "sampled_queries_bitwise_equal_single_query_scan": same,
            "clocks": clocks.summary()}


def ivf_section(ts, args, dev, clocks, peaks):
    """BASELINE.json configs[4] at the largest size one GPU holds with its exact re-score rows: IVF-Flat,
    e4m3 lists, nlist 16384, k' = 100, on a hierarchically clustered corpus (10x more leaf clusters than lists,
    anisotropic low-rank spread + isotropic noise); recall@10 vs the exact search per nprobe, q/s at nprobe 32."""
    from theoremsearch_b200 import synthetic
    free, _ = torch.cuda.mem_get_info(dev)
    per_row = args.dim * 2 + args.dim + 64                       # bf16 rows + e4m3 lists + scales/rows/ids
    rows = int(min(args.ivf_rows, (free - (12 << 30)) // per_row))
    index = ts.TheoremIndex(args.dim, rows, dtype="bf16", device=dev)
    model = synthetic.HierarchicalCorpus(args.dim, n_leaves=10 * args.ivf_nlist, device=dev, seed=0)
    model.fill(index, rows)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.ivf_train(args.ivf_nlist, n_sample=min(rows, 2_000_000), iters=10, seed=0)
    index.ivf_build("fp8")
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    nq_recall = 1024
    q = model.queries(4096, seed=3_000_000)
    _, exact_ids = index.search(q[:nq_recall], 10)
    exact = exact_ids.cpu().numpy()
    curve = {}
    for nprobe in (1, 8, 32, 128):
        _, got = index.ivf_search(q[:nq_recall], 10, nprobe=nprobe, rescore_k=100)
        g = got.cpu().numpy()
        hits = sum(len(set(g[i]) & set(exact[i])) for i in range(nq_recall))
        curve[str(nprobe)] = hits / (10.0 * nq_recall)
    # throughput at nprobe 32, 4096 queries per batch
    index.ivf_search(q, 10, nprobe=32, rescore_k=100)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        e0.record()
        for _ in range(reps):
            index.ivf_search(q, 10, nprobe=32, rescore_k=100)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # unique bytes the list-major scan must read: every (list, group of <= QB probing queries) streams the list once
    cent = ts.build_index(index.ivf_centroids(), dtype="bf16", normalize=False, device=dev)
    _, probes = cent.search(q, 32)
    cnt = torch.bincount(probes.reshape(-1), minlength=args.ivf_nlist)
    sizes = index.ivf_list_sizes()
    row_bytes = ((args.dim + 15) // 16) * 16 + 4
    qbw = {0: 4, 1: 8, 2: 16, 3: 16, 4: 8}.get(ts.get_tunable("ivf.group_mma"), 8)   # queries per (list, group) work item
    unique = int((((cnt + qbw - 1) // qbw) * sizes).sum().item()) * row_bytes
    per_query_equiv = int((cnt * sizes).sum().item()) * row_bytes
    cent.close()
    # single-query latency (device time, mean over 200 distinct queries)
    for i in range(10):
        index.ivf_search(q[i:i + 1], 10, nprobe=32, rescore_k=100)
    torch.cuda.synchronize()
    e0.record()
    for i in range(200):
        index.ivf_search(q[i:i + 1], 10, nprobe=32, rescore_k=100)
    e1.record()
    torch.cuda.synchronize()
    q1_ms = e0.elapsed_time(e1) / 200
    out = {"workload": f"configs[4]: IVF-Flat e4m3 lists + exact re-score, nlist {args.ivf_nlist}, nprobe 32, k' 100, "
                       f"{rows}x{args.dim} hierarchical-cluster corpus ({10 * args.ivf_nlist} leaf clusters) on 1 GPU",
           "rows": rows, "train_build_s": build_s, "recall_at_10": curve["32"], "recall_at_10_by_nprobe": curve,
           "recall_queries": nq_recall, "batch_queries": 4096, "batch_ms": ms, "qps": 4096 / (ms * 1e-3),
           "unique_bytes": unique, "per_query_scan_bytes_if_not_grouped": per_query_equiv,
           "frac_hbm": unique / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
           "queries_per_group": qbw,
           "frac_hbm_note": "unique (list, query-group) bytes / WHOLE batch time (coarse + scan + select + re-score) / measured HBM peak",
           "single_query_ms": q1_ms, "clocks": clocks.summary()}
    index.close()
    return out


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--reference-sample-rows", type=int, default=0,
                    help="reference arm: time a row sample of this size instead of the full corpus (0 = full corpus)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-chain", action="store_true",
                    help="N = 1: plain one-kernel searches instead of the scan + finishing kernel chained by programmatic "
                         "dependent launch (queries are resident and uploaded beforehand either way)")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2] / configs[4] sections (N = 1)")
    ap.add_argument("--batch-nq", type=int, default=4096)
    ap.add_argument("--batch-k", type=int, default=100)
    ap.add_argument("--ivf-rows", type=int, default=40_000_000)
    ap.add_argument("--ivf-nlist", type=int, default=16384)
    ap.add_argument("--exchange", default="pdl", choices=["pdl", "fused", "one-kernel", "nccl"],
                    help="N>1: scan + exchange kernel chained by programmatic dependent launch (default; 'fused' is "
                         "an alias), exchange inside the scan kernel's last CTA, or NCCL all-gather + merge kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.exchange == "fused":
        args.exchange = "pdl"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"configs[1]: single-query exact top-{args.k} over {args.rows}x{args.dim} bf16",
              "device_stream": ("one kernel per query (ts_search)" if args.no_stream_chain else
                                "per query a scan kernel + a finishing kernel chained by programmatic dependent launch; queries "
                                "uploaded beforehand (independent=True), so consecutive scans overlap their tails"),
              "rows": args.rows, "dim": args.dim, "k": args.k,
              "sharding": f"row-sharded x{world}" if world > 1 else "single GPU",
              "arithmetic": "bf16 corpus rows, fp32 query, fp32 products and accumulation",
              "l2": "corpus shard >> 126 MB L2, distinct query per step (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, config)

    # ------------------------------------------------------------------------- our arm
    import torch.distributed as dist

    import theoremsearch_b200 as ts
    from theoremsearch_b200 import synthetic
    from theoremsearch_b200.sharded import ShardedIndex, shard_bounds

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout must carry the JSON line only
        if os.environ.get("NCCL_DEBUG", "WARN").upper() in ("WARN", "VERSION"):   # (the launcher's default) NCCL prints its version banner on STDOUT at these levels
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    clocks = ClockSampler(local_rank)          # NVML initialised here, long before any timed region
    peaks, peak_kind = read_peaks()

    lo, hi = shard_bounds(args.rows, world)[rank]
    index = ts.TheoremIndex(args.dim, hi - lo, dtype="bf16", device=dev)
    synthetic.fill_index(index, lo, hi - lo, seed=0)
    torch.cuda.synchronize()
    sharded = ShardedIndex(index, args.rows) if world > 1 else None
    peer = False
    if sharded is not None:
        if args.exchange != "nccl":
            try:   # raises on EVERY rank if any rank cannot map its peers (the ranks agree inside)
                sharded.enable_peer_exchange(max_nq=1, max_k=max(args.k, 32))
                peer = True
            except ts.TheoremSearchError as e:      # CUDA IPC unavailable in this container: say so, use NCCL
                config["exchange_fallback"] = str(e)
        config["exchange"] = (
            "NCCL all-gather of k packed keys + merge kernel" if not peer else
            "device-initiated: NVLink stores into CUDA-IPC peer memory + flags, inside the scan kernel's last CTA"
            if args.exchange == "one-kernel" else
            "device-initiated: NVLink stores into CUDA-IPC peer memory + flags; scan kernel + exchange kernel chained by "
            "programmatic dependent launch, next query's scan overlaps this query's exchange (queries uploaded "
            "beforehand: TS_SHARDED_INDEPENDENT)")

    total = args.warmup + args.steps
    queries = synthetic.make_queries(total, args.dim, dev)           # replicated: same seed on every rank
    q_host = queries.cpu().numpy()
    torch.cuda.synchronize()
    form = {"one_kernel": True} if args.exchange == "one-kernel" else {"independent": True}

    def one_step_device(i):
        q = queries[i:i + 1]
        if sharded is None:
            return index.search(q, args.k, independent=not args.no_stream_chain)
        if peer:
            return sharded.search(q, args.k, **form)
        return sharded.search(q, args.k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region -------------------------------------------------------------
    for i in range(args.warmup):
        one_step_device(i)
    barrier()
    launches0 = ts.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        # Align the ranks ON THE DEVICE before the first timestamp: the exchange of this untimed query completes
        # only when every rank has pushed its keys, so whatever host skew the barrier left is absorbed here
        # (it is enqueued, not synchronised: the host runs ahead and the timed queries queue up behind it).
        if sharded is not None:
            sharded.search(queries[0:1], args.k, **({"one_kernel": True} if peer else {}))
        launches0 = ts.kernel_launches()
        ev0.record()
        for i in range(args.warmup, total):
            out = one_step_device(i)
        ev1.record()
        barrier()
    clock_summary = clocks.summary()
    ms_total = ev0.elapsed_time(ev1)
    launches = ts.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- per-step device times (diagnostic pass, NOT the timed region: an event between two searches
    # serialises them, so this shows the un-overlapped per-query latency and makes a one-off stall visible)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    if sharded is not None:
        sharded.search(queries[0:1], args.k, **({"one_kernel": True} if peer else {}))
    evs[0].record()
    for j, i in enumerate(range(args.warmup, total)):
        one_step_device(i)
        evs[j + 1].record()
    barrier()
    per_step = sorted(evs[j].elapsed_time(evs[j + 1]) for j in range(args.steps))
    per_step_ms = {"median": statistics.median(per_step), "min": per_step[0], "max": per_step[-1],
                   "p95": per_step[int(0.95 * (len(per_step) - 1))],
                   "note": "separate pass with an event after every step (serialises consecutive searches)"}

    # ---- end-to-end region: host buffers in, host buffers out, per step --------------------
    def one_step_host(i):
        if sharded is None:
            return index.search_host(q_host[i], args.k)
        if peer:
            return sharded.search_host(q_host[i], args.k)
        q = torch.from_numpy(q_host[i:i + 1]).pin_memory().to(dev, non_blocking=True)
        s, ids = sharded.search(q, args.k)
        return s.cpu().numpy(), ids.cpu().numpy()

    for i in range(args.warmup):
        one_step_host(i)
    barrier()
    kernel_ms, lat = [], []
    t0 = time.perf_counter()
    for i in range(args.warmup, total):
        t1 = time.perf_counter()
        out_host = one_step_host(i)
        lat.append(time.perf_counter() - t1)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = args.steps / e2e_s

    # ---- roofline of the dominant kernel (K2 scan), timed live --------------------------------
    shard_bytes = (hi - lo) * args.dim * 2
    # the scan kernel alone on this rank, CUDA events recorded around it on its launch stream (ts_ctx timing)
    for i in range(3):
        index.search_host(q_host[i], args.k, timing=True)
    kernel_ms = []
    for i in range(args.warmup, min(total, args.warmup + 50)):
        index.search_host(q_host[i], args.k, timing=True)
        kernel_ms.append(index.last_kernel_ms)
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = shard_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_topk_kernel (K2)", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, copy read+write)",
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "frac_of_nominal_8TBs": achieved / 8000.0,
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": shard_bytes, "traffic": None,
                "step_frac": shard_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    # DRAM traffic per launch from the committed ncu --set full capture of this kernel on this shape
    for name in ("scan_topk_traffic_r2.json", "scan_topk_traffic_r1.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            with open(tp) as f:
                tr = json.load(f)
            if tr["rows"] == hi - lo and tr["dim"] == args.dim:
                roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roofline["traffic_source"] = tr["source"]
            break

    # ---- parity block (untimed) ------------------------------------------------------------------
    parity = {}
    ok = True
    # planted rows: a stored row used as the query must come back first, with its global id
    for g in range(world):
        own_lo, own_hi = shard_bounds(args.rows, world)[g]
        for r_local in (17, (own_hi - own_lo) // 2):
            qrow = index.get_rows(r_local, 1) if g == rank else torch.empty((1, args.dim), device=dev)
            if world > 1:
                dist.broadcast(qrow, src=g)
            s, ids = (sharded.search(qrow, args.k) if sharded is not None else index.search(qrow, args.k))
            ok &= int(ids[0, 0].item()) == own_lo + r_local and abs(float(s[0, 0].item()) - 1.0) < 1e-2
    parity["planted_rows_first"] = bool(ok)
    same = True
    for i in range(4):
        a = one_step_device(i)
        h = one_step_host(i)
        same &= bool(np.array_equal(a[0].cpu().numpy(), h[0]) and np.array_equal(a[1].cpu().numpy(), h[1]))
    parity["host_path_equals_device_path"] = bool(same)
    if sharded is not None and peer:
        agree = True
        for i in range(4):
            q = queries[i:i + 1]
            a = sharded.search(q, args.k)
            b = sharded.search(q, args.k, one_kernel=True)
            c = sharded.search(q, args.k, independent=True)
            sharded._xchg_saved, sharded._xchg = sharded._xchg, None
            d = sharded.search(q, args.k)                      # NCCL all-gather + K5
            sharded._xchg = sharded._xchg_saved
            for o in (b, c, d):
                agree &= bool(torch.equal(a[0], o[0]) and torch.equal(a[1], o[1]))
        parity["pdl_equals_one_kernel_equals_nccl"] = bool(agree)
        ok &= agree
        parity["peer_exchange_timeouts"] = bool(sharded.peer_exchange_error())
        ok &= not sharded.peer_exchange_error()
    ok &= same
    if world > 1:
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    parity["ok_on_every_rank"] = bool(ok)

    # ---- configs[2] / configs[4] beside the headline (single GPU only) ---------------------------
    batched = ivf = None
    if world == 1 and not args.no_extras:
        try:
            batched = batched_section(ts, index, args, dev, clocks, peaks)
        except Exception as e:   # the headline line must survive a failure here; the error is reported, not hidden
            batched = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle
            cores = os.cpu_count() or 1
            sample = host_rows(args.cpu_sample_rows, args.dim, cores)
            qs = oracle.normalize_f64(oracle.synthetic_queries(32, args.dim))
            r = cpu_reference_arm(sample, qs, args.k, 10, 2)
            pg = cpu_prenormalised_arm(sample, qs, args.k, 10)
            scale = args.rows / args.cpu_sample_rows
            cpu = {"value": 1.0 / (r["s_per_step"] * scale), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"10 queries (+2 warm-up) x {args.cpu_sample_rows} of {args.rows} rows (util.cos_sim + full "
                             f"argsort per query as test_app.py:76-77, torch CPU, {cores} threads); q/s scaled by rows "
                             f"({scale:.0f}x); `bench.py --impl reference` times the full corpus",
                   "prenormalised_dot_topk_value": 1.0 / (pg * scale)}
            del sample

    top1 = int(out[1][0, 0].item())
    if sharded is not None:
        assert not sharded.peer_exchange_error(), "a peer missed the exchange time-out"
        sharded.close()
    index.close()
    del index
    torch.cuda.empty_cache()
    if world == 1 and not args.no_extras:
        try:
            ivf = ivf_section(ts, args, dev, clocks, peaks)
        except Exception as e:
            ivf = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "clocks": clock_summary,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": args.dim * 4,
                    "d2h_bytes_per_step": args.k * 12, "p50_latency_ms": statistics.median(lat) * 1e3,
                    "p95_latency_ms": sorted(lat)[int(0.95 * (len(lat) - 1))] * 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "per_step_ms": per_step_ms, "parity_check": parity,
            "top1_id_last_query": top1,
        }
        if batched is not None:
            line["batched"] = batched
        if ivf is not None:
            line["ivf"] = ivf
        print(json.dumps(line), flush=True)
    # The JSON line is out and flushed; tear down in order. A crash in library teardown at interpreter exit must
    # not cost the measurement, so multi-rank runs leave through os._exit once every rank is past the barrier.
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0 if parity["ok_on_every_rank"] else 1


if __name__ == "__main__":
    sys.exit(main())
