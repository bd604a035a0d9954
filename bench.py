#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

Metric: queries/sec (exact top-10) of the single-query scan over a 10M x 1024 bf16 corpus
(configs[1]); a "step" is ONE query = one full pass over the corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* N = 1: the whole corpus on one B200.  N > 1 (under torchrun): the SAME corpus row-sharded
  over N GPUs (strong scaling); per query every rank scans its shard and the scan kernel's last CTA
  exchanges the k packed keys with the peers over NVLink (stores into CUDA-IPC mapped peer memory,
  sequence flags) and merges the N lists — one kernel per query per GPU, no collective call
  (`--exchange nccl` selects the all-gather + K5 merge form instead).
* `value`   : device-timed (CUDA events, max over ranks), queries already resident in HBM.
* `e2e`     : wall-clock through the public host-buffer API (`TheoremIndex.search_host` ->
              `ts_search_host`): per step a 4 KB pinned H2D query copy and a k*(4+8) B D2H result.
* `roofline`: N*D*2 algorithmic bytes / average duration of the K2 scan kernel, measured live with
              CUDA events recorded around that kernel on its launch stream (ts_ctx timing).
* `cpu_baseline` / `--impl reference`: the reference's CPU path (util.cos_sim + argsort,
  test_app.py:76-77) restated in oracle/oracle.py, on the host's cores, on a bounded row sample.
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
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

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
        except Exception as e:  # NVML missing: report that, never fake numbers
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(corpus_fp32: np.ndarray, queries: np.ndarray, k: int, steps: int, warmup: int,
                      n_full: int):
    """The reference's own CPU path, restated (oracle.reference_single_query_verbatim == test_app.py:75-77:
    cos_sim re-normalises the corpus on every call, then a full argsort).  One step = one query
    over the row sample; throughput is scaled to the full corpus by rows (the path is linear in N)."""
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
    per_query_sample = sum(times) / len(times)
    scale = n_full / corpus_fp32.shape[0]
    # pgvector-shaped variant: corpus normalised once at write time, dot + top-k per query
    dbn = torch.nn.functional.normalize(db, dim=1)
    t0 = time.perf_counter()
    reps = max(3, min(steps, 10))
    for i in range(reps):
        s = torch.mv(dbn, torch.from_numpy(queries[i % len(queries)]))
        torch.topk(s, k)
    pg = (time.perf_counter() - t0) / reps
    return {
        "value": 1.0 / (per_query_sample * scale),
        "ms_per_step_sample": per_query_sample * 1e3,
        "p50_ms_sample": statistics.median(times) * 1e3,
        "prenormalised_dot_topk_value": 1.0 / (pg * scale),
        "scale": scale,
    }


def make_cpu_sample(rows: int, dim: int):
    from oracle import oracle
    x = oracle.synthetic_rows(0, rows, dim, seed=0)
    return oracle.bf16_round(oracle.normalize_f64(x))


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
    ap.add_argument("--cpu-sample-rows", type=int, default=500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: in-kernel NVLink peer exchange (default) or NCCL all-gather + merge kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"configs[1]: single-query exact top-{args.k} over {args.rows}x{args.dim} bf16",
              "rows": args.rows, "dim": args.dim, "k": args.k,
              "sharding": f"row-sharded x{world}" if world > 1 else "single GPU",
              "arithmetic": "bf16 corpus rows, fp32 query, fp32 products and accumulation",
              "l2": "corpus shard >> 126 MB L2, distinct query per step (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        steps = min(args.steps, 20)
        sample = make_cpu_sample(args.cpu_sample_rows, args.dim)
        from oracle import oracle
        qs = oracle.normalize_f64(oracle.synthetic_queries(32, args.dim))
        r = cpu_reference_arm(sample, qs, args.k, steps, min(args.warmup, 3), args.rows)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 / r["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{steps} queries x {args.cpu_sample_rows} of {args.rows} rows "
                                       f"(util.cos_sim + full argsort per query, torch CPU, {cores} threads); "
                                       f"q/s scaled by rows ({r['scale']:.0f}x)",
                             "prenormalised_dot_topk_value": r["prenormalised_dot_topk_value"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return 0

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

    lo, hi = shard_bounds(args.rows, world)[rank]
    index = ts.TheoremIndex(args.dim, hi - lo, dtype="bf16", device=dev)
    synthetic.fill_index(index, lo, hi - lo, seed=0)
    torch.cuda.synchronize()
    sharded = ShardedIndex(index, args.rows) if world > 1 else None
    if sharded is not None:
        fused = args.exchange == "fused"
        if fused:
            try:   # raises on EVERY rank if any rank cannot map its peers (the ranks agree inside)
                sharded.enable_peer_exchange(max_nq=1, max_k=max(args.k, 32))
            except ts.TheoremSearchError as e:      # CUDA IPC unavailable in this container: say so, use NCCL
                fused = False
                config["exchange_fallback"] = str(e)
        config["exchange"] = ("in-kernel NVLink peer stores + flags (ts_search_sharded)" if fused
                              else "NCCL all-gather of k packed keys + merge kernel")

    total = args.warmup + args.steps
    queries = synthetic.make_queries(total, args.dim, dev)           # replicated: same seed on every rank
    q_host = queries.cpu().numpy()

    def one_step_device(i):
        q = queries[i:i + 1]
        if sharded is not None:
            return sharded.search(q, args.k)
        return index.search(q, args.k)

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
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for i in range(args.warmup, total):
            out = one_step_device(i)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ts.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- end-to-end region: host buffers in, host buffers out, per step --------------------
    pin_q = torch.empty((1, args.dim), dtype=torch.float32).pin_memory()
    pin_s = torch.empty((1, args.k), dtype=torch.float32).pin_memory()
    pin_i = torch.empty((1, args.k), dtype=torch.int64).pin_memory()

    def one_step_host(i):
        if sharded is None:
            return index.search_host(q_host[i], args.k, timing=True)
        pin_q.copy_(torch.from_numpy(q_host[i:i + 1]))            # host query -> pinned staging
        q = pin_q.to(dev, non_blocking=True)                      # H2D, 4 KB
        s, ids = sharded.search(q, args.k)
        pin_s.copy_(s, non_blocking=True)                         # D2H, k*(4+8) B
        pin_i.copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pin_s.numpy(), pin_i.numpy()

    for i in range(args.warmup):
        one_step_host(i)
    barrier()
    kernel_ms, lat = [], []
    t0 = time.perf_counter()
    for i in range(args.warmup, total):
        t1 = time.perf_counter()
        one_step_host(i)
        lat.append(time.perf_counter() - t1)
        if sharded is None:
            kernel_ms.append(index.last_kernel_ms)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = args.steps / e2e_s

    # ---- roofline of the dominant kernel (K2 scan), timed live --------------------------------
    peaks, peak_kind = read_peaks()
    shard_bytes = (hi - lo) * args.dim * 2
    if sharded is not None:
        # time the scan kernel alone on this rank through the host ctx (same kernel, same shard)
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
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": shard_bytes, "traffic": None}
    # DRAM traffic per launch from the committed ncu --set full capture of this kernel on this shape
    tp = os.path.join(ROOT, "profiles", "scan_topk_traffic_r1.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tr = json.load(f)
        if tr["rows"] == hi - lo and tr["dim"] == args.dim:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_source"] = tr["source"]

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle
            sample = make_cpu_sample(args.cpu_sample_rows, args.dim)
            qs = oracle.normalize_f64(oracle.synthetic_queries(32, args.dim))
            r = cpu_reference_arm(sample, qs, args.k, 10, 2, args.rows)
            cores = os.cpu_count() or 1
            cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"10 queries x {args.cpu_sample_rows} of {args.rows} rows (util.cos_sim + full argsort "
                             f"per query as test_app.py:76-77, torch CPU, {cores} threads); q/s scaled by rows "
                             f"({r['scale']:.0f}x)",
                   "prenormalised_dot_topk_value": r["prenormalised_dot_topk_value"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": args.dim * 4,
                    "d2h_bytes_per_step": args.k * 12, "p50_latency_ms": statistics.median(lat) * 1e3,
                    "p95_latency_ms": sorted(lat)[int(0.95 * (len(lat) - 1))] * 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "top1_id_last_query": int(out[1][0, 0].item()),
        }
        print(json.dumps(line), flush=True)
    # The JSON line is out and flushed; tear down in order (peer mappings, index, process group). A crash in
    # library teardown at interpreter exit must not cost the measurement, so multi-rank runs leave through
    # os._exit once every rank is past the barrier.
    if sharded is not None:
        assert not sharded.peer_exchange_error(), "a peer missed the in-kernel exchange time-out"
        sharded.close()
    index.close()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
