#!/usr/bin/env python
"""bench.py — approximate-count throughput of the B200 path (and of the CPU
reference arm) on BASELINE.json's synthetic workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl b200|reference]

A *step* is one pass of the hot path (errorCount, reference :531-601) over one
batch: all `lim` query k-mers against the sampled read STARTS (n x sl bases) and
the sampled read ENDS (n x (sl+1) bases, :463) — what the reference does once
per run.  Metric: GCUPS = k x Q x (sum of sampled read lengths) / t / 1e9
(SURVEY.md §8d); `queries_per_s` = 2Q / t rides along.

N > 1 (torchrun, one rank per GPU): weak scaling — every rank holds its own
n-read shard of the synthetic read stream, scans it for all queries and the
per-k-mer count vectors are summed with one small NCCL all-reduce per end.

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: reads per GPU, sampled length, k, lim, generator seed
    "C1": dict(n=10_000, sl=100, k=16, lim=500, seed=1001,
               text="C1: synthetic 10k ONT-like reads with planted adapters, k=16, -sn 10000 -sl 100 -lim 500"),
    "C2": dict(n=100_000, sl=100, k=16, lim=2000, seed=1002,
               text="C2: synthetic 100k ONT-like reads with planted adapters, k=16, -sn 100000 -sl 100 -lim 2000"),
    "C3": dict(n=1_000_000, sl=150, k=20, lim=5000, seed=1003,
               text="C3: synthetic 1M reads, k=20, -sn 1000000 -sl 150 -lim 5000"),
    "C4": dict(n=1_000_000, sl=200, k=32, lim=10000, seed=1004,
               text="C4: synthetic 1M reads, k=32, -sn 1000000 -sl 200 -lim 10000"),
}
PARAM_LC = 1.0          # reference default (:711)
ALGO_OPS_PER_COLUMN = 16  # Myers/Hyyro column update, the figure builder and judge share
L2_FLUSH_BYTES = 256 << 20


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region (B200_PROFILING.md "clocks" line).  NVML from a thread every
# few milliseconds when nvidia_ml_py is importable (the timed region of the default run is ~0.2 s, too short
# for `nvidia-smi -lms`), else the nvidia-smi query of the recipe.
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_id):
        self.rows = []      # (t, sm_mhz, sm_max_mhz, power_w, set(reasons))
        self.proc = None
        self.err = None
        self.stop_flag = threading.Event()
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByUUID(gpu_id.encode() if isinstance(gpu_id, str) else gpu_id)
            smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                    ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                    ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))

            def pump():
                while not self.stop_flag.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        self.rows.append((time.perf_counter(), float(sm), float(smax), pw,
                                          {n for n, b in bits if mask & b}))
                    except Exception as e:  # keep what we have
                        self.err = repr(e)
                        break
                    time.sleep(0.004)

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception as e:
            self.err = repr(e)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_id), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except Exception as e:  # neither available: report, do not invent numbers
            self.err = repr(e)

    def _pump_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                row = (time.perf_counter(), float(parts[0]), float(parts[1]), float(parts[2]),
                       {n for n, v in zip(self.NAMES, parts[3:7]) if v.lower().startswith("active")})
            except ValueError:
                continue
            self.rows.append(row)

    def stop(self, t0, t1):
        if self.source is None:
            return {"error": self.err}
        time.sleep(0.06 if self.source == "nvidia-smi" else 0.01)
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        pad = 0.06 if self.source == "nvidia-smi" else 0.0
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + pad]
        if not rows:
            return {"error": "no clock sample fell inside the timed region", "samples": 0, "source": self.source}
        reasons = set()
        for r in rows:
            reasons |= r[4]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "power_w_max": max(r[3] for r in rows), "reasons": sorted(reasons), "samples": len(rows),
                "source": self.source}


# ------------------------------------------------------------------------------------------
def make_ends(w, first, pinned_torch=None, count=None):
    """Sampled starts and ends of synthetic reads [first, first+n): uint8[n, sl], uint8[n, sl+1]."""
    from approx_counter_b200 import host
    n, sl = (w["n"] if count is None else count), w["sl"]
    if pinned_torch is not None:
        torch = pinned_torch
        t0 = torch.empty((n, sl), dtype=torch.uint8, pin_memory=True)
        t1 = torch.empty((n, sl + 1), dtype=torch.uint8, pin_memory=True)
        a, b = t0.numpy(), t1.numpy()
        host.synth_ends(w["seed"], first, n, sl, False, a)
        host.synth_ends(w["seed"], first, n, sl, True, b)
        return (a, b), (t0, t1)
    return (host.synth_ends(w["seed"], first, n, sl, False), host.synth_ends(w["seed"], first, n, sl, True)), None


def columns_per_step(w, q_start, q_end):
    return q_start * w["n"] * w["sl"] + q_end * w["n"] * (w["sl"] + 1)


CPU_NOTES = {
    "fm": "index-based CPU restatement of errorCount as the reference runs it (:531-601): bidirectional FM index of the "
          "sampled reads built per call (:537-541), then every k-mer searched with the optimal-search-scheme recursion "
          "of SeqAn's find<0,2>(EditDistance) and the reference's delegate/popcount reduction, OpenMP over k-mers "
          "(oracle/fm_index_model.cpp); the reference binary itself needs SeqAn, absent from this image. GCUPS here "
          "is the metric's definition (k x Q x bases / t), not cells computed: an index search is sub-linear",
    "scan": "CPU restatement as a scan (oracle/apc_oracle.c, Myers bit-vector per read + OpenMP over k-mers)",
}


def cpu_run(algo, ends, queries, k, r, threads):
    """One pass of the CPU path over the first r reads of both ends: (wall s, columns, index build s, search s)."""
    from oracle import orc
    cols, build, search = 0, 0.0, 0.0
    t0 = time.perf_counter()
    for sample, km in zip(ends, queries):
        codes, offs = orc.encode_matrix(sample[:r])
        if algo == "fm":
            _, (tb, ts) = orc.fm_index_error_count(codes, offs, km, k, nb_thread=threads, want_seconds=True)
            build += tb
            search += ts
        else:
            orc.error_count(codes, offs, km, k, fast=True, nb_thread=threads)
        cols += len(km) * r * sample.shape[1]
    return time.perf_counter() - t0, cols, build, search


def cpu_leg(w, ends, queries, target_s, threads=0, algo="fm"):
    """Time the CPU path (oracle/) on the first R reads of both ends, R sized for about target_s seconds.
    algo "fm": FM index build + search-scheme search, what the reference's errorCount does; "scan": the
    Myers bit-vector scan with the reference's `omp for schedule(dynamic)` over k-mers (:567)."""
    k = w["k"]
    # torchrun exports OMP_NUM_THREADS=1 to every rank: size the team from the CPUs this
    # process may run on, not from the OpenMP default
    if threads <= 0:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    cpu_run(algo, ends, queries, k, min(w["n"], 64), threads)  # library load, OpenMP team start
    r = min(w["n"], 2048 if algo == "fm" else 64)
    t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)  # calibration: grow until the run is long enough to trust
    while t < 0.3 and r < w["n"]:
        r = min(w["n"], r * 4)
        t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)
    for _ in range(3):  # extrapolate to the target, and again if the first guess came out short (first calls are slow)
        r2 = int(max(32, min(w["n"], target_s * r / max(t, 1e-6))))
        if r2 == r or (r2 < r and t < 2 * target_s):
            break
        r = r2
        t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)
        if t > 0.6 * target_s:
            break
    return {"seconds": t, "columns": cols, "reads": r, "threads": threads, "algo": algo,
            "gcups": k * cols / t / 1e9, "index_build_s": build, "search_s": search,
            "queries_per_s": (len(queries[0]) + len(queries[1])) / t * (r / w["n"])}


def reference_queries(w, ends):
    """Top-`lim` exact k-mers of each end with the CPU restatement (:874, :898) — set-up of the
    reference arm only (the B200 arm uses its own exact stage on the GPU)."""
    from oracle import orc
    thr = orc.adjust_threshold(PARAM_LC, 16, w["k"])
    out = []
    for sample in ends:
        codes, offs = orc.encode_matrix(sample)
        keys, cnts, _ = orc.count_kmers(codes, offs, w["k"], thr)
        km, _ = orc.get_most_frequent(keys, cnts, w["lim"], w["k"])
        out.append(km)
    return out


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build_oracle()
    t_gen = time.perf_counter()
    ends, _ = make_ends(w, 0)
    queries = reference_queries(w, ends)
    log(f"[reference] workload + queries ready in {time.perf_counter() - t_gen:.1f}s")
    # bounded sample per step so that warmup+steps stay within a few minutes
    per_step = max(0.5, min(3.0, 150.0 / max(1, args.steps + args.warmup)))
    first = cpu_leg(w, ends, queries, per_step, algo="fm")
    r = first["reads"]
    k = w["k"]
    builds, searches = [], []

    def step():
        t, _, tb, ts = cpu_run("fm", ends, queries, k, r, first["threads"])
        builds.append(tb)
        searches.append(ts)
        return t

    for _ in range(args.warmup):
        step()
    del builds[:], searches[:]
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    cols = len(queries[0]) * r * ends[0].shape[1] + len(queries[1]) * r * ends[1].shape[1]
    value = k * cols * args.steps / total / 1e9
    sample = (f"first {r} of {w['n']} sampled reads of both ends x all {len(queries[0])}+{len(queries[1])} "
              f"query k-mers per step ({cols:.3g} columns/step; index build {sum(builds) / args.steps:.2f} s + "
              f"search {sum(searches) / args.steps:.2f} s per step)")
    line = {
        "impl": "reference", "metric": "approx_count_gcups", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": w["text"] + ", both ends", "k": k, "reads": w["n"], "sl": w["sl"], "lim": w["lim"],
                   "seed": w["seed"]},
        "queries_per_s": (len(queries[0]) + len(queries[1])) * args.steps / total * (r / w["n"]),
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": first["threads"], "kind": "port",
                         "sample": sample, "algo": "fm-index",
                         "index_build_s_per_step": sum(builds) / args.steps,
                         "search_s_per_step": sum(searches) / args.steps,
                         "search_only_gcups": k * cols * args.steps / max(sum(searches), 1e-9) / 1e9,
                         "note": CPU_NOTES["fm"]},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------
def run_b200(args, w):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (one rank per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from approx_counter_b200 import ApproxCounter, host, allreduce_counts, load
    load()  # fails loudly if libapc.so is missing
    k, n, sl, lim = w["k"], w["n"], w["sl"], w["lim"]
    stream = torch.cuda.Stream(dev)  # everything below runs on this (non-default) stream
    torch.cuda.set_stream(stream)

    # ---- inputs: this rank's shard of the synthetic read stream, in pinned host memory
    if args.scaling == "strong":
        # one job of n reads split over the ranks (contiguous blocks of 32-read tiles)
        from approx_counter_b200 import shard_bounds
        lo, hi = shard_bounds(n, rank, world)
        n_total, first, n = n, lo, hi - lo
    else:
        n_total, first = n * world, rank * n
    (h_start, h_end), pinned = make_ends(w, first, pinned_torch=torch, count=n)
    ends = (h_start, h_end)
    ctxs = [ApproxCounter(local_rank), ApproxCounter(local_rank)]
    for c, s in zip(ctxs, ends):
        c.set_option("scan_variant", args.scan_variant)
        c.set_stream(stream.cuda_stream)
        c.upload_sample_ptr(s.ctypes.data, s.shape[0], s.shape[1])

    # ---- queries: top-`lim` exact k-mers per end from the GPU exact stage (rank 0's shard,
    # broadcast so that every rank scans for the same k-mers)
    thr = host.adjust_threshold(PARAM_LC, 16, k)
    queries, exact_ms = [], []
    for c in ctxs:
        km, ct, nd, hn = c.count_kmers_topn(k, thr, lim)
        exact_ms.append(c.timing()["exact_ms"])
        t = torch.zeros(lim, dtype=torch.int64, device=dev)
        cnt = torch.tensor([len(km)], dtype=torch.int64, device=dev)
        t[: len(km)] = torch.from_numpy(km.view(np.int64)).to(dev)
        if world > 1:
            dist.broadcast(t, 0)
            dist.broadcast(cnt, 0)
        queries.append(t[: int(cnt.item())].cpu().numpy().view(np.uint64).copy())
    q_start, q_end = len(queries[0]), len(queries[1])
    if min(q_start, q_end) == 0:
        raise SystemExit("bench.py: the exact stage returned no query k-mers")
    counts = [torch.zeros(len(q), dtype=torch.int64, device=dev) for q in queries]
    for c, q in zip(ctxs, queries):
        if args.plan_alive_pct >= 0:
            c.set_option("plan_alive_pct", args.plan_alive_pct)
        if args.sg_per_job > 0:
            c.set_option("tiles_per_job", args.sg_per_job)
        c.set_queries(q, k)

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    launches = [0]
    kernel_events = []

    def step(record=False):
        for c, out in zip(ctxs, counts):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            c.scan(out.data_ptr())
            if record:
                e1.record(stream)
                kernel_events.append((e0, e1))
            launches[0] += c.timing_launches()
            allreduce_counts(out)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        flush.zero_()
        step()
    barrier()

    # ---- timed region: K steps, device events around every step (the L2 flush between
    # steps sits outside the event pairs), clocks sampled meanwhile
    gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(gpu_uuid if gpu_uuid.startswith("GPU-") else "GPU-" + gpu_uuid) if rank == 0 else None
    time.sleep(0.12)  # let the sampler produce its first rows; every rank waits alike
    barrier()         # ... and all ranks enter the timed region together
    launches[0] = 0
    step_events = []
    t_mark0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step(record=True)
        e1.record(stream)
        step_events.append((e0, e1))
    barrier()
    t_mark1 = time.perf_counter()
    clocks = sampler.stop(t_mark0, t_mark1) if sampler is not None else None
    if os.environ.get("APC_BS_STATS"):  # A/B library built with -DAPC_BS_STATS (tools/build_ab.sh)
        log("[bs_stats] share of tests after which the deep rows were computed:",
            [round(c.microbench("bs_stats"), 4) for c in ctxs])
    dev_ms = sum(a.elapsed_time(b) for a, b in step_events)
    kern_ms = sum(a.elapsed_time(b) for a, b in kernel_events)
    n_launch = launches[0]
    t = torch.tensor([dev_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, kern_ms = float(t[0]), float(t[1])

    cols_rank = q_start * n * sl + q_end * n * (sl + 1)
    cols_job = q_start * n_total * sl + q_end * n_total * (sl + 1)
    value = k * cols_job * args.steps / (dev_ms / 1e3) / 1e9
    final_counts = [c.cpu().numpy().view(np.uint64).copy() for c in counts]

    # ---- end to end through the C ABI with HOST buffers: per step, per end, H2D of the
    # sampled reads + k-mers, scan, (all-reduce), D2H of the counts
    e2e_steps = max(1, min(args.steps, 20))
    h_q = [torch.from_numpy(q.view(np.int64)).pin_memory() for q in queries]
    h_out = [torch.zeros(len(q), dtype=torch.int64).pin_memory() for q in queries]

    # each end gets its own stream so that the upload of one sample overlaps the scan of the other
    e2e_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for c, es in zip(ctxs, e2e_streams):
        c.set_stream(es.cuda_stream)

    def e2e_step():
        for c, es, s, q, o, d in zip(ctxs, e2e_streams, ends, h_q, h_out, counts):
            with torch.cuda.stream(es):
                c.upload_sample_ptr_async(s.ctypes.data, s.shape[0], s.shape[1])
                if world == 1:
                    c.errorCount_ptr_async(q.data_ptr(), q.numel(), k, o.data_ptr())
                else:
                    c.set_queries_ptr(q.data_ptr(), q.numel(), k)
                    c.scan(d.data_ptr())
                    allreduce_counts(d)
                    o.copy_(d, non_blocking=True)
        torch.cuda.synchronize(dev)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    e2e_value = k * cols_job * e2e_steps / e2e_s / 1e9
    for o, f in zip(h_out, final_counts):
        if not np.array_equal(o.numpy().view(np.uint64), f):
            raise SystemExit("bench.py: end-to-end counts differ from the resident-path counts")
    h2d = int(h_start.nbytes + h_end.nbytes + 8 * (q_start + q_end))
    d2h = int(8 * (q_start + q_end))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (approx_scan_kernel)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    int_peak = ctxs[0].measure_int_peak()
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    alu_peak = sms * 4 * 16 * sm_max * 1e6  # ALU pipe: 16 lanes/clk/SMSP (B300_MICROARCH.md:85)
    # a scan is up to three launches (k-mer quads, pairs, then the ungrouped k-mers): the roofline unit is the scan
    n_scans = max(len(kernel_events), 1)
    per_launch_s = kern_ms / 1e3 / n_scans
    cols_launch = cols_rank / 2.0
    # no kernel can retire more than one instruction per SMSP per clock, and even four k-mers sharing 3k/4 of
    # their rows cost 5 * 7k/16 / 32 > 0.5 lane-operations per column; a faster reading means the event pairs
    # did not contain the kernel (e.g. it ran on another stream)
    issue_peak = sms * 4 * 32 * sm_max * 1e6
    if 0.5 * cols_launch / per_launch_s > issue_peak:
        raise SystemExit("bench.py: implausible kernel time — the timed region did not contain the scan kernel")
    achieved = ALGO_OPS_PER_COLUMN * cols_launch / per_launch_s
    traffic = ncu_alu = None
    try:  # figures from the committed ncu capture of this kernel on this workload (profiles/)
        prof = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")))
        traffic = prof.get(args.workload, {}).get("dram_bytes_per_launch")
        ncu_alu = prof.get(args.workload, {}).get("alu_pipe_pct")
    except Exception:
        pass
    # bit planes: one uint4 per (32-read group, column), columns padded to 16, groups to 32 per super-group;
    # + the k-mers and the counts
    cols_padded = ((sl + 1 + 15) // 16) * 16
    hbm_bytes_launch = ((n + 1023) // 1024) * 32 * cols_padded * 16 + 16 * q_start
    # What the scan plan costs on the ALU pipe: it groups k-mers into units whose shared rows are computed once;
    # a row of a unit costs 5 LOP3 per column and 1024 reads (32 lanes x 32 reads; bitslice_core.cuh).  Dead-row
    # skipping then leaves out the deep rows of a unit in the columns where nothing can reach them, which is
    # data dependent: `planned` is the instruction count WITHOUT skipping (an upper bound of what is executed),
    # the executed share is read from ncu (sm__inst_executed_pipe_alu, profiles/).
    from approx_counter_b200 import plan_queries
    plan_rows, plan_units, plan_reversed = [], [], []
    for q in queries:
        pl = plan_queries(q, k)
        rows = sum(int(u) * (k - int(t) + int(g) * int(t)) for u, t, g in zip(pl["units"], pl["shape_t"], pl["shape_g"]))
        rows += (len(q) - int((pl["units"] * pl["shape_g"]).sum())) * k
        plan_rows.append(rows)
        plan_units.append([int(u) for u in pl["units"]])
        plan_reversed.append(int(pl["reversed"].sum()))
    n_sg_rank = (n + 1023) // 1024
    lop3_lane_ops = sum(5.0 * 32 * rows * n_sg_rank * (2 * ((L + 1) // 2)) for rows, L in zip(plan_rows, (sl, sl + 1)))
    planned = lop3_lane_ops / (kern_ms / 1e3 / max(args.steps, 1))
    roofline = {
        "bound": "int-alu", "kernel": "bs_group_kernel<K,P,G> (one launch per unit shape in use) + bs_scan_kernel<K> (ungrouped k-mers); "
                                      "one scan = up to 13 concurrent launches", "achieved": achieved / 1e12, "peak": alu_peak / 1e12,
        "unit": "Tint-op/s", "frac": achieved / alu_peak, "traffic": traffic, "ncu_alu_pipe_pct": ncu_alu,
        "peak_source": f"ALU pipe nominal = {sms} SM x 4 SMSP x 16 lanes/clk x {sm_max:.0f} MHz (SURVEY.md §8d)",
        "algorithmic_ops_per_column": ALGO_OPS_PER_COLUMN, "columns_per_launch": cols_launch,
        "avg_launch_ms": per_launch_s * 1e3, "scans_timed": n_scans, "launches_timed": n_launch,
        "kernel_share_of_step": kern_ms / dev_ms if dev_ms else None,
        "planned": {"lop3_Tlane_ops_per_s": planned / 1e12, "frac_of_alu_peak": planned / alu_peak,
                    "rows_per_scan": plan_rows, "rows_if_one_kmer_per_warp": [len(q) * k for q in queries],
                    "units_per_shape": plan_units, "kmers_scanned_backwards": plan_reversed,
                    "note": "LOP3 of the scan plan if every unit row were computed in every column (5 per row, column "
                            "and 1024 reads) / time, against the ALU peak; above 1 = what dead-row skipping saves. "
                            "`frac` is against the ALGORITHMIC 16 ops per column; the executed utilisation is "
                            "ncu_alu_pipe_pct"},
        "measured_int_peaks_Tops": {kk: v / 1e12 for kk, v in int_peak.items()},
        "frac_of_measured_lop3_peak": achieved / int_peak["lop3_ops_per_s"],
        "hbm": {"algorithmic_bytes_per_launch": hbm_bytes_launch,
                "achieved_gbs": hbm_bytes_launch / per_launch_s / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"), "note": "the text (bit planes, 0.5 B per base) is re-read from L2 "
                "by every k-mer; the kernel is bound by the ALU pipe (LOP3), HBM is idle"},
        "sm_mhz_during_run": sm_mhz,
    }

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import __graft_entry__ as g
        g.build_oracle()
        leg = cpu_leg(w, ends, queries, target_s=args.cpu_seconds, algo="fm")
        scan = cpu_leg(w, ends, queries, target_s=min(4.0, args.cpu_seconds), algo="scan")
        cpu = {"value": leg["gcups"], "unit": "GCUPS", "cores": leg["threads"], "kind": "port", "algo": "fm-index",
               "sample": f"first {leg['reads']} of {n} sampled reads of both ends x all {q_start}+{q_end} query "
                         f"k-mers ({leg['columns']:.3g} columns, {leg['seconds']:.1f} s wall: index build "
                         f"{leg['index_build_s']:.2f} s + search {leg['search_s']:.2f} s)",
               "index_build_s": leg["index_build_s"], "search_s": leg["search_s"],
               "search_only_gcups": k * leg["columns"] / max(leg["search_s"], 1e-9) / 1e9,
               "scan_port": {"value": scan["gcups"], "unit": "GCUPS", "cores": scan["threads"],
                             "sample": f"first {scan['reads']} reads of both ends, {scan['seconds']:.1f} s wall",
                             "note": CPU_NOTES["scan"]},
               "host_cpus": os.cpu_count(), "note": CPU_NOTES["fm"]}
        # the CPU leg doubles as a spot check of the GPU counts on its sample
        from oracle import orc
        r = min(leg["reads"], 512)
        chk = ApproxCounter(local_rank)
        chk.upload_sample(np.ascontiguousarray(h_start[:r]))
        got = chk.errorCount(queries[0], k)
        codes, offs = orc.encode_matrix(h_start[:r])
        want = orc.error_count(codes, offs, queries[0], k, fast=True)
        chk.close()
        if not np.array_equal(got, want):
            raise SystemExit("bench.py: GPU counts differ from the oracle on the CPU-baseline sample")

    line = {
        "metric": "approx_count_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": w["text"] + ", both ends (start n x sl, end n x (sl+1))", "k": k,
                   "reads_per_gpu": n, "reads_total": n_total, "sl": sl, "lim": lim, "queries": [q_start, q_end], "seed": w["seed"],
                   "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset outside the event pairs)",
                   "parallelism": f"reads sharded over {world} GPU(s), one all-reduce of Q u64 per end"},
        "queries_per_s": (q_start + q_end) * args.steps / (dev_ms / 1e3),
        "columns_per_s": cols_job * args.steps / (dev_ms / 1e3),
        "wall_s_timed_region": t_mark1 - t_mark0,
        "exact_stage_ms": exact_ms,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "path": "apc_upload_sample_async + apc_approx_count_async (C ABI, pinned host buffers, one stream per end), wall clock"},
        "gpu_launches": n_launch,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly one JSON line: park fd 1 on stderr while libraries (NCCL banner,
    # torchrun notices) may print, and write the line to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="C2")
    ap.add_argument("--scan-variant", type=int, default=0,
                    help="kernel A/B: 0 bit-sliced with k-mer pairing (default), 7 bit-sliced, 8 row-packed")
    ap.add_argument("--reads", type=int, default=0, help="override the reads per GPU of the workload (C5 sweep)")
    ap.add_argument("--lim", type=int, default=0, help="override the number of query k-mers (C5 sweep)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: every rank holds its own n-read shard (default); strong: one n-read job split over the ranks")
    ap.add_argument("--plan-alive-pct", type=int, default=-1,
                    help="planner knob (apc_set_option plan_alive_pct): expected share of columns with live deep rows")
    ap.add_argument("--sg-per-job", type=int, default=0,
                    help="tuning knob (apc_set_option tiles_per_job): 1024-read super-groups per job, 0 = auto")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample size, seconds of CPU work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = dict(WORKLOADS[args.workload])
    if args.reads or args.lim:  # C5: the scaling sweep re-sizes a named workload
        w["n"] = args.reads or w["n"]
        w["lim"] = args.lim or w["lim"]
        w["text"] = f"C5 sweep point on {args.workload}: -sn {w['n']} -sl {w['sl']} -lim {w['lim']}, k={w['k']}"
    if args.impl == "reference":
        return run_reference(args, w)
    return run_b200(args, w)


if __name__ == "__main__":
    sys.exit(main())
