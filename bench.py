#!/usr/bin/env python
"""bench.py — approximate-count throughput of the B200 path (and of the CPU
reference arm) on BASELINE.json's synthetic workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--scaling strong|weak] [--impl b200|reference]

A *step* is one pass of the hot path (errorCount, reference :531-601) over one
batch: all `lim` query k-mers against the sampled read STARTS (n x sl bases) and
the sampled read ENDS (n x (sl+1) bases, :463) — what the reference does once
per run.  Metric: GCUPS = k x Q x (sum of sampled read lengths) / t / 1e9
(SURVEY.md §8d); `queries_per_s` = 2Q / t rides along.

Default workload: BASELINE config 3 (1M reads, k=20, sl=150, lim=5000 — the
configuration BASELINE.json assigns to 2/4/8 GPUs; it fits one GPU and a step
is tens of milliseconds, so the timed region is about a second).  N > 1
(torchrun, one rank per GPU): STRONG scaling by default — the one n-read job is
split over the ranks in contiguous blocks of reads, every rank scans its shard
for all queries and the count vectors of both ends are summed with ONE small
NCCL all-reduce per step, issued through the C ABI (apc_allreduce_counts).
The C2 weak-scaling number of round 1 rides along as `c2_weak_value`.

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
def host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to every rank: size teams from the CPUs this process may run on
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_config(w, scaling, world):
    """The `config` object — the SAME keys and values in both arms (the driver compares them)."""
    return {"workload": w["text"] + ", both ends (start n x sl, end n x (sl+1))", "k": w["k"],
            "reads_total": w["n"] * (world if scaling == "weak" else 1), "sl": w["sl"], "lim": w["lim"],
            "seed": w["seed"], "scaling": scaling,
            "l2": f"GPU arm: flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset outside the event pairs)"}


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
    n_have = ends[0].shape[0]
    if threads <= 0:
        threads = host_threads()
    cpu_run(algo, ends, queries, k, min(n_have, 64), threads)  # library load, OpenMP team start
    r = min(n_have, 2048 if algo == "fm" else 64)
    t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)  # calibration: grow until the run is long enough to trust
    while t < 0.3 and r < n_have:
        r = min(n_have, r * 4)
        t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)
    for _ in range(3):  # extrapolate to the target, and again if the first guess came out short (first calls are slow)
        r2 = int(max(32, min(n_have, target_s * r / max(t, 1e-6))))
        if r2 == r or (r2 < r and t < 2 * target_s):
            break
        r = r2
        t, cols, build, search = cpu_run(algo, ends, queries, k, r, threads)
        if t > 0.6 * target_s:
            break
    return {"seconds": t, "columns": cols, "reads": r, "threads": threads, "algo": algo,
            "gcups": k * cols / t / 1e9, "index_build_s": build, "search_s": search}


def ingest_leg(w, name, device):
    """The step before the path on this workload's input file: the host route (parallel parser,
    sampler, upload of both ends) against the device route (apc_ingest_fastx + apc_sample_resident), wall clock of the
    second of two passes, and a byte-for-byte comparison of the samples both leave on the GPU."""
    import mmap
    import tempfile
    from approx_counter_b200 import ApproxCounter, host
    n, sl, fastq = w["n"], w["sl"], name == "C3"
    path = os.path.join(tempfile.gettempdir(), f"apc_bench_{os.getpid()}.{'fq' if fastq else 'fa'}")
    host.synth_write(path, w["seed"], n, sl, fastq=fastq)
    size = os.path.getsize(path)
    c = ApproxCounter(device)
    out = {}
    try:
        same = True
        for _ in range(2):
            t0 = time.perf_counter()
            reads = host.Reads(path)
            t1 = time.perf_counter()
            rows = []
            for bot in (False, True):
                s = reads.sample(n, sl, bot, 7)
                c.upload_sample(s)
                rows.append(s)
            t2 = time.perf_counter()
            reads.close()
            with open(path, "rb") as f:
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
                buf = np.frombuffer(mm, np.uint8)
                t3 = time.perf_counter()
                nrec, _ = c.ingest_fastx(buf)
                t4 = time.perf_counter()
                tm = c.ingest_timing()
                del buf
                mm.close()
            order = host.shuffle_order(nrec, 7)
            t5 = time.perf_counter()
            kernels_ms = 0.0
            for bot in (False, True):
                c.sample_resident(n, sl, bot, order)
                kernels_ms += c.ingest_timing()["sample_ms"]
            t6 = time.perf_counter()
            out = {"file_mb": size / 1e6, "format": "fastq" if fastq else "fasta", "records": int(nrec),
                   "host_ms": (t2 - t0) * 1e3, "host_parse_ms": (t1 - t0) * 1e3, "host_sample_upload_ms": (t2 - t1) * 1e3,
                   "device_ms": (t4 - t3 + t6 - t5) * 1e3, "device_copy_ms": tm["copy_ms"], "device_index_ms": tm["index_ms"],
                   "device_sample_ms": (t6 - t5) * 1e3, "device_sample_kernels_ms": kernels_ms,
                   "shuffle_ms_not_counted": (t5 - t4) * 1e3,
                   "copy_gbs": size / 1e6 / max(tm["copy_ms"], 1e-6), "index_gbs": size / 1e6 / max(tm["index_ms"], 1e-6),
                   "what": "input file of this workload -> both ends' samples resident as scan tiles and bit planes; host: "
                           "parallel parser + sampler + apc_upload_sample; device: apc_ingest_fastx (file bytes through "
                           "page-locked staging, newline and record index kernels) + apc_sample_resident (pick, gather, "
                           "layout kernels); wall clock, second of two passes, page cache warm; the shuffle of the ids "
                           "(host, both routes) is outside both figures' difference"}
        for bot in (False, True):  # the device route's last sample is the end sample: compare both ends once more
            c.sample_resident(n, sl, bot, order)
            same = same and np.array_equal(c.download_sample(), rows[int(bot)])
        out["samples_identical"] = bool(same)
        # the whole per-file pipeline through the C ABI, CUDA context and buffers warm (apc_reserve, as the binary does
        # beside the creation of its context): file bytes (mapped) -> both ends' exact top-lim and approximate counts on
        # the host (reference :819-928 without the export)
        thr = host.adjust_threshold(PARAM_LC, 16, w["k"])
        c.reserve(n, sl + 1, w["k"], w["lim"])
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            buf = np.frombuffer(mm, np.uint8)
            t0 = time.perf_counter()
            nrec, _ = c.ingest_fastx(buf)
            t1 = time.perf_counter()
            phases = {"ingest_ms": (t1 - t0) * 1e3, "shuffle_ms": 0.0, "sample_ms": 0.0, "exact_ms": 0.0, "approx_ms": 0.0}
            for bot in (False, True):
                ta = time.perf_counter()
                order = host.shuffle_order(nrec, 7) if n < nrec else None  # every read wanted: any order gives the same set
                tb = time.perf_counter()
                c.sample_resident(n, sl, bot, order)
                tc = time.perf_counter()
                km, _, _, _ = c.count_kmers_topn(w["k"], thr, w["lim"])
                td = time.perf_counter()
                c.errorCount(km, w["k"])
                te = time.perf_counter()
                phases["shuffle_ms"] += (tb - ta) * 1e3
                phases["sample_ms"] += (tc - tb) * 1e3
                phases["exact_ms"] += (td - tc) * 1e3
                phases["approx_ms"] += (te - td) * 1e3
            phases["total_ms"] = (time.perf_counter() - t0) * 1e3
            del buf
            mm.close()
        out["pipeline_from_file"] = dict(phases, what="apc_ingest_fastx -> per end: apch_shuffle_order (skipped when every read is sampled, as in the binary), apc_sample_resident, "
                                         "apc_exact_topn, apc_approx_count (host in/out); wall clock, one pass, context warm")
        if not same:
            raise SystemExit("bench.py: PARITY FAILURE — device ingest and host ingest leave different samples")
    finally:
        c.close()
        os.unlink(path)
    return out


def reference_queries(w, ends):
    """Top-`lim` exact k-mers of each end with the CPU restatement (:874, :898; all host threads for the count,
    threshold-cut top-N — both held equal to the plain restatement in tests/test_oracle.py).  Set-up of the
    reference arm only: the B200 arm gets its queries from its own exact stage on the GPU."""
    from oracle import orc
    thr = orc.adjust_threshold(PARAM_LC, 16, w["k"])
    out = []
    for sample in ends:
        codes, offs = orc.encode_matrix(sample)
        keys, cnts, _ = orc.count_kmers_mt(codes, offs, w["k"], thr)
        km, _ = orc.get_most_frequent_fast(keys, cnts, w["lim"], w["k"])
        out.append(km)
    return out


def run_reference(args, w):
    """The reference arm: the CPU path on the box's host cores, everything from oracle/ (the product library is
    never loaded in this process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build_oracle()
    from oracle import orc
    world = max(1, args.gpus)
    n_total = w["n"] * (world if args.scaling == "weak" else 1)
    t_gen = time.perf_counter()
    ends = (orc.synth_ends(w["seed"], 0, w["n"], w["sl"], False), orc.synth_ends(w["seed"], 0, w["n"], w["sl"], True))
    queries = reference_queries(w, ends)
    log(f"[reference] workload + queries ready in {time.perf_counter() - t_gen:.1f}s")
    # bounded sample per step so that warmup+steps stay within a few minutes
    per_step = max(0.5, min(3.0, 120.0 / max(1, args.steps + args.warmup)))
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
    sample = (f"first {r} of {n_total} sampled reads of both ends x all {len(queries[0])}+{len(queries[1])} "
              f"query k-mers per step ({cols:.3g} columns/step; index build {sum(builds) / args.steps:.2f} s + "
              f"search {sum(searches) / args.steps:.2f} s per step)")
    line = {
        "impl": "reference", "metric": "approx_count_gcups", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": make_config(w, args.scaling, world),
        "queries_per_s": (len(queries[0]) + len(queries[1])) * args.steps / total * (r / n_total),
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
class Dist:
    """torch.distributed as plumbing: barrier, max over ranks, broadcast of small host arrays."""

    def __init__(self, torch, dev, world):
        self.torch, self.dev, self.world = torch, dev, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=dev)
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def bcast_u64(self, arr, cap):
        """Broadcast a uint64 host array of at most `cap` entries from rank 0."""
        torch = self.torch
        t = torch.zeros(cap + 1, dtype=torch.int64, device=self.dev)
        if arr is not None:
            t[0] = len(arr)
            t[1: 1 + len(arr)] = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).to(self.dev)
        if self.dist:
            self.dist.broadcast(t, 0)
        h = t.cpu().numpy()
        return h[1: 1 + int(h[0])].view(np.uint64).copy()

    def bcast_bytes(self, b, n):
        torch = self.torch
        t = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        if b is not None:
            t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
        if self.dist:
            self.dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


class Job:
    """One workload resident on this rank's GPU: the rank's shard of both sampled ends, the query k-mers of both
    ends, ONE device vector for the counts of both ends (so that the ranks' vectors are summed with a single
    all-reduce per step) — everything through the C ABI (ApproxCounter = ctypes over include/apc.h)."""

    def __init__(self, torch, dev, local_rank, stream, w, first, n, pinned=True, seed=None):
        from approx_counter_b200 import ApproxCounter, host
        self.torch, self.dev, self.w, self.n, self.first, self.stream = torch, dev, w, n, first, stream
        sl = w["sl"]
        seed = w["seed"] if seed is None else seed
        self.pinned = [torch.empty((n, sl + b), dtype=torch.uint8, pin_memory=pinned) for b in (0, 1)]
        self.ends = [t.numpy() for t in self.pinned]
        for b, a in enumerate(self.ends):
            host.synth_ends(seed, first, n, sl, bool(b), a)
        self.ctxs = [ApproxCounter(local_rank), ApproxCounter(local_rank)]
        # the read ends are independent samples: the scan of the `end` sample runs on a second stream beside the scan of
        # the `start` sample (fork / join with events around the caller's stream), which keeps the SMs busy while
        # one scan runs out of jobs
        # ... when both samples' bit planes (0.5 B per base) fit in L2 together; otherwise the two scans only evict each
        # other's text (C3: 2 x 80 MB against 126 MB of L2, measured 46.6 against 46.1 ms per step) and run in turn
        planes_bytes = sum(((n + 1023) // 1024) * 1024 * ((sl + b + 15) // 16) * 16 // 2 for b in (0, 1))
        self.side = torch.cuda.Stream(dev) if planes_bytes <= (48 << 20) else stream
        for c, s, st in zip(self.ctxs, self.ends, (stream, self.side)):
            c.set_stream(st.cuda_stream)
            c.upload_sample_ptr(s.ctypes.data, s.shape[0], s.shape[1])
        torch.cuda.synchronize(dev)
        self.queries = None

    def exact_queries(self, lim):
        """Top-`lim` exact k-mers of each end from the GPU exact stage over what this job holds."""
        from approx_counter_b200 import host
        thr = host.adjust_threshold(PARAM_LC, 16, self.w["k"])
        out, ms = [], []
        for c in self.ctxs:
            km, _, _, _ = c.count_kmers_topn(self.w["k"], thr, lim)
            ms.append(c.timing()["exact_ms"])
            out.append(km)
        return out, ms

    def set_queries(self, queries, options=()):
        torch = self.torch
        self.queries = queries
        self.q = [len(q) for q in queries]
        self.counts = torch.zeros(sum(self.q), dtype=torch.int64, device=self.dev)
        self.ptrs = [self.counts.data_ptr(), self.counts.data_ptr() + 8 * self.q[0]]
        for c, q in zip(self.ctxs, queries):
            for name, value in options:
                c.set_option(name, value)
            c.set_queries(q, self.w["k"])
        torch.cuda.synchronize(self.dev)

    def step(self, kernel_events=None):
        """Scan both ends into the shared count vector, then one all-reduce over the ranks (a no-op on one GPU)."""
        torch = self.torch
        launches = 0
        if kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
        if self.side is not self.stream:
            fork = torch.cuda.Event()
            fork.record(self.stream)
            self.side.wait_event(fork)
        for c, p in zip(self.ctxs, self.ptrs):
            c.scan(p)
            launches += c.timing_launches()
        if self.side is not self.stream:
            join = torch.cuda.Event()
            join.record(self.side)
            self.stream.wait_event(join)
        if kernel_events is not None:
            e1.record(self.stream)
            kernel_events.append((e0, e1))
        self.ctxs[0].allreduce_counts(self.ptrs[0], sum(self.q))
        return launches

    def columns(self, n_reads=None):
        n = self.n if n_reads is None else n_reads
        sl = self.w["sl"]
        return self.q[0] * n * sl + self.q[1] * n * (sl + 1)

    def stats(self):
        a, b = (c.scan_stats() for c in self.ctxs)
        return {k: a[k] + b[k] for k in a}

    def host_counts(self):
        return self.counts.cpu().numpy().view(np.uint64).copy()

    def close(self):
        for c in self.ctxs:
            c.close()


def timed_steps(torch, job, dist, flush, steps, warmup, sampler_factory=None):
    """W untimed + K timed steps, CUDA events on the launching stream around every step (the L2 flush between steps
    sits outside the event pairs), barrier + synchronize on both sides.  Returns device ms (max over ranks),
    kernel ms (scan launches only), launches, wall marks, clocks."""
    stream = job.stream
    for _ in range(max(warmup, 0)):
        flush.zero_()
        job.step()
    dist.barrier()
    job.stats()  # reset the executed-row tally: only the timed steps count
    sampler = sampler_factory() if sampler_factory else None
    if sampler_factory:
        time.sleep(0.12)  # let the sampler produce its first rows; every rank waits alike
    dist.barrier()
    launches = 0
    step_events, kernel_events = [], []
    t0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches += job.step(kernel_events)
        e1.record(stream)
        step_events.append((e0, e1))
    dist.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    dev_ms = sum(a.elapsed_time(b) for a, b in step_events)
    kern_ms = sum(a.elapsed_time(b) for a, b in kernel_events)
    dev_ms, kern_ms = dist.max(dev_ms, kern_ms)
    return {"dev_ms": dev_ms, "kern_ms": kern_ms, "launches": launches, "t0": t0, "t1": t1, "clocks": clocks,
            "stats": job.stats()}


def run_e2e(torch, dev, job, dist, steps):
    """The same metric end to end through the C ABI with HOST buffers: per step and per end the H2D copy of the
    sampled reads (pinned) and of the k-mers, query planning, scan; one all-reduce; D2H of the counts.  The two
    ends run on two streams (the upload of one overlaps the scan of the other) and consecutive steps are
    pipelined two deep with alternating output buffers, as a host loop over batches would do; wall clock around
    the whole loop, every step's counts checked afterwards."""
    k = job.w["k"]
    world = dist.world
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    reduce_stream = torch.cuda.Stream(dev)  # all-reduce + D2H: the next step's uploads and scans do not wait for them
    h_q = [torch.from_numpy(q.view(np.int64)).pin_memory() for q in job.queries]
    nq = sum(job.q)
    h_out = [torch.zeros(nq, dtype=torch.int64).pin_memory() for _ in range(2)]
    d_out = [torch.zeros(nq, dtype=torch.int64, device=dev) for _ in range(2)]
    done = [None, None]
    for c, es in zip(job.ctxs, streams):
        c.set_stream(es.cuda_stream)

    def enqueue(i):
        slot = i & 1
        if done[slot] is not None:
            done[slot].synchronize()  # the output buffers of step i-2 are free again
        base = d_out[slot].data_ptr()
        for e, (c, es, s, q) in enumerate(zip(job.ctxs, streams, job.ends, h_q)):
            with torch.cuda.stream(es):
                c.upload_sample_ptr_async(s.ctypes.data, s.shape[0], s.shape[1])
                c.set_queries_ptr(q.data_ptr(), q.numel(), k)
                c.scan(base + (8 * job.q[0] if e else 0))
            ev = torch.cuda.Event()
            ev.record(es)
            reduce_stream.wait_event(ev)
        with torch.cuda.stream(reduce_stream):
            job.ctxs[0].set_stream(reduce_stream.cuda_stream)
            job.ctxs[0].allreduce_counts(base, nq)
            job.ctxs[0].set_stream(streams[0].cuda_stream)
            h_out[slot].copy_(d_out[slot], non_blocking=True)
        done[slot] = torch.cuda.Event()
        done[slot].record(reduce_stream)

    for i in range(3):
        enqueue(i)
    torch.cuda.synchronize(dev)
    done[0] = done[1] = None
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        enqueue(i)
    torch.cuda.synchronize(dev)
    dist.barrier()
    e2e_s = dist.max(time.perf_counter() - t0)[0]
    for c, st in zip(job.ctxs, (job.stream, job.side)):
        c.set_stream(st.cuda_stream)
    h2d = int(job.ends[0].nbytes + job.ends[1].nbytes + 8 * nq)
    outs = [h.numpy().view(np.uint64).copy() for h in h_out]
    return e2e_s, h2d, int(8 * nq), outs


def run_b200(args, w):
    import torch

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
    dist = Dist(torch, dev, world)

    from approx_counter_b200 import ApproxCounter, host, load, shard_bounds
    load()  # fails loudly if libapc.so is missing
    k, sl, lim = w["k"], w["sl"], w["lim"]
    stream = torch.cuda.Stream(dev)  # everything below runs on this (non-default) stream
    torch.cuda.set_stream(stream)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    options = [("scan_variant", args.scan_variant)]
    if args.plan_alive_pct >= 0:
        options.append(("plan_alive_pct", args.plan_alive_pct))
    if args.sg_per_job > 0:
        options.append(("tiles_per_job", args.sg_per_job))
    if args.shape_mask >= 0:
        options.append(("shape_mask", args.shape_mask))
    if args.graph:
        options.append(("scan_graph", 1))

    # ---- this rank's shard of the synthetic read stream, in pinned host memory
    if args.scaling == "strong":  # one job of n reads split over the ranks (contiguous blocks of 32-read tiles)
        n_total = w["n"]
        lo, hi = shard_bounds(n_total, rank, world)
        first, n = lo, hi - lo
    else:                         # every rank holds its own n-read shard of the stream
        n_total, first, n = w["n"] * world, rank * w["n"], w["n"]
    job = Job(torch, dev, local_rank, stream, w, first, n)

    # ---- the ranks form one NCCL communicator through the C ABI (apc_comm_*): rank 0 creates the id
    if world > 1:
        uid = dist.bcast_bytes(ApproxCounter.comm_unique_id() if rank == 0 else None, 128)
        job.ctxs[0].comm_init_rank(world, rank, uid)

    # ---- queries: the top-`lim` exact k-mers per end of the WHOLE job's sample from the GPU exact stage (rank 0;
    # broadcast so that every rank scans for the same k-mers).  With the sample sharded, rank 0 loads the whole
    # sample once for this set-up step.
    exact_ms = None
    qs = [None, None]
    if rank == 0:
        if world > 1:
            whole = Job(torch, dev, local_rank, stream, w, 0, n_total, pinned=False)
            qs, exact_ms = whole.exact_queries(lim)
            whole.close()
            del whole
        else:
            qs, exact_ms = job.exact_queries(lim)
            qs2, exact_ms2 = job.exact_queries(lim)  # second call: scratch allocated, kernels loaded
            exact_ms = {"first_call": exact_ms, "steady": exact_ms2}
    queries = [dist.bcast_u64(q, lim) for q in qs]
    if min(len(q) for q in queries) == 0:
        raise SystemExit("bench.py: the exact stage returned no query k-mers")
    job.set_queries(queries, options)
    q_start, q_end = job.q

    # ---- timed region
    gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    factory = (lambda: ClockSampler(gpu_uuid if gpu_uuid.startswith("GPU-") else "GPU-" + gpu_uuid)) if rank == 0 else None
    res = timed_steps(torch, job, dist, flush, args.steps, args.warmup, factory)
    dev_ms, kern_ms = res["dev_ms"], res["kern_ms"]
    cols_rank, cols_job = job.columns(), job.columns(n_total)
    value = k * cols_job * args.steps / (dev_ms / 1e3) / 1e9
    final_counts = job.host_counts()

    # ---- parity self-check of the multi-GPU result: rank 0 rescans the union of all shards on ITS GPU alone
    parity = None
    if world > 1:
        if rank == 0:
            tot = np.zeros(q_start + q_end, np.uint64)
            chunk = max(1, min(n_total, 1 << 20))
            for lo in range(0, n_total, chunk):
                u = Job(torch, dev, local_rank, stream, w, lo, min(chunk, n_total - lo), pinned=False)
                u.set_queries(queries, options)
                u.step()
                torch.cuda.synchronize(dev)
                tot += u.host_counts()
                u.close()
                del u
            if not np.array_equal(tot, final_counts):
                raise SystemExit("bench.py: PARITY FAILURE — the all-reduced counts of the sharded scan differ from a "
                                 "single-GPU scan of the same reads")
            parity = f"ok: all-reduced counts of {world} shards == single-GPU scan of all {n_total} reads ({q_start}+{q_end} k-mers)"

    # ---- end to end through the C ABI with host buffers
    e2e_steps = max(2, min(args.steps, 20))
    e2e_s, h2d, d2h, outs = run_e2e(torch, dev, job, dist, e2e_steps)
    e2e_value = k * cols_job * e2e_steps / e2e_s / 1e9
    for o in outs:
        if not np.array_equal(o, final_counts):
            raise SystemExit("bench.py: end-to-end counts differ from the resident-path counts")

    # ---- continuity with round 1: C2, weak scaling (every rank its own 100k-read shard), device-timed only
    c2 = None
    if args.workload != "C2" and not args.no_extras:
        w2 = dict(WORKLOADS["C2"])
        j2 = Job(torch, dev, local_rank, stream, w2, rank * w2["n"], w2["n"])
        if world > 1:
            j2.ctxs[0].comm_init_rank(world, rank, dist.bcast_bytes(ApproxCounter.comm_unique_id() if rank == 0 else None, 128))
        q2 = j2.exact_queries(w2["lim"])[0] if rank == 0 else [None, None]
        j2.set_queries([dist.bcast_u64(q, w2["lim"]) for q in q2], options)
        r2 = timed_steps(torch, j2, dist, flush, 20, 5)
        c2 = {"value": w2["k"] * j2.columns(w2["n"] * world) * 20 / (r2["dev_ms"] / 1e3) / 1e9, "unit": "GCUPS",
              "ms_per_step": r2["dev_ms"] / 20, "scaling": "weak", "reads_per_gpu": w2["n"],
              "workload": w2["text"]}
        j2.close()

    if rank != 0:
        job.close()
        dist.close()
        return 0

    # ---- roofline of the dominant kernels: the integer ALU pipe (LOP3), measured both ways
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    clocks = res["clocks"]
    int_peak = job.ctxs[0].measure_int_peak()
    lop3_peak = int_peak["lop3_ops_per_s"]  # lane-ops/s of a dependent-free LOP3 stream on every SM (peak_kernels.cu)
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    alu_nominal = sms * 4 * 16 * sm_max * 1e6  # ALU pipe: 16 lanes/clk/SMSP (B300_MICROARCH.md:85)
    kern_s = kern_ms / 1e3
    st = res["stats"]  # LOP3 warp instructions of the timed steps on THIS rank (x 32 = lane operations)
    if st["scans"] != 2 * args.steps:
        raise SystemExit(f"bench.py: {st['scans']} scans tallied in the timed region, expected {2 * args.steps}")
    executed = 32.0 * st["lop3_executed"] / kern_s
    planned = 32.0 * st["lop3_planned"] / kern_s
    algorithmic = ALGO_OPS_PER_COLUMN * cols_rank * args.steps / kern_s
    issue_peak = sms * 4 * 32 * sm_max * 1e6
    if executed > issue_peak:
        raise SystemExit("bench.py: implausible kernel time — the timed region did not contain the scan kernels")
    n_scans = 2 * args.steps
    traffic = profiled = None
    try:  # figures from the committed ncu captures (profiles/): not measured in this run
        prof = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")))
        traffic = prof.get(args.workload, {}).get("dram_bytes_per_launch")
        profiled = prof.get(args.workload, {})
    except Exception:
        pass
    cols_padded = ((sl + 1 + 15) // 16) * 16
    hbm_bytes_scan = ((n + 1023) // 1024) * 32 * cols_padded * 16 + 16 * q_start
    roofline = {
        "bound": "int-alu",
        "kernel": "bs_group_kernel<K,P,G> (one launch per unit shape in use) + bs_scan_kernel<K> (ungrouped k-mers); one scan "
                  "= up to 13 concurrent launches; the scans of the two read ends run side by side on two streams when both samples fit in L2",
        "achieved": executed / 1e12, "peak": lop3_peak / 1e12, "unit": "T lane-op/s", "frac": executed / lop3_peak,
        "what": "EXECUTED LOP3 lane-operations of the timed scans (5 per automaton row, text column and 32 reads; the rows "
                "skipped by dead-row skipping are tallied by the kernels at run time and NOT counted) / kernel time, against "
                "the LOP3 peak measured in this run on this GPU (apc_measure_int_peak)",
        "frac_of_nominal_alu_peak": executed / alu_nominal,
        "frac_planned": planned / lop3_peak,
        "frac_algorithmic": algorithmic / lop3_peak,
        "algorithmic_note": "SURVEY.md §8d counts 16 int ops per (k-mer, text position) column (Myers/Hyyro); bit-slicing, "
                            "row sharing between k-mers and dead-row skipping do far fewer, so this figure exceeds 1",
        "peak_source": "measured: dependent-free LOP3 stream on all SMs, this run; nominal ALU pipe = "
                       f"{sms} SM x 4 SMSP x 16 lanes/clk x {sm_max:.0f} MHz = {alu_nominal / 1e12:.2f} T lane-op/s",
        "traffic": traffic,
        "kernel_share_of_step": kern_ms / dev_ms if dev_ms else None,
        "avg_scan_ms": kern_ms / n_scans, "scans_timed": n_scans, "launches_timed": res["launches"],
        "sm_mhz_during_run": sm_mhz,
        "hbm": {"algorithmic_bytes_per_scan": hbm_bytes_scan, "achieved_gbs": hbm_bytes_scan / (kern_s / n_scans) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"), "note": "the text (bit planes, 0.5 B per base) is re-read from L2 by every "
                "unit; the kernels are bound by the ALU pipe (LOP3), HBM is idle"},
        "ncu_profiled": profiled,
    }

    # ---- N = 1 only: CPU baseline on a bounded sample + oracle spot check, the floor, the wide-offset stream
    cpu = floor = wide = ingest = None
    if world == 1 and not args.no_cpu_baseline:
        import __graft_entry__ as g
        g.build_oracle()
        leg = cpu_leg(w, job.ends, queries, target_s=args.cpu_seconds, algo="fm")
        scan = cpu_leg(w, job.ends, queries, target_s=min(4.0, args.cpu_seconds), algo="scan")
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
        # the CPU leg doubles as a check of the GPU counts: a bounded sample of both ends against the oracle
        from oracle import orc
        r = min(n, 4096)
        for e in (0, 1):
            chk = ApproxCounter(local_rank)
            chk.upload_sample(np.ascontiguousarray(job.ends[e][:r]))
            got = chk.errorCount(queries[e], k)
            chk.close()
            codes, offs = orc.encode_matrix(job.ends[e][:r])
            if not np.array_equal(got, orc.fm_index_error_count(codes, offs, queries[e], k)):
                raise SystemExit("bench.py: PARITY FAILURE — GPU counts differ from the oracle on the check sample")
        parity = f"ok: GPU counts == oracle (index-based CPU form) on the first {r} reads of both ends x all {q_start}+{q_end} k-mers"
    if world == 1 and not args.no_extras:
        # the floor: the same sample scanned for as many RANDOM k-mers with one k-mer per warp — no rows shared between
        # k-mers, and unrelated k-mers leave little to skip
        rng = np.random.default_rng(w["seed"])
        rq = [np.unique(rng.integers(0, 1 << 62, 2 * len(q)).astype(np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else (1 << 64) - 1))[: len(q)]
              for q in queries]
        job.set_queries(rq, options + [("shape_mask", 0)])
        fr = timed_steps(torch, job, dist, flush, 3, 3)
        floor = {"value": k * job.columns() * 3 / (fr["dev_ms"] / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": fr["dev_ms"] / 3,
                 "executed_frac_of_lop3_peak": 32.0 * fr["stats"]["lop3_executed"] / (fr["kern_ms"] / 1e3) / lop3_peak,
                 "what": "same sample, as many random k-mers, shape_mask=0 (one k-mer per warp): no row sharing"}
        job.close()
        # the same workload with the adapters at offsets uniform in 0..sl/2 instead of 0..7 (seed bit 40)
        jw = Job(torch, dev, local_rank, stream, w, 0, n, seed=w["seed"] | (1 << 40))
        jw.set_queries(jw.exact_queries(lim)[0], options)
        wr = timed_steps(torch, jw, dist, flush, 5, 3)
        wide = {"value": k * jw.columns() * 5 / (wr["dev_ms"] / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": wr["dev_ms"] / 5,
                "executed_frac_of_lop3_peak": 32.0 * wr["stats"]["lop3_executed"] / (wr["kern_ms"] / 1e3) / lop3_peak,
                "executed_over_planned": wr["stats"]["lop3_executed"] / max(wr["stats"]["lop3_planned"], 1.0),
                "what": "same generator with the adapter offsets uniform in 0..sl/2 (default 0..7), own top-lim queries"}
        jw.close()
        try:
            ingest = ingest_leg(w, args.workload, local_rank)
        except SystemExit:
            raise
        except Exception as e:  # noqa: BLE001 — the leg is an extra: a full /tmp must not cost the bench line
            ingest = {"error": repr(e)}
    else:
        job.close()

    line = {
        "metric": "approx_count_gcups", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": make_config(w, args.scaling, world),
        "reads_per_gpu": n, "queries": [q_start, q_end],
        "parallelism": f"reads sharded over {world} GPU(s); ONE ncclAllReduce of {q_start + q_end} u64 per step through the C ABI "
                       "(apc_allreduce_counts)" if world > 1 else "1 GPU",
        "queries_per_s": (q_start + q_end) * args.steps / (dev_ms / 1e3),
        "columns_per_s": cols_job * args.steps / (dev_ms / 1e3),
        "wall_s_timed_region": res["t1"] - res["t0"],
        "exact_stage_ms": exact_ms,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "frac_of_resident": e2e_value / value,
                "path": "per step: apc_upload_sample_async + apc_set_queries + apc_scan per end (C ABI, pinned host buffers, one "
                        "stream per end), then apc_allreduce_counts + D2H of the counts on a third stream; steps pipelined two "
                        "deep (alternating output buffers); wall clock"},
        "gpu_launches": res["launches"],
        "parity_check": parity,
        "roofline": roofline,
        # flat copies of the figures the review asked to see in the parsed line
        "roofline_frac": roofline["frac"], "roofline_frac_planned": roofline["frac_planned"],
        "roofline_frac_algorithmic": roofline["frac_algorithmic"],
        "lop3_executed_per_step": 32.0 * st["lop3_executed"] / args.steps, "lop3_planned_per_step": 32.0 * st["lop3_planned"] / args.steps,
        "lop3_one_kmer_per_warp_per_step": 32.0 * st["lop3_one_kmer_per_warp"] / args.steps,
        "lop3_top_share_of_executed": st["lop3_top"] / max(st["lop3_executed"], 1.0),
        "lop3_peak_measured_tops": lop3_peak / 1e12,
        "floor_value": floor["value"] if floor else None, "floor": floor,
        "wide_offset_value": wide["value"] if wide else None, "wide_offset": wide,
        "c2_weak_value": c2["value"] if c2 else None, "c2_weak": c2,
        "ingest_device_ms": ingest.get("device_ms") if ingest else None,
        "ingest_host_ms": ingest.get("host_ms") if ingest else None,
        "pipeline_from_file_ms": (ingest.get("pipeline_from_file") or {}).get("total_ms") if ingest else None, "ingest": ingest,
        "cpu_baseline": cpu,
    }
    emit(line)
    dist.close()
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
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="C3")
    ap.add_argument("--scan-variant", type=int, default=0,
                    help="kernel A/B: 0 bit-sliced with k-mer pairing (default), 7 bit-sliced, 8 row-packed")
    ap.add_argument("--reads", type=int, default=0, help="override the reads of the workload (C5 sweep)")
    ap.add_argument("--lim", type=int, default=0, help="override the number of query k-mers (C5 sweep)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong",
                    help="strong (default): one n-read job split over the ranks; weak: every rank holds its own n-read shard")
    ap.add_argument("--plan-alive-pct", type=int, default=-1,
                    help="planner knob (apc_set_option plan_alive_pct): expected share of columns with live deep rows")
    ap.add_argument("--shape-mask", type=lambda x: int(x, 0), default=-1,
                    help="tuning knob (apc_set_option shape_mask): bit s = unit shape s may be used")
    ap.add_argument("--sg-per-job", type=int, default=0,
                    help="tuning knob (apc_set_option tiles_per_job): 1024-read super-groups per job, 0 = auto")
    ap.add_argument("--graph", action="store_true", help="replay repeated scans as a CUDA graph (apc_set_option scan_graph 1)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample size, seconds of CPU work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the floor / wide-offset / C2-weak side measurements")
    args = ap.parse_args()
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the host-side set-up (synthetic reads, the oracle) is OpenMP
    # code: the reference arm (rank 0 alone works) gets every CPU, a B200 rank its share — set before any OpenMP
    # runtime is loaded
    if os.environ.get("WORLD_SIZE") and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        share = 1 if args.impl == "reference" else max(1, int(os.environ["WORLD_SIZE"]))
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // share))
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
