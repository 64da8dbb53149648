#!/usr/bin/env python
"""Host ingest (parallel parser + sampler + upload) against device ingest (apc_ingest_fastx +
apc_sample_resident) on the synthetic input of a BASELINE configuration: wall clock of each phase, the
device times of the ingest kernels, and a byte-for-byte check that both routes leave the same sample.
  python tools/ingest_bench.py C3 [--reps 3]   -> one JSON line per configuration"""
import argparse
import json
import mmap
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approx_counter_b200 import ApproxCounter, host  # noqa: E402

CONFIGS = {"C1": (10_000, 100, 1001, False), "C2": (100_000, 100, 1002, False), "C3": (1_000_000, 150, 1003, True),
           "C4": (1_000_000, 200, 1004, False)}


def wall(f):
    t0 = time.perf_counter()
    r = f()
    return r, (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C2"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--staging", type=int, default=1, help="apc_set_option ingest_staging")
    a = ap.parse_args()
    c = ApproxCounter(0)
    c.set_option("ingest_staging", a.staging)
    for name in a.configs:
        n, sl, seed, fastq = CONFIGS[name]
        path = os.path.join(tempfile.gettempdir(), f"ingest_{name}.{'fq' if fastq else 'fa'}")
        host.synth_write(path, seed, n, sl, fastq=fastq)
        size = os.path.getsize(path)
        best = {}
        same = True
        for rep in range(a.reps):
            row = {}
            reads, row["host_parse_ms"] = wall(lambda: host.Reads(path))
            hs = {}
            t_s = t_u = 0.0
            for bot in (False, True):
                s, t = wall(lambda: reads.sample(n, sl, bot, 7))
                t_s += t
                _, t = wall(lambda: c.upload_sample(s))
                t_u += t
                hs[bot] = s
            row["host_sample_ms"], row["host_upload_ms"] = t_s, t_u
            row["host_total_ms"] = row["host_parse_ms"] + t_s + t_u
            reads.close()
            with open(path, "rb") as f:
                mm, row["dev_map_ms"] = wall(lambda: mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ))
                buf = np.frombuffer(mm, np.uint8)
                (nrec, isfq), row["dev_ingest_ms"] = wall(lambda: c.ingest_fastx(buf))
                tm = c.ingest_timing()
                row["dev_copy_ms"], row["dev_index_ms"] = tm["copy_ms"], tm["index_ms"]
                del buf
                mm.close()
            assert nrec == n and isfq == fastq
            order, row["dev_shuffle_ms"] = wall(lambda: host.shuffle_order(n, 7))
            t_s = t_dev = 0.0
            for bot in (False, True):
                got_n, t = wall(lambda: c.sample_resident(n, sl, bot, order))
                t_s += t
                t_dev += c.ingest_timing()["sample_ms"]
                if rep == 0:
                    same = same and got_n == len(hs[bot]) and np.array_equal(c.download_sample(), hs[bot])
            row["dev_sample_ms"], row["dev_sample_kernels_ms"] = t_s, t_dev
            row["dev_total_ms"] = row["dev_map_ms"] + row["dev_ingest_ms"] + 2 * row["dev_shuffle_ms"] + t_s
            for k2, v in row.items():
                best[k2] = min(best.get(k2, v), v)
        os.unlink(path)
        best = {k2: round(v, 3) for k2, v in best.items()}
        print(json.dumps({"config": name, "reads": n, "sl": sl, "fastq": fastq, "file_mb": round(size / 1e6, 1),
                          "samples_identical": bool(same), "reps": a.reps, "staging": a.staging,
                          "index_gbs_over_file": round(size / 1e6 / max(best["dev_index_ms"], 1e-6), 1),
                          "copy_gbs": round(size / 1e6 / max(best["dev_copy_ms"], 1e-6), 2), **best}), flush=True)
    c.close()


if __name__ == "__main__":
    main()
