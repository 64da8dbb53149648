"""Scratch timing of the scan kernels on random data (not the contract bench).

    python tools/quick_bench.py            bit-sliced vs best row-packed kernel over a range of k
    python tools/quick_bench.py micro      integer-pipe microbenchmarks only (apc_microbench)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from approx_counter_b200 import ApproxCounter

MICRO = ("lop3", "imad", "mixed", "imad_hi", "imad_wide", "lop3_3reg", "blend_3reg", "blend_2reg", "blend_rrr",
         "core_1_256_2", "core_1_256_3", "core_1_256_4", "core_1_128_9", "core_2_256_1", "core_2_256_2",
         "core_2_128_5", "core_3_128_3", "core64_1_256_3")


def run(c, n, L, k, q, variant=0, tpj=0, reps=5):
    """variant: 0 = bit-sliced (default kernel), 8 = best row-packed, 1/2/3/6 = a given packing."""
    rng = np.random.default_rng(1)
    sample = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, L))
    kmers = rng.integers(0, 1 << 62, q).astype(np.uint64) & np.uint64((1 << (2 * k)) - 1)
    c.set_option("scan_variant", variant)
    c.set_option("tiles_per_job", tpj)
    c.upload_sample(sample)
    c.set_queries(kmers, k)
    best = 1e9
    for _ in range(reps):
        c.scan()
        c.sync()
        best = min(best, c.timing()["scan_ms"])
    cols = q * n * L
    return {"n": n, "L": L, "k": k, "q": q, "variant": variant, "tpj": tpj, "ms": round(best, 4),
            "Tcol_s": round(cols / best / 1e9, 3), "GCUPS": round(k * cols / best / 1e6, 1)}


if __name__ == "__main__":
    with ApproxCounter(0) as c:
        print(json.dumps(c.measure_int_peak()))
        for name in MICRO:
            print(name, "%.4g" % c.microbench(name), flush=True)
        if len(sys.argv) > 1 and sys.argv[1] == "micro":
            sys.exit(0)
        for args in [(100000, 100, 16, 2000), (100000, 101, 16, 2000), (10000, 100, 16, 500), (200000, 150, 20, 2000),
                     (200000, 200, 32, 2000), (100000, 100, 10, 2000), (100000, 100, 13, 2000), (200000, 200, 25, 2000),
                     (1000000, 150, 20, 5000)]:
            for variant in (0, 8):
                print(json.dumps(run(c, *args, variant=variant, reps=3)), flush=True)
        if len(sys.argv) > 1 and sys.argv[1] == "pairs":
            sys.exit(0)
        for tpj in (1, 2, 4, 8):
            print(json.dumps(run(c, 100000, 100, 16, 2000, 0, tpj)), flush=True)
