"""Cost of one unit of every shape of the bit-sliced kernel (bs_group_kernel<K, K-T, G>), forwards and
backwards, against one k-mer per warp: calibrates bs_unit_cost() in csrc/bitslice_kernel.cu.

    python tools/shape_bench.py [k ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from approx_counter_b200 import ApproxCounter, plan_queries


def rev(x, k):
    r = 0
    for i in range(k):
        r = (r << 2) | ((x >> (2 * i)) & 3)
    return r


def time_scan(c, kmers, k, reps=4):
    c.set_queries(np.array(kmers, np.uint64), k)
    best = 1e9
    for _ in range(reps):
        c.scan()
        c.sync()
        best = min(best, c.timing()["scan_ms"])
    return best


if __name__ == "__main__":
    ks = [int(a) for a in sys.argv[1:]] or [16, 20, 32]
    rng = np.random.default_rng(3)
    n, L = 200_000, 100
    with ApproxCounter(0) as c:
        c.upload_sample(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, L)))
        for k in ks:
            shapes = plan_queries([0], k)
            singles = [int(x) & ((1 << (2 * k)) - 1) for x in rng.integers(0, 1 << 62, 1200)] if k < 32 else \
                      [int.from_bytes(rng.bytes(8), "little") for _ in range(1200)]
            c.set_option("shape_mask", 0)
            t1 = time_scan(c, singles, k) / len(singles)
            print(json.dumps({"k": k, "shape": "single", "rows": k, "us_per_unit": round(t1 * 1e3, 3)}), flush=True)
            for s in range(len(shapes["units"])):
                t, g = int(shapes["shape_t"][s]), int(shapes["shape_g"][s])
                if g == 0:
                    continue
                n_units = max(64, 4800 // (k - t + g * t) * 4)
                for backwards in (0, 1):
                    kmers = []
                    for u in range(n_units):
                        head = int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * (k - t))) - 1)
                        tails = rng.choice(1 << (2 * t), size=g, replace=(1 << (2 * t)) < g)
                        for tl in tails:
                            v = (head << (2 * t)) | int(tl)
                            kmers.append(rev(v, k) if backwards else v)
                    c.set_option("shape_mask", 1 << s)
                    plan = plan_queries(kmers, k)
                    tu = time_scan(c, kmers, k)
                    units = int(c.timing()["scan_launches"])
                    rows = k - t + g * t
                    print(json.dumps({"k": k, "shape": s, "t": t, "g": g, "rows": rows, "backwards": backwards,
                                      "units": n_units, "launches": units,
                                      "us_per_unit": round(tu / n_units * 1e3, 3),
                                      "row_equiv": round(tu / n_units / t1 * k, 2),
                                      "speedup_vs_singles": round(g * t1 / (tu / n_units), 2)}), flush=True)
            c.set_option("shape_mask", 0xFFFFFFFF)
