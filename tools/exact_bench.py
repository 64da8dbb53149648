"""Steady-state timing of the exact stage (K3-K5) on a synthetic workload; with `check` also
verifies the window bookkeeping (sum of counts + N windows == all windows, filter off)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from approx_counter_b200 import ApproxCounter, host

def main():
    n = int(sys.argv[1]); sl = int(sys.argv[2]); k = int(sys.argv[3]); lim = int(sys.argv[4])
    check = len(sys.argv) > 5 and sys.argv[5] == "check"
    t = time.time(); sample = host.synth_ends(2000 + k, 0, n, sl, True); gen = time.time() - t
    thr = host.adjust_threshold(1.0, 16, k)
    with ApproxCounter(0) as c:
        c.upload_sample(sample)
        up = c.timing()["upload_ms"]
        ms = []
        for _ in range(4):
            km, ct, nd, hn = c.count_kmers_topn(k, thr, lim)
            ms.append(round(c.timing()["exact_ms"], 3))
        out = {"n": n, "L": sl + 1, "k": k, "lim": lim, "windows": n * (sl + 2 - k), "distinct": nd, "had_n": hn,
               "exact_ms": ms, "upload_ms": round(up, 3), "gen_s": round(gen, 1), "top": [int(x) for x in ct[:3]],
               "launches": c.timing()["exact_launches"]}
        if check:
            km2, ct2, nd2, hn2 = c.count_kmers_topn(k, 1e9, 10)          # filter off
            tot = c.solid_kmers(k, 1e9, 1, capacity=8)                    # capacity too small -> retried with n_out
            out["check_sum"] = int(tot[1].sum()) + hn2 == n * (sl + 2 - k)
            out["check_distinct"] = len(tot[0]) == nd2
        print(json.dumps(out), flush=True)

main()
