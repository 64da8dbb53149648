"""A/B of the plane prefetch depth without torch: one C2 end, best of 5 scans per library.
    python tools/pf_ab.py [lib ...]      (paths of libapc variants; default: the in-tree build)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
from approx_counter_b200 import ApproxCounter, host
n, sl, k, lim = 100000, 100, 16, 2000
s = host.synth_ends(1002, 0, n, sl, False)
with ApproxCounter(0) as c:
    c.upload_sample(s)
    km, ct, _, _ = c.count_kmers_topn(k, host.adjust_threshold(1.0, 16, k), lim)
    c.set_queries(km, k)
    best = 1e9
    for _ in range(6):
        c.scan(); c.sync()
        best = min(best, c.timing()["scan_ms"])
    print(json.dumps({"lib": os.environ.get("APC_LIB_PATH", "default"), "scan_ms": round(best, 4),
                      "kGCUPS": round(k * len(km) * n * sl / best / 1e9, 1), "checksum": int(c.get_counts().sum())}))
''' % ROOT

for lib in (sys.argv[1:] or [""]):
    env = dict(os.environ)
    if lib:
        env["APC_LIB_PATH"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
