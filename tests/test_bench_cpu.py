"""bench.py's CPU side (no GPU): the reference arm prints the driver's JSON contract, and the two CPU
restatements it can time agree with each other."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--reads", "2500", "--lim", "50", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, APC_LIB_PATH="/nonexistent/libapc.so"))  # loading the product would raise
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                    # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "approx_count_gcups" and d["unit"] == "GCUPS"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["algo"] == "fm-index" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["index_build_s_per_step"] > 0 and cb["search_s_per_step"] > 0 and "sample" in cb
    assert d["config"]["k"] == 16 and "workload" in d["config"]
    # both arms print the same `config` object (the driver compares them) ...
    sys.path.insert(0, ROOT)
    import bench
    w = dict(bench.WORKLOADS["C1"], n=2500, lim=50)
    w["text"] = f"C5 sweep point on C1: -sn {w['n']} -sl {w['sl']} -lim {w['lim']}, k={w['k']}"
    assert d["config"] == bench.make_config(w, "strong", 1)
    # ... and the reference arm never loads the product library (APC_LIB_PATH above points nowhere)


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_legs_agree(built):
    sys.path.insert(0, ROOT)
    import bench
    from oracle import orc
    w = dict(bench.WORKLOADS["C1"], n=1500, lim=40)
    ends = (orc.synth_ends(w["seed"], 0, w["n"], w["sl"], False), orc.synth_ends(w["seed"], 0, w["n"], w["sl"], True))
    queries = bench.reference_queries(w, ends)
    assert all(len(q) == 40 for q in queries)
    for sample, km in zip(ends, queries):
        codes, offs = orc.encode_matrix(sample)
        assert np.array_equal(orc.fm_index_error_count(codes, offs, km, w["k"]),
                              orc.error_count(codes, offs, km, w["k"], fast=True))
    t, cols, build, search = bench.cpu_run("fm", ends, queries, w["k"], 700, 2)
    assert t > 0 and build > 0 and search > 0 and cols == 40 * 700 * (100 + 101)
    t, cols, build, search = bench.cpu_run("scan", ends, queries, w["k"], 700, 2)
    assert t > 0 and build == 0 and search == 0 and cols == 40 * 700 * (100 + 101)
