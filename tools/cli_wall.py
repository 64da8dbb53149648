"""Wall-clock breakdown of the one-shot drop-in binary on the BASELINE configurations (phase timestamps of -v 2).

    python tools/cli_wall.py [--ingest device] [C1 C2 C3 C4] > gpurun_out/r02_cli_wall.md

Every configuration is run twice (the second run has the input file in the page cache and the CUDA driver warm);
the table reports the second run: process wall clock, then where it went."""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approx_counter_b200 import host
import bench

BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")
FASTQ = {"C3"}
os.makedirs("/tmp/apc_cli", exist_ok=True)


EXTRA = []
if "--ingest" in sys.argv:
    i = sys.argv.index("--ingest")
    EXTRA = ["--ingest", sys.argv[i + 1]]
    del sys.argv[i:i + 2]


def phases(log):
    """[(ms, text)] of the `[x ms]` lines."""
    out = []
    for ln in log.splitlines():
        m = re.match(r"\[([0-9.e+]+) ms\]\s*(.*)", ln)
        if m:
            out.append((float(m.group(1)), m.group(2).strip()))
    return out


def breakdown(ph):
    t = {k: 0.0 for k in ("parse", "wait for CUDA context", "device ingest", "sample", "upload", "exact count + top-N", "approximate count", "export", "release")}
    def at(i):
        return ph[i][0]
    for i, (ms, text) in enumerate(ph[:-1]):
        dt = at(i + 1) - ms
        if text.startswith("Parsing FASTA"):
            t["parse"] += dt
        elif text.startswith("File parsed") or text.startswith("File mapped"):
            t["wait for CUDA context"] += dt
        elif text.startswith("CUDA context ready") and "copying the input" in text:
            t["device ingest"] += dt
        elif text.startswith("Sampling"):
            t["sample"] += dt
        elif text.startswith("Sampled"):
            t["upload"] += dt
        elif text.startswith("Exact k-mer count") or text.startswith("Number of kmer") or text.startswith("Keeping"):
            t["exact count + top-N"] += dt
        elif text.startswith("Exporting exact") or text.startswith("Exporting approximate"):
            t["export"] += dt
        elif text.startswith("Approximate k-mer count"):
            t["approximate count"] += dt
        elif text.startswith("Releasing"):
            t["release"] += dt
    return t


rows = []
for name in (sys.argv[1:] or ["C1", "C2", "C3", "C4"]):
    w = bench.WORKLOADS[name]
    path = f"/tmp/apc_cli/{name}." + ("fq" if name in FASTQ else "fa")
    t0 = time.time()
    host.synth_write(path, w["seed"], w["n"], w["sl"], fastq=name in FASTQ)
    gen = time.time() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        p = subprocess.run([BIN, "-k", str(w["k"]), "-sn", str(w["n"]), "-sl", str(w["sl"]), "-lim", str(w["lim"]), "-v", "2",
                            "-e", f"/tmp/apc_cli/{name}_exact", "-o", f"/tmp/apc_cli/{name}_out", *EXTRA, path],
                           capture_output=True, text=True)
        wall = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr
    ph = phases(p.stdout)
    ctx = [x for _, x in ph if x.startswith("CUDA context ready")]
    rows.append((name, os.path.getsize(path) / 1e6, wall, ph[-1][0] / 1e3, breakdown(ph), ctx[0] if ctx else ""))
    os.unlink(path)

keys = list(rows[0][4])
print("| config | input MB | process wall s | main() s | " + " | ".join(keys) + " |")
print("|---|---|---|---|" + "---|" * len(keys))
for name, mb, wall, main_s, t, ctx in rows:
    print(f"| {name} | {mb:.0f} | {wall:.3f} | {main_s:.3f} | " + " | ".join(f"{t[k] / 1e3:.3f}" for k in keys) + " |")
print()
for name, mb, wall, main_s, t, ctx in rows:
    rest = sum(v for k, v in t.items() if k != "wait for CUDA context") / 1e3
    print(f"* {name}: everything but the wait for the CUDA context: {rest:.3f} s; {ctx}")
print(f"\n(extra arguments: {' '.join(EXTRA) or 'none'})")
print("\n(seconds; both ends summed; second of two runs; `main() s` = last timestamp of the -v 2 log, the rest of the process"
      " wall clock is program start and exit.)")
