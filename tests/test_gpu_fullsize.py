"""Full-size, file-level parity on the BASELINE configurations (north_star: "bit-exact approximate and
exact k-mer count files on all 5 configs"): the drop-in `approx_counter` binary reads the synthetic
FASTA/FASTQ of a configuration at its full size and its four output files are compared byte for byte
with the oracle pipeline — reference main loop :858-952: sample (:867), count_kmers (:874),
get_most_frequent (:898), exportCounter (:910), errorCount (:922), get_most_frequent (:923),
exportCounter (:928).

The oracle's approximate count here is the INDEX-based restatement (oracle/fm_index_model.cpp: own
bidirectional FM index + the search-scheme recursion of SeqAn's find<0,2>, :586), the only CPU form fast
enough at these sizes; the sampled text comes from the oracle's own generator (oracle/synth_reads.c), the
binary reads the same reads from the file written by the product's generator.

C1 and C2 always run (seconds).  C3, C4 and the C5 point (1M reads x 5000 k-mers) take minutes of CPU
time for the oracle and run only with APC_RUN_SLOW=1 (tools/fullsize_slow.sh runs them on the GPU box;
its log is committed under profiles/)."""
import os
import subprocess
import time

import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")

# name: reads, sl, k, lim, seed, fastq (BASELINE.json configs; seeds of SURVEY.md §8d)
CONFIGS = {
    "C1": (10_000, 100, 16, 500, 1001, False),
    "C2": (100_000, 100, 16, 2000, 1002, False),
    "C3": (1_000_000, 150, 20, 5000, 1003, True),
    "C4": (1_000_000, 200, 32, 10_000, 1004, False),
    "C5": (1_000_000, 100, 16, 5000, 2003, False),   # one point of the sweep: -sn 1M -lim 5000
}
SLOW = {"C3", "C4", "C5"}
# the binary parses and samples its input on the GPU by default (--ingest device, apc_ingest_fastx); C2 and C3 (FASTQ)
# are also run with the host parser and sampler
_WANT = {}
CASES = [(name, "device") for name in CONFIGS] + [("C2", "host"), ("C3", "host")]


def oracle_end_files(sample, k, lim, tmp, which, threads):
    codes, offs = orc.encode_matrix(sample)
    thr = orc.adjust_threshold(1.0, 16, k)
    # the plain restatement of count_kmers (:487-519) up to C2; its multi-threaded form (same multiset,
    # tests/test_oracle.py) where the single-threaded one takes minutes per end
    keys, cnts, _ = orc.count_kmers_mt(codes, offs, k, thr) if threads else orc.count_kmers(codes, offs, k, thr)
    tk, tc = orc.get_most_frequent_fast(keys, cnts, lim, k)   # == get_most_frequent (tests/test_oracle.py)
    del keys, cnts
    e, o = tmp / f"want_exact.{which}", tmp / f"want_out.{which}"
    orc.export_counter(tk, tc, k, e)
    approx = orc.fm_index_error_count(codes, offs, tk, k)
    ak, ac = orc.get_most_frequent(tk, approx, lim, k)
    orc.export_counter(ak, ac, k, o)
    return e.read_bytes(), o.read_bytes()


@pytest.mark.parametrize("name,ingest", CASES)
def test_files_match_oracle_at_full_size(built, tmp_path, name, ingest):
    if name in SLOW and os.environ.get("APC_RUN_SLOW") != "1":
        pytest.skip(f"{name} at full size needs minutes of CPU time for the oracle: set APC_RUN_SLOW=1 "
                    "(tools/fullsize_slow.sh; last log in profiles/)")
    from approx_counter_b200 import host
    n, sl, k, lim, seed, fastq = CONFIGS[name]
    path = tmp_path / ("reads.fq" if fastq else "reads.fa")
    host.synth_write(path, seed, n, sl, fastq=fastq)
    t0 = time.perf_counter()
    p = subprocess.run([BIN, "-k", str(k), "-sn", str(n), "-sl", str(sl), "-lim", str(lim), "-lc", "1.0",
                        "-e", str(tmp_path / "exact.txt"), "-o", str(tmp_path / "out.txt"), "-nt", "4", "-v", "2",
                        "--ingest", ingest, str(path)], capture_output=True, text=True, timeout=1800)
    t_bin = time.perf_counter() - t0
    assert p.returncode == 0, p.stderr
    assert ("indexed on the GPU" in p.stdout) == (ingest == "device")
    os.unlink(path)
    t0 = time.perf_counter()
    for which, bot in (("start", False), ("end", True)):
        if (name, which) not in _WANT:   # the oracle's files of a configuration are computed once (minutes for C3-C5)
            _WANT[name, which] = oracle_end_files(orc.synth_ends(seed, 0, n, sl, bot), k, lim, tmp_path, which, name in SLOW)
        want_exact, want_out = _WANT[name, which]
        got_exact = (tmp_path / f"exact.txt_0.{which}").read_bytes()      # `_<run>` always appended (:837)
        got_out = (tmp_path / f"out.txt_0.{which}").read_bytes()
        assert got_exact == want_exact, f"{name} {which}: exact top-{lim} file differs"
        assert got_out == want_out, f"{name} {which}: approximate count file differs"
        assert got_out.count(b"\n") == lim
    print(f"\n[fullsize] {name} (--ingest {ingest}): n={n} sl={sl} k={k} lim={lim} fastq={fastq}: 4 files byte-identical; "
          f"binary {t_bin:.2f} s wall, oracle {time.perf_counter() - t0:.1f} s")
