"""End-to-end GPU parity: the drop-in `approx_counter` binary (same flags and output
files as the reference, :604-669 / :835-955) against the oracle pipeline, byte for
byte, plus the committed golden fixtures through the Python mirror of the API."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def oracle_files(reads, k, sl, lim, param_lc, tmp, solid=0, forbidden=None):
    """What the reference writes when every read is sampled (sn >= #reads, :844-848)."""
    codes, offs = orc.encode(reads)
    thr = orc.adjust_threshold(param_lc, 16, k)
    perm = np.arange(len(reads), dtype=np.uint64)
    out = {}
    for which, bot in (("start", False), ("end", True)):
        sc, so = orc.sample_sequences(codes, offs, perm, len(reads), sl, bot)
        keys, cnts, _ = orc.count_kmers(sc, so, k, thr, forbidden)
        if solid:
            tk, tc = orc.get_solid_kmers(keys, cnts, solid, k)
        else:
            tk, tc = orc.get_most_frequent(keys, cnts, lim, k)
        ek = tmp / f"want_exact.{which}"
        orc.export_counter(tk, tc, k, ek)
        approx = orc.error_count(sc, so, tk, k, fast=True)
        ak, ac = orc.get_most_frequent(tk, approx, lim, k)
        ok = tmp / f"want_out.{which}"
        orc.export_counter(ak, ac, k, ok)
        out[which] = (ek.read_bytes(), ok.read_bytes())
    return out


@pytest.mark.parametrize("fastq,k,sl,lim,n", [(False, 16, 100, 200, 3000), (True, 20, 150, 300, 1500),
                                              (False, 32, 200, 100, 800), (False, 10, 40, 64, 2000)])
def test_binary_matches_oracle_files(built, tmp_path, fastq, k, sl, lim, n):
    from approx_counter_b200 import host
    path = tmp_path / ("reads.fq" if fastq else "reads.fa")
    host.synth_write(path, 4242 + k, n, sl, fastq=fastq)
    r = host.Reads(path)
    reads = [r.seq(i) for i in range(len(r))]
    assert len(reads) == n
    out, exact = tmp_path / "out.txt", tmp_path / "exact.txt"
    want = oracle_files(reads, k, sl, lim, 1.0, tmp_path)
    # the input parsed and sampled on the GPU (the default) and by the host threads
    for ingest in ("host", "device"):
        p = subprocess.run([BIN, "-k", str(k), "-sn", str(n), "-sl", str(sl), "-lim", str(lim), "-lc", "1.0",
                            "-e", str(exact), "-o", str(out), "-nt", "4", "-v", "2", "--ingest", ingest, str(path)],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        assert ("indexed on the GPU" in p.stdout) == (ingest == "device")
        for which in ("start", "end"):
            assert (tmp_path / f"exact.txt_0.{which}").read_bytes() == want[which][0]   # `_<run>` always appended (:837)
            assert (tmp_path / f"out.txt_0.{which}").read_bytes() == want[which][1]
    lines = (tmp_path / "out.txt_0.start").read_text().splitlines()
    assert len(lines) == lim and all(len(x.split("\t")[0]) == k for x in lines)
    assert "Kmer size:" in p.stdout and "Approximate k-mer count" in p.stdout


def test_binary_skip_end_forbidden_and_solid(built, tmp_path):
    from approx_counter_b200 import host
    k, sl, n, lim = 16, 100, 1500, 100
    path = tmp_path / "reads.fa"
    host.synth_write(path, 99, n, sl)
    r = host.Reads(path)
    reads = [r.seq(i) for i in range(n)]
    fk = tmp_path / "forbidden.txt"
    forb = ["AATGTACTTCGTTCAG", "ATGTACTTCGTTCAGT"]
    fk.write_text("\n".join(forb) + "\n")
    out = tmp_path / "o"
    p = subprocess.run([BIN, "-k", str(k), "-sn", str(n), "-sl", str(sl), "-lim", str(lim), "-fk", str(fk),
                        "-se", "-e", str(tmp_path / "e"), "-o", str(out), str(path)], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr
    want = oracle_files(reads, k, sl, lim, 1.0, tmp_path, forbidden=[orc.dna2int(s) for s in forb])
    assert (tmp_path / "o_0.start").read_bytes() == want["start"][1]
    assert (tmp_path / "e_0.start").read_bytes() == want["start"][0]
    assert not (tmp_path / "o_0.end").exists()                       # -se (:943-951)
    p = subprocess.run([BIN, "-k", str(k), "-sn", str(n), "-sl", str(sl), "-lim", "100000", "-sk", "50",
                        "-e", str(tmp_path / "se"), "-o", str(tmp_path / "so"), str(path)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    want = oracle_files(reads, k, sl, 100000, 1.0, tmp_path, solid=50)
    for which in ("start", "end"):
        assert (tmp_path / f"se_0.{which}").read_bytes() == want[which][0]
        assert (tmp_path / f"so_0.{which}").read_bytes() == want[which][1]


def test_binary_config_file_and_multi_run(built, tmp_path):
    from approx_counter_b200 import host
    path = tmp_path / "reads.fa"
    host.synth_write(path, 7, 600, 60)
    conf = tmp_path / "c.conf"
    conf.write_text("# comment\nk = 12\nsl=60\nsn=600\nlim=50\nmr=2\n")
    p = subprocess.run([BIN, "-conf", str(conf), "-lim", "40", "-o", str(tmp_path / "m"), str(path)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    for run in (0, 1):
        for which in ("start", "end"):
            lines = (tmp_path / f"m_{run}.{which}").read_text().splitlines()
            assert len(lines) == 40 and all(len(x.split("\t")[0]) == 12 for x in lines)   # CLI overrides config
    assert (tmp_path / "m_0.start").read_bytes() == (tmp_path / "m_1.start").read_bytes()  # whole set sampled


@pytest.mark.parametrize("case", GOLD["approx"], ids=lambda c: f"k{c['k']}n{len(c['reads'])}")
def test_golden_approx_vectors(counter, case):
    counter.upload_sample(case["reads"])
    got = counter.errorCount([orc.dna2int(s) for s in case["kmers"]], case["k"])
    assert got.tolist() == case["counts"]


@pytest.mark.parametrize("case", GOLD["pipeline"], ids=lambda c: f"k{c['k']}")
def test_golden_pipeline_vectors(counter, case):
    from approx_counter_b200 import host
    k, lim = case["k"], case["lim"]
    counter.upload_sample(case["reads"])
    thr = host.adjust_threshold(case["param_lc"], 16, k)
    km, ct, nd, hn = counter.count_kmers_topn(k, thr, lim)
    assert nd == case["n_distinct"] and hn == case["had_n"]
    assert [[host.int2dna(a, k), int(b)] for a, b in zip(km, ct)] == case["exact"]
    approx = counter.errorCount(km, k)
    ak, ac = host.get_most_frequent(km, approx, lim, k)
    assert [[host.int2dna(a, k), int(b)] for a, b in zip(ak, ac)] == case["approx"]


def test_sharded_counter_single_rank(built):
    """ShardedApproxCounter with world size 1 (no process group): torch stream + torch count tensor."""
    import torch
    from approx_counter_b200 import ShardedApproxCounter
    rng = np.random.default_rng(3)
    sample = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(1000, 100))
    sample[::2, 5:21] = np.frombuffer(b"AATGTACTTCGTTCAG", np.uint8)
    kmers = np.array([orc.dna2int("AATGTACTTCGTTCAG"), orc.dna2int("AATGTACTTCGTTCAT"), 12345], np.uint64)
    s = ShardedApproxCounter(device=0)
    try:
        with torch.cuda.stream(torch.cuda.Stream(0)):
            lo, hi = s.upload_sample(sample)
            assert (lo, hi) == (0, 1000)
            got = s.errorCount(kmers, 16)
        got_default = s.errorCount(kmers, 16)   # legacy default stream
    finally:
        s.close()
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, 16, fast=True)
    assert np.array_equal(got, want) and np.array_equal(got_default, want)


def test_planted_adapter_is_recovered(built, tmp_path):
    """Downstream sanity: a Porechop_ABI-style greedy assembly of the top
    approximate k-mers rebuilds the adapters planted in the synthetic reads."""
    from approx_counter_b200 import host
    k, sl, n = 16, 100, 4000
    path = tmp_path / "reads.fa"
    host.synth_write(path, 2024, n, sl)
    out = tmp_path / "out.txt"
    p = subprocess.run([BIN, "-k", str(k), "-sn", str(n), "-sl", str(sl), "-lim", "300", "-o", str(out), "-v", "0",
                        str(path)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    for which, adapter in (("start", "AATGTACTTCGTTCAGTTACGTATTGCT"), ("end", "GCAATACGTAACTGAACGAAGT")):
        rows = [ln.split("\t") for ln in (tmp_path / f"out.txt_0.{which}").read_text().splitlines()]
        counts = {a: int(b) for a, b in rows}
        adapter_kmers = {adapter[i:i + k] for i in range(len(adapter) - k + 1)}
        top = [a for a, _ in rows[: len(adapter_kmers)]]
        assert len(set(top) & adapter_kmers) >= 0.8 * len(adapter_kmers), (which, top)
        # greedy extension from the best k-mer, both directions, while the next k-mer is frequent
        seq = rows[0][0]
        floor = counts[seq] * 0.5
        grown = True
        while grown:
            grown = False
            best = max(((counts.get(seq[-(k - 1):] + c, 0), c) for c in "ACGT"))
            if best[0] >= floor:
                seq += best[1]
                grown = True
            best = max(((counts.get(c + seq[: k - 1], 0), c) for c in "ACGT"))
            if best[0] >= floor:
                seq = best[1] + seq
                grown = True
            if len(seq) > 80:
                break
        assert adapter in seq or seq in adapter and len(seq) >= len(adapter) - 2, (which, seq)


def test_reference_shim(built, tmp_path):
    """integration/apc_reference_shim.h: the reference's own signatures (count_kmers +
    get_most_frequent, errorCount) over the C ABI, driven like the reference's main loop."""
    from approx_counter_b200 import host
    k, sl, n, lim = 16, 100, 2500, 150
    path = tmp_path / "reads.fa"
    host.synth_write(path, 555, n, sl)
    exe = tmp_path / "shim_demo"
    libdir = os.path.join(ROOT, "approx_counter_b200", "csrc")
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "integration"), os.path.join(ROOT, "integration", "shim_demo.cpp"),
                    "-o", str(exe), "-L", libdir, "-lapc", f"-Wl,-rpath,{libdir}"], check=True)
    r = host.Reads(path)
    want = oracle_files([r.seq(i) for i in range(len(r))], k, sl, lim, 1.0, tmp_path)
    # the host's copy of the reads handed to upload(), then readRecords + sampleSequences on the GPU instead
    for extra in ([], ["device"]):
        p = subprocess.run([str(exe), str(path), str(k), str(sl), str(lim), "1.0", str(tmp_path / "shim")] + extra,
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        for which in ("start", "end"):
            assert (tmp_path / f"shim_0.{which}").read_bytes() == want[which][1]
            os.unlink(tmp_path / f"shim_0.{which}")


def test_binary_reads_a_pipe(built, tmp_path):
    """The default route maps the input; a pipe cannot be mapped and goes to the host parser (which reads it whole):
    same files as from the file itself."""
    from approx_counter_b200 import host
    path = tmp_path / "reads.fa"
    host.synth_write(path, 808, 1200, 60)
    outs = {}
    for name, source in (("file", str(path)), ("pipe", "/dev/stdin")):
        cmd = [BIN, "-k", "12", "-sn", "1200", "-sl", "60", "-lim", "80", "-v", "2", "-o", str(tmp_path / name), source]
        if name == "pipe":
            cat = subprocess.Popen(["cat", str(path)], stdout=subprocess.PIPE)
            p = subprocess.run(cmd, stdin=cat.stdout, capture_output=True, text=True, timeout=300)
            cat.wait()
        else:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        assert ("using the host parser" in p.stdout) == (name == "pipe")
        outs[name] = [(tmp_path / f"{name}_0.{w}").read_bytes() for w in ("start", "end")]
    assert outs["file"] == outs["pipe"] and len(outs["file"][0]) > 0
    p = subprocess.run([BIN, "-k", "12", "-o", str(tmp_path / "none"), str(tmp_path / "missing.fa")], capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 1 and "could not open" in p.stderr
