"""CPU tests of the host-side C++ the drop-in binary runs (include/apc_host.h),
checked against the oracle's restatement of the same reference functions."""
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")


@pytest.fixture(scope="module")
def host(built):
    from approx_counter_b200 import host
    return host


def test_codec_matches_oracle(host):
    rng = np.random.default_rng(0)
    for k in (2, 7, 16, 20, 32):
        for _ in range(50):
            v = int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1)
            s = orc.int2dna(v, k)
            assert host.int2dna(v, k) == s
            assert host.dna2int(s) == v
            assert host.dna2int(s.lower()) == v
    with pytest.raises(ValueError):
        host.dna2int("ACGN")


def test_threshold_and_filter_match_oracle(host):
    rng = np.random.default_rng(1)
    for k in range(3, 33):
        for lc in (0.5, 1.0, 1.5, 2.25):
            thr = host.adjust_threshold(lc, 16, k)
            assert thr == orc.adjust_threshold(lc, 16, k)
            smin = host.lc_min_filtered_sum(k, thr)
            for _ in range(40):
                v = int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1)
                if rng.random() < 0.3:  # make it repetitive
                    v &= int(rng.integers(0, 1 << 16)) * 0x0001000100010001 & ((1 << (2 * k)) - 1)
                assert host.get_complexity(v, k) == orc.get_complexity(v, k)
                assert host.have_low_complexity(v, k, thr) == orc.have_low_complexity(v, k, thr)
                # the integer form the device kernels use
                assert (orc.dimer_sum(v, k) >= smin) == orc.have_low_complexity(v, k, thr)


def test_get_most_frequent_matches_oracle(host):
    rng = np.random.default_rng(2)
    for k in (4, 16, 32):
        keys = np.unique(rng.integers(0, 1 << 62, 3000).astype(np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else (1 << 64) - 1))
        cnts = rng.integers(1, 4, len(keys)).astype(np.uint64)
        for lim in (0, 1, 100, 10 ** 6):
            a = host.get_most_frequent(keys, cnts, lim, k)
            b = orc.get_most_frequent(keys, cnts, lim, k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_export_and_kmer_list(host, tmp_path):
    p = tmp_path / "c.txt"
    assert host.export_counter([host.dna2int("ACGT"), host.dna2int("GGGG")], [7, 1], 4, p)
    assert p.read_text() == "ACGT\t7\nGGGG\t1\n"
    q = tmp_path / "fk.txt"
    q.write_text("ACGT\nNNNN\nacgt\n\nTTTT\r\n")
    assert host.parse_kmer_list(q).tolist() == [host.dna2int("ACGT"), host.dna2int("ACGT"), host.dna2int("TTTT")]
    assert not host.export_counter([1], [1], 4, tmp_path / "no" / "such" / "dir.txt")


FASTA = ">r0 some id\nACGTACGTAC\nGGGG\n>r1\nacgtnnRYAC\n>r2\n\n>r3\nTTTT\n"
FASTQ = "@r0\nACGTACGTACGGGG\n+\nIIIIIIIIIIIIII\n@r1\nacgtnnRYAC\n+r1\n@@@@>>>>++\n@r2\nTTTT\n+\n@III\n"


@pytest.mark.parametrize("mmap_bytes", ["0", "1"])
def test_fasta_fastq_reader(host, tmp_path, monkeypatch, mmap_bytes):
    monkeypatch.setenv("APCH_MMAP_BYTES", mmap_bytes)   # 1: parse from a mapping of the file where its records allow it
    fa, fq = tmp_path / "x.fa", tmp_path / "x.fq"
    fa.write_text(FASTA)
    fq.write_text(FASTQ)
    r = host.Reads(fa)
    assert [r.seq(i) for i in range(len(r))] == [b"ACGTACGTACGGGG", b"acgtnnRYAC", b"", b"TTTT"]
    assert not r.mapped                       # the fixture has a multi-line record
    r = host.Reads(fq)
    assert [r.seq(i) for i in range(len(r))] == [b"ACGTACGTACGGGG", b"acgtnnRYAC", b"TTTT"]
    with pytest.raises(OSError):
        host.Reads(tmp_path / "missing.fa")


def test_sampling_matches_oracle(host, tmp_path):
    rng = np.random.default_rng(4)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(5, 80)))) for _ in range(60)]
    fa = tmp_path / "s.fa"
    fa.write_text("".join(f">r{i}\n{r.decode()}\n" for i, r in enumerate(reads)))
    r = host.Reads(fa)
    cut = 12
    codes, offs = orc.encode(reads)
    ident = np.arange(len(reads), dtype=np.uint64)
    for bot in (False, True):
        got = r.sample(len(reads), cut, bot, seed=3)       # sn >= #reads: the set is all eligible reads
        wc, wo = orc.sample_sequences(codes, offs, ident, len(reads), cut, bot)
        want = np.frombuffer(b"ACGT", np.uint8)[wc].reshape(-1, cut + (1 if bot else 0))
        assert got.shape == want.shape
        assert sorted(map(bytes, got)) == sorted(map(bytes, want))
        few = r.sample(5, cut, bot, seed=3)                # same seed -> same prefix of the walk
        assert np.array_equal(few, got[:5])
    a = r.sample(10, cut, False, seed=1)
    b = r.sample(10, cut, False, seed=2)
    assert not np.array_equal(a, b)


def test_synthetic_reads_are_deterministic(host, tmp_path):
    a = host.synth_ends(1001, 0, 64, 100, False)
    b = host.synth_ends(1001, 32, 32, 100, False)
    assert np.array_equal(a[32:], b)                       # read i depends on (seed, i) only
    e = host.synth_ends(1001, 0, 64, 100, True)
    assert a.shape == (64, 100) and e.shape == (64, 101)
    assert set(np.unique(a)) <= set(b"ACGTN")
    fa = tmp_path / "syn.fa"
    host.synth_write(fa, 1001, 64, 100)
    r = host.Reads(fa)
    assert len(r) == 64
    for i in (0, 17, 63):
        s = r.seq(i)
        assert len(s) >= 200 and s[:100] == bytes(a[i]) and s[-101:] == bytes(e[i])
    fq = tmp_path / "syn.fq"
    host.synth_write(fq, 1001, 64, 100, fastq=True)
    rq = host.Reads(fq)
    assert [rq.seq(i) for i in range(64)] == [r.seq(i) for i in range(64)]
    adapter = b"AATGTACTTCGTTCAG"
    assert sum(adapter in bytes(row) for row in a) > 10    # planted, partly mutated


def run_cli(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=60)


def test_cli_flag_surface_without_gpu(built, tmp_path):
    """Option parsing and validation mirror the reference (:604-669, :691-790) and need no GPU."""
    r = run_cli("--help")
    assert r.returncode == 0
    for flag in ("-lc", "-sn", "-sl", "-nt", "-k", "-lim", "-mr", "-v", "-e", "-conf", "-fk", "-sk", "-se", "-o"):
        assert f"{flag}, --" in r.stdout
    assert run_cli().returncode == 1                        # missing input (:697-698)
    assert run_cli("-zz", "1", "x.fa").returncode == 1      # unknown option
    assert run_cli("-k", "abc", "x.fa").returncode == 1     # not an integer
    fa = tmp_path / "x.fa"
    fa.write_text(">r\nACGT\n")
    r = run_cli("-k", "40", str(fa))                        # uncaught std::invalid_argument (:781-783)
    assert r.returncode != 0 and "kmer size must be between 2 and 32" in r.stderr
    r = run_cli("-k", "16", "-sl", "10", str(fa))           # (:785-787)
    assert r.returncode != 0 and "k <= sl" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = run_cli("-k", "4", "-sl", "10", str(fa))
        assert r.returncode == 2 and "cannot open CUDA device" in r.stderr   # no CPU fallback


def _py_parse(text):
    """Straightforward reference parser (multi-line FASTA, 4-line or multi-line FASTQ)."""
    lines = text.replace("\r", "").split("\n")
    seqs, i = [], 0
    while i < len(lines):
        ln = lines[i].strip()
        if not ln:
            i += 1
            continue
        if ln[0] == ">":
            i += 1
            s = ""
            while i < len(lines) and not lines[i].startswith(">"):
                s += lines[i].replace(" ", "").replace("\t", "")
                i += 1
            seqs.append(s)
        else:
            assert ln[0] == "@"
            i += 1
            s = ""
            while not lines[i].startswith("+"):
                s += lines[i].strip()
                i += 1
            i += 1
            q = 0
            while q < len(s):
                q += len(lines[i])
                i += 1
            seqs.append(s)
    return seqs


@pytest.mark.parametrize("mmap_bytes", ["0", "1"])
@pytest.mark.parametrize("kind", ["fasta_multiline", "fastq_at_quality", "fasta_crlf", "fasta_single", "fastq_crlf"])
def test_parallel_parser_pieces(host, tmp_path, kind, monkeypatch, mmap_bytes):
    """The file is cut into pieces at record boundaries and compacted in place; tiny pieces
    force many cuts through multi-line records, CRLF and quality lines starting with '@'."""
    rng = np.random.default_rng(8)
    recs = []
    for i in range(3000):
        n = int(rng.integers(0, 300))
        s = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, n))
        if kind == "fasta_multiline":
            body = "\n".join(s[j:j + 60] for j in range(0, max(n, 1), 60))
            recs.append(f">r{i} desc > @ +\n{body}\n" + ("\n" if i % 50 == 0 else ""))
        elif kind == "fasta_crlf":
            recs.append(f">r{i}\r\n{s}\r\n")
        elif kind == "fasta_single":
            recs.append(f">r{i} desc > @ +\n{s}\n" + ("\n" if i % 40 == 0 else ""))
        elif kind == "fastq_crlf":
            qual = "".join("@+>I#"[int(x)] for x in rng.integers(0, 5, n))
            recs.append(f"@r{i}\r\n{s}\r\n+r{i}\r\n{qual}\r\n")
        else:
            qual = "".join("@+>I#"[int(x)] for x in rng.integers(0, 5, n))
            if n:
                qual = "@" + qual[1:]          # every quality line starts with '@'
            recs.append(f"@r{i} x\n{s}\n+\n{qual}\n")
    text = "".join(recs)
    path = tmp_path / "x.fx"
    path.write_text(text, newline="")
    want = _py_parse(text)
    # mmap_bytes 1: single-line records are parsed straight from a read-only mapping of the file (no copy);
    # multi-line records make that path give way to the copying parser
    monkeypatch.setenv("APCH_MMAP_BYTES", mmap_bytes)
    for piece in ("1000000000", "4096", "65536"):
        monkeypatch.setenv("APCH_PIECE_BYTES", piece)
        r = host.Reads(path)
        got = [r.seq(i).decode() for i in range(len(r))]
        assert got == want
        assert r.mapped == (mmap_bytes == "1" and kind != "fasta_multiline")


@pytest.mark.parametrize("mmap_bytes", ["0", "1"])
def test_parallel_parser_multiline_fastq(host, tmp_path, monkeypatch, mmap_bytes):
    """Wrapped (multi-line) FASTQ whose quality lines start with '@' and '+' defeats the 4-line boundary
    heuristic; every piece is verified read-only before the in-place compaction and such files are parsed in
    one piece.  Many small random files with tiny pieces (the advisor's reproduction)."""
    monkeypatch.setenv("APCH_MMAP_BYTES", mmap_bytes)
    bad = 0
    for seed in range(120):
        rng = np.random.default_rng(1000 + seed)
        width = int(rng.integers(5, 40))
        recs = []
        for i in range(int(rng.integers(5, 60))):
            n = int(rng.integers(0, 200))
            s = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, n))
            qual = "".join("@+@+I#"[int(x)] for x in rng.integers(0, 6, n))
            wrap = lambda t: "\n".join(t[j:j + width] for j in range(0, max(len(t), 1), width))
            if rng.random() < 0.3:
                recs.append(f"@r{i}\n{s}\n+\n{qual}\n")      # some 4-line records in between
            else:
                recs.append(f"@r{i}\n{wrap(s)}\n+r{i}\n{wrap(qual)}\n")
        text = "".join(recs)
        path = tmp_path / f"m{seed}.fq"
        path.write_text(text, newline="")
        want = _py_parse(text)
        for piece in ("16", "64", "700"):
            monkeypatch.setenv("APCH_PIECE_BYTES", piece)
            try:
                r = host.Reads(path)
                got = [r.seq(i).decode() for i in range(len(r))]
            except OSError:
                got = None
            bad += got != want
    assert bad == 0


def test_oracle_synth_equals_host_synth(host):
    """The CPU side generates its workloads with oracle/synth_reads.c (so that the reference arm never loads
    the product library); the two generators must agree byte for byte."""
    for seed, first, n, sl in ((1001, 0, 300, 100), (1003, 12345, 200, 150), (1004, 999_000, 100, 200), (7, 5, 50, 16),
                               (1002 | 1 << 40, 0, 300, 100)):   # bit 40: adapter offsets uniform in 0..sl/2
        for bot in (False, True):
            assert np.array_equal(orc.synth_ends(seed, first, n, sl, bot), host.synth_ends(seed, first, n, sl, bot))


def test_shuffle_order_is_the_samplers_order(built, tmp_path):
    """apch_shuffle_order (:423-429) hands out the ids apch_sample walks: gathering by hand along it gives the
    sampler's rows (this is what apc_sample_resident does on the device)."""
    from approx_counter_b200 import host
    rng = np.random.default_rng(3)
    reads = [rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(0, 120))).tobytes() for _ in range(300)]
    path = tmp_path / "r.fa"
    path.write_bytes(b"".join(b">r%d\n%s\n" % (i, s) for i, s in enumerate(reads)))
    r = host.Reads(path)
    for seed in (0, 1, 99):
        order = host.shuffle_order(len(reads), seed)
        assert sorted(order.tolist()) == list(range(len(reads)))
        assert np.array_equal(order, host.shuffle_order(len(reads), seed))
        for cut, bot, sn in ((20, False, 50), (20, True, 1000), (35, True, 7)):
            ids = [i for i in order if len(reads[i]) >= 2 * cut][:sn]
            want = [reads[i][len(reads[i]) - 1 - cut:] if bot else reads[i][:cut] for i in ids]
            got = r.sample(sn, cut, bot, seed)
            assert [row.tobytes() for row in got] == want
    assert len(host.shuffle_order(0, 1)) == 0


def test_ingest_test_inputs_mean_what_the_gpu_tests_assume(built, tmp_path):
    """The inputs tests/test_gpu_ingest.py feeds the device parser, through the HOST parser and the plain Python
    parser of this file: the generators really produce the reads they are given, and the edge grammars have the record
    lengths the GPU tests expect (so a GPU-side failure there is the device's, not the test's)."""
    import importlib.util
    from approx_counter_b200 import host
    spec = importlib.util.spec_from_file_location("gpu_ingest_inputs", os.path.join(os.path.dirname(__file__), "test_gpu_ingest.py"))
    gi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gi)

    def host_lens(data, name):
        p = tmp_path / name
        p.write_bytes(data)
        r = host.Reads(p)
        return [r.seq(i) for i in range(len(r))]

    rng = np.random.default_rng(12)
    reads = gi.random_reads(rng, 300, 0, 260)
    for fastq in (False, True):
        for eol, tail in ((b"\n", b"\n"), (b"\r\n", b"\r\n"), (b"\n", b""), (b"\n", b"\n\n \r\n\t\n")):
            data = gi.fastx_bytes(reads, fastq, eol, tail)
            assert host_lens(data, "a.fx") == reads
            assert [s.encode() for s in _py_parse(data.decode())] == reads
    for eol, width, blank_lines, lead in ((b"\n", 60, False, b""), (b"\r\n", 70, False, b""), (b"\n", None, True, b"\n \n"),
                                          (b"\r\n", None, True, b"")):
        rs = list(reads)
        rs[5] = b""
        if blank_lines:
            rs[-1] = b""
        data = lead + gi.wrapped_fasta_bytes(rs, rng, eol, width, blank_lines)
        assert host_lens(data, "w.fa") == rs
    for data, lens in ((b">a\nACGT\nACGT\n>b\nAC\n", [8, 2]), (b">a\nACGT\n\n>b\nAC\n", [4, 2]), (b"\n>a\nACGT\n", [4]),
                       (b">a\nACGT\n>b\n", [4, 0]), (b">a\n>b\nACGT\n", [0, 4]), (b">\n>\n>\n", [0, 0, 0]),
                       (b">a\r\nAC\r\nGT\r\n\r\n>b\r\n", [4, 0])):
        assert [len(s) for s in host_lens(data, "e.fa")] == lens
