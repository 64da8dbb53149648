"""Device-side ingest (reference readRecords :819-825 + sampleSequences :415-476):
apc_ingest_fastx + apc_sample_resident against the host parser and sampler on the same files — record
counts, read lengths and sampled rows byte for byte, then the exact and approximate counts of both
routes — and the grammar the device parser refuses (APC_ERR_FORMAT: those files are the host parser's)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "approx_counter_b200", "csrc", "approx_counter")
APC_ERR_FORMAT = -9


def random_reads(rng, n, lo, hi, with_n=True):
    reads = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = rng.choice(np.frombuffer(b"ACGTacgt" + (b"N" if with_n else b"A"), np.uint8), size=L)
        reads.append(s.tobytes())
    return reads


def fastx_bytes(reads, fastq, eol=b"\n", tail=b"\n", rng=None):
    out = []
    for i, s in enumerate(reads):
        if fastq:
            # quality strings starting with '@' / '+' / '>' are legal: the device parser counts lines, it does not guess
            q = bytes([b"@+>I"[(i + j) % 4] for j in range(len(s))])
            out.append(b"@r%d some text\t%d" % (i, i) + eol + s + eol + b"+" + (b"r%d" % i if i % 3 == 0 else b"") + eol + q)
        else:
            out.append(b">r%d len=%d" % (i, len(s)) + eol + s)
    return eol.join(out) + tail if out else tail


def host_view(path):
    from approx_counter_b200 import host
    r = host.Reads(path)
    return r, np.array([len(r.seq(i)) for i in range(len(r))], np.uint32)


@pytest.mark.parametrize("fastq", [False, True])
@pytest.mark.parametrize("eol,tail", [(b"\n", b"\n"), (b"\r\n", b"\r\n"), (b"\n", b""), (b"\n", b"\n\n \r\n\t\n")])
def test_ingest_and_sampling_match_the_host_route(built, counter, tmp_path, fastq, eol, tail):
    from approx_counter_b200 import host
    rng = np.random.default_rng(31 + fastq + len(tail))
    cut, sn = 40, 700
    reads = random_reads(rng, 1200, 0, 260)      # many reads shorter than 2 * cut, some empty
    reads[0] = reads[0] + b"ACGT" * 30           # a first and a last read that are eligible
    reads[-1] = b"TTGCA" * 40
    data = fastx_bytes(reads, fastq, eol, tail)
    path = tmp_path / ("r.fq" if fastq else "r.fa")
    path.write_bytes(data)
    r, lens = host_view(path)
    n, is_fq = counter.ingest_fastx(data)
    assert (n, is_fq) == (len(reads), fastq) and n == len(r)
    assert np.array_equal(counter.ingest_lengths(), lens)
    assert np.array_equal(lens, np.array([len(s) for s in reads], np.uint32))
    eligible = int((lens >= 2 * cut).sum())
    for seed in (5, 6):
        order = host.shuffle_order(n, seed)
        assert sorted(order.tolist()) == list(range(n))
        for bot in (False, True):
            for want_n in (sn, 10 * n, 1):
                want = r.sample(want_n, cut, bot, seed)
                got_n = counter.sample_resident(want_n, cut, bot, order)
                assert got_n == len(want) == min(want_n, eligible)
                got = counter.download_sample()
                assert got.shape == want.shape and np.array_equal(got, want)
    # against the oracle's restatement of sampleSequences (:415-476) on the same reads and the same shuffled ids
    all_codes, all_offs = orc.encode(reads)
    order = host.shuffle_order(n, 5)
    for bot in (False, True):
        oc, oo = orc.sample_sequences(all_codes, all_offs, order, sn, cut, bot)
        assert counter.sample_resident(sn, cut, bot, order) == len(oo) - 1
        got_codes, got_offs = orc.encode_matrix(counter.download_sample())
        assert np.array_equal(got_codes, oc) and np.array_equal(got_offs, oo)
    # the counts of the gathered sample == the counts of the same rows uploaded from the host
    k = 12
    want = r.sample(sn, cut, True, 5)
    counter.sample_resident(sn, cut, True, host.shuffle_order(n, 5))
    km_d, ct_d, nd_d, hn_d = counter.count_kmers_topn(k, 1e30, 150)
    ap_d = counter.errorCount(km_d, k)
    counter.upload_sample(want)
    km_h, ct_h, nd_h, hn_h = counter.count_kmers_topn(k, 1e30, 150)
    ap_h = counter.errorCount(km_h, k)
    assert np.array_equal(km_d, km_h) and np.array_equal(ct_d, ct_h) and (nd_d, hn_d) == (nd_h, hn_h)
    assert np.array_equal(ap_d, ap_h)
    codes, offs = orc.encode_matrix(want)
    assert np.array_equal(ap_d, orc.error_count(codes, offs, km_d, k, fast=True))


def test_file_order_sampling_and_many_tiles(built, counter, tmp_path):
    """order = NULL walks the file in order; a file of a few thousand 4 KB tiles with lines of every length
    (newline-dense stretches included) exercises the tile prefix sums of the newline index."""
    rng = np.random.default_rng(77)
    reads = random_reads(rng, 20000, 0, 3) + random_reads(rng, 6000, 150, 700) + random_reads(rng, 5000, 0, 2)
    perm = rng.permutation(len(reads))
    reads = [reads[i] for i in perm]
    for fastq in (False, True):
        data = fastx_bytes(reads, fastq)
        assert len(data) > 400 * 4096
        n, _ = counter.ingest_fastx(np.frombuffer(data, np.uint8))
        lens = np.array([len(s) for s in reads], np.uint32)
        assert n == len(reads) and np.array_equal(counter.ingest_lengths(), lens)
        cut = 75
        ids = [i for i in range(n) if lens[i] >= 2 * cut][:4000]
        for bot in (False, True):
            assert counter.sample_resident(4000, cut, bot) == len(ids)
            got = counter.download_sample()
            want = np.array([np.frombuffer(reads[i][len(reads[i]) - 1 - cut:] if bot else reads[i][:cut], np.uint8)
                             for i in ids])
            assert np.array_equal(got, want)


def test_small_and_empty_inputs(built, counter):
    for data in (b"", b"\n\n", b"  \r\n"):
        assert counter.ingest_fastx(data) == (0, False)
        assert counter.sample_resident(10, 5, False) == 0
        assert counter.sample_info()[0] == 0
    assert counter.ingest_fastx(b">only\nACGTACGTAC") == (1, False)
    assert counter.sample_resident(10, 5, True) == 1
    assert counter.download_sample().tobytes() == b"ACGTACGTAC"[10 - 1 - 5:]
    assert counter.sample_resident(10, 5, False) == 1 and counter.download_sample().tobytes() == b"ACGTA"
    assert counter.sample_resident(10, 6, False) == 0          # 10 < 2 * 6 (:461)
    assert counter.sample_resident(10, 0, False) == 0          # cut == 0 takes nothing (:461)
    assert counter.ingest_fastx(b"@q\nACGT\n+\nIIII\n") == (1, True)
    assert counter.ingest_lengths().tolist() == [4]
    assert counter.ingest_fastx(b">a\n\n>b\nAC\n") == (2, False)   # an empty sequence line
    assert counter.ingest_lengths().tolist() == [0, 2]
    with pytest.raises(Exception):
        counter.sample_resident(1, 1, False, order=np.array([0, 2], np.uint32))   # not a permutation
    with pytest.raises(Exception):
        counter.sample_resident(1, 1, False, order=np.array([0], np.uint32))      # wrong length


@pytest.mark.parametrize("data", [
    b">a\nAC GT\n>b\nAC\n",                      # blank inside a sequence
    b">a\nAC\rGT\n>b\nAC\n",                     # CR inside a sequence
    b">a\nACGT\n  AC\n>b\nAC\n",                 # wrapped FASTA with blanks in front of a line
    b"@q\nACGT\n+\nIII\n",                       # quality shorter than the sequence
    b"@q\nACGT\n+\nIIII\n@r\nAC\n+\nIII\n",      # quality longer
    b"@q\nAC\nGT\n+\nII\nII\n",                  # wrapped FASTQ (line count still a multiple of 4 after this record?)
    b"@q\nACGT\n+\nIIII\n@r\nAC\n+\n",           # truncated record
    b"@q\nACGT\nIIII\n+\n",                      # '+' line in the wrong place
    b"ACGT\n",                                   # no header at all
    b"@q\n+CGT\n+\nIIII\n",                      # sequence line starting with '+' (the host reads it as the separator)
])
def test_inputs_outside_the_grammar_are_refused(built, counter, data):
    from approx_counter_b200 import ApcError
    with pytest.raises(ApcError) as e:
        counter.ingest_fastx(data)
    assert e.value.status == APC_ERR_FORMAT
    with pytest.raises(ApcError):
        counter.sample_resident(1, 1, False)     # no resident file after a refused ingest
    assert counter.ingest_fastx(b">a\nACGT\n") == (1, False)   # and the context is still usable


def wrapped_fasta_bytes(reads, rng, eol=b"\n", width=None, blank_lines=False):
    out = []
    for i, s in enumerate(reads):
        w = width or int(rng.integers(1, 90))
        lines = [s[j:j + w] for j in range(0, len(s), w)]
        if blank_lines and i % 5 == 0:
            lines.insert(int(rng.integers(0, len(lines) + 1)), b"")          # a blank line inside the record
        if blank_lines and i % 7 == 0:
            lines.append(b" \t")                                            # a line of blanks in front of the next header
        out.append(b">r%d wrapped" % i + eol + b"".join(l + eol for l in lines))
    return b"".join(out)


@pytest.mark.parametrize("eol,width,blank_lines,lead", [(b"\n", 60, False, b""), (b"\r\n", 70, False, b""),
                                                        (b"\n", None, True, b"\n \n"), (b"\r\n", None, True, b"")])
def test_wrapped_fasta_is_relaid_on_the_device(built, counter, tmp_path, eol, width, blank_lines, lead):
    """FASTA with wrapped sequences, blank lines and headers without a sequence line: the device re-lays the records
    on one line each and indexes that; records and samples equal the host parser's."""
    from approx_counter_b200 import host
    rng = np.random.default_rng(99 + len(eol) + (width or 0))
    reads = random_reads(rng, 900, 0, 400)
    reads[5] = b""
    reads[-1] = b"" if blank_lines else reads[-1]          # the file ends with a header (and blank lines)
    data = lead + wrapped_fasta_bytes(reads, rng, eol, width, blank_lines)
    path = tmp_path / "w.fa"
    path.write_bytes(data)
    r, lens = host_view(path)
    assert np.array_equal(lens, np.array([len(s) for s in reads], np.uint32))
    assert counter.ingest_fastx(data) == (len(reads), False)
    assert np.array_equal(counter.ingest_lengths(), lens)
    cut = 50
    order = host.shuffle_order(len(reads), 8)
    for bot in (False, True):
        want = r.sample(300, cut, bot, 8)
        assert counter.sample_resident(300, cut, bot, order) == len(want) > 0
        assert np.array_equal(counter.download_sample(), want)
    # the same reads on one line each give the same resident state
    single = fastx_bytes(reads, False)
    assert counter.ingest_fastx(single) == (len(reads), False)
    assert np.array_equal(counter.ingest_lengths(), lens)


def test_fasta_edge_grammars_now_taken(built, counter):
    for data, lens in ((b">a\nACGT\nACGT\n>b\nAC\n", [8, 2]), (b">a\nACGT\n\n>b\nAC\n", [4, 2]), (b"\n>a\nACGT\n", [4]),
                       (b">a\nACGT\n>b\n", [4, 0]), (b">a\n>b\nACGT\n", [0, 4]), (b">\n>\n>\n", [0, 0, 0]),
                       (b">a\r\nAC\r\nGT\r\n\r\n>b\r\n", [4, 0])):
        assert counter.ingest_fastx(data) == (len(lens), False), data
        assert counter.ingest_lengths().tolist() == lens, data


@pytest.mark.parametrize("fastq,k,sl,lim,n", [(False, 16, 100, 200, 3000), (True, 20, 150, 300, 1500)])
def test_binary_device_ingest_writes_the_same_files(built, tmp_path, fastq, k, sl, lim, n):
    """--ingest device against --ingest host: same four files for a sample that is a strict subset of the reads
    (same --seed), for a wrapped FASTA as well; the host parser takes over for blanks inside sequence lines."""
    from approx_counter_b200 import host
    path = tmp_path / ("reads.fq" if fastq else "reads.fa")
    host.synth_write(path, 777 + k, n, sl, fastq=fastq)
    files = {}
    for ingest in ("host", "device"):
        p = subprocess.run([BIN, "-k", str(k), "-sn", str(n // 2), "-sl", str(sl), "-lim", str(lim), "--seed", "11",
                            "--ingest", ingest, "-v", "2", "-e", str(tmp_path / f"e_{ingest}"), "-o", str(tmp_path / f"o_{ingest}"),
                            str(path)], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        assert ("indexed on the GPU" in p.stdout) == (ingest == "device")
        assert f"Number of sequences found: {n}." in p.stdout and f"Sampled {n // 2} sequences" in p.stdout
        files[ingest] = [(tmp_path / f"{x}_{ingest}_0.{w}").read_bytes() for x in "eo" for w in ("start", "end")]
        assert all(len(f) > 0 for f in files[ingest])
    assert files["host"] == files["device"]
    if not fastq:
        r = host.Reads(path)
        for name, sep, note in (("wrapped", b"\n", "indexed on the GPU"), ("blanks", b" \n ", "using the host parser")):
            wrapped = tmp_path / f"{name}.fa"
            with open(wrapped, "wb") as f:
                for i in range(len(r)):
                    s = r.seq(i)
                    f.write(b">r%d\n" % i + sep.join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n")
            p = subprocess.run([BIN, "-k", str(k), "-sn", str(n // 2), "-sl", str(sl), "-lim", str(lim), "--seed", "11",
                                "--ingest", "device", "-v", "2", "-e", str(tmp_path / f"e_{name}"), "-o", str(tmp_path / f"o_{name}"),
                                str(wrapped)], capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, p.stderr
            assert note in p.stdout
            assert [(tmp_path / f"{x}_{name}_0.{w}").read_bytes() for x in "eo" for w in ("start", "end")] == files["host"]


def test_staged_copy_of_a_large_file(built, counter, tmp_path):
    """Files of 64 MB and more travel through the two page-locked staging pieces (several pieces, a ragged last
    one): same records and samples as the host route, with and without the staging."""
    from approx_counter_b200 import host
    n, sl = 150_000, 150
    path = tmp_path / "big.fa"
    host.synth_write(path, 31337, n, sl)
    assert os.path.getsize(path) > 4 * (16 << 20)
    data = np.fromfile(path, np.uint8)
    r, lens = host_view(path)
    order = host.shuffle_order(n, 3)
    want = {bot: r.sample(n // 3, sl, bot, 3) for bot in (False, True)}
    try:
        for staging in (1, 0):
            counter.set_option("ingest_staging", staging)
            assert counter.ingest_fastx(data) == (n, False)
            assert np.array_equal(counter.ingest_lengths(), lens)
            for bot in (False, True):
                assert counter.sample_resident(n // 3, sl, bot, order) == n // 3
                assert np.array_equal(counter.download_sample(), want[bot])
    finally:
        counter.set_option("ingest_staging", 1)


def test_peer_upload_of_a_shard(built, counter):
    """apc_upload_sample_peer: rows of one context's resident sample become another context's sample (here two
    contexts on one GPU; across GPUs the same call copies over NVLink — tests/test_gpu_multi.py runs the binary with
    --gpus N through it)."""
    from approx_counter_b200 import ApcError, ApproxCounter, host
    n, sl, k = 5000, 80, 14
    rows = host.synth_ends(4141, 0, n, sl, True)
    data = b"".join(b">r%d\n" % i + b"ACGT" * 25 + rows[i].tobytes() + b"\n" for i in range(n))
    assert counter.ingest_fastx(data) == (n, False)
    assert counter.sample_resident(n, sl, True) == n
    assert np.array_equal(counter.download_sample(), rows)
    km, _, _, _ = counter.count_kmers_topn(k, 1e30, 120)
    whole = counter.errorCount(km, k)
    with ApproxCounter(0) as other:
        total = np.zeros(len(km), np.uint64)
        for first, cnt in ((0, 1024), (1024, 2976), (4000, 1000), (5000, 0)):
            other.upload_sample_peer(counter, first, cnt)
            assert np.array_equal(other.download_sample(), rows[first:first + cnt])
            total += other.errorCount(km, k)
        assert np.array_equal(total, whole)
        with pytest.raises(ApcError):
            other.upload_sample_peer(counter, 4000, 1001)
        with pytest.raises(ApcError):
            other.upload_sample_peer(other, 0, 1)
