"""CPU tests of the oracle (oracle/): known-answer vectors, closed form vs the
literal SeqAn search-scheme model, filter/comparator/sampling restatements.
The reference holds no tests or fixtures of its own, so the
vectors are the hand-checkable ones of SURVEY.md App. B plus tests/golden/."""
import json
import os

import numpy as np
import pytest

from oracle import orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
ACGT = np.frombuffer(b"ACGT", np.uint8)


def test_codec_roundtrip():
    assert orc.dna2int("ACGT") == 0b00011011          # first base most significant (:55-62)
    assert orc.int2dna(0b00011011, 4) == "ACGT"        # (:70-78)
    assert orc.dna2int("T" * 32) == (1 << 64) - 1
    rng = np.random.default_rng(0)
    for k in (2, 5, 16, 31, 32):
        for _ in range(20):
            v = int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1)
            assert orc.dna2int(orc.int2dna(v, k)) == v


@pytest.mark.parametrize("case", GOLD["appendix_b"], ids=lambda c: c["kmer"])
def test_appendix_b(case):
    k, km = case["k"], orc.dna2int(case["kmer"])
    codes, offs = orc.encode(case["reads"])
    assert int(orc.error_count(codes, offs, [km], k)[0]) == case["total"]
    assert int(orc.error_count(codes, offs, [km], k, fast=True)[0]) == case["total"]
    out, flags = orc.seqan_model_error_count(codes, offs, [km], k, want_flags=True)
    assert int(out[0]) == case["total"]
    # per read: flagged at every level e >= d (R0 ⊆ R1 ⊆ R2), contribution 3 - d
    per = flags[0].sum(axis=0).tolist()
    assert per == case["per_read"]
    for r, contrib in enumerate(case["per_read"]):
        d = orc.min_infix_distance(orc.encode([case["reads"][r]])[0], km, k)
        assert max(0, 3 - d) == contrib
        assert flags[0][:, r].tolist() == [int(e >= d) for e in range(3)]


@pytest.mark.parametrize("case", GOLD["approx"], ids=lambda c: f"k{c['k']}n{len(c['reads'])}")
def test_golden_approx(case):
    k = case["k"]
    kmers = [orc.dna2int(s) for s in case["kmers"]]
    codes, offs = orc.encode(case["reads"])
    assert orc.error_count(codes, offs, kmers, k, fast=True).tolist() == case["counts"]
    if len(case["reads"]) <= 30:
        assert orc.error_count(codes, offs, kmers, k).tolist() == case["counts"]
    assert orc.fm_index_error_count(codes, offs, kmers, k).tolist() == case["counts"]   # the index-based form


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("k", [4, 5, 6, 7, 9, 11, 13, 16, 21])
def test_closed_form_equals_seqan_model(k, variant):
    """Σ_r max(0, 3 - d_r) == Σ_e popcount(tcount[e]) of the search-scheme recursion,
    including hits flush with read borders, reads shorter than k and N in the text."""
    rng = np.random.default_rng(1000 * variant + k)
    for trial in range(6):
        L = int(rng.integers(k - 2, 2 * k + 6))
        reads, needle = [], rng.choice(ACGT, size=k).tobytes()
        for r in range(10):
            n = int(rng.integers(max(0, k - 3), L + 1))
            body = bytearray(rng.choice(ACGT, size=n).tobytes())
            if r % 2 == 0 and n >= k:
                m = bytearray(needle)
                for _ in range(int(rng.integers(0, 4))):
                    p = int(rng.integers(0, len(m)))
                    op = int(rng.integers(0, 3))
                    if op == 0:
                        m[p] = int(rng.choice(ACGT))
                    elif op == 1 and len(m) > 1:
                        del m[p]
                    else:
                        m.insert(p, int(rng.choice(ACGT)))
                m = m[:n]
                pos = int(rng.choice([0, n - len(m)]))
                body[pos:pos + len(m)] = m
            if r == 3 and n:
                body[int(rng.integers(0, n))] = ord("N")
            reads.append(bytes(body))
        kmers = [orc.dna2int(needle.decode())] + \
                [int(x) & ((1 << (2 * k)) - 1) for x in rng.integers(0, 1 << 62, 2)]
        codes, offs = orc.encode(reads)
        want = orc.error_count(codes, offs, kmers, k)
        got = orc.seqan_model_error_count(codes, offs, kmers, k, variant=variant)
        assert np.array_equal(got, want), (k, trial, reads)
        assert np.array_equal(orc.error_count(codes, offs, kmers, k, fast=True), want)
        if variant == 0:  # the same recursion over the bidirectional FM index (the index-based CPU baseline)
            assert np.array_equal(orc.fm_index_error_count(codes, offs, kmers, k), want), (k, trial, reads)


def test_count_invariants():
    """3|R0| <= count <= 3n, and |R0| = number of READS holding the k-mer verbatim."""
    rng = np.random.default_rng(3)
    k = 10
    reads = [rng.choice(ACGT, size=60).tobytes() for _ in range(200)]
    needle = reads[0][5:15]
    for r in range(0, 200, 4):
        reads[r] = reads[r][:20] + needle + reads[r][30:]
    codes, offs = orc.encode(reads)
    got = int(orc.error_count(codes, offs, [orc.dna2int(needle.decode())], k, fast=True)[0])
    r0 = sum(needle in r for r in reads)
    assert 3 * r0 <= got <= 3 * len(reads)


def test_filter_thresholds_fp32():
    # smallest dimer sum the fp32 expression (:227-233) rejects, SURVEY.md App. B
    for lc, table in GOLD["filter_thresholds"].items():
        for k, want in table.items():
            k = int(k)
            thr = orc.adjust_threshold(float(lc), 16, k)
            sums = [s for s in range(0, 932, 2) if np.float32(s) / np.float32(2 * (k - 2)) >= np.float32(thr)]
            assert sums[0] == want
    assert orc.adjust_threshold(1.0, 16, 16) == 1.0
    assert abs(orc.adjust_threshold(1.0, 16, 20) - 1.6044445) < 1e-6
    assert abs(orc.adjust_threshold(1.0, 16, 32) - 4.271111) < 1e-6


def test_low_complexity_and_dimer_sum():
    k = 16
    poly_a = 0
    assert orc.dimer_sum(poly_a, k) == 15 * 14
    assert orc.have_low_complexity(poly_a, k, 1.0)
    ac = orc.dna2int("ACACACACACACACAC")  # 8 AC + 7 CA
    assert orc.dimer_sum(ac, k) == 8 * 7 + 7 * 6
    rnd = orc.dna2int("AATGTACTTCGTTCAG")
    assert not orc.have_low_complexity(rnd, k, 1.0)
    assert orc.get_complexity(rnd, k) == pytest.approx(orc.dimer_sum(rnd, k) / 28.0)
    # `>=`: a score exactly on the threshold is filtered (42/28 == 1.5)
    rng = np.random.default_rng(1)
    hit = None
    for _ in range(200000):
        v = int.from_bytes(rng.bytes(4), "little")
        if orc.dimer_sum(v, k) == 42:
            hit = v
            break
    assert hit is not None
    assert orc.have_low_complexity(hit, k, 1.5) and not orc.have_low_complexity(hit, k, 1.5000001)


def test_compare_count_order():
    k = 16
    a, b = orc.dna2int("AATGTACTTCGTTCAG"), orc.dna2int("ACACACACACACACAC")
    keys = np.array([a, b, 5, 7], np.uint64)
    cnts = np.array([3, 3, 9, 3], np.uint64)
    tk, tc = orc.get_most_frequent(keys, cnts, 10, k)
    # count desc, then complexity asc (:301), then k-mer value desc (:297)
    assert tc.tolist() == [9, 3, 3, 3]
    comp = [orc.get_complexity(int(x), k) for x in tk[1:]]
    assert comp == sorted(comp)
    same = np.array([0b0110, 0b1001], np.uint64)  # equal score, equal count
    tk, _ = orc.get_most_frequent(same, np.array([1, 1], np.uint64), 10, 4)
    assert tk.tolist() == sorted(same.tolist(), reverse=True)
    tk, _ = orc.get_most_frequent(keys, cnts, 2, k)
    assert len(tk) == 2


def test_count_kmers_skips_n_and_filters():
    reads = ["ACGTACGTAC", "ACGTNCGTAC", "AAAAAAAAAA", "ACG"]
    codes, offs = orc.encode(reads)
    keys, cnts, had_n = orc.count_kmers(codes, offs, 4, 1e9)
    d = dict(zip(keys.tolist(), cnts.tolist()))
    assert had_n == 4                                    # windows 1..4 of read 1 hold the N
    assert d[orc.dna2int("ACGT")] == 3 and d[orc.dna2int("AAAA")] == 7
    assert sum(d.values()) == 7 + 3 + 7
    keys, cnts, _ = orc.count_kmers(codes, offs, 4, 1.0)  # AAAA: sum=6, 6/4=1.5 >= 1 -> filtered
    assert orc.dna2int("AAAA") not in keys.tolist()
    keys, cnts, _ = orc.count_kmers(codes, offs, 4, 1e9, forbidden=[orc.dna2int("ACGT")])
    assert orc.dna2int("ACGT") not in keys.tolist()


def test_sample_sequences_off_by_one():
    reads = ["".join("ACGT"[(i + j) % 4] for j in range(30 + i)) for i in range(6)] + ["ACGTACGT"]
    codes, offs = orc.encode(reads)
    perm = np.arange(len(reads), dtype=np.uint64)[::-1].copy()
    sc, so = orc.sample_sequences(codes, offs, perm, 4, 10, False)
    lens = np.diff(so).tolist()
    assert lens == [10, 10, 10, 10]                       # prefix(seq, cut) (:466); short read skipped (:461)
    ec, eo = orc.sample_sequences(codes, offs, perm, 4, 10, True)
    assert np.diff(eo).tolist() == [11, 11, 11, 11]       # suffix(seq, len-1-cut) = cut+1 bases (:463)
    first = reads[5]
    assert bytes(np.frombuffer(b"ACGT", np.uint8)[ec[:11]]).decode() == first[-11:]


def test_export_format(tmp_path):
    p = tmp_path / "out.txt"
    assert orc.export_counter([orc.dna2int("ACGT"), orc.dna2int("TTTT")], [12, 3], 4, p)
    assert p.read_text() == "ACGT\t12\nTTTT\t3\n"        # :165


@pytest.mark.parametrize("n,sl,k,lim,seed,variant", [(3000, 100, 16, 150, 1001, 0), (1500, 150, 20, 120, 1003, 0),
                                                     (800, 200, 32, 80, 1004, 0), (1500, 100, 16, 100, 1002, 1),
                                                     (1200, 100, 10, 100, 7, 0)])
def test_closed_form_equals_seqan_model_on_workloads(built, n, sl, k, lim, seed, variant):
    """The literal search-scheme recursion against the closed form on the BASELINE workloads themselves
    (synthetic ONT-like reads, adapters through the error channel, the top-`lim` exact k-mers as queries,
    both ends): the realistic mix of exact, 1-edit and 2-edit hits, hits cut by the read border, N."""
    from approx_counter_b200 import host
    thr = orc.adjust_threshold(1.0, 16, k)
    for bot in (False, True):
        sample = host.synth_ends(seed, 0, n, sl, bot)
        codes, offs = orc.encode_matrix(sample)
        keys, cnts, _ = orc.count_kmers(codes, offs, k, thr)
        top, _ = orc.get_most_frequent(keys, cnts, lim, k)
        rng = np.random.default_rng(seed)
        extra = np.array([int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(10)], np.uint64)
        kmers = np.concatenate([top, extra])
        want = orc.error_count(codes, offs, kmers, k, fast=True)
        got, flags = orc.seqan_model_error_count(codes, offs, kmers, k, variant=variant, want_flags=True)
        assert np.array_equal(got, want)
        # R0 ⊆ R1 ⊆ R2 per k-mer, and every level is used by the workload
        assert (flags[:, 0, :] <= flags[:, 1, :]).all() and (flags[:, 1, :] <= flags[:, 2, :]).all()
        assert flags[:, 0, :].sum() > 0
        if k <= 20:  # the planted adapters (28 / 22 bases) are shorter than k = 32: only exact hits there
            assert (flags[:, 1, :] & ~flags[:, 0, :].astype(bool)).sum() > 0
            assert (flags[:, 2, :] & ~flags[:, 1, :].astype(bool)).sum() > 0


@pytest.mark.parametrize("n,sl,k,lim,seed", [(20000, 100, 16, 500, 1002), (8000, 150, 20, 400, 1003),
                                             (6000, 200, 32, 300, 1004), (12000, 60, 10, 300, 1005)])
def test_fm_index_model_equals_closed_form_at_scale(built, n, sl, k, lim, seed):
    """errorCount as the reference runs it — a bidirectional FM index of the sampled read
    ends searched with the literal optimal-search-scheme recursion (oracle/fm_index_model.cpp) — against the
    closed form (Myers scan) on BASELINE-like workloads, both ends, at sizes the occurrence-list model cannot
    reach: an independent second oracle."""
    from approx_counter_b200 import host
    thr = orc.adjust_threshold(1.0, 16, k)
    for bot in (False, True):
        sample = host.synth_ends(seed, 0, n, sl, bot)
        codes, offs = orc.encode_matrix(sample)
        keys, cnts, _ = orc.count_kmers(codes, offs, k, thr)
        top, _ = orc.get_most_frequent(keys, cnts, lim, k)
        rng = np.random.default_rng(seed)
        extra = np.array([int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(20)], np.uint64)
        kmers = np.concatenate([top, extra])
        want = orc.error_count(codes, offs, kmers, k, fast=True)
        got, (t_build, t_search) = orc.fm_index_error_count(codes, offs, kmers, k, want_seconds=True)
        assert np.array_equal(got, want)
        assert want.max() >= 3 and t_build > 0 and t_search > 0
        if k <= 20:
            assert want.max() > n // 2   # adapter k-mers hit most reads (k = 32 is longer than the planted adapters)


def test_fm_index_model_ragged_and_degenerate():
    rng = np.random.default_rng(77)
    reads = [b"", b"A", b"ACGTNNNNACGT", b"N" * 30, b"ACGT" * 12, b"TTTTTTTTTTTTTTTTTTTT", b"", b"ACGTACGAACGTACGT"]
    reads += [bytes(rng.choice(ACGT, size=int(rng.integers(0, 70)))) for _ in range(200)]
    codes, offs = orc.encode(reads)
    for k in (4, 5, 8, 12, 16, 25, 32):
        kmers = [orc.dna2int(("ACGT" * 8)[:k]), orc.dna2int("T" * k), 0] + \
                [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(5)]
        want = orc.error_count(codes, offs, kmers, k)
        assert np.array_equal(orc.fm_index_error_count(codes, offs, kmers, k), want), k
        assert np.array_equal(orc.fm_index_error_count(codes, offs, kmers, k, nb_thread=1), want), k
    codes, offs = orc.encode([])
    assert orc.fm_index_error_count(codes, offs, [5, 6], 8).tolist() == [0, 0]
    with pytest.raises(ValueError):
        orc.fm_index_error_count(codes, offs, [1], 3)      # the four blocks of the scheme need k >= 4


def _end_distances(pattern, text):
    """d[j] = min edit distance of `pattern` to a substring of `text` ENDING at position j (j = 0..len(text))."""
    m = len(pattern)
    col = list(range(m + 1))
    out = [col[m]]
    for ch in text:
        new = [0] * (m + 1)
        for i in range(1, m + 1):
            new[i] = min(col[i - 1] + (pattern[i - 1] != ch or ch == "N"), col[i] + 1, new[i - 1] + 1)
        col = new
        out.append(col[m])
    return out


def test_split_identity():
    """The identity a two-sided (meet-in-the-middle) scan would rest on: for a k-mer
    AB, min over substrings of d(AB, substring) = min over text positions j of
    [min d(A, substring ending at j)] + [min d(B, substring starting at j)] — the same reversal argument that lets
    suffix-sharing units walk the text backwards.  Checked against the oracle's minimum infix distance."""
    rng = np.random.default_rng(4242)
    for trial in range(60):
        k = int(rng.integers(6, 25))
        L = int(rng.integers(k - 3, 70))
        text = "".join(rng.choice(list("ACGTN"), size=L, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        kmer = "".join(rng.choice(list("ACGT"), size=k))
        if trial % 2:   # plant a noisy copy so that small distances occur
            m = list(kmer)
            for _ in range(int(rng.integers(0, 4))):
                p = int(rng.integers(0, len(m)))
                m[p] = str(rng.choice(list("ACGT")))
            m = "".join(m)[: max(0, L)]
            pos = int(rng.integers(0, max(1, L - len(m) + 1)))
            text = (text[:pos] + m + text[pos + len(m):])[:L]
        for cut in (k // 2, k // 3, k - 2):
            a, b = kmer[:cut], kmer[cut:]
            fwd = _end_distances(a, text)                              # A ending at j
            bwd = _end_distances(b[::-1], text[::-1])[::-1]            # B starting at j
            split = min(x + y for x, y in zip(fwd, bwd))
            codes, _ = orc.encode([text])
            want = orc.min_infix_distance(codes, orc.dna2int(kmer), k)  # clamped at 3
            assert min(split, 3) == want, (kmer, text, cut, split, want)


def test_get_most_frequent_fast_equals_full_sort():
    """The threshold-cut form used at full size (tests/test_gpu_fullsize.py, bench.py's reference arm) returns
    exactly what the comparator sort of all distinct k-mers returns, ties and all."""
    rng = np.random.default_rng(31)
    for k, n in ((16, 300_000), (20, 200_000), (5, 1024)):
        mask = np.uint64((1 << (2 * k)) - 1)
        keys = np.unique(rng.integers(0, 1 << 62, n).astype(np.uint64) & mask)
        cnts = np.ones(len(keys), np.uint64)
        hot = rng.choice(len(keys), min(500, len(keys) // 2), replace=False)
        cnts[hot] = rng.integers(1, 6, len(hot)).astype(np.uint64)
        for lim in (1, 50, 400, 2000):
            a = orc.get_most_frequent_fast(keys, cnts, lim, k)
            b = orc.get_most_frequent(keys, cnts, lim, k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_count_kmers_mt_equals_count_kmers():
    rng = np.random.default_rng(5)
    for k, n, L in ((16, 3000, 100), (20, 500, 151), (32, 300, 201), (4, 200, 30)):
        sample = orc.synth_ends(77 + k, 0, n, L, False)
        sample[rng.integers(0, n, 40), rng.integers(0, L, 40)] = ord("N")
        codes, offs = orc.encode_matrix(sample)
        thr = orc.adjust_threshold(1.0, 16, k)
        a = orc.count_kmers(codes, offs, k, thr)
        for threads in (1, 3, 8):
            b = orc.count_kmers_mt(codes, offs, k, thr, threads)
            assert a[2] == b[2]
            oa, ob = np.argsort(a[0]), np.argsort(b[0])
            assert np.array_equal(a[0][oa], b[0][ob]) and np.array_equal(a[1][oa], b[1][ob])
