"""GPU parity of the exact stage (K3-K5: count_kmers :487-519 + get_most_frequent
:396-405 / get_solid_kmers :372-388) against the CPU oracle, through the C ABI
(apc_exact_topn / apc_exact_solid)."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", np.uint8)
ADAPTER = b"AATGTACTTCGTTCAGTTACGTATTGCT"


def planted(rng, n, L, frac=0.6, p_n=0.001, low_complexity=True):
    s = rng.choice(ACGT, size=(n, L))
    for r in range(n):
        u = rng.random()
        if u < frac and L >= len(ADAPTER) + 8:
            off = int(rng.integers(0, 8))
            s[r, off:off + len(ADAPTER)] = np.frombuffer(ADAPTER, np.uint8)
        elif low_complexity and u > 0.9:
            unit = bytes(rng.choice(ACGT, size=int(rng.integers(1, 4))))
            rep = np.frombuffer((unit * L)[:L], np.uint8)
            a = int(rng.integers(0, L // 2))
            s[r, a:] = rep[a:]
    s[rng.random((n, L)) < p_n] = ord("N")
    return s


def oracle_topn(sample_codes, offs, k, thr, lim, forbidden=None):
    keys, cnts, had_n = orc.count_kmers(sample_codes, offs, k, thr, forbidden)
    tk, tc = orc.get_most_frequent(keys, cnts, lim, k)
    return tk, tc, len(keys), had_n


def check_topn(counter, sample, k, param_lc, lim, forbidden=None, ragged=None):
    thr = orc.adjust_threshold(param_lc, 16, k)
    if ragged is None:
        counter.upload_sample(sample)
        codes, offs = orc.encode_matrix(sample)
    else:
        counter.upload_sample(ragged)
        codes, offs = orc.encode(ragged)
    km, ct, nd, hn = counter.count_kmers_topn(k, thr, lim, forbidden)
    wk, wc, wd, wn = oracle_topn(codes, offs, k, thr, lim, forbidden)
    assert nd == wd
    assert hn == wn
    assert np.array_equal(km, wk)
    assert np.array_equal(ct, wc)
    return km, ct


@pytest.mark.parametrize("k", [2, 3, 4, 8, 12, 15, 16, 17, 20, 24, 31, 32])
def test_topn_all_k(counter, k):
    rng = np.random.default_rng(300 + k)
    sample = planted(rng, 400, 101, p_n=0.002)
    km, ct = check_topn(counter, sample, k, 1.0, 200)
    assert len(km) > 0


@pytest.mark.parametrize("lim", [0, 1, 7, 500, 5000, 10 ** 7])
def test_topn_limits_and_ties(counter, lim):
    # random reads: nearly every k-mer has count 1, so the survivors are decided by
    # (dimer score asc, k-mer value desc) over thousands of tied candidates (:283-301)
    rng = np.random.default_rng(17)
    sample = planted(rng, 600, 100, frac=0.3)
    check_topn(counter, sample, 16, 1.0, lim)


@pytest.mark.parametrize("param_lc", [0.0, 0.2, 0.5, 1.0, 1.5, 3.0, 1e9])
def test_low_complexity_thresholds(counter, param_lc):
    rng = np.random.default_rng(23)
    sample = planted(rng, 300, 100)
    for k in (10, 16, 20):
        check_topn(counter, sample, k, param_lc, 300)


def test_tie_boundary_inside_one_score_class(counter):
    # many distinct k-mers, all count 1 and a handful of dimer sums: the cut falls inside
    # a (count, score) class and must keep the LARGEST k-mer values
    rng = np.random.default_rng(5)
    sample = rng.choice(ACGT, size=(3000, 40))
    for lim in (1, 2, 3, 50, 1000, 2049, 20000):
        check_topn(counter, sample, 8, 10.0, lim)
        check_topn(counter, sample, 16, 1.0, lim)


def test_ragged_short_and_n_reads(counter):
    rng = np.random.default_rng(9)
    reads = [bytes(rng.choice(ACGT, size=int(rng.integers(0, 120)))) for _ in range(500)]
    reads[0] = b""
    reads[1] = b"ACGT"
    reads[2] = b"N" * 50
    reads[3] = ADAPTER
    reads[4] = ADAPTER[:15] + b"N" + ADAPTER[16:]
    check_topn(counter, None, 16, 1.0, 300, ragged=reads)
    check_topn(counter, None, 5, 1.0, 100, ragged=reads)


def test_forbidden_kmers(counter):
    rng = np.random.default_rng(31)
    sample = planted(rng, 300, 100)
    k = 16
    forb = np.array([orc.dna2int(ADAPTER[i:i + k].decode()) for i in range(0, 8)] + [0, 12345], np.uint64)
    km, ct = check_topn(counter, sample, k, 1.0, 100, forbidden=forb)
    assert not set(km.tolist()) & set(forb.tolist())


@pytest.mark.parametrize("solid", [1, 2, 5, 100, 10 ** 6])
def test_solid_kmers(counter, solid):
    rng = np.random.default_rng(41)
    sample = planted(rng, 300, 100)
    k = 16
    thr = orc.adjust_threshold(1.0, 16, k)
    counter.upload_sample(sample)
    km, ct, nd, hn = counter.solid_kmers(k, thr, solid, capacity=64)
    codes, offs = orc.encode_matrix(sample)
    keys, cnts, _ = orc.count_kmers(codes, offs, k, thr)
    wk, wc = orc.get_solid_kmers(keys, cnts, solid, k)
    assert nd == len(keys)
    assert np.array_equal(km, wk)
    assert np.array_equal(ct, wc)


def test_config1_shape_end_sample(counter):
    """10k reads x 101 (the `end` sample is sl+1 bases, :463), k=16, lim=500."""
    rng = np.random.default_rng(77)
    sample = planted(rng, 10000, 101, frac=0.9, p_n=1e-4, low_complexity=False)
    km, ct = check_topn(counter, sample, 16, 1.0, 500)
    assert ct[0] > 1000
    t = counter.timing()
    assert t["exact_ms"] > 0 and t["exact_launches"] >= 5


def test_empty_sample(counter):
    counter.upload_sample(np.zeros((0, 50), np.uint8))
    km, ct, nd, hn = counter.count_kmers_topn(16, 1.0, 10)
    assert len(km) == 0 and nd == 0 and hn == 0
    counter.upload_sample(np.full((5, 10), ord("A"), np.uint8))  # reads shorter than k
    km, ct, nd, hn = counter.count_kmers_topn(16, 1.0, 10)
    assert len(km) == 0 and nd == 0 and hn == 0
