"""Randomised GPU-vs-oracle parity over the whole parameter space of the path: every k,
every packing variant, random read lengths (ragged and uniform), N density, planted
near-matches at read borders.  Seeds are fixed: failures reproduce."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

ACGT = np.frombuffer(b"ACGT", np.uint8)


def edits(rng, s, n):
    m = bytearray(s)
    for _ in range(n):
        op, p = int(rng.integers(0, 3)), int(rng.integers(0, max(1, len(m))))
        if op == 0 and m:
            m[p] = int(rng.choice(ACGT))
        elif op == 1 and len(m) > 1:
            del m[p]
        else:
            m.insert(p, int(rng.choice(ACGT)))
    return bytes(m)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_scan_against_oracle(counter, seed):
    rng = np.random.default_rng(9000 + seed)
    k = int(rng.integers(2, 33))
    n = int(rng.integers(1, 400))
    uniform = bool(rng.integers(0, 2))
    Lmax = int(rng.integers(1, 260))
    p_n = float(rng.choice([0.0, 0.001, 0.05]))
    needles = [bytes(rng.choice(ACGT, size=k)) for _ in range(int(rng.integers(1, 6)))]
    reads = []
    for r in range(n):
        L = Lmax if uniform else int(rng.integers(0, Lmax + 1))
        body = bytearray(rng.choice(ACGT, size=L).tobytes())
        if L and rng.random() < 0.6:
            m = edits(rng, needles[int(rng.integers(0, len(needles)))], int(rng.integers(0, 4)))[:L]
            pos = int(rng.choice([0, L - len(m), rng.integers(0, L - len(m) + 1)]))
            body[pos:pos + len(m)] = m
        for i in range(L):
            if rng.random() < p_n:
                body[i] = ord("N")
        reads.append(bytes(body))
    kmers = [orc.dna2int(x.decode()) for x in needles]
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(int(rng.integers(0, 40)))]
    kmers = np.array(kmers, np.uint64)
    if uniform:
        sample = np.frombuffer(b"".join(reads), np.uint8).reshape(n, Lmax).copy()
        counter.upload_sample(sample)
    else:
        counter.upload_sample(reads)
    codes, offs = orc.encode(reads)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    variants = [0, 1] + ([2] if k <= 16 else []) + ([3] if k <= 10 else []) + ([6] if 12 <= k <= 21 else [])
    for v in variants:
        for tpj in (0, 1, 5):
            counter.set_option("scan_variant", v)
            counter.set_option("tiles_per_job", tpj)
            try:
                got = counter.errorCount(kmers, k)
            finally:
                counter.set_option("scan_variant", 0)
                counter.set_option("tiles_per_job", 0)
            assert np.array_equal(got, want), (seed, k, n, uniform, Lmax, v, tpj)
    counter.set_option("shape_mask", 0x3FFFFF)   # all 22 unit shapes (the default plan of a small sample uses six)
    try:
        got = counter.errorCount(kmers, k)
    finally:
        counter.set_option("shape_mask", 0xFFFFFFFF)
    assert np.array_equal(got, want), (seed, k, n, uniform, Lmax, "all shapes")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_exact_against_oracle(counter, seed):
    rng = np.random.default_rng(7000 + seed)
    k = int(rng.integers(2, 33))
    n = int(rng.integers(1, 600))
    L = int(rng.integers(1, 200))
    alphabet = ACGT[: int(rng.integers(1, 5))]          # small alphabets: heavy ties and low complexity
    sample = rng.choice(alphabet, size=(n, L))
    sample[rng.random((n, L)) < float(rng.choice([0.0, 0.01]))] = ord("N")
    lim = int(rng.choice([0, 1, 3, 50, 1000, 10 ** 6]))
    param_lc = float(rng.choice([0.3, 1.0, 1.5, 100.0]))
    thr = orc.adjust_threshold(param_lc, 16, k)
    counter.upload_sample(sample)
    km, ct, nd, hn = counter.count_kmers_topn(k, thr, lim)
    codes, offs = orc.encode_matrix(sample)
    keys, cnts, had_n = orc.count_kmers(codes, offs, k, thr)
    if k > 2:   # k == 2: the reference comparator is not a strict weak order (NaN score), order of ties undefined
        wk, wc = orc.get_most_frequent(keys, cnts, lim, k)
        assert np.array_equal(km, wk) and np.array_equal(ct, wc), (seed, k, n, L, lim, param_lc)
    else:
        assert sorted(ct.tolist(), reverse=True) == sorted(cnts.tolist(), reverse=True)[:lim]
    assert nd == len(keys) and hn == had_n
