"""Size-independent properties of the approximate count at BASELINE.json's full
sizes (where the oracle would take minutes): linearity over read shards, the
invariants of SURVEY.md §8c, idempotence, variant agreement."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu


def workload(n, sl, seed, bot=False):
    from approx_counter_b200 import host
    return host.synth_ends(seed, 0, n, sl, bot)


# BASELINE.json configs 2, 3 and 4 at their full read counts and lengths (fewer k-mers for C3/C4:
# the properties do not depend on how many k-mers are scanned)
@pytest.mark.parametrize("n,sl,k,lim,seed", [(100_000, 100, 16, 2000, 1002), (1_000_000, 150, 20, 2000, 1003),
                                             (1_000_000, 200, 32, 1000, 1004)])
def test_full_size_properties(counter, n, sl, k, lim, seed):
    from approx_counter_b200 import host
    sample = workload(n, sl, seed, bot=True)            # the `end` sample: sl+1 bases (:463)
    assert sample.shape == (n, sl + 1)
    thr = host.adjust_threshold(1.0, 16, k)
    counter.upload_sample(sample)
    km, ct, nd, hn = counter.count_kmers_topn(k, thr, lim)
    assert len(km) == lim and nd > lim
    assert all(ct[i] >= ct[i + 1] for i in range(lim - 1))
    total = counter.errorCount(km, k)
    again = counter.errorCount(km, k)
    assert np.array_equal(total, again)                  # idempotent
    # 3|R0| <= count <= 3n; exact window count >= |R0| >= count of reads ... (§8c)
    assert (total <= 3 * n).all()
    assert (total >= 3 * np.minimum(ct, 1)).all()
    # linearity: counts over disjoint read shards add up to the whole
    parts = np.zeros(lim, np.uint64)
    cuts = sorted({0, 32 * 1000, 32 * 1777 + 5, n // 2 + 7, n})  # cuts off the tile grid too
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        counter.upload_sample(np.ascontiguousarray(sample[lo:hi]))
        parts += counter.errorCount(km, k)
    assert np.array_equal(parts, total)
    # the resident sub-range option used by the multi-GPU binary gives the same split
    counter.upload_sample(sample)
    counter.set_queries(km, k)
    acc = np.zeros(lim, np.uint64)
    for lo, hi in ((0, 32 * 1500), (32 * 1500, n)):
        counter.set_option("scan_first_read", lo)
        counter.set_option("scan_n_reads", hi - lo)
        counter.scan()
        acc += counter.get_counts()
    counter.set_option("scan_first_read", 0)
    counter.set_option("scan_n_reads", -1)
    assert np.array_equal(acc, total)
    # spot check against the oracle on a slice the CPU finishes quickly
    r = 2000
    counter.upload_sample(np.ascontiguousarray(sample[:r]))
    got = counter.errorCount(km[:64], k)
    codes, offs = orc.encode_matrix(sample[:r])
    assert np.array_equal(got, orc.error_count(codes, offs, km[:64], k, fast=True))
    # every packing variant of the kernel returns the same counts
    for variant in (1, 2, 3, 6):
        try:
            counter.set_option("scan_variant", variant)
            v = counter.errorCount(km[:64], k)
        finally:
            counter.set_option("scan_variant", 0)
        assert np.array_equal(v, got)


def test_exact_stage_checksum_full_size(counter):
    """sum of all exact counts + filtered + N windows = all windows; checked through solid_km=1."""
    from approx_counter_b200 import host
    n, sl, k = 30_000, 100, 16
    sample = workload(n, sl, 1002)
    counter.upload_sample(sample)
    thr = 1e9                                             # filter off
    km, ct, nd, hn = counter.solid_kmers(k, thr, 1, capacity=1 << 24)
    assert len(km) == nd
    assert int(ct.sum()) + hn == n * (sl - k + 1)
    assert len(np.unique(km)) == len(km)
    km2, ct2, nd2, _ = counter.count_kmers_topn(k, thr, 100)
    assert np.array_equal(km2, km[:100]) and np.array_equal(ct2, ct[:100]) and nd2 == nd
