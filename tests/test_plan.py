"""Scan plan of the default (bit-sliced) kernel: apc_plan_queries groups k-mers that share a
prefix — or, reversed, a suffix — into units (host-side dynamic programme, no GPU needed)."""
import numpy as np
import pytest

from approx_counter_b200 import plan_queries

ACGT = b"ACGT"


def dna2int(s):
    v = 0
    for ch in s:
        v = (v << 2) | ACGT.index(ch)
    return v


def family(rng, k, n_var, where):
    """One random k-mer and n_var one-base variants whose difference lies in `where` (range of positions)."""
    base = bytes(rng.choice(np.frombuffer(ACGT, np.uint8), size=k))
    out = [dna2int(base)]
    for _ in range(n_var):
        m = bytearray(base)
        m[int(rng.integers(where[0], where[1]))] = int(rng.choice(np.frombuffer(ACGT, np.uint8)))
        out.append(dna2int(bytes(m)))
    return out


def check_plan(kmers, k):
    kmers = np.asarray(kmers, np.uint64)
    p = plan_queries(kmers, k)
    n = len(kmers)
    assert sorted(p["order"].tolist()) == list(range(n))          # a permutation: every k-mer scanned exactly once
    at = 0
    for s in range(len(p["units"])):
        t, g = int(p["shape_t"][s]), int(p["shape_g"][s])
        if g == 0:
            assert p["units"][s] == 0
            continue
        assert k - t >= 2 and t >= 1 and k - t + g * t <= 60   # kBsMaxRows
        for _ in range(int(p["units"][s])):
            idx = p["order"][at:at + g]
            rev = p["reversed"][at:at + g]
            assert len(set(rev.tolist())) == 1                      # one direction per unit
            vals = [int(kmers[i]) for i in idx]
            if rev[0]:   # members share their LAST k-t bases
                assert len({v & ((1 << (2 * (k - t))) - 1) for v in vals}) == 1
            else:        # members share their FIRST k-t bases
                assert len({v >> (2 * t) for v in vals}) == 1
            at += g
    assert not p["reversed"][at:].any()                            # ungrouped k-mers are scanned forwards
    return p


@pytest.mark.parametrize("k", range(2, 33))
def test_plan_is_a_valid_cover(k):
    rng = np.random.default_rng(900 + k)
    kmers = []
    for _ in range(12):
        kmers += family(rng, k, 9, (max(0, k - 4), k))              # late differences: prefix sharing
        kmers += family(rng, k, 9, (0, min(k, 4)))                  # early differences: suffix sharing
        kmers += family(rng, k, 5, (0, k))
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(40)]
    kmers += kmers[:5]                                              # duplicates
    rng.shuffle(kmers)
    p = check_plan(kmers, k)
    if k >= 8:
        grouped = int((p["units"] * p["shape_g"]).sum())
        assert grouped > len(kmers) // 2
        assert p["reversed"].sum() > 0 and p["reversed"].sum() < grouped


def test_plan_edge_cases():
    assert plan_queries([], 16)["units"].sum() == 0
    assert plan_queries([5], 16)["units"].sum() == 0
    p = plan_queries([7, 7, 7, 7], 16)                              # identical k-mers group like any others
    assert p["units"].sum() >= 1
    check_plan([7, 7, 7, 7], 16)
    check_plan([0, (1 << 64) - 1, 1 << 63, 12345], 32)
    assert plan_queries([1, 2, 3], 2)["units"].sum() == 0           # k = 2: no shape has two shared rows


def test_plan_random_kmers_stay_single():
    rng = np.random.default_rng(5)
    kmers = rng.integers(0, 1 << 62, 500).astype(np.uint64)
    p = check_plan(kmers, 31)
    assert (p["units"] * p["shape_g"]).sum() < 20
