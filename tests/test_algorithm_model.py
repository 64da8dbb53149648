"""CPU model of the bit-sliced scan (csrc/bitslice_core.cuh) driven by the REAL scan plan (apc_plan_queries):
reads packed into (Python) integers one bit per read, units of k-mers sharing a trunk, suffix units stored
reversed and walking the columns backwards, the sticky hit row, and dead-row skipping with its two conditions.
It restates the kernel's algorithm, not its code, and must give the oracle's counts: this is the part of the
hot path the CPU suite can check (the kernels themselves are checked by the -m gpu tests)."""
import numpy as np
import pytest

from approx_counter_b200 import plan_queries
from oracle import orc

ACGT = "ACGT"


def check_row(k):  # bs_check_row_host (apc_internal.h)
    return 14 if k >= 18 else 13 if k >= 16 else k


def reverse_kmer(v, k):
    r = 0
    for i in range(k):
        r = (r << 2) | ((v >> (2 * i)) & 3)
    return r


def bases(v, k):
    return [(v >> (2 * (k - 1 - i))) & 3 for i in range(k)]


class Chain:
    """Rows [first, first + n) of one k-mer: three levels per row, each an integer with one bit per read."""

    def __init__(self, rows, first, hit, all_ones):
        self.rows, self.first, self.hit, self.ALL = rows, first, hit, all_ones
        self.r = [[0, all_ones if first + j < 1 else 0, all_ones if first + j < 2 else 0] for j in range(len(rows))]

    def run(self, j0, j1, carry, eq):
        """The five LOP3 of rows j0..j1-1 for one column; carry = (p0, p1, p2, n0p, n1p) of the row above."""
        p0, p1, p2, n0p, n1p = carry
        for j in range(j0, j1):
            i, e = self.first + j, eq[self.rows[j]]
            o0, o1, o2 = self.r[j]
            if self.hit and j == len(self.rows) - 1:      # sticky hit row
                n0 = (p0 & e) | o0
                n1 = self.ALL if i < 1 else (p1 & e) | o1 | p0 | n0p
                n2 = self.ALL if i < 2 else (p2 & e) | o2 | p1 | n1p
            else:
                n0 = e if i < 1 else p0 & e
                n1 = self.ALL if i < 1 else (p1 & e) | o0 | p0 | n0p
                n2 = self.ALL if i < 2 else (p2 & e) | o1 | p1 | n1p
            self.r[j] = [n0, n1, n2]
            p0, p1, p2, n0p, n1p = o0, o1, o2, n0, n1
        return (p0, p1, p2, n0p, n1p)


def scan_unit(kmers, k, t, reverse, eq_cols, n_reads, stats):
    """One unit: len(kmers) members sharing their first k - t bases (t == k: a single k-mer).  Returns the
    members' counts.  eq_cols[c][b] = reads whose base in column c is b (N and padding: in none)."""
    ALL = (1 << n_reads) - 1
    g, p, m = len(kmers), k - t, check_row(k)
    b = [bases(v, k) for v in kmers]
    assert all(x[:p] == b[0][:p] for x in b)                 # the plan's promise
    trunk = Chain(b[0][:p], 0, False, ALL)
    tails = [Chain(x[p:], p, True, ALL) for x in b]
    deep_zero = True
    cols = range(len(eq_cols) - 1, -1, -1) if reverse else range(len(eq_cols))
    for c in cols:
        eq = eq_cols[c]
        carry = (ALL,) * 5
        if m >= k:                                            # no split
            carry = trunk.run(0, p, carry, eq)
            for tl in tails:
                tl.run(0, t, carry, eq)
            continue
        tm, um = min(m, p), max(0, m - p)                     # the split in trunk and tail coordinates
        before = trunk.r[m - 1][2] if m <= p else 0
        for tl in tails:
            if m > p:
                before |= tl.r[um - 1][2]
        carry = trunk.run(0, tm, carry, eq)
        cg = [carry] * g
        if m > p:
            cg = [tl.run(0, um, carry, eq) for tl in tails]
        after = trunk.r[m - 1][2] if m <= p else 0
        for tl in tails:
            if m > p:
                after |= tl.r[um - 1][2]
        run = (before | after) != 0                           # the warp vote (a): can anything reach the deep rows?
        if not run and not deep_zero:                         # (b): have the deep rows (hit rows apart) drained?
            z = 0
            for j in range(tm, p):
                z |= trunk.r[j][2]
            for tl in tails:
                for j in range(um, t - 1):
                    z |= tl.r[j][2]
            run = z != 0
        deep_zero = not run
        stats[0] += run
        stats[1] += 1
        if run:
            carry = trunk.run(tm, p, carry, eq)
            for i, tl in enumerate(tails):
                tl.run(um, t, carry if m <= p else cg[i], eq)
    return [sum(bin(level).count("1") for level in tl.r[-1]) for tl in tails]


def model_error_count(reads, kmers, k):
    n, L = len(reads), max((len(r) for r in reads), default=0)
    eq_cols = [[0, 0, 0, 0] for _ in range(L)]
    for r, s in enumerate(reads):
        for c, ch in enumerate(s):
            if ch in ACGT:
                eq_cols[c][ACGT.index(ch)] |= 1 << r
    plan = plan_queries(kmers, k)
    out = np.zeros(len(kmers), np.uint64)
    stats = [0, 0]
    at = 0
    for s in range(len(plan["units"])):
        g, t = int(plan["shape_g"][s]), int(plan["shape_t"][s])
        for _ in range(int(plan["units"][s])):
            idx = plan["order"][at:at + g]
            rev = bool(plan["reversed"][at])
            members = [reverse_kmer(int(kmers[i]), k) if rev else int(kmers[i]) for i in idx]
            for i, cnt in zip(idx, scan_unit(members, k, t, rev, eq_cols, n, stats)):
                out[i] = cnt
            at += g
    for i in plan["order"][at:]:                              # ungrouped k-mers: one chain, walked forwards
        out[i] = scan_unit([int(kmers[i])], k, k, False, eq_cols, n, stats)[0]
    return out, plan, stats


def make_case(rng, k, n_reads):
    base = "".join(rng.choice(list(ACGT), size=k + 8))
    reads = []
    for r in range(n_reads):
        L = int(rng.integers(max(1, k - 3), 64))
        s = list(rng.choice(list(ACGT), size=L))
        if r % 3 != 2:
            m = list(base)
            for _ in range(int(rng.integers(0, 4))):
                op, pos = int(rng.integers(0, 3)), int(rng.integers(0, len(m)))
                if op == 0:
                    m[pos] = str(rng.choice(list(ACGT)))
                elif op == 1 and len(m) > 1:
                    del m[pos]
                else:
                    m.insert(pos, str(rng.choice(list(ACGT))))
            m = m[:L]
            pos = int(rng.choice([0, L - len(m), int(rng.integers(0, L - len(m) + 1))]))
            s[pos:pos + len(m)] = m
        if r % 9 == 0:
            s[int(rng.integers(0, L))] = "N"
        reads.append("".join(s))
    reads[5] = ""
    kmers = []
    for off in range(0, 8, 2):
        w = base[off:off + k]
        kmers.append(orc.dna2int(w))
        for lo, hi in ((k - 3, k), (0, 3), (k // 2 - 2, k // 2 + 2)):   # late, early and middle differences
            for _ in range(5):
                v = list(w)
                v[int(rng.integers(max(0, lo), min(k, hi)))] = str(rng.choice(list(ACGT)))
                kmers.append(orc.dna2int("".join(v)))
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(6)]
    return reads, np.array(kmers, np.uint64)


@pytest.mark.parametrize("k", [3, 5, 8, 12, 15, 16, 17, 18, 20, 25, 32])
def test_model_of_the_scan_equals_the_oracle(built, k):
    rng = np.random.default_rng(31000 + k)
    reads, kmers = make_case(rng, k, 150)
    got, plan, stats = model_error_count(reads, kmers, k)
    codes, offs = orc.encode(reads)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)
    if k >= 8:
        assert (plan["units"] > 0).sum() >= 2 and plan["reversed"].sum() > 0   # units in both directions were exercised
    if k >= 16:
        assert 0 < stats[0] < stats[1]                                          # deep rows were both computed and skipped
