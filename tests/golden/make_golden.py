"""Generate tests/golden/*.json.

The reference ships no tests or golden vectors and cannot be built here (SeqAn is
absent), so these fixtures are produced by the CPU restatement in oracle/ and
cross-checked, where small enough, against the literal SeqAn search-scheme model
(oracle/seqan_model.cpp) and, all of them, against the same recursion over an own
bidirectional FM index (oracle/fm_index_model.cpp).  PARITY UNPINNED — they pin the GPU path and the oracle
against regressions, not against a SeqAn build.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

ACGT = np.frombuffer(b"ACGT", np.uint8)
ADAPTER = b"AATGTACTTCGTTCAGTTACGTATTGCT"


def mutate(rng, s, n_edits):
    m = list(s)
    for _ in range(n_edits):
        op = rng.integers(0, 3)
        p = int(rng.integers(0, max(1, len(m))))
        if op == 0 and m:
            m[p] = int(rng.choice(ACGT))
        elif op == 1 and m:
            del m[p]
        else:
            m.insert(p, int(rng.choice(ACGT)))
    return bytes(m)


def approx_case(seed, n, L, k, n_kmers, with_model):
    rng = np.random.default_rng(seed)
    reads = []
    for r in range(n):
        body = bytearray(rng.choice(ACGT, size=int(L if r % 5 else rng.integers(0, L + 1))).tobytes())
        if r % 3 != 2 and len(body) >= len(ADAPTER) + 6:
            m = mutate(rng, ADAPTER, int(rng.integers(0, 4)))
            pos = int(rng.choice([0, len(body) - len(m), rng.integers(0, len(body) - len(m) + 1)]))
            body[pos:pos + len(m)] = m
        if r % 7 == 0 and body:
            body[int(rng.integers(0, len(body)))] = ord("N")
        reads.append(bytes(body).decode())
    src = ADAPTER * 2
    kmers = [orc.dna2int(src[i:i + k].decode()) for i in range(0, len(ADAPTER), 3)]
    kmers += [int(x) & ((1 << (2 * k)) - 1) for x in rng.integers(0, 1 << 62, n_kmers)]
    codes, offs = orc.encode(reads)
    slow = orc.error_count(codes, offs, kmers, k)
    fast = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(slow, fast)
    if with_model:
        model = orc.seqan_model_error_count(codes, offs, kmers, k)
        assert np.array_equal(model, slow), (model, slow)
    # the same recursion over a real bidirectional FM index (oracle/fm_index_model.cpp): every case, any size
    fm = orc.fm_index_error_count(codes, offs, kmers, k)
    assert np.array_equal(fm, slow), (fm, slow)
    return {"k": k, "reads": reads, "kmers": [orc.int2dna(v, k) for v in kmers],
            "counts": [int(c) for c in slow], "checked_against_seqan_model": bool(with_model),
            "checked_against_fm_index_model": True}


def exact_case(seed, n, L, k, param_lc, lim):
    rng = np.random.default_rng(seed)
    reads = []
    for r in range(n):
        body = bytearray(rng.choice(ACGT, size=L).tobytes())
        if r % 2 == 0:
            off = int(rng.integers(0, 8))
            body[off:off + len(ADAPTER)] = ADAPTER
        if r % 9 == 0:
            body[L // 2:] = (b"AC" * L)[: L - L // 2]
        if r % 11 == 0:
            body[int(rng.integers(0, L))] = ord("N")
        reads.append(bytes(body).decode())
    thr = orc.adjust_threshold(param_lc, 16, k)
    codes, offs = orc.encode(reads)
    keys, cnts, had_n = orc.count_kmers(codes, offs, k, thr)
    tk, tc = orc.get_most_frequent(keys, cnts, lim, k)
    approx = orc.error_count(codes, offs, tk, k, fast=True)
    ak, ac = orc.get_most_frequent(tk, approx, lim, k)
    return {"k": k, "param_lc": param_lc, "lim": lim, "reads": reads, "n_distinct": len(keys), "had_n": had_n,
            "exact": [[orc.int2dna(a, k), int(b)] for a, b in zip(tk, tc)],
            "approx": [[orc.int2dna(a, k), int(b)] for a, b in zip(ak, ac)]}


def main():
    out = {
        "appendix_b": [
            {"k": 8, "kmer": "ACGTTGCA", "total": 18,
             "reads": ["TTACGTTGCATT", "TTACGTAGCATT", "TTACGTGCATT", "TTACGTTTGCATT", "TTACCTTGGATT",
                       "TTACGNTGCATT", "GGGGGGGGGGGG", "ACGTTGCA", "CGTTGC", "ACGTTG", "TTAGGTTCCATT"],
             "per_read": [3, 2, 2, 2, 1, 2, 0, 3, 1, 1, 1]},
            {"k": 16, "kmer": "AATGTACTTCGTTCAG", "total": 10,
             "reads": ["GGAATGTACTTCGTTCAGTTACG", "GGAATGTACTTCGTCAGTTACGT", "AATGTACTTAGTTCAGCC",
                       "AATGAACTTAGTTCAGCC", "CCAATGTACTTTCGTTCAAGAA", "AATGTACTTCGTTC", "TTTTTTTTTTTTTTTTTTTT"],
             "per_read": [3, 2, 2, 1, 1, 1, 0]},
        ],
        "filter_thresholds": {"1.0": {"16": 28, "20": 58, "32": 258}, "1.5": {"16": 42, "20": 88, "32": 386}},
        "approx": [
            approx_case(1, 24, 30, 8, 6, True),
            approx_case(2, 20, 36, 12, 4, True),
            approx_case(3, 16, 44, 16, 3, True),
            approx_case(4, 300, 101, 16, 20, False),
            approx_case(5, 200, 151, 20, 12, False),
            approx_case(6, 150, 201, 32, 8, False),
            approx_case(7, 200, 100, 10, 12, False),
        ],
        "pipeline": [
            exact_case(11, 400, 100, 16, 1.0, 50),
            exact_case(12, 300, 101, 20, 1.0, 40),
            exact_case(13, 200, 120, 32, 1.5, 30),
            exact_case(14, 300, 60, 8, 1.0, 25),
        ],
    }
    for v in out["appendix_b"]:  # the hand-checkable vectors must also hold for the oracle
        codes, offs = orc.encode(v["reads"])
        km = [orc.dna2int(v["kmer"])]
        assert int(orc.error_count(codes, offs, km, v["k"])[0]) == v["total"]
        assert int(orc.seqan_model_error_count(codes, offs, km, v["k"])[0]) == v["total"]
        per = [3 - orc.min_infix_distance(orc.encode([r])[0], km[0], v["k"]) for r in v["reads"]]
        assert per == v["per_read"], per
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(HERE, "vectors.json"))


if __name__ == "__main__":
    main()
