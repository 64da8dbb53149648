"""GPU parity of the approximate-count kernel (K1) against the CPU oracle,
through the C ABI (apc_upload_sample / apc_approx_count)."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu

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


def make_sample(rng, n, L, k, p_n=0.002):
    sample = rng.choice(ACGT, size=(n, L))
    for r in range(n):
        if rng.random() < 0.7:
            m = mutate(rng, ADAPTER, int(rng.integers(0, 5)))[: L]
            pos = int(rng.integers(0, L - len(m) + 1)) if rng.random() < 0.7 else int(rng.choice([0, L - len(m)]))
            sample[r, pos:pos + len(m)] = np.frombuffer(m, np.uint8)
    mask = rng.random((n, L)) < p_n
    sample[mask] = ord("N")
    return sample


def make_kmers(rng, k, n_random):
    ks = []
    src = (ADAPTER * 3)
    for i in range(0, len(ADAPTER)):
        ks.append(orc.dna2int(src[i:i + k].decode()))
    for _ in range(n_random):
        ks.append(int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1))
    return np.array(ks, np.uint64)


@pytest.mark.parametrize("k", [4, 8, 10, 11, 12, 15, 16, 17, 20, 21, 22, 27, 31, 32])
def test_parity_all_k(counter, k):
    rng = np.random.default_rng(100 + k)
    n, L = 257, 101
    sample = make_sample(rng, n, L, k)
    kmers = make_kmers(rng, k, 37)
    counter.set_option("scan_variant", 0)
    counter.upload_sample(sample)
    got = counter.errorCount(kmers, k)
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)
    assert got.max() > 0


@pytest.mark.parametrize("k,variant", [(8, 1), (8, 2), (8, 3), (16, 1), (16, 2), (12, 6), (16, 6), (20, 1), (20, 6)])
def test_variants_agree(counter, k, variant):
    rng = np.random.default_rng(7 * k + variant)
    sample = make_sample(rng, 200, 64, k)
    kmers = make_kmers(rng, k, 11)
    counter.upload_sample(sample)
    counter.set_option("scan_variant", variant)
    try:
        got = counter.errorCount(kmers, k)
    finally:
        counter.set_option("scan_variant", 0)
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("L", [1, 7, 15, 16, 17, 31, 32, 33, 100, 101, 150, 151, 200, 201])
def test_read_lengths(counter, L):
    rng = np.random.default_rng(L)
    k = min(16, max(2, L))
    sample = make_sample(rng, 70, L, k) if L >= len(ADAPTER) else rng.choice(ACGT, size=(70, L))
    kmers = make_kmers(rng, k, 5)
    counter.upload_sample(sample)
    got = counter.errorCount(kmers, k)
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)


def test_ragged_and_empty_reads(counter):
    rng = np.random.default_rng(5)
    k = 12
    reads = []
    for i in range(150):
        L = int(rng.integers(0, 90))
        reads.append(bytes(rng.choice(ACGT, size=L)))
    reads[3] = ADAPTER
    reads[4] = b""
    reads[5] = ADAPTER[:11]
    reads[6] = b"N" * 40
    kmers = make_kmers(rng, k, 9)
    counter.upload_sample(reads)
    got = counter.errorCount(kmers, k)
    codes, offs = orc.encode(reads)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)


def test_appendix_b_vectors(counter):
    reads = ["TTACGTTGCATT", "TTACGTAGCATT", "TTACGTGCATT", "TTACGTTTGCATT", "TTACCTTGGATT", "TTACGNTGCATT",
             "GGGGGGGGGGGG", "ACGTTGCA", "CGTTGC", "ACGTTG", "TTAGGTTCCATT"]
    counter.upload_sample(reads)
    assert counter.errorCount([orc.dna2int("ACGTTGCA")], 8).tolist() == [18]
    reads = ["GGAATGTACTTCGTTCAGTTACG", "GGAATGTACTTCGTCAGTTACGT", "AATGTACTTAGTTCAGCC", "AATGAACTTAGTTCAGCC",
             "CCAATGTACTTTCGTTCAAGAA", "AATGTACTTCGTTC", "TTTTTTTTTTTTTTTTTTTT"]
    counter.upload_sample(reads)
    assert counter.errorCount([orc.dna2int("AATGTACTTCGTTCAG")], 16).tolist() == [10]


def test_no_reads_no_kmers(counter):
    counter.upload_sample(np.zeros((0, 50), np.uint8))
    assert counter.errorCount([1, 2, 3], 8).tolist() == [0, 0, 0]
    counter.upload_sample(np.full((10, 50), ord("A"), np.uint8))
    assert len(counter.errorCount([], 8)) == 0


def test_many_kmers_many_reads_split_phase(counter):
    """config-1 shape (10k reads x 100, k=16, 500 k-mers) through the resident API."""
    rng = np.random.default_rng(11)
    n, L, k = 10000, 100, 16
    sample = make_sample(rng, n, L, k, p_n=1e-4)
    kmers = make_kmers(rng, k, 500 - len(ADAPTER))
    counter.upload_sample(sample)
    counter.set_queries(kmers, k)
    counter.scan()
    got = counter.get_counts()
    counter.scan()  # idempotent: counts are overwritten, not accumulated
    again = counter.get_counts()
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    assert np.array_equal(got, want)
    assert np.array_equal(again, want)
    t = counter.timing()
    assert t["scan_ms"] > 0 and t["scan_launches"] in (1, 2)


def test_async_api_two_contexts(built):
    """apc_upload_sample_async + apc_approx_count_async on two contexts (starts / ends),
    page-locked buffers, results valid after apc_sync; repeated with changing inputs."""
    import torch
    from approx_counter_b200 import ApproxCounter
    rng = np.random.default_rng(21)
    k = 16
    ctxs = [ApproxCounter(0), ApproxCounter(0)]
    try:
        for rep in range(3):
            samples = [make_sample(rng, 1500 + 100 * rep, 100 + i, k) for i in range(2)]
            kmers = [make_kmers(rng, k, 40 + rep) for _ in range(2)]
            pin_s = [torch.from_numpy(s).pin_memory() for s in samples]
            pin_q = [torch.from_numpy(q.view(np.int64)).pin_memory() for q in kmers]
            pin_o = [torch.zeros(len(q), dtype=torch.int64).pin_memory() for q in kmers]
            for c, s, q, o in zip(ctxs, pin_s, pin_q, pin_o):
                c.upload_sample_ptr_async(s.data_ptr(), s.shape[0], s.shape[1])
                c.errorCount_ptr_async(q.data_ptr(), q.numel(), k, o.data_ptr())
            for c in ctxs:
                c.sync()
            for s, q, o in zip(samples, kmers, pin_o):
                codes, offs = orc.encode_matrix(s)
                want = orc.error_count(codes, offs, q, k, fast=True)
                assert np.array_equal(o.numpy().view(np.uint64), want)
            t = ctxs[0].timing()
            assert t["upload_ms"] > 0 and t["total_ms"] > 0
    finally:
        for c in ctxs:
            c.close()


def test_abi_error_codes(built):
    """Status codes of the C ABI on a real device: stages out of order, bad arguments."""
    import ctypes as C
    from approx_counter_b200 import ApcError, ApproxCounter, load
    lib = load()
    assert lib.apc_device_count() >= 1
    h = C.c_void_p()
    assert lib.apc_create(10 ** 6, C.byref(h)) == -3                     # APC_ERR_NO_DEVICE
    with ApproxCounter(0) as c:
        with pytest.raises(ApcError, match="NO_SAMPLE"):
            c.errorCount([1, 2], 8)                                      # scan before upload
        with pytest.raises(ApcError, match="NO_SAMPLE"):
            c.count_kmers_topn(8, 1.0, 5)
        c.upload_sample(np.full((4, 20), ord("A"), np.uint8))
        with pytest.raises(ApcError, match="NO_QUERIES"):
            c.scan()                                                     # scan before set_queries
        with pytest.raises(ApcError, match="INVALID"):
            c.errorCount([1], 1)                                         # k < 2 (:781)
        with pytest.raises(ApcError, match="INVALID"):
            c.errorCount([1], 33)                                        # k > 32
        with pytest.raises(ApcError, match="INVALID"):
            c.errorCount([1 << 16], 8)                                   # value wider than 2k bits
        with pytest.raises(ApcError, match="INVALID"):
            c.set_option("no_such_option", 1)
        with pytest.raises(ApcError, match="INVALID"):
            c.set_option("scan_variant", 5)
        assert c.errorCount([0], 8).tolist() == [3 * 4]                  # still usable after the errors: AAAAAAAA
        n, ml, tb = c.sample_info()
        assert (n, ml, tb) == (4, 20, 80)


@pytest.mark.parametrize("k", [4, 5, 9, 16, 17, 20, 25, 32])
def test_prefix_sharing_kmers_are_paired_correctly(counter, k):
    """The default kernel scans two k-mers with a common prefix >= k/2 in one warp (shared rows
    computed once).  Query sets full of such neighbours — one-error variants, duplicates, k-mers
    differing only in the last base — must give the same counts as one k-mer per warp (variant 7),
    as the row-packed kernel (variant 8) and as the oracle."""
    rng = np.random.default_rng(500 + k)
    sample = make_sample(rng, 700, 120, min(k, 16))
    base = rng.choice(ACGT, size=k).tobytes()
    sample[::5, 30:30 + k] = np.frombuffer(base, np.uint8)
    kmers = []
    for i in range(60):
        m = bytearray(base)
        pos = int(rng.integers(k // 2, k)) if i % 3 else int(rng.integers(0, k))   # mostly late differences
        m[pos] = int(rng.choice(ACGT))
        kmers.append(orc.dna2int(bytes(m).decode()))
    kmers += [orc.dna2int(base.decode())] * 3                                     # duplicates
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(15)]
    kmers = np.array(kmers, np.uint64)
    rng.shuffle(kmers)
    counter.upload_sample(sample)
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    for variant in (0, 7, 8):
        counter.set_option("scan_variant", variant)
        try:
            got = counter.errorCount(kmers, k)
            t = counter.timing()
        finally:
            counter.set_option("scan_variant", 0)
        assert np.array_equal(got, want), variant
        if variant == 0:
            assert 2 <= t["scan_launches"] <= 13  # one launch per unit shape in use + the ungrouped k-mers


@pytest.mark.parametrize("k", range(3, 33))
def test_units_of_every_shape_and_direction(counter, k):
    """The default kernel scans units of k-mers sharing a prefix (forwards) or a suffix (backwards over the
    text).  Query sets built to fill every unit shape in both directions, over ragged reads of odd and
    even lengths with N, must match the oracle; so must a sub-range scan of the same plan."""
    from approx_counter_b200 import plan_queries
    rng = np.random.default_rng(7000 + k)
    base = rng.choice(ACGT, size=k).tobytes()
    reads = []
    for r in range(1500):
        L = int(rng.integers(max(1, k - 3), 90))
        row = rng.choice(ACGT, size=L)
        if r % 3 == 0 and L >= k:
            m = mutate(rng, base, int(rng.integers(0, 4)))[:L]
            pos = int(rng.choice([0, L - len(m), int(rng.integers(0, L - len(m) + 1))]))   # flush with both borders too
            row[pos:pos + len(m)] = np.frombuffer(m, np.uint8)
        if r % 11 == 0:
            row[int(rng.integers(0, L))] = ord("N")
        reads.append(row.tobytes())
    reads[17] = b""
    kmers = []
    for lo, hi in ((max(0, k - 3), k), (0, min(3, k)), (max(0, k - 8), k), (0, min(8, k)), (0, k)):
        for _ in range(4):
            b2 = mutate(rng, base, 1)[:k].ljust(k, b"A")
            for _ in range(10):
                m = bytearray(b2)
                for _ in range(int(rng.integers(1, 3))):
                    m[int(rng.integers(lo, hi))] = int(rng.choice(ACGT))
                kmers.append(orc.dna2int(bytes(m).decode()))
    kmers += [orc.dna2int(base.decode())] * 2
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(10)]
    kmers = np.array(kmers, np.uint64)
    rng.shuffle(kmers)
    plan = plan_queries(kmers, k)
    if k >= 10:
        assert (plan["units"] > 0).sum() >= 3 and 0 < plan["reversed"].sum() < len(kmers)
    counter.set_option("scan_variant", 0)
    counter.upload_sample(reads)
    codes, offs = orc.encode(reads)
    want = orc.error_count(codes, offs, kmers, k, fast=True)
    # a sample this small is planned with the small shapes only (kBsSmallShapes); an explicit mask of all 22
    # shapes switches that rule off, so both plans are checked
    try:
        for mask in (0xFFFFFFFF, 0x3FFFFF):
            counter.set_option("shape_mask", mask)
            got = counter.errorCount(kmers, k)
            assert np.array_equal(got, want), hex(mask)
    except Exception:
        counter.set_option("shape_mask", 0xFFFFFFFF)
        raise
    # sub-range of the resident sample (what the multi-GPU binary does), with every shape
    lo, hi = 32 * 7, 32 * 7 + 1001
    counter.set_queries(kmers, k)
    try:
        counter.set_option("scan_first_read", lo)
        counter.set_option("scan_n_reads", hi - lo)
        counter.scan()
        part = counter.get_counts()
    finally:
        counter.set_option("scan_first_read", 0)
        counter.set_option("scan_n_reads", -1)
        counter.set_option("shape_mask", 0xFFFFFFFF)
    codes, offs = orc.encode(reads[lo:hi])
    assert np.array_equal(part, orc.error_count(codes, offs, kmers, k, fast=True))


@pytest.mark.parametrize("k", [16, 17, 18, 20, 24, 27, 32])
def test_deep_rows_wake_and_drain(counter, k):
    """Dead-row skipping (rows >= 13/14 are computed only while level 2 of the row above is set for some read
    of the warp, and until the deep rows have drained): reads made of adapter copies separated by gaps of every
    length — partial copies, copies cut by the read end, back-to-back copies — wake and drain the deep rows over
    and over; repetitive reads keep them awake all the time; a sample without any copy never wakes them."""
    rng = np.random.default_rng(9100 + k)
    motif = (ADAPTER * 2)[:max(k + 6, 28)]
    kmers = []
    for i in range(0, len(motif) - k + 1, 2):
        w = motif[i:i + k]
        kmers.append(orc.dna2int(w.decode()))
        for _ in range(6):
            kmers.append(orc.dna2int(mutate(rng, w, int(rng.integers(1, 3)))[:k].ljust(k, b"C").decode()))
    kmers += [orc.dna2int("A" * k), orc.dna2int(("AC" * k)[:k]), orc.dna2int(("ACG" * k)[:k])]
    kmers += [int.from_bytes(rng.bytes(8), "little") & ((1 << (2 * k)) - 1) for _ in range(8)]
    kmers = np.array(kmers, np.uint64)

    def gappy_read(L):
        out = bytearray()
        while len(out) < L:
            out += rng.choice(ACGT, size=int(rng.integers(0, 45))).tobytes()
            cut = int(rng.integers(k - 4, len(motif) + 1))
            start = int(rng.integers(0, len(motif) - cut + 1))
            out += mutate(rng, motif[start:start + cut], int(rng.integers(0, 4)))
        return bytes(out[:L])

    samples = {
        "gappy": [gappy_read(int(rng.integers(150, 320))) for _ in range(1300)],
        "repetitive": [(b"A" * 200), (b"AC" * 100), (b"ACG" * 70), motif * 6, b"N" * 50 + motif * 3] * 40
                      + [gappy_read(200) for _ in range(100)],
        "quiet": [rng.choice(ACGT, size=int(rng.integers(60, 160))).tobytes() for _ in range(1100)],
    }
    for name, reads in samples.items():
        counter.set_option("scan_variant", 0)
        counter.upload_sample(reads)
        got = counter.errorCount(kmers, k)
        codes, offs = orc.encode(reads)
        want = orc.error_count(codes, offs, kmers, k, fast=True)
        assert np.array_equal(got, want), name
        counter.set_option("shape_mask", 0)            # one k-mer per warp: the same split of the rows, no units
        try:
            assert np.array_equal(counter.errorCount(kmers, k), want), name
        finally:
            counter.set_option("shape_mask", 0xFFFFFFFF)


def test_nccl_communicator_through_the_abi_single_rank(built):
    """apc_comm_unique_id / apc_comm_init_rank / apc_scan_allreduce with a communicator of ONE rank: libnccl is
    loaded by libapc itself and the all-reduce really runs on the context's stream — the driver's single-GPU
    box cannot run the multi-rank tests (tests/test_gpu_multi.py), this keeps the ABI path under its eyes."""
    from approx_counter_b200 import ApproxCounter
    rng = np.random.default_rng(12)
    sample = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(3000, 100))
    sample[::3, 7:23] = np.frombuffer(b"AATGTACTTCGTTCAG", np.uint8)
    kmers = np.array([orc.dna2int("AATGTACTTCGTTCAG"), orc.dna2int("AATGTACTTCGTTCAT"), 99, 12345], np.uint64)
    codes, offs = orc.encode_matrix(sample)
    want = orc.error_count(codes, offs, kmers, 16, fast=True)
    with ApproxCounter(0) as c:
        assert c.comm_info() == (0, 1)
        c.allreduce_counts()                         # no communicator: a no-op, not an error
        c.comm_init_rank(1, 0, ApproxCounter.comm_unique_id())
        assert c.comm_info() == (0, 1)
        c.upload_sample(sample)
        c.set_queries(kmers, 16)
        for _ in range(3):                           # direct launches, graph capture, graph replay
            c.scan_allreduce()
            assert np.array_equal(c.get_counts(), want)
        c.comm_destroy()
        c.scan_allreduce()
        assert np.array_equal(c.get_counts(), want)


def test_scan_graph_replay_and_statistics(built):
    """A scan issued again unchanged is captured in a CUDA graph and replayed; any change (queries, sample,
    options, destination) falls back to direct launches.  The executed-row tally of the kernels is the same
    whichever way the scan was launched and lies between the top rows and the whole plan."""
    from approx_counter_b200 import ApproxCounter
    sample = np.ascontiguousarray(orc.synth_ends(4321, 0, 5000, 100, False))
    codes, offs = orc.encode_matrix(sample)
    adapter = "AATGTACTTCGTTCAGTTACGTATTGCT"
    kmers = sorted({orc.dna2int(adapter[i:i + 16][:p] + c + adapter[i:i + 16][p + 1:])
                    for i in range(8) for p in range(16) for c in "ACGT"})
    kmers = np.array(kmers, np.uint64)
    want = orc.error_count(codes, offs, kmers, 16, fast=True)
    per_scan = []
    for graph in (1, 0):
        with ApproxCounter(0) as c:
            c.set_option("scan_graph", graph)
            c.upload_sample(sample)
            c.set_queries(kmers, 16)
            c.scan_stats()
            for i in range(4):
                c.scan()
                assert np.array_equal(c.get_counts(), want), (graph, i)
            st = c.scan_stats()
            assert st["scans"] == 4
            assert st["lop3_top"] < st["lop3_executed"] < st["lop3_planned"] < st["lop3_one_kmer_per_warp"]
            per_scan.append(st["lop3_executed"] / 4)
            # a change between scans: other queries, then the first set again
            c.set_queries(kmers[::-1].copy(), 16)
            c.scan()
            assert np.array_equal(c.get_counts(), want[::-1])
            c.set_option("scan_n_reads", 2048)
            c.scan()
            c.scan()
            part = orc.error_count(codes[: 2048 * 100], offs[:2049], kmers[::-1].copy(), 16, fast=True)
            assert np.array_equal(c.get_counts(), part)
    assert per_scan[0] == per_scan[1]


def test_reserve_changes_nothing_but_the_first_call_cost(built):
    """apc_reserve (buffers pre-sized, kernels loaded on a dummy sample) leaves the context without sample and
    queries, and every result afterwards is what it is without it."""
    from approx_counter_b200 import ApcError, ApproxCounter, host
    sample = np.ascontiguousarray(orc.synth_ends(99, 0, 6000, 100, True))
    codes, offs = orc.encode_matrix(sample)
    thr = host.adjust_threshold(1.0, 16, 16)
    with ApproxCounter(0) as plain, ApproxCounter(0) as warm:
        warm.reserve(6000, 101, 16, 300)
        with pytest.raises(ApcError):
            warm.scan()                                   # no sample, no queries after reserve
        res = []
        for c in (plain, warm):
            c.upload_sample(sample)
            km, ct, nd, hn = c.count_kmers_topn(16, thr, 300)
            res.append((km, ct, nd, hn, c.errorCount(km, 16), c.timing()["exact_ms"]))
        for a, b in zip(res[0][:5], res[1][:5]):
            assert np.array_equal(a, b)
        assert np.array_equal(res[1][4], orc.error_count(codes, offs, res[1][0], 16, fast=True))
        st = warm.scan_stats()
        assert st["scans"] == 1                           # the dummy run is not in the tally


@pytest.mark.parametrize("k", [16, 19, 20, 27, 32])
def test_every_unit_shape_alone(built, k):
    """Each unit shape on its own (shape_mask = one bit): a family of exactly G k-mers sharing K - T bases, forwards and
    — reversed — backwards, on a sample in which the family's trunk and its variants are planted with errors.  This is
    the direct check of every group-kernel instantiation of these k, the 60-row units included."""
    from approx_counter_b200 import ApproxCounter, plan_queries
    rng = np.random.default_rng(4200 + k)
    shapes = plan_queries(np.array([0], np.uint64), k)
    trunk = rng.choice(ACGT, size=k)
    n, L = 4096, 64 + k
    sample = rng.choice(ACGT, size=(n, L))
    for r in range(n):
        if r % 2 == 0:
            m = bytearray(trunk.tobytes())
            for _ in range(int(rng.integers(0, 3))):        # a tail variant: edits in the second half
                m[int(rng.integers(k // 2, k))] = int(rng.choice(ACGT))
            m = mutate(rng, bytes(m), int(rng.integers(0, 3)))[:L]
            if r % 4 == 0:
                m = m[::-1]                                  # what the backward-walking units look for
            pos = int(rng.choice([0, L - len(m), int(rng.integers(0, L - len(m) + 1))]))
            sample[r, pos:pos + len(m)] = np.frombuffer(m, np.uint8)
    sample[rng.random((n, L)) < 0.002] = ord("N")
    codes, offs = orc.encode_matrix(sample)
    with ApproxCounter(0) as c:
        c.upload_sample(sample)
        tested = 0
        for s in range(len(shapes["shape_g"])):
            t, g = int(shapes["shape_t"][s]), int(shapes["shape_g"][s])
            if g == 0:
                continue
            fam = set()
            while len(fam) < g:                              # G members: the trunk's first K - T bases + distinct tails
                tail = rng.choice(ACGT, size=t)
                fam.add(trunk[: k - t].tobytes() + tail.tobytes())
            fw = [orc.dna2int(x.decode()) for x in sorted(fam)]
            bw = [orc.dna2int(x[::-1].decode()) for x in sorted(fam)]      # share a SUFFIX: scanned reversed
            kmers = np.array(fw + bw + [int(x) for x in rng.integers(0, 1 << 30, 3)], np.uint64)
            want = orc.error_count(codes, offs, kmers, k, fast=True)
            c.set_option("shape_mask", 1 << s)
            got = c.errorCount(kmers, k)
            assert np.array_equal(got, want), (k, s, t, g)
            st = c.scan_stats()
            # the family really ran as units of this shape: fewer planned rows than one k-mer per warp
            assert st["lop3_planned"] < st["lop3_one_kmer_per_warp"], (k, s, t, g)
            tested += 1
        assert tested >= 12
