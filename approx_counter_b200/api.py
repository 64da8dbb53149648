"""Host-side mirror of the reference's interface for the counting path, bound to
libapc's C ABI (include/apc.h).  Names follow /root/reference/approx_counter.cpp:
`errorCount` (:531), `count_kmers` (:487) + `get_most_frequent` (:396).

Nothing here computes on the CPU: every method is a thin ctypes call into the
CUDA library, and raises ApcError / ImportError when that is impossible.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ApcError, ApcScanStats, ApcTiming


def _as_kmers(kmers):
    return np.ascontiguousarray(kmers, dtype=np.uint64)


class ApproxCounter:
    """One GPU's worth of the path: upload the sampled read ends once, then run
    the exact top-N and the approximate count against them."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        h = C.c_void_p()
        st = self._lib.apc_create(int(device), C.byref(h))
        if st != _lib.APC_OK:
            raise ApcError(st, self._lib.apc_strerror(st).decode())
        self._h = h
        self.device = int(device)
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing ---------------------------------------------------------------
    def _check(self, st):
        if st != _lib.APC_OK:
            raise ApcError(st, self._lib.apc_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.apc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream), or
        None for the context's own stream.  0 is torch's spelling of the legacy default
        stream and is passed on as cudaStreamLegacy."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1  # cudaStreamLegacy == (cudaStream_t)0x1
        self._check(self._lib.apc_set_stream(self._h, C.c_void_p(handle)))

    def sync(self):
        self._check(self._lib.apc_sync(self._h))

    def set_option(self, name, value):
        self._check(self._lib.apc_set_option(self._h, name.encode(), int(value)))

    def reserve(self, n_reads, read_len, k, n_kmers=0):
        """One-time set-up for one-shot hosts (apc_reserve): buffers allocated, kernels loaded."""
        self._check(self._lib.apc_reserve(self._h, int(n_reads), int(read_len), int(k), int(n_kmers)))

    # -- sample -------------------------------------------------------------------
    def upload_sample(self, sample):
        """sample: uint8[n, L] ASCII matrix (uniform length), or a list of
        str/bytes (ragged).  Mirrors handing `sample` to errorCount (:922)."""
        if isinstance(sample, np.ndarray):
            if sample.ndim != 2 or sample.dtype != np.uint8:
                raise ValueError("sample matrix must be uint8[n, L]")
            m = np.ascontiguousarray(sample)
            self._check(self._lib.apc_upload_sample(self._h, m.ctypes.data, m.shape[0], m.shape[1]))
            return
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in sample]
        offs = np.zeros(len(bs) + 1, np.uint64)
        if bs:
            offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        raw = np.frombuffer(b"".join(bs) or b"\0", np.uint8)
        self._check(self._lib.apc_upload_sample_ragged(self._h, raw.ctypes.data, offs.ctypes.data, len(bs)))

    def upload_sample_ptr(self, host_ptr, n_reads, read_len):
        """Same, from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._check(self._lib.apc_upload_sample(self._h, C.c_void_p(int(host_ptr)), int(n_reads), int(read_len)))

    def upload_sample_ptr_async(self, host_ptr, n_reads, read_len):
        """Enqueue only; the (page-locked) buffer must stay valid until sync()."""
        self._check(self._lib.apc_upload_sample_async(self._h, C.c_void_p(int(host_ptr)), int(n_reads),
                                                      int(read_len)))

    def sample_info(self):
        n, ml, tb = C.c_uint64(), C.c_uint32(), C.c_uint64()
        self._check(self._lib.apc_sample_info(self._h, C.byref(n), C.byref(ml), C.byref(tb)))
        return n.value, ml.value, tb.value

    # -- ingest on the device (:819-825 + :415-476) ---------------------------------
    def ingest_fastx(self, data):
        """data: the bytes of a FASTA / FASTQ file (bytes, bytearray, mmap or uint8 array).  Copies them to
        HBM and indexes the records there (apc_ingest_fastx) -> (n_records, is_fastq).  Raises ApcError
        with status APC_ERR_FORMAT (-9) for wrapped records and other input the device parser leaves to
        the host parser (host.Reads)."""
        a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        n, fq = C.c_uint64(), C.c_int()
        st = self._lib.apc_ingest_fastx(self._h, a.ctypes.data if a.size else None, a.size, C.byref(n), C.byref(fq))
        if st != _lib.APC_OK:
            raise ApcError(st, self._lib.apc_last_error(self._h).decode())
        self._n_records = n.value
        return n.value, bool(fq.value)

    def ingest_lengths(self, first=0, n=None):
        n = self._n_records - first if n is None else n
        out = np.zeros(n, np.uint32)
        self._check(self._lib.apc_ingest_lengths(self._h, int(first), int(n), out.ctypes.data))
        return out

    def sample_resident(self, nb_sample, cut, bot, order=None):
        """sampleSequences (:415-476) on the device over the ingested file; `order`: the shuffled read ids
        (host.shuffle_order), None = file order.  The context then holds the sample.  -> n_sampled."""
        n = C.c_uint64()
        if order is None:
            st = self._lib.apc_sample_resident(self._h, None, 0, int(nb_sample), int(cut), int(bool(bot)), C.byref(n))
        else:
            o = np.ascontiguousarray(order, np.uint32)
            st = self._lib.apc_sample_resident(self._h, o.ctypes.data, len(o), int(nb_sample), int(cut), int(bool(bot)),
                                               C.byref(n))
        self._check(st)
        return n.value

    def upload_sample_peer(self, src, first_read, n_reads):
        """Rows [first_read, first_read + n_reads) of the sample resident in `src` (another ApproxCounter, normally
        on another GPU) become this context's sample, copied GPU to GPU (apc_upload_sample_peer; asynchronous)."""
        self._check(self._lib.apc_upload_sample_peer(self._h, src._h, int(first_read), int(n_reads)))

    def download_sample(self):
        """ASCII rows of the resident sample -> uint8[n_reads, read_len]."""
        n, ml, _ = self.sample_info()
        out = np.zeros((n, ml), np.uint8)
        self._check(self._lib.apc_download_sample(self._h, out.ctypes.data if out.size else None, out.size))
        return out

    def ingest_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._check(self._lib.apc_ingest_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"copy_ms": a.value, "index_ms": b.value, "sample_ms": c.value}

    # -- exact stage (:487-519 + :396-405) ------------------------------------------
    def count_kmers_topn(self, k, lc_adjusted, lim, forbidden=None):
        """-> (kmers u64[n], counts u64[n], n_distinct, had_n), CompareCount order."""
        fb = None if forbidden is None else _as_kmers(forbidden)
        km = np.zeros(max(1, lim), np.uint64)
        ct = np.zeros(max(1, lim), np.uint64)
        n, nd, hn = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.apc_exact_topn(
            self._h, int(k), float(lc_adjusted), int(lim),
            None if fb is None else fb.ctypes.data, 0 if fb is None else len(fb),
            km.ctypes.data, ct.ctypes.data, C.byref(n), C.byref(nd), C.byref(hn)))
        return km[: n.value].copy(), ct[: n.value].copy(), nd.value, hn.value

    def solid_kmers(self, k, lc_adjusted, solid_km, forbidden=None, capacity=1 << 20):
        fb = None if forbidden is None else _as_kmers(forbidden)
        while True:
            km = np.zeros(max(1, capacity), np.uint64)
            ct = np.zeros(max(1, capacity), np.uint64)
            n, nd, hn = C.c_uint64(), C.c_uint64(), C.c_uint64()
            st = self._lib.apc_exact_solid(
                self._h, int(k), float(lc_adjusted), int(solid_km),
                None if fb is None else fb.ctypes.data, 0 if fb is None else len(fb),
                km.ctypes.data, ct.ctypes.data, capacity, C.byref(n), C.byref(nd), C.byref(hn))
            if st == -7:  # APC_ERR_CAPACITY
                capacity = int(n.value)
                continue
            self._check(st)
            return km[: n.value].copy(), ct[: n.value].copy(), nd.value, hn.value

    # -- approximate stage (:531-601) ---------------------------------------------------
    def errorCount(self, kmers, k):
        """counts[i] = sum_e |reads flagged at error level e| for kmers[i] (host in/out)."""
        km = _as_kmers(kmers)
        out = np.zeros(len(km), np.uint64)
        self._n_kmers = len(km)
        self._check(self._lib.apc_approx_count(self._h, int(k), km.ctypes.data, len(km), out.ctypes.data))
        return out

    def errorCount_ptr(self, kmers_ptr, n_kmers, k, counts_ptr):
        self._n_kmers = int(n_kmers)
        self._check(self._lib.apc_approx_count(self._h, int(k), C.c_void_p(int(kmers_ptr)), int(n_kmers),
                                               C.c_void_p(int(counts_ptr))))

    def errorCount_ptr_async(self, kmers_ptr, n_kmers, k, counts_ptr):
        """Enqueue only; counts are valid after sync()."""
        self._n_kmers = int(n_kmers)
        self._check(self._lib.apc_approx_count_async(self._h, int(k), C.c_void_p(int(kmers_ptr)), int(n_kmers),
                                                     C.c_void_p(int(counts_ptr))))

    def set_queries(self, kmers, k):
        km = _as_kmers(kmers)
        self._n_kmers = len(km)
        self._check(self._lib.apc_set_queries(self._h, int(k), km.ctypes.data, len(km)))

    def set_queries_ptr(self, kmers_ptr, n_kmers, k):
        self._n_kmers = int(n_kmers)
        self._check(self._lib.apc_set_queries(self._h, int(k), C.c_void_p(int(kmers_ptr)), int(n_kmers)))

    def scan(self, d_counts_ptr=None):
        """Launch the scan on the context's stream (asynchronous)."""
        self._check(self._lib.apc_scan(self._h, C.c_void_p(int(d_counts_ptr) if d_counts_ptr else 0)))

    # -- multi-GPU: one NCCL rank per context (apc_comm_*) ---------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte id created on rank 0 and handed to every rank (any channel)."""
        buf = (C.c_uint8 * 128)()
        st = _lib.load().apc_comm_unique_id(buf)
        if st != _lib.APC_OK:
            raise ApcError(st, "apc_comm_unique_id (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init_rank(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.apc_comm_init_rank(self._h, int(n_ranks), int(rank), buf))

    def comm_destroy(self):
        self._check(self._lib.apc_comm_destroy(self._h))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        self._check(self._lib.apc_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def allreduce_counts(self, d_counts_ptr=None, n=0):
        """In-place sum over the ranks on the context's stream (no-op without a communicator)."""
        self._check(self._lib.apc_allreduce_counts(self._h, C.c_void_p(int(d_counts_ptr) if d_counts_ptr else 0), int(n)))

    def scan_allreduce(self, d_counts_ptr=None):
        """errorCount of a sharded sample: scan this rank's shard, then sum over the ranks (asynchronous)."""
        self._check(self._lib.apc_scan_allreduce(self._h, C.c_void_p(int(d_counts_ptr) if d_counts_ptr else 0)))

    def scan_stats(self):
        """LOP3 warp instructions of the scans since the last call (apc_scan_stats_read); synchronises."""
        s = ApcScanStats()
        self._check(self._lib.apc_scan_stats_read(self._h, C.byref(s)))
        return {"scans": int(s.scans), "lop3_executed": s.lop3_executed, "lop3_top": s.lop3_top,
                "lop3_planned": s.lop3_planned, "lop3_one_kmer_per_warp": s.lop3_one_kmer_per_warp}

    def get_counts(self):
        out = np.zeros(self._n_kmers, np.uint64)
        self._check(self._lib.apc_get_counts(self._h, out.ctypes.data))
        return out

    def counts_device_ptr(self):
        return int(self._lib.apc_counts_device_ptr(self._h) or 0)

    def timing(self):
        t = ApcTiming()
        self._check(self._lib.apc_last_timing(self._h, C.byref(t)))
        return {"upload_ms": t.upload_ms, "exact_ms": t.exact_ms, "scan_ms": t.scan_ms,
                "total_ms": t.total_ms, "scan_launches": int(t.scan_launches),
                "exact_launches": int(t.exact_launches)}

    def timing_launches(self):
        """Kernels launched by the last scan (no synchronisation)."""
        return int(self._lib.apc_last_scan_launches(self._h))

    def measure_int_peak(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self._lib.apc_measure_int_peak(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"lop3_ops_per_s": a.value, "imad_ops_per_s": b.value, "mixed_ops_per_s": c.value}


    def microbench(self, name):
        v = C.c_double()
        self._check(self._lib.apc_microbench(self._h, name.encode(), C.byref(v)))
        return v.value


def device_count():
    n = _lib.load().apc_device_count()
    return max(0, n)


PLAN_SHAPES = 22  # APC_PLAN_SHAPES (include/apc.h)


def plan_queries(kmers, k):
    """The scan plan of the default kernel for these k-mers (apc_plan_queries; needs no GPU):
    dict(order, reversed, units, shape_t, shape_g) — see include/apc.h."""
    lib = _lib.load()
    km = _as_kmers(kmers)
    n = len(km)
    order = np.zeros(n, np.uint32)
    rev = np.zeros(n, np.uint8)
    units = np.zeros(PLAN_SHAPES, np.uint32)
    st = np.zeros(PLAN_SHAPES, np.int32)
    sg = np.zeros(PLAN_SHAPES, np.int32)
    rc = lib.apc_plan_queries(int(k), km.ctypes.data, n, order.ctypes.data, rev.ctypes.data, units.ctypes.data,
                              st.ctypes.data, sg.ctypes.data)
    if rc != 0:
        raise ApcError(rc, "apc_plan_queries")
    return {"order": order, "reversed": rev, "units": units, "shape_t": st, "shape_g": sg}
