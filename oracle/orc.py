"""ctypes front end for the CPU oracle (oracle/liborc.so, oracle/libseqan_model.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product
package.  PARITY UNPINNED (see apc_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with its Makefile (gcc/g++, seconds)."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, f)) for f in ("liborc.so", "libseqan_model.so", "libfm_index_model.so"))
    if not need:
        for so, srcs in (("liborc.so", ("apc_oracle.c", "apc_oracle.h", "synth_reads.c")),
                         ("libseqan_model.so", ("seqan_model.cpp",)),
                         ("libfm_index_model.so", ("fm_index_model.cpp",))):
            t = os.path.getmtime(os.path.join(_HERE, so))
            need |= any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in srcs)
    if need:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


_lib = None
_model = None
_fm = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liborc.so"))
        L.orc_dna2int.restype = C.c_uint64
        L.orc_dna2int.argtypes = [_u8p, C.c_int]
        L.orc_int2dna.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.orc_char2code.restype = C.c_uint8
        L.orc_char2code.argtypes = [C.c_char]
        L.orc_adjust_threshold.restype = C.c_float
        L.orc_adjust_threshold.argtypes = [C.c_float, C.c_uint8, C.c_uint8]
        L.orc_have_low_complexity.restype = C.c_int
        L.orc_have_low_complexity.argtypes = [C.c_uint64, C.c_uint8, C.c_float]
        L.orc_get_complexity.restype = C.c_float
        L.orc_get_complexity.argtypes = [C.c_uint64, C.c_uint8]
        L.orc_dimer_sum.restype = C.c_uint64
        L.orc_dimer_sum.argtypes = [C.c_uint64, C.c_uint8]
        L.orc_compare_count.restype = C.c_int
        L.orc_compare_count.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_count_kmers.restype = C.c_uint64
        L.orc_count_kmers.argtypes = [_u8p, _u64p, C.c_uint64, C.c_uint8, C.c_float,
                                      C.c_void_p, C.c_uint64,
                                      C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.c_uint64)]
        L.orc_count_kmers_mt.restype = C.c_uint64
        L.orc_count_kmers_mt.argtypes = [_u8p, _u64p, C.c_uint64, C.c_uint8, C.c_float, C.c_int,
                                         C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint64)),
                                         C.POINTER(C.c_uint64)]
        L.orc_get_most_frequent.restype = C.c_uint64
        L.orc_get_most_frequent.argtypes = [_u64p, _u64p, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_get_solid_kmers.restype = C.c_uint64
        L.orc_get_solid_kmers.argtypes = [_u64p, _u64p, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_error_count.argtypes = [_u8p, _u64p, C.c_uint64, _u64p, C.c_uint64, C.c_uint8, _u64p]
        L.orc_error_count_fast.argtypes = [_u8p, _u64p, C.c_uint64, _u64p, C.c_uint64,
                                           C.c_uint8, C.c_int, _u64p]
        L.orc_min_infix_distance.restype = C.c_int
        L.orc_min_infix_distance.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_uint8]
        L.orc_sample_sequences.restype = C.c_uint64
        L.orc_sample_sequences.argtypes = [_u8p, _u64p, C.c_uint64, _u64p, C.c_uint64,
                                           C.c_uint64, C.c_int, _u8p, _u64p]
        L.orc_export_counter.restype = C.c_int
        L.orc_export_counter.argtypes = [_u64p, _u64p, C.c_uint64, C.c_uint8, C.c_char_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_dimer_sums.argtypes = [_u64p, C.c_uint64, C.c_uint8, C.c_void_p]
        L.orc_synth_ends.restype = C.c_int
        L.orc_synth_ends.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def model():
    global _model
    if _model is None:
        build()
        M = C.CDLL(os.path.join(_HERE, "libseqan_model.so"))
        M.seqan_model_error_count.argtypes = [_u8p, _u64p, C.c_uint64, _u64p, C.c_uint64,
                                              C.c_uint8, C.c_int, _u64p, C.c_void_p]
        _model = M
    return _model


def fm():
    global _fm
    if _fm is None:
        build()
        F = C.CDLL(os.path.join(_HERE, "libfm_index_model.so"))
        F.fm_error_count.restype = C.c_int
        F.fm_error_count.argtypes = [_u8p, _u64p, C.c_uint64, _u64p, C.c_uint64, C.c_uint8, C.c_int, _u64p,
                                     C.POINTER(C.c_double)]
        _fm = F
    return _fm


_CODE = np.full(256, 4, np.uint8)
for _i, _ch in enumerate("ACGT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i


def encode(reads):
    """list of str/bytes -> (codes u8[sum len], offs u64[n+1]) in Dna5 ordinals."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offs = np.zeros(len(bs) + 1, np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    raw = np.frombuffer(b"".join(bs), np.uint8)
    return np.ascontiguousarray(_CODE[raw]), offs


def encode_matrix(ascii_mat):
    """uint8[n, L] ASCII matrix -> (codes, offs) with uniform read length."""
    n, L = ascii_mat.shape
    return (np.ascontiguousarray(_CODE[ascii_mat].reshape(-1)),
            np.arange(n + 1, dtype=np.uint64) * np.uint64(L))


def dna2int(s):
    c, _ = encode([s])
    return int(lib().orc_dna2int(c, len(s)))


def int2dna(v, k):
    buf = C.create_string_buffer(k + 1)
    lib().orc_int2dna(int(v), k, buf)
    return buf.value.decode()


def adjust_threshold(c_old, k_old, k_new):
    return float(lib().orc_adjust_threshold(c_old, k_old, k_new))


def have_low_complexity(kmer, k, thr):
    return bool(lib().orc_have_low_complexity(int(kmer), k, thr))


def get_complexity(kmer, k):
    return float(lib().orc_get_complexity(int(kmer), k))


def dimer_sum(kmer, k):
    return int(lib().orc_dimer_sum(int(kmer), k))


def dimer_sums(kmers, k):
    """orc_dimer_sum (:216-231) over an array."""
    kmers = np.ascontiguousarray(kmers, np.uint64)
    out = np.zeros(len(kmers), np.uint32)
    lib().orc_dimer_sums(kmers, len(kmers), k, out.ctypes.data)
    return out


def count_kmers(codes, offs, k, thr, forbidden=None):
    """-> (keys u64[D], counts u64[D], had_n) in unspecified order."""
    kp, cp = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
    had = C.c_uint64(0)
    fb = None if forbidden is None else np.ascontiguousarray(forbidden, np.uint64)
    n = lib().orc_count_kmers(codes, offs, len(offs) - 1, k, thr,
                              None if fb is None else fb.ctypes.data,
                              0 if fb is None else len(fb),
                              C.byref(kp), C.byref(cp), C.byref(had))
    keys = np.ctypeslib.as_array(kp, shape=(max(n, 1),))[:n].copy()
    cnts = np.ctypeslib.as_array(cp, shape=(max(n, 1),))[:n].copy()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(kp)
    libc.free(cp)
    return keys, cnts, int(had.value)


def count_kmers_mt(codes, offs, k, thr, nb_thread=0):
    """count_kmers with all host threads (same multiset; set-up of the reference arm and full-size tests)."""
    if nb_thread <= 0:
        nb_thread = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kp, cp = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
    had = C.c_uint64(0)
    n = lib().orc_count_kmers_mt(codes, offs, len(offs) - 1, k, thr, int(nb_thread), C.byref(kp), C.byref(cp),
                                 C.byref(had))
    keys = np.ctypeslib.as_array(kp, shape=(max(n, 1),))[:n].copy()
    cnts = np.ctypeslib.as_array(cp, shape=(max(n, 1),))[:n].copy()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(kp)
    libc.free(cp)
    return keys, cnts, int(had.value)


def get_most_frequent(keys, counts, limit, k):
    keys = np.ascontiguousarray(keys, np.uint64).copy()
    counts = np.ascontiguousarray(counts, np.uint64).copy()
    n = lib().orc_get_most_frequent(keys, counts, len(keys), limit, k)
    return keys[:n], counts[:n]


def get_most_frequent_fast(keys, counts, limit, k):
    """get_most_frequent (:396-405) without sorting all D distinct k-mers (minutes at 1e8): a superset of the
    first `limit` entries in CompareCount order (:275-305: count desc, complexity asc, k-mer desc) is cut out
    with integer thresholds — the complexity score is the integer dimer sum over a constant (:227-233), so it
    orders like the sum — and the comparator sort of get_most_frequent then runs on that superset only.
    tests/test_oracle.py holds it equal to the full sort."""
    keys = np.ascontiguousarray(keys, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint64)
    if limit == 0 or len(keys) <= max(4 * limit, 1 << 16):
        return get_most_frequent(keys, counts, limit, k)
    cut = len(counts) - limit
    c_star = np.partition(counts, cut)[cut]                 # limit-th largest count
    above = counts > c_star
    tie = np.flatnonzero(counts == c_star)
    need = limit - int(above.sum())
    sums = dimer_sums(keys[tie], k)
    s_star = np.partition(sums, need - 1)[need - 1]
    keep = np.concatenate([np.flatnonzero(above), tie[sums <= s_star]])
    return get_most_frequent(keys[keep], counts[keep], limit, k)


def get_solid_kmers(keys, counts, solid_km, k):
    keys = np.ascontiguousarray(keys, np.uint64).copy()
    counts = np.ascontiguousarray(counts, np.uint64).copy()
    n = lib().orc_get_solid_kmers(keys, counts, len(keys), solid_km, k)
    return keys[:n], counts[:n]


def error_count(codes, offs, kmers, k, fast=False, nb_thread=0):
    kmers = np.ascontiguousarray(kmers, np.uint64)
    out = np.zeros(len(kmers), np.uint64)
    if fast:
        lib().orc_error_count_fast(codes, offs, len(offs) - 1, kmers, len(kmers), k, nb_thread, out)
    else:
        lib().orc_error_count(codes, offs, len(offs) - 1, kmers, len(kmers), k, out)
    return out


def min_infix_distance(text_codes, kmer, k):
    t = np.ascontiguousarray(text_codes, np.uint8)
    return int(lib().orc_min_infix_distance(t, len(t), int(kmer), k))


def seqan_model_error_count(codes, offs, kmers, k, variant=0, want_flags=False):
    kmers = np.ascontiguousarray(kmers, np.uint64)
    n = len(offs) - 1
    out = np.zeros(len(kmers), np.uint64)
    flags = np.zeros((len(kmers), 3, n), np.uint8) if want_flags else None
    model().seqan_model_error_count(codes, offs, n, kmers, len(kmers), k, variant, out,
                                    None if flags is None else flags.ctypes.data)
    return (out, flags) if want_flags else out


def fm_index_error_count(codes, offs, kmers, k, nb_thread=0, want_seconds=False):
    """errorCount the way the reference runs it (:531-601): build a bidirectional FM index of the reads, search
    every k-mer with the optimal-search-scheme recursion (fm_index_model.cpp).  want_seconds: also return
    (index build s, search s)."""
    kmers = np.ascontiguousarray(kmers, np.uint64)
    out = np.zeros(len(kmers), np.uint64)
    sec = (C.c_double * 2)()
    rc = fm().fm_error_count(codes, offs, len(offs) - 1, kmers, len(kmers), k, nb_thread, out, sec)
    if rc != 0:
        raise ValueError("fm_error_count: text too long for 32-bit positions or k outside [4,32]")
    return (out, (sec[0], sec[1])) if want_seconds else out


def sample_sequences(codes, offs, perm, nb_sample, cut, bot):
    perm = np.ascontiguousarray(perm, np.uint64)
    out_codes = np.zeros(max(1, nb_sample * (cut + 1)), np.uint8)
    out_offs = np.zeros(nb_sample + 1, np.uint64)
    n = lib().orc_sample_sequences(codes, offs, len(offs) - 1, perm, nb_sample, cut,
                                   int(bool(bot)), out_codes, out_offs)
    return out_codes[: int(out_offs[n])].copy(), out_offs[: n + 1].copy()


def export_counter(keys, counts, k, path):
    keys = np.ascontiguousarray(keys, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint64)
    return bool(lib().orc_export_counter(keys, counts, len(keys), k, str(path).encode()))


def num_threads():
    return int(lib().orc_num_threads())


def synth_ends(seed, first, n, sl, bot):
    """Sampled ends of reads [first, first+n) of the synthetic stream `seed` (synth_reads.c):
    uint8[n, sl] (start) or uint8[n, sl+1] (end) — the CPU side's own generator, byte-identical to the
    product's apch_synth_ends (tests/test_host.py)."""
    out = np.empty((int(n), int(sl) + (1 if bot else 0)), np.uint8)
    if lib().orc_synth_ends(int(seed), int(first), int(n), int(sl), int(bool(bot)), out.ctypes.data) != 0:
        raise MemoryError("orc_synth_ends")
    return out
