"""ctypes view of the host-side (CPU) pieces of libapc (include/apc_host.h): the
same C++ the drop-in `approx_counter` binary runs — codec, threshold, CompareCount
order, FASTA/FASTQ reader, sampler, exporters, synthetic reads, CLI entry."""
import ctypes as C

import numpy as np

from . import _lib

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

HOST_SYMBOLS = {
    "apch_dna2int": (C.c_int, [C.c_char_p, C.c_uint32, _u64p]),
    "apch_int2dna": (None, [C.c_uint64, C.c_uint32, C.c_char_p]),
    "apch_adjust_threshold": (C.c_float, [C.c_float, C.c_uint8, C.c_uint8]),
    "apch_get_complexity": (C.c_float, [C.c_uint64, C.c_uint8]),
    "apch_have_low_complexity": (C.c_int, [C.c_uint64, C.c_uint8, C.c_float]),
    "apch_lc_min_filtered_sum": (C.c_uint32, [C.c_uint8, C.c_float]),
    "apch_get_most_frequent": (C.c_uint64, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_int]),
    "apch_export_counter": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint8, C.c_char_p]),
    "apch_parse_kmer_list": (C.c_int64, [C.c_char_p, _vp, C.c_uint64]),
    "apch_reads_load": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "apch_reads_count": (C.c_uint64, [_vp]),
    "apch_reads_length": (C.c_uint64, [_vp, C.c_uint64]),
    "apch_reads_seq": (_vp, [_vp, C.c_uint64]),
    "apch_reads_mapped": (C.c_int, [_vp]),
    "apch_reads_free": (None, [_vp]),
    "apch_sample": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int64, _vp, _u64p]),
    "apch_shuffle_order": (C.c_int, [C.c_uint64, C.c_int64, _vp]),
    "apch_synth_ends": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _vp]),
    "apch_synth_write": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]),
    "apch_cli_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
}

_typed = False


def lib():
    global _typed
    L = _lib.load()
    if not _typed:
        for name, (res, args) in HOST_SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _typed = True
    return L


def dna2int(seq):
    out = C.c_uint64()
    if lib().apch_dna2int(seq.encode(), len(seq), C.byref(out)) != 0:
        raise ValueError(f"not a DNA k-mer of length <= 32: {seq!r}")
    return out.value


def int2dna(value, k):
    buf = C.create_string_buffer(k + 1)
    lib().apch_int2dna(int(value), k, buf)
    return buf.value.decode()


def adjust_threshold(c_old, k_old, k_new):
    return float(lib().apch_adjust_threshold(c_old, k_old, k_new))


def get_complexity(kmer, k):
    return float(lib().apch_get_complexity(int(kmer), k))


def have_low_complexity(kmer, k, thr):
    return bool(lib().apch_have_low_complexity(int(kmer), k, thr))


def lc_min_filtered_sum(k, thr):
    return int(lib().apch_lc_min_filtered_sum(k, thr))


def get_most_frequent(kmers, counts, limit, k):
    km = np.ascontiguousarray(kmers, np.uint64).copy()
    ct = np.ascontiguousarray(counts, np.uint64).copy()
    n = lib().apch_get_most_frequent(km.ctypes.data, ct.ctypes.data, len(km), int(limit), int(k))
    return km[:n], ct[:n]


def export_counter(kmers, counts, k, path):
    km = np.ascontiguousarray(kmers, np.uint64)
    ct = np.ascontiguousarray(counts, np.uint64)
    return bool(lib().apch_export_counter(km.ctypes.data, ct.ctypes.data, len(km), int(k), str(path).encode()))


def parse_kmer_list(path, capacity=1 << 20):
    out = np.zeros(capacity, np.uint64)
    n = lib().apch_parse_kmer_list(str(path).encode(), out.ctypes.data, capacity)
    if n < 0:
        raise OSError(f"cannot read {path}")
    return out[: min(n, capacity)].copy()


class Reads:
    """All records of a FASTA/FASTQ file (reference :819-825)."""

    def __init__(self, path):
        h = _vp()
        if lib().apch_reads_load(str(path).encode(), C.byref(h)) != 0:
            raise OSError(f"cannot parse {path}")
        self._h = h

    def __len__(self):
        return int(lib().apch_reads_count(self._h))

    @property
    def mapped(self):
        """True when the reads are views into a mapping of the file (no copy was made)."""
        return bool(lib().apch_reads_mapped(self._h))

    def seq(self, i):
        n = int(lib().apch_reads_length(self._h, i))
        return C.string_at(lib().apch_reads_seq(self._h, i), n)

    def sample(self, nb_sample, cut, bot, seed=-1):
        """sampleSequences (:415-476) -> uint8[n_sampled, cut (+1 if bot)] ASCII."""
        row = cut + (1 if bot else 0)
        out = np.empty((min(nb_sample, len(self)), row), np.uint8)
        n = C.c_uint64()
        lib().apch_sample(self._h, int(nb_sample), int(cut), int(bool(bot)), int(seed),
                          out.ctypes.data, C.byref(n))
        return out[: n.value]

    def close(self):
        if getattr(self, "_h", None):
            lib().apch_reads_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shuffle_order(n, seed=-1):
    """The shuffled read ids of sampleSequences (:423-429): uint32[n]."""
    out = np.zeros(int(n), np.uint32)
    if lib().apch_shuffle_order(int(n), int(seed), out.ctypes.data) != 0:
        raise MemoryError("apch_shuffle_order")
    return out


def synth_ends(seed, first, n, sl, bot, out=None):
    """Sampled ends of synthetic reads [first, first+n) (SURVEY.md §8d generator):
    uint8[n, sl] (start) or uint8[n, sl+1] (end).  `out` may be a pinned buffer."""
    row = sl + (1 if bot else 0)
    if out is None:
        out = np.empty((n, row), np.uint8)
    assert out.shape == (n, row) and out.dtype == np.uint8 and out.flags.c_contiguous
    lib().apch_synth_ends(int(seed), int(first), int(n), int(sl), int(bool(bot)), out.ctypes.data)
    return out


def synth_write(path, seed, n, sl, fastq=False):
    if lib().apch_synth_write(str(path).encode(), int(seed), int(n), int(sl), int(bool(fastq))) != 0:
        raise OSError(f"cannot write {path}")


def cli_main(argv):
    """Run the drop-in binary's main() in-process; returns its exit code."""
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    return int(lib().apch_cli_main(len(argv), arr))
