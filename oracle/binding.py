"""ctypes binding of oracle/kmn_oracle.c (test infrastructure only; see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libkmn_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_lookup3.so")

SCORING = {"SUM": 0, "MEDIAN": 1, "MIN": 2, "MAX": 3, "AVG": 4}
SCORE_LABEL = ["Score", "MedianScore", "MinScore", "MaxScore", "AvgScore"]


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "kmn_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c99", "-o", _LIB, src, "-lm"])
    if os.path.exists("/root/reference/src/lookup3.h") and (force or not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None
u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_kmer_hash.restype = C.c_uint64
        L.orc_kmer_hash.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_kmer_hash_lookup8.restype = C.c_uint64
        L.orc_kmer_hash_lookup8.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_owner.restype = C.c_uint32
        L.orc_owner.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_quality_table.argtypes = [f64p, C.c_int, C.c_int]
        L.orc_compress_sequence.restype = C.c_uint32
        L.orc_compress_sequence.argtypes = [C.c_char_p, C.c_uint32, u8p, u32p, C.c_char_p]
        L.orc_reverse_complement.argtypes = [u8p, u8p, C.c_uint32]
        L.orc_first_markup_n_or_x.restype = C.c_uint32
        L.orc_read_kmers.restype = C.c_uint32
        L.orc_read_kmers.argtypes = [C.c_char_p, u8p, C.c_uint32, C.c_uint32, C.c_int, f64p, u8p, u8p, f32p, u8p]
        L.orc_spectrum_new.restype = C.c_void_p
        L.orc_spectrum_new.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint32, C.c_uint64, C.c_int]
        L.orc_spectrum_free.argtypes = [C.c_void_p]
        L.orc_spectrum_add_reads.argtypes = [C.c_void_p, C.c_char_p, u8p, u64p, C.c_uint64, u8p, C.c_int, C.c_uint64]
        L.orc_spectrum_purge_min_depth.restype = C.c_uint64
        L.orc_spectrum_purge_min_depth.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_spectrum_stats.argtypes = [C.c_void_p, u64p]
        L.orc_spectrum_size.restype = C.c_uint64
        L.orc_spectrum_size.argtypes = [C.c_void_p]
        L.orc_spectrum_lookup.restype = C.c_uint32
        L.orc_spectrum_lookup.argtypes = [C.c_void_p, u8p]
        L.orc_spectrum_export.restype = C.c_uint64
        L.orc_spectrum_export.argtypes = [C.c_void_p, u8p, u16p, u16p, f32p, u32p]
        L.orc_histogram_bin.restype = C.c_uint32
        L.orc_histogram_bin.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_histogram_bucket_value.restype = C.c_uint32
        L.orc_histogram_bucket_value.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_spectrum_histogram.restype = C.c_uint32
        L.orc_spectrum_histogram.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p, f64p, C.c_uint32]
        L.orc_trim_values.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, u32p, u32p, f32p, u8p]
        L.orc_trim_reads.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint64, u8p, C.c_double, C.c_int,
                                     u32p, u32p, f32p, u8p, C.c_int]
        L.orc_passes_length.restype = C.c_int
        L.orc_passes_length.argtypes = [C.c_float, C.c_uint32, C.c_float]
        L.orc_estimate_raw_kmers.restype = C.c_uint64
        L.orc_estimate_raw_kmers.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def ref_kmer_hash(key: bytes):
    """KmerHasher::getHash computed by the reference's own lookup3.h (oracle/_ref); None if not built."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF):
            build()
        if not os.path.exists(_REF):
            return None
        _ref = C.CDLL(_REF)
        _ref.ref_kmer_hash.restype = C.c_uint64
        _ref.ref_kmer_hash.argtypes = [C.c_char_p, C.c_int]
    return _ref.ref_kmer_hash(key, len(key))


def kmer_hash(key: bytes) -> int:
    return lib().orc_kmer_hash(key, len(key))


def kmer_hash_lookup8(key: bytes) -> int:
    return lib().orc_kmer_hash_lookup8(key, len(key))


def owner(h: int, nranks: int) -> int:
    return lib().orc_owner(h, nranks)


def quality_table(min_quality: int, start: int = 33) -> np.ndarray:
    p = np.zeros(256, dtype=np.float64)
    lib().orc_quality_table(_p(p, f64p), start, min_quality)
    return p


def compress_sequence(bases: bytes):
    n = len(bases)
    out = np.zeros((n + 3) // 4 + 1, dtype=np.uint8)
    mpos = np.zeros(n + 1, dtype=np.uint32)
    mchr = C.create_string_buffer(n + 1)
    nm = lib().orc_compress_sequence(bases, n, _p(out, u8p), _p(mpos, u32p), mchr)
    return out[: (n + 3) // 4].copy(), [(mchr.raw[i : i + 1].decode("latin1"), int(mpos[i])) for i in range(nm)]


def first_markup_n_or_x(markups) -> int:
    for c, p in markups:
        if c in "NX":
            return p + 1
    return 0


def reverse_complement(packed: np.ndarray, length: int) -> np.ndarray:
    out = np.zeros(len(packed) + 1, dtype=np.uint8)
    inp = np.ascontiguousarray(np.concatenate([packed, np.zeros(1, np.uint8)]))
    lib().orc_reverse_complement(_p(inp, u8p), _p(out, u8p), length)
    return out[: (length + 3) // 4].copy()


def read_kmers(bases: bytes, quals: bytes, k: int, min_quality: int = 3, with_ext: bool = True, start: int = 33):
    """Returns (keys[nk,kb] u8, is_fwd[nk] u8, weight[nk] f32, ext[nk,4] u8)."""
    n = len(bases)
    nk = max(0, n - k + 1)
    kb = (k + 3) // 4
    keys = np.zeros((nk + 1, kb), dtype=np.uint8)
    fw = np.zeros(nk + 1, dtype=np.uint8)
    wt = np.zeros(nk + 1, dtype=np.float32)
    ext = np.zeros((nk + 1, 4), dtype=np.uint8)
    p = quality_table(min_quality, start)
    q = np.frombuffer(quals, dtype=np.uint8).copy()
    got = lib().orc_read_kmers(bases, _p(q, u8p), n, k, start, _p(p, f64p), _p(keys, u8p), _p(fw, u8p), _p(wt, f32p),
                               _p(ext, u8p) if with_ext else None)
    assert got == nk or (got == 0 and n < k)
    return keys[:got], fw[:got], wt[:got], ext[:got]


def trim_values(values, k, markup_length, min_score, scoring):
    v = np.ascontiguousarray(values, dtype=np.float32)
    off = C.c_uint32()
    ln = C.c_uint32()
    sc = C.c_float()
    wt = C.c_uint8()
    vv = v if len(v) else np.zeros(1, np.float32)
    lib().orc_trim_values(_p(vv, f32p), len(v), k, markup_length, float(min_score), int(scoring),
                          C.byref(off), C.byref(ln), C.byref(sc), C.byref(wt))
    return off.value, ln.value, sc.value, bool(wt.value)


def passes_length(length, read_length, minimum_length) -> bool:
    return bool(lib().orc_passes_length(float(length), int(read_length), float(minimum_length)))


def histogram_bin(count, zoom_max):
    return lib().orc_histogram_bin(count, zoom_max)


def histogram_bucket_value(idx, zoom_max):
    return lib().orc_histogram_bucket_value(idx, zoom_max)


def estimate_raw_kmers(n_reads, base_count, k):
    return lib().orc_estimate_raw_kmers(n_reads, base_count, k)


def concat_reads(seqs, quals=None):
    """list of bytes -> (bases buffer, quals array, offsets u64[n+1])."""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = b"".join(seqs)
    if quals is None:
        q = np.full(len(bases), ord("I"), dtype=np.uint8)
    else:
        q = np.frombuffer(b"".join(quals), dtype=np.uint8).copy()
    return bases, q, off


class OracleSpectrum:
    """Count table with the reference's value semantics (KmerSpectrum<..>::append / purge / histogram)."""

    def __init__(self, k, min_quality=3, min_kmer_quality=0.10, track_ext=False, threads=1, est_distinct=1 << 16,
                 hash_kind=0, start=33):
        self.k = k
        self.kb = (k + 3) // 4
        self.threads = threads
        self.track_ext = track_ext
        self._h = lib().orc_spectrum_new(k, start, min_quality, min_kmer_quality, int(track_ext), threads, est_distinct, hash_kind)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_spectrum_free(self._h)
            self._h = None

    def add_reads(self, bases, quals, off, discarded=None, batch_reads=0):
        q = np.ascontiguousarray(quals, dtype=np.uint8)
        o = np.ascontiguousarray(off, dtype=np.uint64)
        d = np.ascontiguousarray(discarded, dtype=np.uint8) if discarded is not None else None
        if isinstance(bases, np.ndarray):
            bases = bases.tobytes()
        lib().orc_spectrum_add_reads(self._h, bases, _p(q, u8p), _p(o, u64p), len(o) - 1, _p(d, u8p), self.threads, batch_reads)

    def purge_min_depth(self, min_depth):
        return lib().orc_spectrum_purge_min_depth(self._h, min_depth)

    def stats(self):
        s = np.zeros(5, dtype=np.uint64)
        lib().orc_spectrum_stats(self._h, _p(s, u64p))
        return dict(raw=int(s[0]), raw_good=int(s[1]), unique=int(s[2]), singleton=int(s[3]), purged_singletons=int(s[4]))

    def size(self):
        return lib().orc_spectrum_size(self._h)

    def lookup(self, key: bytes) -> int:
        k = np.frombuffer(key, dtype=np.uint8).copy()
        return lib().orc_spectrum_lookup(self._h, _p(k, u8p))

    def export(self):
        """-> dict(keys[n,kb] u8 sorted by key bytes, count u16, dir u16, wsum f32, ext[n,12] u32|None)."""
        n = self.size()
        keys = np.zeros((n + 1, self.kb), dtype=np.uint8)
        cnt = np.zeros(n + 1, dtype=np.uint16)
        dr = np.zeros(n + 1, dtype=np.uint16)
        ws = np.zeros(n + 1, dtype=np.float32)
        ext = np.zeros((n + 1, 12), dtype=np.uint32) if self.track_ext else None
        got = lib().orc_spectrum_export(self._h, _p(keys, u8p), _p(cnt, u16p), _p(dr, u16p), _p(ws, f32p), _p(ext, u32p))
        assert got == n
        return dict(keys=keys[:n], count=cnt[:n], dir=dr[:n], wsum=ws[:n], ext=ext[:n] if ext is not None else None)

    def histogram(self, zoom_max=256):
        nb = (1 << 16) + 1 + zoom_max + 1
        v = np.zeros(nb, dtype=np.uint64)
        c = np.zeros(nb, dtype=np.uint64)
        w = np.zeros(nb, dtype=np.float64)
        lib().orc_spectrum_histogram(self._h, zoom_max, _p(v, u64p), _p(c, u64p), _p(w, f64p), nb)
        return v, c, w

    def trim_reads(self, bases, off, min_depth, scoring, discarded=None, threads=1):
        o = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(o) - 1
        d = np.ascontiguousarray(discarded, dtype=np.uint8) if discarded is not None else None
        to = np.zeros(n + 1, dtype=np.uint32)
        tl = np.zeros(n + 1, dtype=np.uint32)
        sc = np.zeros(n + 1, dtype=np.float32)
        wt = np.zeros(n + 1, dtype=np.uint8)
        if isinstance(bases, np.ndarray):
            bases = bases.tobytes()
        lib().orc_trim_reads(self._h, bases, _p(o, u64p), n, _p(d, u8p), float(min_depth), int(scoring),
                             _p(to, u32p), _p(tl, u32p), _p(sc, f32p), _p(wt, u8p), threads)
        return to[:n], tl[:n], sc[:n], wt[:n]
