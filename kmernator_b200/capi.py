"""ctypes binding of the C ABI (include/kmernator_b200.h).  No CPU fallback: loading fails loudly when the
CUDA library has not been built, and kmn_create fails loudly when no GPU is present."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkmernator_b200.so")

KMN_VALUE_DIR, KMN_VALUE_DIR_EXT, KMN_VALUE_WEIGHTS = 0, 1, 2
KMN_HASH_LOOKUP3, KMN_HASH_LOOKUP8 = 0, 1
SCORING = {"SUM": 0, "MEDIAN": 1, "MIN": 2, "MAX": 3, "AVG": 4}

EXPORTED = [
    "kmn_default_opts", "kmn_last_error", "kmn_version", "kmn_create", "kmn_destroy", "kmn_reset", "kmn_comm_unique_id",
    "kmn_comm_init", "kmn_count_batch", "kmn_count_finish", "kmn_get_stats", "kmn_purge_min_depth", "kmn_histogram",
    "kmn_lookup", "kmn_trim_batch", "kmn_export", "kmn_debug_kmers", "kmn_sync", "kmn_stream", "kmn_launch_count",
    "kmn_profile_enable", "kmn_profile_read", "kmn_import", "kmn_subtract", "kmn_debug_owner", "kmn_count_batch_2na",
]
PROF_KINDS = ["parse", "insert", "route", "lookup", "trim", "scan", "weight", "subpart"]


class KmnProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_uint64 * 8), ("units", C.c_uint64 * 8)]


class KmnOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("kmer_size", C.c_uint32), ("fastq_start_char", C.c_uint32),
        ("min_quality_score", C.c_uint32), ("min_kmer_quality", C.c_float), ("min_depth", C.c_uint32),
        ("hash_kind", C.c_uint32), ("value_kind", C.c_uint32), ("est_raw_kmers", C.c_uint64),
        ("table_slots", C.c_uint64), ("stage_keys", C.c_uint64), ("slice_bytes", C.c_uint32),
        ("device", C.c_uint32), ("ignore_quality", C.c_uint32), ("reserved", C.c_uint32 * 7),
    ]


class KmnStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("raw_kmers", "raw_good_kmers", "unique_kmers", "singleton_kmers",
                                           "discarded_kmers", "table_slots", "table_partitions", "direct_inserts")]


class KmnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kmernator_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Loads the in-tree CUDA library.  Raises if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("KMN_LIB_VARIANT"):      # developer A/B builds (kmernator_b200/build.py)
        LIB_PATH = LIB_PATH[:-3] + "." + os.environ["KMN_LIB_VARIANT"] + ".so"
    if not os.path.exists(LIB_PATH):
        raise ImportError("kmernator_b200: %s not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    L.kmn_default_opts.argtypes = [C.POINTER(KmnOpts)]
    L.kmn_last_error.restype = C.c_char_p
    L.kmn_last_error.argtypes = [vp]
    L.kmn_version.restype = C.c_char_p
    L.kmn_create.argtypes = [C.POINTER(vp), C.POINTER(KmnOpts)]
    L.kmn_destroy.argtypes = [vp]
    L.kmn_destroy.restype = None
    L.kmn_reset.argtypes = [vp]
    L.kmn_comm_unique_id.argtypes = [vp]
    L.kmn_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.kmn_count_batch.argtypes = [vp, vp, vp, vp, C.c_uint64, vp]
    L.kmn_count_batch_2na.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64]
    L.kmn_count_finish.argtypes = [vp, C.c_int]
    L.kmn_get_stats.argtypes = [vp, C.POINTER(KmnStats)]
    L.kmn_purge_min_depth.argtypes = [vp, C.c_uint32]
    L.kmn_histogram.argtypes = [vp, vp, vp]
    L.kmn_lookup.argtypes = [vp, vp, C.c_uint64, vp]
    L.kmn_trim_batch.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32, C.c_int, vp, vp, vp, vp]
    L.kmn_export.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.kmn_debug_kmers.argtypes = [vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
    L.kmn_sync.argtypes = [vp]
    L.kmn_debug_owner.argtypes = [vp, vp, C.c_uint64, C.c_uint32, vp]
    L.kmn_import.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint64]
    L.kmn_subtract.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kmn_profile_enable.argtypes = [vp, C.c_int]
    L.kmn_profile_read.argtypes = [vp, C.POINTER(KmnProfile)]
    L.kmn_stream.restype = vp
    L.kmn_stream.argtypes = [vp]
    L.kmn_launch_count.restype = C.c_uint64
    L.kmn_launch_count.argtypes = [vp]
    _lib = L
    return L


def _ptr(x):
    """numpy array / torch tensor / bytes / None -> address"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if isinstance(x, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(x)), C.c_void_p).value
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def pack_2na(bases, read_off):
    """ASCII bases + offsets -> (packed u8, packed_off u64, markup_pos u64, markup_chr u8): the reference's TwoBitSequence
    bytes per read (src/TwoBitSequence.cpp:242-269; non-ACGT packs as A and becomes a markup, '.' counts as N :253-260,
    lower case is upper-cased by the reader before packing)."""
    b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.asarray(bases, dtype=np.uint8)
    off = np.asarray(read_off, dtype=np.uint64)
    n = len(off) - 1
    lens = (off[1:] - off[:-1]).astype(np.int64)
    nb = (lens + 3) // 4
    poff = np.zeros(n + 1, dtype=np.uint64)
    poff[1:] = np.cumsum(nb)
    up = b & 0xDF
    code = np.zeros(len(b), dtype=np.uint8)
    code[up == ord("C")] = 1
    code[up == ord("G")] = 2
    code[up == ord("T")] = 3
    valid = (up == ord("A")) | (up == ord("C")) | (up == ord("G")) | (up == ord("T"))
    packed = np.zeros(int(poff[-1]), dtype=np.uint8)
    read_of = np.repeat(np.arange(n), lens)
    pos_in = np.arange(len(b)) - np.repeat(off[:-1].astype(np.int64), lens)
    byte_idx = poff[read_of].astype(np.int64) + pos_in // 4
    np.add.at(packed, byte_idx, (code << (6 - 2 * (pos_in % 4)).astype(np.uint8)).astype(np.uint8))
    mpos = np.nonzero(~valid)[0].astype(np.uint64)
    mchr = b[~valid].copy()
    mchr[mchr == ord(".")] = ord("N")
    return packed, poff, mpos, mchr


class Context:
    """One k-mer spectrum on one GPU (kmn_ctx)."""

    def __init__(self, kmer_size=31, fastq_start_char=33, min_quality_score=3, min_kmer_quality=0.10, min_depth=2,
                 hash_kind=KMN_HASH_LOOKUP3, value_kind=KMN_VALUE_DIR, est_raw_kmers=0, table_slots=0, stage_keys=0,
                 slice_bytes=0, device=0, ignore_quality=False):
        L = load()
        o = KmnOpts()
        L.kmn_default_opts(C.byref(o))
        o.kmer_size, o.fastq_start_char, o.min_quality_score = kmer_size, fastq_start_char, min_quality_score
        o.min_kmer_quality, o.min_depth, o.hash_kind, o.value_kind = min_kmer_quality, min_depth, hash_kind, value_kind
        o.est_raw_kmers, o.table_slots, o.stage_keys, o.device = est_raw_kmers, table_slots, stage_keys, device
        if slice_bytes:
            o.slice_bytes = slice_bytes
        o.ignore_quality = int(ignore_quality)
        self.opts = o
        self.k = kmer_size
        self.kb = (kmer_size + 3) // 4
        self.value_kind = value_kind
        self._h = C.c_void_p()
        self._L = L
        rc = L.kmn_create(C.byref(self._h), C.byref(o))
        if rc:
            self._h = None
            raise KmnError(rc, L.kmn_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise KmnError(rc, self._L.kmn_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.kmn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._ck(self._L.kmn_reset(self._h))

    @staticmethod
    def comm_unique_id():
        buf = np.zeros(128, dtype=np.uint8)
        rc = load().kmn_comm_unique_id(buf.ctypes.data)
        if rc:
            raise KmnError(rc, "kmn_comm_unique_id failed")
        return buf

    def comm_init(self, rank, nranks, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._ck(self._L.kmn_comm_init(self._h, rank, nranks, uid.ctypes.data))

    def count_batch(self, bases, quals, read_off, n_reads=None, discarded=None):
        if n_reads is None:
            n_reads = len(read_off) - 1
        self._keep = (bases, quals, read_off, discarded)
        self._ck(self._L.kmn_count_batch(self._h, _ptr(bases), _ptr(quals), _ptr(read_off), n_reads, _ptr(discarded)))

    def count_batch_2na(self, packed, packed_off, quals, read_off, n_reads=None, discarded=None, markup_pos=None, markup_chr=None):
        """kmn_count_batch_2na: TwoBitSequence-packed bases + markups (see pack_2na)"""
        if n_reads is None:
            n_reads = len(read_off) - 1
        n_mk = 0 if markup_pos is None else len(markup_pos)
        self._keep = (packed, packed_off, quals, read_off, discarded, markup_pos, markup_chr)
        self._ck(self._L.kmn_count_batch_2na(self._h, _ptr(packed), _ptr(packed_off), _ptr(quals), _ptr(read_off), n_reads, _ptr(discarded),
                                             _ptr(markup_pos), _ptr(markup_chr), n_mk))

    def count_finish(self, apply_purge=True):
        self._ck(self._L.kmn_count_finish(self._h, int(apply_purge)))

    def stats(self):
        s = KmnStats()
        self._ck(self._L.kmn_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in KmnStats._fields_}

    def purge_min_depth(self, min_depth):
        self._ck(self._L.kmn_purge_min_depth(self._h, min_depth))

    def histogram(self, with_weights=False):
        h = np.zeros(65536, dtype=np.uint64)
        w = np.zeros(65536, dtype=np.float64) if with_weights else None
        self._ck(self._L.kmn_histogram(self._h, h.ctypes.data, _ptr(w)))
        return (h, w) if with_weights else h

    def lookup(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint8).reshape(-1, self.kb)
        out = np.zeros(len(keys), dtype=np.uint16)
        self._ck(self._L.kmn_lookup(self._h, keys.ctypes.data, len(keys), out.ctypes.data))
        return out

    def trim_batch(self, bases, read_off, min_depth, scoring, n_reads=None, discarded=None, out=None):
        if n_reads is None:
            n_reads = len(read_off) - 1
        if isinstance(scoring, str):
            scoring = SCORING[scoring]
        if out is None:
            out = (np.zeros(n_reads, np.uint32), np.zeros(n_reads, np.uint32), np.zeros(n_reads, np.float32), np.zeros(n_reads, np.uint8))
        self._ck(self._L.kmn_trim_batch(self._h, _ptr(bases), _ptr(read_off), n_reads, _ptr(discarded), min_depth, scoring,
                                        _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3])))
        return out

    def export(self, min_count=1, sort=True):
        """-> dict(keys[n,kb] u8, count u16, dir u16, wsum f32, ext[n,12] u32|None); sorted by key bytes when sort."""
        n = C.c_uint64()
        self._ck(self._L.kmn_export(self._h, min_count, None, None, None, None, None, 0, C.byref(n)))
        n = n.value
        keys = np.zeros((n, self.kb), np.uint8)
        cnt = np.zeros(n, np.uint16)
        dr = np.zeros(n, np.uint16)
        ws = np.zeros(n, np.float32)
        ext = np.zeros((n, 12), np.uint32) if (self.value_kind & KMN_VALUE_DIR_EXT) else None
        if n:
            got = C.c_uint64()
            self._ck(self._L.kmn_export(self._h, min_count, keys.ctypes.data, cnt.ctypes.data, dr.ctypes.data, ws.ctypes.data,
                                        _ptr(ext), n, C.byref(got)))
            assert got.value == n
            if sort:
                order = np.lexsort(keys.T[::-1])
                keys, cnt, dr, ws = keys[order], cnt[order], dr[order], ws[order]
                if ext is not None:
                    ext = ext[order]
        return dict(keys=keys, count=cnt, dir=dr, wsum=ws, ext=ext)

    def debug_owner(self, keys, nranks):
        keys = np.ascontiguousarray(keys, dtype=np.uint8).reshape(-1, self.kb)
        out = np.zeros(len(keys), dtype=np.uint32)
        self._ck(self._L.kmn_debug_owner(self._h, keys.ctypes.data, len(keys), nranks, out.ctypes.data))
        return out

    def import_entries(self, keys, count, dir=None, wsum=None, ext=None):
        """kmn_import: entries as export() returns them"""
        keys = np.ascontiguousarray(keys, dtype=np.uint8).reshape(-1, self.kb)
        count = np.ascontiguousarray(count, dtype=np.uint16)
        dir_ = np.ascontiguousarray(dir, dtype=np.uint16) if dir is not None else None
        wsum = np.ascontiguousarray(wsum, dtype=np.float32) if wsum is not None else None
        ext = np.ascontiguousarray(ext, dtype=np.uint32) if ext is not None else None
        self._ck(self._L.kmn_import(self._h, _ptr(keys), _ptr(count), _ptr(dir_), _ptr(wsum), _ptr(ext), len(count)))

    def subtract(self, other):
        """kmn_subtract: -> (entries removed, instances removed)"""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmn_subtract(self._h, other._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_kmers(self, bases, quals, read_off):
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n_reads = len(read_off) - 1
        n = C.c_uint64()
        self._ck(self._L.kmn_debug_kmers(self._h, _ptr(bases), _ptr(quals), read_off.ctypes.data, n_reads, None, None, None, None, C.byref(n)))
        nn = n
        n = n.value
        keys = np.zeros((n, self.kb), np.uint8)
        fw = np.zeros(n, np.uint8)
        wt = np.zeros(n, np.float32)
        hs = np.zeros(n, np.uint64)
        if n:
            self._ck(self._L.kmn_debug_kmers(self._h, _ptr(bases), _ptr(quals), read_off.ctypes.data, n_reads, keys.ctypes.data,
                                             fw.ctypes.data, wt.ctypes.data, hs.ctypes.data, C.byref(nn)))
        return keys, fw, wt, hs

    def profile_enable(self, on=True):
        self._ck(self._L.kmn_profile_enable(self._h, int(on)))

    def profile_read(self):
        p = KmnProfile()
        self._ck(self._L.kmn_profile_read(self._h, C.byref(p)))
        return {k: dict(ms=p.ms[i], launches=p.launches[i], units=p.units[i]) for i, k in enumerate(PROF_KINDS) if p.launches[i]}

    def sync(self):
        self._ck(self._L.kmn_sync(self._h))

    @property
    def stream(self):
        return self._L.kmn_stream(self._h)

    @property
    def launches(self):
        return self._L.kmn_launch_count(self._h)
