"""Pins the CPU oracle (oracle/) against the reference's own golden fixtures.

Fixtures under tests/golden/ are byte copies of the reference's test data
(/root/reference/test/{1000.fastq, 1000-Filtered*.fastq, phix.mercount.m21, phix.mergraph.m21.D2}),
produced by the reference itself (test/runFilterTests.sh:26,44-76, test/runMeraculousTests.sh:39-75).
"""
import os

import numpy as np
import pytest

import oracle
from oracle import filter_oracle as F

# KmerHasher::getHash vectors generated from the reference's own src/lookup3.h
HASH_VECTORS = [
    ("0000000000000000", "68a314c951b5a5da"),
    ("1b1b1b1b1b1b1b18", "312dafc0b8588ebc"),
    ("0648349ace9df74c", "3fb16c3dfa0035aa"),
    ("24d2e184f20335fc", "a00fcf9dc2ad5258"),
    ("0002fe0f3a40", "549138987bbcf1f3"),
    ("1b" * 15 + "18", "2d9f6196bf6d6c35"),
]
LOOKUP8_VECTORS = [("0000000000000000", "c679d79b45fedb42"), ("1b1b1b1b1b1b1b18", "c69b76952a6b1640")]


def test_hash_vectors():
    for k, h in HASH_VECTORS:
        assert "%016x" % oracle.kmer_hash(bytes.fromhex(k)) == h
    for k, h in LOOKUP8_VECTORS:
        assert "%016x" % oracle.kmer_hash_lookup8(bytes.fromhex(k)) == h


def test_owner_vectors():
    # owner %2 / %4 / %8
    h = oracle.kmer_hash(bytes.fromhex("0648349ace9df74c"))
    assert (h >> 24) & 0x7FFFF == 278010
    assert [oracle.owner(h, n) for n in (2, 4, 8)] == [0, 2, 2]
    h = oracle.kmer_hash(bytes.fromhex("0002fe0f3a40"))
    assert (h >> 24) & 0x7FFFF == 39035
    assert [oracle.owner(h, n) for n in (2, 4, 8)] == [1, 3, 3]
    h = oracle.kmer_hash(bytes.fromhex("1b" * 15 + "18"))
    assert [oracle.owner(h, n) for n in (2, 4, 8)] == [1, 3, 7]


def test_hash_matches_reference_lookup3():
    """oracle/_ref = the reference's own lookup3.h compiled where it lies."""
    if oracle.ref_kmer_hash(b"\0") is None:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt file)")
    rng = np.random.default_rng(7)
    for n in list(range(1, 40)) + [64]:
        for _ in range(20):
            key = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            assert oracle.ref_kmer_hash(key) == oracle.kmer_hash(key)


def test_twobit_literals():
    # test/TwoBitSequenceTest.cpp:60-73 fixture strings fasta1..4 / rev1..4, n1
    fasta = ["ACGTCGTAGTACTACG", "ACGTCGTAGTACTACGA", "ACGTCGTAGTACTACGAC", "ACGTCGTAGTACTACGACT"]
    rev = ["CGTAGTACTACGACGT", "TCGTAGTACTACGACGT", "GTCGTAGTACTACGACGT", "AGTCGTAGTACTACGACGT"]
    for f, r in zip(fasta, rev):
        packed, markups = oracle.compress_sequence(f.encode())
        assert markups == []
        rc = oracle.reverse_complement(packed, len(f))
        exp, _ = oracle.compress_sequence(r.encode())
        assert rc.tobytes() == exp.tobytes()
    packed, markups = oracle.compress_sequence(b"NCGTCGTAGTACTACGACN.")
    assert markups == [("N", 0), ("N", 18), ("N", 19)]
    assert packed[0] == 0b00011011                       # N encoded as A
    assert oracle.first_markup_n_or_x(markups) == 1


def test_kmer_extraction_small_k():
    # test/KmerTest.cpp:253-300: kmers[i].toFasta() == S.substr(i,k) for k=1..12 (non-canonical there;
    # here the canonical key must decode to min(fwd, rc) and is_fwd must say which)
    S = "ACGTCGTAGTACTACGACGTAGCTTAGCCGATTAGC"
    for k in range(1, 13):
        keys, fw, wt, ext = oracle.read_kmers(S.encode(), b"I" * len(S), k)
        assert len(keys) == len(S) - k + 1
        for i in range(len(keys)):
            f = S[i : i + k]
            r = F.revcomp(f)
            dec = "".join("ACGT"[(keys[i][j >> 2] >> (6 - 2 * (j & 3))) & 3] for j in range(k))
            assert dec == min(f, r)
            assert bool(fw[i]) == (f <= r)


def test_quality_table():
    p = oracle.quality_table(3, 33)
    assert p[33 + 2] == 0.0 and p[33 + 3] == 1.0 - 10.0 ** (-0.3)
    assert p[102] == 1.0 - 10.0 ** ((33 - 102) / 10.0) and p[103] == 1.0 and p[255] == 1.0
    p64 = oracle.quality_table(3, 64)
    assert p64[64 + 2] == 0.0 and p64[104] == 1.0 and p64[102] == 1.0 - 10.0 ** (-3.8)


def _load(golden_dir, start):
    recs = F.parse_fastq(open(os.path.join(golden_dir, "1000.fastq")).read())
    base = F.normalise_quals(recs, start=start)
    assert base == 64
    return recs


def test_meraculous_goldens(golden_dir):
    recs = _load(golden_dir, 33)
    counts, graph = F.meraculous_counts(recs)
    assert counts == sorted(set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mercount.m21"))))
    assert graph == sorted(set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mergraph.m21.D2"))))


@pytest.mark.parametrize("threads", [1, 3])
def test_meraculous_goldens_threads(golden_dir, threads):
    recs = _load(golden_dir, 33)
    counts, graph = F.meraculous_counts(recs, threads=threads)
    assert len(counts) == 10802
    assert counts == sorted(set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mercount.m21"))))
    assert graph == sorted(set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mergraph.m21.D2"))))


FILTER_CASES = [
    ("1000-Filtered-0.85.fastq", 0.85, 1, 64),
    ("1000-Filtered-0.85.std.fastq", 0.85, 1, 33),
    ("1000-Filtered-readlength.fastq", 1.0, 1, 64),
    ("1000-Filtered-readlength-both.fastq", 1.0, 2, 64),
    ("1000-Filtered.fastq", 25.0, 1, 64),
]


@pytest.mark.parametrize("fn,minlen,both,start", FILTER_CASES)
def test_filter_goldens(golden_dir, fn, minlen, both, start):
    recs = _load(golden_dir, start)
    out, res, spec = F.filter_reads(recs, k=31, start=start, scoring="MEDIAN", min_read_length=minlen, min_passing_in_pair=both)
    gold = open(os.path.join(golden_dir, fn)).read()
    assert out.split() == gold.split()          # the reference's own check is `diff -w`
    assert out == gold                          # and in fact byte-identical


def test_filter_golden_threaded(golden_dir):
    recs = _load(golden_dir, 64)
    out, _, _ = F.filter_reads(recs, k=31, start=64, scoring="MEDIAN", min_read_length=25.0, threads=3)
    assert out == open(os.path.join(golden_dir, "1000-Filtered.fastq")).read()


def test_trim_values_rules():
    # first-longest run wins; SUM never assigns the score (src/ReadSelector.h:1151-1162)
    v = [5, 5, 0, 7, 7, 0, 9, 9]
    off, ln, sc, was = oracle.trim_values(v, 31, 0, 2, oracle.SCORING["MAX"])
    assert (off, ln, sc, was) == (0, 2 + 30, 5.0, True)
    off, ln, sc, was = oracle.trim_values(v, 31, 0, 2, oracle.SCORING["SUM"])
    assert sc == 0.0
    off, ln, sc, was = oracle.trim_values([0, 0, 0], 31, 0, 2, oracle.SCORING["MAX"])
    assert (off, ln, sc, was) == (0, 0, -1.0, True)
    off, ln, sc, was = oracle.trim_values([3, 4, 5, 6], 31, 0, 2, oracle.SCORING["MEDIAN"])
    assert (off, ln, sc, was) == (0, 4 + 30, 5.0, False)
    off, ln, sc, was = oracle.trim_values([3, 4, 5, 6], 31, 0, 2, oracle.SCORING["AVG"])
    assert sc == 4.5
    # markup at base 33 (markup_length 34) caps numKmers at 34-31=3
    off, ln, sc, was = oracle.trim_values([3, 4, 5, 6, 7, 8], 31, 34, 2, oracle.SCORING["MIN"])
    assert (off, ln, sc, was) == (0, 3 + 30, 3.0, False)


def test_passes_length_fp32():
    assert oracle.passes_length(64, 76, 0.85) is False      # 76*0.85 = 64.6 in fp32
    assert oracle.passes_length(65, 76, 0.85) is True
    assert oracle.passes_length(1, 76, 0.0) is False
    assert oracle.passes_length(25, 76, 25.0) is True


def test_histogram_bins():
    # zoomLogSkip = 7 for zoomMax 255/256 (src/KmerSpectrum.h:947-955)
    assert oracle.histogram_bin(256, 256) == 256
    assert oracle.histogram_bin(257, 256) == 257 and oracle.histogram_bin(511, 256) == 257
    assert oracle.histogram_bin(600, 256) == 258
    assert oracle.histogram_bin(65535, 256) == 264
    assert oracle.histogram_bin(256, 255) == 256 and oracle.histogram_bin(300, 255) == 256


def test_count_saturation_and_stats():
    # 70000 copies of one read: every 31-mer count saturates at 65535 (KmerTrackingData.h:306,427-448)
    seq = b"ACGTTGCAAGGCTTAACCGGATATCGCGATTACGGATCCA"
    n = 70000
    bases, q, off = oracle.concat_reads([seq] * n)
    s = oracle.OracleSpectrum(31)
    s.add_reads(bases, q, off)
    e = s.export()
    assert len(e["count"]) == len(seq) - 30 and (e["count"] == 65535).all()
    st = s.stats()
    assert st["raw"] == n * 10 and st["raw_good"] == n * 10 and st["unique"] == 10 and st["singleton"] == 0


def test_oracle_counts_vs_independent_bruteforce_on_skewed_input():
    """BASELINE config 4 scaled down (skewed abundances: a few genomes sampled 1x..200x): the C restatement against an
    independent pure-Python count of canonical k-mers (strings and a dict; no packing, no hashing, no rolling), for a
    single-word, an exactly-32 and a multi-word k.  Every base has a high quality, so a k-mer counts iff it has no
    non-ACGT character (markup -> weight 0, src/KmerReadUtils.h:218-222)."""
    rng = np.random.default_rng(11)
    genomes = ["".join("ACGT"[c] for c in rng.integers(0, 4, n)) for n in (900, 700, 500, 400)]
    depth = [1, 6, 40, 200]
    comp = str.maketrans("ACGT", "TGCA")
    reads = []
    for g, d in zip(genomes, depth):
        for _ in range(d * len(g) // 100):
            s = int(rng.integers(0, len(g) - 100))
            r = g[s:s + 100]
            if rng.random() < 0.5:
                r = r.translate(comp)[::-1]
            if rng.random() < 0.1:
                p = int(rng.integers(0, 100))
                r = r[:p] + "N" + r[p + 1:]
            reads.append(r)
    bases, quals, off = oracle.concat_reads([r.encode() for r in reads])
    for k in (21, 32, 45):
        want = {}
        for r in reads:
            for i in range(len(r) - k + 1):
                f = r[i:i + k]
                if "N" in f:
                    continue
                rc = f.translate(comp)[::-1]
                c = min(f, rc)
                want[c] = want.get(c, 0) + 1
        s = oracle.OracleSpectrum(k, est_distinct=1 << 14)
        s.add_reads(bases, quals, off)
        e = s.export()
        got = {}
        for key, cnt in zip(e["keys"], e["count"]):
            bits = "".join(format(b, "08b") for b in key.tobytes())
            got["".join("ACGT"[int(bits[2 * j:2 * j + 2], 2)] for j in range(k))] = int(cnt)
        assert got == want
        assert max(want.values()) > 150          # the skew is real: some k-mers are two orders deeper than others


def test_synthetic_read_generator_block_gather_equals_per_read_loop():
    """bench/synth.reads_numpy gathers fixed-length reads in blocks; the bytes must be those of the per-read loop the
    variable-length path still uses (the CPU sample of bench.py and the parity tests share this generator)"""
    import numpy as np
    from bench import synth
    n, L, G, seed = 3000, 150, 100_000, 5
    rng = np.random.default_rng(seed + 77)
    g = synth.genome_codes(G, seed)
    starts = rng.integers(0, G - L, n)
    strand = rng.integers(0, 2, n)
    want = np.empty(n * L, np.uint8)
    for i in range(n):
        seg = g[starts[i]: starts[i] + L]
        want[i * L: (i + 1) * L] = (3 - seg)[::-1] if strand[i] else seg
    bases, quals, off = synth.reads_numpy(n, L, G, seed=seed, err=0.0, lowq=0.0)
    assert (bases == synth.ACGT[want]).all() and (quals == ord("I")).all() and int(off[-1]) == n * L
