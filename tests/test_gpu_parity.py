"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact on integer work.
Run on the B200 box with `pytest -m gpu`."""
import os

import numpy as np
import pytest

import oracle
from bench import synth
from oracle import filter_oracle as F
from tests.gpu_util import assert_tables_equal, oracle_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import kmernator_b200 as K_
    K_.capi.load()
    return K_


def _fixture_reads(golden_dir, start=33):
    recs = F.parse_fastq(open(os.path.join(golden_dir, "1000.fastq")).read())
    F.normalise_quals(recs, start=start)
    return recs


def _tricky_reads():
    """N runs, lower case, '.', other IUPAC markups, low-quality tails, reads shorter than k, empty read."""
    rng = np.random.default_rng(5)
    seqs, quals = [], []
    for i in range(400):
        L = int(rng.integers(0, 260))
        s = bytearray(synth.ACGT[rng.integers(0, 4, L)].tobytes())
        q = bytearray(rng.integers(33 + 2, 33 + 42, L, dtype=np.uint8).tobytes())
        for _ in range(int(rng.integers(0, 4))):
            if L:
                p = int(rng.integers(0, L))
                s[p] = rng.choice(list(b"NnX.RYacgt"))
        if L > 40 and i % 3 == 0:
            q[L - 20:] = b"#" * 20
        if L > 10 and i % 7 == 0:
            q[3] = 33
        seqs.append(bytes(s))
        quals.append(bytes(q))
    return seqs, quals


@pytest.mark.parametrize("k", [1, 4, 15, 21, 31, 32, 33, 47, 48, 49, 63, 64, 65, 96, 97, 128])
def test_kmers_weights_hash(K, k):
    """steps (1)-(3): pack, canonical k-mers, fp32 weight, KmerHasher hash -- bit exact vs oracle"""
    seqs, quals = _tricky_reads()
    bases, q, off = oracle.concat_reads(seqs, quals)
    ctx = K.Context(kmer_size=k, table_slots=4096)
    keys, fw, wt, hs = ctx.debug_kmers(np.frombuffer(bases, np.uint8), q, off)
    pos = 0
    for s_, q_ in zip(seqs, quals):
        ok, ofw, owt, _ = oracle.read_kmers(s_, q_, k, with_ext=False)
        n = len(ok)
        assert (keys[pos:pos + n] == ok).all()
        assert (fw[pos:pos + n] == ofw).all()
        assert (wt[pos:pos + n].view(np.uint32) == owt.view(np.uint32)).all()
        for i in range(0, n, 7):
            assert int(hs[pos + i]) == oracle.kmer_hash(ok[i].tobytes())
        pos += n
    assert pos == len(keys)
    ctx.close()


def test_kmers_lookup8_hash(K):
    seqs, quals = _tricky_reads()
    bases, q, off = oracle.concat_reads(seqs[:50], quals[:50])
    for k in (21, 31, 63):
        ctx = K.Context(kmer_size=k, table_slots=4096, hash_kind=K.capi.KMN_HASH_LOOKUP8)
        keys, fw, wt, hs = ctx.debug_kmers(np.frombuffer(bases, np.uint8), q, off)
        for i in range(0, len(keys), 5):
            assert int(hs[i]) == oracle.kmer_hash_lookup8(keys[i].tobytes())
        ctx.close()


def test_long_read_reseed(K):
    """weight re-seed every 1024 positions (KmerReadUtils.h:204) on a 5 kb read with varied qualities"""
    rng = np.random.default_rng(11)
    L = 5000
    s = synth.ACGT[rng.integers(0, 4, L)].tobytes()
    q = rng.integers(33 + 3, 33 + 41, L, dtype=np.uint8).tobytes()
    bases, qq, off = oracle.concat_reads([s], [q])
    ctx = K.Context(kmer_size=31, table_slots=1 << 14)
    keys, fw, wt, hs = ctx.debug_kmers(np.frombuffer(bases, np.uint8), qq, off)
    ok, ofw, owt, _ = oracle.read_kmers(s, q, 31, with_ext=False)
    assert (keys == ok).all() and (wt.view(np.uint32) == owt.view(np.uint32)).all()
    ctx.close()


@pytest.mark.parametrize("k,start", [(31, 33), (31, 64), (21, 33), (63, 33), (32, 33)])
def test_count_table_1000_fastq(K, golden_dir, k, start):
    """config C1: k-mer -> count table of test/1000.fastq identical to the oracle (which reproduces the goldens)"""
    recs = _fixture_reads(golden_dir, start)
    bases, q, off, disc = F.to_buffers(recs)
    ctx = K.Context(kmer_size=k, fastq_start_char=start, table_slots=1 << 18, value_kind=K.capi.KMN_VALUE_WEIGHTS)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off)
    ctx.count_finish(apply_purge=False)
    g = ctx.export()
    o = oracle_table(bases, q, off, k, start=start).export()
    assert_tables_equal(g, o, check_wsum=True)
    st, ost = ctx.stats(), oracle_table(bases, q, off, k, start=start).stats()
    assert st["raw_kmers"] == ost["raw"] and st["raw_good_kmers"] == ost["raw_good"]
    assert st["unique_kmers"] == ost["unique"] and st["singleton_kmers"] == ost["singleton"]
    ctx.close()


def test_meraculous_golden(K, golden_dir):
    """config C5b on the reference fixture: counts and extension counters equal phix.mercount.m21 / phix.mergraph.m21.D2"""
    recs = _fixture_reads(golden_dir, 33)
    bases, q, off, _ = F.to_buffers(recs)
    ctx = K.Context(kmer_size=21, min_quality_score=2, min_kmer_quality=0.0, table_slots=1 << 18,
                    value_kind=K.capi.KMN_VALUE_DIR_EXT)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off)
    ctx.count_finish(apply_purge=False)
    g = ctx.export(min_count=2)
    idx = "ACGTNX"
    counts, graph = set(), set()
    for key, c, ext in zip(g["keys"], g["count"], g["ext"]):
        km = "".join("ACGT"[(key[i >> 2] >> (6 - 2 * (i & 3))) & 3] for i in range(21))
        rc = F.revcomp(km)
        counts |= {"%s\t%d" % (km, c), "%s\t%d" % (rc, c)}
        re = [0] * 12
        for i, b in enumerate(idx):
            j = idx.index(F.COMP[b])
            re[j] = int(ext[6 + i])
            re[6 + j] = int(ext[i])
        graph |= {km + "\t" + " ".join(str(int(x)) for x in ext) + " 0", rc + "\t" + " ".join(map(str, re)) + " 0"}
    assert counts == set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mercount.m21")))
    assert graph == set(l.rstrip("\n") for l in open(os.path.join(golden_dir, "phix.mergraph.m21.D2")))
    ctx.close()


@pytest.mark.parametrize("k,n_reads,kw", [
    (31, 20000, dict()),
    (31, 20000, dict(stage_keys=1 << 16, slice_bytes=1 << 16)),      # many drains, many partitions
    (31, 6000, dict(stage_keys=1 << 16, slice_bytes=1 << 12)),       # partition count capped by shared memory
    (63, 8000, dict(slice_bytes=1 << 18)),
    (64, 4000, dict()),
    (96, 3000, dict()),
    (21, 8000, dict()),
])
def test_count_table_synthetic(K, k, n_reads, kw):
    """config C2-shaped reads (scaled down): identical table, histogram and stats; multi-batch input"""
    bases, q, off = synth.reads_numpy(n_reads, 150, 40000, seed=3, err=0.002, lowq=0.001, n_rate=0.0005)
    ctx = K.Context(kmer_size=k, table_slots=1 << 22, **kw)
    half = n_reads // 2
    cut = int(off[half])
    ctx.count_batch(bases[:cut], q[:cut], np.ascontiguousarray(off[:half + 1]))
    ctx.count_batch(bases[cut:], q[cut:], np.ascontiguousarray(off[half:] - off[half]))
    ctx.count_finish(apply_purge=False)
    osp = oracle_table(bases, q, off, k, threads=4)
    assert_tables_equal(ctx.export(), osp.export())
    gh = ctx.histogram()
    for zm in (255, 256):                # serial apps print Histogram(256), FilterReads-P MPIHistogram(255)
        ov, oc, ow = osp.histogram(zm)
        gv, gc = np.zeros(len(ov), np.uint64), np.zeros(len(ov), np.uint64)
        for c in np.nonzero(gh)[0]:
            b = oracle.histogram_bin(int(c), zm)
            gv[b] += gh[c]
            gc[b] += gh[c] * np.uint64(c)
        assert (gv == ov).all() and (gc == oc).all()
    exact = np.zeros(65536, np.uint64)
    for c in np.unique(osp.export()["count"]):
        exact[c] = (osp.export()["count"] == c).sum()
    assert (gh == exact).all()
    st, ost = ctx.stats(), osp.stats()
    assert (st["raw_kmers"], st["raw_good_kmers"], st["unique_kmers"], st["singleton_kmers"]) == (ost["raw"], ost["raw_good"], ost["unique"], ost["singleton"])
    # purge: singletons dropped at min_depth 2, count<3 dropped at 3
    for md in (2, 3):
        ctx.purge_min_depth(md)
        osp.purge_min_depth(md)
        assert_tables_equal(ctx.export(), osp.export())
    ctx.close()


@pytest.mark.parametrize("k", [31, 63])
def test_histogram_weight_column(K, k):
    """a9: KmerSpectrum::Histogram (src/KmerSpectrum.h:909-1057) with the weight column -- visits and visitedCount exact,
    visitedWeight within tolerance (a sum of fp32 weightedCounts, order-dependent in the reference too; an entry that was
    promoted from a singleton carries the 1/254 quantisation of its first weight, src/KmerTrackingData.h:641-661, which the
    GPU table does not reproduce for counts >= 2), for both zoomMax values and before / after the singleton purge"""
    bases, q, off = synth.reads_numpy(12000, 150, 25000, seed=21, err=0.004, lowq=0.002, n_rate=0.0005)
    rng = np.random.default_rng(2)
    q = q.copy()
    q[rng.random(len(q)) < 0.3] = 33 + 25          # varied qualities: weights well below 1
    ctx = K.Context(kmer_size=k, table_slots=1 << 22, value_kind=K.capi.KMN_VALUE_WEIGHTS)
    ctx.count_batch(bases, q, off)
    ctx.count_finish(apply_purge=False)
    osp = oracle_table(bases, q, off, k, threads=4)
    for purge in (0, 2):
        if purge:
            ctx.purge_min_depth(purge)
            osp.purge_min_depth(purge)
        gh, gw = ctx.histogram(with_weights=True)
        for zm in (255, 256):
            ov, oc, ow = osp.histogram(zm)
            gv, gc, gws = np.zeros(len(ov), np.uint64), np.zeros(len(ov), np.uint64), np.zeros(len(ov), np.float64)
            for c in np.nonzero(gh)[0]:
                b = oracle.histogram_bin(int(c), zm)
                gv[b] += gh[c]
                gc[b] += gh[c] * np.uint64(c)
                gws[b] += gw[c]
            if purge:                    # the reference keeps counting purged singletons as (1, 1.0) records (:1046-1048)
                ps = osp.stats()["purged_singletons"]
                assert ov[1] == ps and gv[1] == 0
                ov, oc, ow = ov.copy(), oc.copy(), ow.copy()
                ov[1] = oc[1] = 0
                ow[1] = 0.0
            assert (gv == ov).all() and (gc == oc).all()
            assert np.allclose(gws, ow, rtol=3e-3, atol=1e-3)
            assert ow[2:].sum() > 0
    ctx.close()


def test_many_small_batches_spread_over_ctas(K):
    """--batch-size-like input (MPI path of the reference: 8192 reads per batch, src/DistributedFunctions.h:360): a
    long sequence of small batches must fill the per-CTA staging sub-regions evenly (no direct inserts) and count right"""
    n_reads, per = 24000, 400
    bases, q, off = synth.reads_numpy(n_reads, 150, 30000, seed=9, err=0.002, lowq=0.001)
    ctx = K.Context(kmer_size=31, table_slots=1 << 22, stage_keys=1 << 22)
    for r0 in range(0, n_reads, per):
        b0, b1 = int(off[r0]), int(off[r0 + per])
        ctx.count_batch(bases[b0:b1], q[b0:b1], np.ascontiguousarray(off[r0:r0 + per + 1] - off[r0]))
    ctx.count_finish(apply_purge=False)
    assert ctx.stats()["direct_inserts"] == 0
    assert_tables_equal(ctx.export(), oracle_table(bases, q, off, 31, threads=4).export())
    ctx.close()


def test_large_scale_invariants(K):
    """sizes the oracle cannot finish: size-independent properties on 3 M reads generated in HBM --
    every counted instance is in the table exactly once, a second build of the same input gives the same spectrum,
    the min-depth-2 purge removes exactly the singletons"""
    import torch
    n_reads = 3_000_000
    bases, quals, off = synth.reads_torch(n_reads, 150, 7_500_000, seed=5, device="cuda", read_seed=6)
    ctx = K.Context(kmer_size=31, est_raw_kmers=n_reads * 120, table_slots=1 << 27, stage_keys=n_reads * 120 // 5)
    hists = []
    for _ in range(2):
        ctx.reset()
        ctx.count_batch(bases, quals, off, n_reads=n_reads)
        ctx.count_finish(apply_purge=False)
        st = ctx.stats()
        h = ctx.histogram()
        hists.append(h)
        assert st["raw_kmers"] == n_reads * 120 and st["direct_inserts"] == 0
        assert int((h.astype(np.float64) * np.arange(65536)).sum()) == st["raw_good_kmers"] and h[65535] == 0
        assert int(h.sum()) == st["unique_kmers"] and int(h[1]) == st["singleton_kmers"]
    assert (hists[0] == hists[1]).all()
    ctx.purge_min_depth(2)
    hp = ctx.histogram()
    assert hp[1] == 0 and (hp[2:] == hists[0][2:]).all()
    # the lookup pass sees what the table holds: a read made of genomic k-mers (depth 60) keeps its full length
    outs = ctx.trim_batch(bases[: 150 * 100_000], off[: 100_001], 2, "MAX", n_reads=100_000)
    assert (outs[1] == 150).mean() > 0.7 and outs[2].max() <= 65535
    ctx.close()
    del bases, quals, off
    torch.cuda.empty_cache()


def test_count_saturation(K):
    """uint16 saturating count (KmerTrackingData.h:306): 70000 copies of one read -> every count == 65535"""
    seq = b"ACGTTGCAAGGCTTAACCGGATATCGCGATTACGGATCCA"
    n = 70000
    bases, q, off = oracle.concat_reads([seq] * n)
    ctx = K.Context(kmer_size=31, table_slots=1 << 12)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off)
    ctx.count_finish(apply_purge=False)
    g = ctx.export()
    assert len(g["count"]) == 10 and (g["count"] == 65535).all()
    st = ctx.stats()
    assert st["raw_kmers"] == n * 10 and st["unique_kmers"] == 10
    ctx.close()


def test_table_full_is_an_error(K):
    bases, q, off = synth.reads_numpy(3000, 150, 200000, seed=9)
    ctx = K.Context(kmer_size=31, table_slots=1024)
    ctx.count_batch(bases, q, off)
    with pytest.raises(K.KmnError) as ei:
        ctx.count_finish()
    assert ei.value.code == -4
    ctx.close()


def test_lookup_api(K):
    bases, q, off = synth.reads_numpy(3000, 150, 20000, seed=4)
    ctx = K.Context(kmer_size=31, table_slots=1 << 20)
    ctx.count_batch(bases, q, off)
    ctx.count_finish(apply_purge=True)
    osp = oracle_table(bases, q, off, 31)
    osp.purge_min_depth(2)
    e = osp.export()
    got = ctx.lookup(e["keys"])
    assert (got == e["count"]).all()
    absent = np.random.default_rng(1).integers(0, 256, (100, 8), dtype=np.uint8)
    absent[:, 7] &= 0xFC
    exp = np.array([osp.lookup(k.tobytes()) for k in absent], dtype=np.uint16)
    assert (ctx.lookup(absent) == exp).all()
    ctx.close()


@pytest.mark.parametrize("scoring", ["MAX", "MEDIAN", "MIN", "AVG", "SUM"])
@pytest.mark.parametrize("k", [31, 63])
def test_trim_scores(K, scoring, k):
    """lookup pass: (trimOffset, trimLength, score, wasTrimmed) per read identical to the oracle"""
    bases, q, off = synth.reads_numpy(6000, 150, 30000, seed=8, err=0.01, lowq=0.002, n_rate=0.002, var_len=True)
    disc = (np.arange(6000) % 97 == 0).astype(np.uint8)
    ctx = K.Context(kmer_size=k, table_slots=1 << 21)
    ctx.count_batch(bases, q, off, discarded=disc)
    ctx.count_finish(apply_purge=True)
    osp = oracle_table(bases, q, off, k, disc=disc, threads=4)
    osp.purge_min_depth(2)
    for md in (2, 3):
        g = ctx.trim_batch(bases, off, md, scoring, discarded=disc)
        o = osp.trim_reads(bases, off, md, oracle.SCORING[scoring], disc, threads=4)
        for a, b, name in zip(g, o, ("off", "len", "score", "was")):
            assert (a == b).all(), name
    ctx.close()


FILTER_CASES = [
    ("1000-Filtered-0.85.fastq", 0.85, 1, 64),
    ("1000-Filtered-0.85.std.fastq", 0.85, 1, 33),
    ("1000-Filtered-readlength.fastq", 1.0, 1, 64),
    ("1000-Filtered-readlength-both.fastq", 1.0, 2, 64),
    ("1000-Filtered.fastq", 25.0, 1, 64),
]


@pytest.mark.parametrize("fn,minlen,both,start", FILTER_CASES)
def test_filter_goldens_gpu(K, golden_dir, fn, minlen, both, start):
    """config C1: FilterReads on test/1000.fastq with the count and lookup passes on the GPU reproduces the
    reference's golden output byte for byte (test/runFilterTests.sh:26,44-76)"""
    recs = _fixture_reads(golden_dir, start)
    F.artifact_quality_trim(recs, start, 3, minlen)
    bases, q, off, disc = F.to_buffers(recs)
    ctx = K.Context(kmer_size=31, fastq_start_char=start, table_slots=1 << 18)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off, discarded=disc)
    ctx.count_finish(apply_purge=True)
    toff, tlen, score, wast = ctx.trim_batch(np.frombuffer(bases, np.uint8), off, 2, "MEDIAN", discarded=disc)
    res = []
    for i, r in enumerate(recs):
        if r["discarded"]:
            res.append(dict(label="", passes=False, off=0, len=0))
            continue
        label = ("Trim:%d+%d" % (toff[i], tlen[i])) if wast[i] else ""
        label += (" " if label else "") + "MedianScore:%d" % int(score[i] + np.float32(0.5))
        res.append(dict(label=label, passes=bool(score[i] >= 2) and oracle.passes_length(tlen[i], len(r["seq"]), minlen),
                        off=int(toff[i]), len=int(tlen[i])))
    out = []
    for i, j in F.identify_pairs(recs):
        r1, r2 = res[i]["passes"], res[j]["passes"]
        if (r1 and r2) if both >= 2 else (r1 or r2):
            out += [F.format_fastq(recs[i], res[i], start), F.format_fastq(recs[j], res[j], start)]
    assert "".join(out) == open(os.path.join(golden_dir, fn)).read()
    ctx.close()


@pytest.mark.parametrize("k", [21, 31, 63])
def test_owner_rule_on_device(K, k):
    """a5 on one GPU: the device's owner_of(hash) for 1..16 ranks equals getDistributedThreadId
    ((hashlittle2 >> 24) & 0x7ffff) % R (src/Kmer.h:2284-2295) computed by the oracle -- the rule every multi-GPU kernel
    bins by"""
    bases, q, off = synth.reads_numpy(400, 150, 20000, seed=31)
    keys = oracle_table(bases, q, off, k).export()["keys"]
    ctx = K.Context(kmer_size=k, table_slots=4096)
    for nranks in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 100, 1000):       # (kmn_debug_owner answers 0xffffffff when its two evaluations differ)
        got = ctx.debug_owner(keys, nranks)
        want = np.array([oracle.owner(oracle.kmer_hash(x.tobytes()), nranks) for x in keys], dtype=np.uint32)
        assert (got == want).all(), nranks
        assert nranks == 1 or nranks > 16 or len(np.unique(got)) == nranks
    ctx.close()
    ctx8 = K.Context(kmer_size=k, table_slots=4096, hash_kind=K.capi.KMN_HASH_LOOKUP8)
    got = ctx8.debug_owner(keys[:500], 8)
    assert (got == np.array([oracle.owner(oracle.kmer_hash_lookup8(x.tobytes()), 8) for x in keys[:500]], dtype=np.uint32)).all()
    ctx8.close()


@pytest.mark.parametrize("k,vk", [(31, 0), (31, 2), (63, 0), (21, 1)])
def test_import_and_subtract(K, k, vk):
    """f4: kmn_import restores exported entries (the --load-kmer-mmap path) and merges duplicates; kmn_subtract removes the
    k-mers of a second spectrum (--subtract-file / --reference-file, src/KmerSpectrum.h:1582-1589)"""
    bases, q, off = synth.reads_numpy(6000, 150, 30000, seed=12, err=0.003, lowq=0.001)
    a = K.Context(kmer_size=k, table_slots=1 << 21, value_kind=vk)
    a.count_batch(bases, q, off)
    a.count_finish(apply_purge=False)
    ea = a.export()
    b = K.Context(kmer_size=k, table_slots=1 << 21, value_kind=vk)
    half = len(ea["count"]) // 2
    for sl in (slice(0, half), slice(half, None)):
        b.import_entries(ea["keys"][sl], ea["count"][sl], ea["dir"][sl], ea["wsum"][sl], ea["ext"][sl] if ea["ext"] is not None else None)
    eb = b.export()
    assert (eb["keys"] == ea["keys"]).all() and (eb["count"] == ea["count"]).all() and (eb["dir"] == ea["dir"]).all()
    if vk == 2:
        assert np.allclose(eb["wsum"], ea["wsum"], rtol=1e-6)
    if vk == 1:
        assert (eb["ext"] == ea["ext"]).all()
    assert b.stats()["unique_kmers"] == len(ea["count"])
    assert (b.lookup(ea["keys"][:1000]) == ea["count"][:1000]).all()
    b.import_entries(ea["keys"][:100], ea["count"][:100])                    # merge: counts add up
    assert (b.lookup(ea["keys"][:100]).astype(np.int64) == np.minimum(2 * ea["count"][:100].astype(np.int64), 65535)).all()
    # subtract the spectrum of the first 2000 reads (min depth 2) from the full one
    n_sub = 2000
    cut = int(off[n_sub])
    s = K.Context(kmer_size=k, table_slots=1 << 21)
    s.count_batch(bases[:cut], q[:cut], np.ascontiguousarray(off[:n_sub + 1]))
    s.count_finish(apply_purge=True)
    es = s.export()
    removed, inst = a.subtract(s)
    drop = set(x.tobytes() for x in es["keys"])
    keep = np.array([x.tobytes() not in drop for x in ea["keys"]])
    assert removed == int((~keep).sum()) and inst == int(ea["count"][~keep].astype(np.int64).sum())
    e2 = a.export()
    assert (e2["keys"] == ea["keys"][keep]).all() and (e2["count"] == ea["count"][keep]).all()
    for c in (a, b, s):
        c.close()


@pytest.mark.parametrize("k", [31, 63])
def test_count_batch_2na_equals_ascii(K, k):
    """kmn_count_batch_2na (the reference's in-memory read: TwoBitSequence bytes + markups, src/Sequence.h:372-380) counts
    exactly what kmn_count_batch counts on the ASCII form of the same reads -- N / X / '.' markups, ragged lengths, reads
    shorter than k, an empty read, several batches through the staging slots"""
    bases, q, off = synth.reads_numpy(5000, 150, 30000, seed=41, err=0.003, lowq=0.002, n_rate=0.004, var_len=True)
    bases = bases.copy()
    bases[::997] = ord("X")
    bases[5::1201] = ord(".")
    off = off.copy()
    a = K.Context(kmer_size=k, table_slots=1 << 21)
    b = K.Context(kmer_size=k, table_slots=1 << 21)
    a.count_batch(bases, q, off)
    a.count_finish(apply_purge=False)
    step = 700
    for r0 in range(0, 5000, step):
        r1 = min(5000, r0 + step)
        b0, b1 = int(off[r0]), int(off[r1])
        o = np.ascontiguousarray(off[r0:r1 + 1] - off[r0])
        packed, poff, mpos, mchr = K.capi.pack_2na(bases[b0:b1], o)
        b.count_batch_2na(packed, poff, np.ascontiguousarray(q[b0:b1]), o, markup_pos=mpos if len(mpos) else None, markup_chr=mchr if len(mpos) else None)
    b.count_finish(apply_purge=False)
    assert_tables_equal(a.export(), b.export())
    sa, sb = a.stats(), b.stats()
    assert (sa["raw_kmers"], sa["raw_good_kmers"], sa["unique_kmers"]) == (sb["raw_kmers"], sb["raw_good_kmers"], sb["unique_kmers"])
    assert_tables_equal(b.export(), oracle_table(bases, q, off, k, threads=4).export())
    a.close()
    b.close()


@pytest.mark.parametrize("k,vk,n_reads", [(21, 1, 60000), (31, 2, 40000), (21, 0, 60000), (32, 0, 30000)])
def test_hot_kmers_tiny_genome(K, k, vk, n_reads):
    """N1 / config C5b scaled down: a 5386-bp genome at enormous depth -- every staged chunk repeats a handful of keys, the
    shared-memory pre-aggregation path of the insert kernel (and the sub-region / sub-run overflow paths around it) must
    give the oracle's table: counts, directionBias, weights, extension counters"""
    bases, q, off = synth.reads_numpy(n_reads, 150, 5386, seed=0x50, err=0.001, lowq=0.0005)
    kw = dict(min_quality_score=2, min_kmer_quality=0.0) if vk == 1 else {}
    ctx = K.Context(kmer_size=k, table_slots=1 << 20, value_kind=vk, stage_keys=1 << 21, **kw)
    ctx.count_batch(bases, q, off)
    ctx.count_finish(apply_purge=False)
    osp = oracle_table(bases, q, off, k, threads=4, track_ext=(vk == 1), min_quality=kw.get("min_quality_score", 3),
                       min_kmer_quality=kw.get("min_kmer_quality", 0.10))
    g, o = ctx.export(), osp.export()
    assert_tables_equal(g, o, check_wsum=(vk == 2), check_ext=(vk == 1))
    assert int(g["count"].max()) > 500
    st, ost = ctx.stats(), osp.stats()
    assert (st["raw_kmers"], st["raw_good_kmers"], st["unique_kmers"]) == (ost["raw"], ost["raw_good"], ost["unique"])
    ctx.close()


def test_hot_kmers_saturate_with_extensions(K):
    """70000 copies of one read through the pre-aggregated path with extension tracking: every count saturates at 65535
    (src/KmerTrackingData.h:427-448)"""
    seq = b"ACGTTGCAAGGCTTAACCGGATATCGCGATTACGGATCCA"
    n = 70000
    bases, q, off = oracle.concat_reads([seq] * n)
    ctx = K.Context(kmer_size=21, table_slots=1 << 12, value_kind=K.capi.KMN_VALUE_DIR_EXT)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off)
    ctx.count_finish(apply_purge=False)
    g = ctx.export()
    assert len(g["count"]) == 20 and (g["count"] == 65535).all()
    assert ctx.stats()["raw_kmers"] == n * 20
    ctx.close()


@pytest.mark.parametrize("k,mkq,mq", [(31, 0.10, 3), (21, 0.0, 2), (31, 0.5, 3), (31, 0.9, 3), (63, 0.10, 3), (31, 0.05, 10), (32, 0.10, 3), (47, 0.10, 3)])
def test_weight_bound_mixed_qualities(K, k, mkq, mq):
    """phase 1a's bounded reads (a3): reads whose product of all non-zero base probabilities stays above
    min-kmer-quality get their "counted" bits without the recurrence, every other read walks it -- the table and the
    raw/rawGood counters must equal the oracle's for quality mixes on both sides of the bound, with zero-probability
    qualities and markups inside"""
    rng = np.random.default_rng(1000 + k)
    genome = synth.ACGT[rng.integers(0, 4, 3000)]
    seqs, quals = [], []
    for i in range(6000):
        L = int(rng.integers(20, 200))
        st = int(rng.integers(0, len(genome) - L))
        s = bytearray(genome[st:st + L].tobytes())
        kind = i % 6
        q = np.full(L, 33 + 40, np.uint8)
        if kind == 1:                                       # a few mediocre bases: bounded
            for p in rng.integers(0, L, 3):
                q[p] = 33 + int(rng.integers(8, 30))
        elif kind == 2:                                     # many mediocre bases: the bound fails, recurrence decides
            q = rng.integers(33 + mq, 33 + 25, L).astype(np.uint8)
        elif kind == 3:                                     # zero-probability qualities and markups
            for p in rng.integers(0, L, 2):
                q[p] = 33 + int(rng.integers(0, mq + 1))
            s[int(rng.integers(0, L))] = rng.choice(list(b"NnX.acgt"))
        elif kind == 4:                                     # one quality value, markups among the bases
            q[:] = 33 + int(rng.integers(mq, 41))
            s[int(rng.integers(0, L))] = ord("N")
        elif kind == 5:                                     # starts with a zero-probability quality
            q[0] = 33
            q[L // 2] = 33 + 12
        seqs.append(bytes(s))
        quals.append(q.tobytes())
    bases, q, off = oracle.concat_reads(seqs, quals)
    ctx = K.Context(kmer_size=k, table_slots=1 << 16, min_quality_score=mq, min_kmer_quality=mkq)
    ctx.count_batch(np.frombuffer(bases, np.uint8), q, off)
    ctx.count_finish(apply_purge=False)
    osp = oracle_table(bases, q, off, k, min_quality=mq, min_kmer_quality=mkq, threads=4)
    assert_tables_equal(ctx.export(), osp.export())
    st, ost = ctx.stats(), osp.stats()
    assert (st["raw_kmers"], st["raw_good_kmers"], st["unique_kmers"]) == (ost["raw"], ost["raw_good"], ost["unique"])
    ctx.close()
