"""The C++ host mirror (kmernator_b200/host): FilterReads built against the C ABI.

CPU part: the driver builds, exposes the reference's option surface (names, defaults, prefix guessing, README aliases)
and fails loudly without a GPU.  GPU part (-m gpu): the reference's own integration test, test/runFilterTests.sh:26,44-76 --
FilterReads on test/1000.fastq must reproduce the four 1000-Filtered*.fastq goldens (`diff -w`)."""
import os
import subprocess

import pytest

from tests.gpu_util import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def filter_reads():
    from kmernator_b200 import build as kbuild
    from kmernator_b200.host import build as hbuild
    kbuild.build()
    return hbuild.build()


def _run(exe, args, cwd=None):
    return subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, timeout=300)


def _run_ranks(exe, args, world, cwd, comm_dir, gpus=False):
    """one process per rank, as torchrun / mpirun would start them (RANK / WORLD_SIZE / LOCAL_RANK)"""
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r if gpus else 0), KMN_COMM_DIR=comm_dir)
        procs.append(subprocess.Popen([exe] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    return [(p.returncode, o[1]) for p, o in zip(procs, outs)]


def test_filter_reads_p_without_kmers_matches_serial(filter_reads, golden_dir, tmp_path):
    """FilterReads-P with kmer-size 0 needs no GPU (artifact filter + length filter only): three ranks, each on its
    pair-aligned slice of the input, joined in rank order, write exactly what the serial driver writes
    (the reference's criterion for the distributed app, test/runFilterTests.sh:93-116)"""
    exe_p = filter_reads + "-P"
    args = ["--fastq-output-base-quality", "64", "--min-read-length", "25"]
    p = _run(filter_reads, args + ["--out", str(tmp_path / "serial"), "0", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    res = _run_ranks(exe_p, args + ["--out", str(tmp_path / "par"), "0", "1000.fastq"], 3, golden_dir, str(tmp_path / "comm"))
    assert all(rc == 0 for rc, _ in res), res
    assert open(str(tmp_path / "par-1000.fastq")).read() == open(str(tmp_path / "serial-1000.fastq")).read()
    assert not os.path.exists(str(tmp_path / "comm")) and not os.path.exists(str(tmp_path / "par-1000.fastq.rank1"))


def _norm_oracle(golden_dir, seed, depth, minlen=25.0, start=64, both=False):
    """expected output of `FilterReads --max-kmer-output-depth D ... 0 1000.fastq` (no k-mer work: score = read length after
    the artifact quality trim) from the oracle's restatement of pickCoverageNormalizedSubset and an MT19937 stream"""
    from oracle import filter_oracle as F
    import oracle
    recs = F.parse_fastq(open(os.path.join(golden_dir, "1000.fastq")).read())
    F.normalise_quals(recs, start=start)
    F.artifact_quality_trim(recs, start, 3, minlen)
    scores = [0 if r["discarded"] else len(r["seq"]) for r in recs]
    passes = [(not r["discarded"]) and s >= 2 and oracle.passes_length(s, len(r["seq"]), minlen) for r, s in zip(recs, scores)]
    pairs = F.identify_pairs(recs)
    rng = F.MT19937(seed)
    picked = F.pick_coverage_normalized_subset(scores, passes, pairs, depth, True, both, rng)
    text = "".join(F.format_fastq(recs[i], dict(label="", off=0, len=scores[i]), start) for i in picked)
    return text, picked, scores


@pytest.mark.parametrize("seed,depth,both", [(12345, 50, False), (7, 70, False), (99, 40, True)])
def test_random_normalisation_decisions_match_oracle(filter_reads, golden_dir, tmp_path, seed, depth, both):
    """a12: chooseRead / pickCoverageNormalizedSubset (src/ReadSelector.h:661-749) with an injected mt19937 stream make
    the identical keep / drop decisions as the oracle's restatement -- same picked set, same output bytes.  kmer-size 0
    keeps the test on the CPU (score = length of the artifact-filtered read)."""
    out = str(tmp_path / "n")
    args = ["--max-kmer-output-depth", str(depth), "--fastq-output-base-quality", "64", "--min-read-length", "25"]
    if both:
        args += ["--min-passing-in-pair", "2"]
    env = dict(os.environ, KMN_SEED=str(seed))
    p = subprocess.run([filter_reads] + args + ["--out", out, "0", "1000.fastq"], capture_output=True, text=True, cwd=golden_dir, env=env, timeout=600)
    assert p.returncode == 0, p.stderr
    want, picked, scores = _norm_oracle(golden_dir, seed, depth, both=both)
    assert 0 < len(picked) < 1000                     # the depth really subsamples
    assert open("%s-MaxDepth%d-1000.fastq" % (out, depth)).read() == want


def test_random_normalisation_kept_fraction(filter_reads, tmp_path):
    """a12, SURVEY 8: the kept count follows sum min(1, (D+1)/s) -- 20000 unpaired reads of random length, time-seeded
    RNG as in the reference, |kept - expectation| within 5 sigma"""
    import numpy as np
    from oracle import filter_oracle as F
    rng = np.random.default_rng(3)
    lens = rng.integers(60, 251, 20000)
    fq = tmp_path / "len.fastq"
    with open(fq, "w") as f:
        for i, L in enumerate(lens):
            f.write("@r%d\n%s\n+\n%s\n" % (i, "ACGT"[i % 4] * int(L), "I" * int(L)))
    out = str(tmp_path / "k")
    depth = 80
    p = _run(filter_reads, ["--max-kmer-output-depth", str(depth), "--skip-artifact-filter", "1", "--out", out, "0", str(fq)])
    assert p.returncode == 0, p.stderr
    kept = sum(1 for l in open("%s-MaxDepth%d-len.fastq" % (out, depth)).read().split("\n")[0::4] if l)
    exp, var = F.expected_kept_fraction(lens, depth)
    assert abs(kept - exp) <= 5.0 * var ** 0.5 + 1, (kept, exp, var)
    assert kept < 20000


def test_help_lists_reference_options(filter_reads):
    p = _run(filter_reads, ["--help"])
    assert p.returncode == 1                       # apps/FilterReads.cpp:85: `if (!parseOpts) exit(1)`
    for flag, default in [("min-kmer-quality", "0.10"), ("min-depth", "2"), ("min-read-length", "0.40"), ("kmer-scoring-type", "MAX"),
                          ("min-passing-in-pair", "1"), ("max-kmer-output-depth", "-1"), ("min-quality-score", "3"),
                          ("fastq-output-base-quality", "33"), ("batch-size", "100000"), ("skip-artifact-filter", "0"),
                          ("normalization-method", "RANDOM"), ("separate-outputs", "1"), ("kmers-per-bucket", "32"),
                          ("mpi-buffer-size", "33554432"), ("artifact-edit-distance", "2")]:
        assert "--%s arg (=%s)" % (flag, default) in p.stderr, flag


def test_option_errors(filter_reads, golden_dir):
    inp = os.path.join(golden_dir, "10.fastq")
    assert "unrecognised option" in _run(filter_reads, ["--no-such-flag", "1", "31", inp]).stderr
    assert "ambiguous" in _run(filter_reads, ["--min", "1", "31", inp]).stderr            # min-depth, min-read-length, ...
    assert "not implemented by kmernator_b200" in _run(filter_reads, ["--variant-sigmas", "2", "31", inp]).stderr
    assert "at least one input file" in _run(filter_reads, ["31"]).stderr


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_gpu(filter_reads, golden_dir, tmp_path):
    p = _run(filter_reads, ["--out", str(tmp_path / "x"), "31", os.path.join(golden_dir, "10.fastq")])
    assert p.returncode == 1
    assert "no CPU fallback" in p.stderr
    assert not list(tmp_path.iterdir())


CASES = [  # test/runFilterTests.sh:44-70
    ("1000.fastq", "1000-Filtered-0.85.std.fastq", ["--fastq-output-base-quality", "33", "--min-read-length", "0.85"]),
    ("1000.fastq", "1000-Filtered-0.85.fastq", ["--fastq-output-base-quality", "64", "--min-read-length", "0.85"]),
    ("1000.std.fastq", "1000-Filtered-0.85.std.fastq", ["--fastq-output-base-quality", "33", "--min-read-length", "0.85"]),
    ("1000.std.fastq", "1000-Filtered-0.85.fastq", ["--fastq-output-base-quality", "64", "--min-read-length", "0.85"]),
    ("1000.fastq", "1000-Filtered-readlength.fastq", ["--fastq-output-base-quality", "64", "--min-read-length", "1"]),
    ("1000.fastq", "1000-Filtered-readlength-both.fastq", ["--fastq-output-base-quality", "64", "--min-read-length", "1", "--min-passing-in-pair", "2"]),
    ("1000.fastq", "1000-Filtered.fastq", ["--fastq-output-base-quality", "64", "--min-read-length", "25", "--thread", "2"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("inp,good,extra", CASES)
def test_filter_reads_goldens(filter_reads, golden_dir, tmp_path, inp, good, extra):
    out = str(tmp_path / "out")
    # the fixed option set of test/runFilterTests.sh:26 (prefix-guessed --out as in the script)
    args = extra + ["--kmer-scoring-type", "MEDIAN", "--mask-simple-repeats", "0", "--artifact-edit-distance", "1", "--out", out, "31", inp]
    p = _run(filter_reads, args, cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    got = open("%s-MinDepth2-%s" % (out, inp)).read().split()          # diff -w
    want = open(os.path.join(golden_dir, good)).read().split()
    assert got == want


@pytest.mark.gpu
def test_filter_reads_p_two_ranks_reproduce_the_golden(filter_reads, golden_dir, tmp_path):
    """the reference's own test of the distributed app (test/runFilterTests.sh:93-116): FilterReads-P on two ranks (two
    GPUs, table sharded by hash owner, collective lookup pass) writes the golden file of the serial run"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "out")
    args = ["--fastq-output-base-quality", "64", "--min-read-length", "25", "--kmer-scoring-type", "MEDIAN", "--mask-simple-repeats", "0",
            "--artifact-edit-distance", "1", "--out", out, "31", "1000.fastq"]
    res = _run_ranks(filter_reads + "-P", args, 2, golden_dir, str(tmp_path / "comm"), gpus=True)
    assert all(rc == 0 for rc, _ in res), res
    got = open(out + "-MinDepth2-1000.fastq").read().split()
    want = open(os.path.join(golden_dir, "1000-Filtered.fastq")).read().split()
    assert got == want


@pytest.mark.gpu
def test_filter_reads_normalisation_and_aliases(filter_reads, golden_dir, tmp_path):
    """--max-kmer-depth / --min-kmer-depth (README.md:125 spellings) are aliases; RANDOM normalisation with a fixed
    seed keeps every read whose score is <= D and a subset of the others (src/ReadSelector.h:661-749)"""
    out = str(tmp_path / "norm")
    env = dict(os.environ, KMN_SEED="7")
    args = ["--max-kmer-depth", "4", "--min-kmer-depth", "2", "--kmer-scoring-type", "MEDIAN", "--fastq-output-base-quality", "64",
            "--out", out, "31", "1000.fastq"]
    p = subprocess.run([filter_reads] + args, capture_output=True, text=True, cwd=golden_dir, env=env, timeout=600)
    assert p.returncode == 0, p.stderr
    text = open(out + "-MinDepth2-MaxDepth4-1000.fastq").read().split("\n")
    hdrs = [l for l in text[0::4] if l]
    full = open(os.path.join(golden_dir, "1000-Filtered-0.85.fastq")).read().split("\n")[0::4]
    assert 0 < len(hdrs) < len([l for l in full if l])
    p2 = subprocess.run([filter_reads] + args, capture_output=True, text=True, cwd=golden_dir, env=env, timeout=600)
    assert p2.returncode == 0 and open(out + "-MinDepth2-MaxDepth4-1000.fastq").read().split("\n")[0::4][: len(hdrs)] == hdrs


def test_fasta_input_prints_ref_quality(filter_reads, tmp_path):
    """quality-less (FASTA) reads written as FASTQ carry PRINT_REF_QUAL = 33 + 70 = 'g' (src/config.h:140,
    src/Sequence.cpp:744-747)"""
    fa = tmp_path / "in.fasta"
    fa.write_text(">s1\nACGTACGTAC\nGTACGT\n>s2 note\nTTTTGGGGCCCCAAAA\n")
    out = str(tmp_path / "o")
    p = _run(filter_reads, ["--skip-artifact-filter", "1", "--out", out, "0", str(fa)])
    assert p.returncode == 0, p.stderr
    assert open(out + "-in.fastq").read() == "@s1\nACGTACGTACGTACGT\n+\n" + "g" * 16 + "\n@s2 note\nTTTTGGGGCCCCAAAA\n+\n" + "g" * 16 + "\n"


def test_phred_base_detection_examines_first_reads_only(filter_reads, tmp_path):
    """ReadSet::validateFastqStart (src/ReadSet.h:171-194) looks at the first 20000 reads: a high-quality Phred+33 read
    (minimum quality above 33 + 40) far into the file must not flip the base of the whole file"""
    fq = tmp_path / "hq.fastq"
    with open(fq, "w") as f:
        for i in range(20050):
            q = "K" * 20 if i >= 20010 else "I" * 19 + "5"
            f.write("@r%d\n%s\n+\n%s\n" % (i, "ACGT" * 5, q))
    out = str(tmp_path / "o")
    p = _run(filter_reads, ["--skip-artifact-filter", "1", "--out", out, "0", str(fq)])
    assert p.returncode == 0, p.stderr
    lines = open(out + "-hq.fastq").read().split("\n")
    assert lines[3] == "I" * 19 + "5" and lines[4 * 20020 + 3] == "K" * 20


MERA_ARGS = ["--fastq-base-quality", "64", "--thread", "2", "--min-kmer-quality=0", "--min-quality-score=2", "--kmer-size", "21"]


def _check_meraculous(out, golden_dir):
    for mine, good in ((out + ".mercount.m21", "phix.mercount.m21"), (out + ".mergraph.m21.D2", "phix.mergraph.m21.D2")):
        got = sorted(open(mine).read().splitlines())                   # `sort $TEST | diff - $GOOD`
        assert got == open(os.path.join(golden_dir, good)).read().splitlines(), good


@pytest.mark.gpu
def test_meraculous_counter_goldens(filter_reads, golden_dir, tmp_path):
    """the reference's own test of MeraculousCounter (test/runMeraculousTests.sh:39-75): k-mer counts and extension
    counters of test/1000.fastq, sorted, equal phix.mercount.m21 and phix.mergraph.m21.D2 line for line"""
    exe = os.path.join(os.path.dirname(filter_reads), "MeraculousCounter")
    out = str(tmp_path / "m")
    p = _run(exe, MERA_ARGS + ["--out", out, "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    _check_meraculous(out, golden_dir)


@pytest.mark.gpu
def test_meraculous_counter_two_ranks(filter_reads, golden_dir, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(os.path.dirname(filter_reads), "MeraculousCounter")
    out = str(tmp_path / "m")
    res = _run_ranks(exe, MERA_ARGS + ["--out", out, "1000.fastq"], 2, golden_dir, str(tmp_path / "comm"), gpus=True)
    assert all(rc == 0 for rc, _ in res), res
    _check_meraculous(out, golden_dir)


def test_meraculous_counter_requires_kmer_size(filter_reads, golden_dir):
    exe = os.path.join(os.path.dirname(filter_reads), "MeraculousCounter")
    p = _run(exe, ["--out", "/tmp/never", os.path.join(golden_dir, "10.fastq")])
    assert p.returncode == 1 and "can not be 0" in p.stderr


def _contaminated_fastq(path, n=3000, seed=4):
    """reads with adapter / homopolymer fragments (0-2 substitutions, any offset), low-quality heads (the quality-trimmed
    head shifts the screen's scan pointer) and low-quality bases in the middle (second stretch rescued as -qtrim)"""
    import numpy as np
    from oracle import filter_oracle as F
    rng = np.random.default_rng(seed)
    arts = [s for _, s in F.artifact_sequences()]
    with open(path, "w") as f:
        for i in range(n):
            L = int(rng.integers(60, 152))
            s = list("ACGT"[x] for x in rng.integers(0, 4, L))
            q = ["I"] * L
            kind = rng.random()
            if kind < 0.35:
                a = arts[int(rng.integers(0, len(arts)))].replace("N", "A")
                fl = int(rng.integers(24, 41))
                o = int(rng.integers(0, max(1, len(a) - fl + 1)))
                frag = list(a[o:o + fl])
                for _ in range(int(rng.integers(0, 3))):
                    p = int(rng.integers(0, len(frag)))
                    frag[p] = "ACGT"[("ACGT".index(frag[p]) + int(rng.integers(1, 4))) % 4]
                at = int(rng.integers(0, L - len(frag) + 1))
                if rng.random() < 0.6:
                    at &= ~3
                s[at:at + len(frag)] = frag
            if 0.25 < kind < 0.45:
                for p in range(int(rng.integers(1, 13))):
                    q[p] = "#"
            if 0.4 < kind < 0.5:
                q[int(rng.integers(20, L - 20))] = "!"
            if kind > 0.97:
                q = ["#"] * L
            f.write("@r%d%s\n%s\n+\n%s\n" % (i, " c%d" % i if i % 5 == 0 else "", "".join(s), "".join(q)))


@pytest.mark.parametrize("edit,build,minlen", [(2, 2, "0.4"), (1, 2, "0.4"), (2, 0, "0.4"), (1, 1, "40"), (0, 2, "0.4")])
def test_artifact_screen_matches_oracle(filter_reads, tmp_path, edit, build, minlen):
    """f2: FilterKnownOddities with the 24-mer adapter / homopolymer screen (src/FilterKnownOddities.h:190-286,389-541)
    against the oracle's numpy restatement on contaminated reads: same trims, discards and rescued remnants, byte for byte,
    for edits built into the filter and edits searched at run time"""
    from oracle import filter_oracle as F
    fq = tmp_path / "cont.fastq"
    _contaminated_fastq(str(fq))
    out = str(tmp_path / "o")
    p = _run(filter_reads, ["--artifact-edit-distance", str(edit), "--build-artifact-edits-in-filter", str(build), "--min-read-length", minlen,
                            "--out", out, "0", str(fq)])
    assert p.returncode == 0, p.stderr
    recs = F.parse_fastq(open(fq).read())
    n_trim, n_disc, n_rem = F.artifact_filter(recs, 33, 3, float(minlen), 24, edit, build)
    # the rescued "-qtrim" remnants join the read set (and the k-mer count of a k > 0 run) but have no pair entry
    # (ReadSet::append(const Read&), src/ReadSet.cpp:260-262; pairs are identified before the filter,
    # apps/FilterReads.cpp:103,114), so the pair-wise selection never writes them -- in the reference and here
    want = "".join(F.format_fastq(r, dict(label="", off=0, len=len(r["seq"])), 33) for r in recs
                   if not r["discarded"] and len(r["seq"]) > 1 and not r["name"].endswith("-qtrim"))
    got = open(out + "-cont.fastq").read()
    assert got == want
    assert n_trim > 300 and n_disc > 30 and n_rem > 20             # every branch was exercised
    if edit == 2:
        assert ("filter affected (trimmed/removed) %d Reads" % n_trim) in p.stderr


def _parse_map_file(path, kb, singletons=False):
    """reader of the reference's map-file layout (KmerMapByKmerArrayPair::store src/Kmer.h:3138-3155, KmerArrayPair::store
    :960-969): -> (numBuckets, mask, [(bucket, key bytes, value bytes)])"""
    import struct
    data = open(path, "rb").read()
    nb, mask = struct.unpack_from("<QQ", data, 0)
    offs = struct.unpack_from("<%dQ" % nb, data, 16)
    vsz = 1 if singletons else 12
    out = []
    for b in range(nb):
        (n,) = struct.unpack_from("<I", data, offs[b])
        k0 = offs[b] + 4
        v0 = k0 + n * kb
        for i in range(n):
            out.append((b, data[k0 + i * kb: k0 + (i + 1) * kb], data[v0 + i * vsz: v0 + (i + 1) * vsz]))
        if b + 1 < nb:
            assert offs[b + 1] == v0 + n * vsz                          # buckets are contiguous, in index order
    return nb, mask, out


@pytest.mark.gpu
def test_save_and_load_kmer_mmap(filter_reads, golden_dir, tmp_path):
    """f4: --save-kmer-mmap writes the weak map in the reference's file layout (bucket = KmerHasher hash & mask, keys of a
    bucket ascending, TrackingDataWithDirection values); --load-kmer-mmap on that file reproduces the filtered output"""
    import struct
    import numpy as np
    import oracle
    from oracle import filter_oracle as F
    opts = ["--fastq-output-base-quality", "64", "--min-read-length", "25", "--kmer-scoring-type", "MEDIAN"]
    out1, out2 = str(tmp_path / "a"), str(tmp_path / "b")
    p = _run(filter_reads, opts + ["--save-kmer-mmap", "1", "--out", out1, "31", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    nb, mask, ents = _parse_map_file(out1 + "-mmap", 8)
    assert nb & (nb - 1) == 0 and mask == nb - 1
    recs = F.parse_fastq(open(os.path.join(golden_dir, "1000.fastq")).read())
    F.normalise_quals(recs, start=64)
    F.artifact_quality_trim(recs, 64, 3, 25.0)
    bases, q, off, disc = F.to_buffers(recs)
    osp = oracle.OracleSpectrum(31, start=64, est_distinct=1 << 17)
    osp.add_reads(bases, q, off, disc)
    osp.purge_min_depth(2)
    o = osp.export()
    # the reference's sizing: next power of two >= (int)(rawKmers / estimated-depth) / kmers-per-bucket + 1 (src/Kmer.h:2837)
    raw = oracle.estimate_raw_kmers(len(recs), sum(len(r["seq"]) for r in recs), 31)
    want_nb = 1
    while want_nb < int(raw / 20.0) // 32 + 1:
        want_nb *= 2
    assert nb == want_nb
    assert len(ents) == len(o["count"])
    lut = {bytes(k): (int(c), int(d)) for k, c, d in zip(o["keys"], o["count"], o["dir"])}
    last = (-1, b"")
    for b, key, val in ents:
        assert oracle.kmer_hash(key) & mask == b
        assert (b, key) > last                                          # bucket order, then memcmp order inside a bucket
        last = (b, key)
        cnt, w, dr = struct.unpack("<HxxfHxx", val)
        assert cnt == lut[key][0] and dr - lut[key][1] in (0, 1) and 0.0 < w <= cnt
    p = _run(filter_reads, opts + ["--load-kmer-mmap", out1 + "-mmap", "--out", out2, "31", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    assert open(out2 + "-MinDepth2-1000.fastq").read() == open(out1 + "-MinDepth2-1000.fastq").read()
    assert open(out2 + "-MinDepth2-1000.fastq").read().split() == open(os.path.join(golden_dir, "1000-Filtered.fastq")).read().split()


@pytest.mark.gpu
def test_size_history_file(filter_reads, golden_dir, tmp_path):
    """f4: --size-history-file (SizeTracker, src/KmerSpectrum.h:812-900): one (rawKmers, rawGoodKmers, uniqueKmers,
    singletonKmers) sample per 5% growth at batch granularity; the last one is the final state of the spectrum"""
    import oracle
    from oracle import filter_oracle as F
    hist = str(tmp_path / "size.txt")
    p = _run(filter_reads, ["--fastq-output-base-quality", "64", "--batch-size", "100", "--size-history-file", hist, "--out", str(tmp_path / "o"),
                            "31", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    lines = open(hist).read().strip().split("\n")
    assert lines[0] == "rawKmers\trawGoodKmers\tuniqueKmers\tsingletonKmers"
    rows = [tuple(int(x) for x in l.split("\t")) for l in lines[1:]]
    assert len(rows) >= 8            # (SizeTracker::reset's track(0,0,0,0) is below nextToTrack and records nothing, :890-894)
    assert all(a[0] <= b[0] and a[2] <= b[2] for a, b in zip(rows, rows[1:]))
    recs = F.parse_fastq(open(os.path.join(golden_dir, "1000.fastq")).read())
    F.normalise_quals(recs, start=64)
    F.artifact_quality_trim(recs, 64, 3, 0.40)
    bases, q, off, disc = F.to_buffers(recs)
    osp = oracle.OracleSpectrum(31, start=64, est_distinct=1 << 17)
    osp.add_reads(bases, q, off, disc)
    st = osp.stats()
    # the post-build purge has removed the singletons by the time of the final forced sample (apps/FilterReads.cpp:139-141)
    assert rows[-1][:3] == (st["raw"], st["raw_good"], st["unique"]) and rows[-1][3] == 0
    assert any(r[3] > 0 for r in rows[1:-1])


@pytest.mark.gpu
def test_reference_file_subtraction(filter_reads, golden_dir, tmp_path):
    """f4: --reference-file (apps/FilterReads-P.cpp:281-308): k-mers of the reference spectrum leave the main spectrum.
    Subtracting the input from itself empties the spectrum; subtracting an unrelated sequence changes nothing."""
    exe_p = filter_reads + "-P"
    opts = ["--fastq-output-base-quality", "64", "--min-read-length", "25", "--kmer-scoring-type", "MEDIAN"]
    out = str(tmp_path / "s")
    p = _run(exe_p, opts + ["--reference-file", "1000.fastq", "--out", out, "31", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    assert "Subtracted" in p.stderr
    kept = open(out + "-MinDepth2-1000.fastq").read() if os.path.exists(out + "-MinDepth2-1000.fastq") else ""
    assert kept == ""
    other = tmp_path / "other.fasta"
    other.write_text(">x\n" + "ACGTTGCA" * 40 + "\n")
    out2 = str(tmp_path / "t")
    p = _run(exe_p, opts + ["--reference-file", str(other), "--out", out2, "31", "1000.fastq"], cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    assert open(out2 + "-MinDepth2-1000.fastq").read().split() == open(os.path.join(golden_dir, "1000-Filtered.fastq")).read().split()


@pytest.mark.parametrize("paired", [True, False])
def test_byte_range_slicing_tiles_the_file(filter_reads, tmp_path, paired):
    """f1: every rank of FilterReads-P parses only its byte range of the input, cut at record boundaries that never
    split two mates and never mistake a quality line starting with '@' for a header (src/ReadFileReader.h:379-398,
    657-760): for any number of ranks the joined output equals the serial run's"""
    import numpy as np
    rng = np.random.default_rng(8 + paired)
    fq = tmp_path / "in.fastq"
    with open(fq, "w") as f:
        for i in range(700):
            L = int(rng.integers(30, 90))
            s = "".join("ACGT"[x] for x in rng.integers(0, 4, L))
            q = "".join(chr(int(x)) for x in rng.integers(64, 74, L))          # '@' (64) is a frequent quality character here
            if i % 3 == 0:
                q = "@" + q[1:]
            name = ("p%d/%d" % (i // 2, 1 + i % 2)) if paired else ("s%d" % i)
            f.write("@%s\n%s\n+\n%s\n" % (name, s, q))
    args = ["--skip-artifact-filter", "1", "--fastq-base-quality", "33", "--min-read-length", "1"]
    p = _run(filter_reads, args + ["--out", str(tmp_path / "serial"), "0", str(fq)])
    assert p.returncode == 0, p.stderr
    want = open(str(tmp_path / "serial-in.fastq")).read()
    assert want.count("\n") == 4 * 700
    for world in (2, 3, 5, 8):
        out = str(tmp_path / ("par%d" % world))
        res = _run_ranks(filter_reads + "-P", args + ["--out", out, "0", str(fq)], world, str(tmp_path), str(tmp_path / ("comm%d" % world)))
        assert all(rc == 0 for rc, _ in res), res
        assert open(out + "-in.fastq").read() == want, world


@pytest.mark.gpu
def test_table_overflow_rebuilds_into_a_larger_table(filter_reads, golden_dir, tmp_path):
    """the reference's buckets grow as they fill (src/Kmer.h:3095-3110); here a table sized from a hopeless
    --estimated-depth guess overflows, and the adaptor repeats the build into tables four times as large until the
    spectrum fits: same golden output as with the default sizing, and the log says what happened"""
    out = str(tmp_path / "out")
    args = ["--estimated-depth", "1000000", "--estimated-error-rate", "0", "--fastq-output-base-quality", "64", "--min-read-length", "25",
            "--kmer-scoring-type", "MEDIAN", "--mask-simple-repeats", "0", "--artifact-edit-distance", "1", "--out", out, "31", "1000.fastq"]
    p = _run(filter_reads, args, cwd=golden_dir)
    assert p.returncode == 0, p.stderr
    assert "overflowed; rebuilding" in p.stderr
    got = open("%s-MinDepth2-1000.fastq" % out).read().split()
    assert got == open(os.path.join(golden_dir, "1000-Filtered.fastq")).read().split()
