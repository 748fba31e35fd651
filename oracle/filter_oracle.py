"""FilterReads / MeraculousCounter driver-level oracle -- TEST INFRASTRUCTURE ONLY.

Restates, in plain Python + numpy fp32, the host-side glue of the reference around the hot path, on top
of oracle/kmn_oracle.c (counting, lookup, trim).  Each function cites the reference file:line.
Used only on small inputs (the reference's 1000-read fixtures).
"""
import numpy as np

from . import binding as B

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "X": "X"}


def revcomp(s):
    return "".join(COMP.get(c, "N") for c in reversed(s))


def parse_fastq(text):
    """Minimal 4-line FASTQ reader; bases upper-cased like ReadFileReader::nextRead (src/ReadFileReader.h:299-322).
    Name = header up to first whitespace, rest = comment (src/Utils.h:561-598)."""
    lines = text.split("\n")
    recs = []
    for i in range(0, len(lines) - 3, 4):
        if not lines[i].startswith("@"):
            break
        hdr = lines[i][1:]
        name, _, comment = hdr.partition(" ")
        recs.append(dict(name=name, comment=comment, seq=lines[i + 1].upper(), qual=lines[i + 3], discarded=False))
    return recs


def normalise_quals(recs, start=33, input_base=None):
    """Re-express qualities in the internal base `start` (= --fastq-output-base-quality).
    Input base auto-detection: ReadSet::validateFastqStart / Read::validateFastqStart
    (src/ReadSet.h:171-209, src/Sequence.h:455-480): only the MINIMUM quality char of a read is tested
    (`max` is computed with min_element too); min < base or min > base+40 flips 33 <-> 64 for all reads."""
    base = input_base
    if base is None:
        base = 33
        for r in recs[:20000]:
            if not r["qual"]:
                continue
            m = min(r["qual"].encode("latin1"))
            if m < base or m > base + 40:
                base = 64 if base == 33 else 33
    d = start - base
    for r in recs:
        r["qual"] = bytes((c + d) & 0xFF for c in r["qual"].encode("latin1")).decode("latin1")
    return base


def longest_stretch(flags):
    """first-longest run of True; (offset, length)."""
    best = (0, 0)
    st = 0
    n = len(flags)
    for i, f in enumerate(flags):
        if not f:
            if i - st > best[1]:
                best = (st, i - st)
            st = i + 1
    if n - st > best[1]:
        best = (st, n - st)
    return best


def artifact_quality_trim(recs, start, min_quality, min_read_length):
    """Quality-stretch part of FilterKnownOddities (src/FilterKnownOddities.h:407-441,523-533,613-632):
    longest stretch with q >= START+minQuality; if it is not the whole read the read is replaced by the
    stretch (comment AFTrim:a+len) when passesLength, else flagged DISCARDED."""
    n_trim = n_disc = 0
    for r in recs:
        q = r["qual"].encode("latin1")
        a, n = longest_stretch([c >= start + min_quality for c in q])
        if (a, n) != (0, len(q)):
            if n <= 0 or not B.passes_length(n, len(r["seq"]), min_read_length):
                r["discarded"] = True
                n_disc += 1
            else:
                r["seq"] = r["seq"][a : a + n]
                r["qual"] = r["qual"][a : a + n]
                r["comment"] = (r["comment"] + " " if r["comment"] else "") + "AFTrim:%d+%d" % (a, n)
                n_trim += 1
    return n_trim, n_disc


def identify_pairs(recs):
    """Adjacent /1 /2 mates with equal common name (src/ReadSet.cpp:94-118,446-570, src/Utils.h:669-733).
    Returns list of (i, j|None)."""
    pairs = []
    i = 0
    n = len(recs)

    def split(nm):
        if len(nm) > 2 and nm[-2] == "/" and nm[-1] in "12ABFR":
            return nm[:-1], 1 if nm[-1] in "1AF" else 2
        return nm, 0

    while i < n:
        if i + 1 < n:
            c1, n1 = split(recs[i]["name"])
            c2, n2 = split(recs[i + 1]["name"])
            if c1 == c2 and n1 and n2 and n1 != n2:
                pairs.append((i, i + 1))
                i += 2
                continue
        pairs.append((i, None))
        i += 1
    return pairs


def to_buffers(recs):
    seqs = [r["seq"].encode("latin1") for r in recs]
    quals = [r["qual"].encode("latin1") for r in recs]
    bases, q, off = B.concat_reads(seqs, quals)
    disc = np.array([1 if r["discarded"] else 0 for r in recs], dtype=np.uint8)
    return bases, q, off, disc


def filter_reads(recs, k=31, start=33, min_quality=3, min_kmer_quality=0.10, min_depth=2, scoring="MAX",
                 min_read_length=0.40, min_passing_in_pair=1, skip_artifact_filter=False, threads=1):
    """apps/FilterReads.cpp:83-215 + apps/FilterReads.h:158-282 without normalisation.
    Returns (output FASTQ text, per-read results list, spectrum)."""
    if not skip_artifact_filter:
        artifact_quality_trim(recs, start, min_quality, min_read_length)
    bases, q, off, disc = to_buffers(recs)
    n_reads = len(recs)
    est = B.estimate_raw_kmers(n_reads, len(bases), k)
    spec = B.OracleSpectrum(k, min_quality, min_kmer_quality, threads=threads, est_distinct=max(1024, est // 4), start=start)
    spec.add_reads(bases, q, off, disc)
    spec.purge_min_depth(min_depth)
    sc = B.SCORING[scoring]
    toff, tlen, score, wast = spec.trim_reads(bases, off, min_depth, sc, disc, threads=threads)
    res = []
    for i, r in enumerate(recs):
        if r["discarded"]:
            res.append(dict(label="", passes=False, off=0, len=0, score=0.0))
            continue
        label = ""
        if wast[i]:
            label = "Trim:%d+%d" % (toff[i], tlen[i])
        label += (" " if label else "") + "%s:%d" % (B.SCORE_LABEL[sc], int(np.float32(score[i]) + np.float32(0.5)))
        passes = bool(score[i] >= min_depth) and B.passes_length(tlen[i], len(r["seq"]), min_read_length)
        res.append(dict(label=label, passes=passes, off=int(toff[i]), len=int(tlen[i]), score=float(score[i])))
    pairs = identify_pairs(recs)
    has_pairs = any(j is not None for _, j in pairs)
    keep = []
    if has_pairs:                                        # pickAllPassingPairs  src/ReadSelector.h:585-596
        for i, j in pairs:
            r1 = res[i]["passes"]
            r2 = res[j]["passes"] if j is not None else False
            ok = (r1 and r2) if (j is not None and min_passing_in_pair >= 2) else (r1 or r2)
            if ok:
                keep.append(i)
                if j is not None:
                    keep.append(j)
    else:                                                # pickAllPassingReads  src/ReadSelector.h:576-583
        keep = [i for i in range(n_reads) if res[i]["passes"]]
    keep.sort()
    out = []
    for i in keep:
        out.append(format_fastq(recs[i], res[i], start))
    return "".join(out), res, spec


def format_fastq(r, t, start):
    """Read::toFastq with trim (src/Sequence.cpp:296-328,729-770): discarded or trimLength<=1 prints N / START+1."""
    hdr = "@" + r["name"] + (" " + r["comment"] if r["comment"] else "") + (" " + t["label"] if t["label"] else "")
    if r["discarded"] or t["len"] <= 1:
        return hdr + "\nN\n+\n" + chr(start + 1) + "\n"
    s = r["seq"][t["off"] : t["off"] + t["len"]]
    qq = r["qual"][t["off"] : t["off"] + t["len"]]
    return hdr + "\n" + s + "\n+\n" + qq + "\n"


def meraculous_counts(recs, k=21, start=33, min_quality=2, min_kmer_quality=0.0, min_depth=2, threads=1):
    """apps/MeraculousCounter.cpp:110-151 + src/Meraculous.h:107-133: sorted `kmer\\tcount` lines for both
    strands and the extension graph lines.  Returns (mercount_lines, mergraph_lines)."""
    bases, q, off, disc = to_buffers(recs)
    spec = B.OracleSpectrum(k, min_quality, min_kmer_quality, track_ext=True, threads=threads, est_distinct=1 << 16, start=start)
    spec.add_reads(bases, q, off, None)
    e = spec.export()
    counts, graph = [], []
    idx = "ACGTNX"
    for key, c, ext in zip(e["keys"], e["count"], e["ext"]):
        if c < min_depth:
            continue
        km = "".join("ACGT"[(key[i >> 2] >> (6 - 2 * (i & 3))) & 3] for i in range(k))
        rc = revcomp(km)
        counts.append("%s\t%d" % (km, c))
        counts.append("%s\t%d" % (rc, c))
        re = [0] * 12
        for i, b in enumerate(idx):
            j = idx.index(COMP[b])
            re[j] = int(ext[6 + i])
            re[6 + j] = int(ext[i])
        graph.append(km + "\t" + " ".join(str(int(x)) for x in ext) + " 0")
        graph.append(rc + "\t" + " ".join(str(x) for x in re) + " 0")
    return sorted(set(counts)), sorted(set(graph))
