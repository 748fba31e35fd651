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


# ---------------------------------------------------------------------------------------------------------
# a12: RANDOM coverage normalisation (src/ReadSelector.h:661-749).  IntRand is boost::random::mt19937 behind a
# uniform_int_distribution<uint32_t> over its full range (src/Utils.h:1147), i.e. the raw 32-bit outputs of MT19937; the
# reference seeds it from time(NULL), so parity = identical decisions for an injected draw sequence.
# ---------------------------------------------------------------------------------------------------------
class MT19937:
    """Matsumoto & Nishimura's reference generator (init_genrand / genrand_int32), what std::mt19937(seed) and
    boost::random::mt19937(seed) implement; mt19937()() with the default seed 5489 gives 4123659995 as its 10000th output."""

    def __init__(self, seed=5489):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def choose_read(score, target_depth, use_logscale, draw):
    """ReadSelector::chooseRead (src/ReadSelector.h:661-672).  `draw` is called only when score > target_depth."""
    if score <= target_depth:
        return True
    choice = draw() % score
    if use_logscale:
        return choice <= target_depth * np.log(np.float32(score) / np.float32(target_depth))
    return choice <= target_depth


def pick_coverage_normalized_subset(scores, passes, pairs, target_depth, by_pair, both_pass, draw, use_logscale=False):
    """ReadSelector::pickCoverageNormalizedSubset (src/ReadSelector.h:673-749), single thread (pairs in order, one RNG
    stream).  scores[i] = ReadTrimType::score of read i, passes[i] = isPassingRead(i, minimumScore, minimumLength),
    pairs = [(i, j|None)].  Returns the picked read indexes in pick order after optimizePickOrder (ascending)."""
    picked = []
    for i, j in pairs:
        s1 = int(scores[i]) if passes[i] else -1                       # (long) of a float: truncation
        s2 = int(scores[j]) if (j is not None and passes[j]) else -1
        if by_pair:
            p1, p2 = passes[i], (passes[j] if j is not None else False)
            if not ((p1 and p2) if (j is not None and both_pass) else (p1 or p2)):        # isPassingPair :558-568
                continue
            if both_pass and (s1 <= 0 or s2 <= 0):
                continue
            if s1 <= 0 and s2 <= 0:
                continue
            if choose_read(max(s1, s2), target_depth, use_logscale, draw):
                picked.append(i)
                if j is not None:
                    picked.append(j)
        else:
            if s1 > 0 and choose_read(s1, target_depth, use_logscale, draw):
                picked.append(i)
            if s2 > 0 and choose_read(s2, target_depth, use_logscale, draw):
                picked.append(j)
    return sorted(picked)


def expected_kept_fraction(scores, target_depth):
    """sum over reads of P(kept) = min(1, (D+1)/s): rand() % s <= D holds for D+1 of the s residues."""
    s = np.asarray(scores, dtype=np.float64)
    p = np.where(s > target_depth, (target_depth + 1.0) / np.maximum(s, 1.0), 1.0)
    return float(p.sum()), float((p * (1.0 - p)).sum())


# ---------------------------------------------------------------------------------------------------------
# a9: text of KmerSpectrum::Histogram::toString (src/KmerSpectrum.h:987-1035) from the three bucket columns
# ---------------------------------------------------------------------------------------------------------
def format_histogram(visits, visited_count, visited_weight, zoom_max=256):
    """visits/visited_count/visited_weight: bucket arrays as filled by Histogram::addRecord (:964-975); the text is what
    finish() + toString() print: fixed, setprecision(3); integer members print as integers, doubles with 3 decimals."""
    zoom_log_skip = 7                                                  # (uint)(log(zoomMax+1)/log 2 - 1) for 255 and 256
    n = len(visits)
    cum = [0] * n
    count, total_count, total_weight, last = 0, 0.0, 0.0, 0
    for i in range(n - 1, -1, -1):
        count += int(visits[i])
        cum[i] = count
        if visits[i] > 0:
            total_count += float(visited_count[i])
            total_weight += float(visited_weight[i])
            last = max(last, i)

    def f(x):
        return "%.3f" % x

    def div(a, b):
        return a / b if b else float("nan")

    out = ["Counts, Weights and Directions\n"]
    out.append("Counts:\t%d\t%s\t%s\t\n" % (count, f(total_count), f(div(total_count, count))))
    out.append("Weights:\t%d\t%s\t%s\t%s\n" % (count, f(total_weight), f(div(total_weight, count)), f(div(total_weight, total_count))))
    out.append("\n")
    out.append("Bucket\tCumulative\tUnique\t%Unique\tCount\t%Count\tWeight\tQualProb\t%Weight\n")
    for i in range(1, last + 1):
        label = i if i <= zoom_max else int(2.0 ** (i + zoom_log_skip - zoom_max))
        v, c, w = int(visits[i]), int(visited_count[i]), float(visited_weight[i])
        out.append("%d\t%d\t%d\t%s\t%d\t%s\t\t%s\t%s\t%s\t\n" % (
            label, cum[i], v, f(div(100.0 * v, count)), c, f(div(100.0 * c, total_count)), f(w), f(div(w, c)), f(div(100.0 * w, total_weight))))
    return "".join(out)


# ---------------------------------------------------------------------------------------------------------
# f2: FilterKnownOddities with the 24-mer artifact screen (src/FilterKnownOddities.h:190-286,389-541,693-704).
# numpy restatement on uint64-packed k-mers, independent of the host C++ (kmernator_b200/host/kmernator/FilterKnownOddities.h);
# only the list of artifact sequences is shared (host/data/artifacts.inc = the reference's getArtifactFasta, :742-795).
# ---------------------------------------------------------------------------------------------------------
def artifact_sequences():
    import os
    import re
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmernator_b200", "host", "data", "artifacts.inc")
    return re.findall(r'\{"([^"]*)", "([^"]*)"\}', open(inc).read())


_CODE = np.zeros(256, dtype=np.uint64)
for _c, _v in ((b"C", 1), (b"G", 2), (b"T", 3), (b"c", 1), (b"g", 2), (b"t", 3)):
    _CODE[_c[0]] = _v                                                   # everything else packs as A (TwoBitSequence)


def _pack_all(seq, L):
    """uint64 array of every L-mer of seq (first base in the highest used bits)"""
    c = _CODE[np.frombuffer(seq.encode("latin1"), dtype=np.uint8)]
    n = len(c) - L + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    k = np.zeros(n, dtype=np.uint64)
    for i in range(L):
        k = (k << np.uint64(2)) | c[i:i + n]
    return k


def _canonical(k, L):
    rc = np.zeros_like(k)
    f = k.copy()
    for _ in range(L):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (f & np.uint64(3)))
        f >>= np.uint64(2)
    return np.minimum(k, rc)


def _substitutions(k, L):
    """all single-base substitutions of every k-mer in k (3 * L each)"""
    out = []
    for b in range(L):
        sh = np.uint64(2 * (L - 1 - b))
        orig = (k >> sh) & np.uint64(3)
        cleared = k & ~(np.uint64(3) << sh)
        for d in (1, 2, 3):
            out.append(cleared | (((orig + np.uint64(d)) & np.uint64(3)) << sh))
    return np.concatenate(out) if out else k[:0]


class ArtifactFilter:
    def __init__(self, match_length=24, edit_distance=2, build_edits=2):
        self.L = match_length
        seqs = [""] + [s for _, s in artifact_sequences()]
        seqs = [s + s[:match_length] for s in seqs]                      # ReadSet::circularize (src/ReadSet.cpp:120-130)
        self.n_sequences = len(seqs)
        ks = [_canonical(_pack_all(s, match_length), match_length) for s in seqs]
        filt = np.unique(np.concatenate(ks))
        self.num_errors = edit_distance
        for _ in range(edit_distance):                                   # prepareMaps :254-283
            if build_edits == 1 or (build_edits == 2 and len(filt) < 750000):
                self.num_errors -= 1
                filt = np.unique(np.concatenate([filt, _canonical(_substitutions(filt, match_length), match_length)]))
        self.filter = filt

    def _hit(self, k):
        """k: uint64 array of canonical k-mers -> bool array: within the remaining run-time edit distance of the filter"""
        def member(x):
            i = np.searchsorted(self.filter, x)
            i[i >= len(self.filter)] = 0
            return self.filter[i] == x
        hit = member(k)
        cur = [k]
        for _ in range(self.num_errors):                                 # superset of the reference's increasing-position walk: same reachable set
            nxt = _substitutions(np.concatenate(cur), self.L)
            m = member(_canonical(nxt, self.L)).reshape(3 * self.L, -1) if len(nxt) else np.zeros((0, 0), bool)
            per = len(np.concatenate(cur))
            if per:
                mm = m.reshape(3 * self.L, per).any(axis=0)
                # fold the hits of the variants back onto the originating k-mers
                reps = per // len(k)
                hit |= mm.reshape(reps, len(k)).any(axis=0)
            cur = [nxt]
        return hit

    def screen(self, seq, qual, start, min_quality):
        """applyFilterToRead (:389-541) -> (value, minPass, maxPass, secondBest)"""
        L = self.L
        n = len(seq)
        best, second, test = [0, 0], [0, 0], [0, 0]
        if qual is not None:
            for i, c in enumerate(qual):
                test[1] = i
                if c < start + min_quality:
                    if test[1] - test[0] > best[1] - best[0]:
                        best, test = test, best
                    if test[1] - test[0] > second[1] - second[0]:
                        second, test = test, second
                    test = [i + 1, i + 1]
        test[1] = n
        if test[1] - test[0] > best[1] - best[0]:
            best, test = test, best
        if test[1] - test[0] > second[1] - second[0]:
            second, test = test, second
        min_pass, max_pass = (best[0], best[1]) if best[1] > best[0] else (0, 0)
        byte_hops = (max_pass + 3) // 4 - L // 4 - (0 if n % 4 == 0 else 1)
        if byte_hops < 0 or byte_hops > (n + 3) // 4:
            byte_hops = 0
        value = 0
        min_aff, max_aff = max_pass, min_pass
        hops = list(range(min_pass // 4, byte_hops + 1))
        if hops:
            ptrs = np.arange(len(hops)) * 4                              # the scan pointer starts at byte 0 of the read (:466-486)
            ok = ptrs + L <= n
            allk = _pack_all(seq, L)
            if ok.any():
                ks = _canonical(allk[ptrs[ok]], L)
                h = self._hit(ks)
                for hop, hh in zip(np.array(hops)[ok], h):
                    if hh:
                        pos = int(hop) * 4
                        value = 1
                        min_aff = min(min_aff, pos)
                        max_aff = max(max_aff, pos + L)
        if value > 0 and min_aff <= max_aff:
            if (min_aff - min_pass) >= (max_pass - max_aff):
                max_pass = min_aff
            else:
                min_pass = max_aff
        if value == 0 and (max_pass - min_pass) != n:
            value = self.n_sequences
        return value, min_pass, max_pass, tuple(second)


def artifact_filter(recs, start, min_quality, min_read_length, match_length=24, edit_distance=2, build_edits=2, flt=None):
    """FilterKnownOddities::applyFilter (:663-732): trims / discards in place, appends the rescued "-qtrim" remnants.
    Returns (trimmed, discarded, remnants)."""
    flt = flt or ArtifactFilter(match_length, edit_distance, build_edits)
    n_trim = n_disc = 0
    remnants = []
    for r in recs:
        has_q = bool(r["qual"]) and ord(r["qual"][0]) != 0xFF
        q = r["qual"].encode("latin1") if has_q else None
        value, a, b, second = flt.screen(r["seq"], q, start, min_quality)
        if value == 0:
            continue
        L = len(r["seq"])
        if value == flt.n_sequences and B.passes_length(second[1] - second[0], L, min_read_length):
            sl = second[1] - second[0]
            lab = "AFTrim:%d+%d" % (second[0], sl)
            remnants.append(dict(name=r["name"] + "-qtrim", comment=(r["comment"] + "\t" + lab) if r["comment"] else lab,
                                 seq=r["seq"][second[0]:second[1]], qual=r["qual"][second[0]:second[1]], discarded=False))
        n = b - a
        if n <= 0 or not B.passes_length(n, L, min_read_length):
            r["discarded"] = True
            n_disc += 1
        else:
            lab = "AFTrim:%d+%d" % (a, n)
            r["seq"], r["qual"] = r["seq"][a:b], r["qual"][a:b]
            r["comment"] = (r["comment"] + "\t" + lab) if r["comment"] else lab
            n_trim += 1
    recs.extend(remnants)
    return n_trim, n_disc, len(remnants)
