// FilterKnownOddities.h -- the default-on artifact pre-filter that runs immediately before the k-mer spectrum path
// (SURVEY.md section 8f row f2), with the reference's FilterKnownOddities call surface (src/FilterKnownOddities.h:170-741):
//   constructor: the screen's k-mer set -- every canonical 24-mer (--artifact-match-length) of the known artifact
//                sequences, each made circular by its first 24 bases, widened by --artifact-edit-distance substitutions:
//                built into the set while it is small (--build-artifact-edits-in-filter 2: below 750000 entries), the
//                remaining distance is searched at run time                                            :190-286
//   applyFilter(reads): per read (mates independently, :352-386)
//     1. quality: the best and second-best stretch of bases with quality >= START + minQuality          :407-441
//     2. screen: the canonical 24-mers at every 4th base of the kept stretch are looked up; the read keeps the
//        larger side of the hit region                                                                   :444-521
//     3. a read that changed is trimmed in place (comment "AFTrim:<off>+<len>") when what is left passes
//        --min-read-length, else it is DISCARDED; a second good quality stretch is rescued as an extra read
//        "<name>-qtrim" appended to the set                                                              :321-334,523-533,613-632,693-704
// Not implemented: --mask-simple-repeats, --phix-output, --artifact-reference-file, --filter-output (off by default;
// the option parser refuses them).
//
// Two details of the reference are kept because they decide which reads change:
//   * the scan pointer starts at the FIRST byte of the read although the loop variable starts at minPass/4, so with a
//     quality-trimmed head the 24-mer examined in iteration b is the one at byte b - minPass/4 while the hit is
//     recorded at base 4*b (:466-486);
//   * non-ACGT bases of artifacts and reads take part as 'A' (TwoBitSequence packs them so).
#ifndef KMERNATOR_HOST_FILTERKNOWNODDITIES_H
#define KMERNATOR_HOST_FILTERKNOWNODDITIES_H

#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "ReadSelector.h"
#include "ReadSet.h"

class FilterKnownOddities {
public:
    typedef uint64_t Key;                        // canonical k-mer, 2 bits per base, first base in the highest used bits

    struct FilterResults {
        unsigned int value;                      // 0: untouched; index of the artifact; numSequences(): quality trim only
        unsigned int minPass, maxPass;           // the range of the read to keep
        FilterResults() : value(0), minPass(0), maxPass(0) {}
    };

    FilterKnownOddities(int length = (int)kmn_host::asLong("artifact-match-length"), int numErrors = (int)kmn_host::asLong("artifact-edit-distance"))
        : _length((unsigned int)length), _numErrors(numErrors)
    {
        if (_length == 0 || _length > 28 || (_length & 3u)) LOG_THROW("Invalid: FilterKnownOddities must use a multiple of 4 bases, 28 or less");
        static const char *known[][2] = {
#include "../data/artifacts.inc"
        };
        _names.push_back("");                    // index 0 is the signal for no match
        _seqs.push_back("");
        for (size_t i = 0; i < sizeof(known) / sizeof(known[0]); ++i) { _names.push_back(known[i][0]); _seqs.push_back(known[i][1]); }
        for (size_t i = 0; i < _seqs.size(); ++i) _seqs[i] += _seqs[i].substr(0, _length);          // ReadSet::circularize
        prepareMaps();
    }

    unsigned int numSequences() const { return (unsigned int)_seqs.size(); }
    size_t filterSize() const { return _filter.size(); }
    int remainingEdits() const { return _numErrors; }

    // the screen of one read, without changing it
    FilterResults applyFilterToRead(const Read &read, float minimumReadLength, std::pair<long, long> *secondBestOut = NULL) const
    {
        FilterResults res;
        const std::string &seq = read.getFasta();
        const std::string &quals = read.getQuals();
        const unsigned int seqLen = (unsigned int)seq.size();
        const long bytes = ((long)seqLen + 3) / 4;
        unsigned int &minPass = res.minPass, &maxPass = res.maxPass, &value = res.value;
        // quality: best and second-best stretch (the reference's swap sequence, :416-441)
        std::pair<long, long> best(0, 0), secondBest(0, 0), test(0, 0);
        const bool hasQuals = !quals.empty() && (unsigned char)quals[0] != Read::REF_QUAL;
        const int minQual = Read::FASTQ_START_CHAR() + (int)Options::getOptions().getMinQuality();
        if (hasQuals) {
            for (unsigned int i = 0; i < seqLen; ++i) {
                test.second = i;
                if ((int)(unsigned char)quals[i] < minQual) {
                    if (test.second - test.first > best.second - best.first) std::swap(best, test);
                    if (test.second - test.first > secondBest.second - secondBest.first) std::swap(secondBest, test);
                    test.first = test.second = i + 1;
                }
            }
        }
        test.second = seqLen;
        if (test.second - test.first > best.second - best.first) std::swap(best, test);
        if (test.second - test.first > secondBest.second - secondBest.first) std::swap(secondBest, test);
        if (best.second > best.first) { minPass = (unsigned int)best.first; maxPass = (unsigned int)best.second; }
        else { minPass = 0; maxPass = 0; }
        if (secondBestOut) *secondBestOut = secondBest;

        // screen
        const long twoBitLength = _length / 4;
        long byteHops = (((long)maxPass + 3) / 4) - twoBitLength - ((seqLen & 3u) == 0 ? 0 : 1);
        if (byteHops < 0 || byteHops > bytes) byteHops = 0;
        unsigned int minAffected = maxPass, maxAffected = minPass;
        long ptr = 0;                                                     // byte of the read the examined k-mer starts at
        for (long byteHop = minPass / 4; byteHop <= byteHops; ++byteHop, ++ptr) {
            if ((unsigned long)ptr * 4 + _length > seqLen) continue;     // the reference would read past the packed bases here
            const Key least = canonical(pack(seq, (size_t)ptr * 4));
            unsigned int hit = lookup(least);
            if (!hit && _numErrors > 0) hit = lookupEdits(least, 0, _numErrors);
            if (hit) {
                const unsigned int pos = (unsigned int)byteHop * 4;
                value = hit;
                if (minAffected > pos) minAffected = pos;
                if (maxAffected < pos + _length) maxAffected = pos + _length;
            }
        }
        if (value > 0 && minAffected <= maxAffected) {                    // keep the larger side
            if ((long)minAffected - (long)minPass >= (long)maxPass - (long)maxAffected) maxPass = minAffected;
            else minPass = maxAffected;
        }
        if (value == 0 && (maxPass - minPass) != seqLen) value = numSequences();     // quality trim only
        (void)minimumReadLength;
        return res;
    }

    unsigned long applyFilter(ReadSet &reads)
    {
        const float minReadLength = ReadSelectorOptions::getOptions().getMinReadLength();
        unsigned long affected = 0, discarded = 0, trimmedBases = 0;
        std::vector<Read> remnants;
        const ReadSet::ReadSetSizeType n = reads.getSize();
        for (ReadSet::ReadSetSizeType i = 0; i < n; ++i) {
            Read &r = reads.getRead(i);
            std::pair<long, long> second(0, 0);
            const FilterResults res = applyFilterToRead(r, minReadLength, &second);
            if (res.value == 0) continue;
            const unsigned int len = r.getLength();
            if (res.value == numSequences() && ReadSelectorUtil::passesLength((float)(second.second - second.first), len, minReadLength)) {
                // only the quality changed the read: the second good stretch is rescued as a read of its own (:523-533)
                const long slen = second.second - second.first;
                std::ostringstream ss;
                ss << "AFTrim:" << second.first << "+" << slen;
                Read rem(r.getName() + "-qtrim", r.getComment(), r.getFasta().substr((size_t)second.first, (size_t)slen),
                         r.getQuals().substr((size_t)second.first, (size_t)slen));
                rem.addComment(ss.str(), "\t");
                rem.fileNum = r.fileNum;
                remnants.push_back(rem);
            }
            const long passLength = (long)res.maxPass - (long)res.minPass;
            if (passLength <= 0 || !ReadSelectorUtil::passesLength((float)passLength, len, minReadLength)) {
                r.discard();                                              // Recorder::recordDiscard :310-320
                discarded++;
                trimmedBases += len;
            } else {                                                      // Recorder::recordTrim :321-334
                std::ostringstream ss;
                ss << "AFTrim:" << res.minPass << "+" << passLength;
                r.seq = r.seq.substr(res.minPass, (size_t)passLength);
                r.quals = r.quals.substr(res.minPass, (size_t)passLength);
                r.addComment(ss.str(), "\t");
                affected++;
                trimmedBases += len - (unsigned long)passLength;
            }
        }
        for (size_t i = 0; i < remnants.size(); ++i) reads.append(remnants[i]);
        if (!remnants.empty()) LOG_VERBOSE(1, "Rescued " << remnants.size() << " reads from poor quality scores in the middle");
        LOG_VERBOSE(1, "Final Filter Matches to reads:" << affected << "\nDiscarded Reads:" << discarded << "\nTrimmed Reads:" << affected
                       << "\nDiscarded/Trimmed Bases:" << trimmedBases);
        reads.recount();
        return affected;                                                  // the sum of Recorder::readCounts (:716-717)
    }

    // ---- k-mer helpers (public for the tests) ----
    Key pack(const std::string &s, size_t off) const
    {
        Key k = 0;
        for (unsigned int i = 0; i < _length; ++i) k = (k << 2) | code(s[off + i]);
        return k;
    }
    Key canonical(Key fwd) const
    {
        Key rc = 0, f = fwd;
        for (unsigned int i = 0; i < _length; ++i) { rc = (rc << 2) | (3u - (f & 3u)); f >>= 2; }
        return fwd <= rc ? fwd : rc;
    }

private:
    static unsigned int code(char c)
    {
        switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; }
    }
    unsigned int lookup(Key k) const
    {
        std::unordered_map<Key, unsigned int>::const_iterator it = _filter.find(k);
        return it == _filter.end() ? 0u : it->second;
    }
    // any k-mer within `edits` substitutions at positions >= startIdx (Kmer::__permuteBases, src/Kmer.h:1409-1427), canonicalised
    unsigned int lookupEdits(Key k, unsigned int startIdx, int edits) const
    {
        if (edits <= 0) return 0;
        for (unsigned int b = startIdx; b < _length; ++b) {
            const unsigned int sh = 2 * (_length - 1 - b);
            const Key orig = (k >> sh) & 3u;
            for (Key v = 0; v < 4; ++v) {
                if (v == orig) continue;
                const Key m = (k & ~((Key)3 << sh)) | (v << sh);
                if (unsigned int h = lookup(canonical(m))) return h;
                if (edits > 1) if (unsigned int h = lookupEdits(m, b + 1, edits - 1)) return h;
            }
        }
        return 0;
    }
    void insertIfNew(Key k, unsigned int v) { _filter.insert(std::make_pair(k, v)); }      // getOrSetElement: the first value stays
    void prepareMaps()
    {
        for (unsigned int i = 0; i < _seqs.size(); ++i) {
            const std::string &s = _seqs[i];
            if (s.size() < _length) continue;
            for (size_t j = 0; j + _length <= s.size(); ++j) insertIfNew(canonical(pack(s, j)), i);
        }
        const int maxErrors = _numErrors;
        const long buildEdits = kmn_host::asLong("build-artifact-edits-in-filter");
        for (int error = 0; error < maxErrors; ++error) {
            if (buildEdits == 1 || (buildEdits == 2 && _filter.size() < 750000)) {
                _numErrors--;
                std::vector<std::pair<Key, unsigned int> > snapshot(_filter.begin(), _filter.end());
                for (size_t q = 0; q < snapshot.size(); ++q) {
                    const Key k = snapshot[q].first;
                    for (unsigned int b = 0; b < _length; ++b) {
                        const unsigned int sh = 2 * (_length - 1 - b);
                        const Key orig = (k >> sh) & 3u;
                        for (Key v = 0; v < 4; ++v)
                            if (v != orig) insertIfNew(canonical((k & ~((Key)3 << sh)) | (v << sh)), snapshot[q].second);
                    }
                }
            }
        }
        LOG_DEBUG(2, "filter is " << _filter.size() << ".  Remaining edits is:" << _numErrors);
    }

    unsigned int _length;
    int _numErrors;
    std::vector<std::string> _names, _seqs;
    std::unordered_map<Key, unsigned int> _filter;
};

#endif
