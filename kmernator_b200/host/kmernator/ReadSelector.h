// ReadSelector.h -- the reference's ReadSelector call surface on top of the C ABI.
//   RS selector(reads, spectrum.weak)                src/ReadSelector.h:376-420
//   scoreAndTrimReads(minDepth)                      src/ReadSelector.h:1182-1209  -> kmn_trim_batch (GPU)
//   setTrimHeaders labels                            src/ReadSelector.h:1015-1035  (host: string formatting)
//   passesLength / isPassingRead / isPassingPair     src/ReadSelector.h:209-228,550-568
//   pickAllPassingReads / pickAllPassingPairs        src/ReadSelector.h:576-596
//   pickCoverageNormalizedSubset / chooseRead        src/ReadSelector.h:661-749   (RANDOM normalisation, injectable RNG)
//   getOFM / writePicks / optimizePickOrder          src/ReadSelector.h:1212-1262, src/Utils.h:248-464
#ifndef KMERNATOR_HOST_READSELECTOR_H
#define KMERNATOR_HOST_READSELECTOR_H

#include <algorithm>
#include <cmath>
#include <ctime>
#include <fstream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "KmerSpectrum.h"

// key -> output stream, files named <prefix><key><suffix>                 src/Utils.h:248-464 (OfstreamMap)
class OfstreamMap {
public:
    OfstreamMap(const std::string &prefix = "", const std::string &suffix = "") : _prefix(prefix), _suffix(suffix) {}
    std::ostream &getOfstream(const std::string &key)
    {
        std::map<std::string, std::shared_ptr<std::ofstream> >::iterator it = _map.find(key);
        if (it == _map.end()) {
            std::string fn = _prefix + key + _suffix;
            std::shared_ptr<std::ofstream> of(new std::ofstream(fn.c_str()));
            if (!of->good()) LOG_THROW("Could not open " << fn << " for writing");
            LOG_VERBOSE(1, "Writing to " << fn);
            created().push_back(fn);
            it = _map.insert(std::make_pair(key, of)).first;
        }
        return *it->second;
    }
    void clear() { _map.clear(); }
    static std::vector<std::string> &created() { static std::vector<std::string> v; return v; }   // every file opened so far
private:
    std::string _prefix, _suffix;
    std::map<std::string, std::shared_ptr<std::ofstream> > _map;
};

class ReadSelectorUtil {
public:
    // fp32 arithmetic on purpose: readLength * minimumLength is a float product in the reference   src/ReadSelector.h:209-228
    static inline bool passesLength(float length, unsigned int readLength, float minimumLength)
    {
        if (length <= 1.0f) return false;
        if (minimumLength <= 1.0f) return (float)readLength * minimumLength <= length;
        return minimumLength <= length;
    }
};

class ReadSelector {
public:
    typedef OfstreamMap OFM;
    typedef ReadSet::ReadSetSizeType ReadSetSizeType;
    typedef ReadSet::Pair Pair;
    typedef float ScoreType;
    enum KmerScoringType { KS_SUM = 0, KS_MEDIAN = 1, KS_MIN = 2, KS_MAX = 3, KS_AVG = 4 };   // same order as kmn_scoring

    struct ReadTrimType {                                                 // src/ReadSelector.h:259-274
        unsigned int trimOffset, trimLength;
        ScoreType score;
        std::string label;
        bool isAvailable, wasTrimmed;
        ReadTrimType() : trimOffset(0), trimLength(0), score(0), isAvailable(true), wasTrimmed(false) {}
    };

    ReadSelector(const ReadSet &reads, const KmerMapHandle &map) : _reads(reads), _map(map), _lastSortedPick(0)
    {
        // IntRand is mt19937 seeded from time(NULL) in the reference (src/Utils.h:1093-1106,1147): not reproducible
        // run to run.  KMN_SEED fixes the stream for tests.
        const char *s = getenv("KMN_SEED");
        _rng.seed(s ? (unsigned long)strtoul(s, NULL, 10) : (unsigned long)time(NULL) ^ 1ul);
    }
    void setRandomSource(const std::mt19937 &rng) { _rng = rng; }
    const std::vector<ReadTrimType> &getTrims() const { return _trims; }
    const std::vector<Pair> &getPicks() const { return _picks; }

    static KmerScoringType getScoringType()
    {
        std::string s = ReadSelectorOptions::getOptions().getKmerScoringType();
        for (size_t i = 0; i < s.size(); ++i) s[i] = (char)toupper(s[i]);
        if (s == "SUM") return KS_SUM;
        if (s == "MEDIAN") return KS_MEDIAN;
        if (s == "MIN") return KS_MIN;
        if (s == "MAX") return KS_MAX;
        if (s == "AVG") return KS_AVG;
        LOG_THROW("Invalid scoring type!");
    }

    // lookup pass on the GPU; labels here                               src/ReadSelector.h:1182-1209,1015-1035
    void scoreAndTrimReads(ScoreType minimumKmerScore)
    {
        const ReadSetSizeType n = _reads.getSize();
        _trims.assign(n, ReadTrimType());
        const unsigned int k = KmerBaseOptions::getOptions().getKmerSize();
        const KmerScoringType scoring = getScoringType();
        static const char *names[] = {"Sum", "Median", "Min", "Max", "Avg"};
        if (k == 0 || !_map.ctx) {                                        // no k-mer work: whole (artifact-filtered) reads
            for (ReadSetSizeType i = 0; i < n; ++i) {
                _trims[i].trimLength = _reads.getRead(i).isDiscarded() ? 0 : _reads.getRead(i).getLength();
                _trims[i].score = _trims[i].trimLength;
            }
            return;
        }
        ReadSetSizeType batch = Options::getOptions().getBatchSize();
        if (batch == 0) batch = 100000;
        std::string bases, quals;
        std::vector<uint64_t> off;
        std::vector<uint8_t> disc, wasTrimmed;
        std::vector<uint32_t> toff, tlen;
        std::vector<float> score;
        // distributed: the lookup is collective (keys travel to their owners), so every rank makes the same number of
        // calls, with empty batches once it has run out of reads   DistributedReadSelector, src/DistributedFunctions.h:903-1045
        unsigned long nBatches = (n + batch - 1) / batch;
        if (World::instance()) nBatches = World::instance()->allMax(nBatches);
        for (unsigned long bi = 0; bi < nBatches; ++bi) {
            const ReadSetSizeType r0 = std::min<ReadSetSizeType>(n, bi * batch), r1 = std::min<ReadSetSizeType>(n, r0 + batch), m = r1 - r0;
            _reads.concat(r0, r1, bases, quals, off, disc);
            toff.resize(m + 1); tlen.resize(m + 1); score.resize(m + 1); wasTrimmed.resize(m + 1);
            KMN_CHECK(_map.ctx, kmn_trim_batch(_map.ctx, (const uint8_t *)bases.data(), off.data(), m, disc.data(), (uint32_t)minimumKmerScore,
                                               (int)scoring, toff.data(), tlen.data(), score.data(), wasTrimmed.data()));
            for (ReadSetSizeType i = 0; i < m; ++i) {
                ReadTrimType &t = _trims[r0 + i];
                if (disc[i]) continue;                                    // never scored (src/ReadSelector.h:1196-1198)
                t.trimOffset = toff[i]; t.trimLength = tlen[i]; t.score = score[i]; t.wasTrimmed = wasTrimmed[i] != 0;
                std::ostringstream ss;
                if (t.wasTrimmed) ss << "Trim:" << t.trimOffset << "+" << t.trimLength << " ";
                ss << names[scoring] << "Score:" << (int)(t.score + 0.5f);
                t.label = ss.str();
            }
        }
    }

    bool isPassingRead(ReadSetSizeType readIdx) const { return _reads.isValidRead(readIdx); }
    bool isPassingRead(ReadSetSizeType readIdx, ScoreType minimumScore, float minimumLength) const
    {
        if (!isPassingRead(readIdx)) return false;
        const ReadTrimType &t = _trims[readIdx];
        return t.isAvailable && t.score >= minimumScore && ReadSelectorUtil::passesLength((float)t.trimLength, _reads.getRead(readIdx).getLength(), minimumLength);
    }
    bool isPassingPair(const Pair &pair, ScoreType minimumScore, float minimumLength, bool bothPass) const
    {
        bool r1 = isPassingRead(pair.read1, minimumScore, minimumLength), r2 = isPassingRead(pair.read2, minimumScore, minimumLength);
        if (!pair.isSingle() && bothPass) return r1 & r2;
        return r1 | r2;
    }
    bool pickIfNew(ReadSetSizeType readIdx)
    {
        if (!_reads.isValidRead(readIdx) || !_trims[readIdx].isAvailable) return false;
        _trims[readIdx].isAvailable = false;
        _picks.push_back(Pair(readIdx));
        return true;
    }
    int pickAllPassingReads(ScoreType minimumScore = 0.0, float minimumLength = ReadSelectorOptions::getOptions().getMinReadLength())
    {
        int picked = 0;
        for (ReadSetSizeType i = 0; i < _reads.getSize(); i++)
            if (isPassingRead(i, minimumScore, minimumLength) && pickIfNew(i)) picked++;
        optimizePickOrder();
        return picked;
    }
    ReadSetSizeType pickAllPassingPairs(ScoreType minimumScore = 0.0, float minimumLength = ReadSelectorOptions::getOptions().getMinReadLength(), bool bothPass = false)
    {
        ReadSetSizeType picked = 0;
        for (ReadSetSizeType i = 0; i < _reads.getPairSize(); i++) {
            const Pair &pair = _reads.getPair(i);
            if (isPassingPair(pair, minimumScore, minimumLength, bothPass)) {
                if (pickIfNew(pair.read1)) picked++;
                if (pickIfNew(pair.read2)) picked++;
            }
        }
        optimizePickOrder();
        return picked;
    }

    // RANDOM normalisation: a read (or pair, by its larger score) with score s > targetDepth is kept when
    // rand() % s <= targetDepth (or <= targetDepth * log(s / targetDepth))     src/ReadSelector.h:661-749
    bool chooseRead(long score, long targetDepth, bool useLogscale)
    {
        if (score <= targetDepth) return true;
        long choice = (long)(_dist(_rng) % (unsigned long)score);
        if (useLogscale) return choice <= targetDepth * std::log((float)score / (float)targetDepth);
        return choice <= targetDepth;
    }
    ReadSetSizeType pickCoverageNormalizedSubset(long targetDepth, ScoreType minimumScore = 0.0, float minimumLength = ReadSelectorOptions::getOptions().getMinReadLength(),
                                                 bool byPair = false, bool bothPass = false)
    {
        ReadSetSizeType picked = 0;
        const bool useLogscale = ReadSelectorOptions::getOptions().getUseLogscaleAboveMax();
        for (ReadSetSizeType pairIdx = 0; pairIdx < _reads.getPairSize(); pairIdx++) {
            const Pair &pair = _reads.getPair(pairIdx);
            long score1 = (long)(isPassingRead(pair.read1, minimumScore, minimumLength) ? _trims[pair.read1].score : -1);
            long score2 = (long)(isPassingRead(pair.read2, minimumScore, minimumLength) ? _trims[pair.read2].score : -1);
            if (byPair) {
                if (!isPassingPair(pair, minimumScore, minimumLength, bothPass)) continue;
                if (bothPass && (score1 <= 0 || score2 <= 0)) continue;
                if (score1 <= 0 && score2 <= 0) continue;
                if (chooseRead(std::max(score1, score2), targetDepth, useLogscale)) {
                    if (pickIfNew(pair.read1)) picked++;
                    if (pickIfNew(pair.read2)) picked++;
                }
            } else {
                if (score1 > 0 && chooseRead(score1, targetDepth, useLogscale) && pickIfNew(pair.read1)) picked++;
                if (score2 > 0 && chooseRead(score2, targetDepth, useLogscale) && pickIfNew(pair.read2)) picked++;
            }
        }
        optimizePickOrder();
        return picked;
    }

    void optimizePickOrder()
    {
        if (_lastSortedPick >= _picks.size()) return;
        std::sort(_picks.begin() + _lastSortedPick, _picks.end());
        _lastSortedPick = _picks.size();
    }
    OFM getOFM(const std::string &outputFile, const std::string &suffix = "") { return OFM(outputFile, suffix); }
    void writePicks(OFM &ofm, ReadSetSizeType offset = 0, bool byInputFile = ReadSelectorOptions::getOptions().getSeparateOutputs())
    {
        const int fmt = Options::getOptions().getFormatOutput();
        // records are formatted into one buffer per output stream and written in blocks of ~1 MB; the key of a read's stream
        // is looked up once per input file, not once per read
        std::map<unsigned int, std::pair<std::ostream *, std::string> > out;        // input file number (0: one output) -> stream, pending bytes
        for (ReadSetSizeType p = offset; p < _picks.size(); ++p) {
            const ReadSetSizeType ids[2] = {_picks[p].read1, _picks[p].read2};
            for (int q = 0; q < 2; ++q) {
                if (ids[q] == ReadSet::MAX_READ_IDX) continue;
                const ReadTrimType &t = _trims[ids[q]];
                const unsigned int fileKey = byInputFile ? _reads.getReadFileNum(ids[q]) : 0u;
                std::map<unsigned int, std::pair<std::ostream *, std::string> >::iterator it = out.find(fileKey);
                if (it == out.end()) {
                    std::string key;
                    if (byInputFile) key = "-" + _reads.getReadFileNamePrefix(ids[q]);
                    it = out.insert(std::make_pair(fileKey, std::make_pair(&ofm.getOfstream(key), std::string()))).first;
                    it->second.second.reserve((1u << 20) + 4096);
                }
                const Read &r = _reads.getRead(ids[q]);
                std::string &buf = it->second.second;
                if (fmt & 1) buf += r.toFasta(t.trimOffset, t.trimLength, t.label);
                else r.appendFastq(buf, t.trimOffset, t.trimLength, t.label);
                if (buf.size() >= (1u << 20)) { it->second.first->write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
            }
        }
        for (std::map<unsigned int, std::pair<std::ostream *, std::string> >::iterator it = out.begin(); it != out.end(); ++it)
            if (!it->second.second.empty()) it->second.first->write(it->second.second.data(), (std::streamsize)it->second.second.size());
    }

private:
    const ReadSet &_reads;
    KmerMapHandle _map;
    std::vector<ReadTrimType> _trims;
    std::vector<Pair> _picks;
    size_t _lastSortedPick;
    std::mt19937 _rng;
    std::uniform_int_distribution<uint32_t> _dist;
};

#endif
