// ReadSet.h -- host-side input container with the reference's ReadSet / Read call surface for the hot-path drivers
// (SURVEY.md section 8b, next-row f1):
//   ReadSet::appendAllFiles(files)            src/ReadSet.cpp:186-258
//   ReadSet::identifyPairs()                  src/ReadSet.cpp:446-570, src/Utils.h:669-733
//   getSize/getBaseCount/getRead/getPair/getPairSize/hasPairs/getMaxSequenceLength/getReadFileNum
//                                              src/ReadSet.h:359-537
//   Read::getName/getComment/getLength/getFasta/getQuals/isDiscarded/discard
//                                              src/Sequence.h:243-498
// Storage is a plain SoA-friendly vector of records (the reference's compact per-read blob, src/Sequence.h:156-171,
// is out of scope: the GPU consumes concatenated ASCII batches, see concat()).
// FASTQ qualities are re-expressed in Read::FASTQ_START_CHAR (= --fastq-output-base-quality) on load, with the input
// base auto-detected from the minimum quality character of each read (src/ReadSet.h:171-209,694-708,
// src/Sequence.h:455-480).
#ifndef KMERNATOR_HOST_READSET_H
#define KMERNATOR_HOST_READSET_H

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "Log.h"
#include "Options.h"

class Read {
public:
    static int &FASTQ_START_CHAR() { static int c = 33; return c; }        // Read::FASTQ_START_CHAR, src/Sequence.cpp:543-547
    static const unsigned char REF_QUAL = 0xff;                              // Kmernator::REF_QUAL: FASTA reads carry no qualities
    static const unsigned char PRINT_REF_QUAL = 33 + 70;                     // Kmernator::PRINT_REF_QUAL = 103 ('g'), src/config.h:140

    Read() : discarded(false), fileNum(0) {}
    Read(const std::string &n, const std::string &c, const std::string &s, const std::string &q)
        : name(n), comment(c), seq(s), quals(q), discarded(false), fileNum(0) {}

    const std::string &getName() const { return name; }
    const std::string &getComment() const { return comment; }
    unsigned int getLength() const { return (unsigned int)seq.size(); }
    const std::string &getFasta() const { return seq; }
    const std::string &getQuals() const { return quals; }
    bool isDiscarded() const { return discarded; }
    void discard() { discarded = true; }
    // Read::getTrimRead joins an existing comment and the label with a tab (src/Sequence.h:485-496)
    void addComment(const std::string &c, const char *sep = " ") { comment = comment.empty() ? c : comment + sep + c; }

    // Read::toFastq / Sequence::_getFastaString with a trim (src/Sequence.cpp:296-328,729-770): a discarded read, or one
    // trimmed to <= 1 base, is written as "N" with quality START+1
    std::string toFastq(unsigned int trimOffset, unsigned int trimLength, const std::string &label) const
    {
        std::string hdr = "@" + name;
        if (Options::getOptions().getKeepReadComment() && !comment.empty()) hdr += " " + comment;
        if (!label.empty()) hdr += " " + label;
        if (discarded || trimLength <= 1) return hdr + "\nN\n+\n" + std::string(1, (char)(FASTQ_START_CHAR() + 1)) + "\n";
        std::string q = quals.substr(trimOffset, trimLength);
        if (!q.empty() && (unsigned char)q[0] == REF_QUAL) q.assign(trimLength, (char)PRINT_REF_QUAL);   // quality-less reads, src/Sequence.cpp:744-747
        return hdr + "\n" + seq.substr(trimOffset, trimLength) + "\n+\n" + q + "\n";
    }
    std::string toFasta(unsigned int trimOffset, unsigned int trimLength, const std::string &label) const
    {
        std::string hdr = ">" + name;
        if (Options::getOptions().getKeepReadComment() && !comment.empty()) hdr += " " + comment;
        if (!label.empty()) hdr += " " + label;
        if (discarded || trimLength <= 1) return hdr + "\nN\n";
        return hdr + "\n" + seq.substr(trimOffset, trimLength) + "\n";
    }

    std::string name, comment, seq, quals;
    bool discarded;
    unsigned int fileNum;       // 1-based input file number
};

class ReadSet {
public:
    typedef unsigned long ReadSetSizeType;
    static const ReadSetSizeType MAX_READ_IDX = (ReadSetSizeType)-1;
    struct Pair {
        ReadSetSizeType read1, read2;
        Pair() : read1(MAX_READ_IDX), read2(MAX_READ_IDX) {}
        Pair(ReadSetSizeType a, ReadSetSizeType b = MAX_READ_IDX) : read1(a), read2(b) {}
        bool isSingle() const { return read1 == MAX_READ_IDX || read2 == MAX_READ_IDX; }
        bool hasAValidRead() const { return read1 != MAX_READ_IDX || read2 != MAX_READ_IDX; }
        ReadSetSizeType lesser() const { return read1 < read2 ? read1 : read2; }
        bool operator<(const Pair &o) const { return lesser() < o.lesser(); }
    };

    ReadSet() : _baseCount(0), _maxLength(0), _inputBase(0), _deferNormalise(false) {}

    // each rank parses byte range [rank, rank+1)/size of every file in the reference (src/ReadFileReader.h:379-398);
    // here one process feeds one GPU and slices by record count instead
    void appendAllFiles(const OptionsBaseInterface::FileListType &files, int rank = 0, int size = 1)
    {
        unsigned int fileNum = 0;
        for (OptionsBaseInterface::FileListType::const_iterator it = files.begin(); it != files.end(); ++it) appendAnyFile(*it, ++fileNum, rank, size);
        if (!_deferNormalise) normaliseQualities();
    }
    void deferNormalise(bool d = true) { _deferNormalise = d; }       // the distributed driver agrees on the input base first
    void appendAnyFile(const std::string &path, unsigned int fileNum = 1, int rank = 0, int size = 1)
    {
        std::ifstream in(path.c_str());
        if (!in.good()) LOG_THROW("Could not open : " << path);
        std::vector<Read> tmp;
        std::string l1, l2, l3, l4;
        int first = in.peek();
        if (first == '@') {
            while (std::getline(in, l1)) {
                if (l1.empty()) continue;
                if (l1[0] != '@') LOG_THROW("Missing '@' in header of " << path << ": " << l1);
                if (!std::getline(in, l2) || !std::getline(in, l3) || !std::getline(in, l4)) LOG_THROW("Truncated FASTQ record in " << path << ": " << l1);
                stripCR(l1); stripCR(l2); stripCR(l4);
                if (l2.size() != l4.size()) LOG_THROW("Number of bases and quals do not match in " << path << ": " << l1);
                tmp.push_back(makeRead(l1.substr(1), l2, l4, fileNum));
            }
        } else if (first == '>') {
            std::string hdr, seq;
            while (std::getline(in, l1)) {
                stripCR(l1);
                if (!l1.empty() && l1[0] == '>') {
                    if (!hdr.empty()) tmp.push_back(makeRead(hdr, seq, std::string(seq.size(), (char)Read::REF_QUAL), fileNum));
                    hdr = l1.substr(1); seq.clear();
                } else seq += l1;
            }
            if (!hdr.empty()) tmp.push_back(makeRead(hdr, seq, std::string(seq.size(), (char)Read::REF_QUAL), fileNum));
        } else if (first != EOF) LOG_THROW("Unrecognised sequence file format: " << path);
        // contiguous slice of the records; a cut never separates two mates (the reference's readers re-synchronise on
        // record and pair boundaries after seeking, src/ReadFileReader.h:379-398)
        const size_t n = tmp.size();
        size_t a = n * (size_t)rank / (size_t)size, b = n * (size_t)(rank + 1) / (size_t)size;
        if (size > 1) { a = pairAlignedCut(tmp, a); b = pairAlignedCut(tmp, b); }
        for (size_t i = a; i < b; ++i) append(tmp[i]);
    }
    // first index >= cut that does not separate record cut-1 from its mate at cut.  Mates are adjacent records, so the
    // records before the cut are paired off from the start of the file exactly as identifyPairs() does
    static size_t pairAlignedCut(const std::vector<Read> &v, size_t cut)
    {
        if (cut == 0 || cut >= v.size()) return cut;
        size_t i = 0;
        while (i < cut) {
            if (i + 1 < v.size()) {
                std::string c1, c2;
                const int n1 = readNum(v[i], c1), n2 = readNum(v[i + 1], c2);
                if (n1 && n2 && n1 != n2 && c1 == c2) { i += 2; continue; }
            }
            ++i;
        }
        return i;
    }
    void append(const Read &r)
    {
        _reads.push_back(r);
        _baseCount += r.getLength();
        if (r.getLength() > _maxLength) _maxLength = r.getLength();
    }

    ReadSetSizeType getSize() const { return _reads.size(); }
    unsigned long getBaseCount() const { return _baseCount; }
    unsigned int getMaxSequenceLength() const { return _maxLength; }
    const Read &getRead(ReadSetSizeType i) const { return _reads[i]; }
    Read &getRead(ReadSetSizeType i) { return _reads[i]; }
    bool isValidRead(ReadSetSizeType i) const { return i < _reads.size(); }
    ReadSetSizeType getPairSize() const { return _pairs.size(); }
    const Pair &getPair(ReadSetSizeType i) const { return _pairs[i]; }
    bool hasPairs() const { return _pairs.size() != 0 && _pairs.size() < _reads.size(); }   // src/ReadSet.h:526-529
    unsigned int getReadFileNum(ReadSetSizeType i) const { return _reads[i].fileNum; }
    std::string getReadFileNamePrefix(ReadSetSizeType i) const { return Options::getOptions().getInputFileSubstring(getReadFileNum(i) - 1); }
    void recount()
    {
        _baseCount = 0; _maxLength = 0;
        for (size_t i = 0; i < _reads.size(); ++i) { _baseCount += _reads[i].getLength(); if (_reads[i].getLength() > _maxLength) _maxLength = _reads[i].getLength(); }
    }

    // adjacent mates: equal common name (name minus its last char when it ends /1 /2 /A /B /F /R) and different read
    // numbers, or Casava-1.8 comments "1:N:..." / "2:N:..."           src/Utils.h:669-733, src/ReadSet.cpp:94-118,463-478
    ReadSetSizeType identifyPairs()
    {
        _pairs.clear();
        const ReadSetSizeType n = _reads.size();
        ReadSetSizeType i = 0;
        while (i < n) {
            if (i + 1 < n) {
                std::string c1, c2;
                int n1 = readNum(_reads[i], c1), n2 = readNum(_reads[i + 1], c2);
                if (n1 && n2 && n1 != n2 && c1 == c2) { _pairs.push_back(Pair(i, i + 1)); i += 2; continue; }
            }
            _pairs.push_back(Pair(i));
            ++i;
        }
        return _pairs.size();
    }

    // concatenated ASCII batch of reads [r0, r1) for kmn_count_batch / kmn_trim_batch
    void concat(ReadSetSizeType r0, ReadSetSizeType r1, std::string &bases, std::string &quals, std::vector<uint64_t> &off,
                std::vector<uint8_t> &discarded) const
    {
        bases.clear(); quals.clear(); off.assign(1, 0); discarded.clear();
        for (ReadSetSizeType i = r0; i < r1; ++i) {
            bases += _reads[i].seq; quals += _reads[i].quals;
            off.push_back(bases.size());
            discarded.push_back(_reads[i].discarded ? 1 : 0);
        }
    }
    int getInputFastqBase() const { return _inputBase; }

private:
    static void stripCR(std::string &s) { if (!s.empty() && s[s.size() - 1] == '\r') s.erase(s.size() - 1); }
    static Read makeRead(const std::string &hdr, std::string seq, const std::string &qual, unsigned int fileNum)
    {
        // name = header up to the first whitespace, the rest is the comment (src/Utils.h:561-598); bases are upper-cased
        // (ReadFileReader::nextRead src/ReadFileReader.h:299-322)
        size_t sp = hdr.find_first_of(" \t");
        std::string name = hdr.substr(0, sp), comment = sp == std::string::npos ? "" : hdr.substr(sp + 1);
        for (size_t i = 0; i < seq.size(); ++i) if (seq[i] >= 'a' && seq[i] <= 'z') seq[i] = (char)(seq[i] - 32);
        Read r(name, comment, seq, qual);
        r.fileNum = fileNum;
        return r;
    }
    static int readNum(const Read &r, std::string &common)
    {
        const std::string &nm = r.name;
        common = nm;
        if (nm.size() > 2 && nm[nm.size() - 2] == '/') {
            char c = nm[nm.size() - 1];
            common = nm.substr(0, nm.size() - 1);
            if (c == '1' || c == 'A' || c == 'F') return 1;
            if (c == '2' || c == 'B' || c == 'R') return 2;
            common = nm;
        }
        const std::string &cm = r.comment;                       // Casava 1.8: "1:N:0:ACGT"
        if (cm.size() >= 3 && (cm[0] == '1' || cm[0] == '2') && cm[1] == ':' && (cm[2] == 'Y' || cm[2] == 'N')) return cm[0] - '0';
        return 0;
    }
    // auto-detect the input Phred base (33 <-> 64) and re-express every quality in FASTQ_START_CHAR.  As in the reference
    // only the first 20000 reads are examined (ReadSet::validateFastqStart, src/ReadSet.h:171-194: `getSize() < 20000`), a
    // read whose MINIMUM quality lies outside [base, base + 40] flips the assumed base (Read::validateFastqStart tests the
    // minimum twice, src/Sequence.h:455-480), and a flip far into the file is reported.
public:
    int detectInputBase() const
    {
        int base = Options::getOptions().getFastqBaseQuality();
        const size_t n = std::min<size_t>(_reads.size(), 20000);
        for (size_t i = 0; i < n; ++i) {
            const std::string &q = _reads[i].quals;
            if (q.empty() || (unsigned char)q[0] == Read::REF_QUAL) continue;
            unsigned char m = 255;
            for (size_t j = 0; j < q.size(); ++j) if ((unsigned char)q[j] < m) m = (unsigned char)q[j];
            if ((int)m < base || (int)m > base + 40) {
                if (i > 10000) LOG_WARN(1, "expected base-" << base << " fastq but detected the other base only very far into the file, "
                                           "please make sure standard fastq and illumina fastq are not mixed");
                base = (base == 33) ? 64 : 33;
            }
        }
        return base;
    }
    // base = the agreed input base (FilterReads-P: every rank must rescale by the same amount)
    void normaliseQualities(int base = 0)
    {
        if (base == 0) base = detectInputBase();
        _inputBase = base;
        const int d = Read::FASTQ_START_CHAR() - base;
        if (d == 0) return;
        for (size_t i = 0; i < _reads.size(); ++i) {
            std::string &q = _reads[i].quals;
            if (q.empty() || (unsigned char)q[0] == Read::REF_QUAL) continue;
            for (size_t j = 0; j < q.size(); ++j) q[j] = (char)((unsigned char)q[j] + d);
        }
    }

    std::vector<Read> _reads;
    std::vector<Pair> _pairs;
    unsigned long _baseCount;
    unsigned int _maxLength;
    int _inputBase;
    bool _deferNormalise;
};

#endif
