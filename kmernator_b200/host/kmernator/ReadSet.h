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

#include <algorithm>
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
    // the same record appended to a buffer without temporaries (the writer formats millions of them)
    void appendFastq(std::string &out, unsigned int trimOffset, unsigned int trimLength, const std::string &label) const
    {
        out += '@'; out += name;
        if (Options::getOptions().getKeepReadComment() && !comment.empty()) { out += ' '; out += comment; }
        if (!label.empty()) { out += ' '; out += label; }
        if (discarded || trimLength <= 1) { out += "\nN\n+\n"; out += (char)(FASTQ_START_CHAR() + 1); out += '\n'; return; }
        out += '\n';
        out.append(seq, trimOffset, trimLength);
        out += "\n+\n";
        const size_t avail = trimOffset < quals.size() ? std::min<size_t>(trimLength, quals.size() - trimOffset) : 0;
        if (avail && (unsigned char)quals[trimOffset] == REF_QUAL) out.append(trimLength, (char)PRINT_REF_QUAL);
        else out.append(quals, trimOffset, trimLength);
        out += '\n';
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
    // Rank `rank` of `size` parses only its BYTE RANGE of the file: the range starts at the first record boundary at or
    // after fileSize / size * rank and ends at the one at or after fileSize / size * (rank + 1), so the ranks' ranges tile
    // the file without reading each other's part (ReadFileReader::seekToPartition src/ReadFileReader.h:379-398,
    // SequenceStreamParser::seekToNextRecord :657-760).
    void appendAnyFile(const std::string &path, unsigned int fileNum = 1, int rank = 0, int size = 1)
    {
        std::ifstream in(path.c_str(), std::ios::binary);
        if (!in.good()) LOG_THROW("Could not open : " << path);
        const int first = in.peek();
        if (first == EOF) return;
        if (first != '@' && first != '>') LOG_THROW("Unrecognised sequence file format: " << path);
        const char marker = (char)first;
        in.seekg(0, std::ios::end);
        const unsigned long fileSize = (unsigned long)in.tellg();
        unsigned long firstPos = 0, lastPos = fileSize;
        if (size > 1) {
            const unsigned long blockSize = fileSize / (unsigned long)size;
            if (rank + 1 != size) seekToNextRecord(in, blockSize * (unsigned long)(rank + 1), fileSize, marker, lastPos);
            if (!seekToNextRecord(in, blockSize * (unsigned long)rank, fileSize, marker, firstPos)) firstPos = lastPos;
        }
        in.clear();
        in.seekg((std::streamoff)firstPos);
        unsigned long pos = firstPos;
        std::string hdr, seq, qual;
        while (pos < lastPos && readRecord(in, marker, pos, fileSize, hdr, seq, qual, path)) append(makeRead(hdr, seq, qual, fileNum));
    }

private:
    static bool getLine(std::ifstream &in, std::string &l, unsigned long &pos)
    {
        if (!std::getline(in, l)) return false;
        pos += l.size() + (in.eof() ? 0 : 1);
        stripCR(l);
        return true;
    }
    // one record starting at `pos` (a marker line); FASTA sequences may span lines.  false at the end of the file.
    static bool readRecord(std::ifstream &in, char marker, unsigned long &pos, unsigned long fileSize, std::string &hdr, std::string &seq, std::string &qual,
                           const std::string &path)
    {
        std::string l;
        do { if (pos >= fileSize || !getLine(in, l, pos)) return false; } while (l.empty());
        if (l[0] != marker) LOG_THROW("Missing '" << marker << "' in header of " << path << ": " << l);
        hdr = l.substr(1);
        seq.clear(); qual.clear();
        if (marker == '@') {
            std::string plus;
            if (!getLine(in, seq, pos) || !getLine(in, plus, pos) || !getLine(in, qual, pos)) LOG_THROW("Truncated FASTQ record in " << path << ": " << hdr);
            if (seq.size() != qual.size()) LOG_THROW("Number of bases and quals do not match in " << path << ": " << hdr);
        } else {
            while (pos < fileSize && in.peek() != marker && in.peek() != EOF) { if (!getLine(in, l, pos)) break; seq += l; }
            qual.assign(seq.size(), (char)Read::REF_QUAL);
        }
        return true;
    }
    // SequenceStreamParser::seekToNextRecord(minimumPos, byPair = true): the first record boundary at or after minimumPos
    // that does not split two adjacent mates.  false (pos = end of file) when there is none.
    static bool seekToNextRecord(std::ifstream &in, unsigned long minimumPos, unsigned long fileSize, char marker, unsigned long &out)
    {
        out = fileSize;
        in.clear();
        unsigned long pos = 0;
        std::string l;
        if (minimumPos > 0) {
            if (minimumPos - 1 >= fileSize) return false;
            in.seekg((std::streamoff)(minimumPos - 1));
            pos = minimumPos - 1;
            if (in.peek() == '\n') { in.seekg((std::streamoff)minimumPos); pos = minimumPos; }
            else if (!getLine(in, l, pos)) return false;               // the rest of the line minimumPos falls into
        } else { in.seekg(0); out = 0; return true; }
        while (pos < fileSize && in.peek() != EOF && in.peek() != marker) if (!getLine(in, l, pos)) return false;
        if (pos >= fileSize || in.peek() == EOF) return false;
        if (marker == '@') {
            // '@' is a valid quality character: when the NEXT line starts with '@' too, this one was a quality line and the
            // record starts there (:688-707)
            const unsigned long here = pos;
            if (!getLine(in, l, pos)) return false;
            if (pos >= fileSize || in.peek() != marker) { in.clear(); in.seekg((std::streamoff)here); pos = here; }
        }
        if (pos >= fileSize) return false;
        // do not split a pair of adjacent mates (:712-757)
        const unsigned long here1 = pos;
        std::string h1, h2, h3, sq, ql;
        if (!readRecord(in, marker, pos, fileSize, h1, sq, ql, "") || pos >= fileSize) return false;
        const unsigned long here2 = pos;
        if (!readRecord(in, marker, pos, fileSize, h2, sq, ql, "") || pos >= fileSize) return false;
        if (!readRecord(in, marker, pos, fileSize, h3, sq, ql, "")) { out = here1; return true; }
        const Read r1 = makeRead(h1, "", "", 0), r2 = makeRead(h2, "", "", 0), r3 = makeRead(h3, "", "", 0);
        std::string c1, c2, c3;
        const int n1 = readNum(r1, c1), n2 = readNum(r2, c2), n3 = readNum(r3, c3);
        if (n1 && n2 && n1 != n2 && c1 == c2) out = here1;             // a natural pair starts here
        else if (n2 && n3 && n2 != n3 && c2 == c3) out = here2;        // the boundary fell between two mates: one record further
        else out = here1;
        return true;
    }

public:
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
