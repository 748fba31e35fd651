// FilterKnownOddities.h -- the default-on artifact pre-filter that runs immediately before the k-mer spectrum path
// (SURVEY.md section 8f row f2).  Implemented: the quality-stretch trim -- the part that decides the reference's
// 1000-Filtered*.fastq goldens (section 3.4 G):
//   longest stretch with qual >= START+minQuality (first-longest)      src/FilterKnownOddities.h:407-441
//   not the whole read: replaced by the stretch with comment "AFTrim:a+len" when passesLength, else DISCARDED
//                                                                       src/FilterKnownOddities.h:321-334,523-533,613-632
// Not implemented (documented gap): the 24-mer adapter / homopolymer screen with edit distance
// (src/FilterKnownOddities.h:444-521,742-795) and the optional simple-repeat / PhiX screens.
#ifndef KMERNATOR_HOST_FILTERKNOWNODDITIES_H
#define KMERNATOR_HOST_FILTERKNOWNODDITIES_H

#include <sstream>

#include "ReadSelector.h"
#include "ReadSet.h"

class FilterKnownOddities {
public:
    unsigned long applyFilter(ReadSet &reads)
    {
        const int start = Read::FASTQ_START_CHAR();
        const int minQuality = (int)Options::getOptions().getMinQuality();
        const float minReadLength = ReadSelectorOptions::getOptions().getMinReadLength();
        unsigned long affected = 0;
        for (ReadSet::ReadSetSizeType i = 0; i < reads.getSize(); ++i) {
            Read &r = reads.getRead(i);
            const std::string &q = r.getQuals();
            const size_t n = q.size();
            if (n == 0 || (unsigned char)q[0] == Read::REF_QUAL) continue;
            size_t bestOff = 0, bestLen = 0, st = 0;
            for (size_t j = 0; j < n; ++j) {
                if ((int)(unsigned char)q[j] < start + minQuality) {
                    if (j - st > bestLen) { bestLen = j - st; bestOff = st; }
                    st = j + 1;
                }
            }
            if (n - st > bestLen) { bestLen = n - st; bestOff = st; }
            if (bestOff == 0 && bestLen == n) continue;
            affected++;
            if (bestLen == 0 || !ReadSelectorUtil::passesLength((float)bestLen, r.getLength(), minReadLength)) {
                r.discard();
            } else {
                std::ostringstream ss;
                ss << "AFTrim:" << bestOff << "+" << bestLen;
                r.seq = r.seq.substr(bestOff, bestLen);
                r.quals = r.quals.substr(bestOff, bestLen);
                r.addComment(ss.str());
            }
        }
        reads.recount();
        return affected;
    }
};

#endif
