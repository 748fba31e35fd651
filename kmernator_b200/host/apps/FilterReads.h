// FilterReads.h -- what apps/FilterReads.cpp and apps/FilterReads-P.cpp share, as in the reference (apps/FilterReads.h:158-282):
// selectReads() = pick the passing reads / pairs (or the coverage-normalised subset) and write them.
#ifndef KMERNATOR_HOST_APPS_FILTERREADS_H
#define KMERNATOR_HOST_APPS_FILTERREADS_H

#include <fstream>
#include <iostream>

#include "../kmernator/FilterKnownOddities.h"
#include "../kmernator/KmerSpectrum.h"
#include "../kmernator/Options.h"
#include "../kmernator/ReadSelector.h"
#include "../kmernator/ReadSet.h"

typedef KmerSpectrum KS;
typedef ReadSelector RS;

template <typename T> static std::string toStr(T v) { std::ostringstream ss; ss << v; return ss.str(); }

// apps/FilterReads.h:158-282 without the OPTIMAL / partition-by-depth branches (serial-only, experimental)
// outputSuffix: appended to every file name (the distributed driver writes per-rank parts and joins them in rank order)
static long selectReads(unsigned int minDepth, ReadSet &reads, RS &selector, std::string outputFilename, const std::string &outputSuffix = "")
{
    LOG_VERBOSE(1, "selectReads with minDepth " << minDepth << ", minLength " << ReadSelectorOptions::getOptions().getMinReadLength() << ": " << reads.getSize() << " reads");
    long picked = 0;
    const int maximumKmerDepth = ReadSelectorOptions::getOptions().getMaxKmerDepth();
    std::string suffix;
    if (ReadSelectorOptions::getOptions().getSeparateOutputs()) {
        if (KmerBaseOptions::getOptions().getKmerSize() > 0) outputFilename += "-MinDepth" + toStr(minDepth);
        suffix = (Options::getOptions().getFormatOutput() & 1) ? ".fasta" : ".fastq";
    }
    suffix += outputSuffix;
    const float minLen = ReadSelectorOptions::getOptions().getMinReadLength();
    const bool bothPairs = ReadSelectorOptions::getOptions().getBothPairs();
    if (maximumKmerDepth > 0) {
        if (ReadSelectorOptions::getOptions().getSeparateOutputs()) outputFilename += "-MaxDepth" + toStr(maximumKmerDepth);
        RS::OFM ofmap = selector.getOFM(outputFilename, suffix);
        if (ReadSelectorOptions::getOptions().getNormalizationMethod() != "RANDOM")
            LOG_THROW("normalization-method " << ReadSelectorOptions::getOptions().getNormalizationMethod() << " is not implemented (RANDOM only)");
        picked += selector.pickCoverageNormalizedSubset(maximumKmerDepth, minDepth, minLen, reads.hasPairs(), bothPairs);
        if (picked > 0 && !outputFilename.empty()) {
            LOG_VERBOSE(1, "Writing " << picked << " reads to output file(s)");
            selector.writePicks(ofmap, 0);
        }
    } else {
        float tmpMinDepth = (float)minDepth;
        if (KmerBaseOptions::getOptions().getKmerSize() == 0) tmpMinDepth = 0;
        RS::OFM ofmap = selector.getOFM(outputFilename, suffix);
        LOG_VERBOSE(1, "Selecting reads over depth: " << tmpMinDepth);
        if (reads.hasPairs()) picked = selector.pickAllPassingPairs(tmpMinDepth, minLen, bothPairs);
        else picked = selector.pickAllPassingReads(tmpMinDepth, minLen);
        LOG_VERBOSE(2, "At or above coverage: " << tmpMinDepth << " Picked " << picked << " / " << reads.getSize() << " reads");
        if (!outputFilename.empty()) {
            LOG_VERBOSE(1, "Writing " << picked << " reads to output files");
            selector.writePicks(ofmap, 0);
        }
    }
    return picked;
}

#endif
