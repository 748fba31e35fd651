// FilterReads -- drop-in driver for the reference's apps/FilterReads.cpp:83-215 (+ selectReads, apps/FilterReads.h:158-282)
// on the B200 k-mer spectrum path: load -> artifact (quality) filter -> KmerSpectrum build on the GPU -> histogram ->
// purge -> ReadSelector::scoreAndTrimReads on the GPU -> pick -> write "<out>-MinDepth<d>[-MaxDepth<D>]-<inputprefix>.fastq".
// Same option names, positional arguments and output naming as the reference; there is no CPU fallback: without a GPU
// (or without libkmernator_b200.so) it exits 1 with the error of kmn_create.
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
static long selectReads(unsigned int minDepth, ReadSet &reads, RS &selector, std::string outputFilename)
{
    LOG_VERBOSE(1, "selectReads with minDepth " << minDepth << ", minLength " << ReadSelectorOptions::getOptions().getMinReadLength() << ": " << reads.getSize() << " reads");
    long picked = 0;
    const int maximumKmerDepth = ReadSelectorOptions::getOptions().getMaxKmerDepth();
    std::string suffix;
    if (ReadSelectorOptions::getOptions().getSeparateOutputs()) {
        if (KmerBaseOptions::getOptions().getKmerSize() > 0) outputFilename += "-MinDepth" + toStr(minDepth);
        suffix = (Options::getOptions().getFormatOutput() & 1) ? ".fasta" : ".fastq";
    }
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

int main(int argc, char *argv[])
{
    if (!FilterReadsOptions::parseOpts(argc, argv)) return 1;
    Read::FASTQ_START_CHAR() = Options::getOptions().getOutputFastqBaseQuality();
    std::string outputFilename = Options::getOptions().getOutputFile();
    ReadSet reads;
    try {
        OptionsBaseInterface::FileListType &inputs = Options::getOptions().getInputFiles();
        LOG_VERBOSE(1, "Reading Input Files");
        reads.appendAllFiles(inputs);
        LOG_VERBOSE(1, "loaded " << reads.getSize() << " Reads, " << reads.getBaseCount() << " Bases ");
        LOG_VERBOSE(1, "Identifying Pairs: ");
        long numPairs = reads.identifyPairs();
        LOG_VERBOSE(1, "Pairs + single = " << numPairs);

        if (!FilterKnownOdditiesOptions::getOptions().getSkipArtifactFilter()) {
            LOG_VERBOSE(1, "Preparing artifact filter: ");
            FilterKnownOddities filter;
            unsigned long filtered = filter.applyFilter(reads);
            LOG_VERBOSE(1, "filter affected (trimmed/removed) " << filtered << " Reads ");
        }

        KS spectrum(0);
        const unsigned int k = KmerBaseOptions::getOptions().getKmerSize();
        if (k > 0) {
            long rawKmers = KS::estimateRawKmers(reads);
            LOG_DEBUG(1, "targeting " << rawKmers << " raw kmers for reads ");
            spectrum = KS(rawKmers);
            spectrum.buildKmerSpectrumInParts(reads, KmerSpectrumOptions::getOptions().getBuildPartitions(), outputFilename.empty() ? "" : outputFilename + "-mmap");
            spectrum.optimize();
            spectrum.trackSpectrum(true);
            if (Log::isVerbose(1)) {
                kmn_stats st = spectrum.getStats();
                LOG_VERBOSE(1, "Kmer counters: raw " << st.raw_kmers << ", rawGood " << st.raw_good_kmers << ", unique " << st.unique_kmers
                               << ", discarded " << st.discarded_kmers);
                std::cerr << "Kmer Histogram" << std::endl;
                spectrum.printHistograms(std::cerr);
            }
            if (!FilterReadsBaseOptions::getOptions().getHistogramFile().empty()) {
                std::ofstream of(FilterReadsBaseOptions::getOptions().getHistogramFile().c_str());
                spectrum.printHistograms(of);
            }
        }
        unsigned int minDepth = KmerSpectrumOptions::getOptions().getMinDepth();
        if (k > 0) {
            if (minDepth > 1) spectrum.purgeMinDepth(minDepth, true);
            else spectrum.optimize(true);
        }
        if (!outputFilename.empty()) {
            if (k > 0) LOG_VERBOSE(1, "Trimming reads with minDepth: " << minDepth);
            else LOG_VERBOSE(1, "Trimming reads that pass Artifact Filter with length: " << ReadSelectorOptions::getOptions().getMinReadLength());
            RS selector(reads, spectrum.weak);
            selector.scoreAndTrimReads((float)minDepth);
            selectReads(minDepth, reads, selector, outputFilename);
        }
        spectrum.reset();
    } catch (std::exception &e) {
        LOG_ERROR(1, "FilterReads threw an exception!\n\t" << e.what());
        return 1;
    } catch (...) {
        LOG_ERROR(1, "FilterReads threw an error!");
        return 1;
    }
    LOG_VERBOSE(1, "Finished");
    return 0;
}
