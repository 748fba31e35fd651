// FilterReads -- drop-in driver for the reference's apps/FilterReads.cpp:83-215 (+ selectReads, apps/FilterReads.h:158-282)
// on the B200 k-mer spectrum path: load -> artifact (quality) filter -> KmerSpectrum build on the GPU -> histogram ->
// purge -> ReadSelector::scoreAndTrimReads on the GPU -> pick -> write "<out>-MinDepth<d>[-MaxDepth<D>]-<inputprefix>.fastq".
// Same option names, positional arguments and output naming as the reference; there is no CPU fallback: without a GPU
// (or without libkmernator_b200.so) it exits 1 with the error of kmn_create.
#include "FilterReads.h"

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
        const std::string loadMmap = KmerSpectrumOptions::getOptions().getLoadKmerMmap();
        if (k > 0 && !loadMmap.empty()) {                                   // apps/FilterReads.cpp:129-130
            spectrum = KS(KS::estimateRawKmers(reads));
            spectrum.restoreMmap(loadMmap);
        } else if (k > 0) {
            long rawKmers = KS::estimateRawKmers(reads);
            LOG_DEBUG(1, "targeting " << rawKmers << " raw kmers for reads ");
            spectrum = KS(rawKmers);
            const std::string sizeHistoryFile = FilterReadsBaseOptions::getOptions().getSizeHistoryFile();
            spectrum.enableSizeTracking(!sizeHistoryFile.empty());
            spectrum.buildKmerSpectrumInParts(reads, KmerSpectrumOptions::getOptions().getBuildPartitions(), outputFilename.empty() ? "" : outputFilename + "-mmap");
            spectrum.optimize();
            spectrum.trackSpectrum(true);
            if (!sizeHistoryFile.empty()) {                                 // apps/FilterReads.cpp:142-146
                LOG_VERBOSE(1, "Writing size history file to: " << sizeHistoryFile);
                std::ofstream of(sizeHistoryFile.c_str());
                of << spectrum.getSizeTracker().toString();
            }
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
