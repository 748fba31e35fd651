// FilterReads-P -- drop-in driver for the reference's distributed app (apps/FilterReads-P.cpp:102-201,263-325) on the B200
// k-mer spectrum path: one process per GPU (torchrun / mpirun / srun set the rank), every rank loads its slice of the
// input files, the k-mer table is sharded by hash owner inside the library (kmn_comm_init), the lookup pass is collective,
// and the per-rank outputs are joined in rank order, so the result is the file the serial driver writes -- the
// reference's own criterion for this app (test/runFilterTests.sh:93-116).
// Same option names, positional arguments and output naming as FilterReads.  No MPI: rank/size come from the launcher's
// environment and the few host-side agreements go through kmernator/World.h.
#include <algorithm>
#include <set>

#include "../kmernator/World.h"
#include "FilterReads.h"

int main(int argc, char *argv[])
{
    if (!FilterReadsOptions::parseOpts(argc, argv)) return 1;
    Read::FASTQ_START_CHAR() = Options::getOptions().getOutputFastqBaseQuality();
    std::string outputFilename = Options::getOptions().getOutputFile();
    ReadSet reads;
    try {
        World world;
        World::instance() = &world;
        const bool root = world.rank() == 0;
        if (!root && !Options::getOptions().getDebug()) Log::verboseLevel() = 0;       // gathered logs: rank 0 speaks
        OptionsBaseInterface::FileListType &inputs = Options::getOptions().getInputFiles();
        LOG_VERBOSE(1, "Reading Input Files (rank " << world.rank() << " of " << world.size() << ")");
        reads.deferNormalise();
        reads.appendAllFiles(inputs, world.rank(), world.size());
        {   // every rank must rescale its qualities by the same amount: a rank whose slice shows evidence against the
            // configured input base (ReadSet::validateFastqStart, src/ReadSet.h:171-194) makes all ranks switch
            const int dflt = Options::getOptions().getFastqBaseQuality();
            const unsigned long flipped = world.allMax(reads.detectInputBase() != dflt ? 1ul : 0ul);
            reads.normaliseQualities(flipped ? (dflt == 33 ? 64 : 33) : dflt);
        }
        LOG_VERBOSE(1, "loaded " << reads.getSize() << " Reads, " << reads.getBaseCount() << " Bases ");
        long numPairs = reads.identifyPairs();
        LOG_VERBOSE(1, "Pairs + single = " << numPairs);
        // setGlobalReadSetConstants (src/DistributedFunctions.h:73-100): global offset and size of this rank's reads
        const unsigned long globalSize = world.allSum(reads.getSize());
        LOG_VERBOSE(1, "Global reads: " << globalSize);

        if (!FilterKnownOdditiesOptions::getOptions().getSkipArtifactFilter()) {
            FilterKnownOddities filter;
            unsigned long filtered = world.allSum(filter.applyFilter(reads));
            LOG_VERBOSE(1, "filter affected (trimmed/removed) " << filtered << " Reads ");
        }

        KS spectrum(0);
        const unsigned int k = KmerBaseOptions::getOptions().getKmerSize();
        if (k > 0) {
            unsigned long rawKmers = KS::estimateRawKmers(world, reads);
            LOG_DEBUG(1, "targeting " << rawKmers << " raw kmers per rank");
            // --reference-file / --subtract-file: a second spectrum whose k-mers leave the main one
            // (apps/FilterReads-P.cpp:281-308, src/KmerSpectrum.h:1582-1589).  reference files are counted as they are,
            // subtract files go through the artifact filter and the min-depth purge like the input
            KS subtracting(0);
            OptionsBaseInterface::FileListType referenceFiles = FilterReadsBaseOptions::getOptions().getReferenceFiles();
            OptionsBaseInterface::FileListType subtractFiles = FilterReadsBaseOptions::getOptions().getSubtractFiles();
            if (!referenceFiles.empty() || !subtractFiles.empty()) {
                subtracting = KS(world, rawKmers);
                if (!referenceFiles.empty()) {
                    LOG_VERBOSE(1, "Subtracting reference-file set");
                    ReadSet refReads;
                    refReads.appendAllFiles(referenceFiles, world.rank(), world.size());
                    subtracting.buildKmerSpectrum(refReads, false);
                }
                if (!subtractFiles.empty()) {
                    LOG_VERBOSE(1, "Subtracting abundant kmers in subtract-file set");
                    ReadSet subReads;
                    subReads.appendAllFiles(subtractFiles, world.rank(), world.size());
                    subReads.identifyPairs();
                    if (!FilterKnownOdditiesOptions::getOptions().getSkipArtifactFilter()) { FilterKnownOddities f; f.applyFilter(subReads); }
                    subtracting.buildKmerSpectrum(subReads, false);
                }
                subtracting.finishBuild();
                if (!subtractFiles.empty() && KmerSpectrumOptions::getOptions().getMinDepth() > 1)
                    subtracting.purgeMinDepth(KmerSpectrumOptions::getOptions().getMinDepth(), true);
            }
            spectrum = KS(world, rawKmers);
            spectrum.buildKmerSpectrum(reads);
            if (subtracting.weak.ctx) spectrum.subtractReference(subtracting);
            subtracting.reset();
            spectrum.purgeMinDepth(KmerSpectrumOptions::getOptions().getMinDepth());
            // the histogram is all-reduced inside the library: every rank asks, rank 0 prints / writes
            const std::string hist = spectrum.getHistogram(false, 255);               // MPIHistogram(255), src/DistributedFunctions.h:575
            if (root) {
                if (Log::isVerbose(1)) std::cerr << "Collective Kmer Histogram" << std::endl << hist;
                if (!FilterReadsBaseOptions::getOptions().getHistogramFile().empty()) {
                    std::ofstream of(FilterReadsBaseOptions::getOptions().getHistogramFile().c_str());
                    of << hist;
                }
            }
        }
        unsigned int minDepth = KmerSpectrumOptions::getOptions().getMinDepth();
        if (k > 0 && minDepth > 1) spectrum.purgeMinDepth(minDepth, true);
        if (!outputFilename.empty()) {
            std::ostringstream part;
            part << ".rank" << world.rank();
            {
                RS selector(reads, spectrum.weak);
                selector.scoreAndTrimReads((float)minDepth);
                selectReads(minDepth, reads, selector, outputFilename, world.size() > 1 ? part.str() : std::string());
            }
            if (world.size() > 1) {
                // rank-ordered concatenation of the parts (DistributedOfstreamMap, src/DistributedOfstreamMap.h:245-395)
                std::string mine;
                const std::vector<std::string> &made = OfstreamMap::created();
                for (size_t i = 0; i < made.size(); ++i) mine += made[i].substr(0, made[i].size() - part.str().size()) + "\n";
                std::vector<std::string> all = world.allGather(mine);                  // also: every part is closed and complete
                if (root) {
                    std::set<std::string> names;
                    for (size_t r = 0; r < all.size(); ++r) {
                        std::istringstream is(all[r]);
                        std::string l;
                        while (std::getline(is, l)) if (!l.empty()) names.insert(l);
                    }
                    for (std::set<std::string>::const_iterator it = names.begin(); it != names.end(); ++it) {
                        std::ofstream out(it->c_str(), std::ios::binary);
                        if (!out.good()) LOG_THROW("Could not open " << *it << " for writing");
                        for (int r = 0; r < world.size(); ++r) {
                            std::ostringstream pn;
                            pn << *it << ".rank" << r;
                            std::ifstream in(pn.str().c_str(), std::ios::binary);
                            if (in.good()) { if (in.peek() != EOF) out << in.rdbuf(); in.close(); remove(pn.str().c_str()); }
                        }
                        LOG_VERBOSE(1, "Joined " << world.size() << " parts into " << *it);
                    }
                }
            }
        }
        spectrum.reset();
        world.finalize();
        World::instance() = NULL;
    } catch (std::exception &e) {
        LOG_ERROR(1, "FilterReads-P threw an exception!\n\t" << e.what());
        return 1;
    } catch (...) {
        LOG_ERROR(1, "FilterReads-P threw an error!");
        return 1;
    }
    LOG_VERBOSE(1, "Finished");
    return 0;
}
