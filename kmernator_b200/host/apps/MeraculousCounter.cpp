// MeraculousCounter -- drop-in driver for the reference's apps/MeraculousCounter.cpp:110-151 on the B200 k-mer spectrum
// path: load -> KmerSpectrum build on the GPU with extension tracking (ExtensionTrackingData, src/Meraculous.h:79-80) ->
// "<out>.mercount.m<k>" and "<out>.mergraph.m<k>.D<minDepth>" (src/Meraculous.h:107-133).  One process per GPU: with more
// than one rank (RANK / WORLD_SIZE from the launcher) every rank counts its slice of the reads, owns the k-mers the
// reference's hash assigns to it and writes its part; rank 0 joins the parts in rank order.  No artifact filter, no purge
// (buildKmerSpectrum(reads, false)); defaults as in apps/MeraculousCounter.cpp:66-80.
#include <fstream>
#include <iostream>

#include "../kmernator/KmerSpectrum.h"
#include "../kmernator/Options.h"
#include "../kmernator/ReadSet.h"
#include "../kmernator/World.h"

static void meraculousDefaults()
{
    FilterReadsOptions::setDefault("verbose", "2");
    FilterReadsOptions::setDefault("min-quality-score", "2");
    FilterReadsOptions::setDefault("min-kmer-quality", "0");
    FilterReadsOptions::setDefault("kmer-size", "0");                    // "The Kmer size can not be 0": must be given
}

static void joinParts(World &world, const std::string &name)
{
    world.barrier();                                                     // every part is closed and complete
    if (world.rank() != 0) return;
    std::ofstream out(name.c_str(), std::ios::binary);
    if (!out.good()) LOG_THROW("Could not open " << name << " for writing");
    for (int r = 0; r < world.size(); ++r) {
        std::ostringstream pn;
        pn << name << ".rank" << r;
        std::ifstream in(pn.str().c_str(), std::ios::binary);
        if (in.good()) { if (in.peek() != EOF) out << in.rdbuf(); in.close(); remove(pn.str().c_str()); }
    }
}

int main(int argc, char *argv[])
{
    if (!FilterReadsOptions::parseOpts(argc, argv, meraculousDefaults, false)) return 1;
    Read::FASTQ_START_CHAR() = Options::getOptions().getOutputFastqBaseQuality();
    const std::string outputFilename = Options::getOptions().getOutputFile();
    try {
        const unsigned int k = KmerBaseOptions::getOptions().getKmerSize();
        if (k == 0) LOG_THROW("The Kmer size can not be 0");
        World world;
        World::instance() = &world;
        const bool showHistogram = Log::isVerbose(1);                    // decided before the other ranks go quiet: the histogram is collective
        if (world.rank() != 0 && !Options::getOptions().getDebug()) Log::verboseLevel() = 0;
        LOG_VERBOSE(1, "Reading Input Files");
        ReadSet reads;
        reads.deferNormalise();
        reads.appendAllFiles(Options::getOptions().getInputFiles(), world.rank(), world.size());
        {
            const int dflt = Options::getOptions().getFastqBaseQuality();
            const unsigned long flipped = world.allMax(reads.detectInputBase() != dflt ? 1ul : 0ul);
            reads.normaliseQualities(flipped ? (dflt == 33 ? 64 : 33) : dflt);
        }
        KmerSpectrum spectrum(world, KmerSpectrum::estimateRawKmers(world, reads), KMN_VALUE_DIR_EXT);
        spectrum.buildKmerSpectrum(reads);
        const int minDepth = (int)KmerSpectrumOptions::getOptions().getMinDepth();
        if (showHistogram) {
            const std::string hist = spectrum.getHistogram(false, 255);
            if (world.rank() == 0) std::cerr << "Collective Kmer Histogram" << std::endl << hist;
        }
        std::ostringstream n1, n2, part;
        n1 << outputFilename << ".mercount.m" << k;
        n2 << outputFilename << ".mergraph.m" << k << ".D" << minDepth;
        if (world.size() > 1) part << ".rank" << world.rank();
        {
            std::ofstream of((n1.str() + part.str()).c_str());
            if (!of.good()) LOG_THROW("Could not open " << n1.str() << " for writing");
            spectrum.dumpCounts(of, minDepth);
        }
        {
            std::ofstream of((n2.str() + part.str()).c_str());
            if (!of.good()) LOG_THROW("Could not open " << n2.str() << " for writing");
            spectrum.dumpGraphs(of, minDepth);
        }
        if (world.size() > 1) { joinParts(world, n1.str()); joinParts(world, n2.str()); }
        spectrum.reset();
        world.finalize();
        World::instance() = NULL;
        LOG_VERBOSE(1, "Finished");
    } catch (std::exception &e) {
        LOG_ERROR(1, "MeraculousCounter threw an exception! Aborting...\n\t" << e.what());
        return 1;
    } catch (...) {
        LOG_ERROR(1, "MeraculousCounter threw an error!");
        return 1;
    }
    return 0;
}
