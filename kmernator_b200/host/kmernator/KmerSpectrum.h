// KmerSpectrum.h -- the reference's KmerSpectrum call surface on top of the C ABI
// (include/kmernator_b200.h).  The table lives in HBM; this class only drives it.
//   KS::estimateRawKmers(reads)                      src/KmerSpectrum.h:573-584
//   KS spectrum(rawKmers)                            src/KmerSpectrum.h:414-421
//   buildKmerSpectrum / buildKmerSpectrumInParts     src/KmerSpectrum.h:1818-1902,2081-2115
//   purgeMinDepth / getHistogram / printHistograms   src/KmerSpectrum.h:1805-1815,909-1071
//   optimize / trackSpectrum / reset                 src/KmerSpectrum.h:465,1574,533
//   public member `weak` handed to the ReadSelector  apps/FilterReads.cpp:196
// Errors from the C ABI become LoggedException, like LOG_THROW in the reference (src/Log.h:442-456,484).
#ifndef KMERNATOR_HOST_KMERSPECTRUM_H
#define KMERNATOR_HOST_KMERSPECTRUM_H

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/kmernator_b200.h"
#include "Log.h"
#include "Options.h"
#include "ReadSet.h"
#include "World.h"

#define KMN_CHECK(ctx, call)                                                                        \
    do {                                                                                            \
        int rc_ = (call);                                                                           \
        if (rc_ != 0) LOG_THROW(#call << " failed (" << rc_ << "): " << kmn_last_error(ctx));       \
    } while (0)

// the lookup handle the reference passes around as `spectrum.weak` (a KmerMap); here: the device context
struct KmerMapHandle {
    kmn_ctx *ctx;
    KmerMapHandle() : ctx(NULL) {}
};

class KmerSpectrum {
public:
    KmerMapHandle weak;

    // distributed: every rank sizes its table for the largest local estimate, so that all ranks have the same table and
    // staging geometry (DistributedKmerSpectrum::estimateRawKmers(world, reads), src/DistributedFunctions.h:154-161)
    static unsigned long estimateRawKmers(World &world, const ReadSet &reads) { return world.allMax(estimateRawKmers(reads)); }

    // (avgLen - k + 1) * numReads, at least 128                          src/KmerSpectrum.h:573-584
    static unsigned long estimateRawKmers(const ReadSet &reads)
    {
        if (reads.getSize() == 0) return 128;
        unsigned long k = KmerBaseOptions::getOptions().getKmerSize();
        unsigned long avg = reads.getBaseCount() / reads.getSize();
        unsigned long kmers = avg >= k ? (avg - k + 1) * reads.getSize() : 0;
        return kmers < 128 ? 128 : kmers;
    }

    // KS spectrum(world, rawKmers): one context per rank on its local GPU, joined to the library's communicator with an
    // NCCL id that rank 0 creates and the world hands round      src/DistributedFunctions.h:126-131, src/MPIUtils.h:256-391
    KmerSpectrum(World &world, unsigned long rawKmers, unsigned int valueKind = KMN_VALUE_DIR) : KmerSpectrum(rawKmers, world.localRank(), valueKind)
    {
        if (!weak.ctx || world.size() == 1) return;
        std::string id(128, '\0');
        if (world.rank() == 0 && kmn_comm_unique_id(&id[0]) != 0) LOG_THROW("kmn_comm_unique_id failed");
        id = world.broadcast(id);
        if (id.size() != 128) LOG_THROW("bad communicator id from rank 0");
        KMN_CHECK(weak.ctx, kmn_comm_init(weak.ctx, world.rank(), world.size(), id.data()));
    }

    explicit KmerSpectrum(unsigned long rawKmers = 0, int device = 0, unsigned int valueKind = KMN_VALUE_DIR) : _rawKmers(rawKmers)
    {
        if (rawKmers == 0) return;                                       // KS spectrum(0): placeholder (apps/FilterReads.cpp:128)
        kmn_opts o;
        kmn_default_opts(&o);
        o.kmer_size = KmerBaseOptions::getOptions().getKmerSize();
        o.fastq_start_char = (uint32_t)Read::FASTQ_START_CHAR();
        o.min_quality_score = Options::getOptions().getMinQuality();
        o.min_kmer_quality = (float)KmerSpectrumOptions::getOptions().getMinKmerQuality();
        o.min_depth = KmerSpectrumOptions::getOptions().getMinDepth();
        o.value_kind = valueKind | KMN_VALUE_WEIGHTS;                    // the histogram prints the weight columns
        o.est_raw_kmers = rawKmers;
        // the reference sizes weak by est/estimatedDepth and singleton by est*estimatedErrorRate (src/KmerSpectrum.h:414-421);
        // one open-addressing table takes both, at load factor 0.5
        double distinct = (double)rawKmers / KmerSpectrumOptions::getOptions().getEstimatedDepth() +
                          (double)rawKmers * KmerSpectrumOptions::getOptions().getEstimatedErrorRate();
        o.table_slots = (uint64_t)(distinct / 0.5) + 4096;
        o.ignore_quality = Options::getOptions().getIgnoreQual() ? 1 : 0;
        o.device = (uint32_t)device;
        kmn_ctx *c = NULL;
        int rc = kmn_create(&c, &o);
        if (rc != 0) LOG_THROW("kmn_create failed (" << rc << "): " << kmn_last_error(NULL));
        weak.ctx = c;
    }
    ~KmerSpectrum() { reset(); }
    KmerSpectrum(const KmerSpectrum &) = delete;
    KmerSpectrum &operator=(const KmerSpectrum &) = delete;
    KmerSpectrum &operator=(KmerSpectrum &&o)                            // spectrum = KS(rawKmers)  apps/FilterReads.cpp:138
    {
        if (this != &o) { reset(); weak = o.weak; _rawKmers = o._rawKmers; o.weak.ctx = NULL; }
        return *this;
    }
    KmerSpectrum(KmerSpectrum &&o) : weak(o.weak), _rawKmers(o._rawKmers) { o.weak.ctx = NULL; }

    // count pass over the whole ReadSet in --batch-size batches           src/KmerSpectrum.h:2081-2115
    void buildKmerSpectrum(const ReadSet &reads)
    {
        if (!weak.ctx) LOG_THROW("buildKmerSpectrum on an empty KmerSpectrum");
        const ReadSet::ReadSetSizeType n = reads.getSize();
        ReadSet::ReadSetSizeType batch = Options::getOptions().getBatchSize();
        if (batch == 0) batch = 100000;
        std::string bases, quals;
        std::vector<uint64_t> off;
        std::vector<uint8_t> disc;
        // distributed: every rank makes the same number of calls (a rank that has run out of reads passes empty batches),
        // as every MPI rank must keep calling sendReceive in the reference (src/DistributedFunctions.h:436-447)
        unsigned long nBatches = (n + batch - 1) / batch;
        if (World::instance()) nBatches = World::instance()->allMax(nBatches);
        for (unsigned long b = 0; b < nBatches; ++b) {
            const ReadSet::ReadSetSizeType r0 = std::min<ReadSet::ReadSetSizeType>(n, b * batch), r1 = std::min<ReadSet::ReadSetSizeType>(n, r0 + batch);
            reads.concat(r0, r1, bases, quals, off, disc);
            KMN_CHECK(weak.ctx, kmn_count_batch(weak.ctx, (const uint8_t *)bases.data(), (const uint8_t *)quals.data(), off.data(), r1 - r0, disc.data()));
        }
        KMN_CHECK(weak.ctx, kmn_count_finish(weak.ctx, 0));
    }
    // build + post-build purge (src/KmerSpectrum.h:1818-1831); more than one part is never needed in HBM
    void buildKmerSpectrumInParts(const ReadSet &reads, unsigned int /*numParts*/, const std::string & /*mmapPrefix*/ = "")
    {
        buildKmerSpectrum(reads);
        purgeMinDepth(KmerSpectrumOptions::getOptions().getMinDepth());
    }
    void optimize(bool = false) {}
    void trackSpectrum(bool = true) {}
    void purgeMinDepth(long minimumCount, bool = false)
    {
        if (weak.ctx && minimumCount > 1) KMN_CHECK(weak.ctx, kmn_purge_min_depth(weak.ctx, (uint32_t)minimumCount));
    }
    void reset()
    {
        if (weak.ctx) { kmn_destroy(weak.ctx); weak.ctx = NULL; }
    }
    kmn_stats getStats()
    {
        kmn_stats s;
        KMN_CHECK(weak.ctx, kmn_get_stats(weak.ctx, &s));
        return s;
    }

    // KmerSpectrum::Histogram(zoomMax).toString(): counts <= zoomMax have their own bucket, larger counts fall into octaves
    // (zoomLogSkip = 7 for zoomMax 255 and 256); same columns and fixed/setprecision(3)   src/KmerSpectrum.h:909-1035
    std::string getHistogram(bool /*solidOnly*/ = false, unsigned int zoomMax = 256)
    {
        std::vector<uint64_t> hist(65536);
        std::vector<double> wsum(65536);
        KMN_CHECK(weak.ctx, kmn_histogram(weak.ctx, hist.data(), wsum.data()));
        const unsigned int zoomLogSkip = 7;
        struct Elem { unsigned long visits, visitedCount, cumulativeVisits; double visitedWeight; Elem() : visits(0), visitedCount(0), cumulativeVisits(0), visitedWeight(0) {} };
        std::vector<Elem> buckets((1u << 16) + 1 + zoomMax + 1);
        for (unsigned int c = 1; c < 65536; ++c) {
            if (!hist[c]) continue;
            unsigned int idx = c <= zoomMax ? c : (unsigned int)(std::log((double)c) / std::log(2.0) - zoomLogSkip + zoomMax);
            buckets[idx].visits += hist[c];
            buckets[idx].visitedCount += hist[c] * (unsigned long)c;
            buckets[idx].visitedWeight += wsum[c];
        }
        unsigned long count = 0; double totalCount = 0, totalWeightedCount = 0; unsigned int lastBucket = 0;
        for (int i = (int)buckets.size() - 1; i >= 0; --i) {
            buckets[i].cumulativeVisits = count += buckets[i].visits;
            if (buckets[i].visits > 0) {
                totalCount += buckets[i].visitedCount; totalWeightedCount += buckets[i].visitedWeight;
                if ((unsigned int)i > lastBucket) lastBucket = i;
            }
        }
        std::stringstream ss;
        ss << std::fixed << std::setprecision(3);
        ss << "Counts, Weights and Directions" << std::endl;
        ss << "Counts:\t" << count << "\t" << totalCount << "\t" << (totalCount / count) << "\t" << std::endl;
        ss << "Weights:\t" << count << "\t" << totalWeightedCount << "\t" << (totalWeightedCount / count) << "\t" << (totalWeightedCount / totalCount) << std::endl;
        ss << std::endl;
        ss << "Bucket\tCumulative\tUnique\t%Unique\tCount\t%Count\tWeight\tQualProb\t%Weight" << std::endl;
        for (unsigned int i = 1; i < lastBucket + 1; i++) {
            unsigned int label = i <= zoomMax ? i : (unsigned int)std::pow(2.0, (double)(i + zoomLogSkip - zoomMax));
            ss << label << "\t" << buckets[i].cumulativeVisits << "\t" << buckets[i].visits << "\t" << 100.0 * buckets[i].visits / count << "\t";
            ss << buckets[i].visitedCount << "\t" << 100.0 * buckets[i].visitedCount / totalCount << "\t\t";
            ss << buckets[i].visitedWeight << "\t" << buckets[i].visitedWeight / buckets[i].visitedCount << "\t";
            ss << 100.0 * buckets[i].visitedWeight / totalWeightedCount << "\t" << std::endl;
        }
        return ss.str();
    }
    void printHistograms(std::ostream &os, bool solidOnly = false) { os << getHistogram(solidOnly); }

    // MeraculousDistributedKmerSpectrum::dumpCounts / dumpGraphs (src/Meraculous.h:107-133): for every k-mer of this rank's
    // table with count >= minDepth one line for the k-mer and one for its reverse complement -- "<kmer>\t<count>", and
    // "<kmer>\t<6 left + 6 right extension counters A C G T N X> 0" (ExtensionTracking::toTextValues,
    // src/KmerTrackingData.h:153-230; the reverse complement swaps the sides and complements the bases).  The reference
    // writes in bucket order and its test sorts (test/runMeraculousTests.sh:39-75); here: ascending key order.
    void dumpCounts(std::ostream &os, int minDepth) { dump(os, minDepth, false); }
    void dumpGraphs(std::ostream &os, int minDepth) { dump(os, minDepth, true); }

private:
    void dump(std::ostream &os, int minDepth, bool graph)
    {
        if (!weak.ctx) LOG_THROW("dump on an empty KmerSpectrum");
        const unsigned int k = KmerBaseOptions::getOptions().getKmerSize(), kb = (k + 3) / 4;
        uint64_t n = 0;
        const uint32_t md = (uint32_t)std::max(1, minDepth);
        KMN_CHECK(weak.ctx, kmn_export(weak.ctx, md, NULL, NULL, NULL, NULL, NULL, 0, &n));
        if (n == 0) return;
        std::vector<uint8_t> keys((size_t)n * kb);
        std::vector<uint16_t> count(n);
        std::vector<uint32_t> ext(graph ? (size_t)n * 12 : 0);
        uint64_t got = 0;
        KMN_CHECK(weak.ctx, kmn_export(weak.ctx, md, keys.data(), count.data(), NULL, NULL, graph ? ext.data() : NULL, n, &got));
        std::vector<size_t> order(got);
        for (size_t i = 0; i < got; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return memcmp(&keys[a * kb], &keys[b * kb], kb) < 0; });
        static const char B[] = "ACGT";
        static const int COMP[6] = {3, 2, 1, 0, 4, 5};                   // A<->T, C<->G, N, X
        std::string km(k, 'A'), rc(k, 'A');
        for (size_t q = 0; q < got; ++q) {
            const size_t i = order[q];
            for (unsigned int b = 0; b < k; ++b) {
                const int code = (keys[i * kb + (b >> 2)] >> (6 - 2 * (b & 3))) & 3;
                km[b] = B[code]; rc[k - 1 - b] = B[3 - code];
            }
            if (!graph) { os << km << "\t" << count[i] << "\n" << rc << "\t" << count[i] << "\n"; continue; }
            const uint32_t *e = &ext[i * 12];
            uint32_t r[12];
            for (int c = 0; c < 6; ++c) { r[COMP[c]] = e[6 + c]; r[6 + COMP[c]] = e[c]; }
            os << km << "\t";
            for (int c = 0; c < 12; ++c) os << e[c] << " ";
            os << "0\n" << rc << "\t";
            for (int c = 0; c < 12; ++c) os << r[c] << " ";
            os << "0\n";
        }
    }
    unsigned long _rawKmers;
};

#endif
