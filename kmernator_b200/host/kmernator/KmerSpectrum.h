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

#include <fstream>

#include "../../../include/kmernator_b200.h"
#include "KmerHasher.h"
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
        _world = &world;
        _joinWorld();
    }

    explicit KmerSpectrum(unsigned long rawKmers = 0, int device = 0, unsigned int valueKind = KMN_VALUE_DIR) : _rawKmers(rawKmers)
    {
        kmn_default_opts(&_opts);
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
        _opts = o;
        _create();
    }
    ~KmerSpectrum() { reset(); }
    KmerSpectrum(const KmerSpectrum &) = delete;
    KmerSpectrum &operator=(const KmerSpectrum &) = delete;
    KmerSpectrum &operator=(KmerSpectrum &&o)                            // spectrum = KS(rawKmers)  apps/FilterReads.cpp:138
    {
        if (this != &o) { reset(); weak = o.weak; _rawKmers = o._rawKmers; _sizeTracker = o._sizeTracker; _trackSizes = o._trackSizes; _opts = o._opts; _world = o._world; _openBuilds = o._openBuilds; o.weak.ctx = NULL; }
        return *this;
    }
    KmerSpectrum(KmerSpectrum &&o) : weak(o.weak), _opts(o._opts), _world(o._world), _openBuilds(o._openBuilds), _rawKmers(o._rawKmers) { o.weak.ctx = NULL; }

    // count pass over the whole ReadSet in --batch-size batches           src/KmerSpectrum.h:2081-2115
    // finish = false: more read sets follow into the same spectrum (reference files, then subtract files); finishBuild() ends it
    void buildKmerSpectrum(const ReadSet &reads, bool finish = true)
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
        for (int attempt = 0;; ++attempt) {
            for (unsigned long b = 0; b < nBatches; ++b) {
                const ReadSet::ReadSetSizeType r0 = std::min<ReadSet::ReadSetSizeType>(n, b * batch), r1 = std::min<ReadSet::ReadSetSizeType>(n, r0 + batch);
                reads.concat(r0, r1, bases, quals, off, disc);
                KMN_CHECK(weak.ctx, kmn_count_batch(weak.ctx, (const uint8_t *)bases.data(), (const uint8_t *)quals.data(), off.data(), r1 - r0, disc.data()));
                if (_trackSizes) { for (ReadSet::ReadSetSizeType q = r0; q < r1; ++q) { unsigned long l = reads.getRead(q).getLength(); if (l >= _k()) _rawSubmitted += l - _k() + 1; } trackSpectrum(false); }
            }
            if (!finish) { _openBuilds++; return; }
            // The reference's buckets grow as they fill (KmerMapByKmerArrayPair::insert, src/Kmer.h:3095-3110); the table here has a
            // fixed capacity sized from --estimated-depth / --estimated-error-rate.  When that guess was too small (low coverage:
            // nearly every k-mer distinct) the build is repeated into a table four times as large -- the reads are still in
            // memory -- up to the provable bound of one slot pair per k-mer instance.  All ranks decide together.
            const int rc = kmn_count_finish(weak.ctx, 0);
            unsigned long full = rc == KMN_ERR_TABLE_FULL ? 1 : 0;
            if (_world && _world->size() > 1) full = _world->allMax(full);
            if (!full) { if (rc != 0) LOG_THROW("kmn_count_finish failed (" << rc << "): " << kmn_last_error(weak.ctx)); return; }
            const uint64_t bound = 2 * (uint64_t)_rawKmers + 8192;
            if (_openBuilds > 0 || attempt >= 6 || _opts.table_slots >= bound)
                LOG_THROW("count table overflow (" << _opts.table_slots << " slots): " << kmn_last_error(weak.ctx));
            LOG_VERBOSE(1, "count table of " << _opts.table_slots << " slots overflowed; rebuilding with " << std::min<uint64_t>(bound, _opts.table_slots * 4));
            _opts.table_slots = std::min<uint64_t>(bound, _opts.table_slots * 4);
            reset();
            _create();
            _joinWorld();
            _rawSubmitted = 0;
            _sizeTracker = SizeTracker();
        }
    }
    void finishBuild() { KMN_CHECK(weak.ctx, kmn_count_finish(weak.ctx, 0)); _openBuilds = 0; }
    // build + post-build purge (+ --save-kmer-mmap) (src/KmerSpectrum.h:1818-1831); more than one part is never needed in HBM
    void buildKmerSpectrumInParts(const ReadSet &reads, unsigned int /*numParts*/, const std::string &mmapPrefix = "")
    {
        buildKmerSpectrum(reads);
        purgeMinDepth(KmerSpectrumOptions::getOptions().getMinDepth());
        if (KmerSpectrumOptions::getOptions().getSaveKmerMmap() && !mmapPrefix.empty()) storeMmap(mmapPrefix);
    }
    void optimize(bool = false) {}
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

    // ---- persistence: --save-kmer-mmap / --load-kmer-mmap ---------------------------------------------------------
    // The reference writes one file per map (src/KmerSpectrum.h:476-518): "<name>" for the weak map and, when min-depth
    // <= 1, "<name>-singleton".  Layout of a map file (KmerMapByKmerArrayPair::store src/Kmer.h:3138-3155, KmerArrayPair::store
    // :960-969): u64 numBuckets, u64 bucketMask, u64 byte offset of every bucket; then per bucket u32 n, n keys of (k+3)/4
    // bytes, n values.  A key lives in bucket hash & mask (KmerHasher, :2329-2333), keys inside a bucket ascend in memcmp
    // order (:3076-3088).  Values: TrackingDataWithDirection = {u16 count, f32 weightedCount @4, u16 directionBias @8}, 12
    // bytes (src/KmerTrackingData.h:406-407,508); TrackingDataSingleton = one byte (u8)(weight * 254) + 1 (:623,641-661).
    // numBuckets = the next power of two >= estimatedRawKmers / estimated-depth / kmers-per-bucket + 1 (:2837,2224-2229).
    void storeMmap(const std::string &filename)
    {
        if (!weak.ctx) LOG_THROW("storeMmap on an empty KmerSpectrum");
        LOG_VERBOSE(1, "Saving weak kmer spectrum");
        std::vector<uint8_t> keys; std::vector<uint16_t> count, dir; std::vector<float> wsum;
        exportAll(keys, count, dir, wsum);
        writeMapFile(filename, keys, count, dir, wsum, false);
        if (KmerSpectrumOptions::getOptions().getMinDepth() <= 1) {
            LOG_VERBOSE(1, "Saving singleton kmer spectrum");
            writeMapFile(filename + "-singleton", keys, count, dir, wsum, true);
        }
    }
    void restoreMmap(const std::string &filename)
    {
        if (!weak.ctx) LOG_THROW("restoreMmap on an empty KmerSpectrum");
        LOG_VERBOSE(1, "Loading kmer spectrum from saved mmaps: " + filename);
        const bool a = readMapFile(filename, false), b = readMapFile(filename + "-singleton", true);
        if (!a && !b) LOG_THROW("Terribly sorry but there were no kmer spectrum mmap files at: " << filename << "*\n\tCan not continue");
    }
    // k-mers of the subtracting spectrum leave this one (src/KmerSpectrum.h:472-474,1582-1589)
    unsigned long subtractReference(KmerSpectrum &subtracting)
    {
        uint64_t entries = 0, instances = 0;
        KMN_CHECK(weak.ctx, kmn_subtract(weak.ctx, subtracting.weak.ctx, &entries, &instances));
        LOG_VERBOSE(1, "Subtracted " << entries << " kmers (" << instances << " instances) present in the subtracting spectrum");
        return (unsigned long)entries;
    }

    // SizeTracker (src/KmerSpectrum.h:812-900): (rawKmers, rawGoodKmers, uniqueKmers, singletonKmers) whenever rawKmers has
    // grown by 5% since the last sample -- here at batch granularity (the counters are exact after every batch)
    struct SizeTracker {
        struct Element { unsigned long rawKmers, rawGoodKmers, uniqueKmers, singletonKmers; };
        double nextToTrack;
        std::vector<Element> elements;
        SizeTracker() { reset(); }
        void reset() { nextToTrack = 128; elements.clear(); track(0, 0, 0, 0); }
        void track(unsigned long raw, unsigned long rawGood, unsigned long unique, unsigned long single, bool force = false)
        {
            if ((double)raw < nextToTrack && !force) return;
            Element e = {raw, rawGood, unique, single};
            elements.push_back(e);
            while ((double)raw >= nextToTrack) nextToTrack *= 1.05;
        }
        std::string toString() const
        {
            std::ostringstream ss;
            ss << "rawKmers\trawGoodKmers\tuniqueKmers\tsingletonKmers" << std::endl;
            for (size_t i = 0; i < elements.size(); ++i)
                ss << elements[i].rawKmers << "\t" << elements[i].rawGoodKmers << "\t" << elements[i].uniqueKmers << "\t" << elements[i].singletonKmers << std::endl;
            return ss.str();
        }
    };
    SizeTracker &getSizeTracker() { return _sizeTracker; }
    void trackSpectrum(bool force = true)
    {
        if (!weak.ctx || !_trackSizes) return;
        if (!force && (double)_rawSubmitted < _sizeTracker.nextToTrack) return;      // the counters cost a drain and a table scan
        kmn_stats st = getStats();
        _sizeTracker.track(st.raw_kmers, st.raw_good_kmers, st.unique_kmers, st.singleton_kmers, force);
    }
    void enableSizeTracking(bool on = true) { _trackSizes = on; }

    // MeraculousDistributedKmerSpectrum::dumpCounts / dumpGraphs (src/Meraculous.h:107-133): for every k-mer of this rank's
    // table with count >= minDepth one line for the k-mer and one for its reverse complement -- "<kmer>\t<count>", and
    // "<kmer>\t<6 left + 6 right extension counters A C G T N X> 0" (ExtensionTracking::toTextValues,
    // src/KmerTrackingData.h:153-230; the reverse complement swaps the sides and complements the bases).  The reference
    // writes in bucket order and its test sorts (test/runMeraculousTests.sh:39-75); here: ascending key order.
    void dumpCounts(std::ostream &os, int minDepth) { dump(os, minDepth, false); }
    void dumpGraphs(std::ostream &os, int minDepth) { dump(os, minDepth, true); }

private:
    void exportAll(std::vector<uint8_t> &keys, std::vector<uint16_t> &count, std::vector<uint16_t> &dir, std::vector<float> &wsum)
    {
        const unsigned int kb = (KmerBaseOptions::getOptions().getKmerSize() + 3) / 4;
        uint64_t n = 0, got = 0;
        KMN_CHECK(weak.ctx, kmn_export(weak.ctx, 1, NULL, NULL, NULL, NULL, NULL, 0, &n));
        keys.resize((size_t)n * kb); count.resize(n); dir.resize(n); wsum.resize(n);
        if (n) KMN_CHECK(weak.ctx, kmn_export(weak.ctx, 1, keys.data(), count.data(), dir.data(), wsum.data(), NULL, n, &got));
    }
    unsigned long numBuckets() const
    {
        unsigned long want = (unsigned long)((int)((double)_rawKmers / KmerSpectrumOptions::getOptions().getEstimatedDepth())) / (unsigned long)kmn_host::asLong("kmers-per-bucket") + 1;
        if (want > (1ul << 26)) want = 1ul << 26;
        unsigned long p = 1;
        while (p < want) p <<= 1;
        return p;
    }
    void writeMapFile(const std::string &fn, const std::vector<uint8_t> &keys, const std::vector<uint16_t> &count, const std::vector<uint16_t> &dir,
                      const std::vector<float> &wsum, bool singletons) const
    {
        const unsigned int kb = (KmerBaseOptions::getOptions().getKmerSize() + 3) / 4;
        const uint64_t nb = numBuckets(), mask = nb - 1;
        std::vector<std::pair<uint64_t, size_t> > order;                  // (bucket, entry)
        for (size_t i = 0; i < count.size(); ++i)
            if ((count[i] == 1) == singletons) order.push_back(std::make_pair(kmn_host::kmerHash(&keys[i * kb], kb) & mask, i));
        std::sort(order.begin(), order.end(), [&](const std::pair<uint64_t, size_t> &a, const std::pair<uint64_t, size_t> &b) {
            if (a.first != b.first) return a.first < b.first;
            return memcmp(&keys[a.second * kb], &keys[b.second * kb], kb) < 0;
        });
        const size_t vsz = singletons ? 1 : 12;
        std::vector<uint64_t> head(2 + nb);
        head[0] = nb; head[1] = mask;
        std::string body;
        size_t q = 0;
        for (uint64_t b = 0; b < nb; ++b) {
            head[2 + b] = 8 * (2 + nb) + body.size();
            size_t e = q;
            while (e < order.size() && order[e].first == b) ++e;
            const uint32_t n = (uint32_t)(e - q);
            body.append((const char *)&n, 4);
            for (size_t i = q; i < e; ++i) body.append((const char *)&keys[order[i].second * kb], kb);
            for (size_t i = q; i < e; ++i) {
                const size_t x = order[i].second;
                char v[12] = {0};
                if (singletons) v[0] = (char)((unsigned char)((double)wsum[x] * 254.0 + 0.5) + 1);      // the table reports (w8 - 1) / 254
                else { memcpy(v, &count[x], 2); memcpy(v + 4, &wsum[x], 4); memcpy(v + 8, &dir[x], 2); }
                body.append(v, vsz);
            }
            q = e;
        }
        std::ofstream of(fn.c_str(), std::ios::binary);
        if (!of.good()) LOG_THROW("Could not open " << fn << " for writing");
        of.write((const char *)head.data(), (std::streamsize)(head.size() * 8));
        of.write(body.data(), (std::streamsize)body.size());
    }
    bool readMapFile(const std::string &fn, bool singletons)
    {
        std::ifstream in(fn.c_str(), std::ios::binary);
        if (!in.good()) return false;
        std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        if (data.size() < 16) return false;
        const unsigned int kb = (KmerBaseOptions::getOptions().getKmerSize() + 3) / 4;
        uint64_t nb, mask;
        memcpy(&nb, &data[0], 8); memcpy(&mask, &data[8], 8);
        if (nb == 0 || mask != nb - 1 || data.size() < 8 * (2 + nb)) LOG_THROW("not a kmer spectrum map file: " << fn);
        const size_t vsz = singletons ? 1 : 12;
        std::vector<uint8_t> keys; std::vector<uint16_t> count, dir; std::vector<float> wsum;
        for (uint64_t b = 0; b < nb; ++b) {
            uint64_t off; memcpy(&off, &data[8 * (2 + b)], 8);
            if (off + 4 > data.size()) LOG_THROW("truncated kmer spectrum map file: " << fn);
            uint32_t n; memcpy(&n, &data[off], 4);
            const size_t k0 = off + 4, v0 = k0 + (size_t)n * kb;
            if (v0 + (size_t)n * vsz > data.size()) LOG_THROW("truncated kmer spectrum map file: " << fn);
            keys.insert(keys.end(), data.begin() + k0, data.begin() + v0);
            for (uint32_t i = 0; i < n; ++i) {
                const char *v = &data[v0 + (size_t)i * vsz];
                uint16_t c = 1, d = 0; float w;
                if (singletons) w = (float)(((int)(unsigned char)v[0] - 1 + 0.5) / 254.0);   // mid-quantum: re-quantises to the same byte
                else { memcpy(&c, v, 2); memcpy(&w, v + 4, 4); memcpy(&d, v + 8, 2); }
                count.push_back(c); dir.push_back(d); wsum.push_back(w);
            }
        }
        KMN_CHECK(weak.ctx, kmn_import(weak.ctx, keys.data(), count.data(), dir.data(), wsum.data(), NULL, count.size()));
        LOG_VERBOSE(1, "Loaded " << count.size() << " kmers from " << fn);
        return true;
    }
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
    static unsigned long _k() { return KmerBaseOptions::getOptions().getKmerSize(); }
    void _create()
    {
        kmn_ctx *c = NULL;
        int rc = kmn_create(&c, &_opts);
        if (rc != 0) LOG_THROW("kmn_create failed (" << rc << "): " << kmn_last_error(NULL));
        weak.ctx = c;
    }
    void _joinWorld()
    {
        if (!weak.ctx || !_world || _world->size() == 1) return;
        std::string id(128, '\0');
        if (_world->rank() == 0 && kmn_comm_unique_id(&id[0]) != 0) LOG_THROW("kmn_comm_unique_id failed");
        id = _world->broadcast(id);
        if (id.size() != 128) LOG_THROW("bad communicator id from rank 0");
        KMN_CHECK(weak.ctx, kmn_comm_init(weak.ctx, _world->rank(), _world->size(), id.data()));
    }
    kmn_opts _opts;
    World *_world = NULL;
    int _openBuilds = 0;                 // buildKmerSpectrum(reads, finish = false) calls since the last finishBuild
    unsigned long _rawKmers, _rawSubmitted = 0;
    SizeTracker _sizeTracker;
    bool _trackSizes = false;
};

#endif
