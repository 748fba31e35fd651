/*
 * kmernator_b200.h -- C ABI of the B200-native k-mer spectrum hot path.
 *
 * The reference (JGI-Bioinformatics/Kmernator) has no FFI boundary: the path is C++ templates
 * instantiated inside each app.  This header is the boundary a maintainer would bind instead of
 * those templates; each entry point names the reference call it replaces (file:line under the
 * reference tree).  Plain pointers and sizes only; every pointer argument may be a HOST or a
 * DEVICE pointer unless stated otherwise (the library inspects it with cudaPointerGetAttributes).
 *
 * All functions return 0 on success or a negative kmn_status; kmn_last_error() gives the message.
 * One host thread per context.  One context per GPU (one process per GPU under torchrun/mpirun).
 *
 * Key format (everywhere in this ABI): the reference's TwoBitSequence bytes, i.e. (k+3)/4 bytes,
 * 4 bases per byte, first base in bits 7..6, A=0 C=1 G=2 T=3, unused low bits zero
 * (src/TwoBitSequence.cpp:242-269, src/Kmer.h:1347-1357).
 */
#ifndef KMERNATOR_B200_H
#define KMERNATOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmn_ctx kmn_ctx;

typedef enum {
    KMN_OK = 0,
    KMN_ERR_INVALID = -1,     /* bad argument / unsupported option                           */
    KMN_ERR_CUDA = -2,        /* CUDA runtime error (message has the call site)              */
    KMN_ERR_NOMEM = -3,       /* device allocation failed                                    */
    KMN_ERR_TABLE_FULL = -4,  /* count table overflowed and could not grow                   */
    KMN_ERR_COMM = -5,        /* NCCL error / communicator not initialised                   */
    KMN_ERR_STATE = -6        /* call order violated (e.g. lookup before count_finish)       */
} kmn_status;

/* KmerHasher (src/Kmer.h:207-230): the live hash is lookup3 hashlittle2; lookup8 hash2 is the
 * commented-out predecessor (src/Kmer.h:211-213), selectable because north_star names it.        */
typedef enum { KMN_HASH_LOOKUP3_HASHLITTLE2 = 0, KMN_HASH_LOOKUP8_HASH2 = 1 } kmn_hash_kind;

/* Table value kinds (src/KmerTrackingData.h): DIR = TrackingDataWithDirection (FilterReads,
 * apps/FilterReads.cpp:52); DIR_EXT = ExtensionTrackingData (MeraculousCounter, src/Meraculous.h:79-80).
 * KMN_VALUE_WEIGHTS additionally accumulates weightedCount (fp32 sum; order-dependent in the reference too). */
enum { KMN_VALUE_DIR = 0, KMN_VALUE_DIR_EXT = 1, KMN_VALUE_WEIGHTS = 2 };

/* ReadSelector::KmerScoringType (src/ReadSelector.h:240-247), same numeric order */
typedef enum { KMN_SCORE_SUM = 0, KMN_SCORE_MEDIAN = 1, KMN_SCORE_MIN = 2, KMN_SCORE_MAX = 3, KMN_SCORE_AVG = 4 } kmn_scoring;

typedef struct {
    uint32_t struct_size;        /* = sizeof(kmn_opts); for ABI evolution                                           */
    uint32_t kmer_size;          /* positional kmer-size (apps/FilterReads.cpp:68), 1..128                            */
    uint32_t fastq_start_char;   /* Read::FASTQ_START_CHAR = internal quality base = --fastq-output-base-quality
                                    (33|64; src/Sequence.cpp:543-547).  quals passed in must be in this base.        */
    uint32_t min_quality_score;  /* --min-quality-score, default 3 (src/Options.h:329)                               */
    float    min_kmer_quality;   /* --min-kmer-quality, default 0.10 (src/KmerSpectrum.h:92): TrackingData::minimumWeight */
    uint32_t min_depth;          /* --min-depth, default 2 (src/KmerSpectrum.h:92,111)                               */
    uint32_t hash_kind;          /* kmn_hash_kind                                                                    */
    uint32_t value_kind;         /* KMN_VALUE_* flags                                                                */
    uint64_t est_raw_kmers;      /* KmerSpectrum ctor argument (src/KmerSpectrum.h:414-421); 0 = use table_slots     */
    uint64_t table_slots;        /* explicit table capacity (slots); 0 = derive from est_raw_kmers                   */
    uint64_t stage_keys;         /* capacity of the partitioned key staging area (k-mer instances); 0 = auto         */
    uint32_t slice_bytes;        /* target bytes of one table group (the L2-resident unit of the staging); 0 = 64 MiB */
    uint32_t device;             /* CUDA device ordinal                                                              */
    uint32_t ignore_quality;     /* --ignore-quality: every base weight 1                                            */
    uint32_t reserved[7];
} kmn_opts;

typedef struct {                 /* KmerSpectrum counters (src/KmerSpectrum.h:396-402,1590-1650)                     */
    uint64_t raw_kmers;          /* rawKmers: every k-mer instance presented                                         */
    uint64_t raw_good_kmers;     /* rawGoodKmers: instances that passed the weight test                              */
    uint64_t unique_kmers;       /* uniqueKmers: distinct k-mers ever inserted                                       */
    uint64_t singleton_kmers;    /* distinct k-mers whose count is exactly 1 right now                               */
    uint64_t discarded_kmers;    /* TrackingData::discarded                                                          */
    uint64_t table_slots;        /* capacity                                                                         */
    uint64_t table_partitions;   /* L2-sized partitions                                                              */
    uint64_t direct_inserts;     /* instances that bypassed staging (stage overflow)                                  */
} kmn_stats;

void        kmn_default_opts(kmn_opts *o);
const char *kmn_last_error(const kmn_ctx *ctx);   /* ctx may be NULL: error of the last failed kmn_create */
const char *kmn_version(void);

/* KS spectrum(rawKmers)                                  src/KmerSpectrum.h:414-421, DistributedFunctions.h:126-131 */
int  kmn_create(kmn_ctx **out, const kmn_opts *opts);
void kmn_destroy(kmn_ctx *ctx);
/* KmerSpectrum::reset()                                  src/KmerSpectrum.h:533 */
int  kmn_reset(kmn_ctx *ctx);

/* Multi-GPU: one context per rank.  id = 128-byte ncclUniqueId produced by rank 0 with
 * kmn_comm_unique_id and distributed by the caller (torch.distributed / MPI / file).
 * Replaces ScopedMPIComm + MPIAllToAllMessageBuffer    src/MPIUtils.h:256-391, src/MPIBuffer.h:412-1073 */
int  kmn_comm_unique_id(void *id128);
int  kmn_comm_init(kmn_ctx *ctx, int rank, int nranks, const void *id128);

/* spectrum.buildKmerSpectrum(reads): one batch of reads   src/KmerSpectrum.h:1914-2074,2081-2115;
 * DistributedKmerSpectrum::_buildKmerSpectrumMPI          src/DistributedFunctions.h:340-458
 * bases/quals: concatenated ASCII; read_off: n_reads+1 offsets; discarded: nullable per-read flag
 * (Read::isDiscarded, src/KmerReadUtils.h:177-180).  Asynchronous: returns once the batch is staged. */
int  kmn_count_batch(kmn_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *read_off,
                     uint64_t n_reads, const uint8_t *discarded);
/* The same batch in the reference's in-memory form (Read::_data, src/Sequence.h:372-380): per read (len+3)/4
 * TwoBitSequence bytes, 4 bases per byte, first base in bits 7..6 (src/TwoBitSequence.cpp:242-269), the reads' byte
 * strings concatenated (packed_off: n_reads+1 byte offsets), plus the markups of the non-ACGT bases as (position in the
 * concatenated batch, character) pairs -- a packed non-ACGT base reads as 'A' until its markup is applied.  A quarter of
 * the bases' bytes cross PCIe; the library unpacks on the device.  packed_off / read_off: HOST arrays.                  */
int  kmn_count_batch_2na(kmn_ctx *ctx, const uint8_t *packed, const uint64_t *packed_off, const uint8_t *quals,
                         const uint64_t *read_off, uint64_t n_reads, const uint8_t *discarded,
                         const uint64_t *markup_pos, const uint8_t *markup_chr, uint64_t n_markups);
/* end of buildKmerSpectrum: drain staging (+ exchange), then the post-build purge per min_depth
 * (src/KmerSpectrum.h:1825, src/DistributedFunctions.h:559-569) when apply_purge != 0                     */
int  kmn_count_finish(kmn_ctx *ctx, int apply_purge);

int  kmn_get_stats(kmn_ctx *ctx, kmn_stats *out);

/* spectrum.purgeMinDepth(minDepth)                        src/KmerSpectrum.h:1805-1815 */
int  kmn_purge_min_depth(kmn_ctx *ctx, uint32_t min_depth);

/* Exact count histogram: hist[c] = number of distinct k-mers with count c (c in 0..65535), wsum[c] = sum of
 * their weightedCount (zeros unless KMN_VALUE_WEIGHTS).  HOST arrays of 65536 entries.  The reference's zoomed
 * bins (KmerSpectrum::Histogram src/KmerSpectrum.h:909-1057) are a host-side fold of this array.
 * With a communicator the result is all-reduced (MPIHistogram::reduce src/DistributedFunctions.h:495-535). */
int  kmn_histogram(kmn_ctx *ctx, uint64_t *hist65536, double *wsum65536);

/* getElementIfExists(kmer).value().getCount()              src/ReadSelector.h:924-931
 * keys: n * key_bytes reference-format bytes (canonical); counts: n x u16 (0 = absent or purged).
 * With a communicator the call is collective (every rank calls it, n may be 0): keys travel to their owners and the
 * counts come back, as DistributedReadSelector::_batchKmerLookup does (src/DistributedFunctions.h:877-902).  */
int  kmn_lookup(kmn_ctx *ctx, const uint8_t *keys, uint64_t n, uint16_t *counts);

/* ReadSelector::scoreAndTrimReads(minDepth)               src/ReadSelector.h:1182-1209 (+:948-1180),
 * DistributedReadSelector::scoreAndTrimReads               src/DistributedFunctions.h:903-1045
 * outputs (n_reads each; host or device): ReadTrimType fields after setTrimHeaders.                         */
int  kmn_trim_batch(kmn_ctx *ctx, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
                    const uint8_t *discarded, uint32_t min_depth, int scoring,
                    uint32_t *trim_off, uint32_t *trim_len, float *score, uint8_t *was_trimmed);

/* table dump for dumpCounts/dumpGraphs and --save-kmer-mmap   src/Meraculous.h:107-133, src/Kmer.h:3124-3191
 * HOST outputs; any of count/dir/wsum/ext may be NULL.  ext = 12 u32 per k-mer (L A,C,G,T,N,X then R ...).
 * Entries with count < min_count are skipped.  *n_out = entries written (call with cap=0 to size).          */
int  kmn_export(kmn_ctx *ctx, uint32_t min_count, uint8_t *keys, uint16_t *count, uint16_t *dir, float *wsum,
                uint32_t *ext, uint64_t cap, uint64_t *n_out);

/* restore of a saved spectrum: KmerMapByKmerArrayPair(const void *mmap) / KmerSpectrum::restoreMmap
 * (src/Kmer.h:3124-3191, src/KmerSpectrum.h:489-518).  n entries in the format kmn_export writes (dir / wsum / ext
 * may be NULL); an entry whose key is already in the table is merged.  HOST or DEVICE arrays.                    */
int  kmn_import(kmn_ctx *ctx, const uint8_t *keys, const uint16_t *count, const uint16_t *dir, const float *wsum,
                const uint32_t *ext, uint64_t n);

/* spectrum.subtractReference(other): k-mers present in `other` (count >= 1 after its own purge) are removed from this
 * spectrum (src/KmerSpectrum.h:472-474,1582-1589; apps/FilterReads-P.cpp:281-308).  Both contexts on one device; with
 * a communicator both are sharded by the same owner rule, so the call is local.                                   */
int  kmn_subtract(kmn_ctx *ctx, kmn_ctx *other, uint64_t *removed_entries, uint64_t *removed_instances);

/* per-k-mer records of one batch, for parity tests of steps (1)-(3): canonical key bytes, strand, fp32 weight,
 * KmerHasher hash.  HOST outputs sized sum(max(0,len-k+1)).                      src/KmerReadUtils.h:176-248 */
int  kmn_debug_kmers(kmn_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *read_off,
                     uint64_t n_reads, uint8_t *keys, uint8_t *is_fwd, float *weight, uint64_t *hash, uint64_t *n_out);

/* the owner rank the device assigns to canonical keys for `nranks` ranks: ((KmerHasher hash >> 24) & 0x7ffff) % nranks,
 * getDistributedThreadId src/Kmer.h:2284-2295 (parity test of a5 on one GPU).  HOST or DEVICE arrays.               */
int  kmn_debug_owner(kmn_ctx *ctx, const uint8_t *keys, uint64_t n, uint32_t nranks, uint32_t *owner);

int  kmn_sync(kmn_ctx *ctx);

/* Per-kernel device timing (CUDA events on the context's stream around every launch of each kernel class).
 * Off by default (events cost a few microseconds per launch).  kmn_profile_read synchronises, returns the sums
 * accumulated since the last read and clears them.                                                             */
enum { KMN_PROF_PARSE = 0, KMN_PROF_INSERT = 1, KMN_PROF_ROUTE = 2, KMN_PROF_LOOKUP = 3, KMN_PROF_TRIM = 4, KMN_PROF_SCAN = 5,
       KMN_PROF_WEIGHT = 6, KMN_PROF_SUBPART = 7, KMN_PROF_KINDS = 8 };
typedef struct {
    double   ms[KMN_PROF_KINDS];        /* summed device time per kernel class                                  */
    uint64_t launches[KMN_PROF_KINDS];
    uint64_t units[KMN_PROF_KINDS];     /* k-mer instances (parse: upper bound = bases; insert: staged records)  */
} kmn_profile;
int  kmn_profile_enable(kmn_ctx *ctx, int on);
int  kmn_profile_read(kmn_ctx *ctx, kmn_profile *out);
/* the cudaStream_t all kernels of this context are launched on (for CUDA-event timing by the caller)       */
void *kmn_stream(kmn_ctx *ctx);
/* number of kernels launched by this context so far                                                        */
uint64_t kmn_launch_count(const kmn_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
