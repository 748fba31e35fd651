/*
 * kmn_oracle.c -- CPU restatement of Kmernator's k-mer spectrum hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (kmernator_b200/)
 * never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the reference's
 * own fixtures (test/phix.mercount.m21, test/phix.mergraph.m21.D2, the four
 * test/1000-Filtered*.fastq goldens, TwoBitSequenceTest/KmerTest literals) and against
 * oracle/_ref (the reference's own src/lookup3.h compiled where it lies) for the hash.
 *
 * Every function cites the reference file:line (under /root/reference) whose behaviour it restates.
 * Nothing here is copied from the reference: the algorithms are re-expressed on flat C arrays.
 * lookup3/lookup8 are Bob Jenkins' public-domain hashes, restated from their published definition.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_KB 64          /* max key bytes: k <= 256 */
/* Read::FASTQ_START_CHAR: the INTERNAL quality base.  Input qualities are rescaled to it on load
 * (src/ReadSet.h:171-209,694-708) and it equals --fastq-output-base-quality (33 default, or 64):
 * Read::setMinQualityScore(minQuality, outputBase) src/Sequence.cpp:543-547, src/Sequence.h:514.
 * All oracle entry points therefore take `start` and expect quals already expressed in that base. */

/* ------------------------------------------------------------------------------------------
 * A. quality -> probability table            src/Sequence.cpp:522-540, src/config.h:137-141
 * ---------------------------------------------------------------------------------------- */
void orc_quality_table(double *p, int start, int min_quality)
{
    int i;
    for (i = 0; i < 256; i++) p[i] = 0.0;
    for (i = start + min_quality; i < 103 && i < 256; i++)
        p[i] = 1.0 - pow(10.0, (start - i) / 10.0);
    for (i = 103; i < 256; i++) p[i] = 1.0;
}

/* ------------------------------------------------------------------------------------------
 * a1. TwoBitSequence::compressSequence       src/TwoBitSequence.cpp:114-144,242-269
 *     4 bases per byte, first base in bits 7..6; non-ACGT -> A plus a markup (char,pos); '.' -> 'N'
 * ---------------------------------------------------------------------------------------- */
static inline int base_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

uint32_t orc_compress_sequence(const char *bases, uint32_t len, uint8_t *out,
                               uint32_t *markup_pos, char *markup_chr)
{
    uint32_t nm = 0, i;
    uint32_t nbytes = (len + 3) / 4;
    if (out) memset(out, 0, nbytes);
    for (i = 0; i < len; i++) {
        int c = base_code(bases[i]);
        if (c < 0) {
            char m = bases[i] == '.' ? 'N' : bases[i];
            if (markup_pos) markup_pos[nm] = i;
            if (markup_chr) markup_chr[nm] = m;
            nm++;
            c = 0;
        }
        if (out) out[i >> 2] |= (uint8_t)(c << (6 - 2 * (i & 3)));
    }
    return nm;
}

void orc_uncompress_sequence(const uint8_t *in, uint32_t len, char *bases)
{
    uint32_t i;
    for (i = 0; i < len; i++) bases[i] = "ACGT"[(in[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    bases[len] = 0;
}

/* TwoBitSequence::firstMarkupNorX            src/TwoBitSequence.cpp:370-379  (0 = none, else pos+1) */
uint32_t orc_first_markup_n_or_x(const uint32_t *markup_pos, const char *markup_chr, uint32_t nm)
{
    uint32_t i;
    for (i = 0; i < nm; i++)
        if (markup_chr[i] == 'N' || markup_chr[i] == 'X') return markup_pos[i] + 1;
    return 0;
}

/* TwoBitSequence::shiftLeft                  src/TwoBitSequence.cpp:418-477
 * out[j] = bases shifted left by `shift` (0..3) bases; the byte after the last is consulted only
 * when has_extra is set, else zeros are shifted in. */
static void shift_left(const uint8_t *in, uint8_t *out, uint32_t nbytes, int shift, int has_extra)
{
    uint32_t j;
    if (shift == 0) { memmove(out, in, nbytes); return; }
    for (j = 0; j < nbytes; j++) {
        unsigned next = (j + 1 < nbytes || has_extra) ? in[j + 1] : 0;
        out[j] = (uint8_t)((in[j] << (2 * shift)) | (next >> (8 - 2 * shift)));
    }
}

/* TwoBitSequence::reverseComplement          src/TwoBitSequence.cpp:168-177,395-409 */
static uint8_t revcomp_byte(uint8_t b)
{
    /* reverse the four 2-bit fields and complement them */
    uint8_t r = (uint8_t)(((b & 0x03) << 6) | ((b & 0x0c) << 2) | ((b & 0x30) >> 2) | ((b & 0xc0) >> 6));
    return (uint8_t)~r;
}

void orc_reverse_complement(const uint8_t *in, uint8_t *out, uint32_t len)
{
    uint32_t nbytes = (len + 3) / 4, j;
    uint8_t tmp[ORC_MAX_KB + 1];
    uint8_t *t = nbytes <= ORC_MAX_KB ? tmp : (uint8_t *)malloc(nbytes + 1);
    for (j = 0; j < nbytes; j++) t[nbytes - 1 - j] = revcomp_byte(in[j]);
    if (len & 3) shift_left(t, out, nbytes, 4 - (int)(len & 3), 0);
    else memcpy(out, t, nbytes);
    if (t != tmp) free(t);
}

/* ------------------------------------------------------------------------------------------
 * a4. KmerHasher::getHash -> Lookup3::hashlittle2   src/Kmer.h:207-230, src/lookup3.h:120-164,470-644
 *     Bob Jenkins' lookup3 (public domain), byte-wise little-endian formulation (all of the
 *     header's aligned/unaligned variants compute the same function of the key bytes).
 * ---------------------------------------------------------------------------------------- */
#define ROT32(x, k) (((x) << (k)) | ((x) >> (32 - (k))))

static inline uint32_t le32(const uint8_t *p, size_t n)   /* up to 4 bytes, little endian, zero padded */
{
    uint32_t v = 0; size_t i;
    for (i = 0; i < n && i < 4; i++) v |= (uint32_t)p[i] << (8 * i);
    return v;
}

void orc_hashlittle2(const void *key, size_t length, uint32_t *pc, uint32_t *pb)
{
    const uint8_t *k = (const uint8_t *)key;
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + *pc;
    c += *pb;
    while (length > 12) {
        a += le32(k, 4); b += le32(k + 4, 4); c += le32(k + 8, 4);
        a -= c; a ^= ROT32(c, 4);  c += b;
        b -= a; b ^= ROT32(a, 6);  a += c;
        c -= b; c ^= ROT32(b, 8);  b += a;
        a -= c; a ^= ROT32(c, 16); c += b;
        b -= a; b ^= ROT32(a, 19); a += c;
        c -= b; c ^= ROT32(b, 4);  b += a;
        length -= 12; k += 12;
    }
    if (length == 0) { *pc = c; *pb = b; return; }
    a += le32(k, length);
    if (length > 4) b += le32(k + 4, length - 4);
    if (length > 8) c += le32(k + 8, length - 8);
    c ^= b; c -= ROT32(b, 14);
    a ^= c; a -= ROT32(c, 11);
    b ^= a; b -= ROT32(a, 25);
    c ^= b; c -= ROT32(b, 16);
    a ^= c; a -= ROT32(c, 4);
    b ^= a; b -= ROT32(a, 14);
    c ^= b; c -= ROT32(b, 24);
    *pc = c; *pb = b;
}

/* KmerHasher::getHash: seeds pc=0xDEADBEEF, pb=0, result = c | b<<32   src/Kmer.h:207-230 */
uint64_t orc_kmer_hash(const void *key, uint32_t nbytes)
{
    uint32_t pc = 0xDEADBEEFu, pb = 0;
    orc_hashlittle2(key, nbytes, &pc, &pb);
    return (uint64_t)pc | ((uint64_t)pb << 32);
}

/* Alternate (dead code in the reference, named by north_star): KmerHasher::toNumber folded into
 * Lookup8::hash2(&number, 1, 0xDEADBEEF)       src/Kmer.h:191-205,211-213, src/lookup8.h:52-66,170-201 */
static uint64_t to_number(const uint8_t *p, int len)
{
    uint64_t v = 0; int i;
    if (len >= 8) {
        for (i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
        if (len > 8) v += to_number(p + 8, len - 8);
    } else if (len >= 4) { for (i = 0; i < 4; i++) v |= (uint64_t)p[i] << (8 * i); }
    else if (len >= 2)   { v = (uint64_t)p[0] | ((uint64_t)p[1] << 8); }
    else v = p[0];
    return v;
}

uint64_t orc_kmer_hash_lookup8(const void *key, uint32_t nbytes)
{
    uint64_t a, b, c, n = to_number((const uint8_t *)key, (int)nbytes);
    a = b = 0xDEADBEEFull;
    c = 0x9e3779b97f4a7c13ull;
    c += (1ull << 3);
    a += n;
    a -= b; a -= c; a ^= (c >> 43);
    b -= c; b -= a; b ^= (a << 9);
    c -= a; c -= b; c ^= (b >> 8);
    a -= b; a -= c; a ^= (c >> 38);
    b -= c; b -= a; b ^= (a << 23);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 35);
    b -= c; b -= a; b ^= (a << 49);
    c -= a; c -= b; c ^= (b >> 11);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 18);
    c -= a; c -= b; c ^= (b >> 22);
    return c;
}

/* a5. owner / bucket arithmetic               src/Kmer.h:187-188,2284-2295,2329-2333 */
uint32_t orc_owner(uint64_t hash, uint32_t nranks) { return (uint32_t)(((hash >> 24) & 0x7ffffu) % nranks); }
uint64_t orc_bucket(uint64_t hash, uint64_t num_buckets_pow2) { return hash & (num_buckets_pow2 - 1); }

/* ------------------------------------------------------------------------------------------
 * a2+a3. per-read canonical k-mers, weights and extensions
 *        KmerArrayPair::build src/Kmer.h:1323-1375 ; Kmer::buildLeastComplement :356-364 ;
 *        KmerReadUtils::buildWeightedKmers src/KmerReadUtils.h:176-248
 * Outputs (nk = len-k+1 entries, or 0 when len<k):
 *   keys[nk*kb] canonical key bytes ; is_fwd[nk] ; weight[nk] = (float)|w| ; ext[nk*4] = {lbase,lqual,rbase,rqual}
 *   with bases coded A,C,G,T,N,X = 0..5 already swapped/complemented for rc-canonical k-mers.
 * ---------------------------------------------------------------------------------------- */
static inline uint8_t ext_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1;
    case 'G': case 'g': return 2; case 'T': case 't': return 3;
    case 'X': case 'x': return 5; default: return 4;
    }
}
static inline uint8_t ext_comp(uint8_t e) { return e < 4 ? (uint8_t)(3 - e) : e; }

uint32_t orc_read_kmers(const char *bases, const uint8_t *quals, uint32_t len, uint32_t k, int start,
                        const double *p, uint8_t *keys, uint8_t *is_fwd, float *weight, uint8_t *ext)
{
    uint32_t kb = (k + 3) / 4, nk, i, j, nm, mi = 0;
    uint32_t nbytes = (len + 3) / 4;
    uint8_t *packed; uint32_t *mpos; char *mchr, *fasta;
    uint8_t fwd[ORC_MAX_KB + 1], rc[ORC_MAX_KB + 1];
    double w = 0.0, change;
    uint8_t lbase = 5, lqual = 20;                       /* Extension('X', minQuality) */
    static const uint8_t lastmask[4] = {0xff, 0xc0, 0xf0, 0xfc};
    if (len < k || k == 0) return 0;
    nk = len - k + 1;
    packed = (uint8_t *)calloc(nbytes + 2, 1);
    mpos = (uint32_t *)malloc(sizeof(uint32_t) * (len + 1));
    mchr = (char *)malloc(len + 1);
    fasta = (char *)malloc(len + 1);
    nm = orc_compress_sequence(bases, len, packed, mpos, mchr);
    orc_uncompress_sequence(packed, len, fasta);         /* getFastaNoMarkup(): an N neighbour reads as A */
    for (i = 0; i < nk; i++) {
        /* k-mer bytes: shift by (i&3) bases starting at byte i/4, mask the last byte */
        uint32_t b0 = i >> 2;
        shift_left(packed + b0, fwd, kb, (int)(i & 3), b0 + kb < nbytes + 1);
        fwd[kb - 1] &= lastmask[k & 3];
        orc_reverse_complement(fwd, rc, k);
        is_fwd[i] = memcmp(fwd, rc, kb) <= 0;
        memcpy(keys + (size_t)i * kb, is_fwd[i] ? fwd : rc, kb);
        /* weight: double rolling product, re-seeded at i%1024==0 or w==0   KmerReadUtils.h:201-213 */
        if (i % 1024 == 0 || w == 0.0) {
            w = 1.0;
            for (j = 0; j < k; j++) w *= p[quals[i + j]];
        } else {
            change = p[quals[i + k - 1]] / p[quals[i - 1]];
            w *= change;
        }
        while (mi < nm && mpos[mi] < i) mi++;
        if (mi < nm && mpos[mi] < i + k) w = 0.0;          /* any markup inside the window   :214-219 */
        weight[i] = (float)w;                                /* setWeight stores float   KmerTrackingData.h:1019-1021 */
        if (ext) {
            uint8_t rbase = 5, rqual = 20;
            if (i + k < len) { rbase = ext_code(fasta[i + k]); rqual = (uint8_t)(quals[i + k] - start); }
            if (is_fwd[i]) { ext[4*i] = lbase; ext[4*i+1] = lqual; ext[4*i+2] = rbase; ext[4*i+3] = rqual; }
            else { ext[4*i] = ext_comp(rbase); ext[4*i+1] = rqual; ext[4*i+2] = ext_comp(lbase); ext[4*i+3] = lqual; }
            lbase = ext_code(fasta[i]); lqual = (uint8_t)(quals[i] - start);
        }
    }
    free(packed); free(mpos); free(mchr); free(fasta);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * a6. count table with the reference's value semantics
 *     KmerSpectrum::append src/KmerSpectrum.h:1578-1668 ; TrackingData::track src/KmerTrackingData.h:427-448 ;
 *     TrackingDataWithDirection::track :517-529 ; TrackingDataSingleton :613-686 ; ExtensionTracking :153-230
 * The three reference maps (solid/weak/singleton) are one open-addressing table here; count==1 entries
 * are "singletons" and keep the reference's singleton value (quantised weight, no direction).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t cap, n;        /* slots (pow2), used */
    uint32_t kbp;           /* padded key bytes */
    uint8_t *keys;
    uint8_t *used;
    uint16_t *count;
    uint16_t *dir;
    float *wsum;
    uint8_t *single_w;      /* TrackingDataSingleton::_weight */
    uint32_t *ext;          /* 12 per slot, optional */
} orc_shard;

typedef struct {
    uint32_t k, kb, nshards, track_ext, hash_kind;
    int start;
    float min_weight;
    double p[256];
    orc_shard *sh;
    uint64_t raw, raw_good, unique, singleton, purged_singletons;
    uint32_t min_depth_applied;
} orc_spectrum;

static uint64_t spec_hash(const orc_spectrum *s, const uint8_t *key)
{
    return s->hash_kind ? orc_kmer_hash_lookup8(key, s->kb) : orc_kmer_hash(key, s->kb);
}

static void shard_alloc(orc_shard *t, uint64_t cap, uint32_t kbp, int track_ext)
{
    t->cap = cap; t->n = 0; t->kbp = kbp;
    t->keys = (uint8_t *)calloc(cap, kbp);
    t->used = (uint8_t *)calloc(cap, 1);
    t->count = (uint16_t *)calloc(cap, 2);
    t->dir = (uint16_t *)calloc(cap, 2);
    t->wsum = (float *)calloc(cap, 4);
    t->single_w = (uint8_t *)calloc(cap, 1);
    t->ext = track_ext ? (uint32_t *)calloc(cap * 12, 4) : NULL;
}
static void shard_free(orc_shard *t)
{
    free(t->keys); free(t->used); free(t->count); free(t->dir); free(t->wsum); free(t->single_w); free(t->ext);
}

static uint64_t shard_find(const orc_shard *t, const uint8_t *key, uint32_t kb, uint64_t h, int *found)
{
    uint64_t i = (h * 0x9E3779B97F4A7C15ull) >> 20 & (t->cap - 1);
    while (t->used[i]) {
        if (memcmp(t->keys + i * t->kbp, key, kb) == 0) { *found = 1; return i; }
        i = (i + 1) & (t->cap - 1);
    }
    *found = 0;
    return i;
}

static void shard_grow(orc_spectrum *s, orc_shard *t)
{
    orc_shard nt; uint64_t i; int f;
    shard_alloc(&nt, t->cap * 2, t->kbp, s->track_ext);
    for (i = 0; i < t->cap; i++) if (t->used[i]) {
        const uint8_t *key = t->keys + i * t->kbp;
        uint64_t j = shard_find(&nt, key, s->kb, spec_hash(s, key), &f);
        nt.used[j] = 1; memcpy(nt.keys + j * nt.kbp, key, t->kbp);
        nt.count[j] = t->count[i]; nt.dir[j] = t->dir[i]; nt.wsum[j] = t->wsum[i]; nt.single_w[j] = t->single_w[i];
        if (nt.ext) memcpy(nt.ext + j * 12, t->ext + i * 12, 48);
        nt.n++;
    }
    shard_free(t);
    *t = nt;
}

orc_spectrum *orc_spectrum_new(uint32_t k, int start, int min_quality, float min_kmer_quality, int track_ext,
                               uint32_t nshards, uint64_t est_distinct, int hash_kind)
{
    orc_spectrum *s = (orc_spectrum *)calloc(1, sizeof(*s));
    uint32_t i; uint64_t cap = 1024;
    s->k = k; s->kb = (k + 3) / 4; s->nshards = nshards ? nshards : 1; s->track_ext = track_ext;
    s->min_weight = min_kmer_quality; s->hash_kind = hash_kind;
    s->start = start;
    orc_quality_table(s->p, start, min_quality);
    while (cap < 2 * est_distinct / s->nshards) cap <<= 1;
    s->sh = (orc_shard *)calloc(s->nshards, sizeof(orc_shard));
    for (i = 0; i < s->nshards; i++) shard_alloc(&s->sh[i], cap, (s->kb + 7) & ~7u, track_ext);
    return s;
}

void orc_spectrum_free(orc_spectrum *s)
{
    uint32_t i;
    for (i = 0; i < s->nshards; i++) shard_free(&s->sh[i]);
    free(s->sh); free(s);
}

static inline uint32_t shard_of(const orc_spectrum *s, uint64_t h)
{
    return (uint32_t)((h >> 40) % s->nshards);
}

/* one observation; returns 0 discarded, 1 tracked.  Stats are accumulated by the caller. */
typedef struct { uint64_t raw, raw_good, unique, singleton; } orc_delta;

static void track_one(orc_spectrum *s, orc_shard *t, const uint8_t *key, uint64_t h, float weight, int fwd,
                      const uint8_t *ext, orc_delta *d)
{
    uint64_t i; int found;
    d->raw++;
    if (!(weight > s->min_weight)) return;                 /* TrackingData::isDiscard   KmerTrackingData.h:354-364 */
    d->raw_good++;
    if ((t->n + 1) * 10 > t->cap * 6) shard_grow(s, t);
    i = shard_find(t, key, s->kb, h, &found);
    if (!found) {                                           /* new singleton   KmerSpectrum.h:1644-1655, TrackingDataSingleton::track :641-649 */
        t->used[i] = 1; memset(t->keys + i * t->kbp, 0, t->kbp); memcpy(t->keys + i * t->kbp, key, s->kb);
        t->n++;
        t->count[i] = 1; t->dir[i] = 0; t->wsum[i] = 0.0f;
        t->single_w[i] = (uint8_t)((unsigned char)((double)weight * 254.0) + 1);
        d->unique++; d->singleton++;
    } else {
        if (t->count[i] == 1 && t->single_w[i]) {           /* promote singleton to weak   :1630-1641 */
            t->wsum[i] = (float)((t->single_w[i] - 1) / 254.0);
            t->dir[i] = 0;
            t->single_w[i] = 0;
            d->singleton--;
        }
        if (t->count[i] < 65535) {                          /* saturating   TrackingData::track :427-448 */
            t->count[i]++;
            t->wsum[i] += weight;
            if (fwd) t->dir[i]++;                            /* TrackingDataWithDirection::track :517-529 */
        }
    }
    if (t->ext && ext) {                                    /* ExtensionTracking::trackExtension :195-201 */
        if (ext[1] >= 20 || ext[0] >= 4) t->ext[i * 12 + ext[0]]++;
        if (ext[3] >= 20 || ext[2] >= 4) t->ext[i * 12 + 6 + ext[2]]++;
    }
}

/* Serial build in read order: KmerSpectrum::_buildKmerSpectrumSerial src/KmerSpectrum.h:1914-1931.
 * Parallel build (nthreads>1): the T x T buffer matrix with a barrier per batch of
 * _buildKmerSpectrumParallel :1932-2074 -- reads interleaved over threads, each k-mer routed to the
 * thread owning its shard, then each thread drains its own column. */
typedef struct { uint8_t key[ORC_MAX_KB]; float w; uint8_t fwd; uint8_t ext[4]; uint64_t h; } orc_rec;
typedef struct { orc_rec *r; size_t n, cap; } orc_buf;

static void buf_push(orc_buf *b, const orc_rec *r)
{
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 4096; b->r = (orc_rec *)realloc(b->r, b->cap * sizeof(orc_rec)); }
    b->r[b->n++] = *r;
}

void orc_spectrum_add_reads(orc_spectrum *s, const char *bases, const uint8_t *quals, const uint64_t *off,
                            uint64_t n_reads, const uint8_t *discarded, int nthreads, uint64_t batch_reads)
{
    uint32_t T = s->nshards, kb = s->kb;
    uint64_t maxlen = 0, r;
    for (r = 0; r < n_reads; r++) if (off[r + 1] - off[r] > maxlen) maxlen = off[r + 1] - off[r];
    if (maxlen < s->k) return;
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads != T) nthreads = (int)T;         /* one thread per shard */
    if (batch_reads == 0) batch_reads = 100000;             /* --batch-size   src/Options.h:331 */
    if (T == 1) {
        uint8_t *keys = (uint8_t *)malloc((size_t)maxlen * kb);
        uint8_t *fw = (uint8_t *)malloc(maxlen), *ext = (uint8_t *)malloc((size_t)maxlen * 4);
        float *wt = (float *)malloc(sizeof(float) * maxlen);
        orc_delta d = {0, 0, 0, 0};
        for (r = 0; r < n_reads; r++) {
            uint32_t len = (uint32_t)(off[r + 1] - off[r]), nk, i;
            if (discarded && discarded[r]) continue;        /* KmerReadUtils.h:177-180 */
            nk = orc_read_kmers(bases + off[r], quals + off[r], len, s->k, s->start, s->p, keys, fw, wt, ext);
            for (i = 0; i < nk; i++) {
                const uint8_t *key = keys + (size_t)i * kb;
                track_one(s, &s->sh[0], key, spec_hash(s, key), wt[i], fw[i], ext + 4 * i, &d);
            }
        }
        s->raw += d.raw; s->raw_good += d.raw_good; s->unique += d.unique; s->singleton += d.singleton;
        free(keys); free(fw); free(ext); free(wt);
        return;
    }
    {
        orc_buf *bufs = (orc_buf *)calloc((size_t)T * T, sizeof(orc_buf));
        orc_delta *ds = (orc_delta *)calloc(T, sizeof(orc_delta));
        uint64_t b0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
        {
#ifdef _OPENMP
            uint32_t me = (uint32_t)omp_get_thread_num();
#else
            uint32_t me = 0;
#endif
            uint8_t *keys = (uint8_t *)malloc((size_t)maxlen * kb);
            uint8_t *fw = (uint8_t *)malloc(maxlen), *ext = (uint8_t *)malloc((size_t)maxlen * 4);
            float *wt = (float *)malloc(sizeof(float) * maxlen);
            uint64_t bb, rr; uint32_t src;
            for (bb = 0; bb < n_reads; bb += batch_reads) {
                uint64_t be = bb + batch_reads < n_reads ? bb + batch_reads : n_reads;
                for (src = 0; src < T; src++) bufs[(size_t)me * T + src].n = 0;
                for (rr = bb + me; rr < be; rr += T) {
                    uint32_t len = (uint32_t)(off[rr + 1] - off[rr]), nk, i;
                    if (discarded && discarded[rr]) continue;
                    nk = orc_read_kmers(bases + off[rr], quals + off[rr], len, s->k, s->start, s->p, keys, fw, wt, ext);
                    for (i = 0; i < nk; i++) {
                        orc_rec rec;
                        memcpy(rec.key, keys + (size_t)i * kb, kb);
                        rec.w = wt[i]; rec.fwd = fw[i]; memcpy(rec.ext, ext + 4 * i, 4);
                        rec.h = spec_hash(s, rec.key);
                        buf_push(&bufs[(size_t)me * T + shard_of(s, rec.h)], &rec);
                    }
                }
#ifdef _OPENMP
#pragma omp barrier
#endif
                for (src = 0; src < T; src++) {
                    orc_buf *b = &bufs[(size_t)src * T + me];
                    size_t q;
                    for (q = 0; q < b->n; q++)
                        track_one(s, &s->sh[me], b->r[q].key, b->r[q].h, b->r[q].w, b->r[q].fwd, b->r[q].ext, &ds[me]);
                }
#ifdef _OPENMP
#pragma omp barrier
#endif
            }
            free(keys); free(fw); free(ext); free(wt);
        }
        for (b0 = 0; b0 < T; b0++) {
            s->raw += ds[b0].raw; s->raw_good += ds[b0].raw_good; s->unique += ds[b0].unique; s->singleton += ds[b0].singleton;
        }
        for (b0 = 0; b0 < (uint64_t)T * T; b0++) free(bufs[b0].r);
        free(bufs); free(ds);
    }
}

/* a8. purgeMinDepth   src/KmerSpectrum.h:1805-1815 ; post-build src/DistributedFunctions.h:559-569
 * singletons dropped when minDepth>=2, weak entries with count<minDepth dropped when minDepth>2.
 * Entries are tombstoned by zeroing count (lookups then return 0, which is what getValue gives). */
uint64_t orc_spectrum_purge_min_depth(orc_spectrum *s, uint32_t min_depth)
{
    uint64_t purged = 0, i; uint32_t t;
    for (t = 0; t < s->nshards; t++) {
        orc_shard *sh = &s->sh[t];
        for (i = 0; i < sh->cap; i++) if (sh->used[i] && sh->count[i] && sh->count[i] < min_depth) {
            if (sh->count[i] == 1) s->singleton--;   /* purgedSingletons only grows in the periodic purge (:1790-1803), off by default */
            sh->count[i] = 0; sh->single_w[i] = 0; purged++;
        }
    }
    if (min_depth > s->min_depth_applied) s->min_depth_applied = min_depth;
    return purged;
}

void orc_spectrum_stats(const orc_spectrum *s, uint64_t *out5)
{
    out5[0] = s->raw; out5[1] = s->raw_good; out5[2] = s->unique; out5[3] = s->singleton; out5[4] = s->purged_singletons;
}

uint64_t orc_spectrum_size(const orc_spectrum *s)      /* entries with count>0 */
{
    uint64_t n = 0, i; uint32_t t;
    for (t = 0; t < s->nshards; t++) for (i = 0; i < s->sh[t].cap; i++) if (s->sh[t].used[i] && s->sh[t].count[i]) n++;
    return n;
}

/* ReadSelector::getValue   src/ReadSelector.h:924-931 */
uint32_t orc_spectrum_lookup(const orc_spectrum *s, const uint8_t *key)
{
    uint64_t h = spec_hash(s, key), i; int found;
    const orc_shard *t = &s->sh[shard_of(s, h)];
    i = shard_find(t, key, s->kb, h, &found);
    return found ? t->count[i] : 0;
}

typedef struct { const uint8_t *key; uint32_t shard; uint64_t slot; } orc_ref;
static uint32_t g_cmp_kb;
static int cmp_ref(const void *a, const void *b) { return memcmp(((const orc_ref *)a)->key, ((const orc_ref *)b)->key, g_cmp_kb); }

/* export all live entries sorted by key bytes (memcmp order = the in-bucket order of the reference,
 * src/Kmer.h:3076-3088).  ext may be NULL.  Singletons report dir=0, wsum=(w8-1)/254 (KmerTrackingData.h:654-661). */
uint64_t orc_spectrum_export(const orc_spectrum *s, uint8_t *keys, uint16_t *count, uint16_t *dir, float *wsum, uint32_t *ext)
{
    uint64_t n = orc_spectrum_size(s), j = 0, i; uint32_t t;
    orc_ref *refs = (orc_ref *)malloc(sizeof(orc_ref) * (n + 1));
    for (t = 0; t < s->nshards; t++) for (i = 0; i < s->sh[t].cap; i++) if (s->sh[t].used[i] && s->sh[t].count[i]) {
        refs[j].key = s->sh[t].keys + i * s->sh[t].kbp; refs[j].shard = t; refs[j].slot = i; j++;
    }
    g_cmp_kb = s->kb;
    qsort(refs, n, sizeof(orc_ref), cmp_ref);
    for (j = 0; j < n; j++) {
        const orc_shard *sh = &s->sh[refs[j].shard]; i = refs[j].slot;
        if (keys) memcpy(keys + j * s->kb, refs[j].key, s->kb);
        if (count) count[j] = sh->count[i];
        if (dir) dir[j] = sh->single_w[i] ? 0 : sh->dir[i];
        if (wsum) wsum[j] = sh->single_w[i] ? (float)((sh->single_w[i] - 1) / 254.0) : sh->wsum[i];
        if (ext && sh->ext) memcpy(ext + j * 12, sh->ext + i * 12, 48);
    }
    free(refs);
    return n;
}

/* a9. Histogram   src/KmerSpectrum.h:909-1057 (serial zoomMax 256), src/DistributedFunctions.h:575 (MPI zoomMax 255)
 * bins sized (1<<16)+1+zoomMax+1 ; returns the number of bins written. */
uint32_t orc_histogram_bin(uint32_t count, uint32_t zoom_max)
{
    double log_factor = log(2.0);
    uint32_t zoom_log_skip = (uint32_t)(log((double)zoom_max + 1.0) / log_factor - 1.0);
    return count <= zoom_max ? count : (uint32_t)(log((double)count) / log_factor - zoom_log_skip + zoom_max);
}
uint32_t orc_histogram_bucket_value(uint32_t idx, uint32_t zoom_max)
{
    double log_factor = log(2.0);
    uint32_t zoom_log_skip = (uint32_t)(log((double)zoom_max + 1.0) / log_factor - 1.0);
    return idx <= zoom_max ? idx : (uint32_t)pow(2.0, (double)(idx + zoom_log_skip - zoom_max));
}
uint32_t orc_spectrum_histogram(const orc_spectrum *s, uint32_t zoom_max, uint64_t *visits, uint64_t *visited_count,
                                double *visited_weight, uint32_t n_bins)
{
    uint32_t t; uint64_t i;
    memset(visits, 0, 8 * n_bins); memset(visited_count, 0, 8 * n_bins); memset(visited_weight, 0, 8 * n_bins);
    for (t = 0; t < s->nshards; t++) {
        const orc_shard *sh = &s->sh[t];
        for (i = 0; i < sh->cap; i++) if (sh->used[i] && sh->count[i]) {
            uint32_t b = orc_histogram_bin(sh->count[i], zoom_max);
            double w = sh->single_w[i] ? (sh->single_w[i] - 1) / 254.0 : (double)sh->wsum[i];
            if (b >= n_bins) continue;
            visits[b]++; visited_count[b] += sh->count[i]; visited_weight[b] += w;
        }
    }
    for (i = 0; i < s->purged_singletons; i++) {            /* addRecord(1, 1.0)   :1046-1048 */
        uint32_t b = orc_histogram_bin(1, zoom_max);
        visits[b]++; visited_count[b] += 1; visited_weight[b] += 1.0;
    }
    return n_bins;
}

/* ------------------------------------------------------------------------------------------
 * a10. ReadSelector::scoreAndTrimReads per read   src/ReadSelector.h:948-1014,1015-1062,1064-1076,1092-1209
 *   markup_length = firstMarkupNorX (0 none).  scoring: 0 SUM 1 MEDIAN 2 MIN 3 MAX 4 AVG  (enum KmerScoringType src/ReadSelector.h:240-247)
 *   Outputs follow ReadTrimType after setTrimHeaders: trim_off, trim_len (bases), score, was_trimmed.
 * ---------------------------------------------------------------------------------------- */
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

void orc_trim_values(const float *values, uint32_t n_values, uint32_t k, uint32_t markup_length, double min_score,
                     int scoring, uint32_t *trim_off, uint32_t *trim_len, float *score, uint8_t *was_trimmed)
{
    uint32_t num = n_values, i;
    uint32_t best_off = 0, best_len = 0, test_off = 0, test_len = 0;
    float best_score = 0, test_score = 0, sc = 0.0f;      /* ReadTrimType() zero-initialised */
    int trimmed;
    if (markup_length != 0) num = markup_length > k ? markup_length - k : 0;          /* _setNumKmers :1037-1047 */
    if (num > n_values) num = n_values;
    for (i = 0; i < num; i++) {                            /* trimReadByMinimumKmerScore :948-1014 */
        if (values[i] >= min_score) { test_len++; test_score += 1; }
        else {
            if (test_score > best_score) { best_score = test_score; best_off = test_off; best_len = test_len; }
            test_score = 0; test_off += test_len + 1; test_len = 0;
        }
    }
    if (test_score > best_score) { best_off = test_off; best_len = test_len; }
    trimmed = best_len < num;
    if (best_len > 0) {                                    /* scoreReadByScoringType :1092-1180 */
        const float *b = values + best_off;
        if (scoring == 1) {
            float *tmp = (float *)malloc(sizeof(float) * best_len);
            memcpy(tmp, b, sizeof(float) * best_len);
            qsort(tmp, best_len, sizeof(float), cmp_float);
            sc = tmp[best_len / 2];
            free(tmp);
        } else if (scoring == 4) {
            double sum = 0.0; for (i = 0; i < best_len; i++) sum += b[i];
            sc = (float)(sum / (int)best_len);
        } else if (scoring == 2 || scoring == 3) {
            sc = b[0];
            for (i = 1; i < best_len; i++) sc = scoring == 3 ? (sc > b[i] ? sc : b[i]) : (sc < b[i] ? sc : b[i]);
        } else {
            sc = 0.0f;                                      /* KS_SUM never assigns trim.score   :1151-1162 */
        }
        best_len += k - 1;                                  /* setTrimHeaders :1015-1024 */
    } else {
        best_off = 0; sc = -1.0f;
    }
    *trim_off = best_off; *trim_len = best_len; *score = sc; *was_trimmed = (uint8_t)trimmed;
}

void orc_trim_read(const orc_spectrum *s, const char *bases, uint32_t len, double min_score, int scoring,
                   uint32_t *trim_off, uint32_t *trim_len, float *score, uint8_t *was_trimmed)
{
    uint32_t k = s->k, kb = s->kb, nk = len >= k ? len - k + 1 : 0, i, nm, ml;
    uint8_t *keys = (uint8_t *)malloc((size_t)(nk + 1) * kb), *fw = (uint8_t *)malloc(nk + 1);
    float *wt = (float *)malloc(sizeof(float) * (nk + 1)), *vals = (float *)malloc(sizeof(float) * (nk + 1));
    uint8_t *q = (uint8_t *)malloc(len + 1);
    uint32_t *mpos = (uint32_t *)malloc(sizeof(uint32_t) * (len + 1)); char *mchr = (char *)malloc(len + 1);
    memset(q, 'I', len);
    nm = orc_compress_sequence(bases, len, NULL, mpos, mchr);
    ml = orc_first_markup_n_or_x(mpos, mchr, nm);
    orc_read_kmers(bases, q, len, k, s->start, s->p, keys, fw, wt, NULL);
    for (i = 0; i < nk; i++) {                             /* setKmerValues :1064-1076 */
        float v = (float)orc_spectrum_lookup(s, keys + (size_t)i * kb);
        vals[i] = v >= min_score ? v : 0.0f;
    }
    orc_trim_values(vals, nk, k, ml, min_score, scoring, trim_off, trim_len, score, was_trimmed);
    free(keys); free(fw); free(wt); free(vals); free(q); free(mpos); free(mchr);
}

void orc_trim_reads(const orc_spectrum *s, const char *bases, const uint64_t *off, uint64_t n_reads, const uint8_t *discarded,
                    double min_score, int scoring, uint32_t *trim_off, uint32_t *trim_len, float *score, uint8_t *was_trimmed,
                    int nthreads)
{
    int64_t r;
#ifdef _OPENMP
#pragma omp parallel for schedule(guided) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (r = 0; r < (int64_t)n_reads; r++) {
        if (discarded && discarded[r]) { trim_off[r] = 0; trim_len[r] = 0; score[r] = 0.0f; was_trimmed[r] = 0; continue; }
        orc_trim_read(s, bases + off[r], (uint32_t)(off[r + 1] - off[r]), min_score, scoring,
                      &trim_off[r], &trim_len[r], &score[r], &was_trimmed[r]);
    }
}

/* a11. ReadSelectorUtil::passesLength   src/ReadSelector.h:209-228  (fp32 arithmetic) */
int orc_passes_length(float length, uint32_t read_length, float minimum_length)
{
    if (length <= 1.0f) return 0;
    if (minimum_length <= 1.0f) return (float)read_length * minimum_length <= length;
    return minimum_length <= length;
}

/* KmerSpectrum::estimateRawKmers   src/KmerSpectrum.h:573-584 */
uint64_t orc_estimate_raw_kmers(uint64_t n_reads, uint64_t base_count, uint32_t k)
{
    uint64_t avg, per;
    if (base_count == 0 || n_reads == 0) return 128;
    avg = base_count / n_reads;
    per = avg - k + 1;
    if (k > avg) per = 1;
    return per * n_reads;
}
