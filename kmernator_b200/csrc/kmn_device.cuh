// kmn_device.cuh -- device primitives of the k-mer spectrum path (sm_100a).
//
// Key representation: W = ceil(k/32) 64-bit words, base 0 in the top two bits of word 0, unused low
// bits of the last word zero.  The big-endian bytes of these words are exactly the reference's
// TwoBitSequence bytes (src/TwoBitSequence.cpp:242-269, src/Kmer.h:1347-1357), so memcmp order
// (src/Kmer.h:311-313) is lexicographic word order and the hash can be computed from the words.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

namespace kmn {

// ------------------------------------------------------------------------------------------------
// value word of a table slot:  [63] LOCK  [62] READY  [61..32] directionBias  [31..0] count
// (count/dir are kept wider than the reference's uint16 and clamped to 65535 on read: the reference's
//  saturating count, src/KmerTrackingData.h:306,427-448)
// ------------------------------------------------------------------------------------------------
static constexpr u64 VAL_LOCK = 1ull << 63;
static constexpr u64 VAL_READY = 1ull << 62;
static constexpr u64 VAL_COUNT_MASK = 0xffffffffull;
static constexpr u32 MAX_COUNT = 65535u;

template <int W> struct Slot { u64 val; u64 k[W]; };   // W==1: 16 B, stored key is ~key (0 = empty)

// The table is cut into n_parts SLICES of part_slots slots (<= SLICE_SLOTS, so one slice fits in shared memory); a key's
// slice and home slot come from place_hash and probing wraps inside the slice.  2^group_shift consecutive slices form one
// GROUP (~slice_bytes, L2-sized): the unit the level-1 staging is partitioned by.
struct TableView {
    void *slots;          // Slot<W>[n_parts * part_slots]
    float *wsum;          // optional, per slot
    u32 *ext;             // optional, 12 per slot
    u64 part_slots;       // slots per slice
    u32 n_parts;          // slices
    u32 group_shift;      // log2(slices per group)
    __device__ __host__ __forceinline__ u32 n_groups() const { return (n_parts + (1u << group_shift) - 1u) >> group_shift; }
};
static constexpr u32 SLICE_SLOTS = 4096;   // 64 KB of 16-byte slots

struct StageView {        // level-1 staging area of k-mer records, partitioned by (owner rank,) table GROUP (phase 1 -> phase 2)
    u64 *recs;            // [n_owners][n_cta][n_parts][sub_cap][RW]: every phase-1 CTA owns a private sub-region of every
                          // (owner, group), so a flush needs no global atomic (its position follows from a CTA-local counter)
    u32 *count;           // [n_owners][n_parts][n_cta] records written (may exceed sub_cap: the excess was inserted directly)
    u32 sub_cap;          // records per sub-region
    u32 n_cta;            // phase-1 grid size
    u32 n_parts;          // table groups
    u32 pad;
    u32 n_owners;         // 1, or the number of ranks when phase 1 bins by owner as well (multi-GPU push path)
    u32 me;               // this rank (index of the local owner when n_owners > 1)
    u64 *ovf_recs;        // [n_owners][ovf_cap][RW] records of OTHER owners that found their sub-region full (skewed input:
    u32 *ovf_count;       // [n_owners]              one read with very many k-mers, few distinct bins); shipped ungrouped
    u32 ovf_cap;          // 0: no overflow list (a full remote sub-region is then an error)
    u32 pad2;
    __device__ __host__ __forceinline__ u32 local_owner() const { return n_owners > 1 ? me : 0u; }
    __device__ __host__ __forceinline__ size_t sub_index(u32 owner, u32 part, u32 cta) const
    {
        return ((size_t)owner * n_cta + cta) * n_parts + part;     // a CTA's write frontier stays within few pages
    }
    __device__ __host__ __forceinline__ size_t cnt_index(u32 owner, u32 part, u32 cta) const
    {
        return ((size_t)owner * n_parts + part) * n_cta + cta;
    }
};

struct Counters {         // device-side statistics (src/KmerSpectrum.h:1590-1650)
    u64 raw, raw_good, unique, direct, table_full, probe_steps;
};

// ------------------------------------------------------------------------------------------------
// a4. KmerHasher::getHash = lookup3 hashlittle2(key bytes, kb, pc=0xDEADBEEF, pb=0) -> c | b<<32
//     src/Kmer.h:207-230, src/lookup3.h:120-164,470-644 (Bob Jenkins, public domain)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 rot32(u32 x, int k) { return __funnelshift_l(x, x, k); }
__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

template <int W>
__device__ __forceinline__ u64 hash_lookup3(const u64 (&w)[W], int kb)
{
    u32 kk[2 * W + 3];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        kk[2 * i] = bswap32((u32)(w[i] >> 32));
        kk[2 * i + 1] = bswap32((u32)w[i]);
    }
    kk[2 * W] = kk[2 * W + 1] = kk[2 * W + 2] = 0;
    u32 a, b, c;
    a = b = c = 0xdeadbeefu + (u32)kb + 0xDEADBEEFu;
    int len = kb;
#pragma unroll
    for (int blk = 0; blk < (8 * W - 1) / 12; ++blk) {
        if (len > 12) {
            a += kk[3 * blk]; b += kk[3 * blk + 1]; c += kk[3 * blk + 2];
            a -= c; a ^= rot32(c, 4);  c += b;
            b -= a; b ^= rot32(a, 6);  a += c;
            c -= b; c ^= rot32(b, 8);  b += a;
            a -= c; a ^= rot32(c, 16); c += b;
            b -= a; b ^= rot32(a, 19); a += c;
            c -= b; c ^= rot32(b, 4);  b += a;
            len -= 12;
        }
    }
    // tail: 1..12 bytes starting at 32-bit word 3*((kb-1)/12); bytes past kb are zero by construction
    int t = 3 * ((kb - 1) / 12);
    u32 ta = 0, tb = 0, tc = 0;
#pragma unroll
    for (int i = 0; i < 2 * W; ++i) {
        if (i == t) ta = kk[i];
        if (i == t + 1) tb = kk[i];
        if (i == t + 2) tc = kk[i];
    }
    a += ta; b += tb; c += tc;
    c ^= b; c -= rot32(b, 14);
    a ^= c; a -= rot32(c, 11);
    b ^= a; b -= rot32(a, 25);
    c ^= b; c -= rot32(b, 16);
    a ^= c; a -= rot32(c, 4);
    b ^= a; b -= rot32(a, 14);
    c ^= b; c -= rot32(b, 24);
    return (u64)c | ((u64)b << 32);
}

// alternate: KmerHasher::toNumber folded into Lookup8::hash2(&n,1,0xDEADBEEF)   src/Kmer.h:191-205,211-213,
// src/lookup8.h:52-66,170-201 (dead code in the reference; selectable because north_star names it)
__device__ __forceinline__ u64 bswap64(u64 x)
{
    return ((u64)bswap32((u32)x) << 32) | bswap32((u32)(x >> 32));
}
template <int W>
__device__ __forceinline__ u64 hash_lookup8(const u64 (&w)[W], int kb)
{
    u64 n = 0;
    int len = kb;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        u64 le = bswap64(w[i]);
        if (len >= 8) n += le;
        else if (len >= 4) n += le & 0xffffffffull;
        else if (len >= 2) n += le & 0xffffull;
        else if (len >= 1) n += le & 0xffull;
        len -= 8;
    }
    u64 a, b, c;
    a = b = 0xDEADBEEFull;
    c = 0x9e3779b97f4a7c13ull + 8ull;
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

// a5. owner rank: ((hash >> 24) & 0x7ffff) % nranks      src/Kmer.h:187-188,2284-2295
__device__ __forceinline__ u32 owner_of(u64 h, u32 nranks) { return (u32)((h >> 24) & 0x7ffffu) % nranks; }
// the same without a division: x < 2^19 and nranks < 8192, so floor(x / nranks) == umulhi(x, floor((2^32 - 1) / nranks) + 1)
// exactly (the multiplier's excess adds less than 2^-13 to a quotient whose fractional part is at most 1 - 1/nranks).
// magic is computed on the host (owner_magic); nranks >= 2.
__device__ __forceinline__ u32 owner_of_fast(u64 h, u32 nranks, u32 magic)
{
    const u32 x = (u32)(h >> 24) & 0x7ffffu;
    return x - __umulhi(x, magic) * nranks;
}

// ------------------------------------------------------------------------------------------------
// placement inside one GPU (ours, not the reference's: the reference's growable sorted buckets are
// replaced by an open-addressing table cut into L2-sized partitions).  A cheap 64-bit mixer gives
// the partition (high half) and the home slot inside it (low half).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 x)
{
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
template <int W>
__device__ __forceinline__ u64 place_hash(const u64 (&w)[W])
{
    u64 h = mix64(w[0]);
#pragma unroll
    for (int i = 1; i < W; ++i) h = mix64(h ^ w[i]);
    return h;
}
__device__ __forceinline__ u32 part_of(u64 ph, u32 n_parts) { return __umulhi((u32)(ph >> 32), n_parts); }
__device__ __forceinline__ u64 home_slot(u64 ph, u64 part_slots) { return ((u64)(u32)ph * part_slots) >> 32; }

// ------------------------------------------------------------------------------------------------
// canonical k-mer state rolled along a read (a2: KmerArrayPair::build + buildLeastComplement,
// src/Kmer.h:1323-1375,356-364).  fwd/rc are both left-aligned.
// ------------------------------------------------------------------------------------------------
template <int W>
struct Roll {
    u64 f[W], r[W];
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int i = 0; i < W; ++i) f[i] = r[i] = 0;
    }
    // pad = 64*W - 2k (unused low bits of the last word)
    __device__ __forceinline__ void push(u32 code, int pad)
    {
#pragma unroll
        for (int i = 0; i < W - 1; ++i) f[i] = (f[i] << 2) | (f[i + 1] >> 62);
        f[W - 1] = (f[W - 1] << 2) | ((u64)code << pad);
#pragma unroll
        for (int i = W - 1; i > 0; --i) r[i] = (r[i] >> 2) | (r[i - 1] << 62);
        r[0] = (r[0] >> 2) | ((u64)(3u - code) << 62);
        r[W - 1] &= ~((1ull << pad) - 1ull);
    }
    // memcmp(fwd, rc) <= 0  (src/Kmer.h:356-364)
    __device__ __forceinline__ bool fwd_is_least() const
    {
        bool le = true, decided = false;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (!decided && f[i] != r[i]) { le = f[i] < r[i]; decided = true; }
        }
        return le;
    }
};

// ------------------------------------------------------------------------------------------------
// K3. table insert: open addressing, linear probing inside one slice (part_slots is even).
// W==1: slot {val, ~key}; claim by 64-bit atomicCAS on the key word, then one RED on the value word.
// W>=2: slot {val, key[W]}; claim by CAS(val: 0 -> LOCK), write key words, publish with READY.
// add = 1 | (isFwd << 32)
// returns: 0 updated existing, 1 claimed new, -1 partition full
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ld_cg64(const void *p)
{
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// W == 1: two 16-byte slots share one 32-byte sector, so the probe unit is an aligned PAIR of slots fetched with one
// 256-bit load: linear probing that starts at the even slot below the home slot and looks at two slots per sector request.
__device__ __forceinline__ void ld_pair32(const void *p, u64 &v0, u64 &k0, u64 &v1, u64 &k1)
{
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v0), "=l"(k0), "=l"(v1), "=l"(k1) : "l"(p));
}

__device__ __forceinline__ u32 pair_sat(u64 v0, u64 v1) { return ((u32)v0 >= MAX_COUNT ? 1u : 0u) | ((u32)v1 >= MAX_COUNT ? 2u : 0u); }

// PRE: the home probe was loaded by the caller (W == 1: key words of the home pair in pk0/pk1 and "count saturated" bits
// in psat; W >= 2: the value word in pk0), which lets a thread keep the home loads of several records in flight before
// it resolves any of them
template <int W, bool PRE = false>
__device__ __forceinline__ int table_insert(const TableView &t, u32 part, u64 slot0, const u64 (&key)[W], u64 add,
                                            u64 *slot_out, u32 *probes_out, u64 pk0 = 0, u64 pk1 = 0, u32 psat = 0)
{
    Slot<W> *base = reinterpret_cast<Slot<W> *>(t.slots) + (u64)part * t.part_slots;
    const u64 S = t.part_slots;
    if (W == 1) {
        u64 s = slot0 & ~1ull;
        const u64 want = ~key[0];
        for (u64 probes = 0; probes < S; probes += 2) {
            Slot<W> *sl = base + s;
            u64 k0, k1;
            u32 sat;
            if (PRE && probes == 0) { k0 = pk0; k1 = pk1; sat = psat; }
            else { u64 v0, v1; ld_pair32(sl, v0, k0, v1, k1); sat = pair_sat(v0, v1); }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                u64 ck = h ? k1 : k0;
                bool full = (sat >> h) & 1u;
                if (ck == 0) {
                    const u64 old = atomicCAS(&sl[h].k[0], 0ull, want);
                    if (old == 0ull) {
                        atomicAdd(&sl[h].val, add);
                        *slot_out = (u64)part * S + s + h; *probes_out = (u32)probes + h;
                        return 1;
                    }
                    ck = old; full = false;
                }
                if (ck == want) {
                    if (!full) atomicAdd(&sl[h].val, add);
                    *slot_out = (u64)part * S + s + h; *probes_out = (u32)probes + h;
                    return 0;
                }
            }
            s = s + 2 >= S ? 0 : s + 2;
        }
        return -1;
    }
    u64 s = slot0;
    for (u64 probes = 0; probes < S; ++probes) {
        Slot<W> *sl = base + s;
        u64 v = (PRE && probes == 0) ? pk0 : ld_cg64(&sl->val);
        if (v == 0) {
            u64 old = atomicCAS(&sl->val, 0ull, VAL_LOCK);
            if (old == 0ull) {
#pragma unroll
                for (int i = 0; i < W; ++i) sl->k[i] = key[i];
                __threadfence();
                atomicExch(&sl->val, VAL_READY | add);
                *slot_out = (u64)part * S + s; *probes_out = (u32)probes;
                return 1;
            }
            v = old;
        }
        while (!(v & VAL_READY)) {            // another thread is publishing this slot
            __nanosleep(32);
            v = ld_cg64(&sl->val);
        }
        bool eq = true;
#pragma unroll
        for (int i = 0; i < W; ++i) eq = eq && (ld_cg64(&sl->k[i]) == key[i]);
        if (eq) {
            if ((u32)v < MAX_COUNT) atomicAdd(&sl->val, add);
            *slot_out = (u64)part * S + s; *probes_out = (u32)probes;
            return 0;
        }
        s = s + 1 == S ? 0 : s + 1;
    }
    return -1;
}

// read-only probe (K7): returns raw value word (0 when absent); slot index via slot_out
template <int W>
__device__ __forceinline__ u64 table_find(const TableView &t, u32 part, u64 slot0, const u64 (&key)[W], u64 *slot_out)
{
    const Slot<W> *base = reinterpret_cast<const Slot<W> *>(t.slots) + (u64)part * t.part_slots;
    const u64 S = t.part_slots;
    if (W == 1) {
        u64 s = slot0 & ~1ull;
        const u64 want = ~key[0];
        for (u64 probes = 0; probes < S; probes += 2) {
            u64 v0, k0, v1, k1;
            ld_pair32(base + s, v0, k0, v1, k1);
            if (k0 == want) { if (slot_out) *slot_out = (u64)part * S + s; return v0; }
            if (k0 == 0) return 0;
            if (k1 == want) { if (slot_out) *slot_out = (u64)part * S + s + 1; return v1; }
            if (k1 == 0) return 0;
            s = s + 2 >= S ? 0 : s + 2;
        }
        return 0;
    }
    u64 s = slot0;
    for (u64 probes = 0; probes < S; ++probes) {
        const Slot<W> *sl = base + s;
        u64 v = ld_cg64(&sl->val);
        if (v == 0) return 0;
        bool eq = true;
#pragma unroll
        for (int i = 0; i < W; ++i) eq = eq && (ld_cg64(&sl->k[i]) == key[i]);
        if (eq) { if (slot_out) *slot_out = (u64)part * S + s; return v; }
        s = s + 1 == S ? 0 : s + 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// a1. base code of one ASCII character (TwoBitSequence::compressBase src/TwoBitSequence.cpp:114-144):
// returns 0..3 for ACGTacgt; 4 for a markup that is N/X ('.' counts as N, :253-260); 5 for any other markup
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 base_code(u32 c)
{
    u32 u = c & 0xDFu;                         // upper-case letters
    if (u == 'A') return 0;
    if (u == 'C') return 1;
    if (u == 'G') return 2;
    if (u == 'T') return 3;
    if (c == 'N' || c == 'X' || c == '.') return 4;
    return 5;
}

}  // namespace kmn
