// kmn_kernels.cuh -- sm_100a kernels of the k-mer spectrum path.
//
//   count pass  = k_weight_mask     (phase 1a: quality weights along each read -> one "counted" bit per k-mer position)
//               + k_kmer_scatter    (phase 1b: bases -> canonical k-mers -> staging set binned by (owner,) table group;
//                                    records are write-combined in shared-memory rings and leave as whole 32-byte sectors)
//               + k_build_entries / k_slice_split / k_count_slices_ws   (phase 2, k <= 31 without weights or extension
//                                    counters: a second radix pass sorts every group's records by table slice, then each
//                                    64 KB slice is brought into shared memory, counted there and written back)
//               + k_build_worklist / k_insert_staged   (phase 2 of every other flavour, overflow lists: group by group
//                                    into L2-resident slices with global atomics, hot k-mers merged in shared memory first)
//   multi-GPU   = copy engines / k_push_plan + k_push_copy / ncclSend+Recv (records to their owners' receive buffers)
//   lookup pass = k_lookup_vals (+ _dist / _peer / k_lookup_words / k_scatter_answers) + k_trim_score
//   table scans = k_histogram, k_purge, k_export, k_import, k_subtract, k_count_live
//
// Why: on B200 a 64-bit atomic to an HBM-resident table runs at ~20 G/s, the same atomic to a <= 64 MB (L2-resident)
// region at 60-190 G/s (profiles/r01_randacc_microbench.csv), and a 16-byte slot update in shared memory at the rate the
// SM issues it.  So the records are sorted until a slice's worth of them meets its slice in shared memory; every byte of
// HBM traffic on the way is streamed (profiles/r02_summary.md: DRAM bytes = algorithmic bytes for every kernel).
#pragma once
#include "kmn_device.cuh"
#include <cooperative_groups.h>

namespace kmn {

static constexpr int MASK_TPB = 256;       // phase 1a: one warp per 32 reads
static constexpr int SCATTER_MAX_TPB = 1024;           // phase 1b: one thread per read; CTA size and CTAs per SM are chosen at run time
#ifndef KMN_INSERT_MIN_CTAS
#define KMN_INSERT_MIN_CTAS 4
#endif
static constexpr int INSERT_TPB = 256;
#ifndef KMN_INSERT_UNROLL
#define KMN_INSERT_UNROLL 2
#endif
static constexpr int INSERT_UNROLL = KMN_INSERT_UNROLL;
static constexpr int INSERT_CHUNK = INSERT_TPB * INSERT_UNROLL;
static constexpr int INSERT_GROUP = 4;      // chunks per ticket (8 measured no better)
static constexpr int COARSE_SHIFT = 6;      // work-list index: one cell per 64 chunks
#define KMN_MAX_PUSH_RANKS 16               // ranks of the NVLink push path (kernel-parameter arrays of peer pointers)

struct ParseArgs {
    const uint8_t *bases;
    const uint8_t *quals;
    const u64 *read_off;
    const uint8_t *discarded;
    u64 n_reads;
    u64 total_bytes;       // bytes valid behind bases / quals
    const double *ptab;    // 256 doubles: Read::qualityToProbability (host-computed, src/Sequence.cpp:522-540)
    u32 k, kb;
    int pad;               // 64*W - 2k
    float min_weight;
    u32 start_char;
    u32 zero_below;        // qualities below this value have probability 0
    u32 *mask;             // phase 1a -> 1b: bit (read_off[r] + i) set iff k-mer i of read r is counted
    float *wts;            // optional (KMN_VALUE_WEIGHTS): fp32 weight of k-mer i of read r at [read_off[r] + i]
    u32 nranks, rank;
    u32 owner_magic;       // floor((2^32 - 1) / nranks) + 1 (owner_of_fast); nranks >= 2
    u32 use_lookup8;
    u32 l2_hints;          // staging stores carry an L2 evict_last policy
    u32 cta_rot;           // phase 1b: the pieces of this launch are dealt to CTAs starting at this CTA, so that a
                           // sequence of small launches fills the per-CTA sub-regions evenly instead of always the first ones
    u32 piece_shift;       // log2(reads per piece), 5 for large launches; small launches use smaller pieces so that they
                           // still spread over all CTAs (the sub-region capacities assume an even spread)
    u32 ring_R;            // phase 1b: record slots of one bin's shared-memory ring (power of two >= 4; 0 = no rings)
    u32 fast_bound;        // phase 1a: non-uniform reads whose weights are bounded away from min_weight skip the recurrence
    u32 scatter_steps;     // phase 1b: walker steps per round (between two flushes of the rings)
    TableView table;
    StageView stage;
    Counters *ctr;
    // multi-GPU lookup pass: requests are appended to per-destination regions
    u64 *flags;            // [0] += records lost to a full remote-owner sub-region (push path; reported by kmn_count_finish)
    u64 *send_recs;        // [nranks][send_cap][W]
    u64 *send_cursor;      // [nranks]
    u64 send_cap;
};

// ------------------------------------------------------------------------------------------------
// unaligned 8-byte fetch from a byte buffer through two aligned, bounds-guarded 8-byte loads
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ld_nc64(const u64 *p)
{
    u64 v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 load8(const uint8_t *buf, long long idx, long long total)
{
    // bytes buf[idx .. idx+8), zero where out of [0,total)
    const unsigned long long base = (unsigned long long)buf;
    long long a = (long long)(base + idx);
    long long a0 = a & ~7ll;
    int sh = (int)(a & 7) * 8;
    long long lo_lim = (long long)(base & ~7ull);
    long long hi_lim = (long long)((base + (unsigned long long)total + 7ull) & ~7ull);   // exclusive, aligned
    u64 w0 = (a0 >= lo_lim && a0 < hi_lim) ? ld_nc64((const u64 *)a0) : 0ull;
    if (sh == 0) return w0;
    long long a1 = a0 + 8;
    u64 w1 = (a1 >= lo_lim && a1 < hi_lim) ? ld_nc64((const u64 *)a1) : 0ull;
    return (w0 >> sh) | (w1 << (64 - sh));
}

// ------------------------------------------------------------------------------------------------
// record = W key words [+ 1 extra word].  Without the extra word the strand flag sits in bit 0 of the
// last key word (free because k%32 != 0).  Extra word: bit0 strand, bits 8..15 extension byte,
// bits 32..63 fp32 weight.
// ------------------------------------------------------------------------------------------------
template <int W, bool HASX>
struct Rec {
    static constexpr int RW = W + (HASX ? 1 : 0);
    u64 w[RW];
    __device__ __forceinline__ void pack(const u64 (&key)[W], bool fwd, float weight, u32 extbyte)
    {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = key[i];
        if (HASX) w[W] = (u64)(fwd ? 1u : 0u) | ((u64)(extbyte & 0xffu) << 8) | ((u64)__float_as_uint(weight) << 32);
        else w[W - 1] |= (fwd ? 1ull : 0ull);
    }
    __device__ __forceinline__ void unpack(u64 (&key)[W], bool &fwd, float &weight, u32 &extbyte) const
    {
#pragma unroll
        for (int i = 0; i < W; ++i) key[i] = w[i];
        if (HASX) { fwd = w[W] & 1ull; extbyte = (u32)(w[W] >> 8) & 0xffu; weight = __uint_as_float((u32)(w[W] >> 32)); }
        else { fwd = key[W - 1] & 1ull; key[W - 1] &= ~1ull; weight = 0.f; extbyte = 0x3f; }
    }
};

// extension byte: bits 0..2 left code, bits 3..5 right code; codes A,C,G,T,N,X = 0..5, 7 = not counted
// (ExtensionTracking::trackExtension src/KmerTrackingData.h:195-201: counted iff qual>=20 or base is N/X)

template <int W, bool HASX>
__device__ __forceinline__ void insert_record(const TableView &t, const Rec<W, HASX> &rec, u64 &n_unique, u64 &n_full, u64 &n_probes)
{
    u64 key[W]; bool fwd; float weight; u32 eb;
    rec.unpack(key, fwd, weight, eb);
    u64 ph = place_hash<W>(key);
    u32 part = part_of(ph, t.n_parts);
    u64 slot; u32 probes = 0;
    int r = table_insert<W>(t, part, home_slot(ph, t.part_slots), key, 1ull | ((u64)(fwd ? 1u : 0u) << 32), &slot, &probes);
    if (r < 0) { n_full++; return; }
    n_unique += (u64)r;
    n_probes += probes;
    if (HASX) {
        if (t.wsum) atomicAdd(&t.wsum[slot], weight);
        if (t.ext) {
            u32 l = eb & 7u, rr = (eb >> 3) & 7u;
            if (l < 6) atomicAdd(&t.ext[slot * 12 + l], 1u);
            if (rr < 6) atomicAdd(&t.ext[slot * 12 + 6 + rr], 1u);
        }
    }
}

struct LocalCtr { u64 raw, good, unique, full, direct, probes; };

__device__ __forceinline__ void ctr_commit(Counters *ctr, LocalCtr lc)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lc.raw += __shfl_xor_sync(0xffffffffu, lc.raw, o);
        lc.good += __shfl_xor_sync(0xffffffffu, lc.good, o);
        lc.unique += __shfl_xor_sync(0xffffffffu, lc.unique, o);
        lc.full += __shfl_xor_sync(0xffffffffu, lc.full, o);
        lc.direct += __shfl_xor_sync(0xffffffffu, lc.direct, o);
        lc.probes += __shfl_xor_sync(0xffffffffu, lc.probes, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (lc.raw) atomicAdd(&ctr->raw, lc.raw);
        if (lc.good) atomicAdd(&ctr->raw_good, lc.good);
        if (lc.unique) atomicAdd(&ctr->unique, lc.unique);
        if (lc.full) atomicAdd(&ctr->table_full, lc.full);
        if (lc.direct) atomicAdd(&ctr->direct, lc.direct);
        if (lc.probes) atomicAdd(&ctr->probe_steps, lc.probes);
    }
}

// ------------------------------------------------------------------------------------------------
// byte stream over a (possibly unaligned) region of a global buffer, 8 bytes per step, with the aligned word
// of the NEXT step requested one step ahead so its latency is hidden behind the current step's arithmetic.
// ------------------------------------------------------------------------------------------------
struct Stream {
    const u64 *p;       // aligned word holding the current position
    u64 w0, w1;         // words p[0], p[1]
    u64 wn;             // p[2], in flight
    int sh;
    __device__ __forceinline__ static u64 guarded(const u64 *q, const uint8_t *buf, u64 total)
    {
        const unsigned long long base = (unsigned long long)buf;
        const unsigned long long lo = base & ~7ull, hi = (base + total + 7ull) & ~7ull;
        const unsigned long long a = (unsigned long long)q;
        return (a >= lo && a < hi) ? ld_nc64(q) : 0ull;
    }
    __device__ __forceinline__ void init(const uint8_t *buf, long long idx, u64 total)
    {
        const unsigned long long a = (unsigned long long)buf + (unsigned long long)idx;
        p = (const u64 *)(a & ~7ull);
        sh = (int)(a & 7ull) * 8;
        w0 = guarded(p, buf, total);
        w1 = guarded(p + 1, buf, total);
        wn = 0;
    }
    __device__ __forceinline__ void prefetch(const uint8_t *buf, u64 total) { wn = guarded(p + 2, buf, total); }
    __device__ __forceinline__ u64 get() const { return sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0; }
    __device__ __forceinline__ u64 peek() const { return sh ? (w1 >> sh) | (wn << (64 - sh)) : w1; }   // next 8 bytes
    __device__ __forceinline__ void advance() { ++p; w0 = w1; w1 = wn; }
};

// SWAR helpers on 8 packed bytes
__device__ __forceinline__ u64 bytes_eq(u64 x, u64 pat)          // 0x80 in every byte of x equal to the pattern byte
{
    const u64 t = x ^ pat, m = 0x7f7f7f7f7f7f7f7full;
    return ~(((t & m) + m) | t | m);
}

// ------------------------------------------------------------------------------------------------
// per-read walker: rolls the canonical k-mer and the quality weight along one read, 8 bases per step.
// a2 KmerArrayPair::build (src/Kmer.h:1323-1375), a3 KmerReadUtils::buildWeightedKmers
// (src/KmerReadUtils.h:176-248): w re-seeded at i%1024==0 or w==0, otherwise w *= p[q_in]/p[q_out];
// any markup inside the window zeroes w; stored weight is (float)w.
// The per-base body is branch-light: the weight is only touched when something can change it (first window,
// incoming quality != outgoing quality -- x/x is exactly 1.0 --, re-seed boundary, w==0, markup in the window).
// ------------------------------------------------------------------------------------------------
template <int W>
struct Walker {
    Roll<W> roll;
    double w;
    float wf;          // (float)w of the last emitted k-mer
    bool good;         // wf > min_weight
    u64 off;
    u32 len, j;
    int last_bad;      // last position holding a markup (non-ACGT), -1 none
    int last_zero;     // last position whose quality has probability 0, -1 none
    u32 first_nx;      // firstMarkupNorX: position+1 of the first N/X markup, 0 none
    Stream sb, sqi, sqo;

    __device__ __forceinline__ void clear()
    {
        roll.reset(); w = 0.0; wf = 0.f; good = false; off = 0; len = 0; j = 0; last_bad = -1; last_zero = -1; first_nx = 0;
    }
    template <bool NEED_Q>
    __device__ __forceinline__ void begin(const ParseArgs &a, u64 off_, u32 len_)
    {
        clear();
        off = off_; len = len_;
        sb.init(a.bases, (long long)off_, a.total_bytes);
        if (NEED_Q) {
            sqi.init(a.quals, (long long)off_, a.total_bytes);
            sqo.init(a.quals, (long long)off_ - (long long)a.k, a.total_bytes);
        }
    }
};

// One step = up to 8 bases.  EMIT(i, key, fwd, weightf, good, extbyte) is called for every k-mer position i.
// KEYS=false: weights only (the k-mer is not rolled; emit receives a zero key).
template <int W, bool NEED_W, bool EXT, bool KEYS = true, typename EMIT>
__device__ __forceinline__ void walker_step(Walker<W> &s, const ParseArgs &a, const double *ptab, EMIT &&emit)
{
    const u32 k = a.k;
    const u32 j0 = s.j;
    constexpr bool NEED_Q = NEED_W || EXT;
    s.sb.prefetch(a.bases, a.total_bytes);
    if (NEED_Q) { s.sqi.prefetch(a.quals, a.total_bytes); s.sqo.prefetch(a.quals, a.total_bytes); }
    const u64 bw = s.sb.get();
    u64 qin = 0, qout = 0;
    if (NEED_Q) { qin = s.sqi.get(); qout = s.sqo.get(); }

    // base codes, SWAR: x=(c>>1)&3 -> A0 C1 G3 T2 ; code = x ^ (x>>1) -> A0 C1 G2 T3 ; markups -> 0 (packed as A)
    const u64 up = bw & 0xDFDFDFDFDFDFDFDFull;
    const u64 valid = bytes_eq(up, 0x4141414141414141ull) | bytes_eq(up, 0x4343434343434343ull) |
                      bytes_eq(up, 0x4747474747474747ull) | bytes_eq(up, 0x5454545454545454ull);
    u64 codes = (bw >> 1) & 0x0303030303030303ull;
    codes ^= (codes >> 1) & 0x0101010101010101ull;
    codes &= (valid >> 7) * 3ull;
    const u64 qdiff = qin ^ qout;                                   // non-zero byte: incoming quality != outgoing quality
    u64 bnext = 0, qnext = 0;
    if (EXT) { bnext = s.sb.peek(); qnext = s.sqi.peek(); }

    if (NEED_W && !KEYS && !EXT) {
        // weights only: a whole step in which nothing can change the weight -- past the first window, no markup in or
        // entering the window, every incoming quality equal to the outgoing one (x/x == 1.0 exactly), no zero-probability
        // quality, no re-seed boundary, w != 0 -- repeats the previous k-mer's weight eight times.
        const u32 i0 = j0 + 1 - k;
        const u64 zb = 0x0101010101010101ull * a.zero_below;
        const bool has_zero_q = (((qin - zb) & ~qin) & 0x8080808080808080ull) != 0ull;
        if (j0 >= k && j0 + 8 <= s.len && (valid & 0x8080808080808080ull) == 0x8080808080808080ull && qdiff == 0ull && !has_zero_q &&
            (i0 & 1023u) != 0u && (i0 & 1023u) <= 1016u && s.w != 0.0 && s.last_bad < (int)i0) {
#pragma unroll 1
            for (int u = 0; u < 8; ++u) emit(i0 + u, s.roll.f, true, s.wf, s.good, 0x3fu);
            s.sb.advance();
            s.sqi.advance(); s.sqo.advance();
            s.j = j0 + 8;
            return;
        }
    }

    // the weights-only walker (KEYS == false) keeps this loop rolled: unrolled it is 110 KB of code and the kernel
    // stalls on instruction fetch (ncu: "no instruction" was the top stall); the k-mer walkers want it unrolled
    auto per_base = [&](const int u) {
        const u32 j = j0 + u;
        if (j < s.len) {
            if (!((valid >> (8 * u + 7)) & 1ull)) {                 // markup (rare)
                const u32 c = (u32)(bw >> (8 * u)) & 0xffu;
                s.last_bad = (int)j;
                if ((c == 'N' || c == 'X' || c == '.') && s.first_nx == 0) s.first_nx = j + 1;
            }
            const u32 out_code = (u32)(s.roll.f[0] >> 62);          // base leaving the window (left neighbour of the new k-mer)
            if (KEYS) s.roll.push((u32)(codes >> (8 * u)) & 3u, a.pad);
            const u32 i = j + 1 - k;                                // k-mer index (valid when j+1 >= k)
            if (NEED_W) {
                const u32 qi = (u32)(qin >> (8 * u)) & 0xffu;
                const bool first = j < k;
                if (first || qi < a.zero_below) {                   // first window product / zero-probability bookkeeping
                    const double pi = ptab[qi];
                    if (pi == 0.0) s.last_zero = (int)j;
                    if (first) s.w = (j == 0) ? pi : s.w * pi;      // w = p[q0]*p[q1]*... left to right
                }
                if (j + 1 >= k) {
                    const bool touch = i == 0 || ((qdiff >> (8 * u)) & 0xffull) != 0 || (i & 1023u) == 0u || s.w == 0.0 || s.last_bad >= (int)i;
                    if (touch) {
                        if (i > 0) {
                            if ((i & 1023u) == 0u || s.w == 0.0) {
                                if (s.last_zero >= (int)i) s.w = 0.0;      // a zero factor makes the product exactly 0
                                else {
                                    double ww = 1.0;
                                    for (u32 q = 0; q < k; ++q) ww *= ptab[a.quals[s.off + i + q]];
                                    s.w = ww;
                                }
                            } else {
                                const u32 qo = (u32)(qout >> (8 * u)) & 0xffu;
                                if (qi != qo) { const double change = ptab[qi] / ptab[qo]; s.w *= change; }
                            }
                        }
                        if (s.last_bad >= (int)i) s.w = 0.0;               // markup inside [i, i+k)
                        s.wf = (float)s.w;
                        s.good = s.wf > a.min_weight;
                    }
                }
            }
            if (j + 1 >= k) {
                const bool fwd = KEYS ? s.roll.fwd_is_least() : true;
                u32 eb = 0x3f;
                if (EXT) {
                    // left = base i-1 (or X,20), right = base i+k (or X,20); N neighbours read as A   KmerReadUtils.h:224-236
                    const u32 qo = (u32)(qout >> (8 * u)) & 0xffu;
                    u32 lc = (i == 0) ? 5u : out_code, lq = (i == 0) ? 20u : (qo - a.start_char);
                    u32 rc = 5u, rq = 20u;
                    if (j + 1 < s.len) {
                        const u32 nb = (u < 7) ? ((u32)(bw >> (8 * ((u + 1) & 7))) & 0xffu) : ((u32)bnext & 0xffu);
                        const u32 nq = (u < 7) ? ((u32)(qin >> (8 * ((u + 1) & 7))) & 0xffu) : ((u32)qnext & 0xffu);
                        const u32 ncode = base_code(nb);
                        rc = ncode >= 4 ? 0u : ncode;
                        rq = nq - a.start_char;
                    }
                    if (!fwd) { u32 tl = lc, tq = lq; lc = rc < 4 ? 3u - rc : rc; lq = rq; rc = tl < 4 ? 3u - tl : tl; rq = tq; }
                    const u32 le = ((lq & 0xffu) >= 20u || lc >= 4) ? lc : 7u;
                    const u32 re = ((rq & 0xffu) >= 20u || rc >= 4) ? rc : 7u;
                    eb = le | (re << 3);
                }
                // without weights `good` = no markup inside the window (the k-mer's weight would be 0, KmerReadUtils.h:207-217)
                emit(i, fwd ? s.roll.f : s.roll.r, fwd, NEED_W ? s.wf : 1.0f, NEED_W ? s.good : s.last_bad < (int)i, eb);
            }
        }
    };
    if constexpr (KEYS) {
#pragma unroll
        for (int u = 0; u < 8; ++u) per_base(u);
    } else {
#pragma unroll 1
        for (int u = 0; u < 8; ++u) per_base(u);
    }
    s.sb.advance();
    if (NEED_Q) { s.sqi.advance(); s.sqo.advance(); }
    s.j = j0 + 8;
}

// ------------------------------------------------------------------------------------------------
// K2a: phase 1a of the count pass.  Evaluates the reference's sequential weight recurrence (a3,
// src/KmerReadUtils.h:176-248); the only thing the count pass needs from it is one bit per k-mer position --
// "(float)w > minimumWeight" (src/KmerSpectrum.h:1598, src/KmerTrackingData.h:354-364) -- which goes to a bit array
// indexed by the k-mer's first base in the concatenated batch (and, for KMN_VALUE_WEIGHTS, the fp32 weight itself).
//
// Most reads of a modern run are UNIFORM: every base is ACGT and every quality byte is the same value q.  For such a
// read the recurrence never changes the weight: it is seeded with p[q]*p[q]*...*p[q] (k factors, left to right), every
// later step multiplies by p[q]/p[q] (skipped by the reference's own `if change != 1`-free arithmetic: x/x == 1.0
// exactly), and the periodic re-seed (i%1024==0) recomputes the same left-to-right product.  So a warp first
// classifies 32 reads with SWAR compares (one lane per read), uniform reads set their whole bit range at once from a
// 256-entry table of those products, and the others are queued in shared memory and walked 32 at a time -- the
// sequential walker then runs with all lanes busy instead of diverging on every step.
// ------------------------------------------------------------------------------------------------
template <bool WTS>
__device__ __forceinline__ void weight_walk_read(const ParseArgs &a, const double *ptab, u64 r, LocalCtr &lc)
{
    const u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
    const u32 len = (u32)(o1 - o0);
    Walker<1> st;
    st.template begin<true>(a, o0, len);
    u64 curw = ~0ull;
    u32 acc = 0;
    auto emit = [&](u32 i, const u64 (&)[1], bool, float wf, bool good, u32) {
        const u64 gb = o0 + i, w = gb >> 5;
        if (w != curw) { if (acc) atomicOr(&a.mask[curw], acc); acc = 0; curw = w; }
        if (good) { acc |= 1u << (u32)(gb & 31ull); if (WTS) lc.good++; }
        if (WTS) a.wts[gb] = wf;
        lc.raw++;
    };
    while (st.j < st.len) walker_step<1, true, false, false>(st, a, ptab, emit);
    if (acc) atomicOr(&a.mask[curw], acc);
}

// sets bits [b0, b0+n) of the mask (neighbouring reads share the boundary words)
__device__ __forceinline__ void mask_set_range(u32 *mask, u64 b0, u32 n)
{
    u64 w = b0 >> 5;
    u32 sh = (u32)(b0 & 31ull);
    while (n) {
        const u32 take = min(n, 32u - sh);
        const u32 bits = (take == 32u ? 0xffffffffu : ((1u << take) - 1u)) << sh;
        atomicOr(&mask[w], bits);
        n -= take; sh = 0; ++w;
    }
}

// clears bits [b0, b0+n) of the mask
__device__ __forceinline__ void mask_clear_range(u32 *mask, u64 b0, u32 n)
{
    u64 w = b0 >> 5;
    u32 sh = (u32)(b0 & 31ull);
    while (n) {
        const u32 take = min(n, 32u - sh);
        const u32 bits = (take == 32u ? 0xffffffffu : ((1u << take) - 1u)) << sh;
        atomicAnd(&mask[w], ~bits);
        n -= take; sh = 0; ++w;
    }
}
// k-mers [max(0, b-k+1), min(b, n-1)] of a read contain position b: their "counted" bits are cleared
__device__ __forceinline__ void mask_clear_around(u32 *mask, u64 o0, u32 b, u32 k, u32 n)
{
    const u32 lo = b + 1u >= k ? b + 1u - k : 0u;
    const u32 hi = b < n ? b : n - 1u;
    if (hi >= lo) mask_clear_range(mask, o0 + lo, hi - lo + 1u);
}

// lower bound of p^n for 0 < p <= 1 (square-and-multiply rounds differently from a chain of n multiplications)
__device__ __forceinline__ double pow_bound(double p, u32 n)
{
    double r = 1.0, sq = p;
    for (u32 e = n; e; e >>= 1) { if (e & 1u) r *= sq; sq *= sq; }
    return r * (1.0 - 1e-12);
}
// true when a weight >= lb (1 - 3e-13) is certainly counted: (float)w > min_weight
__device__ __forceinline__ bool bound_passes(double lb, float min_weight)
{
    return lb > 1e-30 && lb * (1.0 - 9.5367431640625e-07) > (double)min_weight;
}

// calls f(j, x, m) for the lane's read [pos, pos+len) in the shared buffer: x = bytes j..j+7, m = byte mask of the valid ones
template <typename F>
__device__ __forceinline__ void smem_each_word(const u64 *buf64, u32 pos, u32 len, F &&f)
{
    u32 wi = pos >> 3;
    const u32 sh = (pos & 7u) * 8u;
    u64 w0 = buf64[wi];
    for (u32 j = 0; j < len; j += 8) {
        const u64 w1 = buf64[++wi];
        const u64 x = sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0;
        w0 = w1;
        const u32 nb = len - j;
        f(j, x, nb >= 8 ? ~0ull : ((1ull << (8 * nb)) - 1ull));
    }
}

static constexpr u32 MASK_WBUF = 8192 + 64;      // bytes of one warp's staging buffer: 32 reads of up to 256 bases

// coalesced copy of the 16-byte blocks that hold bytes [b0, b1) of buf into a warp's shared buffer, in two halves: load()
// requests up to STAGE_U blocks per lane at once and keeps them in registers (ten independent 16-byte loads per lane in
// flight -- with one load in flight per lane and 768 threads per SM the kernel ran at the latency of 20 dependent memory
// round trips per round), store() writes them to the buffer (and copies what did not fit into the registers).  Between
// the two the warp works on the buffer's previous contents.  sh0 = offset of byte b0 inside the buffer.
static constexpr int STAGE_U = 10;
struct StageRegs {
    uint4 v[STAGE_U];
    const uint4 *src;
    u32 n_blk, sh0;
    __device__ __forceinline__ void load(const uint8_t *buf, u64 b0, u64 b1, u32 lane)
    {
        const unsigned long long a0 = (unsigned long long)buf + b0, a1 = (unsigned long long)buf + b1;
        const unsigned long long blk0 = a0 & ~15ull;
        src = reinterpret_cast<const uint4 *>(blk0);
        n_blk = (u32)((a1 - blk0 + 15ull) >> 4);
        sh0 = (u32)(a0 - blk0);
#pragma unroll
        for (int u = 0; u < STAGE_U; ++u) {
            const u32 i = (u32)u * 32u + lane;
            if (i < n_blk) v[u] = __ldg(src + i);
        }
    }
    __device__ __forceinline__ void store(uint4 *dst, u32 lane) const
    {
#pragma unroll
        for (int u = 0; u < STAGE_U; ++u) {
            const u32 i = (u32)u * 32u + lane;
            if (i < n_blk) dst[i] = v[u];
        }
        for (u32 i = (u32)STAGE_U * 32u + lane; i < n_blk; i += 32) dst[i] = __ldg(src + i);
    }
};

// true iff some byte of the lane's read [pos, pos+len) in the shared buffer differs from pat (pat_valid: ACGT test instead)
template <bool BASES>
__device__ __forceinline__ bool smem_scan_read(const u64 *buf64, u32 pos, u32 len, u64 qpat)
{
    u32 wi = pos >> 3;
    const u32 sh = (pos & 7u) * 8u;
    u64 w0 = buf64[wi];
    u64 bad = 0;
    for (u32 j = 0; j < len && !bad; j += 8) {
        const u64 w1 = buf64[++wi];
        const u64 x = sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0;
        w0 = w1;
        const u32 nb = len - j;
        const u64 m = nb >= 8 ? ~0ull : ((1ull << (8 * nb)) - 1ull);
        if (BASES) {
            const u64 up = x & 0xDFDFDFDFDFDFDFDFull;
            const u64 valid = bytes_eq(up, 0x4141414141414141ull) | bytes_eq(up, 0x4343434343434343ull) |
                              bytes_eq(up, 0x4747474747474747ull) | bytes_eq(up, 0x5454545454545454ull);
            bad = (~valid & 0x8080808080808080ull) & m;
        } else bad = (x ^ qpat) & m;
    }
    return bad != 0;
}

template <bool WTS>
__global__ void __launch_bounds__(MASK_TPB, 3) k_weight_mask(ParseArgs a)
{
    __shared__ double ptab[256];
    __shared__ float powk[256];                       // (float)(p[q] * p[q] * ... ), k factors, left to right
    __shared__ u64 queue[MASK_TPB / 32][64];          // per-warp queue of non-uniform reads
    extern __shared__ __align__(16) unsigned char mask_smem[];   // [warps][MASK_WBUF] staging buffers
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        const double p = a.ptab[i];
        ptab[i] = p;
        double w = p;
        for (u32 q = 1; q < a.k; ++q) w = w * p;
        powk[i] = (float)w;
    }
    __syncthreads();
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u64 n_warps = (u64)gridDim.x * (MASK_TPB / 32);
    u64 *q = queue[warp];
    uint4 *wbuf = reinterpret_cast<uint4 *>(mask_smem + (size_t)warp * MASK_WBUF);
    const u64 *wbuf64 = reinterpret_cast<const u64 *>(wbuf);
    u32 qn = 0;
    u64 base = ((u64)blockIdx.x * (MASK_TPB / 32) + warp) * 32u;
    // offsets of a round's reads: lane l holds read_off[b + l], e = end of the last read; requested one round ahead
    u64 pre_o0 = 0, pre_B1 = 0;
    auto request_offsets = [&](u64 b, u64 &o, u64 &e) {
        if (b < a.n_reads) {
            const u32 n = (u32)min((u64)32, a.n_reads - b);
            o = lane < n ? __ldg(&a.read_off[b + lane]) : 0ull;
            e = __ldg(&a.read_off[b + n]);
        }
    };
    request_offsets(base, pre_o0, pre_B1);
    while (true) {
        const bool more = base < a.n_reads;             // warp-uniform
        if (more) {
        const u64 r = base + lane;
        const u32 nr = (u32)min((u64)32, a.n_reads - base);
        const u64 o0 = pre_o0;
        const u64 B0 = __shfl_sync(0xffffffffu, o0, 0);
        const u64 B1 = pre_B1;
        const u64 nxt = __shfl_down_sync(0xffffffffu, o0, 1);
        const u64 o1 = lane + 1 < nr ? nxt : B1;
        const u32 len = lane < nr ? (u32)(o1 - o0) : 0u;
        const bool live = lane < nr && len >= a.k && !(a.discarded && a.discarded[r]);
        bool slow = false;
        u64 q0 = 0;
        if (B1 - B0 + 32 <= MASK_WBUF) {
            // the round's bytes go through shared memory with coalesced 16-byte loads: qualities first, then bases
            StageRegs sr;
            sr.load(a.quals, B0, B1, lane);
            request_offsets(base + n_warps * 32u, pre_o0, pre_B1);
            __syncwarp();
            sr.store(wbuf, lane);
            __syncwarp();
            bool bounded = false;
            const u32 n_kmers = live ? len - a.k + 1u : 0u;
            if (live) {
                const u32 pos = sr.sh0 + (u32)(o0 - B0);
                q0 = (wbuf64[pos >> 3] >> ((pos & 7u) * 8u)) & 0xffull;
                const u64 qpat = q0 * 0x0101010101010101ull;
                slow = smem_scan_read<false>(wbuf64, pos, len, qpat);
                if (!WTS && slow && a.fast_bound) {
                    // BOUNDED read: every probability is <= 1, so the weight of a window without a zero-probability quality
                    // is at least LB = the product of ALL non-zero probabilities of the read; the recurrence (a chain of at
                    // most 1024 multiplications and divisions between two re-seeds, relative error < 3e-13) cannot bring
                    // (float)w down to min_weight when LB (1 - 2^-20) > min_weight.  A window WITH a zero-probability
                    // quality has weight exactly 0 (walker_step: zero factor).  So the "counted" bits of such a read are:
                    // no zero-probability quality in [i, i+k) -- no recurrence needed.  (Markups are phase 1b's business.)
                    const double p0 = ptab[q0];
                    if (p0 > 0.0 && p0 <= 1.0) {
                        double lb = pow_bound(p0, len);                // p0^len <= p0^(bytes equal to q0)
                        bool ok = true;
                        u32 nz = 0, z0 = 0, z1 = 0;                    // zero-probability positions (the first two are remembered)
                        smem_each_word(wbuf64, pos, len, [&](u32 j, u64 x, u64 m) {
                            u64 d = ~bytes_eq(x, qpat) & 0x8080808080808080ull & m;
                            while (d) {
                                const u32 bi = (u32)(__ffsll((long long)d) - 1) >> 3;
                                d &= d - 1ull;
                                const double p = ptab[(u32)(x >> (8u * bi)) & 0xffu];
                                if (p == 0.0) { if (nz == 0) z0 = j + bi; else if (nz == 1) z1 = j + bi; ++nz; }
                                else if (p <= 1.0) lb *= p;
                                else ok = false;
                            }
                        });
                        bounded = ok && bound_passes(lb, a.min_weight);
                        if (bounded) {
                            mask_set_range(a.mask, o0, n_kmers);
                            if (nz > 0) mask_clear_around(a.mask, o0, z0, a.k, n_kmers);
                            if (nz > 1) mask_clear_around(a.mask, o0, z1, a.k, n_kmers);
                            if (nz > 2) smem_each_word(wbuf64, pos, len, [&](u32 j, u64 x, u64 m) {
                                u64 d = ~bytes_eq(x, qpat) & 0x8080808080808080ull & m;
                                while (d) {
                                    const u32 bi = (u32)(__ffsll((long long)d) - 1) >> 3;
                                    d &= d - 1ull;
                                    if (j + bi > z1 && ptab[(u32)(x >> (8u * bi)) & 0xffu] == 0.0) mask_clear_around(a.mask, o0, j + bi, a.k, n_kmers);
                                }
                            });
                            lc.raw += n_kmers;
                            slow = false;
                            q0 = 256;                                  // not a uniform read either
                        }
                    }
                }
            }
            if (WTS) {
                // the weight of every k-mer is stored: a markup among the bases zeroes it, the walker has to see the read
                sr.load(a.bases, B0, B1, lane);
                __syncwarp();
                sr.store(wbuf, lane);
                __syncwarp();
                if (live && !slow && smem_scan_read<true>(wbuf64, sr.sh0 + (u32)(o0 - B0), len, 0)) slow = true;
            }
        } else {
            request_offsets(base + n_warps * 32u, pre_o0, pre_B1);
            if (live) {
            // long reads: per-lane streams over global memory
            Stream sb, sq;
            sb.init(a.bases, (long long)o0, a.total_bytes);
            sq.init(a.quals, (long long)o0, a.total_bytes);
            q0 = sq.get() & 0xffull;
            const u64 qpat = q0 * 0x0101010101010101ull;
            u64 bad = 0;
            for (u32 j = 0; j < len && !bad; j += 8) {
                if (WTS) sb.prefetch(a.bases, a.total_bytes);
                sq.prefetch(a.quals, a.total_bytes);
                const u64 bw = WTS ? sb.get() : 0x4141414141414141ull, qw = sq.get();
                const u32 nb = len - j;
                const u64 m = nb >= 8 ? ~0ull : ((1ull << (8 * nb)) - 1ull);
                const u64 up = bw & 0xDFDFDFDFDFDFDFDFull;
                const u64 valid = bytes_eq(up, 0x4141414141414141ull) | bytes_eq(up, 0x4343434343434343ull) |
                                  bytes_eq(up, 0x4747474747474747ull) | bytes_eq(up, 0x5454545454545454ull);
                bad = ((qw ^ qpat) | (~valid & 0x8080808080808080ull)) & m;
                if (WTS) sb.advance();
                sq.advance();
            }
            slow = bad != 0;
            }
        }
        if (live && !slow && q0 < 256) {
            // uniform read: every k-mer has the weight p[q]^k (evaluated left to right)
            const u32 n = len - a.k + 1;
            const float wf = powk[q0];
            lc.raw += n;
            if (wf > a.min_weight) { if (WTS) lc.good += n; mask_set_range(a.mask, o0, n); }
            if (WTS) for (u32 i = 0; i < n; ++i) a.wts[o0 + i] = wf;
        }
        const u32 bal = __ballot_sync(0xffffffffu, slow);
        if (slow) q[qn + __popc(bal & ((1u << lane) - 1u))] = r;
        qn += __popc(bal);
        __syncwarp();
        base += n_warps * 32u;
        }
        // one call site for the sequential walker (its code is large): a full warp of queued reads, or the rest at the end
        if (qn >= 32 || (!more && qn > 0)) {
            const u32 n = qn < 32u ? qn : 32u;
            qn -= n;
            if (lane < n) weight_walk_read<WTS>(a, ptab, q[qn + lane], lc);
            __syncwarp();
        }
        if (!more && qn == 0) break;
    }
    ctr_commit(a.ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// staging write: position pos of this CTA's sub-region of partition `part`, or a direct insert when it is full
// ------------------------------------------------------------------------------------------------
// L2 eviction policies.  The staging stores are 8*RW-byte pieces of sectors that only become complete several
// stores later; a sector evicted half-written costs a DRAM read-modify-write, so these stores ask L2 to keep their
// lines (evict_last) while the streamed inputs are marked evict_first.
__device__ __forceinline__ u64 l2_policy_evict_last()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_normal()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint64(u64 *p, u64 v, u64 policy)
{
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(policy) : "memory");
}

// ------------------------------------------------------------------------------------------------
// K1+K2b (+K5 partition): phase 1b of the count pass.  One thread walks one read and rolls the canonical k-mer
// (a1 TwoBitSequence::compressSequence src/TwoBitSequence.cpp:242-269, a2 KmerArrayPair::build src/Kmer.h:1323-1375,
// buildLeastComplement :356-364) 8 bases per step; k-mers whose "counted" bit (phase 1a) is set become records of this
// CTA's sub-region of their (owner,) table group.
//
// What bounds this kernel on B200 is the SM's store path to L2: it moves one 32-byte SECTOR per ~3.5 cycles whether the
// store fills the sector or only 8 bytes of it (bench/micro/lsu.cu: a shared atomic + one scattered 8-byte store per record
// runs at 64-100 G records/s, the shared atomic + a shared store alone at 940 G/s).  So records are write-combined in shared
// memory: every bin has a ring of R record slots addressed by the record's position in the bin's sub-region (position p
// lives in ring slot p mod R until it is flushed), the position comes from one shared atomicAdd, and after every walker
// step of the CTA (a ROUND: at most 8 records per thread) one thread per bin moves the bin's complete, aligned groups of
// four records to global memory as whole sectors.  A record that finds its ring full (more than R records of one bin in
// a round: skewed input) is stored directly, which is merely slower.
// Multi-GPU (DIST 2): bins are (owner, group) pairs (a5: owner = lookup3 hash, src/Kmer.h:2284-2295), the sub-region
// belongs to the owner's part of the staging set, which travels to that owner as it is.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_sector(u64 *dst, const u64 *src_smem)      // 32 bytes, both 32-byte aligned
{
    const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(src_smem), y = *reinterpret_cast<const ulonglong2 *>(src_smem + 2);
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(x.x), "l"(x.y), "l"(y.x), "l"(y.y) : "memory");
}

template <int W, bool HASX, bool EXT, int DIST>
__global__ void __launch_bounds__(1024, 1) k_kmer_scatter(ParseArgs a)
{
    constexpr int RW = Rec<W, HASX>::RW;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const u32 n_parts = a.stage.n_parts;                               // staging partitions = table groups
    const u32 n_bins = DIST == 2 ? n_parts * a.nranks : n_parts;       // (owner, group) bins on the push path
    const u32 n_pad = (n_bins + 31u) & ~31u;
    const u32 lo = a.stage.local_owner();
    u32 *cnt = reinterpret_cast<u32 *>(smem_raw);                      // [n_bins] records given a position so far (may exceed sub_cap)
    u32 *fl = cnt + n_pad;                                             // [n_bins] positions below fl are in global memory
    u64 *ring = reinterpret_cast<u64 *>(fl + n_pad + 32u);
    const u32 R = a.ring_R;                                            // power of two >= 4, or 0: every record is stored directly
    const u32 sub_cap = a.stage.sub_cap;
    for (u32 i = threadIdx.x; i < n_bins; i += blockDim.x) {
        const u32 o = DIST == 2 ? i / n_parts : lo, g = DIST == 2 ? i - o * n_parts : i;
        const u32 c0 = a.stage.count[a.stage.cnt_index(o, g, blockIdx.x)];
        cnt[i] = c0;
        fl[i] = c0 < sub_cap ? c0 : sub_cap;
    }
    __syncthreads();

    LocalCtr lc{0, 0, 0, 0, 0, 0};
    u64 lost = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 keep = a.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    Walker<W> st;
    st.clear();
    // consecutive groups of 32 reads (one warp) go to different CTAs, so even a small batch spreads evenly over the
    // per-CTA sub-regions / segments while a warp still streams one contiguous piece of the batch
    const u32 vblock = (blockIdx.x + gridDim.x - a.cta_rot % gridDim.x) % gridDim.x;
    // piece q of a sweep goes to warp slot q % T (T slots = CTAs x warps, CTA index fastest), lanes take 32/p pieces of p reads
    const u32 p_shift = a.piece_shift, lane_in = threadIdx.x & 31u;
    const u64 n_slots = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 in_sweep = (((u64)(lane_in >> p_shift) * n_slots + (u64)(threadIdx.x >> 5) * gridDim.x + vblock) << p_shift) + (lane_in & ((1u << p_shift) - 1u));

    // sub-region of bin b in the staging set
    auto sub_base = [&](u32 b) -> u64 * {
        const u32 o = DIST == 2 ? b / n_parts : lo, g = DIST == 2 ? b - o * n_parts : b;
        return a.stage.recs + a.stage.sub_index(o, g, blockIdx.x) * sub_cap * RW;
    };
    // one thread per bin: complete aligned groups of four records leave the ring as whole sectors (everything when `all`)
    auto flush = [&](bool all) {
        for (u32 b = threadIdx.x; b < n_bins; b += blockDim.x) {
            u32 c = cnt[b];
            if (c > sub_cap) c = sub_cap;
            const u32 f = fl[b];
            if (c <= f) continue;
            u32 end, nfl;
            if (c - f > R) { end = f + R; nfl = c; }                   // the ring overflowed: positions >= f + R were stored directly
            else { end = all ? c : (c & ~3u); nfl = end; }
            if (end <= f) continue;
            u64 *g = sub_base(b);
            const u64 *rb = ring + (size_t)b * R * RW;
            u32 pos = f;
            for (; pos < end && (pos & 3u); ++pos) {
#pragma unroll
                for (int q = 0; q < RW; ++q) g[(size_t)pos * RW + q] = rb[(size_t)(pos & (R - 1u)) * RW + q];
            }
            for (; pos + 4u <= end; pos += 4u) {
#pragma unroll
                for (int q = 0; q < RW; ++q) st_sector(g + (size_t)pos * RW + 4 * q, rb + (size_t)(pos & (R - 1u)) * RW + 4 * q);
            }
            for (; pos < end; ++pos) {
#pragma unroll
                for (int q = 0; q < RW; ++q) g[(size_t)pos * RW + q] = rb[(size_t)(pos & (R - 1u)) * RW + q];
            }
            fl[b] = nfl;
        }
    };

    u64 sweep0 = 0, o0 = 0;
    bool active = false, exhausted = false;
    u64 curw = 0, mbits = 0;
    u32 pend = 0, bits8 = 0, ifirst = 0;                               // bits of k-mers ifirst .. ifirst+7 (this step's)
    u32 good32 = 0;                                                    // records emitted by this thread
    auto emit = [&](u32 i, const u64 (&key)[W], bool fwd, float, bool no_markup, u32 eb) {
        // phase 1a's bit speaks for the qualities; without weights it has not looked at the bases, and a markup inside the
        // window makes the weight 0 (KmerReadUtils.h:207-217)
        if (!((bits8 >> (i - ifirst)) & 1u) || !no_markup) return;
        good32++;
        Rec<W, HASX> rec;
        rec.pack(key, fwd, (HASX && a.wts) ? a.wts[o0 + i] : 1.0f, eb);
        const u64 ph = place_hash<W>(key);
        const u32 group = part_of(ph, a.table.n_parts) >> a.table.group_shift;
        u32 own = lo;
        if (DIST != 0) {
            const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
            own = owner_of_fast(h, a.nranks, a.owner_magic);
        }
        const u32 bin = DIST == 2 ? own * n_parts + group : group;
        const u32 p = atomicAdd(&cnt[bin], 1u);
        if (p < sub_cap) {
            if (p - fl[bin] < R) {
                u64 *d = ring + ((size_t)bin * R + (p & (R - 1u))) * RW;
#pragma unroll
                for (int q = 0; q < RW; ++q) d[q] = rec.w[q];
            } else {
                u64 *d = sub_base(bin) + (size_t)p * RW;
#pragma unroll
                for (int q = 0; q < RW; ++q) st_hint64(d + q, rec.w[q], keep);
            }
        } else if (DIST != 2 || own == a.rank) {                       // full sub-region: straight into the table
            insert_record<W, HASX>(a.table, rec, lc.unique, lc.full, lc.probes);
            lc.direct++;
        } else {                                                       // full remote sub-region: the owner's overflow list
            const u32 op = a.stage.ovf_cap ? atomicAdd(&a.stage.ovf_count[own], 1u) : 0u;
            if (op < a.stage.ovf_cap) {
                u64 *d = a.stage.ovf_recs + ((size_t)own * a.stage.ovf_cap + op) * RW;
#pragma unroll
                for (int q = 0; q < RW; ++q) d[q] = rec.w[q];
            } else lost++;
        }
    };

    const u32 steps_per_round = a.ring_R ? max(1u, a.scatter_steps) : 1u;
    while (true) {
      // a ROUND = steps_per_round walker steps between two flushes (fewer block-wide barriers per record; the rings have to
      // hold a round's records of a bin)
      for (u32 sub = 0; sub < steps_per_round; ++sub) {
        if (!active && !exhausted) {                                   // next read of this thread
            while (sweep0 < a.n_reads) {
                const u64 r = sweep0 + in_sweep;
                sweep0 += stride;
                if (r >= a.n_reads) continue;
                const u64 b0 = a.read_off[r], b1 = a.read_off[r + 1];
                const u32 len = (u32)(b1 - b0);
                if (len < a.k || (a.discarded && a.discarded[r])) continue;
                o0 = b0;
                st.template begin<EXT>(a, b0, len);
                // "counted" bits of phase 1a: a 64-bit window (mask words curw, curw+1) plus the following word, which is
                // requested one whole step before it can be needed so that its latency never sits on the critical path
                curw = b0 >> 5;
                mbits = (u64)__ldg(&a.mask[curw]) | ((u64)__ldg(&a.mask[curw + 1]) << 32);
                pend = __ldg(&a.mask[curw + 2]);
                active = true;
                break;
            }
            if (!active) exhausted = true;
        }
        if (active) {                                                  // one step: up to 8 bases, up to 8 records
            ifirst = st.j + 1 >= a.k ? st.j + 1 - a.k : 0u;            // first k-mer this step can emit
            const u64 gb0 = o0 + ifirst;
            if ((gb0 >> 5) != curw) { mbits = (mbits >> 32) | ((u64)pend << 32); ++curw; }
            const u32 nxt = __ldg(&a.mask[curw + 2]);
            bits8 = (u32)(mbits >> (u32)(gb0 & 31ull)) & 0xffu;
            walker_step<W, false, EXT>(st, a, nullptr, emit);
            pend = nxt;
            if (st.j >= st.len) active = false;
        }
      }
        __syncthreads();
        if (R) flush(false);
        if (!__syncthreads_or((active || !exhausted) ? 1 : 0)) break;
    }
    if (R) flush(true);
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_bins; i += blockDim.x) {
        const u32 o = DIST == 2 ? i / n_parts : lo, g = DIST == 2 ? i - o * n_parts : i;
        a.stage.count[a.stage.cnt_index(o, g, blockIdx.x)] = cnt[i];
    }
    if (DIST == 2 && lost) atomicAdd(a.flags, lost);
    if (a.wts == nullptr) lc.good += good32;                           // (with weights phase 1a has seen the bases and counted)
    ctr_commit(a.ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU push path (DIST 2), sender side.  After phase 1 the part of the staging set that belongs to owner o holds
// o's records in n_parts x n_cta sub-regions.  k_push_plan gives every sub-region its place in one contiguous run per
// owner, ordered by group (so the owner's phase 2 can read group g of every source as one slice of the run);
// k_push_copy then writes the runs and their per-group offsets straight into the owners' receive buffers through
// peer pointers: the all-to-all of MPIAllToAllMessageBuffer (src/MPIBuffer.h:588-600,872-892) as coalesced stores over
// NVLink, with no send buffer, no count exchange and no routing pass on the receiver.
// ------------------------------------------------------------------------------------------------
struct PushPeers { u64 *recs[KMN_MAX_PUSH_RANKS]; u32 *meta[KMN_MAX_PUSH_RANKS]; };   // receive buffer + meta of rank o, slot of this rank

__global__ void __launch_bounds__(1024) k_push_plan(StageView st, u64 recv_cap, u32 *run_off, u32 *grp_off, u64 *flags)
{
    // one CTA per owner: exclusive prefix over its sub-regions in (group, cta) order
    const u32 o = blockIdx.x;
    if (o == st.me) return;
    __shared__ u64 carry;
    __shared__ u64 wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const u32 n = st.n_parts * st.n_cta;
    const u32 *count = st.count + (size_t)o * n;
    u32 *ro = run_off + (size_t)o * n;
    u32 *go = grp_off + (size_t)o * (st.n_parts + 1);
    for (u32 base = 0; base < n; base += blockDim.x) {
        const u32 p = base + threadIdx.x;
        u64 c = 0;
        if (p < n) { c = count[p]; if (c > st.sub_cap) c = st.sub_cap; }
        u64 v = c;
        const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, v, d); if (lane >= (u32)d) v += x; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            u64 x = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u64 y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (u32)d) x += y; }
            wsum[lane] = x;
        }
        __syncthreads();
        const u64 incl = v + (warp ? wsum[warp - 1] : 0) + carry;
        if (p < n) {
            u64 excl = incl - c;
            if (excl > recv_cap) excl = recv_cap;                    // the copy clamps at the capacity and reports it
            ro[p] = (u32)excl;
            if (p % st.n_cta == 0) go[p / st.n_cta] = (u32)excl;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        u64 tot = carry;
        if (tot > recv_cap) { atomicAdd(flags + 1, tot - recv_cap); tot = recv_cap; }
        go[st.n_parts] = (u32)tot;
    }
}

__global__ void __launch_bounds__(128, 16) k_push_copy(StageView st, PushPeers peers, u64 recv_cap, const u32 *run_off, const u32 *grp_off, u32 rw)
{
    const u32 n = st.n_parts * st.n_cta;
    const u32 lane = threadIdx.x & 31u;
    const u64 n_warps = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 total = (u64)st.n_owners * n;
    for (u64 e = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < total; e += n_warps) {
        const u32 o = (u32)(e / n), p = (u32)(e - (u64)o * n);
        if (o == st.me) continue;
        const u32 g = p / st.n_cta, c = p - g * st.n_cta;
        u64 cnt = st.count[(size_t)o * n + p];
        if (cnt > st.sub_cap) cnt = st.sub_cap;
        const u64 off = run_off[(size_t)o * n + p];
        if (off + cnt > recv_cap) cnt = recv_cap - off;
        const u64 *src = st.recs + st.sub_index(o, g, c) * st.sub_cap * rw;
        u64 *dst = peers.recs[o] + off * rw;
        const u64 words = cnt * rw;
        u64 i = lane;
        for (; i + 96 < words; i += 128) {                        // four independent 8-byte loads per lane in flight
            const u64 x0 = ld_nc64(src + i), x1 = ld_nc64(src + i + 32), x2 = ld_nc64(src + i + 64), x3 = ld_nc64(src + i + 96);
            dst[i] = x0; dst[i + 32] = x1; dst[i + 64] = x2; dst[i + 96] = x3;
        }
        for (; i < words; i += 32) dst[i] = ld_nc64(src + i);
    }
    // per-group offsets of this rank's run, for the receiver's work list
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < (u64)st.n_owners * (st.n_parts + 1); i += (u64)gridDim.x * blockDim.x) {
        const u32 o = (u32)(i / (st.n_parts + 1));
        if (o == st.me) continue;
        peers.meta[o][i - (u64)o * (st.n_parts + 1)] = grp_off[i];
    }
}

// ------------------------------------------------------------------------------------------------
// kmn_count_batch_2na: the reference keeps a read in memory as TwoBitSequence bytes (4 bases per byte, first base in the
// top two bits, src/TwoBitSequence.cpp:242-269) plus a list of markups for its non-ACGT bases (src/Sequence.h:372-380).
// A batch in that form crosses PCIe at a quarter of the ASCII size; these two kernels turn it back into the ASCII the
// walkers read.  One warp per read, a lane per packed byte.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_2na(const uint8_t *packed, const u64 *packed_off, const u64 *read_off, u64 n_reads, uint8_t *bases)
{
    const u32 lane = threadIdx.x & 31u;
    const u64 n_warps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_reads; r += n_warps) {
        const u64 o0 = read_off[r], len = read_off[r + 1] - o0, p0 = packed_off[r];
        const u64 nbytes = (len + 3) >> 2;
        for (u64 j = lane; j < nbytes; j += 32) {
            const u32 b = packed[p0 + j];
            const u64 at = o0 + 4 * j;
#pragma unroll
            for (u32 q = 0; q < 4; ++q)
                if (4 * j + q < len) bases[at + q] = (uint8_t)((0x54474341u >> (8u * ((b >> (6u - 2u * q)) & 3u))) & 0xffu);   // "ACGT"
        }
    }
}

__global__ void __launch_bounds__(256) k_apply_markups(const u64 *pos, const uint8_t *chr, u64 n, u64 total, uint8_t *bases)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if (pos[i] < total) bases[pos[i]] = chr[i];
}

// number of k-mer positions in reads [0,n): sum max(0, len-k+1) (discarded reads excluded)
__global__ void __launch_bounds__(256) k_count_positions(const u64 *read_off, const uint8_t *discarded, u64 n_reads, u32 k, u64 *total)
{
    u64 mine = 0;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (u64)gridDim.x * blockDim.x) {
        u64 len = read_off[r + 1] - read_off[r];
        if (len >= k && !(discarded && discarded[r])) mine += len - k + 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

// ------------------------------------------------------------------------------------------------
// phase 2 work list: chunk_start[e] = first chunk index of entry e (exclusive scan over the entries of k_build_entries),
// single CTA
// ------------------------------------------------------------------------------------------------
// coarse[c] = the entry that holds chunk c << COARSE_SHIFT: a ticket finds its entry with one load and a bisection over
// the few entries of its cell instead of a bisection over the whole list (which the whole CTA would wait for)
__global__ void k_build_worklist(const u32 *ent_cnt, u32 n_entries, u32 chunk, u64 *chunk_start, u64 *next_item, u32 *coarse)
{
    constexpr u32 IPT = 4;                       // entries per thread and step
    __shared__ u64 carry;
    __shared__ u64 wsum[32];
    if (threadIdx.x == 0) carry = 0;
    if (threadIdx.x < 8) next_item[threadIdx.x] = 0;          // one ticket counter per launch of a split phase 2
    __syncthreads();
    for (u32 base = 0; base < n_entries; base += blockDim.x * IPT) {
        const u32 p0 = base + threadIdx.x * IPT;
        u64 n[IPT], tot = 0;
#pragma unroll
        for (u32 q = 0; q < IPT; ++q) { n[q] = p0 + q < n_entries ? (ent_cnt[p0 + q] + chunk - 1) / chunk : 0; tot += n[q]; }
        u64 v = tot;
        const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (u32)o) v += t; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            u64 t = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u64 q = __shfl_up_sync(0xffffffffu, t, o); if (lane >= (u32)o) t += q; }
            wsum[lane] = t;
        }
        __syncthreads();
        const u64 incl = v + (warp ? wsum[warp - 1] : 0) + carry;
        u64 run = incl - tot;
#pragma unroll
        for (u32 q = 0; q < IPT; ++q) {
            if (p0 + q < n_entries) {
                chunk_start[p0 + q] = run;
                const u64 step = 1ull << COARSE_SHIFT;
                for (u64 m = (run + step - 1) & ~(step - 1); m < run + n[q]; m += step) coarse[m >> COARSE_SHIFT] = p0 + q;
            }
            run += n[q];
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_start[n_entries] = carry;
}

// phase-2 work list entries, group-major: for table group g first the n_cta local sub-regions of the staging set, then
// (multi-GPU) one run per peer rank inside the receive buffer that peer pushed its records of this round into.
// ent_ptr[e] = address of the entry's first record, ent_cnt[e] = its records.
// ------------------------------------------------------------------------------------------------
struct RecvView {
    const u64 *recs;      // [n_src][cap][RW]   this round's receive buffer (written by the peers over NVLink)
    const u32 *meta;      // mode 0: [n_src][n_parts + 1] record offset of every group inside the source's run (exclusive prefix)
                          // mode 1: [n_src][n_parts][n_cta] fill counters of the source's sub-regions
    u64 cap;              // records per source
    u32 n_src;            // ranks (0 = single GPU: no remote entries); the slot of this rank itself is unused
    u32 me;
    u32 mode;             // 0: one run per source, sorted by group (k_push_copy)
                          // 1: verbatim copy of the source's [cta][group][sub_cap] part of its staging set (copy engines),
                          //    followed by the source's overflow list for this rank (ovf_cap records, ungrouped)
    u32 ovf_cap;
    u64 part_recs;        // mode 1: records of the part proper (n_cta * n_parts * sub_cap)
    u64 meta_stride;      // u32 words of meta per source (mode 1: n_parts * n_cta counters + the overflow count)
};

__global__ void __launch_bounds__(256) k_build_entries(StageView st, RecvView rv, u32 rw, u64 *ent_ptr, u32 *ent_cnt)
{
    const u32 n_rem = rv.n_src > 1 ? (rv.n_src - 1) * (rv.mode == 1 ? st.n_cta : 1u) : 0;
    const u32 per_group = st.n_cta + n_rem;
    const u32 n_grouped = st.n_parts * per_group;
    const u32 n_entries = n_grouped + (rv.mode == 1 && rv.n_src > 1 ? rv.n_src - 1 : 0u);   // + one overflow list per peer
    const u32 lo = st.local_owner();
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += gridDim.x * blockDim.x) {
        if (e >= n_grouped) {                                      // ungrouped records (any group): still exact, just not L2-friendly
            u32 src = e - n_grouped;
            if (src >= rv.me) ++src;
            u32 n = rv.meta[(size_t)src * rv.meta_stride + (size_t)st.n_parts * st.n_cta];
            if (n > rv.ovf_cap) n = rv.ovf_cap;
            ent_cnt[e] = n;
            ent_ptr[e] = (u64)(rv.recs + ((size_t)src * rv.cap + rv.part_recs) * rw);
            continue;
        }
        const u32 g = e / per_group, j = e - g * per_group;
        if (j < st.n_cta) {
            u32 n = st.count[st.cnt_index(lo, g, j)];
            if (n > st.sub_cap) n = st.sub_cap;                    // the excess was inserted directly by phase 1
            ent_cnt[e] = n;
            ent_ptr[e] = (u64)(st.recs + st.sub_index(lo, g, j) * st.sub_cap * rw);
        } else if (rv.mode == 1) {
            const u32 jj = j - st.n_cta;
            u32 src = jj / st.n_cta;
            const u32 cta = jj - src * st.n_cta;
            if (src >= rv.me) ++src;
            u32 n = rv.meta[(size_t)src * rv.meta_stride + (size_t)g * st.n_cta + cta];
            if (n > st.sub_cap) n = st.sub_cap;
            ent_cnt[e] = n;
            ent_ptr[e] = (u64)(rv.recs + ((size_t)src * rv.cap + ((size_t)cta * st.n_parts + g) * st.sub_cap) * rw);
        } else {
            u32 src = j - st.n_cta;
            if (src >= rv.me) ++src;
            const u32 *m = rv.meta + (size_t)src * (st.n_parts + 1);
            const u32 o0 = m[g], o1 = m[g + 1];
            ent_cnt[e] = o1 - o0;
            ent_ptr[e] = (u64)(rv.recs + ((size_t)src * rv.cap + o0) * rw);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: phase 2 of the count pass.  Persistent CTAs take (entry, chunk) items in group order from an atomic ticket, so at
// any time the whole GPU works on at most ~2 neighbouring groups whose table slices (slice_bytes each) stay L2-resident.
// Replaces KmerSpectrum::append (src/KmerSpectrum.h:1578-1668) + KmerMapByKmerArrayPair insert/find
// (src/Kmer.h:1491-1544,3095-3110) + TrackingData::track (src/KmerTrackingData.h:427-448,517-529).
//
// The kernel is bound by dependent L2 round trips, not by instruction issue (ncu: 0.35 IPC per scheduler, long-scoreboard
// stalls), so for single-word keys every thread runs its INSERT_UNROLL records through the probe sequence in LOCKSTEP:
// one pass issues the next memory operation of every unresolved record (pair load, or CAS on an empty slot), the next
// pass consumes all the answers.  A chunk then costs as many round trips as its longest probe chain (2-3) instead of
// one chain after the other, and the records of the next chunk are already in flight while this one is resolved.
// ------------------------------------------------------------------------------------------------
template <int W, bool HASX>
__device__ __forceinline__ void track_extras(const TableView &t, u64 slot, float weight, u32 eb)
{
    if (HASX) {
        if (t.wsum) atomicAdd(&t.wsum[slot], weight);
        if (t.ext) {
            const u32 l = eb & 7u, rr = (eb >> 3) & 7u;
            if (l < 6) atomicAdd(&t.ext[slot * 12 + l], 1u);
            if (rr < 6) atomicAdd(&t.ext[slot * 12 + 6 + rr], 1u);
        }
    }
}

template <int W, bool HASX>
__global__ void __launch_bounds__(INSERT_TPB, KMN_INSERT_MIN_CTAS) k_insert_staged(TableView t, const u64 *ent_ptr, const u32 *ent_cnt, u32 n_entries,
                                                                                    const u64 *chunk_start, const u32 *coarse, u64 *next_item,
                                                                                    Counters *ctr, u32 split, u32 n_split)
{
    constexpr int RW = Rec<W, HASX>::RW;
    constexpr int U = INSERT_UNROLL;
    __shared__ u64 s_item;
    __shared__ u32 s_entry;
    // pre-aggregation of hot k-mers (single-word keys): when the records of a ticket repeat few keys (a tiny genome at
    // enormous depth, a repeat, an adapter) every record would fight for the same L2 atomic unit -- the L2 serialises
    // atomics per address.  Such tickets are first counted into a small shared-memory table, and only one update per
    // distinct key and chunk goes to the table in HBM, carrying the whole multiplicity (count, directionBias, weight,
    // extension counters); the count still saturates at 65535 (src/KmerTrackingData.h:427-448).
    constexpr int AGG = 512;                                           // slots of the shared table; a chunk holds INSERT_CHUNK records
    constexpr bool CAN_AGG = W == 1;
    __shared__ u64 agg_key[CAN_AGG ? AGG : 1];                         // ~key, 0 = empty
    __shared__ u32 agg_cnt[CAN_AGG ? AGG : 1], agg_dir[CAN_AGG ? AGG : 1];
    __shared__ float agg_w[CAN_AGG && HASX ? AGG : 1];
    __shared__ u32 agg_ext[CAN_AGG && HASX ? AGG * 12 : 1];
    __shared__ u32 s_hot;
    if (CAN_AGG) {
        for (u32 i = threadIdx.x; i < (u32)AGG; i += INSERT_TPB) {
            agg_key[i] = 0; agg_cnt[i] = 0; agg_dir[i] = 0;
            if (HASX) { agg_w[i] = 0.f; for (int q = 0; q < 12; ++q) agg_ext[i * 12 + q] = 0; }
        }
    }
    u64 n_unique = 0, n_full = 0, n_probes = 0;
    // this launch covers the items [item_lo, total_items) of the work list's `split`-th part (the multi-GPU rounds cut
    // phase 2 into several launches so that the small barrier kernels of the next round are not stuck behind it)
    const u64 all_items = chunk_start[n_entries];
    const u64 per_split = ((all_items + n_split - 1) / n_split + INSERT_GROUP - 1) / INSERT_GROUP * INSERT_GROUP;
    const u64 item_lo = min(all_items, per_split * split);
    const u64 total_items = min(all_items, item_lo + per_split);
    next_item += split;

    // chunk `item` of the work list: first record and record count (entry = hint, moved forward to the item's entry)
    auto locate = [&](u64 item, u32 &entry, const u64 *&src, u32 &cnt) {
        while (__ldg(&chunk_start[entry + 1]) <= item) ++entry;
        const u64 first = (item - __ldg(&chunk_start[entry])) * INSERT_CHUNK;
        const u64 n = __ldg(&ent_cnt[entry]);
        src = reinterpret_cast<const u64 *>(__ldg(&ent_ptr[entry])) + first * RW;
        cnt = (u32)(n - first < (u64)INSERT_CHUNK ? n - first : (u64)INSERT_CHUNK);
    };
    auto load = [&](const u64 *src, u32 cnt, Rec<W, HASX> (&r)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 idx = (u32)u * INSERT_TPB + threadIdx.x;
            if (idx < cnt) {
#pragma unroll
                for (int q = 0; q < RW; ++q) r[u].w[q] = ld_nc64(src + (size_t)idx * RW + q);
            }
        }
    };

    while (true) {
        // one ticket = INSERT_GROUP consecutive chunks; the entry of the first one comes from the coarse index (one load,
        // then a bisection over the entries of that cell), the following ones by stepping
        if (threadIdx.x == 0) {
            const u64 it = item_lo + atomicAdd(next_item, (u64)INSERT_GROUP);
            s_item = it;
            if (it < total_items) {
                const u64 cell = it >> COARSE_SHIFT, n_cells = (all_items + (1ull << COARSE_SHIFT) - 1) >> COARSE_SHIFT;
                u32 lo = __ldg(&coarse[cell]);
                u32 hi = cell + 1 < n_cells ? __ldg(&coarse[cell + 1]) + 1u : n_entries;
                while (hi - lo > 1) { const u32 mid = lo + ((hi - lo) >> 1); if (__ldg(&chunk_start[mid]) <= it) lo = mid; else hi = mid; }
                s_entry = lo;
            }
        }
        __syncthreads();
        const u64 item0 = s_item;
        u32 entry = s_entry;
        __syncthreads();
        if (item0 >= total_items) break;
        Rec<W, HASX> rec[U];
        const u64 *src; u32 cnt;
        locate(item0, entry, src, cnt);
        load(src, cnt, rec);
        bool hot = false;
        if constexpr (CAN_AGG) {
            // a sample of 32 records of the ticket: hot when at least a quarter of them share their key with another one
            if (threadIdx.x < 32) {
                const bool have = threadIdx.x < cnt;
                const u32 act = __ballot_sync(0xffffffffu, have);
                u32 dup = 0;
                if (have) { const u32 peers = __match_any_sync(act, rec[0].w[0] & ~1ull); dup = __popc(peers) > 1 ? 1u : 0u; }
                const u32 n_dup = __popc(__ballot_sync(0xffffffffu, dup != 0));
                if (threadIdx.x == 0) s_hot = (n_dup >= 8u) ? 1u : 0u;
            }
            __syncthreads();
            hot = s_hot != 0;
        }
#pragma unroll 1
        for (u32 g = 0; g < (u32)INSERT_GROUP; ++g) {
            const u64 item = item0 + g;
            if (item >= total_items) break;
            Rec<W, HASX> rec_n[U];
            u32 cnt_n = 0;
            if (g + 1 < (u32)INSERT_GROUP && item + 1 < total_items) { locate(item + 1, entry, src, cnt_n); load(src, cnt_n, rec_n); }

            if (CAN_AGG && hot) {
                if constexpr (CAN_AGG) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if ((u32)u * INSERT_TPB + threadIdx.x >= cnt) continue;
                        u64 key[1]; bool fwd; float weight; u32 eb;
                        rec[u].unpack(key, fwd, weight, eb);
                        const u64 want = ~key[0];
                        u32 h = (u32)(mix64(key[0]) >> 40) & (AGG - 1);
                        bool placed = false;
                        for (int pr = 0; pr < 8 && !placed; ++pr) {
                            u64 k = agg_key[h];
                            if (k == 0ull) k = atomicCAS(&agg_key[h], 0ull, want);
                            if (k == 0ull || k == want) placed = true; else h = (h + 1) & (AGG - 1);
                        }
                        if (placed) {
                            atomicAdd(&agg_cnt[h], 1u);
                            if (fwd) atomicAdd(&agg_dir[h], 1u);
                            if (HASX) {
                                if (t.wsum) atomicAdd(&agg_w[h], weight);
                                if (t.ext) {
                                    const u32 l = eb & 7u, rr = (eb >> 3) & 7u;
                                    if (l < 6) atomicAdd(&agg_ext[h * 12 + l], 1u);
                                    if (rr < 6) atomicAdd(&agg_ext[h * 12 + 6 + rr], 1u);
                                }
                            }
                        } else insert_record<W, HASX>(t, rec[u], n_unique, n_full, n_probes);      // too many distinct keys for the shared table
                    }
                    __syncthreads();
                    for (u32 i = threadIdx.x; i < (u32)AGG; i += INSERT_TPB) {
                        const u64 kk = agg_key[i];
                        if (kk == 0ull) continue;
                        u64 key[1] = {~kk};
                        const u64 ph = place_hash<1>(key);
                        u64 slot; u32 probes = 0;
                        const int r = table_insert<1>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key,
                                                      (u64)agg_cnt[i] | ((u64)agg_dir[i] << 32), &slot, &probes);
                        if (r < 0) n_full += agg_cnt[i];
                        else {
                            n_unique += (u64)r;
                            if (HASX) {
                                if (t.wsum) atomicAdd(&t.wsum[slot], agg_w[i]);
                                if (t.ext) for (int q = 0; q < 12; ++q) if (agg_ext[i * 12 + q]) atomicAdd(&t.ext[slot * 12 + q], agg_ext[i * 12 + q]);
                            }
                        }
                        agg_key[i] = 0; agg_cnt[i] = 0; agg_dir[i] = 0;
                        if (HASX) { agg_w[i] = 0.f; for (int q = 0; q < 12; ++q) agg_ext[i * 12 + q] = 0; }
                    }
                    __syncthreads();
                }
            } else if constexpr (W == 1) {
                // lockstep probing.  state: 0 done, 1 pair load in flight, 2/3 CAS on slot 0/1 of the pair in flight
                u64 want[U], k0[U], k1[U], add[U];
                Slot<1> *sbase[U];
                u32 s[U], pr[U], state[U], sat[U];
                float weight[U]; u32 eb[U];
                const u32 S = (u32)t.part_slots;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    state[u] = 0; pr[u] = 0; sat[u] = 0; k0[u] = k1[u] = 0; want[u] = 0; add[u] = 0; s[u] = 0; sbase[u] = nullptr; weight[u] = 0.f; eb[u] = 0;
                    if ((u32)u * INSERT_TPB + threadIdx.x < cnt) {
                        u64 key[1]; bool fwd;
                        rec[u].unpack(key, fwd, weight[u], eb[u]);
                        const u64 ph = place_hash<1>(key);
                        sbase[u] = reinterpret_cast<Slot<1> *>(t.slots) + (u64)part_of(ph, t.n_parts) * S;
                        s[u] = (u32)home_slot(ph, S) & ~1u;
                        want[u] = ~key[0];
                        add[u] = 1ull | ((u64)(fwd ? 1u : 0u) << 32);
                        state[u] = 1;
                        u64 v0, v1;
                        ld_pair32(sbase[u] + s[u], v0, k0[u], v1, k1[u]);
                        sat[u] = pair_sat(v0, v1);
                    }
                }
                while (true) {
                    bool pending = false;
#pragma unroll
                    for (int u = 0; u < U; ++u) {             // consume the answers
                        if (state[u] == 0) continue;
                        Slot<1> *pair = sbase[u] + s[u];
                        int hit = -1;                          // slot of the pair that now holds this key
                        bool fresh = false;
                        if (state[u] == 1) {
                            if (k0[u] == want[u]) hit = 0;
                            else if (k0[u] == 0) state[u] = 2;
                            else if (k1[u] == want[u]) hit = 1;
                            else if (k1[u] == 0) state[u] = 3;
                            else state[u] = 4;                 // both slots hold other keys: next pair
                        } else {                               // CAS answer is in k0
                            const u64 old = k0[u];
                            const int h = (int)state[u] - 2;
                            if (old == 0 || old == want[u]) { hit = h; fresh = old == 0; sat[u] = 0; }
                            else if (h == 0) {                 // slot 0 went to another key meanwhile; slot 1 as loaded before
                                if (k1[u] == want[u]) hit = 1;
                                else if (k1[u] == 0) state[u] = 3;
                                else state[u] = 4;
                            } else state[u] = 4;
                        }
                        if (hit >= 0) {
                            if (!((sat[u] >> hit) & 1u)) atomicAdd(&pair[hit].val, add[u]);
                            if (fresh) n_unique++;
                            n_probes += pr[u] + (u32)hit;
                            track_extras<1, HASX>(t, (u64)(pair - reinterpret_cast<Slot<1> *>(t.slots)) + (u64)hit, weight[u], eb[u]);
                            state[u] = 0;
                        } else if (state[u] == 4) {
                            pr[u] += 2;
                            s[u] = s[u] + 2 >= S ? 0 : s[u] + 2;
                            if (pr[u] >= S) { n_full++; state[u] = 0; } else state[u] = 1;
                        }
                        pending = pending || state[u] != 0;
                    }
                    if (!pending) break;
#pragma unroll
                    for (int u = 0; u < U; ++u) {             // issue the next operation of every unresolved record
                        if (state[u] == 1) { u64 v0, v1; ld_pair32(sbase[u] + s[u], v0, k0[u], v1, k1[u]); sat[u] = pair_sat(v0, v1); }
                        else if (state[u] >= 2) k0[u] = atomicCAS(&sbase[u][s[u] + (state[u] - 2)].k[0], 0ull, want[u]);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((u32)u * INSERT_TPB + threadIdx.x < cnt) {
                        u64 key[W]; bool fwd; float weight; u32 eb;
                        rec[u].unpack(key, fwd, weight, eb);
                        const u64 ph = place_hash<W>(key);
                        u64 slot; u32 probes = 0;
                        const int r = table_insert<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, 1ull | ((u64)(fwd ? 1u : 0u) << 32), &slot, &probes);
                        if (r < 0) { n_full++; continue; }
                        n_unique += (u64)r; n_probes += probes;
                        track_extras<W, HASX>(t, slot, weight, eb);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) rec[u] = rec_n[u];
            cnt = cnt_n;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
        n_probes += __shfl_xor_sync(0xffffffffu, n_probes, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
        if (n_probes) atomicAdd(&ctr->probe_steps, n_probes);
    }
}

// ------------------------------------------------------------------------------------------------
// K3': phase 2 for single-word keys without an extra record word (k <= 31, the FilterReads value kind): counting in
// SHARED MEMORY.  A scattered access to L2 costs the SM 1.5 (RED) to 3.5 (load, store) cycles per lane, a shared-memory
// access a tenth of that (bench/micro/lsu.cu), so the records of a table group are split once more -- by the 64 KB slice
// of the table their key lives in (k_slice_split, the ring / whole-sector mechanism of phase 1b with the slices of one
// group as bins) -- and k_count_slices then brings one slice at a time into shared memory, probes and counts there
// (LDS.128 probe, ATOMS add, ATOMS CAS to claim), and writes the slice back: the table only ever sees coalesced traffic.
// Same slot layout, same probe sequence (linear from the even slot below the home slot, wrapping inside the slice) and
// the same saturating count as table_insert, so k_insert_staged / table_find / the scans work on the same table.
// ------------------------------------------------------------------------------------------------
// ---- bulk asynchronous copies (TMA, cp.async.bulk) and the mbarriers that track them ----
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tKMN_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra KMN_DONE_%=;\n\tbra KMN_WAIT_%=;\n\tKMN_DONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, u32 bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_a(u32 bar_addr, u32 parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tKMN_WAITA_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra KMN_DONEA_%=;\n\tbra KMN_WAITA_%=;\n\tKMN_DONEA_%=:\n\t}"
                 ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(u32 bar_addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory"); }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ u64 lds64(u32 a) { u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void mbar_arrive(u64 *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


struct SplitArgs {
    TableView table;
    const u64 *ent_ptr;    // phase-2 work list entries, group-major (k_build_entries)
    const u32 *ent_cnt;
    u32 per_group;         // entries per group
    u32 n_groups;          // groups of this launch: [g0, g0 + n_groups); buf and cnt2 are indexed relative to g0
    u32 g0;
    u32 S;                 // a group's entries are cut into S parts; one CTA splits one (group, part) into private sub-runs
    u32 epp;               // entries per part
    u64 *buf;              // [n_groups][S][slices per group][cap2] records
    u32 *cnt2;             // [n_groups][S][slices per group] records of every sub-run
    u32 cap2;              // multiple of 4
    u32 ring_R;            // power of two >= 4
    u32 *ticket;
    Counters *ctr;
};

static constexpr int SPLIT_RPT = 4;        // records per thread and round
static constexpr int SPLIT_TPB = 1024;   // one CTA of 1024 threads per SM, or (SPLIT_TPB2) two CTAs of 512: while one waits at its
static constexpr int SPLIT_TPB2 = 512;   // round barriers the other bins

// The records of the item's entries come into shared memory as bulk copies of one round each (two buffers, the copy of
// the next round runs while this round is binned); threads take their records from there, so the kernel waits for memory
// only at the start of an item.
template <int TPB>
__global__ void __launch_bounds__(TPB, 2048 / TPB / 2) k_slice_split(SplitArgs a)
{
    constexpr int SPLIT_CHUNK = TPB * SPLIT_RPT;                       // records per round = per bulk copy
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ u32 s_item;
    __shared__ __align__(8) u64 bar_full[2];
    const u32 nb = 1u << a.table.group_shift;                          // bins = slices of one group
    const u32 n_pad = (nb + 31u) & ~31u;
    u64 *inbuf = reinterpret_cast<u64 *>(smem_raw);                    // [2][SPLIT_CHUNK] input records
    u32 *cnt = reinterpret_cast<u32 *>(inbuf + 2 * SPLIT_CHUNK);
    u32 *fl = cnt + n_pad;
    u64 *ring = reinterpret_cast<u64 *>(fl + n_pad);
    const u32 R = a.ring_R, cap2 = a.cap2;
    const u32 n_items = a.n_groups * a.S;
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    if (threadIdx.x == 0) { mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u32 ph0 = 0, ph1 = 0, q = 0;                                       // chunks are numbered through all items (buffer = number & 1)
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.ticket, 1u);
        for (u32 i = threadIdx.x; i < nb; i += TPB) { cnt[i] = 0; fl[i] = 0; }
        __syncthreads();
        const u32 item = s_item;
        if (item >= n_items) break;
        const u32 g = a.g0 + item / a.S, part = item % a.S;
        u64 *const obase = a.buf + (size_t)item * nb * cap2;
        auto flush = [&](bool all) {
            for (u32 b = threadIdx.x; b < nb; b += TPB) {
                u32 c = cnt[b];
                if (c > cap2) c = cap2;
                const u32 f = fl[b];
                if (c <= f) continue;
                u32 end, nfl;
                if (c - f > R) { end = f + R; nfl = c; }
                else { end = all ? c : (c & ~3u); nfl = end; }
                if (end <= f) continue;
                u64 *gdst = obase + (size_t)b * cap2;
                const u64 *rb = ring + (size_t)b * R;
                u32 pos = f;
                for (; pos < end && (pos & 3u); ++pos) gdst[pos] = rb[pos & (R - 1u)];
                for (; pos + 4u <= end; pos += 4u) st_sector(gdst + pos, rb + (pos & (R - 1u)));
                for (; pos < end; ++pos) gdst[pos] = rb[pos & (R - 1u)];
                fl[b] = nfl;
            }
        };
        // the item's records = the records of its entries, one after the other, in chunks of SPLIT_CHUNK that do not
        // cross an entry; every thread steps the consumer position (ce, co), thread 0 also the producer's (pe, po)
        const u32 e1 = min(g * a.per_group + (part + 1u) * a.epp, (g + 1u) * a.per_group);
        const u32 e0 = g * a.per_group + part * a.epp;
        // (the size of the entry after the current one is requested when the current one is entered, so that stepping to
        //  it never waits for memory between two block-wide barriers)
        u32 ce = e0, co = 0, cn = e0 < e1 ? __ldg(&a.ent_cnt[e0]) : 0u, cnn = e0 + 1u < e1 ? __ldg(&a.ent_cnt[e0 + 1u]) : 0u;
        u32 pe = e0, po = 0, pn = cn, pnn = cnn;
        auto skip = [&](u32 &e, u32 &o, u32 &n, u32 &nn) {             // first entry at or after e with records left
            while (e < e1) {
                if (o < n) return;
                ++e; o = 0; n = nn;
                nn = e + 1u < e1 ? __ldg(&a.ent_cnt[e + 1u]) : 0u;
            }
            n = 0;
        };
        u64 pptr = 0, pptr_n = 0;                                      // thread 0: address of entry pe / pe + 1
        if (threadIdx.x == 0) {
            pptr = e0 < e1 ? __ldg(&a.ent_ptr[e0]) : 0ull;
            pptr_n = e0 + 1u < e1 ? __ldg(&a.ent_ptr[e0 + 1u]) : 0ull;
        }
        auto produce = [&](u32 qq) {                                   // thread 0: the chunk at (pe, po) -> buffer qq & 1
            while (pe < e1 && po >= pn) {
                ++pe; po = 0; pn = pnn; pptr = pptr_n;
                if (pe + 1u < e1) { pnn = __ldg(&a.ent_cnt[pe + 1u]); pptr_n = __ldg(&a.ent_ptr[pe + 1u]); } else { pnn = 0; pptr_n = 0; }
            }
            if (pe >= e1) return false;
            const u32 n = min((u32)SPLIT_CHUNK, pn - po);
            // entries start on 32-byte boundaries (sub-regions hold multiples of 4 records); a bulk copy moves a multiple of 16 bytes
            const u32 bytes = ((n + 1u) & ~1u) * 8u;
            const u64 *src = reinterpret_cast<const u64 *>(pptr) + po;
            mbar_expect_tx(&bar_full[qq & 1u], bytes);
            bulk_g2s(inbuf + (size_t)(qq & 1u) * SPLIT_CHUNK, src, bytes, &bar_full[qq & 1u]);
            po += n;
            return true;
        };
        u32 q_issue = q;
        if (threadIdx.x == 0) { if (produce(q_issue)) ++q_issue; if (produce(q_issue)) ++q_issue; }
        skip(ce, co, cn, cnn);
        while (ce < e1) {
            const u32 n = min((u32)SPLIT_CHUNK, cn - co);
            if (q & 1u) { mbar_wait(&bar_full[1], ph1); ph1 ^= 1u; } else { mbar_wait(&bar_full[0], ph0); ph0 ^= 1u; }
            const u64 *in = inbuf + (size_t)(q & 1u) * SPLIT_CHUNK;
            u64 rec[SPLIT_RPT];
#pragma unroll
            for (int u = 0; u < SPLIT_RPT; ++u) {
                const u32 idx = (u32)u * TPB + threadIdx.x;
                rec[u] = idx < n ? in[idx] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < SPLIT_RPT; ++u) {
                const u32 idx = (u32)u * TPB + threadIdx.x;
                if (idx >= n) continue;
                const u64 ph = mix64(rec[u] & ~1ull);
                const u32 b = part_of(ph, a.table.n_parts) & (nb - 1u);
                const u32 p = atomicAdd(&cnt[b], 1u);
                if (p < cap2) {
                    if (p - fl[b] < R) ring[(size_t)b * R + (p & (R - 1u))] = rec[u];
                    else obase[(size_t)b * cap2 + p] = rec[u];
                } else {                                               // sub-run full (skewed input): straight into the table
                    Rec<1, false> r; r.w[0] = rec[u];
                    insert_record<1, false>(a.table, r, lc.unique, lc.full, lc.probes);
                    lc.direct++;
                }
            }
            __syncthreads();                                           // the round's records are in the rings; its input buffer is free
            if (threadIdx.x == 0 && produce(q_issue)) ++q_issue;       // (the chunk after next, into the buffer just read)
            flush(false);
            __syncthreads();
            co += n; ++q;
            skip(ce, co, cn, cnn);
        }
        flush(true);
        __syncthreads();
        for (u32 i = threadIdx.x; i < nb; i += TPB) a.cnt2[(size_t)item * nb + i] = min(cnt[i], cap2);
        __syncthreads();
    }
    ctr_commit(a.ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// k_slice_split2: the same second radix pass for two-word keys (33 <= k <= 63: records of two words, the strand flag in
// bit 0 of the second); counted by k_count_slices_w2.  One GPU (KMN_SMEM_COUNT_W2=0 turns both off).
// ------------------------------------------------------------------------------------------------
static constexpr int SPLIT2_RPT = 2;       // records per thread and round: 2048 records = 32 KB per bulk copy
// K32: k = 32 -- one key word that leaves no room for the strand flag, which travels in a second record word
template <bool K32>
__global__ void __launch_bounds__(SPLIT_TPB, 1) k_slice_split2(SplitArgs a)
{
    constexpr int TPB = SPLIT_TPB, CHUNK = TPB * SPLIT2_RPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ u32 s_item;
    __shared__ __align__(8) u64 bar_full[2];
    const u32 nb = 1u << a.table.group_shift;
    const u32 n_pad = (nb + 31u) & ~31u;
    u64 *inbuf = reinterpret_cast<u64 *>(smem_raw);                    // [2][CHUNK][2] input records
    u32 *cnt = reinterpret_cast<u32 *>(inbuf + 2 * CHUNK * 2);
    u32 *fl = cnt + n_pad;
    u64 *ring = reinterpret_cast<u64 *>(fl + n_pad);                   // [nb][R][2]
    const u32 R = a.ring_R, cap2 = a.cap2;
    const u32 n_items = a.n_groups * a.S;
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    if (threadIdx.x == 0) { mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u32 ph0 = 0, ph1 = 0, q = 0;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.ticket, 1u);
        for (u32 i = threadIdx.x; i < nb; i += TPB) { cnt[i] = 0; fl[i] = 0; }
        __syncthreads();
        const u32 item = s_item;
        if (item >= n_items) break;
        const u32 g = a.g0 + item / a.S, part = item % a.S;
        u64 *const obase = a.buf + (size_t)item * nb * cap2 * 2;
        auto flush = [&](bool all) {
            for (u32 b = threadIdx.x; b < nb; b += TPB) {
                u32 c = cnt[b];
                if (c > cap2) c = cap2;
                const u32 f = fl[b];
                if (c <= f) continue;
                u32 end, nfl;
                if (c - f > R) { end = f + R; nfl = c; }
                else { end = all ? c : (c & ~1u); nfl = end; }         // whole sectors = pairs of records
                if (end <= f) continue;
                u64 *gdst = obase + (size_t)b * cap2 * 2;
                const u64 *rb = ring + (size_t)b * R * 2;
                u32 pos = f;
                for (; pos < end && (pos & 1u); ++pos) { gdst[(size_t)pos * 2] = rb[(size_t)(pos & (R - 1u)) * 2]; gdst[(size_t)pos * 2 + 1] = rb[(size_t)(pos & (R - 1u)) * 2 + 1]; }
                for (; pos + 2u <= end; pos += 2u) st_sector(gdst + (size_t)pos * 2, rb + (size_t)(pos & (R - 1u)) * 2);
                for (; pos < end; ++pos) { gdst[(size_t)pos * 2] = rb[(size_t)(pos & (R - 1u)) * 2]; gdst[(size_t)pos * 2 + 1] = rb[(size_t)(pos & (R - 1u)) * 2 + 1]; }
                fl[b] = nfl;
            }
        };
        const u32 e1 = min(g * a.per_group + (part + 1u) * a.epp, (g + 1u) * a.per_group);
        const u32 e0 = g * a.per_group + part * a.epp;
        u32 ce = e0, co = 0, cn = e0 < e1 ? __ldg(&a.ent_cnt[e0]) : 0u, cnn = e0 + 1u < e1 ? __ldg(&a.ent_cnt[e0 + 1u]) : 0u;
        u32 pe = e0, po = 0, pn = cn, pnn = cnn;
        auto skip = [&](u32 &e, u32 &o, u32 &n, u32 &nn) {
            while (e < e1) {
                if (o < n) return;
                ++e; o = 0; n = nn;
                nn = e + 1u < e1 ? __ldg(&a.ent_cnt[e + 1u]) : 0u;
            }
            n = 0;
        };
        u64 pptr = 0, pptr_n = 0;
        if (threadIdx.x == 0) {
            pptr = e0 < e1 ? __ldg(&a.ent_ptr[e0]) : 0ull;
            pptr_n = e0 + 1u < e1 ? __ldg(&a.ent_ptr[e0 + 1u]) : 0ull;
        }
        auto produce = [&](u32 qq) {
            while (pe < e1 && po >= pn) {
                ++pe; po = 0; pn = pnn; pptr = pptr_n;
                if (pe + 1u < e1) { pnn = __ldg(&a.ent_cnt[pe + 1u]); pptr_n = __ldg(&a.ent_ptr[pe + 1u]); } else { pnn = 0; pptr_n = 0; }
            }
            if (pe >= e1) return false;
            const u32 n = min((u32)CHUNK, pn - po);
            const u32 bytes = n * 16u;                                 // (records of 16 bytes: always a multiple of 16)
            const u64 *src = reinterpret_cast<const u64 *>(pptr) + (size_t)po * 2;
            mbar_expect_tx(&bar_full[qq & 1u], bytes);
            bulk_g2s(inbuf + (size_t)(qq & 1u) * CHUNK * 2, src, bytes, &bar_full[qq & 1u]);
            po += n;
            return true;
        };
        u32 q_issue = q;
        if (threadIdx.x == 0) { if (produce(q_issue)) ++q_issue; if (produce(q_issue)) ++q_issue; }
        skip(ce, co, cn, cnn);
        while (ce < e1) {
            const u32 n = min((u32)CHUNK, cn - co);
            if (q & 1u) { mbar_wait(&bar_full[1], ph1); ph1 ^= 1u; } else { mbar_wait(&bar_full[0], ph0); ph0 ^= 1u; }
            const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(inbuf + (size_t)(q & 1u) * CHUNK * 2);
#pragma unroll
            for (int u = 0; u < SPLIT2_RPT; ++u) {
                const u32 idx = (u32)u * TPB + threadIdx.x;
                if (idx >= n) continue;
                const ulonglong2 rec = in[idx];
                u64 key[2] = {rec.x, rec.y & ~1ull};
                const u64 ph = K32 ? mix64(rec.x) : place_hash<2>(key);
                const u32 b = part_of(ph, a.table.n_parts) & (nb - 1u);
                const u32 p = atomicAdd(&cnt[b], 1u);
                if (p < cap2) {
                    u64 *d = (p - fl[b] < R) ? ring + ((size_t)b * R + (p & (R - 1u))) * 2 : obase + ((size_t)b * cap2 + p) * 2;
                    d[0] = rec.x; d[1] = rec.y;
                } else {                                               // sub-run full (skewed input): straight into the table
                    if (K32) { Rec<1, true> r; r.w[0] = rec.x; r.w[1] = rec.y; insert_record<1, true>(a.table, r, lc.unique, lc.full, lc.probes); }
                    else { Rec<2, false> r; r.w[0] = rec.x; r.w[1] = rec.y; insert_record<2, false>(a.table, r, lc.unique, lc.full, lc.probes); }
                    lc.direct++;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0 && produce(q_issue)) ++q_issue;
            flush(false);
            __syncthreads();
            co += n; ++q;
            skip(ce, co, cn, cnn);
        }
        flush(true);
        __syncthreads();
        for (u32 i = threadIdx.x; i < nb; i += TPB) a.cnt2[(size_t)item * nb + i] = min(cnt[i], cap2);
        __syncthreads();
    }
    ctr_commit(a.ctr, lc);
}

static constexpr int COUNT_TPB = 256;
static constexpr int COUNT_U = 4;          // records per thread and batch
static constexpr int COUNT_MAX_S = 8;      // sub-runs per slice (SplitArgs::S)

__global__ void __launch_bounds__(COUNT_TPB, 3) k_count_slices(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, u32 n_groups,
                                                              u32 *ticket, Counters *ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Slot<1> *sl = reinterpret_cast<Slot<1> *>(smem_raw);               // one slice of the table
    __shared__ u32 s_pre[2][32];                                       // exclusive prefix of the S sub-run sizes of this / the next slice
    const u32 nb = 1u << t.group_shift;
    const u32 SL = (u32)t.part_slots;
    u64 n_unique = 0, n_full = 0;
    // slices are dealt round-robin (the work per slice is uniform); the sub-run sizes of a CTA's next slice are requested
    // while it works on the current one (warp 0: one lane per sub-run, prefix by shuffles; s_pre[][S] = total)
    auto load_pre = [&](u32 pi, u32 *dst) {
        if (threadIdx.x < 32) {
            u32 c = (threadIdx.x < S && pi < t.n_parts) ? cnt2[((size_t)(pi >> t.group_shift) * S + threadIdx.x) * nb + (pi & (nb - 1u))] : 0u;
            u32 v = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)threadIdx.x >= d) v += x; }
            if (threadIdx.x < COUNT_MAX_S + 1u) dst[threadIdx.x] = v - c;
        }
    };
    load_pre(blockIdx.x, s_pre[0]);
    u32 it = 0;
    for (u32 pi = blockIdx.x; pi < t.n_parts; pi += gridDim.x, ++it) {
        __syncthreads();                                               // s_pre[it & 1] is complete; the previous slice has left shared memory
        load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);
        u32 pre[COUNT_MAX_S];                                          // start of sub-run p in the slice's concatenated records
#pragma unroll
        for (int p = 0; p < COUNT_MAX_S; ++p) pre[p] = (u32)p < S ? s_pre[it & 1u][p] : 0xffffffffu;
        const u32 total = s_pre[it & 1u][S];
        if (total == 0) continue;                                      // nothing staged for this slice in this drain
        const u32 g = pi >> t.group_shift, j = pi & (nb - 1u);
        const u64 *run0 = buf + (((size_t)g * S) * nb + j) * cap2;     // sub-run p starts at run0 + p * nb * cap2
        const size_t run_stride = (size_t)nb * cap2;
        auto fetch = [&](u32 v0, u64 (&r)[COUNT_U], u32 &have) {       // records v0 + u * TPB + tid of the slice's concatenated sub-runs
            have = 0;
#pragma unroll
            for (int u = 0; u < COUNT_U; ++u) {
                const u32 v = v0 + (u32)u * COUNT_TPB + threadIdx.x;
                if (v < total) {
                    u32 p = 0, start = 0;
#pragma unroll
                    for (int q = 1; q < COUNT_MAX_S; ++q) if (v >= pre[q]) { p = (u32)q; start = pre[q]; }
                    r[u] = ld_nc64(run0 + (size_t)p * run_stride + (v - start));
                    have |= 1u << u;
                }
            }
        };
        u64 cur[COUNT_U], nxt[COUNT_U];
        u32 hc = 0, hn = 0;
        fetch(0, cur, hc);                                             // in flight while the slice is loaded
        uint4 *gsl = reinterpret_cast<uint4 *>(reinterpret_cast<Slot<1> *>(t.slots) + (size_t)pi * SL);
        uint4 *ssl = reinterpret_cast<uint4 *>(sl);
        for (u32 i = threadIdx.x; i < SL; i += COUNT_TPB) ssl[i] = gsl[i];
        __syncthreads();
        for (u32 v0 = 0; v0 < total; v0 += COUNT_TPB * COUNT_U) {
            if (v0 + COUNT_TPB * COUNT_U < total) fetch(v0 + COUNT_TPB * COUNT_U, nxt, hn); else hn = 0;
#pragma unroll
            for (int u = 0; u < COUNT_U; ++u) {
                if (!((hc >> u) & 1u)) continue;
                const u64 rec = cur[u];
                const u64 key1 = rec & ~1ull, want = ~key1;
                u64 key[1] = {key1};
                const u64 ph = place_hash<1>(key);
                u32 s = (u32)home_slot(ph, SL) & ~1u;
                u32 probes = 0;
                for (; probes < SL; ++probes) {
                    const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(&sl[s]);      // {val, key}
                    u64 k = x.y;
                    bool sat = (u32)x.x >= MAX_COUNT;
                    if (k == 0ull) {
                        k = atomicCAS(&sl[s].k[0], 0ull, want);
                        if (k == 0ull) { n_unique++; k = want; }
                        sat = false;
                    }
                    if (k == want) {
                        // count in the low word, directionBias in the high word: two native 32-bit shared atomics (a 64-bit
                        // shared atomicAdd compiles to a compare-and-swap loop); no carry ever crosses the words
                        if (!sat) {
                            u32 *w32 = reinterpret_cast<u32 *>(&sl[s].val);
                            atomicAdd(w32, 1u);
                            if (rec & 1ull) atomicAdd(w32 + 1, 1u);
                        }
                        break;
                    }
                    s = s + 1u == SL ? 0u : s + 1u;
                }
                if (probes >= SL) n_full++;
            }
#pragma unroll
            for (int u = 0; u < COUNT_U; ++u) cur[u] = nxt[u];
            hc = hn;
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < SL; i += COUNT_TPB) gsl[i] = ssl[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_slices with the bulk-copy engine (TMA, cp.async.bulk): the slice, the record chunks and the write-back move
// between global and shared memory as asynchronous bulk copies tracked by mbarriers, issued by one thread, so no thread
// holds loads in registers and the next chunk of records streams in while the current one is counted.
//   shared memory per CTA: the slice (part_slots * 16 B) + two record buffers of COUNT_CHUNK records
// ------------------------------------------------------------------------------------------------
static constexpr int COUNT3_TPB = 512;
static constexpr int COUNT_CHUNK = 1024;   // records per bulk copy (8 KB)
static constexpr int COUNT_NBUF = 4;       // record buffers: a chunk is requested COUNT_NBUF chunks before it is counted


__global__ void __launch_bounds__(COUNT3_TPB, 2) k_count_slices_tma(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, Counters *ctr, u32 slice0, u32 n_sl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 SL = (u32)t.part_slots;
    Slot<1> *sl = reinterpret_cast<Slot<1> *>(smem_raw);               // one slice of the table
    u64 *rbuf = reinterpret_cast<u64 *>(smem_raw + (size_t)SL * 16);   // [COUNT_NBUF][COUNT_CHUNK] record chunks
    __shared__ __align__(8) u64 bar_slice, bar_full[COUNT_NBUF], bar_empty[COUNT_NBUF];
    __shared__ u32 s_pre[2][32];                                       // exclusive prefix of the sub-run sizes of this / the next slice
    const u32 nb = 1u << t.group_shift;
    const u32 lane = threadIdx.x & 31u;
    u64 n_unique = 0, n_full = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar_slice, 1);
        for (int b = 0; b < COUNT_NBUF; ++b) {
            mbar_init(&bar_full[b], 1);                                // completed by the bulk copy's bytes
            mbar_init(&bar_empty[b], COUNT3_TPB / 32);                 // one arrival per warp
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto load_pre = [&](u32 pi, u32 *dst) {
        if (threadIdx.x < 32) {
            u32 c = (threadIdx.x < S && pi < n_sl) ? cnt2[((size_t)(pi >> t.group_shift) * S + threadIdx.x) * nb + (pi & (nb - 1u))] : 0u;
            u32 v = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)threadIdx.x >= d) v += x; }
            if (threadIdx.x < COUNT_MAX_S + 1u) dst[threadIdx.x] = v - c;
        }
    };
    load_pre(blockIdx.x, s_pre[0]);
    // parities of the next completion to wait for (one bit per buffer); the chunks of all slices are numbered through
    // (buffer = number % COUNT_NBUF)
    u32 it = 0, ph_slice = 0, ph_full = 0, ph_empty = 0, qn = 0;
    const u32 sl_addr = smem_u32(sl), sl_end = sl_addr + SL * 16u;
    for (u32 pi = blockIdx.x; pi < n_sl; pi += gridDim.x, ++it) {
        __syncthreads();                                               // s_pre[it & 1] complete; every thread has left the previous slice
        load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);
        const u32 *pre = s_pre[it & 1u];
        const u32 total = pre[S];
        if (total == 0) continue;                                      // nothing staged for this slice in this drain
        const u32 g = pi >> t.group_shift, j = pi & (nb - 1u);
        const u64 *run0 = buf + (((size_t)g * S) * nb + j) * cap2;     // sub-run p starts at run0 + p * nb * cap2
        const size_t run_stride = (size_t)nb * cap2;
        Slot<1> *gsl = reinterpret_cast<Slot<1> *>(t.slots) + (size_t)(slice0 + pi) * SL;
        // the slice's records as chunks of at most COUNT_CHUNK records that do not cross a sub-run: an iterator every thread
        // advances in step (sub-run p, offset o)
        u32 cp = 0, co = 0;                                            // consumer position
        u32 pp = 0, po = 0;                                            // producer position (thread 0), one chunk ahead
        auto skip_empty = [&](u32 &p, u32 &o) { while (p < S && o >= pre[p + 1] - pre[p]) { ++p; o = 0; } };
        skip_empty(cp, co);
        pp = cp; po = co;
        u32 q_issue = qn;                                              // number of the next chunk to issue (thread 0; == qn between slices)
        auto produce = [&]() {                                         // thread 0 only
            skip_empty(pp, po);
            if (pp >= S) return;
            const u32 b = q_issue % COUNT_NBUF;
            // the buffer's previous chunk (COUNT_NBUF numbers back) has been consumed by every warp
            if (q_issue >= COUNT_NBUF) { mbar_wait(&bar_empty[b], (ph_empty >> b) & 1u); ph_empty ^= 1u << b; }
            const u32 n = min((u32)COUNT_CHUNK, pre[pp + 1] - pre[pp] - po);
            const u32 bytes = ((n + 1u) & ~1u) * 8u;                   // bulk copies move multiples of 16 bytes (cap2 is a multiple of 4)
            mbar_expect_tx(&bar_full[b], bytes);
            bulk_g2s(rbuf + (size_t)b * COUNT_CHUNK, run0 + (size_t)pp * run_stride + po, bytes, &bar_full[b]);
            po += n;
            ++q_issue;
        };
        if (threadIdx.x == 0) {
            bulk_wait_read();                                          // the previous slice has been read out of shared memory
            mbar_expect_tx(&bar_slice, SL * 16u);
            bulk_g2s(sl, gsl, SL * 16u, &bar_slice);
            for (int b = 0; b < COUNT_NBUF; ++b) produce();
        }
        mbar_wait(&bar_slice, ph_slice);
        ph_slice ^= 1u;
        while (cp < S) {
            const u32 n = min((u32)COUNT_CHUNK, pre[cp + 1] - pre[cp] - co);
            const u32 b = qn % COUNT_NBUF;
            mbar_wait(&bar_full[b], (ph_full >> b) & 1u);
            ph_full ^= 1u << b;
            const u64 *recs = rbuf + (size_t)b * COUNT_CHUNK;
            for (u32 idx = threadIdx.x; idx < n; idx += COUNT3_TPB) {
                const u64 rec = recs[idx];
                const u64 want = ~(rec & ~1ull);
                const u64 ph = mix64(rec & ~1ull);
                u32 addr = sl_addr + (((u32)(((u64)(u32)ph * SL) >> 32)) & ~1u) * 16u;
                u32 left = SL;
                while (true) {
                    u64 v, k;
                    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(k) : "r"(addr));      // {val, key}
                    bool hit = k == want;
                    if (!hit && k == 0ull) {
                        u64 old;
                        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                        if (old == 0ull) n_unique++;
                        hit = old == 0ull || old == want;
                        v = 0;
                    }
                    if (hit) {
                        // count in the low word, directionBias in the high word: two native 32-bit shared atomics (a 64-bit
                        // shared atomicAdd compiles to a compare-and-swap loop); no carry ever crosses the words
                        if ((u32)v < MAX_COUNT) {
                            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                            if (rec & 1ull) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + 4u) : "memory");
                        }
                        break;
                    }
                    addr += 16u;
                    if (addr == sl_end) addr = sl_addr;
                    if (--left == 0u) { n_full++; break; }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[b]);                 // this warp is done with buffer b
            co += n; ++qn;
            skip_empty(cp, co);
            if (threadIdx.x == 0) produce();                           // refill the buffer just read, COUNT_NBUF chunks ahead (waits for its release)
        }
        // (the releases of the slice's last chunks are waited for by the first chunks of the next slice: chunk q always
        //  waits for the release of chunk q - COUNT_NBUF, across slices)
        __syncthreads();                                               // every warp has counted its records: the slice is final
        if (threadIdx.x == 0) {
            fence_async_smem();                                        // the counts written by all threads are visible to the bulk store
            bulk_s2g(gsl, sl, SL * 16u);
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_slices_ws: the same bulk-copy pipeline with a PRODUCER warp (one lane issues the slice load, the record chunks
// and the write-back) and 16 CONSUMER warps whose lanes never wait for each other's probe sequences: a lane that has
// counted its record takes the next unclaimed record of its warp's share of the current chunk (ballot + popc give the
// idle lanes consecutive records), so a warp-wide probe step always works on 32 live records instead of on the few
// stragglers of a batch (ncu on k_count_slices_tma: 150 warp instructions per 32 records, most of them issued for
// partially finished batches; the probe loop runs E[max of 32 probe lengths] ~ 5.5 times per batch against 1.85 per record).
//   chunk q of a CTA (numbered through all its slices) lives in buffer q % NBUF; its `full` barrier completes for the
//   (q / NBUF)-th time when its bytes have arrived, its `empty` barrier when every consumer warp has taken its share.
// ------------------------------------------------------------------------------------------------
static constexpr int COUNTW_CONSUMERS = 16;                            // consumer warps
static constexpr int COUNTW_TPB = (COUNTW_CONSUMERS + 1) * 32;
#ifndef KMN_COUNTW_CHUNK
#define KMN_COUNTW_CHUNK 1024
#endif
#ifndef KMN_COUNTW_NBUF
#define KMN_COUNTW_NBUF 4
#endif
static constexpr int COUNTW_CHUNK = KMN_COUNTW_CHUNK;                  // records per bulk copy
static constexpr int COUNTW_NBUF = KMN_COUNTW_NBUF;
static constexpr int COUNTW_SHARE = COUNTW_CHUNK / COUNTW_CONSUMERS;   // records of a chunk that belong to one consumer warp

// BALLOT: lane-persistent consumers (a lane that has counted its record takes the warp's next one); otherwise every lane
// counts the records idx = lane, lane + 32, .. of its warp's share one after the other (fewer instructions per record
// although the warp idles on the longest probe sequence of every batch of 32: ncu, 166 -> 204 warp instructions per 32
// records with the ballot loop)
template <int BALLOT>
__global__ void __launch_bounds__(COUNTW_TPB, 2) k_count_slices_ws(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, Counters *ctr, u32 slice0, u32 n_sl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 SL = (u32)t.part_slots;
    Slot<1> *sl = reinterpret_cast<Slot<1> *>(smem_raw);               // one slice of the table
    u64 *rbuf = reinterpret_cast<u64 *>(smem_raw + (size_t)SL * 16);   // [COUNTW_NBUF][COUNTW_CHUNK] record chunks
    __shared__ __align__(8) u64 bar_slice, bar_full[COUNTW_NBUF], bar_empty[COUNTW_NBUF];
    __shared__ u32 s_pre[2][32];                                       // exclusive prefix of the sub-run sizes of this / the next slice
    __shared__ u32 buf_n[COUNTW_NBUF];                                 // records of the chunk in every buffer
    const u32 nb = 1u << t.group_shift;
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool producer = warp == COUNTW_CONSUMERS;
    const u32 lt = (1u << lane) - 1u;
    u64 n_unique = 0, n_full = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar_slice, 1);
        for (int b = 0; b < COUNTW_NBUF; ++b) {
            mbar_init(&bar_full[b], 1);                                // completed by the bulk copy's bytes
            mbar_init(&bar_empty[b], COUNTW_CONSUMERS);                // one arrival per consumer warp
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto load_pre = [&](u32 pi, u32 *dst) {                            // warp 0
        u32 c = (lane < S && pi < n_sl) ? cnt2[((size_t)(pi >> t.group_shift) * S + lane) * nb + (pi & (nb - 1u))] : 0u;
        u32 v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v += x; }
        if (lane < COUNT_MAX_S + 1u) dst[lane] = v - c;
    };
    if (warp == 0) load_pre(blockIdx.x, s_pre[0]);
    u32 it = 0, ph_slice = 0, qn = 0;                                  // qn: number of the slice's first chunk
    // shared-window addresses once, as opaque integers (otherwise the generic-to-shared conversion -- a read of the CTA-id
    // special register -- is rematerialised inside the probe loop)
    auto opaque = [](u32 x) { u32 y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
    const u32 sl_addr = opaque(smem_u32(sl)), sl_end = opaque(sl_addr + SL * 16u);
    const u32 rbuf_a = opaque(smem_u32(rbuf)), full_a = opaque(smem_u32(&bar_full[0])), empty_a = opaque(smem_u32(&bar_empty[0])),
              bufn_a = opaque(smem_u32(&buf_n[0])), slice_bar_a = opaque(smem_u32(&bar_slice));
    u32 n_unique32 = 0, n_full32 = 0;
    Slot<1> *prev_gsl = nullptr;                                       // producer: slice waiting for its write-back
    for (u32 pi = blockIdx.x; pi < n_sl; pi += gridDim.x, ++it) {
        __syncthreads();                                               // s_pre[it & 1] complete; every consumer has left the previous slice
        if (warp == 0) load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);
        const u32 *pre = s_pre[it & 1u];
        const u32 total = pre[S];
        if (total == 0) continue;                                      // nothing staged for this slice in this drain
        u32 n_chunks = 0;
        for (u32 p = 0; p < S; ++p) n_chunks += (pre[p + 1] - pre[p] + COUNTW_CHUNK - 1u) / COUNTW_CHUNK;
        Slot<1> *gsl = reinterpret_cast<Slot<1> *>(t.slots) + (size_t)(slice0 + pi) * SL;
        if (producer) {
            if (lane == 0) {
                const u32 g = pi >> t.group_shift, j = pi & (nb - 1u);
                const u64 *run0 = buf + (((size_t)g * S) * nb + j) * cap2;     // sub-run p starts at run0 + p * nb * cap2
                const size_t run_stride = (size_t)nb * cap2;
                if (prev_gsl) bulk_s2g(prev_gsl, sl, SL * 16u);        // (the consumers fenced their counts before the barrier)
                prev_gsl = gsl;
                u32 pp = 0, po = 0, q = qn;
                auto issue = [&]() {
                    while (pp < S && po >= pre[pp + 1] - pre[pp]) { ++pp; po = 0; }
                    const u32 b = q % COUNTW_NBUF;
                    if (q >= COUNTW_NBUF) mbar_wait(&bar_empty[b], ((q / COUNTW_NBUF) - 1u) & 1u);   // chunk q - NBUF has been taken by every warp
                    const u32 n = min((u32)COUNTW_CHUNK, pre[pp + 1] - pre[pp] - po);
                    const u32 bytes = ((n + 1u) & ~1u) * 8u;           // bulk copies move multiples of 16 bytes (cap2 is a multiple of 4)
                    buf_n[b] = n;
                    mbar_expect_tx(&bar_full[b], bytes);
                    bulk_g2s(rbuf + (size_t)b * COUNTW_CHUNK, run0 + (size_t)pp * run_stride + po, bytes, &bar_full[b]);
                    po += n; ++q;
                };
                // the first chunks need no free slice; then the old slice has to be out of shared memory before the new one lands
                const u32 first = min(n_chunks, (u32)COUNTW_NBUF);
                for (u32 i = 0; i < first; ++i) issue();
                bulk_wait_read();
                mbar_expect_tx(&bar_slice, SL * 16u);
                bulk_g2s(sl, gsl, SL * 16u, &bar_slice);
                for (u32 i = first; i < n_chunks; ++i) issue();
            }
        } else if (BALLOT == 2) {
            // plain consumers, two records of a lane in flight at once: their probe loads are issued back to back, and the
            // warp leaves the loop after the longest of 64 probe sequences instead of twice the longest of 32
            mbar_wait_a(slice_bar_a, ph_slice);
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 q = qn + c, cb = q % COUNTW_NBUF;
                mbar_wait_a(full_a + cb * 8u, (q / COUNTW_NBUF) & 1u);
                const u32 n = lds32(bufn_a + cb * 4u), lo = warp * COUNTW_SHARE, hi = min(n, lo + (u32)COUNTW_SHARE);
                const u32 recs_a = rbuf_a + cb * (COUNTW_CHUNK * 8u);
                for (u32 idx = lo + lane; idx < hi; idx += 64u) {
                    bool bA = true, bB = idx + 32u < hi;
                    const u64 recA = lds64(recs_a + idx * 8u), recB = bB ? lds64(recs_a + (idx + 32u) * 8u) : 0ull;
                    const u64 wantA = ~(recA & ~1ull), wantB = ~(recB & ~1ull);
                    u32 addrA = sl_addr + (((u32)(((u64)(u32)mix64(recA & ~1ull) * SL) >> 32)) & ~1u) * 16u;
                    u32 addrB = sl_addr + (((u32)(((u64)(u32)mix64(recB & ~1ull) * SL) >> 32)) & ~1u) * 16u;
                    u32 leftA = SL, leftB = SL;
                    auto step = [&](u64 v, u64 k, u64 rec, u64 want, u32 &addr, u32 &left, bool &busy) {
                        bool hit = k == want;
                        if (!hit && k == 0ull) {
                            u64 old;
                            asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                            if (old == 0ull) n_unique32++;
                            hit = old == 0ull || old == want;
                            v = 0;
                        }
                        if (hit) {
                            if ((u32)v < MAX_COUNT) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                                if (rec & 1ull) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + 4u) : "memory");
                            }
                            busy = false;
                            return;
                        }
                        addr += 16u;
                        if (addr == sl_end) addr = sl_addr;
                        if (--left == 0u) { n_full32++; busy = false; }
                    };
                    while (bA || bB) {
                        u64 vA, kA, vB, kB;
                        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(vA), "=l"(kA) : "r"(addrA));
                        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(vB), "=l"(kB) : "r"(addrB));
                        if (bA) step(vA, kA, recA, wantA, addrA, leftA, bA);
                        if (bB) step(vB, kB, recB, wantB, addrB, leftB, bB);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + cb * 8u);       // this warp's share of the chunk is counted
            }
            fence_async_smem();
        } else if (!BALLOT) {
            mbar_wait_a(slice_bar_a, ph_slice);
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 q = qn + c, cb = q % COUNTW_NBUF;
                mbar_wait_a(full_a + cb * 8u, (q / COUNTW_NBUF) & 1u);
                const u32 n = lds32(bufn_a + cb * 4u), lo = warp * COUNTW_SHARE, hi = min(n, lo + (u32)COUNTW_SHARE);
                const u32 recs_a = rbuf_a + cb * (COUNTW_CHUNK * 8u);
                for (u32 idx = lo + lane; idx < hi; idx += 32u) {
                    const u64 rec = lds64(recs_a + idx * 8u);
                    const u64 key = rec & ~1ull, want = ~key;
                    u32 addr = sl_addr + (((u32)(((u64)(u32)mix64(key) * SL) >> 32)) & ~1u) * 16u;
                    u32 left = SL;
                    while (true) {
                        u64 v, k;
                        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(k) : "r"(addr));      // {val, key}
                        bool hit = k == want;
                        if (!hit && k == 0ull) {
                            u64 old;
                            asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                            if (old == 0ull) n_unique32++;
                            hit = old == 0ull || old == want;
                            v = 0;
                        }
                        if (hit) {
                            if ((u32)v < MAX_COUNT) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                                if (rec & 1ull) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + 4u) : "memory");
                            }
                            break;
                        }
                        addr += 16u;
                        if (addr == sl_end) addr = sl_addr;
                        if (--left == 0u) { n_full32++; break; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + cb * 8u);       // this warp's share of the chunk is counted
            }
            fence_async_smem();                                        // this thread's counts are visible to the bulk store of the slice
        } else {
            // consumer warp: chunk c of the slice is current, records [wo, wn) of it are this warp's and not yet taken
            u32 c = 0, wo = 0, wn = 0, cur_b = 0, recs_a = rbuf_a;
            auto open_chunk = [&]() -> bool {
                while (c < n_chunks) {
                    const u32 q = qn + c;
                    cur_b = q % COUNTW_NBUF;
                    mbar_wait_a(full_a + cur_b * 8u, (q / COUNTW_NBUF) & 1u);
                    const u32 n = lds32(bufn_a + cur_b * 4u), lo = warp * COUNTW_SHARE;
                    if (lo < n) { wo = lo; wn = min(n, lo + (u32)COUNTW_SHARE); recs_a = rbuf_a + cur_b * (COUNTW_CHUNK * 8u); return true; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(empty_a + cur_b * 8u);     // nothing of this chunk is ours
                    ++c;
                }
                return false;
            };
            bool more = open_chunk();
            mbar_wait_a(slice_bar_a, ph_slice);
            u32 busy = 0, addr = sl_addr, left = 0;
            u64 rec = 0, want = 0;
            while (true) {
                const u32 bm = __ballot_sync(0xffffffffu, busy != 0u);
                u32 took = 0;
                if (more && bm != 0xffffffffu) {
                    const u32 idle = ~bm, n_idle = __popc(idle);
                    const u32 idx = wo + __popc(idle & lt);
                    if (busy == 0u && idx < wn) {
                        rec = lds64(recs_a + idx * 8u);
                        const u64 key = rec & ~1ull;
                        want = ~key;
                        addr = sl_addr + (((u32)(((u64)(u32)mix64(key) * SL) >> 32)) & ~1u) * 16u;
                        left = SL;
                        busy = 1u;
                    }
                    took = min(n_idle, wn - wo);
                    wo += n_idle;
                    if (wo >= wn) {                                    // the warp's share is in registers: the buffer may be refilled
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(empty_a + cur_b * 8u);
                        ++c;
                        more = open_chunk();
                    }
                }
                if ((bm | took) == 0u) break;                          // nothing in flight and nothing left to take
                // one probe step of every live record, without divergent paths: the slot is loaded by all lanes (an idle lane
                // re-reads its last slot), the rare claim of an empty slot sits behind a warp-uniform branch, and the two
                // counter updates are predicated instructions
                u64 v, k;
                asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(k) : "r"(addr));          // {val, key}
                bool hit = busy != 0u && k == want;
                const bool emp = busy != 0u && k == 0ull;
                if (__any_sync(0xffffffffu, emp)) {
                    if (emp) {
                        u64 old;
                        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                        if (old == 0ull) n_unique32++;
                        hit = old == 0ull || old == want;
                        v = 0;
                    }
                }
                // count in the low word, directionBias in the high word: two native 32-bit shared atomics; no carry ever
                // crosses the words
                const u32 pc = (hit && (u32)v < MAX_COUNT) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(addr), "r"(pc) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(addr + 4u), "r"(pc & (u32)rec) : "memory");
                addr += 16u;
                if (addr == sl_end) addr = sl_addr;
                --left;
                if (hit) busy = 0u;
                else if (busy != 0u && left == 0u) { n_full32++; busy = 0u; }
            }
            fence_async_smem();                                        // this thread's counts are visible to the bulk store of the slice
        }
        ph_slice ^= 1u;
        qn += n_chunks;
    }
    __syncthreads();
    if (producer && lane == 0) {
        if (prev_gsl) bulk_s2g(prev_gsl, sl, SL * 16u);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    n_unique = n_unique32; n_full = n_full32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_slices_w2: k_count_slices_ws<0> for two-word keys.  Slot = {val, key[2]} = 24 bytes, so a slice of the (smaller,
// 2048-slot) W = 2 table is 48 KB; records are 16 bytes.  A new key is published as in the global table: CAS val 0 ->
// LOCK, write the key words, exchange val -> READY | first count; a reader that finds LOCK waits for READY.
// ------------------------------------------------------------------------------------------------
static constexpr int COUNT2_CHUNK = 1024, COUNT2_NBUF = 3;             // 16 KB per chunk
static constexpr int COUNT2_SHARE = COUNT2_CHUNK / COUNTW_CONSUMERS;

__global__ void __launch_bounds__(COUNTW_TPB, 2) k_count_slices_w2(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, Counters *ctr, u32 slice0, u32 n_sl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 SL = (u32)t.part_slots;
    const u32 sl_bytes = SL * 24u;
    u64 *rbuf = reinterpret_cast<u64 *>(smem_raw + (size_t)((sl_bytes + 127u) & ~127u));
    __shared__ __align__(8) u64 bar_slice, bar_full[COUNT2_NBUF], bar_empty[COUNT2_NBUF];
    __shared__ u32 s_pre[2][32];
    __shared__ u32 buf_n[COUNT2_NBUF];
    const u32 nb = 1u << t.group_shift;
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool producer = warp == COUNTW_CONSUMERS;
    u64 n_unique = 0, n_full = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar_slice, 1);
        for (int b = 0; b < COUNT2_NBUF; ++b) { mbar_init(&bar_full[b], 1); mbar_init(&bar_empty[b], COUNTW_CONSUMERS); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto load_pre = [&](u32 pi, u32 *dst) {                            // warp 0
        u32 c = (lane < S && pi < n_sl) ? cnt2[((size_t)(pi >> t.group_shift) * S + lane) * nb + (pi & (nb - 1u))] : 0u;
        u32 v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v += x; }
        if (lane < COUNT_MAX_S + 1u) dst[lane] = v - c;
    };
    if (warp == 0) load_pre(blockIdx.x, s_pre[0]);
    u32 it = 0, ph_slice = 0, qn = 0;
    auto opaque = [](u32 x) { u32 y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
    const u32 sl_addr = opaque(smem_u32(smem_raw)), sl_end = opaque(sl_addr + sl_bytes);
    const u32 rbuf_a = opaque(smem_u32(rbuf)), full_a = opaque(smem_u32(&bar_full[0])), empty_a = opaque(smem_u32(&bar_empty[0])),
              bufn_a = opaque(smem_u32(&buf_n[0])), slice_bar_a = opaque(smem_u32(&bar_slice));
    u32 n_unique32 = 0, n_full32 = 0;
    unsigned char *prev_gsl = nullptr;
    for (u32 pi = blockIdx.x; pi < n_sl; pi += gridDim.x, ++it) {
        __syncthreads();
        if (warp == 0) load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);
        const u32 *pre = s_pre[it & 1u];
        const u32 total = pre[S];
        if (total == 0) continue;
        u32 n_chunks = 0;
        for (u32 p = 0; p < S; ++p) n_chunks += (pre[p + 1] - pre[p] + COUNT2_CHUNK - 1u) / COUNT2_CHUNK;
        unsigned char *gsl = reinterpret_cast<unsigned char *>(t.slots) + (size_t)(slice0 + pi) * sl_bytes;
        if (producer) {
            if (lane == 0) {
                const u32 g = pi >> t.group_shift, j = pi & (nb - 1u);
                const u64 *run0 = buf + (((size_t)g * S) * nb + j) * cap2 * 2;
                const size_t run_stride = (size_t)nb * cap2 * 2;
                if (prev_gsl) bulk_s2g(prev_gsl, smem_raw, sl_bytes);
                prev_gsl = gsl;
                u32 pp = 0, po = 0, q = qn;
                auto issue = [&]() {
                    while (pp < S && po >= pre[pp + 1] - pre[pp]) { ++pp; po = 0; }
                    const u32 b = q % COUNT2_NBUF;
                    if (q >= COUNT2_NBUF) mbar_wait(&bar_empty[b], ((q / COUNT2_NBUF) - 1u) & 1u);
                    const u32 n = min((u32)COUNT2_CHUNK, pre[pp + 1] - pre[pp] - po);
                    buf_n[b] = n;
                    mbar_expect_tx(&bar_full[b], n * 16u);
                    bulk_g2s(rbuf + (size_t)b * COUNT2_CHUNK * 2, run0 + (size_t)pp * run_stride + (size_t)po * 2, n * 16u, &bar_full[b]);
                    po += n; ++q;
                };
                const u32 first = min(n_chunks, (u32)COUNT2_NBUF);
                for (u32 i = 0; i < first; ++i) issue();
                bulk_wait_read();
                mbar_expect_tx(&bar_slice, sl_bytes);
                bulk_g2s(smem_raw, gsl, sl_bytes, &bar_slice);
                for (u32 i = first; i < n_chunks; ++i) issue();
            }
        } else {
            mbar_wait_a(slice_bar_a, ph_slice);
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 q = qn + c, cb = q % COUNT2_NBUF;
                mbar_wait_a(full_a + cb * 8u, (q / COUNT2_NBUF) & 1u);
                const u32 n = lds32(bufn_a + cb * 4u), lo = warp * COUNT2_SHARE, hi = min(n, lo + (u32)COUNT2_SHARE);
                const u32 recs_a = rbuf_a + cb * (COUNT2_CHUNK * 16u);
                for (u32 idx = lo + lane; idx < hi; idx += 32u) {
                    const u64 w0 = lds64(recs_a + idx * 16u), w1r = lds64(recs_a + idx * 16u + 8u);
                    const u64 w1 = w1r & ~1ull;
                    u64 key[2] = {w0, w1};
                    const u64 ph = place_hash<2>(key);
                    u32 addr = sl_addr + (u32)(((u64)(u32)ph * SL) >> 32) * 24u;
                    u32 left = SL;
                    const u64 first_val = VAL_READY | 1ull | ((w1r & 1ull) << 32);
                    while (true) {
                        u64 v = lds64(addr);
                        bool mine = false;
                        if (v == 0ull) {
                            u64 old;
                            asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr), "l"(0ull), "l"(VAL_LOCK) : "memory");
                            if (old == 0ull) {
                                asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr + 8u), "l"(w0) : "memory");
                                asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr + 16u), "l"(w1) : "memory");
                                __threadfence_block();
                                asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(first_val) : "memory");
                                n_unique32++;
                                mine = true;
                            } else v = old;
                        }
                        if (mine) break;
                        while (!(v & VAL_READY)) {                     // another thread of the CTA is publishing this slot
                            __nanosleep(20);
                            asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
                        }
                        const u64 k0 = lds64(addr + 8u), k1 = lds64(addr + 16u);
                        if (k0 == w0 && k1 == w1) {
                            if ((u32)v < MAX_COUNT) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                                if (w1r & 1ull) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + 4u) : "memory");
                            }
                            break;
                        }
                        addr += 24u;
                        if (addr == sl_end) addr = sl_addr;
                        if (--left == 0u) { n_full32++; break; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + cb * 8u);
            }
            fence_async_smem();
        }
        ph_slice ^= 1u;
        qn += n_chunks;
    }
    __syncthreads();
    if (producer && lane == 0) {
        if (prev_gsl) bulk_s2g(prev_gsl, smem_raw, sl_bytes);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    n_unique = n_unique32; n_full = n_full32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_slices_k32: k_count_slices_ws<0> for k = 32 -- 16-byte slots as for k <= 31, but 16-byte records (the key fills
// its word, the strand flag sits in bit 0 of the second word).  Two record buffers, so that two CTAs still share an SM.
// ------------------------------------------------------------------------------------------------
static constexpr int COUNT32_NBUF = 2;

__global__ void __launch_bounds__(COUNTW_TPB, 2) k_count_slices_k32(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, Counters *ctr, u32 slice0, u32 n_sl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 SL = (u32)t.part_slots;
    const u32 sl_bytes = SL * 16u;
    u64 *rbuf = reinterpret_cast<u64 *>(smem_raw + (size_t)((sl_bytes + 127u) & ~127u));
    __shared__ __align__(8) u64 bar_slice, bar_full[COUNT32_NBUF], bar_empty[COUNT32_NBUF];
    __shared__ u32 s_pre[2][32];
    __shared__ u32 buf_n[COUNT32_NBUF];
    const u32 nb = 1u << t.group_shift;
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool producer = warp == COUNTW_CONSUMERS;
    u64 n_unique = 0, n_full = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar_slice, 1);
        for (int b = 0; b < COUNT32_NBUF; ++b) { mbar_init(&bar_full[b], 1); mbar_init(&bar_empty[b], COUNTW_CONSUMERS); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto load_pre = [&](u32 pi, u32 *dst) {                            // warp 0
        u32 c = (lane < S && pi < n_sl) ? cnt2[((size_t)(pi >> t.group_shift) * S + lane) * nb + (pi & (nb - 1u))] : 0u;
        u32 v = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v += x; }
        if (lane < COUNT_MAX_S + 1u) dst[lane] = v - c;
    };
    if (warp == 0) load_pre(blockIdx.x, s_pre[0]);
    u32 it = 0, ph_slice = 0, qn = 0;
    auto opaque = [](u32 x) { u32 y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
    const u32 sl_addr = opaque(smem_u32(smem_raw)), sl_end = opaque(sl_addr + sl_bytes);
    const u32 rbuf_a = opaque(smem_u32(rbuf)), full_a = opaque(smem_u32(&bar_full[0])), empty_a = opaque(smem_u32(&bar_empty[0])),
              bufn_a = opaque(smem_u32(&buf_n[0])), slice_bar_a = opaque(smem_u32(&bar_slice));
    u32 n_unique32 = 0, n_full32 = 0;
    unsigned char *prev_gsl = nullptr;
    for (u32 pi = blockIdx.x; pi < n_sl; pi += gridDim.x, ++it) {
        __syncthreads();
        if (warp == 0) load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);
        const u32 *pre = s_pre[it & 1u];
        const u32 total = pre[S];
        if (total == 0) continue;
        u32 n_chunks = 0;
        for (u32 p = 0; p < S; ++p) n_chunks += (pre[p + 1] - pre[p] + COUNT2_CHUNK - 1u) / COUNT2_CHUNK;
        unsigned char *gsl = reinterpret_cast<unsigned char *>(t.slots) + (size_t)(slice0 + pi) * sl_bytes;
        if (producer) {
            if (lane == 0) {
                const u32 g = pi >> t.group_shift, j = pi & (nb - 1u);
                const u64 *run0 = buf + (((size_t)g * S) * nb + j) * cap2 * 2;
                const size_t run_stride = (size_t)nb * cap2 * 2;
                if (prev_gsl) bulk_s2g(prev_gsl, smem_raw, sl_bytes);
                prev_gsl = gsl;
                u32 pp = 0, po = 0, q = qn;
                auto issue = [&]() {
                    while (pp < S && po >= pre[pp + 1] - pre[pp]) { ++pp; po = 0; }
                    const u32 b = q % COUNT32_NBUF;
                    if (q >= COUNT32_NBUF) mbar_wait(&bar_empty[b], ((q / COUNT32_NBUF) - 1u) & 1u);
                    const u32 n = min((u32)COUNT2_CHUNK, pre[pp + 1] - pre[pp] - po);
                    buf_n[b] = n;
                    mbar_expect_tx(&bar_full[b], n * 16u);
                    bulk_g2s(rbuf + (size_t)b * COUNT2_CHUNK * 2, run0 + (size_t)pp * run_stride + (size_t)po * 2, n * 16u, &bar_full[b]);
                    po += n; ++q;
                };
                const u32 first = min(n_chunks, (u32)COUNT32_NBUF);
                for (u32 i = 0; i < first; ++i) issue();
                bulk_wait_read();
                mbar_expect_tx(&bar_slice, sl_bytes);
                bulk_g2s(smem_raw, gsl, sl_bytes, &bar_slice);
                for (u32 i = first; i < n_chunks; ++i) issue();
            }
        } else {
            mbar_wait_a(slice_bar_a, ph_slice);
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 q = qn + c, cb = q % COUNT32_NBUF;
                mbar_wait_a(full_a + cb * 8u, (q / COUNT32_NBUF) & 1u);
                const u32 n = lds32(bufn_a + cb * 4u), lo = warp * COUNT2_SHARE, hi = min(n, lo + (u32)COUNT2_SHARE);
                const u32 recs_a = rbuf_a + cb * (COUNT2_CHUNK * 16u);
                for (u32 idx = lo + lane; idx < hi; idx += 32u) {
                    const u64 key = lds64(recs_a + idx * 16u), fwd = lds64(recs_a + idx * 16u + 8u) & 1ull;
                    const u64 want = ~key;
                    u32 addr = sl_addr + (((u32)(((u64)(u32)mix64(key) * SL) >> 32)) & ~1u) * 16u;
                    u32 left = SL;
                    while (true) {
                        u64 v, k;
                        asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(k) : "r"(addr));      // {val, key}
                        bool hit = k == want;
                        if (!hit && k == 0ull) {
                            u64 old;
                            asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                            if (old == 0ull) n_unique32++;
                            hit = old == 0ull || old == want;
                            v = 0;
                        }
                        if (hit) {
                            if ((u32)v < MAX_COUNT) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                                if (fwd) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + 4u) : "memory");
                            }
                            break;
                        }
                        addr += 16u;
                        if (addr == sl_end) addr = sl_addr;
                        if (--left == 0u) { n_full32++; break; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty_a + cb * 8u);
            }
            fence_async_smem();
        }
        ph_slice ^= 1u;
        qn += n_chunks;
    }
    __syncthreads();
    if (producer && lane == 0) {
        if (prev_gsl) bulk_s2g(prev_gsl, smem_raw, sl_bytes);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    n_unique = n_unique32; n_full = n_full32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// k_count_slices_db: one CTA per SM, TWO slice buffers.  k_count_slices_tma / _ws spend about half of every slice visit
// waiting -- the slice is loaded, counted and written back one after the other, and only the other CTA of the SM fills
// the gaps (both variants take 95 ms on C2 although one issues a third more instructions than the other).  Here the
// producer warp loads slice j+1 and writes slice j-1 back while the 31 consumer warps count slice j, and there is no
// block-wide barrier at all: a consumer warp that has taken its share of slice j goes on to slice j+1.
//   published slice j (empty slices are skipped by the producer) lives in buffer j & 1:
//     bar_ready[j & 1]  completes for the (j >> 1)-th time when the slice has landed (its chunk count is in s_nchunks)
//     bar_done[j & 1]   completes for the (j >> 1)-th time when every consumer warp has counted its share of it
//   record chunks are numbered through all slices of the CTA and live in ring buffer q % NBUF (bar_full / bar_empty as in
//   k_count_slices_ws).  A chunk count of 0xffffffff ends the stream.
// ------------------------------------------------------------------------------------------------
static constexpr int COUNTD_CONSUMERS = 31;
static constexpr int COUNTD_TPB = (COUNTD_CONSUMERS + 1) * 32;
static constexpr int COUNTD_SHARE = 64;                                // records of a chunk that belong to one consumer warp
static constexpr int COUNTD_CHUNK = COUNTD_CONSUMERS * COUNTD_SHARE;   // 1984 records = 15872 bytes per bulk copy
static constexpr int COUNTD_NBUF = 3;

__global__ void __launch_bounds__(COUNTD_TPB, 1) k_count_slices_db(TableView t, const u64 *buf, const u32 *cnt2, u32 S, u32 cap2, Counters *ctr, u32 slice0, u32 n_sl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 SL = (u32)t.part_slots;
    u64 *rbuf = reinterpret_cast<u64 *>(smem_raw + (size_t)SL * 32);   // behind the two slice buffers: [NBUF][CHUNK] record chunks
    __shared__ __align__(8) u64 bar_ready[2], bar_done[2], bar_full[COUNTD_NBUF], bar_empty[COUNTD_NBUF];
    __shared__ u32 s_pre[2][32];                                       // producer: exclusive prefix of the sub-run sizes of this / the next slice
    __shared__ u32 buf_n[COUNTD_NBUF];                                 // records of the chunk in every ring buffer
    __shared__ u32 s_nchunks[2];                                       // chunks of the slice in every slice buffer
    const u32 nb = 1u << t.group_shift;
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_ready[b], 1); mbar_init(&bar_done[b], COUNTD_CONSUMERS); }
        for (int b = 0; b < COUNTD_NBUF; ++b) { mbar_init(&bar_full[b], 1); mbar_init(&bar_empty[b], COUNTD_CONSUMERS); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    auto opaque = [](u32 x) { u32 y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
    const u32 sl_a0 = opaque(smem_u32(smem_raw)), sl_bytes = SL * 16u;
    const u32 rbuf_a = opaque(smem_u32(rbuf)), full_a = opaque(smem_u32(&bar_full[0])), empty_a = opaque(smem_u32(&bar_empty[0])),
              bufn_a = opaque(smem_u32(&buf_n[0])), ready_a = opaque(smem_u32(&bar_ready[0])), done_a = opaque(smem_u32(&bar_done[0])),
              nch_a = opaque(smem_u32(&s_nchunks[0]));
    u32 n_unique32 = 0, n_full32 = 0;
    if (warp == COUNTD_CONSUMERS) {
        // ---------------- producer warp ----------------
        auto load_pre = [&](u32 pi, u32 *dst) {
            u32 c = (lane < S && pi < n_sl) ? cnt2[((size_t)(pi >> t.group_shift) * S + lane) * nb + (pi & (nb - 1u))] : 0u;
            u32 v = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v += x; }
            if (lane < COUNT_MAX_S + 1u) dst[lane] = v - c;
        };
        load_pre(blockIdx.x, s_pre[0]);
        u32 it = 0, j = 0, q = 0;                                      // slices visited, slices published, chunks issued
        Slot<1> *wb[2] = {nullptr, nullptr};                           // global home of the slice in every buffer
        auto retire = [&](u32 b, u32 jj) {                             // lane 0: buffer b holds published slice jj, finished or about to be
            mbar_wait_a(done_a + b * 8u, (jj >> 1) & 1u);
            bulk_s2g(wb[b], smem_raw + (size_t)b * sl_bytes, sl_bytes);    // (the consumers fenced their counts before arriving)
            wb[b] = nullptr;
        };
        for (u32 pi = blockIdx.x; pi < n_sl; pi += gridDim.x, ++it) {
            __syncwarp();
            load_pre(pi + gridDim.x, s_pre[(it + 1u) & 1u]);           // requested one slice ahead
            __syncwarp();
            const u32 *pre = s_pre[it & 1u];
            const u32 total = pre[S];
            if (total == 0) continue;                                  // nothing staged for this slice in this drain
            if (lane == 0) {
                u32 n_chunks = 0;
                for (u32 p = 0; p < S; ++p) n_chunks += (pre[p + 1] - pre[p] + COUNTD_CHUNK - 1u) / COUNTD_CHUNK;
                const u32 b = j & 1u;
                if (wb[b]) { retire(b, j - 2u); bulk_wait_read(); }    // slice j - 2 has left the buffer
                Slot<1> *gsl = reinterpret_cast<Slot<1> *>(t.slots) + (size_t)(slice0 + pi) * SL;
                wb[b] = gsl;
                s_nchunks[b] = n_chunks;
                mbar_expect_tx(&bar_ready[b], sl_bytes);
                bulk_g2s(smem_raw + (size_t)b * sl_bytes, gsl, sl_bytes, &bar_ready[b]);
                const u32 g = pi >> t.group_shift, jn = pi & (nb - 1u);
                const u64 *run0 = buf + (((size_t)g * S) * nb + jn) * cap2;    // sub-run p starts at run0 + p * nb * cap2
                const size_t run_stride = (size_t)nb * cap2;
                u32 pp = 0, po = 0;
                for (u32 i = 0; i < n_chunks; ++i, ++q) {
                    while (pp < S && po >= pre[pp + 1] - pre[pp]) { ++pp; po = 0; }
                    const u32 rb = q % COUNTD_NBUF;
                    if (q >= COUNTD_NBUF) mbar_wait_a(empty_a + rb * 8u, ((q / COUNTD_NBUF) - 1u) & 1u);   // chunk q - NBUF has been taken by every warp
                    const u32 n = min((u32)COUNTD_CHUNK, pre[pp + 1] - pre[pp] - po);
                    const u32 bytes = ((n + 1u) & ~1u) * 8u;           // bulk copies move multiples of 16 bytes (cap2 is a multiple of 4)
                    buf_n[rb] = n;
                    mbar_expect_tx(&bar_full[rb], bytes);
                    bulk_g2s(rbuf + (size_t)rb * COUNTD_CHUNK, run0 + (size_t)pp * run_stride + po, bytes, &bar_full[rb]);
                    po += n;
                }
            }
            ++j;
        }
        __syncwarp();
        if (lane == 0) {
            // end of the stream in the next buffer (once its previous slice is out), then the last slice
            const u32 b = j & 1u;
            if (wb[b]) retire(b, j - 2u);
            s_nchunks[b] = 0xffffffffu;
            mbar_arrive(&bar_ready[b]);
            if (wb[b ^ 1u]) retire(b ^ 1u, j - 1u);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        // ---------------- consumer warps ----------------
        u32 qn = 0;                                                    // number of the current slice's first chunk
        for (u32 j = 0;; ++j) {
            const u32 b = j & 1u;
            mbar_wait_a(ready_a + b * 8u, (j >> 1) & 1u);
            const u32 n_chunks = lds32(nch_a + b * 4u);
            if (n_chunks == 0xffffffffu) break;
            const u32 sl_addr = sl_a0 + b * sl_bytes, sl_end = sl_addr + sl_bytes;
            // chunk c of the slice is current, records [wo, wn) of it are this warp's and not yet taken
            u32 c = 0, wo = 0, wn = 0, cur_b = 0, recs_a = rbuf_a;
            auto open_chunk = [&]() -> bool {
                while (c < n_chunks) {
                    const u32 q = qn + c;
                    cur_b = q % COUNTD_NBUF;
                    mbar_wait_a(full_a + cur_b * 8u, (q / COUNTD_NBUF) & 1u);
                    const u32 n = lds32(bufn_a + cur_b * 4u), lo = warp * COUNTD_SHARE;
                    if (lo < n) { wo = lo; wn = min(n, lo + (u32)COUNTD_SHARE); recs_a = rbuf_a + cur_b * (COUNTD_CHUNK * 8u); return true; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(empty_a + cur_b * 8u);     // nothing of this chunk is ours
                    ++c;
                }
                return false;
            };
            bool more = open_chunk();
            u32 busy = 0, addr = sl_addr, left = 0;
            u64 rec = 0, want = 0;
            while (true) {
                const u32 bm = __ballot_sync(0xffffffffu, busy != 0u);
                u32 took = 0;
                if (more && bm != 0xffffffffu) {
                    const u32 idle = ~bm, n_idle = __popc(idle);
                    const u32 idx = wo + __popc(idle & lt);
                    if (busy == 0u && idx < wn) {
                        rec = lds64(recs_a + idx * 8u);
                        const u64 key = rec & ~1ull;
                        want = ~key;
                        addr = sl_addr + (((u32)(((u64)(u32)mix64(key) * SL) >> 32)) & ~1u) * 16u;
                        left = SL;
                        busy = 1u;
                    }
                    took = min(n_idle, wn - wo);
                    wo += n_idle;
                    if (wo >= wn) {                                    // the warp's share is in registers: the buffer may be refilled
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(empty_a + cur_b * 8u);
                        ++c;
                        more = open_chunk();
                    }
                }
                if ((bm | took) == 0u) break;                          // nothing in flight and nothing left to take
                u64 v, k;
                asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(k) : "r"(addr));          // {val, key}
                bool hit = busy != 0u && k == want;
                const bool emp = busy != 0u && k == 0ull;
                if (__any_sync(0xffffffffu, emp)) {
                    if (emp) {
                        u64 old;
                        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr + 8u), "l"(0ull), "l"(want) : "memory");
                        if (old == 0ull) n_unique32++;
                        hit = old == 0ull || old == want;
                        v = 0;
                    }
                }
                // count in the low word, directionBias in the high word: two native 32-bit shared atomics; no carry ever
                // crosses the words
                const u32 pc = (hit && (u32)v < MAX_COUNT) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(addr), "r"(pc) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(addr + 4u), "r"(pc & (u32)rec) : "memory");
                addr += 16u;
                if (addr == sl_end) addr = sl_addr;
                --left;
                if (hit) busy = 0u;
                else if (busy != 0u && left == 0u) { n_full32++; busy = 0u; }
            }
            fence_async_smem();                                        // this thread's counts are visible to the bulk store of the slice
            __syncwarp();
            if (lane == 0) mbar_arrive_a(done_a + b * 8u);
            qn += n_chunks;
        }
    }
    u64 n_unique = n_unique32, n_full = n_full32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o);
        n_full += __shfl_xor_sync(0xffffffffu, n_full, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_unique) atomicAdd(&ctr->unique, n_unique);
        if (n_full) atomicAdd(&ctr->table_full, n_full);
    }
}

// ------------------------------------------------------------------------------------------------
// table scans
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ bool slot_live(const Slot<W> &s)
{
    if (W == 1) return s.k[0] != 0 && (u32)s.val != 0;
    return (s.val & VAL_READY) && (u32)s.val != 0;
}
__device__ __forceinline__ u32 clamp_count(u64 val) { u32 c = (u32)val; return c > MAX_COUNT ? MAX_COUNT : c; }
__device__ __forceinline__ u32 clamp_dir(u64 val) { u32 d = (u32)(val >> 32) & 0x3fffffffu; return d > MAX_COUNT ? MAX_COUNT : d; }

// weightedCount / directionBias as the reference reports them: a count-1 entry is a TrackingDataSingleton
// (weight quantised to (u8)(w*254)+1, direction 0)   src/KmerTrackingData.h:641-661
__device__ __forceinline__ float report_wsum(u32 count, float wsum)
{
    if (count == 1) { unsigned char q = (unsigned char)((double)wsum * 254.0); return (float)(((int)q + 1 - 1) / 254.0); }
    return wsum;
}

// K6: exact count histogram, 65536 bins (a9: KmerSpectrum::Histogram::set src/KmerSpectrum.h:1036-1056;
// the zoomed/log bins are a host-side fold).  Low counts go through shared memory.
template <int W>
__global__ void __launch_bounds__(256) k_histogram(TableView t, u64 n_slots, u64 *hist, double *whist)
{
    __shared__ u32 sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const Slot<W> *sl = reinterpret_cast<const Slot<W> *>(t.slots);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        Slot<W> s = sl[i];
        if (!slot_live<W>(s)) continue;
        u32 c = clamp_count(s.val);
        if (c < 1024) atomicAdd(&sh[c], 1u); else atomicAdd(&hist[c], 1ull);
        if (whist && t.wsum) atomicAdd(&whist[c], (double)report_wsum(c, t.wsum[i]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
}

// a8: purgeMinDepth (src/KmerSpectrum.h:1805-1815): entries with count < min_depth lose their value; the slot
// keeps its key so probe chains stay intact (count 0 == absent everywhere).
template <int W>
__global__ void __launch_bounds__(256) k_purge(TableView t, u64 n_slots, u32 min_depth, u64 *n_purged)
{
    Slot<W> *sl = reinterpret_cast<Slot<W> *>(t.slots);
    u64 mine = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        u64 v = sl[i].val;
        u32 c = (u32)v;
        if (c != 0 && c < min_depth) {
            sl[i].val = v & VAL_READY;
            if (t.wsum) t.wsum[i] = 0.f;
            if (t.ext) for (int q = 0; q < 12; ++q) t.ext[i * 12 + q] = 0;
            mine++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_purged, mine);
}

template <int W>
__global__ void __launch_bounds__(256) k_count_live(TableView t, u64 n_slots, u32 min_count, u64 *n_live, u64 *n_single)
{
    const Slot<W> *sl = reinterpret_cast<const Slot<W> *>(t.slots);
    u64 live = 0, single = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        Slot<W> s = sl[i];
        if (!slot_live<W>(s)) continue;
        u32 c = clamp_count(s.val);
        if (c >= min_count) live++;
        if (c == 1) single++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { live += __shfl_xor_sync(0xffffffffu, live, o); single += __shfl_xor_sync(0xffffffffu, single, o); }
    if ((threadIdx.x & 31) == 0) { if (live) atomicAdd(n_live, live); if (single) atomicAdd(n_single, single); }
}

// export: compacts live entries (count >= min_count) into flat arrays; keys as reference bytes
template <int W>
__global__ void __launch_bounds__(256) k_export(TableView t, u64 n_slots, u32 min_count, u32 kb, u64 cap, u64 *cursor,
                                                uint8_t *keys, uint16_t *count, uint16_t *dir, float *wsum, u32 *ext)
{
    const Slot<W> *sl = reinterpret_cast<const Slot<W> *>(t.slots);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        Slot<W> s = sl[i];
        if (!slot_live<W>(s)) continue;
        u32 c = clamp_count(s.val);
        if (c < min_count) continue;
        u64 o = atomicAdd(cursor, 1ull);
        if (o >= cap) continue;
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = (W == 1) ? ~s.k[q] : s.k[q];
        if (keys) for (u32 b = 0; b < kb; ++b) keys[o * kb + b] = (uint8_t)(key[b >> 3] >> (56 - 8 * (b & 7)));
        if (count) count[o] = (uint16_t)c;
        if (dir) dir[o] = (uint16_t)(c == 1 ? 0 : clamp_dir(s.val));
        if (wsum) wsum[o] = t.wsum ? report_wsum(c, t.wsum[i]) : 0.f;
        if (ext) for (int q = 0; q < 12; ++q) ext[o * 12 + q] = t.ext ? t.ext[i * 12 + q] : 0u;
    }
}

// kmn_import: entries of a saved spectrum (reference-format key bytes, count, directionBias, weightedCount, extension
// counters) go into the table; an entry whose key is already present is merged (counts add up, saturating)
template <int W>
__global__ void __launch_bounds__(256) k_import(TableView t, const uint8_t *keys, const uint16_t *count, const uint16_t *dir, const float *wsum,
                                                const u32 *ext, u64 n, u32 kb, Counters *ctr)
{
    u64 n_unique = 0, n_full = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = 0;
        for (u32 b = 0; b < kb; ++b) key[b >> 3] |= (u64)keys[i * kb + b] << (56 - 8 * (b & 7));
        const u32 c = count[i];
        if (c == 0) continue;
        const u64 ph = place_hash<W>(key);
        u64 slot; u32 probes = 0;
        const int r = table_insert<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, (u64)c | ((u64)(dir ? dir[i] : 0) << 32), &slot, &probes);
        if (r < 0) { n_full++; continue; }
        n_unique += (u64)r;
        // a count-1 entry reports its weight quantised to n/254 (report_wsum); a quarter quantum keeps floor(w * 254) = n
        if (t.wsum && wsum) atomicAdd(&t.wsum[slot], c == 1 ? wsum[i] + 0.25f / 254.f : wsum[i]);
        if (t.ext && ext) for (int q = 0; q < 12; ++q) if (ext[i * 12 + q]) atomicAdd(&t.ext[slot * 12 + q], ext[i * 12 + q]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n_unique += __shfl_xor_sync(0xffffffffu, n_unique, o); n_full += __shfl_xor_sync(0xffffffffu, n_full, o); }
    if ((threadIdx.x & 31) == 0) { if (n_unique) atomicAdd(&ctr->unique, n_unique); if (n_full) atomicAdd(&ctr->table_full, n_full); }
}

// kmn_subtract: every live entry whose key is present (count >= 1) in the other table loses its value, like a purge:
// KmerSpectrum::append skips k-mers of the subtracting spectrum (src/KmerSpectrum.h:1582-1589); removing them after the
// build leaves the same table.  out[0] += entries removed, out[1] += their counts
template <int W>
__global__ void __launch_bounds__(256) k_subtract(TableView t, u64 n_slots, TableView other, u64 *out)
{
    Slot<W> *sl = reinterpret_cast<Slot<W> *>(t.slots);
    u64 removed = 0, inst = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        Slot<W> s = sl[i];
        if (!slot_live<W>(s)) continue;
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = (W == 1) ? ~s.k[q] : s.k[q];
        const u64 ph = place_hash<W>(key);
        const u64 v = table_find<W>(other, part_of(ph, other.n_parts), home_slot(ph, other.part_slots), key, nullptr);
        if ((u32)v == 0) continue;
        removed++; inst += clamp_count(s.val);
        sl[i].val = s.val & VAL_READY;
        if (t.wsum) t.wsum[i] = 0.f;
        if (t.ext) for (int q = 0; q < 12; ++q) t.ext[i * 12 + q] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { removed += __shfl_xor_sync(0xffffffffu, removed, o); inst += __shfl_xor_sync(0xffffffffu, inst, o); }
    if ((threadIdx.x & 31) == 0 && removed) { atomicAdd(out, removed); atomicAdd(out + 1, inst); }
}

// debug / parity of a5: the owner rank the device assigns to reference-format keys for a given number of ranks -- the same
// hash + owner_of the multi-GPU scatter and lookup kernels use (src/Kmer.h:2284-2295)
template <int W>
__global__ void __launch_bounds__(256) k_debug_owner(const uint8_t *keys, u64 n, u32 kb, u32 nranks, u32 use_lookup8, u32 *owner)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = 0;
        for (u32 b = 0; b < kb; ++b) key[b >> 3] |= (u64)keys[i * kb + b] << (56 - 8 * (b & 7));
        const u64 h = use_lookup8 ? hash_lookup8<W>(key, (int)kb) : hash_lookup3<W>(key, (int)kb);
        // the function phase 1b bins by (division-free); the lookup kernels use the plain remainder -- they must agree
        const u32 fast = nranks > 1 ? owner_of_fast(h, nranks, 0xFFFFFFFFu / nranks + 1u) : 0u;
        owner[i] = fast == owner_of(h, nranks) ? fast : 0xFFFFFFFFu;
    }
}

// kmn_lookup: reference-format key bytes -> u16 count (0 = absent/purged)
template <int W>
__global__ void __launch_bounds__(256) k_lookup_keys(TableView t, const uint8_t *keys, u64 n, u32 kb, uint16_t *out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = 0;
        for (u32 b = 0; b < kb; ++b) key[b >> 3] |= (u64)keys[i * kb + b] << (56 - 8 * (b & 7));
        u64 ph = place_hash<W>(key);
        u64 v = table_find<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, nullptr);
        out[i] = (uint16_t)clamp_count(v);
    }
}

// kmn_lookup with a communicator: a key owned by this rank is answered locally, any other key becomes a request in its
// owner's send region (same protocol as k_lookup_vals_dist: key words out, u16 counts back in request order)
// position of a lookup request in the region of its owner: the lanes of a warp that ask the same owner take their
// positions with ONE atomic (every k-mer of a batch used to add 1 to one of R counters: R addresses serialise in L2 at
// ~2 G atomics/s, which was the whole cost of the multi-GPU lookup pass)
__device__ __forceinline__ u64 request_slot(u64 *cursor, u32 own)
{
    const u32 act = __activemask();
    const u32 grp = __match_any_sync(act, own);
    const u32 lane = threadIdx.x & 31u, leader = (u32)__ffs((int)grp) - 1u;
    u64 base = 0;
    if (lane == leader) base = atomicAdd(&cursor[own], (u64)__popc(grp));
    base = __shfl_sync(grp, base, (int)leader);
    return base + (u64)__popc(grp & ((1u << lane) - 1u));
}

template <int W>
__global__ void __launch_bounds__(256) k_lookup_keys_dist(ParseArgs a, const uint8_t *keys, u64 n, uint16_t *out, u64 *origin)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = 0;
        for (u32 b = 0; b < a.kb; ++b) key[b >> 3] |= (u64)keys[i * a.kb + b] << (56 - 8 * (b & 7));
        const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
        const u32 own = owner_of(h, a.nranks);
        if (own == a.rank) {
            const u64 ph = place_hash<W>(key);
            out[i] = (uint16_t)clamp_count(table_find<W>(a.table, part_of(ph, a.table.n_parts), home_slot(ph, a.table.part_slots), key, nullptr));
        } else {
            const u64 pos = request_slot(a.send_cursor, own);
            if (pos < a.send_cap) {
                u64 *d = a.send_recs + ((size_t)own * a.send_cap + pos) * W;
#pragma unroll
                for (int q = 0; q < W; ++q) d[q] = key[q];
                origin[(size_t)own * a.send_cap + pos] = i;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K7a: lookup pass, part 1.  One thread walks one read (no weights), probes the table for every
// canonical k-mer and writes value(kmer) = count if count >= min_depth else 0 (setKmerValues,
// src/ReadSelector.h:1064-1076) to vals[off+i]; also records firstMarkupNorX per read.
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) k_lookup_vals(ParseArgs a, u32 min_depth, uint16_t *vals, u32 *first_nx)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += stride) {
        u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
        u32 len = (u32)(o1 - o0);
        bool disc = a.discarded && a.discarded[r];
        Walker<W> st;
        st.clear();
        if (!disc && len >= a.k) {
            st.template begin<false>(a, o0, len);
            auto emit = [&](u32 i, const u64 (&key)[W], bool, float, bool, u32) {
                u64 ph = place_hash<W>(key);
                u64 v = table_find<W>(a.table, part_of(ph, a.table.n_parts), home_slot(ph, a.table.part_slots), key, nullptr);
                u32 c = clamp_count(v);
                vals[o0 + i] = (uint16_t)(c >= min_depth ? c : 0u);
            };
            while (st.j < st.len) walker_step<W, false, false>(st, a, nullptr, emit);
        } else if (!disc) {
            // reads shorter than k still need firstMarkupNorX
            for (u32 j = 0; j < len; ++j) { u32 c = base_code(a.bases[o0 + j]); if (c == 4 && st.first_nx == 0) st.first_nx = j + 1; }
        }
        first_nx[r] = st.first_nx;
    }
}

// ------------------------------------------------------------------------------------------------
// K7a (multi-GPU): DistributedReadSelector::_batchKmerLookup (src/DistributedFunctions.h:877-902): a k-mer owned by
// this rank is looked up locally; any other k-mer becomes a request (its key words) in the owner's send region, and the
// place its answer belongs to (index into vals) is remembered in `origin` at the same position.  Requests and
// answers travel as two all-to-alls in the same order, so no request id is needed (the reference sends
// {requestId, k-mer} out and {requestId, score} back, :809-874).
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// multi-GPU lookup pass over peer memory: every rank has the count tables of all ranks mapped (CUDA IPC, NVLink), so a
// k-mer owned by another rank (a5, ((hashlittle2 >> 24) & 0x7ffff) % R) is probed in that rank's table directly: the
// same 32-byte pair loads, travelling over NVLink instead of to local HBM.  This replaces the request/response rounds of
// DistributedReadSelector::_batchKmerLookup (src/DistributedFunctions.h:877-902: 16 B out + 12 B back per k-mer, four
// all-to-all rounds per batch) -- no request buffers, no counts on the host, no collective inside the pass.
// ------------------------------------------------------------------------------------------------
struct PeerTables { void *slots[KMN_MAX_PUSH_RANKS]; };

template <int W>
__global__ void __launch_bounds__(256) k_lookup_vals_peer(ParseArgs a, PeerTables pt, u32 min_depth, uint16_t *vals, u32 *first_nx)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += stride) {
        u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
        u32 len = (u32)(o1 - o0);
        bool disc = a.discarded && a.discarded[r];
        Walker<W> st;
        st.clear();
        if (!disc && len >= a.k) {
            st.template begin<false>(a, o0, len);
            auto emit = [&](u32 i, const u64 (&key)[W], bool, float, bool, u32) {
                const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
                TableView t = a.table;
                t.slots = pt.slots[owner_of(h, a.nranks)];
                const u64 ph = place_hash<W>(key);
                const u32 c = clamp_count(table_find<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, nullptr));
                vals[o0 + i] = (uint16_t)(c >= min_depth ? c : 0u);
            };
            while (st.j < st.len) walker_step<W, false, false>(st, a, nullptr, emit);
        } else if (!disc) {
            for (u32 j = 0; j < len; ++j) { u32 c = base_code(a.bases[o0 + j]); if (c == 4 && st.first_nx == 0) st.first_nx = j + 1; }
        }
        first_nx[r] = st.first_nx;
    }
}

template <int W>
__global__ void __launch_bounds__(256) k_lookup_keys_peer(ParseArgs a, PeerTables pt, const uint8_t *keys, u64 n, uint16_t *out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = 0;
        for (u32 b = 0; b < a.kb; ++b) key[b >> 3] |= (u64)keys[i * a.kb + b] << (56 - 8 * (b & 7));
        const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
        TableView t = a.table;
        t.slots = pt.slots[owner_of(h, a.nranks)];
        const u64 ph = place_hash<W>(key);
        out[i] = (uint16_t)clamp_count(table_find<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, nullptr));
    }
}

template <int W>
__global__ void __launch_bounds__(256) k_lookup_vals_dist(ParseArgs a, u32 min_depth, uint16_t *vals, u32 *first_nx, u64 *origin)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += stride) {
        u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
        u32 len = (u32)(o1 - o0);
        bool disc = a.discarded && a.discarded[r];
        Walker<W> st;
        st.clear();
        if (!disc && len >= a.k) {
            st.template begin<false>(a, o0, len);
            auto emit = [&](u32 i, const u64 (&key)[W], bool, float, bool, u32) {
                const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
                const u32 own = owner_of(h, a.nranks);
                if (own == a.rank) {
                    u64 ph = place_hash<W>(key);
                    u64 v = table_find<W>(a.table, part_of(ph, a.table.n_parts), home_slot(ph, a.table.part_slots), key, nullptr);
                    u32 c = clamp_count(v);
                    vals[o0 + i] = (uint16_t)(c >= min_depth ? c : 0u);
                } else {
                    const u64 pos = request_slot(a.send_cursor, own);
                    if (pos < a.send_cap) {
                        u64 *d = a.send_recs + ((size_t)own * a.send_cap + pos) * W;
#pragma unroll
                        for (int q = 0; q < W; ++q) d[q] = key[q];
                        origin[(size_t)own * a.send_cap + pos] = o0 + i;
                    }
                }
            };
            while (st.j < st.len) walker_step<W, false, false>(st, a, nullptr, emit);
        } else if (!disc) {
            for (u32 j = 0; j < len; ++j) { u32 c = base_code(a.bases[o0 + j]); if (c == 4 && st.first_nx == 0) st.first_nx = j + 1; }
        }
        first_nx[r] = st.first_nx;
    }
}

// owner side: answer the received requests (keys as W words each) in order
template <int W>
__global__ void __launch_bounds__(256) k_lookup_words(TableView t, const u64 *keys, u64 n, uint16_t *out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[W];
#pragma unroll
        for (int q = 0; q < W; ++q) key[q] = keys[i * W + q];
        u64 ph = place_hash<W>(key);
        u64 v = table_find<W>(t, part_of(ph, t.n_parts), home_slot(ph, t.part_slots), key, nullptr);
        out[i] = (uint16_t)clamp_count(v);
    }
}

// requester side: put the answers where they belong (setKmerValues semantics: below min_depth -> 0)
__global__ void __launch_bounds__(256) k_scatter_answers(const uint16_t *resp, const u64 *origin, u64 region_cap, const u64 *counts,
                                                         u32 nranks, u32 min_depth, uint16_t *vals)
{
    for (u32 d = blockIdx.y; d < nranks; d += gridDim.y) {
        const u64 n = counts[d];
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
            const u32 c = resp[(size_t)d * region_cap + i];
            vals[origin[(size_t)d * region_cap + i]] = (uint16_t)(c >= min_depth ? c : 0u);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K7b: lookup pass, part 2.  One warp per read: longest run of vals >= min_depth over [0,numKmers)
// (first-longest wins), score of the run, ReadTrimType after setTrimHeaders.
// ReadSelector::trimReadByMinimumKmerScore / scoreReadByScoringType / setTrimHeaders / _setNumKmers
// (src/ReadSelector.h:948-1047,1092-1180).  KS_SUM never assigns the score (:1151-1162) -> 0.
// ------------------------------------------------------------------------------------------------
struct TrimArgs {
    const u64 *read_off; const uint8_t *discarded; const uint16_t *vals; const u32 *first_nx;
    u64 n_reads; u32 k, min_depth; int scoring;
    u32 *trim_off, *trim_len; float *score; uint8_t *was_trimmed;
};

__global__ void __launch_bounds__(256) k_trim_score(TrimArgs a)
{
    __shared__ u32 hist[8][256];
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u64 wstride = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + warp; r < a.n_reads; r += wstride) {
        if (a.discarded && a.discarded[r]) {          // never scored (src/ReadSelector.h:1196-1198): default ReadTrimType
            if (lane == 0) { a.trim_off[r] = 0; a.trim_len[r] = 0; a.score[r] = 0.f; a.was_trimmed[r] = 0; }
            continue;
        }
        const u64 o0 = a.read_off[r];
        const u32 len = (u32)(a.read_off[r + 1] - o0);
        u32 num = len >= a.k ? len - a.k + 1 : 0;
        const u32 ml = a.first_nx[r];
        if (ml != 0) { u32 capn = ml > a.k ? ml - a.k : 0; if (capn < num) num = capn; }
        const uint16_t *v = a.vals + o0;
        // longest run, first-longest wins; every lane runs the same scalar scan over ballot words
        u32 best_off = 0, best_len = 0, cur_off = 0, cur_len = 0;
        for (u32 base = 0; base < num; base += 32) {
            u32 i = base + lane;
            bool ok = i < num && v[i] >= a.min_depth;
            u32 m = __ballot_sync(0xffffffffu, ok);
            u32 nbits = num - base < 32u ? num - base : 32u;
            u32 pos = 0;
            while (pos < nbits) {
                u32 rest = m >> pos;
                if (rest & 1u) {                       // run of ones
                    u32 ones = (~rest) ? (u32)__ffs(~rest) - 1u : 32u;
                    if (ones > nbits - pos) ones = nbits - pos;
                    if (cur_len == 0) cur_off = base + pos;
                    cur_len += ones; pos += ones;
                } else {                               // run of zeros ends the current run
                    if (cur_len > best_len) { best_len = cur_len; best_off = cur_off; }
                    cur_len = 0;
                    u32 zeros = rest ? (u32)__ffs(rest) - 1u : 32u;
                    if (zeros > nbits - pos) zeros = nbits - pos;
                    pos += zeros;
                }
            }
        }
        if (cur_len > best_len) { best_len = cur_len; best_off = cur_off; }
        const bool trimmed = best_len < num;
        float sc = -1.f;
        u32 out_off = best_off, out_len = 0;
        if (best_len > 0) {
            const uint16_t *b = v + best_off;
            if (a.scoring == 3 || a.scoring == 2) {                 // MAX / MIN
                u32 m = a.scoring == 3 ? 0u : 0xffffffffu;
                for (u32 i = lane; i < best_len; i += 32) { u32 x = b[i]; m = a.scoring == 3 ? max(m, x) : min(m, x); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { u32 x = __shfl_xor_sync(0xffffffffu, m, o); m = a.scoring == 3 ? max(m, x) : min(m, x); }
                sc = (float)m;
            } else if (a.scoring == 4) {                            // AVG: double sum of integers is exact in any order
                u64 s = 0;
                for (u32 i = lane; i < best_len; i += 32) s += b[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                sc = (float)((double)s / (int)best_len);
            } else if (a.scoring == 1) {                            // MEDIAN = sorted[len/2]: two-pass radix select
                u32 *h = hist[warp];
                u32 want = best_len / 2;                            // 0-based rank
                for (u32 i = lane; i < 256; i += 32) h[i] = 0;
                __syncwarp();
                for (u32 i = lane; i < best_len; i += 32) atomicAdd(&h[b[i] >> 8], 1u);
                __syncwarp();
                u32 hi = 0, acc = 0;
                for (u32 q = 0; q < 256; ++q) { u32 c = h[q]; if (acc + c > want) { hi = q; break; } acc += c; }
                __syncwarp();
                for (u32 i = lane; i < 256; i += 32) h[i] = 0;
                __syncwarp();
                for (u32 i = lane; i < best_len; i += 32) if ((u32)(b[i] >> 8) == hi) atomicAdd(&h[b[i] & 0xff], 1u);
                __syncwarp();
                u32 lo = 0;
                for (u32 q = 0; q < 256; ++q) { u32 c = h[q]; if (acc + c > want) { lo = q; break; } acc += c; }
                __syncwarp();
                sc = (float)((hi << 8) | lo);
            } else {
                sc = 0.f;                                           // KS_SUM: score never assigned
            }
            out_len = best_len + a.k - 1;
        } else {
            out_off = 0; sc = -1.f;
        }
        if (lane == 0) { a.trim_off[r] = out_off; a.trim_len[r] = out_len; a.score[r] = sc; a.was_trimmed[r] = trimmed ? 1 : 0; }
    }
}

// ------------------------------------------------------------------------------------------------
// debug / parity: per-k-mer records of a batch (steps 1-3 of the path)
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(128) k_debug_kmers(ParseArgs a, const u64 *kmer_off, uint8_t *keys, uint8_t *is_fwd, float *weight, u64 *hash)
{
    __shared__ double ptab[256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) ptab[i] = a.ptab[i];
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += stride) {
        u64 o0 = a.read_off[r];
        u32 len = (u32)(a.read_off[r + 1] - o0);
        if (len < a.k) continue;
        const u64 ko = kmer_off[r];
        Walker<W> st;
        st.template begin<true>(a, o0, len);
        auto emit = [&](u32 i, const u64 (&key)[W], bool fwd, float wf, bool, u32) {
            for (u32 b = 0; b < a.kb; ++b) keys[(ko + i) * a.kb + b] = (uint8_t)(key[b >> 3] >> (56 - 8 * (b & 7)));
            is_fwd[ko + i] = fwd ? 1 : 0;
            weight[ko + i] = wf;
            hash[ko + i] = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
        };
        while (st.j < st.len) walker_step<W, true, false>(st, a, ptab, emit);
    }
}

}  // namespace kmn
