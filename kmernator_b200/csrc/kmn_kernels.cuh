// kmn_kernels.cuh -- sm_100a kernels of the k-mer spectrum path.
//
//   count pass  = k_weight_mask  (phase 1a: quality weights along each read -> one "counted" bit per k-mer position)
//               + k_kmer_scatter (phase 1b: bases -> canonical k-mers, position-parallel -> partitioned staging)
//               + k_insert_staged (phase 2: per-partition inserts into an L2-resident table slice)
//   lookup pass = k_lookup_vals + k_trim_score
//   table scans = k_histogram, k_purge, k_export, k_count_singletons
//
// Why two phases: on B200 a 64-bit atomic to an HBM-resident table runs at ~20 G/s while the same
// atomic to a <=64 MB (L2-resident) region runs at 60-190 G/s (profiles/r01_randacc_microbench.csv).
// Phase 1 therefore scatters every k-mer into one of n_parts staging regions (each phase-1 CTA owns a private
// sub-region of every partition, addressed by shared-memory counters), and phase 2 walks the regions in order
// so that only one table slice is hot.
#pragma once
#include "kmn_device.cuh"
#include <cooperative_groups.h>

namespace kmn {

static constexpr int MASK_TPB = 256;       // phase 1a: one thread per read
#ifndef KMN_SCATTER_TPB
#define KMN_SCATTER_TPB 1024
#define KMN_SCATTER_CTAS 1
#endif
static constexpr int SCATTER_TPB = KMN_SCATTER_TPB;    // phase 1b: one thread per read, SCATTER_CTAS CTAs per SM
static constexpr int SCATTER_CTAS = KMN_SCATTER_CTAS;
static constexpr int INSERT_TPB = 256;
static constexpr int INSERT_UNROLL = 4;
static constexpr int INSERT_CHUNK = INSERT_TPB * INSERT_UNROLL;
static constexpr int INSERT_GROUP = 4;      // chunks per ticket

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
    u32 use_lookup8;
    u32 l2_hints;          // staging stores carry an L2 evict_last policy
    TableView table;
    StageView stage;
    Counters *ctr;
    // multi-GPU count pass: records owned by other ranks go to this CTA's segment of the destination's send region
    u64 *seg_recs;         // [nranks][n_cta][seg_cap][RW]
    u32 *seg_count;        // [nranks][n_cta]
    u32 seg_cap;
    // multi-GPU lookup pass: requests are appended to per-destination regions
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
#pragma unroll
            for (int u = 0; u < 8; ++u) emit(i0 + u, s.roll.f, true, s.wf, s.good, 0x3fu);
            s.sb.advance();
            s.sqi.advance(); s.sqo.advance();
            s.j = j0 + 8;
            return;
        }
    }

#pragma unroll
    for (int u = 0; u < 8; ++u) {
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
                emit(i, fwd ? s.roll.f : s.roll.r, fwd, NEED_W ? s.wf : 1.0f, NEED_W ? s.good : true, eb);
            }
        }
    }
    s.sb.advance();
    if (NEED_Q) { s.sqi.advance(); s.sqo.advance(); }
    s.j = j0 + 8;
}

// ------------------------------------------------------------------------------------------------
// K2a: phase 1a of the count pass.  One thread walks one read and evaluates the reference's sequential weight
// recurrence (a3, src/KmerReadUtils.h:176-248); the only thing the count pass needs from it is one bit per k-mer
// position -- "(float)w > minimumWeight" (src/KmerSpectrum.h:1598, src/KmerTrackingData.h:354-364) -- which goes to a
// bit array indexed by the k-mer's first base in the concatenated batch (and, for KMN_VALUE_WEIGHTS, the fp32
// weight itself).  Splitting this off leaves phase 1b free of any sequential dependency along the read.
// ------------------------------------------------------------------------------------------------
template <bool WTS>
__global__ void __launch_bounds__(MASK_TPB) k_weight_mask(ParseArgs a)
{
    __shared__ double ptab[256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) ptab[i] = a.ptab[i];
    __syncthreads();
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    const u64 stride = (u64)gridDim.x * blockDim.x;
    Walker<1> st;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += stride) {
        const u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
        const u32 len = (u32)(o1 - o0);
        if (len < a.k || (a.discarded && a.discarded[r])) continue;
        st.template begin<true>(a, o0, len);
        u64 curw = ~0ull;
        u32 acc = 0;
        auto emit = [&](u32 i, const u64 (&)[1], bool, float wf, bool good, u32) {
            const u64 gb = o0 + i, w = gb >> 5;
            if (w != curw) { if (acc) atomicOr(&a.mask[curw], acc); acc = 0; curw = w; }
            if (good) { acc |= 1u << (u32)(gb & 31ull); lc.good++; }
            if (WTS) a.wts[gb] = wf;
            lc.raw++;
        };
        while (st.j < st.len) walker_step<1, true, false, false>(st, a, ptab, emit);
        if (acc) atomicOr(&a.mask[curw], acc);
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

template <int W, bool HASX>
__device__ __forceinline__ void stage_put(const StageView &st, const TableView &tab, u32 *cnt, u32 part, const Rec<W, HASX> &rec, LocalCtr &lc,
                                          u64 policy)
{
    constexpr int RW = Rec<W, HASX>::RW;
    const u32 pos = atomicAdd(&cnt[part], 1u);
    if (pos < st.sub_cap) {
        u64 *d = st.recs + (st.sub_index(part, blockIdx.x) * st.sub_cap + pos) * RW;
#pragma unroll
        for (int q = 0; q < RW; ++q) st_hint64(d + q, rec.w[q], policy);
    } else {
        insert_record<W, HASX>(tab, rec, lc.unique, lc.full, lc.probes);
        lc.direct++;
    }
}

// ------------------------------------------------------------------------------------------------
// K1+K2b (+K5 partition): phase 1b of the count pass.  One thread walks one read and rolls the canonical k-mer
// (a1 TwoBitSequence::compressSequence src/TwoBitSequence.cpp:242-269, a2 KmerArrayPair::build src/Kmer.h:1323-1375,
// buildLeastComplement :356-364) 8 bases per step; k-mers whose "counted" bit (phase 1a) is set go to this CTA's
// sub-region of their table partition -- one shared-memory atomicAdd for the position, one 8*RW-byte store.
// Multi-GPU: records owned by another rank (a5: owner = lookup3 hash, src/Kmer.h:2284-2295) go to that rank's
// send region instead.
// ------------------------------------------------------------------------------------------------
template <int W, bool HASX, bool EXT, bool DIST>
__global__ void __launch_bounds__(SCATTER_TPB, SCATTER_CTAS) k_kmer_scatter(ParseArgs a)
{
    constexpr int RW = Rec<W, HASX>::RW;
    extern __shared__ __align__(16) u32 smem_u32[];
    const u32 n_parts = a.stage.n_parts;                               // staging partitions = table groups
    u32 *cnt = smem_u32;                                               // [n_parts] fill level of this CTA's sub-regions
    u32 *scnt = smem_u32 + ((n_parts + 31u) & ~31u);                   // [nranks] fill level of this CTA's send segments
    for (u32 i = threadIdx.x; i < n_parts; i += blockDim.x) cnt[i] = a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x];
    if (DIST) for (u32 i = threadIdx.x; i < a.nranks; i += blockDim.x) scnt[i] = a.seg_count[(size_t)i * gridDim.x + blockIdx.x];
    __syncthreads();

    LocalCtr lc{0, 0, 0, 0, 0, 0};
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 keep = a.l2_hints ? l2_policy_evict_last() : l2_policy_evict_normal();
    Walker<W> st;
    // consecutive groups of 32 reads (one warp) go to different CTAs, so even a small batch spreads evenly over the
    // per-CTA sub-regions / segments while a warp still streams one contiguous piece of the batch
    for (u64 r = ((u64)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u); r < a.n_reads; r += stride) {
        const u64 o0 = a.read_off[r], o1 = a.read_off[r + 1];
        const u32 len = (u32)(o1 - o0);
        if (len < a.k || (a.discarded && a.discarded[r])) continue;
        st.template begin<EXT>(a, o0, len);
        u64 curw = o0 >> 5;                                            // mask word holding the current position's bit
        u32 mw = __ldg(&a.mask[curw]), mwn = __ldg(&a.mask[curw + 1]);
        auto emit = [&](u32 i, const u64 (&key)[W], bool fwd, float, bool, u32 eb) {
            const u64 gb = o0 + i, w = gb >> 5;
            if (w != curw) { curw = w; mw = mwn; mwn = __ldg(&a.mask[w + 1]); }
            if (!((mw >> (u32)(gb & 31ull)) & 1u)) return;
            Rec<W, HASX> rec;
            rec.pack(key, fwd, (HASX && a.wts) ? a.wts[gb] : 1.0f, eb);
            if (DIST) {
                const u64 h = a.use_lookup8 ? hash_lookup8<W>(key, (int)a.kb) : hash_lookup3<W>(key, (int)a.kb);
                const u32 own = owner_of(h, a.nranks);
                if (own != a.rank) {
                    // one shared atomicAdd per (converged lanes, destination) group
                    namespace cg = cooperative_groups;
                    auto grp = cg::labeled_partition(cg::coalesced_threads(), (int)own);
                    u32 base = 0;
                    if (grp.thread_rank() == 0) base = atomicAdd(&scnt[own], (u32)grp.size());
                    base = grp.shfl(base, 0);
                    const u32 pos = base + grp.thread_rank();
                    if (pos < a.seg_cap) {
                        u64 *d = a.seg_recs + (((size_t)own * gridDim.x + blockIdx.x) * a.seg_cap + pos) * RW;
#pragma unroll
                        for (int q = 0; q < RW; ++q) d[q] = rec.w[q];
                    }
                    return;
                }
            }
            const u64 ph = place_hash<W>(key);
            stage_put<W, HASX>(a.stage, a.table, cnt, part_of(ph, a.table.n_parts) >> a.table.group_shift, rec, lc, keep);
        };
        while (st.j < st.len) walker_step<W, false, EXT>(st, a, nullptr, emit);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_parts; i += blockDim.x) a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x] = cnt[i];
    if (DIST) for (u32 i = threadIdx.x; i < a.nranks; i += blockDim.x) a.seg_count[(size_t)i * gridDim.x + blockIdx.x] = scnt[i];
    ctr_commit(a.ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: the per-CTA send segments of every destination are packed into one contiguous send buffer per
// destination (what the all-to-all ships).  grid = (n_cta, nranks).  send_cursor[d] = records for d; flag != 0 on overflow.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact_send(const u64 *seg_recs, const u32 *seg_count, u32 seg_cap, u32 n_cta, u32 rw,
                                                      u64 *send_recs, u64 send_cap, u64 *send_cursor, u64 *flag)
{
    const u32 d = blockIdx.y, c = blockIdx.x;
    __shared__ u64 part[8];
    __shared__ u64 s_off;
    u64 mine = 0;
    for (u32 i = threadIdx.x; i < c; i += blockDim.x) { const u32 n = seg_count[(size_t)d * n_cta + i]; mine += n < seg_cap ? n : seg_cap; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int i = 0; i < 8; ++i) t += part[i]; s_off = t; }
    __syncthreads();
    const u64 off = s_off;
    u32 n = seg_count[(size_t)d * n_cta + c];
    if (n > seg_cap) { if (threadIdx.x == 0) atomicAdd(flag, 1ull); n = seg_cap; }
    if (off + n > send_cap) { if (threadIdx.x == 0) atomicAdd(flag, 1ull); n = off < send_cap ? (u32)(send_cap - off) : 0u; }
    const u64 *src = seg_recs + ((size_t)d * n_cta + c) * seg_cap * rw;
    u64 *dst = send_recs + ((size_t)d * send_cap + off) * rw;
    for (u64 i = threadIdx.x; i < (u64)n * rw; i += blockDim.x) dst[i] = src[i];
    if (c == n_cta - 1 && threadIdx.x == 0) send_cursor[d] = off + n;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: records received from other ranks are routed into the local staging regions the same way
// (the receiving half of MPIAllToAllMessageBuffer, src/MPIBuffer.h:412-1073).  Same grid as k_kmer_scatter.
// ------------------------------------------------------------------------------------------------
struct RouteArgs {
    const u64 *recs; u64 n_recs;
    TableView table; StageView stage; Counters *ctr;
};

template <int W, bool HASX>
__global__ void __launch_bounds__(SCATTER_TPB, SCATTER_CTAS) k_route_records(RouteArgs a)
{
    constexpr int RW = Rec<W, HASX>::RW;
    extern __shared__ __align__(16) u32 smem_u32[];
    const u32 n_parts = a.stage.n_parts;
    u32 *cnt = smem_u32;
    for (u32 i = threadIdx.x; i < n_parts; i += blockDim.x) cnt[i] = a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x];
    __syncthreads();
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 keep = l2_policy_evict_last();
    for (u64 idx = (u64)blockIdx.x * 32u + (threadIdx.x & 31u) + (u64)(threadIdx.x >> 5) * 32u * gridDim.x; idx < a.n_recs; idx += stride) {
        Rec<W, HASX> rec;
#pragma unroll
        for (int q = 0; q < RW; ++q) rec.w[q] = ld_nc64(a.recs + idx * RW + q);
        u64 key[W]; bool fwd; float wt; u32 eb;
        rec.unpack(key, fwd, wt, eb);
        const u64 ph = place_hash<W>(key);
        stage_put<W, HASX>(a.stage, a.table, cnt, part_of(ph, a.table.n_parts) >> a.table.group_shift, rec, lc, keep);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_parts; i += blockDim.x) a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x] = cnt[i];
    ctr_commit(a.ctr, lc);
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
// phase 2 work list over the (partition, phase-1 CTA) sub-regions in partition-major order:
// chunk_start[e] = first chunk index of sub-region e (exclusive scan), single CTA
// ------------------------------------------------------------------------------------------------
__global__ void k_build_worklist(const u32 *count, u32 sub_cap, u32 n_entries, u32 chunk, u64 *chunk_start, u64 *next_item)
{
    __shared__ u64 carry;
    __shared__ u64 wsum[32];
    if (threadIdx.x == 0) { carry = 0; *next_item = 0; }
    __syncthreads();
    for (u32 base = 0; base < n_entries; base += blockDim.x) {
        u32 p = base + threadIdx.x;
        u64 n = 0;
        if (p < n_entries) { u32 c = count[p]; if (c > sub_cap) c = sub_cap; n = (c + chunk - 1) / chunk; }
        u64 v = n;
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
        u64 incl = v + (warp ? wsum[warp - 1] : 0) + carry;
        if (p < n_entries) chunk_start[p] = incl - n;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_start[n_entries] = carry;
}

// ------------------------------------------------------------------------------------------------
// K3: phase 2 of the count pass.  Persistent CTAs take (sub-region, chunk) items in partition order from an
// atomic ticket, so at any time the whole GPU works on at most ~2 neighbouring partitions whose table
// slices (slice_bytes each) stay L2-resident.  Per record: 16-B slot load, then CAS (new key) or RED (hit).
// Replaces KmerSpectrum::append (src/KmerSpectrum.h:1578-1668) + KmerMapByKmerArrayPair insert/find
// (src/Kmer.h:1491-1544,3095-3110) + TrackingData::track (src/KmerTrackingData.h:427-448,517-529).
// ------------------------------------------------------------------------------------------------
template <int W, bool HASX, bool PRE>
__global__ void __launch_bounds__(INSERT_TPB) k_insert_staged(TableView t, StageView st, const u64 *chunk_start, u64 *next_item, Counters *ctr)
{
    constexpr int RW = Rec<W, HASX>::RW;
    __shared__ u64 s_item;
    __shared__ u32 s_entry;
    u64 n_unique = 0, n_full = 0, n_probes = 0;
    const u32 n_entries = st.n_parts * st.n_cta;
    const u64 total_items = chunk_start[n_entries];
    while (true) {
        // one ticket = INSERT_GROUP consecutive chunks; the sub-region of the first one is found by bisection (the upper
        // levels of the search are the same lines for every ticket and stay in L1), the following ones by stepping
        if (threadIdx.x == 0) {
            const u64 it = atomicAdd(next_item, (u64)INSERT_GROUP);
            s_item = it;
            if (it < total_items) {
                u32 lo = 0, hi = n_entries;
                while (hi - lo > 1) { const u32 mid = lo + ((hi - lo) >> 1); if (__ldg(&chunk_start[mid]) <= it) lo = mid; else hi = mid; }
                s_entry = lo;
            }
        }
        __syncthreads();
        const u64 item0 = s_item;
        u32 entry = s_entry;
        __syncthreads();
        if (item0 >= total_items) break;
#pragma unroll 1
      for (u32 g = 0; g < (u32)INSERT_GROUP; ++g) {
        const u64 item = item0 + g;
        if (item >= total_items) break;
        while (__ldg(&chunk_start[entry + 1]) <= item) ++entry;
        const u32 part = entry / st.n_cta;
        u64 n = st.count[entry]; if (n > st.sub_cap) n = st.sub_cap;
        const u64 first = (item - __ldg(&chunk_start[entry])) * INSERT_CHUNK;
        const u64 *src = st.recs + (st.sub_index(part, entry - part * st.n_cta) * st.sub_cap + first) * RW;
        const u64 cnt = n - first < (u64)INSERT_CHUNK ? n - first : (u64)INSERT_CHUNK;
        Rec<W, HASX> rec[INSERT_UNROLL];
        bool have[INSERT_UNROLL];
#pragma unroll
        for (int u = 0; u < INSERT_UNROLL; ++u) {
            u64 idx = (u64)u * INSERT_TPB + threadIdx.x;
            have[u] = idx < cnt;
            if (have[u]) {
#pragma unroll
                for (int q = 0; q < RW; ++q) rec[u].w[q] = ld_nc64(src + idx * RW + q);
            }
        }
        // PRE: home-slot loads of all records first (independent, all in flight together), then resolve one by one
        u64 key[INSERT_UNROLL][W], home[INSERT_UNROLL], pv[INSERT_UNROLL], pk[INSERT_UNROLL];
        bool fwd[INSERT_UNROLL]; float weight[INSERT_UNROLL]; u32 eb[INSERT_UNROLL]; u32 slice[INSERT_UNROLL];
#pragma unroll
        for (int u = 0; u < INSERT_UNROLL; ++u) {
            pv[u] = pk[u] = 0; home[u] = 0; slice[u] = 0;
            if (have[u]) {
                rec[u].unpack(key[u], fwd[u], weight[u], eb[u]);
                const u64 ph = place_hash<W>(key[u]);
                slice[u] = part_of(ph, t.n_parts);                 // a slice of group `part`
                home[u] = home_slot(ph, t.part_slots);
                const Slot<W> *pbase = reinterpret_cast<const Slot<W> *>(t.slots) + (u64)slice[u] * t.part_slots;
                if (PRE) {
                    if (W == 1) ld_slot16(pbase + home[u], pv[u], pk[u]);
                    else pv[u] = ld_cg64(&pbase[home[u]].val);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < INSERT_UNROLL; ++u) {
            if (have[u]) {
                u64 slot; u32 probes = 0;
                int r = table_insert<W, PRE>(t, slice[u], home[u], key[u], 1ull | ((u64)(fwd[u] ? 1u : 0u) << 32), &slot, &probes, pv[u], pk[u]);
                if (r < 0) { n_full++; continue; }
                n_unique += (u64)r; n_probes += probes;
                if (HASX) {
                    if (t.wsum) atomicAdd(&t.wsum[slot], weight[u]);
                    if (t.ext) {
                        u32 l = eb[u] & 7u, rr = (eb[u] >> 3) & 7u;
                        if (l < 6) atomicAdd(&t.ext[slot * 12 + l], 1u);
                        if (rr < 6) atomicAdd(&t.ext[slot * 12 + 6 + rr], 1u);
                    }
                }
            }
        }
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
// Fast count path (k <= 31, plain TrackingDataWithDirection values): phase 2 in two levels, so that no k-mer instance
// costs an individual L2/HBM transaction.  Measured on B200 (profiles/r01_randacc_microbench.csv): scattered 16-byte
// loads run at ~145 G/s and scattered REDs at ~190 G/s even when L2-resident -- an SM-side limit on uncoalesced sector
// requests -- so one load + one RED per instance caps the insert at ~77 G/s.  Here every instance is moved twice with
// coalesced accesses (level 1: group, level 2: slice) and counted with shared-memory atomics (level 3).
//
// Level 2 (k_subpartition): a CTA takes a tile of SUB_TILE records of one group, counting-sorts it by slice in shared
// memory (histogram -> scan -> scatter), reserves one range per slice bucket with a single global atomicAdd, and copies
// the sorted tile out so that neighbouring lanes write neighbouring records of the same bucket.
// ------------------------------------------------------------------------------------------------
static constexpr int TS_TPB = 256;                   // threads of a tile-sorting CTA
static constexpr int TS_R = 8;                       // records per thread
static constexpr int TS_TILE = TS_TPB * TS_R;        // 2048 records per tile
static constexpr int TS_MAX_BPT = 10;                // bins per thread in the scan: up to 2560 bins
static constexpr u32 TS_NONE = 0xffffffffu;
static constexpr int SUB_TILE = TS_TILE;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *warp_sums /*[TS_TPB/32]*/)
{
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 t = lane < (TS_TPB >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 q = __shfl_up_sync(0xffffffffu, t, o); if (lane >= (u32)o) t += q; }
        if (lane < (TS_TPB >> 5)) warp_sums[lane] = t;             // inclusive over warps
    }
    __syncthreads();
    return incl - v + (warp ? warp_sums[warp - 1] : 0u);
}

// Counting sort of one tile by bin, then a copy-out in sorted order.  Every thread brings up to TS_R records with their
// bins (TS_NONE = no record).  PRE: hist[0..nb) is zero and a barrier has been passed since it was zeroed.
//   reserve(bin, n, &room) -> address of the tile's first record in the bin's output region (called once per non-empty
//                             bin by the thread that owns the bin in the scan); room = records that still fit there
//   overflow(rec, bin) is called for records beyond `room`
// Leaves no barrier pending: the caller must pass a barrier before it touches hist/buf/dst again.
template <typename RESERVE, typename OVERFLOW>
__device__ __forceinline__ void tile_sort_flush(const u64 (&rec)[TS_R], const u32 (&bin)[TS_R], u32 nb, u64 *buf, u64 *dst, u32 *hist, u64 *gptr,
                                                u32 *room, u32 *wsum, u32 *s_total, RESERVE &&reserve, OVERFLOW &&overflow)
{
    u32 rank[TS_R];
#pragma unroll
    for (int u = 0; u < TS_R; ++u) rank[u] = bin[u] != TS_NONE ? atomicAdd(&hist[bin[u]], 1u) : 0u;
    __syncthreads();
    {
        const u32 bpt = (nb + TS_TPB - 1) / TS_TPB;
        const u32 b0 = threadIdx.x * bpt;
        u32 tot = 0;
        for (u32 q = 0; q < bpt; ++q) if (b0 + q < nb) tot += hist[b0 + q];
        u32 run = block_exclusive_scan(tot, wsum);
        for (u32 q = 0; q < bpt; ++q) {
            const u32 b = b0 + q;
            if (b < nb) {
                const u32 c = hist[b];
                hist[b] = run;
                if (c) { u32 rm; gptr[b] = (u64)reserve(b, c, rm); room[b] = rm; }
                run += c;
            }
        }
        if (threadIdx.x == TS_TPB - 1) *s_total = run;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < TS_R; ++u) {
        if (bin[u] != TS_NONE) {
            const u32 pos = hist[bin[u]] + rank[u];
            buf[pos] = rec[u];
            dst[pos] = rank[u] < room[bin[u]] ? gptr[bin[u]] + 8ull * rank[u] : (u64)bin[u];   // bins are small numbers, never an address
        }
    }
    __syncthreads();
    const u32 total = *s_total;
    for (u32 j = threadIdx.x; j < total; j += TS_TPB) {
        const u64 d = dst[j], r = buf[j];
        if (d >= 65536ull) *reinterpret_cast<u64 *>(d) = r;
        else overflow(r, (u32)d);
    }
}

// flat work list of level 2: item_entry[i] = staging sub-region of tile i (tiles of a sub-region are consecutive)
__global__ void __launch_bounds__(256) k_fill_items(const u64 *chunk_start, u32 n_entries, u32 *item_entry)
{
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += gridDim.x * blockDim.x)
        for (u64 i = chunk_start[e]; i < chunk_start[e + 1]; ++i) item_entry[i] = e;
}

__global__ void __launch_bounds__(TS_TPB, 4) k_subpartition(TableView t, StageView st, Stage2View s2, const u64 *chunk_start, const u32 *item_entry,
                                                             Counters *ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);          // [TS_TILE] records, sorted by slice
    u64 *dst = buf + TS_TILE;                              // [TS_TILE] destination address of buf[j]
    const u32 gp = 1u << t.group_shift;
    u64 *gptr = dst + TS_TILE;                             // [gp]
    u32 *hist = reinterpret_cast<u32 *>(gptr + gp);        // [gp]
    u32 *room = hist + gp;                                 // [gp]
    __shared__ u32 s_wsum[TS_TPB / 32];
    __shared__ u32 s_total;
    const u32 n_entries = st.n_parts * st.n_cta;
    const u64 total_items = chunk_start[n_entries];
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    for (u64 item = blockIdx.x; item < total_items; item += gridDim.x) {
        const u32 entry = __ldg(&item_entry[item]);
        const u32 group = entry / st.n_cta;
        u64 n = st.count[entry]; if (n > st.sub_cap) n = st.sub_cap;
        const u64 first = (item - __ldg(&chunk_start[entry])) * TS_TILE;
        const u64 *src = st.recs + st.sub_index(group, entry - group * st.n_cta) * st.sub_cap + first;
        const u32 cnt = (u32)(n - first < (u64)TS_TILE ? n - first : (u64)TS_TILE);
        const u32 slice0 = group << t.group_shift;
        u64 rec[TS_R]; u32 sub[TS_R];
#pragma unroll
        for (int u = 0; u < TS_R; ++u) {
            const u32 idx = (u32)u * TS_TPB + threadIdx.x;
            rec[u] = idx < cnt ? ld_nc64(src + idx) : 0ull;
        }
        for (u32 i = threadIdx.x; i < gp; i += TS_TPB) hist[i] = 0;
        __syncthreads();                                   // also orders the previous tile's copy-out before this tile's writes
#pragma unroll
        for (int u = 0; u < TS_R; ++u) {
            const u32 idx = (u32)u * TS_TPB + threadIdx.x;
            sub[u] = TS_NONE;
            if (idx < cnt) { const u64 key[1] = {rec[u] & ~1ull}; sub[u] = part_of(place_hash<1>(key), t.n_parts) - slice0; }
        }
        tile_sort_flush(rec, sub, gp, buf, dst, hist, gptr, room, s_wsum, &s_total,
            [&](u32 b, u32 c, u32 &rm) -> u64 * {
                const u32 base = atomicAdd(&s2.count[slice0 + b], c);
                rm = base < s2.cap ? s2.cap - base : 0u;
                return s2.recs + (u64)(slice0 + b) * s2.cap + base;
            },
            [&](u64 r, u32) {                              // bucket full: insert directly (nothing else touches the table now)
                Rec<1, false> rr; rr.w[0] = r;
                insert_record<1, false>(t, rr, lc.unique, lc.full, lc.probes);
                lc.direct++;
            });
    }
    ctr_commit(ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// Fast phase 1b (k_kmer_tiles): position-parallel k-mer extraction.  The batch's base bytes are cut into tiles of
// TS_TILE positions; a CTA loads a tile (+ k-1 bytes of halo) and the tile's "counted" bits (phase 1a) into shared memory
// with coalesced loads; every thread packs the 8+k-1 bases of its 8 consecutive positions to 2 bits (SWAR, a1:
// TwoBitSequence::compressSequence src/TwoBitSequence.cpp:242-269), cuts the forward k-mer of each position out of the
// packed window with a funnel shift, gets the reverse complement by bit reversal, takes the smaller (a2: KmerArrayPair::build,
// buildLeastComplement src/Kmer.h:1323-1375,356-364) and hands the records to the tile sort, which writes each group's
// records of the tile as one contiguous run into the CTA's sub-region of that group.  No read structure is needed: the
// "counted" bit of a position is only set where a k-mer of a non-discarded read starts.
// Multi-GPU: a record owned by another rank (a5: src/Kmer.h:2284-2295) sorts into bin n_groups + owner = that rank's send segment.
// ------------------------------------------------------------------------------------------------
struct TileArgs {
    u64 byte0, byte1;      // positions [byte0, byte1) of the batch belong to this launch
    u64 tile0;             // first tile (tile index = position / TS_TILE)
    u64 n_tiles;
};

// 8 ASCII bases (little-endian in w: first base in the low byte) -> 16 bits, first base in the top two bits; markups -> A
__device__ __forceinline__ u32 pack8(u64 w)
{
    const u64 up = w & 0xDFDFDFDFDFDFDFDFull;
    const u64 valid = bytes_eq(up, 0x4141414141414141ull) | bytes_eq(up, 0x4343434343434343ull) |
                      bytes_eq(up, 0x4747474747474747ull) | bytes_eq(up, 0x5454545454545454ull);
    u64 c = (w >> 1) & 0x0303030303030303ull;
    c ^= (c >> 1) & 0x0101010101010101ull;
    c &= (valid >> 7) * 3ull;
    c = ((c << 2) | (c >> 8)) & 0x000F000F000F000Full;            // nibble j = b(2j)<<2 | b(2j+1) at bit 16j
    return (u32)((c * 0x1000010000100001ull) >> 48);
}

template <bool DIST>
__global__ void __launch_bounds__(TS_TPB, 4) k_kmer_tiles(ParseArgs a, TileArgs ta)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);          // [TS_TILE]
    u64 *dst = buf + TS_TILE;                              // [TS_TILE]
    u64 *sb = dst + TS_TILE;                               // [TS_TILE/8 + 6] raw aligned words of the tile's bases (+ halo)
    u32 *smask = reinterpret_cast<u32 *>(sb + TS_TILE / 8 + 6);   // [TS_TILE/32]
    const u32 n_groups = a.stage.n_parts;
    const u32 nb = n_groups + (DIST ? a.nranks : 0u);
    u32 *hist = smask + TS_TILE / 32;                      // [nb]
    u32 *room = hist + nb;                                 // [nb]
    u32 *cnt = room + nb;                                  // [nb] fill level of this CTA's sub-regions / send segments
    u64 *gptr = reinterpret_cast<u64 *>(cnt + nb + (nb & 1u));    // [nb]
    __shared__ u32 s_wsum[TS_TPB / 32];
    __shared__ u32 s_total;
    for (u32 i = threadIdx.x; i < n_groups; i += TS_TPB) cnt[i] = a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x];
    if (DIST) for (u32 i = threadIdx.x; i < a.nranks; i += TS_TPB) cnt[n_groups + i] = a.seg_count[(size_t)i * gridDim.x + blockIdx.x];
    const u32 k = a.k;
    const u64 keymask = ~0ull << a.pad;                    // top 2k bits
    LocalCtr lc{0, 0, 0, 0, 0, 0};
    const unsigned long long gbeg = (unsigned long long)a.bases, gend = gbeg + a.total_bytes;
    for (u64 tile = ta.tile0 + blockIdx.x; tile < ta.tile0 + ta.n_tiles; tile += gridDim.x) {
        const u64 g0 = tile * TS_TILE;
        __syncthreads();                                   // previous tile's copy-out and window reads are complete
        {   // aligned 8-byte words covering bytes [g0, g0 + TS_TILE + 40)
            const unsigned long long A = gbeg + g0, A0 = A & ~7ull;
            for (u32 i = threadIdx.x; i < TS_TILE / 8 + 6; i += TS_TPB) {
                const unsigned long long q = A0 + 8ull * i;
                sb[i] = (q + 8 > (gbeg & ~7ull) && q < ((gend + 7ull) & ~7ull)) ? ld_nc64(reinterpret_cast<const u64 *>(q)) : 0ull;
            }
            for (u32 i = threadIdx.x; i < TS_TILE / 32; i += TS_TPB) smask[i] = __ldg(&a.mask[(g0 >> 5) + i]);
            for (u32 i = threadIdx.x; i < nb; i += TS_TPB) hist[i] = 0;
        }
        __syncthreads();
        u64 rec[TS_R]; u32 bin[TS_R];
#pragma unroll
        for (int u = 0; u < TS_R; ++u) { rec[u] = 0; bin[u] = TS_NONE; }
        const u32 p0 = threadIdx.x * TS_R;                 // first position of this thread inside the tile
        u32 mbits = (smask[p0 >> 5] >> (p0 & 31u)) & 0xffu;
        {   // positions outside [byte0, byte1) belong to another launch of the same batch
            const u64 ap = g0 + p0;
            if (ap + TS_R <= ta.byte0 || ap >= ta.byte1) mbits = 0;
            else if (ap < ta.byte0 || ap + TS_R > ta.byte1) {
#pragma unroll
                for (int u = 0; u < TS_R; ++u) if (ap + u < ta.byte0 || ap + u >= ta.byte1) mbits &= ~(1u << u);
            }
        }
        if (mbits) {
            // the 40 bases starting at position p0: byte offset inside sb = (A & 7) + p0
            const u32 bo = (u32)((gbeg + g0) & 7ull) + p0;
            const u32 wi = bo >> 3, sh = (bo & 7u) * 8u;
            u64 w[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = sb[wi + i];
            u32 v[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) v[i] = pack8(sh ? (w[i] >> sh) | (w[i + 1] << (64u - sh)) : w[i]);
            const u64 hi = ((u64)v[0] << 48) | ((u64)v[1] << 32) | ((u64)v[2] << 16) | (u64)v[3];
            const u64 lo = (u64)v[4] << 48;
#pragma unroll
            for (int u = 0; u < TS_R; ++u) {
                if ((mbits >> u) & 1u) {
                    const u64 f = (u ? (hi << (2 * u)) | (lo >> (64 - 2 * u)) : hi) & keymask;
                    u64 r = __brevll(~(f >> a.pad) & (~0ull >> a.pad));               // reversed bits, left-aligned, pairs swapped
                    r = ((r & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((r & 0x5555555555555555ull) << 1);
                    const bool fwd = f <= r;
                    const u64 key[1] = {fwd ? f : r};
                    rec[u] = key[0] | (fwd ? 1ull : 0ull);
                    u32 b = part_of(place_hash<1>(key), a.table.n_parts) >> a.table.group_shift;
                    if (DIST) {
                        const u64 h = a.use_lookup8 ? hash_lookup8<1>(key, (int)a.kb) : hash_lookup3<1>(key, (int)a.kb);
                        const u32 own = owner_of(h, a.nranks);
                        if (own != a.rank) b = n_groups + own;
                    }
                    bin[u] = b;
                }
            }
        }
        (void)k;
        tile_sort_flush(rec, bin, nb, buf, dst, hist, gptr, room, s_wsum, &s_total,
            [&](u32 b, u32 c, u32 &rm) -> u64 * {
                const u32 base = cnt[b]; cnt[b] = base + c;
                if (!DIST || b < n_groups) {
                    rm = base < a.stage.sub_cap ? a.stage.sub_cap - base : 0u;
                    return a.stage.recs + a.stage.sub_index(b, blockIdx.x) * a.stage.sub_cap + base;
                }
                rm = base < a.seg_cap ? a.seg_cap - base : 0u;
                return a.seg_recs + ((size_t)(b - n_groups) * gridDim.x + blockIdx.x) * a.seg_cap + base;
            },
            [&](u64 r, u32 b) {
                if (b < n_groups) {                        // sub-region full: insert directly
                    Rec<1, false> rr; rr.w[0] = r;
                    insert_record<1, false>(a.table, rr, lc.unique, lc.full, lc.probes);
                    lc.direct++;
                }                                          // a full send segment is reported by k_compact_send (count > capacity)
            });
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_groups; i += TS_TPB) a.stage.count[(size_t)i * a.stage.n_cta + blockIdx.x] = cnt[i];
    if (DIST) for (u32 i = threadIdx.x; i < a.nranks; i += TS_TPB) a.seg_count[(size_t)i * gridDim.x + blockIdx.x] = cnt[n_groups + i];
    ctr_commit(a.ctr, lc);
}

// ------------------------------------------------------------------------------------------------
// Level 3 (k_count_slices): one CTA per table slice.  The slice (<= 64 KB) is brought into shared memory with
// coalesced 16-byte loads (or zero-filled while the table is still clean), the slice's bucket is streamed through it --
// probe with LDS, claim with a 64-bit shared CAS, count with 32-bit shared atomic adds on the two halves of the value
// word -- and the slice is written back whole.  Replaces KmerSpectrum::append (src/KmerSpectrum.h:1578-1668) +
// KmerMapByKmerArrayPair insert/find (src/Kmer.h:1491-1544,3095-3110) + TrackingDataWithDirection::track
// (src/KmerTrackingData.h:427-448,517-529).
// ------------------------------------------------------------------------------------------------
static constexpr int CNT_TPB = 512;

__global__ void __launch_bounds__(CNT_TPB, 3) k_count_slices(TableView t, Stage2View s2, u32 table_clean, Counters *ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sl4 = reinterpret_cast<uint4 *>(smem_raw);
    u64 *sl = reinterpret_cast<u64 *>(smem_raw);               // slot s: sl[2s] = value word, sl[2s+1] = ~key
    const u32 S = (u32)t.part_slots;
    const bool clean = table_clean && ctr->direct == 0;        // no direct insert has touched the table since the reset
    u64 n_unique = 0, n_full = 0;
    for (u32 p = blockIdx.x; p < t.n_parts; p += gridDim.x) {
        u32 n = s2.count[p];
        if (n == 0) continue;                                  // uniform over the CTA
        if (n > s2.cap) n = s2.cap;
        uint4 *g = reinterpret_cast<uint4 *>(t.slots) + (u64)p * S;
        if (clean) { for (u32 i = threadIdx.x; i < S; i += CNT_TPB) sl4[i] = make_uint4(0, 0, 0, 0); }
        else { for (u32 i = threadIdx.x; i < S; i += CNT_TPB) sl4[i] = __ldcs(g + i); }
        __syncthreads();
        const u64 *src = s2.recs + (u64)p * s2.cap;
        for (u32 i = threadIdx.x; i < n; i += CNT_TPB) {
            const u64 rec = ld_nc64(src + i);
            const u64 key[1] = {rec & ~1ull};
            const u64 want = ~key[0];
            u32 s = (u32)home_slot(place_hash<1>(key), S);
            bool done = false;
            for (u32 probes = 0; probes < S; ++probes) {
                u64 ck = *reinterpret_cast<volatile u64 *>(&sl[2 * s + 1]);
                if (ck == 0ull) {
                    const u64 old = atomicCAS(&sl[2 * s + 1], 0ull, want);
                    if (old == 0ull) { n_unique++; ck = want; } else ck = old;
                }
                if (ck == want) {
                    u32 *v32 = reinterpret_cast<u32 *>(&sl[2 * s]);
                    if (*reinterpret_cast<volatile u32 *>(v32) < MAX_COUNT) {
                        atomicAdd(v32, 1u);
                        if (rec & 1ull) atomicAdd(v32 + 1, 1u);
                    }
                    done = true;
                    break;
                }
                s = s + 1 == S ? 0 : s + 1;
            }
            if (!done) n_full++;
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < S; i += CNT_TPB) __stcs(g + i, sl4[i]);
        __syncthreads();
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
                    const u64 pos = atomicAdd(&a.send_cursor[own], 1ull);
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
