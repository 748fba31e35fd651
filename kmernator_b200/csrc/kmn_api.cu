// kmn_api.cu -- host side of the C ABI declared in include/kmernator_b200.h.
// Owns the device memory plan (count table cut into slices and groups, staging set(s) cut by owner and table group, the
// slice-sorted copy of one batch of groups, input staging slots, multi-GPU receive buffers) and launches the kernels of
// kmn_kernels.cuh: phase 1 (k_weight_mask, k_kmer_scatter) and phase 2 (k_build_entries, k_slice_split*, k_count_slices_*
// or k_insert_staged) on the main stream, H2D staging on two copy streams, and on several GPUs the rounds of the push
// path -- records pushed into their owners' receive buffers over NVLink through CUDA-IPC peer memory by the copy engines
// (comm stream, between two small NCCL all-reduces), or ncclSend / ncclRecv of the same parts when peers cannot be mapped.
#include "../../include/kmernator_b200.h"
#include "kmn_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef KMN_WITH_NCCL
#include <nccl.h>
#endif

using namespace kmn;

static thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct kmn_ctx {
    kmn_opts o;
    int W = 1, kb = 0, pad = 0;
    bool hasx = false, ext = false, weights = false;
    int RW = 1;
    int device = 0, n_sms = 0;
    cudaStream_t stream = nullptr;        // main stream: phase 1 (parse), scans, lookup pass; the one kmn_stream() returns
    cudaStream_t s_insert = nullptr;      // phase 2 (insert) runs here so that it overlaps phase 1 of the next sub-batch
    cudaStream_t s_copy = nullptr;        // H2D staging of host inputs, overlapping the kernels of the previous batch
    bool pipeline = false;                // KMN_PIPELINE=1: two staging sets + insert stream (phase 2 of one sub-batch overlaps phase 1 of the
                                          // next; measured slower than one set on one stream on B200: both phases want the same SM resources)
    int n_sets = 1;
    std::string err;
    uint64_t launches = 0;
    // optional per-kernel timing
    bool prof_on = false;
    struct ProfEv { cudaEvent_t a, b; int kind; uint64_t units; };
    std::vector<ProfEv> prof_events;
    kmn_profile prof_acc{};

    // table
    TableView table{};
    uint64_t n_slots = 0;
    size_t slot_bytes = 0;
    // staging: two sets, so that phase 2 drains one while phase 1 fills the other
    struct StageSet {
        StageView v{};
        uint64_t staged_upper = 0;    // upper bound of records currently staged in this set
        cudaEvent_t ev_parsed = nullptr, ev_drained = nullptr;
        bool drain_pending = false;   // a drain was submitted on s_insert and the main stream has not waited for it yet
    } sets[2];
    int cur = 0;
    uint64_t stage_keys = 0;          // record capacity of ONE set (= the sub-batch size of the pipeline)
    u64 *chunk_start = nullptr, *next_item = nullptr;
    u64 *ent_ptr = nullptr; u32 *ent_cnt = nullptr;     // phase-2 work list entries (k_build_entries)
    u32 *coarse = nullptr;                              // entry of every 64th chunk of the work list
    uint64_t n_groups = 0;
    int insert_ctas = 8;              // phase-2 CTAs per SM
    Counters *ctr = nullptr;
    double *ptab = nullptr;
    u64 *scratch = nullptr;           // small device scalars
    // phase-1 launch geometry
    int n_cta = 0;                    // grid of k_kmer_scatter / k_route_records = staging sub-regions per partition
    int scatter_tpb = 512, scatter_ctas = 2;   // phase-1b CTA size and CTAs per SM (KMN_SCATTER_TPB / KMN_SCATTER_CTAS)
    uint32_t ring_R = 0;              // record slots per bin ring in phase 1b (0: no rings, every record stored directly)
    // phase 2 in shared memory (k <= 31 without extra record word): second split by table slice + counting per slice
    bool smem_count = false;
    u64 *l2buf = nullptr;             // [n_groups][S][slices per group][cap2] records, sorted by slice
    u32 *cnt2 = nullptr;              // [n_groups][S][slices per group]
    u32 *tickets = nullptr;           // [2] work tickets of k_slice_split / k_count_slices
    uint32_t split_S = 4, split_cap2 = 0, split_R = 0, split_batches = 1, split_gpb = 0;
    int split_tpb = 1024, split_ctas = 1, count_ctas = 3;
    bool count_tma = true;            // k_count_slices_tma (bulk-copy engine) instead of k_count_slices (KMN_COUNT_TMA=0)
    bool smem_k32 = false;            // shared-memory slices for k = 32 on one GPU (k_slice_split2<true> + k_count_slices_k32)
    bool smem_w2 = false;             // shared-memory slices for two-word keys on one GPU (k_slice_split2 + k_count_slices_w2; KMN_SMEM_COUNT_W2=0: off)
    int count_db = 0;                 // KMN_COUNT_DB=1: k_count_slices_db (one CTA per SM, two slice buffers) instead of k_count_slices_tma
    int count_ws = 2;                 // k_count_slices_ws: 2 = producer warp + plain consumers (default), 1 = lane-persistent consumers, 0 = k_count_slices_tma
    size_t split_smem = 0;
    uint32_t zero_below = 0;
    size_t scatter_smem = 0;
    DevBuf mask, wts;                 // phase 1a -> 1b: "counted" bits (and fp32 weights for KMN_VALUE_WEIGHTS)
    // input staging (host inputs), double-buffered: the copy of batch b+1 overlaps the kernels of batch b
    static constexpr int IN_SLOTS = 4;    // host batches in flight: copies run ahead of the kernels by up to three batches
    DevBuf in_bases[IN_SLOTS], in_quals[IN_SLOTS], in_off[IN_SLOTS], in_disc[IN_SLOTS];
    DevBuf in_packed[IN_SLOTS], in_poff[IN_SLOTS], in_mpos[IN_SLOTS], in_mchr[IN_SLOTS];     // kmn_count_batch_2na
    cudaEvent_t ev_in_ready[IN_SLOTS] = {}, ev_in_ready2[IN_SLOTS] = {}, ev_in_free[IN_SLOTS] = {};
    bool in_used[IN_SLOTS] = {};
    int in_cur = 0;
    cudaStream_t s_copy2 = nullptr;       // the qualities travel on a second copy stream (two DMA engines feed PCIe)
    // lookup pass scratch
    DevBuf vals, first_nx, out_off, out_len, out_score, out_trim, lk_keys, lk_out;
    DevBuf lk_origin, lk_resp_in, lk_resp_out;   // multi-GPU lookup pass: request origins, answers in / out
    uint32_t purged_depth = 0;
    bool finished = false;
    // multi-GPU
    int rank = 0, nranks = 1;
    u64 *send_recs = nullptr, *send_cursor = nullptr, *recv_recs = nullptr, *all_counts = nullptr;
    uint64_t send_cap = 0, recv_cap = 0;
    // multi-GPU push path: phase 1 bins by (owner, group); the other owners' parts are written into their receive
    // buffers over NVLink (peer pointers from CUDA IPC), sorted by group, and inserted from there
    bool p2p = false;                     // more than one rank: phase 1 bins by (owner, group) and the rounds below move the parts
    bool ipc = false;                     // the peers' receive buffers are mapped (CUDA IPC): parts travel by copy engine or by
                                          // k_push_copy over NVLink; otherwise they travel as ncclSend / ncclRecv
    int pending_set = -1, pending_rb = -1;   // the round whose phase 2 has not been submitted yet (it follows the NEXT phase 1)
    cudaStream_t s_comm = nullptr;        // push kernels + the NCCL barriers of a round
    cudaStream_t s_comm2 = nullptr;       // second copy stream: half of the peers, so two copy engines feed NVLink at once
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void *recv_all = nullptr;             // [2 round buffers][nranks sources][push_cap records] + [2][nranks][n_groups+1] offsets
    void *peer_all[KMN_MAX_PUSH_RANKS] = {nullptr};   // recv_all of every peer (cudaIpcOpenMemHandle)
    void *peer_table[KMN_MAX_PUSH_RANKS] = {};   // count tables of all ranks (own pointer at [rank]), mapped with CUDA IPC
    u64 *bar_buf = nullptr;               // [2] operands of the stream-ordered barriers around a peer-memory lookup pass
    bool peer_lookup = false;             // the lookup pass probes the owners' tables over NVLink (no request/response rounds)
    uint64_t push_cap = 0;                // records per (round buffer, source)
    size_t push_meta = 0;                 // u32 words of meta per (round buffer, source)
    bool push_ce = true;                  // transport: copy engines move whole parts (default) / k_push_copy writes sorted runs
    u32 *run_off = nullptr, *grp_off = nullptr;
    u64 *flags = nullptr;                 // [0] records lost to a full remote sub-region, [1] records beyond push_cap
    u64 *d_const = nullptr;               // [0] = 0, [1] = 1, [2] = barrier scratch, [3] = sum of "done" flags
    uint64_t round = 0;
    int round_split = 1;                  // launches per kernel class and round (KMN_ROUND_SPLIT; measured: no gain)
    uint32_t cta_rot = 0;                 // first CTA of the next phase-1b launch
    cudaEvent_t ev_pushed[2] = {nullptr, nullptr}, ev_rb_free[2] = {nullptr, nullptr}, ev_recv[2] = {nullptr, nullptr};
    bool push_pending[2] = {false, false}, rb_busy[2] = {false, false};
#ifdef KMN_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
};

static int fail(kmn_ctx *c, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(ctx, call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? KMN_ERR_NOMEM : KMN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                         \
    } while (0)

// RAII timer around one kernel launch (no-op unless profiling is enabled)
struct ProfScope {
    kmn_ctx *c; cudaEvent_t a = nullptr, b = nullptr; int kind; uint64_t units; cudaStream_t st;
    ProfScope(kmn_ctx *c_, int kind_, uint64_t units_, cudaStream_t st_ = nullptr) : c(c_), kind(kind_), units(units_), st(st_ ? st_ : c_->stream)
    {
        if (!c->prof_on) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~ProfScope()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        c->prof_events.push_back({a, b, kind, units});
    }
};

static int ensure(kmn_ctx *c, DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes && b.p) return 0;
    if (b.p) { CK(c, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    CK(c, cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// dispatch on (W, HASX, EXT, DIST).  -DKMN_ONLY_W1 (developer builds) instantiates k <= 32 only.
#ifdef KMN_ONLY_W1
#define KMN_WIDE_CASES(...)
#else
#define KMN_WIDE_CASES(...)                                            \
    case 2: { constexpr int W_ = 2; __VA_ARGS__; } break;              \
    case 3: { constexpr int W_ = 3; __VA_ARGS__; } break;              \
    case 4: { constexpr int W_ = 4; __VA_ARGS__; } break;
#endif
#define KMN_DISPATCH_W(c, ...)                                         \
    switch ((c)->W) {                                                  \
    case 1: { constexpr int W_ = 1; __VA_ARGS__; } break;              \
    KMN_WIDE_CASES(__VA_ARGS__)                                        \
    default: return fail(c, KMN_ERR_INVALID, "unsupported key width"); \
    }
#define KMN_DISPATCH_X(c, ...)                                   \
    if ((c)->hasx) { constexpr bool X_ = true; __VA_ARGS__; }    \
    else { constexpr bool X_ = false; __VA_ARGS__; }

const char *kmn_version(void) { return "kmernator_b200 0.1 (sm_100a)"; }

void kmn_default_opts(kmn_opts *o)
{
    memset(o, 0, sizeof *o);
    o->struct_size = sizeof *o;
    o->kmer_size = 31;
    o->fastq_start_char = 33;
    o->min_quality_score = 3;       // src/Options.h:329
    o->min_kmer_quality = 0.10f;    // src/KmerSpectrum.h:92
    o->min_depth = 2;               // src/KmerSpectrum.h:92
    o->hash_kind = KMN_HASH_LOOKUP3_HASHLITTLE2;
    o->value_kind = KMN_VALUE_DIR;
    o->slice_bytes = 64u << 20;
}

const char *kmn_last_error(const kmn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void *kmn_stream(kmn_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t kmn_launch_count(const kmn_ctx *ctx) { return ctx ? ctx->launches : 0; }

// staging capacity: stage_keys records per set.  Pipelined (two sets): one set is one sub-batch (phase 2 drains it while
// phase 1 fills the other), so the default aims at ~8 sub-batches over the expected input.  On the multi-GPU push path a
// set is cut by owner rank as well ([owner][cta][group] sub-regions) and one launch fills one set.
static int alloc_stage_sets(kmn_ctx *c)
{
    size_t free_b = 0, total_b = 0;
    for (int si = 0; si < 2; ++si) {
        kmn_ctx::StageSet &st = c->sets[si];
        if (st.v.recs) { CK(c, cudaFree(st.v.recs)); st.v.recs = nullptr; }
        if (st.v.count) { CK(c, cudaFree(st.v.count)); st.v.count = nullptr; }
        if (st.v.ovf_recs) { CK(c, cudaFree(st.v.ovf_recs)); st.v.ovf_recs = nullptr; }
        if (st.v.ovf_count) { CK(c, cudaFree(st.v.ovf_count)); st.v.ovf_count = nullptr; }
        st.v.ovf_cap = 0;
    }
    CK(c, cudaMemGetInfo(&free_b, &total_b));
    const kmn_opts &o = c->o;
    uint64_t sk = o.stage_keys;
    if (!sk) {
        sk = (o.est_raw_kmers ? o.est_raw_kmers : (1ull << 22)) / (c->pipeline ? 8 : 4);
        uint64_t lim = (uint64_t)(0.22 * (double)free_b / (double)(c->RW * 8) / (double)c->n_sets);    // the slice-sorted copy is as large again
        if (sk > lim) sk = lim;
    }
    if (sk < (1ull << 16)) sk = 1ull << 16;
    c->stage_keys = sk;
    const uint64_t n_groups = c->n_groups, n_cta = (uint64_t)c->n_cta, n_own = c->p2p ? (uint64_t)c->nranks : 1;
    const uint64_t per_sub = sk / n_groups / n_cta / n_own;
    // mean + slack for the spread; a multiple of 4 records, so that every sub-region starts on a 32-byte sector boundary
    const uint64_t sub_cap = (per_sub + per_sub / 8 + 8 * (uint64_t)std::sqrt((double)per_sub + 1.0) + 64 + 3) & ~3ull;
    if (sub_cap >= (1ull << 31)) return fail(c, KMN_ERR_INVALID, "staging sub-region too large (%llu records)", (unsigned long long)sub_cap);
    for (int si = 0; si < c->n_sets; ++si) {
        kmn_ctx::StageSet &st = c->sets[si];
        st.v.sub_cap = (u32)sub_cap; st.v.n_cta = (u32)n_cta; st.v.n_parts = (u32)n_groups;
        st.v.n_owners = (u32)n_own; st.v.me = (u32)c->rank;
        CK(c, cudaMalloc((void **)&st.v.recs, (size_t)n_own * n_groups * n_cta * sub_cap * c->RW * 8));
        CK(c, cudaMalloc((void **)&st.v.count, (size_t)n_own * n_groups * n_cta * 4));
        CK(c, cudaMemsetAsync(st.v.count, 0, (size_t)n_own * n_groups * n_cta * 4, c->stream));
        if (n_own > 1 && c->push_ce) {
            // overflow list per owner: records of other owners whose sub-region was full travel ungrouped behind the part
            st.v.ovf_cap = (u32)(std::min<uint64_t>(65536 + sk / n_own / 256, 1u << 30) & ~3ull);   // whole sectors: what follows stays 32-byte aligned
            CK(c, cudaMalloc((void **)&st.v.ovf_recs, (size_t)n_own * st.v.ovf_cap * c->RW * 8));
            CK(c, cudaMalloc((void **)&st.v.ovf_count, (size_t)n_own * 4));
            CK(c, cudaMemsetAsync(st.v.ovf_count, 0, (size_t)n_own * 4, c->stream));
        }
        if (!st.ev_parsed) CK(c, cudaEventCreateWithFlags(&st.ev_parsed, cudaEventDisableTiming));
        if (!st.ev_drained) CK(c, cudaEventCreateWithFlags(&st.ev_drained, cudaEventDisableTiming));
        st.staged_upper = 0;
    }
    // phase-1 shared memory: per (owner, group) bin a position counter, a flush mark and a ring of R record slots
    // (+ send-segment counters, <= 64 ranks).  R = the largest power of two that fits, at most 64; below 4 no rings.
    {
        int dev_smem = 0, sm_smem = 0;
        CK(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        CK(c, cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
        const size_t budget = std::min<size_t>((size_t)dev_smem, ((size_t)sm_smem - 1024u * (size_t)c->scatter_ctas) / (size_t)c->scatter_ctas) - 256;
        const size_t n_bins = (size_t)n_groups * n_own, n_pad = (n_bins + 31) & ~(size_t)31;
        const size_t hdr = (2 * n_pad + 64 + 32) * 4;
        uint32_t R = 64;
        if (const char *e = getenv("KMN_RING")) R = (uint32_t)std::max(0, atoi(e));
        while (R >= 4 && hdr + n_bins * R * c->RW * 8 > budget) R >>= 1;
        if (R < 4 || (R & (R - 1))) R = 0;
        if (hdr > budget) return fail(c, KMN_ERR_INVALID, "too many staging bins (%zu) for the phase-1 shared memory", n_bins);
        c->ring_R = R;
        c->scatter_smem = hdr + n_bins * R * c->RW * 8;
    }
    // shared-memory phase 2: one more split of every group by slice, then counting with the slice in shared memory
    if (c->l2buf) { CK(c, cudaFree(c->l2buf)); c->l2buf = nullptr; }
    if (c->cnt2) { CK(c, cudaFree(c->cnt2)); c->cnt2 = nullptr; }
    // (phase 1 may insert directly into the table when a sub-region overflows, and k_count_slices holds slices in shared
    //  memory: the two must never run at the same time, so the single-GPU pipeline is incompatible and the push path
    //  orders phase 1 behind the drains, see kmn_count_batch)
    c->smem_count = c->W == 1 && !c->hasx && ((c->nranks == 1 && !c->pipeline) || c->p2p) && c->table.part_slots * 16 <= 64 * 1024;
    if (c->smem_w2 && c->nranks == 1 && !c->pipeline && c->table.part_slots * 24 <= 48 * 1024) c->smem_count = true;
    // k = 32: single-word slots, but the strand flag needs a second record word (KMN_SMEM_COUNT_K32=0: L2-resident path)
    c->smem_k32 = c->W == 1 && c->hasx && !c->ext && !c->weights && !(getenv("KMN_SMEM_COUNT_K32") && atoi(getenv("KMN_SMEM_COUNT_K32")) == 0);
    if (c->smem_k32 && c->nranks == 1 && !c->pipeline && c->table.part_slots * 16 <= 64 * 1024) c->smem_count = true; else c->smem_k32 = false;
    const bool w2 = c->smem_count && (c->W == 2 || c->smem_k32);       // 16-byte records
    if (const char *e = getenv("KMN_SMEM_COUNT")) c->smem_count = c->smem_count && atoi(e) != 0;
    if (c->smem_count) {
        int dev_smem = 0;
        CK(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        const size_t nb = (size_t)1 << c->table.group_shift, n_pad = (nb + 31) & ~(size_t)31;
        if (const char *e = getenv("KMN_SPLIT_S")) c->split_S = (uint32_t)std::min(COUNT_MAX_S, std::max(1, atoi(e)));
        if (const char *e = getenv("KMN_SPLIT_TPB")) c->split_tpb = std::min(1024, std::max(64, atoi(e) & ~31));
        if (const char *e = getenv("KMN_SPLIT_CTAS")) c->split_ctas = std::max(1, atoi(e));
        if (c->split_tpb * c->split_ctas > 2048) c->split_ctas = 2048 / c->split_tpb;
        c->split_tpb = SPLIT_TPB; c->split_ctas = 1;
        if (const char *e = getenv("KMN_SPLIT_TPB")) { if (atoi(e) == SPLIT_TPB2) { c->split_tpb = SPLIT_TPB2; c->split_ctas = 2; } }
        const size_t rec_b = w2 ? 16 : 8;
        const size_t in_bytes = 2 * (size_t)c->split_tpb * (w2 ? SPLIT2_RPT : SPLIT_RPT) * rec_b;  // two input buffers of one round each
        int sm_smem2 = 0;
        CK(c, cudaDeviceGetAttribute(&sm_smem2, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
        const size_t budget = std::min<size_t>((size_t)dev_smem, ((size_t)sm_smem2 - 1024u * (size_t)c->split_ctas) / (size_t)c->split_ctas) - 2048;
        uint32_t R = 32;
        while (R >= 4 && in_bytes + 2 * n_pad * 4 + nb * R * rec_b > budget) R >>= 1;
        if (R < 4) c->smem_count = false;
        c->split_R = R;
        c->split_smem = in_bytes + 2 * n_pad * 4 + nb * R * rec_b;
        // records of one drain per sub-run: stage_keys / (slices * S), plus slack for the spread (a sub-run that fills up
        // sends its records straight to the table)
        const uint64_t m2 = sk / c->table.n_parts / c->split_S + 1;
        c->split_cap2 = (uint32_t)((m2 + m2 / 8 + 8 * (uint64_t)std::sqrt((double)m2) + 32 + 3) & ~3ull);
        // the slice-sorted copy holds one BATCH of groups at a time (split, count, next batch): 1/split_batches of a drain
        c->split_batches = 2;
        if (const char *e = getenv("KMN_SPLIT_BATCHES")) c->split_batches = (uint32_t)std::max(1, atoi(e));
        if (const char *e = getenv("KMN_COUNT_TMA")) { if (atoi(e) == 0) c->split_batches = 1; }     // (k_count_slices walks the whole table)
        c->split_batches = (uint32_t)std::min<uint64_t>(c->split_batches, n_groups);
        c->split_gpb = (uint32_t)((n_groups + c->split_batches - 1) / c->split_batches);
        const size_t n_sub2 = (size_t)c->split_gpb * c->split_S * nb;
        if (c->smem_count) {
            if (cudaMalloc((void **)&c->l2buf, n_sub2 * c->split_cap2 * rec_b) != cudaSuccess) { cudaGetLastError(); c->l2buf = nullptr; c->smem_count = false; }
        }
        if (c->smem_count) {
            CK(c, cudaMalloc((void **)&c->cnt2, n_sub2 * 4));
            if (!c->tickets) CK(c, cudaMalloc((void **)&c->tickets, 64));
            if (w2) {
                CK(c, cudaFuncSetAttribute(k_slice_split2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->split_smem));
                CK(c, cudaFuncSetAttribute(k_slice_split2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->split_smem));
                CK(c, cudaFuncSetAttribute(k_count_slices_k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16 + COUNT32_NBUF * COUNT2_CHUNK * 16)));
                CK(c, cudaFuncSetAttribute(k_count_slices_w2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((c->table.part_slots * 24 + 127) & ~(size_t)127) + COUNT2_NBUF * COUNT2_CHUNK * 16)));
            }
            CK(c, cudaFuncSetAttribute(k_slice_split<SPLIT_TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->split_smem));
            CK(c, cudaFuncSetAttribute(k_slice_split<SPLIT_TPB2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->split_smem));
            CK(c, cudaFuncSetAttribute(k_count_slices, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16)));
            CK(c, cudaFuncSetAttribute(k_count_slices_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16 + COUNT_NBUF * COUNT_CHUNK * 8)));
            CK(c, cudaFuncSetAttribute(k_count_slices_ws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16 + COUNTW_NBUF * COUNTW_CHUNK * 8)));
            CK(c, cudaFuncSetAttribute(k_count_slices_ws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16 + COUNTW_NBUF * COUNTW_CHUNK * 8)));
            CK(c, cudaFuncSetAttribute(k_count_slices_ws<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->table.part_slots * 16 + COUNTW_NBUF * COUNTW_CHUNK * 8)));
            if (const char *e = getenv("KMN_COUNT_TMA")) c->count_tma = atoi(e) != 0;
            if (const char *e = getenv("KMN_COUNT_WS")) c->count_ws = atoi(e);
            if (const char *e = getenv("KMN_COUNT_DB")) c->count_db = atoi(e);
            {
                const size_t need = c->table.part_slots * 32 + (size_t)COUNTD_NBUF * COUNTD_CHUNK * 8;
                if (need + 1024 > (size_t)dev_smem) c->count_db = 0;
                else CK(c, cudaFuncSetAttribute(k_count_slices_db, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            }
            if ((c->table.part_slots * 16) % 16 != 0) c->count_tma = false;
        }
    }
    return 0;
}

static int plan_and_alloc(kmn_ctx *c)
{
    const kmn_opts &o = c->o;
    c->W = (int)((o.kmer_size + 31) / 32);
    c->kb = (int)((o.kmer_size + 3) / 4);
    c->pad = 64 * c->W - 2 * (int)o.kmer_size;
    c->ext = (o.value_kind & KMN_VALUE_DIR_EXT) != 0;
    c->weights = (o.value_kind & KMN_VALUE_WEIGHTS) != 0;
    c->hasx = c->ext || c->weights || (o.kmer_size % 32 == 0);
    c->RW = c->W + (c->hasx ? 1 : 0);
    c->slot_bytes = 8 * (size_t)(c->W + 1);

    size_t free_b = 0, total_b = 0;
    CK(c, cudaMemGetInfo(&free_b, &total_b));

    // table capacity: explicit, or the reference's own sizing guess (weak = est/estimatedDepth(20), singleton =
    // est*estimatedErrorRate(0.35), src/KmerSpectrum.h:414-421) at load factor 0.5, bounded by 40% of free memory
    uint64_t slots = o.table_slots;
    const size_t per_slot = c->slot_bytes + (c->weights ? 4 : 0) + (c->ext ? 48 : 0);
    if (!slots) {
        double est = (double)(o.est_raw_kmers ? o.est_raw_kmers : (1ull << 20));
        double distinct = est / 20.0 + est * 0.35;
        slots = (uint64_t)(distinct / 0.5) + 1024;
        uint64_t lim = (uint64_t)(0.40 * (double)free_b / (double)per_slot);
        if (slots > lim) slots = lim;
    }
    if (slots < 1024) slots = 1024;

    // slices of <= SLICE_SLOTS slots (one slice fits in shared memory; probing wraps inside a slice), grouped into
    // L2-sized groups of ~slice_bytes: the level-1 staging is partitioned by group and every phase-1 CTA keeps one
    // 4-byte fill counter per group in shared memory
    int dev_smem = 0;
    CK(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const uint32_t gbytes = o.slice_bytes ? o.slice_bytes : (64u << 20);
    // (two-word keys: slices of 2048 slots = 48 KB so that a slice fits in shared memory twice per SM; KMN_SMEM_COUNT_W2=0: 4096 slots, L2-resident path)
    c->smem_w2 = c->W == 2 && !c->hasx && !(getenv("KMN_SMEM_COUNT_W2") && atoi(getenv("KMN_SMEM_COUNT_W2")) == 0);
    uint64_t part_slots = std::min<uint64_t>(c->smem_w2 ? SLICE_SLOTS / 2 : SLICE_SLOTS, slots) & ~1ull;   // even: W == 1 probes aligned pairs of slots
    uint64_t n_parts = (slots + part_slots - 1) / part_slots;
    if (n_parts >= (1ull << 31)) return fail(c, KMN_ERR_INVALID, "table too large (%llu slices)", (unsigned long long)n_parts);
    slots = part_slots * n_parts;
    uint32_t gshift = 0;
    while (gshift < 11 && (part_slots * c->slot_bytes << (gshift + 1)) <= (uint64_t)gbytes) gshift++;     // <= 2048 slices per group
    if (const char *e = getenv("KMN_SCATTER_TPB")) c->scatter_tpb = std::min(SCATTER_MAX_TPB, std::max(64, atoi(e) & ~31));
    if (const char *e = getenv("KMN_SCATTER_CTAS")) c->scatter_ctas = std::max(1, atoi(e));
    if (c->scatter_tpb * c->scatter_ctas > 2048) c->scatter_ctas = 2048 / c->scatter_tpb;
    const uint64_t p_max = ((size_t)dev_smem / c->scatter_ctas - 2048) / 8 - 64;      // counter + flush mark per bin
    while (((n_parts + (1ull << gshift) - 1) >> gshift) > p_max) gshift++;
    const uint64_t n_groups = (n_parts + (1ull << gshift) - 1) >> gshift;
    c->n_slots = slots;
    c->table.part_slots = part_slots;
    c->table.n_parts = (u32)n_parts;
    c->table.group_shift = gshift;
    c->n_cta = c->n_sms * c->scatter_ctas;

    CK(c, cudaMalloc(&c->table.slots, slots * c->slot_bytes));
    if (c->weights) CK(c, cudaMalloc((void **)&c->table.wsum, slots * 4));
    if (c->ext) CK(c, cudaMalloc((void **)&c->table.ext, slots * 48));

    c->n_groups = n_groups;
    { int r = alloc_stage_sets(c); if (r) return r; }
    {
        const size_t max_entries = (size_t)n_groups * (size_t)c->n_cta * KMN_MAX_PUSH_RANKS + 64;
        CK(c, cudaMalloc((void **)&c->chunk_start, (max_entries + 1) * 8));
        CK(c, cudaMalloc((void **)&c->ent_ptr, max_entries * 8));
        CK(c, cudaMalloc((void **)&c->ent_cnt, max_entries * 4));
        // chunks of one drain <= records of a set / chunk + one partial chunk per entry (sets never grow after this point:
        // the owner-cut sets of the multi-GPU paths hold the same stage_keys)
        const size_t max_chunks = (size_t)(2 * c->stage_keys + (64ull << 20)) / INSERT_CHUNK + max_entries;
        CK(c, cudaMalloc((void **)&c->coarse, ((max_chunks >> COARSE_SHIFT) + 16) * 4));
    }
    CK(c, cudaMalloc((void **)&c->next_item, 64));
    CK(c, cudaMalloc((void **)&c->ctr, sizeof(Counters)));
    CK(c, cudaMalloc((void **)&c->scratch, 64));
    CK(c, cudaMalloc((void **)&c->ptab, 256 * sizeof(double)));

    // Read::qualityToProbability, computed on the host with the same libm the reference uses (src/Sequence.cpp:522-540)
    double p[256];
    for (int i = 0; i < 256; i++) p[i] = 0.0;
    const int start = (int)o.fastq_start_char;
    for (int i = start + (int)o.min_quality_score; i < 103 && i < 256; i++) p[i] = 1.0 - pow(10.0, (start - i) / 10.0);
    for (int i = 103; i < 256; i++) p[i] = 1.0;
    if (o.ignore_quality) for (int i = 0; i < 256; i++) p[i] = 1.0;
    c->zero_below = 0;                      // p[q]==0 exactly for q < zero_below (and for no other q)
    while (c->zero_below < 256 && p[c->zero_below] == 0.0) c->zero_below++;
    for (int i = (int)c->zero_below; i < 256; i++) if (p[i] == 0.0) return fail(c, KMN_ERR_INVALID, "quality table not monotone");
    CK(c, cudaMemcpy(c->ptab, p, sizeof p, cudaMemcpyHostToDevice));
    return 0;
}

static int wait_drains(kmn_ctx *c);
#ifdef KMN_WITH_NCCL
static int flush_pending_insert(kmn_ctx *c);
#endif

int kmn_reset(kmn_ctx *c)
{
    if (!c) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    int r = wait_drains(c); if (r) return r;          // phase-2 work still in flight must not see the cleared table
#ifdef KMN_WITH_NCCL
    r = flush_pending_insert(c); if (r) return r;     // a round whose phase 2 was deferred behind the next phase 1
#endif
    for (int si = 0; si < 2; ++si)                     // nor may a push still read the fill counters cleared below
        if (c->push_pending[si]) { CK(c, cudaStreamWaitEvent(c->stream, c->ev_pushed[si], 0)); c->push_pending[si] = false; }
    CK(c, cudaMemsetAsync(c->table.slots, 0, c->n_slots * c->slot_bytes, c->stream));
    if (c->table.wsum) CK(c, cudaMemsetAsync(c->table.wsum, 0, c->n_slots * 4, c->stream));
    if (c->table.ext) CK(c, cudaMemsetAsync(c->table.ext, 0, c->n_slots * 48, c->stream));
    for (int si = 0; si < c->n_sets; ++si) {
        kmn_ctx::StageSet &st = c->sets[si];
        CK(c, cudaMemsetAsync(st.v.count, 0, (size_t)st.v.n_owners * st.v.n_parts * st.v.n_cta * 4, c->stream));
        if (st.v.ovf_count) CK(c, cudaMemsetAsync(st.v.ovf_count, 0, (size_t)st.v.n_owners * 4, c->stream));
        st.staged_upper = 0;
    }
    if (c->flags) CK(c, cudaMemsetAsync(c->flags, 0, 16, c->stream));
    CK(c, cudaMemsetAsync(c->ctr, 0, sizeof(Counters), c->stream));
    if (c->send_cursor) CK(c, cudaMemsetAsync(c->send_cursor, 0, (size_t)c->nranks * 8, c->stream));
    c->purged_depth = 0;
    c->finished = false;
    return 0;
}

template <int W, bool X>
static int set_smem_attrs(kmn_ctx *c)
{
    const int s = (int)c->scatter_smem;
    if (s <= 48 * 1024) return 0;
    CK(c, cudaFuncSetAttribute(k_kmer_scatter<W, X, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s));
    CK(c, cudaFuncSetAttribute(k_kmer_scatter<W, X, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s));
    if (X) {
        CK(c, cudaFuncSetAttribute(k_kmer_scatter<W, X, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s));
        CK(c, cudaFuncSetAttribute(k_kmer_scatter<W, X, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s));
    }
    return 0;
}

static int apply_smem_attrs(kmn_ctx *c)
{
    const int msm = (int)((MASK_TPB / 32) * MASK_WBUF);
    CK(c, cudaFuncSetAttribute(k_weight_mask<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, msm));
    CK(c, cudaFuncSetAttribute(k_weight_mask<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, msm));
    KMN_DISPATCH_W(c, KMN_DISPATCH_X(c, { int r_ = set_smem_attrs<W_, X_>(c); if (r_) return r_; }));
    return 0;
}

int kmn_create(kmn_ctx **out, const kmn_opts *opts)
{
    if (!out || !opts) return fail(nullptr, KMN_ERR_INVALID, "null argument");
    if (opts->struct_size != sizeof(kmn_opts)) return fail(nullptr, KMN_ERR_INVALID, "kmn_opts size mismatch (%u vs %zu)", opts->struct_size, sizeof(kmn_opts));
    if (opts->kmer_size < 1 || opts->kmer_size > 128) return fail(nullptr, KMN_ERR_INVALID, "kmer_size %u out of range 1..128", opts->kmer_size);
    if (opts->fastq_start_char != 33 && opts->fastq_start_char != 64) return fail(nullptr, KMN_ERR_INVALID, "fastq_start_char must be 33 or 64");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(nullptr, KMN_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e)); }
    if ((int)opts->device >= ndev) return fail(nullptr, KMN_ERR_INVALID, "device %u not present (%d devices)", opts->device, ndev);
    kmn_ctx *c = new kmn_ctx();
    c->o = *opts;
    c->device = (int)opts->device;
    int rc = 0;
    do {
        if (cudaSetDevice(c->device) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "cudaSetDevice failed"); break; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "cudaGetDeviceProperties failed"); break; }
        c->n_sms = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "stream create failed"); break; }
        // tuning knobs (bench experiments only; the defaults are the measured best)
        if (const char *e = getenv("KMN_PIPELINE")) c->pipeline = atoi(e) != 0;
        c->n_sets = c->pipeline ? 2 : 1;
        if (const char *e = getenv("KMN_INSERT_CTAS")) c->insert_ctas = std::max(1, atoi(e));
        if (c->pipeline) {
            if (cudaStreamCreateWithFlags(&c->s_insert, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "stream create failed"); break; }
        } else c->s_insert = c->stream;
        if (cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "stream create failed"); break; }
        if (cudaStreamCreateWithFlags(&c->s_copy2, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(nullptr, KMN_ERR_CUDA, "stream create failed"); break; }
        for (int i = 0; i < kmn_ctx::IN_SLOTS; ++i) {
            cudaEventCreateWithFlags(&c->ev_in_ready[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_in_ready2[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_in_free[i], cudaEventDisableTiming);
        }
        rc = plan_and_alloc(c);
        if (rc) break;
        rc = apply_smem_attrs(c);
        if (rc) break;
        rc = kmn_reset(c);
        if (rc) break;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(c, KMN_ERR_CUDA, "sync failed"); break; }
    } while (0);
    if (rc) {
        if (!c->err.empty()) g_create_error = c->err;
        kmn_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

void kmn_destroy(kmn_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    if (c->s_copy2) cudaStreamSynchronize(c->s_copy2);
    if (c->s_insert) cudaStreamSynchronize(c->s_insert);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    if (c->s_comm2) { cudaStreamSynchronize(c->s_comm2); cudaStreamDestroy(c->s_comm2); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int p = 0; p < KMN_MAX_PUSH_RANKS; ++p) if (c->peer_all[p] && p != c->rank) cudaIpcCloseMemHandle(c->peer_all[p]);
    for (int p = 0; p < KMN_MAX_PUSH_RANKS; ++p) if (c->peer_table[p] && p != c->rank) cudaIpcCloseMemHandle(c->peer_table[p]);
    if (c->bar_buf) cudaFree(c->bar_buf);
    if (c->p2p) { c->send_recs = nullptr; c->recv_recs = nullptr; }     // aliases of recv_all
#ifdef KMN_WITH_NCCL
    if (c->comm) ncclCommDestroy(c->comm);
#endif
    void *ptrs[] = {c->l2buf, c->cnt2, c->tickets, c->sets[0].v.ovf_recs, c->sets[0].v.ovf_count, c->sets[1].v.ovf_recs, c->sets[1].v.ovf_count, c->recv_all, c->run_off, c->grp_off, c->flags, c->d_const, c->ent_ptr, c->ent_cnt, c->coarse,c->table.slots, c->table.wsum, c->table.ext, c->sets[0].v.recs, c->sets[0].v.count, c->sets[1].v.recs, c->sets[1].v.count,
                    c->chunk_start, c->next_item,
                    c->ctr, c->scratch, c->ptab, c->send_recs, c->send_cursor, c->recv_recs, c->all_counts,
                    c->vals.p, c->first_nx.p, c->out_off.p,
                    c->lk_origin.p, c->lk_resp_in.p, c->lk_resp_out.p, c->mask.p, c->wts.p,
                    c->out_len.p, c->out_score.p, c->out_trim.p, c->lk_keys.p, c->lk_out.p};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < kmn_ctx::IN_SLOTS; ++i)
        for (DevBuf *b : {&c->in_bases[i], &c->in_quals[i], &c->in_off[i], &c->in_disc[i], &c->in_packed[i], &c->in_poff[i], &c->in_mpos[i], &c->in_mchr[i]})
            if (b->p) cudaFree(b->p);
    for (auto &st : c->sets) { if (st.ev_parsed) cudaEventDestroy(st.ev_parsed); if (st.ev_drained) cudaEventDestroy(st.ev_drained); }
    for (int i = 0; i < kmn_ctx::IN_SLOTS; ++i) {
        if (c->ev_in_ready[i]) cudaEventDestroy(c->ev_in_ready[i]);
        if (c->ev_in_ready2[i]) cudaEventDestroy(c->ev_in_ready2[i]);
        if (c->ev_in_free[i]) cudaEventDestroy(c->ev_in_free[i]);
    }
    if (c->s_copy2) cudaStreamDestroy(c->s_copy2);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_pushed[i]) cudaEventDestroy(c->ev_pushed[i]);
        if (c->ev_rb_free[i]) cudaEventDestroy(c->ev_rb_free[i]);
        if (c->ev_recv[i]) cudaEventDestroy(c->ev_recv[i]);
    }
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    if (c->s_insert && c->s_insert != c->stream) cudaStreamDestroy(c->s_insert);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---------------------------------------------------------------------------------------------------------
// phase 2: drain the staging regions into the table
// ---------------------------------------------------------------------------------------------------------
// phase 2 of one staging set (plus, on the push path, the receive buffer `rb` of this round) on stream si
static int launch_insert(kmn_ctx *c, const StageView &v, int rb, uint64_t units, cudaStream_t si)
{
    RecvView rv{};
    if (rb >= 0) {
        const size_t R = (size_t)c->nranks;
        rv.recs = (const u64 *)c->recv_all + (size_t)rb * R * c->push_cap * c->RW;
        rv.meta = (const u32 *)((const u64 *)c->recv_all + 2 * R * c->push_cap * c->RW) + (size_t)rb * R * c->push_meta;
        rv.cap = c->push_cap; rv.n_src = (u32)R; rv.me = (u32)c->rank; rv.mode = c->push_ce ? 1u : 0u;
        rv.ovf_cap = v.ovf_cap; rv.part_recs = (u64)v.n_cta * v.n_parts * v.sub_cap; rv.meta_stride = c->push_meta;
    }
    const u32 n_entries = v.n_parts * (v.n_cta + (rv.n_src > 1 ? (rv.n_src - 1) * (rv.mode == 1 ? v.n_cta : 1u) : 0)) +
                          (rv.mode == 1 && rv.n_src > 1 ? rv.n_src - 1 : 0u);
    k_build_entries<<<std::min<u32>((n_entries + 255) / 256, (u32)c->n_sms * 4), 256, 0, si>>>(v, rv, (u32)c->RW, c->ent_ptr, c->ent_cnt);
    c->launches++;
    const u64 *ent_ptr = c->ent_ptr;
    const u32 *ent_cnt = c->ent_cnt;
    u32 n_list = n_entries;
    if (c->smem_count) {
        // grouped entries: split every group by slice, then count slice by slice in shared memory; the ungrouped rest
        // (overflow lists of the peers) goes through the generic insert below, after the slices are back in the table
        const u32 per_group = v.n_cta + (rv.n_src > 1 ? (rv.n_src - 1) * (rv.mode == 1 ? v.n_cta : 1u) : 0);
        const u32 n_grouped = v.n_parts * per_group;
        SplitArgs sa{};
        sa.table = c->table; sa.ent_ptr = c->ent_ptr; sa.ent_cnt = c->ent_cnt; sa.per_group = per_group;
        sa.S = std::min<u32>(c->split_S, per_group); sa.epp = (per_group + sa.S - 1) / sa.S;
        sa.buf = c->l2buf; sa.cnt2 = c->cnt2; sa.cap2 = c->split_cap2; sa.ring_R = c->split_R; sa.ticket = c->tickets; sa.ctr = c->ctr;
        const u32 nbs = 1u << c->table.group_shift;
        for (u32 g0 = 0; g0 < v.n_parts; g0 += c->split_gpb) {
            // one batch of groups: split its records by slice into the slice-sorted copy, then count slice by slice
            sa.g0 = g0; sa.n_groups = std::min<u32>(c->split_gpb, v.n_parts - g0);
            const u32 slice0 = g0 * nbs, n_sl = std::min<u32>(sa.n_groups * nbs, c->table.n_parts - slice0);
            CK(c, cudaMemsetAsync(c->tickets, 0, 8, si));
            {
                ProfScope ps(c, KMN_PROF_SUBPART, g0 == 0 ? units : 0, si);
                if (c->W == 2) k_slice_split2<false><<<c->n_sms, SPLIT_TPB, c->split_smem, si>>>(sa);
                else if (c->smem_k32) k_slice_split2<true><<<c->n_sms, SPLIT_TPB, c->split_smem, si>>>(sa);
                else if (c->split_tpb == SPLIT_TPB2) k_slice_split<SPLIT_TPB2><<<c->n_sms * c->split_ctas, SPLIT_TPB2, c->split_smem, si>>>(sa);
                else k_slice_split<SPLIT_TPB><<<c->n_sms * c->split_ctas, SPLIT_TPB, c->split_smem, si>>>(sa);
            }
            {
                ProfScope ps(c, KMN_PROF_INSERT, g0 == 0 ? units : 0, si);
                const size_t sm_w = c->table.part_slots * 16 + COUNTW_NBUF * COUNTW_CHUNK * 8;
                if (c->smem_k32)
                    k_count_slices_k32<<<c->n_sms * 2, COUNTW_TPB, c->table.part_slots * 16 + COUNT32_NBUF * COUNT2_CHUNK * 16, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->W == 2)
                    k_count_slices_w2<<<c->n_sms * 2, COUNTW_TPB, ((c->table.part_slots * 24 + 127) & ~(size_t)127) + COUNT2_NBUF * COUNT2_CHUNK * 16, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->count_tma && c->count_db)
                    k_count_slices_db<<<c->n_sms, COUNTD_TPB, c->table.part_slots * 32 + (size_t)COUNTD_NBUF * COUNTD_CHUNK * 8, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->count_tma && c->count_ws == 3)
                    k_count_slices_ws<2><<<c->n_sms * 2, COUNTW_TPB, sm_w, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->count_tma && c->count_ws == 2)
                    k_count_slices_ws<0><<<c->n_sms * 2, COUNTW_TPB, sm_w, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->count_tma && c->count_ws)
                    k_count_slices_ws<1><<<c->n_sms * 2, COUNTW_TPB, sm_w, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else if (c->count_tma)
                    k_count_slices_tma<<<c->n_sms * 2, COUNT3_TPB, c->table.part_slots * 16 + COUNT_NBUF * COUNT_CHUNK * 8, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, c->ctr, slice0, n_sl);
                else
                    k_count_slices<<<c->n_sms * c->count_ctas, COUNT_TPB, c->table.part_slots * 16, si>>>(c->table, c->l2buf, c->cnt2, sa.S, sa.cap2, v.n_parts, c->tickets + 1, c->ctr);
            }
            c->launches += 2;
        }
        c->launches += 2;
        CK(c, cudaGetLastError());
        if (n_entries == n_grouped) return 0;
        ent_ptr += n_grouped; ent_cnt += n_grouped; n_list = n_entries - n_grouped;
        units = 0;
    }
    k_build_worklist<<<1, 1024, 0, si>>>(ent_cnt, n_list, (u32)INSERT_CHUNK, c->chunk_start, c->next_item, c->coarse);
    c->launches++;
    const int grid = c->n_sms * c->insert_ctas;
    const u32 n_split = rb >= 0 ? (u32)c->round_split : 1u;
    for (u32 sp = 0; sp < n_split; ++sp) {
        ProfScope ps(c, KMN_PROF_INSERT, sp == 0 ? units : 0, si);
        KMN_DISPATCH_W(c, KMN_DISPATCH_X(c, {
            k_insert_staged<W_, X_><<<grid, INSERT_TPB, 0, si>>>(c->table, ent_ptr, ent_cnt, n_list, c->chunk_start, c->coarse, c->next_item, c->ctr, sp, n_split);
        }));
        c->launches++;
    }
    CK(c, cudaGetLastError());
    return 0;
}

// submit the drain of one staging set on the insert stream (after everything staged so far on the main stream)
static int submit_drain(kmn_ctx *c, int i)
{
    kmn_ctx::StageSet &st = c->sets[i];
    if (st.staged_upper == 0 || c->p2p) return 0;
    cudaStream_t si = c->s_insert;
    if (si != c->stream) {
        CK(c, cudaEventRecord(st.ev_parsed, c->stream));
        CK(c, cudaStreamWaitEvent(si, st.ev_parsed, 0));
    }
    const u32 n_entries = st.v.n_parts * st.v.n_cta;
    { int r = launch_insert(c, st.v, -1, st.staged_upper, si); if (r) return r; }
    CK(c, cudaMemsetAsync(st.v.count, 0, (size_t)n_entries * 4, si));
    if (si != c->stream) {
        CK(c, cudaEventRecord(st.ev_drained, si));
        st.drain_pending = true;
    }
    st.staged_upper = 0;
    return 0;
}

// the main stream waits for every drain submitted so far
static int wait_drains(kmn_ctx *c)
{
    for (auto &st : c->sets)
        if (st.drain_pending) { CK(c, cudaStreamWaitEvent(c->stream, st.ev_drained, 0)); st.drain_pending = false; }
    return 0;
}

// everything staged so far is in the table as far as later work on the main stream is concerned
static int drain(kmn_ctx *c)
{
    int r = submit_drain(c, c->cur); if (r) return r;
    return wait_drains(c);
}

// make room for n more records in the current set: when it would overflow, hand it to phase 2 and switch to the
// other set (waiting, on the main stream, for that set's previous drain)
static int stage_room(kmn_ctx *c, uint64_t n)
{
    if (c->sets[c->cur].staged_upper + n > c->stage_keys && c->sets[c->cur].staged_upper > 0) {
        int r = submit_drain(c, c->cur); if (r) return r;
        if (c->n_sets > 1) c->cur ^= 1;
    }
    kmn_ctx::StageSet &st = c->sets[c->cur];
    if (st.drain_pending) { CK(c, cudaStreamWaitEvent(c->stream, st.ev_drained, 0)); st.drain_pending = false; }
    return 0;
}

static int count_positions(kmn_ctx *c, const u64 *off, const uint8_t *disc, uint64_t n_reads, uint64_t *out)
{
    // on the copy stream: sizing the next sub-batch must not wait behind the kernels queued on the main stream
    CK(c, cudaMemsetAsync(c->scratch + 4, 0, 8, c->s_copy));
    k_count_positions<<<c->n_sms * 4, 256, 0, c->s_copy>>>(off, disc, n_reads, c->o.kmer_size, c->scratch + 4);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, c->scratch + 4, 8, cudaMemcpyDeviceToHost, c->s_copy));
    CK(c, cudaStreamSynchronize(c->s_copy));
    return 0;
}

static void fill_parse_args(kmn_ctx *c, ParseArgs &a, const uint8_t *bases, const uint8_t *quals, const u64 *off,
                            const uint8_t *disc, uint64_t n_reads, uint64_t total_bytes)
{
    memset(&a, 0, sizeof a);
    a.bases = bases; a.quals = quals; a.read_off = off; a.discarded = disc;
    a.n_reads = n_reads; a.total_bytes = total_bytes; a.ptab = c->ptab;
    a.k = c->o.kmer_size; a.kb = (u32)c->kb; a.pad = c->pad;
    a.min_weight = c->o.min_kmer_quality; a.start_char = c->o.fastq_start_char;
    a.zero_below = c->zero_below; a.mask = (u32 *)c->mask.p; a.wts = c->weights ? (float *)c->wts.p : nullptr;
    a.nranks = (u32)c->nranks; a.rank = (u32)c->rank;
    a.owner_magic = c->nranks > 1 ? 0xFFFFFFFFu / (u32)c->nranks + 1u : 0u;
    a.use_lookup8 = c->o.hash_kind == KMN_HASH_LOOKUP8_HASH2;
    a.l2_hints = getenv("KMN_NO_L2_HINTS") ? 0 : 1;
    a.fast_bound = getenv("KMN_NO_WEIGHT_BOUND") ? 0 : 1;
    a.scatter_steps = getenv("KMN_SCATTER_STEPS") ? (u32)std::max(1, atoi(getenv("KMN_SCATTER_STEPS"))) : 2u;
    a.table = c->table; a.stage = c->sets[c->cur].v; a.ctr = c->ctr;
    a.send_recs = c->send_recs; a.send_cursor = c->send_cursor; a.send_cap = c->send_cap;
    a.flags = c->flags;
    // pieces of 32 reads for large launches; smaller pieces when that would leave CTAs without work (at least 2 per CTA)
    uint32_t ps = 5;
    while (ps > 0 && (n_reads >> ps) < 2ull * (uint64_t)std::max(1, c->n_cta)) --ps;
    a.piece_shift = ps;
    a.ring_R = c->ring_R;
    a.cta_rot = c->cta_rot;
    c->cta_rot = (uint32_t)((c->cta_rot + ((n_reads + (1ull << ps) - 1) >> ps)) % (uint64_t)std::max(1, c->n_cta));
}

static int launch_parse(kmn_ctx *c, const ParseArgs &a)
{
    const bool dist = c->nranks > 1;
    {   // phase 1a: weights -> "counted" bits
        ProfScope ps(c, KMN_PROF_WEIGHT, a.n_reads);
        const int grid = (int)std::min<uint64_t>((a.n_reads + MASK_TPB - 1) / MASK_TPB, (uint64_t)c->n_sms * 8);
        const size_t sm = (size_t)(MASK_TPB / 32) * MASK_WBUF;
        if (c->weights) k_weight_mask<true><<<grid, MASK_TPB, sm, c->stream>>>(a);
        else k_weight_mask<false><<<grid, MASK_TPB, sm, c->stream>>>(a);
    }
    c->launches++;
    CK(c, cudaGetLastError());
    {   // phase 1b: k-mers -> staging sub-regions (and send segments / the other owners' parts of the set)
        ProfScope ps(c, KMN_PROF_PARSE, a.n_reads);
        const int grid = c->n_cta;
        const size_t sm = c->scatter_smem;
#define KMN_SCATTER(X_, E_)                                                                          \
        do {                                                                                         \
            if (!dist) k_kmer_scatter<W_, X_, E_, 0><<<grid, c->scatter_tpb, sm, c->stream>>>(a);     \
            else k_kmer_scatter<W_, X_, E_, 2><<<grid, c->scatter_tpb, sm, c->stream>>>(a);           \
        } while (0)
        KMN_DISPATCH_W(c, {
            if (!c->hasx) KMN_SCATTER(false, false);
            else if (c->ext) KMN_SCATTER(true, true);
            else KMN_SCATTER(true, false);
        });
#undef KMN_SCATTER
    }
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
}

#ifdef KMN_WITH_NCCL
// ---------------------------------------------------------------------------------------------------------
// multi-GPU count pass.  Phase 1 bins every record by (owner rank, table group) into the staging set; one phase-1 launch
// is one ROUND: the parts of the other owners travel into their receive buffers, and phase 2 inserts this rank's own part
// plus what the peers sent.  Transport: the peers' receive buffers mapped with CUDA IPC and written over NVLink by the
// copy engines (default) or by k_push_copy (KMN_PUSH=kernel); when the buffers cannot be mapped (KMN_P2P=0, another
// node) the same parts travel as ncclSend / ncclRecv.  Every rank runs the same sequence of rounds whatever its input
// (a rank without reads takes part with empty parts), so no collective ever depends on a rank's own data.
// Set-up: every rank allocates its receive buffers, exports them, and the handles travel through the communicator.
// ---------------------------------------------------------------------------------------------------------
static int nccl_sum_u64(kmn_ctx *c, u64 mine, u64 *out, cudaStream_t st)
{
    CK(c, cudaMemcpyAsync(c->scratch + 6, &mine, 8, cudaMemcpyHostToDevice, st));
    ncclResult_t nr = ncclAllReduce(c->scratch + 6, c->scratch + 7, 1, ncclUint64, ncclSum, c->comm, st);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllReduce failed: %s", ncclGetErrorString(nr));
    CK(c, cudaMemcpyAsync(out, c->scratch + 7, 8, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    return 0;
}

static int setup_push(kmn_ctx *c)
{
    const int R = c->nranks;
    bool want_ipc = R <= KMN_MAX_PUSH_RANKS;
    if (const char *e = getenv("KMN_P2P")) want_ipc = want_ipc && atoi(e) != 0;
    {   // phase 1 keeps a counter, a flush mark and a ring of record slots per (owner, group) bin in shared memory: with many
        // ranks the groups are made coarser (a group is only a binning granularity; the table's slices stay as they are) until
        // the rings hold at least 8 records -- without rings every record costs its own 32-byte sector write (8 ranks, 355
        // groups: phase 1b 212 ms against 124 ms with 2 ranks).  The slice split of phase 2 takes over the finer half of the
        // binning (at most 2048 slices per group).
        int dev_smem = 0, sm_smem = 0;
        CK(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        CK(c, cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
        const size_t budget = std::min<size_t>((size_t)dev_smem, ((size_t)sm_smem - 1024u * (size_t)c->scatter_ctas) / (size_t)c->scatter_ctas) - 256;
        auto ring_slots = [&](uint64_t n_groups) -> uint32_t {           // as alloc_stage_sets sizes them
            const size_t n_bins = (size_t)n_groups * R, n_pad = (n_bins + 31) & ~(size_t)31, hdr = (2 * n_pad + 64 + 32) * 4;
            uint32_t r = 64;
            while (r >= 4 && hdr + n_bins * r * c->RW * 8 > budget) r >>= 1;
            return r < 4 ? 0 : r;
        };
        uint32_t want_ring = 8;
        if (const char *e = getenv("KMN_MIN_RING")) want_ring = (uint32_t)std::max(0, atoi(e));
        while (c->n_groups > 1 && c->table.group_shift < 11 && ring_slots(c->n_groups) < want_ring) {
            c->table.group_shift++;
            c->n_groups = c->table.n_groups();
        }
        while (c->n_groups > 1 && (c->n_groups * (uint64_t)R + 128) * 8 > budget / 2) {        // (the counters themselves must fit)
            c->table.group_shift++;
            c->n_groups = c->table.n_groups();
        }
    }
    const uint64_t G = c->n_groups;
    c->push_ce = true;
    if (const char *e = getenv("KMN_PUSH")) c->push_ce = strcmp(e, "kernel") != 0;
    if (!want_ipc) c->push_ce = true;                       // the NCCL transport ships whole parts like the copy engines do
    // two staging sets cut by owner; the receive buffer of a source is a verbatim copy of its part of a set
    c->pipeline = true; c->n_sets = 2;
    c->p2p = true;
    { int r = alloc_stage_sets(c); if (r) return r; }
    uint64_t cap; size_t meta_words;
    if (c->push_ce) { cap = (uint64_t)c->n_cta * G * c->sets[0].v.sub_cap + c->sets[0].v.ovf_cap; meta_words = (size_t)G * c->n_cta + 1; }
    else { cap = c->stage_keys / (uint64_t)R; cap += cap / 4 + 65536; meta_words = (size_t)G + 1; }
    cap = (cap + 3) & ~3ull;                                // every source's buffer starts on a 32-byte boundary (bulk copies, sector stores)
    const size_t rec_bytes = 2 * (size_t)R * cap * c->RW * 8, meta_bytes = 2 * (size_t)R * meta_words * 4;
    if (cudaMalloc(&c->recv_all, rec_bytes + meta_bytes + 256) != cudaSuccess) {
        cudaGetLastError(); c->recv_all = nullptr;
        return fail(c, KMN_ERR_NOMEM, "receive buffers of the multi-GPU count pass do not fit (%zu bytes); use a smaller stage_keys", rec_bytes + meta_bytes);
    }
    u64 ok = want_ipc ? 1 : 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, c->recv_all) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    std::vector<cudaIpcMemHandle_t> all((size_t)R);
    {
        void *d = nullptr;
        CK(c, cudaMalloc(&d, (size_t)(R + 1) * 64));
        CK(c, cudaMemcpyAsync((char *)d + (size_t)R * 64, &mine, 64, cudaMemcpyHostToDevice, c->stream));
        ncclResult_t nr = ncclAllGather((char *)d + (size_t)R * 64, d, 64, ncclUint8, c->comm, c->stream);
        if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllGather failed: %s", ncclGetErrorString(nr));
        CK(c, cudaMemcpyAsync(all.data(), d, (size_t)R * 64, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaFree(d));
    }
    {   // the ranks write into each other's buffers: the layout (groups, CTAs, sub-region capacity, record width) must be
        // the same on every rank (same options and the same kind of GPU); anything else is a configuration error
        u64 geo[6] = {G, (u64)c->n_cta, (u64)c->sets[0].v.sub_cap, (u64)c->RW, cap, (u64)meta_words};
        void *d = nullptr;
        CK(c, cudaMalloc(&d, (size_t)(R + 1) * sizeof geo));
        CK(c, cudaMemcpyAsync((char *)d + (size_t)R * sizeof geo, geo, sizeof geo, cudaMemcpyHostToDevice, c->stream));
        ncclResult_t nr = ncclAllGather((char *)d + (size_t)R * sizeof geo, d, 6, ncclUint64, c->comm, c->stream);
        if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllGather failed: %s", ncclGetErrorString(nr));
        std::vector<u64> allgeo((size_t)R * 6);
        CK(c, cudaMemcpyAsync(allgeo.data(), d, (size_t)R * sizeof geo, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaFree(d));
        for (int p = 0; p < R; ++p)
            if (memcmp(&allgeo[(size_t)p * 6], geo, sizeof geo) != 0)
                return fail(c, KMN_ERR_INVALID, "rank %d was created with a different table / staging geometry than rank %d (est_raw_kmers, table_slots, "
                                                "stage_keys, kmer size and value kind must agree on all ranks)", p, c->rank);
    }
    u64 all_ok = 0;
    { int r = nccl_sum_u64(c, ok, &all_ok, c->stream); if (r) return r; }
    if (all_ok == (u64)R) {
        for (int p = 0; p < R && ok; ++p) {
            if (p == c->rank) { c->peer_all[p] = c->recv_all; continue; }
            if (cudaIpcOpenMemHandle(&c->peer_all[p], all[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); c->peer_all[p] = nullptr; ok = 0; }
        }
    } else ok = 0;
    { int r = nccl_sum_u64(c, ok, &all_ok, c->stream); if (r) return r; }
    c->ipc = all_ok == (u64)R;
    if (!c->ipc) {                                 // somebody could not map a peer: everybody uses the NCCL transport
        for (int p = 0; p < KMN_MAX_PUSH_RANKS; ++p) { if (p != c->rank && c->peer_all[p]) cudaIpcCloseMemHandle(c->peer_all[p]); c->peer_all[p] = nullptr; }
        cudaGetLastError();
        if (!c->push_ce) return fail(c, KMN_ERR_INVALID, "KMN_PUSH=kernel needs peer-mapped receive buffers");
    }
    // lookup pass over peer memory: the tables are mapped like the receive buffers (all ranks or none)
    {
        // (opt-in, KMN_PEER_LOOKUP=1: measured on 2 x B200, dependent 32-byte reads of a peer's table run at 0.2 G/s per GPU
        //  against 2.4 G/s for the request / response rounds -- profiles/r02_summary.md)
        u64 tok = 0;
        if (const char *e = getenv("KMN_PEER_LOOKUP")) tok = c->ipc && atoi(e) != 0;
        cudaIpcMemHandle_t th;
        memset(&th, 0, sizeof th);
        if (tok && cudaIpcGetMemHandle(&th, c->table.slots) != cudaSuccess) { cudaGetLastError(); tok = 0; }
        std::vector<cudaIpcMemHandle_t> tall((size_t)R);
        void *d = nullptr;
        CK(c, cudaMalloc(&d, (size_t)(R + 1) * 64));
        CK(c, cudaMemcpyAsync((char *)d + (size_t)R * 64, &th, 64, cudaMemcpyHostToDevice, c->stream));
        ncclResult_t nr = ncclAllGather((char *)d + (size_t)R * 64, d, 64, ncclUint8, c->comm, c->stream);
        if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllGather failed: %s", ncclGetErrorString(nr));
        CK(c, cudaMemcpyAsync(tall.data(), d, (size_t)R * 64, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        CK(c, cudaFree(d));
        u64 all_tok = 0;
        { int r = nccl_sum_u64(c, tok, &all_tok, c->stream); if (r) return r; }
        if (all_tok == (u64)R) {
            for (int p = 0; p < R && tok; ++p) {
                if (p == c->rank) { c->peer_table[p] = c->table.slots; continue; }
                if (cudaIpcOpenMemHandle(&c->peer_table[p], tall[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); c->peer_table[p] = nullptr; tok = 0; }
            }
        } else tok = 0;
        { int r = nccl_sum_u64(c, tok, &all_tok, c->stream); if (r) return r; }
        c->peer_lookup = all_tok == (u64)R;
        if (c->peer_lookup && !c->bar_buf) { CK(c, cudaMalloc((void **)&c->bar_buf, 16)); CK(c, cudaMemsetAsync(c->bar_buf, 0, 16, c->stream)); }
        if (!c->peer_lookup) {
            for (int p = 0; p < R; ++p) { if (p != c->rank && c->peer_table[p]) cudaIpcCloseMemHandle(c->peer_table[p]); c->peer_table[p] = nullptr; }
            cudaGetLastError();
        }
    }
    c->push_cap = cap;
    c->push_meta = meta_words;
    c->s_insert = c->stream;                       // phase 1 and phase 2 alternate on one stream; only the transfers overlap them
    {   // the barrier kernels of a round are tiny but sit behind long persistent kernels: give them the first free SM
        int lo_pri = 0, hi_pri = 0;
        CK(c, cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CK(c, cudaStreamCreateWithPriority(&c->s_comm, cudaStreamNonBlocking, hi_pri));
    }
    if (getenv("KMN_TWO_COPY_STREAMS") && c->ipc) {
        CK(c, cudaStreamCreateWithFlags(&c->s_comm2, cudaStreamNonBlocking));
        CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) {
        CK(c, cudaEventCreateWithFlags(&c->ev_pushed[i], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_rb_free[i], cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_recv[i], cudaEventDisableTiming));
    }
    const size_t n_sub = (size_t)G * c->n_cta;
    CK(c, cudaMalloc((void **)&c->run_off, (size_t)R * n_sub * 4));
    CK(c, cudaMalloc((void **)&c->grp_off, (size_t)R * (G + 1) * 4));
    CK(c, cudaMalloc((void **)&c->flags, 16));
    CK(c, cudaMalloc((void **)&c->d_const, 32));
    const u64 consts[4] = {0, 1, 0, 0};
    CK(c, cudaMemcpyAsync(c->d_const, consts, 32, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->flags, 0, 16, c->stream));
    CK(c, cudaMemsetAsync(c->recv_all, 0, rec_bytes + meta_bytes, c->stream));
    { int r = apply_smem_attrs(c); if (r) return r; }
    CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// Transfers of one round: the set `si` (filled by phase 1 on the main stream, possibly empty) is pushed to its owners.
// `done` says this rank has no more input; the sum of the flags comes back when `done_sum` is given.
//   comm stream: plan -> barrier A (the peers' round buffer is free; carries the done flags) -> parts over NVLink / NCCL -> barrier B
static int push_comm(kmn_ctx *c, int si, bool done, u64 *done_sum)
{
    kmn_ctx::StageSet &st = c->sets[si];
    const int R = c->nranks, rb = (int)(c->round & 1);
    const size_t G = (size_t)c->n_groups;
    cudaStream_t sc = c->s_comm;
    CK(c, cudaEventRecord(st.ev_parsed, c->stream));
    CK(c, cudaStreamWaitEvent(sc, st.ev_parsed, 0));
    if (c->rb_busy[rb]) { CK(c, cudaStreamWaitEvent(sc, c->ev_rb_free[rb], 0)); c->rb_busy[rb] = false; }
    if (!c->push_ce) { k_push_plan<<<R, 1024, 0, sc>>>(st.v, c->push_cap, c->run_off, c->grp_off, c->flags); c->launches++; }
    ncclResult_t nr = ncclAllReduce(c->d_const + (done ? 1 : 0), c->d_const + 3, 1, ncclUint64, ncclSum, c->comm, sc);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "push barrier failed: %s", ncclGetErrorString(nr));
    const size_t n_sub = G * c->n_cta, part = n_sub * st.v.sub_cap * c->RW;          // u64 words per owner part
    if (c->push_ce && c->ipc) {
        // copy engines: the part of every other owner (whole sub-region capacity) and its fill counters go to that
        // owner's round buffer as they are; no SM takes part in the transfer
        ProfScope ps(c, KMN_PROF_ROUTE, (uint64_t)(R - 1) * ((done ? 0 : part * 8) + n_sub * 4), sc);   // units = bytes leaving this GPU
        const bool two = c->s_comm2 != nullptr && R > 2;
        if (two) { CK(c, cudaEventRecord(c->ev_fork, sc)); CK(c, cudaStreamWaitEvent(c->s_comm2, c->ev_fork, 0)); }
        for (int q = 1; q < R; ++q) {
            const int p = (c->rank + q) % R;                                             // start with a different peer on every rank
            cudaStream_t scp = (two && (q & 1) == 0) ? c->s_comm2 : sc;
            u64 *base = (u64 *)c->peer_all[p];
            u64 *drec = base + ((size_t)rb * R + c->rank) * c->push_cap * c->RW;
            u32 *dmeta = (u32 *)(base + 2 * (size_t)R * c->push_cap * c->RW) + ((size_t)rb * R + c->rank) * c->push_meta;
            // a finishing round (this rank has no more input) has nothing staged: only the (zero) counters travel
            if (!done) {
                CK(c, cudaMemcpyAsync(drec, st.v.recs + (size_t)p * part, part * 8, cudaMemcpyDeviceToDevice, scp));
                if (st.v.ovf_cap) CK(c, cudaMemcpyAsync(drec + part, st.v.ovf_recs + (size_t)p * st.v.ovf_cap * c->RW, (size_t)st.v.ovf_cap * c->RW * 8, cudaMemcpyDeviceToDevice, scp));
            }
            CK(c, cudaMemcpyAsync(dmeta, st.v.count + (size_t)p * n_sub, n_sub * 4, cudaMemcpyDeviceToDevice, scp));
            if (st.v.ovf_cap) CK(c, cudaMemcpyAsync(dmeta + n_sub, st.v.ovf_count + p, 4, cudaMemcpyDeviceToDevice, scp));
        }
        if (two) { CK(c, cudaEventRecord(c->ev_join, c->s_comm2)); CK(c, cudaStreamWaitEvent(sc, c->ev_join, 0)); }
    } else if (c->push_ce) {
        // NCCL transport: the same parts and counters as send / receive pairs (MPIAllToAllMessageBuffer::sendReceive,
        // src/MPIBuffer.h:588-600).  The receiver cannot know whether a sender is in a finishing round, so the parts always travel.
        ProfScope ps(c, KMN_PROF_ROUTE, (uint64_t)(R - 1) * (part * 8 + n_sub * 4), sc);
        u64 *mybase = (u64 *)c->recv_all;
        ncclGroupStart();
        for (int q = 1; q < R; ++q) {
            const int p = (c->rank + q) % R;
            u64 *rrec = mybase + ((size_t)rb * R + p) * c->push_cap * c->RW;
            u32 *rmeta = (u32 *)(mybase + 2 * (size_t)R * c->push_cap * c->RW) + ((size_t)rb * R + p) * c->push_meta;
            ncclSend(st.v.recs + (size_t)p * part, part, ncclUint64, p, c->comm, sc);
            ncclRecv(rrec, part, ncclUint64, p, c->comm, sc);
            if (st.v.ovf_cap) {
                ncclSend(st.v.ovf_recs + (size_t)p * st.v.ovf_cap * c->RW, (size_t)st.v.ovf_cap * c->RW, ncclUint64, p, c->comm, sc);
                ncclRecv(rrec + part, (size_t)st.v.ovf_cap * c->RW, ncclUint64, p, c->comm, sc);
                ncclSend(st.v.ovf_count + p, 1, ncclUint32, p, c->comm, sc);
                ncclRecv(rmeta + n_sub, 1, ncclUint32, p, c->comm, sc);
            }
            ncclSend(st.v.count + (size_t)p * n_sub, n_sub, ncclUint32, p, c->comm, sc);
            ncclRecv(rmeta, n_sub, ncclUint32, p, c->comm, sc);
        }
        nr = ncclGroupEnd();
        if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "nccl exchange of the round failed: %s", ncclGetErrorString(nr));
    } else {
        PushPeers pp;
        memset(&pp, 0, sizeof pp);
        for (int p = 0; p < R; ++p) {
            u64 *base = (u64 *)c->peer_all[p];
            pp.recs[p] = base + ((size_t)rb * R + c->rank) * c->push_cap * c->RW;
            pp.meta[p] = (u32 *)(base + 2 * (size_t)R * c->push_cap * c->RW) + ((size_t)rb * R + c->rank) * c->push_meta;
        }
        ProfScope ps(c, KMN_PROF_ROUTE, 0, sc);
        k_push_copy<<<c->n_sms * 2, 128, 0, sc>>>(st.v, pp, c->push_cap, c->run_off, c->grp_off, (u32)c->RW);
        c->launches++;
    }
    CK(c, cudaGetLastError());
    {   // the other owners' fill counters of this set are consumed
        if (c->rank > 0) CK(c, cudaMemsetAsync(st.v.count, 0, (size_t)c->rank * n_sub * 4, sc));
        if (c->rank + 1 < R) CK(c, cudaMemsetAsync(st.v.count + (size_t)(c->rank + 1) * n_sub, 0, (size_t)(R - 1 - c->rank) * n_sub * 4, sc));
        if (st.v.ovf_count) CK(c, cudaMemsetAsync(st.v.ovf_count, 0, (size_t)R * 4, sc));
    }
    CK(c, cudaEventRecord(c->ev_pushed[si], sc));
    c->push_pending[si] = true;
    nr = ncclAllReduce(c->d_const, c->d_const + 2, 1, ncclUint64, ncclSum, c->comm, sc);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "push barrier failed: %s", ncclGetErrorString(nr));
    CK(c, cudaEventRecord(c->ev_recv[rb], sc));
    c->round++;
    if (done_sum) {
        CK(c, cudaMemcpyAsync(done_sum, c->d_const + 3, 8, cudaMemcpyDeviceToHost, sc));
        CK(c, cudaStreamSynchronize(sc));
    }
    return 0;
}

// Phase 2 of a round on the main stream, once the peers' parts have arrived: this rank's own part of set `si` plus round
// buffer `rb`.  It is submitted AFTER the next round's phase 1, so the main stream runs P1(r+1), I(r), P1(r+2), I(r+1), ...
// while the transfers of round r+1 overlap I(r): phase 1 and phase 2 never run at the same time (phase 1 may insert
// directly into the table when a sub-region overflows, and phase 2 holds table slices in shared memory).
static int push_insert(kmn_ctx *c, int si, int rb)
{
    kmn_ctx::StageSet &st = c->sets[si];
    const size_t G = (size_t)c->n_groups;
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_recv[rb], 0));
    { int r = launch_insert(c, st.v, rb, c->stage_keys, c->stream); if (r) return r; }
    CK(c, cudaMemsetAsync(st.v.count + (size_t)c->rank * G * c->n_cta, 0, G * c->n_cta * 4, c->stream));
    CK(c, cudaEventRecord(c->ev_rb_free[rb], c->stream));
    c->rb_busy[rb] = true;
    return 0;
}

static int flush_pending_insert(kmn_ctx *c)
{
    if (c->pending_set < 0) return 0;
    const int si = c->pending_set, rb = c->pending_rb;
    c->pending_set = c->pending_rb = -1;
    return push_insert(c, si, rb);
}

// one round: phase 1 has been submitted into set `si`; ship it, then run the previous round's phase 2
static int push_round(kmn_ctx *c, int si, bool done, u64 *done_sum)
{
    const int rb = (int)(c->round & 1);
    int r = push_comm(c, si, done, done_sum); if (r) return r;
    r = flush_pending_insert(c); if (r) return r;
    c->pending_set = si; c->pending_rb = rb;
    return 0;
}

// phase 1 may write into set `si` again: its previous push is complete (its previous insert precedes on the main stream)
static int push_set_ready(kmn_ctx *c, int si)
{
    if (c->push_pending[si]) { CK(c, cudaStreamWaitEvent(c->stream, c->ev_pushed[si], 0)); c->push_pending[si] = false; }
    return 0;
}
#endif

int kmn_comm_unique_id(void *id128)
{
#ifdef KMN_WITH_NCCL
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (ncclGetUniqueId(&id) != ncclSuccess) return KMN_ERR_COMM;
    memcpy(id128, &id, 128);
    return 0;
#else
    (void)id128;
    return KMN_ERR_COMM;
#endif
}

int kmn_comm_init(kmn_ctx *c, int rank, int nranks, const void *id128)
{
    if (!c) return KMN_ERR_INVALID;
#ifdef KMN_WITH_NCCL
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return fail(c, KMN_ERR_INVALID, "bad rank %d / %d (at most 64 ranks)", rank, nranks);
    if (c->sets[0].staged_upper || c->sets[1].staged_upper) return fail(c, KMN_ERR_STATE, "kmn_comm_init after counting started");
    c->rank = rank; c->nranks = nranks;
    if (nranks == 1) return 0;
    CK(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t nr = ncclCommInitRank(&c->comm, nranks, id, rank);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclCommInitRank failed: %s", ncclGetErrorString(nr));
    CK(c, cudaMalloc((void **)&c->send_cursor, (size_t)nranks * 8));
    CK(c, cudaMalloc((void **)&c->all_counts, (size_t)nranks * nranks * 8));
    CK(c, cudaMemsetAsync(c->send_cursor, 0, (size_t)nranks * 8, c->stream));
    { int r = setup_push(c); if (r) return r; }
    // request / answer regions of the lookup pass: it runs after the count pass, so they reuse the two round buffers
    c->send_cap = c->stage_keys / 2 / (uint64_t)nranks * 3 / 2 + 65536;
    const size_t half = (size_t)nranks * c->push_cap * c->RW;
    if ((size_t)nranks * c->send_cap * c->RW > half) c->send_cap = half / ((size_t)nranks * c->RW);
    c->recv_cap = c->send_cap * (uint64_t)(nranks - 1);
    c->send_recs = (u64 *)c->recv_all;
    c->recv_recs = (u64 *)c->recv_all + half;
    return 0;
#else
    (void)rank; (void)nranks; (void)id128;
    return fail(c, KMN_ERR_COMM, "library built without NCCL");
#endif
}

// ---------------------------------------------------------------------------------------------------------
// count pass
// ---------------------------------------------------------------------------------------------------------
struct BatchPtrs {
    const uint8_t *bases, *quals, *disc;
    const u64 *off;
    uint64_t total_bytes;
    bool off_on_host;
    int slot = -1;                 // input staging slot used for host inputs (-1: everything was already on the device)
};

// Host inputs are copied into one of two device staging slots on the copy stream, so the copy of this batch overlaps
// the kernels of the previous one; the main stream waits for the copy, the copy waits until the slot's previous
// batch has been consumed.
static int stage_inputs(kmn_ctx *c, const uint8_t *bases, const uint8_t *quals, const uint64_t *read_off, uint64_t n_reads,
                        const uint8_t *discarded, bool need_quals, BatchPtrs &bp)
{
    bp.off_on_host = !is_device_ptr(read_off);
    u64 first = 0, last = 0;
    if (bp.off_on_host) { first = read_off[0]; last = read_off[n_reads]; }
    else {
        CK(c, cudaMemcpyAsync(&first, read_off, 8, cudaMemcpyDeviceToHost, c->s_copy));
        CK(c, cudaMemcpyAsync(&last, read_off + n_reads, 8, cudaMemcpyDeviceToHost, c->s_copy));
        CK(c, cudaStreamSynchronize(c->s_copy));
    }
    if (first != 0) return fail(c, KMN_ERR_INVALID, "read_off[0] must be 0");
    if (need_quals && !quals) return fail(c, KMN_ERR_INVALID, "quals is required");
    bp.total_bytes = last;
    const bool h_bases = !is_device_ptr(bases), h_quals = need_quals && !is_device_ptr(quals), h_disc = discarded && !is_device_ptr(discarded);
    bp.off = reinterpret_cast<const u64 *>(read_off); bp.bases = bases; bp.quals = need_quals ? quals : nullptr; bp.disc = discarded;
    if (!(bp.off_on_host || h_bases || h_quals || h_disc)) return 0;
    const int j = c->in_cur;
    c->in_cur = (c->in_cur + 1) % kmn_ctx::IN_SLOTS;
    bp.slot = j;
    if (c->in_used[j]) { CK(c, cudaStreamWaitEvent(c->s_copy, c->ev_in_free[j], 0)); CK(c, cudaStreamWaitEvent(c->s_copy2, c->ev_in_free[j], 0)); }
    cudaStream_t sc = c->s_copy;
    if (bp.off_on_host) {
        int r = ensure(c, c->in_off[j], (n_reads + 1) * 8); if (r) return r;
        CK(c, cudaMemcpyAsync(c->in_off[j].p, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, sc));
        bp.off = (const u64 *)c->in_off[j].p;
    }
    if (h_bases) {
        int r = ensure(c, c->in_bases[j], last + 16); if (r) return r;
        CK(c, cudaMemcpyAsync(c->in_bases[j].p, bases, last, cudaMemcpyHostToDevice, sc));
        bp.bases = (const uint8_t *)c->in_bases[j].p;
    }
    if (h_quals) {
        int r = ensure(c, c->in_quals[j], last + 16); if (r) return r;
        CK(c, cudaMemcpyAsync(c->in_quals[j].p, quals, last, cudaMemcpyHostToDevice, c->s_copy2));
        bp.quals = (const uint8_t *)c->in_quals[j].p;
    }
    if (h_disc) {
        int r = ensure(c, c->in_disc[j], n_reads + 16); if (r) return r;
        CK(c, cudaMemcpyAsync(c->in_disc[j].p, discarded, n_reads, cudaMemcpyHostToDevice, sc));
        bp.disc = (const uint8_t *)c->in_disc[j].p;
    }
    CK(c, cudaEventRecord(c->ev_in_ready[j], sc));
    CK(c, cudaEventRecord(c->ev_in_ready2[j], c->s_copy2));
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_in_ready[j], 0));
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_in_ready2[j], 0));
    return 0;
}

// the kernels reading this batch's staging slot have been launched on the main stream: mark the slot reusable after
// them, and return to the caller only once its host buffers have been read (it may overwrite them right away)
static int release_inputs(kmn_ctx *c, const BatchPtrs &bp)
{
    if (bp.slot < 0) return 0;
    CK(c, cudaEventRecord(c->ev_in_free[bp.slot], c->stream));
    c->in_used[bp.slot] = true;
    CK(c, cudaEventSynchronize(c->ev_in_ready[bp.slot]));
    CK(c, cudaEventSynchronize(c->ev_in_ready2[bp.slot]));
    return 0;
}

static int count_staged(kmn_ctx *c, BatchPtrs &bp, const uint64_t *read_off, uint64_t n_reads, const uint8_t *discarded);

int kmn_count_batch(kmn_ctx *c, const uint8_t *bases, const uint8_t *quals, const uint64_t *read_off, uint64_t n_reads,
                    const uint8_t *discarded)
{
    if (!c) return KMN_ERR_INVALID;
    if (n_reads && (!bases || !read_off)) return fail(c, KMN_ERR_INVALID, "null input");
    CK(c, cudaSetDevice(c->device));
    c->finished = false;
    if (n_reads == 0) return 0;          // a rank without reads still takes part in every round: kmn_count_finish runs empty ones
    BatchPtrs bp;
    int r = stage_inputs(c, bases, quals, read_off, n_reads, discarded, true, bp);
    if (r) return r;
    return count_staged(c, bp, read_off, n_reads, discarded);
}

int kmn_count_batch_2na(kmn_ctx *c, const uint8_t *packed, const uint64_t *packed_off, const uint8_t *quals, const uint64_t *read_off,
                        uint64_t n_reads, const uint8_t *discarded, const uint64_t *markup_pos, const uint8_t *markup_chr, uint64_t n_markups)
{
    if (!c) return KMN_ERR_INVALID;
    if (n_reads && (!packed || !packed_off || !read_off || !quals)) return fail(c, KMN_ERR_INVALID, "null input");
    if (n_markups && (!markup_pos || !markup_chr)) return fail(c, KMN_ERR_INVALID, "null markup arrays");
    if (is_device_ptr(read_off) || is_device_ptr(packed_off)) return fail(c, KMN_ERR_INVALID, "kmn_count_batch_2na takes host offset arrays");
    CK(c, cudaSetDevice(c->device));
    c->finished = false;
    if (n_reads == 0) return 0;
    if (read_off[0] != 0 || packed_off[0] != 0) return fail(c, KMN_ERR_INVALID, "read_off[0] and packed_off[0] must be 0");
    const u64 total = read_off[n_reads], ptotal = packed_off[n_reads];
    const int j = c->in_cur;
    c->in_cur = (c->in_cur + 1) % kmn_ctx::IN_SLOTS;
    if (c->in_used[j]) { CK(c, cudaStreamWaitEvent(c->s_copy, c->ev_in_free[j], 0)); CK(c, cudaStreamWaitEvent(c->s_copy2, c->ev_in_free[j], 0)); }
    int r;
    if ((r = ensure(c, c->in_off[j], (n_reads + 1) * 8)) || (r = ensure(c, c->in_poff[j], (n_reads + 1) * 8)) || (r = ensure(c, c->in_packed[j], ptotal + 16)) ||
        (r = ensure(c, c->in_bases[j], total + 16)) || (r = ensure(c, c->in_quals[j], total + 16))) return r;
    cudaStream_t sc = c->s_copy;
    CK(c, cudaMemcpyAsync(c->in_off[j].p, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, sc));
    CK(c, cudaMemcpyAsync(c->in_poff[j].p, packed_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, sc));
    CK(c, cudaMemcpyAsync(c->in_packed[j].p, packed, ptotal, cudaMemcpyDefault, sc));
    CK(c, cudaMemcpyAsync(c->in_quals[j].p, quals, total, cudaMemcpyDefault, c->s_copy2));
    if (discarded) {
        if ((r = ensure(c, c->in_disc[j], n_reads + 16))) return r;
        CK(c, cudaMemcpyAsync(c->in_disc[j].p, discarded, n_reads, cudaMemcpyDefault, sc));
    }
    if (n_markups) {
        if ((r = ensure(c, c->in_mpos[j], n_markups * 8)) || (r = ensure(c, c->in_mchr[j], n_markups + 16))) return r;
        CK(c, cudaMemcpyAsync(c->in_mpos[j].p, markup_pos, n_markups * 8, cudaMemcpyDefault, sc));
        CK(c, cudaMemcpyAsync(c->in_mchr[j].p, markup_chr, n_markups, cudaMemcpyDefault, sc));
    }
    CK(c, cudaEventRecord(c->ev_in_ready[j], sc));
    CK(c, cudaEventRecord(c->ev_in_ready2[j], c->s_copy2));
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_in_ready[j], 0));
    CK(c, cudaStreamWaitEvent(c->stream, c->ev_in_ready2[j], 0));
    // packed bases -> ASCII in the slot's base buffer, markups on top
    k_unpack_2na<<<c->n_sms * 8, 256, 0, c->stream>>>((const uint8_t *)c->in_packed[j].p, (const u64 *)c->in_poff[j].p, (const u64 *)c->in_off[j].p, n_reads, (uint8_t *)c->in_bases[j].p);
    c->launches++;
    if (n_markups) {
        k_apply_markups<<<(unsigned)std::min<u64>((n_markups + 255) / 256, (u64)c->n_sms * 4), 256, 0, c->stream>>>((const u64 *)c->in_mpos[j].p, (const uint8_t *)c->in_mchr[j].p, n_markups, total, (uint8_t *)c->in_bases[j].p);
        c->launches++;
    }
    CK(c, cudaGetLastError());
    BatchPtrs bp;
    bp.bases = (const uint8_t *)c->in_bases[j].p; bp.quals = (const uint8_t *)c->in_quals[j].p; bp.off = (const u64 *)c->in_off[j].p;
    bp.disc = discarded ? (const uint8_t *)c->in_disc[j].p : nullptr;
    bp.total_bytes = total; bp.off_on_host = true; bp.slot = j;
    return count_staged(c, bp, read_off, n_reads, is_device_ptr(discarded) ? nullptr : discarded);
}

// phase 1 of a staged batch (inputs on the device; read_off / discarded: the host copies when there are any)
static int count_staged(kmn_ctx *c, BatchPtrs &bp, const uint64_t *read_off, uint64_t n_reads, const uint8_t *discarded)
{
    int r;
    // phase 1a -> 1b scratch: one bit per base position of the batch (+ one fp32 per position for KMN_VALUE_WEIGHTS)
    const size_t mask_bytes = (bp.total_bytes / 32 + 4) * 4;
    r = ensure(c, c->mask, mask_bytes); if (r) return r;
    CK(c, cudaMemsetAsync(c->mask.p, 0, mask_bytes, c->stream));
    if (c->weights) { r = ensure(c, c->wts, (bp.total_bytes + 4) * 4); if (r) return r; }
    // A launch may stage at most `limit` instances (one staging set; multi-GPU: half of it, the other half takes the
    // records received from the peers and the same bound sizes the send regions).  The exact number of k-mer
    // positions of a read range is computed on the device; ranges that do not fit are halved.
    const uint64_t limit = std::max<uint64_t>(c->stage_keys, 1);
    const bool host_sizes = bp.off_on_host && (!discarded || !is_device_ptr(discarded));
    if (!host_sizes && bp.slot >= 0) { CK(c, cudaStreamSynchronize(c->s_copy)); CK(c, cudaStreamSynchronize(c->s_copy2)); }   // count_positions reads the staged copies
    struct Range { uint64_t r0, r1; };
    std::vector<Range> todo;
    todo.push_back({0, n_reads});
    while (!todo.empty()) {
        Range rg = todo.back();
        todo.pop_back();
        uint64_t npos = 0;
        if (host_sizes) {                                                      // host offsets: no device round trip
            const uint32_t k = c->o.kmer_size;
            for (uint64_t q = rg.r0; q < rg.r1; ++q) {
                uint64_t len = read_off[q + 1] - read_off[q];
                if (len >= k && !(discarded && discarded[q])) npos += len - k + 1;
            }
        } else {
            r = count_positions(c, bp.off + rg.r0, bp.disc ? bp.disc + rg.r0 : nullptr, rg.r1 - rg.r0, &npos);
            if (r) return r;
        }
        if (npos > limit && rg.r1 - rg.r0 > 1) {
            // ceil(npos / limit) pieces of equal read count (every piece is measured again: read lengths may be uneven)
            const uint64_t nr = rg.r1 - rg.r0, parts = std::min<uint64_t>(nr, (npos + limit - 1) / limit);
            for (uint64_t q = parts; q-- > 0;)      // stack order: the first piece is processed first
                todo.push_back({rg.r0 + nr * q / parts, rg.r0 + nr * (q + 1) / parts});
            continue;
        }
        if (npos == 0) continue;
#ifdef KMN_WITH_NCCL
        if (c->p2p) {                                       // one launch = one round: fill a set, push it, insert
            r = push_set_ready(c, c->cur); if (r) return r;
            const uint64_t nr = rg.r1 - rg.r0, ns = (uint64_t)std::max(1, c->round_split);
            for (uint64_t sp = 0; sp < ns; ++sp) {            // same set, several launches (see round_split)
                const uint64_t q0 = rg.r0 + nr * sp / ns, q1 = rg.r0 + nr * (sp + 1) / ns;
                if (q1 == q0) continue;
                ParseArgs a;
                fill_parse_args(c, a, bp.bases, bp.quals, bp.off + q0, bp.disc ? bp.disc + q0 : nullptr, q1 - q0, bp.total_bytes);
                r = launch_parse(c, a); if (r) return r;
            }
            r = push_round(c, c->cur, false, nullptr); if (r) return r;
            c->cur ^= 1;
            continue;
        }
#endif
        r = stage_room(c, npos); if (r) return r;
        ParseArgs a;
        fill_parse_args(c, a, bp.bases, bp.quals, bp.off + rg.r0, bp.disc ? bp.disc + rg.r0 : nullptr, rg.r1 - rg.r0, bp.total_bytes);
        r = launch_parse(c, a); if (r) return r;
        c->sets[c->cur].staged_upper += npos;
    }
    return release_inputs(c, bp);
}

static int launch_purge(kmn_ctx *c, uint32_t min_depth)
{
    CK(c, cudaMemsetAsync(c->scratch, 0, 8, c->stream));
    {
        ProfScope ps(c, KMN_PROF_SCAN, c->n_slots);
        KMN_DISPATCH_W(c, { k_purge<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->n_slots, min_depth, c->scratch); });
    }
    c->launches++;
    CK(c, cudaGetLastError());
    if (min_depth > c->purged_depth) c->purged_depth = min_depth;
    return 0;
}

int kmn_count_finish(kmn_ctx *c, int apply_purge)
{
    if (!c) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    int r = 0;
#ifdef KMN_WITH_NCCL
    if (c->p2p && !c->finished) {
        // empty rounds until every rank has said it is done (a rank that finishes early keeps receiving and inserting)
        while (true) {
            r = push_set_ready(c, c->cur); if (r) return r;
            u64 n_done = 0;
            r = push_round(c, c->cur, true, &n_done); if (r) return r;
            c->cur ^= 1;
            if (n_done == (u64)c->nranks) break;
        }
        r = flush_pending_insert(c); if (r) return r;
        u64 fl[2] = {0, 0};
        CK(c, cudaMemcpyAsync(fl, c->flags, 16, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        if (fl[0] || fl[1]) return fail(c, KMN_ERR_COMM, "push path overflow: %llu records beyond a remote sub-region, %llu beyond the receive buffer "
                                        "(skewed k-mer distribution); use smaller batches or KMN_P2P=0", (unsigned long long)fl[0], (unsigned long long)fl[1]);
    }
#endif
    r = drain(c);
    if (r) return r;
    Counters h;
    CK(c, cudaMemcpyAsync(&h, c->ctr, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (h.table_full) return fail(c, KMN_ERR_TABLE_FULL, "count table overflow: %llu k-mer instances could not be inserted (capacity %llu slots); "
                                  "raise table_slots / est_raw_kmers", (unsigned long long)h.table_full, (unsigned long long)c->n_slots);
    // post-build purge: singletons dropped when minDepth>=2, count<minDepth dropped when minDepth>2
    // (src/KmerSpectrum.h:1825,1805-1815; src/DistributedFunctions.h:559-569)
    if (apply_purge && c->o.min_depth > 1) { r = launch_purge(c, c->o.min_depth); if (r) return r; }
    c->finished = true;
    return 0;
}

int kmn_purge_min_depth(kmn_ctx *c, uint32_t min_depth)
{
    if (!c) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    int r = drain(c); if (r) return r;
    if (min_depth > 1) { r = launch_purge(c, min_depth); if (r) return r; }
    CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int kmn_get_stats(kmn_ctx *c, kmn_stats *out)
{
    if (!c || !out) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    int r = drain(c); if (r) return r;
    CK(c, cudaMemsetAsync(c->scratch, 0, 16, c->stream));
    KMN_DISPATCH_W(c, { k_count_live<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->n_slots, 1, c->scratch, c->scratch + 1); });
    c->launches++;
    Counters h; u64 ls[2];
    CK(c, cudaMemcpyAsync(&h, c->ctr, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(ls, c->scratch, 16, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    out->raw_kmers = h.raw; out->raw_good_kmers = h.raw_good; out->unique_kmers = h.unique; out->singleton_kmers = ls[1];
    out->discarded_kmers = h.raw - h.raw_good; out->table_slots = c->n_slots; out->table_partitions = c->table.n_groups();
    out->direct_inserts = h.direct;
    return 0;
}

int kmn_histogram(kmn_ctx *c, uint64_t *hist, double *wsum)
{
    if (!c || !hist) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->nranks > 1 && c->p2p && !c->finished)          // the round buffers are still in use by the count pass (header: KMN_ERR_STATE)
        return fail(c, KMN_ERR_STATE, "%s before kmn_count_finish on a multi-GPU context", __func__);
    int r = drain(c); if (r) return r;
    DevBuf &hb = c->lk_out;
    r = ensure(c, hb, 65536 * 16); if (r) return r;
    u64 *dh = (u64 *)hb.p; double *dw = (double *)(dh + 65536);
    CK(c, cudaMemsetAsync(dh, 0, 65536 * 16, c->stream));
    KMN_DISPATCH_W(c, { k_histogram<W_><<<c->n_sms * 4, 256, 0, c->stream>>>(c->table, c->n_slots, dh, wsum ? dw : nullptr); });
    c->launches++;
    CK(c, cudaGetLastError());
#ifdef KMN_WITH_NCCL
    if (c->nranks > 1) {   // MPIHistogram::reduce (src/DistributedFunctions.h:495-535)
        ncclGroupStart();
        ncclAllReduce(dh, dh, 65536, ncclUint64, ncclSum, c->comm, c->stream);
        ncclAllReduce(dw, dw, 65536, ncclDouble, ncclSum, c->comm, c->stream);
        if (ncclGroupEnd() != ncclSuccess) return fail(c, KMN_ERR_COMM, "histogram all-reduce failed");
    }
#endif
    CK(c, cudaMemcpyAsync(hist, dh, 65536 * 8, cudaMemcpyDeviceToHost, c->stream));
    if (wsum) CK(c, cudaMemcpyAsync(wsum, dw, 65536 * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

#ifdef KMN_WITH_NCCL
static int allreduce_max_u64(kmn_ctx *c, u64 mine, u64 *out);
static int lookup_exchange(kmn_ctx *c, uint32_t min_depth, uint16_t *vals);
// Barrier in stream order (no host synchronisation): what follows on c->stream starts when every rank's stream has reached
// its own call.  A peer-memory lookup pass sits between two of them: the first says "every table is final" (the count pass
// and purges of all ranks precede it on their streams), the second "nobody reads my table any more" (so a reset, purge or
// the next count pass behind it on the stream is safe).
static int stream_barrier(kmn_ctx *c)
{
    ncclResult_t nr = ncclAllReduce(c->bar_buf, c->bar_buf + 1, 1, ncclUint64, ncclSum, c->comm, c->stream);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllReduce failed: %s", ncclGetErrorString(nr));
    return 0;
}
static PeerTables peer_tables(const kmn_ctx *c)
{
    PeerTables pt;
    memset(&pt, 0, sizeof pt);
    for (int p = 0; p < c->nranks && p < KMN_MAX_PUSH_RANKS; ++p) pt.slots[p] = c->peer_table[p];
    return pt;
}
#endif

int kmn_lookup(kmn_ctx *c, const uint8_t *keys, uint64_t n, uint16_t *counts)
{
    if (!c || (n && (!keys || !counts))) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->nranks > 1 && c->p2p && !c->finished)          // the round buffers are still in use by the count pass (header: KMN_ERR_STATE)
        return fail(c, KMN_ERR_STATE, "%s before kmn_count_finish on a multi-GPU context", __func__);
    int r = drain(c); if (r) return r;
#ifdef KMN_WITH_NCCL
    if (c->nranks > 1 && c->peer_lookup) {
        // collective (n may be 0): the owners' tables are probed over peer memory between two stream-ordered barriers
        r = stream_barrier(c); if (r) return r;
        if (n) {
            const uint8_t *dk = keys;
            if (!is_device_ptr(keys)) {
                r = ensure(c, c->lk_keys, n * c->kb); if (r) return r;
                CK(c, cudaMemcpyAsync(c->lk_keys.p, keys, n * c->kb, cudaMemcpyHostToDevice, c->stream));
                dk = (const uint8_t *)c->lk_keys.p;
            }
            uint16_t *dout = counts;
            const bool out_host = !is_device_ptr(counts);
            if (out_host) { r = ensure(c, c->lk_out, n * 2); if (r) return r; dout = (uint16_t *)c->lk_out.p; }
            ParseArgs a;
            fill_parse_args(c, a, nullptr, nullptr, nullptr, nullptr, 0, 0);
            const PeerTables pt = peer_tables(c);
            KMN_DISPATCH_W(c, { k_lookup_keys_peer<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(a, pt, dk, n, dout); });
            c->launches++;
            CK(c, cudaGetLastError());
            if (out_host) CK(c, cudaMemcpyAsync(counts, dout, n * 2, cudaMemcpyDeviceToHost, c->stream));
        }
        r = stream_barrier(c); if (r) return r;
        CK(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (c->nranks > 1) {
        // collective: every rank calls it (n may be 0); keys go to their owners in rounds of at most send_cap keys,
        // as DistributedReadSelector::_batchKmerLookup does per batch (src/DistributedFunctions.h:877-902)
        const u64 per = std::max<u64>(1, c->send_cap);
        u64 rounds = 0;
        r = allreduce_max_u64(c, (n + per - 1) / per, &rounds); if (r) return r;
        if (rounds == 0) return 0;
        r = ensure(c, c->lk_origin, (size_t)c->nranks * c->send_cap * 8); if (r) return r;
        r = ensure(c, c->lk_resp_in, (size_t)c->nranks * c->send_cap * 2); if (r) return r;
        const uint8_t *dk = keys;
        if (n && !is_device_ptr(keys)) {
            r = ensure(c, c->lk_keys, n * c->kb); if (r) return r;
            CK(c, cudaMemcpyAsync(c->lk_keys.p, keys, n * c->kb, cudaMemcpyHostToDevice, c->stream));
            dk = (const uint8_t *)c->lk_keys.p;
        }
        uint16_t *dout = counts;
        const bool out_host = n && !is_device_ptr(counts);
        if (out_host) { r = ensure(c, c->lk_out, n * 2); if (r) return r; dout = (uint16_t *)c->lk_out.p; }
        ParseArgs a;
        fill_parse_args(c, a, nullptr, nullptr, nullptr, nullptr, 0, 0);
        for (u64 i = 0; i < rounds; ++i) {
            const u64 k0 = std::min<u64>(n, i * per), k1 = std::min<u64>(n, (i + 1) * per);
            if (k1 > k0) {
                KMN_DISPATCH_W(c, { k_lookup_keys_dist<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(a, dk + k0 * c->kb, k1 - k0, dout + k0, (u64 *)c->lk_origin.p); });
                c->launches++;
                CK(c, cudaGetLastError());
            }
            r = lookup_exchange(c, 0, k1 > k0 ? dout + k0 : nullptr); if (r) return r;
        }
        if (out_host) CK(c, cudaMemcpyAsync(counts, dout, n * 2, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
#endif
    if (n == 0) return 0;
    const uint8_t *dk = keys;
    if (!is_device_ptr(keys)) {
        r = ensure(c, c->lk_keys, n * c->kb); if (r) return r;
        CK(c, cudaMemcpyAsync(c->lk_keys.p, keys, n * c->kb, cudaMemcpyHostToDevice, c->stream));
        dk = (const uint8_t *)c->lk_keys.p;
    }
    uint16_t *dout = counts;
    bool out_host = !is_device_ptr(counts);
    if (out_host) { r = ensure(c, c->lk_out, n * 2); if (r) return r; dout = (uint16_t *)c->lk_out.p; }
    KMN_DISPATCH_W(c, { k_lookup_keys<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, dk, n, (u32)c->kb, dout); });
    c->launches++;
    CK(c, cudaGetLastError());
    if (out_host) CK(c, cudaMemcpyAsync(counts, dout, n * 2, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}


// ---------------------------------------------------------------------------------------------------------
// multi-GPU lookup pass (DistributedReadSelector::scoreAndTrimReads, src/DistributedFunctions.h:903-1045): one round =
// requests out (keys to their owner), answers back (u16 counts in request order).  Every rank takes part in every
// round; a rank without work sends nothing.
// ---------------------------------------------------------------------------------------------------------
#ifdef KMN_WITH_NCCL
static int allreduce_max_u64(kmn_ctx *c, u64 mine, u64 *out)
{
    CK(c, cudaMemcpyAsync(c->scratch + 6, &mine, 8, cudaMemcpyHostToDevice, c->stream));
    ncclResult_t nr = ncclAllReduce(c->scratch + 6, c->scratch + 7, 1, ncclUint64, ncclMax, c->comm, c->stream);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllReduce failed: %s", ncclGetErrorString(nr));
    CK(c, cudaMemcpyAsync(out, c->scratch + 7, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// requests are in send_recs[dest][0..send_cursor[dest]) as W words per key, origins in lk_origin at the same positions
static int lookup_exchange(kmn_ctx *c, uint32_t min_depth, uint16_t *vals)
{
    const int R = c->nranks, W = c->W;
    ncclResult_t nr = ncclAllGather(c->send_cursor, c->all_counts, (size_t)R, ncclUint64, c->comm, c->stream);
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "ncclAllGather failed: %s", ncclGetErrorString(nr));
    std::vector<u64> counts((size_t)R * R);
    CK(c, cudaMemcpyAsync(counts.data(), c->all_counts, counts.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    u64 recv_total = 0;
    for (int p = 0; p < R; ++p) {
        if (p == c->rank) continue;
        if (counts[(size_t)c->rank * R + p] > c->send_cap) return fail(c, KMN_ERR_COMM, "lookup request region overflow");
        recv_total += counts[(size_t)p * R + c->rank];
    }
    if (recv_total * W > c->recv_cap * (u64)c->RW) return fail(c, KMN_ERR_COMM, "lookup receive region overflow");
    int r = ensure(c, c->lk_resp_out, recv_total * 2 + 16); if (r) return r;
    ncclGroupStart();
    u64 roff = 0;
    for (int p = 0; p < R; ++p) {
        if (p == c->rank) continue;
        u64 ns = counts[(size_t)c->rank * R + p], nrv = counts[(size_t)p * R + c->rank];
        if (ns) ncclSend(c->send_recs + (size_t)p * c->send_cap * W, ns * W, ncclUint64, p, c->comm, c->stream);
        if (nrv) ncclRecv(c->recv_recs + roff * W, nrv * W, ncclUint64, p, c->comm, c->stream);
        roff += nrv;
    }
    nr = ncclGroupEnd();
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "lookup request all-to-all failed: %s", ncclGetErrorString(nr));
    if (recv_total) {
        ProfScope ps(c, KMN_PROF_LOOKUP, recv_total);
        KMN_DISPATCH_W(c, { k_lookup_words<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->recv_recs, recv_total, (uint16_t *)c->lk_resp_out.p); });
        c->launches++;
        CK(c, cudaGetLastError());
    }
    ncclGroupStart();
    roff = 0;
    for (int p = 0; p < R; ++p) {
        if (p == c->rank) continue;
        u64 ns = counts[(size_t)c->rank * R + p], nrv = counts[(size_t)p * R + c->rank];
        if (nrv) ncclSend((uint16_t *)c->lk_resp_out.p + roff, nrv * 2, ncclUint8, p, c->comm, c->stream);
        if (ns) ncclRecv((uint16_t *)c->lk_resp_in.p + (size_t)p * c->send_cap, ns * 2, ncclUint8, p, c->comm, c->stream);
        roff += nrv;
    }
    nr = ncclGroupEnd();
    if (nr != ncclSuccess) return fail(c, KMN_ERR_COMM, "lookup answer all-to-all failed: %s", ncclGetErrorString(nr));
    if (vals) {
        dim3 grid(c->n_sms * 2, (unsigned)std::min(R, 8));
        k_scatter_answers<<<grid, 256, 0, c->stream>>>((const uint16_t *)c->lk_resp_in.p, (const u64 *)c->lk_origin.p, c->send_cap,
                                                       c->send_cursor, (u32)R, min_depth, vals);
        c->launches++;
        CK(c, cudaGetLastError());
    }
    CK(c, cudaMemsetAsync(c->send_cursor, 0, (size_t)R * 8, c->stream));
    return 0;
}
#endif

int kmn_trim_batch(kmn_ctx *c, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads, const uint8_t *discarded,
                   uint32_t min_depth, int scoring, uint32_t *trim_off, uint32_t *trim_len, float *score, uint8_t *was_trimmed)
{
    if (!c) return KMN_ERR_INVALID;
    if (scoring < 0 || scoring > 4) return fail(c, KMN_ERR_INVALID, "Invalid scoring type!");
    CK(c, cudaSetDevice(c->device));
    if (c->nranks > 1 && c->p2p && !c->finished)          // the round buffers are still in use by the count pass (header: KMN_ERR_STATE)
        return fail(c, KMN_ERR_STATE, "%s before kmn_count_finish on a multi-GPU context", __func__);
    int r = drain(c); if (r) return r;
#ifdef KMN_WITH_NCCL
    if (c->nranks > 1 && c->peer_lookup) { r = stream_barrier(c); if (r) return r; }       // every rank's table is final
    if (c->nranks > 1 && c->peer_lookup && n_reads == 0) return stream_barrier(c);         // (nothing to look up here)
    if (c->nranks > 1 && n_reads == 0) {           // a rank without reads still serves the other ranks' requests
        u64 rounds = 0;
        r = allreduce_max_u64(c, 0, &rounds); if (r) return r;
        for (u64 i = 0; i < rounds; ++i) { r = lookup_exchange(c, min_depth, nullptr); if (r) return r; }
        return 0;
    }
#endif
    if (n_reads == 0) return 0;
    if (!bases || !read_off || !trim_off || !trim_len || !score || !was_trimmed) return fail(c, KMN_ERR_INVALID, "null argument");
    BatchPtrs bp;
    r = stage_inputs(c, bases, nullptr, read_off, n_reads, discarded, false, bp);
    if (r) return r;
    r = ensure(c, c->vals, bp.total_bytes * 2 + 64); if (r) return r;
    r = ensure(c, c->first_nx, n_reads * 4); if (r) return r;
    ParseArgs a;
    fill_parse_args(c, a, bp.bases, nullptr, bp.off, bp.disc, n_reads, bp.total_bytes);
    const int grid = c->n_sms * 8;
    if (c->nranks > 1 && c->peer_lookup) {
#ifdef KMN_WITH_NCCL
        {
            ProfScope ps(c, KMN_PROF_LOOKUP, bp.total_bytes);
            const PeerTables pt = peer_tables(c);
            KMN_DISPATCH_W(c, { k_lookup_vals_peer<W_><<<grid, 256, 0, c->stream>>>(a, pt, min_depth, (uint16_t *)c->vals.p, (u32 *)c->first_nx.p); });
            c->launches++;
            CK(c, cudaGetLastError());
        }
        r = stream_barrier(c); if (r) return r;                                            // nobody reads this rank's table any more
#endif
    } else if (c->nranks > 1) {
#ifdef KMN_WITH_NCCL
        // rounds: read ranges whose k-mer positions fit one request region even if every k-mer had the same owner
        r = ensure(c, c->lk_origin, (size_t)c->nranks * c->send_cap * 8); if (r) return r;
        r = ensure(c, c->lk_resp_in, (size_t)c->nranks * c->send_cap * 2); if (r) return r;
        std::vector<u64> hoff;
        const u64 *ho = reinterpret_cast<const u64 *>(read_off);
        if (!bp.off_on_host) {
            hoff.resize(n_reads + 1);
            CK(c, cudaMemcpyAsync(hoff.data(), bp.off, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
            CK(c, cudaStreamSynchronize(c->stream));
            ho = hoff.data();
        }
        std::vector<uint64_t> cuts(1, 0);
        u64 acc = 0;
        for (uint64_t q = 0; q < n_reads; ++q) {
            u64 len = ho[q + 1] - ho[q];
            u64 np = len >= c->o.kmer_size ? len - c->o.kmer_size + 1 : 0;
            if (np > c->send_cap) return fail(c, KMN_ERR_INVALID, "read %llu has more k-mers than the request region (%llu)", (unsigned long long)q, (unsigned long long)c->send_cap);
            if (acc + np > c->send_cap) { cuts.push_back(q); acc = 0; }
            acc += np;
        }
        cuts.push_back(n_reads);
        u64 rounds = 0;
        r = allreduce_max_u64(c, cuts.size() - 1, &rounds); if (r) return r;
        for (u64 i = 0; i < rounds; ++i) {
            if (i + 1 < cuts.size()) {
                ParseArgs ai = a;
                ai.read_off = bp.off + cuts[i]; ai.discarded = bp.disc ? bp.disc + cuts[i] : nullptr; ai.n_reads = cuts[i + 1] - cuts[i];
                ProfScope ps(c, KMN_PROF_LOOKUP, ai.n_reads);
                KMN_DISPATCH_W(c, { k_lookup_vals_dist<W_><<<grid, 256, 0, c->stream>>>(ai, min_depth, (uint16_t *)c->vals.p, (u32 *)c->first_nx.p + cuts[i], (u64 *)c->lk_origin.p); });
                c->launches++;
                CK(c, cudaGetLastError());
            }
            r = lookup_exchange(c, min_depth, (uint16_t *)c->vals.p); if (r) return r;
        }
#else
        return fail(c, KMN_ERR_COMM, "library built without NCCL");
#endif
    } else {
        ProfScope ps(c, KMN_PROF_LOOKUP, bp.total_bytes);
        KMN_DISPATCH_W(c, { k_lookup_vals<W_><<<grid, 256, 0, c->stream>>>(a, min_depth, (uint16_t *)c->vals.p, (u32 *)c->first_nx.p); });
        c->launches++;
        CK(c, cudaGetLastError());
    }
    TrimArgs t;
    t.read_off = bp.off; t.discarded = bp.disc; t.vals = (const uint16_t *)c->vals.p; t.first_nx = (const u32 *)c->first_nx.p;
    t.n_reads = n_reads; t.k = c->o.kmer_size; t.min_depth = min_depth; t.scoring = scoring;
    const bool host_out = !is_device_ptr(trim_off);
    if (host_out) {
        r = ensure(c, c->out_off, n_reads * 4); if (r) return r;
        r = ensure(c, c->out_len, n_reads * 4); if (r) return r;
        r = ensure(c, c->out_score, n_reads * 4); if (r) return r;
        r = ensure(c, c->out_trim, n_reads); if (r) return r;
        t.trim_off = (u32 *)c->out_off.p; t.trim_len = (u32 *)c->out_len.p; t.score = (float *)c->out_score.p; t.was_trimmed = (uint8_t *)c->out_trim.p;
    } else { t.trim_off = trim_off; t.trim_len = trim_len; t.score = score; t.was_trimmed = was_trimmed; }
    {
        ProfScope ps(c, KMN_PROF_TRIM, n_reads);
        k_trim_score<<<c->n_sms * 8, 256, 0, c->stream>>>(t);
    }
    c->launches++;
    CK(c, cudaGetLastError());
    if (host_out) {
        CK(c, cudaMemcpyAsync(trim_off, t.trim_off, n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemcpyAsync(trim_len, t.trim_len, n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemcpyAsync(score, t.score, n_reads * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemcpyAsync(was_trimmed, t.was_trimmed, n_reads, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
    }
    return release_inputs(c, bp);
}

int kmn_export(kmn_ctx *c, uint32_t min_count, uint8_t *keys, uint16_t *count, uint16_t *dir, float *wsum, uint32_t *ext,
               uint64_t cap, uint64_t *n_out)
{
    if (!c || !n_out) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->nranks > 1 && c->p2p && !c->finished)          // the round buffers are still in use by the count pass (header: KMN_ERR_STATE)
        return fail(c, KMN_ERR_STATE, "%s before kmn_count_finish on a multi-GPU context", __func__);
    int r = drain(c); if (r) return r;
    if (min_count < 1) min_count = 1;
    CK(c, cudaMemsetAsync(c->scratch, 0, 16, c->stream));
    KMN_DISPATCH_W(c, { k_count_live<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->n_slots, min_count, c->scratch, c->scratch + 1); });
    c->launches++;
    u64 n_live = 0;
    CK(c, cudaMemcpyAsync(&n_live, c->scratch, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *n_out = n_live;
    if (cap == 0 || n_live == 0) return 0;
    if (cap < n_live) return fail(c, KMN_ERR_INVALID, "export capacity %llu < %llu entries", (unsigned long long)cap, (unsigned long long)n_live);
    uint8_t *dk = nullptr; uint16_t *dc = nullptr, *dd = nullptr; float *dw = nullptr; u32 *de = nullptr;
    if (keys) CK(c, cudaMalloc((void **)&dk, n_live * c->kb));
    if (count) CK(c, cudaMalloc((void **)&dc, n_live * 2));
    if (dir) CK(c, cudaMalloc((void **)&dd, n_live * 2));
    if (wsum) CK(c, cudaMalloc((void **)&dw, n_live * 4));
    if (ext) CK(c, cudaMalloc((void **)&de, n_live * 48));
    CK(c, cudaMemsetAsync(c->scratch, 0, 8, c->stream));
    KMN_DISPATCH_W(c, { k_export<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->n_slots, min_count, (u32)c->kb, n_live, c->scratch, dk, dc, dd, dw, de); });
    c->launches++;
    CK(c, cudaGetLastError());
    if (keys) CK(c, cudaMemcpyAsync(keys, dk, n_live * c->kb, cudaMemcpyDeviceToHost, c->stream));
    if (count) CK(c, cudaMemcpyAsync(count, dc, n_live * 2, cudaMemcpyDeviceToHost, c->stream));
    if (dir) CK(c, cudaMemcpyAsync(dir, dd, n_live * 2, cudaMemcpyDeviceToHost, c->stream));
    if (wsum) CK(c, cudaMemcpyAsync(wsum, dw, n_live * 4, cudaMemcpyDeviceToHost, c->stream));
    if (ext) CK(c, cudaMemcpyAsync(ext, de, n_live * 48, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(dk); cudaFree(dc); cudaFree(dd); cudaFree(dw); cudaFree(de);
    return 0;
}

int kmn_import(kmn_ctx *c, const uint8_t *keys, const uint16_t *count, const uint16_t *dir, const float *wsum, const uint32_t *ext, uint64_t n)
{
    if (!c || (n && (!keys || !count))) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    int r = drain(c); if (r) return r;
    if (n == 0) { c->finished = true; return 0; }
    uint8_t *dk = nullptr; uint16_t *dc = nullptr, *dd = nullptr; float *dw = nullptr; u32 *de = nullptr;
    CK(c, cudaMalloc((void **)&dk, n * c->kb)); CK(c, cudaMalloc((void **)&dc, n * 2));
    CK(c, cudaMemcpyAsync(dk, keys, n * c->kb, cudaMemcpyDefault, c->stream));
    CK(c, cudaMemcpyAsync(dc, count, n * 2, cudaMemcpyDefault, c->stream));
    if (dir) { CK(c, cudaMalloc((void **)&dd, n * 2)); CK(c, cudaMemcpyAsync(dd, dir, n * 2, cudaMemcpyDefault, c->stream)); }
    if (wsum && c->table.wsum) { CK(c, cudaMalloc((void **)&dw, n * 4)); CK(c, cudaMemcpyAsync(dw, wsum, n * 4, cudaMemcpyDefault, c->stream)); }
    if (ext && c->table.ext) { CK(c, cudaMalloc((void **)&de, n * 48)); CK(c, cudaMemcpyAsync(de, ext, n * 48, cudaMemcpyDefault, c->stream)); }
    KMN_DISPATCH_W(c, { k_import<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, dk, dc, dd, dw, de, n, (u32)c->kb, c->ctr); });
    c->launches++;
    CK(c, cudaGetLastError());
    Counters h;
    CK(c, cudaMemcpyAsync(&h, c->ctr, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(dk); cudaFree(dc); cudaFree(dd); cudaFree(dw); cudaFree(de);
    if (h.table_full) return fail(c, KMN_ERR_TABLE_FULL, "count table overflow during import: %llu entries did not fit (capacity %llu slots)",
                                  (unsigned long long)h.table_full, (unsigned long long)c->n_slots);
    c->finished = true;
    return 0;
}

int kmn_subtract(kmn_ctx *c, kmn_ctx *other, uint64_t *removed_entries, uint64_t *removed_instances)
{
    if (!c || !other || c == other) return KMN_ERR_INVALID;
    if (c->device != other->device || c->o.kmer_size != other->o.kmer_size) return fail(c, KMN_ERR_INVALID, "kmn_subtract needs two spectra of the same k on the same device");
    CK(c, cudaSetDevice(c->device));
    int r = drain(c); if (r) return r;
    r = drain(other); if (r) return r;
    CK(c, cudaStreamSynchronize(other->stream));
    CK(c, cudaMemsetAsync(c->scratch, 0, 16, c->stream));
    KMN_DISPATCH_W(c, { k_subtract<W_><<<c->n_sms * 8, 256, 0, c->stream>>>(c->table, c->n_slots, other->table, c->scratch); });
    c->launches++;
    CK(c, cudaGetLastError());
    u64 out[2];
    CK(c, cudaMemcpyAsync(out, c->scratch, 16, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (removed_entries) *removed_entries = out[0];
    if (removed_instances) *removed_instances = out[1];
    return 0;
}

int kmn_debug_owner(kmn_ctx *c, const uint8_t *keys, uint64_t n, uint32_t nranks, uint32_t *owner)
{
    if (!c || !keys || !owner || nranks == 0) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (n == 0) return 0;
    uint8_t *dk = nullptr; u32 *dout = nullptr;
    CK(c, cudaMalloc((void **)&dk, n * c->kb)); CK(c, cudaMalloc((void **)&dout, n * 4));
    CK(c, cudaMemcpyAsync(dk, keys, n * c->kb, cudaMemcpyDefault, c->stream));
    KMN_DISPATCH_W(c, { k_debug_owner<W_><<<c->n_sms * 4, 256, 0, c->stream>>>(dk, n, (u32)c->kb, nranks, c->o.hash_kind == KMN_HASH_LOOKUP8_HASH2, dout); });
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(owner, dout, n * 4, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(dk); cudaFree(dout);
    return 0;
}

int kmn_debug_kmers(kmn_ctx *c, const uint8_t *bases, const uint8_t *quals, const uint64_t *read_off, uint64_t n_reads,
                    uint8_t *keys, uint8_t *is_fwd, float *weight, uint64_t *hash, uint64_t *n_out)
{
    if (!c || !bases || !quals || !read_off || !n_out) return KMN_ERR_INVALID;
    if (is_device_ptr(read_off)) return fail(c, KMN_ERR_INVALID, "kmn_debug_kmers needs host offsets");
    CK(c, cudaSetDevice(c->device));
    const uint32_t k = c->o.kmer_size;
    std::vector<u64> koff(n_reads + 1, 0);
    for (uint64_t r = 0; r < n_reads; ++r) {
        u64 len = read_off[r + 1] - read_off[r];
        koff[r + 1] = koff[r] + (len >= k ? len - k + 1 : 0);
    }
    const u64 nk = koff[n_reads];
    *n_out = nk;
    if (!keys || nk == 0) return 0;
    BatchPtrs bp;
    int r = stage_inputs(c, bases, quals, read_off, n_reads, nullptr, true, bp); if (r) return r;
    u64 *dko; uint8_t *dk, *df; float *dw; u64 *dh;
    CK(c, cudaMalloc((void **)&dko, (n_reads + 1) * 8));
    CK(c, cudaMalloc((void **)&dk, nk * c->kb)); CK(c, cudaMalloc((void **)&df, nk));
    CK(c, cudaMalloc((void **)&dw, nk * 4)); CK(c, cudaMalloc((void **)&dh, nk * 8));
    CK(c, cudaMemcpyAsync(dko, koff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    ParseArgs a;
    fill_parse_args(c, a, bp.bases, bp.quals, bp.off, nullptr, n_reads, bp.total_bytes);
    KMN_DISPATCH_W(c, { k_debug_kmers<W_><<<c->n_sms * 4, 128, 0, c->stream>>>(a, dko, dk, df, dw, dh); });
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(keys, dk, nk * c->kb, cudaMemcpyDeviceToHost, c->stream));
    if (is_fwd) CK(c, cudaMemcpyAsync(is_fwd, df, nk, cudaMemcpyDeviceToHost, c->stream));
    if (weight) CK(c, cudaMemcpyAsync(weight, dw, nk * 4, cudaMemcpyDeviceToHost, c->stream));
    if (hash) CK(c, cudaMemcpyAsync(hash, dh, nk * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(dko); cudaFree(dk); cudaFree(df); cudaFree(dw); cudaFree(dh);
    return release_inputs(c, bp);
}

int kmn_profile_enable(kmn_ctx *c, int on)
{
    if (!c) return KMN_ERR_INVALID;
    c->prof_on = on != 0;
    return 0;
}

int kmn_profile_read(kmn_ctx *c, kmn_profile *out)
{
    if (!c || !out) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->s_comm) CK(c, cudaStreamSynchronize(c->s_comm));
    if (c->s_insert != c->stream) CK(c, cudaStreamSynchronize(c->s_insert));
    for (auto &e : c->prof_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        c->prof_acc.ms[e.kind] += ms; c->prof_acc.launches[e.kind]++; c->prof_acc.units[e.kind] += e.units;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    c->prof_events.clear();
    *out = c->prof_acc;
    memset(&c->prof_acc, 0, sizeof c->prof_acc);
    return 0;
}

int kmn_sync(kmn_ctx *c)
{
    if (!c) return KMN_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->s_copy));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->s_comm) CK(c, cudaStreamSynchronize(c->s_comm));
    if (c->s_insert != c->stream) CK(c, cudaStreamSynchronize(c->s_insert));
    return 0;
}

