// LSU cost model microbenchmark for the count pass (round 2).
// The count pass is bound by per-lane costs of scattered memory operations in the SM's load/store unit, so every design
// decision of phase 1 (binning) and phase 2 (insert) is a count of such operations per k-mer record.  This program measures
// them one by one on the target GPU:
//   smem side (phase 1):  atoms      : shared atomicAdd with return to NB spread counters (the bin-position claim)
//                         atoms_sts  : + one 8-byte shared store into the bin's ring
//                         atoms_stg  : + one scattered 8-byte global store (round-1 scatter: one frontier per (CTA, bin))
//                         ring       : atoms_sts + per-round cooperative flush of the rings as coalesced segments
//   L2 side (phase 2):    pair_red   : 32-byte slot-pair load + 64-bit RED (round-1 insert)
//                         cas_red    : 64-bit CAS on the key word + RED
//                         cas128_red : 128-bit CAS on the two key words of a pair + RED
//                         atom_only  : one returning 64-bit atomicAdd per record
// Output: CSV, G records/s and cycles per lane at the measured SM clock.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o lsu lsu.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef unsigned long long u64;
typedef unsigned int u32;

__device__ __forceinline__ u64 mix64(u64 x)
{
    x *= 0x9E3779B97F4A7C15ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    return x;
}

// ---------------------------------------------------------------------------------------------------------
// phase-1 side.  Every thread produces 8 records per round (one walker step), n_rounds rounds.
// MODE 0 atoms, 1 atoms_sts, 2 atoms_stg, 3 ring (atoms + sts + barrier + flush of all rings + barrier)
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void k_bin(u32 n_bins, u32 R, u32 n_rounds, u64 *out, u32 sub_cap, u64 *sink)
{
    extern __shared__ __align__(16) unsigned char smem[];
    u32 *cnt = reinterpret_cast<u32 *>(smem);
    u32 *base = cnt + ((n_bins + 31u) & ~31u);
    u64 *ring = reinterpret_cast<u64 *>(base + ((n_bins + 31u) & ~31u));
    for (u32 i = threadIdx.x; i < n_bins; i += blockDim.x) { cnt[i] = 0; base[i] = 0; }
    __syncthreads();
    u64 *cbase = out + (size_t)blockIdx.x * n_bins * sub_cap;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 acc = 0;
    for (u32 r = 0; r < n_rounds; ++r) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const u64 rec = mix64((tid * n_rounds + r) * 8 + u + 12345);
            const u32 b = __umulhi((u32)(rec >> 32), n_bins);
            const u32 p = atomicAdd(&cnt[b], 1u);
            if (MODE == 0) acc += p;
            if (MODE == 1) ring[(size_t)b * R + (p % R)] = rec;
            if (MODE == 2) cbase[(size_t)b * sub_cap + (p % sub_cap)] = rec;
            if (MODE == 3) {
                if (p < R) ring[(size_t)b * R + p] = rec;
                else { const u32 pos = base[b] + p; if (pos < sub_cap) cbase[(size_t)b * sub_cap + pos] = rec; }
            }
        }
        if (MODE == 3) {
            __syncthreads();
            // flush every ring as one coalesced segment: 8 lanes per bin, 4 bins per warp and pass
            const u32 lane = threadIdx.x & 31u, sub = lane >> 3, l8 = lane & 7u;
            const u32 warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
            for (u32 b0 = warp * 4; b0 < n_bins; b0 += n_warps * 4) {
                const u32 b = b0 + sub;
                if (b < n_bins) {
                    const u32 c = cnt[b], bs = base[b];
                    const u32 n = c < R ? c : R;
                    for (u32 i = l8; i < n; i += 8) {
                        const u32 pos = bs + i;
                        if (pos < sub_cap) cbase[(size_t)b * sub_cap + pos] = ring[(size_t)b * R + i];
                    }
                    __syncwarp(0xffu << (sub * 8));
                    if (l8 == 0) { base[b] = bs + c; cnt[b] = 0; }
                }
            }
            __syncthreads();
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

// ---------------------------------------------------------------------------------------------------------
// phase-2 side.  U records in flight per thread, random slots inside a region of n_pairs 32-byte slot pairs.
// pair = {val0, key0, val1, key1} (round-1 layout) ; for cas128: {key0, key1, val0, val1}
// ---------------------------------------------------------------------------------------------------------
template <int OP, int U>
__global__ void __launch_bounds__(256) k_ins(u64 *tab, u64 n_pairs, u64 nops, u64 seed, u64 *sink)
{
    const u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x, nth = gridDim.x * (u64)blockDim.x;
    u64 acc = 0;
    for (u64 i = tid; i < nops; i += nth * U) {
        u64 idx[U], key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const u64 h = mix64((i + u * nth) ^ seed); idx[u] = __umul64hi(h, n_pairs); key[u] = h | 1ull; }
        if (OP == 0) {
            u64 v0[U], k0[U], v1[U], k1[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v0[u]), "=l"(k0[u]), "=l"(v1[u]), "=l"(k1[u]) : "l"(tab + idx[u] * 4));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k0[u] != key[u] * 3) atomicAdd(&tab[idx[u] * 4 + ((key[u] >> 1) & 1) * 2], 1ull);
                else acc += v0[u] + v1[u] + k1[u];
            }
        } else if (OP == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u64 want = idx[u] + 1;       // one key per slot: every CAS after the first is a "hit"
                const u64 old = atomicCAS(&tab[idx[u] * 4 + 1], 0ull, want);
                if (old == 0ull || old == want) atomicAdd(&tab[idx[u] * 4], 1ull); else acc += old;
            }
        } else if (OP == 2) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u64 want = idx[u] + 1;
                u64 o0, o1;
                asm volatile("{\n\t.reg .b128 c, d, e;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 d, {%5, %6};\n\t"
                             "atom.global.cas.b128 e, [%2], c, d;\n\tmov.b128 {%0, %1}, e;\n\t}"
                             : "=l"(o0), "=l"(o1) : "l"(tab + idx[u] * 4), "l"(0ull), "l"(0ull), "l"(want), "l"(0ull) : "memory");
                if (o0 == 0ull || o0 == want) atomicAdd(&tab[idx[u] * 4 + 2], 1ull); else acc += o0 + o1;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += atomicAdd(&tab[idx[u] * 4], 1ull);
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <typename F>
static float time_ms(F &&launch, int reps = 3)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return best;
}

int main()
{
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz * 1e-6;
    fprintf(stderr, "device %s SMs=%d clock=%.3f GHz\n", p.name, sms, ghz);
    u64 *sink; CK(cudaMalloc(&sink, 8));
    printf("bench,variant,param,threads_per_sm,Grec_per_s,cyc_per_lane\n");

    // ---- phase-1 side ----
    {
        const u32 sub_cap = 4096;
        const u32 bins_list[] = {355, 710, 1420, 2840};
        u64 *out; CK(cudaMalloc(&out, (size_t)sms * 2 * 2840 * sub_cap * 8));
        for (u32 nb : bins_list) {
            for (int tpb : {512, 1024}) {
                for (int ctas : {1, 2}) {
                    if (tpb * ctas > 2048) continue;
                    const size_t budget = (size_t)(ctas == 1 ? 200 : 100) * 1024;
                    const size_t hdr = 2 * (size_t)((nb + 31u) & ~31u) * 4;
                    if (hdr + (size_t)nb * 8 * 4 > budget) continue;
                    u32 R = (u32)((budget - hdr) / ((size_t)nb * 8));
                    if (R > 64) R = 64;
                    const size_t smem = hdr + (size_t)nb * R * 8;
                    const u32 rounds = 4096 / 8;      // 4096 records per thread
                    const double recs = (double)sms * ctas * tpb * rounds * 8;
                    auto run = [&](auto kern, const char *name) {
                        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        float ms = time_ms([&] { kern<<<sms * ctas, tpb, smem>>>(nb, R, rounds, out, sub_cap, sink); });
                        const double g = recs / (ms * 1e-3) / 1e9;
                        printf("bin,%s,bins=%u;R=%u;ctas=%d,%d,%.2f,%.3f\n", name, nb, R, ctas, tpb * ctas, g, sms * ghz / g);
                        fflush(stdout);
                    };
                    run(k_bin<0>, "atoms");
                    run(k_bin<1>, "atoms_sts");
                    run(k_bin<2>, "atoms_stg");
                    run(k_bin<3>, "ring");
                }
            }
        }
        CK(cudaFree(out));
    }
    // ---- phase-2 side ----
    {
        const u64 max_bytes = 128ull << 20;
        u64 *tab; CK(cudaMalloc(&tab, max_bytes));
        const u64 nops = 1ull << 30;
        for (u64 mb : {8ull, 16ull, 32ull, 64ull, 128ull}) {
            const u64 n_pairs = (mb << 20) / 32;
            for (int bps : {4, 8}) {
                auto run = [&](auto kern, const char *name, int u) {
                    CK(cudaMemset(tab, 0, max_bytes));
                    int rep = 0;
                    float ms = time_ms([&] { kern<<<sms * bps, 256>>>(tab, n_pairs, nops, 777 + rep++, sink); });
                    const double g = nops / (ms * 1e-3) / 1e9;
                    printf("ins,%s,MiB=%llu;U=%d;ctas=%d,%d,%.2f,%.3f\n", name, mb, u, bps, bps * 256, g, sms * ghz / g);
                    fflush(stdout);
                };
                run(k_ins<0, 2>, "pair_red", 2);
                run(k_ins<0, 4>, "pair_red", 4);
                run(k_ins<1, 2>, "cas_red", 2);
                run(k_ins<2, 2>, "cas128_red", 2);
                run(k_ins<3, 2>, "atom_only", 2);
            }
        }
    }
    return 0;
}
