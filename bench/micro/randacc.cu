// Random-access microbenchmark for the count-table design.
// Measures, on one B200, the practical ceilings the k-mer count/lookup kernels are judged against:
//   red    : 64-bit atomicAdd without return (RED) to uniformly random 16-B slots
//   ldred  : 16-B slot load + compare + RED            (the "hit" path of the insert kernel)
//   casadd : 64-bit atomicCAS on the key word + RED on the value word (the "claim" path)
//   ld16   : 16-B random read, ld.global.nc            (the lookup path)
// for table sizes from L2-resident to HBM-resident.  Output: one CSV line per (op, size).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o randacc randacc.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

struct __align__(16) Slot { unsigned long long key; unsigned long long val; };

template <int OP, int UNROLL>
__global__ void __launch_bounds__(256) k_rand(Slot* __restrict__ tab, uint64_t nslots, uint64_t nops, uint64_t seed,
                                              unsigned long long* __restrict__ sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t nthreads = gridDim.x * (uint64_t)blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t i = tid; i < nops; i += nthreads * UNROLL) {
        uint64_t idx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint64_t h = mix64((i + u * nthreads) ^ seed);
            idx[u] = __umul64hi(h, nslots);
        }
        if (OP == 0) {          // RED only
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops) atomicAdd(&tab[idx[u]].val, 1ULL);
        } else if (OP == 1) {   // load 16 B, compare, RED
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops)
                    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(&tab[idx[u]]));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops) {
                    if (v[u].x != 0xffffffffu) atomicAdd(&tab[idx[u]].val, 1ULL);
                    else acc += v[u].y;
                }
        } else if (OP == 2) {   // CAS on key, RED on value
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops) {
                    unsigned long long want = idx[u] | 1ULL;
                    unsigned long long old = atomicCAS(&tab[idx[u]].key, 0ULL, want);
                    if (old == 0ULL || old == want) atomicAdd(&tab[idx[u]].val, 1ULL);
                    else acc += old;
                }
        } else {                // 16-B read-only
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(&tab[idx[u]]));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (i + u * nthreads < nops) acc += v[u].x + v[u].z;
        }
    }
    if (acc == 0x123456789ULL) *sink = acc;
}

template <int OP, int UNROLL>
static double run(Slot* tab, uint64_t nslots, uint64_t nops, unsigned long long* sink, int blocks) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_rand<OP, UNROLL><<<blocks, 256>>>(tab, nslots, nops / 8, 1, sink);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_rand<OP, UNROLL><<<blocks, 256>>>(tab, nslots, nops, 1234567 + rep, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return nops / (best * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    size_t freeb, totb; CK(cudaMemGetInfo(&freeb, &totb));
    fprintf(stderr, "device %s SMs=%d L2=%d MB free=%.1f GB\n", p.name, sms, p.l2CacheSize >> 20, freeb / 1e9);
    uint64_t max_bytes = (argc > 1) ? strtoull(argv[1], 0, 10) << 20 : (32ULL << 30);
    uint64_t nops = (argc > 2) ? strtoull(argv[2], 0, 10) : (1ULL << 30);
    Slot* tab; CK(cudaMalloc(&tab, max_bytes));
    CK(cudaMemset(tab, 0, max_bytes));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    printf("op,unroll,table_MiB,blocks_per_sm,Gops_per_s,GBps_64B_per_op\n");
    uint64_t sizes_mb[] = {8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096, 16384, 32768};
    for (uint64_t smb : sizes_mb) {
        uint64_t bytes = smb << 20; if (bytes > max_bytes) break;
        uint64_t nslots = bytes / sizeof(Slot);
        for (int bps : {4, 8}) {
            int blocks = sms * bps;
            double g;
            g = run<0, 1>(tab, nslots, nops, sink, blocks); printf("red,1,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 64);
            g = run<0, 4>(tab, nslots, nops, sink, blocks); printf("red,4,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 64);
            g = run<1, 1>(tab, nslots, nops, sink, blocks); printf("ldred,1,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 64);
            g = run<1, 4>(tab, nslots, nops, sink, blocks); printf("ldred,4,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 64);
            g = run<2, 1>(tab, nslots, nops, sink, blocks); printf("casadd,1,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 64);
            g = run<3, 1>(tab, nslots, nops, sink, blocks); printf("ld16,1,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 32);
            g = run<3, 4>(tab, nslots, nops, sink, blocks); printf("ld16,4,%llu,%d,%.2f,%.1f\n", (unsigned long long)smb, bps, g, g * 32);
            fflush(stdout);
        }
    }
    return 0;
}
