// Microbenchmark: DRAM bytes and time per random 16-byte table gather on B200, for the load flavours the
// probe kernel could use.  Run under ncu to read dram__bytes_read.sum per kernel (profiles/ evidence).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
template <int V>
__device__ __forceinline__ uint4 load16(const uint4 *p)
{
    uint4 v;
    if (V == 0) v = *p;
    else if (V == 1) v = __ldcg(p);
    else if (V == 2) v = __ldcs(p);
    else if (V == 3) v = __ldg(p);
    else if (V == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == 5) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (V == 6) { unsigned long long a, b, c, d; const uint4 *q = (const uint4 *)((uintptr_t)p & ~(uintptr_t)31);
        asm volatile("ld.global.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(q)); v = make_uint4((unsigned)a, (unsigned)b, (unsigned)c, (unsigned)d); }
    else { const uint2 *q = (const uint2 *)p; uint2 a = __ldcg(q), b = __ldcg(q + 1); v = make_uint4(a.x, a.y, b.x, b.y); }
    return v;
}
template <int V>
__global__ void __launch_bounds__(256) gather(const uint4 *tab, uint64_t nslots, uint64_t n, unsigned *out)
{
    unsigned acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint64_t s = mix(i * 0x9E3779B97F4A7C15ULL + V) % nslots;
        uint4 v = load16<V>(tab + s);
        acc += v.x ^ v.z;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// gather followed by a fire-and-forget RED on the same entry (the probe's saturated-hit pattern)
__global__ void __launch_bounds__(256) gather_red(uint4 *tab, uint64_t nslots, uint64_t n, unsigned *out)
{
    unsigned acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint64_t s = mix(i * 0x9E3779B97F4A7C15ULL + 99) % nslots;
        uint4 v = __ldcg(tab + s);
        atomicAdd((int *)&tab[s].z, 1);
        acc += v.x;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int V> void run(const char *name, uint4 *tab, uint64_t nslots, uint64_t n, unsigned *out)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<V><<<148 * 8, 256>>>(tab, nslots, n / 8, out);
    cudaEventRecord(a); gather<V><<<148 * 8, 256>>>(tab, nslots, n, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %7.2f G gathers/s  (16B: %.0f GB/s useful)\n", name, ms, n / ms / 1e6, n * 16.0 / ms / 1e6);
}
int main(int argc, char **argv)
{
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set limit %d -> %s\n", gran, cudaGetErrorString(e)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    uint64_t nslots = (4ull << 30) / 16, n = 256ull << 20;
    uint4 *tab; unsigned *out; cudaMalloc(&tab, nslots * 16); cudaMemset(tab, 0, nslots * 16); cudaMalloc(&out, 4);
    run<0>("ld.global (default .ca)", tab, nslots, n, out);
    run<1>("__ldcg", tab, nslots, n, out);
    run<2>("__ldcs", tab, nslots, n, out);
    run<3>("__ldg (.nc)", tab, nslots, n, out);
    run<4>("L1::no_allocate", tab, nslots, n, out);
    run<5>("ld.cv", tab, nslots, n, out);
    run<6>("256-bit ld L2::evict_first", tab, nslots, n, out);
    run<7>("2 x 8B __ldcg", tab, nslots, n, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); gather_red<<<148 * 8, 256>>>(tab, nslots, n, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %7.2f G gathers/s\n", "__ldcg + RED", ms, n / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
