// What would locality buy?  Random 16-byte gather + 64-bit atomic on a table region of varying size: when the
// region fits the 126 MB L2 the line is served from L2, otherwise every access costs a 128-byte DRAM line.
// Also times a one-pass 256-way partition of 16-byte operation records (what bucketing the probes by table
// region would cost per pass).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__global__ void __launch_bounds__(256) probe(uint4 *tab, uint64_t nslots, uint64_t n, int do_atomic, unsigned *out)
{
    unsigned acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint64_t s = mix(i * 0x9E3779B97F4A7C15ULL) % nslots;
        uint4 v; asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(tab + s));
        if (do_atomic) atomicAdd(reinterpret_cast<unsigned long long *>(&tab[s].z), 0x0000000100000001ull);
        acc += v.x;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// one pass of a 256-way scatter: histogram is assumed known (uniform), each op goes to bucket (slot >> shift)
__global__ void __launch_bounds__(256) scatter(const uint4 *in, uint4 *outp, unsigned *cursors, uint64_t n, uint64_t per_bucket)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint4 v = in[i];
        unsigned b = (unsigned)(mix(i) & 255u);
        unsigned pos = atomicAdd(&cursors[b], 1u);
        if (pos < per_bucket) outp[(uint64_t)b * per_bucket + pos] = v;
    }
}
int main()
{
    uint64_t maxslots = (8ull << 30) / 16, n = 256ull << 20;
    uint4 *tab; unsigned *out; cudaMalloc(&tab, maxslots * 16); cudaMemset(tab, 0, maxslots * 16); cudaMalloc(&out, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int atom = 0; atom < 2; atom++)
        for (uint64_t mb : {8ull, 32ull, 64ull, 96ull, 128ull, 256ull, 1024ull, 8192ull})
        {
            uint64_t ns = (mb << 20) / 16;
            probe<<<148 * 8, 256>>>(tab, ns, n / 4, atom, out);
            cudaEventRecord(a); probe<<<148 * 8, 256>>>(tab, ns, n, atom, out); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("%s region %5llu MB: %7.2f ms  %6.1f G ops/s\n", atom ? "gather+atomic64" : "gather         ", (unsigned long long)mb, ms, n / ms / 1e6);
        }
    // scatter cost
    uint64_t m = 128ull << 20, per = (m / 256) * 5 / 4;
    uint4 *in, *outp; unsigned *cur; cudaMalloc(&in, m * 16); cudaMalloc(&outp, per * 256 * 16); cudaMalloc(&cur, 1024);
    cudaMemset(in, 1, m * 16);
    for (int rep = 0; rep < 2; rep++)
    {
        cudaMemset(cur, 0, 1024);
        cudaEventRecord(a); scatter<<<148 * 8, 256>>>(in, outp, cur, m, per); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("256-way scatter of %llu M 16-B records (global cursors): %7.2f ms  %6.1f G rec/s  %6.0f GB/s r+w\n", (unsigned long long)(m >> 20), ms, m / ms / 1e6, m * 32.0 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
