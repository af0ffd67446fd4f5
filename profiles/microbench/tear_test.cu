// Do 16-byte vector loads observe a 64-bit RED on bytes 8..15 atomically?  Writers add (+1,+1) to the
// (count,aux) word of a few slots; readers check that count - aux stays 0 for three load flavours.
#include <cuda_runtime.h>
#include <cstdio>
struct __align__(16) Slot { unsigned long long key; int count; unsigned aux; };
__global__ void k(Slot *tab, int nslots, unsigned long long *bad, int iters)
{
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    bool writer = (tid & 1);
    for (int it = 0; it < iters; it++)
    {
        Slot *s = &tab[(tid * 7 + it) % nslots];
        if (writer)
            atomicAdd(reinterpret_cast<unsigned long long *>(&s->count), 0x0000000100000001ull);
        else
        {
            uint4 a;
            asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(s));
            if (a.z != a.w) atomicAdd(&bad[0], 1ull);
            unsigned long long k2, w2;
            asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(k2), "=l"(w2) : "l"(s));
            if ((unsigned)w2 != (unsigned)(w2 >> 32)) atomicAdd(&bad[1], 1ull);
            unsigned long long w;
            asm volatile("ld.global.u64 %0, [%1];" : "=l"(w) : "l"(&s->count));
            if ((unsigned)w != (unsigned)(w >> 32)) atomicAdd(&bad[2], 1ull);
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(s));
            if (a.z != a.w) atomicAdd(&bad[3], 1ull);
        }
    }
}
int main()
{
    Slot *tab; unsigned long long *bad, h[4];
    int nslots = 64;
    cudaMalloc(&tab, nslots * sizeof(Slot)); cudaMemset(tab, 0, nslots * sizeof(Slot));
    cudaMalloc(&bad, 32); cudaMemset(bad, 0, 32);
    k<<<148 * 8, 256>>>(tab, nslots, bad, 2000);
    cudaDeviceSynchronize();
    cudaMemcpy(h, bad, 32, cudaMemcpyDeviceToHost);
    printf("torn reads: v4.u32 %llu  v2.u64 %llu  scalar u64 %llu  cg.v4.u32 %llu   (%s)\n", h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
