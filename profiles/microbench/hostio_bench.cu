/*
 * hostio_bench.cu -- what the GPU box's HOST side can do (round 2, end-to-end design input).
 *
 *   memcpy      page-cache (tmpfs mmap) -> pinned buffer, T threads
 *   nlcount     AVX2 newline count over the mmap, T threads
 *   write       write() of T private files on tmpfs, 4 MB per call, from a pinned buffer
 *   pwrite1     pwrite() by T threads into ONE tmpfs file at disjoint offsets
 *   mmapw1      T threads memcpy into a MAP_SHARED mapping of ONE fresh tmpfs file (ftruncate'd) at disjoint offsets
 *   register    cudaHostRegister(ReadOnly) of the tmpfs mmap (would allow DMA straight from the page cache)
 *   h2d / d2h   cudaMemcpyAsync pinned <-> device
 *   zc_read     kernel reading mapped pinned memory (16-byte loads); zc_write: kernel storing to it
 * Build: nvcc -O2 -arch=sm_100a -Xcompiler -mavx2,-pthread hostio_bench.cu -o hostio_bench
 */
#include <cuda_runtime.h>
#include <fcntl.h>
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct
{
    int id, nthreads, mode;
    const char *src;
    char *dst;
    size_t bytes;
    const char *dir;
    int fd;
    uint64_t result;
} job;

static uint64_t count_nl(const char *p, size_t n)
{
    const __m256i nl = _mm256_set1_epi8('\n');
    uint64_t c = 0;
    size_t i = 0;
    for (; i + 64 <= n; i += 64)
    {
        __m256i a = _mm256_loadu_si256((const __m256i *)(p + i)), b = _mm256_loadu_si256((const __m256i *)(p + i + 32));
        c += __builtin_popcount((unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, nl))) +
             __builtin_popcount((unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, nl)));
    }
    for (; i < n; i++)
        c += p[i] == '\n';
    return c;
}

static void *worker(void *a)
{
    job *j = (job *)a;
    size_t per = (j->bytes / j->nthreads) & ~(size_t)4095, lo = per * j->id, n = per;
    switch (j->mode)
    {
    case 0:
        memcpy(j->dst + lo, j->src + lo, n);
        break;
    case 1:
        j->result = count_nl(j->src + lo, n);
        break;
    case 2:
    {
        char name[256];
        snprintf(name, sizeof name, "%s/hostio_w%d.bin", j->dir, j->id);
        int fd = open(name, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        for (size_t o = 0; o < n; o += 4u << 20)
        {
            size_t m = n - o < (4u << 20) ? n - o : (4u << 20);
            if (write(fd, j->dst + lo + o, m) != (ssize_t)m)
                perror("write");
        }
        close(fd);
        unlink(name);
        break;
    }
    case 3:
        for (size_t o = 0; o < n; o += 4u << 20)
        {
            size_t m = n - o < (4u << 20) ? n - o : (4u << 20);
            if (pwrite(j->fd, j->dst + lo + o, m, (off_t)(lo + o)) != (ssize_t)m)
                perror("pwrite");
        }
        break;
    case 4:
        memcpy((char *)j->src + lo, j->dst + lo, n); /* src = the writable mapping of the output file here */
        break;
    }
    return NULL;
}

static double run(int mode, int T, const char *src, char *dst, size_t bytes, const char *dir, int fd)
{
    pthread_t th[256];
    job jobs[256];
    double t0 = now();
    for (int i = 0; i < T; i++)
    {
        jobs[i] = (job){i, T, mode, src, dst, bytes, dir, fd, 0};
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (int i = 0; i < T; i++)
        pthread_join(th[i], NULL);
    return now() - t0;
}

__global__ void k_zc_read(const uint4 *p, size_t n, unsigned *sink)
{
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        uint4 v = p[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345)
        *sink = acc;
}
__global__ void k_zc_write(uint4 *p, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4((unsigned)i, 1, 2, 3);
}

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "/dev/shm";
    size_t bytes = (argc > 2 ? (size_t)atol(argv[2]) : 2048) << 20;
    char name[256];
    snprintf(name, sizeof name, "%s/hostio_src.bin", dir);
    int fd = open(name, O_RDWR | O_CREAT | O_TRUNC, 0644);
    char *fill = (char *)malloc(1 << 20);
    for (int i = 0; i < (1 << 20); i++)
        fill[i] = (i % 83 == 82) ? '\n' : "ACGT"[(i * 7 + i / 13) & 3];
    for (size_t o = 0; o < bytes; o += 1 << 20)
        if (write(fd, fill, 1 << 20) != (1 << 20))
            perror("fill");
    const char *src = (const char *)mmap(NULL, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    printf("online cpus %ld, bytes %zu MB, dir %s\n", ncpu, bytes >> 20, dir);
    char *pinned = NULL;
    double t0 = now();
    if (cudaHostAlloc((void **)&pinned, bytes, cudaHostAllocMapped) != cudaSuccess)
    {
        printf("cudaHostAlloc failed\n");
        return 1;
    }
    printf("cudaHostAlloc %zu MB: %.3f s\n", bytes >> 20, now() - t0);
    memset(pinned, 1, bytes);
    int Ts[] = {1, 2, 4, 8, 16, 32, 64};
    for (int mode = 0; mode < 5; mode++)
    {
        const char *nm[] = {"memcpy mmap->pinned", "nlcount mmap", "write T files", "pwrite one file", "memcpy into mmap of 1 file"};
        int ofd = -1;
        char oname[256];
        snprintf(oname, sizeof oname, "%s/hostio_one.bin", dir);
        for (unsigned ti = 0; ti < sizeof Ts / sizeof *Ts; ti++)
        {
            int T = Ts[ti];
            if (T > 2 * ncpu)
                break;
            if (mode == 3)
                ofd = open(oname, O_WRONLY | O_CREAT | O_TRUNC, 0644);
            double best = 1e9;
            if (mode == 4)
            { /* what a writer would do per step: grow the file, map the new range, fill it on T threads, unmap */
                ofd = open(oname, O_RDWR | O_CREAT | O_TRUNC, 0644);
                double t0 = now();
                if (ftruncate(ofd, (off_t)bytes))
                    perror("ftruncate");
                char *m = (char *)mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, ofd, 0);
                run(4, T, m, pinned, bytes, dir, ofd);
                munmap(m, bytes);
                best = now() - t0;
                close(ofd);
                unlink(oname);
            }
            else
            for (int rep = 0; rep < (mode >= 2 ? 1 : 2); rep++)
            {
                double s = run(mode, T, src, pinned, bytes, dir, ofd);
                if (s < best)
                    best = s;
            }
            if (mode == 3)
            {
                close(ofd);
                unlink(oname);
            }
            printf("%-22s T=%2d  %.3f s  %.2f GB/s\n", nm[mode], T, best, bytes / best / 1e9);
        }
    }
    /* pinning the page cache mapping itself */
    for (size_t mb = 64; mb <= 1024 && (mb << 20) <= bytes; mb *= 4)
    {
        t0 = now();
        cudaError_t e = cudaHostRegister((void *)src, mb << 20, cudaHostRegisterReadOnly | cudaHostRegisterMapped);
        double s = now() - t0;
        printf("cudaHostRegister(ReadOnly) %4zu MB of the tmpfs mmap: %s, %.3f s (%.2f GB/s)\n", mb, cudaGetErrorString(e), s,
               (mb << 20) / s / 1e9);
        if (e == cudaSuccess)
        {
            t0 = now();
            cudaHostUnregister((void *)src);
            printf("  unregister %.3f s\n", now() - t0);
        }
        else
            cudaGetLastError();
    }
    /* PCIe */
    char *dev = NULL;
    cudaMalloc((void **)&dev, bytes);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int dir2 = 0; dir2 < 2; dir2++)
        for (int rep = 0; rep < 2; rep++)
        {
            cudaEventRecord(e0, st);
            if (dir2 == 0)
                cudaMemcpyAsync(dev, pinned, bytes, cudaMemcpyHostToDevice, st);
            else
                cudaMemcpyAsync(pinned, dev, bytes, cudaMemcpyDeviceToHost, st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s pinned %zu MB: %.2f ms  %.2f GB/s\n", dir2 ? "D2H" : "H2D", bytes >> 20, ms, bytes / ms / 1e6);
        }
    /* both directions at once */
    {
        cudaStream_t s2;
        cudaStreamCreate(&s2);
        char *dev2 = NULL;
        cudaMalloc((void **)&dev2, bytes / 2);
        cudaEventRecord(e0, st);
        cudaMemcpyAsync(dev, pinned, bytes / 2, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(pinned + bytes / 2, dev2, bytes / 2, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s2);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("H2D + D2H concurrently, %zu MB each: %.2f ms  %.2f GB/s each way\n", bytes >> 21, ms, bytes / 2 / ms / 1e6);
    }
    /* pageable H2D straight from the mmap */
    cudaEventRecord(e0, st);
    cudaMemcpyAsync(dev, src, bytes / 4, cudaMemcpyHostToDevice, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("H2D pageable (the mmap) %zu MB: %.2f ms  %.2f GB/s\n", bytes >> 22, ms, bytes / 4 / ms / 1e6);
    unsigned *sink;
    cudaMalloc((void **)&sink, 4);
    char *dpin = NULL;
    cudaHostGetDevicePointer((void **)&dpin, pinned, 0);
    for (int blocks = 148; blocks <= 148 * 8; blocks *= 8)
    {
        cudaEventRecord(e0, st);
        k_zc_read<<<blocks, 256, 0, st>>>((const uint4 *)dpin, bytes / 16, sink);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("zero-copy read  %4d CTAs: %.2f ms  %.2f GB/s\n", blocks, ms, bytes / ms / 1e6);
        cudaEventRecord(e0, st);
        k_zc_write<<<blocks, 256, 0, st>>>((uint4 *)dpin, bytes / 16);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("zero-copy write %4d CTAs: %.2f ms  %.2f GB/s\n", blocks, ms, bytes / ms / 1e6);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    munmap((void *)src, bytes);
    close(fd);
    unlink(name);
    return 0;
}
