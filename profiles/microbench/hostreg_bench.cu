/*
 * hostreg_bench.cu -- can the GPU's copy engines reach page-cache pages directly (no staging copy, no write())?
 *   input side : mmap of a tmpfs file (MAP_SHARED / MAP_PRIVATE, PROT_READ [+WRITE]) + cudaHostRegister, then H2D from it
 *   output side: ftruncate + mmap(MAP_SHARED, PROT_WRITE) of a fresh tmpfs file + cudaHostRegister, then D2H into it
 * Build: nvcc -O2 -arch=sm_100a hostreg_bench.cu -o hostreg_bench
 */
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : "/dev/shm";
    size_t bytes = (argc > 2 ? (size_t)atol(argv[2]) : 512) << 20;
    char name[256];
    snprintf(name, sizeof name, "%s/hostreg_src.bin", dir);
    int fd = open(name, O_RDWR | O_CREAT | O_TRUNC, 0644);
    char *fill = (char *)malloc(1 << 20);
    memset(fill, 'A', 1 << 20);
    for (size_t o = 0; o < bytes; o += 1 << 20)
        if (write(fd, fill, 1 << 20) != (1 << 20))
            perror("fill");
    char *dev = NULL;
    cudaMalloc((void **)&dev, bytes);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int ro = 0, cur = 0;
    cudaDeviceGetAttribute(&ro, cudaDevAttrHostRegisterReadOnlySupported, 0);
    cudaDeviceGetAttribute(&cur, cudaDevAttrHostRegisterSupported, 0);
    printf("host register supported %d, read-only supported %d, pageableMemoryAccess %d\n", cur, ro, prop.pageableMemoryAccess);
    struct
    {
        const char *what;
        int prot, flags;
        unsigned reg;
    } in[] = {{"input MAP_SHARED  R  default", PROT_READ, MAP_SHARED, cudaHostRegisterDefault},
              {"input MAP_SHARED  R  readonly", PROT_READ, MAP_SHARED, cudaHostRegisterReadOnly},
              {"input MAP_SHARED  RW default", PROT_READ | PROT_WRITE, MAP_SHARED, cudaHostRegisterDefault},
              {"input MAP_PRIVATE R  readonly", PROT_READ, MAP_PRIVATE, cudaHostRegisterReadOnly},
              {"input MAP_PRIVATE RW default", PROT_READ | PROT_WRITE, MAP_PRIVATE, cudaHostRegisterDefault}};
    for (unsigned i = 0; i < sizeof in / sizeof *in; i++)
    {
        char *m = (char *)mmap(NULL, bytes, in[i].prot, in[i].flags, fd, 0);
        double t0 = now();
        cudaError_t e = cudaHostRegister(m, bytes, in[i].reg);
        double t1 = now();
        printf("%-32s register %zu MB: %s, %.3f s (%.2f GB/s)\n", in[i].what, bytes >> 20, cudaGetErrorString(e), t1 - t0,
               bytes / (t1 - t0) / 1e9);
        if (e == cudaSuccess)
        {
            t0 = now();
            cudaMemcpyAsync(dev, m, bytes, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            t1 = now();
            printf("    H2D from it: %.2f GB/s\n", bytes / (t1 - t0) / 1e9);
            t0 = now();
            cudaHostUnregister(m);
            printf("    unregister %.3f s\n", now() - t0);
        }
        else
            cudaGetLastError();
        munmap(m, bytes);
    }
    /* output side */
    for (size_t mb = 16; mb <= 256; mb *= 4)
    {
        size_t n = mb << 20;
        snprintf(name, sizeof name, "%s/hostreg_out.bin", dir);
        int ofd = open(name, O_RDWR | O_CREAT | O_TRUNC, 0644);
        double t0 = now();
        if (ftruncate(ofd, (off_t)n))
            perror("ftruncate");
        char *m = (char *)mmap(NULL, n, PROT_READ | PROT_WRITE, MAP_SHARED, ofd, 0);
        double t1 = now();
        cudaError_t e = cudaHostRegister(m, n, cudaHostRegisterDefault);
        double t2 = now();
        printf("output MAP_SHARED RW %4zu MB: map %.4f s, register %s %.3f s (%.2f GB/s)\n", mb, t1 - t0, cudaGetErrorString(e),
               t2 - t1, n / (t2 - t1) / 1e9);
        if (e == cudaSuccess)
        {
            cudaMemsetAsync(dev, 'C', n, st);
            cudaStreamSynchronize(st);
            t1 = now();
            cudaMemcpyAsync(m, dev, n, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            t2 = now();
            printf("    D2H into it: %.2f GB/s\n", n / (t2 - t1) / 1e9);
            t1 = now();
            cudaHostUnregister(m);
            t2 = now();
            munmap(m, n);
            char c[4] = {0};
            if (pread(ofd, c, 1, (off_t)(n - 1)) != 1)
                perror("pread");
            printf("    unregister %.3f s, last byte of the file reads '%c'\n", t2 - t1, c[0]);
        }
        else
        {
            cudaGetLastError();
            munmap(m, n);
        }
        close(ofd);
        unlink(name);
    }
    snprintf(name, sizeof name, "%s/hostreg_src.bin", dir);
    close(fd);
    unlink(name);
    return 0;
}
