/*
 * hostbench_main.c -- DEVELOPMENT TOOL ONLY: drives the library entry points the way bench.py does
 * (inputs held in memory, outputs under /dev/shm) against the null engine, and prints the host
 * pipeline's stage times.  usage: nk_hostlib fwd.fastq rev.fastq [partitions] [repeats]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/nk_b200.h"

/* optional sampling profile (NK_HOSTBENCH_PROF=file): instruction pointers at 1 kHz of process CPU time, to be
 * resolved with addr2line on a -no-pie / -static build */
#define _GNU_SOURCE 1
#include <signal.h>
#include <sys/time.h>
#include <ucontext.h>
#define NK_MAX_SAMPLES (1 << 20)
static unsigned long long *nk_samples;
static volatile long nk_n_samples;
static void nk_on_prof(int sig, siginfo_t *si, void *uc)
{
    (void)sig;
    (void)si;
    long i = __atomic_fetch_add(&nk_n_samples, 1, __ATOMIC_RELAXED);
    if (i < NK_MAX_SAMPLES)
        nk_samples[i] = (unsigned long long)((ucontext_t *)uc)->uc_mcontext.gregs[16]; /* REG_RIP */
}
static void nk_prof_start(void)
{
    nk_samples = calloc(NK_MAX_SAMPLES, sizeof *nk_samples);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = nk_on_prof;
    sa.sa_flags = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, NULL);
    struct itimerval tv = {{0, 1000}, {0, 1000}};
    setitimer(ITIMER_PROF, &tv, NULL);
}
static void nk_prof_stop(const char *path)
{
    struct itimerval tv = {{0, 0}, {0, 0}};
    setitimer(ITIMER_PROF, &tv, NULL);
    FILE *o = fopen(path, "w");
    long n = nk_n_samples < NK_MAX_SAMPLES ? nk_n_samples : NK_MAX_SAMPLES;
    for (long i = 0; i < n; i++)
        fprintf(o, "%llx\n", nk_samples[i]);
    fclose(o);
    /* the address map, so that samples inside shared libraries can be attributed to their module */
    char mp[512];
    snprintf(mp, sizeof mp, "%s.maps", path);
    FILE *mi = fopen("/proc/self/maps", "r"), *mo = fopen(mp, "w");
    if (mi && mo)
        for (int ch; (ch = fgetc(mi)) != EOF;)
            fputc(ch, mo);
    if (mi)
        fclose(mi);
    if (mo)
        fclose(mo);
}

static char *slurp(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    char *b = malloc(*n + 1);
    if (fread(b, 1, *n, f) != *n) { perror("read"); exit(1); }
    b[*n] = 0;
    fclose(f);
    return b;
}
static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s fwd rev [partitions] [repeats]\n", argv[0]); return 1; }
    size_t nf, nr;
    char *f = slurp(argv[1], &nf), *r = slurp(argv[2], &nr);
    int parts = argc > 3 ? atoi(argv[3]) : 8, reps = argc > 4 ? atoi(argv[4]) : 3;
    for (int it = 0; it < reps; it++)
    {
        nk_config c;
        memset(&c, 0, sizeof c);
        c.k = 25; c.depth = 100; c.coverage = 0.9f; c.canonical = 1; c.in_fastq = c.out_fastq = 1;
        c.partitions = parts; c.n_forward_files = 1; c.have_reverse = 1; c.out_dir = "/dev/shm/nk_hostbench"; c.n_devices = 1;
        c.memory_gb = getenv("NK_HB_MEMORY") ? atoi(getenv("NK_HB_MEMORY")) : 0;
        if (system("rm -rf /dev/shm/nk_hostbench && mkdir -p /dev/shm/nk_hostbench")) return 1;
        nk_ctx *x;
        if (nk_create(&c, &x)) { fprintf(stderr, "%s\n", nk_create_error()); return 1; }
        double t0 = now();
        nk_seed_buffer(x, f, nf, 3000001);
        nk_seed_buffer(x, r, nr, 3000001);
        nk_seed_finish(x);
        double t1 = now();
        if (getenv("NK_HOSTBENCH_PROF") && it == reps - 1)
            nk_prof_start();
        int rc = nk_process_paired(x, f, nf, r, nr);
        if (getenv("NK_HOSTBENCH_PROF") && it == reps - 1)
            nk_prof_stop(getenv("NK_HOSTBENCH_PROF"));
        double t2 = now();
        nk_totals t;
        nk_totals_get(x, &t);
        nk_finish(x);
        printf("rc %d seed %.3f process %.3f (index %.3f device %.3f write %.3f) processed %llu printed %llu\n", rc, t1 - t0,
               t2 - t1, t.index_seconds, t.device_seconds, t.write_seconds, (unsigned long long)t.processed,
               (unsigned long long)t.printed);
        nk_destroy(x);
    }
    if (system("rm -rf /dev/shm/nk_hostbench")) return 1;
    return 0;
}
