/*
 * nk_host.c -- C host side of the B200 k-mer normalisation path (nk_* in include/nk_b200.h).
 *
 * What the reference does per file in main() and its worker threads (normalise_kmers_multi_large.c,
 * "C:n") is restated here around the device engine:
 *   partition byte ranges      calculate_thread_positions[_from_records], count_records_seqfile  C:1240-1320
 *   record indexer             the read_line x4/x2 loop of process_thread_chunk_*                C:394-409, C:1605-1631
 *   length gate                is_valid_sequence_* (N->A and the alphabet gate run on the GPU)    C:1404-1457
 *   staging                    sequence lines only, 16-byte aligned, into page-locked buffers
 *   writer                     accepted records, N->A in the sequence line, fq->fa rewrite        C:1649-1666, C:852-876
 *   seeding                    seed_kmer_hash's own line splitter and "> K" rule                  C:1322-1373
 *   CLI                        parse_arguments / main: same flags, files, stdout lines, exit codes C:520-745, C:2223-2455
 * One pipeline thread per GPU; partitions are independent, so no collective is needed (README:68).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <getopt.h>
#include <locale.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/nk_b200.h"

/* ------------------------------------------------------------------ small helpers */

typedef struct
{
    const char *data;
    size_t size;
} nk_buf;

static inline char nk_at(const nk_buf *f, size_t i) { return i < f->size ? f->data[i] : '\0'; } /* past EOF reads as NUL */

static double nk_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef void (*nk_task_fn)(int index, void *arg);

/* Persistent worker pool shared by every context of the process: the pipeline stages hand it index ranges many
 * times per step, and creating / joining threads for each of them (round 1) cost more than some of the stages.
 * A caller works on its own job too, so nested and concurrent jobs always make progress. */
typedef struct nk_job
{
    nk_task_fn fn;
    void *arg;
    int n, limit;  /* tasks; most workers that may serve it besides the caller */
    int next, done, helpers;
    struct nk_job *link;
} nk_job;

static struct
{
    pthread_mutex_t mu;
    pthread_cond_t work, finished;
    nk_job *jobs;
    int n_threads, started;
} nk_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0};

static int nk_host_threads(void);

static void nk_job_work(nk_job *j)
{
    for (;;)
    {
        int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n)
            break;
        j->fn(i, j->arg);
        if (__atomic_add_fetch(&j->done, 1, __ATOMIC_ACQ_REL) == j->n)
        {
            pthread_mutex_lock(&nk_pool.mu);
            pthread_cond_broadcast(&nk_pool.finished);
            pthread_mutex_unlock(&nk_pool.mu);
        }
    }
}

static void *nk_pool_worker(void *unused)
{
    (void)unused;
    pthread_mutex_lock(&nk_pool.mu);
    for (;;)
    {
        nk_job *j = nk_pool.jobs;
        while (j && (__atomic_load_n(&j->next, __ATOMIC_RELAXED) >= j->n || j->helpers >= j->limit))
            j = j->link;
        if (!j)
        {
            pthread_cond_wait(&nk_pool.work, &nk_pool.mu);
            continue;
        }
        j->helpers++;
        pthread_mutex_unlock(&nk_pool.mu);
        nk_job_work(j);
        pthread_mutex_lock(&nk_pool.mu);
        j->helpers--;
        pthread_cond_broadcast(&nk_pool.finished); /* the owner may be waiting for its helpers to let go */
    }
    return NULL;
}

static void nk_pool_start(void)
{
    pthread_mutex_lock(&nk_pool.mu);
    if (!nk_pool.started)
    {
        nk_pool.started = 1;
        int want = nk_host_threads();
        if (want > 256)
            want = 256;
        for (int t = 0; t < want; t++)
        {
            pthread_t th;
            if (pthread_create(&th, NULL, nk_pool_worker, NULL) == 0)
            {
                pthread_detach(th);
                nk_pool.n_threads++;
            }
        }
    }
    pthread_mutex_unlock(&nk_pool.mu);
}

/* run fn(0..n-1) on up to nthreads threads (the caller is one of them) */
static void nk_parallel_for(int n, int nthreads, nk_task_fn fn, void *arg)
{
    if (nthreads > n)
        nthreads = n;
    if (nthreads <= 1)
    {
        for (int i = 0; i < n; i++)
            fn(i, arg);
        return;
    }
    nk_pool_start();
    nk_job job = {fn, arg, n, nthreads - 1, 0, 0, 0, NULL};
    pthread_mutex_lock(&nk_pool.mu);
    job.link = nk_pool.jobs;
    nk_pool.jobs = &job;
    pthread_cond_broadcast(&nk_pool.work);
    pthread_mutex_unlock(&nk_pool.mu);
    nk_job_work(&job);
    pthread_mutex_lock(&nk_pool.mu);
    while (__atomic_load_n(&job.done, __ATOMIC_ACQUIRE) < n || job.helpers > 0)
        pthread_cond_wait(&nk_pool.finished, &nk_pool.mu);
    for (nk_job **pp = &nk_pool.jobs; *pp; pp = &(*pp)->link)
        if (*pp == &job)
        {
            *pp = job.link;
            break;
        }
    pthread_mutex_unlock(&nk_pool.mu);
}

/* NKB200_TRACE=file: one line per pipeline stage and step (engine, step, stage, start, end in seconds), for the
 * timeline plots under profiles/ */
static struct
{
    int on;
    FILE *f;
    pthread_mutex_t mu;
    double t0;
} nk_tr = {-1, NULL, PTHREAD_MUTEX_INITIALIZER, 0};

static void nk_trace(int engine, int step, const char *stage, double a, double b)
{
    if (nk_tr.on == 0)
        return;
    pthread_mutex_lock(&nk_tr.mu);
    if (nk_tr.on < 0)
    {
        const char *path = getenv("NKB200_TRACE");
        char name[1024];
        if (path && *path)
            snprintf(name, sizeof name, "%s.%d", path, (int)getpid()); /* one file per process (rank) */
        nk_tr.f = path && *path ? fopen(name, "a") : NULL;
        nk_tr.on = nk_tr.f != NULL;
    }
    if (nk_tr.on)
    {
        fprintf(nk_tr.f, "%d %d %s %.6f %.6f\n", engine, step, stage, a, b);
        fflush(nk_tr.f);
    }
    pthread_mutex_unlock(&nk_tr.mu);
}

/* Page-locking memory costs ~0.4 s per GB on these hosts, and a context's step buffers are GBs: the large blocks of a
 * destroyed context are kept for the next one of the same shape in this process (bench.py and any caller that runs
 * several files' worth of contexts).  At most NK_PIN_CACHE blocks are held; NKB200_NO_PIN_CACHE=1 turns it off. */
#define NK_PIN_CACHE 48
static struct
{
    pthread_mutex_t mu;
    void *ptr[NK_PIN_CACHE];
    size_t size[NK_PIN_CACHE];
} nk_pins = {PTHREAD_MUTEX_INITIALIZER, {0}, {0}};

static void *nk_pinned_get(size_t bytes)
{
    void *p = NULL;
    pthread_mutex_lock(&nk_pins.mu);
    for (int i = 0; i < NK_PIN_CACHE && !p; i++)
        if (nk_pins.ptr[i] && nk_pins.size[i] == bytes)
        {
            p = nk_pins.ptr[i];
            nk_pins.ptr[i] = NULL;
        }
    pthread_mutex_unlock(&nk_pins.mu);
    return p ? p : nkd_alloc_pinned(bytes);
}

static void nk_pinned_put(void *p, size_t bytes)
{
    if (!p)
        return;
    const char *off = getenv("NKB200_NO_PIN_CACHE");
    if (bytes >= ((size_t)8 << 20) && !(off && *off && strcmp(off, "0") != 0))
    {
        pthread_mutex_lock(&nk_pins.mu);
        for (int i = 0; i < NK_PIN_CACHE; i++)
            if (!nk_pins.ptr[i])
            {
                nk_pins.ptr[i] = p;
                nk_pins.size[i] = bytes;
                p = NULL;
                break;
            }
        pthread_mutex_unlock(&nk_pins.mu);
    }
    if (p)
        nkd_free_pinned(p);
}

/* boolean environment switches: unset, empty and "0" mean off */
static int nk_env_on(const char *name)
{
    const char *s = getenv(name);
    return s && *s && strcmp(s, "0") != 0;
}

static int nk_host_threads(void)
{
    const char *s = getenv("NKB200_THREADS");
    if (s && atoi(s) > 0)
        return atoi(s);
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------ capacity (C:416-422, C:676-684) */

static uint64_t nk_pow4_wrapping(int k)
{
    uint64_t v = 1;
    for (int i = 0; i < k && i < 64; i++)
        v *= 4;
    return v;
}

static uint64_t nk_capacity_unclamped(int memory_gb, int partitions)
{
    if (memory_gb <= 0)
        return NK_DEFAULT_SLOTS;
    /* float32 on purpose: the capacity decides which k-mers collide, so it is part of the results */
    size_t bytes = (size_t)((float)memory_gb * 1073741824);
    float total = (float)(bytes / 16);
    size_t per = (size_t)(total / (float)partitions);
    return (per % 2 == 0) ? per + 1 : per;
}

uint64_t nk_initial_capacity(int memory_gb, int partitions, int k)
{
    uint64_t cap = nk_capacity_unclamped(memory_gb, partitions > 0 ? partitions : 1);
    uint64_t lim = nk_pow4_wrapping(k);
    return lim < cap ? lim : cap;
}

/* ------------------------------------------------------------------ partition byte ranges */

typedef struct
{
    char msg[256];
    int failed;
} nk_diag;

/* find_thread_exact_end, C:1199-1236 */
static int nk_boundary_before(const nk_buf *f, size_t lo, size_t hi, int fastq, size_t *out, nk_diag *d)
{
    if (!fastq)
    {
        for (size_t i = hi; i > lo; i--)
            if (nk_at(f, i) == '>')
            {
                *out = i - 1;
                return 0;
            }
    }
    else
    {
        int nl = 0, plus = 0;
        for (size_t i = hi; i > lo; i--)
        {
            if (nk_at(f, i) != '\n')
                continue;
            nl++;
            if (nk_at(f, i + 1) == '+')
                plus = 1;
            else if (plus && nk_at(f, i + 1) == '@')
            {
                *out = i;
                return 0;
            }
            if (nl == 7)
            {
                snprintf(d->msg, sizeof d->msg, "ERROR: after 7 lines, I couldn't find the + and @ headers near this chunk %'zu", i);
                d->failed = 1;
                return -1;
            }
        }
    }
    snprintf(d->msg, sizeof d->msg, "ERROR: i couldn't find the start of sequence before this chunk end %'zu", hi);
    d->failed = 1;
    return -1;
}

/* calculate_thread_positions, C:1240-1262: starts[1] is never assigned and the last end is overwritten,
 * so partition 1 re-reads partition 0's bytes and the file tail is dropped */
static int nk_ranges_by_size(const nk_buf *f, int p, int fastq, uint64_t *st, uint64_t *en, nk_diag *d)
{
    size_t chunk = f->size / (size_t)p;
    if (chunk <= (size_t)NK_MAX_LINE * 4)
    {
        snprintf(d->msg, sizeof d->msg, "Error: input too small to split by size across %d partitions", p);
        d->failed = 1;
        return -1;
    }
    size_t approx = chunk - (size_t)NK_MAX_LINE * 4, e;
    st[0] = 0;
    if (nk_boundary_before(f, 0, approx, fastq, &e, d))
        return -1;
    en[0] = e;
    en[p - 1] = f->size - 1;
    for (int t = 1; t < p; t++)
    {
        size_t s = en[t - 1] + 1;
        if (nk_boundary_before(f, s, s + approx, fastq, &e, d))
            return -1;
        en[t] = e;
        if (t < p - 1)
            st[t + 1] = en[t] + 1;
    }
    return 0;
}

/* ---- byte scanning: 64 bytes at a time, AVX2 when the CPU has it, SSE2 otherwise ---- */
#include <immintrin.h>

/* bit i of *nl / *nul is set when p[i] is '\n' / NUL, for i < 64 */
typedef void (*nk_mask64_fn)(const char *p, uint64_t *nl, uint64_t *nul);

__attribute__((target("avx2"))) static void nk_mask64_avx2(const char *p, uint64_t *nl, uint64_t *nul)
{
    const __m256i a = _mm256_loadu_si256((const __m256i *)p), b = _mm256_loadu_si256((const __m256i *)(p + 32));
    const __m256i n = _mm256_set1_epi8('\n'), z = _mm256_setzero_si256();
    *nl = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, n)) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, n)) << 32);
    *nul = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z)) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z)) << 32);
}

static void nk_mask64_sse2(const char *p, uint64_t *nl, uint64_t *nul)
{
    const __m128i n = _mm_set1_epi8('\n'), z = _mm_setzero_si128();
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 4; i++)
    {
        __m128i v = _mm_loadu_si128((const __m128i *)(p + 16 * i));
        a |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, n)) << (16 * i);
        b |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, z)) << (16 * i);
    }
    *nl = a;
    *nul = b;
}

static nk_mask64_fn nk_mask64_pick(void)
{
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? nk_mask64_avx2 : nk_mask64_sse2;
}
static nk_mask64_fn nk_mask64 = NULL;

static inline void nk_mask64_tail(const char *p, size_t n, uint64_t *nl, uint64_t *nul)
{ /* fewer than 64 bytes left: never touch memory past the mapping */
    char tmp[64];
    memset(tmp, 1, sizeof tmp);
    memcpy(tmp, p, n);
    nk_mask64(tmp, nl, nul);
    if (n < 64)
    {
        *nl &= ((uint64_t)1 << n) - 1;
        *nul &= ((uint64_t)1 << n) - 1;
    }
}

/* number of '\n' in [p, p+n) */
static uint64_t nk_count_newlines(const char *p, size_t n)
{
    uint64_t c = 0, nl, nul;
    size_t i = 0;
    for (; i + 64 <= n; i += 64)
    {
        nk_mask64(p + i, &nl, &nul);
        c += (uint64_t)__builtin_popcountll(nl);
    }
    if (i < n)
    {
        nk_mask64_tail(p + i, n - i, &nl, &nul);
        c += (uint64_t)__builtin_popcountll(nl);
    }
    return c;
}

/* offset of the k-th (1-based) '\n' in [p, p+n), or n when there are fewer */
static size_t nk_kth_newline(const char *p, size_t n, uint64_t k)
{
    uint64_t nl, nul;
    size_t i = 0;
    for (; i < n; i += 64)
    {
        if (i + 64 <= n)
            nk_mask64(p + i, &nl, &nul);
        else
            nk_mask64_tail(p + i, n - i, &nl, &nul);
        uint64_t c = (uint64_t)__builtin_popcountll(nl);
        if (c < k)
        {
            k -= c;
            continue;
        }
        while (--k)
            nl &= nl - 1;
        return i + (size_t)__builtin_ctzll(nl);
    }
    return n;
}

/* streaming newline iterator over [pos, end) of a buffer */
typedef struct
{
    const char *data;
    size_t end;     /* scan limit (file size) */
    size_t blk;     /* offset of the 64-byte block the masks describe */
    uint64_t nl;    /* unread newline bits of that block */
    int nul_seen;   /* a NUL byte was seen in a block touched since the last reset */
} nk_nliter;

static inline void nk_nliter_seek(nk_nliter *it, const char *data, size_t pos, size_t end)
{
    it->data = data;
    it->end = end;
    it->blk = pos & ~(size_t)63;
    it->nul_seen = 0;
    it->nl = 0;
    if (it->blk < end)
    {
        uint64_t nul;
        if (it->blk + 64 <= end)
            nk_mask64(data + it->blk, &it->nl, &nul);
        else
            nk_mask64_tail(data + it->blk, end - it->blk, &it->nl, &nul);
        uint64_t keep = ~(uint64_t)0 << (pos & 63);
        it->nl &= keep;
        it->nul_seen = (nul & keep) != 0;
    }
}

/* offset of the next '\n', or SIZE_MAX at the end of the buffer */
static inline size_t nk_nliter_next(nk_nliter *it)
{
    for (;;)
    {
        if (it->nl)
        {
            size_t r = it->blk + (size_t)__builtin_ctzll(it->nl);
            it->nl &= it->nl - 1;
            return r;
        }
        it->blk += 64;
        if (it->blk >= it->end)
            return SIZE_MAX;
        uint64_t nul;
        if (it->blk + 64 <= it->end)
            nk_mask64(it->data + it->blk, &it->nl, &nul);
        else
            nk_mask64_tail(it->data + it->blk, it->end - it->blk, &it->nl, &nul);
        it->nul_seen |= nul != 0;
    }
}

/* per-file line index: newline counts of fixed chunks, built on all host cores.  The partitioner looks up
 * "the n-th line end after this offset" in it (C:1265-1300), and so does the raw-text pipeline when it cuts a
 * partition's byte range into steps of whole records. */
#ifndef NK_LI_CHUNK_BYTES
#define NK_LI_CHUNK_BYTES (256 << 10) /* the emulation build of the tests uses 4 KB: many chunks in small files */
#endif
#define NK_LI_CHUNK ((size_t)NK_LI_CHUNK_BYTES)
typedef struct
{
    const nk_buf *f;
    int nchunks;       /* chunks of the whole file */
    int c_lo, c_hi;    /* chunks that were counted: offsets outside [c_lo, c_hi) * NK_LI_CHUNK cannot be looked up */
    uint64_t *cum;     /* cum[i] = newlines in chunks [c_lo, i) for c_lo <= i <= c_hi */
} nk_lineidx;

typedef struct
{
    nk_lineidx *li;
    int base; /* first chunk of this call */
} nk_lineidx_span;

static void nk_lineidx_task(int i, void *a)
{
    nk_lineidx_span *sp = a;
    nk_lineidx *li = sp->li;
    int c = sp->base + i;
    size_t lo = (size_t)c * NK_LI_CHUNK, hi = lo + NK_LI_CHUNK;
    if (hi > li->f->size)
        hi = li->f->size;
    li->cum[c + 1] = nk_count_newlines(li->f->data + lo, hi - lo);
}

/* an index that starts at the chunk holding byte_lo and covers nothing yet */
static void nk_lineidx_open(nk_lineidx *li, const nk_buf *f, size_t byte_lo)
{
    if (!nk_mask64)
        nk_mask64 = nk_mask64_pick();
    li->f = f;
    li->nchunks = (int)((f->size + NK_LI_CHUNK - 1) / NK_LI_CHUNK);
    li->c_lo = (int)(byte_lo / NK_LI_CHUNK);
    if (li->c_lo > li->nchunks)
        li->c_lo = li->nchunks;
    li->c_hi = li->c_lo;
    li->cum = calloc((size_t)li->nchunks + 2, sizeof(uint64_t)); /* pages of chunks never counted stay untouched */
}

/* counts the chunks between the covered ones and byte_hi (one owner at a time: readers are the owner) */
static void nk_lineidx_extend(nk_lineidx *li, int threads, size_t byte_hi)
{
    if (byte_hi > li->f->size)
        byte_hi = li->f->size;
    int hi = (int)((byte_hi + NK_LI_CHUNK - 1) / NK_LI_CHUNK);
    if (hi > li->nchunks)
        hi = li->nchunks;
    if (hi <= li->c_hi || !li->cum)
        return;
    nk_lineidx_span sp = {li, li->c_hi};
    nk_parallel_for(hi - li->c_hi, threads, nk_lineidx_task, &sp);
    for (int i = li->c_hi; i < hi; i++)
        li->cum[i + 1] += li->cum[i];
    /* published last: another thread may look things up in the part that is already counted while this one goes on
     * (the reverse file's index while the first partitions are being worked on, nk_roll_thread) */
    __atomic_store_n(&li->c_hi, hi, __ATOMIC_RELEASE);
}

static inline int nk_lineidx_hi(const nk_lineidx *li) { return __atomic_load_n(&li->c_hi, __ATOMIC_ACQUIRE); }

/* counts of the chunks that overlap [byte_lo, byte_hi) */
static void nk_lineidx_build_range(nk_lineidx *li, const nk_buf *f, int threads, size_t byte_lo, size_t byte_hi)
{
    nk_lineidx_open(li, f, byte_lo);
    nk_lineidx_extend(li, threads, byte_hi);
}
static void nk_lineidx_build_n(nk_lineidx *li, const nk_buf *f, int threads) { nk_lineidx_build_range(li, f, threads, 0, f->size); }
static void nk_lineidx_build(nk_lineidx *li, const nk_buf *f) { nk_lineidx_build_n(li, f, nk_host_threads()); }

/* the same index from per-chunk counts somebody else made (nk_count_chunk_lines: ranks of a multi-process
 * launch count a share of the chunks each and exchange the counts) */
static void nk_lineidx_from_counts(nk_lineidx *li, const nk_buf *f, const uint32_t *counts)
{
    li->f = f;
    li->nchunks = (int)((f->size + NK_LI_CHUNK - 1) / NK_LI_CHUNK);
    li->c_lo = 0;
    li->c_hi = li->nchunks;
    li->cum = calloc((size_t)li->nchunks + 2, sizeof(uint64_t));
    for (int i = 0; i < li->nchunks; i++)
        li->cum[i + 1] = li->cum[i] + counts[i];
}

static void nk_lineidx_free(nk_lineidx *li)
{
    free(li->cum);
    li->cum = NULL;
}

static uint64_t nk_lineidx_total(const nk_lineidx *li) { return li->cum ? li->cum[nk_lineidx_hi(li)] : 0; }

/* offset of the newline with index g (0-based, counted from the first counted chunk), or SIZE_MAX */
static size_t nk_lineidx_find(const nk_lineidx *li, uint64_t g)
{
    const int c_hi = nk_lineidx_hi(li);
    if (c_hi == li->c_lo || g >= li->cum[c_hi])
        return SIZE_MAX;
    int lo = li->c_lo, hi = c_hi - 1; /* last chunk whose cum <= g */
    while (lo < hi)
    {
        int mid = (lo + hi + 1) / 2;
        if (li->cum[mid] <= g)
            lo = mid;
        else
            hi = mid - 1;
    }
    size_t base = (size_t)lo * NK_LI_CHUNK, n = NK_LI_CHUNK;
    if (base + n > li->f->size)
        n = li->f->size - base;
    return base + nk_kth_newline(li->f->data + base, n, g - li->cum[lo] + 1);
}

/* newlines in [first counted chunk, pos) */
static uint64_t nk_lineidx_before(const nk_lineidx *li, size_t pos)
{
    const int c_hi = nk_lineidx_hi(li);
    if (c_hi == li->c_lo)
        return 0;
    if (pos > li->f->size)
        pos = li->f->size;
    int c = (int)(pos / NK_LI_CHUNK);
    if (c < li->c_lo)
        return 0;
    if (c >= c_hi)
        return li->cum[c_hi];
    return li->cum[c] + nk_count_newlines(li->f->data + (size_t)c * NK_LI_CHUNK, pos - (size_t)c * NK_LI_CHUNK);
}

static uint64_t nk_records_from_lines(const nk_buf *f, uint64_t lines, int fastq)
{ /* count_records_seqfile, C:1302-1320 */
    if (f->size > 0 && f->data[f->size - 1] != '\n')
        lines++;
    return fastq ? lines / 4 : lines / 2;
}

uint64_t nk_count_records(const char *data, size_t size, int fastq)
{
    nk_buf f = {data, size};
    nk_lineidx li;
    nk_lineidx_build(&li, &f);
    uint64_t r = nk_records_from_lines(&f, nk_lineidx_total(&li), fastq);
    nk_lineidx_free(&li);
    return r;
}

/* calculate_thread_positions_from_records, C:1265-1300: partition t ends at the `want`-th newline counted
 * from its own start; found through the line index instead of a sequential scan of the file */
static void nk_ranges_by_records(const nk_lineidx *li, int p, int fastq, uint64_t records, uint64_t *st, uint64_t *en)
{
    const nk_buf *f = li->f;
    uint64_t per = records / (uint64_t)p;
    if (p < 2 || per < 1 || f->size < 1)
        return;
    int want = (int)(fastq ? per * 4 : per * 2);
    st[0] = 0;
    en[p - 1] = f->size - 1;
    for (int t = 0; t < p - 1; t++)
    {
        uint64_t before = nk_lineidx_before(li, st[t]);
        size_t pos = want > 0 ? nk_lineidx_find(li, before + (uint64_t)want - 1) : SIZE_MAX;
        if (pos != SIZE_MAX)
        {
            en[t] = pos;
            st[t + 1] = pos + 1;
        }
    }
}

int nk_partition_ranges(const char *data, size_t size, int partitions, int fastq, int mode, uint64_t records,
                        uint64_t *starts, uint64_t *ends)
{
    nk_buf f = {data, size};
    nk_diag d = {{0}, 0};
    memset(starts, 0, sizeof(uint64_t) * (size_t)partitions);
    memset(ends, 0, sizeof(uint64_t) * (size_t)partitions);
    if (partitions == 1)
    {
        ends[0] = size - 1;
        return NK_OK;
    }
    if (mode == 0)
        return nk_ranges_by_size(&f, partitions, fastq, starts, ends, &d) ? NK_EDATA : NK_OK;
    nk_lineidx li;
    nk_lineidx_build(&li, &f);
    nk_ranges_by_records(&li, partitions, fastq, records, starts, ends);
    nk_lineidx_free(&li);
    return NK_OK;
}

/* ------------------------------------------------------------------ line reader (C:394-409) */

/* One line as read_line() sees it: ends at '\n', at a NUL (EOF included) or after 1023 chars.
 * *more = 0 when the byte after the consumed span is NUL (read_line returns NULL).
 * *plain = 0 unless the line is simply "text\n" (the writer's verbatim fast path). */
static inline size_t nk_take_line(const nk_buf *f, size_t pos, uint32_t *len, int *more, int *plain)
{
    size_t avail = pos < f->size ? f->size - pos : 0;
    size_t lim = avail < (size_t)(NK_MAX_LINE - 1) ? avail : (size_t)(NK_MAX_LINE - 1);
    const char *p = f->data + pos;
    const char *nl = lim ? memchr(p, '\n', lim) : NULL;
    size_t n = nl ? (size_t)(nl - p) : lim;
    const char *z = n ? memchr(p, 0, n) : NULL;
    if (z)
    {
        *plain = 0;
        *len = (uint32_t)(z - p);
        *more = 0;
        return pos + (size_t)(z - p);
    }
    pos += n;
    if (nl)
        pos++;
    else
    {
        *plain = 0;
        if (nk_at(f, pos) == '\n')
            pos++;
    }
    *len = (uint32_t)n;
    *more = nk_at(f, pos) != '\0';
    return pos;
}

/* ------------------------------------------------------------------ context */

typedef struct
{
    uint64_t start;    /* file offset of the record's first byte */
    uint32_t nbytes;   /* bytes consumed by its lines */
    uint16_t seq_rel;  /* offset of the sequence line */
    uint16_t seq_len;
    uint32_t plain;    /* bit 0: all lines are "text\n", the record can be copied verbatim; bits 8..: lines that were read */
} nk_span;

typedef struct
{
    size_t fp, fe, rp, re;
    int done;
    nk_nliter itf, itr; /* newline iterators positioned at fp / rp */
} nk_cursor;

typedef struct
{
    nkd_read *reads; /* page-locked */
    size_t n_reads, n_records;
    size_t seq_lo, seq_hi, seq_end; /* region of the step-wide sequence buffer */
    nk_span *spans;                 /* stride records (fwd, rev) */
    uint32_t ops;
    int64_t fatal_record; /* length-gate passed but the engine reported a non-DNA byte */
    /* raw-text steps: the file ranges [f0, f1) / [r0, r1) of the step's raw_n records */
    size_t f0, f1, r0, r1;
    uint32_t raw_n;
} nk_pstep;

typedef struct
{
    uint8_t *dst;
    const char *src;
    size_t n;
} nk_copy;

typedef struct
{
    uint8_t *seq;    /* page-locked, n_parts regions */
    uint8_t *accept; /* page-locked */
    nk_pstep *ps;    /* per device-local partition */
    nkd_segment *segs;
    size_t n_records;
    /* raw-text steps (page-locked): record text in, accepted records' text out */
    uint8_t *raw, *out;
    size_t raw_cap, out_cap; /* allocated sizes (the blocks go back to the process-wide pinned cache) */
    size_t raw_bytes;
    nkd_raw_segment *rsegs;
    nkd_raw_result *rres;
    int *rseg_li; /* engine-local partition of each raw segment */
    int n_rsegs;
    nk_copy *copies;
    int n_copies;
} nk_stepbuf;

typedef struct
{
    int gid;       /* global partition id (file names, C:2286) */
    int dev, lidx; /* device slot and engine-local index */
    FILE *out_f, *out_r;
    char *wbuf_f, *wbuf_r;
    nk_cursor cur;
    uint64_t processed, printed, skipped;
    double t_start;
    uint64_t last_processed;
    /* raw-text pipeline: whole records [raw_next, raw_total) of the partition's ranges are still to be staged;
     * line_f / line_r = index of the partition's first line end in the files' line indexes; the GPU stage moves
     * commit_* past every step it has finished, which is where the host parser takes over if it has to */
    uint64_t raw_total, raw_next, line_f, line_r;
    size_t raw_fp, raw_rp, commit_fp, commit_rp;
    /* the line indexes the builder looks records up in: the context's (whole files, counted before the pipelines
     * start: raw_known) or the partition's own, which its engine's builder extends a step ahead of itself so that
     * the GPU does not wait for a count of the whole input.  raw_total is "plenty" until the range's end is covered. */
    nk_lineidx *lf, *lr;
    nk_lineidx own_f, own_r;
    int raw_known;
    int ready; /* its byte ranges are known (under nk_ctx.roll_mu while the reverse file is still being counted) */
    int active; /* part of the wave that is being processed (all partitions, unless the tables do not fit the GPU) */
} nk_part;

#define NK_NBUF 3 /* index step i+1, run step i on the GPU and write step i-1 at the same time */

typedef struct
{
    int ordinal;
    int lead; /* index of the first engine on the same GPU: it alone builds the seed table */
    pthread_mutex_t step_lock; /* of the lead engine: serialises a GPU's device steps when serial_steps is set */
    nkd_engine *eng;
    int n_parts;
    int *parts; /* indices into ctx->part */
    nk_stepbuf sb[NK_NBUF];
    int have_parsed_bufs, have_raw_bufs; /* staging is allocated when a pipeline of that kind first runs */
    uint64_t table_budget;               /* bytes of HBM for this engine's tables, 0 = all of them fit */
    size_t max_step_bytes;
    double index_s, device_s, write_s;
    double t_stage, t_run, t_fetch; /* NKB200_TIMES=1: where the raw-text device steps spend their host time */
    uint64_t h2d, d2h;
    int rc;
    char err[512];
} nk_dev;

struct nk_ctx
{
    nk_config cfg;
    char err[768];
    int depth_part;
    uint64_t cap0;
    int part_first, n_local;
    nk_part *part;
    int n_dev;
    nk_dev *dev;
    int threads;
    uint32_t step_pairs, step_ops, step_bytes;
    int dev_group; /* partitions per device launch group, 0 = all resident partitions */
    int serial_steps; /* NKB200_SERIAL_STEPS: one device step at a time per GPU (the host pipelines still overlap) */
    int raw_mode;     /* steps go to the device as raw record text (default); NKB200_HOST_PARSE=1 parses on the host */
    uint32_t raw_part_bytes; /* raw text per partition, mate and step */
    nk_lineidx lif, lir;     /* line indexes of the files being processed */
    /* the reverse file is counted while the first partitions are already being worked on (nk_roll_thread) */
    int rolling;
    uint64_t count_route[3]; /* inputs whose line ends were counted up front / alongside the first steps / by the step builders */
    pthread_t roll_th;
    pthread_mutex_t roll_mu;
    pthread_cond_t roll_cv;
    uint64_t raw_steps, parsed_steps;
    uint64_t seed_raw_records, seed_parsed_records; /* seed records taken from raw text on the device / parsed here */
    uint64_t waves; /* passes over disjoint sets of partitions in nk_process_* so far (1 per file when all tables fit) */
    /* seeding */
    uint8_t *seed_seq[2]; /* two staging buffers: one is parsed into while the GPUs seed from the other */
    nkd_read *seed_reads[2];
    uint64_t *seed_pos[2];
    size_t seed_cap_reads, seed_cap_bytes, seed_cap_ops;
    int seeded;
    /* totals */
    nk_totals tot;
    uint64_t file_max_used;
    /* per-file state shared with the device pipelines */
    nk_buf ff, rf;
    int paired;
    uint64_t *fs, *fe, *rs, *re;
    int finished;
};

static char g_create_err[768];
const char *nk_create_error(void) { return g_create_err; }
const char *nk_last_error(const nk_ctx *c) { return c ? c->err : g_create_err; }

static int nk_fail(nk_ctx *c, int code, const char *fmt, ...)
{
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER; /* engines dump their tables concurrently */
    va_list ap;
    va_start(ap, fmt);
    pthread_mutex_lock(&mu);
    vsnprintf(c ? c->err : g_create_err, 768, fmt, ap);
    pthread_mutex_unlock(&mu);
    va_end(ap);
    return code;
}

static char *nk_out_name(const char *dir, const char *base, int k, int depth_part, int t, const char *suffix) /* C:834-850 */
{
    size_t n = strlen(base) + 80 + (dir ? strlen(dir) + 1 : 0);
    char *s = malloc(n);
    if (t >= 0)
        snprintf(s, n, "%s%s%s.k%d_norm%d_thread%d.%s", dir ? dir : "", dir ? "/" : "", base, k, depth_part, t, suffix);
    else
        snprintf(s, n, "%s%s%s.k%d_norm%d.%s", dir ? dir : "", dir ? "/" : "", base, k, depth_part, suffix);
    return s;
}

#define NK_WBUF (4u << 20)

static void nk_free_stepbuf(nk_stepbuf *sb, int n_parts)
{
    if (sb->ps)
        for (int i = 0; i < n_parts; i++)
        {
            nkd_free_pinned(sb->ps[i].reads);
            free(sb->ps[i].spans);
        }
    free(sb->ps);
    free(sb->segs);
    nkd_free_pinned(sb->seq);
    nkd_free_pinned(sb->accept);
    nk_pinned_put(sb->raw, sb->raw_cap);
    nk_pinned_put(sb->out, sb->out_cap);
    free(sb->rsegs);
    free(sb->rres);
    free(sb->rseg_li);
    free(sb->copies);
    memset(sb, 0, sizeof *sb);
}

void nk_destroy(nk_ctx *c)
{
    if (!c)
        return;
    if (nk_env_on("NKB200_TIMES"))
        for (int d = 0; d < c->n_dev; d++)
            fprintf(stderr, "[nk] engine %d: raw steps stage %.3f s, run %.3f s, fetch %.3f s\n", d, c->dev[d].t_stage,
                    c->dev[d].t_run, c->dev[d].t_fetch);
    for (int d = 0; d < c->n_dev; d++)
    {
        nk_dev *dv = &c->dev[d];
        for (int b = 0; b < NK_NBUF; b++)
            nk_free_stepbuf(&dv->sb[b], dv->n_parts);
        if (dv->eng)
            nkd_destroy(dv->eng);
        free(dv->parts);
    }
    for (int i = 0; i < c->n_local; i++)
    {
        if (c->part[i].out_f)
            fclose(c->part[i].out_f);
        if (c->part[i].out_r)
            fclose(c->part[i].out_r);
        free(c->part[i].wbuf_f);
        free(c->part[i].wbuf_r);
    }
    for (int b = 0; b < 2; b++)
    {
        nkd_free_pinned(c->seed_seq[b]);
        nkd_free_pinned(c->seed_reads[b]);
        free(c->seed_pos[b]);
    }
    free(c->fs);
    free(c->fe);
    free(c->rs);
    free(c->re);
    free(c->part);
    free(c->dev);
    free(c);
}

static int nk_alloc_raw_bufs(nk_ctx *c, nk_dev *dv);
static void nk_copy_task(int i, void *a);
static int nk_have_avx2 = -1; /* streaming-store staging copies (nk_copy_stream_avx2); chosen once in nk_create */

static void nk_alloc_raw_task(int d, void *a)
{
    nk_ctx *c = a;
    c->dev[d].rc = nk_alloc_raw_bufs(c, &c->dev[d]);
}

int nk_create(const nk_config *cfg, nk_ctx **out)
{
    *out = NULL;
    /* the limits parse_arguments enforces, C:704-743 (k = 32 is rejected as in the reference: a 64-bit key
     * would also collide with the claim tag in bit 63 of the key field) */
    if (cfg->partitions <= 0 || cfg->partitions > NK_MAX_PARTITIONS)
        return nk_fail(NULL, NK_EINVAL, "Error: CPU count (%d) must be a positive integer and up to %d", cfg->partitions, NK_MAX_PARTITIONS);
    if (cfg->k < 5 || cfg->k > 31)
        return nk_fail(NULL, NK_EINVAL, "Error: Only kmer sizes (%d) of 5 to 31 are supported", cfg->k);
    if (cfg->coverage > 1 || cfg->coverage < 0.001)
        return nk_fail(NULL, NK_EINVAL, "Error: Coverage (%3.f) is the proportion of the sequence covered by high kmers and must be between 0 and 1", cfg->coverage);
    if (cfg->depth < 2 || cfg->depth / cfg->partitions < 2)
        return nk_fail(NULL, NK_EINVAL, "Error: Depth (%d) must be at least 2 x number of CPUs", cfg->depth);
    if (!cfg->in_fastq && cfg->out_fastq)
        return nk_fail(NULL, NK_EINVAL, "Error: cannot request an output format of FASTQ when input is FASTA");
    nk_ctx *c = calloc(1, sizeof *c);
    if (!c)
        return nk_fail(NULL, NK_ENOMEM, "Memory allocation failed");
    c->cfg = *cfg;
    c->depth_part = cfg->depth / cfg->partitions; /* C:674 */
    c->cap0 = nk_initial_capacity(cfg->memory_gb, cfg->partitions, cfg->k);
    c->part_first = cfg->part_count > 0 ? cfg->part_first : 0;
    c->n_local = cfg->part_count > 0 ? cfg->part_count : cfg->partitions;
    if (c->part_first < 0 || c->part_first + c->n_local > cfg->partitions)
    {
        free(c);
        return nk_fail(NULL, NK_EINVAL, "partition slice outside 0..%d", cfg->partitions);
    }
    if ((cfg->merged_table || cfg->merged_output) && c->n_local != cfg->partitions)
    {
        free(c);
        return nk_fail(NULL, NK_EINVAL, "merged table / merged output need a context that owns all %d partitions", cfg->partitions);
    }
    c->threads = nk_host_threads();
    /* process-wide choices of CPU features, made here while a single thread runs */
    if (!nk_mask64)
        nk_mask64 = nk_mask64_pick();
    if (nk_have_avx2 < 0)
    {
        __builtin_cpu_init();
        nk_have_avx2 = __builtin_cpu_supports("avx2") && !nk_env_on("NKB200_PLAIN_MEMCPY");
    }
    int ndev_avail = nkd_device_count();
    if (ndev_avail <= 0)
    {
        free(c);
        return nk_fail(NULL, NK_ENODEVICE, "no CUDA device: the B200 path has no CPU fallback");
    }
    /* Partition t lives on GPU t mod G: no data moves between GPUs on the hot path.  A GPU's partitions are
     * dealt to (by default) up to four engines, each with its own stream, scratch lists and pipeline thread, so that
     * one engine's list passes (classify, sort, rank, commit) and host round trips overlap the other's table
     * probing on the same GPU. */
    /* default: up to four engines, each with at least two partitions (measured at 8, 4, 2 and 1 partitions per GPU,
     * profiles/r02_ab_engines.txt: an engine with a single partition needs steps so large that the ordered slow path
     * grows faster than the overlap between engines pays) */
    int epg = getenv("NKB200_ENGINES_PER_GPU") ? atoi(getenv("NKB200_ENGINES_PER_GPU")) : 0;
    if (epg < 0 || epg > 8)
        epg = 1;
    int n_gpus = cfg->n_devices > 0 ? cfg->n_devices : 1;
    if (n_gpus > c->n_local)
        n_gpus = c->n_local;
    int *first_eng = calloc((size_t)n_gpus + 1, sizeof(int)), *n_eng = calloc((size_t)n_gpus, sizeof(int));
    for (int g = 0; g < n_gpus; g++)
    {
        int cnt = (c->n_local - g + n_gpus - 1) / n_gpus; /* partitions i with i mod n_gpus == g */
        int want = epg ? epg : (cnt / 2 > 4 ? 4 : (cnt / 2 < 1 ? 1 : cnt / 2));
        n_eng[g] = cnt < want ? cnt : want;
        first_eng[g + 1] = first_eng[g] + n_eng[g];
    }
    c->n_dev = first_eng[n_gpus];
    c->dev = calloc((size_t)c->n_dev, sizeof *c->dev);
    c->part = calloc((size_t)c->n_local, sizeof *c->part);
    c->fs = calloc((size_t)cfg->partitions, sizeof(uint64_t));
    c->fe = calloc((size_t)cfg->partitions, sizeof(uint64_t));
    c->rs = calloc((size_t)cfg->partitions, sizeof(uint64_t));
    c->re = calloc((size_t)cfg->partitions, sizeof(uint64_t));
    for (int g = 0; g < n_gpus; g++)
        for (int d = first_eng[g]; d < first_eng[g + 1]; d++)
        {
            c->dev[d].ordinal = cfg->devices ? cfg->devices[g] : g;
            c->dev[d].lead = first_eng[g];
            pthread_mutex_init(&c->dev[d].step_lock, NULL);
            c->dev[d].parts = calloc((size_t)c->n_local, sizeof(int));
        }
    for (int i = 0; i < c->n_local; i++)
    {
        nk_part *p = &c->part[i];
        int g = i % n_gpus;
        p->gid = c->part_first + i;
        p->dev = first_eng[g] + (i / n_gpus) % n_eng[g];
        nk_dev *dv = &c->dev[p->dev];
        p->lidx = dv->n_parts;
        dv->parts[dv->n_parts++] = i;
    }
    free(first_eng);
    free(n_eng);
    /* step sizing: records per partition per step; operations and bytes follow from 150-base reads,
     * longer reads simply end a partition's batch earlier */
    int max_dev_parts = 0;
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].n_parts > max_dev_parts)
            max_dev_parts = c->dev[d].n_parts;
    uint32_t sp = cfg->step_pairs;
    if (!sp)
    {
        const char *e = getenv("NKB200_STEP_PAIRS");
        if (e && atoi(e) > 0)
            sp = (uint32_t)atoi(e);
        else
        { /* about 37 M operations per launch (measured: profiles/r01_sweep_engines.txt; launches below ~20 M
           * operations waste their list chunks).  Round 1 gave a GPU's only engine twice that; with one partition per
           * engine (8 GPUs for -p 8) that meant steps of 262,144 pairs, where the ordered slow path is 25 % of the kernel
           * time and five steps leave little to overlap: 131,072 is 4 % faster on the device and 8 % end to end
           * (profiles/r02_ab_engines.txt) */
            uint64_t target = 48ull << 20;
            sp = 262144u;
            while (sp > 2048 && (uint64_t)sp * 288u * (uint64_t)max_dev_parts > target)
                sp /= 2;
        }
    }
    if (sp < 16)
        sp = 16;
    c->dev_group = getenv("NKB200_GROUP") ? atoi(getenv("NKB200_GROUP")) : 0;
    c->serial_steps = nk_env_on("NKB200_SERIAL_STEPS");
    c->raw_mode = !nk_env_on("NKB200_HOST_PARSE");
    int grp = c->dev_group > 0 && c->dev_group < max_dev_parts ? c->dev_group : max_dev_parts;
    if (grp < max_dev_parts)
        c->raw_mode = 0; /* launch groups exist for the parsed path only (an experiment knob) */
    c->step_pairs = sp;
    /* operations and bytes a partition's share of a step may hold: a 150-base pair is ~660 bytes of FASTQ and at
     * most 292 operations; a raw-text step is cut so that bytes / 2 (FASTQ: sequence + quality) stays below step_ops */
    c->step_ops = sp * 336u < 4096u ? 4096u : sp * 336u;
    c->step_bytes = (sp * 2u * 176u + 4096u) & ~15u;
    c->raw_part_bytes = (sp * 384u + 16384u) & ~15u;
    for (int d = 0; d < c->n_dev; d++)
    {
        nk_dev *dv = &c->dev[d];
        nkd_config ec;
        memset(&ec, 0, sizeof ec);
        ec.device = dv->ordinal;
        ec.k = cfg->k;
        ec.canonical = cfg->canonical;
        ec.depth_per_part = c->depth_part;
        ec.coverage = cfg->coverage;
        ec.n_parts = dv->n_parts;
        ec.capacity0 = c->cap0;
        int eg = grp < dv->n_parts ? grp : dv->n_parts; /* partitions in one device launch group */
        ec.max_step_reads = (uint64_t)eg * sp * 2u + 16;
        ec.max_step_bytes = (uint64_t)dv->n_parts * c->step_bytes + 64;
        ec.max_step_ops = (uint64_t)eg * c->step_ops + 64;
        ec.max_raw_bytes = c->raw_mode ? (uint64_t)dv->n_parts * 2u * (c->raw_part_bytes + 16u) : 0;
        dv->max_step_bytes = (size_t)ec.max_step_bytes;
        int rc = nkd_create(&ec, &dv->eng);
        if (rc)
        {
            nk_fail(NULL, rc, "%s", dv->eng ? nkd_last_error(dv->eng) : "engine allocation failed");
            nk_destroy(c);
            return rc;
        }
        for (int b = 0; b < NK_NBUF; b++)
        {
            nk_stepbuf *sb = &dv->sb[b];
            sb->ps = calloc((size_t)dv->n_parts, sizeof *sb->ps);
            sb->segs = calloc((size_t)dv->n_parts, sizeof *sb->segs);
            if (!sb->ps || !sb->segs)
            {
                nk_fail(NULL, NK_ENOMEM, "Memory allocation failed (staging buffers)");
                nk_destroy(c);
                return NK_ENOMEM;
            }
        }
    }
    /* the raw-text staging is pinned now, outside any timed region and on all engines at once (page-locking a GB
     * takes ~0.4 s); the host parser's staging only if it is ever needed */
    if (c->raw_mode)
    {
        nk_parallel_for(c->n_dev, c->n_dev, nk_alloc_raw_task, c);
        for (int d = 0; d < c->n_dev; d++)
            if (c->dev[d].rc)
            {
                nk_fail(NULL, NK_ENOMEM, "Memory allocation failed (staging buffers)");
                nk_destroy(c);
                return NK_ENOMEM;
            }
    }
    /* seed staging: as large as the smallest engine's step */
    int min_dev_parts = max_dev_parts;
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].n_parts < min_dev_parts)
            min_dev_parts = c->dev[d].n_parts;
    if (min_dev_parts > grp)
        min_dev_parts = grp;
    c->seed_cap_reads = (size_t)min_dev_parts * sp * 2u;
    c->seed_cap_bytes = (size_t)min_dev_parts * c->step_bytes;
    c->seed_cap_ops = (size_t)min_dev_parts * c->step_ops;
    if (c->seed_cap_ops > (1u << 27))
        c->seed_cap_ops = 1u << 27; /* one table: stay below 2^28 operations per step */
    for (int b = 0; b < 2; b++)
    {
        c->seed_seq[b] = nkd_alloc_pinned(c->seed_cap_bytes + 64);
        c->seed_reads[b] = nkd_alloc_pinned(c->seed_cap_reads * sizeof(nkd_read));
        c->seed_pos[b] = malloc(c->seed_cap_reads * sizeof(uint64_t));
        if (!c->seed_seq[b] || !c->seed_reads[b] || !c->seed_pos[b])
        {
            nk_fail(NULL, NK_ENOMEM, "Memory allocation failed (seed staging)");
            nk_destroy(c);
            return NK_ENOMEM;
        }
        memset(c->seed_seq[b], 0, c->seed_cap_bytes + 64);
    }
    *out = c;
    return NK_OK;
}

/* Page-locked staging of an engine's NK_NBUF steps, allocated when a pipeline of that kind first runs (pinning
 * memory costs ~0.4 s per GB): parsed steps hold sequence lines + read descriptors + accept flags, raw-text steps
 * hold record text in and the accepted records' text out. */
static int nk_alloc_parsed_bufs(nk_ctx *c, nk_dev *dv)
{
    if (dv->have_parsed_bufs)
        return NK_OK;
    const uint32_t sp = c->step_pairs;
    for (int b = 0; b < NK_NBUF; b++)
    {
        nk_stepbuf *sb = &dv->sb[b];
        sb->seq = nkd_alloc_pinned(dv->max_step_bytes + 64);
        sb->accept = nkd_alloc_pinned((size_t)dv->n_parts * sp + 16);
        if (!sb->seq || !sb->accept)
            return NK_ENOMEM;
        memset(sb->seq, 0, dv->max_step_bytes + 64);
        for (int i = 0; i < dv->n_parts; i++)
        {
            sb->ps[i].reads = nkd_alloc_pinned((size_t)sp * 2u * sizeof(nkd_read));
            sb->ps[i].spans = malloc((size_t)sp * 2u * sizeof(nk_span));
            sb->ps[i].seq_lo = (size_t)i * c->step_bytes;
            sb->ps[i].seq_end = sb->ps[i].seq_lo + c->step_bytes;
            if (!sb->ps[i].reads || !sb->ps[i].spans)
                return NK_ENOMEM;
        }
    }
    dv->have_parsed_bufs = 1;
    return NK_OK;
}

#define NK_COPY_PIECE ((size_t)1 << 20) /* bytes one pool task copies into the staging buffer */

static int nk_alloc_raw_bufs(nk_ctx *c, nk_dev *dv)
{
    if (dv->have_raw_bufs)
        return NK_OK;
    size_t cap = (size_t)dv->n_parts * 2u * ((size_t)c->raw_part_bytes + 16u) + 64;
    for (int b = 0; b < NK_NBUF; b++)
    {
        nk_stepbuf *sb = &dv->sb[b];
        sb->raw_cap = cap;
        sb->out_cap = cap + (size_t)dv->n_parts * 4u * c->step_pairs; /* fq->fa may add "/1" per record */
        sb->raw = nk_pinned_get(sb->raw_cap);
        sb->out = nk_pinned_get(sb->out_cap);
        sb->rsegs = calloc((size_t)dv->n_parts, sizeof *sb->rsegs);
        sb->rres = calloc((size_t)dv->n_parts, sizeof *sb->rres);
        sb->rseg_li = calloc((size_t)dv->n_parts, sizeof *sb->rseg_li);
        sb->copies = calloc((size_t)dv->n_parts * 2u * ((size_t)c->raw_part_bytes / NK_COPY_PIECE + 2), sizeof *sb->copies);
        if (!sb->raw || !sb->out || !sb->rsegs || !sb->rres || !sb->rseg_li || !sb->copies)
            return NK_ENOMEM;
    }
    dv->have_raw_bufs = 1;
    return NK_OK;
}

/* every partition's outputs are opened "w" once, after seeding (C:2286-2302 follows C:2244-2250), and appended to
 * across input files */
static int nk_open_outputs(nk_ctx *c)
{
    const nk_config *cfg = &c->cfg;
    for (int i = 0; i < c->n_local; i++)
    {
        nk_part *p = &c->part[i];
        char *n = nk_out_name(cfg->out_dir, "output_forward", cfg->k, c->depth_part, p->gid, "fastq");
        p->out_f = fopen(n, "w");
        if (!p->out_f)
        {
            int rc = nk_fail(c, NK_EIO, "Error opening file to write: %s", n);
            free(n);
            return rc;
        }
        free(n);
        p->wbuf_f = malloc(NK_WBUF);
        setvbuf(p->out_f, p->wbuf_f, _IOFBF, NK_WBUF);
        if (cfg->have_reverse)
        {
            n = nk_out_name(cfg->out_dir, "output_reverse", cfg->k, c->depth_part, p->gid, "fastq");
            p->out_r = fopen(n, "w");
            if (!p->out_r)
            {
                int rc = nk_fail(c, NK_EIO, "Error opening file to write: %s", n);
                free(n);
                return rc;
            }
            free(n);
            p->wbuf_r = malloc(NK_WBUF);
            setvbuf(p->out_r, p->wbuf_r, _IOFBF, NK_WBUF);
        }
    }
    return NK_OK;
}

/* ------------------------------------------------------------------ seeding (C:1322-1373) */

typedef struct
{
    nk_ctx *c;
    int buf;
    size_t n_reads, bytes;
    const nk_buf *f;
    int64_t inv[64];
    int rc[64];
    int result;
    pthread_t th;
    int running;
} nk_seed_job;

static void nk_seed_task(int d, void *a)
{
    nk_seed_job *j = a;
    nk_dev *dv = &j->c->dev[d];
    j->inv[d] = -1;
    j->rc[d] = NK_OK;
    if (dv->lead != d)
        return; /* engines sharing a GPU copy the lead engine's seed table in nk_seed_finish */
    j->rc[d] = nkd_seed_step(dv->eng, j->c->seed_seq[j->buf], j->bytes, j->c->seed_reads[j->buf], j->n_reads, &j->inv[d]);
}

static void nk_scrub_copy(char *dst, const char *src, size_t n)
{
    for (size_t i = 0; i < n; i++)
        dst[i] = src[i] == 'N' ? 'A' : src[i];
    dst[n] = 0;
}

/* one batch of seed reads on every GPU (each builds the same seed table) */
static void *nk_seed_flush_thread(void *a)
{
    nk_seed_job *j = a;
    nk_ctx *c = j->c;
    j->result = NK_OK;
    if (nk_env_on("NKB200_DEBUG"))
        fprintf(stderr, "[nk] seed flush: %zu reads, %zu bytes\n", j->n_reads, j->bytes);
    nk_parallel_for(c->n_dev, c->n_dev, nk_seed_task, j);
    for (int d = 0; d < c->n_dev; d++)
        if (j->rc[d] && !j->result)
            j->result = nk_fail(c, j->rc[d], "%s", nkd_last_error(c->dev[d].eng));
    if (!j->result && j->inv[0] >= 0)
    { /* is_valid_sequence_single's abort, C:1416-1420 */
        size_t i = (size_t)j->inv[0];
        const nkd_read *rd = &c->seed_reads[j->buf][i];
        char *s = malloc((size_t)rd->len + 1);
        nk_scrub_copy(s, j->f->data + c->seed_pos[j->buf][i], rd->len);
        nk_fail(c, NK_EDATA, "FATAL: FWD sequence does not appear to be a DNA sequence\n%s\n", s);
        free(s);
        j->result = NK_EDATA;
    }
    return NULL;
}

static int nk_seed_join(nk_seed_job *j)
{
    if (!j->running)
        return NK_OK;
    pthread_join(j->th, NULL);
    j->running = 0;
    return j->result;
}

/* seed_kmer_hash with the records parsed here (the byte-exact path; what the device declines ends up here) from
 * file offset `start` on */
static int nk_seed_buffer_parsed(nk_ctx *c, const char *data, size_t size, size_t start, int records_to_seed)
{
    double t0 = nk_now();
    if (!nk_mask64)
        nk_mask64 = nk_mask64_pick();
    nk_buf f = {data, size};
    int per = c->cfg.in_fastq ? 4 : 2, k = c->cfg.k;
    int done = 0, rc = NK_OK, cur = 0;
    size_t n_reads = 0, bytes = 0, ops = 0;
    nk_seed_job jobs[2];
    memset(jobs, 0, sizeof jobs);
    /* seed_kmer_hash splits on '\n' only and needs every line of a record terminated (C:1334-1344); the GPUs
     * seed from one staging buffer while the next batch is parsed into the other */
    nk_nliter it;
    nk_nliter_seek(&it, data, start, size);
    size_t rec_start = start;
    for (;;)
    {
        size_t q = rec_start, seq_start = 0, seq_len = 0;
        int ok = 1;
        for (int i = 0; i < per; i++)
        {
            size_t nl = nk_nliter_next(&it);
            if (nl == SIZE_MAX)
            {
                ok = 0;
                break;
            }
            if (i == 1)
            {
                seq_start = q;
                seq_len = nl - q;
            }
            q = nl + 1;
        }
        if (!ok)
            break; /* no further complete record */
        rec_start = q;
        size_t slen = it.nul_seen ? strnlen(data + seq_start, seq_len) : seq_len; /* strlen semantics, C:1347 */
        if (slen <= (size_t)k) /* strictly longer than K */
            continue;
        if (slen >= NK_MAX_LINE)
        {
            rc = nk_fail(c, NK_EDATA, "seed record with a sequence line of %zu chars (limit %d)", slen, NK_MAX_LINE - 1);
            break;
        }
        size_t need = (slen + 15) & ~(size_t)15, nops = slen - (size_t)k + 1;
        if (n_reads + 1 > c->seed_cap_reads || bytes + need > c->seed_cap_bytes || ops + nops > c->seed_cap_ops)
        {
            /* hand the full buffer to the GPUs (after the previous hand-off finished) and switch */
            rc = nk_seed_join(&jobs[cur ^ 1]);
            if (rc)
                break;
            jobs[cur].c = c;
            jobs[cur].buf = cur;
            jobs[cur].n_reads = n_reads;
            jobs[cur].bytes = bytes;
            jobs[cur].f = &f;
            if (pthread_create(&jobs[cur].th, NULL, nk_seed_flush_thread, &jobs[cur]) != 0)
            {
                rc = nk_fail(c, NK_EINTERNAL, "cannot start the seeding thread");
                break;
            }
            jobs[cur].running = 1;
            cur ^= 1;
            n_reads = bytes = ops = 0;
        }
        uint8_t *dst = c->seed_seq[cur] + bytes;
        memcpy(dst, data + seq_start, slen);
        memset(dst + slen, 0, need - slen);
        nkd_read *rd = &c->seed_reads[cur][n_reads];
        rd->seq_off = (uint32_t)bytes;
        rd->op_base = (uint32_t)ops;
        rd->len = (uint16_t)slen;
        rd->part = 0;
        rd->reserved = 0;
        c->seed_pos[cur][n_reads] = seq_start;
        n_reads++;
        bytes += need;
        ops += nops;
        c->seed_parsed_records++;
        if (++done == records_to_seed)
            break;
    }
    /* batches must reach the table in order: wait for the one in flight, then run the last one */
    int rc2 = nk_seed_join(&jobs[cur ^ 1]);
    if (!rc)
        rc = rc2;
    if (!rc && n_reads)
    {
        jobs[cur].c = c;
        jobs[cur].buf = cur;
        jobs[cur].n_reads = n_reads;
        jobs[cur].bytes = bytes;
        jobs[cur].f = &f;
        nk_seed_flush_thread(&jobs[cur]);
        rc = jobs[cur].result;
    }
    c->tot.seed_seconds += nk_now() - t0;
    return rc;
}

/* ---- seeding on raw record text: the host cuts the head of the file into pieces of whole records (parallel
 * line count), copies them into the lead engines' page-locked step buffer, and the device does the rest
 * (nkd_seed_raw): 6 M records no longer wait for a single-threaded parse */

typedef struct
{
    nk_ctx *c;
    const uint8_t *raw;
    size_t text_bytes, pos; /* the piece: data[pos, pos + text_bytes) */
    uint32_t n_records, limit;
    uint32_t taken[64];
    int64_t inv[64];
    int rc[64];
    pthread_t th;
} nk_seedraw_job;

static void nk_seedraw_task(int d, void *a)
{
    nk_seedraw_job *j = a;
    nk_dev *dv = &j->c->dev[d];
    j->rc[d] = NK_OK;
    j->inv[d] = -1;
    j->taken[d] = 0;
    if (dv->lead != d)
        return;
    j->rc[d] = nkd_seed_raw(dv->eng, j->raw, j->text_bytes, j->n_records, j->c->cfg.in_fastq ? 4 : 2, j->limit, &j->taken[d],
                            &j->inv[d]);
}

static void *nk_seedraw_thread(void *a)
{
    nk_seedraw_job *j = a;
    nk_parallel_for(j->c->n_dev, j->c->n_dev, nk_seedraw_task, j);
    return NULL;
}

static void nk_seedraw_forget(nk_ctx *c)
{
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].lead == d)
            nkd_upload_raw(c->dev[d].eng, NULL, 0);
}

/* returns NK_OK with *consumed = bytes of the file that were dealt with and *seeded = records taken; stops early
 * (without error) at text the device declines, which the caller gives to the host parser.
 * Two step buffers take turns: while the GPUs insert one piece, the host counts and copies the next one and sends it
 * ahead on the upload stream (nkd_upload_raw), so the pieces reach the table in order without the device waiting. */
static int nk_seed_buffer_raw(nk_ctx *c, const char *data, size_t size, int records_to_seed, size_t *consumed, int *seeded)
{
    const int per = c->cfg.in_fastq ? 4 : 2;
    nk_buf f = {data, size};
    size_t cap = (size_t)c->dev[0].n_parts * 2u * ((size_t)c->raw_part_bytes + 16u);
    for (int d = 0; d < c->n_dev; d++)
    { /* every lead engine takes the same pieces: the smallest step buffer decides their size */
        size_t cd = (size_t)c->dev[d].n_parts * 2u * ((size_t)c->raw_part_bytes + 16u);
        if (c->dev[d].lead == d && cd < cap)
            cap = cd;
    }
    uint32_t max_records = (uint32_t)c->seed_cap_reads;
    size_t pos = 0; /* where the next piece starts */
    *consumed = 0;
    *seeded = 0;
    if (c->n_dev > 64)
        return NK_OK;
    const int ahead = !nk_env_on("NKB200_NO_PREFETCH");
    nk_seedraw_job jobs[2], *inflight = NULL;
    memset(jobs, 0, sizeof jobs);
    int rc = NK_OK, slot = 0, exhausted = 0;
    for (;;)
    {
        /* 1. the next piece, into the buffer the piece in flight does not use.  How many records are still wanted is
         * known only when that piece is done: prepare on the assumption that all of its records count, and nothing
         * at all if that would already be enough. */
        int64_t need = (int64_t)records_to_seed - *seeded - (inflight ? (int64_t)inflight->n_records : 0);
        nk_seedraw_job *next = NULL;
        if (need > 0 && pos < size && !exhausted)
        {
            nk_stepbuf *sb = &c->dev[0].sb[slot]; /* free until processing starts */
            /* as many whole records as fit the buffer, the read limit and the operation limit (a sequence line holds
             * at most as many k-mers as it has bytes; FASTQ spends half of a record's bytes on it) */
            size_t want_bytes = cap - 64;
            size_t ops_bytes = c->cfg.in_fastq ? c->seed_cap_ops * 2 : c->seed_cap_ops;
            if (ops_bytes < want_bytes)
                want_bytes = ops_bytes;
            size_t hi = pos + want_bytes < size ? pos + want_bytes : size;
            nk_lineidx li = {0};
            nk_lineidx_build_range(&li, &f, c->threads, pos, hi);
            uint64_t before = nk_lineidx_before(&li, pos), upto = nk_lineidx_before(&li, hi);
            uint64_t recs = (upto - before) / (uint64_t)per;
            if (recs > max_records)
                recs = max_records;
            if (recs > (uint64_t)need + (uint64_t)need / 8 + 4096)
                recs = (uint64_t)need + (uint64_t)need / 8 + 4096; /* a few more than needed: some may be too short to count */
            size_t end = recs ? nk_lineidx_find(&li, before + recs * (uint64_t)per - 1) : SIZE_MAX;
            nk_lineidx_free(&li);
            if (!recs || end == SIZE_MAX)
                exhausted = 1; /* no whole record left (the tail goes to the host parser, which knows what to do with it) */
            else
            {
                end += 1;
                size_t n = end - pos, n16 = (n + 15) & ~(size_t)15;
                int n_copies = 0;
                for (size_t o = 0; o < n; o += NK_COPY_PIECE)
                {
                    nk_copy *cp = &sb->copies[n_copies++];
                    cp->dst = sb->raw + o;
                    cp->src = data + pos + o;
                    cp->n = n - o < NK_COPY_PIECE ? n - o : NK_COPY_PIECE;
                }
                nk_parallel_for(n_copies, c->threads, nk_copy_task, sb->copies);
                memset(sb->raw + n, ' ', n16 - n);
                next = &jobs[slot];
                next->c = c;
                next->raw = sb->raw;
                next->text_bytes = n;
                next->pos = pos;
                next->n_records = (uint32_t)recs;
                if (ahead) /* travels while the piece in flight is being inserted */
                    for (int d = 0; d < c->n_dev; d++)
                        if (c->dev[d].lead == d)
                            nkd_upload_raw(c->dev[d].eng, sb->raw, n16);
            }
        }
        /* 2. the piece in flight: its outcome decides whether the prepared one is wanted */
        if (inflight)
        {
            nk_seedraw_job *j = inflight;
            pthread_join(j->th, NULL);
            inflight = NULL;
            int declined = 0;
            for (int d = 0; d < c->n_dev; d++)
            {
                if (j->rc[d] == NK_EIRREGULAR)
                    declined = 1;
                else if (j->rc[d] && !rc)
                    rc = nk_fail(c, j->rc[d], "%s", nkd_last_error(c->dev[d].eng));
            }
            if (rc || declined)
                break; /* declined: nothing of that piece was inserted, the host parser continues at its start */
            if (j->inv[0] >= 0)
            { /* is_valid_sequence_single's abort, C:1416-1420: the text of record inv's sequence line */
                const char *t = data + j->pos;
                size_t n = j->text_bytes;
                size_t q = j->inv[0] ? nk_kth_newline(t, n, (uint64_t)per * (uint64_t)j->inv[0]) + 1 : 0;
                q += nk_kth_newline(t + q, n - q, 1) + 1;
                size_t len = nk_kth_newline(t + q, n - q, 1);
                char *txt = malloc(len + 1);
                nk_scrub_copy(txt, t + q, len);
                rc = nk_fail(c, NK_EDATA, "FATAL: FWD sequence does not appear to be a DNA sequence\n%s\n", txt);
                free(txt);
                break;
            }
            *seeded += (int)j->taken[0];
            *consumed = j->pos + j->text_bytes;
            if (*seeded >= records_to_seed)
                break;
            if (!next && !exhausted && pos < size)
                continue; /* nothing was prepared because that piece might have sufficed; it did not */
        }
        if (!next)
            break;
        /* 3. hand the prepared piece to the GPUs */
        next->limit = (uint32_t)(records_to_seed - *seeded);
        if (pthread_create(&next->th, NULL, nk_seedraw_thread, next) != 0)
        {
            rc = nk_fail(c, NK_EINTERNAL, "cannot start the seeding thread");
            break;
        }
        inflight = next;
        slot ^= 1;
        pos = next->pos + next->text_bytes;
    }
    if (inflight)
        pthread_join(inflight->th, NULL);
    nk_seedraw_forget(c);
    return rc;
}

int nk_seed_buffer(nk_ctx *c, const char *data, size_t size, int records_to_seed)
{
    if (c->seeded)
        return nk_fail(c, NK_EINVAL, "nk_seed_buffer after nk_seed_finish");
    double t0 = nk_now();
    size_t consumed = 0;
    int seeded = 0, rc = NK_OK;
    if (c->raw_mode && !nk_env_on("NKB200_HOST_SEED") && c->dev[0].have_raw_bufs)
        rc = nk_seed_buffer_raw(c, data, size, records_to_seed, &consumed, &seeded);
    c->tot.seed_seconds += nk_now() - t0;
    c->seed_raw_records += (uint64_t)seeded;
    /* the rest (text the device declined, the unterminated tail of a file, or everything when raw text is off) */
    if (!rc && seeded < records_to_seed && consumed < size)
        rc = nk_seed_buffer_parsed(c, data, size, consumed, records_to_seed - seeded);
    return rc;
}

#define NK_DUMP_CHUNK (2u << 20) /* entries formatted per device pass */

/* print_kmer_table (C:354-385): slot order, stored keys only.  The lines are formatted on the GPU
 * (nkd_dump_text); the host only appends the text to the file. */
static int nk_write_dump(nk_ctx *c, nkd_engine *e, int part, uint64_t entries, const char *tag, int gid)
{
    char base[32];
    snprintf(base, sizeof base, "output_kmer%s", tag);
    char *name = nk_out_name(c->cfg.out_dir, base, c->cfg.k, c->depth_part, gid, "tsv");
    FILE *o = fopen(name, "w");
    if (!o)
    {
        int rc = nk_fail(c, NK_EIO, "cannot open %s", name);
        free(name);
        return rc;
    }
    free(name);
    setvbuf(o, NULL, _IONBF, 0);
    size_t per = (size_t)c->cfg.k + (part == NKD_PART_MERGED ? 22u : 13u);
    uint64_t chunk = entries < NK_DUMP_CHUNK ? (entries ? entries : 1) : NK_DUMP_CHUNK;
    char *text = nkd_alloc_pinned(chunk * per);
    int rc = text ? NK_OK : nk_fail(c, NK_ENOMEM, "Memory allocation failed (table dump)");
    for (uint64_t at = 0; !rc && at < entries; at += chunk)
    {
        uint64_t n = entries - at < chunk ? entries - at : chunk;
        size_t bytes = 0;
        rc = nkd_dump_text(e, part, at, n, text, chunk * per, &bytes);
        if (rc)
            nk_fail(c, rc, "%s", nkd_last_error(e));
        else if (bytes && fwrite(text, 1, bytes, o) != bytes)
            rc = nk_fail(c, NK_EIO, "error writing the k-mer table: %s", strerror(errno));
    }
    nkd_free_pinned(text);
    if (fclose(o) != 0 && !rc)
        rc = nk_fail(c, NK_EIO, "error closing the k-mer table: %s", strerror(errno));
    return rc;
}

/* all partitions' stored k-mers once, ascending, counts summed (the TODO of C:25-26):
 * output_kmer_merged.k{K}_norm{D}.tsv.  Partitions on the first GPU are compacted in place, the
 * others on their own GPU and handed over through the host. */
static int nk_write_merged_table(nk_ctx *c)
{
    uint64_t total = 0;
    for (int i = 0; i < c->n_local; i++)
    {
        nkd_part_stats st;
        nkd_part_stats_get(c->dev[c->part[i].dev].eng, c->part[i].lidx, &st);
        total += st.used;
    }
    nkd_engine *e0 = c->dev[0].eng;
    int rc = nkd_merge_begin(e0, total);
    if (rc)
        return nk_fail(c, rc, "%s", nkd_last_error(e0));
    for (int i = 0; i < c->n_local && !rc; i++)
    {
        nk_part *p = &c->part[i];
        if (p->dev == 0)
        {
            if ((rc = nkd_merge_add_part(e0, p->lidx)) != 0)
                nk_fail(c, rc, "%s", nkd_last_error(e0));
            continue;
        }
        nkd_engine *e = c->dev[p->dev].eng;
        nkd_part_stats st;
        nkd_part_stats_get(e, p->lidx, &st);
        uint64_t n = 0, cap = st.used ? st.used : 1;
        uint64_t *keys = nkd_alloc_pinned(cap * sizeof *keys);
        int64_t *vals = nkd_alloc_pinned(cap * sizeof *vals);
        if (!keys || !vals)
            rc = nk_fail(c, NK_ENOMEM, "Memory allocation failed (merged table)");
        else if ((rc = nkd_compact(e, p->lidx, keys, vals, cap, &n)) != 0)
            nk_fail(c, rc, "%s", nkd_last_error(e));
        else if ((rc = nkd_merge_add(e0, keys, vals, n)) != 0)
            nk_fail(c, rc, "%s", nkd_last_error(e0));
        nkd_free_pinned(keys);
        nkd_free_pinned(vals);
    }
    uint64_t distinct = 0;
    if (!rc && (rc = nkd_merge_finish(e0, &distinct)) != 0)
        nk_fail(c, rc, "%s", nkd_last_error(e0));
    if (!rc)
        rc = nk_write_dump(c, e0, NKD_PART_MERGED, distinct, "_merged", -1);
    nkd_merge_begin(e0, 0); /* drop the merge buffers */
    return rc;
}

/* one Trinity-ready file per mate: the partitions' outputs in partition order (the reference leaves the
 * concatenation to the user): output_forward.k{K}_norm{D}.fastq / output_reverse.... */
static int nk_concat_outputs(nk_ctx *c, const char *base)
{
    char *name = nk_out_name(c->cfg.out_dir, base, c->cfg.k, c->depth_part, -1, "fastq");
    int out = open(name, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (out < 0)
    {
        int rc = nk_fail(c, NK_EIO, "Error opening file to write: %s", name);
        free(name);
        return rc;
    }
    free(name);
    int rc = NK_OK;
    char *buf = malloc(NK_WBUF);
    for (int i = 0; i < c->n_local && !rc; i++)
    {
        char *pn = nk_out_name(c->cfg.out_dir, base, c->cfg.k, c->depth_part, c->part[i].gid, "fastq");
        int in = open(pn, O_RDONLY);
        if (in < 0)
            rc = nk_fail(c, NK_EIO, "cannot reopen %s", pn);
        for (ssize_t got = 0; !rc && (got = read(in, buf, NK_WBUF)) != 0;)
        {
            if (got < 0)
            {
                if (errno == EINTR)
                    continue;
                rc = nk_fail(c, NK_EIO, "error reading %s: %s", pn, strerror(errno));
                break;
            }
            for (ssize_t done = 0; done < got && !rc;)
            {
                ssize_t w = write(out, buf + done, (size_t)(got - done));
                if (w < 0 && errno != EINTR)
                    rc = nk_fail(c, NK_EIO, "error writing the merged output: %s", strerror(errno));
                else if (w > 0)
                    done += w;
            }
        }
        if (in >= 0)
            close(in);
        free(pn);
    }
    free(buf);
    if (close(out) != 0 && !rc)
        rc = nk_fail(c, NK_EIO, "error closing the merged output: %s", strerror(errno));
    return rc;
}

/* Do the tables of all partitions fit their GPU, with room to grow?  If not, every engine of that GPU gets a budget
 * and nk_process_* works on its partitions in waves (SURVEY 8.B row e; partitions are independent, C:1841-1880).
 * NKB200_TABLE_BUDGET_MB forces a budget per engine (tests). */
static int nk_set_table_budgets(nk_ctx *c)
{
    const char *forced = getenv("NKB200_TABLE_BUDGET_MB");
    for (int d = 0; d < c->n_dev; d++)
    {
        nk_dev *dv = &c->dev[d];
        if (dv->lead != d)
            continue;
        nkd_part_stats st;
        nkd_seed_stats(dv->eng, &st);
        uint64_t table = st.capacity * 16, free_b = 0, total_b = 0, parts_on_gpu = 0, engines = 0;
        for (int o = 0; o < c->n_dev; o++)
            if (c->dev[o].lead == d)
            {
                parts_on_gpu += (uint64_t)c->dev[o].n_parts;
                engines++;
            }
        uint64_t budget = 0;
        if (forced && atoll(forced) > 0)
            budget = (uint64_t)atoll(forced) << 20;
        else if (nkd_device_memory(dv->ordinal, &free_b, &total_b) == NK_OK)
        {
            /* resident for good needs every table plus room for each to grow once (x1.5) and a re-hash in flight */
            uint64_t want = parts_on_gpu * table * 3 / 2 + engines * table * 3 / 2;
            if (want > free_b - free_b / 10)
                budget = (free_b - free_b / 10) / engines;
        }
        for (int o = 0; o < c->n_dev; o++)
            if (c->dev[o].lead == d)
            {
                c->dev[o].table_budget = budget;
                int rc = nkd_set_table_budget(c->dev[o].eng, budget);
                if (rc)
                    return nk_fail(c, rc, "%s", nkd_last_error(c->dev[o].eng));
            }
        if (budget && c->cfg.verbose)
            printf("B200: GPU %d holds %llu partitions of %.2f GB tables in %.1f GB: working in waves, %.1f GB per engine\n", dv->ordinal,
                   (unsigned long long)parts_on_gpu, (double)table / 1e9, (double)free_b / 1e9, (double)budget / 1e9);
    }
    return NK_OK;
}

static void nk_seed_finish_task(int d, void *a)
{
    nk_ctx *c = a;
    c->dev[d].rc = c->dev[d].lead == d ? nkd_seed_finish(c->dev[d].eng) : NK_OK;
}

static void nk_seed_adopt_task(int d, void *a)
{
    nk_ctx *c = a;
    c->dev[d].rc = c->dev[d].lead == d ? NK_OK : nkd_seed_finish_from(c->dev[d].eng, c->dev[c->dev[d].lead].eng);
}

int nk_seed_finish(nk_ctx *c)
{
    if (c->seeded)
        return nk_fail(c, NK_EINVAL, "nk_seed_finish called twice");
    double t0 = nk_now();
    if (c->cfg.dump_tables && c->part_first == 0)
    { /* C:2251-2252 */
        nkd_part_stats st;
        nkd_seed_stats(c->dev[0].eng, &st);
        int rc = nk_write_dump(c, c->dev[0].eng, NKD_PART_SEED, st.capacity, "_seeds", -1);
        if (rc)
            return rc;
    }
    int brc = nk_set_table_budgets(c);
    if (brc)
        return brc;
    nk_parallel_for(c->n_dev, c->n_dev, nk_seed_adopt_task, c); /* while the lead engines still hold the seed table */
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].rc)
            return nk_fail(c, c->dev[d].rc, "%s", nkd_last_error(c->dev[d].eng));
    nk_parallel_for(c->n_dev, c->n_dev, nk_seed_finish_task, c);
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].rc)
            return nk_fail(c, c->dev[d].rc, "%s", nkd_last_error(c->dev[d].eng));
    int orc = nk_open_outputs(c);
    if (orc)
        return orc;
    c->seeded = 1;
    c->tot.seed_seconds += nk_now() - t0;
    return NK_OK;
}

/* ------------------------------------------------------------------ indexer */

typedef struct
{
    nk_ctx *c;
    nk_dev *dv;
    nk_stepbuf *sb;
} nk_step_job;

/* Fill one partition's share of a step: the worker loop's record reader (C:1605-1631) up to the
 * step's record/byte/operation budget. */
static void nk_index_task(int li, void *a)
{
    nk_step_job *j = a;
    nk_ctx *c = j->c;
    nk_part *p = &c->part[j->dv->parts[li]];
    nk_pstep *ps = &j->sb->ps[li];
    uint8_t *seqbuf = j->sb->seq;
    const int per = c->cfg.in_fastq ? 4 : 2, k = c->cfg.k, paired = c->paired;
    const nk_buf *ff = &c->ff, *rf = &c->rf;
    nk_cursor *cur = &p->cur;
    ps->n_reads = ps->n_records = 0;
    ps->ops = 0;
    ps->seq_hi = ps->seq_lo;
    size_t pos = ps->seq_lo;
    while (p->active && !cur->done && cur->fp < cur->fe && (!paired || cur->rp < cur->re))
    {
        if (ps->n_records >= c->step_pairs || pos + 2 * NK_MAX_LINE > ps->seq_end || ps->ops + 2 * NK_MAX_LINE > c->step_ops)
            break;
        nk_span sf = {cur->fp, 0, 0, 0, 1}, sr = {cur->rp, 0, 0, 0, 1};
        size_t fp = cur->fp, rp = cur->rp;
        int more = 1, complete = 1, plain_f = 1, plain_r = 1, lines_read = per;
        /* fast path: every line of the record ends in '\n' within 1023 chars and no NUL is near, so
         * read_line (C:394-409) consumes exactly "text\n" each time; positions come from the SIMD iterator */
        int fast = 1;
        {
            size_t q = fp;
            for (int i = 0; i < per && fast; i++)
            {
                size_t nl = nk_nliter_next(&cur->itf);
                if (nl == SIZE_MAX || nl - q >= (size_t)NK_MAX_LINE)
                    fast = 0;
                else
                {
                    if (i == 1)
                    {
                        sf.seq_rel = (uint16_t)(q - cur->fp);
                        sf.seq_len = (uint16_t)(nl - q);
                    }
                    q = nl + 1;
                }
            }
            if (fast && !cur->itf.nul_seen)
                fp = q;
            else
                fast = 0;
            if (paired && fast)
            {
                q = rp;
                for (int i = 0; i < per && fast; i++)
                {
                    size_t nl = nk_nliter_next(&cur->itr);
                    if (nl == SIZE_MAX || nl - q >= (size_t)NK_MAX_LINE)
                        fast = 0;
                    else
                    {
                        if (i == 1)
                        {
                            sr.seq_rel = (uint16_t)(q - cur->rp);
                            sr.seq_len = (uint16_t)(nl - q);
                        }
                        q = nl + 1;
                    }
                }
                if (fast && !cur->itr.nul_seen)
                    rp = q;
                else
                    fast = 0;
            }
        }
        if (fast)
            more = nk_at(ff, fp) != '\0' && (!paired || nk_at(rf, rp) != '\0');
        else
        { /* anything unusual: the byte-exact line reader, then re-aim the iterators */
            fp = cur->fp;
            rp = cur->rp;
            for (int i = 0; i < per; i++)
            {
                uint32_t lf = 0, lr = 0;
                int mf = 1, mr = 1;
                size_t f0 = fp, r0 = rp;
                fp = nk_take_line(ff, fp, &lf, &mf, &plain_f);
                if (paired)
                    rp = nk_take_line(rf, rp, &lr, &mr, &plain_r);
                if (i == 1)
                {
                    sf.seq_rel = (uint16_t)(f0 - cur->fp);
                    sf.seq_len = (uint16_t)lf;
                    sr.seq_rel = (uint16_t)(r0 - cur->rp);
                    sr.seq_len = (uint16_t)lr;
                }
                if (!mf || !mr)
                { /* read_line gave up (a NUL byte, the end of the file): with both sequence lines in, the reference
                   * still scores and counts the record before it stops (C:1629, C:1733) */
                    more = 0;
                    complete = (i >= 1);
                    if (i < per - 1)
                    { /* printed line by line, as far as it was read (both mates stop at the same line) */
                        plain_f = plain_r = 0;
                        lines_read = i + 1;
                    }
                    break;
                }
            }
            nk_nliter_seek(&cur->itf, ff->data, fp, ff->size);
            if (paired)
                nk_nliter_seek(&cur->itr, rf->data, rp, rf->size);
        }
        if (!complete)
        { /* cut before its sequence lines: not scored (the reference scores stale stack bytes here, C:1616-1629) */
            cur->done = 1;
            break;
        }
        sf.nbytes = (uint32_t)(fp - cur->fp);
        sr.nbytes = (uint32_t)(rp - cur->rp);
        sf.plain = (uint32_t)plain_f | ((uint32_t)lines_read << 8);
        sr.plain = (uint32_t)plain_r | ((uint32_t)lines_read << 8);
        cur->fp = fp;
        cur->rp = rp;
        if (!more)
            cur->done = 1;
        if ((int)sf.seq_len < k || (paired && (int)sr.seq_len < k))
            continue; /* dropped silently: no counter moves, the table is untouched, C:1430-1443 */
        for (int m = 0; m < (paired ? 2 : 1); m++)
        {
            const nk_span *s = m ? &sr : &sf;
            const nk_buf *src = m ? rf : ff;
            size_t need = ((size_t)s->seq_len + 15) & ~(size_t)15;
            memcpy(seqbuf + pos, src->data + s->start + s->seq_rel, s->seq_len);
            memset(seqbuf + pos + s->seq_len, 0, need - s->seq_len);
            nkd_read *rd = &ps->reads[ps->n_reads];
            rd->seq_off = (uint32_t)pos;
            rd->op_base = ps->ops;
            rd->len = s->seq_len;
            rd->part = (uint16_t)p->lidx;
            rd->reserved = 0;
            ps->spans[ps->n_reads] = *s;
            ps->n_reads++;
            ps->ops += (uint32_t)(s->seq_len - k + 1);
            pos += need;
        }
        ps->n_records++;
    }
    ps->seq_hi = pos;
}

/* ------------------------------------------------------------------ writer (C:1649-1666, C:852-876) */

static void nk_emit_lines(FILE *o, const nk_buf *f, const nk_span *s, int per, char *tmp)
{
    if (s->plain & 1u)
    { /* verbatim bytes, N -> A inside the sequence line (C:1426-1427 mutates the buffer that is printed) */
        const char *src = f->data + s->start;
        const char *seq = src + s->seq_rel;
        if (!memchr(seq, 'N', s->seq_len))
        {
            fwrite(src, 1, s->nbytes, o);
            return;
        }
        memcpy(tmp, src, s->nbytes);
        for (uint32_t i = 0; i < s->seq_len; i++)
            if (tmp[s->seq_rel + i] == 'N')
                tmp[s->seq_rel + i] = 'A';
        fwrite(tmp, 1, s->nbytes, o);
        return;
    }
    size_t pos = s->start; /* cut, unterminated or NUL-holding lines: print each as the reference's "%s\n" */
    if ((int)(s->plain >> 8) < per)
        per = (int)(s->plain >> 8); /* the record reader stopped early: only the lines it read are printed */
    for (int i = 0; i < per; i++)
    {
        uint32_t len;
        int more, plain = 1;
        size_t p0 = pos;
        pos = nk_take_line(f, pos, &len, &more, &plain);
        memcpy(tmp, f->data + p0, len);
        if (i == 1)
            for (uint32_t b = 0; b < len; b++)
                if (tmp[b] == 'N')
                    tmp[b] = 'A';
        tmp[len] = '\n';
        fwrite(tmp, 1, len + 1, o);
        if (!more)
            break;
    }
}

static void nk_emit_fasta(FILE *o, const nk_buf *f, const nk_span *s, int fwd, char *tmp)
{ /* fastq_to_fasta, C:852-876 */
    uint32_t hlen;
    int more, plain = 1;
    nk_take_line(f, s->start, &hlen, &more, &plain);
    const char *hdr = f->data + s->start;
    size_t n = 0;
    tmp[n++] = '>';
    if (hlen > 1)
    {
        memcpy(tmp + n, hdr + 1, hlen - 1);
        n += hlen - 1;
    }
    const char a = '/', b = fwd ? '1' : '2';
    if (hlen < 2 || hdr[hlen - 2] != a || hdr[hlen - 1] != b)
    {
        tmp[n++] = a;
        tmp[n++] = b;
    }
    tmp[n++] = '\n';
    const char *seq = f->data + s->start + s->seq_rel;
    for (uint32_t i = 0; i < s->seq_len; i++)
        tmp[n++] = seq[i] == 'N' ? 'A' : seq[i];
    tmp[n++] = '\n';
    fwrite(tmp, 1, n, o);
}

typedef struct
{
    nk_ctx *c;
    nk_dev *dv;
    nk_stepbuf *sb;
    size_t *rec_base; /* first record of each partition inside the step's accept array */
} nk_write_job;

/* one task per (partition, mate): the forward and reverse outputs of a partition are independent files */
static void nk_write_task(int idx, void *a)
{
    nk_write_job *j = a;
    nk_ctx *c = j->c;
    const int paired = c->paired, stride = paired ? 2 : 1;
    const int li = idx / stride, mate = idx % stride;
    nk_part *p = &c->part[j->dv->parts[li]];
    nk_pstep *ps = &j->sb->ps[li];
    const uint8_t *acc = j->sb->accept + j->rec_base[li];
    const int per = c->cfg.in_fastq ? 4 : 2;
    const int to_fasta = c->cfg.in_fastq && !c->cfg.out_fastq;
    FILE *out = mate ? p->out_r : p->out_f;
    const nk_buf *src = mate ? &c->rf : &c->ff;
    char tmp[4 * NK_MAX_LINE + 16];
    size_t nrec = ps->n_records;
    if (ps->fatal_record >= 0 && (size_t)ps->fatal_record < nrec)
        nrec = (size_t)ps->fatal_record; /* the reference stops at the first non-DNA record */
    uint64_t printed = 0;
    for (size_t r = 0; r < nrec; r++)
    {
        if (!acc[r])
            continue;
        printed++;
        const nk_span *sp = &ps->spans[r * (size_t)stride + (size_t)mate];
        if (to_fasta)
        {
            if (paired) /* single-end fq->fa prints nothing although it counts as printed, C:1995-1999 */
                nk_emit_fasta(out, src, sp, mate == 0, tmp);
        }
        else
            nk_emit_lines(out, src, sp, per, tmp);
    }
    if (mate == 0)
    {
        p->processed += nrec;
        p->printed += printed;
        p->skipped += nrec - printed;
    }
}

/* ------------------------------------------------------------------ per-GPU pipeline */

typedef struct
{
    nk_ctx *c;
    nk_dev *dv;
    /* three stages over NK_NBUF staging buffers: indexer (this thread) -> GPU thread -> writer thread */
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int built, completed, written; /* steps finished by each stage */
    int total;                     /* number of steps once the indexer ran dry, else -1 */
    int abort_rc;                  /* first error of any stage */
    int64_t first_invalid[NK_NBUF];
    int t_index, t_write;
    int raw; /* 1: steps are raw record text parsed on the device; 0: parsed here */
    /* raw text is sent ahead of its step (nkd_upload_raw) by whichever of the two threads sees the chance first:
     * the GPU stage right after it has staged the step before, or the builder when it finishes a step later than
     * that.  staged = steps whose nkd_stage_raw has returned; uploaded = steps whose upload has been claimed */
    int staged, uploaded, prefetch;
} nk_pipe;

/* step `u` is built and the step before it has been staged (its device buffer is the free one): send it ahead */
static void nk_try_upload(nk_pipe *pp, int u)
{
    if (!pp->raw || !pp->prefetch)
        return;
    pthread_mutex_lock(&pp->mu);
    int go = pp->built > u && pp->staged == u && pp->uploaded < u && !pp->abort_rc;
    if (go)
        pp->uploaded = u;
    pthread_mutex_unlock(&pp->mu);
    if (go)
    {
        nk_stepbuf *sb = &pp->dv->sb[u % NK_NBUF];
        if (sb->raw_bytes)
            nkd_upload_raw(pp->dv->eng, sb->raw, sb->raw_bytes);
    }
}

static int nk_gpu_step(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int64_t *first_invalid)
{
    /* One host step holds a batch of every resident partition; the device takes it `group` partitions at a
     * time: partitions are independent, and fewer tables per launch means a larger share of each table's
     * hot k-mers stays in the 126 MB L2 while the launch runs. */
    int group = c->dev_group > 0 && c->dev_group < dv->n_parts ? c->dev_group : dv->n_parts;
    sb->n_records = 0;
    *first_invalid = -1;
    int rc = NK_OK;
    double t0 = nk_now();
    for (int i0 = 0; i0 < dv->n_parts && !rc; i0 += group)
    {
        int nseg = 0;
        size_t recs = 0;
        for (int i = i0; i < dv->n_parts && i < i0 + group; i++)
        {
            nk_pstep *ps = &sb->ps[i];
            sb->segs[nseg].reads = ps->reads;
            sb->segs[nseg].n_reads = ps->n_reads;
            sb->segs[nseg].seq_lo = ps->seq_lo;
            sb->segs[nseg].seq_hi = ps->seq_hi;
            sb->segs[nseg].trusted = 1; /* built by nk_index_task under the nkd_read rules */
            sb->segs[nseg].part = (uint32_t)i;
            sb->segs[nseg].ops = ps->ops;
            nseg++;
            recs += ps->n_records;
        }
        int64_t inv = -1;
        if (recs)
        {
            /* experiment knob: with serial_steps the engines of a GPU take turns, so that every kernel has the GPU
             * to itself (clean per-kernel times) while the host stages of the pipelines still interleave */
            pthread_mutex_t *turn = c->serial_steps ? &c->dev[dv->lead].step_lock : NULL;
            if (turn)
                pthread_mutex_lock(turn);
            rc = nkd_stage_segments(dv->eng, sb->seq, sb->segs, nseg, c->paired);
            if (!rc)
                rc = nkd_run(dv->eng);
            if (!rc)
                rc = nkd_fetch(dv->eng, sb->accept + sb->n_records, recs, &inv);
            if (turn)
                pthread_mutex_unlock(turn);
        }
        if (inv >= 0 && *first_invalid < 0)
            *first_invalid = (int64_t)sb->n_records + inv;
        sb->n_records += recs;
    }
    dv->device_s += nk_now() - t0;
    if (rc)
        snprintf(dv->err, sizeof dv->err, "%s", nkd_last_error(dv->eng));
    return rc;
}

/* ------------------------------------------------------------------ raw-text steps
 *
 * The device reads the records itself (nkd_stage_raw): the host cuts each partition's byte range into steps of
 * whole records with the line index, copies the bytes into the page-locked step buffer on all pool threads, and
 * later write()s the text of the accepted records that comes back.  What the reference's worker does per record
 * (C:1605-1674) happens on the GPU; the byte-exact host parser above remains for text the device declines. */

/* file offset of the first byte of record j of a partition (j >= 1): one past its (per*j)-th line end */
static size_t nk_record_start(const nk_lineidx *li, uint64_t line0, int per, uint64_t j)
{
    size_t nl = nk_lineidx_find(li, line0 + (uint64_t)per * j - 1);
    return nl == SIZE_MAX ? SIZE_MAX : nl + 1;
}

/* whole records the raw-text pipeline may take from a byte range: those that start before `end` (the worker
 * loop's test, C:1605) and whose every line is terminated inside the file */
static uint64_t nk_raw_records_in(const nk_lineidx *li, size_t start, size_t end, int per, uint64_t line0)
{
    if (start >= end)
        return 0;
    uint64_t upto = nk_lineidx_before(li, end - 1) - line0; /* line ends in [start, end-1) */
    uint64_t n = 1 + upto / (uint64_t)per;
    uint64_t have = nk_lineidx_total(li) - line0; /* line ends from start to the end of the counted range */
    if (n > have / (uint64_t)per)
        n = have / (uint64_t)per;
    return n;
}

/* The staging buffer is written once and next read by the GPU's copy engine: streaming stores keep it out of the
 * caches and spare the read-for-ownership of every destination line (the host side of this path is bound by
 * memory bandwidth, profiles/r02_host_traffic.md). */
__attribute__((target("avx2"))) static void nk_copy_stream_avx2(uint8_t *dst, const char *src, size_t n)
{
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > n)
        head = n;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128)
    {
        __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 32), b);
        _mm256_stream_si256((__m256i *)(dst + i + 64), c);
        _mm256_stream_si256((__m256i *)(dst + i + 96), d);
    }
    _mm_sfence();
    memcpy(dst + i, src + i, n - i);
}

static void nk_copy_task(int i, void *a)
{
    const nk_copy *cp = &((const nk_copy *)a)[i];
    if (nk_have_avx2 > 0)
        nk_copy_stream_avx2(cp->dst, cp->src, cp->n);
    else
        memcpy(cp->dst, cp->src, cp->n);
}

#define NK_RAW_PLENTY ((uint64_t)1 << 62) /* raw_total of a partition whose range has not been counted to its end */
static const size_t nk_range_slack = 8u * NK_MAX_LINE + 2; /* the record that starts before a range's end runs past it */

/* A partition with its own line indexes: count far enough ahead of its position for one step of `byte_room` bytes
 * per file.  Near the end of the range the rest is counted and the number of whole records becomes exact (the same
 * rule as nk_raw_records_in on an index of the whole file).  Returns the records known to be available. */
static uint64_t nk_part_reach(nk_ctx *c, nk_part *p, int threads, uint64_t byte_room)
{
    const int per = c->cfg.in_fastq ? 4 : 2, paired = c->paired;
    if (p->raw_known)
        return p->raw_total - p->raw_next;
    /* chunk boundaries at least one chunk past a full window */
    size_t need_f = ((p->raw_fp + byte_room) / NK_LI_CHUNK + 2) * NK_LI_CHUNK,
           need_r = ((p->raw_rp + byte_room) / NK_LI_CHUNK + 2) * NK_LI_CHUNK;
    int full = need_f >= p->cur.fe || (paired && need_r >= p->cur.re);
    if (full)
    {
        nk_lineidx_extend(p->lf, threads, p->cur.fe + nk_range_slack);
        uint64_t n = nk_raw_records_in(p->lf, p->cur.fp, p->cur.fe, per, p->line_f);
        if (paired)
        {
            nk_lineidx_extend(p->lr, threads, p->cur.re + nk_range_slack);
            uint64_t nr = nk_raw_records_in(p->lr, p->cur.rp, p->cur.re, per, p->line_r);
            n = nr < n ? nr : n;
        }
        p->raw_total = n > p->raw_next ? n : p->raw_next;
        p->raw_known = 1;
        return p->raw_total - p->raw_next;
    }
    /* everything counted lies before the ends of the ranges: every whole record in it belongs to the partition */
    nk_lineidx_extend(p->lf, threads, need_f);
    uint64_t n = (nk_lineidx_total(p->lf) - p->line_f) / (uint64_t)per;
    if (paired)
    {
        nk_lineidx_extend(p->lr, threads, need_r);
        uint64_t nr = (nk_lineidx_total(p->lr) - p->line_r) / (uint64_t)per;
        n = nr < n ? nr : n;
    }
    return n > p->raw_next ? n - p->raw_next : 0;
}

/* One step: for every partition of the engine the next (at most step_pairs) records of both files, as two windows
 * of the step buffer.  Returns the number of records staged. */
static size_t nk_build_step_raw(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int threads)
{
    double t0 = nk_now();
    const int per = c->cfg.in_fastq ? 4 : 2, paired = c->paired;
    size_t at = 0, total = 0;
    sb->n_rsegs = 0;
    sb->n_copies = 0;
    /* the step's room (records, bytes, operations) is shared by the partitions that still have records: when only a
     * wave of them is being worked on, or most have run dry, each takes a larger share and the launches stay large */
    int busy = 0;
    char ready[NK_MAX_PARTITIONS];
    memset(ready, 1, sizeof ready);
    if (c->rolling)
    { /* partitions whose reverse range is still being located (nk_roll_thread) join the steps as they are released;
       * with nothing else to do, wait for the next one */
        pthread_mutex_lock(&c->roll_mu);
        for (;;)
        {
            int work = 0, pending = 0;
            for (int li = 0; li < dv->n_parts && li < NK_MAX_PARTITIONS; li++)
            {
                nk_part *p = &c->part[dv->parts[li]];
                ready[li] = (char)p->ready;
                if (!p->active)
                    continue;
                if (!p->ready)
                    pending = 1;
                else if (p->raw_total > p->raw_next)
                    work = 1;
            }
            if (work || !pending)
                break;
            pthread_cond_wait(&c->roll_cv, &c->roll_mu);
        }
        pthread_mutex_unlock(&c->roll_mu);
    }
    for (int li = 0; li < dv->n_parts; li++)
    {
        nk_part *p = &c->part[dv->parts[li]];
        busy += ready[li] && p->active && p->raw_total > p->raw_next;
    }
    uint64_t share = busy ? (uint64_t)dv->n_parts / (uint64_t)busy : 1;
    if (share * c->step_pairs > 262144)
        share = 262144 / c->step_pairs ? 262144 / c->step_pairs : 1; /* operations per partition and step stay below 2^28 */
    const uint64_t quota = (uint64_t)c->step_pairs * share, byte_room = (uint64_t)c->raw_part_bytes * share,
                   op_room = (uint64_t)c->step_ops * share;
    for (int li = 0; li < dv->n_parts; li++)
    {
        nk_part *p = &c->part[dv->parts[li]];
        nk_pstep *ps = &sb->ps[li];
        ps->raw_n = 0;
        ps->fatal_record = -1;
        uint64_t left = ready[li] && p->active && p->raw_total > p->raw_next ? nk_part_reach(c, p, threads, byte_room) : 0;
        if (!left)
        {
            if (ready[li] && p->active && !p->raw_known)
                p->raw_total = p->raw_next; /* not one whole record within reach: not regular text */
            continue;
        }
        uint64_t n = left < quota ? left : quota;
        size_t f1 = 0, r1 = 0;
        for (;;)
        {
            f1 = nk_record_start(p->lf, p->line_f, per, p->raw_next + n);
            r1 = paired ? nk_record_start(p->lr, p->line_r, per, p->raw_next + n) : 0;
            if (f1 == SIZE_MAX || r1 == SIZE_MAX)
            { /* cannot happen for records counted by nk_raw_records_in; leave the rest to the host parser */
                n = 0;
                break;
            }
            size_t wf = f1 - p->raw_fp, wr = paired ? r1 - p->raw_rp : 0, big = wf > wr ? wf : wr;
            size_t ops_bound = c->cfg.in_fastq ? (wf + wr) / 2 : wf + wr;
            if ((big <= byte_room && ops_bound <= op_room) || n == 1)
            {
                if (big > byte_room)
                    n = 0; /* one record larger than the window: not regular text */
                break;
            }
            double shrink = 0.95 * (double)byte_room / (double)big, s2 = 0.95 * (double)op_room / (double)ops_bound;
            if (s2 < shrink)
                shrink = s2;
            uint64_t m = (uint64_t)((double)n * shrink);
            n = m >= n ? n - 1 : (m < 1 ? 1 : m);
        }
        if (n == 0)
        {
            p->raw_total = p->raw_next;
            continue;
        }
        ps->raw_n = (uint32_t)n;
        ps->f0 = p->raw_fp;
        ps->f1 = f1;
        ps->r0 = p->raw_rp;
        ps->r1 = r1;
        nkd_raw_segment *g = &sb->rsegs[sb->n_rsegs];
        sb->rseg_li[sb->n_rsegs++] = li;
        g->part = (uint32_t)p->lidx;
        g->n_records = (uint32_t)n;
        for (int m = 0; m < (paired ? 2 : 1); m++)
        {
            const nk_buf *src = m ? &c->rf : &c->ff;
            size_t lo = m ? ps->r0 : ps->f0, hi = m ? ps->r1 : ps->f1;
            if (m)
            {
                g->rev_off = (uint32_t)at;
                g->rev_bytes = (uint32_t)(hi - lo);
            }
            else
            {
                g->fwd_off = (uint32_t)at;
                g->fwd_bytes = (uint32_t)(hi - lo);
            }
            for (size_t o = lo; o < hi; o += NK_COPY_PIECE)
            {
                nk_copy *cp = &sb->copies[sb->n_copies++];
                cp->dst = sb->raw + at + (o - lo);
                cp->src = src->data + o;
                cp->n = hi - o < NK_COPY_PIECE ? hi - o : NK_COPY_PIECE;
            }
            at += hi - lo;
            size_t pad = (16 - (at & 15)) & 15;
            memset(sb->raw + at, ' ', pad); /* neither a line end nor a NUL */
            at += pad;
        }
        if (!paired)
            g->rev_off = g->rev_bytes = 0;
        p->raw_next += n;
        p->raw_fp = f1;
        p->raw_rp = r1;
        total += n;
    }
    sb->raw_bytes = at;
    sb->n_records = total;
    double t1 = nk_now();
    if (sb->n_copies)
        nk_parallel_for(sb->n_copies, threads, nk_copy_task, sb->copies);
    dv->index_s += nk_now() - t0;
    nk_trace((int)(dv - c->dev), -1, "cut", t0, t1);
    nk_trace((int)(dv - c->dev), -1, "copy", t1, nk_now());
    return total;
}

static int nk_emit_mode(const nk_ctx *c)
{
    if (c->cfg.in_fastq && !c->cfg.out_fastq)
        return c->paired ? 1 : 2; /* single-end fq->fa prints nothing although it counts as printed, C:1995-1999 */
    return 0;
}

static int nk_gpu_step_raw(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, nk_pipe *pp, int step, int64_t *first_invalid)
{
    *first_invalid = -1;
    double t0 = nk_now();
    pthread_mutex_t *turn = c->serial_steps ? &c->dev[dv->lead].step_lock : NULL;
    if (turn)
        pthread_mutex_lock(turn);
    int rc = nkd_stage_raw(dv->eng, sb->raw, sb->raw_bytes, sb->rsegs, sb->n_rsegs, c->paired, c->cfg.in_fastq ? 4 : 2);
    if (!rc && pp)
    { /* the following step's bytes travel while this one runs */
        pthread_mutex_lock(&pp->mu);
        pp->staged = step + 1;
        pthread_mutex_unlock(&pp->mu);
        nk_try_upload(pp, step + 1);
    }
    double t1 = nk_now();
    if (!rc)
        rc = nkd_run(dv->eng);
    double t2 = nk_now();
    if (!rc)
        rc = nkd_fetch_raw_slot(dv->eng, nk_emit_mode(c), sb->out, (size_t)-1, sb->rres, first_invalid, (int)(sb - dv->sb));
    double t3 = nk_now();
    dv->t_stage += t1 - t0;
    dv->t_run += t2 - t1;
    dv->t_fetch += t3 - t2;
    int label = (int)__atomic_load_n(&c->raw_steps, __ATOMIC_RELAXED); /* other engines' device threads count too */
    nk_trace((int)(dv - c->dev), label, "stage", t0, t1);
    nk_trace((int)(dv - c->dev), label, "run", t1, t2);
    nk_trace((int)(dv - c->dev), label, "fetch", t2, t3);
    if (turn)
        pthread_mutex_unlock(turn);
    dv->device_s += nk_now() - t0;
    if (rc)
        snprintf(dv->err, sizeof dv->err, "%s", nkd_last_error(dv->eng));
    else
        for (int s = 0; s < sb->n_rsegs; s++)
        { /* this step is done: the host parser, if it is ever needed, continues behind it */
            nk_pstep *ps = &sb->ps[sb->rseg_li[s]];
            nk_part *p = &c->part[dv->parts[sb->rseg_li[s]]];
            p->commit_fp = ps->f1;
            p->commit_rp = ps->r1;
        }
    return rc;
}

/* the reference's FATAL text for record `rec` of a raw-text step (C:1445-1454) */
static int nk_report_invalid_raw(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int64_t rec)
{
    const int per = c->cfg.in_fastq ? 4 : 2;
    size_t base = 0;
    for (int s = 0; s < sb->n_rsegs; s++)
    {
        nk_pstep *ps = &sb->ps[sb->rseg_li[s]];
        if ((size_t)rec >= base + ps->raw_n)
        {
            base += ps->raw_n;
            continue;
        }
        size_t r = (size_t)rec - base;
        for (int m = 0; m < (c->paired ? 2 : 1); m++)
        {
            const nk_buf *f = m ? &c->rf : &c->ff;
            size_t lo = m ? ps->r0 : ps->f0, hi = m ? ps->r1 : ps->f1;
            size_t q = r ? lo + nk_kth_newline(f->data + lo, hi - lo, (uint64_t)per * r) + 1 : lo;
            q += nk_kth_newline(f->data + q, hi - q, 1) + 1; /* the sequence line follows the header */
            size_t len = nk_kth_newline(f->data + q, hi - q, 1);
            int bad = 0;
            for (size_t b = 0; b < len; b++)
                if (!strchr("ACGTN", f->data[q + b]) || f->data[q + b] == 0)
                    bad = 1;
            if (bad)
            {
                char *txt = malloc(len + 1);
                nk_scrub_copy(txt, f->data + q, len);
                snprintf(dv->err, sizeof dv->err, "FATAL: %s sequence does not appear to be a DNA sequence\n%s\n", m ? "REV" : "FWD", txt);
                free(txt);
                return NK_EDATA;
            }
        }
        snprintf(dv->err, sizeof dv->err, "FATAL: sequence does not appear to be a DNA sequence");
        return NK_EDATA;
    }
    return NK_EINTERNAL;
}

typedef struct
{
    nk_ctx *c;
    nk_dev *dv;
    nk_stepbuf *sb;
    int io_error;
} nk_rawwrite_job;

/* one task per (segment, mate): the accepted records' text goes to the partition's file in one piece */
static void nk_rawwrite_task(int idx, void *a)
{
    nk_rawwrite_job *j = a;
    const int stride = j->c->paired ? 2 : 1, s = idx / stride, mate = idx % stride;
    nk_part *p = &j->c->part[j->dv->parts[j->sb->rseg_li[s]]];
    const nkd_raw_result *r = &j->sb->rres[s];
    FILE *out = mate ? p->out_r : p->out_f;
    uint64_t off = mate ? r->rev_off : r->fwd_off, n = mate ? r->rev_bytes : r->fwd_bytes;
    if (n && out && fwrite(j->sb->out + off, 1, (size_t)n, out) != (size_t)n)
        j->io_error = 1;
    if (mate == 0)
    {
        p->processed += r->processed;
        p->printed += r->printed;
        p->skipped += r->processed - r->printed;
    }
}

static int nk_write_step_raw(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int threads)
{
    double t0 = nk_now();
    if (nkd_fetch_wait(dv->eng, (int)(sb - dv->sb)) != NK_OK) /* the text is still arriving on the engine's copy stream */
    {
        snprintf(dv->err, sizeof dv->err, "%s", nkd_last_error(dv->eng));
        return NK_ENODEVICE;
    }
    double t1 = nk_now();
    nk_rawwrite_job job = {c, dv, sb, 0};
    nk_parallel_for(sb->n_rsegs * (c->paired ? 2 : 1), threads, nk_rawwrite_task, &job);
    dv->write_s += nk_now() - t0;
    nk_trace((int)(dv - c->dev), -1, "d2h_wait", t0, t1);
    nk_trace((int)(dv - c->dev), -1, "write", t1, nk_now());
    if (job.io_error)
    {
        snprintf(dv->err, sizeof dv->err, "error writing the output files: %s", strerror(errno));
        return NK_EIO;
    }
    return NK_OK;
}

static void nk_pipe_fail(nk_pipe *pp, int rc)
{
    pthread_mutex_lock(&pp->mu);
    if (!pp->abort_rc)
        pp->abort_rc = rc;
    pthread_cond_broadcast(&pp->cv);
    pthread_mutex_unlock(&pp->mu);
}

static void *nk_gpu_thread(void *a)
{
    nk_pipe *pp = a;
    for (int step = 0;; step++)
    {
        pthread_mutex_lock(&pp->mu);
        while (pp->built <= step && pp->total < 0 && !pp->abort_rc)
            pthread_cond_wait(&pp->cv, &pp->mu);
        int stop = pp->abort_rc || (pp->built <= step);
        pthread_mutex_unlock(&pp->mu);
        if (stop)
            break;
        nk_stepbuf *sb = &pp->dv->sb[step % NK_NBUF];
        int rc = pp->raw ? nk_gpu_step_raw(pp->c, pp->dv, sb, pp, step, &pp->first_invalid[step % NK_NBUF])
                         : nk_gpu_step(pp->c, pp->dv, sb, &pp->first_invalid[step % NK_NBUF]);
        if (rc)
        {
            nk_pipe_fail(pp, rc);
            break;
        }
        __atomic_add_fetch(pp->raw ? &pp->c->raw_steps : &pp->c->parsed_steps, 1, __ATOMIC_RELAXED);
        pthread_mutex_lock(&pp->mu);
        pp->completed = step + 1;
        pthread_cond_broadcast(&pp->cv);
        pthread_mutex_unlock(&pp->mu);
    }
    return NULL;
}

static size_t nk_build_step(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int threads)
{
    double t0 = nk_now();
    nk_step_job job = {c, dv, sb};
    nk_parallel_for(dv->n_parts, threads, nk_index_task, &job);
    size_t n = 0;
    for (int i = 0; i < dv->n_parts; i++)
    {
        sb->ps[i].fatal_record = -1;
        n += sb->ps[i].n_reads;
    }
    dv->index_s += nk_now() - t0;
    return n;
}

/* turn the engine's "first invalid record" into the reference's FATAL text (C:1445-1454) */
static int nk_report_invalid(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int64_t rec)
{
    size_t base = 0;
    for (int i = 0; i < dv->n_parts; i++)
    {
        nk_pstep *ps = &sb->ps[i];
        if ((size_t)rec < base + ps->n_records)
        {
            size_t r = (size_t)rec - base;
            ps->fatal_record = (int64_t)r;
            int stride = c->paired ? 2 : 1;
            for (int m = 0; m < stride; m++)
            {
                const nk_span *s = &ps->spans[r * (size_t)stride + (size_t)m];
                const nk_buf *f = m ? &c->rf : &c->ff;
                const char *q = f->data + s->start + s->seq_rel;
                int bad = 0;
                for (uint32_t b = 0; b < s->seq_len; b++)
                    if (!strchr("ACGTN", q[b]) || q[b] == 0)
                        bad = 1;
                if (bad)
                {
                    char *txt = malloc((size_t)s->seq_len + 1);
                    nk_scrub_copy(txt, q, s->seq_len);
                    snprintf(dv->err, sizeof dv->err, "FATAL: %s sequence does not appear to be a DNA sequence\n%s\n", m ? "REV" : "FWD", txt);
                    free(txt);
                    return NK_EDATA;
                }
            }
            snprintf(dv->err, sizeof dv->err, "FATAL: sequence does not appear to be a DNA sequence");
            return NK_EDATA;
        }
        base += ps->n_records;
    }
    return NK_EINTERNAL;
}

static void nk_write_step(nk_ctx *c, nk_dev *dv, nk_stepbuf *sb, int threads)
{
    double t0 = nk_now();
    size_t rec_base[NK_MAX_PARTITIONS];
    size_t b = 0;
    for (int i = 0; i < dv->n_parts; i++)
    {
        rec_base[i] = b;
        b += sb->ps[i].n_records;
    }
    nk_write_job job = {c, dv, sb, rec_base};
    nk_parallel_for(dv->n_parts * (c->paired ? 2 : 1), threads, nk_write_task, &job);
    dv->write_s += nk_now() - t0;
}

static void *nk_writer_thread(void *a)
{
    nk_pipe *pp = a;
    nk_ctx *c = pp->c;
    nk_dev *dv = pp->dv;
    for (int step = 0;; step++)
    {
        pthread_mutex_lock(&pp->mu);
        while (pp->completed <= step && !(pp->total >= 0 && step >= pp->total) && !pp->abort_rc)
            pthread_cond_wait(&pp->cv, &pp->mu);
        int stop = pp->completed <= step; /* ran dry or aborted before this step finished on the GPU */
        pthread_mutex_unlock(&pp->mu);
        if (stop)
            break;
        nk_stepbuf *sb = &dv->sb[step % NK_NBUF];
        int rc = NK_OK;
        if (pp->first_invalid[step % NK_NBUF] >= 0)
            rc = pp->raw ? nk_report_invalid_raw(c, dv, sb, pp->first_invalid[step % NK_NBUF])
                         : nk_report_invalid(c, dv, sb, pp->first_invalid[step % NK_NBUF]);
        if (pp->raw)
        {
            int wrc = nk_write_step_raw(c, dv, sb, pp->t_write);
            if (!rc)
                rc = wrc;
        }
        else
            nk_write_step(c, dv, sb, pp->t_write);
        if (rc)
        {
            nk_pipe_fail(pp, rc);
            break;
        }
        pthread_mutex_lock(&pp->mu);
        pp->written = step + 1;
        pthread_cond_broadcast(&pp->cv);
        pthread_mutex_unlock(&pp->mu);
    }
    return NULL;
}

/* per-GPU pipeline: this thread indexes records into the next free staging buffer */
static void *nk_device_pipeline(void *a)
{
    nk_pipe *pp = a;
    nk_ctx *c = pp->c;
    nk_dev *dv = pp->dv;
    int threads = c->threads / c->n_dev;
    if (threads < 2)
        threads = 2;
    /* one indexing task per partition, one writing task per partition and mate: give the indexer a thread
     * per partition when there are enough cores and the writer the rest */
    pp->t_index = dv->n_parts < threads - 1 ? dv->n_parts : (threads * 2 + 2) / 3;
    if (pp->t_index >= threads)
        pp->t_index = threads - 1;
    pp->t_write = threads - pp->t_index;
    if (dv->n_parts >= threads - 1)
        /* many partitions per engine (-p 64): which stage is the heavier one depends on the share of records
         * that is kept, so both stages get all of this pipeline's threads and the OS balances them (20 M pairs
         * at -p 64: 1.36 -> 0.98 s) */
        pp->t_index = pp->t_write = threads;
    if (pp->raw)
    { /* copying in and writing out are plain byte moves: every pool thread this pipeline may use helps with both */
        pp->t_index = threads;
        pp->t_write = threads;
    }
    if (dv->rc)
        return NULL;
    if ((pp->raw ? nk_alloc_raw_bufs(c, dv) : nk_alloc_parsed_bufs(c, dv)) != NK_OK)
    {
        dv->rc = NK_ENOMEM;
        snprintf(dv->err, sizeof dv->err, "Memory allocation failed (staging buffers)");
        return NULL;
    }
    pthread_mutex_init(&pp->mu, NULL);
    pthread_cond_init(&pp->cv, NULL);
    pp->built = pp->completed = pp->written = 0;
    pp->staged = 0;
    pp->uploaded = 0; /* step 0 is copied by its own nkd_stage_raw */
    pp->prefetch = !nk_env_on("NKB200_NO_PREFETCH");
    pp->total = -1;
    pp->abort_rc = 0;
    pthread_t gth, wth;
    if (pthread_create(&gth, NULL, nk_gpu_thread, pp) != 0 || pthread_create(&wth, NULL, nk_writer_thread, pp) != 0)
    {
        dv->rc = NK_EINTERNAL;
        snprintf(dv->err, sizeof dv->err, "cannot start the device threads");
        return NULL;
    }
    for (int step = 0;; step++)
    {
        pthread_mutex_lock(&pp->mu);
        while (step - pp->written >= NK_NBUF && !pp->abort_rc)
            pthread_cond_wait(&pp->cv, &pp->mu);
        int stop = pp->abort_rc != 0;
        pthread_mutex_unlock(&pp->mu);
        if (stop)
            break;
        size_t n = pp->raw ? nk_build_step_raw(c, dv, &dv->sb[step % NK_NBUF], pp->t_index)
                           : nk_build_step(c, dv, &dv->sb[step % NK_NBUF], pp->t_index);
        pthread_mutex_lock(&pp->mu);
        if (n == 0)
            pp->total = step;
        else
            pp->built = step + 1;
        pthread_cond_broadcast(&pp->cv);
        pthread_mutex_unlock(&pp->mu);
        if (n == 0)
            break;
        nk_try_upload(pp, step);
    }
    pthread_mutex_lock(&pp->mu);
    if (pp->total < 0)
        pp->total = pp->built; /* aborted: let the other stages drain */
    pthread_cond_broadcast(&pp->cv);
    pthread_mutex_unlock(&pp->mu);
    pthread_join(gth, NULL);
    pthread_join(wth, NULL);
    if (pp->abort_rc && pp->abort_rc != NK_EIRREGULAR) /* declined raw text is not an error: the host parser continues */
        dv->rc = pp->abort_rc;
    pthread_mutex_destroy(&pp->mu);
    pthread_cond_destroy(&pp->cv);
    return NULL;
}

/* lf / lr: line indexes of the whole files.  Already built ones (cum != NULL) are used; if the record-count
 * partitioner needs them they are built here and left to the caller, who frees them. */
static int nk_plan(const nk_buf *ff, const nk_buf *rf, int paired, int P, int fastq, int threads, uint64_t *fs,
                   uint64_t *fe, uint64_t *rs, uint64_t *re, nk_diag *d, nk_lineidx *lf, nk_lineidx *lr, int *defer_rev)
{
    const int may_defer = defer_rev && *defer_rev;
    if (defer_rev)
        *defer_rev = 0;
    if (!nk_mask64)
        nk_mask64 = nk_mask64_pick();
    memset(fs, 0, sizeof(uint64_t) * (size_t)P);
    memset(fe, 0, sizeof(uint64_t) * (size_t)P);
    if (rs)
        memset(rs, 0, sizeof(uint64_t) * (size_t)P);
    if (re)
        memset(re, 0, sizeof(uint64_t) * (size_t)P);
    if (P == 1)
    { /* C:1796-1803 */
        fe[0] = ff->size - 1;
        if (paired)
            re[0] = rf->size - 1;
    }
    else if (!paired || ff->size == rf->size)
    { /* C:1807-1813, C:2142 */
        if (nk_ranges_by_size(ff, P, fastq, fs, fe, d) || (paired && nk_ranges_by_size(rf, P, fastq, rs, re, d)))
            return NK_EDATA;
    }
    else
    { /* C:1815-1828: the forward file's record count is applied to both files */
        if (!lf->cum)
            nk_lineidx_build_n(lf, ff, threads);
        uint64_t recs = nk_records_from_lines(ff, nk_lineidx_total(lf), fastq);
        nk_ranges_by_records(lf, P, fastq, recs, fs, fe);
        if (may_defer && !lr->cum)
        { /* the caller counts the reverse file and places its boundaries as it goes (nk_roll_thread) */
            *defer_rev = 1;
            return NK_OK;
        }
        if (!lr->cum)
            nk_lineidx_build_n(lr, rf, threads);
        nk_ranges_by_records(lr, P, fastq, recs, rs, re);
    }
    return NK_OK;
}

int nk_plan_ranges(const char *fwd, size_t fwd_size, const char *rev, size_t rev_size, int partitions, int fastq,
                   int threads, uint64_t *fwd_starts, uint64_t *fwd_ends, uint64_t *rev_starts, uint64_t *rev_ends,
                   char *errbuf, size_t errbuf_size)
{
    if (partitions < 1 || partitions > NK_MAX_PARTITIONS || !fwd || fwd_size == 0 || (rev && rev_size == 0))
        return NK_EINVAL;
    nk_buf ff = {fwd, fwd_size}, rf = {rev, rev_size};
    nk_diag d = {{0}, 0};
    nk_lineidx lf = {0}, lr = {0};
    int rc = nk_plan(&ff, &rf, rev != NULL, partitions, fastq, threads > 0 ? threads : nk_host_threads(), fwd_starts,
                     fwd_ends, rev_starts, rev_ends, &d, &lf, &lr, NULL);
    nk_lineidx_free(&lf);
    nk_lineidx_free(&lr);
    if (rc && errbuf && errbuf_size)
        snprintf(errbuf, errbuf_size, "%s", d.msg);
    return rc;
}

/* ---- the reverse file counted while the first partitions are being worked on
 *
 * calculate_thread_positions_from_records (C:1265-1300) places partition t+1 of a file behind the `want`-th line end
 * counted from partition t's start, with `want` taken from the FORWARD file's record count (C:1815-1828): the forward
 * file must be counted to its end before any boundary is known, the reverse file need not be.  This thread counts it
 * front to back on the pool, places one boundary after the other exactly as nk_ranges_by_records would on a complete
 * index (a longer index gives the same answer for the same line), and releases each partition to its engine's step
 * builder as soon as its range is covered. */
#ifndef NK_ROLL_CHUNKS
#define NK_ROLL_CHUNKS 256 /* index chunks counted per round (64 MB); the emulation build of the tests uses a few */
#endif

static void nk_roll_release(nk_ctx *c, int gid)
{
    const int per = c->cfg.in_fastq ? 4 : 2;
    for (int i = 0; i < c->n_local; i++)
    {
        nk_part *p = &c->part[i];
        if (p->gid != gid)
            continue;
        p->cur.rp = c->rs[gid];
        p->cur.re = c->re[gid];
        p->commit_rp = p->raw_rp = p->cur.rp;
        uint64_t n = 0;
        if (p->cur.fp < p->cur.fe && p->cur.rp < p->cur.re)
        {
            p->line_r = nk_lineidx_before(&c->lir, p->cur.rp);
            n = nk_raw_records_in(p->lf, p->cur.fp, p->cur.fe, per, p->line_f);
            uint64_t nr = nk_raw_records_in(&c->lir, p->cur.rp, p->cur.re, per, p->line_r);
            n = nr < n ? nr : n;
        }
        pthread_mutex_lock(&c->roll_mu);
        p->raw_total = n;
        p->ready = 1;
        pthread_cond_broadcast(&c->roll_cv);
        pthread_mutex_unlock(&c->roll_mu);
    }
}

static void *nk_roll_thread(void *a)
{
    nk_ctx *c = a;
    double t0 = nk_now();
    nk_lineidx *li = &c->lir;
    const nk_buf *f = &c->rf;
    const int P = c->cfg.partitions, fastq = c->cfg.in_fastq;
    uint64_t recs = nk_records_from_lines(&c->ff, nk_lineidx_total(&c->lif), fastq);
    uint64_t per = recs / (uint64_t)P;
    uint64_t *st = c->rs, *en = c->re;
    const size_t round = (size_t)NK_ROLL_CHUNKS * NK_LI_CHUNK;
    if (P < 2 || per < 1 || f->size < 1)
    { /* nk_ranges_by_records leaves every range empty */
        for (int t = 0; t < P; t++)
            nk_roll_release(c, t);
        return NULL;
    }
    int want = (int)(fastq ? per * 4 : per * 2);
    st[0] = 0;
    en[P - 1] = f->size - 1;
    for (int t = 0; t < P - 1; t++)
    {
        size_t pos = SIZE_MAX;
        if (want > 0)
        {
            nk_lineidx_extend(li, c->threads, st[t] + 1); /* lines before st[t] (behind us, unless st[t] stayed 0) */
            uint64_t g = nk_lineidx_before(li, st[t]) + (uint64_t)want - 1;
            while (nk_lineidx_total(li) <= g && nk_lineidx_hi(li) < li->nchunks)
                nk_lineidx_extend(li, c->threads, (size_t)nk_lineidx_hi(li) * NK_LI_CHUNK + round);
            pos = nk_lineidx_find(li, g);
        }
        if (pos != SIZE_MAX)
        {
            en[t] = pos;
            st[t + 1] = pos + 1;
        }
        /* the record that starts before the end of the range runs past it */
        nk_lineidx_extend(li, c->threads, en[t] + nk_range_slack + 1);
        nk_roll_release(c, t);
    }
    nk_lineidx_extend(li, c->threads, f->size);
    nk_roll_release(c, P - 1);
    nk_trace(-1, -1, "count_rev", t0, nk_now());
    return NULL;
}

/* one pipeline per engine, raw-text or host-parsed steps; returns the first error */
static int nk_run_pipelines(nk_ctx *c, int raw)
{
    nk_pipe *pipes = calloc((size_t)c->n_dev, sizeof *pipes);
    pthread_t *th = calloc((size_t)c->n_dev, sizeof *th);
    char *started = calloc((size_t)c->n_dev, 1);
    if (!pipes || !th || !started)
    {
        free(pipes);
        free(th);
        free(started);
        return nk_fail(c, NK_ENOMEM, "Memory allocation failed (pipelines)");
    }
    for (int dd = 0; dd < c->n_dev; dd++)
    {
        pipes[dd].c = c;
        pipes[dd].dv = &c->dev[dd];
        pipes[dd].raw = raw;
        c->dev[dd].rc = NK_OK;
    }
    for (int dd = 1; dd < c->n_dev; dd++)
    {
        if (pthread_create(&th[dd], NULL, nk_device_pipeline, &pipes[dd]) == 0)
            started[dd] = 1;
        else
        {
            c->dev[dd].rc = NK_EINTERNAL;
            snprintf(c->dev[dd].err, sizeof c->dev[dd].err, "cannot start the pipeline thread of engine %d", dd);
        }
    }
    nk_device_pipeline(&pipes[0]);
    for (int dd = 1; dd < c->n_dev; dd++)
        if (started[dd])
            pthread_join(th[dd], NULL);
    free(pipes);
    free(th);
    free(started);
    int rc = NK_OK;
    double mi = 0, md = 0, mw = 0;
    for (int dd = 0; dd < c->n_dev; dd++)
    {
        nk_dev *dv = &c->dev[dd];
        if (dv->rc && !rc)
            rc = nk_fail(c, dv->rc, "%s", dv->err);
        if (dv->index_s > mi)
            mi = dv->index_s;
        if (dv->device_s > md)
            md = dv->device_s;
        if (dv->write_s > mw)
            mw = dv->write_s;
        dv->index_s = dv->device_s = dv->write_s = 0;
    }
    c->tot.index_seconds += mi;
    c->tot.device_seconds += md;
    c->tot.write_seconds += mw;
    return rc;
}

static int nk_process(nk_ctx *c, const char *fwd, size_t fsize, const char *rev, size_t rsize, int paired,
                      const uint64_t *plan[4], const uint32_t *counts_f, const uint32_t *counts_r)
{
    if (!c->seeded)
        return nk_fail(c, NK_EINVAL, "nk_process_* before nk_seed_finish");
    if (fsize == 0 || (paired && rsize == 0))
        return nk_fail(c, NK_EIO, "Error memory mapping input files");
    double t0 = nk_now();
    int P = c->cfg.partitions, fastq = c->cfg.in_fastq, per = fastq ? 4 : 2;
    c->ff.data = fwd;
    c->ff.size = fsize;
    c->rf.data = rev;
    c->rf.size = rsize;
    c->paired = paired;
    if (!nk_mask64)
        nk_mask64 = nk_mask64_pick();
    memset(&c->lif, 0, sizeof c->lif);
    memset(&c->lir, 0, sizeof c->lir);
    if (counts_f)
        nk_lineidx_from_counts(&c->lif, &c->ff, counts_f);
    if (counts_r && paired)
        nk_lineidx_from_counts(&c->lir, &c->rf, counts_r);
    int rc = NK_OK;
    if (plan)
    {
        memcpy(c->fs, plan[0], sizeof(uint64_t) * (size_t)P);
        memcpy(c->fe, plan[1], sizeof(uint64_t) * (size_t)P);
        memset(c->rs, 0, sizeof(uint64_t) * (size_t)P);
        memset(c->re, 0, sizeof(uint64_t) * (size_t)P);
        if (paired)
        {
            memcpy(c->rs, plan[2], sizeof(uint64_t) * (size_t)P);
            memcpy(c->re, plan[3], sizeof(uint64_t) * (size_t)P);
        }
        for (int t = 0; t < P; t++)
            if (c->fe[t] >= fsize || c->fs[t] > fsize || (paired && (c->re[t] >= rsize || c->rs[t] > rsize)))
                rc = nk_fail(c, NK_EINVAL, "nk_process_planned: range outside the file");
    }
    else
    {
        nk_diag d = {{0}, 0};
        /* raw-text steps: the reverse file's boundaries may be placed while the first partitions are worked on.
         * Opt-in (NKB200_ROLLING_COUNT=1): measured on one B200 it ends 2 % sooner, but the engines' step spans then
         * start with one partition each and the GPU-busy time grows by 5 % (profiles/r02_ab_engines.txt). */
        int defer = c->raw_mode && paired && nk_env_on("NKB200_ROLLING_COUNT") && !nk_env_on("NKB200_EAGER_COUNT");
        if (nk_plan(&c->ff, &c->rf, paired, P, fastq, c->threads, c->fs, c->fe, c->rs, c->re, &d, &c->lif, &c->lir, &defer))
            rc = nk_fail(c, NK_EDATA, "%s", d.msg);
        else if (defer)
        {
            nk_lineidx_open(&c->lir, &c->rf, 0);
            if (!c->lir.cum)
                rc = nk_fail(c, NK_ENOMEM, "Memory allocation failed (line index)");
            else
                c->rolling = 1;
        }
    }
    int raw_work = 0;
    if (!rc && c->raw_mode)
    {
        /* line indexes over the byte ranges of this context's partitions (kept from planning when it made them) */
        size_t flo = SIZE_MAX, fhi = 0, rlo = SIZE_MAX, rhi = 0;
        for (int i = 0; i < c->n_local; i++)
        {
            int g = c->part[i].gid;
            if (c->fs[g] < c->fe[g] && (!paired || c->rs[g] < c->re[g]))
            {
                flo = c->fs[g] < flo ? c->fs[g] : flo;
                fhi = c->fe[g] > fhi ? c->fe[g] : fhi;
                rlo = c->rs[g] < rlo ? c->rs[g] : rlo;
                rhi = c->re[g] > rhi ? c->re[g] : rhi;
            }
        }
        /* Ranges that came without a count of the files (split by size, one partition, a caller's plan) are counted
         * by the engines' step builders as they go: the GPUs start at once.  NKB200_EAGER_COUNT=1 counts first. */
        if (nk_env_on("NKB200_EAGER_COUNT"))
        {
            if (!c->lif.cum && flo != SIZE_MAX)
                nk_lineidx_build_range(&c->lif, &c->ff, c->threads, flo, fhi + nk_range_slack);
            if (paired && !c->lir.cum && rlo != SIZE_MAX)
                nk_lineidx_build_range(&c->lir, &c->rf, c->threads, rlo, rhi + nk_range_slack);
        }
    }
    for (int i = 0; i < c->n_local && !rc; i++)
    {
        nk_part *p = &c->part[i];
        p->cur.fp = c->fs[p->gid];
        p->cur.fe = c->fe[p->gid];
        p->cur.rp = c->rs[p->gid];
        p->cur.re = c->re[p->gid];
        p->cur.done = 0;
        p->commit_fp = p->raw_fp = p->cur.fp;
        p->commit_rp = p->raw_rp = p->cur.rp;
        p->raw_total = p->raw_next = 0;
        p->raw_known = 1;
        p->ready = 1;
        p->lf = p->lr = NULL;
        memset(&p->own_f, 0, sizeof p->own_f);
        memset(&p->own_r, 0, sizeof p->own_r);
        if (c->rolling)
        { /* the reverse side follows when nk_roll_thread releases the partition */
            p->lf = &c->lif;
            p->lr = &c->lir;
            p->line_f = nk_lineidx_before(&c->lif, p->cur.fp);
            p->ready = 0;
            p->raw_total = NK_RAW_PLENTY;
        }
        else if (c->raw_mode && p->cur.fp < p->cur.fe && (!paired || p->cur.rp < p->cur.re))
        {
            int own = 0;
            for (int m = 0; m < (paired ? 2 : 1); m++)
            {
                nk_lineidx *whole = m ? &c->lir : &c->lif, *mine = m ? &p->own_r : &p->own_f;
                size_t at = m ? p->cur.rp : p->cur.fp;
                if (!whole->cum)
                { /* the partition's own index: the chunk of its first byte for now */
                    nk_lineidx_open(mine, m ? &c->rf : &c->ff, at);
                    if (!mine->cum)
                    {
                        rc = nk_fail(c, NK_ENOMEM, "Memory allocation failed (line index)");
                        break;
                    }
                    nk_lineidx_extend(mine, 1, at + 1);
                    own = 1;
                }
                *(m ? &p->lr : &p->lf) = whole->cum ? whole : mine;
                *(m ? &p->line_r : &p->line_f) = nk_lineidx_before(whole->cum ? whole : mine, at);
            }
            if (rc)
                break;
            if (own)
            {
                p->raw_known = 0;
                p->raw_total = NK_RAW_PLENTY;
                raw_work = 1;
            }
            else
            {
                uint64_t n = nk_raw_records_in(p->lf, p->cur.fp, p->cur.fe, per, p->line_f);
                if (paired)
                {
                    uint64_t nr = nk_raw_records_in(p->lr, p->cur.rp, p->cur.re, per, p->line_r);
                    n = nr < n ? nr : n;
                }
                p->raw_total = n;
                raw_work |= n > 0;
            }
        }
        p->t_start = nk_now();
        p->last_processed = p->processed;
    }
    if (!rc && c->raw_mode)
    {
        int own = 0;
        for (int i = 0; i < c->n_local; i++)
            own |= !c->part[i].raw_known;
        c->count_route[c->rolling ? 1 : own ? 2 : 0]++;
    }
    int roll_running = 0;
    if (!rc && c->rolling)
    {
        pthread_mutex_init(&c->roll_mu, NULL);
        pthread_cond_init(&c->roll_cv, NULL);
        if (pthread_create(&c->roll_th, NULL, nk_roll_thread, c) == 0)
            roll_running = 1;
        else
            nk_roll_thread(c); /* no thread: count here and now */
    }
    c->tot.index_seconds += nk_now() - t0;
    nk_trace(-1, -1, "plan", t0, nk_now());
    /* Waves: an engine whose tables do not fit its share of the GPU works on `per_wave` of its partitions at a time,
     * to the end of their byte ranges; the tables of the others wait in host memory (the engine parks and fetches
     * them as the steps name other partitions).  One wave when everything fits. */
    int n_waves = 1;
    int per_wave[256];
    for (int d = 0; d < c->n_dev && d < 256; d++)
    {
        nk_dev *dv = &c->dev[d];
        per_wave[d] = dv->n_parts;
        if (!dv->table_budget)
            continue;
        uint64_t table = 0;
        for (int i = 0; i < dv->n_parts; i++)
        {
            nkd_part_stats st;
            if (nkd_part_stats_get(dv->eng, i, &st) == NK_OK && st.capacity * 16 > table)
                table = st.capacity * 16;
        }
        /* a wave's tables, each with room to grow once, plus one re-hash in flight */
        uint64_t fit = table ? dv->table_budget / table : (uint64_t)dv->n_parts;
        int r = fit > 3 ? (int)((fit - 2) * 2 / 3) : 1;
        if (r < 1)
            r = 1;
        int waves = (dv->n_parts + r - 1) / r;
        per_wave[d] = (dv->n_parts + waves - 1) / waves; /* even waves */
        if (waves > n_waves)
            n_waves = waves;
    }
    for (int w = 0; w < n_waves && !rc; w++)
    {
        raw_work = roll_running; /* partitions are still being released: their record counts are not ours to read yet */
        for (int i = 0; i < c->n_local; i++)
        {
            nk_part *p = &c->part[i];
            int d = p->dev < 256 ? p->dev : 255;
            p->active = p->lidx / per_wave[d] == w;
            if (!roll_running)
                raw_work |= p->active && p->raw_total > p->raw_next;
        }
        c->waves++;
        if (raw_work)
            rc = nk_run_pipelines(c, 1);
        if (roll_running)
        { /* every partition has been released by now, or the pipelines stopped early */
            pthread_join(c->roll_th, NULL);
            roll_running = 0;
        }
        /* whatever the raw-text pipeline did not take -- text the device declined (NUL bytes, lines of 1024+ chars),
         * a last record cut short by the end of the file -- goes through the byte-exact host parser from where each
         * partition stands */
        int parsed_work = 0;
        for (int i = 0; i < c->n_local && !rc; i++)
        {
            nk_part *p = &c->part[i];
            if (!p->active)
                continue;
            p->cur.fp = p->commit_fp;
            p->cur.rp = p->commit_rp;
            nk_nliter_seek(&p->cur.itf, c->ff.data, p->cur.fp, c->ff.size);
            if (paired)
                nk_nliter_seek(&p->cur.itr, c->rf.data, p->cur.rp, c->rf.size);
            if (p->cur.fp < p->cur.fe && (!paired || p->cur.rp < p->cur.re))
                parsed_work = 1;
        }
        if (!rc && parsed_work)
            rc = nk_run_pipelines(c, 0);
    }
    nk_trace(-1, -1, "pipelines", t0, nk_now());
    if (roll_running)
        pthread_join(c->roll_th, NULL);
    if (c->rolling)
    {
        pthread_mutex_destroy(&c->roll_mu);
        pthread_cond_destroy(&c->roll_cv);
        c->rolling = 0;
    }
    nk_lineidx_free(&c->lif);
    nk_lineidx_free(&c->lir);
    for (int i = 0; i < c->n_local; i++)
    {
        nk_lineidx_free(&c->part[i].own_f);
        nk_lineidx_free(&c->part[i].own_r);
    }
    /* reporting totals are sums of the partitions' cumulative counters, C:1897-1909 */
    uint64_t pr = 0, pt = 0, sk = 0, mu = 0;
    for (int i = 0; i < c->n_local; i++)
    {
        nk_part *p = &c->part[i];
        pr += p->processed;
        pt += p->printed;
        sk += p->skipped;
        nkd_part_stats st;
        if (nkd_part_stats_get(c->dev[p->dev].eng, p->lidx, &st) == NK_OK && st.used > mu)
            mu = st.used;
    }
    c->tot.processed = pr;
    c->tot.printed = pt;
    c->tot.skipped = sk;
    c->file_max_used = mu;
    if (mu > c->tot.max_used)
        c->tot.max_used = mu;
    c->tot.process_seconds += nk_now() - t0;
    return rc;
}

int nk_process_paired(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size)
{
    return nk_process(c, fwd, fwd_size, rev, rev_size, 1, NULL, NULL, NULL);
}
int nk_process_single(nk_ctx *c, const char *fwd, size_t fwd_size)
{
    return nk_process(c, fwd, fwd_size, NULL, 0, 0, NULL, NULL, NULL);
}
/* The per-chunk line-end counts every planning step starts from (count_records_seqfile's scan, C:1302-1320, cut
 * into NK_LINE_CHUNK-byte chunks).  The ranks of a multi-process launch count a share of the chunks each,
 * exchange the counts, and hand all of them to nk_process_indexed, so that the files are scanned once per node
 * instead of once per rank. */
size_t nk_line_chunk_bytes(void) { return NK_LI_CHUNK; }
int nk_count_chunk_lines(const char *data, size_t size, size_t chunk_first, size_t n_chunks, uint32_t *counts, int threads)
{
    nk_buf f = {data, size};
    size_t total = (size + NK_LI_CHUNK - 1) / NK_LI_CHUNK;
    if (!data || chunk_first > total || n_chunks > total - chunk_first)
        return NK_EINVAL;
    if (n_chunks == 0)
        return NK_OK;
    nk_lineidx li = {0};
    nk_lineidx_build_range(&li, &f, threads > 0 ? threads : nk_host_threads(), chunk_first * NK_LI_CHUNK,
                           (chunk_first + n_chunks) * NK_LI_CHUNK);
    for (size_t i = 0; i < n_chunks; i++)
        counts[i] = (uint32_t)(li.cum[chunk_first + i + 1] - li.cum[chunk_first + i]);
    nk_lineidx_free(&li);
    return NK_OK;
}
int nk_process_indexed(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size,
                       const uint32_t *fwd_counts, const uint32_t *rev_counts)
{
    if (!fwd_counts || (rev && !rev_counts))
        return nk_fail(c, NK_EINVAL, "nk_process_indexed: missing line counts");
    return nk_process(c, fwd, fwd_size, rev, rev ? rev_size : 0, rev != NULL, NULL, fwd_counts, rev_counts);
}
int nk_process_planned(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size,
                       const uint64_t *fwd_starts, const uint64_t *fwd_ends, const uint64_t *rev_starts,
                       const uint64_t *rev_ends)
{
    const uint64_t *plan[4] = {fwd_starts, fwd_ends, rev_starts, rev_ends};
    if (!fwd_starts || !fwd_ends || (rev && (!rev_starts || !rev_ends)))
        return nk_fail(c, NK_EINVAL, "nk_process_planned: missing range arrays");
    return nk_process(c, fwd, fwd_size, rev, rev ? rev_size : 0, rev != NULL, plan, NULL, NULL);
}

static int nk_span_cmp(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return x < y ? -1 : x > y;
}

/* length of the union of the step spans of every engine on one GPU (nkd_run_spans) */
static double nk_gpu_busy_ms(nk_ctx *c, int ordinal, double fallback)
{
    size_t total = 0, engines = 0;
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].ordinal == ordinal)
        {
            size_t n = 0;
            nkd_run_spans(c->dev[d].eng, NULL, 0, &n);
            total += n;
            engines++;
        }
    if (engines < 2 || !total)
        return fallback;
    float *sp = malloc(total * 2 * sizeof *sp);
    if (!sp)
        return fallback;
    size_t at = 0;
    for (int d = 0; d < c->n_dev; d++)
        if (c->dev[d].ordinal == ordinal)
        {
            size_t n = 0;
            nkd_run_spans(c->dev[d].eng, sp + 2 * at, total - at, &n);
            at += n < total - at ? n : total - at;
        }
    qsort(sp, at, 2 * sizeof *sp, nk_span_cmp);
    double busy = 0, lo = sp[0], hi = sp[1];
    for (size_t i = 1; i < at; i++)
    {
        if (sp[2 * i] > hi)
        {
            busy += hi - lo;
            lo = sp[2 * i];
            hi = sp[2 * i + 1];
        }
        else if (sp[2 * i + 1] > hi)
            hi = sp[2 * i + 1];
    }
    busy += hi - lo;
    free(sp);
    return busy;
}

int nk_totals_get(nk_ctx *c, nk_totals *out)
{
    *out = c->tot;
    out->run_ms = out->probe_ms = 0;
    out->launches = out->probe_launches = out->probe_touches = out->h2d_bytes = out->d2h_bytes = 0;
    out->ops = out->touches = out->slow_events = out->expansions = 0;
    out->pend_events = out->open_ops = 0;
    out->hot_hits = 0;
    memset(out->class_ms, 0, sizeof out->class_ms);
    for (int d = 0; d < c->n_dev; d++)
    {
        nkd_run_stats rs;
        nkd_run_stats_get(c->dev[d].eng, &rs);
        /* engines sharing a GPU run concurrently: that GPU's busy time is the union of their step spans */
        int first = 1;
        for (int o = 0; o < d; o++)
            if (c->dev[o].ordinal == c->dev[d].ordinal)
                first = 0;
        if (first)
            out->run_ms += nk_gpu_busy_ms(c, c->dev[d].ordinal, rs.run_ms);
        out->probe_ms += rs.probe_ms;
        out->launches += rs.launches;
        out->probe_launches += rs.probe_launches;
        out->probe_touches += rs.probe_touches;
        out->h2d_bytes += rs.h2d_bytes;
        out->d2h_bytes += rs.d2h_bytes;
        out->hot_hits += rs.hot_hits;
        out->pend_events += rs.pend_events;
        out->open_ops += rs.open_ops;
        for (int k = 0; k < 10; k++)
            out->class_ms[k] += rs.class_ms[k];
    }
    out->engines = (uint64_t)c->n_dev;
    out->raw_steps = c->raw_steps;
    out->parsed_steps = c->parsed_steps;
    if (c->seeded)
        for (int i = 0; i < c->n_local; i++)
        {
            nkd_part_stats st;
            if (nkd_part_stats_get(c->dev[c->part[i].dev].eng, c->part[i].lidx, &st) == NK_OK)
            {
                out->ops += st.ops;
                out->touches += st.touches;
                out->slow_events += st.slow_events;
                out->expansions += st.expansions;
            }
        }
    return NK_OK;
}

int nk_partition_stats(nk_ctx *c, int partition, nkd_part_stats *out)
{
    int i = partition - c->part_first;
    if (i < 0 || i >= c->n_local)
        return nk_fail(c, NK_EINVAL, "partition %d is not owned by this context", partition);
    nk_part *p = &c->part[i];
    int rc = nkd_part_stats_get(c->dev[p->dev].eng, p->lidx, out);
    if (rc)
        return nk_fail(c, rc, "%s", nkd_last_error(c->dev[p->dev].eng));
    out->processed = p->processed;
    out->printed = p->printed;
    out->skipped = p->skipped;
    return NK_OK;
}

static void nk_dump_task(int d, void *a)
{
    nk_ctx *c = a;
    nk_dev *dv = &c->dev[d];
    dv->rc = NK_OK;
    for (int i = 0; i < dv->n_parts && !dv->rc; i++)
    {
        nk_part *p = &c->part[dv->parts[i]];
        nkd_part_stats st;
        nkd_part_stats_get(dv->eng, p->lidx, &st);
        dv->rc = nk_write_dump(c, dv->eng, p->lidx, st.capacity, "", p->gid);
    }
}

int nk_finish(nk_ctx *c)
{ /* C:2398-2413 */
    if (c->finished)
        return NK_OK;
    c->finished = 1;
    int rc = NK_OK;
    for (int i = 0; i < c->n_local; i++)
    {
        nk_part *p = &c->part[i];
        if (p->out_f && fclose(p->out_f) != 0 && !rc)
            rc = nk_fail(c, NK_EIO, "error closing output: %s", strerror(errno));
        p->out_f = NULL;
        if (p->out_r && fclose(p->out_r) != 0 && !rc)
            rc = nk_fail(c, NK_EIO, "error closing output: %s", strerror(errno));
        p->out_r = NULL;
    }
    double t0 = nk_now();
    if (c->cfg.dump_tables && c->seeded && !rc)
    { /* every engine formats and writes its own partitions' tables; engines work side by side */
        nk_parallel_for(c->n_dev, c->n_dev, nk_dump_task, c);
        for (int d = 0; d < c->n_dev && !rc; d++)
            rc = c->dev[d].rc;
        if (c->cfg.verbose)
            printf("B200: k-mer tables written in %.3f s\n", nk_now() - t0);
    }
    t0 = nk_now();
    if (c->seeded && !rc && c->cfg.merged_table)
    {
        rc = nk_write_merged_table(c);
        if (c->cfg.verbose)
            printf("B200: merged k-mer table written in %.3f s\n", nk_now() - t0);
    }
    if (c->seeded && !rc && c->cfg.merged_output)
    {
        rc = nk_concat_outputs(c, "output_forward");
        if (!rc && c->cfg.have_reverse)
            rc = nk_concat_outputs(c, "output_reverse");
    }
    return rc;
}

/* ------------------------------------------------------------------ CLI (C:492-832, C:2223-2455) */

typedef struct
{
    char **fwd;
    int nfwd;
    char **rev;
    int nrev;
    int single, memory, debug;
    nk_config cfg;
} nk_cli;

static void nk_usage(void)
{
    fprintf(stderr,
            "Usage:\n\n\t\tMandatory:\n"
            "\t\t* --forward|-f file1 [file2+]\tList of forward (read1) sequence files\n"
            "\t\t* --reverse|-r file1 [file2+]\tList of reverse (read2) sequence files\n\n"
            "\t\tOptional:\n"
            "\t\t[--single|-s]\t\t\t\tdata are single ended; --forward files without a --reverse partner are single-end\n"
            "\t\t[--ksize|-k (integer 5-31; def. 15)]\tk-mer size\n"
            "\t\t[--depth|-d (integer; def. 100)]\tcount at which a k-mer is high coverage; at least 2 x partitions\n"
            "\t\t[--coverage|-g (float 0-1; def. 0.9)]\tproportion of a read's k-mers that must be high coverage to drop it\n"
            "\t\t[--canonical|-c]\t\t\tmerge k-mers with their reverse complement\n"
            "\t\t[--filetype|-t (fq|fa; def. fq)]\tinput format\n"
            "\t\t[--outformat|-o (fq|fa; def. fq)]\toutput format\n"
            "\t\t[--memory_start|-m (integer Gb)]\tinitial table memory across all partitions\n"
            "\t\t[--cpu|-p (int; def 1)]\t\t\tnumber of partitions (the reference's threads); fixed by the user, spread over the GPUs\n"
            "\t\t[--verbose|-e] [--debug|-b level] [--print|-P] [--version|-v] [--help|-h]\n"
            "\t\tB200 placement: --gpus N (or NKB200_GPUS) uses N GPUs of this node; results do not depend on N\n"
            "\t\tB200 extras:    --merged-table\talso write output_kmer_merged.*.tsv: all threads' kmers, sorted, counts summed\n"
            "\t\t                --merged-output\talso write output_forward/output_reverse.*.fastq without _thread: all threads concatenated\n\n");
}

static int nk_is_fa(const char *s) { return !strcasecmp(s, "fa") || !strcasecmp(s, "fasta") || !strcasecmp(s, "fsa") || !strcasecmp(s, "fas"); }
static int nk_is_fq(const char *s) { return !strcasecmp(s, "fq") || !strcasecmp(s, "fastq") || !strcasecmp(s, "fsq"); }

/* -f/-r take every following argument up to the next one starting with '-', C:747-832 */
static void nk_add_files(char ***list, int *n, char *first, char **argv, int *idx)
{
    char *cur = first;
    for (;;)
    {
        if (access(cur, R_OK) == 0)
        {
            *list = realloc(*list, (size_t)(*n + 1) * sizeof(char *));
            (*list)[(*n)++] = strdup(cur);
        }
        else
            fprintf(stderr, "Warning: File '%s' does not exist or is not readable. Skipping.\n", cur);
        if (argv[*idx] == NULL || argv[*idx][0] == '-')
            break;
        cur = argv[(*idx)++];
    }
}

static int nk_parse(nk_cli *a, int argc, char **argv, int *gpus)
{
    memset(a, 0, sizeof *a);
    nk_config *c = &a->cfg;
    c->coverage = 0.9;
    c->partitions = 1;
    c->k = 15;
    c->depth = 100;
    c->in_fastq = c->out_fastq = 1;
    static struct option lo[] = {{"forward", 1, 0, 'f'}, {"reverse", 1, 0, 'r'}, {"ksize", 1, 0, 'k'}, {"depth", 1, 0, 'd'},
                                 {"coverage", 1, 0, 'g'}, {"filetype", 1, 0, 't'}, {"outformat", 1, 0, 'o'}, {"cpu", 1, 0, 'p'},
                                 {"memory_start", 1, 0, 'm'}, {"debug", 1, 0, 'b'}, {"verbose", 0, 0, 'e'}, {"help", 0, 0, 'h'},
                                 {"canonical", 0, 0, 'c'}, {"version", 0, 0, 'v'}, {"single", 0, 0, 's'}, {"print", 0, 0, 'P'},
                                 {"gpus", 1, 0, 1000}, {"merged-table", 0, 0, 1001}, {"merged-output", 0, 0, 1002},
                                 {0, 0, 0, 0}};
    int o;
    optind = 1;
    while ((o = getopt_long(argc, argv, "f:r:k:d:g:t:o:p:m:b:ehcvsP", lo, NULL)) != -1)
    {
        switch (o)
        {
        case 1000:
            *gpus = atoi(optarg);
            break;
        case 1001:
            c->merged_table = 1;
            break;
        case 1002:
            c->merged_output = 1;
            break;
        case 'P':
            c->dump_tables = 1;
            break;
        case 's':
            a->single = 1;
            break;
        case 'c':
            c->canonical = 1;
            break;
        case 'm':
            a->memory = atoi(optarg);
            if (a->memory < 1)
            {
                printf("Memory cannot be less than 1 Gb %'d\n", a->memory);
                return 0;
            }
            break;
        case 'b':
            a->debug = atoi(optarg);
            break;
        case 'h':
            nk_usage();
            exit(EXIT_SUCCESS);
        case 'p':
            c->partitions = atoi(optarg);
            break;
        case 'f':
            nk_add_files(&a->fwd, &a->nfwd, optarg, argv, &optind);
            break;
        case 'r':
            nk_add_files(&a->rev, &a->nrev, optarg, argv, &optind);
            break;
        case 'k':
            c->k = atoi(optarg);
            break;
        case 'd':
            c->depth = atoi(optarg);
            break;
        case 'g':
            c->coverage = atof(optarg);
            break;
        case 'v':
            printf("%d\n", NK_VERSION);
            exit(EXIT_SUCCESS);
        case 'e':
            c->verbose = 1;
            break;
        case 't':
            if (nk_is_fa(optarg))
                c->in_fastq = 0;
            else if (nk_is_fq(optarg))
                c->in_fastq = 1;
            else
            {
                printf("Input file format must be either fa or fq, not %s\n", optarg);
                return 0;
            }
            break;
        case 'o':
            if (nk_is_fa(optarg))
                c->out_fastq = 0;
            else if (nk_is_fq(optarg))
                c->out_fastq = 1;
            else
            {
                printf("Output file format must be either fa or fq, not %s\n", optarg);
                return 0;
            }
            break;
        default:
            fprintf(stderr, "Unexpected error in option processing\n");
            return 0;
        }
    }
    if (c->verbose)
    {
        printf("\nVERSION: %d, CMD: ", NK_VERSION);
        for (int i = 0; i < argc; i++)
            printf("%s ", argv[i]);
        printf("\n\n");
    }
    if (c->partitions <= 0)
    { /* the reference divides by the CPU count before checking it (C:674): report the check's message */
        fprintf(stderr, "Error: CPU count (%d) must be a positive integer and up to %d\n", c->partitions, NK_MAX_PARTITIONS);
        return 0;
    }
    c->memory_gb = a->memory;
    int depth_part = c->depth / c->partitions;
    uint64_t cap = nk_capacity_unclamped(a->memory, c->partitions);
    float mem_part = (float)cap * 16 / 1073741824;
    uint64_t lim = nk_pow4_wrapping(c->k);
    int mem_total = a->memory;
    if (lim < cap)
    {
        cap = lim;
        mem_part = (float)cap * 16 / 1073741824;
        mem_total = (int)(mem_part * c->partitions);
    }
    printf("Initial hash table size set to %'zu (maximum for k=%d is %'zu); memory ~ %'0.2f Gb for each of %d threads (~ %'d Gb total))\n\n",
           (size_t)cap, c->k, (size_t)lim, mem_part, c->partitions, mem_total);
    if (a->nfwd == 0 || (a->nrev == 0 && !a->single))
    {
        fprintf(stderr, "Error: no fwd (%d) or reverse (%d) files provided\n", a->nfwd, a->nrev);
        return 0;
    }
    if (!c->in_fastq && c->out_fastq)
    {
        fprintf(stderr, "Error: cannot request an output format of FASTQ when input is FASTA\n");
        return 0;
    }
    if (!a->single && a->nfwd != a->nrev)
    {
        fprintf(stderr, "Error: Number of forward (%d) and reverse files (%d) must match\n", a->nfwd, a->nrev);
        return 0;
    }
    if (c->partitions > NK_MAX_PARTITIONS)
    {
        fprintf(stderr, "Error: CPU count (%d) must be a positive integer and up to %d\n", c->partitions, NK_MAX_PARTITIONS);
        return 0;
    }
    if (c->k < 5 || c->k > 31)
    {
        fprintf(stderr, "Error: Only kmer sizes (%d) of 5 to 31 are supported\n", c->k);
        return 0;
    }
    if (c->coverage > 1 || c->coverage < 0.001)
    {
        fprintf(stderr, "Error: Coverage (%3.f) is the proportion of the sequence covered by high kmers and must be between 0 and 1\n", c->coverage);
        return 0;
    }
    if (c->depth < 2)
    {
        fprintf(stderr, "Error: Depth (%d) is the number of times a kmer needs to be found before being flagged as high coverage, it must be above 1\n", c->depth);
        return 0;
    }
    if (depth_part < 2)
    {
        fprintf(stderr, "Error: Depth (%d) must be at least 2 x number of CPUs (for performance reasons; but this version of the program is written to normalise to 50+\n", c->depth);
        return 0;
    }
    return 1;
}

typedef struct
{
    void *map;
    size_t size;
} nk_map;

static int nk_map_file(const char *path, nk_map *m)
{ /* mmap_file, C:424-461 */
    m->map = NULL;
    m->size = 0;
    int fd = open(path, O_RDONLY);
    if (fd < 0)
    {
        perror("Error opening file");
        return -1;
    }
    struct stat sb;
    if (fstat(fd, &sb) < 0)
    {
        perror("Error getting file size");
        close(fd);
        return -1;
    }
    m->size = (size_t)sb.st_size;
    if (m->size == 0)
    {
        fprintf(stderr, "Error mapping file: Invalid argument\n");
        close(fd);
        return -1;
    }
    m->map = mmap(NULL, m->size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m->map == MAP_FAILED)
    {
        perror("Error mapping file");
        m->map = NULL;
        m->size = 0;
        return -1;
    }
    madvise(m->map, m->size, MADV_SEQUENTIAL);
    return 0;
}

static void nk_unmap(nk_map *m)
{
    if (m->map)
        munmap(m->map, m->size);
    m->map = NULL;
    m->size = 0;
}

static int nk_main_run(nk_cli *a, int argc, char **argv);

static void nk_cli_free(nk_cli *a)
{
    for (int i = 0; i < a->nfwd; i++)
        free(a->fwd[i]);
    for (int i = 0; i < a->nrev; i++)
        free(a->rev[i]);
    free(a->fwd);
    free(a->rev);
    a->fwd = a->rev = NULL;
    a->nfwd = a->nrev = 0;
}

int nk_main(int argc, char **argv)
{
    nk_cli a;
    int rc = nk_main_run(&a, argc, argv);
    nk_cli_free(&a);
    return rc;
}

static int nk_main_run(nk_cli *a, int argc, char **argv)
{
    setlocale(LC_ALL, "");
    int gpus = getenv("NKB200_GPUS") ? atoi(getenv("NKB200_GPUS")) : 1;
    if (!nk_parse(a, argc, argv, &gpus))
    {
        nk_usage();
        return 1;
    }
    nk_config *cfg = &a->cfg;
    cfg->n_forward_files = a->nfwd;
    cfg->have_reverse = a->nrev != 0;
    int avail = nkd_device_count();
    if (gpus < 1)
        gpus = 1;
    if (avail > 0 && gpus > avail)
        gpus = avail;
    cfg->n_devices = gpus;
    nk_ctx *c = NULL;
    int rc = nk_create(cfg, &c);
    if (rc)
    {
        fprintf(stderr, "%s\n", nk_create_error());
        return 1;
    }
    double t_seed = nk_now();
    int want = 1 + (int)(3e6 / a->nfwd); /* C:2242 */
    for (int i = 0; i < a->nfwd && !rc; i++)
    {
        for (int m = 0; m < 2 && !rc; m++)
        {
            if (m == 1 && i >= a->nrev)
                break;
            const char *path = m ? a->rev[i] : a->fwd[i];
            if (cfg->verbose)
                printf("Seeding hash table with up to %'d records from file %s\n", want, path);
            nk_map mp;
            if (nk_map_file(path, &mp) < 0)
                continue; /* the reference walks a zero-length mapping here */
            rc = nk_seed_buffer(c, mp.map, mp.size, want);
            nk_unmap(&mp);
        }
    }
    if (!rc)
        rc = nk_seed_finish(c);
    if (rc)
    {
        fprintf(stderr, "%s\n", nk_last_error(c));
        nk_destroy(c);
        return 1;
    }
    if (cfg->verbose)
        printf("Seeding took %.2f seconds on %d GPU(s)\n", nk_now() - t_seed, gpus);

    time_t start_time = time(NULL); /* the reference's clock starts after seeding, C:2308 */
    double t_proc = nk_now();
    int mapped_ok = 1;
    for (int i = 0; i < a->nfwd; i++)
    {
        int paired = i < a->nrev;
        nk_map mf = {0}, mr = {0};
        int okf = nk_map_file(a->fwd[i], &mf) == 0, okr = 1;
        if (paired)
        {
            printf("Processing file pair %d of %d: %s and %s\n", i + 1, a->nfwd, a->fwd[i], a->rev[i]);
            okr = nk_map_file(a->rev[i], &mr) == 0;
        }
        else
            printf("Processing single-ended file %d of %d: %s\n", i + 1, a->nfwd, a->fwd[i]);
        if (!okf || !okr)
        { /* the reference jumps to cleanup and still returns 0, C:2318-2321 */
            fprintf(stderr, "Error memory mapping input files\n");
            mapped_ok = 0;
            nk_unmap(&mf);
            nk_unmap(&mr);
            break;
        }
        char lead = cfg->in_fastq ? '@' : '>';
        const char *kind = cfg->in_fastq ? "FASTQ" : "FASTA";
        if (((const char *)mf.map)[0] != lead)
        {
            fprintf(stderr, "Input %s file %s starts with %c which is not expected\n", kind, a->fwd[i], ((const char *)mf.map)[0]);
            return 1;
        }
        if (paired && ((const char *)mr.map)[0] != lead)
        {
            fprintf(stderr, "Input %s file %s starts with %c which is not expected\n", kind, a->rev[i], ((const char *)mr.map)[0]);
            return 1;
        }
        /* what the reference's worker keeps from the start of a file for the percentages of its report line, C:1592-1594 */
        uint64_t before[NK_MAX_PARTITIONS], before_printed[NK_MAX_PARTITIONS], before_skipped[NK_MAX_PARTITIONS],
            before_used[NK_MAX_PARTITIONS];
        for (int t = 0; t < c->n_local; t++)
        {
            nkd_part_stats st0;
            before[t] = c->part[t].processed;
            before_printed[t] = c->part[t].printed;
            before_skipped[t] = c->part[t].skipped;
            before_used[t] = nk_partition_stats(c, c->part[t].gid, &st0) == NK_OK ? st0.used : 0;
        }
        double tf = nk_now();
        rc = paired ? nk_process_paired(c, mf.map, mf.size, mr.map, mr.size) : nk_process_single(c, mf.map, mf.size);
        double dt = nk_now() - tf;
        nk_unmap(&mf);
        nk_unmap(&mr);
        if (rc == NK_EDATA)
        { /* FATAL / partition errors: message + exit(EXIT_FAILURE) */
            const char *m = nk_last_error(c);
            if (!strncmp(m, "ERROR", 5))
                printf("%s\n", m);
            else
                fprintf(stderr, "%s\n", m);
            nk_finish(c);
            nk_destroy(c);
            return 1;
        }
        if (rc)
        {
            fprintf(stderr, "%s\n", nk_last_error(c));
            fprintf(stderr, "Error processing files\n");
            nk_destroy(c);
            return 1;
        }
        for (int t = 0; t < c->n_local; t++)
        { /* the per-thread line of C:1752 (rates are ours) */
            nk_part *p = &c->part[t];
            nkd_part_stats st;
            nk_partition_stats(c, p->gid, &st);
            /* growth since the start of this file, as the reference reports it when no 60 s progress report
             * intervened (C:1746-1749); the rate's own percentage compares with such a report and stays 0 */
            float gp = before_printed[t] ? (float)(p->printed - before_printed[t]) / (int)before_printed[t] : 0;
            float gs = before_skipped[t] ? (float)(p->skipped - before_skipped[t]) / (int)before_skipped[t] : 0;
            float gk = before_used[t] ? (float)(st.used - before_used[t]) / before_used[t] : 0;
            printf("Thread %d - Processing rate: %'.0f (%+.2f%%) sequences/s, processed %'zu pairs, printed: %'zu (%+.2f%%), skipped: %'zu (%+.2f%%), Unique kmers (all sequences; this thread): %'zu (%+.2f%%)\n",
                   p->gid, dt > 0 ? (double)(p->processed - before[t]) / dt : 0.0, 0.0, (size_t)p->processed, (size_t)p->printed, gp * 100,
                   (size_t)p->skipped, gs * 100, (size_t)st.used, gk * 100);
        }
        printf("Cumulative file statistics: Processed %'zu, Printed %'zu, Skipped %'zu, Cumulative Max Unique Kmers in a thread: %'zu\n",
               (size_t)c->tot.processed, (size_t)c->tot.printed, (size_t)c->tot.skipped, (size_t)c->file_max_used);
    }
    rc = nk_finish(c);
    if (rc)
    {
        fprintf(stderr, "%s\n", nk_last_error(c));
        nk_destroy(c);
        return 1;
    }
    if (mapped_ok)
    {
        printf("\n--- Final Report ---\n");
        printf("Processed Records: %'zu\n", (size_t)c->tot.processed);
        printf("Printed Records: %'zu\n", (size_t)c->tot.printed);
        printf("Skipped Records: %'zu\n", (size_t)c->tot.skipped);
        printf("Cumulative Max unique kmers in any thread: %'zu\n", (size_t)c->tot.max_used);
    }
    double total_runtime = difftime(time(NULL), start_time);
    printf("Total runtime: %.2f seconds\n", total_runtime);
    if (c->tot.processed > 0)
    {
        double fine = nk_now() - t_proc;
        double rate = (double)c->tot.processed / (total_runtime > 0 ? total_runtime : fine);
        printf("Overall processing rate: %'.0f %s per second\n", rate, a->nrev ? "sequence pairs" : "sequences");
    }
    else
        printf("No data processed\n");
    if (cfg->verbose)
    {
        printf("B200: seed %.3f s, process %.3f s (index %.3f, device %.3f, write %.3f) on %d GPU(s), %d host threads\n",
               c->tot.seed_seconds, c->tot.process_seconds, c->tot.index_seconds, c->tot.device_seconds, c->tot.write_seconds, gpus, c->threads);
        printf("B200: %llu device steps on raw record text, %llu on host-parsed records\n", (unsigned long long)c->raw_steps,
               (unsigned long long)c->parsed_steps);
        printf("B200: %llu seed records taken from raw text on the device, %llu parsed by the host\n",
               (unsigned long long)c->seed_raw_records, (unsigned long long)c->seed_parsed_records);
        printf("B200: line ends counted before the first step for %llu input(s), reverse file alongside the first steps for %llu, "
               "by the step builders for %llu\n", (unsigned long long)c->count_route[0], (unsigned long long)c->count_route[1],
               (unsigned long long)c->count_route[2]);
        uint64_t ev = 0, ld = 0;
        for (int d = 0; d < c->n_dev; d++)
        {
            uint64_t e1 = 0, l1 = 0;
            nkd_residency_stats(c->dev[d].eng, NULL, &e1, &l1);
            ev += e1;
            ld += l1;
        }
        printf("B200: %llu wave(s) of partitions, %llu tables parked in host memory, %llu brought in\n", (unsigned long long)c->waves,
               (unsigned long long)ev, (unsigned long long)ld);
    }
    nk_destroy(c);
    return 0;
}
