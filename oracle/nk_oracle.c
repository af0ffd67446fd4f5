/*
 * nk_oracle.c -- CPU ORACLE for the k-mer coverage-normalisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nomalise_kmers_multi_large_b200/ may
 * include, link, import or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * It is a from-the-spec restatement of what
 * /root/reference/normalise_kmers_multi_large.c ("C:n" below) computes, written
 * sequentially: partitions are processed one after another, which is
 * result-identical because the reference's threads share no mutable state
 * (README:68) once the canonical-buffer race (C:1177) is removed.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks this program
 * byte-for-byte (outputs, counters, -P dumps) against the reference binary
 * built by oracle/Makefile into oracle/_ref/, and against the md5 goldens the
 * survey recorded ("Golden vectors").
 *
 * Deliberate, documented deviations (reference behaviour is undefined there):
 *   D1 bytes past EOF read as '\0' (reference relies on the zero page tail and
 *      faults when size is a multiple of the page size, C:397)
 *   D2 a record that read_line cuts short (a NUL byte, or the end of the file inside
 *      the record) is scored, counted and ends the partition as in the reference
 *      (C:1616-1631, C:1733) when both sequence lines were read; if it is accepted,
 *      only the lines that were read are printed (the reference prints whatever its
 *      stack holds for the others).  Cut before the sequence lines, it is not
 *      scored (reference: stale stack bytes are scored)
 *   D3 a final record dropped by the length gate ends the partition (reference
 *      dereferences NULL, C:1622-1631)
 *   D4 pure single-end with -p > 1 runs (reference tests an uninitialised FILE*,
 *      C:2153)
 *   D5 --canonical uses the race-free minimum (reference shares a static
 *      buffer between threads, C:1177)
 *   D6 equal-size partitioning with size/p <= 4096 is an error (reference wraps
 *      a size_t, C:1243)
 *
 * Build:  see oracle/Makefile  (libnk_oracle.so for ctypes, nk_oracle CLI).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <getopt.h>
#include <locale.h>
#include <time.h>
#include <unistd.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>

#define NKO_DEFAULT_SLOTS 67108879ULL /* C:137 */
#define NKO_LINE_MAX 1024             /* C:139 */
#define NKO_MAX_PARTS 256             /* C:142 */
#define NKO_VERSION 20240823          /* C:1   */

/* ------------------------------------------------------------------ table */

typedef struct
{
    uint64_t key; /* 0 = empty                         C:159 */
    int32_t count;
    uint32_t pad;
} nko_slot;

typedef struct
{
    nko_slot *slot;
    size_t cap, used;
    /* instrumentation (not in the reference) */
    uint64_t n_ops, n_touches, n_walk_ops, n_expansions;
} nko_table;

static void die(const char *msg)
{
    fprintf(stderr, "%s\n", msg);
    exit(EXIT_FAILURE);
}

nko_table *nko_table_new(size_t cap)
{
    nko_table *t = calloc(1, sizeof *t);
    if (!t)
        die("Memory allocation failed");
    t->slot = calloc(cap, sizeof(nko_slot)); /* C:898 */
    if (!t->slot)
        die("Memory allocation failed");
    t->cap = cap;
    return t;
}

void nko_table_free(nko_table *t)
{
    if (t)
    {
        free(t->slot);
        free(t);
    }
}

nko_table *nko_table_clone(const nko_table *src) /* C:908-927 */
{
    nko_table *t = nko_table_new(src->cap);
    memcpy(t->slot, src->slot, src->cap * sizeof(nko_slot));
    t->used = src->used;
    return t;
}

size_t nko_table_cap(const nko_table *t) { return t->cap; }
size_t nko_table_used(const nko_table *t) { return t->used; }
uint64_t nko_table_ops(const nko_table *t) { return t->n_ops; }
uint64_t nko_table_touches(const nko_table *t) { return t->n_touches; }
uint64_t nko_table_walk_ops(const nko_table *t) { return t->n_walk_ops; }
uint64_t nko_table_expansions(const nko_table *t) { return t->n_expansions; }
uint64_t nko_slot_key(const nko_table *t, size_t i) { return t->slot[i].key; }
int32_t nko_slot_count(const nko_table *t, size_t i) { return t->slot[i].count; }
void nko_table_export(const nko_table *t, uint64_t *keys, int32_t *counts)
{
    for (size_t i = 0; i < t->cap; i++)
    {
        keys[i] = t->slot[i].key;
        counts[i] = t->slot[i].count;
    }
}

/* growth: x1.5 in double, truncated; re-insert stored keys in old-slot order
 * with LINEAR probing; ghost counts (key==0,count>0) are dropped   C:1055-1108 */
void nko_expand(nko_table *t)
{
    size_t ncap = (size_t)((double)t->cap + (double)t->cap * 0.5);
    if (ncap <= t->cap)
        return;
    nko_slot *ns = calloc(ncap, sizeof(nko_slot));
    if (!ns)
        die("Error: Memory allocation failed to expand local hash table");
    size_t moved = 0;
    for (size_t i = 0; i < t->cap; i++)
    {
        if (t->slot[i].key == 0)
            continue;
        size_t j = t->slot[i].key % ncap;
        while (ns[j].key != 0)
            j = (j + 1) % ncap;
        ns[j] = t->slot[i];
        moved++;
    }
    free(t->slot);
    t->slot = ns;
    t->cap = ncap;
    t->used = moved;
    t->n_expansions++;
}

/* one table operation; returns the slot whose count the caller tests  C:929-1053 */
size_t nko_store(nko_table *t, uint64_t x, int init)
{
    if ((double)t->used >= (double)t->cap * 0.8) /* before EVERY op, C:933 */
        nko_expand(t);
    t->n_ops++;
    t->n_touches++;
    size_t i = x % t->cap;
    nko_slot *s = &t->slot[i];
    if (s->key == 0)
    { /* claim the home slot            C:948-971 */
        s->key = x;
        s->count = init ? 0 : 1;
        t->used++;
        return i;
    }
    if (s->key == x)
    { /* hit at home                    C:972-1008 */
        if (!init)
            s->count++;
        return i;
    }
    /* home taken by another key: cumulative-quadratic walk; every slot landed on
     * is incremented (foreign, empty or own); the key is never written C:1009-1048 */
    t->n_walk_ops++;
    int c = 0;
    while (t->slot[i].key != 0 && t->slot[i].key != x)
    {
        c++;
        i = (i + (size_t)(c * c)) % t->cap;
        t->n_touches++;
        if (init)
            t->slot[i].count = 0;
        else
            t->slot[i].count++;
    }
    return i;
}

/* ------------------------------------------------------------------ codec */

static inline int base_code(unsigned char b) /* C:150-153: anything else packs as 0 */
{
    switch (b)
    {
    case 'C':
        return 1;
    case 'G':
        return 2;
    case 'T':
        return 3;
    default:
        return 0;
    }
}

uint64_t nko_encode(const char *w, int k) /* C:1118-1126 */
{
    uint64_t x = 0;
    for (int i = 0; i < k; i++)
        x = (x << 2) | (uint64_t)base_code((unsigned char)w[i]);
    return x;
}

void nko_decode(uint64_t x, int k, char *out) /* C:1128-1136 */
{
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    for (int i = k - 1; i >= 0; i--)
    {
        out[i] = sym[x & 3];
        x >>= 2;
    }
    out[k] = 0;
}

/* encoding of the reverse complement of the window whose encoding is x */
uint64_t nko_revcomp(uint64_t x, int k)
{
    uint64_t r = 0;
    for (int i = 0; i < k; i++)
    {
        r = (r << 2) | (3 - (x & 3));
        x >>= 2;
    }
    return r;
}

/* strcmp order on ACGT strings == numeric order of encodings  C:1175-1180 */
uint64_t nko_window_key(const char *w, int k, int canonical)
{
    uint64_t f = nko_encode(w, k);
    if (!canonical)
        return f;
    uint64_t r = nko_revcomp(f, k);
    return f < r ? f : r;
}

/* keys of every window of seq (0 = window the reference ignores, C:1483) */
int nko_window_keys(const char *seq, int len, int k, int canonical, uint64_t *out)
{
    int n = len - k + 1;
    for (int i = 0; i < n; i++)
        out[i] = nko_window_key(seq + i, k, canonical);
    return n < 0 ? 0 : n;
}

/* ------------------------------------------------------------------ scoring */

/* C:1459-1499 */
void nko_score(nko_table *t, const char *seq, int len, int k, int canonical, int depth,
               int *high, int *total)
{
    *high = 0;
    *total = 0;
    for (int i = 0; i + k <= len; i++)
    {
        uint64_t x = nko_window_key(seq + i, k, canonical);
        if (x == 0)
            continue;
        (*total)++;
        size_t s = nko_store(t, x, 0);
        if (t->slot[s].count >= depth)
            (*high)++;
    }
}

/* C:1501-1537 */
static void seed_sequence(nko_table *t, const char *seq, int len, int k, int canonical)
{
    for (int i = 0; i + k <= len; i++)
    {
        uint64_t x = nko_window_key(seq + i, k, canonical);
        if (x)
            nko_store(t, x, 1);
    }
}

/* N -> A in place; returns 0 when a byte outside ACGT remains   C:475-486, C:1144-1158 */
static int scrub_and_check(char *s, int len)
{
    int ok = 1;
    for (int i = 0; i < len; i++)
    {
        if (s[i] == 'N')
            s[i] = 'A';
        else if (s[i] != 'A' && s[i] != 'C' && s[i] != 'G' && s[i] != 'T')
            ok = 0;
    }
    return ok;
}

/* float32 ratio and strict '<' against float32 coverage   C:1641-1646 */
int nko_keep_mate(int high, int total, float coverage)
{
    float r = total > 0 ? (float)high / (float)total : 0.0f;
    return r < coverage;
}

/* ------------------------------------------------------------------ capacity */

static size_t pow4_wrapping(int k) /* C:297-308 (wraps to 0 at k = 32, as size_t does) */
{
    size_t lim = 1;
    for (int i = 0; i < k && i < 64; i++)
        lim *= 4;
    return lim;
}

static size_t capacity_unclamped(int memory_gb, int parts) /* C:416-422, float32 arithmetic */
{
    if (memory_gb <= 0)
        return NKO_DEFAULT_SLOTS;
    size_t bytes = (size_t)((float)memory_gb * 1073741824);
    float total = (float)(bytes / 16);
    size_t per = (size_t)(total / (float)parts);
    return (per % 2 == 0) ? per + 1 : per;
}

size_t nko_capacity(int memory_gb, int parts, int k) /* C:676-684 */
{
    size_t cap = capacity_unclamped(memory_gb, parts);
    size_t lim = pow4_wrapping(k);
    return lim < cap ? lim : cap;
}

/* ------------------------------------------------------------------ input  */

typedef struct
{
    const char *data;
    size_t size;
    void *map;
    size_t maplen;
} nko_file;

static inline char at(const nko_file *f, size_t i) { return i < f->size ? f->data[i] : '\0'; } /* D1 */

static int file_open(nko_file *f, const char *path)
{
    memset(f, 0, sizeof *f);
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
    f->size = (size_t)sb.st_size;
    if (f->size)
    {
        f->map = mmap(NULL, f->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (f->map == MAP_FAILED)
        {
            perror("Error mapping file");
            close(fd);
            return -1;
        }
        f->maplen = f->size;
        f->data = f->map;
    }
    close(fd);
    return 0;
}

static void file_close(nko_file *f)
{
    if (f->map)
        munmap(f->map, f->maplen);
    memset(f, 0, sizeof *f);
}

/* the worker's line reader: stops at '\n', '\0' or 1023 chars; sets *more = 0
 * when the byte after the consumed span is '\0'                       C:394-409 */
static size_t take_line(const nko_file *f, size_t pos, char *buf, int *len, int *more)
{
    int n = 0;
    while (at(f, pos) != '\n' && at(f, pos) != '\0' && n < NKO_LINE_MAX - 1)
        buf[n++] = at(f, pos++);
    buf[n] = 0;
    if (at(f, pos) == '\n')
        pos++;
    *len = n;
    *more = at(f, pos) != '\0';
    return pos;
}

/* backwards search for a record boundary                              C:1199-1236 */
static size_t boundary_before(const nko_file *f, size_t lo, size_t hi, int fastq)
{
    if (!fastq)
    {
        for (size_t i = hi; i > lo; i--)
            if (at(f, i) == '>')
                return i - 1;
    }
    else
    {
        int nl = 0, plus = 0;
        for (size_t i = hi; i > lo; i--)
        {
            if (at(f, i) != '\n')
                continue;
            nl++;
            if (at(f, i + 1) == '+')
                plus = 1;
            else if (plus && at(f, i + 1) == '@')
                return i;
            if (nl == 7)
            {
                printf("ERROR: after 7 lines, I couldn't find the + and @ headers near this chunk %'zu\n", i);
                exit(EXIT_FAILURE);
            }
        }
    }
    printf("ERROR: i couldn't find the start of sequence before this chunk end %'zu\n", hi);
    exit(EXIT_FAILURE);
}

/* byte-size partitioner with its observable quirks (starts[1] stays 0, the
 * last end is overwritten)                                            C:1240-1262 */
void nko_ranges_by_size(const nko_file *f, int p, int fastq, size_t *st, size_t *en)
{
    size_t chunk = f->size / (size_t)p;
    if (chunk <= (size_t)NKO_LINE_MAX * 4)
        die("Error: input too small to split by size across this many partitions"); /* D6 */
    size_t approx = chunk - (size_t)NKO_LINE_MAX * 4;
    st[0] = 0;
    en[0] = boundary_before(f, 0, approx, fastq);
    en[p - 1] = f->size - 1;
    for (int t = 1; t < p; t++)
    {
        size_t s = en[t - 1] + 1;
        en[t] = boundary_before(f, s, s + approx, fastq);
        if (t < p - 1)
            st[t + 1] = en[t] + 1;
    }
}

size_t nko_count_records(const nko_file *f, int fastq) /* C:1302-1320 */
{
    size_t lines = 0;
    for (size_t i = 0; i < f->size; i++)
        if (f->data[i] == '\n')
            lines++;
    if (f->size > 0 && f->data[f->size - 1] != '\n')
        lines++;
    return fastq ? lines / 4 : lines / 2;
}

/* record-count partitioner                                            C:1265-1300 */
void nko_ranges_by_records(const nko_file *f, int p, int fastq, size_t records, size_t *st, size_t *en)
{
    size_t per = records / (size_t)p;
    if (p < 2 || per < 1 || f->size < 1)
        return;
    int want = (int)(fastq ? per * 4 : per * 2);
    st[0] = 0;
    en[p - 1] = f->size - 1;
    for (int t = 0; t < p - 1; t++)
    {
        size_t seen = 0;
        for (size_t i = st[t]; i < f->size; i++)
        {
            if (f->data[i] != '\n')
                continue;
            if (++seen == (size_t)want)
            {
                en[t] = i;
                st[t + 1] = i + 1;
                break;
            }
        }
    }
}

/* ------------------------------------------------------------------ run state */

typedef struct
{
    char **fwd;
    int nfwd;
    char **rev;
    int nrev;
    int k, depth, depth_part, parts, memory, canonical, single, dump, verbose, debug;
    int in_fastq, out_fastq;
    float coverage;
    size_t cap0;
} nko_cfg;

typedef struct
{
    nko_table *table;
    size_t processed, printed, skipped;
    FILE *out_f, *out_r;
} nko_part;

static char *out_name(const char *base, int k, int depth_part, int t, const char *suffix) /* C:834-850 */
{
    char *s = malloc(strlen(base) + 64);
    if (t >= 0)
        sprintf(s, "%s.k%d_norm%d_thread%d.%s", base, k, depth_part, t, suffix);
    else
        sprintf(s, "%s.k%d_norm%d.%s", base, k, depth_part, suffix);
    return s;
}

static void dump_table(const nko_cfg *c, const nko_table *t, const char *tag, int part) /* C:354-385 */
{
    char base[32];
    snprintf(base, sizeof base, "output_kmer%s", tag);
    char *name = out_name(base, c->k, c->depth_part, part, "tsv");
    FILE *o = fopen(name, "w");
    if (!o)
        die("cannot open kmer dump");
    char kmer[40];
    for (size_t i = 0; i < t->cap; i++)
    {
        if (!t->slot[i].key)
            continue;
        nko_decode(t->slot[i].key, c->k, kmer);
        fprintf(o, "%s\t%d\n", kmer, t->slot[i].count);
    }
    fclose(o);
    free(name);
}

/* fq -> fa header rewrite                                             C:852-876 */
static void emit_fasta(FILE *o, const char *hdr, const char *seq, int fwd)
{
    const char *sfx = fwd ? "/1" : "/2";
    size_t n = strlen(hdr);
    fputc('>', o);
    if (n > 0)
        fputs(hdr + 1, o);
    if (n < 2 || strcmp(hdr + n - 2, sfx) != 0)
        fputs(sfx, o);
    fputc('\n', o);
    fputs(seq, o);
    fputc('\n', o);
}

static void seed_from_file(const nko_cfg *c, nko_table *t, const char *path, int want) /* C:1322-1373 */
{
    nko_file f;
    if (file_open(&f, path) < 0 || f.size == 0)
        return;
    int per = c->in_fastq ? 4 : 2;
    int line = 0, done = 0;
    size_t line_start = 0, seq_start = 0, seq_len = 0;
    char *buf = malloc(f.size + 1);
    for (size_t i = 0; i < f.size; i++)
    {
        if (f.data[i] != '\n')
            continue;
        if (line == 1)
        {
            seq_start = line_start;
            seq_len = i - line_start;
        }
        line++;
        line_start = i + 1;
        if (line < per)
            continue;
        line = 0;
        size_t slen = strnlen(f.data + seq_start, seq_len);
        if (slen > (size_t)c->k) /* strictly longer than K, C:1347 */
        {
            memcpy(buf, f.data + seq_start, slen);
            if (!scrub_and_check(buf, (int)slen))
            {
                buf[slen] = 0;
                fprintf(stderr, "FATAL: FWD sequence does not appear to be a DNA sequence\n%s\n\n", buf);
                exit(EXIT_FAILURE);
            }
            seed_sequence(t, buf, (int)slen, c->k, c->canonical);
            if (++done == want)
                break;
        }
    }
    free(buf);
    file_close(&f);
}

/* one partition's pass over its byte range of one (pair of) file(s)   C:1568-1770, C:1921-2111 */
static void run_range(const nko_cfg *c, nko_part *p, const nko_file *ff, size_t fs, size_t fe,
                      const nko_file *rf, size_t rs, size_t re)
{
    int per = c->in_fastq ? 4 : 2;
    int paired = rf != NULL;
    char fl[4][NKO_LINE_MAX], rl[4][NKO_LINE_MAX];
    int fn[4], rn[4];
    size_t fp = fs, rp = rs;
    while (fp < fe && (!paired || rp < re))
    {
        int more = 1, complete = 1;
        for (int i = 0; i < per; i++)
        {
            int mf = 1, mr = 1;
            fp = take_line(ff, fp, fl[i], &fn[i], &mf);
            if (paired)
                rp = take_line(rf, rp, rl[i], &rn[i], &mr);
            if (!mf || !mr)
            {
                more = 0;
                complete = (i >= 1); /* both sequence lines are in: the reference scores the record, C:1629 */
                for (int j = i + 1; j < per; j++)
                {
                    fl[j][0] = rl[j][0] = '\0';
                    fn[j] = rn[j] = -1; /* not read: not printed (D2) */
                }
                break;
            }
        }
        if (!complete)
            break; /* D2 */
        /* N->A first, then the length gate, then the alphabet gate  C:1424-1457 */
        int okf = scrub_and_check(fl[1], fn[1]);
        int okr = paired ? scrub_and_check(rl[1], rn[1]) : 1;
        if (fn[1] < c->k || (paired && rn[1] < c->k))
        {
            if (!more)
                break; /* D3 */
            continue;
        }
        if (!okf)
        {
            fprintf(stderr, "FATAL: FWD sequence does not appear to be a DNA sequence\n%s\n\n", fl[1]);
            exit(EXIT_FAILURE);
        }
        if (!okr)
        {
            fprintf(stderr, "FATAL: REV sequence does not appear to be a DNA sequence\n%s\n\n", rl[1]);
            exit(EXIT_FAILURE);
        }
        int hf = 0, tf = 0, hr = 0, tr = 0;
        nko_score(p->table, fl[1], fn[1], c->k, c->canonical, c->depth_part, &hf, &tf);
        if (paired)
            nko_score(p->table, rl[1], rn[1], c->k, c->canonical, c->depth_part, &hr, &tr);
        p->processed++;
        int keep = nko_keep_mate(hf, tf, c->coverage) && (!paired || nko_keep_mate(hr, tr, c->coverage));
        if (keep)
        {
            if (c->in_fastq && !c->out_fastq)
            {
                if (paired) /* single-end fq->fa writes nothing, C:1995-1999 */
                {
                    emit_fasta(p->out_f, fl[0], fl[1], 1);
                    emit_fasta(p->out_r, rl[0], rl[1], 0);
                }
            }
            else
            {
                for (int i = 0; i < per && fn[i] >= 0; i++)
                {
                    fprintf(p->out_f, "%s\n", fl[i]);
                    if (paired)
                        fprintf(p->out_r, "%s\n", rl[i]);
                }
            }
            p->printed++;
        }
        else
            p->skipped++;
        if (!more)
            break;
    }
}

static void usage(void)
{
    fprintf(stderr, "Usage: nk_oracle -f fwd [fwd2..] -r rev [rev2..] [-s] [-k 5-31] [-d depth] [-g coverage] [-c]\n"
                    "                 [-t fq|fa] [-o fq|fa] [-m GB] [-p partitions] [-e] [-b level] [-P] [-v]\n");
}

static int is_fa(const char *s) { return !strcasecmp(s, "fa") || !strcasecmp(s, "fasta") || !strcasecmp(s, "fsa") || !strcasecmp(s, "fas"); }
static int is_fq(const char *s) { return !strcasecmp(s, "fq") || !strcasecmp(s, "fastq") || !strcasecmp(s, "fsq"); }

static void add_files(char ***list, int *n, char *first, char **argv, int *idx) /* C:747-832 */
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

static int parse(nko_cfg *c, int argc, char **argv) /* C:520-745 */
{
    memset(c, 0, sizeof *c);
    c->coverage = 0.9;
    c->parts = 1;
    c->k = 15;
    c->depth = 100;
    c->in_fastq = c->out_fastq = 1;
    static struct option lo[] = {
        {"forward", 1, 0, 'f'}, {"reverse", 1, 0, 'r'}, {"ksize", 1, 0, 'k'}, {"depth", 1, 0, 'd'}, {"coverage", 1, 0, 'g'}, {"filetype", 1, 0, 't'}, {"outformat", 1, 0, 'o'}, {"cpu", 1, 0, 'p'}, {"memory_start", 1, 0, 'm'}, {"debug", 1, 0, 'b'}, {"verbose", 0, 0, 'e'}, {"help", 0, 0, 'h'}, {"canonical", 0, 0, 'c'}, {"version", 0, 0, 'v'}, {"single", 0, 0, 's'}, {"print", 0, 0, 'P'}, {0, 0, 0, 0}};
    int o;
    optind = 1;
    while ((o = getopt_long(argc, argv, "f:r:k:d:g:t:o:p:m:b:ehcvsP", lo, NULL)) != -1)
    {
        switch (o)
        {
        case 'P':
            c->dump = 1;
            break;
        case 's':
            c->single = 1;
            break;
        case 'c':
            c->canonical = 1;
            break;
        case 'm':
            c->memory = atoi(optarg);
            if (c->memory < 1)
            {
                printf("Memory cannot be less than 1 Gb %'d\n", c->memory);
                return 0;
            }
            break;
        case 'b':
            c->debug = atoi(optarg);
            break;
        case 'h':
            usage();
            exit(EXIT_SUCCESS);
        case 'p':
            c->parts = atoi(optarg);
            break;
        case 'f':
            add_files(&c->fwd, &c->nfwd, optarg, argv, &optind);
            break;
        case 'r':
            add_files(&c->rev, &c->nrev, optarg, argv, &optind);
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
            printf("%d\n", NKO_VERSION);
            exit(EXIT_SUCCESS);
        case 'e':
            c->verbose = 1;
            break;
        case 't':
            if (is_fa(optarg))
                c->in_fastq = 0;
            else if (is_fq(optarg))
                c->in_fastq = 1;
            else
            {
                printf("Input file format must be either fa or fq, not %s\n", optarg);
                return 0;
            }
            break;
        case 'o':
            if (is_fa(optarg))
                c->out_fastq = 0;
            else if (is_fq(optarg))
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
    if (c->parts <= 0)
    { /* the reference divides by cpus before validating it; avoid SIGFPE */
        fprintf(stderr, "Error: CPU count (%d) must be a positive integer and up to %d\n", c->parts, NKO_MAX_PARTS);
        return 0;
    }
    c->depth_part = c->depth / c->parts; /* C:674 */
    c->cap0 = capacity_unclamped(c->memory, c->parts);
    float mem_part = (float)c->cap0 * 16 / 1073741824;
    size_t lim = pow4_wrapping(c->k);
    int mem_total = c->memory;
    if (lim < c->cap0)
    {
        c->cap0 = lim;
        mem_part = (float)c->cap0 * 16 / 1073741824;
        mem_total = (int)(mem_part * c->parts);
    }
    printf("Initial hash table size set to %'zu (maximum for k=%d is %'zu); memory ~ %'0.2f Gb for each of %d threads (~ %'d Gb total))\n\n",
           c->cap0, c->k, lim, mem_part, c->parts, mem_total);
    if (c->nfwd == 0 || (c->nrev == 0 && !c->single))
    {
        fprintf(stderr, "Error: no fwd (%d) or reverse (%d) files provided\n", c->nfwd, c->nrev);
        return 0;
    }
    if (!c->in_fastq && c->out_fastq)
    {
        fprintf(stderr, "Error: cannot request an output format of FASTQ when input is FASTA\n");
        return 0;
    }
    if (!c->single && c->nfwd != c->nrev)
    {
        fprintf(stderr, "Error: Number of forward (%d) and reverse files (%d) must match\n", c->nfwd, c->nrev);
        return 0;
    }
    if (c->parts > NKO_MAX_PARTS)
    {
        fprintf(stderr, "Error: CPU count (%d) must be a positive integer and up to %d\n", c->parts, NKO_MAX_PARTS);
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
    if (c->depth_part < 2)
    {
        fprintf(stderr, "Error: Depth (%d) must be at least 2 x number of CPUs (for performance reasons; but this version of the program is written to normalise to 50+\n", c->depth);
        return 0;
    }
    return 1;
}

/* whole program, sequential over partitions                           C:2223-2455 */
int nko_main(int argc, char **argv)
{
    setlocale(LC_ALL, "");
    nko_cfg c;
    if (!parse(&c, argc, argv))
    {
        usage();
        return 1;
    }
    nko_table *seed = nko_table_new(c.cap0);
    int want = 1 + (int)(3e6 / c.nfwd); /* C:2242 */
    for (int i = 0; i < c.nfwd; i++)
    {
        seed_from_file(&c, seed, c.fwd[i], want);
        if (i < c.nrev)
            seed_from_file(&c, seed, c.rev[i], want);
    }
    if (c.dump)
        dump_table(&c, seed, "_seeds", -1);
    uint64_t seed_ops = seed->n_ops, seed_touches = seed->n_touches;

    nko_part *part = calloc((size_t)c.parts, sizeof *part);
    for (int t = 0; t < c.parts; t++)
    {
        part[t].table = nko_table_clone(seed);
        char *n = out_name("output_forward", c.k, c.depth_part, t, "fastq");
        part[t].out_f = fopen(n, "w");
        if (!part[t].out_f)
            die("Error opening file to write");
        free(n);
        if (c.nrev)
        {
            n = out_name("output_reverse", c.k, c.depth_part, t, "fastq");
            part[t].out_r = fopen(n, "w");
            if (!part[t].out_r)
                die("Error opening file to write");
            free(n);
        }
    }
    nko_table_free(seed);

    time_t t0 = time(NULL);
    size_t tot_proc = 0, tot_print = 0, tot_skip = 0, max_used_all = 0;
    int map_failed = 0;
    size_t *fs = calloc((size_t)c.parts, sizeof(size_t)), *fe = calloc((size_t)c.parts, sizeof(size_t));
    size_t *rs = calloc((size_t)c.parts, sizeof(size_t)), *re = calloc((size_t)c.parts, sizeof(size_t));
    for (int i = 0; i < c.nfwd; i++)
    {
        int paired = i < c.nrev;
        nko_file ff, rf;
        if (paired)
            printf("Processing file pair %d of %d: %s and %s\n", i + 1, c.nfwd, c.fwd[i], c.rev[i]);
        else
            printf("Processing single-ended file %d of %d: %s\n", i + 1, c.nfwd, c.fwd[i]);
        if (file_open(&ff, c.fwd[i]) < 0 || (paired && file_open(&rf, c.rev[i]) < 0) || ff.size == 0 || (paired && rf.size == 0))
        {
            fprintf(stderr, "Error memory mapping input files\n");
            map_failed = 1;
            break;
        }
        char lead = c.in_fastq ? '@' : '>';
        if (ff.data[0] != lead)
        {
            fprintf(stderr, "Input %s file %s starts with %c which is not expected\n", c.in_fastq ? "FASTQ" : "FASTA", c.fwd[i], ff.data[0]);
            exit(EXIT_FAILURE);
        }
        if (paired && rf.data[0] != lead)
        {
            fprintf(stderr, "Input %s file %s starts with %c which is not expected\n", c.in_fastq ? "FASTQ" : "FASTA", c.rev[i], rf.data[0]);
            exit(EXIT_FAILURE);
        }
        memset(fs, 0, sizeof(size_t) * (size_t)c.parts);
        memset(fe, 0, sizeof(size_t) * (size_t)c.parts);
        memset(rs, 0, sizeof(size_t) * (size_t)c.parts);
        memset(re, 0, sizeof(size_t) * (size_t)c.parts);
        if (c.parts == 1)
        { /* C:1796-1803 */
            fe[0] = ff.size - 1;
            if (paired)
                re[0] = rf.size - 1;
        }
        else if (!paired)
            nko_ranges_by_size(&ff, c.parts, c.in_fastq, fs, fe); /* C:2142 */
        else if (ff.size == rf.size)
        { /* C:1807-1813 */
            nko_ranges_by_size(&ff, c.parts, c.in_fastq, fs, fe);
            nko_ranges_by_size(&rf, c.parts, c.in_fastq, rs, re);
        }
        else
        { /* C:1815-1828: the forward record count drives both files */
            size_t recs = nko_count_records(&ff, c.in_fastq);
            nko_ranges_by_records(&ff, c.parts, c.in_fastq, recs, fs, fe);
            nko_ranges_by_records(&rf, c.parts, c.in_fastq, recs, rs, re);
        }
        size_t max_used = 0;
        tot_proc = tot_print = tot_skip = 0;
        for (int t = 0; t < c.parts; t++)
        {
            run_range(&c, &part[t], &ff, fs[t], fe[t], paired ? &rf : NULL, rs[t], re[t]);
            printf("Thread %d - processed %'zu pairs, printed: %'zu, skipped: %'zu, Unique kmers (all sequences; this thread): %'zu\n",
                   t, part[t].processed, part[t].printed, part[t].skipped, part[t].table->used);
            tot_proc += part[t].processed;
            tot_print += part[t].printed;
            tot_skip += part[t].skipped;
            if (part[t].table->used > max_used)
                max_used = part[t].table->used;
        }
        if (max_used > max_used_all)
            max_used_all = max_used;
        printf("Cumulative file statistics: Processed %'zu, Printed %'zu, Skipped %'zu, Cumulative Max Unique Kmers in a thread: %'zu\n",
               tot_proc, tot_print, tot_skip, max_used);
        file_close(&ff);
        if (paired)
            file_close(&rf);
    }
    uint64_t ops = 0, touches = 0, walks = 0, expansions = 0;
    for (int t = 0; t < c.parts; t++)
    {
        fclose(part[t].out_f);
        if (c.nrev)
            fclose(part[t].out_r);
        if (c.dump)
            dump_table(&c, part[t].table, "", t);
        ops += part[t].table->n_ops;
        touches += part[t].table->n_touches;
        walks += part[t].table->n_walk_ops;
        expansions += part[t].table->n_expansions;
        nko_table_free(part[t].table);
    }
    if (!map_failed)
    {
        printf("\n--- Final Report ---\n");
        printf("Processed Records: %'zu\n", tot_proc);
        printf("Printed Records: %'zu\n", tot_print);
        printf("Skipped Records: %'zu\n", tot_skip);
        printf("Cumulative Max unique kmers in any thread: %'zu\n", max_used_all);
    }
    double dt = difftime(time(NULL), t0);
    printf("Total runtime: %.2f seconds\n", dt);
    if (tot_proc > 0)
        printf("Overall processing rate: %'.0f %s per second\n", tot_proc / dt, c.nrev ? "sequence pairs" : "sequences");
    else
        printf("No data processed\n");
    /* oracle-only instrumentation (ops and slot touches) */
    printf("ORACLE seed_ops=%llu seed_touches=%llu ops=%llu touches=%llu walk_ops=%llu expansions=%llu\n",
           (unsigned long long)seed_ops, (unsigned long long)seed_touches, (unsigned long long)ops,
           (unsigned long long)touches, (unsigned long long)walks, (unsigned long long)expansions);
    return 0;
}

#ifndef NKO_LIBRARY
int main(int argc, char **argv)
{
    return nko_main(argc, argv);
}
#endif
