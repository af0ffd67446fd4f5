/*
 * nk_synth.c -- seeded synthetic paired-end RNA-Seq generator for the benchmarks and parity tests
 * (SURVEY.md 8(d)): T transcripts of length U[400,4000] over uniform ACGT; expression proportional to
 * lognormal(0, sigma=2) x length (polar-method normals with libm-free log/exp); fragment length N(300,50) clipped to [L, transcript]; read 1 =
 * fragment[:L], read 2 = revcomp(fragment)[:L]; 0.5 % substitutions; one N in 1 % of read 1; quality 'I';
 * names of variable width so that the two files differ in size (the reference then takes its
 * record-count partitioner, C:1815-1828) unless --equal is given (C:1807-1813 path).
 *
 * Library (ctypes) and CLI:  nk_synth -n pairs -o prefix [-s seed] [-t transcripts] [-L readlen] [--equal] [--fasta]
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
    uint64_t s[2];
} rng_t;

static inline uint64_t rng_next(rng_t *r)
{ /* xoroshiro128+ */
    uint64_t s0 = r->s[0], s1 = r->s[1], res = s0 + s1;
    s1 ^= s0;
    r->s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
    r->s[1] = (s1 << 37) | (s1 >> 27);
    return res;
}
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }
/* No libm on purpose: the bytes this generator writes are part of committed golden vectors
 * (tests/golden), so they must not depend on a host's exp/log/cos implementation.  Only + - * / and exact
 * operations are used (IEEE double, no contraction). */
static double det_log(double x)
{ /* x = m * 2^e with m in [sqrt(1/2), sqrt(2)); log m = 2 atanh((m-1)/(m+1)) by its series */
    int e = 0;
    while (x >= 1.4142135623730951)
    {
        x *= 0.5;
        e++;
    }
    while (x < 0.7071067811865476)
    {
        x *= 2.0;
        e--;
    }
    double z = (x - 1.0) / (x + 1.0), z2 = z * z, term = z, sum = 0.0;
    for (int i = 1; i <= 41; i += 2)
    {
        sum += term / (double)i;
        term *= z2;
    }
    return 2.0 * sum + (double)e * 0.6931471805599453;
}
static double rng_normal(rng_t *r)
{ /* Marsaglia polar method; sqrt is the correctly rounded hardware instruction */
    for (;;)
    {
        double u = 2.0 * rng_unit(r) - 1.0, v = 2.0 * rng_unit(r) - 1.0, q = u * u + v * v;
        if (q >= 1.0 || q < 1e-300)
            continue;
        return u * __builtin_sqrt(-2.0 * det_log(q) / q);
    }
}
static double det_exp(double x)
{ /* exp(x) = 2^k * exp(rem), rem in [-ln2/2, ln2/2], Taylor to degree 14 (error < 1e-16) */
    const double ln2 = 0.6931471805599453;
    double kf = x / ln2;
    long k = (long)(kf >= 0 ? kf + 0.5 : kf - 0.5);
    double rem = x - (double)k * ln2, term = 1.0, sum = 1.0;
    for (int i = 1; i <= 14; i++)
    {
        term = term * rem / (double)i;
        sum += term;
    }
    double scale = 1.0;
    for (long i = 0; i < (k < 0 ? -k : k); i++)
        scale *= 2.0;
    return k >= 0 ? sum * scale : sum / scale;
}
static void rng_seed(rng_t *r, uint64_t seed)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 2; i++)
    {
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        r->s[i] = x ^ (x >> 31);
    }
}

typedef struct
{
    char *buf;
    size_t len, cap;
} out_t;

static void out_reserve(out_t *o, size_t extra)
{
    if (o->len + extra <= o->cap)
        return;
    size_t nc = o->cap ? o->cap : (1u << 20);
    while (nc < o->len + extra)
        nc += nc / 2;
    o->buf = realloc(o->buf, nc);
    o->cap = nc;
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char b) { return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : 'N'; }

static size_t put_name(char *dst, char lead, uint64_t idx, rng_t *r, int equal, int mate, uint32_t extra_fixed)
{
    size_t n = (size_t)sprintf(dst, "%csyn.%llu ", lead, (unsigned long long)idx);
    uint32_t extra = equal ? extra_fixed : 1 + rng_below(r, 20);
    for (uint32_t i = 0; i < extra; i++)
        dst[n++] = (char)('a' + rng_below(r, 26));
    dst[n++] = '/';
    dst[n++] = (char)('0' + mate);
    dst[n++] = '\n';
    return n;
}

/* Generates n_pairs records into two malloc'd buffers (caller frees with nk_synth_free). */
int nk_synth_generate(uint64_t n_pairs, uint64_t seed, uint32_t n_transcripts, uint32_t read_len, int equal_sizes,
                      int fasta, char **fwd, size_t *fwd_size, char **rev, size_t *rev_size)
{
    rng_t r;
    rng_seed(&r, seed);
    if (n_transcripts == 0)
    { /* keep table growth realistic: about 500 pairs per transcript, at least 200 transcripts */
        n_transcripts = (uint32_t)(n_pairs / 500);
        if (n_transcripts < 200)
            n_transcripts = 200;
    }
    if (read_len < 20)
        read_len = 150;
    uint32_t *tlen = malloc(sizeof(uint32_t) * n_transcripts);
    size_t *toff = malloc(sizeof(size_t) * (n_transcripts + 1));
    double *cum = malloc(sizeof(double) * n_transcripts);
    size_t total = 0;
    uint32_t min_len = read_len + 250 > 400 ? read_len + 250 : 400;
    for (uint32_t t = 0; t < n_transcripts; t++)
    {
        tlen[t] = min_len + rng_below(&r, 4000 - 400 + 1);
        toff[t] = total;
        total += tlen[t];
    }
    toff[n_transcripts] = total;
    char *genome = malloc(total);
    for (size_t i = 0; i < total; i++)
        genome[i] = BASES[rng_next(&r) >> 62];
    double acc = 0;
    for (uint32_t t = 0; t < n_transcripts; t++)
    {
        acc += det_exp(2.0 * rng_normal(&r)) * (double)tlen[t];
        cum[t] = acc;
    }
    out_t of = {0}, orv = {0};
    size_t per_rec = (size_t)read_len * 2 + 64;
    out_reserve(&of, n_pairs * per_rec / 8 + (1u << 20));
    out_reserve(&orv, n_pairs * per_rec / 8 + (1u << 20));
    char *r1 = malloc(read_len + 1), *r2 = malloc(read_len + 1);
    char lead = fasta ? '>' : '@';
    for (uint64_t i = 0; i < n_pairs; i++)
    {
        double x = rng_unit(&r) * acc;
        uint32_t lo = 0, hi = n_transcripts - 1;
        while (lo < hi)
        {
            uint32_t mid = (lo + hi) / 2;
            if (cum[mid] < x)
                lo = mid + 1;
            else
                hi = mid;
        }
        uint32_t L = tlen[lo];
        int frag = (int)(300.0 + 50.0 * rng_normal(&r));
        if (frag < (int)read_len)
            frag = (int)read_len;
        if (frag > (int)L)
            frag = (int)L;
        uint32_t start = rng_below(&r, L - (uint32_t)frag + 1);
        const char *f = genome + toff[lo] + start;
        for (uint32_t b = 0; b < read_len; b++)
        {
            r1[b] = f[b];
            r2[b] = comp(f[frag - 1 - (int)b]);
        }
        for (uint32_t b = 0; b < read_len; b++)
        { /* 0.5 % substitutions per base, independently in both mates */
            if (rng_below(&r, 200) == 0)
                r1[b] = BASES[rng_below(&r, 4)];
            if (rng_below(&r, 200) == 0)
                r2[b] = BASES[rng_below(&r, 4)];
        }
        if (rng_below(&r, 100) == 0)
            r1[rng_below(&r, read_len)] = 'N';
        out_reserve(&of, per_rec);
        out_reserve(&orv, per_rec);
        uint32_t fixed = 1 + (uint32_t)(i % 20);
        of.len += put_name(of.buf + of.len, lead, i, &r, equal_sizes, 1, fixed);
        orv.len += put_name(orv.buf + orv.len, lead, i, &r, equal_sizes, 2, fixed);
        memcpy(of.buf + of.len, r1, read_len);
        of.len += read_len;
        of.buf[of.len++] = '\n';
        memcpy(orv.buf + orv.len, r2, read_len);
        orv.len += read_len;
        orv.buf[orv.len++] = '\n';
        if (!fasta)
        {
            of.buf[of.len++] = '+';
            of.buf[of.len++] = '\n';
            memset(of.buf + of.len, 'I', read_len);
            of.len += read_len;
            of.buf[of.len++] = '\n';
            orv.buf[orv.len++] = '+';
            orv.buf[orv.len++] = '\n';
            memset(orv.buf + orv.len, 'I', read_len);
            orv.len += read_len;
            orv.buf[orv.len++] = '\n';
        }
    }
    free(r1);
    free(r2);
    free(genome);
    free(tlen);
    free(toff);
    free(cum);
    *fwd = of.buf;
    *fwd_size = of.len;
    *rev = orv.buf;
    *rev_size = orv.len;
    return 0;
}

void nk_synth_free(char *p) { free(p); }

#ifndef NK_SYNTH_LIBRARY
int main(int argc, char **argv)
{
    uint64_t n = 100000, seed = 1;
    uint32_t nt = 0, L = 150;
    int equal = 0, fasta = 0;
    const char *prefix = "synth";
    for (int i = 1; i < argc; i++)
    {
        if (!strcmp(argv[i], "-n") && i + 1 < argc)
            n = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc)
            seed = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc)
            nt = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-L") && i + 1 < argc)
            L = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc)
            prefix = argv[++i];
        else if (!strcmp(argv[i], "--equal"))
            equal = 1;
        else if (!strcmp(argv[i], "--fasta"))
            fasta = 1;
        else
        {
            fprintf(stderr, "usage: nk_synth -n pairs -o prefix [-s seed] [-t transcripts] [-L readlen] [--equal] [--fasta]\n");
            return 1;
        }
    }
    char *f, *r;
    size_t fs, rs;
    nk_synth_generate(n, seed, nt, L, equal, fasta, &f, &fs, &r, &rs);
    char name[4096];
    const char *ext = fasta ? "fasta" : "fastq";
    snprintf(name, sizeof name, "%s_1.%s", prefix, ext);
    FILE *o = fopen(name, "w");
    if (!o || fwrite(f, 1, fs, o) != fs)
        return 2;
    fclose(o);
    snprintf(name, sizeof name, "%s_2.%s", prefix, ext);
    o = fopen(name, "w");
    if (!o || fwrite(r, 1, rs, o) != rs)
        return 2;
    fclose(o);
    fprintf(stderr, "wrote %llu pairs: %zu + %zu bytes\n", (unsigned long long)n, fs, rs);
    return 0;
}
#endif
