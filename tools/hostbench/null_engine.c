/*
 * null_engine.c -- DEVELOPMENT TOOL ONLY: an nkd_* engine that computes nothing (accepts every record,
 * or a pseudo-random ~48 % when NK_NULL_ACCEPT is set) so that the C host pipeline of nk_host.c --
 * line indexing, staging, writing -- can be timed and profiled on a machine without a GPU.
 * Never linked into the product library or the tests' parity paths.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/nk_b200.h"

struct nkd_engine
{
    nkd_config cfg;
    size_t n_records;
    int paired;
    int pct;
};

int nkd_create(const nkd_config *c, nkd_engine **out)
{
    nkd_engine *e = calloc(1, sizeof *e);
    e->cfg = *c;
    e->pct = getenv("NK_NULL_ACCEPT") ? atoi(getenv("NK_NULL_ACCEPT")) : 100;
    *out = e;
    return 0;
}
void nkd_destroy(nkd_engine *e) { free(e); }
const char *nkd_last_error(const nkd_engine *e) { (void)e; return ""; }
int nkd_seed_step(nkd_engine *e, const uint8_t *s, size_t sb, const nkd_read *r, size_t n, int64_t *inv)
{
    (void)e; (void)s; (void)sb; (void)r; (void)n;
    if (inv) *inv = -1;
    return 0;
}
int nkd_seed_finish(nkd_engine *e) { (void)e; return 0; }
int nkd_seed_finish_from(nkd_engine *e, nkd_engine *s) { (void)e; (void)s; return 0; }
int nkd_seed_stats(nkd_engine *e, nkd_part_stats *st) { memset(st, 0, sizeof *st); st->capacity = e->cfg.capacity0; return 0; }
int nkd_stage_segments(nkd_engine *e, const uint8_t *seq, const nkd_segment *segs, int n, int paired)
{
    (void)seq;
    size_t reads = 0;
    for (int i = 0; i < n; i++) reads += segs[i].n_reads;
    e->paired = paired;
    e->n_records = paired ? reads / 2 : reads;
    return 0;
}
int nkd_run(nkd_engine *e) { (void)e; return 0; }
int nkd_fetch(nkd_engine *e, uint8_t *accept, size_t n, int64_t *inv)
{
    static uint32_t x = 12345;
    for (size_t i = 0; i < n; i++)
    {
        x = x * 1664525u + 1013904223u;
        accept[i] = (x >> 8) % 100 < (uint32_t)e->pct;
    }
    if (inv) *inv = -1;
    return 0;
}
int nkd_run_spans(nkd_engine *e, float *s, size_t c, size_t *n) { (void)e; (void)s; (void)c; if (n) *n = 0; return 0; }
int nkd_run_stats_get(nkd_engine *e, nkd_run_stats *o) { (void)e; memset(o, 0, sizeof *o); return 0; }
int nkd_part_stats_get(nkd_engine *e, int p, nkd_part_stats *o) { (void)p; memset(o, 0, sizeof *o); o->capacity = e->cfg.capacity0; return 0; }
int nkd_dump_text(nkd_engine *e, int p, uint64_t f, uint64_t n, char *t, size_t c, size_t *b) { (void)e; (void)p; (void)f; (void)n; (void)t; (void)c; *b = 0; return 0; }
int nkd_compact(nkd_engine *e, int p, uint64_t *k, int64_t *v, uint64_t c, uint64_t *n) { (void)e; (void)p; (void)k; (void)v; (void)c; *n = 0; return 0; }
int nkd_merge_begin(nkd_engine *e, uint64_t m) { (void)e; (void)m; return 0; }
int nkd_merge_add_part(nkd_engine *e, int p) { (void)e; (void)p; return 0; }
int nkd_merge_add(nkd_engine *e, const uint64_t *k, const int64_t *v, uint64_t n) { (void)e; (void)k; (void)v; (void)n; return 0; }
int nkd_merge_finish(nkd_engine *e, uint64_t *n) { (void)e; *n = 0; return 0; }
void *nkd_alloc_pinned(size_t b) { void *p = malloc(b ? b : 16); if (p) memset(p, 0, b); return p; }
void nkd_free_pinned(void *p) { free(p); }
int nkd_device_count(void) { return 1; }
