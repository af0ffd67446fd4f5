/*
 * nk_engine_api.h -- the extern "C" nkd_* entry points (include/nk_b200.h) over NkEngine<NK_BACKEND>.
 * Included once by nk_engine.cu (NK_BACKEND = CudaBackend) and once by tests/emu/nk_emu.cpp
 * (NK_BACKEND = the test-only CPU emulation).
 */
#ifndef NK_ENGINE_API_H
#define NK_ENGINE_API_H

#include "nk_orchestrate.h"

struct nkd_engine
{
    NkEngine<NK_BACKEND> e;
};

static_assert(sizeof(nkd_read) == sizeof(NkRead), "nkd_read layout");

/* a CUDA error anywhere in the call turns the result into NK_ENODEVICE (there is nothing to fall back to) */
static inline int nkd_done(nkd_engine *h, int rc)
{
    std::string m;
    if (h->e.be.failed(m))
    {
        if (rc == NK_OK)
        {
            h->e.err = m;
            rc = NK_ENODEVICE;
        }
    }
    return rc;
}

extern "C" {

int nkd_create(const nkd_config *cfg, nkd_engine **out)
{
    if (!cfg || !out)
        return NK_EINVAL;
    nkd_engine *h = new nkd_engine();
    int rc = h->e.create(*cfg);
    rc = nkd_done(h, rc);
    *out = h; /* kept on failure so that nkd_last_error can be read; caller still destroys it */
    return rc;
}

void nkd_destroy(nkd_engine *h)
{
    if (!h)
        return;
    h->e.be.enter();
    h->e.destroy();
    delete h;
}

const char *nkd_last_error(const nkd_engine *h) { return h ? h->e.err.c_str() : "null engine"; }

int nkd_seed_step(nkd_engine *h, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads,
                  int64_t *first_invalid)
{
    h->e.be.enter();
    return nkd_done(h, h->e.seed_step(seq, seq_bytes, reads, n_reads, first_invalid));
}
int nkd_stage_segments(nkd_engine *h, const uint8_t *seq_base, const nkd_segment *segs, int n_segs, int paired)
{
    h->e.be.enter();
    if (!h->e.seeded)
        return h->e.fail(NK_EINVAL, "nkd_stage_segments before nkd_seed_finish");
    return nkd_done(h, h->e.stage_segments(seq_base, segs, n_segs, paired, h->e.cfg.n_parts, false));
}
int nkd_seed_finish(nkd_engine *h) {
    h->e.be.enter(); return nkd_done(h, h->e.seed_finish()); }
int nkd_seed_finish_from(nkd_engine *h, nkd_engine *src)
{
    h->e.be.enter();
    if (!src || src == h)
        return h->e.fail(NK_EINVAL, "nkd_seed_finish_from: bad source engine");
    return nkd_done(h, h->e.seed_finish_from(src->e));
}
int nkd_seed_stats(nkd_engine *h, nkd_part_stats *out)
{
    *out = h->e.seed.st;
    out->capacity = h->e.seed.cap;
    out->used = h->e.seed.used;
    return NK_OK;
}
int nkd_seed_export(nkd_engine *h, uint64_t *keys, int32_t *counts, uint64_t capacity)
{
    h->e.be.enter();
    return nkd_done(h, h->e.export_table(h->e.seed, keys, counts, capacity));
}
int nkd_stage(nkd_engine *h, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads, int paired)
{
    h->e.be.enter();
    if (!h->e.seeded)
        return h->e.fail(NK_EINVAL, "nkd_stage before nkd_seed_finish");
    return nkd_done(h, h->e.stage(seq, seq_bytes, reads, n_reads, paired, h->e.cfg.n_parts, false));
}
int nkd_run(nkd_engine *h) {
    h->e.be.enter(); return nkd_done(h, h->e.run_step()); }
int nkd_fetch(nkd_engine *h, uint8_t *accept, size_t n_records, int64_t *first_invalid)
{
    h->e.be.enter();
    return nkd_done(h, h->e.fetch(accept, n_records, first_invalid));
}
int nkd_set_table_budget(nkd_engine *h, uint64_t bytes)
{
    if (h->e.seeded)
        return h->e.fail(NK_EINVAL, "nkd_set_table_budget after nkd_seed_finish");
    h->e.table_budget = bytes;
    return NK_OK;
}
int nkd_residency_stats(nkd_engine *h, uint64_t *resident_parts, uint64_t *evictions, uint64_t *loads)
{
    uint64_t r = 0;
    for (auto &p : h->e.parts)
        r += p.tab != nullptr;
    if (resident_parts)
        *resident_parts = r;
    if (evictions)
        *evictions = h->e.evictions;
    if (loads)
        *loads = h->e.loads;
    return NK_OK;
}
int nkd_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes)
{
    return NK_BACKEND::device_memory(device, free_bytes, total_bytes) ? NK_OK : NK_ENODEVICE;
}
int nkd_stage_raw(nkd_engine *h, const uint8_t *raw, size_t raw_bytes, const nkd_raw_segment *segs, int n_segs, int paired,
                  int lines_per_record)
{
    h->e.be.enter();
    return nkd_done(h, h->e.stage_raw(raw, raw_bytes, segs, n_segs, paired, lines_per_record));
}
int nkd_seed_raw(nkd_engine *h, const uint8_t *raw, size_t text_bytes, uint32_t n_records, int lines_per_record, uint32_t limit,
                 uint32_t *taken, int64_t *first_invalid)
{
    h->e.be.enter();
    unsigned t = 0;
    int rc = nkd_done(h, h->e.seed_raw(raw, text_bytes, n_records, lines_per_record, limit, &t, first_invalid));
    if (taken)
        *taken = t;
    return rc;
}
int nkd_upload_raw(nkd_engine *h, const uint8_t *raw, size_t raw_bytes)
{
    h->e.be.enter();
    return nkd_done(h, h->e.upload_raw(raw, raw_bytes));
}
int nkd_fetch_raw_slot(nkd_engine *h, int emit_mode, uint8_t *out, size_t out_cap, nkd_raw_result *results,
                       int64_t *first_invalid, int slot)
{
    h->e.be.enter();
    if (slot < 0 || slot >= NKD_FETCH_SLOTS)
        return h->e.fail(NK_EINVAL, "nkd_fetch_raw_slot: bad slot");
    return nkd_done(h, h->e.fetch_raw(emit_mode, out, out_cap, results, first_invalid, slot));
}
int nkd_fetch_raw(nkd_engine *h, int emit_mode, uint8_t *out, size_t out_cap, nkd_raw_result *results, int64_t *first_invalid)
{
    /* not through nkd_fetch_raw_slot: an exported name may resolve into another library that exports it too (the tests
     * load the CUDA library and the CPU emulation of the engine side by side) */
    h->e.be.enter();
    return nkd_done(h, h->e.fetch_raw(emit_mode, out, out_cap, results, first_invalid, 0));
}
int nkd_fetch_wait(nkd_engine *h, int slot)
{
    h->e.be.enter();
    if (slot < 0 || slot >= NKD_FETCH_SLOTS)
        return h->e.fail(NK_EINVAL, "nkd_fetch_wait: bad slot");
    h->e.be.copy_wait(slot);
    return nkd_done(h, NK_OK);
}
int nkd_last_run_ms(nkd_engine *h, float *total_ms, float *probe_ms)
{
    if (total_ms)
        *total_ms = h->e.last_total_ms;
    if (probe_ms)
        *probe_ms = h->e.last_probe_ms;
    return NK_OK;
}
int nkd_run_stats_get(nkd_engine *h, nkd_run_stats *out)
{
    *out = h->e.rs;
    out->launches = h->e.be.launches;
    out->h2d_bytes = h->e.h2d_bytes + h->e.upload_total;
    out->d2h_bytes = h->e.d2h_bytes;
    return NK_OK;
}
int nkd_run_spans(nkd_engine *h, float *spans, size_t cap_spans, size_t *n_spans)
{
    size_t n = h->e.spans.size() / 2;
    if (n_spans)
        *n_spans = n;
    if (spans)
        memcpy(spans, h->e.spans.data(), sizeof(float) * 2 * (n < cap_spans ? n : cap_spans));
    return NK_OK;
}
int nkd_part_stats_get(nkd_engine *h, int part, nkd_part_stats *out)
{
    if (part < 0 || part >= (int)h->e.parts.size())
        return h->e.fail(NK_EINVAL, "no such partition");
    *out = h->e.parts[part].st;
    out->capacity = h->e.parts[part].cap;
    out->used = h->e.parts[part].used;
    return NK_OK;
}
int nkd_export(nkd_engine *h, int part, uint64_t *keys, int32_t *counts, uint64_t capacity)
{
    h->e.be.enter();
    if (part < 0 || part >= (int)h->e.parts.size())
        return h->e.fail(NK_EINVAL, "no such partition");
    int rc = h->e.make_resident(std::vector<int>{part});
    if (rc)
        return nkd_done(h, rc);
    return nkd_done(h, h->e.export_table(h->e.parts[part], keys, counts, capacity));
}
int nkd_dump_text(nkd_engine *h, int part, uint64_t first, uint64_t n, char *text, size_t text_cap, size_t *bytes)
{
    h->e.be.enter();
    size_t b = 0;
    int rc = nkd_done(h, h->e.dump_text(part, first, n, text, text_cap, &b));
    if (bytes)
        *bytes = b;
    return rc;
}
int nkd_compact(nkd_engine *h, int part, uint64_t *keys, int64_t *counts, uint64_t cap_entries, uint64_t *n)
{
    h->e.be.enter();
    uint64_t m = 0;
    int rc = nkd_done(h, h->e.compact(part, nullptr, nullptr, keys, counts, cap_entries, &m));
    if (n)
        *n = m;
    return rc;
}
int nkd_merge_begin(nkd_engine *h, uint64_t max_entries)
{
    h->e.be.enter();
    return nkd_done(h, h->e.merge_begin(max_entries));
}
int nkd_merge_add_part(nkd_engine *h, int part)
{
    h->e.be.enter();
    if (part < 0)
        return h->e.fail(NK_EINVAL, "no such partition");
    return nkd_done(h, h->e.merge_add_part(part));
}
int nkd_merge_add(nkd_engine *h, const uint64_t *keys, const int64_t *counts, uint64_t n)
{
    h->e.be.enter();
    return nkd_done(h, h->e.merge_add(keys, counts, n));
}
int nkd_merge_finish(nkd_engine *h, uint64_t *n_unique)
{
    h->e.be.enter();
    return nkd_done(h, h->e.merge_finish(n_unique));
}
int nkd_read_scores(nkd_engine *h, uint32_t *high, uint32_t *total, size_t n_reads)
{
    h->e.be.enter();
    if (n_reads != h->e.n_reads)
        return h->e.fail(NK_EINVAL, "nkd_read_scores: read count differs from the staged step");
    h->e.be.d2h(high, h->e.d_high, n_reads * sizeof(uint32_t));
    h->e.be.d2h(total, h->e.d_total, n_reads * sizeof(uint32_t));
    h->e.be.sync();
    return nkd_done(h, NK_OK);
}
int nkd_extract_keys(nkd_engine *h, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads,
                     uint64_t *keys_out, size_t n_ops, uint8_t *invalid_out)
{
    h->e.be.enter();
    return nkd_done(h, h->e.extract_keys(seq, seq_bytes, reads, n_reads, keys_out, n_ops, invalid_out));
}

} /* extern "C" */

#endif
