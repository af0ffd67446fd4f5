/*
 * nk_emu.cpp -- TEST-ONLY CPU emulation backend for NkEngine (never part of the product library).
 *
 * It executes the same per-operation functions the sm_100a kernels execute (nk_core.h) and the same
 * step orchestration (nk_orchestrate.h), but every "launch" visits its operations in a seeded random
 * order, which is the freedom the GPU has.  Passing the oracle comparison under many seeds shows the
 * parallel restatement is order-independent and equals the reference's sequential semantics.
 * Built by tests/emu/Makefile into tests/emu/libnk_emu.so and loaded by tests/ only.
 */
#include <algorithm>
#include <numeric>
#include <random>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../nomalise_kmers_multi_large_b200/csrc/nk_core.h"

struct EmuBackend
{
    std::mt19937_64 rng;
    unsigned long long launches = 0;
    int init(int, std::string &)
    {
        const char *s = getenv("NK_EMU_SEED");
        rng.seed(s ? strtoull(s, nullptr, 10) : 12345ull);
        return 0;
    }
    void shutdown() {}
    void enter() {}
    const void *device_view_of_host(const void *) { return nullptr; }
    bool failed(std::string &) { return false; }
    void *alloc(size_t n) { return malloc(n ? n : 1); }
    void release(void *p) { free(p); }
    void zero(void *p, size_t n) { memset(p, 0, n); }
    void h2d(void *d, const void *h, size_t n) { memcpy(d, h, n); }
    void put_small(void *d, const void *h, size_t n) { memcpy(d, h, n); }
    void d2h(void *h, const void *d, size_t n) { memcpy(h, d, n); }
    void d2d(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    void sync() {}
    bool prepare_sort(size_t, std::string &) { return true; }
    unsigned min_list_entries() const { return 0; }
    unsigned slow_hole_margin() const { return 0; }
    void chunk_sizes(unsigned *c, unsigned, unsigned, unsigned, unsigned, unsigned) { c[0] = c[1] = c[2] = c[3] = c[4] = 0; }
    void begin_timer(int) {}
    void end_timer(int) {}
    void reset_timer(int) {}
    void timer_spans(int, std::vector<float> &) {}
    float timer_ms(int) { return 0.f; }
    float gap_ms(int, int) { return 0.f; }

    std::vector<unsigned> order(size_t n)
    {
        std::vector<unsigned> v(n);
        std::iota(v.begin(), v.end(), 0u);
        std::shuffle(v.begin(), v.end(), rng);
        return v;
    }

    void probe(const NkRun &P)
    {
        struct Op
        {
            unsigned read, w;
        };
        std::vector<Op> ops;
        for (unsigned r = 0; r < P.n_reads; r++)
        {
            const NkRead &rd = P.reads[r];
            if (P.mode != NK_MODE_COUNT)
                for (unsigned i = 0; i < rd.len; i++)
                {
                    unsigned char b = P.seq[rd.seq_off + i];
                    if (b != 'A' && b != 'C' && b != 'G' && b != 'T' && b != 'N')
                        P.invalid[r] = 1;
                }
            for (unsigned w = 0; w + P.k <= rd.len; w++)
                ops.push_back({r, w});
        }
        std::shuffle(ops.begin(), ops.end(), rng);
        for (auto &op : ops)
        {
            const NkRead &rd = P.reads[op.read];
            unsigned part = (P.mode == NK_MODE_SEED || P.mode == NK_MODE_KEYS) ? 0u : rd.part;
            const NkPart &pd = P.parts[part];
            unsigned t = rd.op_base + op.w;
            if (t < pd.lo || t >= pd.hi)
                continue;
            unsigned long long key = nk_window_key_ascii(P.seq + rd.seq_off + op.w, P.k, P.canonical);
            if (P.mode == NK_MODE_KEYS)
            {
                P.keys_out[t] = key;
                continue;
            }
            if (key == 0)
                continue;
            P.ctr->real_ops[part] += 1;
            if (P.mode == NK_MODE_COUNT)
                continue;
            P.total[op.read] += (unsigned)P.delta;
            int high = 0;
            unsigned hot = 0;
            unsigned tch = nk_probe_op(P, pd, part, key, t, op.read, high, hot);
            P.ctr->touches[part] += tch;
            P.ctr->probe_touches += tch;
            P.ctr->hot_hits += hot;
            P.high[op.read] += (unsigned)high;
        }
    }
    void hot_flush(const NkRun &P)
    {
        for (unsigned i : order((size_t)P.hot_mask + 1))
            nk_hot_flush_op(P, i);
    }
    void hot_clear(const NkRun &P, unsigned part1)
    {
        for (unsigned i = 0; i <= P.hot_mask; i++)
            nk_hot_clear_op(P, i, part1);
    }
    void prepare_claims(const NkRun &P)
    {
        for (unsigned idx : order(std::min(P.ctr->n_open, P.open_cap)))
            nk_prepare_claim_op(P, idx);
    }
    void open_ops(const NkRun &P)
    {
        unsigned n = std::min(P.ctr->n_open, P.open_cap);
        for (unsigned idx : order(n))
        {
            int high = 0, claimed = 0;
            unsigned touches = nk_open_op(P, idx, high, claimed);
            P.ctr->touches[P.open[idx].part] += touches;
            P.ctr->claims[P.open[idx].part] += (unsigned)claimed;
            P.high[P.open[idx].read] += (unsigned)high;
        }
    }
    void apply(const NkRun &P, unsigned n)
    {
        for (unsigned i : order(n))
            nk_apply_op(P, i);
    }
    void classify(const NkRun &P, unsigned n)
    {
        for (unsigned i : order(n))
        {
            NkPend r;
            unsigned long long g;
            int x;
            if (nk_classify_op(P, i, r, g, x))
                emit(P, g, r, x, 1);
        }
    }
    void classify_claimed(const NkRun &P, unsigned n)
    {
        for (unsigned i : order(n))
        {
            NkPend r;
            unsigned long long g;
            int x;
            if (nk_classify_claimed_op(P, i, r, g, x))
                emit(P, g, r, x, 0);
        }
    }
    void emit(const NkRun &P, unsigned long long g, const NkPend &r, int x, int kind)
    {
        unsigned si = P.ctr->n_slow++;
        if (si < P.slow_cap)
            nk_slow_write(P, si, g, r, x, kind);
        else
            P.ctr->overflow |= NK_OVF_SLOW;
    }
    void sort_pairs(unsigned long long *kin, unsigned long long *kout, unsigned long long *vin, unsigned long long *vout,
                    unsigned n)
    {
        std::vector<unsigned> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return kin[a] < kin[b]; });
        for (unsigned i = 0; i < n; i++)
        {
            kout[i] = kin[idx[i]];
            vout[i] = vin[idx[i]];
        }
    }
    void rank(const NkRun &P, const unsigned long long *keys, const unsigned long long *vals, unsigned n)
    {
        for (unsigned i : order(n))
            nk_rank_op(P, keys, vals, n, i);
    }
    void commit(const NkRun &P, unsigned n)
    {
        for (unsigned i : order(n))
            nk_commit_op(P, i);
    }
    void untag(const NkRun &P, unsigned n)
    {
        for (unsigned i : order(n))
            nk_untag_op(P, i);
    }
    void rehash(const NkSlot *old_tab, unsigned long long cap, NkSlot *nt, unsigned long long ncap, unsigned long long nmagic)
    {
        for (unsigned i : order((size_t)cap))
            nk_rehash_place_op(old_tab, i, nt, ncap, nmagic);
        for (unsigned long long j = 0; j < ncap; j++)
            nk_rehash_fill_op(old_tab, nt, j);
    }
    void decide(const NkRun &P, unsigned n_records, int paired, float coverage, unsigned char *accept)
    {
        for (unsigned r : order(n_records))
            nk_decide_op(P, r, paired, coverage, accept);
    }
    /* raw record text: the same per-record / per-entry functions as the kernels, entries visited in shuffled order */
    bool prepare_scan(size_t, std::string &) { return true; }
    static bool device_memory(int, uint64_t *free_bytes, uint64_t *total_bytes)
    { /* NK_EMU_HBM_MB lets the CPU tests drive the host pipeline's wave scheduling */
        const char *s = getenv("NK_EMU_HBM_MB");
        uint64_t b = (s && atoll(s) > 0 ? (uint64_t)atoll(s) : 1024ull * 180) << 20;
        if (free_bytes)
            *free_bytes = b;
        if (total_bytes)
            *total_bytes = b;
        return true;
    }
    void upload(void *d, const void *h, size_t n) { memcpy(d, h, n); }
    void upload_fence() {}
    void copy_fence() {}
    void copy_out(void *h, const void *d, size_t n, int) { memcpy(h, d, n); }
    void copy_wait(int) {}
    void raw_index(const NkRaw &R)
    {
        raw_records(R);
        raw_number(R);
    }
    void raw_limit(const NkRaw &R, unsigned limit)
    {
        for (unsigned i : order(R.n_records))
            nk_seed_flag_op(R, i);
        unsigned run = 0;
        for (unsigned i = 0; i < R.n_records; i++)
        {
            R.outoff[i] = run;
            run += R.outlen[i];
        }
        R.outoff[R.n_records] = run;
        for (unsigned i : order(R.n_records))
            nk_seed_clip_op(R, i, limit);
    }
    void raw_records(const NkRaw &R)
    {
        unsigned n = 0;
        for (unsigned p = 0; p < R.raw_bytes; p++)
        {
            if (R.raw[p] == 0)
                *R.flags |= NK_RAW_NUL;
            if (R.raw[p] == '\n')
            {
                if (n < R.nlpos_cap)
                    R.nlpos[n] = p;
                n++;
            }
        }
        R.flags[1] = n;
        for (unsigned i : order(R.n_records))
            nk_raw_record_op(R, i);
    }
    void raw_number(const NkRaw &R)
    {
        const unsigned n_reads = R.n_records * R.stride;
        unsigned run = 0;
        for (unsigned j = 0; j < n_reads; j++)
        {
            R.opscan[j] = run;
            run += R.nops[j];
        }
        R.opscan[n_reads] = run;
        for (unsigned j : order(n_reads))
            nk_raw_opbase_op(R, j);
    }
    void raw_emit(const NkRaw &R)
    {
        const unsigned n_out = R.n_records * R.stride;
        for (unsigned e : order(n_out))
        {
            unsigned wi, mate, rec;
            int counted, printed;
            R.outlen[e] = nk_emit_len_op(R, e, wi, mate, rec, counted, printed);
            if (!mate)
            {
                R.summary[6u * wi + 4u] += (unsigned)counted;
                R.summary[6u * wi + 5u] += (unsigned)printed;
            }
        }
        unsigned run = 0;
        for (unsigned e = 0; e < n_out; e++)
        {
            R.outoff[e] = run;
            run += R.outlen[e];
        }
        R.outoff[n_out] = run;
        for (unsigned e : order(n_out))
        {
            const unsigned o = R.outoff[e], len = R.outoff[e + 1] - o;
            if (!len)
                continue;
            unsigned wi, mate, rec;
            int counted, printed;
            nk_emit_len_op(R, e, wi, mate, rec, counted, printed);
            const NkRawWin w = R.wins[wi];
            const NkRawRec x = nk_raw_record(R, w, rec - w.rec0, (int)mate);
            for (unsigned b = 0; b < len; b++)
                R.out[o + b] = nk_emit_byte(R, x, mate, b, len);
        }
        for (unsigned w : order(R.n_wins))
            nk_emit_summary_op(R, w);
    }
    /* table dump and merged table: the same tile sizes / offsets / per-entry formatter as the kernels */
    bool dump_scan(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, int k, int text,
                   unsigned long long *tile)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE, run = 0;
        for (unsigned long long t = 0; t < tiles; t++)
        {
            unsigned long long u = 0;
            for (unsigned long long i = t * NK_DUMP_TILE; i < std::min<unsigned long long>(n, (t + 1) * NK_DUMP_TILE); i++)
            {
                unsigned long long key;
                long long val;
                nk_dump_entry(src, lo + i, key, val);
                u += text ? nk_dump_len(key, val, k) : (key != 0);
            }
            tile[t] = run;
            run += u;
        }
        tile[tiles] = run;
        return true;
    }
    void dump_text(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, int k, const unsigned long long *tile,
                   char *text)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        for (unsigned t : order((size_t)tiles))
        {
            unsigned long long at = tile[t];
            for (unsigned long long i = (unsigned long long)t * NK_DUMP_TILE;
                 i < std::min<unsigned long long>(n, ((unsigned long long)t + 1) * NK_DUMP_TILE); i++)
            {
                unsigned long long key;
                long long val;
                nk_dump_entry(src, lo + i, key, val);
                unsigned len = nk_dump_len(key, val, k);
                if (len)
                    nk_dump_format(key, val, k, text + at, len);
                at += len;
            }
        }
    }
    void dump_pairs(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, const unsigned long long *tile,
                    unsigned long long *keys_out, long long *vals_out)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        for (unsigned t : order((size_t)tiles))
        {
            unsigned long long at = tile[t];
            for (unsigned long long i = (unsigned long long)t * NK_DUMP_TILE;
                 i < std::min<unsigned long long>(n, ((unsigned long long)t + 1) * NK_DUMP_TILE); i++)
            {
                unsigned long long key;
                long long val;
                nk_dump_entry(src, lo + i, key, val);
                if (key)
                {
                    keys_out[at] = key;
                    vals_out[at++] = val;
                }
            }
        }
    }
    bool merge_pairs(unsigned long long *keys, long long *vals, unsigned long long n, int, unsigned long long *,
                     long long *, unsigned long long *n_out)
    {
        std::vector<unsigned> idx = order((size_t)n);
        std::sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return keys[a] < keys[b]; });
        std::vector<unsigned long long> k2;
        std::vector<long long> v2;
        for (unsigned i : idx)
        {
            if (!k2.empty() && k2.back() == keys[i])
                v2.back() += vals[i];
            else
            {
                k2.push_back(keys[i]);
                v2.push_back(vals[i]);
            }
        }
        std::copy(k2.begin(), k2.end(), keys);
        std::copy(v2.begin(), v2.end(), vals);
        *n_out = k2.size();
        return true;
    }
};

#define NK_BACKEND EmuBackend
#include "../../nomalise_kmers_multi_large_b200/csrc/nk_engine_api.h"

extern "C" void *nkd_alloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 16); }
extern "C" void nkd_free_pinned(void *p) { free(p); }
/* NK_EMU_DEVICES=n lets the CPU tests drive the host pipeline's multi-GPU placement (one emulated engine per "GPU") */
extern "C" int nkd_device_count(void)
{
    const char *s = getenv("NK_EMU_DEVICES");
    return s && atoi(s) > 0 ? atoi(s) : 1;
}
