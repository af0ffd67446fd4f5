/*
 * nk_orchestrate.h -- step orchestration of one device engine, templated on a backend.
 *
 * The product instantiates NkEngine<CudaBackend> (nk_engine.cu).  tests/emu instantiates it with
 * a CPU backend that executes the per-operation functions of nk_core.h in shuffled order, to check
 * the parallel algorithm against the oracle without a GPU.  No CPU backend is compiled into the
 * product library.
 *
 * A backend provides: alloc/release/zero/h2d/d2h/d2d/sync, the launches probe/open_ops/apply/
 * classify/sort_pairs/rank/commit/untag/rehash/decide, and begin_timer/end_timer.
 */
#ifndef NK_ORCHESTRATE_H
#define NK_ORCHESTRATE_H

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "nk_core.h"
#include "../../include/nk_b200.h"

struct NkTable
{
    NkSlot *tab = nullptr;
    uint64_t cap = 0, used = 0, thr = 0, magic = 0;
    nkd_part_stats st{};
    /* table residency (engines with a table budget): a partition's table is in HBM (tab), parked in host memory
     * (host), or still identical to the seed table (fresh) */
    void *host = nullptr;
    bool fresh = false;
    uint64_t last_use = 0;
};

/* smallest used for which the reference's `used >= capacity * 0.8` (double) holds, C:933 */
static inline uint64_t nk_expand_threshold(uint64_t cap)
{
    double lim = (double)cap * 0.8;
    uint64_t u = (uint64_t)lim;
    while ((double)u < lim)
        u++;
    while (u > 0 && (double)(u - 1) >= lim)
        u--;
    return u;
}
static inline uint64_t nk_grown_capacity(uint64_t cap) { return (uint64_t)((double)cap + (double)cap * 0.5); } /* C:1058 */
static inline uint64_t nk_magic(uint64_t cap) { return ~0ull / cap; }

static inline double nk_env_double(const char *name, double dflt)
{
    const char *s = getenv(name);
    return s && *s ? atof(s) : dflt;
}

template <class B>
class NkEngine
{
  public:
    B be;
    nkd_config cfg{};
    std::string err;

    NkTable seed;
    std::vector<NkTable> parts;
    bool seeded = false;
    /* Partitions are independent (C:1841-1880): when their tables do not fit the GPU together, the host works on them in
     * waves and the tables of the others wait in host memory.  0 = no budget, every table stays in HBM. */
    uint64_t table_budget = 0, use_clock = 0, evictions = 0, loads = 0;
    unsigned fresh_left = 0;

    /* step buffers */
    const unsigned char *seq_view = nullptr; /* where the kernels read the step's sequence bytes from */
    unsigned char *d_seq = nullptr;
    NkRead *d_reads = nullptr;
    unsigned *d_high = nullptr, *d_total = nullptr;
    unsigned char *d_invalid = nullptr, *d_accept = nullptr;
    NkOpen *d_open = nullptr;
    NkPend *d_pend = nullptr, *d_spec = nullptr;
    NkClaim *d_claim = nullptr;
    unsigned long long *d_skey[2] = {nullptr, nullptr}, *d_sval[2] = {nullptr, nullptr};
    unsigned long long *d_keys_out = nullptr;
    NkCounters *d_ctr = nullptr;
    NkPart *d_parts = nullptr;
    unsigned *d_bloom = nullptr;
    unsigned bloom_words = 0;
    NkHot *d_hot = nullptr; /* hot table (nk_core.h): saturated home hits stay in L2 */
    unsigned hot_entries = 0;
    bool hot_dirty = false; /* some scoring has run since the sums were last folded into the tables */
    unsigned open_cap = 0, pend_cap = 0, claim_cap = 0, spec_cap = 0, slow_cap = 0;

    NkCounters h_ctr{};
    std::vector<NkPart> h_parts;
    std::vector<NkClaim> h_claims;

    /* staged step */
    size_t n_reads = 0, n_records = 0;
    int paired = 0;
    std::vector<unsigned> T;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    uint64_t upload_total = 0; /* bytes sent ahead by nkd_upload_raw (kept apart: another thread adds to it) */
    nkd_run_stats rs{};
    std::vector<float> spans; /* (start, end) ms of every scoring step on the device clock */
    bool staged = false, ran = false;
    float last_total_ms = 0, last_probe_ms = 0;

    /* merged table (all partitions, sorted by k-mer, counts summed) */
    unsigned long long *d_mkeys[2] = {nullptr, nullptr};
    long long *d_mvals[2] = {nullptr, nullptr};
    uint64_t merge_cap = 0, merge_n = 0, merged_n = 0;
    bool merged_ready = false;

    /* raw record text path (nkd_stage_raw / nkd_fetch_raw) */
    unsigned char *d_raw = nullptr, *d_out = nullptr;
    unsigned char *d_raw_next = nullptr; /* second text buffer: the next step's bytes arrive while this one runs */
    const uint8_t *uploaded = nullptr;   /* host buffer whose bytes d_raw_next holds (or is receiving) */
    size_t uploaded_bytes = 0;
    std::atomic<bool> raw_ready{false};  /* raw_prepare has run */
    std::mutex up_mu; /* nkd_upload_raw may come from another host thread than the one that runs the steps */
    unsigned *d_tile = nullptr, *d_nlpos = nullptr, *d_nops = nullptr, *d_opscan = nullptr, *d_tout = nullptr;
    unsigned *d_rflags = nullptr, *d_outlen = nullptr, *d_outoff = nullptr;
    unsigned long long *d_summary = nullptr;
    NkRawWin *d_wins = nullptr;
    std::vector<NkRawWin> h_wins;
    uint64_t raw_cap = 0, raw_reads_cap = 0, raw_lines_cap = 0;
    NkRaw raw{}; /* the staged raw step */
    bool raw_check = false; /* the staged raw step's parse flags have not been looked at yet */
    uint64_t raw_lines_expected = 0;
    unsigned h_rflags[4] = {0, 0, 0, 0};
    bool raw_staged = false;

    bool debug = getenv("NKB200_DEBUG") && *getenv("NKB200_DEBUG") && strcmp(getenv("NKB200_DEBUG"), "0") != 0;

    int fail(int code, const std::string &m)
    {
        err = m;
        return code;
    }

    int create(const nkd_config &c)
    {
        cfg = c;
        if (c.k < 5 || c.k > 31 || c.n_parts < 1 || c.n_parts > NK_MAX_PARTITIONS || c.depth_per_part < 2 ||
            c.capacity0 < 1 || c.capacity0 >= 0xFFFFFFFFull)
            return fail(NK_EINVAL, "nkd_create: bad configuration");
        if (c.max_step_ops >= (1ull << NK_T_BITS) || c.max_step_reads >= (1ull << 31))
            return fail(NK_EINVAL, "nkd_create: step limits too large (ops per step must stay below 2^28)");
        int rc = be.init(c.device, err);
        if (rc)
            return rc;
        if (!alloc_table(seed, c.capacity0))
            return fail(NK_ENOMEM, "nkd_create: cannot allocate the seed table");
        uint64_t ops = std::max<uint64_t>(c.max_step_ops, 1024);
        open_cap = (unsigned)std::min<double>(4e9, ops * nk_env_double("NKB200_OPEN_FRAC", 1.0) + 1024);
        pend_cap = (unsigned)std::min<double>(4e9, ops * nk_env_double("NKB200_PEND_FRAC", 2.0) + 1024);
        /* small steps: every warp of a launch may hold two reserved chunks per list, so a list shorter than
         * a few times that is all holes and the step would only shrink its windows and retry */
        if (!getenv("NKB200_OPEN_FRAC") && !getenv("NKB200_PEND_FRAC"))
        {
            open_cap = std::max(open_cap, be.min_list_entries());
            pend_cap = std::max(pend_cap, be.min_list_entries());
        }
        claim_cap = open_cap;
        spec_cap = open_cap;
        /* every listed event can turn into one slow record; chunked reservation wastes at most a chunk per warp */
        slow_cap = (unsigned)std::min<double>(4e9, 1.25 * ((double)pend_cap + (double)spec_cap) + (1 << 20) + be.slow_hole_margin());
        bool ok = true;
        ok &= dalloc(d_seq, c.max_step_bytes + 64);
        ok &= dalloc(d_reads, c.max_step_reads + 1);
        ok &= dalloc(d_high, c.max_step_reads + 1);
        ok &= dalloc(d_total, c.max_step_reads + 1);
        ok &= dalloc(d_invalid, c.max_step_reads + 1);
        ok &= dalloc(d_accept, c.max_step_reads + 1);
        ok &= dalloc(d_open, open_cap);
        ok &= dalloc(d_pend, pend_cap);
        ok &= dalloc(d_spec, spec_cap);
        ok &= dalloc(d_claim, claim_cap);
        for (int i = 0; i < 2; i++)
        {
            ok &= dalloc(d_skey[i], slow_cap);
            ok &= dalloc(d_sval[i], slow_cap);
        }
        ok &= dalloc(d_ctr, 1);
        ok &= dalloc(d_parts, NK_MAX_PARTITIONS);
        /* filter: ~8 bits per operation of a step, 64 MB at most so that it stays in the 126 MB L2 */
        bloom_words = 1u << 15;
        while (bloom_words < (1u << 24) && (uint64_t)bloom_words * 32u < ops * 8u)
            bloom_words <<= 1;
        ok &= dalloc(d_bloom, bloom_words);
        /* hot table (nk_core.h): OFF by default.  Measured on the benchmark (profiles/r02_hot_table.txt): with 2^20 / 2^24
         * entries per engine it absorbs 31 % / 49 % of all operations, and k_probe_score does not get faster (202 -> 210
         * ms): the lines of hot saturated counters were L2 hits already, the DRAM traffic comes from the cold ones.
         * NKB200_HOT_ENTRIES=n (rounded down to a power of two) turns it on; the tests keep it covered. */
        {
            const char *e = getenv("NKB200_HOT_ENTRIES");
            long long want = e && *e ? atoll(e) : 0;
            hot_entries = 0;
            if (want > 0)
            {
                hot_entries = 1;
                while ((long long)hot_entries * 2 <= want && hot_entries < (1u << 26))
                    hot_entries *= 2;
                ok &= dalloc(d_hot, hot_entries);
                if (d_hot)
                    be.zero(d_hot, (size_t)hot_entries * sizeof(NkHot));
            }
        }
        if (!ok)
            return fail(NK_ENOMEM, "nkd_create: cannot allocate step scratch");
        if (!be.prepare_sort(slow_cap, err))
            return NK_ENOMEM;
        h_parts.resize(NK_MAX_PARTITIONS);
        T.assign(NK_MAX_PARTITIONS, 0);
        return NK_OK;
    }

    void destroy()
    {
        be.sync();
        if (getenv("NKB200_TIMES") && *getenv("NKB200_TIMES") && strcmp(getenv("NKB200_TIMES"), "0") != 0)
            fprintf(stderr, "[nkd] span gaps ms: parse->probe %.1f, probe->open %.1f, open->classify %.1f, classify->sort %.1f, "
                            "decide->emit %.1f, sort->decide %.1f; run %.1f ms\n",
                    gaps[0], gaps[1], gaps[2], gaps[3], gaps[4], gaps[5], rs.run_ms);
        be.release(seed.tab);
        for (auto &p : parts)
        {
            be.release(p.tab);
            free(p.host);
            p.host = nullptr;
        }
        be.release(d_seq);
        be.release(d_reads);
        be.release(d_high);
        be.release(d_total);
        be.release(d_invalid);
        be.release(d_accept);
        be.release(d_open);
        be.release(d_pend);
        be.release(d_spec);
        be.release(d_claim);
        for (int i = 0; i < 2; i++)
        {
            be.release(d_skey[i]);
            be.release(d_sval[i]);
        }
        be.release(d_keys_out);
        be.release(d_ctr);
        be.release(d_parts);
        be.release(d_bloom);
        be.release(d_hot);
        be.release(d_raw);
        be.release(d_raw_next);
        be.release(d_out);
        be.release(d_tile);
        be.release(d_nlpos);
        be.release(d_nops);
        be.release(d_opscan);
        be.release(d_tout);
        be.release(d_rflags);
        be.release(d_outlen);
        be.release(d_outoff);
        be.release(d_summary);
        be.release(d_wins);
        merge_release();
        be.shutdown();
    }

    template <class X>
    bool dalloc(X *&p, size_t n)
    {
        p = (X *)be.alloc(n * sizeof(X));
        return p != nullptr;
    }

    bool alloc_table(NkTable &t, uint64_t cap)
    {
        t.tab = (NkSlot *)be.alloc(cap * sizeof(NkSlot));
        if (!t.tab)
            return false;
        be.zero(t.tab, cap * sizeof(NkSlot));
        t.cap = cap;
        t.used = 0;
        t.thr = nk_expand_threshold(cap);
        t.magic = nk_magic(cap);
        t.st.capacity = cap;
        return true;
    }

    /* ------------------------------------------------------------ table residency */

    /* fold the hot table's pending sums into the tables: before anything reads exact counts or moves a table */
    void flush_hot()
    {
        if (!d_hot || !hot_dirty || parts.empty())
            return;
        std::vector<NkTable *> tabs;
        for (auto &p : parts)
            tabs.push_back(&p);
        std::vector<unsigned> none(tabs.size(), 0);
        upload_parts(tabs, none, none);
        be.hot_flush(make_run(NK_MODE_SCORE, +1, 0));
        hot_dirty = false;
    }

    int evict_part(int p)
    {
        NkTable &t = parts[p];
        if (!t.tab)
            return NK_OK;
        flush_hot();
        size_t bytes = (size_t)t.cap * sizeof(NkSlot);
        t.host = malloc(bytes);
        if (!t.host)
            return fail(NK_ENOMEM, "cannot park a partition table in host memory");
        be.d2h(t.host, t.tab, bytes);
        be.sync();
        be.release(t.tab);
        t.tab = nullptr;
        d2h_bytes += bytes;
        evictions++;
        if (debug)
            fprintf(stderr, "[nkd] partition %d parked in host memory (%zu MB)\n", p, bytes >> 20);
        return NK_OK;
    }

    int load_part(int p)
    {
        NkTable &t = parts[p];
        if (t.tab)
            return NK_OK;
        size_t bytes = (size_t)t.cap * sizeof(NkSlot);
        t.tab = (NkSlot *)be.alloc(bytes);
        if (!t.tab)
            return fail(NK_ENOMEM, "Memory allocation failed (partition table)");
        if (t.host)
        {
            be.h2d(t.tab, t.host, bytes);
            be.sync();
            free(t.host);
            t.host = nullptr;
            h2d_bytes += bytes;
        }
        else if (t.fresh)
        { /* copy_hash_table, C:908-927: first use of this partition */
            be.d2d(t.tab, seed.tab, bytes);
            t.fresh = false;
            if (--fresh_left == 0)
            {
                be.sync();
                be.release(seed.tab);
                seed.tab = nullptr;
            }
        }
        else
            return fail(NK_EINTERNAL, "partition table lost");
        loads++;
        return NK_OK;
    }

    /* the partitions in `need` get their tables into HBM; least recently used others make room */
    int make_resident(const std::vector<int> &need)
    {
        if (!table_budget)
            return NK_OK;
        std::vector<char> wanted(parts.size(), 0);
        uint64_t resident = seed.tab ? seed.cap * sizeof(NkSlot) : 0, incoming = 0, largest = 0;
        for (int p : need)
            wanted[p] = 1;
        for (size_t p = 0; p < parts.size(); p++)
        {
            uint64_t b = parts[p].cap * sizeof(NkSlot);
            if (parts[p].tab)
                resident += b;
            else if (wanted[p])
                incoming += b;
            if (wanted[p] && b > largest)
                largest = b;
        }
        /* room for one table growing while the others stay (old and new table coexist during a re-hash, C:1070) */
        uint64_t headroom = largest + largest / 2;
        while (resident + incoming + headroom > table_budget)
        {
            int victim = -1;
            for (size_t p = 0; p < parts.size(); p++)
                if (parts[p].tab && !wanted[p] && (victim < 0 || parts[p].last_use < parts[victim].last_use))
                    victim = (int)p;
            if (victim < 0)
                break; /* everything resident is needed: the allocation below decides */
            resident -= parts[victim].cap * sizeof(NkSlot);
            int rc = evict_part(victim);
            if (rc)
                return rc;
        }
        use_clock++;
        for (int p : need)
        {
            int rc = load_part(p);
            if (rc)
                return rc;
            parts[p].last_use = use_clock;
        }
        return NK_OK;
    }

    /* expand_local_hash_table, C:1055-1108 */
    int expand(NkTable &t)
    {
        uint64_t ncap = nk_grown_capacity(t.cap);
        if (ncap <= t.cap)
            return NK_OK;
        if (ncap >= 0xFFFFFFFFull)
            return fail(NK_ENOMEM, "table growth beyond 2^32 slots is not supported on one partition");
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        if (debug)
            be.sync();
        const bool is_part = !parts.empty() && &t >= parts.data() && &t < parts.data() + parts.size();
        if (is_part)
            flush_hot(); /* the re-hash carries the counts over (C:1089) and moves the slots */
        double t0 = now();
        NkSlot *nt = (NkSlot *)be.alloc(ncap * sizeof(NkSlot));
        if (!nt)
        {
            char m[160];
            snprintf(m, sizeof m, "Error: Memory allocation failed to expand local hash table, from %llu to %llu",
                     (unsigned long long)t.cap, (unsigned long long)ncap);
            return fail(NK_ENOMEM, m);
        }
        double t1 = now();
        be.zero(nt, ncap * sizeof(NkSlot));
        be.rehash(t.tab, t.cap, nt, ncap, nk_magic(ncap));
        if (is_part && d_hot)
            be.hot_clear(make_run(NK_MODE_SCORE, +1, 0), (unsigned)(&t - parts.data()) + 1u);
        be.sync();
        double t2 = now();
        be.release(t.tab);
        if (debug)
            fprintf(stderr, "[nkd] expand %llu -> %llu slots: alloc %.2f ms, zero+rehash %.2f ms, free %.2f ms\n",
                    (unsigned long long)t.cap, (unsigned long long)ncap, t1 - t0, t2 - t1, now() - t2);
        t.tab = nt;
        t.cap = ncap;
        t.thr = nk_expand_threshold(ncap);
        t.magic = nk_magic(ncap);
        t.st.capacity = ncap;
        t.st.expansions++;
        return NK_OK;
    }

    NkRun make_run(int mode, int delta, int record)
    {
        NkRun P{};
        P.seq = seq_view ? seq_view : d_seq;
        P.reads = d_reads;
        P.n_reads = (unsigned)n_reads;
        P.parts = d_parts;
        P.k = cfg.k;
        P.canonical = cfg.canonical;
        P.depth = cfg.depth_per_part;
        P.mode = mode;
        P.delta = delta;
        P.record = record;
        P.high = d_high;
        P.total = d_total;
        P.invalid = d_invalid;
        P.open = d_open;
        P.open_cap = open_cap;
        P.pend = d_pend;
        P.pend_cap = pend_cap;
        P.spec = d_spec;
        P.spec_cap = spec_cap;
        P.claim = d_claim;
        P.claim_cap = claim_cap;
        P.slow_key = d_skey[0];
        P.slow_val = d_sval[0];
        P.slow_cap = slow_cap;
        P.ctr = d_ctr;
        P.keys_out = d_keys_out;
        P.bloom = d_bloom;
        P.bloom_words = bloom_words;
        P.hot = d_hot;
        P.hot_mask = hot_entries ? hot_entries - 1 : 0;
        be.chunk_sizes(P.chunk, pend_cap, open_cap, claim_cap, slow_cap, spec_cap);
        return P;
    }

    void upload_parts(std::vector<NkTable *> &tabs, const std::vector<unsigned> &lo, const std::vector<unsigned> &hi)
    {
        uint64_t g = 0;
        for (size_t p = 0; p < tabs.size(); p++)
        {
            NkPart &d = h_parts[p];
            d.tab = tabs[p]->tab;
            d.cap = tabs[p]->cap;
            d.magic = tabs[p]->magic;
            d.gbase = g;
            d.lo = lo[p];
            d.hi = hi[p];
            if (tabs[p]->tab) /* tables parked in host memory take no part in this step */
                g += tabs[p]->cap;
        }
        /* through kernel parameters, not the copy engine: a step's small control blocks must not queue behind the next
         * step's text, which the upload stream is sending while this step runs (that cost 2.5 ms per step) */
        be.put_small(d_parts, h_parts.data(), tabs.size() * sizeof(NkPart));
    }

    void fetch_counters()
    {
        be.d2h(&h_ctr, d_ctr, sizeof(NkCounters));
        if (raw_check)
            be.d2h(h_rflags, d_rflags, sizeof h_rflags);
        be.sync();
    }

    /* 0, or why the staged raw step cannot be scored as it stands (looked at once per step) */
    int raw_verdict()
    {
        if (!raw_check)
            return NK_OK;
        raw_check = false;
        if (h_rflags[1] == raw_lines_expected && !(h_rflags[0] & (NK_RAW_NUL | NK_RAW_LONG | NK_RAW_SHAPE)))
            return NK_OK;
        if (h_rflags[1] == raw_lines_expected && (h_rflags[0] & (NK_RAW_NUL | NK_RAW_LONG)))
            return fail(NK_EIRREGULAR, "raw text needs the host parser (NUL byte or a line of 1024+ chars)");
        return fail(NK_EINVAL, "nkd_stage_raw: a window does not hold the announced number of complete records");
    }

    /* The sequential semantics of all operations [0,T[p]) of every table in tabs.
     * mode: NK_MODE_SCORE or NK_MODE_SEED. */
    int run_ops(int mode, std::vector<NkTable *> &tabs)
    {
        size_t np = tabs.size();
        std::vector<unsigned> lo(np, 0), hi(np);
        /* after an overflow the windows regrow by doubling instead of jumping back to the whole remainder, which
         * would overflow again and repeat the halving cascade for every few operations of progress */
        std::vector<unsigned> regrow(np, 0); /* 0 = no limit; else the size of the last window that was tried */
        for (size_t p = 0; p < np; p++)
            hi[p] = T[p];
        uint64_t gsum = 0;
        for (auto *t : tabs)
            gsum += t->tab ? t->cap : 0;
        for (;;)
        {
            bool live = false;
            for (size_t p = 0; p < np; p++)
                live |= lo[p] < T[p];
            if (!live)
                break;
            /* the load-factor test precedes every operation (C:933): a table at its threshold grows
             * as soon as one more non-ignored window arrives */
            bool need_count = false;
            for (size_t p = 0; p < np; p++)
                need_count |= (lo[p] < hi[p] && tabs[p]->used >= tabs[p]->thr);
            if (need_count)
            {
                upload_parts(tabs, lo, hi);
                be.zero(d_ctr, sizeof(NkCounters));
                be.probe(make_run(NK_MODE_COUNT, 0, 0));
                fetch_counters();
                if (int bad = raw_verdict())
                    return bad; /* nothing has been changed yet */
                for (size_t p = 0; p < np; p++)
                    if (lo[p] < hi[p] && tabs[p]->used >= tabs[p]->thr && h_ctr.real_ops[p] > 0)
                    {
                        int rc = expand(*tabs[p]);
                        if (rc)
                            return rc;
                    }
                gsum = 0;
                for (auto *t : tabs)
                    gsum += t->tab ? t->cap : 0;
            }
            if (gsum >= (1ull << 34))
                return fail(NK_ENOMEM, "tables of one device exceed 2^34 slots");

            /* forward run: probe (claim-independent part), then the deferred operations */
            upload_parts(tabs, lo, hi);
            be.zero(d_ctr, sizeof(NkCounters));
            if (mode == NK_MODE_SCORE)
                be.zero(d_bloom, (size_t)bloom_words * sizeof(unsigned));
            NkRun F = make_run(mode, +1, 1);
            if (debug)
            {
                be.sync();
                fprintf(stderr, "[nkd] mode %d launching probe: reads %u, p0 [%u,%u)\n", mode, F.n_reads, lo[0], hi[0]);
            }
            be.begin_timer(1);
            hot_dirty = hot_dirty || mode == NK_MODE_SCORE; /* from here on the hot table may hold sums again */
            be.probe(F);
            be.end_timer(1);
            if (debug)
            {
                be.d2h(&h_ctr, d_ctr, 32);
                be.sync();
                fprintf(stderr, "[nkd] probe done: n_open %u n_pend %u ovf %x\n", h_ctr.n_open, h_ctr.n_pend, h_ctr.overflow);
            }
            be.begin_timer(2);
            if (mode == NK_MODE_SCORE)
                be.prepare_claims(F);
            be.open_ops(F);
            be.end_timer(2);
            if (debug)
            {
                be.d2h(&h_ctr, d_ctr, 32);
                be.sync();
                fprintf(stderr, "[nkd] open done (n_open %u)\n", h_ctr.n_open);
            }
            fetch_counters();
            if (int bad = raw_verdict())
            { /* forget this run (the same replay as for an overflow) and hand the text back */
                NkRun U = make_run(mode, -1, 0);
                be.probe(U);
                be.open_ops(U);
                be.untag(U, std::min(h_ctr.n_open, open_cap));
                be.sync();
                return bad;
            }
            if (debug)
            {
                fprintf(stderr, "[nkd] mode %d fwd: open %u pend %u spec %u claim %u ovf %x |", mode, h_ctr.n_open,
                        h_ctr.n_pend, h_ctr.n_spec, h_ctr.n_claim, h_ctr.overflow);
                for (size_t p = 0; p < np && p < 4; p++)
                    fprintf(stderr, " p%zu [%u,%u) of %u used %llu/%llu cap %llu claims %u real %llu", p, lo[p], hi[p], T[p],
                            (unsigned long long)tabs[p]->used, (unsigned long long)tabs[p]->thr,
                            (unsigned long long)tabs[p]->cap, h_ctr.claims[p], (unsigned long long)h_ctr.real_ops[p]);
                fprintf(stderr, "\n");
            }
            /* A walk that hits the watchdog usually means the speculative pass claimed more slots than the table
             * has free ones (a step full of new k-mers): the growth the reference would have done in the middle of
             * the step has not happened yet.  Treated like a list overflow: undo, halve the windows, and let the
             * threshold cut of a shorter window grow the table in time. */
            bool walk = (h_ctr.overflow & NK_OVF_WALK) != 0;
            bool ovf = walk || (h_ctr.overflow & (NK_OVF_OPEN | NK_OVF_PEND | NK_OVF_CLAIM)) != 0;
            bool cut = false;
            std::vector<unsigned> nhi(hi);
            if (ovf)
            {
                for (size_t p = 0; p < np; p++)
                    if (hi[p] > lo[p])
                    {
                        unsigned w = hi[p] - lo[p];
                        if (w > 1)
                        {
                            nhi[p] = lo[p] + w / 2;
                            regrow[p] = w / 2;
                            cut = true;
                        }
                    }
                if (!cut)
                    return walk ? fail(NK_EINTERNAL, "probe walk exceeded the supported length (table degenerate)")
                                : fail(NK_ENOMEM, "step scratch too small for a single operation");
            }
            else
            {
                bool fetched = false;
                for (size_t p = 0; p < np; p++)
                {
                    NkTable &t = *tabs[p];
                    if (t.used >= t.thr || t.used + h_ctr.claims[p] < t.thr)
                        continue;
                    /* the claim that brings `used` to the threshold; the next non-ignored window grows the table */
                    if (!fetched)
                    {
                        unsigned nc = std::min(h_ctr.n_claim, claim_cap);
                        h_claims.resize(nc);
                        if (nc)
                            be.d2h(h_claims.data(), d_claim, (size_t)nc * sizeof(NkClaim));
                        be.sync();
                        fetched = true;
                    }
                    std::vector<unsigned> times;
                    for (auto &c : h_claims)
                        if (c.slot != NK_HOLE && c.part == p)
                            times.push_back(c.t);
                    uint64_t m = t.thr - t.used;
                    std::nth_element(times.begin(), times.begin() + (m - 1), times.end());
                    unsigned tstar = times[m - 1] + 1;
                    if (tstar < hi[p])
                    {
                        nhi[p] = tstar;
                        cut = true;
                    }
                }
            }
            if (debug && (cut || ovf))
                fprintf(stderr, "[nkd] cut=%d ovf=%d -> undo and retry with shorter windows\n", (int)cut, (int)ovf);
            /* Partitions whose window must be shortened are abandoned for this round: their part of the run is
             * replayed with -1 (the decisions are stable) and their claim tags are forgotten.  The others commit
             * from this very run: every list consumer skips entries of partitions whose window is empty. */
            std::vector<unsigned> clo(lo), chi(hi); /* windows the commit works on */
            if (cut)
            {
                std::vector<unsigned> ulo(np), uhi(np);
                for (size_t p = 0; p < np; p++)
                {
                    bool dead = ovf || nhi[p] != hi[p];
                    ulo[p] = dead ? lo[p] : hi[p];
                    uhi[p] = hi[p];
                    if (dead)
                        chi[p] = clo[p]; /* empty: nothing of it is committed */
                }
                upload_parts(tabs, ulo, uhi);
                NkRun U = make_run(mode, -1, 0);
                be.begin_timer(8);
                be.probe(U);
                be.open_ops(U);
                be.untag(U, std::min(h_ctr.n_open, open_cap));
                be.end_timer(8);
                if (ovf)
                { /* nothing can be trusted when a list overflowed: retry everything with halved windows */
                    hi = nhi;
                    continue;
                }
                upload_parts(tabs, clo, chi);
            }
            /* commit */
            if (mode == NK_MODE_SCORE)
            {
                unsigned np_ = std::min(h_ctr.n_pend, pend_cap), ns_ = std::min(h_ctr.n_spec, spec_cap);
                if (ns_)
                {
                    be.begin_timer(3);
                    be.apply(F, ns_);
                    be.end_timer(3);
                }
                if (np_ || ns_)
                {
                    be.begin_timer(4);
                    if (np_)
                        be.classify(F, np_);
                    if (ns_)
                        be.classify_claimed(F, ns_);
                    be.end_timer(4);
                    be.d2h(&h_ctr, d_ctr, 32);
                    be.sync();
                    if ((h_ctr.overflow & NK_OVF_SLOW) || h_ctr.n_slow > slow_cap)
                        return fail(NK_EINTERNAL, "slow-path list overflow (scratch sizing bug)");
                    if (h_ctr.n_slow)
                    {
                        be.begin_timer(5);
                        be.sort_pairs(d_skey[0], d_skey[1], d_sval[0], d_sval[1], h_ctr.n_slow);
                        be.rank(F, d_skey[1], d_sval[1], h_ctr.n_slow);
                        be.end_timer(5);
                    }
                }
            }
            if (h_ctr.n_claim)
            {
                be.begin_timer(6);
                be.commit(F, std::min(h_ctr.n_claim, claim_cap));
                be.end_timer(6);
            }
            for (size_t p = 0; p < np; p++)
            {
                NkTable &t = *tabs[p];
                if (chi[p] == clo[p] && hi[p] != lo[p])
                { /* abandoned this round: run its shortened window next */
                    hi[p] = nhi[p];
                    continue;
                }
                t.used += h_ctr.claims[p];
                t.st.used = t.used;
                t.st.ops += h_ctr.real_ops[p];
                t.st.touches += h_ctr.touches[p];
                lo[p] = hi[p];
                hi[p] = T[p];
                if (regrow[p])
                {
                    regrow[p] = regrow[p] > (1u << 30) ? regrow[p] : regrow[p] * 2;
                    if (T[p] - lo[p] > regrow[p])
                        hi[p] = lo[p] + regrow[p];
                    else
                        regrow[p] = 0; /* the remainder fits the regrown window: back to normal */
                }
            }
            if (mode == NK_MODE_SCORE && np > 0)
            {
                tabs[0]->st.slow_events += h_ctr.n_slow; /* device-wide figure, kept on partition 0 */
                rs.probe_touches += h_ctr.probe_touches;
                rs.hot_hits += h_ctr.hot_hits;
                rs.probe_launches++;
                rs.pend_events += h_ctr.n_pend;
                rs.open_ops += h_ctr.n_open;
                rs.slow_events += h_ctr.n_slow;
            }
        }
        return NK_OK;
    }

    std::vector<unsigned short> rec_part; /* partition of every record of the staged step */

    int stage(const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t nr, int is_paired, int n_tabs,
              bool ignore_part)
    {
        nkd_segment s{reads, nr, 0, seq_bytes, 0, 0, 0};
        return stage_segments(seq, &s, 1, is_paired, n_tabs, ignore_part);
    }

    int stage_segments(const uint8_t *seq, const nkd_segment *segs, int n_segs, int is_paired, int n_tabs,
                       bool ignore_part)
    {
        size_t nr = 0;
        for (int s = 0; s < n_segs; s++)
            nr += segs[s].n_reads;
        if (nr > cfg.max_step_reads)
            return fail(NK_EINVAL, "step exceeds the read limit given to nkd_create");
        std::fill(T.begin(), T.end(), 0u);
        int stride = is_paired ? 2 : 1;
        rec_part.resize(nr / stride + 1);
        size_t at = 0;
        const unsigned char *view = (const unsigned char *)be.device_view_of_host(seq);
        seq_view = view;
        for (int s = 0; s < n_segs; s++)
        {
            const nkd_segment &g = segs[s];
            if (g.seq_hi > cfg.max_step_bytes || g.seq_lo > g.seq_hi || (g.seq_lo & 15u))
                return fail(NK_EINVAL, "step segment exceeds the byte limit given to nkd_create");
            if (is_paired && (g.n_reads & 1))
                return fail(NK_EINVAL, "paired step with an odd number of reads");
            if (g.trusted && !ignore_part)
            {
                if ((int)g.part >= n_tabs)
                    return fail(NK_EINVAL, "segment names a partition that is not resident");
                if (g.ops > T[g.part])
                    T[g.part] = g.ops;
                for (size_t i = (at + stride - 1) / stride; i < (at + g.n_reads + stride - 1) / stride; i++)
                    rec_part[i] = (unsigned short)g.part;
            }
            else
            for (size_t i = 0; i < g.n_reads; i++)
            {
                const nkd_read &rd = g.reads[i];
                unsigned p = ignore_part ? 0 : rd.part;
                if ((int)p >= n_tabs)
                    return fail(NK_EINVAL, "read names a partition that is not resident");
                if (rd.len >= NK_MAX_LINE)
                    return fail(NK_EINVAL, "read longer than 1023 bases (the reference cuts lines there, C:397)");
                if ((int)rd.len < cfg.k)
                    return fail(NK_EINVAL, "read shorter than k in a step (the caller drops those, C:1430-1443)");
                if ((rd.seq_off & 15u) || rd.seq_off < g.seq_lo || (size_t)rd.seq_off + rd.len > g.seq_hi)
                    return fail(NK_EINVAL, "read offset not 16-byte aligned or outside its segment");
                unsigned end = rd.op_base + (unsigned)(rd.len - cfg.k + 1);
                if (end > T[p])
                    T[p] = end;
                if ((at + i) % stride == 0)
                    rec_part[(at + i) / stride] = (unsigned short)p;
            }
            if (g.n_reads)
            {
                /* the last 16-byte chunk of a read may extend past seq_hi: the buffers are padded */
                size_t hi = std::min<size_t>((g.seq_hi + 15) & ~(size_t)15, cfg.max_step_bytes);
                if (!view)
                    be.h2d(d_seq + g.seq_lo, seq + g.seq_lo, hi - g.seq_lo);
                be.h2d(d_reads + at, g.reads, g.n_reads * sizeof(nkd_read));
                h2d_bytes += (hi - g.seq_lo) + g.n_reads * sizeof(nkd_read);
            }
            at += g.n_reads;
        }
        uint64_t tot = 0;
        std::vector<int> need;
        for (int p = 0; p < n_tabs; p++)
        {
            if (T[p] >= (1u << NK_T_BITS))
                return fail(NK_EINVAL, "a partition has 2^28 or more operations in one step");
            tot += T[p];
            if (T[p])
                need.push_back(p);
        }
        if (tot > cfg.max_step_ops)
            return fail(NK_EINVAL, "step has more operations than max_step_ops");
        if (seeded && !ignore_part)
        {
            int rc = make_resident(need);
            if (rc)
                return rc;
        }
        n_reads = nr;
        paired = is_paired;
        n_records = nr / stride;
        staged = true;
        raw_staged = false;
        raw_check = false;
        ran = false;
        return NK_OK;
    }

    /* ------------------------------------------------------------ steps handed over as raw record text */

    /* scratch of the raw path, allocated on first use: a context that only ever stages parsed reads does not pay */
    int raw_prepare()
    {
        if (raw_ready.load(std::memory_order_acquire)) /* nkd_upload_raw asks from another thread than the steps' */
            return NK_OK;
        if (d_raw)
            return fail(NK_ENOMEM, "nkd_stage_raw: cannot allocate the raw-text scratch");
        if (!cfg.max_raw_bytes)
            return fail(NK_EINVAL, "nkd_stage_raw: the engine was created with max_raw_bytes = 0");
        raw_cap = (cfg.max_raw_bytes + 15) & ~15ull;
        raw_reads_cap = cfg.max_step_reads + 2;
        raw_lines_cap = raw_reads_cap * 4 + 16;
        if (raw_cap >= 0xFFFFFFF0ull)
            return fail(NK_EINVAL, "nkd_stage_raw: a step's raw text must stay below 4 GiB");
        bool ok = true;
        ok &= dalloc(d_raw, raw_cap + 64);
        ok &= dalloc(d_raw_next, raw_cap + 64);
        ok &= dalloc(d_out, raw_cap + 2 * raw_reads_cap + 64);
        ok &= dalloc(d_tile, raw_cap / NK_RAW_TILE + 4);
        ok &= dalloc(d_nlpos, raw_lines_cap);
        ok &= dalloc(d_nops, raw_reads_cap + 1);
        ok &= dalloc(d_opscan, raw_reads_cap + 1);
        ok &= dalloc(d_outlen, raw_reads_cap + 1);
        ok &= dalloc(d_outoff, raw_reads_cap + 1);
        ok &= dalloc(d_tout, NK_MAX_PARTITIONS);
        ok &= dalloc(d_rflags, 4);
        ok &= dalloc(d_summary, 6 * NK_MAX_PARTITIONS);
        ok &= dalloc(d_wins, NK_MAX_PARTITIONS);
        if (!ok || !be.prepare_scan((size_t)std::max<uint64_t>(raw_reads_cap + 1, raw_cap / NK_RAW_TILE + 4), err))
            return fail(NK_ENOMEM, "nkd_stage_raw: cannot allocate the raw-text scratch");
        raw_ready.store(true, std::memory_order_release);
        return NK_OK;
    }

    int upload_raw(const uint8_t *host_raw, size_t raw_bytes)
    {
        if (!host_raw)
        { /* forget what was sent ahead: the buffer it came from is about to be reused for something else */
            std::lock_guard<std::mutex> lock(up_mu);
            uploaded = nullptr;
            return NK_OK;
        }
        int rc = raw_prepare();
        if (rc)
            return rc;
        if ((raw_bytes & 15u) || raw_bytes > raw_cap)
            return fail(NK_EINVAL, "nkd_upload_raw: bad size");
        std::lock_guard<std::mutex> lock(up_mu);
        be.upload(d_raw_next, host_raw, raw_bytes);
        uploaded = host_raw;
        uploaded_bytes = raw_bytes;
        upload_total += raw_bytes;
        return NK_OK;
    }

    int stage_raw(const uint8_t *host_raw, size_t raw_bytes, const nkd_raw_segment *segs, int n_segs, int is_paired, int per)
    {
        if (!seeded)
            return fail(NK_EINVAL, "nkd_stage_raw before nkd_seed_finish");
        int rc = raw_prepare();
        if (rc)
            return rc;
        if ((per != 2 && per != 4) || n_segs < 1 || n_segs > (int)parts.size())
            return fail(NK_EINVAL, "nkd_stage_raw: bad segment count or lines per record");
        const unsigned stride = is_paired ? 2u : 1u;
        if (raw_bytes & 15u)
            return fail(NK_EINVAL, "nkd_stage_raw: raw_bytes must be a multiple of 16 (pad with spaces)");
        size_t bytes16 = raw_bytes;
        if (bytes16 > raw_cap)
            return fail(NK_EINVAL, "nkd_stage_raw: step exceeds max_raw_bytes");
        h_wins.assign((size_t)n_segs, NkRawWin{});
        std::fill(T.begin(), T.end(), 0u);
        std::vector<char> seen(parts.size(), 0);
        uint64_t recs = 0, lines = 0, outs = 0, prev_end = 0;
        for (int s = 0; s < n_segs; s++)
        {
            const nkd_raw_segment &g = segs[s];
            NkRawWin &w = h_wins[s];
            if (g.part >= parts.size() || seen[g.part])
                return fail(NK_EINVAL, "nkd_stage_raw: one segment per resident partition");
            seen[g.part] = 1;
            /* windows follow each other at the next 16-byte boundary, so that the copies below cover every byte
             * the line scan reads */
            if (g.n_records == 0 || g.fwd_off != ((prev_end + 15) & ~15ull) || (uint64_t)g.fwd_off + g.fwd_bytes > raw_bytes ||
                g.fwd_bytes == 0)
                return fail(NK_EINVAL, "nkd_stage_raw: forward window misplaced");
            prev_end = (uint64_t)g.fwd_off + g.fwd_bytes;
            if (is_paired)
            {
                if (g.rev_off != ((prev_end + 15) & ~15ull) || (uint64_t)g.rev_off + g.rev_bytes > raw_bytes || g.rev_bytes == 0)
                    return fail(NK_EINVAL, "nkd_stage_raw: reverse window misplaced");
                prev_end = (uint64_t)g.rev_off + g.rev_bytes;
            }
            w.f_off = g.fwd_off;
            w.f_bytes = g.fwd_bytes;
            w.r_off = is_paired ? g.rev_off : 0;
            w.r_bytes = is_paired ? g.rev_bytes : 0;
            w.n_records = g.n_records;
            w.part = g.part;
            w.f_line0 = (unsigned)lines;
            lines += (uint64_t)per * g.n_records;
            w.r_line0 = (unsigned)lines;
            if (is_paired)
                lines += (uint64_t)per * g.n_records;
            w.rec0 = (unsigned)recs;
            w.out0 = (unsigned)outs;
            recs += g.n_records;
            outs += (uint64_t)stride * g.n_records;
        }
        if (recs * stride > cfg.max_step_reads || lines + 1 > raw_lines_cap)
            return fail(NK_EINVAL, "nkd_stage_raw: step exceeds the read limit given to nkd_create");
        if (((prev_end + 15) & ~15ull) != raw_bytes)
            return fail(NK_EINVAL, "nkd_stage_raw: raw_bytes must end the last window (rounded up to 16)");
        {
            std::vector<int> need;
            for (int s = 0; s < n_segs; s++)
                need.push_back((int)segs[s].part);
            int rc2 = make_resident(need);
            if (rc2)
                return rc2;
        }
        /* windows go over as they are; the gaps keep whatever the host put there (neither '\n' nor NUL) */
        bool prefetched;
        {
            std::lock_guard<std::mutex> lock(up_mu);
            prefetched = uploaded == host_raw && uploaded_bytes == raw_bytes;
            uploaded = nullptr;
            if (prefetched)
            { /* the bytes were sent ahead (nkd_upload_raw): take that buffer once its copy has landed */
                std::swap(d_raw, d_raw_next);
                be.upload_fence();
            }
        }
        for (int s = 0; s < n_segs && !prefetched; s++)
        {
            const NkRawWin &w = h_wins[s];
            unsigned lo = w.f_off, hi = (w.f_off + w.f_bytes + 15u) & ~15u;
            be.h2d(d_raw + lo, host_raw + lo, hi - lo);
            h2d_bytes += hi - lo;
            if (is_paired)
            {
                lo = w.r_off;
                hi = (w.r_off + w.r_bytes + 15u) & ~15u;
                be.h2d(d_raw + lo, host_raw + lo, hi - lo);
                h2d_bytes += hi - lo;
            }
        }
        be.put_small(d_wins, h_wins.data(), (size_t)n_segs * sizeof(NkRawWin));
        be.zero(d_rflags, 4 * sizeof(unsigned));
        be.zero(d_nlpos, (size_t)(lines + 1) * sizeof(unsigned));
        raw = NkRaw{};
        raw.raw = d_raw;
        raw.raw_bytes = (unsigned)bytes16;
        raw.wins = d_wins;
        raw.n_wins = (unsigned)n_segs;
        raw.n_records = (unsigned)recs;
        raw.stride = stride;
        raw.per = (unsigned)per;
        raw.k = cfg.k;
        raw.min_len = cfg.k;
        raw.tile = d_tile;
        raw.nlpos = d_nlpos;
        raw.nlpos_cap = (unsigned)lines;
        raw.reads = d_reads;
        raw.nops = d_nops;
        raw.opscan = d_opscan;
        raw.t_out = d_tout;
        raw.flags = d_rflags;
        raw.accept = d_accept;
        raw.outlen = d_outlen;
        raw.outoff = d_outoff;
        raw.out = d_out;
        raw.summary = d_summary;
        raw.ctr = d_ctr;
        be.begin_timer(0); /* the step's device span starts with its parsing */
        be.reset_timer(9);
        be.begin_timer(9);
        be.raw_index(raw);
        be.end_timer(9);
        /* No round trip here: what the parse found (NUL bytes, long lines, a window that does not hold its records) is
         * read together with the step's counters after the first forward run (run_ops), which is undone if the text has
         * to go to the host parser after all.  Until then every partition's operation count is bounded by its bytes
         * (a sequence line has at most as many k-mers as it has characters). */
        for (int s = 0; s < n_segs; s++)
        {
            uint64_t bound = (uint64_t)h_wins[s].f_bytes + h_wins[s].r_bytes;
            if (bound >= (1u << NK_T_BITS))
                return fail(NK_EINVAL, "a partition has 2^28 or more bytes of record text in one step");
            T[h_wins[s].part] = (unsigned)bound;
        }
        raw_check = true;
        raw_lines_expected = lines;
        seq_view = d_raw;
        n_reads = (size_t)recs * stride;
        n_records = (size_t)recs;
        paired = is_paired;
        staged = true;
        raw_staged = true;
        ran = false;
        return NK_OK;
    }

    /* seed_kmer_hash (C:1322-1373) on raw record text: the first `limit` records of this piece of a file whose
     * sequence is longer than K are inserted with count 0 (sequence_to_hash_zero, C:1501-1537); *taken = how many */
    int seed_raw(const uint8_t *host_raw, size_t text_bytes, unsigned n_records, int per, unsigned limit, unsigned *taken,
                 int64_t *first_invalid)
    {
        if (seeded)
            return fail(NK_EINVAL, "nkd_seed_raw after nkd_seed_finish");
        int rc = raw_prepare();
        if (rc)
            return rc;
        const size_t bytes16 = (text_bytes + 15) & ~(size_t)15;
        const uint64_t lines = (uint64_t)per * n_records;
        if ((per != 2 && per != 4) || n_records == 0 || text_bytes == 0 || bytes16 > raw_cap || n_records > cfg.max_step_reads ||
            lines + 1 > raw_lines_cap)
            return fail(NK_EINVAL, "nkd_seed_raw: piece exceeds the limits given to nkd_create");
        h_wins.assign(1, NkRawWin{});
        h_wins[0].f_bytes = (unsigned)text_bytes;
        h_wins[0].n_records = n_records;
        bool prefetched;
        {
            std::lock_guard<std::mutex> lock(up_mu);
            prefetched = uploaded == host_raw && uploaded_bytes == bytes16;
            uploaded = nullptr;
            if (prefetched)
            { /* the piece was sent ahead while the previous one was being inserted (nkd_upload_raw) */
                std::swap(d_raw, d_raw_next);
                be.upload_fence();
            }
        }
        if (!prefetched)
        {
            be.h2d(d_raw, host_raw, bytes16);
            h2d_bytes += bytes16;
        }
        be.put_small(d_wins, h_wins.data(), sizeof(NkRawWin));
        be.zero(d_rflags, 4 * sizeof(unsigned));
        be.zero(d_nlpos, (size_t)(lines + 1) * sizeof(unsigned));
        raw = NkRaw{};
        raw.raw = d_raw;
        raw.raw_bytes = (unsigned)bytes16;
        raw.wins = d_wins;
        raw.n_wins = 1;
        raw.n_records = n_records;
        raw.stride = 1;
        raw.per = (unsigned)per;
        raw.k = cfg.k;
        raw.min_len = cfg.k + 1; /* strictly longer than K, C:1347 */
        raw.tile = d_tile;
        raw.nlpos = d_nlpos;
        raw.nlpos_cap = (unsigned)lines;
        raw.reads = d_reads;
        raw.nops = d_nops;
        raw.opscan = d_opscan;
        raw.t_out = d_tout;
        raw.flags = d_rflags;
        raw.outlen = d_outlen;
        raw.outoff = d_outoff;
        be.raw_records(raw);
        be.raw_limit(raw, limit);
        be.raw_number(raw);
        unsigned h_flags[4] = {0, 0, 0, 0}, h_t = 0, h_q = 0;
        be.d2h(h_flags, d_rflags, sizeof h_flags);
        be.d2h(&h_t, d_tout, sizeof h_t);
        be.d2h(&h_q, d_outoff + n_records, sizeof h_q);
        be.sync();
        if (h_flags[1] != lines || (h_flags[0] & (NK_RAW_NUL | NK_RAW_LONG | NK_RAW_SHAPE)))
        {
            if (h_flags[1] == lines && (h_flags[0] & (NK_RAW_NUL | NK_RAW_LONG)))
                return fail(NK_EIRREGULAR, "raw text needs the host parser (NUL byte or a line of 1024+ chars)");
            return fail(NK_EINVAL, "nkd_seed_raw: the piece does not hold the announced number of complete records");
        }
        if (h_t >= (1u << NK_T_BITS))
            return fail(NK_EINVAL, "a seed piece has 2^28 or more operations");
        std::fill(T.begin(), T.end(), 0u);
        T[0] = h_t;
        seq_view = d_raw;
        n_reads = n_records;
        raw_check = false;
        paired = 0;
        be.zero(d_invalid, n_reads + 1);
        std::vector<NkTable *> tabs{&seed};
        rc = run_ops(NK_MODE_SEED, tabs);
        if (rc)
            return rc;
        if (taken)
            *taken = h_q < limit ? h_q : limit;
        if (first_invalid)
        { /* the reference aborts at the first seed record that is not DNA (C:1349-1350) */
            be.zero(d_ctr, sizeof(NkCounters));
            be.decide(make_run(NK_MODE_SEED, 0, 0), (unsigned)n_reads, 0, cfg.coverage, d_accept);
            be.d2h(&h_ctr, d_ctr, 32);
            be.sync();
            *first_invalid = h_ctr.inv_max ? (int64_t)(NK_TMAX - h_ctr.inv_max) : -1;
        }
        return NK_OK;
    }

    int fetch_raw(int emit_mode, uint8_t *out, size_t out_cap, nkd_raw_result *results, int64_t *first_invalid, int slot)
    {
        if (!ran || !raw_staged)
            return fail(NK_EINVAL, "nkd_fetch_raw without nkd_stage_raw + nkd_run");
        if (emit_mode < 0 || emit_mode > 2)
            return fail(NK_EINVAL, "nkd_fetch_raw: bad emit mode");
        raw.emit_mode = emit_mode;
        raw.ctr = d_ctr; /* k_decide left the first non-DNA record there: the text kernels read it on the device */
        be.reset_timer(10);
        be.begin_timer(10);
        be.zero(d_summary, 6 * (size_t)raw.n_wins * sizeof(unsigned long long));
        be.copy_fence(); /* the previous step's text must have left d_out */
        be.raw_emit(raw);
        be.end_timer(10);
        be.end_timer(0);
        std::vector<unsigned long long> sm(6 * (size_t)raw.n_wins);
        be.d2h(sm.data(), d_summary, sm.size() * sizeof(unsigned long long));
        be.d2h(&h_ctr, d_ctr, 32);
        be.sync();
        int64_t inv = h_ctr.inv_max ? (int64_t)(NK_TMAX - h_ctr.inv_max) : -1;
        uint64_t total = 0;
        for (unsigned w = 0; w < raw.n_wins; w++)
            total = std::max<uint64_t>(total, std::max(sm[6 * w] + sm[6 * w + 1], sm[6 * w + 2] + sm[6 * w + 3]));
        if (total > out_cap)
            return fail(NK_EINVAL, "nkd_fetch_raw: output buffer too small");
        be.copy_out(out, d_out, (size_t)total, slot); /* on the copy stream: nkd_fetch_wait(slot) before `out` is read */
        d2h_bytes += total + sm.size() * 8 + 32;
        finish_timers();
        for (unsigned w = 0; w < raw.n_wins; w++)
        {
            nkd_raw_result &r = results[w];
            r.fwd_off = sm[6 * w];
            r.fwd_bytes = sm[6 * w + 1];
            r.rev_off = sm[6 * w + 2];
            r.rev_bytes = sm[6 * w + 3];
            r.processed = sm[6 * w + 4];
            r.printed = sm[6 * w + 5];
            nkd_part_stats &st = parts[h_wins[w].part].st;
            st.processed += r.processed;
            st.printed += r.printed;
            st.skipped += r.processed - r.printed;
        }
        if (first_invalid)
            *first_invalid = inv;
        staged = false;
        raw_staged = false;
        ran = false;
        return NK_OK;
    }

    double gaps[6] = {0, 0, 0, 0, 0, 0};
    void finish_timers()
    {
        /* where a step's span is not covered by the kernel classes: parse -> probe, probe -> open, open -> classify,
         * classify -> sort, decide -> emit (NKB200_TIMES prints the totals when the engine goes away) */
        gaps[0] += be.gap_ms(9, 1);
        gaps[1] += be.gap_ms(1, 2);
        gaps[2] += be.gap_ms(2, 4);
        gaps[3] += be.gap_ms(4, 5);
        gaps[4] += be.gap_ms(7, 10);
        gaps[5] += be.gap_ms(5, 7) + be.gap_ms(6, 7) * 0;
        be.timer_spans(0, spans);
        last_total_ms = be.timer_ms(0);
        last_probe_ms = be.timer_ms(1);
        rs.run_ms += last_total_ms;
        rs.probe_ms += last_probe_ms;
        rs.class_ms[0] += last_probe_ms;
        for (int t = 2; t <= 10; t++)
            rs.class_ms[t - 1] += be.timer_ms(t);
    }

    int seed_step(const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t nr, int64_t *first_invalid)
    {
        if (seeded)
            return fail(NK_EINVAL, "nkd_seed_step after nkd_seed_finish");
        int rc = stage(seq, seq_bytes, reads, nr, 0, 1, true);
        if (rc)
            return rc;
        be.zero(d_invalid, n_reads + 1);
        std::vector<NkTable *> tabs{&seed};
        rc = run_ops(NK_MODE_SEED, tabs);
        staged = false;
        if (rc)
            return rc;
        if (first_invalid)
        { /* the reference aborts at the first seed record that is not DNA (C:1349-1350) */
            be.zero(d_ctr, sizeof(NkCounters));
            be.decide(make_run(NK_MODE_SEED, 0, 0), (unsigned)n_reads, 0, cfg.coverage, d_accept);
            be.d2h(&h_ctr, d_ctr, 32);
            be.sync();
            *first_invalid = h_ctr.inv_max ? (int64_t)(NK_TMAX - h_ctr.inv_max) : -1;
        }
        return rc;
    }

    /* copy_hash_table from another engine's seed table on the same GPU: engines that share a GPU seed once */
    int seed_finish_from(NkEngine &src)
    {
        if (seeded)
            return fail(NK_EINVAL, "nkd_seed_finish called twice");
        if (!src.seed.tab || src.seeded)
            return fail(NK_EINVAL, "nkd_seed_finish_from: the source engine no longer holds its seed table");
        if (src.cfg.device != cfg.device || src.cfg.k != cfg.k || src.cfg.canonical != cfg.canonical)
            return fail(NK_EINVAL, "nkd_seed_finish_from: engines differ in GPU or k-mer settings");
        src.be.sync(); /* the source's seed steps are complete */
        be.release(seed.tab);
        seed = src.seed;
        seed.tab = nullptr;
        if (table_budget)
        { /* tables come into being when their partition is first used: this engine keeps its own seed table till then */
            seed.tab = (NkSlot *)be.alloc(seed.cap * sizeof(NkSlot));
            if (!seed.tab)
                return fail(NK_ENOMEM, "Memory allocation failed (seed table copy)");
            be.d2d(seed.tab, src.seed.tab, seed.cap * sizeof(NkSlot));
            be.sync();
            return seed_finish_lazy();
        }
        parts.resize(cfg.n_parts);
        for (int p = 0; p < cfg.n_parts; p++)
        {
            NkTable &t = parts[p];
            t = seed;
            t.st = nkd_part_stats{};
            t.st.capacity = seed.cap;
            t.st.used = seed.used;
            t.tab = (NkSlot *)be.alloc(seed.cap * sizeof(NkSlot));
            if (!t.tab)
                return fail(NK_ENOMEM, "Memory allocation failed (partition table copy)");
            be.d2d(t.tab, src.seed.tab, seed.cap * sizeof(NkSlot));
        }
        be.sync();
        seeded = true;
        return NK_OK;
    }

    int seed_finish_lazy()
    {
        parts.assign(cfg.n_parts, NkTable{});
        for (int p = 0; p < cfg.n_parts; p++)
        {
            NkTable &t = parts[p];
            t = seed;
            t.tab = nullptr;
            t.st = nkd_part_stats{};
            t.st.capacity = seed.cap;
            t.st.used = seed.used;
            t.fresh = true;
        }
        fresh_left = (unsigned)cfg.n_parts;
        be.sync();
        seeded = true;
        return NK_OK;
    }

    int seed_finish()
    {
        if (seeded)
            return fail(NK_EINVAL, "nkd_seed_finish called twice");
        if (table_budget)
            return seed_finish_lazy();
        parts.resize(cfg.n_parts);
        for (int p = 0; p < cfg.n_parts; p++)
        {
            NkTable &t = parts[p];
            t = seed;
            t.st = nkd_part_stats{};
            t.st.capacity = seed.cap;
            t.st.used = seed.used;
            if (p == cfg.n_parts - 1)
            { /* the last partition takes the seed table itself */
                seed.tab = nullptr;
                break;
            }
            t.tab = (NkSlot *)be.alloc(seed.cap * sizeof(NkSlot));
            if (!t.tab)
                return fail(NK_ENOMEM, "Memory allocation failed (partition table copy)");
            be.d2d(t.tab, seed.tab, seed.cap * sizeof(NkSlot));
        }
        be.sync();
        seeded = true;
        return NK_OK;
    }

    int run_step()
    {
        if (!seeded)
            return fail(NK_EINVAL, "nkd_run before nkd_seed_finish");
        if (!staged)
            return fail(NK_EINVAL, "nkd_run without nkd_stage");
        if (!raw_staged)
            be.begin_timer(0); /* a raw step's span started in stage_raw and ends in fetch_raw */
        for (int t = 1; t <= 8; t++)
            be.reset_timer(t);
        be.zero(d_high, (n_reads + 1) * sizeof(unsigned));
        be.zero(d_total, (n_reads + 1) * sizeof(unsigned));
        be.zero(d_invalid, n_reads + 1);
        std::vector<NkTable *> tabs;
        for (auto &p : parts)
            tabs.push_back(&p);
        int rc = run_ops(NK_MODE_SCORE, tabs);
        if (rc)
            return rc;
        be.zero(d_ctr, sizeof(NkCounters));
        be.begin_timer(7);
        be.decide(make_run(NK_MODE_SCORE, 0, 0), (unsigned)n_records, paired, cfg.coverage, d_accept);
        be.end_timer(7);
        if (!raw_staged)
            be.end_timer(0);
        ran = true;
        return NK_OK;
    }

    int fetch(uint8_t *accept, size_t nrec, int64_t *first_invalid)
    {
        if (!ran || raw_staged)
            return fail(NK_EINVAL, "nkd_fetch without nkd_stage + nkd_run");
        if (nrec != n_records)
            return fail(NK_EINVAL, "nkd_fetch: record count differs from the staged step");
        be.d2h(accept, d_accept, nrec);
        be.d2h(&h_ctr, d_ctr, 32);
        d2h_bytes += nrec + 32;
        be.sync();
        finish_timers();
        int64_t inv = h_ctr.inv_max ? (int64_t)(NK_TMAX - h_ctr.inv_max) : -1;
        if (first_invalid)
            *first_invalid = (inv >= 0 && (size_t)inv < nrec) ? inv : -1;
        for (size_t r = 0; r < nrec; r++)
        {
            if (inv >= 0 && (size_t)inv < nrec && r >= (size_t)inv)
                break; /* the reference stops at the first non-DNA record */
            nkd_part_stats &st = parts[rec_part[r]].st;
            st.processed++;
            if (accept[r])
                st.printed++;
            else
                st.skipped++;
        }
        staged = false;
        ran = false;
        return NK_OK;
    }

    int export_table(const NkTable &t, uint64_t *keys, int32_t *counts, uint64_t capacity)
    {
        flush_hot();
        if (!t.tab)
            return fail(NK_EINVAL, "table not resident");
        if (capacity != t.cap)
            return fail(NK_EINVAL, "export: capacity differs from the table's");
        const size_t chunk = 1u << 20;
        std::vector<NkSlot> buf(chunk);
        for (uint64_t o = 0; o < t.cap; o += chunk)
        {
            size_t n = (size_t)std::min<uint64_t>(chunk, t.cap - o);
            be.d2h(buf.data(), t.tab + o, n * sizeof(NkSlot));
            be.sync();
            for (size_t i = 0; i < n; i++)
            {
                keys[o + i] = buf[i].key;
                counts[o + i] = buf[i].count;
            }
        }
        return NK_OK;
    }

    /* ------------------------------------------------------------ table dump and merged table */

    void merge_release()
    {
        for (int i = 0; i < 2; i++)
        {
            be.release(d_mkeys[i]);
            be.release(d_mvals[i]);
            d_mkeys[i] = nullptr;
            d_mvals[i] = nullptr;
        }
        merge_cap = merge_n = merged_n = 0;
        merged_ready = false;
    }

    /* entries of a dump source: NKD_PART_SEED, NKD_PART_MERGED or a partition index */
    int dump_source(int part, NkDumpSrc &src, uint64_t &limit)
    {
        flush_hot();
        src = NkDumpSrc{nullptr, nullptr, nullptr};
        if (part == NKD_PART_MERGED)
        {
            if (!merged_ready)
                return fail(NK_EINVAL, "no merged table: call nkd_merge_finish first");
            src.keys = d_mkeys[0];
            src.vals = d_mvals[0];
            limit = merged_n;
            return NK_OK;
        }
        const NkTable *t = part == NKD_PART_SEED ? &seed : (part >= 0 && part < (int)parts.size() ? &parts[part] : nullptr);
        if (!t)
            return fail(NK_EINVAL, "no such partition");
        if (part >= 0 && !t->tab)
        {
            int rc = make_resident(std::vector<int>{part});
            if (rc)
                return rc;
        }
        if (!t->tab)
            return fail(NK_EINVAL, "table not resident");
        src.tab = t->tab;
        limit = t->cap;
        return NK_OK;
    }

    /* print_kmer_table's lines (C:368-380) for entries [lo, lo+n) of a source, formatted on the device */
    int dump_text(int part, uint64_t lo, uint64_t n, char *text, size_t text_cap, size_t *bytes)
    {
        NkDumpSrc src;
        uint64_t limit = 0;
        int rc = dump_source(part, src, limit);
        if (rc)
            return rc;
        if (lo > limit || n > limit - lo)
            return fail(NK_EINVAL, "nkd_dump_text: range beyond the table");
        *bytes = 0;
        if (n == 0)
            return NK_OK;
        uint64_t tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        unsigned long long *d_tile = (unsigned long long *)be.alloc((tiles + 1) * sizeof(unsigned long long));
        if (!d_tile || !be.dump_scan(src, lo, n, cfg.k, 1, d_tile))
        {
            be.release(d_tile);
            return fail(NK_ENOMEM, "nkd_dump_text: allocation failed");
        }
        unsigned long long total = 0;
        be.d2h(&total, d_tile + tiles, sizeof total);
        be.sync();
        if (total > text_cap)
        {
            be.release(d_tile);
            return fail(NK_EINVAL, "nkd_dump_text: text buffer too small");
        }
        if (total)
        {
            char *d_text = (char *)be.alloc(total);
            if (!d_text)
            {
                be.release(d_tile);
                return fail(NK_ENOMEM, "nkd_dump_text: allocation failed");
            }
            be.dump_text(src, lo, n, cfg.k, d_tile, d_text);
            be.d2h(text, d_text, total);
            be.sync();
            be.release(d_text);
            d2h_bytes += total;
        }
        be.release(d_tile);
        *bytes = (size_t)total;
        return NK_OK;
    }

    /* stored (k-mer, count) pairs of a table in slot order, packed on the device; dst on the device
     * (d_keys/d_vals) or on the host (keys/vals) */
    int compact(int part, unsigned long long *d_keys, long long *d_vals, uint64_t *keys, int64_t *vals, uint64_t cap_entries,
                uint64_t *n_out)
    {
        NkDumpSrc src;
        uint64_t limit = 0;
        int rc = dump_source(part, src, limit);
        if (rc)
            return rc;
        uint64_t tiles = (limit + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        unsigned long long *d_tile = (unsigned long long *)be.alloc((tiles + 1) * sizeof(unsigned long long));
        if (!d_tile || !be.dump_scan(src, 0, limit, cfg.k, 0, d_tile))
        {
            be.release(d_tile);
            return fail(NK_ENOMEM, "nkd_compact: allocation failed");
        }
        unsigned long long total = 0;
        be.d2h(&total, d_tile + tiles, sizeof total);
        be.sync();
        *n_out = total;
        if (total > cap_entries)
        {
            be.release(d_tile);
            return fail(NK_EINVAL, "nkd_compact: more stored k-mers than the destination holds");
        }
        bool to_host = d_keys == nullptr;
        if (to_host && total)
        {
            d_keys = (unsigned long long *)be.alloc(total * sizeof(unsigned long long));
            d_vals = (long long *)be.alloc(total * sizeof(long long));
            if (!d_keys || !d_vals)
            {
                be.release(d_keys);
                be.release(d_vals);
                be.release(d_tile);
                return fail(NK_ENOMEM, "nkd_compact: allocation failed");
            }
        }
        if (total)
            be.dump_pairs(src, 0, limit, d_tile, d_keys, d_vals);
        if (to_host && total)
        {
            be.d2h(keys, d_keys, total * sizeof(unsigned long long));
            be.d2h(vals, d_vals, total * sizeof(long long));
            be.sync();
            be.release(d_keys);
            be.release(d_vals);
            d2h_bytes += total * 16;
        }
        be.release(d_tile);
        return NK_OK;
    }

    int merge_begin(uint64_t max_entries)
    {
        merge_release();
        if (max_entries >= (1ull << 31))
            return fail(NK_EINVAL, "nkd_merge_begin: more than 2^31 entries");
        merge_cap = std::max<uint64_t>(max_entries, 1);
        bool ok = true;
        for (int i = 0; i < 2; i++)
        {
            ok &= dalloc(d_mkeys[i], merge_cap);
            ok &= dalloc(d_mvals[i], merge_cap);
        }
        if (!ok)
        {
            merge_release();
            return fail(NK_ENOMEM, "nkd_merge_begin: allocation failed");
        }
        return NK_OK;
    }
    int merge_add_part(int part)
    {
        if (!merge_cap)
            return fail(NK_EINVAL, "nkd_merge_add_part before nkd_merge_begin");
        uint64_t n = 0;
        int rc = compact(part, d_mkeys[0] + merge_n, d_mvals[0] + merge_n, nullptr, nullptr, merge_cap - merge_n, &n);
        if (!rc)
            merge_n += n;
        return rc;
    }
    int merge_add(const uint64_t *keys, const int64_t *vals, uint64_t n)
    {
        if (!merge_cap)
            return fail(NK_EINVAL, "nkd_merge_add before nkd_merge_begin");
        if (n > merge_cap - merge_n)
            return fail(NK_EINVAL, "nkd_merge_add: more entries than nkd_merge_begin announced");
        be.h2d(d_mkeys[0] + merge_n, keys, n * sizeof(uint64_t));
        be.h2d(d_mvals[0] + merge_n, vals, n * sizeof(int64_t));
        be.sync();
        h2d_bytes += n * 16;
        merge_n += n;
        return NK_OK;
    }
    int merge_finish(uint64_t *n_unique)
    {
        if (!merge_cap)
            return fail(NK_EINVAL, "nkd_merge_finish before nkd_merge_begin");
        merged_n = 0;
        if (merge_n)
        {
            unsigned long long *d_n = (unsigned long long *)be.alloc(sizeof(unsigned long long));
            if (!d_n || !be.merge_pairs(d_mkeys[0], d_mvals[0], merge_n, 2 * cfg.k, d_mkeys[1], d_mvals[1], d_n))
            {
                be.release(d_n);
                return fail(NK_ENOMEM, "nkd_merge_finish: allocation failed");
            }
            unsigned long long nu = 0;
            be.d2h(&nu, d_n, sizeof nu);
            be.sync();
            be.release(d_n);
            merged_n = nu;
        }
        merged_ready = true;
        if (n_unique)
            *n_unique = merged_n;
        return NK_OK;
    }

    int extract_keys(const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t nr, uint64_t *keys_out,
                     size_t n_ops, uint8_t *invalid_out)
    {
        int rc = stage(seq, seq_bytes, reads, nr, 0, 1, true);
        if (rc)
            return rc;
        if (n_ops != T[0])
            return fail(NK_EINVAL, "nkd_extract_keys: n_ops differs from the reads' window count");
        be.release(d_keys_out);
        d_keys_out = (unsigned long long *)be.alloc((n_ops + 1) * sizeof(unsigned long long));
        if (!d_keys_out)
            return fail(NK_ENOMEM, "nkd_extract_keys: allocation failed");
        be.zero(d_keys_out, (n_ops + 1) * sizeof(unsigned long long));
        be.zero(d_invalid, n_reads + 1);
        be.zero(d_ctr, sizeof(NkCounters));
        NkTable dummy = seed.tab ? seed : parts[0];
        std::vector<NkTable *> tabs{&dummy};
        std::vector<unsigned> lo(1, 0), hi(1, T[0]);
        upload_parts(tabs, lo, hi);
        be.probe(make_run(NK_MODE_KEYS, 0, 0));
        be.d2h(keys_out, d_keys_out, n_ops * sizeof(unsigned long long));
        if (invalid_out)
            be.d2h(invalid_out, d_invalid, nr);
        be.sync();
        staged = false;
        return NK_OK;
    }
};

#endif /* NK_ORCHESTRATE_H */
