/*
 * nk_core.h -- per-operation logic of the exact parallel emulation of the reference's
 * sequential k-mer table (normalise_kmers_multi_large.c, "C:n").
 *
 * Everything here is __host__ __device__ so that the sm_100a kernels (nk_engine.cu) and the
 * shuffled-order CPU emulation used ONLY by tests/ (tests/emu/nk_emu.cpp) execute the same
 * statements.  The product library contains the CUDA instantiation only.
 *
 * The reference processes one partition's k-mer windows strictly in order (C:1464, C:1559-1563):
 *     store(x): i = x % cap
 *               empty        -> claim: key=x, count=1, used++                    C:948-971
 *               key == x     -> count++                                          C:972-1008
 *               otherwise    -> walk i=(i+c*c)%cap, count++ on EVERY landed slot,
 *                               stop at an empty slot or at x; x is never stored  C:1009-1048
 *     then tests count[returned slot] >= depth                                   C:1494-1497
 *
 * Parallel restatement for one step of T operations with times t = 0..T-1:
 *   claim time   a slot that is empty at step start can only be taken by the earliest
 *                operation whose home it is -> 64-bit atomicMax of TAG|~t|open-index on the key field
 *   occupancy    "slot s is occupied at time t" = stored before the step, or claim time < t
 *   counters     every slot has a counter `count`; a slot claimed inside the step has a second
 *                one (`aux`, the count after its claim, offset 1) because the claim resets it (C:963)
 *   saturation   counts only grow, the test is a threshold: a counter whose value at step start
 *                is >= depth-1 makes every test on it true -> fire-and-forget RED
 *   pending      events on the other counters are listed, applied after all probing, and the
 *                tests on counters that end < depth are all false
 *   slow path    only counters that cross depth inside the step need their events ranked by
 *                time: sort by (slot, counter, t) and rank inside the segment
 */
#ifndef NK_CORE_H
#define NK_CORE_H

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define NK_HD __host__ __device__ __forceinline__
#else
#define NK_HD static inline
#endif

#define NK_LINE_SPLIT 1024u /* read_line cuts a line after 1023 chars (C:139, C:397): lines this long are split */
#define NK_TAG (1ull << 63)
#define NK_TMAX 0x7FFFFFFFu
#define NK_T_BITS 28 /* operations per partition per step < 2^28 (slow-path sort key budget) */

/* 16-byte table entry: same footprint as the reference's kmer_t {u64 hash; int count; pad} (C:157-161);
 * the pad carries step-local state and is 0 between steps. */
struct
#if defined(__CUDACC__)
    __align__(16)
#endif
        NkSlot
{
    unsigned long long key; /* 0 empty; bit 63 set = claim attempt inside the current step */
    int count;
    unsigned aux;
};

struct NkPart
{
    NkSlot *tab;
    unsigned long long cap;
    unsigned long long magic; /* floor((2^64-1)/cap) */
    unsigned long long gbase; /* first global slot number (slow-path sort key) */
    unsigned lo, hi;          /* operations [lo,hi) of this partition are live in this run */
    unsigned pad0, pad1;
};

struct NkRead /* layout == nkd_read (include/nk_b200.h) */
{
    unsigned seq_off, op_base;
    unsigned short len, part;
    unsigned reserved;
};

struct NkOpen /* an operation that met a slot which was empty at step start */
{
    unsigned long long key;
    unsigned t, read, slot, part;
    unsigned c;     /* walk step at which it stopped (0 = at home) */
    unsigned flags; /* bit0: stopped at its home slot (claim candidate) */
};

struct NkPend /* an increment of a counter that was not saturated at step start */
{
    unsigned slot;
    unsigned tw; /* t<<2 | which<<1 | terminal */
    int base;    /* special list: counter value at step start (1 for the post-claim counter); else unused */
    unsigned read;
};

struct NkClaim
{
    unsigned long long key;
    unsigned slot, t, part, pad;
};

#define NK_OVF_OPEN 1u
#define NK_OVF_PEND 2u
#define NK_OVF_CLAIM 4u
#define NK_OVF_WALK 8u
#define NK_OVF_SLOW 16u

struct NkCounters
{
    unsigned n_open, n_pend, n_claim, n_slow;
    unsigned overflow;
    unsigned inv_max; /* max over invalid records of NK_TMAX - record index; 0 = none */
    unsigned n_spec;
    unsigned pad;
    unsigned long long probe_touches; /* slots visited by k_probe (all partitions) */
    unsigned long long hot_hits;      /* of which: home hits on saturated counters served by the L2-resident hot table */
    unsigned long long touches[256]; /* per partition, slots visited */
    unsigned long long real_ops[256];
    unsigned claims[256];
};

/* Lists (pending, open, claim, slow) are appended to through per-warp chunks: a warp reserves `chunk`
 * entries with ONE global atomic and hands them out through a shared-memory cursor, because a single
 * append counter hit by every warp serialises the whole kernel at the L2 atomic unit.  Unused entries
 * of a chunk are written as holes and skipped by every consumer. */
enum
{
    NK_LIST_PEND = 0,
    NK_LIST_OPEN = 1,
    NK_LIST_CLAIM = 2,
    NK_LIST_SLOW = 3,
    NK_LIST_SPEC = 4, /* events on slots claimed inside the step (their counters need the apply round) */
    NK_NLISTS = 5
};
#define NK_HOLE 0xFFFFFFFFu

struct NkWarpCur /* two chunks per list: the second one is reserved ahead so a read never runs dry */
{
    unsigned base[NK_NLISTS], cap[NK_NLISTS], base2[NK_NLISTS], cap2[NK_NLISTS], used[NK_NLISTS];
};

enum
{
    NK_MODE_SCORE = 0,
    NK_MODE_SEED = 1,
    NK_MODE_COUNT = 2,
    NK_MODE_KEYS = 3
};

struct NkRun
{
    const unsigned char *seq;
    const NkRead *reads;
    unsigned n_reads;
    const NkPart *parts;
    int k, canonical, depth, mode;
    int delta;  /* +1 forward, -1 undo */
    int record; /* 0 in undo runs: no list appends, no claim attempts */
    unsigned *high, *total;
    unsigned char *invalid;
    NkOpen *open;
    unsigned open_cap;
    NkPend *pend;
    unsigned pend_cap;
    NkPend *spec;
    unsigned spec_cap;
    NkClaim *claim;
    unsigned claim_cap;
    unsigned long long *slow_key, *slow_val;
    unsigned slow_cap;
    NkCounters *ctr;
    unsigned long long *keys_out; /* NK_MODE_KEYS */
    unsigned chunk[NK_NLISTS];    /* entries a warp reserves at a time (device only) */
    struct NkHot *hot;            /* hot table (see nk_hot_*), or null */
    unsigned hot_mask;            /* entries - 1 (power of two) */
    unsigned *bloom;              /* one bit per hashed global slot: "this counter reached depth in this run" */
    unsigned bloom_words;         /* power of two */
    NkWarpCur *wcur;              /* this warp's cursors in shared memory (device only) */
};

/* ---------------------------------------------------------------- primitives */

#if defined(__CUDA_ARCH__)
#define NK_DEVICE_CODE 1
#else
#define NK_DEVICE_CODE 0
#endif

NK_HD unsigned long long nk_atomic_max64(unsigned long long *p, unsigned long long v)
{
#if NK_DEVICE_CODE
    return atomicMax(p, v);
#else
    unsigned long long o = *p;
    if (v > o)
        *p = v;
    return o;
#endif
}
NK_HD unsigned nk_atomic_max32(unsigned *p, unsigned v)
{
#if NK_DEVICE_CODE
    return atomicMax(p, v);
#else
    unsigned o = *p;
    if (v > o)
        *p = v;
    return o;
#endif
}
NK_HD void nk_red_add32(int *p, int v)
{
#if NK_DEVICE_CODE
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
/* count and aux live in one 8-byte word (count low, aux high): +1/+1 or -1/-1 in ONE atomic */
#define NK_BOTH_PLUS 0x0000000100000001ull
#define NK_BOTH_MINUS 0xFFFFFFFEFFFFFFFFull
NK_HD unsigned long long nk_atomic_add64(unsigned long long *p, unsigned long long v)
{
#if NK_DEVICE_CODE
    return atomicAdd(p, v);
#else
    unsigned long long o = *p;
    *p += v;
    return o;
#endif
}
NK_HD void nk_red_add64(unsigned long long *p, unsigned long long v)
{
#if NK_DEVICE_CODE
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
NK_HD void nk_red_or32(unsigned *p, unsigned v)
{
#if NK_DEVICE_CODE
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
/* L2-resident filter of the counters that reached depth inside the current run (no false negatives) */
NK_HD void nk_bloom_pos(const NkRun &P, unsigned long long gslot, unsigned &word, unsigned &bit)
{
    unsigned long long h = gslot * 0x9E3779B97F4A7C15ull;
    word = (unsigned)(h >> 37) & (P.bloom_words - 1u);
    bit = 1u << ((unsigned)(h >> 32) & 31u);
}
NK_HD void nk_bloom_set(const NkRun &P, unsigned long long gslot)
{
    unsigned w, b;
    nk_bloom_pos(P, gslot, w, b);
    nk_red_or32(&P.bloom[w], b);
}
NK_HD bool nk_bloom_test(const NkRun &P, unsigned long long gslot)
{
    unsigned w, b;
    nk_bloom_pos(P, gslot, w, b);
    return (P.bloom[w] & b) != 0;
}
/* reserve one entry of list `list` (global counter gctr).  Device: every lane bumps its warp's cursor with its
 * own shared-memory atomic.  No warp aggregation on purpose: appends happen in divergent code, and a
 * __activemask()/__shfl_sync(mask) group can contain a lane that branches away and then waits at the
 * warp barrier of the chunk rotation forever (seen on B200 in k_open). */
NK_HD unsigned nk_list_append(const NkRun &P, int list, unsigned *gctr)
{
#if NK_DEVICE_CODE
    NkWarpCur *wc = P.wcur;
    if (wc == nullptr)
        return atomicAdd(gctr, 1u);
    unsigned my = atomicAdd(&wc->used[list], 1u);
    if (my < wc->cap[list])
        return wc->base[list] + my;
    my -= wc->cap[list];
    if (my < wc->cap2[list])
        return wc->base2[list] + my;
    return atomicAdd(gctr, 1u); /* both chunks exhausted inside one unit of work: rare */
#else
    (void)P;
    (void)list;
    return (*gctr)++;
#endif
}

#if defined(__CUDACC__)
/* Converged-warp call before a unit of work: when the first chunk is used up, the second becomes the first
 * and a fresh one is reserved with ONE global atomic.  No holes are produced here. */
__device__ __forceinline__ void nk_chunk_rotate(const NkRun &P, int list, unsigned *gctr, unsigned gcap)
{
    NkWarpCur *wc = P.wcur;
    const unsigned lane = threadIdx.x & 31;
    for (int round = 0; round < 2; round++)
    {
        /* The decision must be a warp vote: a lane that skipped the rotation may start appending (and bump
         * `used`) while a slower lane is still looking at the cursor; a per-lane test then sends that lane
         * into the barrier below alone, where it pairs with the others' next barrier and the warp deadlocks. */
        if (!__any_sync(0xFFFFFFFFu, wc->used[list] >= wc->cap[list]))
            break;
        if (lane == 0)
        {
            unsigned over = wc->used[list] - wc->cap[list];
            unsigned ch = P.chunk[list];
            unsigned b = atomicAdd(gctr, ch);
            unsigned c = b >= gcap ? 0u : (gcap - b < ch ? gcap - b : ch);
            if (c < ch)
                atomicOr(&P.ctr->overflow, list == NK_LIST_PEND ? NK_OVF_PEND : list == NK_LIST_OPEN ? NK_OVF_OPEN : NK_OVF_CLAIM);
            wc->base[list] = wc->base2[list];
            wc->cap[list] = wc->cap2[list];
            wc->used[list] = over < wc->cap2[list] ? over : wc->cap2[list];
            wc->base2[list] = b;
            wc->cap2[list] = c;
        }
        __syncwarp();
    }
}
/* end of kernel: the unused entries of both chunks become holes; hole(idx) writes one hole record */
template <class HoleFn>
__device__ __forceinline__ void nk_chunk_close(const NkRun &P, int list, HoleFn hole)
{
    NkWarpCur *wc = P.wcur;
    const unsigned lane = threadIdx.x & 31;
    __syncwarp();
    unsigned used = wc->used[list], cap = wc->cap[list], cap2 = wc->cap2[list];
    unsigned u1 = used < cap ? used : cap, u2 = used > cap ? (used - cap < cap2 ? used - cap : cap2) : 0u;
    for (unsigned i = u1 + lane; i < cap; i += 32)
        hole(wc->base[list] + i);
    for (unsigned i = u2 + lane; i < cap2; i += 32)
        hole(wc->base2[list] + i);
    __syncwarp();
}
#endif

NK_HD NkSlot nk_load_slot(const NkSlot *p)
{
#if NK_DEVICE_CODE
    /* Plain ld.global: measured on B200 (profiles/microbench) L1-bypassing loads (.cg/.cv/no_allocate) gather
     * at half the rate (18.9 vs 35.4 G/s).  A stale L1 line is harmless here: counts that matter are either
     * untouched during the launch or only need ">= depth-1", and an unseen claim TAG reads as "empty at step
     * start", which is what a TAG means.  asm volatile keeps the compiler from merging re-reads in the walk loop. */
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    NkSlot s;
    s.key = ((unsigned long long)v.y << 32) | v.x;
    s.count = (int)v.z;
    s.aux = v.w;
    return s;
#else
    return *p;
#endif
}
NK_HD unsigned long long nk_mulhi64(unsigned long long a, unsigned long long b)
{
#if NK_DEVICE_CODE
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}
/* x % d with d's reciprocal m = floor((2^64-1)/d): the reference's per-probe 64-bit modulo (C:936, C:1028) */
NK_HD unsigned long long nk_mod(unsigned long long x, unsigned long long d, unsigned long long m)
{
    unsigned long long q = nk_mulhi64(x, m);
    unsigned long long r = x - q * d;
    while (r >= d)
        r -= d;
    return r;
}

/* is operation time t of this partition inside the window the current launch works on? */
NK_HD bool nk_live(const NkPart &pd, unsigned t) { return t >= pd.lo && t < pd.hi; }
NK_HD bool nk_is_real(unsigned long long f) { return f != 0 && !(f & NK_TAG); }
NK_HD unsigned nk_tag_time(unsigned long long f) { return NK_TMAX - (unsigned)((f >> 32) & NK_TMAX); }
NK_HD unsigned nk_tag_open(unsigned long long f) { return (unsigned)f; }
NK_HD unsigned long long nk_make_tag(unsigned t, unsigned idx)
{
    return NK_TAG | ((unsigned long long)(NK_TMAX - t) << 32) | idx;
}

/* ---------------------------------------------------------------- codec (C:1118-1126, C:1160-1180) */

/* encoding of the reverse complement of a k-mer: complement, reverse the 2-bit groups */
NK_HD unsigned long long nk_revcomp(unsigned long long x, int k)
{
    x = ~x;
#if NK_DEVICE_CODE
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = __builtin_bswap64(x);
#endif
    x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
    return x >> (64 - 2 * k);
}

/* scalar reference of the packed path: A0 C1 G2 T3, anything else (incl. N) 0, MSB-first */
NK_HD unsigned long long nk_window_key_ascii(const unsigned char *s, int k, int canonical)
{
    unsigned long long x = 0;
    for (int i = 0; i < k; i++)
    {
        unsigned b = s[i];
        unsigned code = (b == 'C') ? 1u : (b == 'G') ? 2u
                                      : (b == 'T')   ? 3u
                                                     : 0u;
        x = (x << 2) | code;
    }
    if (canonical)
    {
        unsigned long long r = nk_revcomp(x, k);
        if (r < x)
            x = r;
    }
    return x;
}

/* ---------------------------------------------------------------- hot table
 *
 * Most operations of a skewed data set are home hits on counters that passed depth-1 long ago: their test is true
 * whatever the exact count (saturation, above) and their increments commute.  Such a counter may therefore be kept
 * in two places: the table's count, and a `pending` sum in a small table of hot keys that stays in L2 (16 MB per
 * engine).  A hit there needs no DRAM access at all; the table's count lags behind by `pending` until the sums are
 * folded back in (nk_hot_flush_op) -- before anything reads exact counts: a re-hash (which also moves the slots, so
 * the hot table is cleared), parking a table in host memory, the -P dump, exports.  Exact by construction: an entry
 * is only created by an operation that saw the key stored at its home slot with a step-start count >= depth-1, stored
 * keys do not move between re-hashes, and counts never fall. */
struct
#if defined(__CUDACC__)
    __align__(16)
#endif
        NkHot
{
    unsigned long long key; /* 0 = free */
    unsigned part1;         /* partition + 1 (written after the key was claimed; 0 = not valid yet) */
    int pending;            /* increments not yet added to the table's count */
};

NK_HD unsigned nk_hot_index(const NkRun &P, unsigned long long key, unsigned part)
{
    unsigned long long h = (key ^ ((unsigned long long)(part + 1u) << 56)) * 0x9E3779B97F4A7C15ull;
    return (unsigned)(h >> 40) & P.hot_mask;
}
NK_HD unsigned long long nk_atomic_cas64(unsigned long long *p, unsigned long long expect, unsigned long long v)
{
#if NK_DEVICE_CODE
    return atomicCAS(p, expect, v);
#else
    unsigned long long o = *p;
    if (o == expect)
        *p = v;
    return o;
#endif
}
/* true when the operation was absorbed: the counter is saturated, the test is true, the increment is noted */
NK_HD bool nk_hot_hit(const NkRun &P, unsigned long long key, unsigned part, int &high_acc)
{
    NkHot *e = &P.hot[nk_hot_index(P, key, part)];
#if NK_DEVICE_CODE
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(e));
    const unsigned long long ekey = ((unsigned long long)v.y << 32) | v.x;
    const unsigned epart = v.z;
#else
    const unsigned long long ekey = e->key;
    const unsigned epart = e->part1;
#endif
    if (ekey != key || epart != part + 1u)
        return false;
    nk_red_add32(&e->pending, P.delta);
    high_acc += P.delta;
    return true;
}
/* called by an operation that has just seen `key` stored at its home slot with a saturated counter */
NK_HD void nk_hot_install(const NkRun &P, unsigned long long key, unsigned part)
{
    NkHot *e = &P.hot[nk_hot_index(P, key, part)];
    if (e->key == 0 && nk_atomic_cas64(&e->key, 0ull, key) == 0ull)
        e->part1 = part + 1u;
}
/* fold entry i's pending increments into the table (every table of the engine must be resident and described by P.parts) */
NK_HD void nk_hot_flush_op(const NkRun &P, unsigned i)
{
    NkHot *e = &P.hot[i];
    if (e->key == 0 || e->part1 == 0 || e->pending == 0)
        return;
    const NkPart &pd = P.parts[e->part1 - 1u];
    if (!pd.tab)
        return; /* parked: its sums were folded in before it left */
    unsigned long long home = nk_mod(e->key, pd.cap, pd.magic);
    nk_red_add32(&pd.tab[home].count, e->pending);
    e->pending = 0;
}
NK_HD void nk_hot_clear_op(const NkRun &P, unsigned i, unsigned part1)
{
    NkHot *e = &P.hot[i];
    if (part1 == 0 || e->part1 == part1)
    {
        e->key = 0;
        e->part1 = 0;
        e->pending = 0;
    }
}

/* ---------------------------------------------------------------- events */

/* Increment of the ordinary counter of a slot that is NOT claimed inside this step (stored before the step, or
 * a ghost that stays empty).  Unsaturated increments bump count and aux together in one 64-bit atomic, so
 * count - aux never exceeds the counter's value at step start and equals it whenever aux was 0 then.  aux is
 * zeroed when a counter is found to have reached depth (nk_classify_op) -- from then on it is saturated and aux
 * stays 0 -- and when a slot is claimed (nk_prepare_claim_op).  A stale aux on a still-cold counter only makes it
 * look less saturated, which costs a listing but never changes a result: the test outcome of a listed event comes
 * from the counter's final value (below depth: false) or from time ranking with base = final - events listed.
 * The atomic's return value tells the event that takes the counter to `depth`: it marks the run's filter, so that
 * k_classify only revisits counters that reached depth.  The same holds for the -1 replay of an abandoned run. */
NK_HD void nk_event(const NkRun &P, const NkPart &pd, unsigned slot, int count, unsigned aux, int terminal, unsigned t,
                    unsigned read, int &high_acc)
{
    if (count - (int)aux >= P.depth - 1)
    {
        nk_red_add32(&pd.tab[slot].count, P.delta);
        if (terminal)
            high_acc += P.delta;
        return;
    }
    unsigned long long *word = reinterpret_cast<unsigned long long *>(&pd.tab[slot].count);
    if (P.delta < 0)
    {
        nk_red_add64(word, NK_BOTH_MINUS);
        return;
    }
    unsigned long long old = nk_atomic_add64(word, NK_BOTH_PLUS);
    if ((int)(unsigned)old + 1 >= P.depth)
        nk_bloom_set(P, pd.gbase + slot);
    if (P.record)
    {
        unsigned idx = nk_list_append(P, NK_LIST_PEND, &P.ctr->n_pend);
        if (idx < P.pend_cap)
        {
            NkPend r;
            r.slot = slot;
            r.tw = (t << 2) | (unsigned)terminal;
            r.base = 0;
            r.read = read;
            P.pend[idx] = r;
        }
        else
            nk_red_or32(&P.ctr->overflow, NK_OVF_PEND);
    }
}

/* Increment of a counter of a slot that IS claimed inside this step: which 0 = the ghost count before the
 * claim, which 1 = aux, the count after the claim (value 1 + aux, C:963).  base = value at step start. These
 * counters are not touched until all probing is done (nk_apply_op), so base is what every reader sees. */
NK_HD void nk_event_claimed(const NkRun &P, const NkPart &pd, unsigned slot, int which, int base, int terminal, unsigned t,
                            unsigned read, int &high_acc)
{
    if (base >= P.depth - 1)
    {
        int *ctr = which ? reinterpret_cast<int *>(&pd.tab[slot].aux) : &pd.tab[slot].count;
        nk_red_add32(ctr, P.delta);
        if (terminal)
            high_acc += P.delta;
    }
    else if (P.record)
    {
        unsigned idx = nk_list_append(P, NK_LIST_SPEC, &P.ctr->n_spec);
        if (idx < P.spec_cap)
        {
            NkPend r;
            r.slot = slot;
            r.tw = (t << 2) | ((unsigned)which << 1) | (unsigned)terminal;
            r.base = base;
            r.read = read;
            P.spec[idx] = r;
        }
        else
            nk_red_or32(&P.ctr->overflow, NK_OVF_PEND);
    }
}

NK_HD void nk_defer(const NkRun &P, unsigned long long key, unsigned t, unsigned read, unsigned part, unsigned slot,
                    unsigned c, unsigned at_home, NkSlot *home_slot)
{
    if (!P.record)
        return;
    unsigned idx = nk_list_append(P, NK_LIST_OPEN, &P.ctr->n_open);
    if (idx >= P.open_cap)
    {
        nk_red_or32(&P.ctr->overflow, NK_OVF_OPEN);
        return;
    }
    NkOpen o;
    o.key = key;
    o.t = t;
    o.read = read;
    o.slot = slot;
    o.part = part;
    o.c = c;
    o.flags = at_home;
    P.open[idx] = o;
    if (at_home) /* claim attempt: earliest operation wins, payload finds its key later */
        nk_atomic_max64(&home_slot->key, nk_make_tag(t, idx));
}

#define NK_MAX_WALK 4096 /* watchdog: at load <= 0.8 a legitimate walk this long has probability ~0.8^4096 */

/* phase 1: everything about operation (key,t) that does not depend on in-step claims.
 * Returns the number of slots visited (touches).  */
NK_HD unsigned nk_probe_op(const NkRun &P, const NkPart &pd, unsigned part, unsigned long long key, unsigned t,
                           unsigned read, int &high_acc, unsigned &hot_hits)
{
    if (P.mode == NK_MODE_SCORE && P.hot && nk_hot_hit(P, key, part, high_acc))
    { /* a home hit on a saturated counter (C:972-1008 with a count that no test can tell from the exact one) */
        hot_hits++;
        return 1;
    }
    unsigned long long i = nk_mod(key, pd.cap, pd.magic);
    NkSlot e = nk_load_slot(&pd.tab[i]);
    if (P.mode == NK_MODE_SEED)
    { /* init mode changes no count (all are 0): only claims are observable (C:963, C:994-1000, C:1046) */
        if (!nk_is_real(e.key))
            nk_defer(P, key, t, read, part, (unsigned)i, 0, 1, &pd.tab[i]);
        return 1;
    }
    if (e.key == key)
    {
        nk_event(P, pd, (unsigned)i, e.count, e.aux, 1, t, read, high_acc);
        if (P.mode == NK_MODE_SCORE && P.hot && P.record && e.count - (int)e.aux >= P.depth - 1)
            nk_hot_install(P, key, part); /* from now on this key's operations stay in L2 */
        return 1;
    }
    if (!nk_is_real(e.key))
    {
        nk_defer(P, key, t, read, part, (unsigned)i, 0, 1, &pd.tab[i]);
        return 1;
    }
    unsigned touches = 1;
    for (unsigned c = 1;; c++)
    {
        if (c > NK_MAX_WALK)
        {
            nk_red_or32(&P.ctr->overflow, NK_OVF_WALK);
            break;
        }
        i = nk_mod(i + (unsigned long long)c * c, pd.cap, pd.magic);
        e = nk_load_slot(&pd.tab[i]);
        if (!nk_is_real(e.key))
        { /* empty at step start: whether it is still empty at time t is decided in phase 2 */
            nk_defer(P, key, t, read, part, (unsigned)i, c, 0, nullptr);
            break;
        }
        touches++;
        int term = e.key == key;
        nk_event(P, pd, (unsigned)i, e.count, e.aux, term, t, read, high_acc);
        if (term)
            break;
    }
    return touches;
}

/* phase 2: finish a deferred operation now that every claim time is known */
NK_HD unsigned nk_open_op(const NkRun &P, unsigned idx, int &high_acc, int &claimed)
{
    NkOpen o = P.open[idx];
    claimed = 0;
    if (o.flags == NK_HOLE)
        return 0;
    const NkPart &pd = P.parts[o.part];
    if (!nk_live(pd, o.t))
        return 0;
    unsigned long long i = o.slot;
    unsigned c = o.c;
    unsigned touches = 0;
    bool landed = true; /* slot i has been reached but not evaluated yet */
    if (o.flags & 1u)
    {
        unsigned long long f = nk_load_slot(&pd.tab[i]).key;
        unsigned tc = nk_tag_time(f);
        if (tc == o.t)
        { /* this operation stores the key: count becomes 1 (0 when seeding), test is false (depth >= 2) */
            if (P.record)
            {
                unsigned ci = nk_list_append(P, NK_LIST_CLAIM, &P.ctr->n_claim);
                if (ci < P.claim_cap)
                {
                    NkClaim cl;
                    cl.key = o.key;
                    cl.slot = (unsigned)i;
                    cl.t = o.t;
                    cl.part = o.part;
                    cl.pad = 0;
                    P.claim[ci] = cl;
                }
                else
                    nk_red_or32(&P.ctr->overflow, NK_OVF_CLAIM);
                claimed = 1; /* the caller adds it to ctr->claims[o.part] */
            }
            return 0;
        }
        if (P.mode == NK_MODE_SEED)
            return 0;
        /* an earlier operation of this step owns the home slot */
        unsigned long long owner = P.open[nk_tag_open(f)].key;
        if (owner == o.key)
        {
            nk_event_claimed(P, pd, (unsigned)i, 1, 1, 1, o.t, o.read, high_acc);
            return 0;
        }
        landed = false; /* collision at home: walk */
    }
    for (;;)
    {
        if (!landed)
        {
            c++;
            if (c > NK_MAX_WALK)
            {
                nk_red_or32(&P.ctr->overflow, NK_OVF_WALK);
                break;
            }
            i = nk_mod(i + (unsigned long long)c * c, pd.cap, pd.magic);
        }
        landed = false;
        touches++;
        NkSlot e = nk_load_slot(&pd.tab[i]);
        if (nk_is_real(e.key))
        {
            int term = e.key == o.key;
            nk_event(P, pd, (unsigned)i, e.count, e.aux, term, o.t, o.read, high_acc);
            if (term)
                break;
            continue;
        }
        if (e.key == 0)
        { /* stays empty for the whole step: the walk ends on a ghost counter (C:1015, C:1043-1044) */
            nk_event(P, pd, (unsigned)i, e.count, e.aux, 1, o.t, o.read, high_acc);
            break;
        }
        if (nk_tag_time(e.key) > o.t)
        { /* still empty at time t, claimed later in this step */
            nk_event_claimed(P, pd, (unsigned)i, 0, e.count, 1, o.t, o.read, high_acc);
            break;
        }
        /* claimed earlier in this step by another key (a walker never meets its own key here:
         * keys are stored at their home slot only, and this walker's home holds a different key) */
        nk_event_claimed(P, pd, (unsigned)i, 1, 1, 0, o.t, o.read, high_acc);
    }
    return touches;
}

/* slow-path records: key = (global slot, counter, t, terminal); value = (x, read, kind) where kind 1 means x is
 * the counter's FINAL value (base = x - number of events of the segment) and kind 0 means x is its base */
NK_HD void nk_slow_write(const NkRun &P, unsigned si, unsigned long long gslot, const NkPend &r, int x, int kind)
{
    unsigned t = r.tw >> 2;
    P.slow_key[si] = (gslot << (NK_T_BITS + 2)) | ((unsigned long long)((r.tw >> 1) & 1u) << (NK_T_BITS + 1)) |
                     ((unsigned long long)t << 1) | (r.tw & 1u);
    P.slow_val[si] = ((unsigned long long)(unsigned)x << 32) | ((unsigned long long)r.read << 1) | (unsigned)kind;
}

/* ordinary counters: all increments are in.  A counter that never reached depth (not in the filter, or a false
 * positive whose final count is below depth) had only false tests.  Otherwise its events are ranked by time
 * (returns true: the caller appends a slow record) and aux goes back to 0. */
NK_HD bool nk_classify_op(const NkRun &P, unsigned idx, NkPend &r, unsigned long long &gslot, int &x)
{
    r = P.pend[idx];
    if (r.slot == NK_HOLE)
        return false;
    const NkPart &pd = P.parts[P.reads[r.read].part];
    if (!nk_live(pd, r.tw >> 2))
        return false;
    gslot = pd.gbase + r.slot;
    if (!nk_bloom_test(P, gslot))
        return false;
    NkSlot *s = &pd.tab[r.slot];
    NkSlot e = nk_load_slot(s);
    if (e.count < P.depth)
        return false;
    if (e.aux != 0)
        s->aux = 0;
    x = e.count;
    return true;
}

/* between probe and the deferred operations: the operation that claims a slot zeroes its aux, which becomes
 * the slot's post-claim counter (the ghost's stale aux, if any, has no meaning any more) */
NK_HD void nk_prepare_claim_op(const NkRun &P, unsigned idx)
{
    NkOpen o = P.open[idx];
    if (o.flags == NK_HOLE || !(o.flags & 1u))
        return;
    NkSlot *s = &P.parts[o.part].tab[o.slot];
    if ((s->key & NK_TAG) && nk_tag_time(s->key) == o.t && s->aux != 0)
        s->aux = 0;
}

/* counters of slots claimed inside the step: apply the listed increments ... */
NK_HD void nk_apply_op(const NkRun &P, unsigned idx)
{
    NkPend r = P.spec[idx];
    if (r.slot == NK_HOLE)
        return;
    const NkPart &pd = P.parts[P.reads[r.read].part];
    if (!nk_live(pd, r.tw >> 2))
        return;
    int *ctr = (r.tw & 2u) ? reinterpret_cast<int *>(&pd.tab[r.slot].aux) : &pd.tab[r.slot].count;
    nk_red_add32(ctr, 1);
}

/* ... then classify them the same way */
NK_HD bool nk_classify_claimed_op(const NkRun &P, unsigned idx, NkPend &r, unsigned long long &gslot, int &x)
{
    r = P.spec[idx];
    if (r.slot == NK_HOLE)
        return false;
    const NkPart &pd = P.parts[P.reads[r.read].part];
    if (!nk_live(pd, r.tw >> 2))
        return false;
    NkSlot e = nk_load_slot(&pd.tab[r.slot]);
    long long v = (r.tw & 2u) ? 1ll + (long long)e.aux : (long long)e.count;
    gslot = pd.gbase + r.slot;
    x = r.base;
    return v >= P.depth;
}

/* phase 5: rank of event i inside its (slot,counter) segment of the time-sorted list */
NK_HD void nk_rank_op(const NkRun &P, const unsigned long long *keys, const unsigned long long *vals, unsigned n,
                      unsigned i)
{
    unsigned long long key = keys[i];
    if (!(key & 1) || key == ~0ull)
        return; /* only the terminal landing is tested (C:1494); all-ones = hole */
    unsigned long long seg = key >> (NK_T_BITS + 1);
    unsigned lo = 0, hi = i; /* first index whose segment is >= seg */
    while (lo < hi)
    {
        unsigned mid = (lo + hi) >> 1;
        if ((keys[mid] >> (NK_T_BITS + 1)) < seg)
            lo = mid + 1;
        else
            hi = mid;
    }
    unsigned long long val = vals[i];
    long long base = (long long)(int)(val >> 32);
    if (val & 1)
    { /* x is the final value: base = final - events of this counter in the step */
        unsigned a = i, b = n; /* first index whose segment is > seg */
        while (a < b)
        {
            unsigned mid = (a + b) >> 1;
            if ((keys[mid] >> (NK_T_BITS + 1)) <= seg)
                a = mid + 1;
            else
                b = mid;
        }
        base -= (long long)(a - lo);
    }
    long long after = base + (long long)(i - lo + 1);
    if (after >= P.depth)
    {
        unsigned read = (unsigned)((val & 0xFFFFFFFFull) >> 1);
#if NK_DEVICE_CODE
        atomicAdd(&P.high[read], 1u);
#else
        P.high[read] += 1u;
#endif
    }
}

/* phase 6: store the claimed keys (C:962-965) */
NK_HD void nk_commit_op(const NkRun &P, unsigned idx)
{
    NkClaim cl = P.claim[idx];
    if (cl.slot == NK_HOLE || !nk_live(P.parts[cl.part], cl.t))
        return;
    NkSlot *s = &P.parts[cl.part].tab[cl.slot];
    NkSlot n;
    n.key = cl.key;
    n.count = (P.mode == NK_MODE_SEED) ? 0 : (int)(1u + s->aux);
    n.aux = 0;
    *s = n;
}

/* undo: forget the claim attempts of an abandoned run */
NK_HD void nk_untag_op(const NkRun &P, unsigned idx)
{
    NkOpen o = P.open[idx];
    if (o.flags == NK_HOLE || !(o.flags & 1u) || !nk_live(P.parts[o.part], o.t))
        return;
    NkSlot *s = &P.parts[o.part].tab[o.slot];
    if (s->key & NK_TAG)
        s->key = 0;
    s->aux = 0;
}

/* ---------------------------------------------------------------- growth (C:1055-1108) */

/* Sequential re-insertion in old-slot order with linear probing == linear probing with
 * priority "lower old index wins" (history-independent layout).  aux of the new table holds
 * 0xFFFFFFFF - old index of the current occupant while placing. */
NK_HD void nk_rehash_place_op(const NkSlot *old_tab, unsigned long long old_i, NkSlot *new_tab, unsigned long long ncap,
                              unsigned long long nmagic)
{
    unsigned long long key = old_tab[old_i].key;
    if (key == 0)
        return;
    unsigned cur = 0xFFFFFFFFu - (unsigned)old_i;
    unsigned long long j = nk_mod(key, ncap, nmagic);
    for (;;)
    {
        unsigned prev = nk_atomic_max32(&new_tab[j].aux, cur);
        if (prev == 0)
            return;
        if (prev < cur)
            cur = prev; /* displaced a later key: carry it on */
        j = (j + 1 == ncap) ? 0 : j + 1;
    }
}

NK_HD void nk_rehash_fill_op(const NkSlot *old_tab, NkSlot *new_tab, unsigned long long j)
{
    unsigned a = new_tab[j].aux;
    if (a == 0)
        return;
    NkSlot s = old_tab[0xFFFFFFFFu - a];
    s.aux = 0; /* ghost counts of empty slots are not carried (C:1079) */
    new_tab[j] = s;
}

/* ---------------------------------------------------------------- decision (C:1641-1646, C:1988-1992) */

NK_HD int nk_keep_mate(unsigned high, unsigned total, float coverage)
{
#if NK_DEVICE_CODE
    float r = total > 0 ? __fdiv_rn((float)high, (float)total) : 0.0f;
#else
    float r = total > 0 ? (float)high / (float)total : 0.0f;
#endif
    return r < coverage;
}

/* one record: both mates must stay below the coverage (C:1646); a record with a non-ACGTN byte is
 * reported instead (the reference aborts on it, C:1445-1454) */
NK_HD void nk_decide_op(const NkRun &P, unsigned rec, int paired, float coverage, unsigned char *accept)
{
    unsigned r0 = paired ? 2 * rec : rec;
    bool bad = P.invalid[r0] || (paired && P.invalid[r0 + 1]);
    if (bad)
        nk_atomic_max32(&P.ctr->inv_max, NK_TMAX - rec);
    int keep = nk_keep_mate(P.high[r0], P.total[r0], coverage);
    if (paired)
        keep = keep && nk_keep_mate(P.high[r0 + 1], P.total[r0 + 1], coverage);
    if (P.reads[r0].len == 0)
        keep = 2; /* raw-text steps: the record failed the length gate on the device and does not count (C:1430-1443) */
    accept[rec] = (unsigned char)keep;
}

/* ---------------------------------------------------------------- table dump (C:354-385, C:1128-1136) */

/* where a dump reads its (k-mer, count) entries: a table in slot order (print_kmer_table), or the
 * merged arrays (all partitions, sorted by k-mer, counts summed: the TODO of C:25-26) */
struct NkDumpSrc
{
    const NkSlot *tab;
    const unsigned long long *keys;
    const long long *vals;
};

NK_HD void nk_dump_entry(const NkDumpSrc &s, unsigned long long i, unsigned long long &key, long long &val)
{
    if (s.tab)
    {
        NkSlot e = s.tab[i];
        key = e.key;
        val = (long long)e.count;
    }
    else
    {
        key = s.keys[i];
        val = s.vals[i];
    }
}

#define NK_DUMP_MAXLEN 54 /* 31 bases, tab, sign, 19 digits, newline, and one spare */

/* bytes of "KMER\tcount\n" (fprintf "%s\t%d\n", C:378); empty slots print nothing (C:372) */
NK_HD unsigned nk_dump_len(unsigned long long key, long long val, int k)
{
    if (key == 0)
        return 0;
    unsigned long long m = val < 0 ? 0ull - (unsigned long long)val : (unsigned long long)val;
    unsigned d = 1;
    while (m >= 10)
    {
        m /= 10;
        d++;
    }
    return (unsigned)k + 2u + d + (val < 0 ? 1u : 0u);
}

/* decode_kmer_plain (C:1128-1136): first base in the top bits, A C G T = 0 1 2 3 */
NK_HD void nk_dump_format(unsigned long long key, long long val, int k, char *out, unsigned len)
{
    for (int b = k - 1; b >= 0; b--)
    {
        out[b] = (char)((0x54474341u >> ((unsigned)(key & 3) * 8u)) & 0xFFu); /* "ACGT" */
        key >>= 2;
    }
    out[k] = '\t';
    out[len - 1] = '\n';
    unsigned long long m = val < 0 ? 0ull - (unsigned long long)val : (unsigned long long)val;
    unsigned at = len - 2;
    do
    {
        out[at--] = (char)('0' + (unsigned)(m % 10));
        m /= 10;
    } while (m);
    if (val < 0)
        out[at] = '-';
}

#define NK_DUMP_TILE 512 /* entries whose text one block assembles in shared memory per pass */

/* ---------------------------------------------------------------- raw record text on the device
 *
 * The host hands a step over as raw FASTQ/FASTA bytes: per partition one window of the forward file and one of
 * the reverse file holding the same number of complete records (process_thread_chunk_*'s read_line x4/x2 loop,
 * C:1605-1631, C:394-409).  The device finds the line ends, derives every record's sequence line, applies the
 * length gate (C:1430-1443), scores, and assembles the accepted records' output text (C:1649-1666, C:852-876)
 * so that the host only copies bytes in and write()s bytes out.  A window with a NUL byte or a line of 1024+
 * chars (where read_line would split differently) is reported back and that work goes through the host parser. */

struct NkRawWin /* device copy of one nkd_raw_segment */
{
    unsigned f_off, f_bytes, r_off, r_bytes; /* byte ranges in the step's raw buffer (r_bytes = 0: single-end) */
    unsigned n_records, part;
    unsigned f_line0, r_line0; /* index of the window's first line end in the step-wide list */
    unsigned rec0;             /* first record in the step's record numbering */
    unsigned out0;             /* first entry of the window in the output-length array: forward records, then reverse */
    unsigned pad0, pad1;
};

#define NK_RAW_TILE 4096u /* bytes of raw text whose line ends one block counts / lists per pass */
#define NK_RAW_NUL 1u     /* a NUL byte inside the step's text */
#define NK_RAW_LONG 2u    /* a line of NK_LINE_SPLIT or more chars */
#define NK_RAW_SHAPE 4u   /* a window does not end on the line end of its last record */

enum
{
    NK_EMIT_VERBATIM = 0, /* fq->fq, fa->fa: the record's lines, N->A in the sequence line (C:1426, C:1661-1665) */
    NK_EMIT_FASTA = 1,    /* fq->fa, paired: fastq_to_fasta (C:852-876) */
    NK_EMIT_NONE = 2      /* fq->fa, single-end: counted as printed, nothing written (C:1995-1999) */
};

struct NkRaw
{
    const unsigned char *raw;
    unsigned raw_bytes; /* multiple of 16 */
    const NkRawWin *wins;
    unsigned n_wins, n_records, stride /* mates per record */, per /* lines per record */;
    int k, emit_mode;
    int min_len; /* length gate: k when scoring (C:1430-1443), k + 1 when seeding (strlen > K, C:1347) */
    unsigned *tile; /* line ends per tile, then their exclusive scan; [n_tiles] = total */
    unsigned *nlpos;
    unsigned nlpos_cap;
    NkRead *reads;
    unsigned *nops, *opscan; /* operations per read and their exclusive scan ([n_reads] = total) */
    unsigned *t_out;         /* operations per window (= per partition of the step) */
    unsigned *flags;         /* NK_RAW_* */
    const unsigned char *accept;
    unsigned *outlen, *outoff; /* per (window, mate, record) and its exclusive scan */
    unsigned char *out;
    unsigned long long *summary; /* per window: fwd offset, fwd bytes, rev offset, rev bytes, processed, printed */
    const NkCounters *ctr;       /* inv_max = NK_TMAX - first record with a non-DNA byte (k_decide), 0 = none */
};

/* window of record i: the last window whose rec0 <= i */
NK_HD unsigned nk_raw_window_of(const NkRaw &R, unsigned i)
{
    unsigned lo = 0, hi = R.n_wins - 1;
    while (lo < hi)
    {
        unsigned mid = (lo + hi + 1) >> 1;
        if (R.wins[mid].rec0 <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

struct NkRawRec /* one mate of one record, as byte offsets into the raw buffer */
{
    unsigned start, end; /* first byte, one past its last line end */
    unsigned hdr_len, seq_off, seq_len;
    unsigned longest; /* longest of its lines */
};

NK_HD NkRawRec nk_raw_record(const NkRaw &R, const NkRawWin &w, unsigned r, int mate)
{
    const unsigned line0 = (mate ? w.r_line0 : w.f_line0) + R.per * r;
    const unsigned woff = mate ? w.r_off : w.f_off;
    NkRawRec x;
    x.start = r == 0 ? woff : R.nlpos[line0 - 1] + 1u;
    const unsigned e0 = R.nlpos[line0], e1 = R.nlpos[line0 + 1];
    x.hdr_len = e0 - x.start;
    x.seq_off = e0 + 1u;
    x.seq_len = e1 - e0 - 1u;
    x.longest = x.hdr_len > x.seq_len ? x.hdr_len : x.seq_len;
    unsigned prev = e1;
    for (unsigned l = 2; l < R.per; l++)
    {
        unsigned e = R.nlpos[line0 + l];
        if (e - prev - 1u > x.longest)
            x.longest = e - prev - 1u;
        prev = e;
    }
    x.end = prev + 1u;
    return x;
}

/* record i of the step: sequence lines of its mates, the length gate, operations per read */
NK_HD void nk_raw_record_op(const NkRaw &R, unsigned i)
{
    const unsigned wi = nk_raw_window_of(R, i);
    const NkRawWin w = R.wins[wi];
    const unsigned r = i - w.rec0;
    NkRawRec m[2];
    bool kept = true;
    unsigned flags = 0;
    for (unsigned s = 0; s < R.stride; s++)
    {
        m[s] = nk_raw_record(R, w, r, (int)s);
        if (m[s].longest >= NK_LINE_SPLIT)
            flags |= NK_RAW_LONG;
        if ((int)m[s].seq_len < R.min_len)
            kept = false; /* either mate shorter than K: the record vanishes, C:1430-1443 */
        if (r + 1 == w.n_records && m[s].end != (s ? w.r_off + w.r_bytes : w.f_off + w.f_bytes))
            flags |= NK_RAW_SHAPE;
    }
    if (flags)
        nk_red_or32(R.flags, flags);
    for (unsigned s = 0; s < R.stride; s++)
    {
        NkRead rd;
        rd.seq_off = m[s].seq_off;
        rd.op_base = 0;
        rd.len = (unsigned short)(kept && !(flags & NK_RAW_LONG) ? m[s].seq_len : 0u);
        rd.part = (unsigned short)w.part;
        rd.reserved = 0;
        R.reads[R.stride * i + s] = rd;
        R.nops[R.stride * i + s] = rd.len ? (unsigned)rd.len - (unsigned)R.k + 1u : 0u;
    }
}

/* seeding takes the first `limit` records that pass the length gate (C:1347-1356): flag them, rank them (scan), and
 * drop the ones beyond the limit */
NK_HD void nk_seed_flag_op(const NkRaw &R, unsigned i) { R.outlen[i] = R.nops[i] ? 1u : 0u; }
NK_HD void nk_seed_clip_op(const NkRaw &R, unsigned i, unsigned limit)
{
    if (R.outoff[i] >= limit && R.nops[i])
    {
        R.reads[i].len = 0;
        R.nops[i] = 0;
    }
}

/* read j: its first operation's number inside its partition's step; the window's first read reports the total */
NK_HD void nk_raw_opbase_op(const NkRaw &R, unsigned j)
{
    const unsigned i = j / R.stride;
    const unsigned wi = nk_raw_window_of(R, i);
    const NkRawWin w = R.wins[wi];
    const unsigned first = R.stride * w.rec0;
    R.reads[j].op_base = R.opscan[j] - R.opscan[first];
    if (j == first)
        R.t_out[wi] = R.opscan[first + R.stride * w.n_records] - R.opscan[first];
}

/* output bytes of entry e = (window, mate, record); returns the window through wi and whether the record counts */
NK_HD unsigned nk_emit_len_op(const NkRaw &R, unsigned e, unsigned &wi, unsigned &mate, unsigned &rec, int &counted,
                              int &printed)
{
    unsigned lo = 0, hi = R.n_wins - 1; /* last window whose out0 <= e */
    while (lo < hi)
    {
        unsigned mid = (lo + hi + 1) >> 1;
        if (R.wins[mid].out0 <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    wi = lo;
    const NkRawWin w = R.wins[wi];
    unsigned off = e - w.out0;
    mate = off >= w.n_records ? 1u : 0u;
    const unsigned r = off - mate * w.n_records;
    rec = w.rec0 + r;
    const unsigned char a = R.accept[rec];
    /* the reference stops at the first non-DNA record (C:1445-1454): later records of that partition do not count */
    const unsigned inv_rec = R.ctr->inv_max ? NK_TMAX - R.ctr->inv_max : NK_TMAX;
    const bool cut = inv_rec >= w.rec0 && inv_rec < w.rec0 + w.n_records && rec >= inv_rec;
    counted = (a != 2 && !cut) ? 1 : 0;
    printed = (a == 1 && !cut) ? 1 : 0;
    if (!printed || R.emit_mode == NK_EMIT_NONE)
        return 0;
    const NkRawRec x = nk_raw_record(R, w, r, (int)mate);
    if (R.emit_mode == NK_EMIT_VERBATIM)
        return x.end - x.start;
    /* ">" + header[1:] + "/1"|"/2" unless it already ends so + "\n" + sequence + "\n" */
    const unsigned char *h = R.raw + x.start;
    const unsigned char tag = mate ? '2' : '1';
    const bool has = x.hdr_len >= 2 && h[x.hdr_len - 2] == '/' && h[x.hdr_len - 1] == tag;
    return 1u + (x.hdr_len > 1 ? x.hdr_len - 1u : 0u) + (has ? 0u : 2u) + 1u + x.seq_len + 1u;
}

/* byte b of an entry's output text */
NK_HD unsigned char nk_emit_byte(const NkRaw &R, const NkRawRec &x, unsigned mate, unsigned b, unsigned len)
{
    if (R.emit_mode == NK_EMIT_VERBATIM)
    {
        const unsigned p = x.start + b;
        unsigned char c = R.raw[p];
        return (c == 'N' && p >= x.seq_off && p < x.seq_off + x.seq_len) ? (unsigned char)'A' : c;
    }
    const unsigned hl = x.hdr_len > 1 ? x.hdr_len - 1u : 0u;
    const unsigned seq_at = len - x.seq_len - 1u; /* the sequence and its line end close the text */
    if (b == 0)
        return '>';
    if (b <= hl)
        return R.raw[x.start + b];
    if (b >= seq_at)
    {
        if (b == len - 1u)
            return '\n';
        unsigned char c = R.raw[x.seq_off + (b - seq_at)];
        return c == 'N' ? (unsigned char)'A' : c;
    }
    if (b == seq_at - 1u)
        return '\n';
    return (b == hl + 1u) ? (unsigned char)'/' : (unsigned char)(mate ? '2' : '1'); /* the two suffix bytes */
}

/* per window: where its forward / reverse text lies in the step's output and what it counted */
NK_HD void nk_emit_summary_op(const NkRaw &R, unsigned wi)
{
    const NkRawWin w = R.wins[wi];
    unsigned long long *s = R.summary + 6u * wi;
    const unsigned f0 = R.outoff[w.out0], f1 = R.outoff[w.out0 + w.n_records];
    s[0] = f0;
    s[1] = f1 - f0;
    s[2] = f1;
    s[3] = R.stride == 2 ? R.outoff[w.out0 + 2u * w.n_records] - f1 : 0u;
}

#endif /* NK_CORE_H */
