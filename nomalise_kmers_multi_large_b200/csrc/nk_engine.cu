/*
 * nk_engine.cu -- sm_100a kernels and the CUDA backend of the device engine (nkd_* in include/nk_b200.h).
 *
 * Kernels (all integer, HBM/L2-latency bound; no tensor-core work exists on this path):
 *   k_probe    warp per read: coalesced 16-byte loads of the ASCII sequence, 2-bit packing in shared
 *              memory, per-lane window extraction (+ reverse-complement minimum), then the table probe
 *              of nk_core.h: one 16-byte L2-only gather per visited slot, fire-and-forget REDs on
 *              saturated counters                       [encode_kmer_plain C:1118, get_canonical_kmer
 *                                                        C:1175, is_valid_sequence_* C:1404-1457,
 *                                                        store_kmer C:929-1053, sequence_to_hash C:1459]
 *   k_open     thread per deferred operation (met a slot that was empty at step start)
 *   k_apply / k_classify / k_rank   pending increments, threshold classification, time ranking
 *   k_commit / k_untag              store claimed keys / forget an abandoned run
 *   k_rehash_place / k_rehash_fill  exact table growth                  [expand_local_hash_table C:1055]
 *   k_decide   float32 ratio test per record                                          [C:1641-1646]
 *   k_dump_measure / k_dump_text / k_dump_pairs   -P table dump formatted on the device, compaction
 *                                                 for the merged table        [print_kmer_table C:354]
 * Library calls: cub::DeviceRadixSort for the (rare) time-ordered slow path; cub scan / sort /
 * reduce-by-key in the table dump and the merged table (not on the scoring path).
 */
#include <cuda_runtime.h>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda/std/functional>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "nk_core.h"
#include "../../include/nk_b200.h"

#define NK_WARPS 8
#define NK_THREADS (NK_WARPS * 32)
#define NK_WORDS 68 /* 64 packed words (1024 bases) + zero padding read by the last windows */

/* 16 ASCII bases -> 32 bits, first base in the top bits.  A0 C1 G2 T3; N packs as A (C:1406); bytes
 * beyond nb are ignored.  bad collects bytes outside ACGTN (C:1144-1158). */
__device__ __forceinline__ unsigned nk_pack16(uint4 v, int nb, unsigned &bad)
{
    unsigned in[4] = {v.x, v.y, v.z, v.w};
    unsigned out = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        int n = nb - 4 * j;
        unsigned mask = n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << (8 * n)) - 1u));
        unsigned x = in[j] & mask;
        unsigned ok = __vcmpeq4(x, 0x41414141u) | __vcmpeq4(x, 0x43434343u) | __vcmpeq4(x, 0x47474747u) |
                      __vcmpeq4(x, 0x54545454u) | __vcmpeq4(x, 0x4E4E4E4Eu);
        bad |= (~ok) & mask;
        unsigned codes = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
        unsigned y = __byte_perm(codes, 0, 0x0123);
        unsigned p = (y | (y >> 6) | (y >> 12) | (y >> 18)) & 0xFFu;
        out = (out << 8) | p;
    }
    return out;
}

__device__ __forceinline__ unsigned long long nk_window_key_packed(const unsigned *words, int w, int k, int canonical)
{
    int wi = w >> 4, s = (w & 15) * 2;
    unsigned long long hi = ((unsigned long long)words[wi] << 32) | words[wi + 1];
    unsigned long long v = s ? ((hi << s) | ((unsigned long long)words[wi + 2] >> (32 - s))) : hi;
    unsigned long long x = v >> (64 - 2 * k);
    if (canonical)
    {
        unsigned long long r = nk_revcomp(x, k);
        x = r < x ? r : x;
    }
    return x;
}

__device__ __forceinline__ void nk_cur_init(NkWarpCur *wc, unsigned lane)
{
    if (lane < NK_NLISTS)
    {
        wc->base[lane] = wc->base2[lane] = 0;
        wc->cap[lane] = wc->cap2[lane] = 0;
        wc->used[lane] = 0;
    }
    __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void nk_probe_body(NkRun P)
{
    P.mode = MODE; /* compile-time mode: the per-operation code specialises */
    __shared__ unsigned s_words[NK_WARPS][NK_WORDS];
    __shared__ NkWarpCur s_cur[NK_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned *words = s_words[warp];
    P.wcur = &s_cur[warp];
    nk_cur_init(P.wcur, lane);
    auto pend_hole = [&](unsigned i) { P.pend[i].slot = NK_HOLE; };
    auto open_hole = [&](unsigned i) { P.open[i].flags = NK_HOLE; };
    const unsigned nwarps = gridDim.x * NK_WARPS;
    const bool one_part = (MODE == NK_MODE_SEED || MODE == NK_MODE_KEYS);
    /* per-partition counters are accumulated in registers and flushed when the partition changes:
     * reads are partition-major, so a warp flushes a handful of times per launch */
    unsigned acc_part = 0xFFFFFFFFu, acc_real = 0, acc_touch = 0, acc_hot = 0;
    auto flush = [&]() {
        if (lane == 0 && acc_part != 0xFFFFFFFFu)
        {
            if (acc_real)
                atomicAdd(&P.ctr->real_ops[acc_part], (unsigned long long)acc_real);
            if (acc_touch)
            {
                atomicAdd(&P.ctr->touches[acc_part], (unsigned long long)acc_touch);
                atomicAdd(&P.ctr->probe_touches, (unsigned long long)acc_touch);
            }
            if (acc_hot)
                atomicAdd(&P.ctr->hot_hits, (unsigned long long)acc_hot);
        }
        acc_real = acc_touch = acc_hot = 0;
    };
    for (unsigned r = blockIdx.x * NK_WARPS + warp; r < P.n_reads; r += nwarps)
    {
        const uint4 rraw = __ldg(reinterpret_cast<const uint4 *>(P.reads + r));
        const unsigned seq_off = rraw.x, op_base = rraw.y, len = rraw.z & 0xFFFFu;
        const unsigned part = one_part ? 0u : (rraw.z >> 16);
        if (len == 0)
            continue; /* raw-text steps: a record that failed the length gate keeps its place but has no windows */
        const NkPart pd = P.parts[part];
        const int nwin = (int)len - P.k + 1;
        if (op_base + (unsigned)nwin <= pd.lo || op_base >= pd.hi)
            continue;
        if (part != acc_part)
        {
            flush();
            acc_part = part;
        }
        if (P.record && (MODE == NK_MODE_SCORE || MODE == NK_MODE_SEED))
        {
            if (MODE == NK_MODE_SCORE)
                nk_chunk_rotate(P, NK_LIST_PEND, &P.ctr->n_pend, P.pend_cap);
            nk_chunk_rotate(P, NK_LIST_OPEN, &P.ctr->n_open, P.open_cap);
        }
        const unsigned nchunks = (len + 15u) >> 4;
        /* parsed steps stage every sequence at a 16-byte boundary; in raw-text steps it starts wherever its line
         * starts, and a lane assembles its 16 bases from the two aligned chunks they straddle (shift is warp-uniform) */
        const unsigned shift = seq_off & 15u;
        const uint4 *chunks = reinterpret_cast<const uint4 *>(P.seq + (seq_off - shift));
        unsigned bad = 0;
        for (unsigned c = lane; c < nchunks + 2u; c += 32u)
        {
            unsigned w = 0;
            if (c < nchunks)
            {
                uint4 v = __ldg(chunks + c);
                if (shift)
                {
                    const uint4 u = __ldg(chunks + c + 1);
                    const unsigned sh = (shift & 3u) * 8u;
                    unsigned a0, a1, a2, a3, a4;
                    switch (shift >> 2)
                    {
                    case 0:
                        a0 = v.x, a1 = v.y, a2 = v.z, a3 = v.w, a4 = u.x;
                        break;
                    case 1:
                        a0 = v.y, a1 = v.z, a2 = v.w, a3 = u.x, a4 = u.y;
                        break;
                    case 2:
                        a0 = v.z, a1 = v.w, a2 = u.x, a3 = u.y, a4 = u.z;
                        break;
                    default:
                        a0 = v.w, a1 = u.x, a2 = u.y, a3 = u.z, a4 = u.w;
                    }
                    v.x = __funnelshift_r(a0, a1, sh);
                    v.y = __funnelshift_r(a1, a2, sh);
                    v.z = __funnelshift_r(a2, a3, sh);
                    v.w = __funnelshift_r(a3, a4, sh);
                }
                w = nk_pack16(v, (int)len - 16 * (int)c, bad);
            }
            words[c] = w;
        }
        __syncwarp();
        if (MODE != NK_MODE_COUNT && __any_sync(0xFFFFFFFFu, bad != 0) && lane == 0)
            P.invalid[r] = 1;
        unsigned n_real = 0, touches = 0, hot = 0;
        int high = 0;
        for (int w0 = 0; w0 < nwin; w0 += 32)
        {
            const int w = w0 + (int)lane;
            const unsigned t = op_base + (unsigned)w;
            const bool live = w < nwin && t >= pd.lo && t < pd.hi;
            if (!live)
                continue;
            const unsigned long long key = nk_window_key_packed(words, w, P.k, P.canonical);
            if (MODE == NK_MODE_KEYS)
            {
                P.keys_out[t] = key;
                continue;
            }
            if (key == 0) /* all-A window (or all-T under --canonical): ignored entirely, C:1483 */
                continue;
            n_real++;
            if (MODE != NK_MODE_COUNT)
                touches += nk_probe_op(P, pd, part, key, t, r, high, hot);
        }
        __syncwarp();
        if (MODE != NK_MODE_KEYS)
        {
            n_real = __reduce_add_sync(0xFFFFFFFFu, n_real);
            acc_real += n_real;
            if (MODE != NK_MODE_COUNT)
            {
                acc_touch += __reduce_add_sync(0xFFFFFFFFu, touches);
                if (MODE == NK_MODE_SCORE && P.hot)
                    acc_hot += __reduce_add_sync(0xFFFFFFFFu, hot);
                high = __reduce_add_sync(0xFFFFFFFFu, high);
                if (lane == 0)
                {
                    if (n_real)
                        atomicAdd(&P.total[r], (unsigned)(P.delta * (int)n_real));
                    if (high)
                        atomicAdd(&P.high[r], (unsigned)high);
                }
            }
        }
        __syncwarp();
    }
    flush();
    if (P.record && (MODE == NK_MODE_SCORE || MODE == NK_MODE_SEED))
    {
        nk_chunk_close(P, NK_LIST_PEND, pend_hole);
        nk_chunk_close(P, NK_LIST_OPEN, open_hole);
    }
}

/* distinct symbols per mode so that profiles separate scoring from seeding */
__global__ void __launch_bounds__(NK_THREADS) k_probe_score(const NkRun P) { nk_probe_body<NK_MODE_SCORE>(P); }
__global__ void __launch_bounds__(NK_THREADS) k_probe_seed(const NkRun P) { nk_probe_body<NK_MODE_SEED>(P); }
__global__ void __launch_bounds__(NK_THREADS) k_probe_count(const NkRun P) { nk_probe_body<NK_MODE_COUNT>(P); }
__global__ void __launch_bounds__(NK_THREADS) k_probe_keys(const NkRun P) { nk_probe_body<NK_MODE_KEYS>(P); }

__global__ void __launch_bounds__(256) k_open(NkRun P)
{
    __shared__ NkWarpCur s_cur[8];
    __shared__ unsigned s_touch[NK_MAX_PARTITIONS], s_claims[NK_MAX_PARTITIONS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    P.wcur = &s_cur[warp];
    nk_cur_init(P.wcur, lane);
    for (unsigned i = threadIdx.x; i < NK_MAX_PARTITIONS; i += blockDim.x)
        s_touch[i] = s_claims[i] = 0;
    __syncthreads();
    auto pend_hole = [&](unsigned i) { P.pend[i].slot = NK_HOLE; };
    auto spec_hole = [&](unsigned i) { P.spec[i].slot = NK_HOLE; };
    auto claim_hole = [&](unsigned i) { P.claim[i].slot = NK_HOLE; };
    const unsigned n = min(P.ctr->n_open, P.open_cap);
    for (unsigned base = (blockIdx.x * 8u + warp) * 32u; base < n; base += gridDim.x * 256u)
    {
        if (P.record)
        {
            if (P.mode == NK_MODE_SCORE)
            {
                nk_chunk_rotate(P, NK_LIST_PEND, &P.ctr->n_pend, P.pend_cap);
                nk_chunk_rotate(P, NK_LIST_SPEC, &P.ctr->n_spec, P.spec_cap);
            }
            nk_chunk_rotate(P, NK_LIST_CLAIM, &P.ctr->n_claim, P.claim_cap);
        }
        const unsigned i = base + lane;
        if (i < n)
        {
            int high = 0, claimed = 0;
            unsigned touches = nk_open_op(P, i, high, claimed);
            if (touches | (unsigned)claimed | (unsigned)high)
            {
                const unsigned part = P.open[i].part;
                if (touches)
                    atomicAdd(&s_touch[part], touches);
                if (claimed)
                    atomicAdd(&s_claims[part], 1u);
                if (high)
                    atomicAdd(&P.high[P.open[i].read], (unsigned)high);
            }
        }
        __syncwarp();
    }
    if (P.record)
    {
        nk_chunk_close(P, NK_LIST_PEND, pend_hole);
        nk_chunk_close(P, NK_LIST_SPEC, spec_hole);
        nk_chunk_close(P, NK_LIST_CLAIM, claim_hole);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < NK_MAX_PARTITIONS; i += blockDim.x)
    {
        if (s_touch[i])
            atomicAdd(&P.ctr->touches[i], (unsigned long long)s_touch[i]);
        if (s_claims[i])
            atomicAdd(&P.ctr->claims[i], s_claims[i]);
    }
}

__global__ void __launch_bounds__(256) k_prepare_claims(const NkRun P)
{
    const unsigned n = min(P.ctr->n_open, P.open_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        nk_prepare_claim_op(P, i);
}
__global__ void __launch_bounds__(256) k_apply(const NkRun P, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        nk_apply_op(P, i);
}
/* Classification appends only at converged points, so the warp aggregates with a ballot: one global atomic
 * per chunk of NK_SLOW_CHUNK entries, no shared cursor, and warps with nothing to emit reserve nothing. */
#define NK_SLOW_CHUNK 256u
template <bool CLAIMED>
__device__ __forceinline__ void nk_classify_body(NkRun &P, unsigned n)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned cbase = 0, cused = 0, ccap = 0; /* warp-uniform */
    for (unsigned base = (blockIdx.x * 8u + warp) * 32u; base < n; base += gridDim.x * 256u)
    {
        const unsigned i = base + lane;
        NkPend r;
        unsigned long long gslot = 0;
        int x = 0;
        bool emit = false;
        if (i < n)
            emit = CLAIMED ? nk_classify_claimed_op(P, i, r, gslot, x) : nk_classify_op(P, i, r, gslot, x);
        const unsigned votes = __ballot_sync(0xFFFFFFFFu, emit);
        const unsigned cnt = __popc(votes);
        if (cnt == 0)
            continue;
        if (cused + cnt > ccap)
        {
            for (unsigned h = cused + lane; h < ccap; h += 32)
                P.slow_key[cbase + h] = ~0ull; /* tail of the old chunk: holes */
            unsigned b = 0;
            if (lane == 0)
                b = atomicAdd(&P.ctr->n_slow, NK_SLOW_CHUNK);
            b = __shfl_sync(0xFFFFFFFFu, b, 0);
            cbase = b;
            cused = 0;
            ccap = b >= P.slow_cap ? 0u : (P.slow_cap - b < NK_SLOW_CHUNK ? P.slow_cap - b : NK_SLOW_CHUNK);
            if (ccap < cnt && lane == 0)
                atomicOr(&P.ctr->overflow, NK_OVF_SLOW);
        }
        if (emit)
        {
            unsigned my = cused + __popc(votes & ((1u << lane) - 1u));
            if (my < ccap)
                nk_slow_write(P, cbase + my, gslot, r, x, CLAIMED ? 0 : 1);
        }
        cused += cnt;
    }
    for (unsigned h = (cused < ccap ? cused : ccap) + lane; h < ccap; h += 32)
        P.slow_key[cbase + h] = ~0ull;
}
__global__ void __launch_bounds__(256) k_classify(NkRun P, unsigned n) { nk_classify_body<false>(P, n); }
__global__ void __launch_bounds__(256) k_classify_claimed(NkRun P, unsigned n) { nk_classify_body<true>(P, n); }
__global__ void __launch_bounds__(256) k_rank(const NkRun P, const unsigned long long *keys, const unsigned long long *vals,
                                             unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        nk_rank_op(P, keys, vals, n, i);
}
__global__ void __launch_bounds__(256) k_commit(const NkRun P, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        nk_commit_op(P, i);
}
__global__ void __launch_bounds__(256) k_untag(const NkRun P, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        nk_untag_op(P, i);
}
__global__ void __launch_bounds__(256) k_rehash_place(const NkSlot *old_tab, unsigned long long cap, NkSlot *nt,
                                                     unsigned long long ncap, unsigned long long nmagic)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < cap;
         i += (unsigned long long)gridDim.x * blockDim.x)
        nk_rehash_place_op(old_tab, i, nt, ncap, nmagic);
}
__global__ void __launch_bounds__(256) k_rehash_fill(const NkSlot *old_tab, NkSlot *nt, unsigned long long ncap)
{
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < ncap;
         j += (unsigned long long)gridDim.x * blockDim.x)
        nk_rehash_fill_op(old_tab, nt, j);
}
__global__ void __launch_bounds__(256) k_decide(const NkRun P, unsigned n_records, int paired, float coverage,
                                               unsigned char *accept)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_records; i += gridDim.x * blockDim.x)
        nk_decide_op(P, i, paired, coverage, accept);
}

/* ------------------------------------------------------------------ table dump (C:354-385) */

/* pass 1: text bytes (or stored entries) of every tile of NK_DUMP_TILE consecutive entries */
__global__ void __launch_bounds__(256) k_dump_measure(const NkDumpSrc src, unsigned long long lo, unsigned long long n,
                                                      int k, int text, unsigned long long *tile_units)
{
    typedef cub::BlockReduce<unsigned, 256> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        unsigned u = 0;
#pragma unroll
        for (int j = 0; j < NK_DUMP_TILE / 256; j++)
        {
            unsigned long long i = tile * NK_DUMP_TILE + threadIdx.x * (NK_DUMP_TILE / 256) + j;
            if (i < n)
            {
                unsigned long long key;
                long long val;
                nk_dump_entry(src, lo + i, key, val);
                u += text ? nk_dump_len(key, val, k) : (key != 0);
            }
        }
        unsigned sum = Reduce(tmp).Sum(u);
        if (threadIdx.x == 0)
            tile_units[tile] = sum;
        __syncthreads();
    }
}

/* pass 2: every thread formats its entries into the block's shared buffer at its scanned offset; the
 * block then copies the tile's text to its place in the output with coalesced stores */
__global__ void __launch_bounds__(256) k_dump_text(const NkDumpSrc src, unsigned long long lo, unsigned long long n, int k,
                                                   const unsigned long long *tile_off, char *text)
{
    typedef cub::BlockScan<unsigned, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ char buf[NK_DUMP_TILE * NK_DUMP_MAXLEN];
    constexpr int PER = NK_DUMP_TILE / 256;
    unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        unsigned long long key[PER];
        long long val[PER];
        unsigned len[PER], mine = 0;
#pragma unroll
        for (int j = 0; j < PER; j++)
        {
            unsigned long long i = tile * NK_DUMP_TILE + threadIdx.x * PER + j;
            key[j] = 0;
            val[j] = 0;
            if (i < n)
                nk_dump_entry(src, lo + i, key[j], val[j]);
            len[j] = nk_dump_len(key[j], val[j], k);
            mine += len[j];
        }
        unsigned at;
        Scan(tmp).ExclusiveSum(mine, at);
#pragma unroll
        for (int j = 0; j < PER; j++)
        {
            if (len[j])
                nk_dump_format(key[j], val[j], k, buf + at, len[j]);
            at += len[j];
        }
        __syncthreads();
        unsigned long long o = tile_off[tile];
        unsigned total = (unsigned)(tile_off[tile + 1] - o);
        for (unsigned b = threadIdx.x; b < total; b += 256)
            text[o + b] = buf[b];
        __syncthreads();
    }
}

/* pass 2 of a compaction: stored (key, count) pairs in slot order */
__global__ void __launch_bounds__(256) k_dump_pairs(const NkDumpSrc src, unsigned long long lo, unsigned long long n,
                                                    const unsigned long long *tile_off, unsigned long long *keys_out,
                                                    long long *vals_out)
{
    typedef cub::BlockScan<unsigned, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    constexpr int PER = NK_DUMP_TILE / 256;
    unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        unsigned long long key[PER];
        long long val[PER];
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < PER; j++)
        {
            unsigned long long i = tile * NK_DUMP_TILE + threadIdx.x * PER + j;
            key[j] = 0;
            val[j] = 0;
            if (i < n)
                nk_dump_entry(src, lo + i, key[j], val[j]);
            mine += key[j] != 0;
        }
        unsigned at;
        Scan(tmp).ExclusiveSum(mine, at);
        unsigned long long o = tile_off[tile] + at;
#pragma unroll
        for (int j = 0; j < PER; j++)
            if (key[j])
            {
                keys_out[o] = key[j];
                vals_out[o] = val[j];
                o++;
            }
        __syncthreads();
    }
}


/* ------------------------------------------------------------------ raw record text (C:1605-1631, C:394-409 on the device) */

/* newline / NUL bits of 16 bytes: bit i of the result is byte i */
__device__ __forceinline__ unsigned nk_eq_mask16(uint4 v, unsigned pattern)
{
    unsigned in[4] = {v.x, v.y, v.z, v.w}, m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        unsigned t = __vcmpeq4(in[j], pattern) & 0x08040201u;
        m |= ((t | (t >> 8) | (t >> 16) | (t >> 24)) & 0xFu) << (4 * j);
    }
    return m;
}

/* pass 1: line ends per tile of NK_RAW_TILE bytes; NUL bytes anywhere are reported */
__global__ void __launch_bounds__(256) k_raw_count(const NkRaw R, unsigned n_tiles)
{
    typedef cub::BlockReduce<unsigned, 256> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const unsigned at = tile * NK_RAW_TILE + threadIdx.x * 16u;
        unsigned c = 0;
        if (at < R.raw_bytes)
        {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(R.raw + at));
            c = __popc(nk_eq_mask16(v, 0x0A0A0A0Au));
            if (nk_eq_mask16(v, 0u))
                atomicOr(R.flags, NK_RAW_NUL);
        }
        unsigned sum = Reduce(tmp).Sum(c);
        if (threadIdx.x == 0)
            R.tile[tile] = sum;
        __syncthreads();
    }
}

/* pass 2: the positions, in order, at the tile's scanned offset */
__global__ void __launch_bounds__(256) k_raw_positions(const NkRaw R, unsigned n_tiles)
{
    typedef cub::BlockScan<unsigned, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        R.flags[1] = R.tile[n_tiles];
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const unsigned at = tile * NK_RAW_TILE + threadIdx.x * 16u;
        unsigned m = 0;
        if (at < R.raw_bytes)
            m = nk_eq_mask16(__ldg(reinterpret_cast<const uint4 *>(R.raw + at)), 0x0A0A0A0Au);
        unsigned before;
        Scan(tmp).ExclusiveSum((unsigned)__popc(m), before);
        unsigned idx = R.tile[tile] + before;
        while (m)
        {
            unsigned b = __ffs(m) - 1u;
            m &= m - 1u;
            if (idx < R.nlpos_cap)
                R.nlpos[idx] = at + b;
            idx++;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_raw_records(const NkRaw R)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < R.n_records; i += gridDim.x * blockDim.x)
        nk_raw_record_op(R, i);
}
__global__ void __launch_bounds__(256) k_seed_flag(const NkRaw R)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < R.n_records; i += gridDim.x * blockDim.x)
        nk_seed_flag_op(R, i);
}
__global__ void __launch_bounds__(256) k_seed_clip(const NkRaw R, unsigned limit)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < R.n_records; i += gridDim.x * blockDim.x)
        nk_seed_clip_op(R, i, limit);
}
__global__ void __launch_bounds__(256) k_raw_opbase(const NkRaw R)
{
    const unsigned n = R.n_records * R.stride;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        nk_raw_opbase_op(R, j);
}

/* output bytes per (window, mate, record) and the windows' processed / printed counters (C:1667, C:1672) */
__global__ void __launch_bounds__(256) k_emit_measure(const NkRaw R, unsigned n_out)
{
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned base = (blockIdx.x * 8u + (threadIdx.x >> 5)) * 32u; base < n_out; base += gridDim.x * 256u)
    {
        const unsigned e = base + lane;
        unsigned wi = 0xFFFFFFFFu, mate = 0, rec = 0, len = 0;
        int counted = 0, printed = 0;
        if (e < n_out)
        {
            len = nk_emit_len_op(R, e, wi, mate, rec, counted, printed);
            R.outlen[e] = len;
            if (mate)
                counted = printed = 0; /* a record counts once */
        }
        const unsigned w0 = __shfl_sync(0xFFFFFFFFu, wi, 0);
        if (__all_sync(0xFFFFFFFFu, wi == w0))
        {
            const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)counted), p = __reduce_add_sync(0xFFFFFFFFu, (unsigned)printed);
            if (lane == 0 && w0 != 0xFFFFFFFFu)
            {
                if (c)
                    atomicAdd(&R.summary[6u * w0 + 4u], (unsigned long long)c);
                if (p)
                    atomicAdd(&R.summary[6u * w0 + 5u], (unsigned long long)p);
            }
        }
        else if (wi != 0xFFFFFFFFu)
        {
            if (counted)
                atomicAdd(&R.summary[6u * wi + 4u], 1ull);
            if (printed)
                atomicAdd(&R.summary[6u * wi + 5u], 1ull);
        }
    }
}

/* warp per entry: the accepted record's text at its scanned offset */
__global__ void __launch_bounds__(256) k_emit_copy(const NkRaw R, unsigned n_out)
{
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned e = blockIdx.x * 8u + (threadIdx.x >> 5); e < n_out; e += gridDim.x * 8u)
    {
        const unsigned o = R.outoff[e], len = R.outoff[e + 1] - o;
        if (len == 0)
            continue;
        unsigned lo = 0, hi = R.n_wins - 1;
        while (lo < hi)
        {
            unsigned mid = (lo + hi + 1) >> 1;
            if (R.wins[mid].out0 <= e)
                lo = mid;
            else
                hi = mid - 1;
        }
        const NkRawWin w = R.wins[lo];
        const unsigned off = e - w.out0, mate = off >= w.n_records ? 1u : 0u;
        const NkRawRec x = nk_raw_record(R, w, off - mate * w.n_records, (int)mate);
        for (unsigned b = lane; b < len; b += 32u)
            R.out[o + b] = nk_emit_byte(R, x, mate, b, len);
    }
}
__global__ void k_emit_summary(const NkRaw R)
{
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < R.n_wins; w += gridDim.x * blockDim.x)
        nk_emit_summary_op(R, w);
}

__global__ void __launch_bounds__(256) k_hot_flush(const NkRun P)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i <= P.hot_mask; i += gridDim.x * blockDim.x)
        nk_hot_flush_op(P, i);
}
__global__ void __launch_bounds__(256) k_hot_clear(const NkRun P, unsigned part1)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i <= P.hot_mask; i += gridDim.x * blockDim.x)
        nk_hot_clear_op(P, i, part1);
}

/* Clearing and small control blocks without the copy engines: while the next step's text is on its way (one long H2D
 * transfer on the upload stream) a cudaMemsetAsync or a small cudaMemcpyAsync of this step would wait behind it. */
__global__ void __launch_bounds__(256) k_zero(uint4 *p, size_t n16, unsigned char *tail, unsigned n_tail)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0 && threadIdx.x < n_tail)
        tail[threadIdx.x] = 0;
}
struct NkBlob
{
    unsigned w[768]; /* 3 KB: kernel parameters may hold 4 KB */
};
__global__ void __launch_bounds__(256) k_put_small(unsigned *dst, const NkBlob blob, unsigned n_words)
{
    for (unsigned i = threadIdx.x; i < n_words; i += blockDim.x)
        dst[i] = blob.w[i];
}

/* ------------------------------------------------------------------ backend */

/* boolean environment switches: unset, empty and "0" mean off (the same rule as nk_env_on in nk_host.c) */
static bool nk_env_flag(const char *name)
{
    const char *s = getenv(name);
    return s && *s && strcmp(s, "0") != 0;
}

struct CudaBackend
{
    int dev = -1, sms = 148;
    unsigned long long launches = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t sync_ev = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    std::string cuda_err;
    struct Timer
    {
        std::vector<cudaEvent_t> ev; /* start/stop pairs */
        size_t used = 0;
    } timers[11];
    /* the accepted records' text leaves on its own stream, so the engine's next step starts without waiting for it */
    cudaStream_t copy_stream = nullptr, up_stream = nullptr;
    cudaEvent_t up_done = nullptr;
    cudaEvent_t copy_done[NKD_FETCH_SLOTS] = {}, emit_ready = nullptr, last_copy = nullptr;

    bool ok(cudaError_t e, const char *what)
    {
        if (e == cudaSuccess)
            return true;
        if (cuda_err.empty())
            cuda_err = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    }
    /* the current device is per host thread: every API entry selects this engine's GPU */
    void enter()
    {
        if (dev >= 0)
            cudaSetDevice(dev);
    }
    bool failed(std::string &msg)
    {
        ok(cudaGetLastError(), "kernel launch");
        if (cuda_err.empty())
            return false;
        msg = cuda_err;
        return true;
    }
    int init(int device, std::string &err)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        {
            err = "no CUDA device: the B200 engine has no CPU fallback";
            return NK_ENODEVICE;
        }
        if (device < 0 || device >= n)
        {
            err = "CUDA device ordinal out of range";
            return NK_ENODEVICE;
        }
        dev = device;
        if (!ok(cudaSetDevice(dev), "cudaSetDevice") ||
            !ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        {
            err = cuda_err;
            return NK_ENODEVICE;
        }
        if (!nk_env_flag("NKB200_SPIN_SYNC") &&
            cudaEventCreateWithFlags(&sync_ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)
        {
            cudaGetLastError();
            sync_ev = nullptr;
        }
        ok(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < NKD_FETCH_SLOTS; i++)
            ok(cudaEventCreateWithFlags(&copy_done[i], cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreate");
        ok(cudaEventCreateWithFlags(&emit_ready, cudaEventDisableTiming), "cudaEventCreate");
        ok(cudaStreamCreateWithFlags(&up_stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ok(cudaEventCreateWithFlags(&up_done, cudaEventDisableTiming), "cudaEventCreate");
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (getenv("NKB200_COPY_PIECE_MB") && atoi(getenv("NKB200_COPY_PIECE_MB")) > 0)
            piece = (size_t)atoi(getenv("NKB200_COPY_PIECE_MB")) << 20;
        epoch(dev);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
        {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        return NK_OK;
    }
    void shutdown()
    {
        if (dev < 0)
            return;
        cudaSetDevice(dev);
        for (auto &t : timers)
            for (auto e : t.ev)
                cudaEventDestroy(e);
        if (sync_ev)
            cudaEventDestroy(sync_ev);
        sync_ev = nullptr;
        if (sort_tmp)
            cudaFreeAsync(sort_tmp, stream);
        if (scan_tmp)
            cudaFreeAsync(scan_tmp, stream);
        if (copy_stream)
        {
            cudaStreamSynchronize(copy_stream);
            cudaStreamDestroy(copy_stream);
        }
        copy_stream = nullptr;
        for (int i = 0; i < NKD_FETCH_SLOTS; i++)
            if (copy_done[i])
                cudaEventDestroy(copy_done[i]);
        if (emit_ready)
            cudaEventDestroy(emit_ready);
        if (up_stream)
        {
            cudaStreamSynchronize(up_stream);
            cudaStreamDestroy(up_stream);
        }
        up_stream = nullptr;
        if (up_done)
            cudaEventDestroy(up_done);
        if (stream)
        {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        stream = nullptr;
    }
    /* stream-ordered allocation from the device's pool, which keeps freed memory: table growth frees and
     * allocates GBs, and cudaFree/cudaMalloc were seen to take 100-500 ms each there */
    void *alloc(size_t n)
    {
        cudaSetDevice(dev);
        void *p = nullptr;
        if (cudaMallocAsync(&p, n ? n : 16, stream) != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void release(void *p)
    {
        if (p)
        {
            cudaSetDevice(dev);
            cudaFreeAsync(p, stream);
        }
    }
    /* page-locked host memory is mapped into the device address space (UVA): the probe kernel can read the
     * staged sequence bytes over PCIe with coalesced 16-byte loads instead of waiting for a bulk copy */
    const void *device_view_of_host(const void *p)
    {
        if (nk_env_flag("NKB200_NO_ZEROCOPY"))
            return nullptr;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    }
    void zero(void *p, size_t n)
    {
        if (!n)
            return;
        unsigned char *b = (unsigned char *)p;
        size_t head = (16 - ((uintptr_t)b & 15)) & 15;
        if (head > n)
            head = n;
        if (head) /* allocations are 256-byte aligned: only interior pointers get here */
            ok(cudaMemsetAsync(b, 0, head, stream), "cudaMemsetAsync");
        size_t n16 = (n - head) / 16, tail = (n - head) & 15;
        unsigned g = grid_for(n16 ? n16 : 1, 256 * 8);
        k_zero<<<g, 256, 0, stream>>>((uint4 *)(b + head), n16, b + head + n16 * 16, (unsigned)tail);
        launches++;
    }
    /* n bytes (a multiple of 4, 4-byte aligned destination) as kernel parameters */
    void put_small(void *d, const void *h, size_t n)
    {
        if ((n & 3) || ((uintptr_t)d & 3))
        {
            h2d(d, h, n);
            return;
        }
        const size_t words = n / 4;
        for (size_t at = 0; at < words; at += 768)
        {
            NkBlob b;
            size_t m = words - at < 768 ? words - at : 768;
            memcpy(b.w, (const unsigned *)h + at, m * 4);
            k_put_small<<<1, 256, 0, stream>>>((unsigned *)d + at, b, (unsigned)m);
            launches++;
        }
    }
    void h2d(void *d, const void *h, size_t n) { ok(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, stream), "H2D copy"); }
    void d2h(void *h, const void *d, size_t n) { ok(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, stream), "D2H copy"); }
    void d2d(void *d, const void *s, size_t n) { ok(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, stream), "D2D copy"); }
    /* Waiting on a blocking-sync event puts the host thread to sleep instead of spinning in the driver: the
     * engines' threads wait most of the time, and the cores they would burn are the ones the host stages
     * (record indexing, output writing) are short of.  NKB200_SPIN_SYNC=1 restores the spinning wait. */
    void sync()
    {
        if (!sync_ev)
            ok(cudaStreamSynchronize(stream), "stream synchronize");
        else if (ok(cudaEventRecord(sync_ev, stream), "event record"))
            ok(cudaEventSynchronize(sync_ev), "event synchronize");
    }

    /* entries a warp reserves per global atomic: large enough to make the atomics negligible, small enough
     * that the holes of (SMs x 8 x 8) warps stay a small fraction of the list */
    static unsigned env_chunk(const char *name, unsigned dflt)
    {
        const char *e = getenv(name);
        return e && atoi(e) >= 32 ? (unsigned)atoi(e) : dflt;
    }
    unsigned min_list_entries() const { return (unsigned)sms * 8u * 8u * 2u * 32u * 3u; }
    /* entries the classification kernels may leave as holes: every warp of k_classify and k_classify_claimed can
     * abandon the tail of one NK_SLOW_CHUNK reservation */
    unsigned slow_hole_margin() const { return 2u * (unsigned)sms * 8u * 8u * NK_SLOW_CHUNK; }
    void chunk_sizes(unsigned *c, unsigned pend_cap, unsigned open_cap, unsigned claim_cap, unsigned slow_cap,
                     unsigned spec_cap)
    {
        unsigned warps = (unsigned)sms * 8u * 8u;
        auto pick = [&](unsigned cap, unsigned mx) {
            unsigned v = cap / (4u * warps);
            return v < 32u ? 32u : (v > mx ? mx : v);
        };
        c[NK_LIST_PEND] = pick(pend_cap, env_chunk("NKB200_CHUNK_PEND", 384));
        c[NK_LIST_OPEN] = pick(open_cap, env_chunk("NKB200_CHUNK_OPEN", 192));
        c[NK_LIST_CLAIM] = pick(claim_cap, 32);
        c[NK_LIST_SLOW] = pick(slow_cap, 32);
        c[NK_LIST_SPEC] = pick(spec_cap, 64);
    }

    bool prepare_sort(size_t n, std::string &err)
    {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                        (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n, 0, 64,
                                        stream);
        sort_tmp = alloc(bytes);
        sort_tmp_bytes = bytes;
        if (!sort_tmp)
        {
            err = "cannot allocate sort scratch";
            return false;
        }
        return true;
    }

    void begin_timer(int i)
    {
        Timer &t = timers[i];
        if (t.used + 2 > t.ev.size())
        {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            t.ev.push_back(a);
            t.ev.push_back(b);
        }
        cudaEventRecord(t.ev[t.used], stream);
    }
    void end_timer(int i)
    {
        Timer &t = timers[i];
        cudaEventRecord(t.ev[t.used + 1], stream);
        t.used += 2;
    }
    void reset_timer(int i) { timers[i].used = 0; }
    /* (start, end) of timer i's recorded pairs in ms since this GPU's process-wide epoch event: engines that
     * share a GPU overlap, and their busy time is the union of these spans, not the sum */
    void timer_spans(int i, std::vector<float> &out)
    {
        cudaEvent_t e0 = epoch(dev);
        Timer &t = timers[i];
        for (size_t j = 0; j + 1 < t.used; j += 2)
        {
            float a = 0, b = 0;
            if (e0 && cudaEventElapsedTime(&a, e0, t.ev[j]) == cudaSuccess &&
                cudaEventElapsedTime(&b, e0, t.ev[j + 1]) == cudaSuccess)
            {
                out.push_back(a);
                out.push_back(b);
            }
            else
                cudaGetLastError();
        }
    }
    static cudaEvent_t epoch(int device)
    {
        static std::mutex mu;
        static cudaEvent_t ev[64];
        if (device < 0 || device >= 64)
            return nullptr;
        std::lock_guard<std::mutex> lock(mu);
        if (!ev[device])
        {
            cudaSetDevice(device);
            if (cudaEventCreate(&ev[device]) != cudaSuccess || cudaEventRecord(ev[device], 0) != cudaSuccess ||
                cudaEventSynchronize(ev[device]) != cudaSuccess)
            {
                cudaGetLastError();
                ev[device] = nullptr;
            }
        }
        return ev[device];
    }
    /* ms between the last end of timer i and the first start of timer j of the current step (idle or untimed work) */
    float gap_ms(int i, int j)
    {
        Timer &a = timers[i], &b = timers[j];
        float ms = 0;
        if (a.used >= 2 && b.used >= 2 && cudaEventElapsedTime(&ms, a.ev[a.used - 1], b.ev[0]) == cudaSuccess)
            return ms;
        cudaGetLastError();
        return 0;
    }
    float timer_ms(int i)
    {
        Timer &t = timers[i];
        float sum = 0;
        for (size_t j = 0; j + 1 < t.used; j += 2)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.ev[j], t.ev[j + 1]) == cudaSuccess)
                sum += ms;
        }
        t.used = 0;
        return sum;
    }

    /* grids are multiples of the SM count: persistent grid-stride loops, 8 CTAs of 256 threads per SM */
    unsigned grid_for(unsigned long long n, unsigned per_block)
    {
        unsigned long long need = (n + per_block - 1) / per_block;
        unsigned long long cap = (unsigned long long)sms * 8ull;
        if (need >= cap)
            return (unsigned)cap;
        unsigned long long g = ((need + sms - 1) / sms) * sms;
        return (unsigned)(g ? g : sms);
    }

    void probe(const NkRun &P)
    {
        if (!P.n_reads)
            return;
        unsigned g = grid_for(P.n_reads, NK_WARPS);
        launches++;
        switch (P.mode)
        {
        case NK_MODE_SCORE:
            k_probe_score<<<g, NK_THREADS, 0, stream>>>(P);
            break;
        case NK_MODE_SEED:
            k_probe_seed<<<g, NK_THREADS, 0, stream>>>(P);
            break;
        case NK_MODE_COUNT:
            k_probe_count<<<g, NK_THREADS, 0, stream>>>(P);
            break;
        default:
            k_probe_keys<<<g, NK_THREADS, 0, stream>>>(P);
        }
    }
    void hot_flush(const NkRun &P) { k_hot_flush<<<grid_for((unsigned long long)P.hot_mask + 1, 256), 256, 0, stream>>>(P), launches++; }
    void hot_clear(const NkRun &P, unsigned part1)
    {
        k_hot_clear<<<grid_for((unsigned long long)P.hot_mask + 1, 256), 256, 0, stream>>>(P, part1), launches++;
    }
    void prepare_claims(const NkRun &P) { k_prepare_claims<<<sms * 8, 256, 0, stream>>>(P), launches++; }
    void open_ops(const NkRun &P) { k_open<<<sms * 8, 256, 0, stream>>>(P), launches++; }
    void apply(const NkRun &P, unsigned n) { k_apply<<<grid_for(n, 256), 256, 0, stream>>>(P, n), launches++; }
    void classify(const NkRun &P, unsigned n) { k_classify<<<grid_for(n, 256), 256, 0, stream>>>(P, n), launches++; }
    void classify_claimed(const NkRun &P, unsigned n)
    {
        k_classify_claimed<<<grid_for(n, 256), 256, 0, stream>>>(P, n), launches++;
    }
    void sort_pairs(unsigned long long *kin, unsigned long long *kout, unsigned long long *vin, unsigned long long *vout,
                    unsigned n)
    {
        size_t bytes = sort_tmp_bytes;
        ok(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, (const unsigned long long *)kin, kout,
                                           (const unsigned long long *)vin, vout, (int)n, 0, 64, stream),
           "radix sort");
        launches++;
    }
    void rank(const NkRun &P, const unsigned long long *keys, const unsigned long long *vals, unsigned n)
    {
        k_rank<<<grid_for(n, 256), 256, 0, stream>>>(P, keys, vals, n), launches++;
    }
    void commit(const NkRun &P, unsigned n) { k_commit<<<grid_for(n, 256), 256, 0, stream>>>(P, n), launches++; }
    void untag(const NkRun &P, unsigned n)
    {
        if (n)
            k_untag<<<grid_for(n, 256), 256, 0, stream>>>(P, n), launches++;
    }
    void rehash(const NkSlot *old_tab, unsigned long long cap, NkSlot *nt, unsigned long long ncap, unsigned long long nmagic)
    {
        k_rehash_place<<<grid_for(cap, 256), 256, 0, stream>>>(old_tab, cap, nt, ncap, nmagic), launches++;
        k_rehash_fill<<<grid_for(ncap, 256), 256, 0, stream>>>(old_tab, nt, ncap), launches++;
    }
    void decide(const NkRun &P, unsigned n_records, int paired, float coverage, unsigned char *accept)
    {
        if (n_records)
            k_decide<<<grid_for(n_records, 256), 256, 0, stream>>>(P, n_records, paired, coverage, accept), launches++;
    }

    static bool device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes)
    {
        size_t f = 0, t = 0;
        if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess)
        {
            cudaGetLastError();
            return false;
        }
        /* what the stream-ordered pool holds but has handed back counts as free as well */
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
        {
            unsigned long long reserved = 0, used = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
                f += (size_t)(reserved - used);
        }
        if (free_bytes)
            *free_bytes = f;
        if (total_bytes)
            *total_bytes = t;
        return true;
    }
    /* the next step's text goes to the device on its own stream ... */
    /* Large transfers go in pieces: a step's small reads (counters, flags: a handful of round trips per step) share the
     * copy engines and the PCIe link with them, and behind one 100+ MB copy each of them waited milliseconds. */
    size_t piece = 4u << 20;
    void upload(void *d, const void *h, size_t n)
    {
        for (size_t at = 0; at < n; at += piece)
            ok(cudaMemcpyAsync((char *)d + at, (const char *)h + at, n - at < piece ? n - at : piece, cudaMemcpyHostToDevice, up_stream),
               "H2D copy");
        ok(cudaEventRecord(up_done, up_stream), "event record");
    }
    /* ... and the engine's stream picks it up when it has landed */
    void upload_fence() { ok(cudaStreamWaitEvent(stream, up_done, 0), "stream wait"); }
    /* the engine's stream waits until the last text transfer has left the device buffer */
    void copy_fence()
    {
        if (last_copy)
            ok(cudaStreamWaitEvent(stream, last_copy, 0), "stream wait");
    }
    /* the caller has synchronised the engine's stream: the text is complete in d_out */
    void copy_out(void *h, const void *d, size_t n, int slot)
    {
        for (size_t at = 0; at < n; at += piece)
            ok(cudaMemcpyAsync((char *)h + at, (const char *)d + at, n - at < piece ? n - at : piece, cudaMemcpyDeviceToHost,
                               copy_stream),
               "D2H copy");
        ok(cudaEventRecord(copy_done[slot], copy_stream), "event record");
        last_copy = copy_done[slot];
    }
    void copy_wait(int slot) { ok(cudaEventSynchronize(copy_done[slot]), "event synchronize"); }

    /* raw record text: line ends -> records -> operation numbering */
    void *scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    bool prepare_scan(size_t n, std::string &err)
    {
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned *)nullptr, (unsigned *)nullptr, (long long)n, stream);
        if (bytes <= scan_tmp_bytes)
            return true;
        release(scan_tmp);
        scan_tmp = alloc(bytes);
        scan_tmp_bytes = scan_tmp ? bytes : 0;
        if (!scan_tmp)
            err = "cannot allocate scan scratch";
        return scan_tmp != nullptr;
    }
    void scan_u32(unsigned *data, size_t n)
    {
        size_t bytes = scan_tmp_bytes;
        ok(cub::DeviceScan::ExclusiveSum(scan_tmp, bytes, data, data, (long long)n, stream), "exclusive scan");
        launches++;
    }
    void scan_u32(const unsigned *in, unsigned *out, size_t n)
    {
        size_t bytes = scan_tmp_bytes;
        ok(cub::DeviceScan::ExclusiveSum(scan_tmp, bytes, in, out, (long long)n, stream), "exclusive scan");
        launches++;
    }
    void raw_records(const NkRaw &R)
    {
        const unsigned n_tiles = (R.raw_bytes + NK_RAW_TILE - 1) / NK_RAW_TILE, n_reads = R.n_records * R.stride;
        zero(R.tile + n_tiles, sizeof(unsigned));
        k_raw_count<<<grid_for(n_tiles, 1), 256, 0, stream>>>(R, n_tiles), launches++;
        scan_u32(R.tile, (size_t)n_tiles + 1);
        k_raw_positions<<<grid_for(n_tiles, 1), 256, 0, stream>>>(R, n_tiles), launches++;
        zero(R.nops + n_reads, sizeof(unsigned));
        k_raw_records<<<grid_for(R.n_records, 256), 256, 0, stream>>>(R), launches++;
    }
    /* seeding: only the first `limit` records that passed the length gate stay */
    void raw_limit(const NkRaw &R, unsigned limit)
    {
        zero(R.outlen + R.n_records, sizeof(unsigned));
        k_seed_flag<<<grid_for(R.n_records, 256), 256, 0, stream>>>(R), launches++;
        scan_u32(R.outlen, R.outoff, (size_t)R.n_records + 1);
        k_seed_clip<<<grid_for(R.n_records, 256), 256, 0, stream>>>(R, limit), launches++;
    }
    void raw_number(const NkRaw &R)
    {
        const unsigned n_reads = R.n_records * R.stride;
        scan_u32(R.nops, R.opscan, (size_t)n_reads + 1);
        k_raw_opbase<<<grid_for(n_reads, 256), 256, 0, stream>>>(R), launches++;
    }
    void raw_index(const NkRaw &R)
    {
        raw_records(R);
        raw_number(R);
    }
    /* accepted records' text, forward then reverse per window, and the windows' counters */
    void raw_emit(const NkRaw &R)
    {
        const unsigned n_out = R.n_records * R.stride;
        zero(R.outlen + n_out, sizeof(unsigned));
        k_emit_measure<<<grid_for(n_out, 256), 256, 0, stream>>>(R, n_out), launches++;
        scan_u32(R.outlen, R.outoff, (size_t)n_out + 1);
        k_emit_copy<<<grid_for(n_out, 8), 256, 0, stream>>>(R, n_out), launches++;
        k_emit_summary<<<1, 256, 0, stream>>>(R), launches++;
    }

    /* table dump: tile sizes, their exclusive scan in place (tile[n_tiles] = total), then the writers */
    bool dump_scan(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, int k, int text,
                   unsigned long long *d_tile)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        zero(d_tile + tiles, sizeof(unsigned long long));
        k_dump_measure<<<grid_for(tiles, 1), 256, 0, stream>>>(src, lo, n, k, text, d_tile), launches++;
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_tile, d_tile, (long long)(tiles + 1), stream);
        void *tmp = alloc(bytes);
        if (!tmp)
            return false;
        ok(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_tile, d_tile, (long long)(tiles + 1), stream), "tile scan");
        launches++;
        release(tmp);
        return true;
    }
    void dump_text(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, int k,
                   const unsigned long long *d_tile, char *d_text)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        k_dump_text<<<grid_for(tiles, 1), 256, 0, stream>>>(src, lo, n, k, d_tile, d_text), launches++;
    }
    void dump_pairs(const NkDumpSrc &src, unsigned long long lo, unsigned long long n, const unsigned long long *d_tile,
                    unsigned long long *keys_out, long long *vals_out)
    {
        unsigned long long tiles = (n + NK_DUMP_TILE - 1) / NK_DUMP_TILE;
        k_dump_pairs<<<grid_for(tiles, 1), 256, 0, stream>>>(src, lo, n, d_tile, keys_out, vals_out), launches++;
    }
    /* merged table: sort the concatenated (k-mer, count) pairs of all partitions by k-mer and sum the counts
     * of equal k-mers; *d_n_out receives the number of distinct k-mers */
    bool merge_pairs(unsigned long long *keys, long long *vals, unsigned long long n, int key_bits,
                     unsigned long long *keys_tmp, long long *vals_tmp, unsigned long long *d_n_out)
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, keys, keys_tmp, vals, vals_tmp, (long long)n, 0, key_bits, stream);
        cub::DeviceReduce::ReduceByKey(nullptr, b2, keys_tmp, keys, vals_tmp, vals, d_n_out, cuda::std::plus<>(), (long long)n,
                                       stream);
        void *tmp = alloc(b1 > b2 ? b1 : b2);
        if (!tmp)
            return false;
        ok(cub::DeviceRadixSort::SortPairs(tmp, b1, keys, keys_tmp, vals, vals_tmp, (long long)n, 0, key_bits, stream),
           "merge sort");
        ok(cub::DeviceReduce::ReduceByKey(tmp, b2, keys_tmp, keys, vals_tmp, vals, d_n_out, cuda::std::plus<>(), (long long)n,
                                          stream),
           "merge reduce");
        launches += 2;
        release(tmp);
        return true;
    }
};

#define NK_BACKEND CudaBackend
#include "nk_engine_api.h"

extern "C" void *nkd_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void nkd_free_pinned(void *p)
{
    if (p)
        cudaFreeHost(p);
}
extern "C" int nkd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}
