/*
 * nk_b200.h -- C ABI of the B200-native k-mer coverage-normalisation path.
 *
 * Drop-in boundary for the hot path of normalise_kmers_multi_large.c ("C:n" =
 * line n of the reference's single source file).  The reference has no
 * plugin/FFI API; its seam is the set of C functions main() calls per file
 * (C:2239-2409).  Each entry point below names the reference function it
 * replaces.  Plain C types only; all buffers are caller-owned host memory.
 *
 * Two layers, both exported from libnk_b200.so:
 *   nk_*   host pipeline: byte ranges -> record index -> pinned staging ->
 *          device steps -> accepted records written per partition.  This is
 *          what a maintainer of the reference binds.
 *   nkd_*  device engines: HBM-resident per-partition tables, one CUDA stream,
 *          step scratch and the sm_100a kernels (probe / resolve / decide /
 *          rehash / dump).  The host pipeline drives up to four per GPU.  Used
 *          by nk_*, by the parity tests and by bench.py's kernel-only figures.
 *
 * Every function returns 0 on success and a negative NK_E* code otherwise;
 * nk_last_error() / nkd_last_error() give the message.  There is no CPU
 * fallback: without a CUDA device nkd_create() fails with NK_ENODEVICE.
 */
#ifndef NK_B200_H
#define NK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NK_VERSION 20240823 /* same version number the reference prints, C:1 */

#define NK_MAX_PARTITIONS 256 /* C:142 */
#define NK_MAX_LINE 1024      /* C:139: lines are cut at 1023 chars */
#define NK_DEFAULT_SLOTS 67108879ULL /* C:137 */

enum
{
    NK_OK = 0,
    NK_EINVAL = -1,    /* bad argument / configuration (reference: usage + exit 1) */
    NK_ENODEVICE = -2, /* no CUDA device or CUDA error */
    NK_ENOMEM = -3,    /* host or device allocation failed (reference: exit 1, C:1073) */
    NK_EDATA = -4,     /* input is not DNA / malformed (reference: FATAL + exit 1, C:1418) */
    NK_EIO = -5,       /* cannot open / write a file */
    NK_EINTERNAL = -6,
    NK_EIRREGULAR = -7 /* nkd_stage_raw only: the text holds a NUL byte or a line of 1024+ chars, where read_line
                          (C:394-409) splits differently; nothing was staged, use the host-parsed step instead */
};

/* ---------------------------------------------------------------- device engine */

/* One read of a step.  Reads of a step are ordered exactly as the reference's
 * worker visits them (C:1605-1631): partition-major, record order, forward mate
 * before reverse mate.  op_base numbers the read's first k-mer window within
 * its partition's step (window i of the read is operation op_base + i), which
 * is the sequential order store_kmer() sees (C:1464, C:1559-1563). */
typedef struct
{
    uint32_t seq_off; /* byte offset of the sequence line in the step's buffer, 16-aligned */
    uint32_t op_base;
    uint16_t len;  /* bases (>= k; shorter records are dropped by the caller, C:1430-1443) */
    uint16_t part; /* engine-local partition index */
    uint32_t reserved;
} nkd_read;

typedef struct
{
    uint64_t capacity; /* slots (C:167) */
    uint64_t used;     /* stored k-mers (C:166) */
    uint64_t processed, printed, skipped; /* C:179-181 */
    uint64_t ops;        /* table operations issued (non-zero-key windows) */
    uint64_t touches;    /* slots visited: ops + extra probe steps */
    uint64_t expansions; /* C:1055 calls that grew the table */
    uint64_t slow_events; /* events that needed time-ordered resolution */
} nkd_part_stats;

typedef struct nkd_engine nkd_engine;

typedef struct
{
    int device;          /* CUDA ordinal */
    int k;               /* 5..31 (C:724) */
    int canonical;       /* C:1472-1476 */
    int depth_per_part;  /* cfg.depth_per_cpu, C:674 */
    float coverage;      /* cfg.coverage as float32, C:216 */
    int n_parts;         /* partitions this engine holds */
    uint64_t capacity0;  /* cfg.initial_hash_size, C:676-684 */
    uint64_t max_step_reads; /* staging limits of one step */
    uint64_t max_step_bytes;
    uint64_t max_step_ops;
    uint64_t max_raw_bytes; /* raw record text of one step (nkd_stage_raw); 0 = that path is not used */
} nkd_config;

/* init_hash_table (C:890): allocates the zeroed seed table of capacity0 slots and the step scratch */
int nkd_create(const nkd_config *cfg, nkd_engine **out);
void nkd_destroy(nkd_engine *e);
const char *nkd_last_error(const nkd_engine *e);

/* Table budget (bytes of HBM for this engine's partition tables, 0 = unlimited; set before nkd_seed_finish).
 * The reference's threads each own a table and share nothing (C:1841-1880, README:68), so partitions can be worked
 * on in any order: with a budget, a table comes into being (copy_hash_table, C:908) when its partition is first
 * staged, and the least recently used tables of partitions that are not part of the current step wait in host
 * memory.  Results do not depend on the budget. */
int nkd_set_table_budget(nkd_engine *e, uint64_t bytes);
int nkd_residency_stats(nkd_engine *e, uint64_t *resident_parts, uint64_t *evictions, uint64_t *loads);
int nkd_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);

/* sequence_to_hash_zero over a batch of seed reads (C:1501-1537, C:1352); reads[i].part is ignored.
 * n_ops = op_base + windows of the last read. */
int nkd_seed_step(nkd_engine *e, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads,
                  int64_t *first_invalid);
/* seed_kmer_hash (C:1322-1373) on a piece of a file as raw record text: `raw` holds n_records complete records
 * (text_bytes bytes, padded with spaces to a multiple of 16 in a page-locked buffer); the device finds the lines
 * and inserts, with count 0, the first `limit` records whose sequence line is longer than k.  *taken = how many
 * of them this piece had.  NK_EIRREGULAR when the text needs the host parser (nkd_seed_step). */
int nkd_seed_raw(nkd_engine *e, const uint8_t *raw, size_t text_bytes, uint32_t n_records, int lines_per_record,
                 uint32_t limit, uint32_t *taken, int64_t *first_invalid);
/* copy_hash_table for every resident partition (C:908-927, C:2279); frees the seed table */
int nkd_seed_finish(nkd_engine *e);
/* the same for an engine that shares its GPU with `src`: its partitions are copied from src's seed table, so
 * a GPU driven by several engines (one stream each) is seeded once.  Call before nkd_seed_finish(src). */
int nkd_seed_finish_from(nkd_engine *e, nkd_engine *src);
int nkd_seed_stats(nkd_engine *e, nkd_part_stats *out);
/* print_kmer_table's data source for the "_seeds" dump (C:2251-2252): slot-ordered table copy */
int nkd_seed_export(nkd_engine *e, uint64_t *keys, int32_t *counts, uint64_t capacity);

/* process_sequence_pair / process_sequence_single + the print decision for a step of records
 * (C:1539-1566, C:1641-1674, C:1988-2014).  stage = host->device copy of the step's inputs,
 * run = kernels only, fetch = device->host copy of the accept flags (1 byte per record:
 * 1 = emit, 0 = skip) and first_invalid (record index of the first non-ACGT sequence or -1,
 * which the caller turns into the reference's FATAL exit, C:1445-1454). */
int nkd_stage(nkd_engine *e, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads, int paired);
/* the same for a step gathered from several host segments (one per partition in the host pipeline):
 * segment s holds reads[s][0..n_reads[s]) whose seq_off index the step-wide buffer `seq_base`, of which
 * only [seq_lo[s], seq_hi[s]) is copied.  Read/record numbering is the concatenation of the segments. */
typedef struct
{
    const nkd_read *reads;
    size_t n_reads;
    size_t seq_lo, seq_hi;
    /* optional fast path for callers that build the segment themselves (the host pipeline): when trusted is
     * non-zero every read of the segment belongs to partition `part`, satisfies the nkd_read rules, and the
     * partition's operation count in this step is `ops`; the per-read validation pass is skipped */
    int trusted;
    uint32_t part, ops;
} nkd_segment;
int nkd_stage_segments(nkd_engine *e, const uint8_t *seq_base, const nkd_segment *segs, int n_segs, int paired);
/* A step handed over as raw record text (the worker loop's read_line x4 / x2 per mate, C:1605-1631, done on
 * the device): per partition one window of the forward file and, when paired, one of the reverse file, each
 * holding exactly n_records complete records (every line ended by '\n').  The windows lie in one page-locked
 * step buffer at 16-byte aligned offsets; bytes between windows must be neither '\n' nor NUL.  The device finds
 * the line ends, applies the length gate (C:1430-1443: a record with a mate shorter than k vanishes), numbers
 * the operations and scores as nkd_run does.  At most one segment per partition. */
typedef struct
{
    uint32_t part; /* engine-local partition */
    uint32_t n_records;
    uint32_t fwd_off, fwd_bytes;
    uint32_t rev_off, rev_bytes;
} nkd_raw_segment;
/* what came of a segment: where its accepted records' text lies in the output buffer (forward file's
 * records in order, then the reverse file's: the bytes process_thread_chunk_* prints, C:1649-1666, N->A in
 * sequence lines, fastq_to_fasta for emit_mode 1) and its counters (C:1667, C:1672). */
typedef struct
{
    uint64_t fwd_off, fwd_bytes, rev_off, rev_bytes;
    uint64_t processed, printed;
} nkd_raw_result;
/* Optional: start copying the NEXT step's buffer to the device while the current step runs (its own stream, a
 * second device buffer); nkd_stage_raw of that very buffer then finds the bytes in place.  The buffer must stay
 * untouched until that step's nkd_stage_raw has returned.  nkd_seed_raw takes a piece sent ahead the same way.
 * raw = NULL forgets what was sent ahead (before the host buffer is reused for other text of the same size). */
int nkd_upload_raw(nkd_engine *e, const uint8_t *raw, size_t raw_bytes);
/* returns NK_EIRREGULAR (and stages nothing) when the text needs the byte-exact host parser */
int nkd_stage_raw(nkd_engine *e, const uint8_t *raw, size_t raw_bytes, const nkd_raw_segment *segs, int n_segs,
                  int paired, int lines_per_record);
/* after nkd_run: emit_mode 0 = records verbatim (fq->fq, fa->fa), 1 = FASTQ records as FASTA (paired fq->fa),
 * 2 = nothing (single-end fq->fa, C:1995-1999).  out = page-locked buffer of out_cap bytes. */
int nkd_fetch_raw(nkd_engine *e, int emit_mode, uint8_t *out, size_t out_cap, nkd_raw_result *results,
                  int64_t *first_invalid);
/* nkd_fetch_raw returns once the results are known; the text itself is still on its way to `out` on a copy
 * stream (so that the engine's next step does not wait for it).  Call this before reading `out`.  slot = the
 * value passed to nkd_fetch_raw_slot (nkd_fetch_raw uses slot 0); an engine keeps NKD_FETCH_SLOTS transfers apart. */
#define NKD_FETCH_SLOTS 4
int nkd_fetch_raw_slot(nkd_engine *e, int emit_mode, uint8_t *out, size_t out_cap, nkd_raw_result *results,
                       int64_t *first_invalid, int slot);
int nkd_fetch_wait(nkd_engine *e, int slot);
/* page-locked host memory for the staging buffers (cudaMallocHost / cudaFreeHost) */
void *nkd_alloc_pinned(size_t bytes);
void nkd_free_pinned(void *p);
int nkd_device_count(void);
int nkd_run(nkd_engine *e);
int nkd_fetch(nkd_engine *e, uint8_t *accept, size_t n_records, int64_t *first_invalid);
/* time of the last nkd_run on the device, from CUDA events on the engine's stream */
int nkd_last_run_ms(nkd_engine *e, float *total_ms, float *probe_ms);
/* cumulative since nkd_create: what bench.py's roofline and gpu_launches are computed from */
typedef struct
{
    uint64_t launches;       /* kernels launched (all kinds, incl. the slow path's radix sort as 1) */
    uint64_t probe_launches; /* k_probe launches in scoring steps */
    double run_ms;           /* sum of nkd_run device times (CUDA events) */
    double probe_ms;         /* sum of k_probe device times (CUDA events) */
    uint64_t probe_touches;  /* slots visited inside k_probe (the rest are visited by k_open) */
    uint64_t h2d_bytes, d2h_bytes;
    /* CUDA-event time per kernel class in scoring steps: 0 probe, 1 open, 2 apply, 3 classify,
     * 4 sort+rank, 5 commit, 6 decide, 7 growth/undo, 8 raw text -> reads (line ends, records, operation
     * numbering), 9 accepted records' text (measure, scan, copy) */
    double class_ms[10];
    uint64_t pend_events, open_ops, slow_events; /* list entries incl. chunk holes */
    uint64_t hot_hits; /* operations served by the L2-resident table of hot saturated k-mers (no table access) */
} nkd_run_stats;
int nkd_run_stats_get(nkd_engine *e, nkd_run_stats *out);

/* (start, end) of every scoring step in ms on the GPU's clock since a process-wide per-GPU epoch: engines that
 * share one GPU (several streams) overlap, so that GPU's busy time is the union of their spans.  spans holds
 * 2 floats per step; *n_spans = steps recorded so far. */
int nkd_run_spans(nkd_engine *e, float *spans, size_t cap_spans, size_t *n_spans);
int nkd_part_stats_get(nkd_engine *e, int part, nkd_part_stats *out);
/* print_kmer_table's data source (C:354-385): slot-ordered copy of partition's table */
int nkd_export(nkd_engine *e, int part, uint64_t *keys, int32_t *counts, uint64_t capacity);

/* print_kmer_table on the device (C:354-385): the "KMER\tcount\n" lines of entries [first, first+n) of a
 * table, in slot order, formatted by a kernel; only the text crosses PCIe.  part = partition index,
 * NKD_PART_SEED (the seed table, before nkd_seed_finish) or NKD_PART_MERGED (after nkd_merge_finish;
 * entries are then the distinct k-mers in ascending order).  text must hold n * (k + 13) bytes for a
 * table, n * (k + 22) for the merged table; *bytes = text length. */
#define NKD_PART_SEED (-1)
#define NKD_PART_MERGED (-2)
int nkd_dump_text(nkd_engine *e, int part, uint64_t first, uint64_t n, char *text, size_t text_cap, size_t *bytes);
/* stored (k-mer, count) pairs of a table in slot order, compacted on the device; *n = their number (= used) */
int nkd_compact(nkd_engine *e, int part, uint64_t *keys, int64_t *counts, uint64_t cap_entries, uint64_t *n);
/* merged table across partitions (the author's TODO, C:25-26): every stored k-mer once, ascending, with
 * its counts summed over the partitions.  begin(total entries) -> add_part (a partition of this engine)
 * / add (pairs compacted on another GPU) -> finish -> nkd_dump_text(e, NKD_PART_MERGED, ...). */
int nkd_merge_begin(nkd_engine *e, uint64_t max_entries);
int nkd_merge_add_part(nkd_engine *e, int part);
int nkd_merge_add(nkd_engine *e, const uint64_t *keys, const int64_t *counts, uint64_t n);
int nkd_merge_finish(nkd_engine *e, uint64_t *n_unique);

/* test hook: per-read (high, total) of the step just run (sequence_to_hash's two outputs, C:1459-1499);
 * valid between nkd_run and the next nkd_stage */
int nkd_read_scores(nkd_engine *e, uint32_t *high, uint32_t *total, size_t n_reads);

/* test hooks: the codec alone (encode_kmer_plain / get_canonical_kmer, C:1118-1126, C:1175-1180).
 * keys_out receives one key per window of every read (0 = ignored window), in op order. */
int nkd_extract_keys(nkd_engine *e, const uint8_t *seq, size_t seq_bytes, const nkd_read *reads, size_t n_reads,
                     uint64_t *keys_out, size_t n_ops, uint8_t *invalid_out);

/* ---------------------------------------------------------------- host pipeline */

typedef struct
{
    int k;              /* -k */
    int depth;          /* -d */
    float coverage;     /* -g */
    int canonical;      /* -c */
    int in_fastq;       /* -t: 1 fastq, 0 fasta */
    int out_fastq;      /* -o */
    int memory_gb;      /* -m, 0 = default capacity */
    int partitions;     /* -p */
    int dump_tables;    /* -P */
    int verbose;        /* -e */
    int n_forward_files; /* cfg.forward_file_count: sets records_to_seed, C:2242 */
    int have_reverse;   /* cfg.reverse_file_count != 0: open output_reverse files, C:2294 */
    const char *out_dir; /* NULL = current directory (the reference always writes to CWD) */
    /* placement (not in the reference): which GPUs, and which slice of the partitions this
     * process owns (part_first .. part_first+part_count-1; 0,0 = all) */
    int n_devices;
    const int *devices; /* NULL = 0..n_devices-1 */
    int part_first, part_count;
    uint32_t step_pairs; /* records per partition per step, 0 = default */
    /* extras the reference leaves to the user (C:25-26, README): written by nk_finish next to the
     * per-partition files; both need a context that owns every partition */
    int merged_table;  /* output_kmer_merged.k{K}_norm{D}.tsv: all partitions' k-mers, ascending, counts summed */
    int merged_output; /* output_forward/output_reverse.k{K}_norm{D}.fastq: partitions concatenated in order */
} nk_config;

typedef struct nk_ctx nk_ctx;

/* parse_arguments' derived values + init_hash_table (C:674-684, C:890) */
int nk_create(const nk_config *cfg, nk_ctx **out);
void nk_destroy(nk_ctx *c);
const char *nk_last_error(const nk_ctx *c);
/* error text of the most recent failed nk_create */
const char *nk_create_error(void);

/* cfg.initial_hash_size for (-m, -p, -k): memoryGB2capacity in float32 and the 4^k clamp (C:416-422, C:676-684) */
uint64_t nk_initial_capacity(int memory_gb, int partitions, int k);

/* seed_kmer_hash (C:1322-1373) on an in-memory file image */
int nk_seed_buffer(nk_ctx *c, const char *data, size_t size, int records_to_seed);
/* the per-thread copy_hash_table loop + "_seeds" dump + output files opened "w" (C:2251-2302) */
int nk_seed_finish(nk_ctx *c);

/* multithreaded_process_files_paired (C:1772-1920) / _single (C:2113-2217) on in-memory file images:
 * partition byte ranges, score every record, append accepted records to the partition outputs,
 * accumulate counters. */
int nk_process_paired(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size);
int nk_process_single(nk_ctx *c, const char *fwd, size_t fwd_size);

/* The byte ranges multithreaded_process_files_* computes before it starts its threads (C:1796-1838, C:2133-2143):
 * whole file for one partition, calculate_thread_positions for equal sizes / single-end, the record-count
 * partitioner otherwise.  Each array has `partitions` entries; rev may be NULL (single-end).  Split out so that a
 * multi-process launch (one rank per GPU) computes the plan once instead of once per rank. */
int nk_plan_ranges(const char *fwd, size_t fwd_size, const char *rev, size_t rev_size, int partitions, int fastq,
                   int threads, uint64_t *fwd_starts, uint64_t *fwd_ends, uint64_t *rev_starts, uint64_t *rev_ends,
                   char *errbuf, size_t errbuf_size);
/* nk_process_paired / nk_process_single with a plan from nk_plan_ranges (rev arrays ignored for single-end) */
int nk_process_planned(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size,
                       const uint64_t *fwd_starts, const uint64_t *fwd_ends, const uint64_t *rev_starts,
                       const uint64_t *rev_ends);

/* Planning for launches with one process per GPU: the byte ranges and the raw-text steps both start from
 * the number of line ends in fixed chunks of the files (count_records_seqfile's scan, C:1302-1320).  Each
 * rank counts a share of the chunks (nk_count_chunk_lines), the ranks exchange the counts, and every rank
 * passes all of them to nk_process_indexed, which then does what nk_process_paired / nk_process_single do
 * without scanning the files again.  counts has ceil(size / nk_line_chunk_bytes()) entries per file. */
size_t nk_line_chunk_bytes(void);
int nk_count_chunk_lines(const char *data, size_t size, size_t chunk_first, size_t n_chunks, uint32_t *counts,
                         int threads);
int nk_process_indexed(nk_ctx *c, const char *fwd, size_t fwd_size, const char *rev, size_t rev_size,
                       const uint32_t *fwd_counts, const uint32_t *rev_counts);

typedef struct
{
    uint64_t processed, printed, skipped; /* reporting.total_* (C:198-205) */
    uint64_t max_used;                    /* reporting.max_total_kmers */
    double seed_seconds, process_seconds, index_seconds, device_seconds, write_seconds;
    uint64_t h2d_bytes, d2h_bytes;
    /* device-side figures of this context (scoring steps only).  run_ms: per GPU the union of its engines'
     * step spans (nkd_run_spans), summed over the GPUs; the others are sums over the engines */
    double run_ms, probe_ms;
    uint64_t launches, probe_launches;
    uint64_t ops, touches, probe_touches, slow_events, expansions;
    double class_ms[10]; /* see nkd_run_stats */
    uint64_t pend_events, open_ops;
    uint64_t engines; /* engines (stream + scratch + pipeline thread) of this context, over all its GPUs */
    /* device steps by kind: raw record text parsed on the device (the default) / records parsed by the host
     * (NKB200_HOST_PARSE=1, or text the device declined: NUL bytes, lines of 1024+ chars, a cut last record) */
    uint64_t raw_steps, parsed_steps;
    uint64_t hot_hits; /* see nkd_run_stats */
} nk_totals;

int nk_totals_get(nk_ctx *c, nk_totals *out);
int nk_partition_stats(nk_ctx *c, int partition, nkd_part_stats *out);
/* closes outputs; writes output_kmer.k{K}_norm{D}_thread{t}.tsv when dump_tables (C:2398-2413) */
int nk_finish(nk_ctx *c);

/* the byte ranges the reference gives each partition (C:1240-1300), for tests: starts/ends have
 * `partitions` entries.  mode 0 = by size (calculate_thread_positions), 1 = by records. */
int nk_partition_ranges(const char *data, size_t size, int partitions, int fastq, int mode, uint64_t records,
                        uint64_t *starts, uint64_t *ends);
uint64_t nk_count_records(const char *data, size_t size, int fastq);

/* the whole program: same argv contract, stdout lines and exit status as the reference's main (C:2223-2455) */
int nk_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif /* NK_B200_H */
