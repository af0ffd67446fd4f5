"""Host-pipeline binding (nk_* in include/nk_b200.h): the per-file driver of the reference,
multithreaded_process_files_paired/_single (C:1772-1920, C:2113-2217), over in-memory file images."""
from __future__ import annotations

import ctypes as C
import subprocess

import numpy as np

from . import capi


class Pipeline:
    """nk_ctx: seed -> process files -> finish; outputs are the reference's per-partition files in out_dir."""

    def __init__(self, k=15, depth=100, coverage=0.9, canonical=False, in_fastq=True, out_fastq=True, memory_gb=0,
                 partitions=1, dump_tables=False, n_forward_files=1, have_reverse=True, out_dir=".", devices=(0,),
                 part_first=0, part_count=0, step_pairs=0, merged_table=False, merged_output=False, lib=None):
        self.lib = lib if lib is not None else capi.load_library()
        self._devs = (C.c_int * len(devices))(*devices)
        self._out_dir = str(out_dir).encode()
        self.cfg = capi.PipelineConfig(k, depth, coverage, int(canonical), int(in_fastq), int(out_fastq), memory_gb,
                                       partitions, int(dump_tables), 0, n_forward_files, int(have_reverse),
                                       self._out_dir, len(devices), self._devs, part_first, part_count, step_pairs,
                                       int(merged_table), int(merged_output))
        self.h = C.c_void_p()
        rc = self.lib.nk_create(C.byref(self.cfg), C.byref(self.h))
        if rc != capi.NK_OK:
            raise capi.NkError(rc, self.lib.nk_create_error().decode())

    def _check(self, rc):
        if rc != capi.NK_OK:
            raise capi.NkError(rc, self.lib.nk_last_error(self.h).decode())

    @staticmethod
    def _ptr(buf):
        """address + size of a bytes-like / numpy / mmap object without copying"""
        if hasattr(buf, "ctypes"):
            return buf.ctypes.data, buf.nbytes
        mv = memoryview(buf)
        if mv.readonly:
            return C.cast(C.c_char_p(bytes(buf) if not isinstance(buf, bytes) else buf), C.c_void_p).value, mv.nbytes
        return C.addressof((C.c_char * mv.nbytes).from_buffer(buf)), mv.nbytes

    def seed(self, buf, records_to_seed):
        """seed_kmer_hash (C:1322-1373)"""
        p, n = self._ptr(buf)
        self._check(self.lib.nk_seed_buffer(self.h, p, n, records_to_seed))

    def seed_finish(self):
        self._check(self.lib.nk_seed_finish(self.h))

    def process_paired(self, fwd, rev):
        pf, nf = self._ptr(fwd)
        pr, nr = self._ptr(rev)
        self._check(self.lib.nk_process_paired(self.h, pf, nf, pr, nr))

    def process_planned(self, fwd, rev, plan):
        """nk_process_planned: plan = uint64 array of shape (4, partitions) from plan_ranges()"""
        pf, nf = self._ptr(fwd)
        pr, nr = self._ptr(rev) if rev is not None else (None, 0)
        plan = np.ascontiguousarray(plan, dtype=np.uint64)
        self._check(self.lib.nk_process_planned(self.h, pf, nf, pr, nr, plan[0].ctypes.data, plan[1].ctypes.data,
                                                plan[2].ctypes.data, plan[3].ctypes.data))

    def process_indexed(self, fwd, rev, fwd_counts, rev_counts):
        """nk_process_indexed: like process_paired / process_single, with the per-chunk line-end counts of the files
        (count_chunk_lines, gathered from all ranks) so that the files are not scanned again for planning"""
        pf, nf = self._ptr(fwd)
        pr, nr = self._ptr(rev) if rev is not None else (None, 0)
        fc = np.ascontiguousarray(fwd_counts, dtype=np.uint32)
        rc_ = np.ascontiguousarray(rev_counts, dtype=np.uint32) if rev is not None else None
        self._check(self.lib.nk_process_indexed(self.h, pf, nf, pr, nr, fc.ctypes.data,
                                                rc_.ctypes.data if rc_ is not None else None))

    def process_single(self, fwd):
        pf, nf = self._ptr(fwd)
        self._check(self.lib.nk_process_single(self.h, pf, nf))

    def totals(self):
        t = capi.Totals()
        self._check(self.lib.nk_totals_get(self.h, C.byref(t)))
        return t.as_dict()

    def partition_stats(self, partition):
        st = capi.PartStats()
        self._check(self.lib.nk_partition_stats(self.h, partition, C.byref(st)))
        return st.as_dict()

    def finish(self):
        self._check(self.lib.nk_finish(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.nk_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def plan_ranges(fwd, rev, partitions, fastq=True, threads=0, lib=None):
    """nk_plan_ranges: the byte ranges of every partition (C:1796-1838); returns a (4, partitions) uint64 array."""
    lib = lib if lib is not None else capi.load_library()
    plan = np.zeros((4, partitions), dtype=np.uint64)
    pf, nf = Pipeline._ptr(fwd)
    pr, nr = Pipeline._ptr(rev) if rev is not None else (None, 0)
    err = C.create_string_buffer(512)
    rc = lib.nk_plan_ranges(pf, nf, pr, nr, partitions, int(fastq), threads, plan[0].ctypes.data, plan[1].ctypes.data,
                            plan[2].ctypes.data, plan[3].ctypes.data, err, 512)
    if rc != capi.NK_OK:
        raise capi.NkError(rc, err.value.decode())
    return plan


def count_chunk_lines(buf, rank=0, world=1, threads=0, lib=None):
    """nk_count_chunk_lines for this rank's share of a file's chunks: returns (counts of the share as uint32,
    number of chunks of the whole file, chunks per share).  Shares are equal blocks, the last one may be shorter."""
    lib = lib if lib is not None else capi.load_library()
    p, n = Pipeline._ptr(buf)
    chunk = lib.nk_line_chunk_bytes()
    n_chunks = (n + chunk - 1) // chunk
    share = (n_chunks + world - 1) // world
    first = min(n_chunks, rank * share)
    mine = max(0, min(share, n_chunks - first))
    counts = np.zeros(max(mine, 1), dtype=np.uint32)
    rc = lib.nk_count_chunk_lines(p, n, first, mine, counts.ctypes.data, threads)
    if rc != capi.NK_OK:
        raise capi.NkError(rc, "nk_count_chunk_lines")
    return counts[:mine], n_chunks, share


def run_cli(args, cwd=None, env=None, **kw):
    """Run the drop-in command-line program (same argv contract as the reference binary)."""
    if not capi.CLI_PATH.exists():
        raise FileNotFoundError(f"{capi.CLI_PATH} is missing: run __graft_entry__.build()")
    return subprocess.run([str(capi.CLI_PATH)] + [str(a) for a in args], cwd=cwd, env=env, **kw)
