"""ctypes binding of the C ABI in include/nk_b200.h (libnk_b200.so).

Mirrors the reference's call seam (normalise_kmers_multi_large.c, "C:n"):
``Engine`` = the per-GPU table engine (store_kmer / sequence_to_hash / the print decision,
C:929-1053, C:1459-1499, C:1641-1646), ``Pipeline`` = the per-file host driver
(multithreaded_process_files_paired/_single, C:1772-1920, C:2113-2217).

There is no CPU fallback: loading fails loudly when the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "csrc" / "libnk_b200.so"
CLI_PATH = PKG_DIR / "csrc" / "normalise_kmers_multi_large_b200"

NK_OK = 0
ERRORS = {-1: "NK_EINVAL", -2: "NK_ENODEVICE", -3: "NK_ENOMEM", -4: "NK_EDATA", -5: "NK_EIO", -6: "NK_EINTERNAL",
          -7: "NK_EIRREGULAR"}


class NkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class ReadDesc(C.Structure):
    """nkd_read"""
    _fields_ = [("seq_off", C.c_uint32), ("op_base", C.c_uint32), ("len", C.c_uint16), ("part", C.c_uint16),
                ("reserved", C.c_uint32)]


READ_DTYPE = np.dtype([("seq_off", "<u4"), ("op_base", "<u4"), ("len", "<u2"), ("part", "<u2"), ("reserved", "<u4")])


class PartStats(C.Structure):
    """nkd_part_stats"""
    _fields_ = [(n, C.c_uint64) for n in
                ("capacity", "used", "processed", "printed", "skipped", "ops", "touches", "expansions", "slow_events")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class RunStats(C.Structure):
    """nkd_run_stats"""
    _fields_ = [("launches", C.c_uint64), ("probe_launches", C.c_uint64), ("run_ms", C.c_double),
                ("probe_ms", C.c_double), ("probe_touches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("class_ms", C.c_double * 10), ("pend_events", C.c_uint64),
                ("open_ops", C.c_uint64), ("slow_events", C.c_uint64), ("hot_hits", C.c_uint64)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "class_ms" else getattr(self, n)) for n, _ in self._fields_}


class EngineConfig(C.Structure):
    """nkd_config"""
    _fields_ = [("device", C.c_int), ("k", C.c_int), ("canonical", C.c_int), ("depth_per_part", C.c_int),
                ("coverage", C.c_float), ("n_parts", C.c_int), ("capacity0", C.c_uint64),
                ("max_step_reads", C.c_uint64), ("max_step_bytes", C.c_uint64), ("max_step_ops", C.c_uint64),
                ("max_raw_bytes", C.c_uint64)]


class RawSegment(C.Structure):
    """nkd_raw_segment"""
    _fields_ = [(n, C.c_uint32) for n in ("part", "n_records", "fwd_off", "fwd_bytes", "rev_off", "rev_bytes")]


class RawResult(C.Structure):
    """nkd_raw_result"""
    _fields_ = [(n, C.c_uint64) for n in ("fwd_off", "fwd_bytes", "rev_off", "rev_bytes", "processed", "printed")]


class PipelineConfig(C.Structure):
    """nk_config"""
    _fields_ = [("k", C.c_int), ("depth", C.c_int), ("coverage", C.c_float), ("canonical", C.c_int),
                ("in_fastq", C.c_int), ("out_fastq", C.c_int), ("memory_gb", C.c_int), ("partitions", C.c_int),
                ("dump_tables", C.c_int), ("verbose", C.c_int), ("n_forward_files", C.c_int),
                ("have_reverse", C.c_int), ("out_dir", C.c_char_p), ("n_devices", C.c_int),
                ("devices", C.POINTER(C.c_int)), ("part_first", C.c_int), ("part_count", C.c_int),
                ("step_pairs", C.c_uint32), ("merged_table", C.c_int), ("merged_output", C.c_int)]


class Totals(C.Structure):
    """nk_totals"""
    _fields_ = [("processed", C.c_uint64), ("printed", C.c_uint64), ("skipped", C.c_uint64), ("max_used", C.c_uint64),
                ("seed_seconds", C.c_double), ("process_seconds", C.c_double), ("index_seconds", C.c_double),
                ("device_seconds", C.c_double), ("write_seconds", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("run_ms", C.c_double), ("probe_ms", C.c_double),
                ("launches", C.c_uint64), ("probe_launches", C.c_uint64), ("ops", C.c_uint64),
                ("touches", C.c_uint64), ("probe_touches", C.c_uint64), ("slow_events", C.c_uint64),
                ("expansions", C.c_uint64), ("class_ms", C.c_double * 10), ("pend_events", C.c_uint64),
                ("open_ops", C.c_uint64), ("engines", C.c_uint64), ("raw_steps", C.c_uint64),
                ("parsed_steps", C.c_uint64), ("hot_hits", C.c_uint64)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "class_ms" else getattr(self, n)) for n, _ in self._fields_}


ENGINE_SYMBOLS = ["nkd_create", "nkd_destroy", "nkd_last_error", "nkd_seed_step", "nkd_seed_finish", "nkd_seed_stats",
                  "nkd_seed_export", "nkd_stage", "nkd_run", "nkd_fetch", "nkd_last_run_ms", "nkd_part_stats_get",
                  "nkd_export", "nkd_extract_keys", "nkd_stage_segments", "nkd_alloc_pinned", "nkd_free_pinned",
                  "nkd_device_count", "nkd_run_stats_get", "nkd_read_scores", "nkd_dump_text", "nkd_compact",
                  "nkd_merge_begin", "nkd_merge_add_part", "nkd_merge_add", "nkd_merge_finish", "nkd_run_spans", "nkd_seed_finish_from",
                  "nkd_stage_raw", "nkd_fetch_raw", "nkd_fetch_raw_slot", "nkd_fetch_wait",
                  "nkd_upload_raw", "nkd_set_table_budget", "nkd_residency_stats", "nkd_device_memory", "nkd_seed_raw"]
PART_SEED, PART_MERGED = -1, -2
PIPELINE_SYMBOLS = ["nk_create", "nk_destroy", "nk_last_error", "nk_create_error", "nk_initial_capacity",
                    "nk_seed_buffer", "nk_seed_finish", "nk_process_paired", "nk_process_single", "nk_totals_get",
                    "nk_partition_stats", "nk_finish", "nk_partition_ranges", "nk_count_records", "nk_main",
                    "nk_plan_ranges", "nk_process_planned", "nk_line_chunk_bytes", "nk_count_chunk_lines",
                    "nk_process_indexed"]


def _declare_engine(lib):
    vp, u8p, sz = C.c_void_p, C.c_void_p, C.c_size_t
    lib.nkd_create.argtypes = [C.POINTER(EngineConfig), C.POINTER(vp)]
    lib.nkd_destroy.argtypes = [vp]
    lib.nkd_destroy.restype = None
    lib.nkd_last_error.argtypes = [vp]
    lib.nkd_last_error.restype = C.c_char_p
    lib.nkd_seed_step.argtypes = [vp, u8p, sz, vp, sz, C.POINTER(C.c_int64)]
    lib.nkd_seed_finish.argtypes = [vp]
    lib.nkd_seed_stats.argtypes = [vp, C.POINTER(PartStats)]
    lib.nkd_seed_export.argtypes = [vp, vp, vp, C.c_uint64]
    lib.nkd_stage.argtypes = [vp, u8p, sz, vp, sz, C.c_int]
    lib.nkd_run.argtypes = [vp]
    lib.nkd_fetch.argtypes = [vp, u8p, sz, C.POINTER(C.c_int64)]
    lib.nkd_last_run_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.nkd_part_stats_get.argtypes = [vp, C.c_int, C.POINTER(PartStats)]
    lib.nkd_export.argtypes = [vp, C.c_int, vp, vp, C.c_uint64]
    lib.nkd_extract_keys.argtypes = [vp, u8p, sz, vp, sz, vp, sz, u8p]
    lib.nkd_run_stats_get.argtypes = [vp, C.POINTER(RunStats)]
    lib.nkd_read_scores.argtypes = [vp, vp, vp, sz]
    lib.nkd_seed_finish_from.argtypes = [vp, vp]
    lib.nkd_run_spans.argtypes = [vp, vp, sz, C.POINTER(sz)]
    lib.nkd_dump_text.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, vp, sz, C.POINTER(sz)]
    lib.nkd_compact.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.nkd_merge_begin.argtypes = [vp, C.c_uint64]
    lib.nkd_merge_add_part.argtypes = [vp, C.c_int]
    lib.nkd_merge_add.argtypes = [vp, vp, vp, C.c_uint64]
    lib.nkd_merge_finish.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.nkd_device_count.restype = C.c_int
    lib.nkd_stage_raw.argtypes = [vp, u8p, sz, C.POINTER(RawSegment), C.c_int, C.c_int, C.c_int]
    lib.nkd_fetch_raw.argtypes = [vp, C.c_int, u8p, sz, C.POINTER(RawResult), C.POINTER(C.c_int64)]
    lib.nkd_fetch_raw_slot.argtypes = [vp, C.c_int, u8p, sz, C.POINTER(RawResult), C.POINTER(C.c_int64), C.c_int]
    lib.nkd_fetch_wait.argtypes = [vp, C.c_int]
    lib.nkd_upload_raw.argtypes = [vp, u8p, sz]
    lib.nkd_seed_raw.argtypes = [vp, u8p, sz, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]
    lib.nkd_set_table_budget.argtypes = [vp, C.c_uint64]
    lib.nkd_residency_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.nkd_device_memory.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib


def _declare_pipeline(lib):
    vp, sz = C.c_void_p, C.c_size_t
    lib.nk_create.argtypes = [C.POINTER(PipelineConfig), C.POINTER(vp)]
    lib.nk_destroy.argtypes = [vp]
    lib.nk_destroy.restype = None
    lib.nk_last_error.argtypes = [vp]
    lib.nk_last_error.restype = C.c_char_p
    lib.nk_create_error.restype = C.c_char_p
    lib.nk_initial_capacity.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.nk_initial_capacity.restype = C.c_uint64
    lib.nk_seed_buffer.argtypes = [vp, vp, sz, C.c_int]
    lib.nk_seed_finish.argtypes = [vp]
    lib.nk_process_paired.argtypes = [vp, vp, sz, vp, sz]
    lib.nk_process_single.argtypes = [vp, vp, sz]
    lib.nk_totals_get.argtypes = [vp, C.POINTER(Totals)]
    lib.nk_partition_stats.argtypes = [vp, C.c_int, C.POINTER(PartStats)]
    lib.nk_finish.argtypes = [vp]
    lib.nk_partition_ranges.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_uint64, vp, vp]
    lib.nk_count_records.argtypes = [vp, sz, C.c_int]
    lib.nk_count_records.restype = C.c_uint64
    lib.nk_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.nk_plan_ranges.argtypes = [vp, sz, vp, sz, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_char_p, sz]
    lib.nk_process_planned.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, vp]
    lib.nk_line_chunk_bytes.restype = sz
    lib.nk_count_chunk_lines.argtypes = [vp, sz, sz, sz, vp, C.c_int]
    lib.nk_process_indexed.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    return lib


_LIB = None


def load_library(path: os.PathLike | None = None):
    """Load libnk_b200.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(f"{p} is missing: build it with `make -C {PKG_DIR / 'csrc'}` "
                                "(there is no CPU fallback for the CUDA engine)")
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    _declare_engine(lib)
    if hasattr(lib, "nk_create"):
        _declare_pipeline(lib)
    if path is None:
        _LIB = lib
    return lib


def pack_reads(seqs, parts=None, k=15):
    """Lay reads out the way nkd_stage expects them: sequence bytes 16-byte aligned, op_base numbered
    per partition in visiting order (window i of a read is operation op_base+i, C:1464)."""
    n = len(seqs)
    descs = np.zeros(n, dtype=READ_DTYPE)
    offs, pos = [], 0
    for s in seqs:
        offs.append(pos)
        pos += (len(s) + 15) & ~15
    buf = np.zeros(pos + 16, dtype=np.uint8)
    next_op = {}
    for i, s in enumerate(seqs):
        b = np.frombuffer(s if isinstance(s, (bytes, bytearray)) else s.encode(), dtype=np.uint8)
        buf[offs[i]:offs[i] + len(b)] = b
        p = int(parts[i]) if parts is not None else 0
        descs[i] = (offs[i], next_op.get(p, 0), len(b), p, 0)
        next_op[p] = next_op.get(p, 0) + len(b) - k + 1
    return buf, descs, next_op


class Engine:
    """One per-GPU engine (nkd_*)."""

    def __init__(self, k=15, canonical=False, depth_per_part=100, coverage=0.9, n_parts=1, capacity0=67108879,
                 max_step_reads=1 << 16, max_step_bytes=1 << 24, max_step_ops=1 << 22, device=0, lib=None,
                 max_raw_bytes=1 << 25):
        self.lib = lib if lib is not None else load_library()
        self.cfg = EngineConfig(device, k, int(canonical), depth_per_part, coverage, n_parts, capacity0,
                                max_step_reads, max_step_bytes, max_step_ops, max_raw_bytes)
        self.h = C.c_void_p()
        rc = self.lib.nkd_create(C.byref(self.cfg), C.byref(self.h))
        if rc != NK_OK:
            msg = self.lib.nkd_last_error(self.h).decode() if self.h else "allocation failed"
            self.close()
            raise NkError(rc, msg)
        self._keep = None

    def _check(self, rc):
        if rc != NK_OK:
            raise NkError(rc, self.lib.nkd_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.nkd_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_table_budget(self, nbytes):
        self._check(self.lib.nkd_set_table_budget(self.h, nbytes))

    def residency(self):
        r, e, l = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.nkd_residency_stats(self.h, C.byref(r), C.byref(e), C.byref(l))
        return {"resident": r.value, "evictions": e.value, "loads": l.value}

    def seed_step(self, buf: np.ndarray, descs: np.ndarray):
        inv = C.c_int64(-1)
        self._check(self.lib.nkd_seed_step(self.h, buf.ctypes.data, buf.size, descs.ctypes.data, len(descs), C.byref(inv)))
        return inv.value

    def seed_raw(self, text: bytes, n_records, limit, lines_per_record=4, ahead=None):
        """nkd_seed_raw: returns (records taken, first invalid record or -1).  ahead: "use" sends the piece ahead with
        nkd_upload_raw first, "forget" sends it and takes that back (nkd_upload_raw(NULL)), "other" sends a different
        buffer ahead (the piece itself then travels with the call)."""
        raw = np.frombuffer(text + b" " * ((-len(text)) % 16), dtype=np.uint8).copy()
        if ahead in ("use", "forget"):
            self._check(self.lib.nkd_upload_raw(self.h, raw.ctypes.data, raw.size))
            if ahead == "forget":
                self._check(self.lib.nkd_upload_raw(self.h, None, 0))
        elif ahead == "other":
            self._other = np.full(raw.size + 16, ord("\n"), dtype=np.uint8)
            self._check(self.lib.nkd_upload_raw(self.h, self._other.ctypes.data, self._other.size))
        taken, inv = C.c_uint32(0), C.c_int64(-1)
        self._check(self.lib.nkd_seed_raw(self.h, raw.ctypes.data, len(text), n_records, lines_per_record, limit,
                                          C.byref(taken), C.byref(inv)))
        return taken.value, inv.value

    def seed_finish(self):
        self._check(self.lib.nkd_seed_finish(self.h))

    def seed_finish_from(self, src):
        """copy_hash_table from another engine's seed table on the same GPU (call before src.seed_finish())"""
        self._check(self.lib.nkd_seed_finish_from(self.h, src.h))

    def run_spans(self):
        n = C.c_size_t(0)
        self.lib.nkd_run_spans(self.h, None, 0, C.byref(n))
        buf = np.zeros(2 * max(1, n.value), np.float32)
        self.lib.nkd_run_spans(self.h, buf.ctypes.data, n.value, C.byref(n))
        return buf[:2 * n.value].reshape(-1, 2)

    def seed_stats(self):
        st = PartStats()
        self._check(self.lib.nkd_seed_stats(self.h, C.byref(st)))
        return st.as_dict()

    def seed_export(self):
        cap = self.seed_stats()["capacity"]
        keys, counts = np.empty(cap, np.uint64), np.empty(cap, np.int32)
        self._check(self.lib.nkd_seed_export(self.h, keys.ctypes.data, counts.ctypes.data, cap))
        return keys, counts

    def stage(self, buf, descs, paired):
        self._keep = (buf, descs)  # nkd_fetch reads descs[i].part
        self._check(self.lib.nkd_stage(self.h, buf.ctypes.data, buf.size, descs.ctypes.data, len(descs), int(paired)))
        self._nrec = len(descs) // 2 if paired else len(descs)

    def run(self):
        self._check(self.lib.nkd_run(self.h))

    def fetch(self):
        acc = np.empty(self._nrec, np.uint8)
        inv = C.c_int64(-1)
        self._check(self.lib.nkd_fetch(self.h, acc.ctypes.data, self._nrec, C.byref(inv)))
        return acc, inv.value

    def step(self, buf, descs, paired):
        self.stage(buf, descs, paired)
        self.run()
        return self.fetch()

    def step_raw(self, windows, paired, lines_per_record=4, emit_mode=0):
        """nkd_stage_raw + nkd_run + nkd_fetch_raw.  windows = [(part, n_records, fwd_bytes, rev_bytes or None)];
        returns ([(fwd_text, rev_text, processed, printed)], first_invalid)."""
        segs = (RawSegment * len(windows))()
        parts_bytes, at = [], 0
        for i, (part, n, fwd, rev) in enumerate(windows):
            segs[i].part, segs[i].n_records = part, n
            for mate, text in enumerate((fwd, rev) if paired else (fwd,)):
                if mate == 0:
                    segs[i].fwd_off, segs[i].fwd_bytes = at, len(text)
                else:
                    segs[i].rev_off, segs[i].rev_bytes = at, len(text)
                pad = (-len(text)) % 16
                parts_bytes.append(text + b" " * pad)
                at += len(text) + pad
        raw = np.frombuffer(b"".join(parts_bytes), dtype=np.uint8).copy()
        self._check(self.lib.nkd_stage_raw(self.h, raw.ctypes.data, raw.size, segs, len(windows), int(paired),
                                           lines_per_record))
        self._check(self.lib.nkd_run(self.h))
        out = np.zeros(raw.size + 4 * sum(w[1] for w in windows) + 64, np.uint8)
        res = (RawResult * len(windows))()
        inv = C.c_int64(-1)
        self._check(self.lib.nkd_fetch_raw(self.h, emit_mode, out.ctypes.data, out.size, res, C.byref(inv)))
        self._check(self.lib.nkd_fetch_wait(self.h, 0))
        got = []
        for r in res:
            got.append((out[r.fwd_off:r.fwd_off + r.fwd_bytes].tobytes(), out[r.rev_off:r.rev_off + r.rev_bytes].tobytes(),
                        int(r.processed), int(r.printed)))
        return got, inv.value

    def read_scores(self, n_reads):
        hi, tot = np.empty(n_reads, np.uint32), np.empty(n_reads, np.uint32)
        self._check(self.lib.nkd_read_scores(self.h, hi.ctypes.data, tot.ctypes.data, n_reads))
        return hi, tot

    def last_run_ms(self):
        a, b = C.c_float(), C.c_float()
        self.lib.nkd_last_run_ms(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run_stats(self):
        rs = RunStats()
        self._check(self.lib.nkd_run_stats_get(self.h, C.byref(rs)))
        return rs.as_dict()

    def part_stats(self, part):
        st = PartStats()
        self._check(self.lib.nkd_part_stats_get(self.h, part, C.byref(st)))
        return st.as_dict()

    def export(self, part):
        cap = self.part_stats(part)["capacity"]
        keys, counts = np.empty(cap, np.uint64), np.empty(cap, np.int32)
        self._check(self.lib.nkd_export(self.h, part, keys.ctypes.data, counts.ctypes.data, cap))
        return keys, counts

    def dump_text(self, part, first, n):
        """print_kmer_table's lines for entries [first, first+n), formatted on the device (C:354-385)"""
        buf = np.empty(max(1, n * (self.cfg.k + (22 if part == PART_MERGED else 13))), np.uint8)
        got = C.c_size_t(0)
        self._check(self.lib.nkd_dump_text(self.h, part, first, n, buf.ctypes.data, buf.size, C.byref(got)))
        return buf[:got.value].tobytes()

    def compact(self, part, cap_entries):
        keys, counts = np.empty(max(1, cap_entries), np.uint64), np.empty(max(1, cap_entries), np.int64)
        n = C.c_uint64(0)
        self._check(self.lib.nkd_compact(self.h, part, keys.ctypes.data, counts.ctypes.data, cap_entries, C.byref(n)))
        return keys[:n.value], counts[:n.value]

    def merge(self, parts=(), extra=None, max_entries=0):
        """merged table: partitions of this engine plus optional (keys, counts) arrays; returns distinct k-mers"""
        self._check(self.lib.nkd_merge_begin(self.h, max_entries))
        for p in parts:
            self._check(self.lib.nkd_merge_add_part(self.h, p))
        if extra is not None:
            k = np.ascontiguousarray(extra[0], np.uint64)
            v = np.ascontiguousarray(extra[1], np.int64)
            self._check(self.lib.nkd_merge_add(self.h, k.ctypes.data, v.ctypes.data, len(k)))
        n = C.c_uint64(0)
        self._check(self.lib.nkd_merge_finish(self.h, C.byref(n)))
        return n.value

    def extract_keys(self, buf, descs, n_ops):
        keys = np.empty(n_ops, np.uint64)
        inv = np.zeros(len(descs), np.uint8)
        self._check(self.lib.nkd_extract_keys(self.h, buf.ctypes.data, buf.size, descs.ctypes.data, len(descs),
                                              keys.ctypes.data, n_ops, inv.ctypes.data))
        return keys, inv
