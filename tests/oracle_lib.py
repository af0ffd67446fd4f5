"""ctypes binding of oracle/libnk_oracle.so -- TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "libnk_oracle.so"
ORACLE_CLI = ORACLE_DIR / "nk_oracle"
REF_BIN = ORACLE_DIR / "_ref" / "nkml"
REF_BIN_TLS = ORACLE_DIR / "_ref" / "nkml_tls"


def build_oracle():
    if not ORACLE_LIB.exists() or not ORACLE_CLI.exists() or \
            ORACLE_LIB.stat().st_mtime < (ORACLE_DIR / "nk_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_LIB))
        vp = C.c_void_p
        L.nko_table_new.restype = vp
        L.nko_table_new.argtypes = [C.c_size_t]
        L.nko_table_free.argtypes = [vp]
        L.nko_table_clone.restype = vp
        L.nko_table_clone.argtypes = [vp]
        for f in ("cap", "used"):
            getattr(L, f"nko_table_{f}").restype = C.c_size_t
            getattr(L, f"nko_table_{f}").argtypes = [vp]
        for f in ("ops", "touches", "walk_ops", "expansions"):
            getattr(L, f"nko_table_{f}").restype = C.c_uint64
            getattr(L, f"nko_table_{f}").argtypes = [vp]
        L.nko_table_export.argtypes = [vp, vp, vp]
        L.nko_store.restype = C.c_size_t
        L.nko_store.argtypes = [vp, C.c_uint64, C.c_int]
        L.nko_expand.argtypes = [vp]
        L.nko_slot_count.restype = C.c_int32
        L.nko_slot_count.argtypes = [vp, C.c_size_t]
        L.nko_score.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nko_window_keys.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, vp]
        L.nko_keep_mate.argtypes = [C.c_int, C.c_int, C.c_float]
        L.nko_capacity.restype = C.c_size_t
        L.nko_capacity.argtypes = [C.c_int, C.c_int, C.c_int]
        L.nko_encode.restype = C.c_uint64
        L.nko_encode.argtypes = [C.c_char_p, C.c_int]
        L.nko_revcomp.restype = C.c_uint64
        L.nko_revcomp.argtypes = [C.c_uint64, C.c_int]
        _lib = L
    return _lib


class OracleTable:
    """The reference's per-thread hash_table_t with sequential semantics (C:929-1108)."""

    def __init__(self, cap=None, _h=None):
        self.L = lib()
        self.h = _h if _h is not None else self.L.nko_table_new(cap)

    def clone(self):
        return OracleTable(_h=self.L.nko_table_clone(self.h))

    def __del__(self):
        if self.h:
            self.L.nko_table_free(self.h)
            self.h = None

    cap = property(lambda s: s.L.nko_table_cap(s.h))
    used = property(lambda s: s.L.nko_table_used(s.h))
    ops = property(lambda s: s.L.nko_table_ops(s.h))
    touches = property(lambda s: s.L.nko_table_touches(s.h))
    expansions = property(lambda s: s.L.nko_table_expansions(s.h))

    def store(self, key, init=False):
        return self.L.nko_store(self.h, int(key), int(init))

    def seed(self, seq: bytes, k, canonical):
        """sequence_to_hash_zero (C:1501-1537); N is scrubbed by the caller (C:1406)."""
        seq = seq.replace(b"N", b"A")
        keys = window_keys(seq, k, canonical)
        for x in keys:
            if x:
                self.L.nko_store(self.h, int(x), 1)

    def score(self, seq: bytes, k, canonical, depth):
        seq = seq.replace(b"N", b"A")
        hi, tot = C.c_int(), C.c_int()
        self.L.nko_score(self.h, seq, len(seq), k, int(canonical), depth, C.byref(hi), C.byref(tot))
        return hi.value, tot.value

    def export(self):
        keys, counts = np.empty(self.cap, np.uint64), np.empty(self.cap, np.int32)
        self.L.nko_table_export(self.h, keys.ctypes.data, counts.ctypes.data)
        return keys, counts


def window_keys(seq: bytes, k, canonical):
    n = len(seq) - k + 1
    out = np.zeros(max(n, 0), np.uint64)
    if n > 0:
        lib().nko_window_keys(seq, len(seq), k, int(canonical), out.ctypes.data)
    return out


def keep_mate(high, total, coverage):
    return bool(lib().nko_keep_mate(high, total, coverage))


def oracle_records(table: OracleTable, records, k, canonical, depth, coverage, paired):
    """The worker loop's scoring + decision for a list of records (C:1605-1674): returns accept flags."""
    acc = []
    for rec in records:
        if paired:
            f, r = rec
            if len(f) < k or len(r) < k:
                continue
            hf, tf = table.score(f, k, canonical, depth)
            hr, tr = table.score(r, k, canonical, depth)
            acc.append(keep_mate(hf, tf, coverage) and keep_mate(hr, tr, coverage))
        else:
            if len(rec) < k:
                continue
            hf, tf = table.score(rec, k, canonical, depth)
            acc.append(keep_mate(hf, tf, coverage))
    return np.array(acc, dtype=np.uint8)
