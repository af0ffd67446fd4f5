"""The C-ABI library must load without a GPU and export every symbol include/nk_b200.h declares."""
import ctypes
import re
from pathlib import Path

from nomalise_kmers_multi_large_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "nk_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nkd?_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    if not capi.LIB_PATH.exists():
        __graft_entry__.build()
    lib = ctypes.CDLL(str(capi.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(capi.ENGINE_SYMBOLS + capi.PIPELINE_SYMBOLS) <= set(names)


def test_capacity_maths_matches_reference_values():
    """memoryGB2capacity in float32 + 4^k clamp (C:416-422, C:676-684); values verified in SURVEY 8.A."""
    lib = capi.load_library()
    f = lib.nk_initial_capacity
    assert f(0, 1, 15) == 67108879
    assert f(1, 8, 15) == 8388609
    assert f(1, 64, 15) == 1048577
    assert f(1, 200, 15) == 335545
    assert f(64, 64, 25) == 67108865
    assert f(0, 1, 5) == 1024
    from tests import oracle_lib as ol
    for m, p, k in [(0, 1, 31), (3, 7, 13), (17, 5, 12), (1, 1, 10), (2, 256, 21)]:
        assert f(m, p, k) == ol.lib().nko_capacity(m, p, k)


def test_no_device_means_loud_failure_not_fallback():
    """Without a CUDA device the engine refuses to start (there is no CPU path in the product)."""
    lib = capi.load_library()
    if lib.nkd_device_count() > 0:
        return
    try:
        capi.Engine(capacity0=1009, lib=lib)
    except capi.NkError as e:
        assert e.code == -2
    else:
        raise AssertionError("engine started without a GPU")


def test_partition_ranges_match_oracle_cli_counts(tmp_path):
    """calculate_thread_positions* (C:1240-1300) through the C ABI: ranges tile the file as the reference's do."""
    import numpy as np
    from tests import cli_cases as cc
    lib = capi.load_library()
    f, _ = cc.synth(tmp_path, "s", 4000, seed=3)
    data = f.read_bytes()
    buf = ctypes.create_string_buffer(data, len(data))
    recs = lib.nk_count_records(buf, len(data), 1)
    assert recs == 4000
    for p in (2, 3, 8):
        st, en = np.zeros(p, np.uint64), np.zeros(p, np.uint64)
        assert lib.nk_partition_ranges(buf, len(data), p, 1, 1, recs, st.ctypes.data, en.ctypes.data) == 0
        assert st[0] == 0 and en[-1] == len(data) - 1
        for t in range(p - 1):
            assert data[int(en[t])] == 10 and st[t + 1] == en[t] + 1
            assert data[int(st[t]):int(en[t]) + 1].count(b"\n") == (recs // p) * 4
        st2, en2 = np.zeros(p, np.uint64), np.zeros(p, np.uint64)
        assert lib.nk_partition_ranges(buf, len(data), p, 1, 0, 0, st2.ctypes.data, en2.ctypes.data) == 0
        assert st2[0] == 0 and (p < 2 or st2[1] == 0)          # starts[1] is never assigned
        assert en2[-1] < len(data) - 1                          # ... and the file tail is dropped
