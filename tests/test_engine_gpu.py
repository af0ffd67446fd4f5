"""B200 parity tests of the device engine through the C ABI (libnk_b200.so) against the oracle."""
import numpy as np
import pytest

from nomalise_kmers_multi_large_b200 import capi
from tests import engine_cases as ec
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu

CASES = [
    dict(k=15, canonical=False, depth=3, cap0=4099, paired=True, n_parts=2),
    dict(k=15, canonical=True, depth=2, cap0=2003, paired=False, n_parts=3),
    dict(k=25, canonical=True, depth=4, cap0=8191, paired=True, n_parts=1),
    dict(k=5, canonical=False, depth=2, cap0=1024, paired=True, n_parts=2),
    dict(k=31, canonical=False, depth=5, cap0=1009, paired=True, n_parts=2),
    dict(k=21, canonical=True, depth=12, cap0=3001, paired=True, n_parts=4),
    dict(k=7, canonical=True, depth=3, cap0=16384, paired=False, n_parts=2),
    # reads up to the 1023-base line limit (C:397): several packing iterations per read, more events per read than
    # two list chunks hold (global-atomic fallback of the chunked appends)
    dict(k=27, canonical=True, depth=3, cap0=60013, paired=True, n_parts=2, read_len=(400, 1023), genome_len=20000,
         records_per_step=24, n_seed_reads=30),
    dict(k=25, canonical=True, depth=12, cap0=1000003, paired=True, n_parts=8, genome_len=200000,
         records_per_step=400, read_len=(150, 150), n_seed_reads=2000),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"k{c['k']}c{int(c['canonical'])}d{c['depth']}cap{c['cap0']}")
@pytest.mark.parametrize("seed", [1, 2])
def test_engine_matches_oracle(cuda_lib, case, seed):
    kw = dict(steps=3, records_per_step=60)
    kw.update(case)
    info = ec.run_case(cuda_lib, seed=seed, **kw)
    assert info["ops"] > 0


def test_engine_is_deterministic_under_tight_scratch(cuda_lib):
    """Scratch lists sized to the bare minimum, many repeats: any lost or duplicated list entry (a race in the
    chunked appends) shows up as a differing score."""
    for rep in range(12):
        ec.run_case(cuda_lib, seed=1, k=15, canonical=False, depth=3, cap0=4099, paired=True, n_parts=2, steps=2,
                    records_per_step=60)


def test_engines_share_one_seed_table(cuda_lib):
    """several engines (streams) per GPU seed once: nkd_seed_finish_from"""
    ec.run_shared_seed_case(cuda_lib, device_clock=True)


def test_step_that_overfills_the_table(cuda_lib):
    """more new k-mers in one step than the table has free slots: shorter windows, growth in time, same results"""
    info = ec.run_case(cuda_lib, seed=126076854, k=7, canonical=False, depth=6, coverage=0.5, n_parts=1, cap0=257,
                       genome_len=6000, n_seed_reads=20, steps=4, records_per_step=40, paired=True, read_len=(100, 160), err=0.0)
    assert info["expansions"] >= 3


def test_engine_scratch_overflow_is_exact(cuda_lib, monkeypatch):
    monkeypatch.setenv("NKB200_OPEN_FRAC", "0.02")
    monkeypatch.setenv("NKB200_PEND_FRAC", "0.03")
    info = ec.run_case(cuda_lib, seed=3, k=15, canonical=True, depth=3, cap0=4099, n_parts=2, steps=2, records_per_step=80)
    assert info["ops"] > 0


@pytest.mark.parametrize("k,canonical", [(5, False), (15, False), (15, True), (25, True), (31, False), (31, True)])
def test_codec_matches_oracle(cuda_lib, k, canonical):
    """2-bit packing, window extraction and the reverse-complement minimum (C:1118-1126, C:1160-1180),
    N -> A (C:1406) and the alphabet gate (C:1144-1158), every read length from k to 1023."""
    rng = np.random.default_rng(k * 2 + int(canonical))
    lens = list(range(k, k + 40)) + [150, 151, 255, 256, 257, 511, 512, 513, 1000, 1023]
    seqs = []
    for L in lens:
        s = ec.BASES[rng.integers(0, 4, L)].copy()
        if L % 3 == 0:
            s[rng.integers(0, L)] = ord("N")
        if L % 7 == 0:
            s[: min(L, 40)] = ord("A")      # zero keys
        if L % 11 == 0:
            s[-min(L, 40):] = ord("T")      # zero keys under --canonical
        seqs.append(s.tobytes())
    bad = [3, 17]
    seqs[3] = seqs[3][:5] + b"a" + seqs[3][6:]
    seqs[17] = seqs[17][:-1] + b"R"
    buf, descs, nops = capi.pack_reads(seqs, None, k)
    with capi.Engine(k=k, canonical=canonical, depth_per_part=2, capacity0=1009, max_step_reads=len(seqs) + 8,
                     max_step_bytes=buf.size + 64, max_step_ops=nops[0] + 64, lib=cuda_lib) as eng:
        keys, inv = eng.extract_keys(buf, descs, nops[0])
    want = np.concatenate([ol.window_keys(s.replace(b"N", b"A"), k, canonical) for s in seqs])
    ok = np.ones(len(seqs), bool)
    ok[bad] = False
    for i, s in enumerate(seqs):
        if not ok[i]:
            continue
        a = int(descs["op_base"][i])
        n = len(s) - k + 1
        assert np.array_equal(keys[a:a + n], want[a:a + n]), f"read {i} len {len(s)}"
    assert list(np.flatnonzero(inv)) == bad


def test_engine_random_configurations(cuda_lib):
    """the first 60 configurations of the randomised engine stress (tests/engine_cases.random_case, same fixed seed as
    the emulation test): capacities from 257 slots, k 5..31, 1-5 partitions, ragged read lengths"""
    import random
    rnd = random.Random(20240823)
    for i in range(60):
        cfg = ec.random_case(rnd)
        rnd.randrange(1 << 30)   # the emulation test draws its launch-order seed here; keep the streams aligned
        try:
            ec.run_case(cuda_lib, **cfg)
        except Exception as e:
            raise AssertionError(f"case {i} {cfg}: {e!r}") from e


@pytest.mark.parametrize("entries", ["16", "65536"])
def test_hot_table_of_saturated_counters_is_exact_gpu(cuda_lib, monkeypatch, entries):
    """the same with the sm_100a kernels (k_probe_score's hot-table branch, k_hot_flush, k_hot_clear)"""
    monkeypatch.setenv("NKB200_HOT_ENTRIES", entries)
    info = ec.run_case(cuda_lib, seed=3, k=15, canonical=True, depth=3, cap0=1031, n_parts=2, genome_len=600, steps=4,
                       records_per_step=120)
    assert info["expansions"] >= 1
    raw = ec.run_raw_case(cuda_lib, seed=4, k=15, depth=2, cap0=257, n_parts=2, genome_len=500, steps=3)
    assert raw["processed"] > 0
