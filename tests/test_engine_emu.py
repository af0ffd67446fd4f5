"""CPU checks of the parallel table algorithm: the per-operation code the sm_100a kernels run
(nk_core.h) and the step orchestration (nk_orchestrate.h), executed in shuffled order by the
test-only emulation backend, must equal the oracle's sequential table bit for bit."""
import os

import pytest

from tests import engine_cases as ec

CASES = [
    dict(k=15, canonical=False, depth=3, cap0=4099, paired=True, n_parts=2),
    dict(k=15, canonical=True, depth=2, cap0=2003, paired=False, n_parts=3),
    dict(k=25, canonical=True, depth=4, cap0=8191, paired=True, n_parts=1),
    dict(k=5, canonical=False, depth=2, cap0=1024, paired=True, n_parts=2),       # 4^k clamp (C:678-684)
    dict(k=31, canonical=False, depth=5, cap0=1009, paired=True, n_parts=2),      # growth inside steps
    dict(k=21, canonical=True, depth=12, cap0=3001, paired=True, n_parts=1),
    dict(k=7, canonical=True, depth=3, cap0=16384, paired=False, n_parts=2),
    # reads up to the 1023-base line limit (C:397): several packing iterations per read, more events per read than
    # two list chunks hold (global-atomic fallback of the chunked appends)
    dict(k=27, canonical=True, depth=3, cap0=60013, paired=True, n_parts=2, read_len=(400, 1023), genome_len=20000,
         records_per_step=24, n_seed_reads=30),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"k{c['k']}c{int(c['canonical'])}d{c['depth']}cap{c['cap0']}")
@pytest.mark.parametrize("seed", [1, 2])
def test_emulated_engine_matches_oracle(emu_lib, case, seed, monkeypatch):
    monkeypatch.setenv("NK_EMU_SEED", str(seed * 7919))
    kw = dict(steps=3, records_per_step=60)
    kw.update(case)
    info = ec.run_case(emu_lib, seed=seed, **kw)
    assert info["ops"] > 0


def test_emulated_engine_scratch_overflow_is_exact(emu_lib, monkeypatch):
    """Lists too small for a step: the run is undone (replayed with -1), the window halved, results unchanged."""
    monkeypatch.setenv("NKB200_OPEN_FRAC", "0.02")
    monkeypatch.setenv("NKB200_PEND_FRAC", "0.03")
    info = ec.run_case(emu_lib, seed=3, k=15, canonical=True, depth=3, cap0=4099, n_parts=2, steps=2, records_per_step=80)
    assert info["ops"] > 0


def test_emulated_engines_share_one_seed_table(emu_lib):
    ec.run_shared_seed_case(emu_lib)


def test_emulated_step_that_overfills_the_table(emu_lib, monkeypatch):
    """A step whose new k-mers outnumber the table's free slots: the speculative pass fills the table, walks hit the
    watchdog, and the step must fall back to shorter windows (growing the table in time) instead of failing."""
    monkeypatch.setenv("NK_EMU_SEED", "271928365")
    info = ec.run_case(emu_lib, seed=126076854, k=7, canonical=False, depth=6, coverage=0.5, n_parts=1, cap0=257,
                       genome_len=6000, n_seed_reads=20, steps=4, records_per_step=40, paired=True, read_len=(100, 160), err=0.0)
    assert info["expansions"] >= 3


def test_emulated_engine_random_configurations(emu_lib, monkeypatch):
    """a fixed-seed slice of tools/stress/stress_engine_emu.py: random k, depth, capacities from 257 slots, partitions,
    read shapes, list sizes and launch orders, each compared with the oracle after every step"""
    import random
    rnd = random.Random(20240823)
    for i in range(120):
        cfg = ec.random_case(rnd)
        monkeypatch.setenv("NK_EMU_SEED", str(rnd.randrange(1 << 30)))
        if i % 5 == 0:
            monkeypatch.setenv("NKB200_OPEN_FRAC", "0.05")
            monkeypatch.setenv("NKB200_PEND_FRAC", "0.1")
        else:
            monkeypatch.delenv("NKB200_OPEN_FRAC", raising=False)
            monkeypatch.delenv("NKB200_PEND_FRAC", raising=False)
        try:
            ec.run_case(emu_lib, **cfg)
        except Exception as e:
            raise AssertionError(f"case {i} {cfg}: {e!r}") from e


@pytest.mark.parametrize("entries", ["16", "4096"])
def test_hot_table_of_saturated_counters_is_exact(emu_lib, monkeypatch, entries):
    """NKB200_HOT_ENTRIES: home hits on saturated counters are kept as sums in a small side table and folded into the
    table's counts before anything reads them (growth, -P dump, export): every slot's count still equals the
    oracle's.  16 entries = constant collisions, 4096 = most hot k-mers fit."""
    monkeypatch.setenv("NKB200_HOT_ENTRIES", entries)
    info = ec.run_case(emu_lib, seed=3, k=15, canonical=True, depth=3, cap0=1031, n_parts=2, genome_len=600, steps=4,
                       records_per_step=120)
    assert info["expansions"] >= 1
    raw = ec.run_raw_case(emu_lib, seed=4, k=15, depth=2, cap0=257, n_parts=2, genome_len=500, steps=3)
    assert raw["processed"] > 0
