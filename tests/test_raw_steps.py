"""Steps as raw record text (nkd_stage_raw / nkd_fetch_raw, include/nk_b200.h): the worker loop's record reader,
length gate, scoring and printing (C:1605-1674, C:852-876) done on the device, against the oracle.  The same
scenarios run on the CPU emulation of the engine (host-side logic, per-record functions of nk_core.h) and, marked
gpu, on the sm_100a kernels."""
import ctypes as C

import numpy as np
import pytest

from nomalise_kmers_multi_large_b200 import capi
from tests import engine_cases as ec

CASES = {
    "fq_paired": dict(k=15, depth=3, n_parts=2),
    "fq_paired_canonical_growth": dict(k=25, canonical=True, depth=2, cap0=257, n_parts=3, read_len=(20, 150)),
    "fq_to_fa": dict(k=21, depth=4, n_parts=2, emit_mode=1, read_len=(15, 100)),
    "single_end": dict(k=17, depth=3, n_parts=2, paired=False),
    "single_end_fq_to_fa_prints_nothing": dict(k=17, depth=3, n_parts=1, paired=False, emit_mode=2),
    "fasta_in_out": dict(k=15, depth=5, n_parts=2, fastq=False, coverage=0.5),
    "k5_long_reads": dict(k=5, depth=40, n_parts=1, read_len=(5, 1000), records_per_step=30, genome_len=6000),
    "k31_one_record_steps": dict(k=31, depth=2, n_parts=4, records_per_step=1, steps=6, read_len=(25, 60)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_raw_steps_match_oracle_emu(emu_lib, name):
    t = ec.run_raw_case(emu_lib, seed=11, **CASES[name])
    assert t["processed"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_raw_steps_match_oracle_gpu(cuda_lib, name):
    for seed in (11, 12):
        t = ec.run_raw_case(cuda_lib, seed=seed, **CASES[name])
        assert t["processed"] > 0


def _declined(lib, text_f, text_r, n):
    eng = capi.Engine(k=15, depth_per_part=3, n_parts=1, capacity0=257, max_step_reads=64, max_step_bytes=1 << 12,
                      max_step_ops=1 << 14, max_raw_bytes=1 << 16, lib=lib)
    try:
        eng.seed_finish()
        with pytest.raises(capi.NkError) as e:
            eng.step_raw([(0, n, text_f, text_r)], True)
        return e.value.code
    finally:
        eng.close()


REC = b"@a\nACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIII\n"


def check_declined(lib):
    """text the record reader of the reference splits differently goes back to the host parser (NK_EIRREGULAR = -7);
    windows that do not hold the announced records are the caller's bug (NK_EINVAL)"""
    nul = REC.replace(b"GTAC", b"GT\0C", 1)
    long_line = b"@a\n" + b"ACGT" * 300 + b"\n+\n" + b"I" * 1200 + b"\n"
    assert _declined(lib, REC + nul, REC + REC, 2) == -7
    assert _declined(lib, REC + long_line, REC + REC, 2) == -7
    assert _declined(lib, REC + REC, REC + REC, 3) == -1            # fewer line ends than announced
    assert _declined(lib, REC + REC + b"@b\n", REC + REC, 2) == -1   # window does not end on its last record


def test_raw_text_the_device_declines_emu(emu_lib):
    check_declined(emu_lib)


@pytest.mark.gpu
def test_raw_text_the_device_declines_gpu(cuda_lib):
    check_declined(cuda_lib)


def check_seed_raw(lib, ahead=(None, None)):
    """nkd_seed_raw: the first `limit` records with a sequence longer than k are inserted with count 0
    (seed_kmer_hash, C:1322-1373); shorter ones do not count; the table equals the oracle's slot for slot"""
    from tests import oracle_lib as ol
    rng = np.random.default_rng(7)
    genome = ec.make_genome(rng, 4000)
    k, cap0 = 15, 1031
    eng = capi.Engine(k=k, canonical=True, depth_per_part=3, n_parts=1, capacity0=cap0, max_step_reads=1024,
                      max_step_bytes=1 << 12, max_step_ops=1 << 17, max_raw_bytes=1 << 18, lib=lib)
    try:
        otab = ol.OracleTable(cap0)
        total_taken = 0
        for (piece, limit), how in zip(((300, 120), (200, 1000)), ahead):      # the first piece holds more than `limit`, the second fewer
            seqs = [ec.sample_read(rng, genome, 5, 90, n_rate=0.05) for _ in range(piece)]
            text = b"".join(b"@s%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in enumerate(seqs))
            good = [s for s in seqs if len(s) > k][:limit]
            for s in good:
                otab.seed(s, k, True)
            taken, inv = eng.seed_raw(text, piece, limit, ahead=how)
            assert (taken, inv) == (len(good), -1)
            total_taken += taken
        st = eng.seed_stats()
        assert (st["capacity"], st["used"]) == (otab.cap, otab.used) and otab.cap > cap0   # the table grew while seeding
        ek, ec_ = eng.seed_export()
        okk, okc = otab.export()
        assert np.array_equal(ek, okk) and np.array_equal(ec_, okc)
        # a record that is not DNA is reported by its index in the piece
        bad = b"@a\n" + b"ACGT" * 8 + b"\n+\n" + b"I" * 32 + b"\n@b\nACGTACGTACGTACGTACGTnACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIII\n"
        assert eng.seed_raw(bad, 2, 10)[1] == 1
    finally:
        eng.close()


def test_seeding_from_raw_text_emu(emu_lib):
    check_seed_raw(emu_lib)


@pytest.mark.parametrize("ahead", [("use", "use"), ("use", "forget"), ("other", "use"), ("forget", "other")])
def test_seed_pieces_sent_ahead_emu(emu_lib, ahead):
    """nkd_upload_raw before nkd_seed_raw (the host pipeline sends the next piece while the previous one is inserted):
    a piece found in place, one taken back, and one overtaken by another buffer all give the same table"""
    check_seed_raw(emu_lib, ahead)


@pytest.mark.gpu
def test_seeding_from_raw_text_gpu(cuda_lib):
    check_seed_raw(cuda_lib)
