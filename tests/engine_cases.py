"""Shared scenarios: drive a device engine (CUDA build or the test-only emulation) and the oracle's
sequential table with the same reads and compare everything observable: accept flags in record
order, the whole table (keys AND counts of every slot, ghost counts included), used, capacity."""
import numpy as np

from nomalise_kmers_multi_large_b200 import capi
from tests import oracle_lib as ol

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(rng, n, repeat_frac=0.2):
    g = BASES[rng.integers(0, 4, n)].copy()
    # low-complexity stretches give zero keys (poly-A; poly-T under --canonical) and hot k-mers
    for _ in range(int(n * repeat_frac / 40)):
        p = int(rng.integers(0, n - 40))
        g[p:p + 40] = BASES[int(rng.integers(0, 4))]
    return g


def sample_read(rng, genome, lo, hi, err=0.01, n_rate=0.01):
    L = int(rng.integers(lo, hi + 1))
    p = int(rng.integers(0, len(genome) - L))
    r = genome[p:p + L].copy()
    e = rng.random(L) < err
    r[e] = BASES[rng.integers(0, 4, int(e.sum()))]
    if rng.random() < n_rate:
        r[int(rng.integers(0, L))] = ord("N")
    return r.tobytes()


def revcomp(s: bytes) -> bytes:
    return s.translate(bytes.maketrans(b"ACGTN", b"TGCAN"))[::-1]


def table_text(keys, counts, k):
    """print_kmer_table's output for a slot-ordered table (C:354-385)"""
    out = []
    for key, c in zip(keys.tolist(), counts.tolist()):
        if key:
            out.append("".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k)) + "\t%d\n" % c)
    return "".join(out).encode()


def check_dumps(eng, otabs, k):
    """device-formatted -P dump, compaction and the merged table against the oracle's tables"""
    sums = {}
    for p, ot in enumerate(otabs):
        okk, okc = ot.export()
        want = table_text(okk, okc, k)
        assert eng.dump_text(p, 0, ot.cap) == want, f"dump text differs, part {p}"
        cut = (ot.cap // 3) | 1  # ranges need not be tile-aligned
        assert eng.dump_text(p, 0, cut) + eng.dump_text(p, cut, ot.cap - cut) == want, f"ranged dump differs, part {p}"
        ck, cc_ = eng.compact(p, ot.used)
        nz = okk != 0
        assert np.array_equal(ck, okk[nz]) and np.array_equal(cc_, okc[nz].astype(np.int64)), f"compaction differs, part {p}"
        for key, c in zip(okk[nz].tolist(), okc[nz].tolist()):
            sums[key] = sums.get(key, 0) + c
    # merged table: all partitions, plus partition 0 once more through the host path, with large counts
    k0, c0 = otabs[0].export()
    nz = k0 != 0
    big = c0[nz].astype(np.int64) + (1 << 33)
    for key, c in zip(k0[nz].tolist(), big.tolist()):
        sums[key] += c
    total = sum(ot.used for ot in otabs) + int(nz.sum())
    n = eng.merge(range(len(otabs)), extra=(k0[nz], big), max_entries=total)
    assert n == len(sums)
    ks = np.array(sorted(sums), np.uint64)
    want = table_text(ks, np.array([sums[x] for x in ks.tolist()], np.int64), k)
    assert eng.dump_text(capi.PART_MERGED, 0, n) == want, "merged table differs"


def run_case(lib, *, seed=1, k=15, canonical=False, depth=3, coverage=0.9, n_parts=2, cap0=4099, genome_len=3000,
             n_seed_reads=200, steps=4, records_per_step=150, paired=True, read_len=(40, 120), err=0.01,
             max_step_ops=None):
    rng = np.random.default_rng(seed)
    genome = make_genome(rng, genome_len)
    lo, hi = read_len
    lo = max(lo, k)  # records shorter than k never reach the engine (C:1430-1443)
    ops_bound = (records_per_step * n_parts * 2 + n_seed_reads) * (hi - k + 1) + 64
    eng = capi.Engine(k=k, canonical=canonical, depth_per_part=depth, coverage=coverage, n_parts=n_parts,
                      capacity0=cap0, max_step_reads=4 * (records_per_step * n_parts * 2 + n_seed_reads) + 8,
                      max_step_bytes=(records_per_step * n_parts * 2 + n_seed_reads) * (hi + 32) + 64,
                      max_step_ops=max_step_ops or ops_bound, lib=lib)
    try:
        # ---- seeding: one shared table, count 0 (C:1322-1373, C:1501-1537)
        seed_reads = [sample_read(rng, genome, max(lo, k + 1), hi, err) for _ in range(n_seed_reads)]
        otab = ol.OracleTable(cap0)
        half = n_seed_reads // 2
        for chunk in (seed_reads[:half], seed_reads[half:]):
            if chunk:
                buf, descs, _ = capi.pack_reads(chunk, None, k)
                eng.seed_step(buf, descs)
            for s in chunk:
                otab.seed(s, k, canonical)
        st = eng.seed_stats()
        assert (st["capacity"], st["used"]) == (otab.cap, otab.used), ("seed table", st, otab.cap, otab.used)
        ek, ec = eng.seed_export()
        okk, okc = otab.export()
        assert np.array_equal(ek, okk), "seed keys differ"
        assert np.array_equal(ec, okc), "seed counts differ"
        assert eng.dump_text(capi.PART_SEED, 0, otab.cap) == table_text(okk, okc, k), "seed dump text differs"
        eng.seed_finish()
        otabs = [otab.clone() for _ in range(n_parts)]
        # ---- steps
        info = {"slow_events": 0, "expansions": 0, "ops": 0, "touches": 0}
        for step in range(steps):
            seqs, parts, per_part = [], [], []
            for p in range(n_parts):
                recs = []
                for _ in range(records_per_step):
                    f = sample_read(rng, genome, lo, hi, err)
                    if paired:
                        recs.append((f, revcomp(sample_read(rng, genome, lo, hi, err))))
                        seqs += [recs[-1][0], recs[-1][1]]
                        parts += [p, p]
                    else:
                        recs.append(f)
                        seqs.append(f)
                        parts.append(p)
                per_part.append(recs)
            buf, descs, _ = capi.pack_reads(seqs, parts, k)
            acc, inv = eng.step(buf, descs, paired)
            assert inv == -1
            got_hi, got_tot = eng.read_scores(len(seqs))
            # sequence_to_hash's (high, total) of every read, in the reference's order (C:1459-1499, C:1559-1563)
            want_scores = []
            for p in range(n_parts):
                for rec in per_part[p]:
                    for mate in (rec if paired else (rec,)):
                        want_scores.append(otabs[p].score(mate, k, canonical, depth))
            got_scores = list(zip(got_hi.tolist(), got_tot.tolist()))
            bad = [i for i in range(len(seqs)) if got_scores[i] != want_scores[i]]
            assert not bad, f"(high,total) differ at step {step}: reads {bad[:6]} got {[got_scores[i] for i in bad[:6]]} want {[want_scores[i] for i in bad[:6]]}"
            stride = 2 if paired else 1
            want = np.array([int(all(ol.keep_mate(h, t, coverage) for h, t in want_scores[i * stride:(i + 1) * stride]))
                             for i in range(len(seqs) // stride)], dtype=np.uint8)
            assert np.array_equal(acc, want), f"accept flags differ at step {step}: {np.flatnonzero(acc != want)[:10]}"
            for p in range(n_parts):
                st = eng.part_stats(p)
                assert (st["capacity"], st["used"]) == (otabs[p].cap, otabs[p].used), (step, p, st)
                ek, ec = eng.export(p)
                okk, okc = otabs[p].export()
                assert np.array_equal(ek, okk), f"keys differ step {step} part {p}"
                bad = np.flatnonzero(ec != okc)
                assert bad.size == 0, f"counts differ step {step} part {p}: slots {bad[:8]} got {ec[bad[:8]]} want {okc[bad[:8]]}"
        check_dumps(eng, otabs, k)
        for p in range(n_parts):
            st = eng.part_stats(p)
            info["slow_events"] += st["slow_events"]
            info["expansions"] += st["expansions"]
            info["ops"] += st["ops"]
            info["touches"] += st["touches"]
            assert st["ops"] == otabs[p].ops, ("ops", st["ops"], otabs[p].ops)  # clones count from 0
            assert st["touches"] == otabs[p].touches, ("touches", st["touches"], otabs[p].touches)
        return info
    finally:
        eng.close()


def run_shared_seed_case(lib, *, k=15, canonical=True, depth=3, cap0=4099, seed=5, device_clock=False):
    """Two engines on one GPU: the second takes its partitions from the first one's seed table
    (nkd_seed_finish_from) and must then behave exactly like an engine that seeded itself."""
    rng = np.random.default_rng(seed)
    genome = make_genome(rng, 2500)
    seed_reads = [sample_read(rng, genome, 40, 100) for _ in range(150)]
    mk = lambda: capi.Engine(k=k, canonical=canonical, depth_per_part=depth, coverage=0.9, n_parts=2, capacity0=cap0,
                             max_step_reads=4096, max_step_bytes=1 << 20, max_step_ops=1 << 20, lib=lib)
    lead, follower = mk(), mk()
    try:
        buf, descs, _ = capi.pack_reads(seed_reads, None, k)
        lead.seed_step(buf, descs)
        otab = ol.OracleTable(cap0)
        for s_ in seed_reads:
            otab.seed(s_, k, canonical)
        follower.seed_finish_from(lead)
        lead.seed_finish()
        otabs = [otab.clone() for _ in range(2)]
        seqs, parts, want = [], [], []
        for p in range(2):
            for _ in range(100):
                f, r = sample_read(rng, genome, 40, 100), revcomp(sample_read(rng, genome, 40, 100))
                seqs += [f, r]
                parts += [p, p]
                scores = [otabs[p].score(m, k, canonical, depth) for m in (f, r)]  # both mates always touch the table
                want.append(int(all(ol.keep_mate(h, t, 0.9) for h, t in scores)))
        buf, descs, _ = capi.pack_reads(seqs, parts, k)
        for eng in (follower, lead):
            acc, inv = eng.step(buf, descs, True)
            assert inv == -1 and acc.tolist() == want
            for p in range(2):
                ek, ec_ = eng.export(p)
                okk, okc = otabs[p].export()
                assert np.array_equal(ek, okk) and np.array_equal(ec_, okc)
        spans = follower.run_spans()   # one (start, end) per scoring step on the GPU's clock; the emulation has no clock
        assert spans.shape == ((1, 2) if device_clock else (0, 2))
        assert not device_clock or spans[0, 1] > spans[0, 0] >= 0
    finally:
        follower.close()
        lead.close()


def random_case(rnd):
    """One random engine configuration (tools/stress/stress_engine_emu.py and the randomised emu test share it)."""
    k = rnd.choice([5, 7, 11, 15, 21, 25, 31])
    return dict(seed=rnd.randrange(1 << 30), k=k, canonical=rnd.random() < 0.5, depth=rnd.choice([2, 3, 4, 6, 12, 40]),
                coverage=rnd.choice([0.5, 0.9, 0.96, 1.0]), n_parts=rnd.choice([1, 2, 3, 5]),
                cap0=rnd.choice([257, 1031, 4099, 16411, 65537]), genome_len=rnd.choice([400, 1500, 6000]),
                n_seed_reads=rnd.choice([0, 20, 200]), steps=rnd.choice([1, 2, 4]), records_per_step=rnd.choice([5, 40, 150]),
                paired=rnd.random() < 0.7, read_len=rnd.choice([(k, k + 3), (40, 120), (100, 160), (k, 300)]),
                err=rnd.choice([0.0, 0.01, 0.05]))


def fasta_of(header: bytes, seq: bytes, tag: bytes) -> bytes:
    """fastq_to_fasta (C:852-876) on the already N->A scrubbed sequence"""
    h = b">" + header[1:]
    if len(h) < 2 or not h.endswith(b"/" + tag):
        h += b"/" + tag
    return h + b"\n" + seq + b"\n"


def run_raw_case(lib, *, seed=1, k=15, canonical=False, depth=3, coverage=0.9, n_parts=2, cap0=4099, genome_len=3000,
                 steps=3, records_per_step=120, paired=True, fastq=True, emit_mode=0, read_len=(10, 120), err=0.01):
    """Steps handed over as raw record text (nkd_stage_raw / nkd_fetch_raw): the device finds the lines, applies
    the length gate, scores and assembles the accepted records' text.  Compared with the oracle's sequential
    table driven by the worker loop's rules (C:1605-1674): text of every partition and mate, processed / printed,
    and every slot of every table after every step."""
    rng = np.random.default_rng(seed)
    genome = make_genome(rng, genome_len)
    lo, hi = read_len
    per_rec = 2 * (hi + 40) + 8
    eng = capi.Engine(k=k, canonical=canonical, depth_per_part=depth, coverage=coverage, n_parts=n_parts, capacity0=cap0,
                      max_step_reads=2 * records_per_step * n_parts + 64, max_step_bytes=1 << 16,
                      max_step_ops=2 * records_per_step * n_parts * (hi + 1) + 4096,
                      max_raw_bytes=2 * records_per_step * n_parts * per_rec + 4096, lib=lib)
    try:
        eng.seed_finish()
        otabs = [ol.OracleTable(cap0) for _ in range(n_parts)]
        serial = 0
        totals = {"processed": 0, "printed": 0, "dropped": 0}
        for step in range(steps):
            windows, want = [], []
            for p in range(n_parts):
                n = int(rng.integers(1, records_per_step + 1))
                texts, recs = [[], []], []
                for _ in range(n):
                    mates = []
                    for m in range(2 if paired else 1):
                        s = sample_read(rng, genome, lo, hi, err, n_rate=0.05)
                        if m:
                            s = revcomp(s)
                        serial += 1
                        name = b"r%d %s" % (serial, b"x" * int(rng.integers(0, 9)))
                        if rng.random() < 0.5:
                            name += b"/%d" % (m + 1)   # fastq_to_fasta keeps an existing /1 or /2
                        lead = b"@" if fastq else b">"
                        rec = lead + name + b"\n" + s + b"\n" + ((b"+\n" + b"I" * len(s) + b"\n") if fastq else b"")
                        texts[m].append(rec)
                        mates.append((lead + name, s))
                    recs.append(mates)
                windows.append((p, n, b"".join(texts[0]), b"".join(texts[1]) if paired else None))
                out = [[], []]
                processed = printed = 0
                for mates in recs:
                    if any(len(s) < k for _, s in mates):
                        totals["dropped"] += 1
                        continue  # the pair vanishes: no counter, no table access (C:1430-1443)
                    scores = [otabs[p].score(s, k, canonical, depth) for _, s in mates]
                    processed += 1
                    if all(ol.keep_mate(h, t, coverage) for h, t in scores):
                        printed += 1
                        for m, (hdr, s) in enumerate(mates):
                            s2 = s.replace(b"N", b"A")
                            if emit_mode == 0:
                                out[m].append(hdr + b"\n" + s2 + b"\n" + ((b"+\n" + b"I" * len(s) + b"\n") if fastq else b""))
                            elif emit_mode == 1:
                                out[m].append(fasta_of(hdr, s2, b"%d" % (m + 1)))
                want.append((b"".join(out[0]), b"".join(out[1]), processed, printed))
                totals["processed"] += processed
                totals["printed"] += printed
            got, inv = eng.step_raw(windows, paired, lines_per_record=4 if fastq else 2, emit_mode=emit_mode)
            assert inv == -1
            for p in range(n_parts):
                assert got[p][2:] == want[p][2:], (step, p, "processed/printed", got[p][2:], want[p][2:])
                assert got[p][0] == want[p][0], (step, p, "forward text differs")
                assert got[p][1] == want[p][1], (step, p, "reverse text differs")
                ek, ec = eng.export(p)
                okk, okc = otabs[p].export()
                assert np.array_equal(ek, okk), f"keys differ step {step} part {p}"
                assert np.array_equal(ec, okc), f"counts differ step {step} part {p}"
        return totals
    finally:
        eng.close()
