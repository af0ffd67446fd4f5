"""CLI-level parity: run two command-line programs with the same argv in scratch directories and compare
every output_* file byte for byte plus the counters they print (the reference's observable contract,
SURVEY 8.B row (b))."""
import ctypes
import hashlib
import os
import re
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SYNTH_LIB = ROOT / "tools" / "libnk_synth.so"


def synth(tmp: Path, name: str, n_pairs: int, seed=1, transcripts=0, read_len=150, equal=False, fasta=False):
    """Write seeded synthetic paired files (tools/nk_synth.c); returns (fwd_path, rev_path)."""
    subprocess.run(["make", "-C", str(ROOT / "tools")], check=True, capture_output=True)
    L = ctypes.CDLL(str(SYNTH_LIB))
    f, r = ctypes.c_void_p(), ctypes.c_void_p()
    fs, rs = ctypes.c_size_t(), ctypes.c_size_t()
    L.nk_synth_generate.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                    ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
    L.nk_synth_free.argtypes = [ctypes.c_void_p]
    L.nk_synth_generate(n_pairs, seed, transcripts, read_len, int(equal), int(fasta), ctypes.byref(f), ctypes.byref(fs),
                        ctypes.byref(r), ctypes.byref(rs))
    ext = "fasta" if fasta else "fastq"
    pf, pr = tmp / f"{name}_1.{ext}", tmp / f"{name}_2.{ext}"
    pf.write_bytes(ctypes.string_at(f, fs.value))
    pr.write_bytes(ctypes.string_at(r, rs.value))
    L.nk_synth_free(f)
    L.nk_synth_free(r)
    return pf, pr


def mutate_lengths(path: Path, out: Path, seed=3, fastq=True):
    """Trim reads to variable lengths (some below k) so that the length gate and ragged windows are exercised."""
    rng = np.random.default_rng(seed)
    lines = path.read_bytes().split(b"\n")
    per = 4 if fastq else 2
    res = []
    for i in range(0, len(lines) - 1, per):
        rec = lines[i:i + per]
        keep = int(rng.integers(8, len(rec[1]) + 1))
        rec[1] = rec[1][:keep]
        if fastq:
            rec[3] = rec[3][:keep]
        res += rec
    out.write_bytes(b"\n".join(res) + b"\n")
    return out


COUNTER_RE = re.compile(r"Cumulative file statistics: Processed ([\d,]+), Printed ([\d,]+), Skipped ([\d,]+), "
                        r"Cumulative Max Unique Kmers in a thread: ([\d,]+)")


def run_cli(binary, args, cwd: Path, env=None, timeout=1800):
    cwd.mkdir(parents=True, exist_ok=True)
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([str(binary)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, env=e,
                       timeout=timeout)
    files = {f.name: hashlib.md5(f.read_bytes()).hexdigest() for f in sorted(cwd.glob("output_*"))}
    counters = [tuple(int(x.replace(",", "")) for x in m) for m in COUNTER_RE.findall(p.stdout)]
    final = {}
    for key in ("Processed Records", "Printed Records", "Skipped Records", "Cumulative Max unique kmers in any thread"):
        m = re.search(key + r": ([\d,]+)", p.stdout)
        if m:
            final[key] = int(m.group(1).replace(",", ""))
    return {"rc": p.returncode, "files": files, "counters": counters, "final": final, "stdout": p.stdout,
            "stderr": p.stderr}


def stdout_contract(res):
    """The program's stdout with the parts that depend on timing masked (rates, run times) and the per-thread lines in
    thread order (the reference prints them in completion order): every other line is part of the observable contract
    (SURVEY 8.B row b: 'Initial hash table size ...', 'Processing file pair ...', per-thread counts, cumulative file
    statistics, final report)."""
    lines = []
    for l in res["stdout"].splitlines():
        if l.startswith("B200:"):
            continue   # the product's own extra lines under -e
        l = re.sub(r"Processing rate: (?:[\d,]+|inf|-?nan) \([+-]?(?:[\d.]+|inf|-?nan)%\)", "Processing rate: R", l)
        l = re.sub(r"Total runtime: [\d.]+ seconds", "Total runtime: T seconds", l)
        l = re.sub(r"Overall processing rate: [\d,]+ ", "Overall processing rate: R ", l)
        lines.append(l)
    threads = sorted((l for l in lines if l.startswith("Thread ")), key=lambda l: int(l.split()[1]))
    it = iter(threads)
    return [next(it) if l.startswith("Thread ") else l for l in lines]


def assert_same_stdout(a, b, what=""):
    sa, sb = stdout_contract(a), stdout_contract(b)
    assert sa == sb, (what, "stdout differs", [(x, y) for x, y in zip(sa, sb) if x != y][:4], len(sa), len(sb))


def assert_same(a, b, what=""):
    assert a["rc"] == b["rc"], (what, "exit status", a["rc"], b["rc"], a["stderr"][-500:], b["stderr"][-500:])
    assert a["counters"] == b["counters"], (what, "counters", a["counters"], b["counters"])
    assert a["final"] == b["final"], (what, "final report", a["final"], b["final"])
    assert sorted(a["files"]) == sorted(b["files"]), (what, "file sets", sorted(a["files"]), sorted(b["files"]))
    diff = [n for n in a["files"] if a["files"][n] != b["files"][n]]
    assert not diff, (what, "files differ", diff[:10])
    assert a["files"], (what, "no outputs produced")


def standard_cases(tmp: Path, n_pairs: int):
    """(name, argv) pairs covering SURVEY 8.B rows a1-a12 and BASELINE configs 2-5 at test scale."""
    f, r = synth(tmp, "s", n_pairs, seed=1)
    f2, r2 = synth(tmp, "t", max(200, n_pairs // 2), seed=2, read_len=100)
    fe, re_ = synth(tmp, "e", n_pairs, seed=4, equal=True)
    fa, ra = synth(tmp, "a", max(200, n_pairs // 2), seed=5, fasta=True)
    fv = mutate_lengths(f, tmp / "v_1.fastq", seed=6)
    rv = mutate_lengths(r, tmp / "v_2.fastq", seed=7)
    return [
        ("canonical_p8", ["-f", f, "-r", r, "-k", 25, "-c", "-p", 8, "-d", 100, "-m", 1]),          # config 2 shape
        ("stranded_k31_fa_growth", ["-f", f, "-r", r, "-k", 31, "-g", 0.96, "-o", "fa", "-m", 1, "-p", 64, "-d", 256]),  # config 4
        ("dump_p4_k15", ["-f", f, "-r", r, "-k", 15, "-p", 4, "-d", 16, "-m", 1, "-P"]),
        ("equal_sizes_F6", ["-f", fe, "-r", re_, "-k", 21, "-p", 3, "-d", 12, "-m", 1]),
        ("single_end", ["-f", f, "-s", "-k", 17, "-p", 2, "-d", 8, "-m", 1]),
        ("single_end_fq2fa_empty", ["-f", f, "-s", "-k", 17, "-p", 1, "-d", 8, "-m", 1, "-o", "fa"]),
        ("fasta_in_out_mixed", ["-t", "fa", "-o", "fa", "-s", "-f", fa, ra, "-r", ra, "-k", 15, "-p", 2, "-d", 50, "-g", 0.5, "-m", 1, "-P"]),
        ("multi_file", ["-f", f, f2, "-r", r, r2, "-k", 25, "-c", "-p", 2, "-d", 40, "-m", 1]),
        ("ragged_lengths", ["-f", fv, "-r", rv, "-k", 25, "-p", 2, "-d", 8, "-m", 1, "-g", 1.0]),
        ("tiny_k5", ["-f", f2, "-r", r2, "-k", 5, "-p", 1, "-d", 400, "-m", 1]),
        ("one_partition_default_depth", ["-f", f2, "-r", r2, "-k", 20, "-m", 1]),
        ("p64_canonical_config3_shape", ["-f", f, "-r", r, "-k", 25, "-c", "-p", 64, "-d", 256, "-m", 1]),   # configs[2] flags
        ("depth_coverage_sweep_point", ["-f", f2, "-r", r2, "-k", 15, "-p", 4, "-d", 400, "-g", 0.96, "-m", 1, "-P"]),
    ]


def check_merged_extras(binary, oracle_cli, args, tmp: Path, env=None):
    """--merged-table / --merged-output (the reference leaves both to the user):
    the per-partition files stay byte-identical to the oracle's, the merged k-mer table is every dumped
    k-mer once, ascending, counts summed over the partitions, and the merged read files are the
    partitions' files concatenated in partition order."""
    want = run_cli(oracle_cli, list(args) + ["-P"], tmp / "oracle")
    got = run_cli(binary, list(args) + ["-P", "--merged-table", "--merged-output"], tmp / "got", env=env)
    assert got["rc"] == 0, got["stderr"][-500:]
    extra = sorted(set(got["files"]) - set(want["files"]))
    per_part = {n: h for n, h in got["files"].items() if n not in extra}
    assert per_part == want["files"], "per-partition files changed"
    out = tmp / "got"
    threads = lambda base: sorted(out.glob(base + ".*_thread*"), key=lambda f: int(re.search(r"_thread(\d+)", f.name).group(1)))
    sums = {}
    for f in threads("output_kmer"):
        for line in f.read_text().splitlines():
            kmer, count = line.split("\t")
            sums[kmer] = sums.get(kmer, 0) + int(count)
    merged = [f for f in extra if f.startswith("output_kmer_merged.")]
    assert len(merged) == 1, extra
    text = (out / merged[0]).read_text()
    assert text == "".join("%s\t%d\n" % kv for kv in sorted(sums.items())), "merged table differs"
    for base in ("output_forward", "output_reverse"):
        parts = threads(base)
        if not parts:
            continue
        name = parts[0].name.replace("_thread0.", ".")
        assert name in extra, (name, extra)
        assert (out / name).read_bytes() == b"".join(f.read_bytes() for f in parts), base
    return len(sums)
