#!/usr/bin/env python
"""bench.py -- read pairs/s of the k-mer coverage-normalisation hot path on B200 (BASELINE.json metric).

A "step" is one pass of the path over the whole workload, from freshly seeded tables (seeding is redone,
untimed, before every step because scoring mutates the tables; the reference's own rate clock also starts
after seeding, C:2308).  Default workload = BASELINE.json configs[1] ("c2"): 10 M synthetic 150-base read pairs
with transcriptome-skewed coverage, -k 25 --canonical -p 8 -d 100, default table capacity (67,108,879 slots per
partition, growing x1.5), on one B200.  With --gpus N the same fixed partitions are spread over N ranks (a
contiguous slice per rank, no data-path collective; the ranks count the files' line ends together and
all-gather the counts -- 100 KB of control data) -- strong scaling, results identical for every N.
--workload c3 = configs[2]'s flags (-k 25 -c -p 64 -d 256) on --pairs pairs.

  value   pairs / GPU-busy time of the pass: CUDA events around every device step (record parsing, scoring,
          assembling the accepted records' text); a GPU's engines (streams) overlap, so their step spans are
          merged (union) on the GPU clock.  Inputs are resident in HBM when a step's span starts.
  e2e     pairs / wall time of nk_process_paired on HOST buffers (the mmap'd files): planning, staging copies,
          H2D, kernels, D2H, write() of the accepted records.  Bytes are those of the timed region only.
  roofline  k_probe: (20 B x slots it visits + sequence bytes + 1 B/pair) / its CUDA-event time vs measured HBM peak;
          roofline.isolated = the same from one extra untimed pass in which k_probe has the GPU to itself
  golden  before anything is timed, one pass's output files are hashed and compared with the reference's own
          output on this very workload (tests/golden/bench_*.json, made by tests/golden/make_bench_golden.py);
          a mismatch aborts the run instead of printing a number.
  cpu_baseline   the reference binary (oracle/_ref) on the host cores, same flags, bounded sample (first 100 k pairs)
  --impl reference  the reference binary on the FULL workload, once per invocation (the timing is reused for
          every --steps; a run takes minutes), processing phase timed from its own per-thread completion lines
          with the sleep(1) stagger removed; its seeding time and "Total runtime" are reported beside it.
"""
import argparse
import hashlib
import json
import os
import pty
import re
import select
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

COVERAGE, READ_LEN, SEED = 0.9, 150, 1
WORKLOADS = {
    "c2": dict(pairs=10_000_000, k=25, depth=100, parts=8, transcripts=20000,
               what="BASELINE.json configs[1]"),
    "c3": dict(pairs=40_000_000, k=25, depth=256, parts=64, transcripts=80000,
               what="BASELINE.json configs[2] flags; 200 M pairs in the original, --pairs sets the size"),
}
SEED_RECORDS = 1 + 3_000_000  # 1 + 3e6 / forward_file_count, C:2242
CLASSES = ["probe", "open", "apply", "classify", "sort_rank", "commit", "decide", "growth_undo", "parse", "emit"]
METRIC = "read pairs/sec (k-mer coverage normalisation, processing phase)"
# measured on this pool's boxes with profiles/microbench/hostio_bench.cu (profiles/r02_hostio.txt)
PCIE_H2D_GBS, HOST_SCAN_GBS = 55.0, 116.0


def workload_of(args):
    w = dict(WORKLOADS[args.workload])
    if args.pairs:
        w["pairs"] = args.pairs
    w["memory"] = args.memory
    w["flags"] = f"-k {w['k']} -c -p {w['parts']} -d {w['depth']} -g {COVERAGE}" + (f" -m {args.memory}" if args.memory else "")
    return w


def config_of(w):
    """identical in both arms: the workload, nothing about how an arm ran it"""
    return {"workload": f"{w['pairs']} synthetic {READ_LEN}bp PE pairs, transcriptome-skewed ({w['transcripts']} transcripts, "
                        f"lognormal sigma 2, 0.5% errors; tools/nk_synth -s {SEED}), {w['flags']}, "
                        + ("default capacity 67108879 slots/partition" if not w["memory"] else f"-m {w['memory']}")
                        + f" ({w['what']})",
            "pairs": w["pairs"], "partitions": w["parts"], "flags": w["flags"]}


def golden_of(args, w):
    path = ROOT / "tests" / "golden" / f"bench_{args.workload}_{w['pairs'] // 1_000_000}M.json"
    if path.exists() and not w["memory"] and w["pairs"] % 1_000_000 == 0:
        return path, json.loads(path.read_text())
    return path, None


def shm_dir():
    d = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    d = d / "nkb200_bench"
    d.mkdir(parents=True, exist_ok=True)
    return d


def generate(n_pairs, transcripts, tag):
    """Seeded synthetic dataset (tools/nk_synth.c); files on tmpfs so every rank can map them."""
    subprocess.run(["make", "-C", str(ROOT / "tools")], check=True, capture_output=True)
    d = shm_dir()
    stem = f"{tag}_{n_pairs}_{transcripts}"
    pf, pr = d / f"{stem}_1.fastq", d / f"{stem}_2.fastq"
    done = d / f"{stem}.done"
    if not done.exists():
        synth = str(ROOT / "tools" / "nk_synth")
        if n_pairs <= 20_000_000:
            subprocess.run([synth, "-n", str(n_pairs), "-s", str(SEED), "-t", str(transcripts), "-L", str(READ_LEN),
                            "-o", str(d / stem)], check=True, capture_output=True)
        else:
            # large workloads: 10 M-pair blocks generated side by side (block b uses seed SEED + b over the same
            # transcript count) and concatenated in block order; still a pure function of (n_pairs, transcripts)
            blocks = [(b, min(10_000_000, n_pairs - b * 10_000_000)) for b in range((n_pairs + 9_999_999) // 10_000_000)]
            procs = [subprocess.Popen([synth, "-n", str(n), "-s", str(SEED + b), "-t", str(transcripts), "-L", str(READ_LEN),
                                       "-o", str(d / f"{stem}.b{b}")], stdout=subprocess.DEVNULL) for b, n in blocks]
            assert all(p.wait() == 0 for p in procs), "nk_synth failed"
            for mate, dst in (("1", pf), ("2", pr)):
                with open(dst, "wb") as out:
                    for b, _ in blocks:
                        part = d / f"{stem}.b{b}_{mate}.fastq"
                        with open(part, "rb") as src:
                            shutil.copyfileobj(src, out, 1 << 24)
                        part.unlink()
        done.write_text("ok")
    return pf, pr


def map_file(path):
    return np.memmap(path, dtype=np.uint8, mode="r")


def md5_of(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.active = [], None, False
        self.gpu = gpu_index
        if shutil.which("nvidia-smi"):
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            if self.active:
                self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ the reference on host cores

def reference_binary():
    tls, port = ROOT / "oracle" / "_ref" / "nkml_tls", ROOT / "oracle" / "nk_oracle"
    if tls.exists():
        return tls, "reference"
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "oracle"], check=True, capture_output=True)
    return port, "port"


def time_reference(w, n_pairs, golden=None):
    """Run the reference (thread-local canonical buffer variant) with the workload's flags on its first
    n_pairs pairs (all of them for --impl reference).  Its threads start 1 s apart (sleep(1), C:1879); each prints a
    completion line, so a pty gives per-thread work times W_t = finish_t - (T0 + t).  rate = pairs / max_t W_t."""
    binary, kind = reference_binary()
    parts = w["parts"]
    pf, pr = generate(n_pairs, w["transcripts"], "bench" if n_pairs == w["pairs"] else "sample")
    work = Path(tempfile.mkdtemp(prefix="ref_", dir=shm_dir()))
    argv = [str(binary), "-f", str(pf), "-r", str(pr), "-k", str(w["k"]), "-c", "-p", str(parts), "-d", str(w["depth"])]
    if w["memory"]:
        argv += ["-m", str(w["memory"])]
    master, slave = pty.openpty()
    t_launch = time.perf_counter()
    p = subprocess.Popen(argv, cwd=work, stdout=slave, stderr=subprocess.DEVNULL)
    os.close(slave)
    stamps, buf = [], b""
    while True:
        r, _, _ = select.select([master], [], [], 1.0)
        if r:
            try:
                chunk = os.read(master, 65536)
            except OSError:
                chunk = b""
            if not chunk:
                break
            now = time.perf_counter()
            buf += chunk
            while b"\n" in buf:
                line, buf = buf.split(b"\n", 1)
                stamps.append((now, line.decode(errors="replace").strip()))
        elif p.poll() is not None:
            break
    p.wait()
    t_end = time.perf_counter()
    os.close(master)
    verified = None
    if golden is not None and kind == "reference":
        names = sorted(f.name for f in work.glob("output_*"))
        with ThreadPoolExecutor(8) as ex:
            got = dict(zip(names, ex.map(lambda n: md5_of(work / n), names)))
        verified = got == golden["files_md5"]
    shutil.rmtree(work, ignore_errors=True)
    t0 = next((t for t, l in stamps if l.startswith("Processing file pair")), None)
    done = {}
    for t, l in stamps:
        m = re.match(r"Thread (\d+) - (Processing rate|processed)", l)
        if m:
            done[int(m.group(1))] = t

    def num(key):
        for _, l in stamps:
            if l.startswith(key):
                m = re.search(r"([\d,]+(?:\.\d+)?)", l.split(":", 1)[1])
                return float(m.group(1).replace(",", "")) if m else 0.0
        return 0.0
    processed = int(num("Processed Records"))
    own_runtime = num("Total runtime")
    if t0 is None or len(done) < parts or processed == 0:
        raise RuntimeError("could not parse the reference's output:\n" + "\n".join(l for _, l in stamps[-20:]))
    if kind == "reference":
        work_s = [done[t] - (t0 + t) for t in range(parts)]      # thread t is created t seconds after T0
        threads = parts
    else:                                                       # the port runs partitions back to back, one thread
        order = sorted(done.values())
        work_s = [sum(b - a for a, b in zip([t0] + order[:-1], order))]
        threads = 1
    proc_s = max(work_s)
    whole = "the whole workload" if n_pairs == w["pairs"] else f"first {n_pairs} pairs of the workload"
    return {"value": processed / proc_s, "unit": "pairs/s", "cores": threads, "kind": kind,
            "sample": f"{whole}, {w['flags']}, default capacity; processing phase only (max per-thread work time "
                      f"{proc_s:.2f} s, sleep(1) stagger removed); seeding took {t0 - t_launch:.1f} s single-threaded, "
                      f"whole run {t_end - t_launch:.1f} s wall, the reference's own 'Total runtime' {own_runtime:.0f} s "
                      f"(includes {parts} s of sleep), {os.cpu_count()} host cores visible",
            "seed_s": t0 - t_launch, "process_s": proc_s, "pairs": processed, "wall_s": t_end - t_launch,
            "own_total_runtime_s": own_runtime, "outputs_match_golden": verified}


# ------------------------------------------------------------------ our arm

def dist_setup():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def all_max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(dist, local, xs):
    if dist is None:
        return list(xs)
    import torch
    t = torch.tensor(list(xs), dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def barrier(dist, local):
    import torch
    torch.cuda.synchronize(local)
    if dist is not None:
        dist.barrier()


def probe_traffic(algorithmic_bytes_per_launch):
    """ncu dram__bytes of one k_probe launch of the shape that ran (profiles/probe_traffic.json lists the captured
    shapes by their algorithmic bytes); None when no capture is within 15 % of this run's launches"""
    tf = ROOT / "profiles" / "probe_traffic.json"
    if not tf.exists() or algorithmic_bytes_per_launch <= 0:
        return None
    best = None
    for shape in json.loads(tf.read_text()).get("shapes", []):
        a = shape.get("algorithmic_bytes_per_launch")
        if a and abs(a - algorithmic_bytes_per_launch) / algorithmic_bytes_per_launch < 0.15:
            if best is None or abs(a - algorithmic_bytes_per_launch) < abs(best[0] - algorithmic_bytes_per_launch):
                best = (a, shape.get("dram_bytes_per_launch"))
    return best[1] if best else None


def run_ours(args):
    import torch
    from nomalise_kmers_multi_large_b200 import Pipeline, capi, count_chunk_lines
    w = workload_of(args)
    parts = w["parts"]
    rank, world, local, dist = dist_setup()
    if world > 1 and "NKB200_THREADS" not in os.environ:   # ranks share the box's host cores
        os.environ["NKB200_THREADS"] = str(max(2, (os.cpu_count() or 2) // world))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    capi.load_library()
    if rank == 0:
        generate(w["pairs"], w["transcripts"], "bench")
    if dist is not None:
        dist.barrier()
    pf, pr = generate(w["pairs"], w["transcripts"], "bench")
    fwd, rev = map_file(pf), map_file(pr)
    # blocked placement: rank r owns partitions [r*P/N, (r+1)*P/N); placement cannot change results (README:68)
    assert parts % world == 0, "--gpus must divide the fixed partition count"
    per_rank = parts // world
    first_part = rank * per_rank
    if args.slice:   # development aid: this process works on a slice of the partitions only (what one rank of a larger launch sees)
        first_part, per_rank = (int(x) for x in args.slice.split(","))
    out_dir = Path(tempfile.mkdtemp(prefix=f"out_r{rank}_", dir=shm_dir()))
    sampler = ClockSampler(local) if rank == 0 else None
    gold_path, golden = golden_of(args, w)
    steps = []

    def one_pass(timed, verify=False):
        ctx = Pipeline(k=w["k"], depth=w["depth"], coverage=COVERAGE, canonical=True, partitions=parts,
                       memory_gb=w["memory"], n_forward_files=1, have_reverse=True, out_dir=out_dir,
                       devices=(local,), part_first=first_part, part_count=per_rank)
        t_seed = time.perf_counter()
        ctx.seed(fwd, SEED_RECORDS)
        ctx.seed(rev, SEED_RECORDS)
        ctx.seed_finish()
        seed_s = time.perf_counter() - t_seed
        before = ctx.totals()
        barrier(dist, local)
        if sampler:
            sampler.active = timed
        t0 = time.perf_counter()
        t_plan = 0.0
        if world > 1:
            # every rank counts the line ends of its share of the files' chunks; the counts (4 bytes per 256 KB of
            # input) are all-gathered, and each rank plans its own partitions from all of them (nk_process_indexed)
            counts = []
            for buf in (fwd, rev):
                mine, n_chunks, share = count_chunk_lines(buf, rank, world)
                pad = torch.zeros(share, dtype=torch.int32, device=f"cuda:{local}")
                pad[:len(mine)] = torch.from_numpy(mine.view(np.int32)).to(pad.device)
                allc = torch.empty(share * world, dtype=torch.int32, device=pad.device)
                dist.all_gather_into_tensor(allc, pad)
                counts.append(allc.cpu().numpy().view(np.uint32)[:n_chunks].copy())
            t_plan = time.perf_counter() - t0
            ctx.process_indexed(fwd, rev, counts[0], counts[1])
        else:
            ctx.process_paired(fwd, rev)
        ctx.finish()
        torch.cuda.synchronize(local)
        wall = time.perf_counter() - t0
        if sampler:
            sampler.active = False
        barrier(dist, local)
        tot = ctx.totals()
        agg = {k: v for k, v in tot.items() if k != "class_ms"}
        for i, name in enumerate(CLASSES):
            agg["ms_" + name] = tot["class_ms"][i]
        agg["h2d_bytes"] -= before["h2d_bytes"]       # seeding traffic is outside the timed region
        agg["d2h_bytes"] -= before["d2h_bytes"]
        agg["wall_s"], agg["seed_s"], agg["count_s"] = wall, seed_s, t_plan
        ctx.close()
        ok = True
        if verify and golden is not None:
            names = [f"output_{m}.k{w['k']}_norm{w['depth'] // parts}_thread{t}.fastq" for m in ("forward", "reverse")
                     for t in range(first_part, first_part + per_rank)]
            with ThreadPoolExecutor(8) as ex:
                got = dict(zip(names, ex.map(lambda n: md5_of(out_dir / n), names)))
            bad = [n for n in names if got[n] != golden["files_md5"].get(n)]
            ok = not bad
            if bad:
                print(f"[bench] rank {rank}: outputs differ from the reference's ({gold_path.name}): {bad[:4]}", file=sys.stderr)
        return agg, ok

    # untimed: one pass whose outputs are compared with the reference's own output on this workload
    first, ok = one_pass(False, verify=True)
    all_ok = all_sum(dist, local, [0.0 if ok else 1.0])[0] == 0.0
    printed = all_sum(dist, local, [first["printed"]])[0]
    if golden is not None and (not all_ok or (not args.slice and int(printed) != golden["final"]["Printed Records"])):
        if rank == 0:
            print(json.dumps({"error": "outputs differ from the reference's on this workload; no value reported",
                              "golden": gold_path.name, "printed": printed}), flush=True)
        sys.exit(1)
    for it in range(1, args.warmup + args.steps):
        agg, _ = one_pass(it >= args.warmup)
        if it >= args.warmup:
            steps.append(agg)
    if args.warmup == 0:
        steps.insert(0, first)
    # k_probe timed without the GPU's other engines running beside it: one extra, untimed pass with one engine per GPU
    isolated = None
    if not args.no_isolated_probe:
        prev = os.environ.get("NKB200_ENGINES_PER_GPU")
        os.environ["NKB200_ENGINES_PER_GPU"] = "1"
        isolated, _ = one_pass(False)
        if prev is None:
            del os.environ["NKB200_ENGINES_PER_GPU"]
        else:
            os.environ["NKB200_ENGINES_PER_GPU"] = prev
    if sampler:
        sampler.stop()
    shutil.rmtree(out_dir, ignore_errors=True)
    steps = steps[-args.steps:]
    n = n_ = len(steps)
    # per-step maxima over ranks (device time and wall), sums of the counted quantities
    dev_ms = sum(all_max(dist, local, s["run_ms"]) for s in steps)
    wall_s = sum(all_max(dist, local, s["wall_s"]) for s in steps)
    keys = ["processed", "printed", "skipped", "launches", "probe_launches", "ops", "touches", "probe_touches",
            "slow_events", "expansions", "h2d_bytes", "d2h_bytes", "probe_ms", "index_seconds", "device_seconds",
            "write_seconds", "seed_s", "count_s", "hot_hits", "pend_events", "open_ops", "engines", "raw_steps", "parsed_steps"] + ["ms_" + c for c in CLASSES]
    sums = dict(zip(keys, all_sum(dist, local, [sum(s[k] for s in steps) for k in keys])))
    ikeys = ["probe_ms", "probe_launches", "probe_touches", "processed", "run_ms"] + ["ms_" + c for c in CLASSES]
    iso = dict(zip(ikeys, all_sum(dist, local, [isolated[k] for k in ikeys]))) if isolated else None
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pairs = sums["processed"] / n
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # k_probe's algorithmic bytes: 16-B entry read + 4-B count write per visited slot, the sequence bytes, 1 B/pair
    probe_bytes = 20.0 * sums["probe_touches"] + (2 * READ_LEN + 1) * sums["processed"]
    probe_s = sums["probe_ms"] / 1e3 / max(world, 1)      # ranks run concurrently: average per-rank kernel time
    achieved = probe_bytes / probe_s / 1e9 / max(world, 1) if probe_s > 0 else 0.0   # per GPU
    alg_per_launch = probe_bytes / max(sums["probe_launches"], 1)
    input_bytes = float(fwd.size + rev.size)
    e2e_s = wall_s / n
    line = {
        "metric": METRIC,
        "value": pairs / (dev_ms / n / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / n, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_of(w),
        "details": {"partitions_per_gpu": parts / world,
                    "l2": "tables (GBs per partition) and step inputs far exceed the 126 MB L2; no flush needed",
                    "seeding": "redone untimed before every step (the reference's rate clock starts after seeding, C:2308)",
                    "seed_s_per_step": sums["seed_s"] / n / max(world, 1),
                    "engines_per_gpu": sums["engines"] / n / max(world, 1),
                    "device_steps": {"raw_text": sums["raw_steps"] / n, "host_parsed": sums["parsed_steps"] / n},
                    "device_time": "per GPU, the union of its engines' step spans on the GPU clock (CUDA events); max over ranks",
                    "golden": (f"outputs of an untimed pass equal the reference's on this workload ({gold_path.name})"
                               if golden is not None else "no reference output recorded for this workload size")},
        "e2e": {"value": pairs / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": sums["h2d_bytes"] / n,
                "d2h_bytes_per_step": sums["d2h_bytes"] / n, "ms_per_step": e2e_s * 1e3,
                # per rank: busy time of the slowest engine's pipeline stages (they overlap), averaged over the ranks
                "host_s_per_step": {"count_lines_allgather": sums["count_s"] / n / world,
                                    "stage_copy": sums["index_seconds"] / n / world,
                                    "device_calls": sums["device_seconds"] / n / world,
                                    "write": sums["write_seconds"] / n / world}},
        # H2D ingest against what PCIe and the host can do (north_star; SURVEY 8(d) "ingest bound")
        "ingest": {"h2d_gbs": sums["h2d_bytes"] / n / e2e_s / 1e9, "h2d_gbs_per_gpu": sums["h2d_bytes"] / n / e2e_s / 1e9 / world,
                   "pcie_peak_gbs": PCIE_H2D_GBS, "frac_of_pcie_per_gpu": sums["h2d_bytes"] / n / e2e_s / 1e9 / world / PCIE_H2D_GBS,
                   "host_read_gbs": input_bytes / e2e_s / 1e9, "host_scan_peak_gbs": HOST_SCAN_GBS,
                   "d2h_gbs": sums["d2h_bytes"] / n / e2e_s / 1e9,
                   "note": "pcie_peak = pinned cudaMemcpyAsync on this pool's boxes (Gen5 x16), host_scan_peak = AVX2 line-end count "
                           "over the page cache on 16 cores; both from profiles/r02_hostio.txt"},
        "gpu_launches": int(sums["launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_probe", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": probe_traffic(alg_per_launch), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_per_launch,
                     "avg_launch_ms": sums["probe_ms"] / max(sums["probe_launches"], 1),
                     # share of the kernels' stream time (engines overlap, so not of the GPU-busy time)
                     "share_of_step": sums["probe_ms"] / max(sum(sums["ms_" + c] for c in CLASSES), 1e-9),
                     "touches_per_op": sums["touches"] / max(sums["ops"], 1),
                     "note": "avg_launch_ms is measured while the GPU's other engines run their kernels concurrently"},
        "kernel_ms_per_step": {c: sums["ms_" + c] / n_ / max(world, 1) for c in CLASSES},
        "clocks": sampler.summary() if sampler else None,
        "counters": {"printed": sums["printed"] / n, "skipped": sums["skipped"] / n, "ops": sums["ops"] / n,
                     "touches": sums["touches"] / n, "slow_events": sums["slow_events"] / n,
                     "pending_list_entries": sums["pend_events"] / n, "open_list_entries": sums["open_ops"] / n,
                     "expansions_in_scoring": sums["expansions"] / n,
                     "hot_table_hits": sums["hot_hits"] / n},
    }
    if iso and iso["probe_ms"] > 0:
        ibytes = 20.0 * iso["probe_touches"] + (2 * READ_LEN + 1) * iso["processed"]
        iach = ibytes / (iso["probe_ms"] / 1e3) / 1e9      # summed over ranks on both sides: per-GPU rate
        ialg = ibytes / max(iso["probe_launches"], 1)
        line["roofline"]["isolated"] = {
            "achieved": iach, "frac": iach / peak, "avg_launch_ms": iso["probe_ms"] / max(iso["probe_launches"], 1),
            "algorithmic_bytes_per_launch": ialg, "traffic": probe_traffic(ialg),
            "device_ms": iso["run_ms"] / max(world, 1), "share_of_step": iso["probe_ms"] / max(iso["run_ms"], 1e-9),
            "kernel_ms": {c: iso["ms_" + c] / max(world, 1) for c in CLASSES},
            "how": "one extra untimed pass with one engine per GPU (NKB200_ENGINES_PER_GPU=1): k_probe has the GPU to itself"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = time_reference(w, args.sample_pairs)
        except Exception as e:  # the baseline is a reported figure; its absence must not void the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w = workload_of(args)
    _, golden = golden_of(args, w)
    pairs = args.ref_pairs or w["pairs"]
    r = time_reference(w, pairs, golden if pairs == w["pairs"] else None)   # once: every step would repeat the same minutes
    cfg = config_of(w)
    if pairs != w["pairs"]:
        cfg = dict(cfg, workload=f"bounded sample: first {pairs} pairs of: " + cfg["workload"], pairs=pairs)
    line = {"impl": "reference", "metric": METRIC,
            "value": r["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["process_s"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cfg,
            "details": {"measured_runs": 1,
                        "why_one_run": "one full-size run of the reference takes minutes (single-threaded seeding of 6 M records "
                                       "alone); its timing stands for every requested step",
                        "processing_s": r["process_s"], "seeding_s": r["seed_s"], "wall_s": r["wall_s"],
                        "reference_own_total_runtime_s": r["own_total_runtime_s"],
                        "outputs_match_golden": r["outputs_match_golden"], "host_cores_visible": os.cpu_count(),
                        "threads": r["cores"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override the workload's size")
    ap.add_argument("--sample-pairs", type=int, default=100_000, help="size of the cpu_baseline sample of the b200 arm")
    ap.add_argument("--ref-pairs", type=int, default=0, help="--impl reference on a prefix instead of the whole workload")
    ap.add_argument("--memory", type=int, default=0, help="-m for the tables (0 = reference default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slice", default="", help="first,count: work on these partitions only (development aid, not a bench line)")
    ap.add_argument("--no-isolated-probe", action="store_true",
                    help="skip the extra untimed pass that times k_probe with one engine per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
