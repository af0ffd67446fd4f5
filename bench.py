#!/usr/bin/env python
"""bench.py -- read pairs/s of the k-mer coverage-normalisation hot path on B200 (BASELINE.json metric).

A "step" is one pass of the path over the whole workload, from freshly seeded tables (seeding is redone,
untimed, before every step because scoring mutates the tables; the reference's own rate clock also starts
after seeding, C:2308).  The workload is BASELINE.json configs[1]: 10 M synthetic 150-base read pairs with
transcriptome-skewed coverage, -k 25 --canonical -p 8 -d 100, default table capacity (67,108,879 slots per
partition, growing x1.5), on one B200.  With --gpus N the same fixed 8 partitions are spread over N ranks
(a contiguous slice per rank, no data-path collective) -- strong scaling, results identical for every N.

  value   pairs / GPU-busy time of the pass: CUDA events around every device step; a GPU's engines (streams)
          overlap, so their step spans are merged (union) on the GPU clock.  The kernels read the step's
          sequence bytes from pinned host memory; with the bytes copied to HBM first (NKB200_NO_ZEROCOPY=1)
          the device time is the same (profiles/README.md)
  e2e     pairs / wall time of nk_process_paired on HOST buffers (partitioning, record indexing, pinned
          staging, H2D, kernels, D2H, writing the accepted records)
  roofline  k_probe: (20 B x slots it visits + sequence bytes + 1 B/pair) / its CUDA-event time vs measured HBM peak;
          roofline.isolated = the same from one extra untimed pass in which k_probe has the GPU to itself
  cpu_baseline / --impl reference  the reference binary (oracle/_ref) on the box's host cores, same flags,
          bounded sample, timed from its own per-thread completion lines with the sleep(1) stagger removed.
"""
import argparse
import ctypes
import json
import mmap
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
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K, DEPTH, COVERAGE, PARTS, READ_LEN, TRANSCRIPTS, SEED = 25, 100, 0.9, 8, 150, 20000, 1
SEED_RECORDS = 1 + 3_000_000  # 1 + 3e6 / forward_file_count, C:2242
CLASSES = ["probe", "open", "apply", "classify", "sort_rank", "commit", "decide", "growth_undo", "parse", "emit"]


def shm_dir():
    d = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    d = d / "nkb200_bench"
    d.mkdir(parents=True, exist_ok=True)
    return d


def generate(n_pairs, tag):
    """Seeded synthetic dataset (tools/nk_synth.c); files on tmpfs so every rank can map them."""
    subprocess.run(["make", "-C", str(ROOT / "tools")], check=True, capture_output=True)
    d = shm_dir()
    pf, pr = d / f"{tag}_{n_pairs}_1.fastq", d / f"{tag}_{n_pairs}_2.fastq"
    done = d / f"{tag}_{n_pairs}.done"
    if not done.exists():
        subprocess.run([str(ROOT / "tools" / "nk_synth"), "-n", str(n_pairs), "-s", str(SEED), "-t", str(TRANSCRIPTS),
                        "-L", str(READ_LEN), "-o", str(d / f"{tag}_{n_pairs}")], check=True, capture_output=True)
        done.write_text("ok")
    return pf, pr


def map_file(path):
    return np.memmap(path, dtype=np.uint8, mode="r")


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


def time_reference(n_pairs):
    """Run the reference (thread-local canonical buffer variant) with the benchmark's flags on the first
    n_pairs of the workload.  Its threads start 1 s apart (sleep(1), C:1879); each prints a completion line, so a
    pty gives per-thread work times W_t = finish_t - (T0 + t).  rate = n_pairs / max_t W_t."""
    binary, kind = reference_binary()
    pf, pr = generate(n_pairs, "sample")
    work = Path(tempfile.mkdtemp(prefix="ref_", dir=shm_dir()))
    argv = [str(binary), "-f", str(pf), "-r", str(pr), "-k", str(K), "-c", "-p", str(PARTS), "-d", str(DEPTH)]
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
    shutil.rmtree(work, ignore_errors=True)
    t0 = next((t for t, l in stamps if l.startswith("Processing file pair")), None)
    done = {}
    for t, l in stamps:
        m = re.match(r"Thread (\d+) - (Processing rate|processed)", l)
        if m:
            done[int(m.group(1))] = t
    processed = next((int(re.sub(r"[^\d]", "", l.split(":")[1])) for _, l in stamps if l.startswith("Processed Records")), 0)
    if t0 is None or len(done) < PARTS or processed == 0:
        raise RuntimeError("could not parse the reference's output:\n" + "\n".join(l for _, l in stamps[-20:]))
    if kind == "reference":
        work_s = [done[t] - (t0 + t) for t in range(PARTS)]      # thread t is created t seconds after T0
        threads = PARTS
    else:                                                       # the port runs partitions back to back, one thread
        order = sorted(done.values())
        work_s = [sum(b - a for a, b in zip([t0] + order[:-1], order))]
        threads = 1
    proc_s = max(work_s)
    return {"value": processed / proc_s, "unit": "pairs/s", "cores": threads, "kind": kind,
            "sample": f"first {n_pairs} pairs of the workload, -k {K} -c -p {PARTS} -d {DEPTH}, default capacity; "
                      f"processing phase only (max per-thread work time {proc_s:.2f} s, sleep(1) stagger removed); "
                      f"seeding took {t0 - t_launch:.1f} s single-threaded, whole run {t_end - t_launch:.1f} s",
            "seed_s": t0 - t_launch, "process_s": proc_s, "pairs": processed}


# ------------------------------------------------------------------ our arm

def dist_setup(n_gpus):
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


def run_ours(args):
    import torch
    from nomalise_kmers_multi_large_b200 import Pipeline, capi, plan_ranges
    rank, world, local, dist = dist_setup(args.gpus)
    if world > 1 and "NKB200_THREADS" not in os.environ:   # ranks share the box's host cores
        os.environ["NKB200_THREADS"] = str(max(2, (os.cpu_count() or 2) // world))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    capi.load_library()
    if rank == 0:
        generate(args.pairs, "bench")
    if dist is not None:
        dist.barrier()
    pf, pr = generate(args.pairs, "bench")
    fwd, rev = map_file(pf), map_file(pr)
    # blocked placement: rank r owns partitions [r*P/N, (r+1)*P/N); placement cannot change results (README:68)
    assert PARTS % world == 0, "--gpus must divide the fixed partition count"
    per_rank = PARTS // world
    out_dir = Path(tempfile.mkdtemp(prefix=f"out_r{rank}_", dir=shm_dir()))
    sampler = ClockSampler(local) if rank == 0 else None
    steps = []
    def one_pass(timed):
        ctxs = [Pipeline(k=K, depth=DEPTH, coverage=COVERAGE, canonical=True, partitions=PARTS,
                         memory_gb=args.memory, n_forward_files=1, have_reverse=True, out_dir=out_dir,
                         devices=(local,), part_first=rank * per_rank, part_count=per_rank)]
        t_seed = time.perf_counter()
        for c in ctxs:
            c.seed(fwd, SEED_RECORDS)
            c.seed(rev, SEED_RECORDS)
            c.seed_finish()
        seed_s = time.perf_counter() - t_seed
        barrier(dist, local)
        if sampler:
            sampler.active = timed
        t0 = time.perf_counter()
        if world > 1:
            # the byte ranges are computed once (rank 0, all host cores) and handed to the ranks: 256 bytes of control
            # data, the counterpart of the reference's main thread computing them before it starts its workers
            plan = torch.zeros((4, PARTS), dtype=torch.int64, device=f"cuda:{local}")
            if rank == 0:
                plan.copy_(torch.from_numpy(plan_ranges(fwd, rev, PARTS, True, os.cpu_count() or 0).view(np.int64)))
            dist.broadcast(plan, src=0)
            plan_np = plan.cpu().numpy().view(np.uint64)
        for c in ctxs:
            if world > 1:
                c.process_planned(fwd, rev, plan_np)
            else:
                c.process_paired(fwd, rev)
            c.finish()
        torch.cuda.synchronize(local)
        wall = time.perf_counter() - t0
        if sampler:
            sampler.active = False
        barrier(dist, local)
        tot = [c.totals() for c in ctxs]
        agg = {k: sum(t[k] for t in tot) for k in tot[0] if k != "class_ms"}
        for i, name in enumerate(CLASSES):
            agg["ms_" + name] = sum(t["class_ms"][i] for t in tot)
        agg["wall_s"], agg["seed_s"] = wall, seed_s
        for c in ctxs:
            c.close()
        return agg

    for it in range(args.warmup + args.steps):
        agg = one_pass(it >= args.warmup)
        if it >= args.warmup:
            steps.append(agg)
    # k_probe timed without the GPU's other engines running beside it: one extra, untimed pass with one engine per GPU
    isolated = None
    if not args.no_isolated_probe:
        prev = os.environ.get("NKB200_ENGINES_PER_GPU")
        os.environ["NKB200_ENGINES_PER_GPU"] = "1"
        isolated = one_pass(False)
        if prev is None:
            del os.environ["NKB200_ENGINES_PER_GPU"]
        else:
            os.environ["NKB200_ENGINES_PER_GPU"] = prev
    if sampler:
        sampler.stop()
    shutil.rmtree(out_dir, ignore_errors=True)
    n = n_ = len(steps)
    # per-step maxima over ranks (device time and wall), sums of the counted quantities
    dev_ms = sum(all_max(dist, local, s["run_ms"]) for s in steps)
    wall_s = sum(all_max(dist, local, s["wall_s"]) for s in steps)
    keys = ["processed", "printed", "skipped", "launches", "probe_launches", "ops", "touches", "probe_touches",
            "slow_events", "expansions", "h2d_bytes", "d2h_bytes", "probe_ms", "index_seconds", "device_seconds",
            "write_seconds", "seed_s", "pend_events", "open_ops", "engines"] + ["ms_" + n for n in CLASSES]
    sums = dict(zip(keys, all_sum(dist, local, [sum(s[k] for s in steps) for k in keys])))
    ikeys = ["probe_ms", "probe_launches", "probe_touches", "processed", "run_ms"]
    iso = dict(zip(ikeys, all_sum(dist, local, [isolated[k] for k in ikeys]))) if isolated else None
    if rank != 0:
        return
    pairs = sums["processed"] / n
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # k_probe's algorithmic bytes: 16-B entry read + 4-B count write per visited slot, the sequence bytes, 1 B/pair
    probe_bytes = 20.0 * sums["probe_touches"] + (2 * READ_LEN + 1) * sums["processed"]
    probe_s = sums["probe_ms"] / 1e3 / max(world, 1)      # ranks run concurrently: average per-rank kernel time
    achieved = probe_bytes / probe_s / 1e9 / max(world, 1) if probe_s > 0 else 0.0   # per GPU
    traffic = None
    tf = ROOT / "profiles" / "probe_traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
    line = {
        "metric": "read pairs/sec (k-mer coverage normalisation, processing phase)",
        "value": pairs / (dev_ms / n / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / n, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.pairs} synthetic 150bp PE pairs, transcriptome-skewed (20000 transcripts, lognormal "
                               f"sigma 2, 0.5% errors), -k {K} --canonical -p {PARTS} -d {DEPTH} -g {COVERAGE}, "
                               + ("default capacity 67108879 slots/partition" if not args.memory else f"-m {args.memory}"),
                   "partitions": PARTS, "partitions_per_gpu": PARTS / world, "pairs": int(pairs),
                   "l2": "tables (GBs per partition) and step inputs far exceed the 126 MB L2; no flush needed",
                   "seeding": "redone untimed before every step (the reference's rate clock starts after seeding, C:2308)",
                   "seed_s_per_step": sums["seed_s"] / n / max(world, 1),
                   "engines_per_gpu": sums["engines"] / n / max(world, 1),
                   "device_time": "per GPU, the union of its engines' step spans on the GPU clock (CUDA events); "
                                  "max over ranks"},
        "e2e": {"value": pairs / (wall_s / n), "unit": "pairs/s", "h2d_bytes_per_step": sums["h2d_bytes"] / n,
                "d2h_bytes_per_step": sums["d2h_bytes"] / n, "ms_per_step": wall_s / n * 1e3,
                "host_s_per_step": {"index": sums["index_seconds"] / n, "device_calls": sums["device_seconds"] / n,
                                    "write": sums["write_seconds"] / n}},
        "gpu_launches": int(sums["launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_probe", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": probe_bytes / max(sums["probe_launches"], 1),
                     "avg_launch_ms": sums["probe_ms"] / max(sums["probe_launches"], 1),
                     # share of the kernels' stream time (engines overlap, so not of the GPU-busy time)
                     "share_of_step": sums["probe_ms"] / max(sum(sums["ms_" + c] for c in CLASSES), 1e-9),
                     "touches_per_op": sums["touches"] / max(sums["ops"], 1),
                     "note": "avg_launch_ms is measured while the GPU's other engines run their kernels concurrently"},
        "kernel_ms_per_step": {n: sums["ms_" + n] / n_ / max(world, 1) for n in CLASSES},
        "clocks": sampler.summary() if sampler else None,
        "counters": {"printed": sums["printed"] / n, "skipped": sums["skipped"] / n, "ops": sums["ops"] / n,
                     "touches": sums["touches"] / n, "slow_events": sums["slow_events"] / n,
                     "pending_list_entries": sums["pend_events"] / n, "open_list_entries": sums["open_ops"] / n,
                     "expansions_in_scoring": sums["expansions"] / n},
    }
    if iso and iso["probe_ms"] > 0:
        ibytes = 20.0 * iso["probe_touches"] + (2 * READ_LEN + 1) * iso["processed"]
        iach = ibytes / (iso["probe_ms"] / 1e3) / 1e9      # summed over ranks on both sides: per-GPU rate
        line["roofline"]["isolated"] = {
            "achieved": iach, "frac": iach / peak, "avg_launch_ms": iso["probe_ms"] / max(iso["probe_launches"], 1),
            "algorithmic_bytes_per_launch": ibytes / max(iso["probe_launches"], 1),
            "device_ms": iso["run_ms"] / max(world, 1), "share_of_step": iso["probe_ms"] / max(iso["run_ms"], 1e-9),
            "how": "one extra untimed pass with one engine per GPU (NKB200_ENGINES_PER_GPU=1): k_probe has the GPU to itself"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = time_reference(args.sample_pairs)
        except Exception as e:  # the baseline is a reported figure; its absence must not void the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    runs = []
    for it in range(args.warmup + args.steps):
        n = args.sample_pairs if it >= args.warmup else max(20000, args.sample_pairs // 5)
        r = time_reference(n)
        if it >= args.warmup:
            runs.append(r)
    pairs = sum(r["pairs"] for r in runs)
    secs = sum(r["process_s"] for r in runs)
    base = dict(runs[-1])
    base["value"] = pairs / secs
    line = {"impl": "reference", "metric": "read pairs/sec (k-mer coverage normalisation, processing phase)",
            "value": pairs / secs, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / len(runs) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"bounded sample: first {args.sample_pairs} pairs of the {args.pairs}-pair workload per step, "
                                   f"-k {K} --canonical -p {PARTS} -d {DEPTH} -g {COVERAGE}, default capacity; the reference's "
                                   f"pthreads path on the host cores ({os.cpu_count()} visible, {base['cores']} used = -p)",
                       "partitions": PARTS},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": pairs / secs, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--sample-pairs", type=int, default=100_000)
    ap.add_argument("--memory", type=int, default=0, help="-m for the tables (0 = reference default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-isolated-probe", action="store_true",
                    help="skip the extra untimed pass that times k_probe with one engine per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
