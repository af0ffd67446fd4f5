"""Seeding alone on the benchmark workload (for profiling): seed both files, finish, report the time.
usage: python tools/seed_only.py [passes]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from nomalise_kmers_multi_large_b200 import Pipeline

w = bench.WORKLOADS["c2"]
pf, pr = bench.generate(w["pairs"], w["transcripts"], "c2")
fwd, rev = bench.map_file(pf), bench.map_file(pr)
out = bench.shm_dir() / "seed_only"
out.mkdir(exist_ok=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx = Pipeline(k=w["k"], depth=w["depth"], coverage=bench.COVERAGE, canonical=True, partitions=w["parts"],
                   memory_gb=0, n_forward_files=1, have_reverse=True, out_dir=out, devices=(0,))
    t = [time.perf_counter()]
    ctx.seed(fwd, bench.SEED_RECORDS); t.append(time.perf_counter())
    ctx.seed(rev, bench.SEED_RECORDS); t.append(time.perf_counter())
    ctx.seed_finish(); t.append(time.perf_counter())
    print(f"pass {i}: seed fwd {t[1]-t[0]:.3f} s, rev {t[2]-t[1]:.3f} s, finish {t[3]-t[2]:.3f} s", flush=True)
    ctx.close()
