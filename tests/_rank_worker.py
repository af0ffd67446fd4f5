"""Worker for tests/test_multirank_gloo.py: one rank of the N>1 path on CPU (gloo).  Each rank owns a contiguous
slice of the fixed -p partitions (as bench.py --gpus N does), runs the host pipeline on it, and the three counters
are summed with all_reduce -- the only cross-rank traffic the path has (C:1897-1909)."""
import ctypes
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nomalise_kmers_multi_large_b200 import capi  # noqa: E402
from nomalise_kmers_multi_large_b200.pipeline import Pipeline, count_chunk_lines, plan_ranges  # noqa: E402


def main():
    fwd_path, rev_path, out_dir, parts, k, depth, libpath = sys.argv[1:8]
    mode = sys.argv[8] if len(sys.argv) > 8 else "planned"
    parts, k, depth = int(parts), int(k), int(depth)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ctypes.CDLL(libpath)
    capi._declare_engine(lib)
    capi._declare_pipeline(lib)
    per = parts // world
    fwd, rev = np.fromfile(fwd_path, dtype=np.uint8), np.fromfile(rev_path, dtype=np.uint8)
    with Pipeline(k=k, depth=depth, canonical=True, partitions=parts, memory_gb=1, out_dir=out_dir, devices=(0,),
                  part_first=rank * per, part_count=per, step_pairs=64, lib=lib) as p:
        p.seed(fwd, 3000001)
        p.seed(rev, 3000001)
        p.seed_finish()
        if mode == "indexed":
            # as bench.py --gpus N does: every rank counts the line ends of its share of the files' chunks, the counts
            # are all-gathered, and each rank plans its own partitions from all of them
            counts = []
            for buf in (fwd, rev):
                mine, n_chunks, share = count_chunk_lines(buf, rank, world, threads=2, lib=lib)
                pad = torch.zeros(share, dtype=torch.int32)
                pad[:len(mine)] = torch.from_numpy(mine.view(np.int32))
                allc = [torch.empty(share, dtype=torch.int32) for _ in range(world)]
                dist.all_gather(allc, pad)
                counts.append(torch.cat(allc).numpy().view(np.uint32)[:n_chunks].copy())
            p.process_indexed(fwd, rev, counts[0], counts[1])
        else:
            # the plan (byte ranges, C:1796-1838) is computed once on rank 0 and broadcast
            plan = torch.zeros((4, parts), dtype=torch.int64)
            if rank == 0:
                plan.copy_(torch.from_numpy(plan_ranges(fwd, rev, parts, True, 2, lib=lib).view(np.int64)))
            dist.broadcast(plan, src=0)
            p.process_planned(fwd, rev, plan.numpy().view(np.uint64))
        p.finish()
        t = p.totals()
    sums = torch.tensor([t["processed"], t["printed"], t["skipped"]], dtype=torch.int64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    mx = torch.tensor([t["max_used"]], dtype=torch.int64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        Path(out_dir, "totals.json").write_text(json.dumps({"processed": int(sums[0]), "printed": int(sums[1]),
                                                            "skipped": int(sums[2]), "max_used": int(mx[0])}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
