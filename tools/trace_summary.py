#!/usr/bin/env python
"""Summarise an NKB200_TRACE file (engine, step, stage, start, end): per pass and engine, the busy time of every
pipeline stage, the span they cover, and the time the GPU stage spent waiting between its calls."""
import collections
import sys

rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
rows = [(int(e), int(s), st, float(a), float(b)) for e, s, st, a, b in rows]
passes, cur = [], []
for r in rows:
    cur.append(r)
    if r[2] == "pipelines":
        passes.append(cur)
        cur = []
for i, p in enumerate(passes):
    plan = next(r for r in p if r[2] == "plan")
    pipe = next(r for r in p if r[2] == "pipelines")
    t0 = plan[3]
    print(f"pass {i}: plan {1e3 * (plan[4] - plan[3]):.1f} ms, whole {1e3 * (pipe[4] - t0):.1f} ms")
    eng = collections.defaultdict(lambda: collections.defaultdict(list))
    for e, s, st, a, b in p:
        if e >= 0:
            eng[e][st].append((a - t0, b - t0))
    for e in sorted(eng):
        parts = []
        for st in ("cut", "copy", "stage", "run", "fetch", "d2h_wait", "write"):
            iv = eng[e].get(st, [])
            if iv:
                parts.append(f"{st} {1e3 * sum(b - a for a, b in iv):.0f} [{1e3 * iv[0][0]:.0f}..{1e3 * iv[-1][1]:.0f}]")
        gpu = sorted(eng[e].get("stage", []) + eng[e].get("run", []) + eng[e].get("fetch", []))
        idle = sum(max(0.0, gpu[j + 1][0] - gpu[j][1]) for j in range(len(gpu) - 1))
        print(f"  engine {e}: " + ", ".join(parts) + f", gpu-stage waits {1e3 * idle:.0f} ms")
