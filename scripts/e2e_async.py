#!/usr/bin/env python
"""Development helper (GPU box): times the asynchronous pinned-buffer entry (device packing) against the host-packing
entry and the resident kernels on config 2.   python scripts/e2e_async.py [jobs] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
jobs, qpool, tpool = config2_jobs(n, seed=12345)
ctx = B.KswB200(0)
cfg = B.make_cfg()
rb = ctx.upload(cfg, jobs, qpool, tpool)
ms = ctx.run_timed(rb, 4)
cells = float(ctx.download_cells(rb).astype(np.int64).sum())
ref = ctx.download(rb)
rb.free()
print(f"resident kernels: {ms[1:].mean():.2f} ms -> {cells / ms[1:].mean() / 1e6:.0f} GCUPS")
pj, pq, pt = B.pinned_copy(jobs), B.pinned_copy(qpool), B.pinned_copy(tpool)
pr = B.PinnedArray(n, B.RES_DT)
for name in ("async", "host"):
    ts = []
    for s in range(steps + 1):
        pr.a[:] = 0
        t0 = time.perf_counter()
        if name == "async":
            ctx.extend_batch_async(cfg, pj.a, pq.a, pt.a, pr.a); ctx.wait()
        else:
            ctx.extend_batch(cfg, jobs, qpool, tpool, out=pr.a)
        ts.append(time.perf_counter() - t0)
    same = all((pr.a[f] == ref[f]).all() for f in B.RES_DT.names)
    t = float(np.mean(ts[1:]))
    print(f"{name}: {[round(x * 1e3, 1) for x in ts]} ms -> {cells / t / 1e9:.0f} GCUPS e2e, h2d/d2h {ctx.last_transfer()}, identical to resident: {same}")
