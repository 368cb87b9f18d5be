#!/usr/bin/env python
"""Development helper (GPU box): long-query jobs (beyond the s16x2 kernel's 512 columns) through the resident kernel
path — the warp-cooperative int32 kernel against the thread-per-job generic kernel (KSW_B200_DISABLE_WARP=1).
   python scripts/bench_long.py [jobs] [qlen]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ql = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
rng = np.random.default_rng(3)
qs, ts = [], []
for _ in range(n):
    L = int(rng.integers(ql // 2, ql + 1))
    t = rng.integers(0, 4, L + 60).astype(np.uint8)
    q = K.mutate(rng, t, 0.08, 0.02, max_indel=6)[:L]
    qs.append(q.astype(np.uint8)); ts.append(t)
b = K._pools_from_lists(qs, ts, rng.integers(19, 60, n), np.full(n, 100), K.make_cfg())
want, cells = K.run_oracle(b, threads=os.cpu_count(), want_cells=True)
ctx = B.KswB200(0)
rb = ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
ms = ctx.run_timed(rb, 5)[1:]
got = ctx.download(rb)
print(f"{n} jobs, qlen {ql // 2}..{ql}, {cells.sum() / 1e9:.2f} G visited cells, {rb.info()}: {ms.mean():.2f} ms -> "
      f"{cells.sum() / ms.mean() / 1e6:.1f} GCUPS; mismatch: {K.first_mismatch(want, got.view(K.RES_DT))}; DISABLE_WARP={os.environ.get('KSW_B200_DISABLE_WARP')}")
