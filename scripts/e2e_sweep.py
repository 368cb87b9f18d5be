"""e2e tuning helper: extend_batch on host buffers for several chunk sizes / pack-thread counts."""
import sys, os, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
jobs, q, t = config2_jobs(n, seed=1)
ctx = B.KswB200(0)
cfg = B.make_cfg()
ref = None
out = np.zeros(n, dtype=B.RES_DT)
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
    for thr in (8, 16, 32):
        ctx.set_chunk_jobs(chunk); ctx.set_pack_threads(thr)
        ctx.extend_batch(cfg, jobs, q, t)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); r = ctx.extend_batch(cfg, jobs, q, t, out=out); ts.append(time.perf_counter() - t0)
        if ref is None: ref = r.copy()
        ok = all((r[f] == ref[f]).all() for f in B.RES_DT.names)
        print(f"chunk={chunk} threads={thr}: best {min(ts)*1e3:.1f} ms -> {n/min(ts)/1e6:.1f} M ext/s same={ok}", flush=True)
