#!/usr/bin/env python
"""Development helper (GPU box): the asynchronous pinned-buffer entry on config 2 under several settings of the
pipeline's knobs (host-lane staging sets, its own upload stream, chunk schedule, pack threads), one context per
setting in one process.   python scripts/e2e_knobs.py [jobs] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
WARM = 6          # the lanes' pace estimates settle over the first calls of a context
jobs, qpool, tpool = config2_jobs(n, seed=12345)
cfg = B.make_cfg()
ctx = B.KswB200(0)
rb = ctx.upload(cfg, jobs, qpool, tpool)
ms = ctx.run_timed(rb, 4)
ref = ctx.download(rb)
rb.free()
ctx.close()
print(f"resident kernels: {ms[1:].mean():.2f} ms", flush=True)
pj, pq, pt = B.pinned_copy(jobs), B.pinned_copy(qpool), B.pinned_copy(tpool)
pr = B.PinnedArray(n, B.RES_DT)
KEYS = ("KSW_B200_PRE_PRIO", "KSW_B200_ASYNC_CTAS", "KSW_B200_FAST_CTAS", "KSW_B200_BLOCKSYNC", "KSW_B200_HSLOTS", "KSW_B200_HUP", "KSW_B200_CHUNK", "KSW_B200_LEAD", "KSW_B200_HYBRID", "KSW_B200_PACK_WORDS")
SETTINGS = [
    {},
    {"KSW_B200_ASYNC_CTAS": "12"},
    {"KSW_B200_ASYNC_CTAS": "0"},
    {"KSW_B200_CHUNK": "524288"},
    {"KSW_B200_HYBRID": "0"},
    {},
]
if len(sys.argv) > 3:
    import json
    SETTINGS = json.load(open(sys.argv[3])) if os.path.exists(sys.argv[3]) else json.loads(sys.argv[3])
for st in SETTINGS:
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in st.items():
        if k.startswith("KSW_"):
            os.environ[k] = v
    c = B.KswB200(0, pack_threads=int(st["pack_threads"])) if "pack_threads" in st else B.KswB200(0)
    ts = []
    for s in range(steps + WARM):
        pr.a[:] = 0
        t0 = time.perf_counter()
        c.extend_batch_async(cfg, pj.a, pq.a, pt.a, pr.a); c.wait()
        ts.append(time.perf_counter() - t0)
    same = all((pr.a[f] == ref[f]).all() for f in B.RES_DT.names)
    t = np.array(ts[WARM:]) * 1e3
    print(f"{st}: mean {t.mean():.2f} ms min {t.min():.2f} max {t.max():.2f}  h2d {c.last_transfer()[0] / 1e9:.3f} GB  identical: {same}", flush=True)
    c.close()
