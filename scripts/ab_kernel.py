#!/usr/bin/env python
"""Development helper (GPU box): A/B of library builds on the resident kernel path of config 2.
   python scripts/ab_kernel.py [jobs] [lib.so ...]      (no libs: the in-tree one + build/variants/*.so)
Each build runs in its own process (KSW_B200_LIB): 5 timed launches, bit-exact check of the first 200 k jobs."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "tests")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs
n = %(n)d
jobs, qpool, tpool = config2_jobs(n, seed=12345)
ctx = B.KswB200(0)
cfg = B.make_cfg()
rb = ctx.upload(cfg, jobs, qpool, tpool)
for _ in range(3): ctx.run(rb)
ctx.sync()
ms = ctx.run_timed(rb, 5)
cells = ctx.download_cells(rb).astype(np.int64)
got = ctx.download(rb)
ns = min(n, 200000)
want, ocells = K.run_oracle(K.Batch(K.make_cfg(), jobs[:ns], qpool, tpool), threads=os.cpu_count(), want_cells=True)
ok = all((want[f] == got[f][:ns]).all() for f in B.RES_DT.names) and bool((ocells == cells[:ns]).all())
print(json.dumps({"lib": os.environ.get("KSW_B200_LIB", "in-tree"), "ms": float(ms.mean()), "ms_min": float(ms.min()),
                  "gcups": float(cells.sum() / ms.mean() / 1e6), "bit_exact": ok}))
'''

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    libs = sys.argv[2:] or [""] + sorted(glob.glob(os.path.join(ROOT, "bwa_mem_quickassist_b200/build/variants/libksw_b200_*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib: env["KSW_B200_LIB"] = os.path.abspath(lib)
        else: env.pop("KSW_B200_LIB", None)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "n": n}], env=env, capture_output=True, text=True, timeout=600)
        print((r.stdout.strip().splitlines() or ["?"])[-1] if r.returncode == 0 else f"FAILED {lib}: {r.stderr[-400:]}", flush=True)

if __name__ == "__main__":
    main()
