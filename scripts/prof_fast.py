#!/usr/bin/env python
"""Development helper (GPU box): uploads N config-2 jobs and launches the resident kernels a few times — the process
ncu attaches to (ncu -k regex:ksw_fast --launch-skip 1 --launch-count 1 --set full ...).
   python scripts/prof_fast.py [jobs] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
it = int(sys.argv[2]) if len(sys.argv) > 2 else 3
jobs, qpool, tpool = config2_jobs(n, seed=12345)
ctx = B.KswB200(0)
rb = ctx.upload(B.make_cfg(), jobs, qpool, tpool)
ms = ctx.run_timed(rb, it)
cells = ctx.download_cells(rb).astype(np.int64).sum()
print(f"{n} jobs, {cells} visited cells, ms per launch {ms}, {cells / ms[1:].mean() / 1e6:.1f} GCUPS")
