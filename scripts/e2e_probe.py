import sys, os, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import bwa_mem_quickassist_b200 as B
from bwa_mem_quickassist_b200.synth import config2_jobs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
jobs, q, t = config2_jobs(n, seed=1)
ctx = B.KswB200(0)
cfg = B.make_cfg()
out = np.zeros(n, dtype=B.RES_DT)
for rep in range(4):
    t0 = time.perf_counter(); ctx.extend_batch(cfg, jobs, q, t, out=out); dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3:.1f} ms -> {n/dt/1e6:.1f} M ext/s", flush=True)
t0 = time.perf_counter(); out[:] = 0; print("zeroing out: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
