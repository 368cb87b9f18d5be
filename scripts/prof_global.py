"""Profile target: a few calls of ksw_b200_global_batch on N jobs of the bwa_gen_cigar2 shape (150 bp, w = 35)."""
import os, sys, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = 150
rng = np.random.default_rng(7)
q = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
t = q.copy()
pos = rng.random((n, L)) < 0.01
t[pos] = (t[pos] + rng.integers(1, 4, size=int(pos.sum()), dtype=np.uint8)) & 3
jobs = np.zeros(n, dtype=K.GJOB_DT)
jobs["q_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
jobs["t_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
jobs["qlen"], jobs["tlen"], jobs["w"] = L, L, 35
ctx = B.KswB200(0)
for _ in range(3):
    t0 = time.perf_counter(); res, cig = ctx.global_batch(K.make_cfg(), jobs, q.reshape(-1), t.reshape(-1)); dt = time.perf_counter() - t0
print("ms", 1e3 * dt, "band cells", n * L * 71)
