"""Times the banded global alignment + backtrace path (ksw_b200_global_batch) next to the reference's own ksw_global2
on the host cores, on the job shape bwa_gen_cigar2 produces for 150 bp reads (bwa.c:118-132: w = 35 for default scoring).
  python scripts/bench_global.py [jobs] [read_len]
Unit: band cells per second = sum over jobs of tlen * min(qlen, 2w+1) (every cell of the band is computed; no trimming)."""
import json, os, sys, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rng = np.random.default_rng(7)
# query = read, target = reference window: 1 % substitutions, one 1-3 bp indel in 10 % of the jobs
q = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
t = q.copy()
pos = rng.random((n, L)) < 0.01
t[pos] = (t[pos] + rng.integers(1, 4, size=int(pos.sum()), dtype=np.uint8)) & 3
tl = np.full(n, L, dtype=np.int64)
tpool = np.empty(n * (L + 3), dtype=np.uint8)
toff = np.arange(n, dtype=np.int64) * (L + 3)
T2 = np.concatenate([t, rng.integers(0, 4, size=(n, 3), dtype=np.uint8)], axis=1)
ind = np.flatnonzero(rng.random(n) < 0.10)
for k in ind[:200000]:
    d = int(rng.integers(1, 4)); p = int(rng.integers(10, L - 10))
    if rng.random() < 0.5:   # deletion from the read's point of view: the window is longer
        T2[k, p + d:L + d] = t[k, p:L]; tl[k] = L + d
    else:
        T2[k, p:L - d] = t[k, p + d:L]; tl[k] = L - d
tpool[:] = T2.reshape(-1)
jobs = np.zeros(n, dtype=K.GJOB_DT)
jobs["q_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
jobs["t_off"] = toff
jobs["qlen"], jobs["tlen"] = L, tl
max_gap = max(int((((L + 1) >> 1) * 1 - 6) / 1 + 1.), 1)
w = np.maximum((max_gap + np.abs(tl - L) + 1) >> 1, np.abs(tl - L) + 3)
jobs["w"] = w
cfg = K.make_cfg()
cells = float((tl * np.minimum(L, 2 * w + 1)).sum())
ctx = B.KswB200(0)
ctx.global_batch(cfg, jobs[:1000], q.reshape(-1), tpool)
times = []
for _ in range(3):
    t0 = time.perf_counter(); res, cig = ctx.global_batch(cfg, jobs, q.reshape(-1), tpool); times.append(time.perf_counter() - t0)
dt = min(times)
ns = min(n, 200000)
b = K.GBatch(cfg, jobs[:ns], q.reshape(-1), tpool)
threads = os.cpu_count() or 1
t0 = time.perf_counter(); want = K.run_global_ref(b, threads=threads) if K.have_ref() else K.run_global_oracle(b, threads=threads); dtc = time.perf_counter() - t0
ok = K.global_mismatch((res[:ns], cig), want) is None
cells_s = float((tl[:ns] * np.minimum(L, 2 * w[:ns] + 1)).sum())
print(json.dumps({"what": "ksw_b200_global_batch, host buffers in, host results out (pack + H2D + kernel + D2H)",
                  "jobs": n, "read_len": L, "band_cells_per_job": cells / n, "ms": 1e3 * dt, "jobs_per_s": n / dt,
                  "gcups_band_cells": cells / dt / 1e9, "cigar_ops_per_job": float(res["n_cigar"].mean()),
                  "cpu": {"kind": "reference" if K.have_ref() else "port", "threads": threads, "jobs": ns,
                          "jobs_per_s": ns / dtc, "gcups_band_cells": cells_s / dtc / 1e9},
                  "bit_exact_vs_cpu_sample": bool(ok)}))
