"""Times the mate-rescue local alignment (ksw_b200_align_batch = the reference's ksw_align2) next to the reference's own SSE2
code on the host cores, on the job shape mem_matesw produces (bwamem_pair.c:128-150): a 150 bp read against a rescue window
of 400-900 bp that holds a mutated copy of it in 70 % of the jobs; xtra = KSW_XSUBO | KSW_XSTART | KSW_XBYTE | 19.
  python scripts/bench_align.py [jobs] [read_len]
Unit: forward-pass cells per second = sum of qlen * tlen (the second pass over the reversed prefixes comes on top)."""
import json, os, sys, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rng = np.random.default_rng(3)
tl = rng.integers(400, 901, n)
toff = np.concatenate([[0], np.cumsum(tl)[:-1]])
tpool = rng.integers(0, 4, int(tl.sum()) + 16).astype(np.uint8)
q = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
has = rng.random(n) < 0.7
pos = (rng.random(n) * (tl - L - 1)).astype(np.int64)
mut = q.copy()
m = rng.random((n, L)) < 0.02
mut[m] = (mut[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
idx = np.flatnonzero(has)
cols = (toff[idx] + pos[idx])[:, None] + np.arange(L)[None, :]
tpool[cols] = mut[idx]
jobs = np.zeros(n, dtype=K.AJOB_DT)
jobs["q_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
jobs["t_off"] = toff
jobs["qlen"], jobs["tlen"] = L, tl
jobs["xtra"] = K.KSW_XSUBO | K.KSW_XSTART | (K.KSW_XBYTE if L < 250 else 0) | 19
cfg = K.make_cfg()
qpool = np.concatenate([q.reshape(-1), np.zeros(16, np.uint8)])
cells = float((tl * L).sum())
ctx = B.KswB200(0)
ctx.align_batch(cfg, jobs[:1000], qpool, tpool)
times = []
for _ in range(3):
    t0 = time.perf_counter(); res = ctx.align_batch(cfg, jobs, qpool, tpool); times.append(time.perf_counter() - t0)
dt = min(times)
ns = min(n, 50000)
b = K.ABatch(cfg, jobs[:ns], qpool, tpool)
threads = os.cpu_count() or 1
t0 = time.perf_counter(); want = K.run_align_ref(b, threads=threads) if K.have_ref() else K.run_align_oracle(b, threads=threads); dtc = time.perf_counter() - t0
ok = K.align_mismatch(res[:ns], want) is None
print(json.dumps({"what": "ksw_b200_align_batch, host buffers in, host results out", "jobs": n, "read_len": L, "mean_tlen": float(tl.mean()),
                  "ms": dt * 1e3, "jobs_per_s": n / dt, "gcups_forward_cells": cells / dt / 1e9, "rescued_fraction": float((res["score"] >= 19).mean()),
                  "cpu": {"kind": "reference" if K.have_ref() else "port", "threads": threads, "jobs": ns, "jobs_per_s": ns / dtc,
                          "gcups_forward_cells": float((tl[:ns] * L).sum()) / dtc / 1e9},
                  "bit_exact_vs_cpu_sample": bool(ok)}))
