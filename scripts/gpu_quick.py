"""Quick GPU sanity + timing (development helper, not the bench)."""
import sys, time, os
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B

ctx = B.KswB200(0)
for which in (0, 1):
    ops, ms = ctx.dpx_peak(which)
    print(f"dpx_peak which={which}: {ops/1e12:.2f} T lane-ops/s ({ms:.3f} ms)")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
b = K.gen_config2(n, seed=1)
t0 = time.time(); want, cells = K.run_oracle(b, threads=os.cpu_count(), want_cells=True); t1 = time.time()
print(f"oracle: {n} jobs {cells.sum()/1e9:.2f} Gcells in {t1-t0:.2f}s on {os.cpu_count()} threads -> {cells.sum()/(t1-t0)/1e9:.2f} GCUPS")
t0 = time.time(); got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool); t1 = time.time()
print("e2e first call %.3fs" % (t1 - t0), "mismatch:", K.first_mismatch(want, got.view(K.RES_DT)))
t0 = time.time(); got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool); t1 = time.time()
print("e2e second call %.3fs -> %.2f M ext/s" % (t1 - t0, n / (t1 - t0) / 1e6))
rb = ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
print(rb.info())
ms = ctx.run_timed(rb, 5)
print("kernel ms:", ms, "-> %.1f GCUPS (visited), %.2f M ext/s" % (cells.sum() / (ms[1:].mean() * 1e-3) / 1e9, n / (ms[1:].mean() * 1e-3) / 1e6))
got = ctx.download(rb)
print("mismatch after timed:", K.first_mismatch(want, got.view(K.RES_DT)))
