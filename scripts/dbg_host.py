import os, sys
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B
ctx = B.KswB200(0)
def host(b, tag):
    want = K.run_oracle(b)
    got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
    bad = np.zeros(b.n, bool)
    for f in B.RES_DT.names: bad |= want[f] != got[f]
    print(tag, "host path mismatches:", int(bad.sum()), "of", b.n, "first idx", np.flatnonzero(bad)[:8], flush=True)
    if bad.any():
        i = int(np.flatnonzero(bad)[0]); print("  job", b.jobs[i], "want", want[i], "got", got[i])
def asy(b, tag):
    want = K.run_oracle(b)
    pj, pq, pt = B.pinned_copy(b.jobs), B.pinned_copy(b.qpool), B.pinned_copy(b.tpool)
    pr = B.PinnedArray(b.n, B.RES_DT); pr.a[:] = 0
    ctx.extend_batch_async(b.cfg, pj.a, pq.a, pt.a, pr.a); ctx.wait()
    got = pr.a.copy()
    bad = np.zeros(b.n, bool)
    for f in B.RES_DT.names: bad |= want[f] != got[f]
    print(tag, "async mismatches:", int(bad.sum()), "of", b.n, "first idx", np.flatnonzero(bad)[:8], flush=True)
    if bad.any():
        i = int(np.flatnonzero(bad)[0]); print("  job", b.jobs[i], "want", want[i], "got", got[i])
adv = K.gen_adversarial(); c2 = K.gen_config2(60000, seed=3); fz = K.gen_fuzz(20000, seed=4, n_frac=0.05)
host(adv, "adv before"); host(c2, "c2 before"); host(fz, "fuzz before")
asy(adv, "adv"); asy(c2, "c2"); asy(fz, "fuzz")
host(adv, "adv after"); host(c2, "c2 after"); host(fz, "fuzz after")
ctx.set_chunk_jobs(7001)
host(c2, "c2 chunked"); asy(c2, "c2 chunked"); asy(fz, "fuzz chunked"); host(fz, "fuzz chunked")
