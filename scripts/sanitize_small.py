"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel variant on a few thousand jobs."""
import sys, os
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import kswtest as K
import bwa_mem_quickassist_b200 as B
ctx = B.KswB200(0)
for name, b in (("adversarial", K.gen_adversarial()), ("boundaries", K.gen_boundaries()), ("fuzz", K.gen_fuzz(1500, seed=3, max_q=700)),
                ("config2", K.gen_config2(3000, seed=4))):
    got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
    mm = K.first_mismatch(K.run_oracle(b), got.view(K.RES_DT))
    print(name, b.n, "mismatch:", mm)
    assert mm is None
cs = K.gen_chains(150, seed=5)
assert K.regs_equal(K.run_chain_gpu(ctx, cs), K.run_chain_oracle(cs)[:2])
print("chains ok")
# round 2: warp-cooperative kernel (long queries), async pinned entry (device packing), global alignment (both kernels),
# mate-rescue alignment (packed / int32 / literal forms)
import numpy as np
b = K.gen_fuzz(300, seed=6, max_q=1500)
got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
assert K.first_mismatch(K.run_oracle(b), got.view(K.RES_DT)) is None
print("long queries ok")
b = K.gen_config2(5000, seed=7)
qp, tp, jp = B.pinned_copy(b.qpool), B.pinned_copy(b.tpool), B.pinned_copy(b.jobs)
res = B.PinnedArray((b.n,), B.RES_DT)
ctx.extend_batch_async(b.cfg, jp.a.astype(B.JOB_DT, copy=False), qp.a, tp.a, res.a); ctx.wait()
assert K.first_mismatch(K.run_oracle(b), res.a.view(K.RES_DT)) is None
print("async entry ok")
g = K.gen_global(1500, seed=8, max_q=200)
want = K.run_global_oracle(g)
assert K.global_mismatch(ctx.global_batch(g.cfg, g.jobs, g.qpool, g.tpool), want) is None
os.environ["KSW_B200_GLOBAL_FAST"] = "0"
assert K.global_mismatch(ctx.global_batch(g.cfg, g.jobs, g.qpool, g.tpool), want) is None
del os.environ["KSW_B200_GLOBAL_FAST"]
print("global ok")
a = K.gen_align(600, seed=9, max_q=250, max_t=500)
want = K.run_align_oracle(a)
for env in ({}, {"KSW_B200_ALIGN_INT32": "1"}, {"KSW_B200_ALIGN_LITERAL": "1"}):
    os.environ.update(env)
    assert K.align_mismatch(ctx.align_batch(a.cfg, a.jobs, a.qpool, a.tpool), want) is None
    for k in env: del os.environ[k]
print("align ok")
