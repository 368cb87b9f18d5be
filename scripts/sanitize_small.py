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
