"""Profile target: one upload of N config-2 jobs, then a few launches of the fast kernel."""
import sys, os
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import kswtest as K
import bwa_mem_quickassist_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = B.KswB200(0)
b = K.gen_config2(n, seed=1)
rb = ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
ms = ctx.run_timed(rb, reps)
print("ms", ms)
