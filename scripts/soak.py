"""Randomised soak on the GPU: extension (default kernels, then the opt-in pair kernel) and global alignment against the
oracle with random scoring schemes, lengths and bands, until the time budget is spent.
  python scripts/soak.py [seconds] [seed0]"""
import os, sys, time
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "tests"), os.path.join(os.path.dirname(__file__), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = B.KswB200(0)
t_end = time.time() + budget
n_ext = n_glob = rounds = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed)
    a = int(rng.integers(1, 5)); b = int(rng.integers(1, 10))
    cfg = K.make_cfg(a=a, b=b, o_del=int(rng.integers(0, 13)), e_del=int(rng.integers(1, 5)), o_ins=int(rng.integers(0, 13)),
                     e_ins=int(rng.integers(1, 5)), zdrop=int(rng.choice([-1, 0, 1, 20, 100, 400])), end_bonus=int(rng.integers(0, 12)))
    max_q = int(rng.choice([30, 100, 124, 128, 200, 400, 700]))
    bt = K.gen_fuzz(int(rng.integers(2000, 12000)), seed=seed, cfg=cfg, max_q=max_q, h0_max=int(rng.choice([20, 100, 250, 1000])))
    want = K.run_oracle(bt, threads=os.cpu_count())
    for pair in ("0", "1"):
        os.environ["KSW_B200_PAIR"] = pair
        got = ctx.extend_batch(bt.cfg, bt.jobs, bt.qpool, bt.tpool)
        mm = K.first_mismatch(want, got.view(K.RES_DT))
        assert mm is None, ("extend", seed, pair, mm)
    n_ext += bt.n
    g = K.gen_global(int(rng.integers(1000, 6000)), seed=seed, cfg=cfg, max_q=int(rng.choice([20, 150, 300, 600])))
    mm = K.global_mismatch(ctx.global_batch(g.cfg, g.jobs, g.qpool, g.tpool), K.run_global_oracle(g, threads=os.cpu_count()))
    assert mm is None, ("global", seed, mm)
    n_glob += g.n
    seed += 1; rounds += 1
print(f"soak ok: {rounds} rounds, {n_ext} extension jobs x 2 kernels, {n_glob} global alignments, all bit-exact")
