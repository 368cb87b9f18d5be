"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle, bit-exact on all six
outputs (score, qle, tle, gtle, gscore, max_off) for every job.  Integer work: tolerance is zero."""
import os
import subprocess
import sys

import numpy as np
import pytest

import kswtest as K

pytestmark = pytest.mark.gpu


def _check(ctx, b: K.Batch, expect_fast=None):
    want = K.run_oracle(b)
    got = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
    mm = K.first_mismatch(want, got.view(K.RES_DT))
    assert mm is None, f"first mismatch at job {mm[0]} ({mm[1]} jobs differ): {mm[2]}; job={b.jobs[mm[0]]}"
    if expect_fast is not None:
        rb = ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
        info = rb.info()
        rb.free()
        assert (info["n_fast"] > 0) == expect_fast, info


def test_adversarial(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_adversarial(), expect_fast=True)


def test_fuzz_default_scoring(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_fuzz(30000, seed=11))


def test_fuzz_nondefault_asymmetric_gaps(gpu_ctx, oracle_built):
    cfg = K.make_cfg(a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, end_bonus=9)
    _check(gpu_ctx, K.gen_fuzz(20000, seed=12, cfg=cfg))


def test_fuzz_zdrop_disabled_bwasw_style(gpu_ctx, oracle_built):
    # bwasw calls ksw_extend with zdrop=-1, end_bonus=0 (bwtsw2_aux.c:133,161)
    _check(gpu_ctx, K.gen_fuzz(10000, seed=13, cfg=K.make_cfg(zdrop=-1, end_bonus=0)))


def test_fuzz_long_queries_mix_fast_and_generic(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_fuzz(3000, seed=14, max_q=900))


def test_int16_overflow_goes_generic(gpu_ctx, oracle_built):
    # h0 + qlen*a beyond the s16 budget must be routed to the int32 kernel and stay exact
    b = K.gen_fuzz(2000, seed=15, h0_max=40000)
    _check(gpu_ctx, b)
    cfg = K.make_cfg(a=100, b=120, o_del=200, e_del=30, o_ins=250, e_ins=20, zdrop=3000, end_bonus=50)
    _check(gpu_ctx, K.gen_fuzz(2000, seed=16, cfg=cfg, h0_max=3000))


def test_class_boundaries(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_boundaries())
    _check(gpu_ctx, K.gen_boundaries(cfg=K.make_cfg(a=3, b=5, o_del=9, e_del=2, o_ins=4, e_ins=3, zdrop=200, end_bonus=7)))


def test_general_matrix(gpu_ctx, oracle_built):
    rng = np.random.default_rng(5)
    mat = rng.integers(-9, 8, 25).astype(np.int8)
    mat[[0, 6, 12, 18]] = [5, 6, 7, 4]
    _check(gpu_ctx, K.gen_fuzz(10000, seed=17, cfg=K.make_cfg(mat=mat, o_del=5, e_del=2, o_ins=3, e_ins=2, zdrop=50)))


def test_config2_sample(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_config2(200000, seed=20), expect_fast=True)


def test_visited_cell_counts_match_oracle(gpu_ctx, oracle_built):
    for b in (K.gen_config2(50000, seed=41), K.gen_fuzz(5000, seed=42, max_q=700)):
        _, cells = K.run_oracle(b, want_cells=True)
        rb = gpu_ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
        gpu_ctx.run(rb)
        got = gpu_ctx.download_cells(rb)
        rb.free()
        assert (got.astype(np.int64) == cells).all()


def test_golden_vectors(gpu_ctx):
    for name, (b, want) in K.load_golden().items():
        got = gpu_ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
        assert K.first_mismatch(want, got.view(K.RES_DT)) is None, name


def test_empty_and_single(gpu_ctx, oracle_built):
    b = K.gen_fuzz(1, seed=3)
    _check(gpu_ctx, b)
    empty = K.Batch(b.cfg, b.jobs[:0].copy(), b.qpool, b.tpool)
    assert gpu_ctx.extend_batch(empty.cfg, empty.jobs, empty.qpool, empty.tpool).shape[0] == 0


def test_usage_errors_are_reported_and_context_stays_usable(gpu_ctx, oracle_built):
    import bwa_mem_quickassist_b200 as B
    b = K.gen_fuzz(50, seed=51)
    bad = b.jobs.copy(); bad["qlen"][7] = 0                      # the reference would write out of bounds (ksw.c:394)
    with pytest.raises(B.KswB200Error, match="qlen < 1"):
        gpu_ctx.extend_batch(b.cfg, bad, b.qpool, b.tpool)
    cfg6 = K.make_cfg(); cfg6.m = 6                               # no reference caller passes anything but 5
    with pytest.raises(B.KswB200Error, match="m == 5"):
        gpu_ctx.extend_batch(cfg6, b.jobs, b.qpool, b.tpool)
    _check(gpu_ctx, b)                                            # the context still works


def test_multi_context_sharding(oracle_built):
    # one context per visible GPU (two contexts on the same GPU when only one is visible): same results as one context
    import bwa_mem_quickassist_b200 as B
    n_gpu = max(1, B.load_library().ksw_b200_device_count())
    ctxs = [B.KswB200(d % n_gpu) for d in range(max(2, min(n_gpu, 8)))]
    b = K.gen_fuzz(20000, seed=61, max_q=400)
    got = B.extend_batch_multi(ctxs, b.cfg, b.jobs, b.qpool, b.tpool)
    assert K.first_mismatch(K.run_oracle(b), got.view(K.RES_DT)) is None
    for c in ctxs:
        c.close()


def test_concurrent_contexts_with_different_query_lengths(oracle_built):
    # regression: host threads (one context each) launching the fast kernel with different shared-memory sizes at the same
    # time must not disturb each other (the dynamic shared-memory ceiling is a per-device kernel attribute)
    import threading
    import bwa_mem_quickassist_b200 as B
    batches = [K.gen_fuzz(1500, seed=70 + i, max_q=mq, h0_max=80) for i, mq in enumerate((20, 60, 120, 250, 500, 30, 400, 100))]
    wants = [K.run_oracle(b) for b in batches]
    errs = []

    def work(i):
        try:
            ctx = B.KswB200(0, pack_threads=1)
            for _ in range(25):
                got = ctx.extend_batch(batches[i].cfg, batches[i].jobs, batches[i].qpool, batches[i].tpool)
                mm = K.first_mismatch(wants[i], got.view(K.RES_DT))
                if mm is not None:
                    errs.append((i, mm)); break
            ctx.close()
        except Exception as e:          # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:3]


def test_scalar_dropins(gpu_ctx, oracle_built):
    import bwa_mem_quickassist_b200 as B
    b = K.gen_fuzz(40, seed=21)
    want = K.run_oracle(b)
    mat = K.cfg_mat(b.cfg)
    for k in range(b.n):
        j = b.jobs[k]
        q = b.qpool[int(j["q_off"]): int(j["q_off"]) + int(j["qlen"])]
        t = b.tpool[int(j["t_off"]): int(j["t_off"]) + int(j["tlen"])]
        got = B.ksw_extend2(int(j["qlen"]), q, int(j["tlen"]), t, 5, mat, 6, 1, 6, 1, int(j["w"]), 5, 100, int(j["h0"]))
        assert got == tuple(int(want[f][k]) for f in K.RES_DT.names)
        got = B.ksw_extend(int(j["qlen"]), q, int(j["tlen"]), t, 5, mat, 6, 1, int(j["w"]), 5, 100, int(j["h0"]))
        assert got == tuple(int(want[f][k]) for f in K.RES_DT.names)


def test_generic_kernel_alone(oracle_built):
    """Same jobs with the fast kernel disabled (separate process: the switch is read once)."""
    code = (
        "import sys; sys.path[:0]=['tests','.']\n"
        "import kswtest as K, bwa_mem_quickassist_b200 as B\n"
        "ctx=B.KswB200(0)\n"
        "for b in (K.gen_adversarial(), K.gen_fuzz(8000, seed=31), K.gen_config2(20000, seed=32)):\n"
        "    rb=ctx.upload(b.cfg,b.jobs,b.qpool,b.tpool); assert rb.info()['n_fast']==0; rb.free()\n"
        "    mm=K.first_mismatch(K.run_oracle(b), ctx.extend_batch(b.cfg,b.jobs,b.qpool,b.tpool).view(K.RES_DT))\n"
        "    assert mm is None, mm\n"
        "print('ok')\n")
    env = dict(os.environ, KSW_B200_DISABLE_FAST="1")
    out = subprocess.run([sys.executable, "-c", code], cwd=K.ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_size_independent_properties_full_rows(gpu_ctx, oracle_built):
    """At sizes the oracle is not run on: identical sequences must score h0+qlen with qle=tle=qlen and
    gscore=score (closed form), and splitting/permuting a batch must not change any job's result."""
    n, L = 300000, 101
    rng = np.random.default_rng(77)
    t = rng.integers(0, 4, (n, L), dtype=np.uint8)
    jobs = np.zeros(n, dtype=K.JOB_DT)
    jobs["q_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
    jobs["t_off"] = jobs["q_off"]
    jobs["qlen"] = L; jobs["tlen"] = L; jobs["w"] = 100
    jobs["h0"] = rng.integers(19, 101, n)
    cfg = K.make_cfg()
    r = gpu_ctx.extend_batch(cfg, jobs, t.reshape(-1), t.reshape(-1))
    assert (r["score"] == jobs["h0"] + L).all() and (r["qle"] == L).all() and (r["tle"] == L).all()
    assert (r["gscore"] == r["score"]).all() and (r["gtle"] == L).all() and (r["max_off"] == 0).all()
    perm = rng.permutation(n)
    r2 = gpu_ctx.extend_batch(cfg, jobs[perm], t.reshape(-1), t.reshape(-1))
    for f in K.RES_DT.names:
        assert (r2[f] == r[f][perm]).all()


# ------------------------------------------------------------------ the pair kernel (two jobs per lane), opt-in
@pytest.fixture()
def pair_kernel_on():
    """KSW_B200_PAIR is read per launch: class 0 (qlen <= 124, scores < 512, no N in the query) goes to ksw_pair_kernel."""
    os.environ["KSW_B200_PAIR"] = "1"
    yield
    os.environ.pop("KSW_B200_PAIR", None)


def test_pair_kernel_config2_and_cells(gpu_ctx, oracle_built, pair_kernel_on):
    b = K.gen_config2(150000, seed=61)
    _, cells = K.run_oracle(b, want_cells=True)
    _check(gpu_ctx, b, expect_fast=True)
    rb = gpu_ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
    gpu_ctx.run(rb)
    got = gpu_ctx.download_cells(rb)
    rb.free()
    assert (got.astype(np.int64) == cells).all()


def test_pair_kernel_fuzz_and_adversarial(gpu_ctx, oracle_built, pair_kernel_on):
    _check(gpu_ctx, K.gen_adversarial())
    _check(gpu_ctx, K.gen_boundaries())
    _check(gpu_ctx, K.gen_fuzz(30000, seed=62, max_q=124, h0_max=120))                  # mostly class 0, N in queries and targets
    _check(gpu_ctx, K.gen_fuzz(20000, seed=63, max_q=300))                                # class 0 next to the other classes
    _check(gpu_ctx, K.gen_fuzz(15000, seed=64, max_q=124, w_choices=(1, 3, 20, 300), h0_max=40,
                               cfg=K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3, zdrop=15, end_bonus=2)))
    _check(gpu_ctx, K.gen_fuzz(15000, seed=65, max_q=124, cfg=K.make_cfg(zdrop=-1)))
    _check(gpu_ctx, K.gen_fuzz(3, seed=66, max_q=60, n_frac=0.0))                         # fewer jobs than one lane's two slots x 32
    for name, (b, want) in K.load_golden().items():
        got = gpu_ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
        assert K.first_mismatch(want, got.view(K.RES_DT)) is None, name


def test_warp_cooperative_kernel_long_queries(gpu_ctx, oracle_built):
    """Queries longer than the s16x2 kernel's 512 columns and int16-unsafe jobs of 129+ columns run on the
    warp-cooperative int32 kernel (ksw_warp.cu: one job per warp, F by a max-plus scan over the lanes)."""
    b = K.gen_fuzz(1500, seed=41, max_q=3000, w_choices=(5, 50, 100, 400, 1500))
    rb = gpu_ctx.upload(b.cfg, b.jobs, b.qpool, b.tpool)
    info = rb.info()
    rb.free()
    assert info["n_generic"] > 300 and info["n_fast"] > 100, info        # both routes populated
    _check(gpu_ctx, b)
    # bwasw-style scoring (no z-drop), asymmetric gaps, and very long queries
    _check(gpu_ctx, K.gen_fuzz(400, seed=42, max_q=2500, cfg=K.make_cfg(zdrop=-1, end_bonus=0), w_choices=(30, 200, 2500)))
    _check(gpu_ctx, K.gen_fuzz(400, seed=43, max_q=2000, cfg=K.make_cfg(a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, end_bonus=9)))
    _check(gpu_ctx, K.gen_fuzz(12, seed=44, max_q=20000, w_choices=(100, 3000), related=1.0))
    # carried-in scores beyond int16 with medium queries: warp kernel (129+) and thread kernel (shorter) side by side
    _check(gpu_ctx, K.gen_fuzz(1500, seed=45, max_q=400, h0_max=60000))
    cfg = K.make_cfg(a=100, b=120, o_del=200, e_del=30, o_ins=250, e_ins=20, zdrop=3000, end_bonus=50)
    _check(gpu_ctx, K.gen_fuzz(800, seed=46, cfg=cfg, max_q=700, h0_max=3000))
