"""Local alignment with start positions and second-best score: the reference's ksw_align2 (bwa-0.7.8/ksw.c:329-354, the
striped SSE2 kernels ksw_u8 / ksw_i16), which mem_matesw calls for mate rescue (bwamem_pair.c:150).  SURVEY §8(f) rank 4.
CPU tests pin the restatement (oracle/ksw_align_oracle.c) to the compiled reference; GPU tests compare the kernel with it."""
import numpy as np
import pytest
import kswtest as K

X = K.KSW_XBYTE, K.KSW_XSTOP, K.KSW_XSUBO, K.KSW_XSTART
ALL_FLAGS = [0, K.KSW_XBYTE, K.KSW_XSTART, K.KSW_XSUBO, K.KSW_XSUBO | K.KSW_XSTART, K.KSW_XBYTE | K.KSW_XSUBO | K.KSW_XSTART,
             K.KSW_XSTOP, K.KSW_XSTOP | K.KSW_XSTART, K.KSW_XBYTE | K.KSW_XSTOP | K.KSW_XSUBO | K.KSW_XSTART]
CFGS = [None, K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3), K.make_cfg(a=1, b=1, o_del=1, e_del=2, o_ins=0, e_ins=1),
        K.make_cfg(a=3, b=2, o_del=0, e_del=1, o_ins=0, e_ins=1), K.make_cfg(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1)]


def test_align_oracle_matches_compiled_reference(oracle_built):
    if not K.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    n = 0
    for seed, cfg in enumerate(CFGS):
        b = K.gen_align(700, seed=700 + seed, cfg=cfg, max_q=150 if seed % 2 else 250)               # mem_matesw's flags
        mm = K.align_mismatch(K.run_align_oracle(b), K.run_align_ref(b))
        assert mm is None, ("matesw", seed, mm, b.jobs[mm[0]])
        b = K.gen_align(700, seed=720 + seed, cfg=cfg, max_q=120, max_t=400, flags=ALL_FLAGS)         # every flag combination
        mm = K.align_mismatch(K.run_align_oracle(b), K.run_align_ref(b))
        assert mm is None, ("flags", seed, mm, b.jobs[mm[0]])
        n += 2 * b.n
    assert n >= 7000


def test_align_oracle_matches_golden_vectors(oracle_built):
    """The committed fixture (tests/golden/make_golden.py: outputs of the reference's own ksw_align2) — runs anywhere."""
    n = 0
    for name, (b, want) in K.load_align_golden().items():
        mm = K.align_mismatch(K.run_align_oracle(b), want)
        assert mm is None, (name, mm)
        n += b.n
    assert n >= 2000


def test_align_lazy_f_loop_closed_form_equals_the_loop(oracle_built):
    """The claim the GPU kernel rests on, checked on the CPU: for o_ins >= 1 the reference's lazy-F loop, with
    its early stop and its 16-round limit, gives the same results as H' = max(H, carry propagated all the way).  With o_ins = 0
    it does not (the early stop becomes observable), which is why the kernel keeps the literal loop for that case."""
    n = 0
    for seed, cfg in enumerate([None, K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3), K.make_cfg(a=1, b=1, o_del=3, e_del=2, o_ins=1, e_ins=1),
                                K.make_cfg(a=5, b=3, o_del=2, e_del=1, o_ins=1, e_ins=4), K.make_cfg(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1)]):
        for b in (K.gen_align(1200, seed=800 + seed, cfg=cfg, max_q=250), K.gen_align(800, seed=820 + seed, cfg=cfg, max_q=60, max_t=300, flags=ALL_FLAGS)):
            mm = K.align_mismatch(K.run_align_oracle_closed_form(b), K.run_align_oracle(b))
            assert mm is None, (seed, mm, b.jobs[mm[0]])
            n += b.n
    assert n >= 10000
    # and the counter-example class: zero gap-open for insertions
    b = K.gen_align(3000, seed=840, cfg=K.make_cfg(a=3, b=2, o_del=0, e_del=1, o_ins=0, e_ins=1), max_q=120, max_t=400)
    assert K.align_mismatch(K.run_align_oracle_closed_form(b), K.run_align_oracle(b)) is not None


def test_align_oracle_known_answers(oracle_built):
    """A read embedded once, exactly: score = qlen * a, ends and starts where it was put; embedded twice: the second-best
    score is the same and points at the other copy."""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, 80).astype(np.uint8)
    t = rng.integers(0, 4, 600).astype(np.uint8)
    t[100:180] = q
    t[400:480] = q
    jobs = np.zeros(1, dtype=K.AJOB_DT)
    jobs[0] = (0, 0, 80, 600, K.KSW_XSUBO | K.KSW_XSTART | K.KSW_XBYTE | 19, 0)
    r = K.run_align_oracle(K.ABatch(K.make_cfg(), jobs, np.concatenate([q, np.zeros(16, np.uint8)]), np.concatenate([t, np.zeros(16, np.uint8)])))[0]
    assert (int(r["score"]), int(r["te"]), int(r["qe"]), int(r["tb"]), int(r["qb"])) == (80, 179, 79, 100, 0)
    assert int(r["score2"]) == 80 and int(r["te2"]) == 479


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_align_golden_vectors(gpu_ctx):
    for name, (b, want) in K.load_align_golden().items():
        mm = K.align_mismatch(gpu_ctx.align_batch(b.cfg, b.jobs, b.qpool, b.tpool), want)
        assert mm is None, (name, mm)


@pytest.mark.gpu
def test_gpu_align_matesw_shaped_jobs(gpu_ctx, oracle_built):
    for seed, cfg in enumerate(CFGS):
        b = K.gen_align(3000, seed=740 + seed, cfg=cfg, max_q=150 if seed % 2 else 250)
        mm = K.align_mismatch(gpu_ctx.align_batch(b.cfg, b.jobs, b.qpool, b.tpool), K.run_align_oracle(b))
        assert mm is None, (seed, mm, b.jobs[mm[0]])


@pytest.mark.gpu
def test_gpu_align_every_flag_combination(gpu_ctx, oracle_built):
    for seed, cfg in enumerate(CFGS):
        b = K.gen_align(3000, seed=760 + seed, cfg=cfg, max_q=120, max_t=400, flags=ALL_FLAGS)
        mm = K.align_mismatch(gpu_ctx.align_batch(b.cfg, b.jobs, b.qpool, b.tpool), K.run_align_oracle(b))
        assert mm is None, (seed, mm, b.jobs[mm[0]])


@pytest.mark.gpu
def test_gpu_align_three_kernel_forms_agree(gpu_ctx, oracle_built, monkeypatch):
    """The packed kernel (two SSE lanes per thread in s16x2 registers, closed-form lazy-F loop: the default when o_ins >= 1), the
    int32 kernel with the closed form (KSW_B200_ALIGN_INT32=1) and the int32 kernel running the lazy-F loop literally
    (KSW_B200_ALIGN_LITERAL=1) on the same jobs; scorings with o_ins >= 1 only."""
    for seed, cfg in ((780, None), (781, K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3)), (782, K.make_cfg(a=1, b=1, o_del=3, e_del=2, o_ins=1, e_ins=1))):
        b = K.gen_align(2000, seed=seed, cfg=cfg, max_q=250)
        want = K.run_align_oracle(b)
        for name, env in (("packed", {}), ("int32", {"KSW_B200_ALIGN_INT32": "1"}), ("literal", {"KSW_B200_ALIGN_LITERAL": "1"})):
            monkeypatch.delenv("KSW_B200_ALIGN_INT32", raising=False)
            monkeypatch.delenv("KSW_B200_ALIGN_LITERAL", raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            mm = K.align_mismatch(gpu_ctx.align_batch(b.cfg, b.jobs, b.qpool, b.tpool), want)
            assert mm is None, (name, seed, mm, b.jobs[mm[0]])
    monkeypatch.delenv("KSW_B200_ALIGN_INT32", raising=False)
    monkeypatch.delenv("KSW_B200_ALIGN_LITERAL", raising=False)


@pytest.mark.gpu
def test_gpu_align_edges(gpu_ctx, oracle_built):
    """Empty batch, empty targets, one-base queries, queries on every vector-length boundary, a long query."""
    res = gpu_ctx.align_batch(K.make_cfg(), np.zeros(0, dtype=K.AJOB_DT), np.zeros(16, np.uint8), np.zeros(16, np.uint8))
    assert res.shape[0] == 0
    rng = np.random.default_rng(11)
    shapes = [(1, 0), (1, 1), (1, 40), (7, 0), (8, 30), (9, 30), (15, 64), (16, 64), (17, 64), (31, 90), (32, 90), (33, 90), (249, 700), (250, 700),
              (1000, 1500), (4096, 300)]
    qs, ts = [], []
    jobs = np.zeros(2 * len(shapes), dtype=K.AJOB_DT)
    qo = to = 0
    for k, (ql, tl) in enumerate(shapes + shapes):
        q = rng.integers(0, 4, ql).astype(np.uint8); t = rng.integers(0, 4, tl).astype(np.uint8)
        if tl > ql + 2: t[1:1 + ql] = q
        xtra = K.KSW_XSUBO | K.KSW_XSTART | 5 | (K.KSW_XBYTE if (k < len(shapes) and ql < 250) else 0)
        jobs[k] = (qo, to, ql, tl, xtra, 0)
        qs.append(q); ts.append(t); qo += ql; to += tl
    b = K.ABatch(K.make_cfg(), jobs, np.concatenate(qs + [np.zeros(16, np.uint8)]), np.concatenate(ts + [np.zeros(16, np.uint8)]))
    mm = K.align_mismatch(gpu_ctx.align_batch(b.cfg, b.jobs, b.qpool, b.tpool), K.run_align_oracle(b))
    assert mm is None, (mm, b.jobs[mm[0]])
    from bwa_mem_quickassist_b200 import KswB200Error
    bad = jobs[:1].copy(); bad["qlen"] = 0
    with pytest.raises(KswB200Error):
        gpu_ctx.align_batch(b.cfg, bad, b.qpool, b.tpool)


@pytest.mark.gpu
def test_gpu_align_through_the_shared_queue(oracle_built):
    """Eight threads submit slices of one job set to a queue at the same time: merged batches, everyone's own results."""
    import threading
    import bwa_mem_quickassist_b200 as B
    b = K.gen_align(4000, seed=790, max_q=250)
    want = K.run_align_oracle(b)
    q = B.KswQueue(0)
    got = np.zeros(b.n, dtype=K.ARES_DT)
    errs = []

    def work(t):
        try:
            for lo in range(t * 500, (t + 1) * 500, 125):
                sl = b.jobs[lo:lo + 125]
                got[lo:lo + 125] = q.align_batch(b.cfg, sl, b.qpool, b.tpool)
        except Exception as e:                                     # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    assert K.align_mismatch(got, want) is None
    st = q.stats()
    assert st["submissions"] == 32 and st["batches"] <= 32
    q.close()


@pytest.mark.gpu
def test_gpu_align_scalar_dropin(oracle_built):
    """ksw_align2 with the reference's signature (kswr_t by value), one job per call."""
    import bwa_mem_quickassist_b200 as B
    b = K.gen_align(30, seed=795, max_q=200, max_t=500)
    want = K.run_align_oracle(b)
    for k in range(b.n):
        j = b.jobs[k]
        q = b.qpool[int(j["q_off"]):int(j["q_off"]) + int(j["qlen"])]
        t = b.tpool[int(j["t_off"]):int(j["t_off"]) + int(j["tlen"])]
        r = B.ksw_align2(int(j["qlen"]), q, int(j["tlen"]), t, 5, K.cfg_mat(b.cfg), b.cfg.o_del, b.cfg.e_del, b.cfg.o_ins, b.cfg.e_ins, int(j["xtra"]))
        assert tuple(r[f] for f in ("score", "te", "qe", "score2", "te2", "tb", "qb")) == tuple(int(want[f][k]) for f in ("score", "te", "qe", "score2", "te2", "tb", "qb")), k
