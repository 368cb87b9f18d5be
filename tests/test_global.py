"""Banded global alignment with backtrace (ksw_global2, the CIGAR generator).
CPU: the oracle restatement against the golden vectors of the compiled reference, against the reference itself on
fresh fuzz, and against properties a CIGAR must have.  GPU (-m gpu): the CUDA path through the C ABI against the
oracle and the golden vectors, bit-exact in score and in every CIGAR operation."""
import numpy as np
import pytest

import kswtest as K


def _score_of_cigar(b, k, cig):
    """Score of the alignment a CIGAR describes, recomputed from the scoring scheme."""
    j = b.jobs[k]
    q = b.qpool[int(j["q_off"]):int(j["q_off"]) + int(j["qlen"])]
    t = b.tpool[int(j["t_off"]):int(j["t_off"]) + int(j["tlen"])]
    mat = K.cfg_mat(b.cfg).reshape(5, 5).astype(np.int64)
    x = y = 0
    s = 0
    for op_len in cig:
        op, ln = op_len & 0xf, op_len >> 4
        if op == 0:
            s += int(mat[t[y:y + ln], q[x:x + ln]].sum()); x += ln; y += ln
        elif op == 1:
            s -= b.cfg.o_ins + b.cfg.e_ins * ln; x += ln
        else:
            s -= b.cfg.o_del + b.cfg.e_del * ln; y += ln
    return s, x, y


def test_global_oracle_matches_golden_vectors(oracle_built):
    n = 0
    for name, (b, want) in K.load_global_golden().items():
        assert K.global_mismatch(K.run_global_oracle(b), want) is None, name
        n += b.n
    assert n > 3000


def test_global_oracle_matches_compiled_reference_on_fresh_fuzz(oracle_built):
    if not K.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for seed, cfg in ((501, None), (502, K.make_cfg(a=3, b=2, o_del=5, e_del=3, o_ins=1, e_ins=1)),
                      (503, K.make_cfg(a=1, b=9, o_del=2, e_del=1, o_ins=2, e_ins=1))):
        b = K.gen_global(2500, seed=seed, cfg=cfg)
        assert K.global_mismatch(K.run_global_oracle(b), K.run_global_ref(b)) is None


def test_global_cigar_properties(oracle_built):
    b = K.gen_global(1500, seed=504)
    res, pool = K.run_global_oracle(b)
    for k, cig in enumerate(K.cigars(res, pool)):
        s, x, y = _score_of_cigar(b, k, cig)
        assert (x, y) == (int(b.jobs["qlen"][k]), int(b.jobs["tlen"][k]))      # consumes both sequences completely
        assert s == int(res["score"][k])                                       # and scores what the DP says
        ops = [c & 0xf for c in cig]
        assert all(a != b_ for a, b_ in zip(ops, ops[1:]))                     # runs are maximal (push_cigar merges)


def test_global_identical_sequences_give_one_match_run(oracle_built):
    rng = np.random.default_rng(3)
    q = rng.integers(0, 4, 77).astype(np.uint8)
    jobs = np.zeros(1, dtype=K.GJOB_DT)
    jobs["qlen"], jobs["tlen"], jobs["w"] = 77, 77, 3
    res, pool = K.run_global_oracle(K.GBatch(K.make_cfg(), jobs, q, q.copy()))
    assert int(res["score"][0]) == 77 and K.cigars(res, pool)[0] == (77 << 4,)


def test_global_fast_kernel_source_emulated_matches_oracle(oracle_built):
    """ksw_gfast_core.h compiled for the CPU (software DPX): scores and CIGARs of every eligible job equal the oracle's."""
    total = 0
    cases = [(621, None, dict(max_q=180, w_extra=(0, 1, 2, 3))), (622, None, dict(max_q=400, w_extra=(3, 10, 50, 200), indel=(0.03, 0.08))),
             (623, None, dict(max_q=16, w_extra=(0, 1, 5), n_frac=0.2)), (624, None, dict(max_q=1000, w_extra=(0, 7))),
             (625, K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3), dict(max_q=250)),
             (626, K.make_cfg(a=1, b=1, o_del=1, e_del=2, o_ins=2, e_ins=1), dict(max_q=40)),
             (627, K.make_cfg(a=3, b=2, o_del=0, e_del=0, o_ins=0, e_ins=0), dict(max_q=60))]
    for name, (b, want) in K.load_global_golden().items():
        res, pool, nf = K.run_global_emu(b)
        sel = np.flatnonzero(res["score"] != np.iinfo(np.int32).min)
        assert (res["score"][sel] == want[0]["score"][sel]).all(), name
        ca, cb = K.cigars(res, pool), K.cigars(*want)
        assert all(ca[k] == cb[k] for k in sel), name
        total += nf
    for seed, cfg, kw in cases:
        b = K.gen_global(1500 if kw["max_q"] < 500 else 200, seed=seed, cfg=cfg, **kw)
        want = K.run_global_oracle(b)
        res, pool, nf = K.run_global_emu(b)
        sel = np.flatnonzero(res["score"] != np.iinfo(np.int32).min)
        assert sel.size == nf
        bad = sel[res["score"][sel] != want[0]["score"][sel]]
        assert bad.size == 0, (seed, int(bad[0]), b.jobs[int(bad[0])], int(res["score"][bad[0]]), int(want[0]["score"][bad[0]]))
        ca, cb = K.cigars(res, pool), K.cigars(*want)
        for k in sel:
            assert ca[k] == cb[k], (seed, int(k), b.jobs[int(k)], ca[k], cb[k])
        total += nf
    assert total > 5000


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_global_golden_and_fuzz(gpu_ctx, oracle_built):
    for name, (b, want) in K.load_global_golden().items():
        got = gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool)
        assert K.global_mismatch(got, want) is None, name
    for seed, cfg, mq in ((601, None, 250), (602, K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3), 250),
                          (603, None, 900), (604, K.make_cfg(a=1, b=1, o_del=1, e_del=2, o_ins=2, e_ins=1), 40)):
        b = K.gen_global(6000 if mq < 500 else 800, seed=seed, cfg=cfg, max_q=mq)
        got = gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool)
        mm = K.global_mismatch(got, K.run_global_oracle(b))
        assert mm is None, mm


@pytest.mark.gpu
def test_gpu_global_many_jobs_span_several_chunks(gpu_ctx, oracle_built):
    b = K.gen_global(300000, seed=605, max_q=60, w_extra=(0, 2))               # > 2^18 jobs: two chunks, rebased CIGAR offsets
    got = gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool)
    want = K.run_global_oracle(b)
    assert (got[0]["score"] == want[0]["score"]).all() and (got[0]["n_cigar"] == want[0]["n_cigar"]).all()
    for k in np.random.default_rng(1).integers(0, b.n, 3000):
        a = got[1][int(got[0]["cigar_off"][k]):int(got[0]["cigar_off"][k]) + int(got[0]["n_cigar"][k])]
        w = want[1][int(want[0]["cigar_off"][k]):int(want[0]["cigar_off"][k]) + int(want[0]["n_cigar"][k])]
        assert (a == w).all()
    assert int(got[0]["n_cigar"].sum()) == got[1].shape[0]                    # the pool is dense


@pytest.mark.gpu
def test_gpu_global_empty_batch_and_scalar_dropin(gpu_ctx, oracle_built):
    import bwa_mem_quickassist_b200 as B
    res, pool = gpu_ctx.global_batch(K.make_cfg(), np.zeros(0, dtype=K.GJOB_DT), np.zeros(1, np.uint8), np.zeros(1, np.uint8))
    assert res.shape[0] == 0 and pool.shape[0] == 0
    b = K.gen_global(40, seed=606, max_q=120)
    want_res, want_pool = K.run_global_oracle(b)
    for k, cig in enumerate(K.cigars(want_res, want_pool)):
        j = b.jobs[k]
        q = b.qpool[int(j["q_off"]):int(j["q_off"]) + int(j["qlen"])]
        t = b.tpool[int(j["t_off"]):int(j["t_off"]) + int(j["tlen"])]
        sc, got = B.ksw_global2(int(j["qlen"]), q, int(j["tlen"]), t, 5, K.cfg_mat(b.cfg), b.cfg.o_del, b.cfg.e_del,
                                b.cfg.o_ins, b.cfg.e_ins, int(j["w"]))
        assert sc == int(want_res["score"][k]) and tuple(int(x) for x in got) == cig


@pytest.mark.gpu
def test_gpu_global_both_kernels_agree_with_oracle(gpu_ctx, oracle_built, monkeypatch):
    """The s16x2 kernel (no direction matrix: the backtrack recomputes the reference's bits from H) and the int32 kernel on the
    same jobs: narrow and wide bands, N bases, heavy indels, lengths on every quad boundary."""
    for seed, kw in ((611, dict(max_q=180, w_extra=(0, 1, 2, 3))), (612, dict(max_q=400, w_extra=(3, 10, 50, 200), indel=(0.03, 0.08))),
                     (613, dict(max_q=16, w_extra=(0, 1, 5), n_frac=0.2)), (614, dict(max_q=1000, w_extra=(0, 7)))):
        b = K.gen_global(4000 if kw["max_q"] < 500 else 600, seed=seed, **kw)
        want = K.run_global_oracle(b)
        monkeypatch.delenv("KSW_B200_GLOBAL_FAST", raising=False)
        mm = K.global_mismatch(gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool), want)
        assert mm is None, ("s16x2", seed, mm)
        monkeypatch.setenv("KSW_B200_GLOBAL_FAST", "0")
        mm = K.global_mismatch(gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool), want)
        assert mm is None, ("int32", seed, mm)


def _degenerate_batch():
    """Empty query or empty target, single bases, w = 0 on equal lengths: the band still holds the end cell."""
    rng = np.random.default_rng(9)
    shapes = [(0, 0, 0), (0, 5, 5), (5, 0, 5), (1, 1, 0), (1, 1, 3), (7, 7, 0), (1, 9, 8), (9, 1, 8), (3, 3, 50), (0, 1, 1), (1, 0, 1)]
    qs, ts = [], []
    jobs = np.zeros(len(shapes), dtype=K.GJOB_DT)
    qo = to = 0
    for k, (ql, tl, w) in enumerate(shapes):
        qs.append(rng.integers(0, 4, ql).astype(np.uint8)); ts.append(rng.integers(0, 4, tl).astype(np.uint8))
        jobs[k] = (qo, to, ql, tl, w, 0)
        qo += ql; to += tl
    return K.GBatch(K.make_cfg(), jobs, np.concatenate(qs + [np.zeros(1, np.uint8)]), np.concatenate(ts + [np.zeros(1, np.uint8)]))


def test_global_degenerate_shapes_oracle_vs_reference(oracle_built):
    if not K.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    b = _degenerate_batch()
    assert K.global_mismatch(K.run_global_oracle(b, threads=1), K.run_global_ref(b, threads=1)) is None


@pytest.mark.gpu
def test_gpu_global_degenerate_shapes(gpu_ctx, oracle_built):
    b = _degenerate_batch()
    mm = K.global_mismatch(gpu_ctx.global_batch(b.cfg, b.jobs, b.qpool, b.tpool), K.run_global_oracle(b, threads=1))
    assert mm is None, mm
