"""GPU tests of the asynchronous batched entry for page-locked callers (ksw_b200_extend_batch_async / ksw_b200_wait,
SURVEY.md 8(b)): packing, binning and routing run on the device there, so every case is checked against the oracle
AND against the host-packed entry — bit-exact, tolerance zero."""
import numpy as np
import pytest

import bwa_mem_quickassist_b200 as B
import kswtest as K

pytestmark = pytest.mark.gpu


def _run_async(ctx, b: K.Batch):
    pj, pq, pt = B.pinned_copy(b.jobs), B.pinned_copy(b.qpool), B.pinned_copy(b.tpool)
    pr = B.PinnedArray(b.n, B.RES_DT)
    pr.a[:] = 0
    ctx.extend_batch_async(b.cfg, pj.a, pq.a, pt.a, pr.a)
    ctx.wait()
    out = pr.a.copy()
    for p in (pj, pq, pt, pr):
        p.close()
    return out


def _check(ctx, b: K.Batch):
    want = K.run_oracle(b)
    got = _run_async(ctx, b)
    mm = K.first_mismatch(want, got.view(K.RES_DT))
    assert mm is None, f"async entry: first mismatch at job {mm[0]} ({mm[1]} jobs differ): {mm[2]}; job={b.jobs[mm[0]]}"
    host = ctx.extend_batch(b.cfg, b.jobs, b.qpool, b.tpool)
    assert all((host[f] == got[f]).all() for f in B.RES_DT.names)


def test_async_adversarial_and_boundaries(gpu_ctx, oracle_built):
    _check(gpu_ctx, K.gen_adversarial())
    _check(gpu_ctx, K.gen_boundaries())
    _check(gpu_ctx, K.gen_boundaries(cfg=K.make_cfg(a=3, b=5, o_del=9, e_del=2, o_ins=4, e_ins=3, zdrop=200, end_bonus=7)))


def test_async_fuzz_with_n_all_classes(gpu_ctx, oracle_built):
    # N in queries and targets (class-0 jobs that hold one move to class 1 on the device), long queries (generic kernel)
    _check(gpu_ctx, K.gen_fuzz(30000, seed=21, n_frac=0.05))
    _check(gpu_ctx, K.gen_fuzz(4000, seed=22, max_q=900))
    _check(gpu_ctx, K.gen_fuzz(8000, seed=23, cfg=K.make_cfg(a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, end_bonus=9)))
    _check(gpu_ctx, K.gen_fuzz(2000, seed=24, h0_max=40000))


def test_async_config2_several_chunks(gpu_ctx, oracle_built):
    b = K.gen_config2(50000, seed=25)
    gpu_ctx.set_chunk_jobs(7001)                     # 8 ragged chunks through the three pipeline slots
    try:
        _check(gpu_ctx, b)
    finally:
        gpu_ctx.set_chunk_jobs(1 << 20)


def test_async_shared_and_unordered_slices(gpu_ctx, oracle_built):
    # jobs that share sequence slices and come in an order unrelated to the pools' (as mem_chain2aln's do)
    b = K.gen_fuzz(6000, seed=26)
    rng = np.random.default_rng(5)
    jobs = b.jobs.copy()
    jobs["t_off"][1::2] = jobs["t_off"][0::2][: len(jobs["t_off"][1::2])]
    jobs["tlen"][1::2] = np.minimum(jobs["tlen"][1::2], jobs["tlen"][0::2][: len(jobs["tlen"][1::2])])
    jobs = jobs[rng.permutation(len(jobs))]
    gpu_ctx.set_chunk_jobs(1000)
    try:
        _check(gpu_ctx, K.Batch(b.cfg, jobs, b.qpool, b.tpool))
    finally:
        gpu_ctx.set_chunk_jobs(1 << 20)


def test_async_empty_and_errors(gpu_ctx, oracle_built):
    b = K.gen_fuzz(10, seed=27)
    pj, pq, pt = B.pinned_copy(b.jobs), B.pinned_copy(b.qpool), B.pinned_copy(b.tpool)
    pr = B.PinnedArray(b.n, B.RES_DT)
    gpu_ctx.extend_batch_async(b.cfg, pj.a[:0], pq.a, pt.a, pr.a[:0])        # empty batch
    gpu_ctx.wait()
    with pytest.raises(B.KswB200Error, match="page-locked"):                  # pageable result array
        gpu_ctx.extend_batch_async(b.cfg, pj.a, pq.a, pt.a, np.zeros(b.n, dtype=B.RES_DT))
    bad = pj.a.copy(); bad["qlen"][3] = 0
    pb = B.pinned_copy(bad)
    gpu_ctx.extend_batch_async(b.cfg, pb.a, pq.a, pt.a, pr.a)
    with pytest.raises(B.KswB200Error, match="qlen < 1"):
        gpu_ctx.wait()
    far = pj.a.copy(); far["q_off"][2] = b.qpool.nbytes
    pf = B.pinned_copy(far)
    gpu_ctx.extend_batch_async(b.cfg, pf.a, pq.a, pt.a, pr.a)
    with pytest.raises(B.KswB200Error, match="past the end"):
        gpu_ctx.wait()
    gpu_ctx.extend_batch_async(b.cfg, pj.a, pq.a, pt.a, pr.a)                # the context is still usable
    gpu_ctx.wait()
    assert K.first_mismatch(K.run_oracle(b), pr.a.copy().view(K.RES_DT)) is None
