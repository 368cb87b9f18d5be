"""GPU tests of the extension against a reference kept on the device (ksw_b200_ref_set / ksw_b200_extend_batch_ref,
SURVEY.md 8(f) rank 3): targets are named by coordinate in the doubled reference space and sliced from the resident
2-bit .pac by the GPU.  Checked bit for bit against the same jobs with host-materialised slices (b200_get_ref_slice =
bns_get_seq, bntseq.c:355-376) through the oracle, and through the chain driver against the golden regions."""
import ctypes as C

import numpy as np
import pytest

import bwa_mem_quickassist_b200 as B
import kswtest as K

pytestmark = pytest.mark.gpu


def _ref_slice(lib, l_pac, pac, beg, end):
    out = np.zeros(max(abs(end - beg), 1), dtype=np.uint8)
    lib.b200_get_ref_slice.restype = C.c_int64
    lib.b200_get_ref_slice.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    n = lib.b200_get_ref_slice(l_pac, pac.ctypes.data_as(C.c_void_p), beg, end, out.ctypes.data_as(C.c_void_p))
    return out[:n]


def test_ref_jobs_match_materialised_slices(gpu_ctx, oracle_built):
    rng = np.random.default_rng(77)
    l_pac = 50_000
    genome = rng.integers(0, 4, l_pac).astype(np.uint8)
    pac = K.pack_pac(genome)
    gpu_ctx.ref_set(pac, l_pac)
    n = 6000
    reads, qoffs, at = [], [], 0
    rj = np.zeros(n, dtype=B.RJOB_DT)
    qs, ts = [], []
    for k in range(n):
        ql = int(rng.integers(1, 260))
        tl = int(rng.integers(0, 2 * ql + 40))
        strand = int(rng.integers(0, 2))
        lo = strand * l_pac
        beg = int(rng.integers(lo, lo + l_pac - tl)) if tl < l_pac else lo
        t_fwd = _ref_slice(gpu_ctx.lib, l_pac, pac, beg, beg + tl)            # what bns_get_seq returns for [beg, beg+tl)
        assert len(t_fwd) == tl
        down = bool(rng.integers(0, 2))
        # the read: a noisy copy of the target (so that real extensions happen), sometimes with N
        src = t_fwd[::-1] if down else t_fwd
        q = K.mutate(rng, src, 0.03, 0.01, max_indel=5)[:ql] if tl else np.zeros(0, np.uint8)
        if len(q) < ql:
            q = np.concatenate([q, rng.integers(0, 4, ql - len(q)).astype(np.uint8)])
        if rng.random() < 0.1:
            q = np.where(rng.random(ql) < 0.03, 4, q).astype(np.uint8)
        q_down = bool(rng.integers(0, 2))
        stored = q[::-1] if q_down else q                                       # how the read sits in the pool
        reads.append(stored)
        rj[k]["q_off"] = at + (ql - 1 if q_down else 0)
        rj[k]["q_step"] = -1 if q_down else 1
        rj[k]["t_pos"] = beg + tl - 1 if down else beg
        rj[k]["t_step"] = -1 if down else 1
        rj[k]["qlen"], rj[k]["tlen"] = ql, tl
        rj[k]["h0"] = int(rng.integers(0, 120))
        rj[k]["w"] = int(rng.choice([5, 30, 100, 200]))
        at += ql
        qs.append(q); ts.append(src)
    qpool = np.concatenate(reads)
    cfg = K.make_cfg()
    got = gpu_ctx.extend_batch_ref(cfg, rj, qpool)
    want = K.run_oracle(K._pools_from_lists(qs, ts, rj["h0"], rj["w"], cfg))
    mm = K.first_mismatch(want, got.view(K.RES_DT))
    assert mm is None, f"first mismatch at job {mm[0]} ({mm[1]} jobs differ): {mm[2]}; job={rj[mm[0]]}"


def test_ref_rejects_bad_jobs(gpu_ctx, oracle_built):
    l_pac = 1000
    pac = K.pack_pac(np.random.default_rng(1).integers(0, 4, l_pac).astype(np.uint8))
    gpu_ctx.ref_set(pac, l_pac)
    rj = np.zeros(2, dtype=B.RJOB_DT)
    rj["qlen"], rj["tlen"], rj["q_step"], rj["t_step"], rj["w"] = 10, 20, 1, 1, 50
    rj[1]["t_pos"] = l_pac - 5                                                  # bridges the strand boundary
    with pytest.raises(B.KswB200Error, match="bridging"):
        gpu_ctx.extend_batch_ref(K.make_cfg(), rj, np.zeros(10, np.uint8))
    rj[1]["t_pos"] = 0
    gpu_ctx.extend_batch_ref(K.make_cfg(), rj, np.zeros(10, np.uint8))          # and the context is still usable


def test_chain_driver_in_device_reference_mode(gpu_ctx, oracle_built, monkeypatch):
    monkeypatch.setenv("KSW_B200_REF", "1")
    for name, (cs, want) in K.load_chain_golden().items():
        assert K.regs_equal(K.run_chain_gpu(gpu_ctx, cs), want), name
        rnd = K.run_chain_driver(gpu_ctx.lib, cs, ctx=gpu_ctx.ctx, rounds=True)
        assert K.regs_equal(rnd[:2], want), ("rounds", name)


def test_shared_queue_merges_submissions_of_many_threads(oracle_built):
    """ksw_b200_queue_t (cross-thread coalescing): twelve host threads submit extension and
    global-alignment batches at the same time; every thread must get exactly the results of its own jobs, and the
    server must have run fewer GPU batches than there were submissions."""
    import threading
    rng = np.random.default_rng(5)
    l_pac = 40_000
    genome = rng.integers(0, 4, l_pac).astype(np.uint8)
    pac = K.pack_pac(genome)
    q = B.KswQueue(0)
    q.ref_set(pac, l_pac)
    cfg = K.make_cfg()
    n_thr, rounds = 12, 4

    def make(seed):
        r = np.random.default_rng(seed)
        n = int(r.integers(200, 900))
        rj = np.zeros(n, dtype=B.RJOB_DT)
        reads, qs, ts, at = [], [], [], 0
        for k in range(n):
            ql, tl = int(r.integers(20, 160)), int(r.integers(10, 260))
            beg = int(r.integers(0, l_pac - tl))
            t = genome[beg:beg + tl]
            qq = K.mutate(r, t, 0.02, 0.005, max_indel=3)[:ql]
            if len(qq) < ql:
                qq = np.concatenate([qq, r.integers(0, 4, ql - len(qq)).astype(np.uint8)])
            reads.append(qq); qs.append(qq); ts.append(t)
            rj[k]["q_off"], rj[k]["q_step"], rj[k]["t_pos"], rj[k]["t_step"] = at, 1, beg, 1
            rj[k]["qlen"], rj[k]["tlen"], rj[k]["h0"], rj[k]["w"] = ql, tl, int(r.integers(0, 80)), 100
            at += ql
        want = K.run_oracle(K._pools_from_lists(qs, ts, rj["h0"], rj["w"], cfg), threads=2)
        g = K.gen_global(int(r.integers(50, 300)), seed=seed + 1000, max_q=150)
        gwant = K.run_global_oracle(g, threads=2)
        return rj, np.concatenate(reads), want, g, gwant

    work = [[make(100 * t + i) for i in range(rounds)] for t in range(n_thr)]
    errors = []

    def body(t):
        try:
            for rj, qpool, want, g, gwant in work[t]:
                got = q.extend_ref(cfg, rj, qpool)
                mm = K.first_mismatch(want, got.view(K.RES_DT))
                assert mm is None, ("extension", t, mm)
                mm = K.global_mismatch(q.global_batch(g.cfg, g.jobs, g.qpool, g.tpool), gwant)
                assert mm is None, ("global", t, mm)
        except Exception as e:                       # noqa: BLE001 — reported by the main thread
            errors.append(e)

    th = [threading.Thread(target=body, args=(t,)) for t in range(n_thr)]
    [x.start() for x in th]
    [x.join() for x in th]
    st = q.stats()
    q.close()
    assert not errors, errors[0]
    assert st["submissions"] == 2 * n_thr * rounds and st["batches"] < st["submissions"], st
