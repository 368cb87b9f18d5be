"""mem_chain2aln level: the oracle restatement against the golden regions produced by
the reference's own mem_chain2aln, and (GPU) the product's batched plan/run/replay driver against both."""
import ctypes as C

import numpy as np
import pytest

import kswtest as K


def test_oracle_chain2aln_matches_golden(oracle_built):
    g = K.load_chain_golden()
    assert set(g) >= {"default", "narrow_w", "hi_indel250", "asym", "with_n"}
    for name, (cs, want) in g.items():
        got = K.run_chain_oracle(cs)
        assert K.regs_equal(got[:2], want), name
    # the narrow-band set must exercise the MAX_BAND_TRY retry (bwamem.c:818-829): a->w doubles
    assert int(g["narrow_w"][1][0]["w"].max()) == 24


@pytest.mark.skipif(not K.have_bwa_ref(), reason="oracle/_ref/libbwa_ref.so not built")
def test_oracle_chain2aln_matches_reference_on_fresh_sets(oracle_built):
    for seed, kw in ((21, {}), (22, dict(opt=K.make_ext_opt(w=10), indel=0.012, max_indel=16)),
                     (23, dict(sub=0.05, indel=0.02, max_indel=12, read_lens=(250,))),
                     (24, dict(opt=K.make_ext_opt(a=3, b=2, o_del=9, e_del=1, o_ins=3, e_ins=2, pen_clip5=0, pen_clip3=11, zdrop=25)))):
        cs = K.gen_chains(250, seed=seed, **kw)
        assert K.regs_equal(K.run_chain_oracle(cs)[:2], K.run_chain_ref(cs)), seed


def test_host_driver_both_schedulers_against_golden_cpu(oracle_built):
    """bwamem_ext.c (speculate-and-replay, and the rounds scheduler) linked against an oracle-backed stub of the GPU
    entry: the host logic is checked on the CPU against the regions of the reference's own mem_chain2aln."""
    lib = K.ext_emu_lib()
    for name, (cs, want) in K.load_chain_golden().items():
        spec = K.run_chain_driver(lib, cs)
        assert K.regs_equal(spec[:2], want), ("speculative", name)
        jobs_before = C.c_int64.in_dll(lib, "ext_stub_jobs").value
        rnd = K.run_chain_driver(lib, cs, rounds=True)
        assert K.regs_equal(rnd[:2], want), ("rounds", name)
        # the rounds scheduler runs exactly the DP calls the sequential reference makes
        assert rnd[2] == K.run_chain_oracle(cs)[2], name
        assert C.c_int64.in_dll(lib, "ext_stub_jobs").value - jobs_before == rnd[2]


def test_ref_slice_known_answers():
    import bwa_mem_quickassist_b200 as B
    lib = B.load_library()
    lib.b200_get_ref_slice.restype = C.c_int64
    lib.b200_get_ref_slice.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    rng = np.random.default_rng(2)
    L = 1003
    genome = rng.integers(0, 4, L).astype(np.uint8)
    pac = K.pack_pac(genome)
    dbl = np.concatenate([genome, 3 - genome[::-1]]).astype(np.uint8)      # bntseq.c:365-369
    out = np.zeros(2 * L + 8, dtype=np.uint8)
    for beg, end in [(0, L), (5, 77), (L - 3, L), (L, 2 * L), (L + 11, L + 400), (2 * L - 9, 2 * L), (700, 300), (-5, 40),
                     (2 * L - 5, 2 * L + 30)]:
        n = lib.b200_get_ref_slice(L, K._ptr(pac), beg, end, K._ptr(out))
        b, e = min(beg, end), max(beg, end)
        b, e = max(b, 0), min(e, 2 * L)
        assert n == e - b and (out[:n] == dbl[b:e]).all(), (beg, end)
    # a window that bridges the forward/reverse boundary yields nothing (bntseq.c:374)
    assert lib.b200_get_ref_slice(L, K._ptr(pac), L - 10, L + 10, K._ptr(out)) == 0


@pytest.mark.gpu
def test_batched_driver_matches_golden_and_oracle(gpu_ctx, oracle_built):
    for name, (cs, want) in K.load_chain_golden().items():
        got = K.run_chain_gpu(gpu_ctx, cs)
        assert K.regs_equal(got, want), name
    for seed, kw in ((31, {}), (32, dict(opt=K.make_ext_opt(w=10), indel=0.012, max_indel=16)),
                     (33, dict(sub=0.05, indel=0.02, max_indel=12, read_lens=(250,))), (34, dict(n_frac=0.05))):
        cs = K.gen_chains(600, seed=seed, **kw)
        want = K.run_chain_oracle(cs)
        assert K.regs_equal(K.run_chain_gpu(gpu_ctx, cs), want[:2]), seed
        rnd = K.run_chain_driver(gpu_ctx.lib, cs, ctx=gpu_ctx.ctx, rounds=True)       # strategy B on the GPU
        assert K.regs_equal(rnd[:2], want[:2]) and rnd[2] == want[2], seed


def test_host_driver_device_reference_mode_cpu(oracle_built, monkeypatch):
    """SURVEY.md 8(f) rank 3 on the host side: in device-reference mode the driver materialises no reference window and
    names every target by its coordinate in the doubled reference space (ksw_b200_rjob_t).  Linked against the stub,
    which slices the .pac the way bns_get_seq does, it must still produce the reference's regions."""
    monkeypatch.setenv("KSW_B200_REF", "1")
    lib = K.ext_emu_lib()
    for name, (cs, want) in K.load_chain_golden().items():
        assert K.regs_equal(K.run_chain_driver(lib, cs)[:2], want), ("speculative, device reference", name)
        assert K.regs_equal(K.run_chain_driver(lib, cs, rounds=True)[:2], want), ("rounds, device reference", name)
