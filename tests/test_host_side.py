"""CPU tests of the product's host side: the packer and the fast kernel's per-lane source compiled for the
CPU (tests/emu) against the golden vectors and the oracle; the C-ABI library loads and exports every symbol
include/ksw_b200.h declares; the Python binding refuses to run without a GPU (no silent fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import kswtest as K


def test_fast_lane_source_matches_golden_vectors(oracle_built):
    for name, (b, want) in K.load_golden().items():
        got, n_fast = K.run_emu(b)
        assert n_fast > 0, name
        sel = got["score"] != np.iinfo(np.int32).min
        assert sel.sum() == n_fast
        assert K.first_mismatch(want[sel], got[sel]) is None, name


def test_fast_lane_source_fuzz_vs_oracle(oracle_built):
    cases = [K.gen_fuzz(4000, seed=301), K.gen_config2(4000, seed=302),
             K.gen_fuzz(3000, seed=303, cfg=K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3, zdrop=15, end_bonus=2)),
             K.gen_fuzz(1500, seed=304, max_q=512, w_choices=(3, 64, 300))]
    rng = np.random.default_rng(1)
    mat = rng.integers(-6, 5, 25).astype(np.int8); mat[[0, 6, 12, 18]] = [3, 4, 5, 2]
    cases.append(K.gen_fuzz(3000, seed=305, cfg=K.make_cfg(mat=mat, o_del=3, e_del=1, o_ins=2, e_ins=2, zdrop=40)))
    for b in cases:
        want = K.run_oracle(b)
        got, n_fast = K.run_emu(b)
        sel = got["score"] != np.iinfo(np.int32).min
        assert n_fast == sel.sum() and n_fast > 0
        mm = K.first_mismatch(want[sel], got[sel])
        assert mm is None, mm


def test_packer_simd_path_writes_the_same_bytes_as_the_word_path(oracle_built):
    """ksw_pack.cpp packs 64 byte codes per step where the CPU has AVX-512BW; that path, the word-at-a-time path and
    a numpy restatement of the layout (ksw_dev.cuh: base k of a sequence in word k/16 at bits 2(k%16), N as 0 + mask
    bit) must agree on every byte: job records, 2-bit pool, N side pool, class counts."""
    cases = [K.gen_fuzz(3000, seed=311), K.gen_fuzz(600, seed=312, max_q=700), K.gen_config2(2000, seed=313),
             K.gen_fuzz(1500, seed=314, max_q=70, n_frac=0.3), K.gen_adversarial()]
    for b in cases:
        dj_s, pool_s, nm_s, cn_s = K.run_packer(b, force_words=False)
        dj_w, pool_w, nm_w, cn_w = K.run_packer(b, force_words=True)
        assert cn_s == cn_w
        assert np.array_equal(dj_s, dj_w) and np.array_equal(pool_s, pool_w) and np.array_equal(nm_s, nm_w)
        # independent check of the pool: every job's query and target words
        seq_off, flags, nmask_off = dj_s[:, 0].astype(np.int64), dj_s[:, 6], dj_s[:, 7].astype(np.int64)
        for k in range(0, b.n, 7):
            j = b.jobs[k]
            at = seq_off[k] * 4
            noff = nmask_off[k]
            for bit, pool_in, off, ln in ((1, b.qpool, int(j["q_off"]), int(j["qlen"])), (2, b.tpool, int(j["t_off"]), int(j["tlen"]))):
                codes = pool_in[off:off + ln].astype(np.uint32)
                isn = codes > 3
                pad = np.zeros((-ln) % 16, dtype=np.uint32)
                c2 = np.concatenate([np.where(isn, 0, codes), pad]).reshape(-1, 16)
                want = (c2 << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1).astype(np.uint32)
                assert np.array_equal(pool_s[at:at + len(want)], want), (k, bit)
                at += len(want)
                assert bool(flags[k] & bit) == bool(isn.any()), (k, bit)
                if isn.any():
                    mpad = np.zeros((-ln) % 32, dtype=np.uint32)
                    mw = (np.concatenate([isn.astype(np.uint32), mpad]).reshape(-1, 32) << np.arange(32, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
                    assert np.array_equal(nm_s[noff:noff + len(mw)], mw), (k, bit)
                    noff += len(mw)
            units = int(dj_s[k + 1, 0]) - int(seq_off[k]) if k + 1 < b.n else (len(pool_s) // 4 - int(seq_off[k]))
            assert not pool_s[at:(int(seq_off[k]) + units) * 4].any(), k      # padding words are zero


def test_packer_routes_out_of_range_jobs_to_generic(oracle_built):
    b = K.gen_fuzz(300, seed=306, h0_max=60000)
    got, n_fast = K.run_emu(b)
    big = b.jobs["h0"].astype(np.int64) + b.jobs["qlen"] > 20000
    assert big.any()
    assert (got["score"][big] == np.iinfo(np.int32).min).all()          # not taken by the s16 kernel
    long_q = K.gen_fuzz(50, seed=307, max_q=2000)
    got2, _ = K.run_emu(long_q)
    assert (got2["score"][long_q.jobs["qlen"] > 512] == np.iinfo(np.int32).min).all()


def test_fast_lane_source_on_class_boundaries(oracle_built):
    for cfg in (K.make_cfg(), K.make_cfg(a=3, b=5, o_del=9, e_del=2, o_ins=4, e_ins=3, zdrop=200, end_bonus=7)):
        b = K.gen_boundaries(cfg=cfg)
        want = K.run_oracle(b)
        got, n_fast = K.run_emu(b)
        sel = got["score"] != np.iinfo(np.int32).min
        assert 0 < n_fast < b.n and K.run_emu.last_keyed > 0          # all three routes are exercised
        mm = K.first_mismatch(want[sel], got[sel])
        assert mm is None, mm
        # the largest score the keyed class may see is exactly its bound
        a = int(cfg.mat[0])
        assert (want["score"][sel] <= 20000).all()


def _header_functions():
    txt = open(os.path.join(K.ROOT, "include", "ksw_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ksw_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    import bwa_mem_quickassist_b200 as B
    lib = B.load_library()
    names = _header_functions()
    assert "ksw_extend" in names and "ksw_extend2" in names and "ksw_b200_extend_batch" in names
    for nm in names:
        assert hasattr(lib, nm), f"libksw_b200.so does not export {nm}"


def test_band_clamp_matches_oracle(oracle_built):
    import bwa_mem_quickassist_b200 as B
    lib, orc = B.load_library(), K.oracle_lib()
    rng = np.random.default_rng(3)
    for _ in range(3000):
        a, bb = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        mat = K.fill_scmat(a, bb)
        args = [int(rng.integers(1, 400))]
        od, ed, oi, ei = (int(rng.integers(0, 12)), int(rng.integers(1, 5)), int(rng.integers(0, 12)), int(rng.integers(1, 5)))
        w, eb = int(rng.integers(1, 300)), int(rng.integers(0, 20))
        got = lib.ksw_b200_clamp_w(args[0], mat.ctypes.data_as(C.c_void_p), od, ed, oi, ei, w, eb)
        want = orc.ksw_oracle_clamp_w(args[0], 5, mat.ctypes.data_as(C.c_void_p), od, ed, oi, ei, w, eb)
        assert got == want


def test_no_cpu_fallback_without_gpu():
    import bwa_mem_quickassist_b200 as B
    lib = B.load_library()
    if lib.ksw_b200_device_count() > 0:
        pytest.skip("a GPU is visible: the no-device behaviour cannot be shown here")
    with pytest.raises(B.KswB200Error):
        B.KswB200(0)


def test_product_does_not_touch_oracle():
    """The product sources must not reference oracle/ or the emulation."""
    pkg = os.path.join(K.ROOT, "bwa_mem_quickassist_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                s = open(os.path.join(dp, f)).read()
                assert "libksw_oracle" not in s and "ksw_oracle_" not in s and "libksw_ref" not in s, f


def test_lane_source_under_address_and_ub_sanitizers(oracle_built):
    """compute-sanitizer is closed on the GPU pool, so the packer and the kernel's per-lane source run under ASan + UBSan on
    the host with every array at exactly the size the launcher reserves (tests/emu/asan_fuzz.cpp)."""
    import subprocess
    emu = os.path.join(K.ROOT, "tests", "emu")
    subprocess.run(["make", "-C", emu, "asan_fuzz", "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
    for seed in (11, 12):
        out = subprocess.run([os.path.join(emu, "asan_fuzz"), "4000", str(seed)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "0 mismatches" in out.stdout, out.stdout + out.stderr[-2000:]
