"""CPU tests of the PAIR kernel's per-lane source (bwa_mem_quickassist_b200/csrc/ksw_pair_core.h: two extension jobs
per lane, one in each half of the s16x2 registers) compiled for the CPU with software DPX (tests/emu), against the
golden vectors of the compiled reference and against the oracle.  The emulated lanes refill a slot as soon as its job
ends, so jobs meet partners at every phase: overlapping bands, disjoint bands, a fresh job beside a half-finished one,
and a lone job with the other slot empty."""
import numpy as np
import pytest

import kswtest as K

NONE = np.iinfo(np.int32).min


def _check(b, want=None, **kw):
    if want is None:
        want, wcells = K.run_oracle(b, want_cells=True)
    else:
        _, wcells = K.run_oracle(b, want_cells=True)
    got, cells, n_pair, rows = K.run_pair_emu(b, **kw)
    sel = got["score"] != NONE
    assert sel.sum() == n_pair
    mm = K.first_mismatch(want[sel], got[sel])
    assert mm is None, mm
    assert (wcells[sel] == cells[sel]).all()
    return sel, n_pair


def test_pair_lane_matches_golden_vectors(oracle_built):
    total = 0
    for name, (b, want) in K.load_golden().items():
        _, n_pair = _check(b, want)
        total += n_pair
    assert total > 1000


@pytest.mark.parametrize("lanes,order", [(1, 0), (3, 0), (2, 1)])
def test_pair_lane_fuzz_vs_oracle(oracle_built, lanes, order):
    cases = [K.gen_fuzz(4000, seed=401, max_q=124), K.gen_config2(3000, seed=402),
             K.gen_fuzz(3000, seed=403, max_q=124, cfg=K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3, zdrop=15, end_bonus=2)),
             K.gen_fuzz(2000, seed=404, max_q=124, w_choices=(1, 3, 64, 300), h0_max=60),
             K.gen_fuzz(2000, seed=405, max_q=124, cfg=K.make_cfg(zdrop=-1))]
    rng = np.random.default_rng(1)
    mat = rng.integers(-6, 5, 25).astype(np.int8); mat[[0, 6, 12, 18]] = [3, 4, 5, 2]
    cases.append(K.gen_fuzz(3000, seed=406, max_q=100, h0_max=80, cfg=K.make_cfg(mat=mat, o_del=3, e_del=1, o_ins=2, e_ins=2, zdrop=40)))
    for b in cases:
        sel, n_pair = _check(b, lanes=lanes, order=order)
        assert n_pair > b.n // 4


def test_pair_lane_never_takes_a_query_with_n(oracle_built):
    b = K.gen_fuzz(3000, seed=407, max_q=100, h0_max=50, n_frac=0.05)
    sel, _ = _check(b)
    has_qn = np.array([(b.qpool[int(j["q_off"]):int(j["q_off"]) + int(j["qlen"])] == 4).any() for j in b.jobs])
    assert has_qn.any() and not (sel & has_qn).any()
    # class 0 is N-free altogether (the keyed one-job-per-lane kernel skips the target's N mask): no target N either
    has_tn = np.array([(b.tpool[int(j["t_off"]):int(j["t_off"]) + int(j["tlen"])] == 4).any() for j in b.jobs])
    assert has_tn.any() and not (sel & has_tn).any()


def test_pair_lane_adversarial_and_boundaries(oracle_built):
    for cfg in (K.make_cfg(), K.make_cfg(a=3, b=5, o_del=9, e_del=2, o_ins=4, e_ins=3, zdrop=200, end_bonus=7)):
        _check(K.gen_adversarial(cfg=cfg), lanes=1)
        _check(K.gen_boundaries(cfg=cfg), lanes=2)
