"""CPU tests of the oracle (oracle/ksw_oracle.c): pinned to the committed golden vectors (outputs of the
reference's own ksw.c) and, where the compiled reference is present, differentially against it."""
import numpy as np
import pytest

import kswtest as K


@pytest.fixture(scope="module")
def golden(oracle_built):
    return K.load_golden()


def test_oracle_matches_golden_vectors(golden):
    assert set(golden) >= {"adversarial", "fuzz_default", "fuzz_asym", "fuzz_bwasw", "config2", "highindel250"}
    for name, (b, want) in golden.items():
        got = K.run_oracle(b, threads=2)
        assert K.first_mismatch(want, got) is None, name


@pytest.mark.skipif(not K.have_ref(), reason="oracle/_ref not built (reference sources not mounted)")
def test_oracle_matches_compiled_reference_on_fresh_fuzz(oracle_built):
    for b in (K.gen_fuzz(6000, seed=201), K.gen_config2(20000, seed=202),
              K.gen_fuzz(3000, seed=203, cfg=K.make_cfg(a=3, b=5, o_del=2, e_del=3, o_ins=9, e_ins=2, zdrop=10, end_bonus=0)),
              K.gen_fuzz(1500, seed=204, max_q=1200, w_choices=(5, 100, 400))):
        assert K.first_mismatch(K.run_ref(b, threads=4), K.run_oracle(b, threads=4)) is None


@pytest.mark.skipif(not K.have_ref(), reason="oracle/_ref not built")
def test_golden_fixture_is_reproducible_from_reference(golden):
    for name, (b, want) in golden.items():
        assert K.first_mismatch(want, K.run_ref(b, threads=2)) is None, name


def test_known_answers_closed_form(oracle_built):
    # identical sequences, tlen == qlen == L: score = h0 + L*a, ends at (L, L), gscore == score
    cfg = K.make_cfg()
    L = 50
    seq = np.random.default_rng(4).integers(0, 4, L).astype(np.uint8)
    jobs = np.zeros(3, dtype=K.JOB_DT)
    jobs["qlen"] = L; jobs["tlen"] = L; jobs["w"] = 100; jobs["h0"] = [8, 19, 100]
    r = K.run_oracle(K.Batch(cfg, jobs, seq, seq))
    assert list(r["score"]) == [8 + L, 19 + L, 100 + L]
    assert (r["qle"] == L).all() and (r["tle"] == L).all() and (r["gscore"] == r["score"]).all()
    # tlen == 0: nothing runs; the reference returns h0 with all-zero ends and gscore -1 (ksw.c:408-410,469-475)
    jobs2 = np.zeros(1, dtype=K.JOB_DT)
    jobs2["qlen"] = 10; jobs2["tlen"] = 0; jobs2["w"] = 100; jobs2["h0"] = 33
    r2 = K.run_oracle(K.Batch(cfg, jobs2, seq, seq))
    assert tuple(int(r2[f][0]) for f in K.RES_DT.names) == (33, 0, 0, 0, -1, 0)
    # negative h0 is clamped to 0 (ksw.c:384)
    # negative h0 is clamped to 0 (ksw.c:384): same answer as h0 == 0
    jobs3 = jobs[:2].copy(); jobs3["h0"] = [-5, 0]
    r3 = K.run_oracle(K.Batch(cfg, jobs3, seq, seq))
    assert all(int(r3[f][0]) == int(r3[f][1]) for f in K.RES_DT.names)
    # periodic sequence with h0 == 0: row 0 scores 1 on every 4th column (no zero guard, ksw.c:430), the
    # last of them wins the tie (ksw.c:434) and the band collapses onto it (ksw.c:463-466) -> score 2
    per = (np.arange(L) % 4).astype(np.uint8)
    jobs4 = jobs[:1].copy(); jobs4["h0"] = 0
    r4 = K.run_oracle(K.Batch(cfg, jobs4, per, per))
    assert tuple(int(r4[f][0]) for f in K.RES_DT.names) == (2, 50, 2, 2, 2, 48)


def test_visited_cells_counter(oracle_built):
    b = K.gen_config2(2000, seed=9)
    _, cells = K.run_oracle(b, want_cells=True)
    nominal = b.jobs["qlen"].astype(np.int64) * b.jobs["tlen"]
    assert (cells > 0).all() and (cells <= nominal).all()
    # with w >= qlen, no z-drop hit and no zero in the band a perfect match visits a growing band
    assert 0.5 < cells.sum() / nominal.sum() < 0.95
