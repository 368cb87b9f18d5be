"""Whole-program parity (BASELINE.json: "the resulting SAM must be byte-identical to stock bwa mem apart from the
@PG line").  CPU: the harness itself (fork == stock, as SURVEY.md §0 found).  GPU: `bwa mem` with pass 1 bound to the
B200 library (integration/bwamem_b200_glue.c) against stock 0.7.8, same -t, SE and PE, several -b / -t."""
import os
import re

import pytest

import samtest as S

need_bins = pytest.mark.skipif(not S.have_binaries(S.BWA_STOCK, S.BWA_FORK),
                               reason="oracle/_ref bwa binaries not built (reference sources not mounted at build time)")


@pytest.fixture(scope="module")
def small_se(tmp_path_factory):
    d = tmp_path_factory.mktemp("se")
    fa = str(d / "ref.fa")
    g = S.write_genome(fa, 300000, seed=1)
    S.bwa_index(fa)
    fq = str(d / "r.fq")
    S.write_reads_se(fq, g, 3000, 100, seed=2, n_rate=0.02)
    return d, fa, fq


@need_bins
def test_harness_fork_equals_stock_cpu(small_se):
    d, fa, fq = small_se
    S.bwa_mem(S.BWA_STOCK, fa, [fq], str(d / "stock.sam"), threads=2)
    S.bwa_mem(S.BWA_FORK, fa, [fq], str(d / "fork.sam"), threads=2, extra=["-b", "64"])
    ok, why = S.sam_equal(str(d / "stock.sam"), str(d / "fork.sam"))
    assert ok, why
    assert len(S.sam_body(str(d / "stock.sam"))) > 3000


need_b200 = pytest.mark.skipif(not S.have_binaries(S.BWA_STOCK, S.BWA_B200), reason="integration/_bin/bwa_b200 not built")


@pytest.mark.gpu
@need_b200
def test_se100_config1_byte_identical(tmp_path):
    # BASELINE.json configs[0]: single-end 100 bp, 1 % subst, 0.1 % indel, 1 Mbp reference
    fa = str(tmp_path / "ref.fa")
    g = S.write_genome(fa, 1_000_000, seed=12345)
    S.bwa_index(fa)
    fq = str(tmp_path / "r.fq")
    S.write_reads_fast([fq], g, 60000, 100, seed=12346, sub=0.01, indel=0.001)
    S.bwa_mem(S.BWA_STOCK, fa, [fq], str(tmp_path / "stock.sam"), threads=4)
    for tag, extra, thr in (("b1", [], 4), ("b5000", ["-b", "5000"], 4), ("t1", ["-b", "100000"], 1)):
        out = str(tmp_path / f"b200_{tag}.sam")
        if thr != 4:
            S.bwa_mem(S.BWA_STOCK, fa, [fq], str(tmp_path / "stock_t.sam"), threads=thr)
        S.bwa_mem(S.BWA_B200, fa, [fq], out, threads=thr, extra=extra)
        ok, why = S.sam_equal(str(tmp_path / ("stock.sam" if thr == 4 else "stock_t.sam")), out)
        assert ok, (tag, why)
    # every CIGAR through the miss path of the redirected ksw_global2 call (one GPU job per look-up): same SAM
    fq_small = str(tmp_path / "small.fq")
    with open(fq, "rb") as src, open(fq_small, "wb") as dst:
        dst.writelines(src.readlines()[:4 * 3000])
    S.bwa_mem(S.BWA_STOCK, fa, [fq_small], str(tmp_path / "stock_small.sam"), threads=4)
    err = S.bwa_mem(S.BWA_B200, fa, [fq_small], str(tmp_path / "b200_miss.sam"), threads=4, env=dict(os.environ, KSW_B200_CIGAR="2"))
    ok, why = S.sam_equal(str(tmp_path / "stock_small.sam"), str(tmp_path / "b200_miss.sam"))
    assert ok, ("miss path", why)
    m = re.findall(r"global alignments so far: (\d+) computed ahead, (\d+) hits, (\d+) misses", err)
    assert m and int(m[-1][0]) == 0 and int(m[-1][1]) == 0 and int(m[-1][2]) > 300, m
    # the opt-in pair kernel (two jobs per lane) must give the same SAM
    out = str(tmp_path / "b200_pair.sam")
    S.bwa_mem(S.BWA_B200, fa, [fq], out, threads=4, extra=["-b", "20000"], env=dict(os.environ, KSW_B200_PAIR="1"))
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), out)
    assert ok, ("pair kernel", why)


@pytest.mark.gpu
@need_b200
def test_pe150_byte_identical(tmp_path):
    fa = str(tmp_path / "ref.fa")
    g = S.write_genome(fa, 2_000_000, seed=21, n_contigs=3)
    S.bwa_index(fa)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    S.write_reads_fast([f1, f2], g, 40000, 150, seed=22, ins_mean=375, ins_sd=37)
    S.bwa_mem(S.BWA_STOCK, fa, [f1, f2], str(tmp_path / "stock.sam"), threads=4)
    S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200.sam"), threads=4, extra=["-b", "3000"])
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200.sam"))
    assert ok, why
    # the rounds scheduler (strategy B: only the seeds the reference extends) must give the same SAM
    env = dict(os.environ, KSW_B200_SCHED="rounds")
    S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_rounds.sam"), threads=4, extra=["-b", "3000"], env=env)
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200_rounds.sam"))
    assert ok, ("rounds", why)
    # the CIGAR look-ahead (GPU ksw_global2 behind the reference's bwa_gen_cigar2) must have served pass 2 ...
    err = S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_again.sam"), threads=4, extra=["-b", "3000"])
    m = re.findall(r"global alignments so far: (\d+) computed ahead, (\d+) hits, (\d+) misses", err)
    assert m, err[-500:]
    ahead, hits, misses = (int(x) for x in m[-1])
    assert ahead > 10000 and hits >= ahead * 0.9 and misses < ahead * 0.01, m[-1]    # (most 150 bp alignments need no DP at all: bwa.c:110)
    # ... and switching it off (pass 2 entirely on the host, as in the reference) must give the same SAM
    env = dict(os.environ, KSW_B200_CIGAR="0")
    S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_nocig.sam"), threads=4, extra=["-b", "3000"], env=env)
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200_nocig.sam"))
    assert ok, ("cigar look-ahead off", why)


@pytest.mark.gpu
@need_b200
def test_pe250_high_indel_byte_identical(tmp_path):
    # BASELINE.json configs[3] shape: 2x250, 3 % subst, 2 % indel events of length U[1,12] (wide bands, z-drop heavy)
    fa = str(tmp_path / "ref.fa")
    g = S.write_genome(fa, 2_000_000, seed=31)
    S.bwa_index(fa)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    S.write_reads_fast([f1, f2], g, 15000, 250, seed=32, sub=0.03, indel=0.004, indel_max=12)
    # plus reads with several indel events each (the per-base simulator)
    f1b, f2b = str(tmp_path / "s1.fq"), str(tmp_path / "s2.fq")
    S.write_reads_pe(f1b, f2b, g, 3000, 250, seed=33, sub=0.03, indel=0.02, indel_max=12)
    for a_, b_ in ((f1, f1b), (f2, f2b)):
        with open(a_, "ab") as out, open(b_, "rb") as src:
            out.write(src.read())
    S.bwa_mem(S.BWA_STOCK, fa, [f1, f2], str(tmp_path / "stock.sam"), threads=4)
    S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200.sam"), threads=4)
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200.sam"))
    assert ok, why
    env = dict(os.environ, KSW_B200_SCHED="rounds")
    S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_rounds.sam"), threads=4, env=env)
    ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200_rounds.sam"))
    assert ok, ("rounds", why)


@pytest.mark.gpu
@need_b200
def test_pe_mate_rescue_byte_identical(tmp_path):
    """Pairs in which one mate cannot be seeded (18 % extra substitutions) or is junk: pass 2 finds it by mate rescue (mem_matesw ->
    ksw_align2), which the B200-bound build computes ahead on the GPU.  The SAM must not change, most look-ups
    must hit, and with the look-ahead off (the reference's own SSE2 ksw_align2) the SAM is the same again."""
    import re
    fa = str(tmp_path / "ref.fa")
    g = S.write_genome(fa, 2_000_000, seed=41)
    S.bwa_index(fa)
    for L, n in ((150, 30000), (250, 12000)):                                   # byte kernel (150 * 1 < 250) and 16-bit kernel
        f1, f2 = str(tmp_path / f"r1_{L}.fq"), str(tmp_path / f"r2_{L}.fq")
        S.write_reads_fast([f1, f2], g, n, L, seed=42 + L, sub=0.01, indel=0.002, indel_max=3, rescue_frac=0.2, junk_frac=0.05)
        S.bwa_mem(S.BWA_STOCK, fa, [f1, f2], str(tmp_path / "stock.sam"), threads=4)
        err = S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200.sam"), threads=4)
        ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200.sam"))
        assert ok, (L, why)
        mm = re.findall(r"mate-rescue alignments: look-ahead [\d.]+ thread-s, (\d+) computed ahead, (\d+) hits, (\d+) misses", err)
        assert mm, err[-600:]
        ahead, hits, misses = (int(x) for x in mm[-1])
        assert ahead > n // 10 and hits > n // 10 and misses <= hits // 50, (L, ahead, hits, misses)
        S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_off.sam"), threads=4, env=dict(os.environ, KSW_B200_RESCUE="0"))
        ok, why = S.sam_equal(str(tmp_path / "stock.sam"), str(tmp_path / "b200_off.sam"))
        assert ok, ("rescue look-ahead off", L, why)
        if L == 150:
            # the options that change what mem_sam_pe does with rescue: -S (no rescue), -P (rescue but no pairing), -I (given
            # insert-size distribution instead of mem_pestat's), -m (fewer rescue rounds)
            for extra in (["-S"], ["-P"], ["-I", "380,40"], ["-m", "1"]):
                S.bwa_mem(S.BWA_STOCK, fa, [f1, f2], str(tmp_path / "stock_x.sam"), threads=4, extra=extra)
                S.bwa_mem(S.BWA_B200, fa, [f1, f2], str(tmp_path / "b200_x.sam"), threads=4, extra=extra)
                ok, why = S.sam_equal(str(tmp_path / "stock_x.sam"), str(tmp_path / "b200_x.sam"))
                assert ok, (extra, why)
