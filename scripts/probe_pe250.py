"""Dev helper: phase timers of the B200-bound bwa mem on the config-4 shape (PE250 high-indel)."""
import os, sys, tempfile, time
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "ref.fa")
    g = S.write_genome(fa, 10_000_000, seed=1)
    S.bwa_index(fa)
    reads = [os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")]
    S.write_reads_fast(reads, g, 600_000, 250, seed=2, sub=0.03, indel=0.002, indel_max=12)
    for env in ({}, {"KSW_B200_CIGAR": "0"}):
        t0 = time.perf_counter()
        err = S.bwa_mem(S.BWA_B200, fa, reads, os.path.join(d, "o.sam"), threads=16, extra=["-b", "1"], env=dict(os.environ, **env))
        print(env, "wall", round(time.perf_counter() - t0, 2))
        for ln in err.splitlines():
            if "Processed" in ln: print("   ", ln[:420])
