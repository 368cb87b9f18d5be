"""Debug helper: the rescue-heavy PE150 set through the B200-bound build, stderr summary lines printed."""
import os, sys, tempfile
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "ref.fa")
    g = S.write_genome(fa, 10_000_000, seed=1)
    S.bwa_index(fa)
    reads = [os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")]
    S.write_reads_fast(reads, g, n, 150, seed=2, sub=0.01, indel=0.001, indel_max=1, rescue_frac=0.15, junk_frac=0.03)
    for env in ({}, {"KSW_B200_RESCUE": "0"}):
        err = S.bwa_mem(S.BWA_B200, fa, reads, os.path.join(d, "o.sam"), threads=16, env=dict(os.environ, **env))
        print(env)
        for ln in err.splitlines():
            if "Processed" in ln or "queue" in ln: print(ln[:900])
    err = S.bwa_mem(S.BWA_STOCK, fa, reads, os.path.join(d, "o.sam"), threads=16)
    for ln in err.splitlines():
        if "Processed" in ln: print(ln[:300])
