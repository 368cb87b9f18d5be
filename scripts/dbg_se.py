"""Debug helper: the SE100 config through the B200-bound build (stderr summary lines), twice, then stock."""
import os, sys, tempfile
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in (sys.argv[2:] or ["", ""])]
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "ref.fa")
    g = S.write_genome(fa, 1_000_000, seed=1)
    S.bwa_index(fa)
    reads = [os.path.join(d, "r.fq")]
    S.write_reads_fast(reads, g, n, 100, seed=2, sub=0.01, indel=0.001, indel_max=1)
    for env in variants:
        err = S.bwa_mem(S.BWA_B200, fa, reads, os.path.join(d, "o.sam"), threads=16, env=dict(os.environ, **env))
        print(env)
        for ln in err.splitlines():
            if "Processed" in ln: print(ln[:520])
    err = S.bwa_mem(S.BWA_STOCK, fa, reads, os.path.join(d, "o.sam"), threads=16)
    for ln in err.splitlines():
        if "Processed" in ln: print(ln[:100])
