"""Debug helper: a PE set through the B200-bound build with several environment variants, stderr summary lines printed.
   python scripts/dbg_rescue.py <pairs> <read_len> <rescue_frac> <sub> <indel> <indel_max> [ENV=VAL,ENV=VAL ...]"""
import os, sys, tempfile
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
n = int(sys.argv[1]); L = int(sys.argv[2]); rf = float(sys.argv[3]); sub = float(sys.argv[4]); indel = float(sys.argv[5]); imax = int(sys.argv[6])
variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in (sys.argv[7:] or [""])]
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "ref.fa")
    g = S.write_genome(fa, int(os.environ.get("DBG_GENOME", "10000000")), seed=1)
    S.bwa_index(fa)
    reads = [os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")]
    S.write_reads_fast(reads, g, n, L, seed=2, sub=sub, indel=indel, indel_max=imax, rescue_frac=rf, junk_frac=0.03 if rf else 0.0)
    for env in variants:
        err = S.bwa_mem(S.BWA_B200, fa, reads, os.path.join(d, "o.sam"), threads=int(os.environ.get("DBG_THREADS", "16")), env=dict(os.environ, **env))
        print(env)
        for ln in err.splitlines():
            if "Processed" in ln or "queue" in ln: print(ln[:900])
