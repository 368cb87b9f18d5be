"""Dev helper: first-chunk (cold start) and steady-state chunk times of stock vs B200-bound bwa mem, SE100."""
import os, sys, tempfile, time
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "ref.fa")
    g = S.write_genome(fa, 1_000_000, seed=1)
    S.bwa_index(fa)
    reads = [os.path.join(d, "r.fq")]
    S.write_reads_fast(reads, g, 4_000_000, 100, seed=2, sub=0.01, indel=0.001, indel_max=1)
    for binary, extra in ((S.BWA_STOCK, []), (S.BWA_B200, ["-b", "1"])):
        t0 = time.perf_counter()
        err = S.bwa_mem(binary, fa, reads, os.path.join(d, "o.sam"), threads=16, extra=extra)
        print(os.path.basename(binary), "wall", round(time.perf_counter() - t0, 2))
        for ln in err.splitlines():
            if "Processed" in ln: print("   ", ln[:330])
