#!/usr/bin/env python
"""BASELINE config 5 (throughput sweep): synthetic 2x150 bp reads against a synthetic genome, stock `bwa mem -t <all host
cores>` once, then the B200-bound build with the read batches sharded over 1 / 2 / 4 / 8 GPUs (KSW_B200_GPUS); wall
time, steady-state reads/s and the SAM diff (minus @PG) at every N.  Data, index and the stock run are made once.
Seeding stays on the host (north star), so the whole program is host-bound and flat in N: this is the record of that,
and of SAM identity with the batches spread over several GPUs.
   python scripts/bench_sweep.py --pairs 4000000 --genome 50000000 --gpus 1,2,4,8 --out profiles/x.json"""
import argparse, json, os, subprocess, sys, tempfile, time

sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S
from bench_bwamem import chunk_times, timed


def same_sam(a, b):
    """cmp of the two files without their @PG lines (C tools: the files are GBs)"""
    r = subprocess.run(["bash", "-c", f"cmp <(grep -v '^@PG' {a}) <(grep -v '^@PG' {b})"], capture_output=True, text=True)
    return r.returncode == 0, r.stdout.strip()[:200]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4_000_000)
    ap.add_argument("--genome", type=int, default=50_000_000)
    ap.add_argument("--length", type=int, default=150)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rows = {"config": f"config5-shape: {2 * a.pairs} reads 2x{a.length} bp vs {a.genome / 1e6:.0f} Mbp genome, -t {a.threads}", "runs": []}
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "ref.fa")
        t0 = time.perf_counter()
        g = S.write_genome(fa, a.genome, seed=1)
        S.bwa_index(fa)
        reads = [os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")]
        S.write_reads_fast(reads, g, a.pairs, a.length, seed=2, sub=0.01, indel=0.001, indel_max=1)
        rows["setup_s"] = round(time.perf_counter() - t0, 1)
        stock_sam = os.path.join(d, "stock.sam")
        t_stock, e_stock = timed(S.BWA_STOCK, fa, reads, stock_sam, a.threads)
        cs = chunk_times(e_stock)
        ss = sum(r for r, _ in cs[1:]) / max(sum(t for _, t in cs[1:]), 1e-9) if len(cs) > 1 else None
        rows["stock"] = {"wall_s": round(t_stock, 2), "reads_per_s": round(2 * a.pairs / t_stock), "steady_reads_per_s": round(ss) if ss else None}
        print(json.dumps(rows["stock"]), flush=True)
        for n in [int(x) for x in a.gpus.split(",")]:
            out = os.path.join(d, f"b200_{n}.sam")
            env = dict(os.environ, KSW_B200_GPUS=str(n))
            t_b, e_b = timed(S.BWA_B200, fa, reads, out, a.threads, env=env)
            cb = chunk_times(e_b)
            sb = sum(r for r, _ in cb[1:]) / max(sum(t for _, t in cb[1:]), 1e-9) if len(cb) > 1 else None
            ok, why = same_sam(stock_sam, out)
            os.remove(out)
            row = {"n_gpus": n, "wall_s": round(t_b, 2), "reads_per_s": round(2 * a.pairs / t_b), "steady_reads_per_s": round(sb) if sb else None,
                   "sam_identical_minus_PG": bool(ok), "diff": why}
            print(json.dumps(row), flush=True)
            rows["runs"].append(row)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
