#!/usr/bin/env python
"""Whole-program side-by-side: stock `bwa mem -t N` (bwa 0.7.8 from the tarball) and the
B200-bound build on synthetic data of the BASELINE.json shapes, same flags, same -t; wall time, reads/s, and the
SAM diff (minus @PG).  Seeding/chaining/SAM stay on the host (Amdahl: SURVEY.md §6), so this is a report, not the
metric.  Usage: python scripts/bench_bwamem.py [--scale 1.0] [--threads N] [--out profiles/x.json]"""
import argparse, json, os, sys, tempfile, time

sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import samtest as S


def timed(binary, fa, reads, out, threads, extra=(), env=None):
    t0 = time.perf_counter()
    err = S.bwa_mem(binary, fa, reads, out, threads=threads, extra=extra, env=env)
    return time.perf_counter() - t0, err


def chunk_times(err):
    """(reads, real seconds) of every mem_process_seqs chunk, from the reference's own log line (bwamem.c:1320)."""
    out = []
    for ln in err.splitlines():
        if "Processed" in ln and "real sec" in ln:
            f = ln.split()
            out.append((int(f[f.index("Processed") + 1]), float(f[f.index("real") - 1])))
    return out


def cigar_stats(err):
    """(computed ahead, hits, misses) of the CIGAR look-ahead, from the glue's last chunk line."""
    import re
    m = None
    for ln in err.splitlines():
        x = re.search(r"global alignments so far: (\d+) computed ahead, (\d+) hits, (\d+) misses", ln)
        m = x or m
    return [int(v) for v in m.groups()] if m else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--sched", default="speculate", choices=["speculate", "rounds"], help="KSW_B200_SCHED of the B200-bound build")
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="substring of the config name to run alone (e.g. PE150)")
    ap.add_argument("--genome", type=int, default=0, help="override the genome length of the selected configs (bp)")
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads (SE) / pairs (PE) of the selected configs")
    ap.add_argument("--no-cigar-off", action="store_true", help="skip the extra run with the CIGAR look-ahead off")
    ap.add_argument("--rescue-off", action="store_true", help="one more run with the mate-rescue look-ahead off (KSW_B200_RESCUE=0)")
    ap.add_argument("--gpus", type=int, default=0, help="KSW_B200_GPUS of the B200-bound build (0 = all visible)")
    a = ap.parse_args()
    sc = a.scale
    cfgs = [
        ("config1 SE100 1Mbp 1%sub 0.1%indel", dict(genome=1_000_000, pe=False, n=int(4_000_000 * sc), L=100, sub=0.01, indel=0.001, imax=1)),
        ("config3-shape PE150 (10 Mbp genome)", dict(genome=10_000_000, pe=True, n=int(1_500_000 * sc), L=150, sub=0.01, indel=0.001, imax=1)),
        ("config4-shape PE250 high-indel (10 Mbp genome)", dict(genome=10_000_000, pe=True, n=int(600_000 * sc), L=250, sub=0.03, indel=0.002, imax=12)),
        # one mate of 15 % of the pairs cannot be seeded, 3 % are junk: mem_matesw's local alignments dominate pass 2
        ("rescue PE150, 15 % of the pairs need mate rescue (10 Mbp genome)", dict(genome=10_000_000, pe=True, n=int(1_000_000 * sc), L=150, sub=0.01, indel=0.001, imax=1, rescue=0.15, junk=0.03)),
    ]
    if a.only:
        cfgs = [(n, c) for n, c in cfgs if a.only in n]
    if a.genome:
        cfgs = [(n.replace("10 Mbp", f"{a.genome / 1e6:.0f} Mbp"), dict(c, genome=a.genome)) for n, c in cfgs]
    if a.reads:
        cfgs = [(n, dict(c, n=a.reads)) for n, c in cfgs]
    rows = []
    with tempfile.TemporaryDirectory() as d:
        for name, c in cfgs:
            fa = os.path.join(d, "ref.fa")
            g = S.write_genome(fa, c["genome"], seed=1)
            S.bwa_index(fa)
            if c["pe"]:
                reads = [os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")]
                S.write_reads_fast(reads, g, c["n"], c["L"], seed=2, sub=c["sub"], indel=c["indel"], indel_max=c["imax"],
                                   rescue_frac=c.get("rescue", 0.0), junk_frac=c.get("junk", 0.0))
                n_reads = 2 * c["n"]
            else:
                reads = [os.path.join(d, "r.fq")]
                S.write_reads_fast(reads, g, c["n"], c["L"], seed=2, sub=c["sub"], indel=c["indel"], indel_max=c["imax"])
                n_reads = c["n"]
            t_stock, e_stock = timed(S.BWA_STOCK, fa, reads, os.path.join(d, "stock.sam"), a.threads)
            env = dict(os.environ, KSW_B200_SCHED=a.sched) if a.sched == "rounds" else dict(os.environ)
            if a.gpus:
                env["KSW_B200_GPUS"] = str(a.gpus)
            t_b200, e_b200 = timed(S.BWA_B200, fa, reads, os.path.join(d, "b200.sam"), a.threads, extra=["-b", str(a.batch)], env=env)
            ok, why = S.sam_equal(os.path.join(d, "stock.sam"), os.path.join(d, "b200.sam"))
            # the same build with the CIGAR look-ahead off (pass 2 entirely on the host, as before): isolates its effect
            if a.no_cigar_off:
                t_b200_0, e_b200_0, ok0 = 0.0, "", None
            else:
                env0 = dict(env or os.environ, KSW_B200_CIGAR="0")
                t_b200_0, e_b200_0 = timed(S.BWA_B200, fa, reads, os.path.join(d, "b200_0.sam"), a.threads, extra=["-b", str(a.batch)], env=env0)
                ok0, _ = S.sam_equal(os.path.join(d, "stock.sam"), os.path.join(d, "b200_0.sam"))
            rescue = None
            if c["pe"]:
                import re
                mm = re.findall(r"mate-rescue alignments: look-ahead ([\d.]+) thread-s, (\d+) computed ahead, (\d+) hits, (\d+) misses", e_b200)
                rescue = {"lookahead_thread_s": float(mm[-1][0]), "computed_ahead": int(mm[-1][1]), "hits": int(mm[-1][2]), "misses": int(mm[-1][3])} if mm else None
                if a.rescue_off and rescue is not None:
                    env1 = dict(env or os.environ, KSW_B200_RESCUE="0")
                    t1, e1 = timed(S.BWA_B200, fa, reads, os.path.join(d, "b200_1.sam"), a.threads, extra=["-b", str(a.batch)], env=env1)
                    ok1, _ = S.sam_equal(os.path.join(d, "stock.sam"), os.path.join(d, "b200_1.sam"))
                    c1 = chunk_times(e1)
                    rescue.update(off_wall_s=round(t1, 3), off_sam_identical_minus_PG=bool(ok1),
                                  off_steady_reads_per_s=round(sum(r for r, _ in c1[1:]) / max(sum(t for _, t in c1[1:]), 1e-9)) if len(c1) > 1 else None)
            cb0 = chunk_times(e_b200_0)
            sb0 = sum(r for r, _ in cb0[1:]) / max(sum(t for _, t in cb0[1:]), 1e-9) if len(cb0) > 1 else None
            cs, cb = chunk_times(e_stock), chunk_times(e_b200)
            # steady state = chunks after the first (the first B200 chunk pays CUDA context creation, which a real run
            # hides behind loading a GB-sized index)
            ss = sum(r for r, _ in cs[1:]) / max(sum(t for _, t in cs[1:]), 1e-9) if len(cs) > 1 else None
            sb = sum(r for r, _ in cb[1:]) / max(sum(t for _, t in cb[1:]), 1e-9) if len(cb) > 1 else None
            row = {"config": name, "sched": a.sched, "reads": n_reads, "threads": a.threads, "stock_wall_s": round(t_stock, 3), "b200_wall_s": round(t_b200, 3),
                   "stock_reads_per_s": round(n_reads / t_stock), "b200_reads_per_s": round(n_reads / t_b200),
                   "stock_chunks": cs, "b200_chunks": cb,
                   "stock_steady_reads_per_s": round(ss) if ss else None, "b200_steady_reads_per_s": round(sb) if sb else None,
                   "sam_identical_minus_PG": bool(ok),
                   "cigar_lookahead": {"computed_ahead_hits_misses": cigar_stats(e_b200),
                                       "off_wall_s": round(t_b200_0, 3), "off_steady_reads_per_s": round(sb0) if sb0 else None,
                                       "off_sam_identical_minus_PG": ok0},
                   "mate_rescue_lookahead": rescue,
                   "n_gpus": a.gpus or "all visible", "genome_bp": c["genome"]}
            print(json.dumps(row), flush=True)
            rows.append(row)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
