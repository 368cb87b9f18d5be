#!/usr/bin/env python
"""Replays extension-job batches harvested from a real `bwa mem` run (KSW_B200_DUMP, see bwamem_ext.c) through the
kernels: job-shape statistics, kernel-only GCUPS / ext/s on the real job mix, bit-exact check against the oracle.
Usage: python scripts/bench_jobs.py dump.bin [max_batches]"""
import ctypes as C, os, struct, sys, time
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")]
import numpy as np
import kswtest as K
import bwa_mem_quickassist_b200 as B


from bwa_mem_quickassist_b200 import jobdump


def read_batches(path, limit):
    return [K.Batch(*b) for b in jobdump.read_batches(path, limit)]


def merge(batches):
    return K.Batch(*jobdump.merge([(b.cfg, b.jobs, b.qpool, b.tpool) for b in batches]))


def main():
    bs = read_batches(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9)
    big = merge(bs)
    n = big.n
    ql, tl = big.jobs["qlen"], big.jobs["tlen"]
    print(f"{len(bs)} pass batches, {n} jobs; qlen mean {ql.mean():.1f} max {ql.max()}, tlen mean {tl.mean():.1f} max {tl.max()}, h0 mean {big.jobs['h0'].mean():.1f}")
    t0 = time.perf_counter(); want, cells = K.run_oracle(big, threads=os.cpu_count(), want_cells=True); t_cpu = time.perf_counter() - t0
    print(f"oracle: {cells.sum() / 1e9:.3f} Gcells visited ({cells.mean():.0f}/job, {cells.sum() / (ql.astype(np.int64) * tl).sum():.2f} of qlen*tlen) "
          f"in {t_cpu:.2f}s on {os.cpu_count()} threads -> {cells.sum() / t_cpu / 1e9:.2f} GCUPS")
    ctx = B.KswB200(0)
    rb = ctx.upload(big.cfg, big.jobs, big.qpool, big.tpool)
    print(rb.info())
    ms = ctx.run_timed(rb, 6)[1:]
    got = ctx.download(rb)
    print(f"kernel: {ms.mean():.3f} ms -> {cells.sum() / ms.mean() / 1e6:.1f} GCUPS visited, {n / ms.mean() / 1e3:.1f} M ext/s; mismatch vs oracle: {K.first_mismatch(want, got.view(K.RES_DT))}")
    out = np.zeros(n, dtype=B.RES_DT)
    ctx.extend_batch(big.cfg, big.jobs, big.qpool, big.tpool, out=out)
    t0 = time.perf_counter(); ctx.extend_batch(big.cfg, big.jobs, big.qpool, big.tpool, out=out); dt = time.perf_counter() - t0
    print(f"e2e: {dt * 1e3:.2f} ms -> {n / dt / 1e6:.1f} M ext/s")


if __name__ == "__main__":
    main()
