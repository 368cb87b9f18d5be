"""Synthetic extension workloads of the shapes BASELINE.json names (no datasets are reachable).

config 2: n jobs, qlen = tlen = 101, target uniform random, query = target with 1 % substitutions and
0.1 % single-base indels, h0 ~ U[19,100], w = 100, default scoring.  Fully
vectorised so that 10 M jobs are generated in seconds; deterministic for a given (seed, n, chunk).
"""
from __future__ import annotations

import numpy as np

from .ksw import JOB_DT


def config2_jobs(n: int, seed: int = 12345, qlen: int = 101, tlen: int = 101, sub: float = 0.01,
                 indel: float = 0.001, w: int = 100, h0_lo: int = 19, h0_hi: int = 100, chunk: int = 1 << 20):
    """Returns (jobs[JOB_DT], qpool[uint8], tpool[uint8])."""
    qpool = np.empty(n * qlen, dtype=np.uint8)
    tpool = np.empty(n * tlen, dtype=np.uint8)
    jobs = np.zeros(n, dtype=JOB_DT)
    jobs["q_off"] = np.arange(n, dtype=np.uint64) * np.uint64(qlen)
    jobs["t_off"] = np.arange(n, dtype=np.uint64) * np.uint64(tlen)
    jobs["qlen"], jobs["tlen"], jobs["w"] = qlen, tlen, w
    L = max(qlen, tlen) + 2
    col = np.arange(L, dtype=np.int32)[None, :]
    for ci, b in enumerate(range(0, n, chunk)):
        e = min(n, b + chunk)
        m = e - b
        rng = np.random.default_rng([seed, ci])
        base = rng.integers(0, 4, size=(m, L), dtype=np.uint8)
        tpool[b * tlen:e * tlen] = base[:, :tlen].reshape(-1)
        q = base.copy()
        # substitutions: Binomial number of positions, each moved to one of the 3 other bases
        k = int(rng.binomial(m * L, sub))
        pos = rng.integers(0, m * L, size=k)
        flat = q.reshape(-1)
        flat[pos] = (flat[pos] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
        # at most one single-base indel event per query (rate indel*qlen per job)
        rows = np.flatnonzero(rng.random(m) < indel * qlen)
        if rows.size:
            p = rng.integers(1, qlen - 1, size=rows.size).astype(np.int32)[:, None]
            ins = (rng.random(rows.size) < 0.5)[:, None]
            shift = np.where(ins, np.where(col > p, -1, 0), np.where(col >= p, 1, 0))
            idx = np.clip(col + shift, 0, L - 1)
            sub_rows = np.take_along_axis(q[rows], idx, axis=1)
            newbase = rng.integers(0, 4, size=rows.size, dtype=np.uint8)
            ins_rows = np.flatnonzero(ins[:, 0])
            sub_rows[ins_rows, p[ins_rows, 0]] = newbase[ins_rows]
            q[rows] = sub_rows
        qpool[b * qlen:e * qlen] = q[:, :qlen].reshape(-1)
        jobs["h0"][b:e] = rng.integers(h0_lo, h0_hi + 1, size=m)
    return jobs, qpool, tpool
