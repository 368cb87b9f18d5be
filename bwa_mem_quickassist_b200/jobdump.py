"""Reader of the extension-job trace the host driver writes (KSW_B200_DUMP, csrc/bwamem_ext.c: the batched counterpart of
the reference's `-v 4` extension trace, bwa-0.7.8/bwamem.c:821-828) and the generator of the *real job mixes* the
benchmark reports: the B200-bound `bwa mem` (integration/_bin/bwa_b200) is run on synthetic reads of the BASELINE
shapes with the trace on, and every job of both extension passes is replayed through the kernels.  No data is
committed; the mixes are regenerated from seeds."""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import struct
import subprocess
import tempfile

import numpy as np

from .ksw import Cfg, JOB_DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BWA_B200 = os.path.join(ROOT, "integration", "_bin", "bwa_b200")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0], dtype=np.uint8)

# the three read shapes of BASELINE.json configs 1 / 3 / 4: (paired, read length, substitution rate,
# indel events per base, longest indel)
MIXES = {
    "se100": dict(pe=False, length=100, sub=0.01, indel=0.001, indel_max=1),
    "pe150": dict(pe=True, length=150, sub=0.01, indel=0.001, indel_max=1),
    "pe250hi": dict(pe=True, length=250, sub=0.03, indel=0.02, indel_max=12),
}


def read_batches(path: str, limit: int = 10 ** 9):
    """-> list of (cfg, jobs[JOB_DT], qpool, tpool), one per extension pass of one worker batch"""
    out = []
    with open(path, "rb") as f:
        while len(out) < limit:
            magic = f.read(4)
            if len(magic) < 4:
                break
            if magic != b"KSWJ":
                raise ValueError(f"{path}: not a job trace")
            cfg = Cfg.from_buffer_copy(f.read(C.sizeof(Cfg)))
            n, nq, nt = struct.unpack("<3Q", f.read(24))
            jobs = np.frombuffer(f.read(JOB_DT.itemsize * n), dtype=JOB_DT).copy()
            q = np.frombuffer(f.read(nq), dtype=np.uint8).copy()
            t = np.frombuffer(f.read(nt), dtype=np.uint8).copy()
            out.append((cfg, jobs, q, t))
    return out


def merge(batches, end_bonus=None):
    """the passes of many worker batches as ONE batch (optionally only those with the given end_bonus, i.e. one side)"""
    sel = [b for b in batches if end_bonus is None or b[0].end_bonus == end_bonus]
    if not sel:
        raise ValueError("no job batches in the trace (was the run made with KSW_B200_DUMP and KSW_B200_REF=0?)")
    jobs, qs, ts, qo, to = [], [], [], 0, 0
    for cfg, j, q, t in sel:
        j = j.copy(); j["q_off"] += qo; j["t_off"] += to
        jobs.append(j); qs.append(q); ts.append(t); qo += len(q); to += len(t)
    return sel[0][0], np.concatenate(jobs), np.concatenate(qs), np.concatenate(ts)


def _fastq(path, prefix, reads):
    n, L = reads.shape
    names = np.char.add(prefix, np.char.zfill(np.arange(n).astype("U9"), 9)).astype("S")
    wn = names.dtype.itemsize
    buf = np.full((n, 1 + wn + 1 + L + 3 + L + 1), ord("\n"), dtype=np.uint8)
    buf[:, 0] = ord("@")
    buf[:, 1:1 + wn] = np.frombuffer(names.tobytes(), dtype=np.uint8).reshape(n, wn)
    o = 1 + wn + 1
    buf[:, o:o + L] = ACGT[reads]
    buf[:, o + L + 1] = ord("+")
    buf[:, o + L + 3:o + L + 3 + L] = ord("I")
    buf.tofile(path)


def _mutate(rng, frag, L, sub, indel, indel_max):
    """substitutions at rate `sub` per base, at most one indel event per read (probability indel * L), length U[1, indel_max]"""
    n, W = frag.shape
    q = frag.copy()
    k = int(rng.binomial(n * W, sub))
    pos = rng.integers(0, n * W, size=k)
    flat = q.reshape(-1)
    flat[pos] = (flat[pos] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    rows = np.flatnonzero(rng.random(n) < min(1.0, indel * L))
    if rows.size:
        col = np.arange(W, dtype=np.int32)[None, :]
        p = rng.integers(1, L - 1, size=rows.size).astype(np.int32)[:, None]
        ln = rng.integers(1, indel_max + 1, size=rows.size).astype(np.int32)[:, None]
        ins = (rng.random(rows.size) < 0.5)[:, None]
        shift = np.where(ins, np.where(col >= p + ln, -ln, 0), np.where(col >= p, ln, 0))
        idx = np.clip(col + shift, 0, W - 1)
        sub_rows = np.take_along_axis(q[rows], idx, axis=1)
        rnd = rng.integers(0, 4, size=sub_rows.shape, dtype=np.uint8)
        q[rows] = np.where(ins & (col >= p) & (col < p + ln), rnd, sub_rows)
    return q[:, :L]


def write_synthetic(dirname: str, genome_len: int, n_reads: int, seed: int, pe: bool, length: int, sub: float, indel: float,
                    indel_max: int):
    """uniform-random genome + reads sampled from both strands (PE: FR pairs, insert ~ N(2.5 L, 0.25 L)) -> (fasta, [fastq ...])"""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, genome_len).astype(np.uint8)
    fa = os.path.join(dirname, "ref.fa")
    with open(fa, "wb") as f:
        f.write(b">chr1\n")
        body = ACGT[g]
        w = 80
        full = (len(body) // w) * w
        lines = np.full((full // w, w + 1), ord("\n"), dtype=np.uint8)
        lines[:, :w] = body[:full].reshape(-1, w)
        f.write(lines.tobytes())
        if full < len(body):
            f.write(body[full:].tobytes() + b"\n")
    pad = 2 * indel_max + 4
    W = length + pad
    col = np.arange(W, dtype=np.int64)[None, :]
    if pe:
        m = n_reads // 2
        isz = np.maximum(length + 10, rng.normal(2.5 * length, 0.25 * length, m).astype(np.int64))
        start = rng.integers(0, genome_len - isz.max() - W - 2, size=m)
        rev = rng.random(m) < 0.5
        f1 = g[start[:, None] + col]
        f2 = COMP[g[(start + isz)[:, None] - 1 - col]]
        a = np.where(rev[:, None], f2, f1)
        b = np.where(rev[:, None], f1, f2)
        paths = [os.path.join(dirname, "r1.fq"), os.path.join(dirname, "r2.fq")]
        _fastq(paths[0], "p", _mutate(rng, a, length, sub, indel, indel_max))
        _fastq(paths[1], "p", _mutate(rng, b, length, sub, indel, indel_max))
    else:
        start = rng.integers(0, genome_len - W - 2, size=n_reads)
        rev = rng.random(n_reads) < 0.5
        fw = g[start[:, None] + col]
        rc = COMP[g[(start + W)[:, None] - 1 - col]]
        paths = [os.path.join(dirname, "r1.fq")]
        _fastq(paths[0], "r", _mutate(rng, np.where(rev[:, None], rc, fw), length, sub, indel, indel_max))
    return fa, paths


def harvest(mix: str, n_reads: int, genome_len: int = 2_000_000, seed: int = 1, threads: int | None = None, keep: str | None = None):
    """Runs the B200-bound `bwa mem` on a synthetic set of shape `mix` with the job trace on and returns the pass batches.
    Needs integration/_bin/bwa_b200 and a GPU (the extension runs on it; the trace is what the kernels were given)."""
    if not (os.path.exists(BWA_B200) and os.access(BWA_B200, os.X_OK)):
        raise FileNotFoundError(f"{BWA_B200} not built (make -C integration, needs the reference sources)")
    d = keep or tempfile.mkdtemp(prefix="ksw_mix_")
    try:
        fa, reads = write_synthetic(d, genome_len, n_reads, seed, **MIXES[mix])
        subprocess.run([BWA_B200, "index", fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        # KSW_B200_REF=0: the host driver materialises the reference windows (with the reference resident on the device a job
        # is a coordinate, and there is nothing to trace for offline replay)
        env = dict(os.environ, KSW_B200_DUMP=os.path.join(d, "jobs"), KSW_B200_REF="0")
        t = threads or min(os.cpu_count() or 4, 16)
        with open(os.devnull, "wb") as nul:
            subprocess.run([BWA_B200, "mem", "-t", str(t), fa, *reads], check=True, stdout=nul, stderr=subprocess.PIPE, env=env)
        out = []
        for f in sorted(glob.glob(os.path.join(d, "jobs.*.bin"))):
            out += read_batches(f)
        return out
    finally:
        if not keep:
            shutil.rmtree(d, ignore_errors=True)
