"""Whole-program check support: synthetic genome + reads, run stock / fork / B200-bound `bwa mem`, compare SAM.
Methodology of the reference's own A/B harness (pipeline.sh vs pipeline_ref.sh: same command, DUT vs REF SAM)
and of NEWS:27-28 (whole-SAM identity).  Test infrastructure only."""
from __future__ import annotations

import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
BWA_STOCK = os.path.join(REFDIR, "bwa_stock")
BWA_FORK = os.path.join(REFDIR, "bwa_fork")
BWA_B200 = os.path.join(ROOT, "integration", "_bin", "bwa_b200")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def have_binaries(*paths) -> bool:
    return all(os.path.exists(p) and os.access(p, os.X_OK) for p in paths)


def write_genome(path: str, length: int, seed: int, n_contigs: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, length).astype(np.uint8)
    with open(path, "w") as f:
        per = length // n_contigs
        for c in range(n_contigs):
            seg = g[c * per:(c + 1) * per if c < n_contigs - 1 else length]
            f.write(f">chr{c + 1}\n")
            s = ACGT[seg].tobytes().decode()
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + "\n")
    return g


def _mutate(rng, seq, sub, indel_events, indel_max):
    """substitutions per base; `indel_events` per-base rate of an insertion or deletion of length U[1,indel_max]"""
    out = []
    i, n = 0, len(seq)
    r = rng.random(2 * n + 4)
    k = 0
    while i < n:
        x = r[k]; k += 1
        if x < indel_events / 2:
            i += int(rng.integers(1, indel_max + 1))
            continue
        if x < indel_events:
            out.extend(rng.integers(0, 4, int(rng.integers(1, indel_max + 1))).tolist())
        c = int(seq[i])
        if r[k] < sub:
            c = (c + int(rng.integers(1, 4))) & 3
        k += 1
        out.append(c); i += 1
    return np.array(out, dtype=np.uint8)


def _fq(f, name, codes):
    s = ACGT[codes].tobytes().decode()
    f.write(f"@{name}\n{s}\n+\n{'I' * len(s)}\n")


def write_reads_se(path, genome, n, length, seed, sub=0.01, indel=0.001, indel_max=1, n_rate=0.0):
    rng = np.random.default_rng(seed)
    G = len(genome)
    with open(path, "w") as f:
        for r in range(n):
            p = int(rng.integers(0, G - length - 20))
            frag = genome[p:p + length + 20]
            if rng.random() < 0.5:
                frag = COMP[frag[::-1]]
            rd = _mutate(rng, frag, sub, indel, indel_max)[:length]
            s = ACGT[rd].tobytes().decode()
            if n_rate and rng.random() < 0.2:
                s = "".join("N" if rng.random() < n_rate else ch for ch in s)
            f.write(f"@r{r}\n{s}\n+\n{'I' * len(s)}\n")


def write_reads_pe(path1, path2, genome, n_pairs, length, seed, sub=0.01, indel=0.001, indel_max=1, ins_mean=None, ins_sd=None):
    rng = np.random.default_rng(seed)
    G = len(genome)
    ins_mean = ins_mean or 2.5 * length
    ins_sd = ins_sd or 0.25 * length
    with open(path1, "w") as f1, open(path2, "w") as f2:
        for r in range(n_pairs):
            isz = max(length + 10, int(rng.normal(ins_mean, ins_sd)))
            p = int(rng.integers(0, G - isz - 40))
            frag = genome[p:p + isz + 40]
            if rng.random() < 0.5:
                frag = COMP[frag[::-1]]
            a = _mutate(rng, frag[:length + 30], sub, indel, indel_max)[:length]
            b = _mutate(rng, COMP[frag[:isz][::-1]][:length + 30], sub, indel, indel_max)[:length]
            _fq(f1, f"p{r}", a)
            _fq(f2, f"p{r}", b)


def run(cmd, stdout_path=None, env=None, timeout=1800):
    with open(stdout_path, "wb") if stdout_path else open(os.devnull, "wb") as out:
        p = subprocess.run(cmd, stdout=out, stderr=subprocess.PIPE, env=env, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed ({p.returncode}): {p.stderr.decode()[-2000:]}")
    return p.stderr.decode()


def bwa_index(fa):
    run([BWA_STOCK, "index", fa])


def bwa_mem(binary, fa, reads, out_sam, threads=4, extra=(), env=None):
    return run([binary, "mem", "-t", str(threads), *extra, fa, *reads], stdout_path=out_sam, env=env)


def sam_body(path):
    """SAM lines without @PG (the only line allowed to differ: main.c:66-68 puts the command line there)."""
    with open(path, "rb") as f:
        return [ln for ln in f if not ln.startswith(b"@PG")]


def sam_equal(a, b):
    """Streams both files (multi-GB SAMs of the BASELINE-size runs must not be held in memory)."""
    def body(path):
        with open(path, "rb") as f:
            for ln in f:
                if not ln.startswith(b"@PG"):
                    yield ln
    ia, ib = body(a), body(b)
    i = 0
    while True:
        x, y = next(ia, None), next(ib, None)
        if x is None and y is None:
            return True, None
        if x != y:
            if x is None or y is None:
                return False, (i, b"<length differs>", b"one file ends at line %d" % i)
            return False, (i, x[:300], y[:300])
        i += 1


# ------------------------------------------------------------------ vectorised read simulator (millions of reads in seconds)
def _mutate_matrix(rng, frag, L, sub, indel, indel_max):
    """frag: (n, L+pad) uint8 fragments (read direction).  Returns (n, L): substitutions at rate `sub` per base and at
    most one indel event per read (probability indel*L) of length U[1,indel_max]."""
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
        # deletion from the read: read[c] = frag[c + ln] for c >= p; insertion: read[c] = frag[c - ln] for c >= p + ln,
        # random bases in [p, p + ln)
        shift = np.where(ins, np.where(col >= p + ln, -ln, 0), np.where(col >= p, ln, 0))
        idx = np.clip(col + shift, 0, W - 1)
        sub_rows = np.take_along_axis(q[rows], idx, axis=1)
        rnd = rng.integers(0, 4, size=sub_rows.shape, dtype=np.uint8)
        inside = ins & (col >= p) & (col < p + ln)
        sub_rows = np.where(inside, rnd, sub_rows)
        q[rows] = sub_rows
    return q[:, :L]


def _write_fastq_fixed(path, prefix, reads):
    """reads: (n, L) uint8 codes; names are fixed width so the whole file is one numpy fill."""
    n, L = reads.shape
    names = np.char.add(prefix, np.char.zfill(np.arange(n).astype("U9"), 9)).astype("S")
    wn = names.dtype.itemsize
    rec = 1 + wn + 1 + L + 3 + L + 1
    buf = np.full((n, rec), ord("\n"), dtype=np.uint8)
    buf[:, 0] = ord("@")
    buf[:, 1:1 + wn] = np.frombuffer(names.tobytes(), dtype=np.uint8).reshape(n, wn)
    o = 1 + wn + 1
    buf[:, o:o + L] = ACGT[reads]
    buf[:, o + L + 1] = ord("+")
    buf[:, o + L + 3:o + L + 3 + L] = ord("I")
    buf.tofile(path)


def write_reads_fast(paths, genome, n, length, seed, sub=0.01, indel=0.001, indel_max=1, ins_mean=None, ins_sd=None, chunk=500_000,
                     rescue_frac=0.0, rescue_sub=0.18, junk_frac=0.0):
    """SE if len(paths) == 1 else PE (FR orientation, insert ~ N(ins_mean, ins_sd)).  PE only: in a fraction `rescue_frac`
    of the pairs one mate (either, at random) gets `rescue_sub` more substitutions per base — too many for a 19 bp seed to survive in most of them, so
    the aligner finds it only by mate rescue (mem_matesw) — and in `junk_frac` of the pairs one mate is random sequence."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    pe = len(paths) == 2
    pad = 2 * indel_max + 4
    ins_mean = ins_mean or 2.5 * length
    ins_sd = ins_sd or 0.25 * length
    for p in paths:
        open(p, "wb").close()
    done = 0
    while done < n:
        m = min(chunk, n - done)
        W = length + pad
        col = np.arange(W, dtype=np.int64)[None, :]
        if pe:
            isz = np.maximum(length + 10, rng.normal(ins_mean, ins_sd, m).astype(np.int64))
            start = rng.integers(0, G - isz.max() - W - 2, size=m)
            rev = rng.random(m) < 0.5
            # fragment on the forward strand is genome[start, start+isz); read 1 from its 5' end, read 2 = revcomp of its 3' end
            f1 = genome[start[:, None] + col]
            f2 = COMP[genome[(start + isz)[:, None] - 1 - col]]
            # flipped pairs: the fragment is taken from the reverse strand, i.e. the roles swap
            a = np.where(rev[:, None], f2, f1)
            b = np.where(rev[:, None], f1, f2)
            r1 = _mutate_matrix(rng, a, length, sub, indel, indel_max)
            r2 = _mutate_matrix(rng, b, length, sub, indel, indel_max)
            if rescue_frac > 0 or junk_frac > 0:
                u = rng.random(m)
                which = rng.random(m) < 0.5
                hard = u < rescue_frac
                junk = (u >= rescue_frac) & (u < rescue_frac + junk_frac)
                for r, sel in ((r1, which), (r2, ~which)):
                    rows = np.flatnonzero(hard & sel)
                    if rows.size:
                        mut = rng.random((rows.size, length)) < rescue_sub
                        blk = r[rows]
                        blk[mut] = (blk[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
                        r[rows] = blk
                    rows = np.flatnonzero(junk & sel)
                    if rows.size:
                        r[rows] = rng.integers(0, 4, size=(rows.size, length), dtype=np.uint8)
            for path, r in ((paths[0], r1), (paths[1], r2)):
                tmp = path + ".part"
                _write_fastq_fixed(tmp, f"p{done // chunk:03d}_", r)
                with open(path, "ab") as out, open(tmp, "rb") as src:
                    out.write(src.read())
                os.remove(tmp)
        else:
            start = rng.integers(0, G - W - 2, size=m)
            rev = rng.random(m) < 0.5
            fw = genome[start[:, None] + col]
            rc = COMP[genome[(start + W)[:, None] - 1 - col]]
            frag = np.where(rev[:, None], rc, fw)
            r = _mutate_matrix(rng, frag, length, sub, indel, indel_max)
            tmp = paths[0] + ".part"
            _write_fastq_fixed(tmp, f"r{done // chunk:03d}_", r)
            with open(paths[0], "ab") as out, open(tmp, "rb") as src:
                out.write(src.read())
            os.remove(tmp)
        done += m
