"""Shared test support: ctypes bindings for the oracle / compiled reference / product C-ABI,
and seeded generators for extension jobs (SURVEY.md §8c job classes).

Test infrastructure only.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libksw_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libksw_ref.so")

JOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"),
                   ("h0", "<i4"), ("w", "<i4")])
RES_DT = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                   ("gscore", "<i4"), ("max_off", "<i4")])
assert JOB_DT.itemsize == 32 and RES_DT.itemsize == 24


class Cfg(C.Structure):
    """Layout shared by oracle_cfg_t, ref_cfg_t and ksw_b200_cfg_t."""
    _fields_ = [("mat", C.c_int8 * 25), ("m", C.c_int32), ("o_del", C.c_int32), ("e_del", C.c_int32),
                ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32)]


def fill_scmat(a: int, b: int) -> np.ndarray:
    """Scoring matrix as the reference builds it (bwa-0.7.8/bwa.c:77-86)."""
    mat = np.full((5, 5), -1, dtype=np.int8)
    for i in range(4):
        for j in range(4):
            mat[i, j] = a if i == j else -b
    return mat.reshape(25)


def make_cfg(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, mat=None) -> Cfg:
    cfg = Cfg()
    m = fill_scmat(a, b) if mat is None else np.asarray(mat, dtype=np.int8).reshape(25)
    for i in range(25):
        cfg.mat[i] = int(m[i])
    cfg.m = 5
    cfg.o_del, cfg.e_del, cfg.o_ins, cfg.e_ins = o_del, e_del, o_ins, e_ins
    cfg.zdrop, cfg.end_bonus = zdrop, end_bonus
    return cfg


def cfg_mat(cfg: Cfg) -> np.ndarray:
    return np.array([cfg.mat[i] for i in range(25)], dtype=np.int8)


def build_oracle() -> None:
    """Compile oracle/ (and oracle/_ref when /root/reference is mounted)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


_libs: dict = {}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def oracle_lib():
    if "oracle" not in _libs:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.ksw_oracle_extend_batch.restype = C.c_int
        lib.ksw_oracle_extend_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int]
        lib.ksw_oracle_clamp_w.restype = C.c_int
        lib.ksw_oracle_clamp_w.argtypes = [C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 6
        _libs["oracle"] = lib
    return _libs["oracle"]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    if "ref" not in _libs:
        lib = C.CDLL(REF_SO)
        lib.ksw_ref_extend_batch.restype = C.c_int
        lib.ksw_ref_extend_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int]
        _libs["ref"] = lib
    return _libs["ref"]


@dataclass
class Batch:
    cfg: Cfg
    jobs: np.ndarray      # JOB_DT
    qpool: np.ndarray     # uint8 codes 0..4
    tpool: np.ndarray     # uint8 codes 0..4

    @property
    def n(self) -> int:
        return int(self.jobs.shape[0])

    def take(self, idx) -> "Batch":
        return Batch(self.cfg, np.ascontiguousarray(self.jobs[idx]), self.qpool, self.tpool)


def run_oracle(b: Batch, threads: int = 8, want_cells: bool = False):
    res = np.zeros(b.n, dtype=RES_DT)
    cells = np.zeros(b.n, dtype=np.int64) if want_cells else None
    oracle_lib().ksw_oracle_extend_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool),
                                         _ptr(res), _ptr(cells) if want_cells else None, threads)
    return (res, cells) if want_cells else res


def run_ref(b: Batch, threads: int = 8):
    res = np.zeros(b.n, dtype=RES_DT)
    ref_lib().ksw_ref_extend_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool),
                                   _ptr(res), threads)
    return res


def first_mismatch(a: np.ndarray, b: np.ndarray):
    """Index and field-wise description of the first differing job, or None."""
    bad = np.zeros(a.shape[0], dtype=bool)
    for f in RES_DT.names:
        bad |= a[f] != b[f]
    if not bad.any():
        return None
    k = int(np.flatnonzero(bad)[0])
    return k, int(bad.sum()), {f: (int(a[f][k]), int(b[f][k])) for f in RES_DT.names}


# ------------------------------------------------------------------ generators

def _pools_from_lists(qs, ts, h0, w, cfg) -> Batch:
    n = len(qs)
    qlen = np.array([len(x) for x in qs], dtype=np.int64)
    tlen = np.array([len(x) for x in ts], dtype=np.int64)
    jobs = np.zeros(n, dtype=JOB_DT)
    jobs["q_off"] = np.concatenate([[0], np.cumsum(qlen)[:-1]]) if n else 0
    jobs["t_off"] = np.concatenate([[0], np.cumsum(tlen)[:-1]]) if n else 0
    jobs["qlen"], jobs["tlen"] = qlen, tlen
    jobs["h0"], jobs["w"] = h0, w
    qpool = np.concatenate(qs).astype(np.uint8) if n else np.zeros(0, np.uint8)
    tpool = np.concatenate(ts).astype(np.uint8) if n else np.zeros(0, np.uint8)
    return Batch(cfg, jobs, np.ascontiguousarray(qpool), np.ascontiguousarray(tpool))


def mutate(rng: np.random.Generator, seq: np.ndarray, sub: float, indel: float, max_indel: int = 1) -> np.ndarray:
    """Copy of seq with substitutions and short indels (per-base rates)."""
    out = []
    i, n = 0, len(seq)
    r = rng.random(n * 2 + 8)
    k = 0
    while i < n:
        x = r[k]; k += 1
        if x < indel / 2:                      # deletion from the copy
            i += int(rng.integers(1, max_indel + 1))
            continue
        if x < indel:                          # insertion into the copy
            out.extend(rng.integers(0, 4, int(rng.integers(1, max_indel + 1))).tolist())
        c = int(seq[i])
        if r[k] < sub:
            c = (c + int(rng.integers(1, 4))) & 3
        k += 1
        out.append(c)
        i += 1
    return np.array(out, dtype=np.uint8)


def gen_config2(n: int, seed: int = 12345, qlen: int = 101, tlen: int = 101, sub=0.01, indel=0.001,
                cfg: Cfg | None = None, w: int = 100, h0_lo: int = 19, h0_hi: int = 100) -> Batch:
    """BASELINE.json config 2 shape (the generator bench.py uses): target uniform random, query = target
    with 1 % substitutions and 0.1 % single-base indels, h0 ~ U[19,100], w=100, default scoring."""
    from bwa_mem_quickassist_b200.synth import config2_jobs
    jobs, q, t = config2_jobs(n, seed=seed, qlen=qlen, tlen=tlen, sub=sub, indel=indel, w=w, h0_lo=h0_lo, h0_hi=h0_hi)
    return Batch(cfg or make_cfg(), jobs.astype(JOB_DT), q, t)


def gen_fuzz(n: int, seed: int, cfg: Cfg | None = None, max_q: int = 250, n_frac: float = 0.02,
             related: float = 0.8, w_choices=(1, 2, 5, 10, 30, 50, 100, 150, 200), h0_max: int = 250,
             t_has_n: bool = True) -> Batch:
    """Random jobs: qlen in [1,max_q], tlen in [0, 2*qlen+50], mixture of related and unrelated
    pairs, queries (and optionally targets) with N."""
    rng = np.random.default_rng(seed)
    cfg = cfg or make_cfg()
    qs, ts = [], []
    for _ in range(n):
        ql = int(rng.integers(1, max_q + 1))
        mode = rng.random()
        if mode < related:
            tl = int(rng.integers(max(1, ql // 2), 2 * ql + 50))
            t = rng.integers(0, 4, tl).astype(np.uint8)
            sub = float(rng.choice([0.0, 0.01, 0.03, 0.1, 0.25]))
            ind = float(rng.choice([0.0, 0.001, 0.02, 0.08]))
            q = mutate(rng, t, sub, ind, max_indel=int(rng.choice([1, 4, 12, 40])))
            if len(q) >= ql:
                q = q[:ql]
            else:
                q = np.concatenate([q, rng.integers(0, 4, ql - len(q)).astype(np.uint8)])
        else:
            tl = int(rng.integers(0, 2 * ql + 50))
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = rng.integers(0, 4, ql).astype(np.uint8)
            if rng.random() < 0.3 and tl > 0:      # low-complexity: many ties and 0+match restarts
                t = np.full(tl, t[0], dtype=np.uint8)
                q = np.where(rng.random(ql) < 0.9, t[0], q).astype(np.uint8)
        if rng.random() < 0.3:
            q = np.where(rng.random(ql) < n_frac, 4, q).astype(np.uint8)
        if t_has_n and rng.random() < 0.1 and len(t):
            t = np.where(rng.random(len(t)) < n_frac, 4, t).astype(np.uint8)
        qs.append(q); ts.append(t)
    h0 = rng.integers(0, h0_max + 1, n)
    h0[rng.random(n) < 0.05] = 0
    w = rng.choice(np.array(w_choices), n)
    return _pools_from_lists(qs, ts, h0, w, cfg)


def gen_adversarial(seed: int = 7, cfg: Cfg | None = None) -> Batch:
    """Hand-built classes from SURVEY.md §7.3(2): empty rows, beg==qlen, tie rules, end growth,
    tlen<qlen, qlen==1, active w clamp, z-drop on both branches, long indels, all-N."""
    rng = np.random.default_rng(seed)
    cfg = cfg or make_cfg()
    qs, ts, h0s, ws = [], [], [], []

    def add(q, t, h0, w):
        qs.append(np.asarray(q, dtype=np.uint8)); ts.append(np.asarray(t, dtype=np.uint8))
        h0s.append(h0); ws.append(w)

    for ql in (1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 63, 64, 65, 127, 128, 129, 250, 251):
        t = rng.integers(0, 4, ql + 40)
        for h0 in (0, 1, 6, 7, 8, 19, 100, 250):
            for w in (1, 3, 100):
                add(t[:ql], t, h0, w)                       # perfect match, target longer
                add(t[:ql], t[:max(1, ql // 2)], h0, w)     # tlen < qlen
                add(t[:ql], t[:0], h0, w)                   # tlen == 0
        add(np.full(ql, 4), t, 30, 100)                     # all-N query
        add(t[:ql], np.full(ql + 5, 4), 30, 100)            # all-N target
        add(np.full(ql, 2), np.full(2 * ql + 10, 2), 5, 100)   # homopolymer: ties everywhere
        add(np.full(ql, 2), np.full(2 * ql + 10, 1), 50, 100)  # all mismatch: decays to m==0 / zdrop
    # long deletions / insertions relative to the target: band and z-drop on both branches
    for gap in (1, 5, 20, 60, 90, 120):
        for ql in (60, 101, 150, 250):
            t = rng.integers(0, 4, ql + gap + 60)
            cut = ql // 2
            qd = np.concatenate([t[:cut], t[cut + gap:cut + gap + (ql - cut)]])          # deletion in query
            add(qd, t, 40, 100); add(qd, t, 40, 30); add(qd, t, 200, 200)
            qi = np.concatenate([t[:cut], rng.integers(0, 4, gap), t[cut:]])[:ql]          # insertion in query
            add(qi, t, 40, 100); add(qi, t, 40, 30); add(qi, t, 200, 200)
    # periodic sequences: equal scores on many diagonals (mj / max_ie ties, end growth)
    for period in (1, 2, 3, 5):
        unit = rng.integers(0, 4, period)
        for ql in (10, 50, 101):
            q = np.tile(unit, ql // period + 1)[:ql]
            t = np.tile(unit, (2 * ql) // period + 1)[:2 * ql]
            for h0 in (0, 3, 19, 60):
                add(q, t, h0, 100); add(q, t, h0, 7)
    return _pools_from_lists(qs, ts, np.array(h0s), np.array(ws), cfg)


# ------------------------------------------------------------------ golden fixture
GOLDEN = os.path.join(ROOT, "tests", "golden", "ksw_extend_golden.npz")


def load_golden():
    """{name: (Batch, reference results)} from the committed fixture (made by tests/golden/make_golden.py
    from the compiled reference)."""
    z = np.load(GOLDEN)
    names = sorted({k.split(".")[0] for k in z.files})
    out = {}
    for nm in names:
        cfg = Cfg.from_buffer_copy(z[f"{nm}.cfg"].tobytes())
        b = Batch(cfg, z[f"{nm}.jobs"].astype(JOB_DT), z[f"{nm}.qpool"], z[f"{nm}.tpool"])
        out[nm] = (b, z[f"{nm}.res"].astype(RES_DT))
    return out


# ------------------------------------------------------------------ CPU emulation of the fast kernel source
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libfast_emu.so")


def emu_lib():
    if "emu" not in _libs:
        subprocess.run(["make", "-C", EMU_DIR, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(EMU_SO)
        lib.ksw_fast_emu_batch.restype = C.c_int
        lib.ksw_fast_emu_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int]
        _libs["emu"] = lib
    return _libs["emu"]


def run_emu(b: Batch, threads: int = 4):
    """Results of the product packer + fast-kernel lane code compiled for the CPU.  Jobs the packer routes
    to the generic kernel come back with score == INT32_MIN.  Returns (res, n_fast)."""
    res = np.zeros(b.n, dtype=RES_DT)
    nf = C.c_int64(0)
    rc = emu_lib().ksw_fast_emu_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res),
                                      C.byref(nf), threads)
    assert rc == 0, rc
    return res, int(nf.value)
