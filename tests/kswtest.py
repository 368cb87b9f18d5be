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
# banded global alignment with backtrace (ksw_global2): job / result records shared by oracle, reference shim and product
GJOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"), ("reserved", "<i4")])
GRES_DT = np.dtype([("score", "<i4"), ("n_cigar", "<i4"), ("cigar_off", "<i8")])
assert GJOB_DT.itemsize == 32 and GRES_DT.itemsize == 16
AJOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"), ("xtra", "<i4"), ("reserved", "<i4")])
ARES_DT = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4"), ("reserved", "<i4")])
assert AJOB_DT.itemsize == 32 and ARES_DT.itemsize == 32
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000      # ksw.h:6-9


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
        lib.ksw_oracle_global_batch.restype = C.c_int
        lib.ksw_oracle_global_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_int]
        lib.ksw_oracle_align_batch.restype = C.c_int
        lib.ksw_oracle_align_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
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
        lib.ksw_ref_global_batch.restype = C.c_int
        lib.ksw_ref_global_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int]
        if hasattr(lib, "ksw_ref_align_batch"):
            lib.ksw_ref_align_batch.restype = C.c_int
            lib.ksw_ref_align_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
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


GLOBAL_GOLDEN = os.path.join(ROOT, "tests", "golden", "ksw_global_golden.npz")


def load_global_golden():
    """{name: (GBatch, (res, cigar pool))} from the committed fixture made from the compiled reference's ksw_global2."""
    z = np.load(GLOBAL_GOLDEN)
    out = {}
    for nm in sorted({k.split(".")[0] for k in z.files}):
        cfg = Cfg.from_buffer_copy(z[f"{nm}.cfg"].tobytes())
        b = GBatch(cfg, z[f"{nm}.jobs"].astype(GJOB_DT), z[f"{nm}.qpool"], z[f"{nm}.tpool"])
        out[nm] = (b, (z[f"{nm}.res"].astype(GRES_DT), z[f"{nm}.cigar"]))
    return out


# ------------------------------------------------------------------ CPU emulation of the fast kernel source
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libfast_emu.so")
EMU_SO = os.environ.get("KSW_EMU_SO", EMU_SO)


def emu_lib():
    if "emu" not in _libs:
        subprocess.run(["make", "-C", EMU_DIR, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(EMU_SO)
        lib.ksw_fast_emu_batch.restype = C.c_int
        lib.ksw_fast_emu_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_void_p]
        lib.ksw_pair_emu_batch.restype = C.c_int
        lib.ksw_pair_emu_batch.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _libs["emu"] = lib
    return _libs["emu"]


def run_global_emu(b):
    """(res, pool, n_fast) from the CPU build of the fast global-alignment kernels' source (ksw_gfast_core.h); jobs outside
    that kernel's eligibility keep score = INT32_MIN."""
    lib = emu_lib()
    lib.ksw_gfast_emu_batch.restype = C.c_int
    lib.ksw_gfast_emu_batch.argtypes = [C.POINTER(Cfg), C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(C.c_int64)]
    res = np.zeros(b.n, dtype=GRES_DT)
    cap = b.jobs["qlen"].astype(np.int64) + b.jobs["tlen"] + 2
    res["cigar_off"] = np.concatenate([[0], np.cumsum(cap)[:-1]]) if b.n else 0
    pool = np.zeros(int(cap.sum()) + 1, dtype=np.uint32)
    nf = C.c_int64(0)
    rc = lib.ksw_gfast_emu_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res), _ptr(pool), C.byref(nf))
    assert rc == 0
    return res, pool, int(nf.value)


def run_pair_emu(b: Batch, lanes: int = 3, order: int = 0):
    """Results of the product packer + the PAIR kernel's lane code (two jobs per lane) compiled for the CPU.
    Jobs outside class 0 come back with score == INT32_MIN.  Returns (res, cells, n_pair, lane_rows)."""
    res = np.zeros(b.n, dtype=RES_DT)
    cells = np.zeros(b.n, dtype=np.uint32)
    npair = C.c_int64(0)
    steps = C.c_int64(0)
    rc = emu_lib().ksw_pair_emu_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res),
                                      _ptr(cells), lanes, order, C.byref(npair), C.byref(steps))
    assert rc == 0, rc
    return res, cells, int(npair.value), int(steps.value)


def run_emu(b: Batch, threads: int = 4):
    """Results of the product packer + fast-kernel lane code compiled for the CPU.  Jobs the packer routes
    to the generic kernel come back with score == INT32_MIN.  Returns (res, n_fast)."""
    res = np.zeros(b.n, dtype=RES_DT)
    nf = C.c_int64(0)
    nk = C.c_int64(0)
    rc = emu_lib().ksw_fast_emu_batch(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res),
                                      C.byref(nf), threads, C.byref(nk))
    assert rc == 0, rc
    run_emu.last_keyed = int(nk.value)
    return res, int(nf.value)


def run_packer(b: Batch, force_words: bool, threads: int = 3):
    """What the product's host packer writes for a batch: (job records as raw uint32[n, 8], 2-bit pool words, N side
    pool words, jobs per kernel class), through the 64-byte SIMD path or the word-at-a-time path."""
    cap = int((b.jobs["qlen"].astype(np.int64) + b.jobs["tlen"] + 94).sum() // 4 + 64)
    dj = np.zeros((max(b.n, 1), 8), dtype=np.uint32)
    pool = np.full(cap, 0xdeadbeef, dtype=np.uint32)
    ncap = int((b.jobs["qlen"].astype(np.int64) + b.jobs["tlen"] + 64).sum() // 32 + 2 * b.n + 8)
    nm = np.zeros(ncap, dtype=np.uint32)
    pw, nw = C.c_int64(0), C.c_int64(0)
    cn = (C.c_int64 * 6)()
    rc = emu_lib().ksw_pack_emu(C.byref(b.cfg), C.c_int64(b.n), _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), int(force_words), threads,
                                _ptr(dj), _ptr(pool), C.c_int64(cap), C.byref(pw), _ptr(nm), C.c_int64(ncap), C.byref(nw), cn)
    assert rc == 0, rc
    return dj[:b.n], pool[:pw.value].copy(), nm[:nw.value].copy(), [int(x) for x in cn]


# ------------------------------------------------------------------ chains -> regions (mem_chain2aln level)
SEED_DT = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4")])                      # mem_seed_t
REG_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                   ("sub", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"), ("seedcov", "<i4"),
                   ("secondary", "<i4"), ("hash", "<u8")])                                   # mem_alnreg_t
assert SEED_DT.itemsize == 16 and REG_DT.itemsize == 64


class ExtOpt(C.Structure):
    """b200_ext_opt_t / o_opt_t / flat_opt_t: the mem_opt_t fields the extension path reads."""
    _fields_ = [("a", C.c_int), ("b", C.c_int), ("o_del", C.c_int), ("e_del", C.c_int), ("o_ins", C.c_int),
                ("e_ins", C.c_int), ("pen_clip5", C.c_int), ("pen_clip3", C.c_int), ("w", C.c_int), ("zdrop", C.c_int),
                ("mat", C.c_int8 * 25)]


def make_ext_opt(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=100, zdrop=100) -> ExtOpt:
    o = ExtOpt()
    o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins = a, b, o_del, e_del, o_ins, e_ins
    o.pen_clip5, o.pen_clip3, o.w, o.zdrop = pen_clip5, pen_clip3, w, zdrop
    m = fill_scmat(a, b)
    for i in range(25):
        o.mat[i] = int(m[i])
    return o


@dataclass
class ChainSet:
    opt: ExtOpt
    l_pac: int
    pac: np.ndarray          # uint8, 2-bit packed forward strand (bntseq.c:191)
    read_off: np.ndarray     # int64
    read_len: np.ndarray     # int32
    qpool: np.ndarray        # uint8 codes
    chain_read: np.ndarray   # int32, non-decreasing
    chain_seed0: np.ndarray  # int64
    chain_nseeds: np.ndarray  # int32
    seeds: np.ndarray        # SEED_DT

    @property
    def n_reads(self):
        return int(self.read_len.shape[0])

    @property
    def n_chains(self):
        return int(self.chain_read.shape[0])

    def flat_args(self):
        return [self.l_pac, _ptr(self.pac), self.n_reads, _ptr(self.read_off), _ptr(self.read_len), _ptr(self.qpool),
                self.n_chains, _ptr(self.chain_read), _ptr(self.chain_seed0), _ptr(self.chain_nseeds), _ptr(self.seeds)]


def pack_pac(genome: np.ndarray) -> np.ndarray:
    """2-bit packing as the reference stores .pac: base l in byte l>>2 at bits ((~l)&3)<<1 (bntseq.c:191)."""
    n = len(genome)
    g = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
    g[:n] = genome
    g = g.reshape(-1, 4)
    pac = (g[:, 0] << 6) | (g[:, 1] << 4) | (g[:, 2] << 2) | g[:, 3]
    return np.ascontiguousarray(np.concatenate([pac, np.zeros(1, np.uint8)]).astype(np.uint8))


def gen_chains(n_reads: int, seed: int, l_pac: int = 30000, read_lens=(100, 150, 250), sub=0.02, indel=0.003,
               max_indel=8, opt: ExtOpt | None = None, k: int = 19, n_frac=0.0) -> ChainSet:
    """Synthetic reads against a random genome with a few repeats; seeds = maximal exact matches (>= k) found with
    a k-mer index over the doubled (forward + reverse-complement) coordinate space of the reference; chains =
    greedy groups of same-strand seeds on nearby diagonals.  Any such chain is a legal mem_chain2aln input."""
    rng = np.random.default_rng(seed)
    opt = opt or make_ext_opt()
    genome = rng.integers(0, 4, l_pac).astype(np.uint8)
    for _ in range(6):                                   # repeats -> several chains per read, contained seeds
        ln = int(rng.integers(60, 400)); a = int(rng.integers(0, l_pac - ln)); b = int(rng.integers(0, l_pac - ln))
        seg = genome[a:a + ln].copy()
        flip = rng.random(ln) < 0.01
        seg[flip] = (seg[flip] + 1) & 3
        genome[b:b + ln] = seg
    dbl = np.concatenate([genome, (3 - genome[::-1])]).astype(np.uint8)
    L2 = 2 * l_pac
    # k-mer index over the doubled space (windows that straddle l_pac are skipped)
    codes = np.zeros(L2 - k + 1, dtype=np.uint64)
    for i in range(k):
        codes = (codes << np.uint64(2)) | dbl[i:L2 - k + 1 + i].astype(np.uint64)
    index: dict = {}
    for pos, c in enumerate(codes.tolist()):
        if pos < l_pac < pos + k:
            continue
        index.setdefault(c, []).append(pos)
    reads, chain_read, chain_seed0, chain_nseeds, seeds = [], [], [], [], []
    for r in range(n_reads):
        L = int(rng.choice(read_lens))
        strand = int(rng.integers(0, 2))
        p0 = int(rng.integers(0, l_pac - L - 1)) + strand * l_pac
        read = mutate(rng, dbl[p0:p0 + L + 40], sub, indel, max_indel)[:L]
        if len(read) < L:
            read = np.concatenate([read, rng.integers(0, 4, L - len(read)).astype(np.uint8)])
        if n_frac and rng.random() < 0.3:
            read = np.where(rng.random(L) < n_frac, 4, read).astype(np.uint8)
        found = set()
        rc = np.zeros(max(L - k + 1, 0), dtype=np.uint64)
        ok = np.ones(max(L - k + 1, 0), dtype=bool)
        for i in range(k):
            part = read[i:L - k + 1 + i]
            ok &= part < 4
            rc = (rc << np.uint64(2)) | (part & 3).astype(np.uint64)
        for q, (c, good) in enumerate(zip(rc.tolist(), ok.tolist())):
            if not good:
                continue
            for pos in index.get(c, ()):
                qb, rb = q, pos
                lim_lo = 0 if pos < l_pac else l_pac
                lim_hi = l_pac if pos < l_pac else L2
                while qb > 0 and rb > lim_lo and read[qb - 1] == dbl[rb - 1]:
                    qb -= 1; rb -= 1
                qe, re_ = q + k, pos + k
                while qe < L and re_ < lim_hi and read[qe] == dbl[re_]:
                    qe += 1; re_ += 1
                found.add((rb, qb, qe - qb))
        sl = sorted(found)
        # greedy chaining: same strand, diagonal within 120, reference distance within 1000
        chains: list = []
        for (rb, qb, ln) in sl:
            placed = False
            for ch in chains:
                rb0, qb0, _ = ch[-1]
                if (rb0 < l_pac) == (rb < l_pac) and abs((rb - qb) - (rb0 - qb0)) < 120 and abs(rb - rb0) < 1000:
                    ch.append((rb, qb, ln)); placed = True
                    break
            if not placed:
                chains.append([(rb, qb, ln)])
        chains.sort(key=lambda ch: -sum(s[2] for s in ch))
        reads.append(read)
        for ch in chains[:6]:
            chain_read.append(r); chain_seed0.append(len(seeds)); chain_nseeds.append(len(ch))
            seeds.extend(ch)
    rl = np.array([len(x) for x in reads], dtype=np.int32)
    ro = np.concatenate([[0], np.cumsum(rl)[:-1]]).astype(np.int64)
    sd = np.zeros(len(seeds), dtype=SEED_DT)
    if seeds:
        arr = np.array(seeds, dtype=np.int64)
        sd["rbeg"], sd["qbeg"], sd["len"] = arr[:, 0], arr[:, 1], arr[:, 2]
    return ChainSet(opt, l_pac, pack_pac(genome), ro, rl, np.ascontiguousarray(np.concatenate(reads).astype(np.uint8)),
                    np.array(chain_read, dtype=np.int32), np.array(chain_seed0, dtype=np.int64),
                    np.array(chain_nseeds, dtype=np.int32), sd)


def _run_chain_flat(fn, cs: ChainSet, extra_head=(), want_calls=False):
    cap = int(cs.seeds.shape[0]) + 8
    out = np.zeros(cap, dtype=REG_DT)
    out_read = np.zeros(cap, dtype=np.int32)
    n_out = C.c_int64(0)
    n_calls = C.c_int64(0)
    args = list(extra_head) + [C.byref(cs.opt)] + cs.flat_args() + [C.c_int64(cap), _ptr(out), _ptr(out_read), C.byref(n_out)]
    if want_calls:
        args.append(C.byref(n_calls))
    rc = fn(*args)
    assert rc == 0, rc
    n = int(n_out.value)
    return (out[:n].copy(), out_read[:n].copy(), int(n_calls.value)) if want_calls else (out[:n].copy(), out_read[:n].copy())


_FLAT_TAIL = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
              C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]


def run_chain_oracle(cs: ChainSet):
    fn = oracle_lib().oracle_chain2aln_flat
    fn.restype = C.c_int
    fn.argtypes = _FLAT_TAIL + [C.c_void_p]
    return _run_chain_flat(fn, cs, want_calls=True)


BWA_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libbwa_ref.so")


def have_bwa_ref() -> bool:
    return os.path.exists(BWA_REF_SO)


def run_chain_ref(cs: ChainSet):
    if "bwaref" not in _libs:
        _libs["bwaref"] = C.CDLL(BWA_REF_SO)
    fn = _libs["bwaref"].ref_chain2aln_flat
    fn.restype = C.c_int
    fn.argtypes = _FLAT_TAIL
    return _run_chain_flat(fn, cs)


def run_chain_gpu(ctx, cs: ChainSet):
    """The product's batched driver (b200_chain2aln_flat) on a KswB200 context."""
    fn = ctx.lib.b200_chain2aln_flat
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] + _FLAT_TAIL
    return _run_chain_flat(fn, cs, extra_head=[ctx.ctx])


def regs_equal(a, b):
    (ra, ia), (rb, ib) = a, b
    if ra.shape != rb.shape or not (ia == ib).all():
        return False
    return all((ra[f] == rb[f]).all() for f in REG_DT.names)


CHAIN_GOLDEN = os.path.join(ROOT, "tests", "golden", "chain2aln_golden.npz")


def load_chain_golden():
    """{name: (ChainSet, (regions, region_read))} made by tests/golden/make_golden.py from the reference's own
    mem_chain2aln."""
    z = np.load(CHAIN_GOLDEN)
    out = {}
    for nm in sorted({k.split(".")[0] for k in z.files}):
        opt = ExtOpt.from_buffer_copy(z[f"{nm}.opt"].tobytes())
        cs = ChainSet(opt, int(z[f"{nm}.l_pac"][0]), z[f"{nm}.pac"], z[f"{nm}.read_off"], z[f"{nm}.read_len"], z[f"{nm}.qpool"],
                      z[f"{nm}.chain_read"], z[f"{nm}.chain_seed0"], z[f"{nm}.chain_nseeds"], z[f"{nm}.seeds"].astype(SEED_DT))
        out[nm] = (cs, (z[f"{nm}.regs"].astype(REG_DT), z[f"{nm}.reg_read"]))
    return out


def gen_boundaries(seed: int = 99, cfg: Cfg | None = None) -> Batch:
    """Jobs that sit exactly on the packer's class boundaries: the keyed bound (h0 + qlen*a + o_del + e_del == 511 and
    512), the query-length classes (124/125, 128/129, 256/257, 512/513) and the int16 budget (h0 + qlen*a == 20000 and
    20001), with perfect matches so that the largest scores are really reached, plus near-perfect variants."""
    rng = np.random.default_rng(seed)
    cfg = cfg or make_cfg()
    a = int(cfg.mat[0]); B = cfg.o_del + cfg.e_del
    qs, ts, h0s, ws = [], [], [], []
    for ql in (1, 3, 4, 5, 123, 124, 125, 127, 128, 129, 255, 256, 257, 511, 512, 513):
        t = rng.integers(0, 4, ql + 30).astype(np.uint8)
        q = t[:ql].copy()
        q2 = q.copy()
        if ql > 8:
            q2[ql // 2] = (q2[ql // 2] + 1) & 3
        for h0 in sorted({0, 1, max(0, 511 - B - ql * a - 1), max(0, 511 - B - ql * a), max(0, 511 - B - ql * a + 1),
                          max(0, 20000 - ql * a - 1), max(0, 20000 - ql * a), max(0, 20000 - ql * a + 1), 32000}):
            for qq in (q, q2):
                qs.append(qq); ts.append(t); h0s.append(h0); ws.append(100)
                qs.append(qq); ts.append(t[:ql]); h0s.append(h0); ws.append(1000)
    return _pools_from_lists(qs, ts, np.array(h0s), np.array(ws), cfg)


# ------------------------------------------------------------------ host-side driver against an oracle-backed stub (no GPU)
EXT_EMU_SO = os.path.join(EMU_DIR, "libext_emu.so")


def ext_emu_lib():
    if "extemu" not in _libs:
        subprocess.run(["make", "-C", EMU_DIR, "libext_emu.so", "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
        _libs["extemu"] = C.CDLL(EXT_EMU_SO)
    return _libs["extemu"]


def run_chain_driver(lib, cs: ChainSet, ctx=None, rounds: bool = False):
    """b200_chain2aln_flat / _flat_rounds of `lib` (the product library with a GPU context, or the stub-backed test
    build with ctx=None).  Returns (regions, region_read, n_jobs) — n_jobs only in rounds mode."""
    if rounds:
        fn = lib.b200_chain2aln_flat_rounds
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] + _FLAT_TAIL + [C.c_void_p]
        regs, rr, nj = _run_chain_flat(fn, cs, extra_head=[ctx], want_calls=True)
        return regs, rr, nj
    fn = lib.b200_chain2aln_flat
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] + _FLAT_TAIL
    regs, rr = _run_chain_flat(fn, cs, extra_head=[ctx])
    return regs, rr, None


# ------------------------------------------------------------------ banded global alignment with backtrace (ksw_global2)
@dataclass
class GBatch:
    cfg: Cfg
    jobs: np.ndarray      # GJOB_DT
    qpool: np.ndarray
    tpool: np.ndarray

    @property
    def n(self) -> int:
        return int(self.jobs.shape[0])


def _run_global(fn, b: GBatch, threads: int):
    res = np.zeros(b.n, dtype=GRES_DT)
    cap = b.jobs["qlen"].astype(np.int64) + b.jobs["tlen"] + 2
    res["cigar_off"] = np.concatenate([[0], np.cumsum(cap)[:-1]]) if b.n else 0
    pool = np.zeros(int(cap.sum()) + 1, dtype=np.uint32)
    rc = fn(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res), _ptr(pool), threads)
    assert rc == 0
    return res, pool


def run_global_oracle(b: GBatch, threads: int = 8):
    """(res[GRES_DT], cigar pool) from the CPU restatement oracle/ksw_global_oracle.c."""
    return _run_global(oracle_lib().ksw_oracle_global_batch, b, threads)


def run_global_ref(b: GBatch, threads: int = 8):
    """The same from the reference's own ksw_global2 (oracle/_ref/libksw_ref.so)."""
    return _run_global(ref_lib().ksw_ref_global_batch, b, threads)


def cigars(res: np.ndarray, pool: np.ndarray):
    """Per-job CIGARs as a list of tuples of uint32 operations."""
    return [tuple(int(x) for x in pool[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])]) for r in res]


def global_mismatch(a, b):
    """First job whose (score, CIGAR) differs between two (res, pool) results, or None."""
    (ra, pa), (rb, pb) = a, b
    bad = np.flatnonzero((ra["score"] != rb["score"]) | (ra["n_cigar"] != rb["n_cigar"]))
    if bad.size:
        k = int(bad[0])
        return k, (int(ra["score"][k]), int(ra["n_cigar"][k])), (int(rb["score"][k]), int(rb["n_cigar"][k]))
    ca, cb = cigars(ra, pa), cigars(rb, pb)
    for k, (x, y) in enumerate(zip(ca, cb)):
        if x != y:
            return k, x, y
    return None


def gen_global(n: int, seed: int, cfg: Cfg | None = None, max_q: int = 250, n_frac: float = 0.01,
               w_extra=(0, 1, 3, 10, 50), indel=(0.0, 0.005, 0.03), sub=(0.0, 0.01, 0.05, 0.2)) -> GBatch:
    """Jobs of the shape bwa_gen_cigar2 produces (bwa.c:118-132): the target is the reference window of an alignment,
    the query the read segment; the band always holds the end cell: w >= |tlen - qlen| (the reference reads
    direction cells it never wrote otherwise).  Includes unrelated pairs, N, qlen/tlen of 1, and w = 0."""
    rng = np.random.default_rng(seed)
    cfg = cfg or make_cfg()
    qs, ts, ws = [], [], []
    for _ in range(n):
        ql = int(rng.integers(1, max_q + 1))
        mode = rng.random()
        if mode < 0.85:
            q = rng.integers(0, 4, ql).astype(np.uint8)
            t = mutate(rng, q, float(rng.choice(sub)), float(rng.choice(indel)), max_indel=int(rng.choice([1, 3, 10])))
            if len(t) == 0:
                t = q[:1].copy()
        else:
            q = rng.integers(0, 4, ql).astype(np.uint8)
            t = rng.integers(0, 4, int(rng.integers(1, ql + 30))).astype(np.uint8)
        if rng.random() < 0.2:
            q = np.where(rng.random(len(q)) < n_frac, 4, q).astype(np.uint8)
        if rng.random() < 0.1:
            t = np.where(rng.random(len(t)) < n_frac, 4, t).astype(np.uint8)
        qs.append(q); ts.append(t)
        ws.append(abs(len(t) - len(q)) + int(rng.choice(np.array(w_extra))))
    jobs = np.zeros(n, dtype=GJOB_DT)
    qlen = np.array([len(x) for x in qs], dtype=np.int64)
    tlen = np.array([len(x) for x in ts], dtype=np.int64)
    jobs["q_off"] = np.concatenate([[0], np.cumsum(qlen)[:-1]]) if n else 0
    jobs["t_off"] = np.concatenate([[0], np.cumsum(tlen)[:-1]]) if n else 0
    jobs["qlen"], jobs["tlen"], jobs["w"] = qlen, tlen, ws
    return GBatch(cfg, jobs, np.ascontiguousarray(np.concatenate(qs)), np.ascontiguousarray(np.concatenate(ts)))


# ------------------------------------------------------------------ local alignment (ksw_align2, mate rescue)
@dataclass
class ABatch:
    cfg: Cfg
    jobs: np.ndarray      # AJOB_DT
    qpool: np.ndarray
    tpool: np.ndarray

    @property
    def n(self) -> int:
        return int(self.jobs.shape[0])


def _run_align(fn, b: ABatch, threads: int):
    res = np.zeros(b.n, dtype=ARES_DT)
    rc = fn(C.byref(b.cfg), b.n, _ptr(b.jobs), _ptr(b.qpool), _ptr(b.tpool), _ptr(res), threads)
    assert rc == 0
    return res


def run_align_oracle(b: ABatch, threads: int = 8):
    """ARES_DT records from the CPU restatement oracle/ksw_align_oracle.c."""
    return _run_align(oracle_lib().ksw_oracle_align_batch, b, threads)


def run_align_oracle_closed_form(b: ABatch, threads: int = 8):
    """The restatement with the lazy-F loop replaced by its closed form (what the GPU kernel computes when o_ins >= 1)."""
    lib = oracle_lib()
    lib.ksw_oracle_align_batch_closed_form.restype = C.c_int
    lib.ksw_oracle_align_batch_closed_form.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return _run_align(lib.ksw_oracle_align_batch_closed_form, b, threads)


def run_align_ref(b: ABatch, threads: int = 8):
    """The same from the reference's own ksw_align2 (oracle/_ref/libksw_ref.so)."""
    return _run_align(ref_lib().ksw_ref_align_batch, b, threads)


def align_mismatch(a: np.ndarray, b: np.ndarray):
    for f in ("score", "te", "qe", "score2", "te2", "tb", "qb"):
        bad = np.flatnonzero(a[f] != b[f])
        if bad.size:
            k = int(bad[0])
            return k, f, tuple(int(a[x][k]) for x in ARES_DT.names[:7]), tuple(int(b[x][k]) for x in ARES_DT.names[:7])
    return None


def gen_align(n: int, seed: int, cfg: Cfg | None = None, max_q: int = 250, max_t: int = 900, flags=None, min_seed: int = 19,
              n_frac: float = 0.01, copies=(0, 1, 1, 1, 2, 3)) -> ABatch:
    """Jobs of the shape mem_matesw produces (bwamem_pair.c:128-150): the query is a read (or its reverse complement), the
    target a reference window that holds 0, 1 or several mutated copies of it (the repeats feed the second-best list);
    xtra as mem_matesw sets it — KSW_XSUBO | KSW_XSTART | (qlen * a < 250 ? KSW_XBYTE : 0) | min_seed_len * a — unless
    `flags` gives a list of flag words to draw from (then the low 16 bits are drawn too)."""
    rng = np.random.default_rng(seed)
    cfg = cfg or make_cfg()
    a = int(cfg_mat(cfg)[0])
    qs, ts = [], []
    jobs = np.zeros(n, dtype=AJOB_DT)
    qo = to = 0
    for k in range(n):
        ql = int(rng.integers(1, max_q + 1)) if rng.random() < 0.2 else int(rng.integers(max(1, max_q // 3), max_q + 1))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        tl = int(rng.integers(0, max_t + 1)) if rng.random() < 0.1 else int(rng.integers(max_t // 4, max_t + 1))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        for _ in range(int(rng.choice(copies))):
            c = q.copy()
            sub = rng.random(ql) < float(rng.choice([0.0, 0.02, 0.08, 0.2]))
            c[sub] = (c[sub] + rng.integers(1, 4, int(sub.sum()))) & 3
            if rng.random() < 0.4 and ql > 8:
                p = int(rng.integers(2, ql - 2)); d = int(rng.integers(1, 6))
                c = np.concatenate([c[:p], rng.integers(0, 4, d).astype(np.uint8), c[p:]]) if rng.random() < 0.5 else np.concatenate([c[:p], c[p + d:]])
            if rng.random() < 0.3:
                c = c[int(rng.integers(0, max(1, len(c) // 2))):]                 # a partial copy
            if len(c) and tl > len(c):
                p = int(rng.integers(0, tl - len(c)))
                t[p:p + len(c)] = c
        q[rng.random(ql) < n_frac] = 4
        if tl: t[rng.random(tl) < n_frac] = 4
        if flags is None:
            xtra = KSW_XSUBO | KSW_XSTART | (KSW_XBYTE if ql * a < 250 else 0) | (min_seed * a)
        else:
            xtra = int(rng.choice(flags)) | int(rng.choice([0, 1, 10, 19, 30, 60, 200]))
            if (xtra & KSW_XBYTE) and ql * max(a, 1) >= 250: xtra &= ~KSW_XBYTE        # byte overflow is outside the domain
        jobs[k] = (qo, to, ql, tl, xtra, 0)
        qs.append(q); ts.append(t); qo += ql; to += tl
    return ABatch(cfg, jobs, np.concatenate(qs + [np.zeros(16, np.uint8)]), np.concatenate(ts + [np.zeros(16, np.uint8)]))


ALIGN_GOLDEN = os.path.join(ROOT, "tests", "golden", "ksw_align_golden.npz")


def load_align_golden():
    """{name: (ABatch, res)} from the committed fixture made from the compiled reference's ksw_align2."""
    z = np.load(ALIGN_GOLDEN)
    out = {}
    for nm in sorted({k.split(".")[0] for k in z.files}):
        cfg = Cfg.from_buffer_copy(z[f"{nm}.cfg"].tobytes())
        out[nm] = (ABatch(cfg, z[f"{nm}.jobs"].astype(AJOB_DT), z[f"{nm}.qpool"], z[f"{nm}.tpool"]), z[f"{nm}.res"].astype(ARES_DT))
    return out
