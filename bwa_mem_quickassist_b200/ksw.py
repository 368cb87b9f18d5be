"""ctypes binding of include/ksw_b200.h (the drop-in boundary of the ksw_extend path).

Names and argument meaning mirror the reference's C interface (bwa-0.7.8/ksw.h:107-108 for the
scalar calls; the batched entry replaces the loop over ksw_extend2 in mem_chain2aln,
bwa-0.7.8/bwamem.c:826,854).  No algorithm lives here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

JOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"),
                   ("h0", "<i4"), ("w", "<i4")])            # ksw_b200_job_t
RES_DT = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                   ("gscore", "<i4"), ("max_off", "<i4")])  # ksw_b200_res_t
RJOB_DT = np.dtype([("q_off", "<u8"), ("t_pos", "<i8"), ("qlen", "<i4"), ("tlen", "<i4"), ("h0", "<i4"), ("w", "<i4"),
                    ("q_step", "i1"), ("t_step", "i1"), ("reserved", "i1", (6,))])   # ksw_b200_rjob_t
# banded global alignment with backtrace
GJOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"),
                    ("reserved", "<i4")])                   # ksw_b200_gjob_t
GRES_DT = np.dtype([("score", "<i4"), ("n_cigar", "<i4"), ("cigar_off", "<i8")])   # ksw_b200_gres_t
AJOB_DT = np.dtype([("q_off", "<u8"), ("t_off", "<u8"), ("qlen", "<i4"), ("tlen", "<i4"), ("xtra", "<i4"), ("reserved", "<i4")])   # ksw_b200_ajob_t
ARES_DT = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4"),
                    ("reserved", "<i4")])                                          # ksw_b200_ares_t (kswr_t, ksw.h:30-36)
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000   # ksw.h:6-9


class KswB200Error(RuntimeError):
    pass


class Cfg(C.Structure):                                      # ksw_b200_cfg_t
    _fields_ = [("mat", C.c_int8 * 25), ("m", C.c_int32), ("o_del", C.c_int32), ("e_del", C.c_int32),
                ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32)]


def make_cfg(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, mat=None) -> Cfg:
    """Scoring as `bwa mem` builds it: bwa_fill_scmat (bwa-0.7.8/bwa.c:77-86) + mem_opt_init defaults
    (bwa-0.7.8/bwamem.c:45-75)."""
    cfg = Cfg()
    if mat is None:
        k = 0
        for i in range(4):
            for j in range(4):
                cfg.mat[k] = a if i == j else -b
                k += 1
            cfg.mat[k] = -1
            k += 1
        for j in range(5):
            cfg.mat[k] = -1
            k += 1
    else:
        m = np.asarray(mat, dtype=np.int8).reshape(25)
        for i in range(25):
            cfg.mat[i] = int(m[i])
    cfg.m = 5
    cfg.o_del, cfg.e_del, cfg.o_ins, cfg.e_ins = o_del, e_del, o_ins, e_ins
    cfg.zdrop, cfg.end_bonus = zdrop, end_bonus
    return cfg


def lib_path() -> str:
    # KSW_B200_LIB: development override (A/B builds of the same library); the product is the in-tree libksw_b200.so
    return os.environ.get("KSW_B200_LIB") or os.path.join(_HERE, "libksw_b200.so")


_lib = None


def load_library():
    """Loads libksw_b200.so.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise KswB200Error(f"{p} is missing: run ./build.sh (or __graft_entry__.build()); "
                           "there is no CPU fallback for the extension path")
    lib = C.CDLL(p)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.ksw_b200_device_count.restype = i32
    lib.ksw_b200_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.ksw_b200_ctx_destroy.argtypes = [vp]
    lib.ksw_b200_ctx_destroy.restype = None
    lib.ksw_b200_strerror.argtypes = [vp]
    lib.ksw_b200_strerror.restype = C.c_char_p
    lib.ksw_b200_ctx_set_pack_threads.argtypes = [vp, i32]
    lib.ksw_b200_ctx_set_chunk_jobs.argtypes = [vp, i64]
    lib.ksw_b200_ctx_launch_count.argtypes = [vp]
    lib.ksw_b200_ctx_launch_count.restype = i64
    lib.ksw_b200_ctx_sync.argtypes = [vp]
    lib.ksw_b200_extend_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.ksw_b200_extend_batch_multi.argtypes = [i32, vp, vp, i64, vp, vp, vp, vp]
    lib.ksw_b200_host_alloc.argtypes = [C.c_size_t]
    lib.ksw_b200_host_alloc.restype = vp
    lib.ksw_b200_host_free.argtypes = [vp]
    lib.ksw_b200_host_free.restype = None
    lib.ksw_b200_host_register.argtypes = [vp, C.c_size_t]
    lib.ksw_b200_host_unregister.argtypes = [vp]
    lib.ksw_b200_extend_batch_async.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp, C.c_size_t, vp]
    lib.ksw_b200_wait.argtypes = [vp]
    lib.ksw_b200_queue_create.argtypes = [i32, C.POINTER(vp)]
    lib.ksw_b200_queue_destroy.argtypes = [vp]
    lib.ksw_b200_queue_destroy.restype = None
    lib.ksw_b200_queue_ref_set.argtypes = [vp, vp, i64]
    lib.ksw_b200_queue_extend_ref.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.ksw_b200_queue_global.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(i64)]
    lib.ksw_b200_queue_strerror.argtypes = [vp]
    lib.ksw_b200_queue_strerror.restype = C.c_char_p
    lib.ksw_b200_queue_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.ksw_b200_queue_stats.restype = None
    lib.ksw_b200_ref_set.argtypes = [vp, vp, i64]
    lib.ksw_b200_extend_batch_ref.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.ksw_b200_batch_upload.argtypes = [vp, vp, i64, vp, vp, vp, C.POINTER(vp)]
    lib.ksw_b200_batch_run.argtypes = [vp, vp]
    lib.ksw_b200_batch_run_timed.argtypes = [vp, vp, i32, vp]
    lib.ksw_b200_batch_run_timed2.argtypes = [vp, vp, i32, vp, vp]
    lib.ksw_b200_batch_download.argtypes = [vp, vp, vp]
    lib.ksw_b200_batch_download_cells.argtypes = [vp, vp, vp]
    lib.ksw_b200_ctx_last_transfer.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.ksw_b200_batch_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.ksw_b200_batch_free.argtypes = [vp, vp]
    lib.ksw_b200_batch_free.restype = None
    lib.ksw_b200_dpx_peak.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.ksw_b200_clamp_w.argtypes = [i32, vp] + [i32] * 6
    scalar = [i32, vp, i32, vp, i32, vp]
    lib.ksw_extend.argtypes = scalar + [i32] * 6 + [C.POINTER(i32)] * 5
    lib.ksw_extend2.argtypes = scalar + [i32] * 8 + [C.POINTER(i32)] * 5
    lib.ksw_b200_global_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(i64)]
    lib.ksw_global.argtypes = scalar + [i32] * 3 + [C.POINTER(i32), C.POINTER(C.POINTER(C.c_uint32))]
    lib.ksw_global2.argtypes = scalar + [i32] * 5 + [C.POINTER(i32), C.POINTER(C.POINTER(C.c_uint32))]
    _lib = lib
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """A page-locked host array from ksw_b200_host_alloc, viewed as numpy (`.a`); freed with close() or by the GC."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
        self.nbytes = n * dtype.itemsize
        self.ptr = self.lib.ksw_b200_host_alloc(self.nbytes)
        if not self.ptr:
            raise KswB200Error(f"ksw_b200_host_alloc({self.nbytes}) failed")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.a = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes).view(dtype).reshape(shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.a = None
            self.lib.ksw_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_copy(a: np.ndarray) -> PinnedArray:
    """`a` copied into page-locked memory (what a C caller would have filled in place)."""
    a = np.ascontiguousarray(a)
    p = PinnedArray(a.shape, a.dtype)
    p.a[...] = a
    return p


class ResidentBatch:
    def __init__(self, owner: "KswB200", handle, n: int):
        self.owner, self.handle, self.n = owner, handle, n

    def info(self):
        nf, ng, pb = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.owner.lib.ksw_b200_batch_info(self.handle, C.byref(nf), C.byref(ng), C.byref(pb))
        return {"n_fast": nf.value, "n_generic": ng.value, "packed_bytes": pb.value}

    def free(self):
        if self.handle:
            self.owner.lib.ksw_b200_batch_free(self.owner.ctx, self.handle)
            self.handle = None


class KswB200:
    """One extension context == one (host thread, GPU) pair, as the C ABI defines it."""

    def __init__(self, device: int = 0, pack_threads: int | None = None):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.ksw_b200_ctx_create(device, C.byref(self.ctx))
        if rc != 0:
            raise KswB200Error(f"ksw_b200_ctx_create(device={device}) failed with code {rc}: no usable B200; "
                               "the extension path has no CPU fallback")
        if pack_threads:
            self.lib.ksw_b200_ctx_set_pack_threads(self.ctx, int(pack_threads))

    def set_chunk_jobs(self, n: int):
        self.lib.ksw_b200_ctx_set_chunk_jobs(self.ctx, int(n))

    def set_pack_threads(self, n: int):
        self.lib.ksw_b200_ctx_set_pack_threads(self.ctx, int(n))

    def close(self):
        if self.ctx:
            self.lib.ksw_b200_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise KswB200Error(f"{what} failed ({rc}): {self.lib.ksw_b200_strerror(self.ctx).decode()}")

    @staticmethod
    def _norm(jobs, qpool, tpool):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
        return jobs, qpool, tpool

    def extend_batch(self, cfg: Cfg, jobs, qpool, tpool, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host results out (pack + H2D + kernels + D2H inside the call).
        `out` (RES_DT, len(jobs)) lets a caller reuse its result array, as a C caller would."""
        jobs, qpool, tpool = self._norm(jobs, qpool, tpool)
        if out is not None:
            assert out.dtype == RES_DT and out.shape[0] == jobs.shape[0] and out.flags.c_contiguous
            res = out
        else:
            res = np.zeros(jobs.shape[0], dtype=RES_DT)
        self._check(self.lib.ksw_b200_extend_batch(self.ctx, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool),
                                                   _p(tpool), _p(res)), "ksw_b200_extend_batch")
        return res

    def extend_batch_async(self, cfg: Cfg, jobs: np.ndarray, qpool: np.ndarray, tpool: np.ndarray, res: np.ndarray):
        """ksw_b200_extend_batch_async: all four arrays must live in page-locked memory (PinnedArray.a or registered)
        and stay alive until wait() returns.  Packing runs on the device; returns immediately."""
        assert jobs.dtype == JOB_DT and res.dtype == RES_DT and qpool.dtype == np.uint8 and tpool.dtype == np.uint8
        assert jobs.flags.c_contiguous and res.flags.c_contiguous and qpool.flags.c_contiguous and tpool.flags.c_contiguous
        assert res.shape[0] == jobs.shape[0]
        self._cfg_keepalive = cfg
        self._check(self.lib.ksw_b200_extend_batch_async(self.ctx, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), qpool.nbytes,
                                                         _p(tpool), tpool.nbytes, _p(res)), "ksw_b200_extend_batch_async")

    def wait(self):
        self._check(self.lib.ksw_b200_wait(self.ctx), "ksw_b200_wait")

    def ref_set(self, pac: np.ndarray, l_pac: int):
        """ksw_b200_ref_set: the forward-strand 2-bit .pac (l_pac/4+1 bytes) goes to this context's device."""
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        assert pac.nbytes >= l_pac // 4 + 1
        self._pac_keepalive = pac
        self._check(self.lib.ksw_b200_ref_set(self.ctx, _p(pac), int(l_pac)), "ksw_b200_ref_set")

    def extend_batch_ref(self, cfg: Cfg, rjobs, qpool) -> np.ndarray:
        """ksw_b200_extend_batch_ref: targets are runs of the doubled reference space, sliced on the device."""
        rjobs = np.ascontiguousarray(rjobs, dtype=RJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        res = np.zeros(rjobs.shape[0], dtype=RES_DT)
        self._check(self.lib.ksw_b200_extend_batch_ref(self.ctx, C.byref(cfg), rjobs.shape[0], _p(rjobs), _p(qpool), qpool.nbytes,
                                                       _p(res)), "ksw_b200_extend_batch_ref")
        return res

    def global_batch(self, cfg: Cfg, jobs, qpool, tpool):
        """ksw_b200_global_batch: banded global alignment + backtrace of every job (GJOB_DT).  Returns
        (res[GRES_DT], cigar_pool[uint32]); job k's CIGAR is cigar_pool[res[k].cigar_off : +res[k].n_cigar],
        one operation per word, len << 4 | op (0 = M, 1 = I, 2 = D) as in the reference."""
        jobs = np.ascontiguousarray(jobs, dtype=GJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
        res = np.zeros(jobs.shape[0], dtype=GRES_DT)
        pool = C.POINTER(C.c_uint32)()
        total = C.c_int64(0)
        self._check(self.lib.ksw_b200_global_batch(self.ctx, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), _p(tpool),
                                                   _p(res), C.byref(pool), C.byref(total)), "ksw_b200_global_batch")
        cig = np.ctypeslib.as_array(pool, shape=(total.value,)).copy() if total.value else np.zeros(0, np.uint32)
        return res, cig

    def align_batch(self, cfg: Cfg, jobs, qpool, tpool):
        """ksw_b200_align_batch: the reference's ksw_align2 (local alignment, start positions, second-best score) for every
        job (AJOB_DT; xtra = KSW_X* flags | threshold as mem_matesw passes them).  Returns res[ARES_DT]."""
        jobs = np.ascontiguousarray(jobs, dtype=AJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
        res = np.zeros(jobs.shape[0], dtype=ARES_DT)
        self._check(self.lib.ksw_b200_align_batch(self.ctx, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), _p(tpool), _p(res)),
                    "ksw_b200_align_batch")
        return res

    def upload(self, cfg: Cfg, jobs, qpool, tpool) -> ResidentBatch:
        jobs, qpool, tpool = self._norm(jobs, qpool, tpool)
        h = C.c_void_p()
        self._check(self.lib.ksw_b200_batch_upload(self.ctx, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool),
                                                   _p(tpool), C.byref(h)), "ksw_b200_batch_upload")
        return ResidentBatch(self, h, int(jobs.shape[0]))

    def run(self, batch: ResidentBatch):
        self._check(self.lib.ksw_b200_batch_run(self.ctx, batch.handle), "ksw_b200_batch_run")

    def run_timed(self, batch: ResidentBatch, iters: int) -> np.ndarray:
        ms = np.zeros(iters, dtype=np.float32)
        self._check(self.lib.ksw_b200_batch_run_timed(self.ctx, batch.handle, iters, _p(ms)),
                    "ksw_b200_batch_run_timed")
        return ms

    def run_timed2(self, batch: ResidentBatch, iters: int):
        """(ms per step, ms of the extension kernels alone per step)"""
        ms = np.zeros(iters, dtype=np.float32)
        ext = np.zeros(iters, dtype=np.float32)
        self._check(self.lib.ksw_b200_batch_run_timed2(self.ctx, batch.handle, iters, _p(ms), _p(ext)),
                    "ksw_b200_batch_run_timed2")
        return ms, ext

    def download(self, batch: ResidentBatch) -> np.ndarray:
        res = np.zeros(batch.n, dtype=RES_DT)
        self._check(self.lib.ksw_b200_batch_download(self.ctx, batch.handle, _p(res)), "ksw_b200_batch_download")
        return res

    def download_cells(self, batch: ResidentBatch) -> np.ndarray:
        cells = np.zeros(batch.n, dtype=np.uint32)
        self._check(self.lib.ksw_b200_batch_download_cells(self.ctx, batch.handle, _p(cells)),
                    "ksw_b200_batch_download_cells")
        return cells

    def last_transfer(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.ksw_b200_ctx_last_transfer(self.ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def sync(self):
        self._check(self.lib.ksw_b200_ctx_sync(self.ctx), "ksw_b200_ctx_sync")

    def launch_count(self) -> int:
        return int(self.lib.ksw_b200_ctx_launch_count(self.ctx))

    def dpx_peak(self, which: int = 0):
        ops, ms = C.c_double(0), C.c_float(0)
        self._check(self.lib.ksw_b200_dpx_peak(self.ctx, which, C.byref(ops), C.byref(ms)), "ksw_b200_dpx_peak")
        return ops.value, ms.value


class KswQueue:
    """ksw_b200_queue_t: one submission queue per GPU shared by any number of host threads (the calls block and release
    the GIL, so Python threads submit concurrently)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.q = C.c_void_p()
        rc = self.lib.ksw_b200_queue_create(device, C.byref(self.q))
        if rc != 0:
            raise KswB200Error(f"ksw_b200_queue_create(device={device}) failed with code {rc}")

    def _check(self, rc, what):
        if rc != 0:
            raise KswB200Error(f"{what} failed ({rc}): {self.lib.ksw_b200_queue_strerror(self.q).decode()}")

    def ref_set(self, pac: np.ndarray, l_pac: int):
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        self._pac_keepalive = pac
        self._check(self.lib.ksw_b200_queue_ref_set(self.q, _p(pac), int(l_pac)), "ksw_b200_queue_ref_set")

    def extend_ref(self, cfg: Cfg, rjobs, qpool) -> np.ndarray:
        rjobs = np.ascontiguousarray(rjobs, dtype=RJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        res = np.zeros(rjobs.shape[0], dtype=RES_DT)
        self._check(self.lib.ksw_b200_queue_extend_ref(self.q, C.byref(cfg), rjobs.shape[0], _p(rjobs), _p(qpool), qpool.nbytes, _p(res)),
                    "ksw_b200_queue_extend_ref")
        return res

    def global_batch(self, cfg: Cfg, jobs, qpool, tpool):
        jobs = np.ascontiguousarray(jobs, dtype=GJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
        res = np.zeros(jobs.shape[0], dtype=GRES_DT)
        pool = C.POINTER(C.c_uint32)()
        total = C.c_int64(0)
        self._check(self.lib.ksw_b200_queue_global(self.q, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), _p(tpool), _p(res),
                                                   C.byref(pool), C.byref(total)), "ksw_b200_queue_global")
        cig = np.ctypeslib.as_array(pool, shape=(total.value,)).copy() if total.value else np.zeros(0, np.uint32)
        if pool:
            libc = C.CDLL(None)
            libc.free.argtypes = [C.c_void_p]
            libc.free(pool)
        return res, cig

    def align_batch(self, cfg: Cfg, jobs, qpool, tpool):
        jobs = np.ascontiguousarray(jobs, dtype=AJOB_DT)
        qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
        tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
        res = np.zeros(jobs.shape[0], dtype=ARES_DT)
        self._check(self.lib.ksw_b200_queue_align(self.q, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), _p(tpool), _p(res)),
                    "ksw_b200_queue_align")
        return res

    def stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.ksw_b200_queue_stats(self.q, C.byref(a), C.byref(b))
        return {"batches": a.value, "submissions": b.value}

    def close(self):
        if self.q:
            self.lib.ksw_b200_queue_destroy(self.q)
            self.q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def extend_batch_multi(ctxs, cfg: Cfg, jobs, qpool, tpool) -> np.ndarray:
    """ksw_b200_extend_batch_multi over a list of KswB200 contexts (one per GPU)."""
    lib = load_library()
    jobs, qpool, tpool = KswB200._norm(jobs, qpool, tpool)
    res = np.zeros(jobs.shape[0], dtype=RES_DT)
    arr = (C.c_void_p * len(ctxs))(*[c.ctx for c in ctxs])
    rc = lib.ksw_b200_extend_batch_multi(len(ctxs), arr, C.byref(cfg), jobs.shape[0], _p(jobs), _p(qpool), _p(tpool), _p(res))
    if rc != 0:
        raise KswB200Error(f"ksw_b200_extend_batch_multi failed ({rc}): " + "; ".join(
            lib.ksw_b200_strerror(c.ctx).decode() for c in ctxs))
    return res


def ksw_extend2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
    """Scalar drop-in with the reference's argument order (ksw.h:108).  Returns
    (score, qle, tle, gtle, gscore, max_off)."""
    lib = load_library()
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    mt = np.ascontiguousarray(mat, dtype=np.int8)
    out = [C.c_int(0) for _ in range(5)]
    sc = lib.ksw_extend2(qlen, _p(q), tlen, _p(t), m, _p(mt), o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                         *[C.byref(x) for x in out])
    return (sc,) + tuple(x.value for x in out)


def ksw_extend(qlen, query, tlen, target, m, mat, gapo, gape, w, end_bonus, zdrop, h0):
    """Scalar drop-in (ksw.h:107)."""
    return ksw_extend2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, end_bonus, zdrop, h0)


def ksw_global2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w):
    """Scalar drop-in with the reference's argument order (ksw.h:84).  Returns (score, cigar[uint32]); the array the
    library malloc'd for the caller is freed here with libc free(), exactly what a C caller has to do."""
    lib = load_library()
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    mt = np.ascontiguousarray(mat, dtype=np.int8)
    n = C.c_int(0)
    cig = C.POINTER(C.c_uint32)()
    sc = lib.ksw_global2(qlen, _p(q), tlen, _p(t), m, _p(mt), o_del, e_del, o_ins, e_ins, w, C.byref(n), C.byref(cig))
    out = np.array([cig[k] for k in range(n.value)], dtype=np.uint32)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(cig, C.c_void_p))
    return sc, out


class _Kswr(C.Structure):                                # kswr_t, ksw.h:30-36
    _fields_ = [(n, C.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")]


def ksw_align2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra):
    """Scalar drop-in with the reference's argument order (ksw.h:63; qry = NULL).  Returns the kswr_t fields as a dict."""
    lib = load_library()
    lib.ksw_align2.restype = _Kswr
    lib.ksw_align2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    mt = np.ascontiguousarray(mat, dtype=np.int8)
    r = lib.ksw_align2(qlen, _p(q), tlen, _p(t), m, _p(mt), o_del, e_del, o_ins, e_ins, xtra, None)
    return {n: getattr(r, n) for n, _ in _Kswr._fields_}
