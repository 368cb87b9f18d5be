#!/usr/bin/env python
"""bench.py — the ksw_extend hot path on B200, per the driver's contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--jobs J] [--impl reference]

Workload (BASELINE.json configs[1]): J = 10 M synthetic 101 bp query/target pairs per GPU, w = 100,
default scoring, h0 ~ U[19,100] (bwa_mem_quickassist_b200/synth.py).  One *step* = one pass of the
batched extension over the whole batch.

  value   GCUPS over *visited* DP cells (sum over executed rows of end-beg, ksw.c:418-421; counted by
          the kernels themselves and cross-checked against the oracle on the CPU sample), inputs
          already packed and resident in HBM; a step = binning (key kernel + radix sort) + the extension
          kernels, timed with CUDA events on the launching stream.
  e2e     the same metric through the C-ABI call a host program makes with HOST buffers
          (ksw_b200_extend_batch_async + ksw_b200_wait on page-locked job / sequence / result arrays):
          raw H2D, 2-bit packing on the device (some chunks on idle host threads), binning, kernels,
          D2H into the caller's result array, every step.  `e2e_pageable` is the same through
          ksw_b200_extend_batch on ordinary (pageable) arrays: host packing into pinned staging.
  roofline      cell-update rate against the DPX issue peak measured live by the library's probe
                kernel (peak_CUPS = lane-ops/s x 2 cells / 8 issue slots); the path is
                integer-issue bound, so the HBM figure is reported beside it as a sanity line.
  cpu_baseline  the reference's own ksw_extend2 (oracle/_ref, compiled unmodified from the reference
                sources) on all host cores over a bounded sample of the same jobs; it is also the
                bit-exact check of the GPU results for that sample.

With N > 1 (torchrun, one rank per GPU) every rank runs the same per-GPU workload (weak scaling; the
path shards by jobs with no exchange step, so there is no data-path collective); timing is the max
over ranks.  `--impl reference` times only the CPU reference (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

METRIC = "ksw_extend GCUPS (visited cells)"
UNIT = "GCUPS"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--jobs", type=int, default=10_000_000, help="jobs per GPU (config 2: 10 M)")
    ap.add_argument("--e2e-jobs", type=int, default=0, help="jobs per e2e step (default: same as --jobs)")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="jobs in the CPU baseline sample")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--real-mix-reads", type=int, default=200_000,
                    help="reads per harvested real job mix (0 = skip the real_mix leg)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(n, seed):
    from bwa_mem_quickassist_b200.synth import config2_jobs
    return config2_jobs(n, seed=seed)


def cpu_reference(cfg, jobs, qpool, tpool, threads):
    """Times the reference's own ksw_extend2 (oracle/_ref) — or the oracle port if the compiled
    reference did not travel — over the given jobs on `threads` host threads."""
    import kswtest as K
    b = K.Batch(cfg, jobs, qpool, tpool)
    kind = "reference" if K.have_ref() else "port"
    t0 = time.perf_counter()
    res = K.run_ref(b, threads=threads) if kind == "reference" else K.run_oracle(b, threads=threads)
    dt = time.perf_counter() - t0
    return kind, res, dt


def visited_cells_oracle(cfg, jobs, qpool, tpool, threads):
    import kswtest as K
    _, cells = K.run_oracle(K.Batch(cfg, jobs, qpool, tpool), threads=threads, want_cells=True)
    return cells


def real_mixes(ctx, peak_gcups, n_reads):
    """Job mixes harvested from the B200-bound `bwa mem` on synthetic reads of the BASELINE shapes (every job of both
    extension passes, bwa_mem_quickassist_b200/jobdump.py), replayed through the resident kernel path: GCUPS over
    visited cells, fraction of the DPX peak, bit-exact flag against the oracle for every job."""
    import kswtest as K
    from bwa_mem_quickassist_b200 import jobdump
    out = {}
    for mix in ("se100", "pe150", "pe250hi"):
        try:
            t0 = time.perf_counter()
            batches = jobdump.harvest(mix, n_reads if mix != "pe250hi" else max(n_reads // 2, 2), genome_len=2_000_000, seed=7)
            cfg, jobs, qpool, tpool = jobdump.merge(batches)
            t_h = time.perf_counter() - t0
            rb = ctx.upload(cfg, jobs, qpool, tpool)
            info = rb.info()
            ms = ctx.run_timed(rb, 5)[1:]
            cells = ctx.download_cells(rb).astype(np.int64)
            got = ctx.download(rb)
            rb.free()
            kcfg = K.Cfg.from_buffer_copy(bytes(cfg))            # same layout, the checker's own ctypes class
            want, ocells = K.run_oracle(K.Batch(kcfg, jobs, qpool, tpool), threads=os.cpu_count() or 1, want_cells=True)
            ok = all((want[f] == got[f]).all() for f in want.dtype.names) and bool((ocells == cells).all())
            g = float(cells.sum()) / float(ms.mean()) / 1e6
            out[mix] = {"jobs": int(len(jobs)), "reads": int(n_reads if mix != "pe250hi" else max(n_reads // 2, 2)),
                        "mean_qlen": float(jobs["qlen"].mean()), "mean_tlen": float(jobs["tlen"].mean()),
                        "visited_cells_per_job": float(cells.mean()), "ms": float(ms.mean()), "gcups": g,
                        "ext_per_s": len(jobs) / float(ms.mean()) * 1e3, "frac_of_dpx_peak": g / peak_gcups,
                        "fast_jobs": info["n_fast"], "generic_jobs": info["n_generic"], "bit_exact": bool(ok),
                        "harvest_s": t_h}
        except Exception as e:                                       # e.g. integration/_bin/bwa_b200 did not travel
            out[mix] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    return out


def run_reference_arm(a):
    """`--impl reference`: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import kswtest as K
    K.build_oracle() if not os.path.exists(K.ORACLE_SO) else None
    threads = os.cpu_count() or 1
    n = min(a.cpu_sample, a.jobs)
    jobs, qpool, tpool = make_workload(n, a.seed)
    cfg = K.make_cfg()
    cells = int(visited_cells_oracle(cfg, jobs, qpool, tpool, threads).sum())
    times = []
    kind = "port"
    for s in range(a.warmup + a.steps):
        kind, _, dt = cpu_reference(cfg, jobs, qpool, tpool, threads)
        if s >= a.warmup:
            times.append(dt)
    tot = sum(times)
    gcups = cells * len(times) / tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"config2: {n} x (qlen=101,tlen=101) w=100 default scoring, bounded sample of the "
                               f"{a.jobs}-job batch", "jobs_per_step": n},
        "ext_per_s": n * len(times) / tot,
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} jobs per step, all host threads"},
        "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def capture_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, torch.distributed
    notices), at the C level as well, so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
    written to the saved descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    a = parse()
    capture_stdout()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines ("NCCL version ...") go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_

    import bwa_mem_quickassist_b200 as B
    import kswtest as K

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    ctx = B.KswB200(local, pack_threads=max(1, min(32, (os.cpu_count() or 8) // max(local_world, 1))))   # ranks share the host cores
    cfg = B.make_cfg()
    jobs, qpool, tpool = make_workload(a.jobs, a.seed + rank)      # every rank its own shard of reads
    n = a.jobs

    # ---- live DPX issue peak (the roofline denominator), measured before the timed region
    lane_ops, _ = ctx.dpx_peak(0)
    peak_gcups = lane_ops * 2.0 / 8.0 / 1e9

    # ---- resident path: pack + upload once, time the kernels
    launches0 = ctx.launch_count()
    rb = ctx.upload(cfg, jobs, qpool, tpool)
    info = rb.info()
    for _ in range(a.warmup):
        ctx.run(rb)
    ctx.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    ms, ms_ext = ctx.run_timed2(rb, a.steps)                       # CUDA events on the ctx stream: per step, and its extension kernels alone
    ctx.sync()
    l1 = ctx.launch_count()
    barrier()
    clocks = sampler.stop()
    t_dev = float(ms.sum()) * 1e-3
    t_max = max_over_ranks(t_dev)
    cells_job = ctx.download_cells(rb)
    cells = float(cells_job.astype(np.int64).sum())
    res_gpu = ctx.download(rb)
    nominal = float((jobs["qlen"].astype(np.int64) * jobs["tlen"]).sum())
    cells_all = sum_over_ranks(cells)
    nominal_all = sum_over_ranks(nominal)
    value = cells_all * a.steps / t_max / 1e9
    kernel_launches = l1 - l0
    rb.free()

    # ---- end to end through the C ABI with host buffers (copies, packing, binning, kernels, results inside)
    ne = a.e2e_jobs or n
    ej, eq, et = (jobs, qpool, tpool) if ne == n else (jobs[:ne], qpool, tpool)
    e_cells = float(cells_job[:ne].astype(np.int64).sum())
    # (1) page-locked caller buffers, asynchronous entry: what a host program that owns its buffers would call
    pj, pq, pt = B.pinned_copy(ej), B.pinned_copy(eq), B.pinned_copy(et)
    pr = B.PinnedArray(ne, B.RES_DT)
    for _ in range(max(a.warmup, 1)):
        ctx.extend_batch_async(cfg, pj.a, pq.a, pt.a, pr.a); ctx.wait()
    pr.a[:] = 0                                                    # so that the check below sees only timed-step results
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ctx.extend_batch_async(cfg, pj.a, pq.a, pt.a, pr.a)
        ctx.wait()
    t_e2e = time.perf_counter() - t0
    barrier()
    t_e2e_max = max_over_ranks(t_e2e)
    h2d, d2h = ctx.last_transfer()
    e2e_value = sum_over_ranks(e_cells) * a.steps / t_e2e_max / 1e9
    same = all((pr.a[f] == res_gpu[f][:ne]).all() for f in B.RES_DT.names)
    # (2) pageable caller buffers, synchronous entry (host packing into the library's pinned staging)
    res_e2e = np.zeros(ne, dtype=B.RES_DT)
    ctx.extend_batch(cfg, ej, eq, et, out=res_e2e)
    res_e2e[:] = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ctx.extend_batch(cfg, ej, eq, et, out=res_e2e)
    t_pg = time.perf_counter() - t0
    barrier()
    t_pg_max = max_over_ranks(t_pg)
    h2d_pg, d2h_pg = ctx.last_transfer()
    pg_value = sum_over_ranks(e_cells) * a.steps / t_pg_max / 1e9
    same = same and all((res_e2e[f] == res_gpu[f][:ne]).all() for f in B.RES_DT.names)
    for p in (pj, pq, pt, pr):
        p.close()

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only) + bit-exact check of that sample
    cpu = None
    parity = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        ns = min(a.cpu_sample, n)
        kind, res_cpu, dt = cpu_reference(K.make_cfg(), jobs[:ns], qpool, tpool, threads)
        ocells = visited_cells_oracle(K.make_cfg(), jobs[:ns], qpool, tpool, threads)
        ok = all((res_cpu[f] == res_gpu[f][:ns]).all() for f in B.RES_DT.names)
        ok_cells = bool((ocells == cells_job[:ns].astype(np.int64)).all())
        parity = {"sample_jobs": ns, "bit_exact": bool(ok), "cells_match_oracle": ok_cells, "e2e_equals_resident": bool(same)}
        n1 = min(ns, 100_000)                                    # and one core, as BASELINE.md 3.1(a) asks
        _, _, dt1 = cpu_reference(K.make_cfg(), jobs[:n1], qpool, tpool, 1)
        cpu = {"value": float(ocells.sum()) / dt / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"first {ns} jobs of the batch, {threads} host threads, one pass",
               "ext_per_s": ns / dt,
               "one_core": {"value": float(ocells[:n1].sum()) / dt1 / 1e9, "unit": UNIT, "ext_per_s": n1 / dt1, "sample": f"first {n1} jobs, 1 thread"}}

    mixes = None
    if rank == 0 and world == 1 and a.real_mix_reads > 0:
        mixes = real_mixes(ctx, peak_gcups, a.real_mix_reads)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = float(info["packed_bytes"] + 24 * n + 4 * n)      # job records + 2-bit pool in, results + cell counts out
        per_gpu_gcups = value / world
        # the dominant kernel's own launch duration (an event sits between the binning kernels and the extension launch of
        # every timed step): what the roofline figure divides by; `value` keeps the whole step
        kernel_ms = float(ms_ext.mean())
        kernel_gcups = cells / kernel_ms / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (int16 pairs; int32 kernel for out-of-range jobs)", "data": "synthetic",
            "config": {"workload": f"config2: {n} jobs/GPU x (qlen=101,tlen=101), w=100, zdrop=100, end_bonus=5, "
                                   "a=1 b=4 o=6 e=1, h0~U[19,100]", "jobs_per_gpu": n, "l2_policy": "inputs larger than L2 "
                                   f"({info['packed_bytes'] / 1e6:.0f} MB packed per GPU)",
                       "fast_jobs": info["n_fast"], "generic_jobs": info["n_generic"]},
            "ext_per_s": n * world * a.steps / t_max,
            "gcups_nominal_qlen_x_tlen": nominal_all * a.steps / t_max / 1e9,
            "visited_cells_per_job": cells / n,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ext_per_s": ne * world * a.steps / t_e2e_max, "jobs_per_step_per_gpu": ne,
                    "ms_per_step": 1e3 * t_e2e_max / a.steps,
                    "what": "ksw_b200_extend_batch_async + ksw_b200_wait on page-locked host byte-code buffers: raw H2D + "
                            "packing on the device (idle host threads pack some chunks) + binning + kernels + D2H into the caller's array"},
            "e2e_pageable": {"value": pg_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_pg), "d2h_bytes_per_step": int(d2h_pg),
                             "ms_per_step": 1e3 * t_pg_max / a.steps,
                             "what": "ksw_b200_extend_batch on pageable host buffers: host packing into pinned staging + H2D + "
                                     "binning + kernels + D2H + copy into the caller's array"},
            "gpu_launches": int(kernel_launches),
            "value_includes": "per step: binning (key kernel + radix sort) + extension kernels, all on the GPU",
            "roofline": {"bound": "dpx_issue", "achieved": kernel_gcups, "peak": peak_gcups, "unit": "GCUPS/GPU",
                         "frac": kernel_gcups / peak_gcups,
                         "kernel": "ksw_fast_kernel<KEYED>" if info["n_generic"] == 0 else "ksw_fast_kernel + int32 kernels",
                         "kernel_ms": kernel_ms, "step_ms": float(ms.mean()),
                         "achieved_basis": "this rank's visited cells per step / the extension kernels' own duration (CUDA events "
                                           "around them inside every timed step); the step also holds the binning kernels",
                         "frac_of_whole_step": per_gpu_gcups / peak_gcups,
                         # DRAM bytes per launch are not measurable from inside this run: the figure is the ncu --set full
                         # capture of the same kernel build on 400 k jobs (profiles/r2_ncu_fast_kernel_summary.txt:
                         # 40.79 MB read + 1.48 MB write = 105.7 B/job), scaled to this launch's job count
                         "traffic": 105.7 * n,
                         "traffic_source": "ncu capture of the same kernel (400k jobs, dram__bytes_read.sum+dram__bytes_write.sum = 105.7 B/job) "
                                           "x this launch's jobs; not measured in this run",
                         "peak_source": f"live DPX probe: {lane_ops / 1e12:.2f} T lane-ops/s x 2 cells / 8 issue slots "
                                        "(SURVEY.md 8d); not in MEASURED_PEAKS.json"},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes * a.steps / t_dev / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes * a.steps / t_dev / 1e9 / hbm_peak, "traffic": 105.7 * n,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
            "real_mix": mixes,
        }
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
