"""Generates tests/golden/ksw_extend_golden.npz from the REFERENCE's own ksw_extend2, ksw_global_golden.npz from its
ksw_global2, ksw_align_golden.npz from its ksw_align2 (oracle/_ref/libksw_ref.so = bwa-0.7.8/ksw.c compiled unmodified by oracle/Makefile) and
chain2aln_golden.npz from its mem_chain2aln (oracle/_ref/libbwa_ref.so).

Run in the build container (where /root/reference is mounted):  python tests/golden/make_golden.py
The reference ships no known-answer vectors for this path, so these outputs of the
compiled reference are the golden vectors; the oracle restatement, the CPU emulation of the kernel
source and the CUDA path are all checked against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import kswtest as K  # noqa: E402


def main():
    K.build_oracle()
    assert K.have_ref(), "oracle/_ref/libksw_ref.so missing: /root/reference must be mounted to regenerate"
    sets = {
        "adversarial": K.gen_adversarial(seed=7),
        "fuzz_default": K.gen_fuzz(3000, seed=101),
        "fuzz_asym": K.gen_fuzz(1500, seed=102, cfg=K.make_cfg(a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, end_bonus=9)),
        "fuzz_bwasw": K.gen_fuzz(1000, seed=103, cfg=K.make_cfg(zdrop=-1, end_bonus=0)),
        "config2": K.gen_config2(3000, seed=104),
        "highindel250": K.gen_fuzz(1000, seed=105, max_q=250, related=1.0, w_choices=(100,), h0_max=120),
    }
    out = {}
    for name, b in sets.items():
        res = K.run_ref(b, threads=4)
        out[f"{name}.jobs"] = b.jobs
        out[f"{name}.qpool"] = b.qpool
        out[f"{name}.tpool"] = b.tpool
        out[f"{name}.cfg"] = np.frombuffer(bytes(b.cfg), dtype=np.uint8).copy()
        out[f"{name}.res"] = res
        print(name, b.n, "jobs")
    path = os.path.join(HERE, "ksw_extend_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # banded global alignment with backtrace: scores and CIGARs from the reference's own ksw_global2
    gsets = {
        "default": K.gen_global(1500, seed=201),
        "asym": K.gen_global(800, seed=202, cfg=K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3)),
        "cheap_gaps": K.gen_global(800, seed=203, cfg=K.make_cfg(a=1, b=1, o_del=1, e_del=2, o_ins=2, e_ins=1)),
        "short": K.gen_global(600, seed=204, max_q=12, w_extra=(0, 1, 2)),
    }
    out = {}
    for name, b in gsets.items():
        res, pool = K.run_global_ref(b, threads=4)
        cig = np.concatenate([pool[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])] for r in res])
        dense = res.copy()
        dense["cigar_off"] = np.concatenate([[0], np.cumsum(res["n_cigar"].astype(np.int64))[:-1]])
        out[f"{name}.jobs"] = b.jobs
        out[f"{name}.qpool"] = b.qpool
        out[f"{name}.tpool"] = b.tpool
        out[f"{name}.cfg"] = np.frombuffer(bytes(b.cfg), dtype=np.uint8).copy()
        out[f"{name}.res"] = dense
        out[f"{name}.cigar"] = cig.astype(np.uint32)
        print("global", name, b.n, "jobs", len(cig), "operations")
    path = os.path.join(HERE, "ksw_global_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # local alignment (mate rescue): kswr_t records from the reference's own ksw_align2 (striped SSE2 kernels)
    flags = [0, K.KSW_XBYTE, K.KSW_XSTART, K.KSW_XSUBO, K.KSW_XSUBO | K.KSW_XSTART, K.KSW_XBYTE | K.KSW_XSUBO | K.KSW_XSTART,
             K.KSW_XSTOP, K.KSW_XSTOP | K.KSW_XSTART, K.KSW_XBYTE | K.KSW_XSTOP | K.KSW_XSUBO | K.KSW_XSTART]
    asets = {
        "matesw150": K.gen_align(500, seed=301, max_q=150, max_t=700),
        "matesw250": K.gen_align(300, seed=302, max_q=250, max_t=900),
        "flags": K.gen_align(600, seed=303, max_q=100, max_t=300, flags=flags),
        "asym": K.gen_align(300, seed=304, cfg=K.make_cfg(a=2, b=7, o_del=0, e_del=1, o_ins=11, e_ins=3), max_q=120, max_t=400),
        "zero_open": K.gen_align(300, seed=305, cfg=K.make_cfg(a=3, b=2, o_del=0, e_del=1, o_ins=0, e_ins=1), max_q=80, max_t=300, flags=flags),
    }
    out = {}
    for name, b in asets.items():
        res = K.run_align_ref(b, threads=4)
        out[f"{name}.cfg"] = np.frombuffer(bytes(b.cfg), dtype=np.uint8).copy()
        out[f"{name}.jobs"] = b.jobs
        out[f"{name}.qpool"] = b.qpool
        out[f"{name}.tpool"] = b.tpool
        out[f"{name}.res"] = res
        print("align", name, b.n, "jobs")
    path = os.path.join(HERE, "ksw_align_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # mem_chain2aln level: regions from the reference's own mem_chain2aln (oracle/_ref/libbwa_ref.so)
    assert K.have_bwa_ref()
    csets = {
        "default": K.gen_chains(300, seed=11),
        "narrow_w": K.gen_chains(300, seed=12, opt=K.make_ext_opt(w=12), indel=0.01, max_indel=14),
        "hi_indel250": K.gen_chains(200, seed=13, sub=0.04, indel=0.02, max_indel=12, read_lens=(250,)),
        "asym": K.gen_chains(200, seed=14, opt=K.make_ext_opt(a=2, b=5, o_del=4, e_del=2, o_ins=8, e_ins=1, pen_clip5=7,
                                                                pen_clip3=3, w=60, zdrop=40)),
        "with_n": K.gen_chains(200, seed=15, n_frac=0.03),
    }
    out = {}
    for name, cs in csets.items():
        regs, reg_read = K.run_chain_ref(cs)
        for f in ("pac", "read_off", "read_len", "qpool", "chain_read", "chain_seed0", "chain_nseeds", "seeds"):
            out[f"{name}.{f}"] = getattr(cs, f)
        out[f"{name}.l_pac"] = np.array([cs.l_pac], dtype=np.int64)
        out[f"{name}.opt"] = np.frombuffer(bytes(cs.opt), dtype=np.uint8).copy()
        out[f"{name}.regs"] = regs
        out[f"{name}.reg_read"] = reg_read
        print(name, cs.n_reads, "reads", cs.n_chains, "chains", len(regs), "regions")
    path = os.path.join(HERE, "chain2aln_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
