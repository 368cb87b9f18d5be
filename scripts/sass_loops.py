#!/usr/bin/env python
"""sass_loops.py <file.sass> [function-substring] — static view of the loops of a kernel in a `cuobjdump -sass` dump:
for every backward branch, the body length and its instruction mix by issue pipe (ALU / FMA / LSU / other), which is
what bounds the extension kernels.  Used before spending GPU time on a variant."""
import re, sys

ALU = ("VIADDMNMX", "VIMNMX", "VIADD", "PRMT", "LOP3", "IADD3", "ISETP", "SEL", "SHF", "LEA", "IABS", "PLOP3", "IMNMX", "POPC", "FLO", "BREV", "SGXT", "BMSK", "P2R", "R2P", "VABSDIFF", "ICMP", "MOV", "CS2R")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")
LSU = ("LDS", "STS", "LDG", "STG", "LD.", "ST.", "ATOM", "RED", "LDC", "LDL", "STL", "SHFL")

def pipe(op):
    if op.startswith("IMAD") or op.startswith(FMA): return "FMA"
    if op.startswith(LSU): return "LSU"
    if op.startswith(ALU): return "ALU"
    return "OTH"

def main():
    txt = open(sys.argv[1]).read().split("Function : ")
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for fn in txt[1:]:
        name = fn.split("\n", 1)[0]
        if want not in name: continue
        ins = []
        for ln in fn.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if not m: continue
            addr = int(m.group(1), 16); text = m.group(2).strip()
            body = re.sub(r"^@!?U?P\d+\s+", "", text)
            ins.append((addr, body.split()[0], text))
        print(f"== {name[:100]}  ({len(ins)} instructions)")
        amap = {a: i for i, (a, _, _) in enumerate(ins)}
        for i, (a, op, text) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if not m: continue
                t = int(m.group(1), 16)
                if t in amap and amap[t] <= i:
                    body = ins[amap[t]:i + 1]
                    cnt = {"ALU": 0, "FMA": 0, "LSU": 0, "OTH": 0}
                    ops = {}
                    for _, o, _ in body:
                        cnt[pipe(o)] += 1
                        k = o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("VIADDMNMX", "VIMNMX")) and "." in o else "")
                        ops[k] = ops.get(k, 0) + 1
                    print(f"loop {t:#06x}..{a:#06x}: {len(body):4d} instr  ALU {cnt['ALU']:3d} FMA {cnt['FMA']:3d} LSU {cnt['LSU']:3d} other {cnt['OTH']:3d}   "
                          + " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))

if __name__ == "__main__":
    main()
