"""Trimmed SASS evidence for profiles/: per hot kernel the opcode histogram (DMMA, UBLKCP, SYNCS, UTMA*, BAR, LDS/STS ...) and the
body of its densest DMMA loop (the backward branch whose body holds the most DMMAs).  usage: python tools/sass_summary.py lib.so > out.txt"""
import re
import subprocess
import sys

lib = sys.argv[1]
want = ["k2_col_kernelILb0ELi7", "k2_real_kernelILb0", "k2_cplx_kernelILb1", "k3_tma_kernelILb1ELi4", "k2_stream_kernelILb0", "k5_gemm_kernelILb1ELi128", "k4_jmn_kernelILb0", "k1_eig_kernel"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if not any(w in name for w in want):
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    hist = {}
    for _, t in ins:
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0] if not op.startswith(("DMMA", "UBLKCP", "SYNCS", "UTMA")) else op
        hist[op] = hist.get(op, 0) + 1
    key = ["DMMA.8x8x4", "UBLKCP.S.G", "UBLKCP.G.S", "UTMASTG.2D", "UTMACMDFLUSH", "SYNCS.EXCH.64", "SYNCS.ARRIVE.TRANS64", "SYNCS.ARRIVE.TRANS64.A1T0",
           "SYNCS.PHASECHK.TRANS64.TRYWAIT", "BAR", "FENCE", "MEMBAR", "USETMAXREG", "LDS", "STS", "LDG", "STG", "LDL", "STL", "DMUL", "DFMA", "DADD", "SHFL", "R2UR"]
    print(f"== {name}  ({len(ins)} instructions)")
    print("   " + "  ".join(f"{k}:{hist[k]}" for k in key if k in hist))
    # densest DMMA loop: backward branches
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]: i + 1]
                nd = sum("DMMA" in x for _, x in body)
                if nd and (best is None or nd / len(body) > best[0]):
                    best = (nd / len(body), body, nd)
    if best:
        print(f"   densest DMMA loop: {best[2]} DMMA in {len(best[1])} instructions")
        for a, t in best[1]:
            print(f"      /*{a:05x}*/ {t}")
    print()
