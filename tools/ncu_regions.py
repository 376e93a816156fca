"""Region view of an `ncu --page source --csv` dump: the SASS is cut at every backward-branch target / barrier-ish
instruction into regions; for each region: samples (share), executed warp-instructions, DMMA count and the dominant
stall reasons.  usage: python tools/ncu_regions.py src.csv [min_share_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) > 5]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in body:
    ins.append({"src": r[1].strip(), "samp": int(r[4] or 0), "exec": int(r[5] or 0),
                "st": {h: int(r[i] or 0) for i, h in st_cols}})
tot = sum(x["samp"] for x in ins) or 1
# cut regions where the executed count changes by more than 2x (loop boundaries) or at SYNCS/BAR/USETMAXREG
regions = []
cur = []
prev = None
for x in ins:
    e = x["exec"]
    cut = False
    if prev is not None and (e > 2 * prev + 8 or prev > 2 * e + 8):
        cut = True
    if cut and cur:
        regions.append(cur)
        cur = []
    cur.append(x)
    prev = e
if cur:
    regions.append(cur)
pos = 0
for reg in regions:
    s = sum(x["samp"] for x in reg)
    if 100.0 * s / tot >= minshare:
        e = sum(x["exec"] for x in reg)
        nd = sum(1 for x in reg if "DMMA" in x["src"])
        agg = {}
        for x in reg:
            for h, v in x["st"].items():
                agg[h] = agg.get(h, 0) + v
        top = ", ".join(f"{h[6:]} {100 * v / max(s, 1):.0f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4])
        ops = {}
        for x in reg:
            t = x["src"].split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        topo = ",".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        print(f"[{pos:5d}-{pos + len(reg):5d}) samp {100 * s / tot:5.1f}%  exec/instr {e // max(len(reg), 1):8d}  dmma {nd:3d}  {topo}  | {top}")
    pos += len(reg)
