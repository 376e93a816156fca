"""Per-instruction view of an `ncu --page source --csv` dump: regions of consecutive SASS with their share of
samples and executed instructions.  usage: python tools/ncu_hot.py src.csv [window]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))[2:]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ins = [(r[1].strip(), int(r[4] or 0), int(r[5] or 0)) for r in rows if len(r) > 5]
tot_s = sum(s for _, s, _ in ins) or 1
tot_e = sum(e for _, _, e in ins) or 1
print(f"{len(ins)} instructions, {tot_s} samples, {tot_e} warp-instructions executed")
for i in range(0, len(ins), win):
    blk = ins[i:i + win]
    s = sum(x[1] for x in blk)
    e = sum(x[2] for x in blk)
    nd = sum(1 for x in blk if x[0].startswith("DMMA"))
    ops = {}
    for x in blk:
        op = x[0].split()[0] if not x[0].startswith("@") else x[0].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    top = ",".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    if s * 200 > tot_s or e * 200 > tot_e:
        print(f"[{i:5d}-{i + len(blk):5d}) samples {100 * s / tot_s:5.1f}%  exec {100 * e / tot_e:5.1f}%  dmma {nd:2d}  {top}")
