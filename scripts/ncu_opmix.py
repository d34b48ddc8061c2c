"""Dynamic SASS opcode mix + hottest instructions from an .ncu-rep source page.  Usage: ncu_opmix.py REP [kernel-substr]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the file holds one block per launch: a line with the kernel name, a header line, then instructions
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Address":
        cur = {"hdr": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0].startswith("0x") or (cur is not None and r and r[0].isdigit()):
        cur["rows"].append(r)
    elif r and cur is None:
        pass
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = {n: i for i, n in enumerate(b["hdr"])}
mix = collections.Counter()
tot = 0
hot = []
for r in b["rows"]:
    try:
        n = int(float(r[h["Instructions Executed"]] or 0))
    except ValueError:
        continue
    ops = r[h["Source"]].split()
    op = ops[1] if ops and ops[0].startswith("@") else (ops[0] if ops else "?")
    mix[op.split(".")[0]] += n
    tot += n
    hot.append((int(float(r[h["# Samples"]] or 0)), n, r[h["Source"]][:90]))
print("total warp-instructions:", tot)
for op, n in mix.most_common(28):
    print(f"  {op:14s} {n:14d} {100*n/tot:6.2f} %")
print("hottest by stall samples:")
for s, n, src in sorted(hot, reverse=True)[:25]:
    print(f"  {s:7d} {n:12d}  {src}")
