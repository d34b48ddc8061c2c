"""List backward branches (loops) of a kernel's SASS with their instruction counts and opcode mix.
Usage: sass_loops.py <lib.so> <kernel-name-substring>"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(name[:120], "total", len(ins))
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*$", t)
        m2 = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]: i + 1]
                mix = collections.Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for _, x in body)
                if len(body) > 20:
                    print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instrs", mix.most_common(14))
