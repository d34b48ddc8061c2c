"""Print the SASS of one kernel between two addresses (hex).  Usage: sass_dump.py <lib.so> <kernel-substring> <from> <to>"""
import re, subprocess, sys
lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", out)[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    print(name[:140])
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(f"{m.group(1)}  {m.group(2).strip()}")
    break
