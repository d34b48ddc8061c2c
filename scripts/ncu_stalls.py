"""Per-instruction view of an ncu report (captured with --set full --import-source on): where the stall samples sit.

usage: python scripts/ncu_stalls.py <report.ncu-rep> [top N instructions, default 25] [kernel-name substring]

For every kernel in the report (once per distinct name):
  * the warp-stall sample totals by reason,
  * code regions -- instructions grouped by their execution count (a loop body shares one count): share of the samples
    and of the executed instructions, which tells loops (lean / general / a-plane items) apart without symbols,
  * the N instructions with the most samples, in address order, with their two main stall reasons.
This is how the findings quoted in DESIGN.md section 5 were read (e.g. profiles/r2o_llg_stall_regions.txt)."""
import collections
import csv
import io
import subprocess
import sys


def kernels(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    out, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            out.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return out


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    seen = set()
    for k in kernels(rep):
        if pat not in k["name"] or k["name"] in seen:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        i_s, i_e = h.index("# Samples"), h.index("Instructions Executed")
        reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[i_s]) for r in k["rows"]) or 1
        tot_e = sum(int(r[i_e]) for r in k["rows"]) or 1
        print("==", k["name"][:120])
        print(f"   samples {tot}, executed warp instructions {tot_e}, SASS instructions {len(k['rows'])}")
        agg = {c[6:]: sum(int(r[h.index(c)]) for r in k["rows"]) for c in reasons}
        print("   stalls:", ", ".join(f"{n} {100 * v / tot:.1f}%" for n, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
        regions = collections.OrderedDict()
        for r in k["rows"]:
            d = regions.setdefault(int(r[i_e]), [0, 0, 0])
            d[0] += 1
            d[1] += int(r[i_s])
            d[2] += int(r[i_e])
        print("   regions (instructions sharing an execution count):")
        for e, (n, sm, ex) in sorted(regions.items(), key=lambda t: -t[1][1])[:8]:
            print(f"     executed {e:9d} x {n:5d} instr: samples {100 * sm / tot:5.1f}%  instructions {100 * ex / tot_e:5.1f}%")
        top = sorted(enumerate(k["rows"]), key=lambda t: -int(t[1][i_s]))[:ntop]
        for i, r in sorted(top):
            ss = sorted(((c[6:], int(r[h.index(c)])) for c in reasons if int(r[h.index(c)]) > 0), key=lambda t: -t[1])[:2]
            print(f"   {i:5d} {100 * int(r[i_s]) / tot:5.1f}% x{r[i_e]:>8} {str(ss):44s} {r[1].strip()[:80]}")


if __name__ == "__main__":
    main()
