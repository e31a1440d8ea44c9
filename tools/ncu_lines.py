"""Per-source-line stall samples of an .ncu-rep (cuda,sass view).  usage: python tools/ncu_lines.py rep [top] [lo hi]"""
import csv, io, subprocess, sys


def I(x):
    try:
        return int(x)
    except Exception:
        return 0


rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10**9)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
lines = []
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        cur = dict(line=int(r[0]), src=r[1].strip(), samples=I(r[4]), inst=I(r[7]), sass=[])
        lines.append(cur)
    elif cur is not None:
        cur["sass"].append((r[3].strip(), I(r[4]), I(r[7])))
tot = sum(l["samples"] for l in lines)
print("total samples", tot)
sel = [l for l in lines if lo <= l["line"] <= hi]
for l in sorted(sel, key=lambda l: -l["samples"])[:top]:
    print("%5d %7d %5.1f%% inst=%-12d %s" % (l["line"], l["samples"], 100.0 * l["samples"] / tot, l["inst"], l["src"][:110]))
if len(sys.argv) > 4:
    print("--- SASS of the range (op, samples, executed)")
    for l in sorted(sel, key=lambda l: l["line"]):
        for s in l["sass"]:
            if "LDL" in s[0] or "STL" in s[0]:
                print(l["line"], s)
