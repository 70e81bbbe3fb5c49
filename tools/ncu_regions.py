"""Instruction share per source line of one kernel of an ncu report, sorted by line (tools only):
    python tools/ncu_regions.py rep.ncu-rep kernel-substring [min_pct]"""
import csv, io, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Function Name": cur = {"name": r[1], "lines": []}; blocks.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0].isdigit(): cur["lines"].append(r)
for b in blocks:
    if sub not in b["name"] or not b["lines"]: continue
    h = b["hdr"]; ii = h.index("Instructions Executed")
    tot = sum(int(r[ii]) for r in b["lines"])
    print(b["name"], tot)
    for r in sorted(b["lines"], key=lambda r: int(r[0])):
        pct = 100 * int(r[ii]) / tot
        if pct >= min_pct: print(f"{int(r[0]):5d} {pct:5.1f}% | {r[1].strip()[:120]}")
    break
