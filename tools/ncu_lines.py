"""Per-source-line summary of an ncu report (tools only): python tools/ncu_lines.py rep.ncu-rep [kernel-substring] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Function Name": cur = {"name": r[1], "lines": []}; blocks.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0].isdigit(): cur["lines"].append(r)
seen = set()
for b in blocks:
    if sub not in b["name"] or b["name"] in seen or not b["lines"]: continue
    seen.add(b["name"])
    h = b["hdr"]; ii = h.index("Instructions Executed"); si = h.index("# Samples"); ti = h.index("Thread Instructions Executed")
    tot_i = sum(int(r[ii]) for r in b["lines"]); tot_s = sum(int(r[si]) for r in b["lines"]) or 1
    print(f"== {b['name']}: warp-inst {tot_i}, samples {tot_s}")
    for r in sorted(b["lines"], key=lambda r: -int(r[si]))[:top]:
        print(f"{int(r[0]):5d} inst {100*int(r[ii])/tot_i:5.1f}%  samp {100*int(r[si])/tot_s:5.1f}%  thr/inst {int(r[ti])/max(int(r[ii]),1):4.1f} | {r[1].strip()[:110]}")
