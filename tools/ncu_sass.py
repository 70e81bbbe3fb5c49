"""Per-SASS-instruction view of one kernel of an ncu report (tools only): execution multiplicity per warp, stall samples.
    python tools/ncu_sass.py rep.ncu-rep kernel-regex n_warps [top]"""
import csv, io, subprocess, sys, collections
rep, kre, W = sys.argv[1], sys.argv[2], int(sys.argv[3]); top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ii = h.index("Instructions Executed"); si = h.index("# Samples"); ti = h.index("Avg. Threads Executed")
body = [r for r in rows[hi + 1:] if len(r) > ii and r[ii].isdigit()]
tot = sum(int(r[ii]) for r in body); ts = sum(int(r[si]) for r in body) or 1
print(f"{len(body)} SASS instructions, {tot} warp-inst = {tot / W:.0f} per warp, {ts} samples")
c = collections.Counter()
for r in body: c[round(int(r[ii]) / W * 2) / 2] += int(r[ii])
print("share of executed instructions by per-warp multiplicity:", {k: round(100 * v / tot, 1) for k, v in sorted(c.items()) if v / tot > 0.005})
with open("/tmp/sass_dump.txt", "w") as f:
    for n, r in enumerate(body): f.write(f"{n:5d} {int(r[ii]) / W:5.2f} s{int(r[si]):5d} t{float(r[ti]):4.1f} {r[1]}\n")
for n, r in sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:top]:
    print(f"{100 * int(r[si]) / ts:4.1f}% #{n:5d} x{int(r[ii]) / W:5.2f} t{float(r[ti]):4.1f} {r[1][:90]}")
