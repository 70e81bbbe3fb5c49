"""Instruction / sample share per region of compose_kernel.cu for the k-th k_compose launch of an ncu report (tools only):
    python tools/ncu_regions.py rep.ncu-rep [launch-index] [top-lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Function Name": cur = {"name": r[1], "lines": []}; blocks.append(cur)
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0].isdigit(): cur["lines"].append(r)
blocks = [b for b in blocks if len(b["lines"]) > 200 and any("sprite_texel" in r[1] for r in b["lines"])]
b = blocks[which]
h = b["hdr"]; ii = h.index("Instructions Executed"); si = h.index("# Samples"); ti = h.index("Thread Instructions Executed")
tot = sum(int(r[ii]) for r in b["lines"]); ts = sum(int(r[si]) for r in b["lines"]) or 1
print(f"launch {which} of {len(blocks)}: warp-inst {tot}, samples {ts}")
# regions of compose_kernel.cu by the first code line of each stage
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixelmek-3d_b200", "csrc", "compose_kernel.cu")).read().split("\n")
keys = [("k_compose(const __grid_constant__", "prologue"), ("const int s0 = p.sprite_offset[view];", "warp sprite list"), ("const int cx = p.w / 2, cy = p.h / 2;", "convergence"),
        ("const int lx = lane & 15, ly = lane >> 4;", "lane setup"), ("for (int it = 0; it < WARP_ROWS / 2", "row loop head"), ("unsigned s_got = 0u;", "sprites f2b"),
        ("const unsigned s_base =", "bg select"), ("} else if (fast_ok && y >= fs_max", "floor fast"),
        ("} else if (fast_ok && y < sky_below)", "sky"), ("if (!done) {", "mixed rows"), ("if ((s_base >> j) & 1u) px[j] = s_top[j];", "combine"),
        ("const unsigned repaint", "repaint"), ("uint32_t *o = out32", "store"), ("__device__ __forceinline__ uint64_t word_hash", "after")]
bounds = [(0, "helpers")]
for k, name in keys:
    for n, line in enumerate(src):
        if k in line: bounds.append((n + 1, name)); break
for n, (lo, name) in enumerate(bounds):
    hi = bounds[n + 1][0] - 1 if n + 1 < len(bounds) else 10 ** 9
    a = sum(int(r[ii]) for r in b["lines"] if lo <= int(r[0]) <= hi); s = sum(int(r[si]) for r in b["lines"] if lo <= int(r[0]) <= hi)
    print(f"{lo:4d}- inst {100*a/tot:5.1f}% samples {100*s/ts:5.1f}%  {name}")
for r in sorted(b["lines"], key=lambda r: -int(r[si]))[:top]:
    print(f"{int(r[0]):5d} inst {100*int(r[ii])/tot:5.1f}%  samp {100*int(r[si])/ts:5.1f}%  thr/inst {int(r[ti])/max(int(r[ii]),1):4.1f} | {r[1].strip()[:110]}")
