"""Per-region instruction and stall-sample shares: python tools/ncu_regions.py REPORT LIB KERNEL file:lo-hi=name ..."""
import sys, os, runpy, io, contextlib
rep, lib, kname = sys.argv[1:4]
regions = []
for spec in sys.argv[4:]:
    loc, name = spec.split("=")
    fn, rng = loc.split(":")
    lo, hi = (int(v) for v in rng.split("-"))
    regions.append((fn, lo, hi, name))
sys.argv = ["ncu_lines.py", rep, lib, kname, "0"]
with contextlib.redirect_stdout(io.StringIO()):
    g = runpy.run_path(os.path.join(os.path.dirname(__file__), "ncu_lines.py"))
body, ix, base, linemap, stall_cols = g["body"], g["ix"], g["base"], g["linemap"], g["stall_cols"]
import collections
acc = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
cur = "other"
for r in body:
    off = int(r[ix["Address"]], 16) - base
    (fn, ln), _ = linemap.get(off, (("?", 0), ""))
    for rf, lo, hi, name in regions:
        if fn == rf and lo <= ln <= hi:
            cur = name
            break
    else:
        if fn == regions[0][0]: cur = "other"
    a = acc[cur]
    a[0] += float(r[ix["Instructions Executed"]] or 0)
    a[1] += float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    for c in stall_cols:
        v = float(r[ix[c]] or 0)
        if v: a[2][c[6:]] += v
ti = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
print(f"total warp-instructions {ti:.0f}, stall samples {ts:.0f}")
for name, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    top = ", ".join(f"{k} {v / a[1] * 100:.0f}%" for k, v in a[2].most_common(3)) if a[1] else ""
    print(f"{name:12s} inst {a[0] / ti * 100:5.1f}%  samples {a[1] / ts * 100:5.1f}%   [{top}]")
