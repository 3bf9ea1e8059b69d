"""Attribute the warp-stall samples of an .ncu-rep to CUDA source lines (ncu's CSV source page is SASS only).
usage: python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [TOP_N]
Joins ncu's per-SASS-instruction samples with `nvdisasm --print-line-info` of the cubin embedded in LIB.so."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernel_full = rows[0][1]
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(body[0][ix["Address"]], 16)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
# mangled-name fragment to find: use template args from kernel_full e.g. sweep_blocked_kernel<(int)2, (int)4, (int)3>
m = re.search(r"(\w+)<(.*)>\(", kernel_full)
frag = kname
if m:
    parts = []
    for a in [x.strip() for x in m.group(2).split(",")]:
        mi = re.match(r"\(int\)(\d+)$", a); mb = re.match(r"\(bool\)(\d+)$", a)
        if mi: parts.append("Li%sE" % mi.group(1))
        elif mb: parts.append("Lb%sE" % mb.group(1))
        elif a == "double": parts.append("d")
        else: parts.append("%d%s" % (len(a), a))          # a named type, e.g. double2 -> 7double2
    frag = m.group(1) + "I" + "".join(parts)
linemap = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if frag not in out: continue
    insec, cur = False, None
    for line in out.splitlines():
        if line.startswith("//-") and ".text." in line:
            insec = frag in line
            continue
        if not insec: continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if mm:
            if "inlined at" in mm.group(3) and cur is not None and False: pass
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if mm and cur: linemap[int(mm.group(1), 16)] = (cur, mm.group(2).strip())
agg = collections.defaultdict(lambda: collections.Counter())
instr = collections.Counter()
tot = 0
for r in body:
    off = int(r[ix["Address"]], 16) - base
    s = float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    key = linemap.get(off, (("?", 0), ""))[0]
    tot += s
    agg[key]["samples"] += s
    instr[key] += float(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols:
        v = float(r[ix[c]] or 0)
        if v: agg[key][c] += v
print(f"kernel: {kernel_full}\ntotal stall samples {tot:.0f}; instructions executed {sum(instr.values()):.0f}")
srcs = {}
for key, cnt in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    fn, ln = key
    if fn not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", fn)
        srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[fn][ln - 1].strip()[:90] if 0 < ln <= len(srcs[fn]) else ""
    reasons = ", ".join(f"{c[6:]} {v / cnt['samples'] * 100:.0f}%" for c, v in cnt.most_common(4) if c != "samples")
    print(f"{cnt['samples'] / tot * 100:5.1f}%  inst {instr[key] / sum(instr.values()) * 100:4.1f}%  {fn}:{ln:<4} {text}\n          [{reasons}]")

if os.environ.get("SHARED_CONFLICTS") and "L1 Wavefronts Shared Excessive" in ix:
    exc, wav = collections.Counter(), collections.Counter()
    for r in body:
        off = int(r[ix["Address"]], 16) - base
        key = linemap.get(off, (("?", 0), ""))[0]
        exc[key] += float(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        wav[key] += float(r[ix["L1 Wavefronts Shared"]] or 0)
    te, tw = sum(exc.values()), sum(wav.values())
    print(f"\nshared-memory wavefronts {tw:.0f}, excessive (bank conflicts) {te:.0f} = {te / max(tw, 1) * 100:.1f} %; lines with the most excessive wavefronts:")
    for key, v in exc.most_common(12):
        fn, ln = key
        if fn not in srcs:
            pth = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", fn)
            srcs[fn] = open(pth).read().splitlines() if os.path.exists(pth) else []
        text = srcs[fn][ln - 1].strip()[:100] if 0 < ln <= len(srcs[fn]) else ""
        print(f"{v / max(te, 1) * 100:5.1f}%  ({v:.0f} of {wav[key]:.0f} wavefronts)  {fn}:{ln:<4} {text}")

if os.environ.get("SASS_RANGE"):
    lo, hi = (int(v) for v in os.environ["SASS_RANGE"].split("-"))
    print(f"\nSASS with most samples attributed to lines {lo}-{hi} (and inlined headers in between):")
    sel = []
    inside = False
    for r in body:
        off = int(r[ix["Address"]], 16) - base
        (fn, ln), txt = linemap.get(off, (("?", 0), ""))
        if fn == os.environ.get("SASS_FILE", "sweep_blocked.cu"): inside = lo <= ln <= hi
        if inside:
            s = float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
            reasons = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            sel.append((s, off, fn, ln, r[ix["Source"]].strip(), reasons, float(r[ix["Instructions Executed"]] or 0)))
    ssum = sum(x[0] for x in sel); isum = sum(x[6] for x in sel)
    print(f"  range total: {ssum:.0f} samples ({ssum / tot * 100:.1f}%), {isum:.0f} warp-instructions ({len(sel)} SASS instructions)")
    for s, off, fn, ln, txt, reasons, ie in sorted(sel, reverse=True)[:int(os.environ.get("SASS_TOP", "40"))]:
        print(f"  {s:6.0f}  {off:06x} {fn}:{ln:<4} {txt[:70]:70s} {reasons[0][1]} {reasons[0][0]:.0f}, {reasons[1][1]} {reasons[1][0]:.0f}  ie={ie:.0f}")
