"""Stall samples of one profiled kernel by CUDA source line and by device function: joins the SASS rows of an
.ncu-rep source page with nvdisasm's line table of the SAME build of librg_b200.so.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [mangled-substring, default rg_kernelILi10ELb0] [top N]"""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else "rg_kernelILi10ELb0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
# every Taylor order is its own translation unit: take the object of the profiled instantiation (the cubins inside the
# linked library all carry the same name and overwrite each other on extraction)
mo = re.search(r"rg_kernelILi(\d+)ELb(\d)", sub)
obj = os.environ.get("NCU_LINES_OBJ") or (os.path.join(root, "rg_b200", "csrc", "build", "inst_%s_%s.o" % (mo.group(1), mo.group(2)))
                                           if mo else os.path.join(root, "rg_b200", "librg_b200.so"))
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
lines = []
for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):   # one cubin per translation unit
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    cur, on = None, False
    for l in dis:
        if l.startswith(".text."):
            on = sub in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2)))
    if lines:
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r)
ix = {k: i for i, k in enumerate(rows[h])}
data = rows[h + 1:]
assert len(data) == len(lines), (len(data), len(lines), "the report and the library are different builds")
def num(x):
    try: return float(x)
    except ValueError: return 0.0
# enclosing function of every source line
funcs = {}
for fn in ("rg_device.cuh", "rg_multiply.cuh", "rg_b200.cu"):
    cur_f, tab = "?", []
    for n, l in enumerate(open(os.path.join(root, "rg_b200", "csrc", fn)), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", l)
        if m: cur_f = m.group(1)
        tab.append(cur_f)
    funcs[fn] = tab
by_line, by_func, stall_f = Counter(), Counter(), {}
wave_f, ideal_f, wave_l = Counter(), Counter(), Counter()   # shared-memory wavefronts by function / line
stall_cols = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
tot = 0.0
for (off, loc, sass), r in zip(lines, data):
    s = num(r[ix["# Samples"]]); tot += s
    by_line[loc] += s
    f = funcs.get(loc[0], ["?"] * 100000)[loc[1] - 1] if loc else "?"
    by_func[f] += s
    if "L1 Wavefronts Shared" in ix:
        wave_f[f] += num(r[ix["L1 Wavefronts Shared"]]); ideal_f[f] += num(r[ix["L1 Wavefronts Shared Ideal"]])
        wave_l[loc] += num(r[ix["L1 Wavefronts Shared"]])
    d = stall_f.setdefault(f, Counter())
    for k in stall_cols: d[k] += num(r[ix[k]])
print("total samples %d" % tot)
print("-- by device function (innermost inlined frame)")
for f, s in by_func.most_common(16):
    top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(s, 1)) for k, v in stall_f[f].most_common(4))
    print("%-18s %6.2f%%   %s" % (f, 100 * s / tot, top))
if wave_f:
    wt = sum(wave_f.values())
    print("-- shared-memory wavefronts by device function (total %.3g, ideal %.3g)" % (wt, sum(ideal_f.values())))
    for f, v in wave_f.most_common(10):
        print("%-22s %6.2f%%   x%.2f of ideal" % (f, 100 * v / wt, v / max(ideal_f[f], 1)))
    print("-- shared-memory wavefronts by source line")
    for loc, v in wave_l.most_common(12):
        fn, ln = loc if loc else ("?", 0)
        path = os.path.join(root, "rg_b200", "csrc", fn)
        text = open(path).read().splitlines()[ln - 1].strip() if loc and os.path.exists(path) else ""
        print("%-16s %5d %6.2f%%  %s" % (fn, ln, 100 * v / wt, text[:100]))
print("-- by source line")
for (loc, s) in by_line.most_common(topn):
    fn, ln = loc if loc else ("?", 0)
    path = os.path.join(root, "rg_b200", "csrc", fn)
    text = open(path).read().splitlines()[ln - 1].strip() if loc and os.path.exists(path) else ""
    print("%-16s %5d %6.2f%%  %s" % (fn, ln, 100 * s / tot, text[:100]))
