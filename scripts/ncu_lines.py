#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <mangled kernel name> [top]
Joins `ncu --page source --csv` (per-address executed counts / stall samples, in address order) with
`nvdisasm -g` line info of the same kernel extracted from the built libsgmpf.so."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(PKG, "libsgmpf.so")], cwd=tmp, capture_output=True)
sass, start = None, None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):          # one cubin per translation unit
    lines = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    hit = [i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":")]
    if hit:
        sass, start = lines, hit[0]
        break
if sass is None:
    sys.exit("kernel %s not found in libsgmpf.so" % kern)
cur, seq = None, []
for l in sass[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
st = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
# several kernels in the report: take the segment whose kernel name matches argv[4] (substring), default the first
want = sys.argv[4] if len(sys.argv) > 4 else None
pick = 0
if want:
    for n, i in enumerate(st):
        if want in " ".join(rows[i + 1]) or (i + 2 < len(rows) and want in " ".join(rows[i + 2])):
            pick = n
            break
seg = rows[st[pick] + 1:st[pick + 1]] if pick + 1 < len(st) else rows[st[pick] + 1:]
hdr, data = seg[0], [r for r in seg[1:] if len(r) > 5]
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
print("sass instrs: nvdisasm", len(seq), "ncu", len(data))
inst, samp = collections.Counter(), collections.Counter()
for k in range(min(len(seq), len(data))):
    key = seq[k] or ("?", 0)
    inst[key] += int(data[k][iE] or 0)
    samp[key] += int(data[k][iS] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
src = {}
for fn in os.listdir(os.path.join(PKG, "csrc")):
    src[fn] = open(os.path.join(PKG, "csrc", fn)).read().split("\n")
print("total warp instructions", ti, "samples", ts)
for (f, ln), c in inst.most_common(top):
    text = src[f][ln - 1].strip()[:90] if f in src and ln - 1 < len(src[f]) else ""
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0 * c / ti, 100.0 * samp[(f, ln)] / ts, f, ln, text))
