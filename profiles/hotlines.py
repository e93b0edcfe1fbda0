#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: hotlines.py <report.ncu-rep> <kernel regex> <library.so> [min_pct]
Joins `ncu --page source --csv` (per SASS instruction: samples, instructions executed) with
`nvdisasm -g` line info of the same cubin by instruction order."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kre, so = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[1])))
hdr = rows[1]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
inst = [(r[iSrc].strip(), int(r[iS] or 0), int(r[iI] or 0)) for r in rows[2:] if len(r) > iI]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r"^\s*\.text\.[^\n]*(%s)[^\n]*:\n" % kre, txt, re.M)
    if m:
        lines = txt[m.end():].split("\n")
        break
cur, seq = None, []
for ln in lines:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        seq.append((cur, m.group(1).strip()))
    if ln.startswith("//---------------------") and seq:
        break
n = min(len(seq), len(inst))
agg = {}
for k in range(n):
    key = seq[k][0]
    a = agg.setdefault(key, [0, 0])
    a[0] += inst[k][1]; a[1] += inst[k][2]
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print("kernel %s: %d SASS instructions, %d samples, %d warp-instructions executed" % (kre, n, ts, ti))
srcs = {}
for (key, (s, i)) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if key is None or (100.0 * s / ts < min_pct and 100.0 * i / ti < min_pct):
        continue
    fn, ln = key
    if fn not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", fn)
        srcs[fn] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[fn][ln - 1].strip()[:100] if 0 < ln <= len(srcs[fn]) else ""
    print("%5.1f%% samples %5.1f%% inst  %s:%d  %s" % (100.0 * s / ts, 100.0 * i / ti, fn, ln, text))
