#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (SASS level) with `nvdisasm -g` line info and aggregate executed
warp-instructions / stall samples / shared wavefronts per source line.  usage: join_sass_lines.py src.csv cubin kernel"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kern = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
sec = dis[dis.index(".text." + kern + ":"):]
nxt = sec.find("//--------------------- .text.", 10)
sec = sec[:nxt] if nxt > 0 else sec
line_of = {}
cur = ("?", 0)
stack = None
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, isamp, iwf = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    off = int(r[ia], 16) - base
    (loc, ins) = line_of.get(off, (("?", 0), ""))
    a = agg[loc]
    a[0] += int(r[ie]); a[1] += int(r[isamp]); a[2] += int(r[iwf] or 0); a[3] += 1
    tot[0] += int(r[ie]); tot[1] += int(r[isamp]); tot[2] += int(r[iwf] or 0)
print("total warp-instr %d  samples %d  shared wavefronts %d" % tuple(tot))
print("%-22s %6s %12s %7s %8s %7s %12s" % ("file:line", "sass", "warp-instr", "%instr", "samples", "%samp", "smem-wavefr"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[5]) if len(sys.argv) > 5 else 1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%-22s %6d %12d %6.1f%% %8d %6.1f%% %12d" % ("%s:%d" % loc, a[3], a[0], 100.0 * a[0] / tot[0], a[1], 100.0 * a[1] / tot[1], a[2]))

# ---- phase totals for fused.cuh kernels (line ranges of the current source; informational)
def phase(loc):
    f, l = loc
    if f == "fft.cuh": return "fft butterflies (fft.cuh)"
    if f == "fused.cuh":
        if 20 <= l <= 62: return "fft stage load/store/twiddle/index"
        if 63 <= l <= 80: return "fft dispatch"
        if 85 <= l <= 116: return "prologue"
        if 117 <= l <= 134: return "gather"
        if 135 <= l <= 168: return "grad"
        if 169 <= l <= 181: return "c2r loop/barriers"
        if 182 <= l <= 226: return "stress"
        if 227 <= l <= 238: return "r2c loop/barriers"
        if 239 <= l <= 282: return "quad-pre"
        if l >= 283: return "quad-post+scatter"
    return f
ph = collections.defaultdict(lambda: [0, 0, 0])
for loc, a in agg.items():
    p = ph[phase(loc)]
    p[0] += a[0]; p[1] += a[1]; p[2] += a[2]
print()
for k, p in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print("%-40s instr %5.1f%%  samples %5.1f%%  smem wavefronts %5.1f%%" % (k, 100.0 * p[0] / tot[0], 100.0 * p[1] / tot[1], 100.0 * p[2] / max(tot[2], 1)))
