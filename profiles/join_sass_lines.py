#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (SASS level) with `nvdisasm -gi` line info.

usage: join_sass_lines.py src.csv cubin mangled_kernel [top_n] [sort: instr|samples|smem]

Prints (1) totals, (2) the top source lines (innermost frame) and (3) a per-phase table, where a phase is the
range between two `// @phase <name>` markers of the kernel's own source file (outermost inline frame)."""
import collections, csv, re, subprocess, sys

src_csv, cubin, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sort = {"instr": 0, "samples": 1, "smem": 2}[sys.argv[5] if len(sys.argv) > 5 else "samples"]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
sec = dis[dis.index(".text." + kern + ":"):]
nxt = sec.find("//--------------------- .text.", 10)
sec = sec[:nxt] if nxt > 0 else sec
_src = {}
def src_line(path, n):
    if path not in _src:
        try:
            _src[path] = open(path).read().split("\n")
        except OSError:
            _src[path] = []
    L = _src[path]
    return L[n - 1] if 0 < n <= len(L) else ""
def pick_outer(chain):
    """outermost frame that is not merely the call of an inlined element body (so phases inside it are seen)"""
    for f in reversed(chain):
        if "fused_pass<" not in src_line(*f) and "fused_element<" not in src_line(*f):
            return f
    return chain[-1]
loc_of, chain, fresh = {}, [("?", 0)], True
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*)", ln)
    if m:
        loc_of[int(m.group(1), 16)] = (chain[0], pick_outer(chain), m.group(2))
        fresh = True
# phase markers of the outermost file
marks = {}
def phases_of(path):
    if path not in marks:
        ms = []
        try:
            for n, l in enumerate(open(path), 1):
                m = re.search(r"@phase\s+(.+?)\s*$", l)
                if m:
                    ms.append((n, m.group(1)))
        except OSError:
            pass
        marks[path] = ms
    return marks[path]
def phase(outer):
    name = "(unmarked) " + outer[0].split("/")[-1]
    for n, nm in phases_of(outer[0]):
        if outer[1] >= n:
            name = nm
    return name
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, isamp, iwf = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
ph = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    inner, outer, ins = loc_of.get(int(r[ia], 16) - base, (("?", 0), ("?", 0), ""))
    v = (int(r[ie]), int(r[isamp]), int(r[iwf] or 0))
    a = agg[(inner[0].split("/")[-1], inner[1])]
    p = ph[phase(outer)]
    for k in range(3):
        a[k] += v[k]; p[k] += v[k]; tot[k] += v[k]
    a[3] += 1
print("total warp-instr %d  samples %d  shared wavefronts %d" % tuple(tot))
print("%-24s %6s %12s %7s %8s %7s %12s" % ("file:line (innermost)", "sass", "warp-instr", "%instr", "samples", "%samp", "smem-wavefr"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][sort])[:topn]:
    print("%-24s %6d %12d %6.1f%% %8d %6.1f%% %12d" % ("%s:%d" % loc, a[3], a[0], 100.0 * a[0] / tot[0], a[1], 100.0 * a[1] / max(tot[1], 1), a[2]))
print()
print("%-44s %8s %9s %9s" % ("phase (outermost frame)", "%instr", "%samples", "%smem-wf"))
for k, p in sorted(ph.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %7.1f%% %8.1f%% %8.1f%%" % (k, 100.0 * p[0] / tot[0], 100.0 * p[1] / max(tot[1], 1), 100.0 * p[2] / max(tot[2], 1)))

if len(sys.argv) > 6:      # opcode mix of one phase
    want = sys.argv[6]
    ops = collections.Counter()
    for r in rows[2:]:
        inner, outer, ins = loc_of.get(int(r[ia], 16) - base, (("?", 0), ("?", 0), ""))
        if phase(outer) == want:
            op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
            ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG")) and "." in op else "")] += int(r[ie])
    t = sum(ops.values())
    print("\nopcode mix of phase '%s' (%d warp-instr)" % (want, t))
    for op, n in ops.most_common(25):
        print("  %-14s %10d %5.1f%%" % (op, n, 100.0 * n / t))

# ---- stall reasons per phase (sampled)
names = hdr[30:47]
st = collections.defaultdict(lambda: [0] * len(names))
for r in rows[2:]:
    inner, outer, ins = loc_of.get(int(r[ia], 16) - base, (("?", 0), ("?", 0), ""))
    a = st[phase(outer)]
    for k in range(len(names)):
        a[k] += int(r[30 + k] or 0)
keep = [k for k in range(len(names)) if sum(a[k] for a in st.values()) > 0.01 * max(tot[1], 1)]
print("\nstall samples per phase: " + " ".join("%9s" % names[k].replace("stall_", "")[:9] for k in keep))
for ph_name, a in sorted(st.items(), key=lambda kv: -sum(kv[1])):
    print("%-28s" % ph_name[:28] + " ".join("%9d" % a[k] for k in keep))
