#!/usr/bin/env python
"""SASS listing per hot kernel (north star: "a SASS listing per kernel").  Dumps the sm_100a SASS of every kernel of
libaxisem3d_b200.so with cuobjdump, writes one gzip'd listing per hot kernel under profiles/sass/ and a summary table
(profiles/r2_sass_summary.md): instruction count, registers / spills from the resource usage, and the opcode mix that shows
what the code is made of -- packed FP32 (FFMA2 / FADD2 / FMUL2), shared-memory traffic (LDS / STS), cp.async gathers
(LDGSTS), bulk-async copies (UBLKCP), scatter atomics (RED), barriers."""
import collections
import gzip
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "axisem3d_b200", "libaxisem3d_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
HOT = ["k_elem3d_fused", "k_elem1d", "k_newmark_solid", "k_newmark_fluid", "k_grad3d", "k_fft3d_v2", "k_quad3d", "k_sf_couple", "k_halo_put",
       "k_halo_wait_add", "k_mass3d", "k_source"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
os.makedirs(OUT, exist_ok=True)
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("void ", "")
    if not any(h in short for h in HOT):
        continue
    if "true>" in short and ("k_fft3d_v2" in short or "k_grad3d" in short or "k_quad3d" in short or "k_elem1d" in short):
        continue          # particle-relabelling instances: same code plus the 9-component branch
    if "k_fft3d_v2" in short and (", 512," in short or ", 1024," in short):
        continue          # same body as the 256-thread instance
    ops = collections.Counter()
    n = 0
    for ln in b.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1).split(".")[0]] += 1
            n += 1
    fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")
    with gzip.open(os.path.join(OUT, fn + ".sass.gz"), "wt") as f:
        f.write("Function : " + b)
    u = usage.get(name, (0, 0, 0, 0))
    pick = lambda *ks: sum(ops[k] for k in ks)
    rows.append((short, n, u[0], u[1], pick("FFMA2", "FADD2", "FMUL2"), pick("FFMA", "FADD", "FMUL"), pick("LDS"), pick("STS"), pick("LDGSTS"),
                 pick("UBLKCP"), pick("RED", "ATOM", "ATOMG"), pick("LDG"), pick("STG"), pick("LDL"), pick("STL"), pick("BAR")))
rows.sort(key=lambda r: -r[1])
with open(os.path.join(ROOT, "profiles", "r2_sass_summary.md"), "w") as f:
    f.write("# SASS of the hot kernels (sm_100a), static instruction counts\n\n")
    f.write("Full listings: `profiles/sass/<kernel>.sass.gz` (cuobjdump -sass; written by profiles/scripts/make_sass.py).\n")
    f.write("`FFMA2/FADD2/FMUL2` are the packed two-lane FP32 forms of sm_100a (`fma.rn.f32x2` ...), `LDGSTS` = cp.async, `UBLKCP` = 1-D bulk\n")
    f.write("async copy (TMA engine, in-kernel Newmark stage loads), `RED` = the stiffness scatter (`red.global.add.v2.f32`).\n\n")
    f.write("| kernel | SASS instr | regs | stack B | packed FP32 | scalar FP32 | LDS | STS | LDGSTS | UBLKCP | RED/ATOM | LDG | STG | LDL | STL | BAR |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| `%s` | %s |\n" % (r[0], " | ".join(str(x) for x in r[1:])))
print("wrote", len(rows), "listings")
