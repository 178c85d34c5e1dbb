#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel counts of the sm_100a mnemonics that matter for this path
(UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, FADD2/FMUL2/FFMA2 = packed FP32, BAR/ATOM for the rest)
from `cuobjdump -sass libtpb200.so`, plus the TMA issue sequence and the phase-1 filter loop of the
headline instantiation k_interact_tiles<3, 3, float, float, Wendland C2, ContinuityDensity>.
usage: python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "trixiparticles.jl_b200", "libtpb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
MNEMONICS = ["UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA2", "LDS.128", "STS.U16", "BAR", "ATOM", "RED", "MUFU", "LDTM", "UTCMMA"]
kernels, cur, arch = collections.OrderedDict(), None, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        kernels[cur].append(line.rstrip())
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)}: {len(kernels)} kernels, arch {sorted(arch)}")
tot = collections.Counter()
rows = []
for k, lines in kernels.items():
    c = collections.Counter()
    for l in lines:
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn) + r"\b", l):
                c[mn] += 1
    tot.update(c)
    rows.append((k, len(lines), c))
print("totals: " + ", ".join(f"{mn} {tot[mn]}" for mn in MNEMONICS))
print()
print("hot-path kernels (3-D, Float32 fields and coordinates, Wendland C2, ContinuityDensity, free-slip wall):")
want = ["k_interact_tilesILi3ELi3EffLi0ELi0ELb0E", "k_adami_tilesILi3ELi3EffLi0ELb0E", "k_summation_tilesILi3ELi3EffLi0E",
        "k_scan_cells_tiles", "k_post_scanILi3Eff", "k_reorder_fluidILi3EffLi0E", "k_cell_countILi3Ef", "k_driftILi3Eff",
        "k_halo_packIff", "k_halo_installIff", "k_adaptive_constsIf"]
for w in want:
    for k, n, c in rows:
        if w in k:
            name = demangle(k).split("(")[0]
            print(f"  {name[:86]:86s} {n:6d} instr  " + " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn]))
            break
print()
hot = next((k for k in kernels if "k_interact_tilesILi3ELi3EffLi0ELi0ELb0E" in k), None)
if hot:
    lines = kernels[hot]
    strip = lambda l: re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip()
    idx = [i for i, l in enumerate(lines) if "UBLKCP" in l]
    if idx:
        print("TMA issue (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes -> UBLKCP, armed by SYNCS.ARRIVE.TRANS64):")
        for l in lines[max(0, idx[0] - 6): idx[0] + 8]:
            print("   ", strip(l))
        print()
    # the phase-1 loop: the densest run of LDS.128 + FADD2
    best, best_i = -1, 0
    for i in range(0, len(lines) - 60):
        sc = sum(("LDS.128" in l) + ("FADD2" in l) + ("FMUL2" in l) for l in lines[i:i + 60])
        if sc > best:
            best, best_i = sc, i
    print("phase-1 filter loop (broadcast LDS.128 of candidate records, FADD2/FMUL2 distance test, predicated STS.U16 append):")
    for l in lines[best_i: best_i + 60]:
        print("   ", strip(l))
