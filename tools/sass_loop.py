#!/usr/bin/env python
"""Static opcode mix of a kernel's SASS (whole function or its hottest backward-branch loop).

    python tools/sass_loop.py [--lib path.so] --kernel fwdc_kernelIdLi5ELi0E [--loop]

--loop restricts the count to the largest loop (the span of the longest backward branch).  A static count is a proxy
only (slow paths that never execute are included); the dynamic mix comes from ncu (profiles/).
"""
import argparse, collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "differentiable_tube_mpc_b200", "libdtmpc_b200.so"))
ap.add_argument("--kernel", required=True)
ap.add_argument("--loop", action="store_true")
ap.add_argument("--dump", action="store_true")
a = ap.parse_args()
out = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
ins = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);")
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if a.kernel not in name:
        continue
    rows = [(int(m.group(1), 16), m.group(2).strip()) for m in ins.finditer(b)]
    lo, hi = 0, rows[-1][0]
    if a.loop:
        best = (0, 0, 0)
        for addr, txt in rows:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", txt)
            if m and int(m.group(1), 16) < addr and addr - int(m.group(1), 16) > best[0]:
                best = (addr - int(m.group(1), 16), int(m.group(1), 16), addr)
        _, lo, hi = best
    sel = [(ad, t) for ad, t in rows if lo <= ad <= hi]
    c = collections.Counter()
    for _, t in sel:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        c[t.split()[0].split(".")[0]] += 1
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
    print(f"{name}: {len(sel)} instructions in [{lo:#x}, {hi:#x}], fp64 {fp64}")
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(28)))
    if a.dump:
        for ad, t in sel:
            print(f"  {ad:05x}  {t}")
