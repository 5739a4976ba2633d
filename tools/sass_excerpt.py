"""Write profiles/<tag>_sass_excerpt.txt: the instructions of the built library that show the quad kernel's design.
usage: sass_excerpt.py <tag>"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
lib = os.path.join(ROOT, "rte-ecckd_b200", "lib", "libecckd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); fn[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line): fn[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
name = next(n for n in fn if "quad_kernelIdLb0ELi1ELb" in n)            # <double, LW, NV = 1, ...>: the production longwave kernel
dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
L = fn[name]
pats = ["UTMASTG", "UTMACMDFLUSH", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "USETMAXREG", "FENCE.VIEW.ASYNC", "LDG.E.128.CONSTANT", "DFMA", "DMUL",
        "F2F.F64.F32", "STS.64", "LDS.128", "LDS.64", "STG.E.EF.64", "LDL", "STL"]
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.txt"), "w") as f:
    f.write(f"cuobjdump -sass rte-ecckd_b200/lib/libecckd_b200.so, function\n  {dem[:200]}\n(the production longwave kernel; tools/sass_excerpt.py)\n")
    f.write("instructions that show the design: TMA tensor stores (UTMASTG), bulk-group commit (UTMACMDFLUSH), mbarrier operations (SYNCS.*),\n"
            "register re-allocation between warp roles (USETMAXREG), the proxy fence (FENCE.VIEW.ASYNC), 16-byte predicated cache refills\n"
            "(LDG.E.128.CONSTANT); LDL / STL = local-memory (spill) traffic, which the kernel must not have in its loops\n\ncounts (whole function, all warp roles):\n")
    for p in pats:
        rx = re.compile(r"\b" + re.escape(p))
        f.write("  %-22s %d\n" % (p, sum(1 for l in L if rx.search(l))))
    f.write("\nexcerpt (first occurrence of each):\n")
    for p in pats[:6]:
        for l in L:
            if re.search(r"\b" + re.escape(p), l): f.write(l + "\n"); break
    f.write("\nthe refills of the register cache (four columns per warp, every load predicated on its column's key), staged behind the\narithmetic that last reads each table:\n")
    cuts = [i for i, l in enumerate(L) if "USETMAXREG" in l]
    n = 0
    for l in L[cuts[-1]:]:
        if "LDG.E.128.CONSTANT" in l and l.strip().split()[1].startswith("@"):
            f.write(l + "\n"); n += 1
            if n >= 28: break
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.txt")).read()[:1800])
