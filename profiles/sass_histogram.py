"""SASS opcode histogram per kernel of the shipped library (cuobjdump -sass): which kernels really carry tcgen05 / TMA / bulk
copies.   python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "regcn_b200", "libregcn_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
keep = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA", "MUFU", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "REDUX"]
kern, hist, total = None, {}, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for k in keep:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
print("SASS opcode histogram per kernel of regcn_b200/libregcn_b200.so (cuobjdump -sass, sm_100a; profiles/sass_histogram.py).")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.\n")
for k in sorted(total, key=lambda k: -total[k]):
    name = demangle(k).split("(")[0]
    print(f"{total[k]:6d} instr  {name}")
    print("        " + str({o: hist[k][o] for o in keep if hist[k][o]}))
