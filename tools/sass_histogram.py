"""Per-kernel SASS opcode histogram of a built library (cuobjdump -sass), for profiles/.
Shows which machine idioms each kernel really uses: HMMA (mma.sync tensor-core products), FHFMA
(fp16 x fp16 + fp32), FADD2/FFMA2/FMUL2 (packed fp32x2, sm_100), LDGSTS (cp.async), UTMALDG /
UTCHMMA / LDTM (TMA, tcgen05, tensor memory), LDG/LDS/STS/SHFL counts.
Usage: python tools/sass_histogram.py approximatenn_b200/libann_b200_f32.so > profiles/r2_sass_f32.txt"""
import collections, re, subprocess, sys

path = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
special = ("HMMA", "FHFMA", "FADD2", "FFMA2", "FMUL2", "HFMA2", "HMNMX2", "LDGSTS", "UTMALDG", "UTMASTG", "UTCHMMA",
           "LDTM", "STTM", "LDG", "STG", "LDS", "STS", "LDC", "SHFL", "VOTE", "ATOMS", "ATOMG", "RED", "LDL", "STL", "MUFU", "BAR")
print("library:", path)
tot = collections.Counter()
for k, h in hist.items():
    n = sum(h.values())
    tot.update(h)
    print(f"\n{k}   ({n} instructions)")
    print("   notable: " + ", ".join(f"{o} {h[o]}" for o in special if h[o]))
    print("   top:     " + ", ".join(f"{o} {c}" for o, c in h.most_common(10)))
print("\nwhole library: " + ", ".join(f"{o} {tot[o]}" for o in special if tot[o]))
