"""Per-kernel SASS instruction counts of libbfl.so (cuobjdump -sass): what the hot kernels are made of.
    python tools/sass_counts.py > profiles/sass_r2.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bayesflare_b200", "libbfl.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", name)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
cols = ["DFMA", "DMUL", "DADD", "FFMA", "FFMA2", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMG", "SHFL", "BAR", "F2F", "F2I", "I2F"]
print("SASS of bayesflare_b200/libbfl.so (%s), static instruction counts per kernel; LDGSTS = cp.async; no tensor-core or TMA"
      " opcodes by design (FP64 vector pipe, north_star)" % arch)
print("%-58s %6s " % ("kernel", "total") + " ".join("%6s" % c for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) < 40:
        continue
    print("%-58s %6d " % (k[:58], sum(c.values())) + " ".join("%6d" % c[x] for x in cols))
    tot.update(c)
print("%-58s %6d " % ("ALL KERNELS", sum(tot.values())) + " ".join("%6d" % tot[x] for x in cols))
absent = [x for x in ("UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCHMMA", "UTCQMMA", "HMMA", "DMMA", "HGMMA") if tot[x] == 0]
print("opcodes absent: " + ", ".join(absent))
