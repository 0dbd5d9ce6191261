"""Per-kernel SASS instruction-class counts of the CUDA library (evidence that tcgen05 / TMA / TMEM / mbarrier / vector
red / peer loads are what the kernels actually execute):  python tools/sass_summary.py > profiles/sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dgl_b200", "lib", "libdglb200_cuda.so")
CLASSES = [("UTCHMMA", r"UTC[A-Z]*MMA"), ("LDTM", r"LDTM"), ("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"), ("UBLKCP(bulk copy)", r"UBLKCP"), ("UTCBAR", r"UTCBAR"),
           ("MULTIMEM", r"MULTIMEM|\.MMEM|LDGMC|REDGMC|STGMC"),
           ("SYNCS(mbarrier)", r"SYNCS"), ("LDGSTS(cp.async)", r"LDGSTS"), ("HMMA(legacy)", r"\bHMMA"), ("REDG", r"\bREDG"),
           ("ATOMG", r"\bATOMG"), ("ATOMS", r"\bATOMS"), ("SHFL", r"\bSHFL"), ("LDG.128", r"LDG\.E\.128|LDG\.E\.[A-Z.]*128"),
           ("STG.128", r"STG\.E\.128|STG\.E\.[A-Z.]*128"), ("MEMBAR/FENCE", r"MEMBAR|FENCE"), ("MUFU.EX2", r"MUFU\.EX2")]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = {}
cur = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for name, pat in CLASSES:
        if re.search(pat, line):
            counts[cur][name] += 1
names = list(counts)
try:
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
except Exception:  # noqa: BLE001
    dem = names
print(f"# {os.path.basename(so)}: {len(names)} kernels; columns = instruction classes present in the SASS of each kernel")
for n, d in zip(names, dem):
    c = counts[n]
    short = re.sub(r"\(.*", "", d)
    print(f"{short[:90]:90s} " + " ".join(f"{k}={v}" for k, v in c.items() if v))
