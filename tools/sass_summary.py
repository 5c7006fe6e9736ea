#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): DMMA (fp64 mma.sync),
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG/UBLKCP (TMA), LDGSTS (cp.async), plus the
peer-memory evidence of the multi-GPU pull kernel (system-scope loads/stores).  Runs without a GPU:
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rust-la_b200", "libla_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
PAT = [("DMMA", r"\bDMMA"), ("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
       ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"),
       ("LD.SYS", r"\bLDG?\.E[\w.]*\.SYS"), ("ST.SYS", r"\bSTG?\.E[\w.]*\.SYS"), ("HMMA", r"\bHMMA"), ("DFMA", r"\bDFMA"),
       ("FFMA", r"\bFFMA")]
arch = set(re.findall(r"arch = (sm_\w+)", txt))
kern = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            kern[cur][name] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}   arch: {', '.join(sorted(arch))}")
print(f"# {'kernel':<70s} " + " ".join(f"{n:>8s}" for n, _ in PAT))
tot = collections.Counter()
for k, c in kern.items():
    name = demangle(k)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name).replace("la::<unnamed>::", "").replace("void ", "")[:70]
    print(f"  {name:<70s} " + " ".join(f"{c[n]:8d}" for n, _ in PAT))
    tot.update(c)
print(f"  {'TOTAL':<70s} " + " ".join(f"{tot[n]:8d}" for n, _ in PAT))
