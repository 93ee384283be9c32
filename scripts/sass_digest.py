"""Per-kernel SASS digest of libdadd_b200.so: counts of the Blackwell-native mnemonics (tcgen05 MMA, TMA loads / stores,
TMEM loads / stores, MUFU.EX2) in every kernel that has any, from `cuobjdump -sass`.

    python scripts/sass_digest.py > profiles/r02_sass_digest.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "progressive_stable_diffusion_b200", "libdadd_b200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "MUFU.EX2", "HMMA", "LDGSTS", "SYNCS"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name:
        for k in KEYS:
            if re.search(r"\b" + re.escape(k), line):
                counts[name][k] += 1
print(f"SASS digest of {os.path.relpath(LIB, ROOT)} (sm_100a); columns: " + " ".join(KEYS))
tot = collections.Counter()
for n, c in counts.items():
    tot.update(c)
    if not any(c[k] for k in KEYS[:7] + ["HMMA"]):
        continue
    d = re.sub(r"\(.*", "", demangle(n))
    d = re.sub(r"^void ", "", d)
    print(f"{d[:110]:110s} " + " ".join(f"{c[k]:5d}" for k in KEYS))
print(f"{'TOTAL (all ' + str(len(counts)) + ' kernels)':110s} " + " ".join(f"{tot[k]:5d}" for k in KEYS))
