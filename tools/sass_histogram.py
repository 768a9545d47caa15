"""cuobjdump -sass libqoa_b200.so | python tools/sass_histogram.py : per-kernel opcode histogram (the evidence that the
TMA / cp.async / I2IP paths are in the built code: UTMASTG, UBLKCP, LDGSTS, I2IP, ...)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

kernels = OrderedDict()
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
names = list(kernels)
try:
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
except OSError:
    dem = names
for name, d in zip(names, dem):
    c = kernels[name]
    total = sum(c.values())
    if total < 40:
        continue
    print(f"== {d[:150]}\n   {total} instructions")
    key = [k for k in c if any(t in k for t in ("UTMASTG", "UBLKCP", "LDGSTS", "I2IP", "UTMACMDFLUSH", "SHFL", "REDUX", "VIADDMNMX", "VIMNMX", "DFMA"))]
    print("   of interest: " + ", ".join(f"{k} x{c[k]}" for k in sorted(key)))
    print("   top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
