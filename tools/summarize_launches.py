"""Turns an ncu launch list (--metrics gpu__time_duration.sum --csv) into the per-kernel summary committed under
profiles/ (run on the CPU side).  Usage: summarize_launches.py launches.csv "<command that was profiled>" """
import csv
import re
import sys
from collections import OrderedDict


def main(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    name_i, val_i, unit_i = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[name_i]).replace("qoa::", "").strip()
        ms = float(r[val_i].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[unit_i], 1e-6)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] = max(a[2], ms)
    total = sum(a[1] for a in agg.values())
    print(f"ncu launch list summary (command: {command})")
    print("per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
    for name, (n, ms, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:58s} launches={n:3d} mean={ms / n:9.3f} ms  max={mx:9.3f} ms  total={ms:9.2f} ms  share={100 * ms / total:5.1f}%")
    print(f"\n{sum(a[0] for a in agg.values())} launches, {total:.2f} ms of kernel time")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
