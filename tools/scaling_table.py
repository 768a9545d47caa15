"""Builds the DESIGN.md scaling table from profiles/<round>_bench_n{1,2,4,8}.json (run on the CPU side)."""
import json
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = []
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"profiles/{R}_bench_n{n}.json").read().strip().splitlines()[-1])
    except OSError:
        continue
    e, s = d["encode"], d["scaling_strong"]
    rows.append((n, d["value"], d["roofline"]["frac"], s["decode"]["value"], d["e2e"]["value"], d["e2e"]["host_ceiling_msamples_s"],
                 e["value"], s["encode"]["value"], e["e2e"]["value"], d["clocks"]["sm_mhz"]))
print("| GPUs | decode, weak (8192 streams/GPU) | roofline frac per GPU | decode, strong (8192 streams total) | e2e decode | bare-copy ceiling of the host | encode, weak (4096 streams/GPU) | encode, strong (4096 streams total) | e2e encode | median SM MHz |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]:,.0f} | {r[2]:.3f} | {r[3]:,.0f} | {r[4]:,.0f} | {r[5]:,.0f} | {r[6]:,.0f} | {r[7]:,.0f} | {r[8]:,.0f} | {r[9]:.0f} |")
b = rows[0]
print()
print("Speed-up over one GPU: " + "; ".join(
    f"N = {r[0]}: weak decode {r[1] / b[1]:.2f}×, strong decode {r[3] / b[3]:.2f}×, weak encode {r[6] / b[6]:.2f}×, strong encode {r[7] / b[7]:.2f}×, e2e decode {r[4] / b[4]:.2f}×"
    for r in rows[1:]))
