"""Stereo/mono decode kernel variants: parity against the oracle on adversarial streams, then device-resident timing of
the headline shape.  Variants are chosen by the environment (QOA_B200_TMAP = 2 / 1 / 0: both loops / general loop only /
no tensor stores), so run this once per variant:  QOA_B200_TMAP=1 python tools/dec_variants.py [streams] [steps] [ch]"""
import os, sys, time, statistics, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qoaudio_b200 as q
from oracle import binding as ob

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
CH = int(sys.argv[3]) if len(sys.argv) > 3 else 2
RATE, N = 44100, 2646000
eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("QOA_B200_")) or "default"

def enc_size(ch, n):
    return 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch

size = enc_size(CH, N); stride = (size + 15) // 16 * 16; out_per = N * CH
slot = (out_per + 5120 * CH - 1) // (5120 * CH) * (5120 * CH)
qoa = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
pcm = torch.zeros(S * slot, dtype=torch.int16, device="cuda")
eng.gen_qoa_dev(qoa.data_ptr(), stride, S, 0, CH, RATE, N, 1, stream)
in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
lens = np.full(S, size, dtype=np.uint64)
out_off = np.arange(S, dtype=np.uint64) * np.uint64(slot)
caps = np.full(S, out_per, dtype=np.uint64)

def step(ev=None):
    plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, lens, q._native.DECODE_ALL, stream)
    if ev: ev[0].record()
    st = plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
    if ev: ev[1].record()
    plan.close()
    return st

for _ in range(3): st = step()
torch.cuda.synchronize()
assert (st == 0).all()
# parity: random headers' LMS + random payload (weights and predictions wrap all over the i32 range)
bad = 0
for s in sorted(set([0, 1, S // 2, S - 1])):
    blob = qoa[s * stride:s * stride + size].cpu().numpy().tobytes()
    got = pcm[s * slot:s * slot + out_per].cpu().numpy()
    ref = ob.decode_all(blob)[1]
    if not np.array_equal(got, ref):
        bad += 1
        i = int(np.argmax(got != ref))
        print(f"[{tag}] MISMATCH stream {s} first at sample {i} (frame {i // (5120 * CH)}, in-frame {i % (5120 * CH)}): got {got[i:i+6]} ref {ref[i:i+6]}")
print(f"[{tag}] parity vs oracle on 4 random-payload streams: {'OK' if not bad else 'FAILED'}")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                       stdout=subprocess.PIPE, text=True)
time.sleep(0.2)
for k in range(K): step(ev[k])
torch.cuda.synchronize()
smi.terminate()
rows = [r.split(",") for r in smi.stdout.read().strip().splitlines() if "," in r]
clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
ms = [a.elapsed_time(b) for a, b in ev]
alg = S * (size + 2 * out_per)
m = statistics.mean(ms)
print(f"[{tag}] ch={CH} streams={S} steps={K}: kernel mean {m:.3f} ms  min {min(ms):.3f}  last5 {statistics.mean(ms[-5:]):.3f}  "
      f"-> {alg / m / 1e6:.0f} GB/s = {alg / m / 1e6 / 6549.1:.3f} of 6549; {S * out_per / m / 1e9:.3f} Tsamples/s; "
      f"clk median {statistics.median(clk) if clk else None} min {min(clk) if clk else None} MHz, power max {max(pw) if pw else None} W")
