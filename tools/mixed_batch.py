"""Device-resident decode throughput for a mixed batch (BASELINE.json configs[4]: mono / stereo / 5.1 streams of 1-120 s
shuffled into one arena).  Development aid: prints the batch rate and what each class reaches alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qoaudio_b200 as q

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(5)
chs = rng.choice([1, 2, 6], size=n_streams, p=[0.3, 0.5, 0.2])
secs = rng.integers(1, 121, size=n_streams)
ns = secs * 44100
sizes = np.array([8 + ((n + 5119) // 5120) * (8 + 16 * c) + ((n + 19) // 20) * 8 * c for n, c in zip(ns, chs)], dtype=np.uint64)
strides = (sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16)
in_off = np.concatenate([[0], np.cumsum(strides)[:-1]]).astype(np.uint64)
caps = (ns * chs).astype(np.uint64)
slots = (caps + np.uint64(7)) // np.uint64(8) * np.uint64(8)
out_off = np.concatenate([[0], np.cumsum(slots)[:-1]]).astype(np.uint64)
qoa = torch.empty(int(strides.sum()), dtype=torch.uint8, device="cuda")
pcm = torch.empty(int(slots.sum()) + 8, dtype=torch.int16, device="cuda")
for i in range(n_streams):
    eng.gen_qoa_dev(qoa.data_ptr() + int(in_off[i]), int(strides[i]), 1, i, int(chs[i]), 44100, int(ns[i]), 1, stream)
torch.cuda.synchronize()


def run(sel, label):
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan = eng.decode_plan_dev(qoa.data_ptr(), in_off[sel], sizes[sel], 1, stream)
        e0.record()
        plan.run_dev(pcm.data_ptr(), out_off[sel], caps[sel], stream)
        e1.record()
        torch.cuda.synchronize()
        plan.close()
        if it:
            best = min(best, e0.elapsed_time(e1))
    samples = int(caps[sel].sum())
    print(f"{label:10s} streams={int(sel.sum()):5d} samples={samples / 1e9:7.3f} G  {best:8.3f} ms  {samples / best / 1e6:8.1f} Gsamples/s")
    return best


total = run(np.ones(n_streams, dtype=bool), "mixed")
parts = sum(run(chs == c, f"{c} ch only") for c in (1, 2, 6))
print(f"sum of classes {parts:.3f} ms vs mixed {total:.3f} ms")
