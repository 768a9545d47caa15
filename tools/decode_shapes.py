"""Device-resident decode throughput for different channel counts (development aid): Msamples/s per shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qoaudio_b200 as q

eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
shapes = [(1, 4096, 5120 * 100), (2, 4096, 5120 * 100), (3, 2048, 5120 * 100), (4, 2048, 5120 * 100), (6, 1024, 5120 * 100),
          (8, 1024, 5120 * 100)]
if len(sys.argv) > 1:  # ch:streams:samples ...
    shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]]
for shape in shapes:
    ch, n_streams, n = shape[:3]
    ostride = shape[3] if len(shape) > 3 else n * ch
    size = 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch
    stride = (size + 15) // 16 * 16
    qoa = torch.empty(n_streams * stride, dtype=torch.uint8, device="cuda")
    pcm = torch.empty(n_streams * ostride, dtype=torch.int16, device="cuda")
    eng.gen_qoa_dev(qoa.data_ptr(), stride, n_streams, 0, ch, 44100, n, 1, stream)
    in_off = np.arange(n_streams, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(n_streams, size, dtype=np.uint64)
    out_off = np.arange(n_streams, dtype=np.uint64) * np.uint64(ostride)
    caps = np.full(n_streams, n * ch, dtype=np.uint64)
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, lens, 1, stream)
        e0.record()
        plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
        e1.record()
        torch.cuda.synchronize()
        plan.close()
        if it:
            best = min(best, e0.elapsed_time(e1))
    samples = n_streams * n * ch
    print(f"ch={ch} streams={n_streams} n={n} ostride={ostride}: {best:8.3f} ms  {samples / best / 1e3:10.0f} Msamples/s  "
          f"{(n_streams * size + 2 * samples) / best / 1e6:7.0f} GB/s algorithmic")
    del qoa, pcm
