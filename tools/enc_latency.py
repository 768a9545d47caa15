"""Device-resident encode time against the number of streams (development aid: where does the latency-mode kernel stop
paying?).  Run once per variant: QOA_B200_ENC_FAST=0 / 1 python tools/enc_latency.py [seconds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qoaudio_b200 as q

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
ch, rate = int(os.environ.get("CH", "2")), 44100
n = int(secs * rate)
size = 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch
for streams in [int(v) for v in os.environ.get("STREAMS", "1,8,64,296,592,1184,2368,4096").split(",")]:
    pcm = torch.empty(streams * n * ch, dtype=torch.int16, device="cuda")
    out = torch.empty(streams * (size + 16), dtype=torch.uint8, device="cuda")
    eng.gen_pcm_dev(pcm.data_ptr(), streams, 0, ch, n, 0, 1, "music", stream)
    descs = [(ch, rate, n)] * streams
    pcm_off = [i * n * ch for i in range(streams)]
    out_off = [i * (size + 16) for i in range(streams)]
    best = 1e9
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st, _ = eng.encode_all_batch_dev(descs, pcm.data_ptr(), pcm_off, [n * ch] * streams, out.data_ptr(), out_off, [size] * streams, stream)
        e1.record()
        torch.cuda.synchronize()
        assert not any(st)
        best = min(best, e0.elapsed_time(e1))
    print(f"streams={streams:5d} warps={streams:5d}: {best:9.3f} ms  {streams * n * ch / best / 1e3:9.1f} Msamples/s")
