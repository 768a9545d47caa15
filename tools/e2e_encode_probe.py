"""Times qoa_b200_encode_all_batch with pinned host buffers for a few batch shapes (development aid)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qoaudio_b200 as q
N = q._native
eng = q.Engine(0)
def run(ES, EF, ch=2, rate=48000):
    per = EF * 5120 * ch
    osz = 8 + EF * (8 + 16 * ch) + EF * 256 * 8 * ch
    h_pcm = torch.randint(-3000, 3000, (ES * per,), dtype=torch.int16).pin_memory()
    h_enc = torch.empty(ES * osz, dtype=torch.uint8).pin_memory()
    descs = (N.Desc * ES)(*[N.Desc(ch, rate, EF * 5120)] * ES)
    pp = (C.c_void_p * ES)(*[h_pcm.data_ptr() + 2 * i * per for i in range(ES)])
    pl = (C.c_size_t * ES)(*([per] * ES))
    op_ = (C.c_void_p * ES)(*[h_enc.data_ptr() + i * osz for i in range(ES)])
    oc = (C.c_size_t * ES)(*([osz] * ES)); ol = (C.c_size_t * ES)(); st = (C.c_int32 * ES)()
    def step():
        rc = N.lib().qoa_b200_encode_all_batch(eng.handle, descs, pp, pl, op_, oc, ol, st, ES); assert rc == 0
    step(); step()
    t0 = time.perf_counter()
    for _ in range(3): step()
    dt = (time.perf_counter() - t0) / 3
    print(f"streams={ES} frames={EF}: {dt*1e3:8.2f} ms  {ES*per/dt/1e6:9.0f} Msamples/s  (PCM {ES*per*2/1e6:.0f} MB)")
for shape in ((4096, 8), (4096, 24), (1024, 96), (64, 1536)):
    run(*shape)
