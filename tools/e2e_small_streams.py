"""End-to-end decode of many small streams with pinned, NON-uniformly spaced host buffers (development aid: where is
the crossover between the copy kernel and one cudaMemcpyAsync per buffer?).  QOA_B200_COPY_KERNEL_MAX_KB=<n> python ..."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qoaudio_b200 as q
N = q._native
eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
ch, rate, S = 2, 44100, int(os.environ.get('S', '4096'))
for n in [int(v) for v in os.environ.get('NS', '4096,8192,16384,32768,65536,131072').split(',')]:          # samples per channel -> 16 KB .. 512 KB of PCM per stream
    size = 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch
    stride = (size + 15) // 16 * 16
    d = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
    eng.gen_qoa_dev(d.data_ptr(), stride, S, 0, ch, rate, n, 1, stream)
    torch.cuda.synchronize()
    gaps = np.random.default_rng(1).integers(1, 9, S) * 16             # irregular spacing: no strided runs
    in_pos = np.concatenate([[0], np.cumsum(stride + gaps)[:-1]])
    out_pos = np.concatenate([[0], np.cumsum(n * ch * 2 + gaps)[:-1]])
    h_in = torch.zeros(int(in_pos[-1] + stride + 64), dtype=torch.uint8).pin_memory()
    h_out = torch.empty(int(out_pos[-1] + n * ch * 2 + 64), dtype=torch.uint8).pin_memory()
    src = d.cpu().numpy()
    hin = h_in.numpy()
    for i in range(S):
        hin[in_pos[i]:in_pos[i] + size] = src[i * stride:i * stride + size]
    qp = (C.c_void_p * S)(*[h_in.data_ptr() + int(p) for p in in_pos])
    ql = (C.c_size_t * S)(*([size] * S))
    op_ = (C.c_void_p * S)(*[h_out.data_ptr() + int(p) for p in out_pos])
    oc = (C.c_size_t * S)(*([n * ch] * S))
    info = (N.StreamInfo * S)()
    st = (C.c_int32 * S)()
    def step():
        rc = N.lib().qoa_b200_decode_all_batch(eng.handle, qp, ql, S, 1, op_, oc, info, st)
        assert rc == 0
    step(); step()
    t0 = time.perf_counter()
    for _ in range(3): step()
    dt = (time.perf_counter() - t0) / 3
    print(f"PCM per stream {n * ch * 2 // 1024:4d} KB: {dt * 1e3:8.2f} ms  {S * n * ch / dt / 1e6:8.0f} Msamples/s  D2H {S * n * ch * 2 / dt / 1e9:5.1f} GB/s")
