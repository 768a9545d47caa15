"""Host <-> device copy ceiling of this box at 1..N ranks (development aid; the e2e legs of bench.py are bound by it).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_multi.py

Every rank moves the bytes of one bench e2e decode step (128 streams: 0.27 GB in, 1.35 GB out) with bare cudaMemcpyAsync
on two streams, all ranks at once, for three kinds of host memory: cudaHostAlloc (what torch.pin_memory gives),
cudaHostAlloc write-combined (H2D source only) and cudaHostRegister-ed pageable memory.  Rank 0 prints the aggregate
GB/s (bytes of all ranks / slowest rank's time) for H2D alone, D2H alone and both directions together, and the same
bytes through qoa_b200_decode_all_batch for comparison."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
import qoaudio_b200 as q
from qoaudio_b200.dist import Group

grp = Group(device=torch.device("cuda", local))
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
IN_BYTES, OUT_BYTES = 128 * 2137488, 128 * 10584000
REPS = 8


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p.value


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        fn()
    torch.cuda.synchronize()
    return grp.max(time.perf_counter() - t0) / REPS


d_in = torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT_BYTES, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
H2D, D2H = 1, 2
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


def report(name, h_in, h_out):
    def h2d():
        rt.cudaMemcpyAsync(d_in.data_ptr(), h_in, IN_BYTES, H2D, s1.cuda_stream)

    def d2h():
        rt.cudaMemcpyAsync(h_out, d_out.data_ptr(), OUT_BYTES, D2H, s2.cuda_stream)

    t_in, t_out = timed(h2d), timed(d2h)
    t_both = timed(lambda: (h2d(), d2h()))
    if rank == 0:
        print(f"{name:34s} ranks={world}: H2D {IN_BYTES * world / t_in / 1e9:6.1f} GB/s   D2H {OUT_BYTES * world / t_out / 1e9:6.1f} GB/s   "
              f"both {(IN_BYTES + OUT_BYTES) * world / t_both / 1e9:6.1f} GB/s ({t_both * 1e3:.1f} ms per step)", flush=True)
    return t_both


a = host_alloc(IN_BYTES, 1)   # cudaHostAllocPortable
b = host_alloc(OUT_BYTES, 1)
t_bare = report("cudaHostAlloc (portable)", a, b)
wc = host_alloc(IN_BYTES, 1 | 4)  # + cudaHostAllocWriteCombined
report("write-combined H2D source", wc, b)
pg_in = np.empty(IN_BYTES, dtype=np.uint8)
pg_out = np.empty(OUT_BYTES, dtype=np.uint8)
pg_in[:] = 1
pg_out[:] = 1
assert rt.cudaHostRegister(C.c_void_p(pg_in.ctypes.data), C.c_size_t(IN_BYTES), C.c_uint(1)) == 0
assert rt.cudaHostRegister(C.c_void_p(pg_out.ctypes.data), C.c_size_t(OUT_BYTES), C.c_uint(1)) == 0
report("cudaHostRegister (pageable, numpy)", pg_in.ctypes.data, pg_out.ctypes.data)

# the same bytes through the library's one-call host path (128 streams of the headline shape, pinned buffers)
eng = q.Engine(local)
size, out_per = 2137488, 5292000
stride = (size + 15) // 16 * 16
qoa = torch.empty(128 * stride, dtype=torch.uint8, device="cuda")
eng.gen_qoa_dev(qoa.data_ptr(), stride, 128, rank * 128, 2, 44100, 2646000, 1, torch.cuda.current_stream().cuda_stream)
h_in = qoa.cpu().pin_memory()
h_out = torch.empty(128 * out_per, dtype=torch.int16).pin_memory()
N = q._native
in_ptrs = (C.c_void_p * 128)(*[h_in.data_ptr() + i * stride for i in range(128)])
in_lens = (C.c_size_t * 128)(*([size] * 128))
out_ptrs = (C.c_void_p * 128)(*[h_out.data_ptr() + 2 * i * out_per for i in range(128)])
out_caps = (C.c_size_t * 128)(*([out_per] * 128))
stat = (C.c_int32 * 128)()


def lib_step():
    rc = N.lib().qoa_b200_decode_all_batch(eng.handle, in_ptrs, in_lens, 128, N.DECODE_ALL, out_ptrs, out_caps, None, stat)
    assert rc == 0


t_lib = timed(lib_step)
if rank == 0:
    print(f"{'qoa_b200_decode_all_batch':34s} ranks={world}: {(IN_BYTES + OUT_BYTES) * world / t_lib / 1e9:6.1f} GB/s ({t_lib * 1e3:.1f} ms per step) "
          f"= {t_bare / t_lib:.2f} of the bare-copy rate; {128 * out_per * world / t_lib / 1e9:.1f} Gsamples/s", flush=True)
grp.close()
