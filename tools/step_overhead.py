"""Where the host time of one decode step (index pre-pass + decode launch) goes (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import qoaudio_b200 as q
N = q._native
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = q.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
ch, rate, n = 2, 44100, 2646000
size = 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch
stride = (size + 15) // 16 * 16
per = n * ch
slot = (per + 10239) // 10240 * 10240
qoa = torch.empty(S * stride, dtype=torch.uint8, device="cuda")
pcm = torch.empty(S * slot, dtype=torch.int16, device="cuda")
eng.gen_qoa_dev(qoa.data_ptr(), stride, S, 0, ch, rate, n, 1, stream)
in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
lens = np.full(S, size, dtype=np.uint64)
out_off = np.arange(S, dtype=np.uint64) * np.uint64(slot)
caps = np.full(S, per, dtype=np.uint64)
status = np.zeros(S, dtype=np.int32)
L = N.lib()
T = {"create": 0.0, "run": 0.0, "destroy": 0.0, "py_plan": 0.0}
K = 30
for it in range(K + 3):
    if it == 3:
        torch.cuda.synchronize(); T = {k: 0.0 for k in T}; t_all = time.perf_counter()
    t0 = time.perf_counter()
    h = C.c_void_p()
    rc = L.qoa_b200_decode_plan_create_dev(eng.handle, C.c_void_p(qoa.data_ptr()), C.c_void_p(in_off.ctypes.data), C.c_void_p(lens.ctypes.data), S, 1, C.c_void_p(stream), C.byref(h))
    t1 = time.perf_counter()
    rc |= L.qoa_b200_decode_plan_run_dev(h, C.c_void_p(pcm.data_ptr()), C.c_void_p(out_off.ctypes.data), C.c_void_p(caps.ctypes.data), C.c_void_p(status.ctypes.data), C.c_void_p(stream))
    t2 = time.perf_counter()
    L.qoa_b200_decode_plan_destroy(h)
    t3 = time.perf_counter()
    assert rc == 0
    T["create"] += t1 - t0; T["run"] += t2 - t1; T["destroy"] += t3 - t2
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / K
print(f"streams={S}: wall per step {wall * 1e3:.3f} ms; host: create {T['create'] / K * 1e3:.3f} ms (includes the index kernels + 2 syncs), "
      f"run_dev {T['run'] / K * 1e3:.3f} ms, destroy {T['destroy'] / K * 1e3:.3f} ms")
# the same through the python Engine wrappers (what bench.py calls)
t_all = time.perf_counter()
for it in range(K):
    plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, lens, 1, stream)
    st = plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
    plan.close()
torch.cuda.synchronize()
print(f"through qoaudio_b200.Engine: wall per step {(time.perf_counter() - t_all) / K * 1e3:.3f} ms")
