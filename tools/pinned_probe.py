"""H2D rate of pinned buffers allocated the way bench.py does (ikb_host_alloc), many buffers, in a Python process."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ik_b200 import _capi as capi
import bench
torch.cuda.init()
rt = C.CDLL("libcudart.so")
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
n = 30932992
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
s = torch.cuda.Stream()
bufs = [bench.pinned_array((n,), np.uint8) for _ in range(10)]
for b in bufs:
    b[:] = 1
for i, b in enumerate(bufs):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for _ in range(5):
            rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), C.c_void_p(b.ctypes.data), n, 1, C.c_void_p(s.cuda_stream))
        e1.record()
    torch.cuda.synchronize()
    print("buffer %d  h2d %.1f GB/s" % (i, n * 5 / 1e6 / e0.elapsed_time(e1)))
