"""Host link under load: every visible GPU copies the e2e step's 30.9 MB (H2D) and 12.9 MB (D2H) from / to its own pinned
buffers -- first one GPU at a time, then all at once (one process per GPU, barrier via files)."""
import glob, os, subprocess, sys, time

if len(sys.argv) > 1:
    import ctypes as C
    import torch
    sys.path.insert(0, os.getcwd())
    from ik_b200 import _capi as capi
    g, n, tag = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    torch.cuda.set_device(g)
    dev = torch.device("cuda", g)
    def pinned(nbytes):
        p = capi.lib.ikb_host_alloc(nbytes)
        return (C.c_char * nbytes).from_address(p), p
    nin, nout = 30932992, 12910592
    hin, pin_ = pinned(nin); hout, pout = pinned(nout)
    din = torch.empty(nin, dtype=torch.uint8, device=dev); dout = torch.empty(nout, dtype=torch.uint8, device=dev)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d(): assert rt.cudaMemcpyAsync(din.data_ptr(), pin_, nin, 1, s1.cuda_stream) == 0
    def d2h(): assert rt.cudaMemcpyAsync(pout, dout.data_ptr(), nout, 2, s2.cuda_stream) == 0
    for _ in range(3): h2d(); d2h()
    torch.cuda.synchronize()
    open("/tmp/probe_ready_%s_%d" % (tag, g), "w").close()
    t_wait = time.time()
    while len(glob.glob("/tmp/probe_ready_%s_*" % tag)) < n and time.time() - t_wait < 20: time.sleep(0.001)
    res = []
    for mode in ("h2d", "d2h", "both"):
        t0 = time.perf_counter()
        for _ in range(40):
            if mode != "d2h": h2d()
            if mode != "h2d": d2h()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 40
        res.append("%s %.1f GB/s" % (mode, ((nin if mode != "d2h" else 0) + (nout if mode != "h2d" else 0)) / dt / 1e9))
    print("gpu %d of %d active: %s" % (g, n, ", ".join(res)), flush=True)
    sys.exit(0)

import torch
ng = torch.cuda.device_count()
print(subprocess.run("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'; nvidia-smi topo -m | head -12; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
for group in ([[g] for g in range(min(ng, 2))] + ([list(range(2))] if ng >= 2 else []) + ([list(range(4))] if ng >= 4 else []) + ([list(range(ng))] if ng > 4 else [])):
    tag = "x".join(map(str, group)) + "_%d" % os.getpid()
    ps = [subprocess.Popen([sys.executable, __file__, str(g), str(len(group)), tag]) for g in group]
    for p in ps: p.wait()
