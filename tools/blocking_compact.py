"""Blocking host call with the compact wire format (pinned buffers): python tools/blocking_compact.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import ctypes as C
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
from ik_b200 import _capi as capi


def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * n).from_address(capi.lib.ikb_host_alloc(max(n, 1))), dtype=dtype).reshape(shape)


B = 65536
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model(); names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device="cuda:0"), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses, qstar)
h_c = pinned((pb.compact_target_size, B), np.float64); h_c[:] = pb.compact_targets(tg).T
q0 = pinned((23,), np.float64); q0[:] = W.standing_configuration(m, W.CASSIE_STANDING)
out = {"q": pinned((23, B), np.float64), "success": pinned((B,), np.uint8), "iters": pinned((B,), np.int32), "resid": pinned((B,), np.float64)}
for env in ({}, {"IKB_HOST_PIPELINE": "0"}):
    os.environ.pop("IKB_HOST_PIPELINE", None); os.environ.update(env)
    for _ in range(3):
        ik.dls_batch_host(pb, q0, h_c, None, "f64", "soa", out, compact=True, outputs=("q", "success"))
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ik.dls_batch_host(pb, q0, h_c, None, "f64", "soa", out, compact=True, outputs=("q", "success")); ts.append((time.perf_counter() - t0) * 1e3)
    print("blocking host call, compact wire format, B=%d %s: median %.3f ms (min %.3f)" % (B, env or "(sliced pipeline)", np.median(ts), min(ts)))
