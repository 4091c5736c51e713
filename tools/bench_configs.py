#!/usr/bin/env python3
"""Supplementary measurements for the BASELINE.json configs that are NOT the bench.py headline (they are parity-test
cases; these numbers go into DESIGN.md only): Cassie 4,096 FP64, humanoid 262,144, manipulator 1,048,576 per GPU.

    python tools/bench_configs.py [--scale 1.0]   ->  one JSON line per config (device-resident, CUDA-event timed)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (smoke runs)")
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    import torch

    import ik_b200 as ik
    from ik_b200 import workloads as W

    dev = torch.device("cuda:0")
    demo = ik.dls_parameters(max_iterations=200, step_length=0.1, damping=0.1)  # cassie.cpp:107-109 (secondary set)
    cases = [("cassie feet+pelvis B=4096", W.cassie_feet_pelvis_problem, 4096, "standing", ("f64", "f32")),
             ("cassie feet+pelvis B=65536, demo parameters (200 / 0.1 / 0.1)", W.cassie_feet_pelvis_problem, 65536, "standing", ("f64",), demo),
             ("cassie demo task set (cassie.cpp:43-81) B=65536", W.cassie_demo_problem, 65536, "standing", ("f64", "f32")),
             ("cassie feet+pelvis B=65536", W.cassie_feet_pelvis_problem, 65536, "standing", ("f64", "f32")),
             ("humanoid 5 Full tasks B=262144", W.humanoid_problem, 262144, "near", ("f64", "f32")),
             ("manipulator 1 Full task B=1048576", W.manipulator_problem, 1048576, "near", ("f64", "f32"))]
    for case in cases:
        name, make, B, start, dtypes = case[:5]
        prm = case[5] if len(case) > 5 else None
        B = max(64, int(B * args.scale))
        pb = make()
        pb.finalize(0)
        m = pb.model()
        names = W.task_frames(pb)
        qstar = W.sample_configurations(m, B)
        poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names)
                             for i in range(0, B, 65536)], dim=1)
        poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
        tg = W.targets_from_frame_poses(pb, poses)
        q0 = (np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)) if start == "standing"
              else W.near_start(m, qstar))
        for dt in dtypes:
            tdt = torch.float64 if dt == "f64" else torch.float32
            q0_d = torch.tensor(q0.T.copy(), dtype=tdt, device=dev)
            tg_d = torch.tensor(tg.T.copy(), dtype=tdt, device=dev)
            out = ik.dls_batch(pb, q0_d, tg_d, prm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                out = ik.dls_batch(pb, q0_d, tg_d, prm, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            ok = out["success"].sum().item()
            # the same batches as a stream through the pipelined queue (4 per kernel pair)
            queue = ik.SolveQueue(pb, 8, 4)
            outs = [None] * 4
            for k in range(4):
                _, outs[k] = queue.submit(q0_d, tg_d, prm, outs[k])
            queue.drain()
            nq_steps = max(4, args.steps // 4 * 4)
            e0.record()
            last = None
            for k in range(nq_steps):
                last, outs[k % 4] = queue.submit(q0_d, tg_d, prm, outs[k % 4])
            queue.flush()
            queue.wait_on_stream(last)
            e1.record()
            torch.cuda.synchronize()
            queue.drain()
            qms = e0.elapsed_time(e1) / nq_steps
            print(json.dumps({"config": name, "dtype": dt, "kernel": pb.kernel_name(dt), "batch": B, "ms_per_batch": ms,
                              "converged_solves_per_s": ok / (ms * 1e-3), "queue_ms_per_batch": qms,
                              "queue_converged_solves_per_s": ok / (qms * 1e-3), "converged_fraction": ok / B,
                              "mean_iterations": out["iters"].float().mean().item()}))
            del queue


if __name__ == "__main__":
    main()
