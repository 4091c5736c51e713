#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ (committed, small).

The reference (dazzmo/ik) cannot run here -- Pinocchio / Eigen are absent and its own tests pin nothing (SURVEY.md 4,
8c) -- so these vectors come from the CPU ORACLE (oracle/ik_oracle.c, a restatement of the reference path).  They pin
(a) the oracle against regressions, (b) the product's URDF flattener (exact topology) and (c) the CUDA path on the GPU
box, where /root/reference and this script's inputs are not needed.

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ik_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.bridge import make_workload, oracle_model, oracle_problem_like  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def solve_case(name, robot, ff, make, B, seed, params, **wk):
    pb = make()
    om = oracle_model(robot, ff)
    opb = oracle_problem_like(pb, om)
    q0, tg, qstar = make_workload(pb, om, B, seed=seed, **wk)
    prm = O.params(*params)
    q, ok, it, res = O.dls_batch(opb, q0, tg, prm)
    e0 = np.stack([opb.evaluate(q0[b], tg[b])[0] for b in range(B)])
    J0 = np.stack([opb.evaluate(q0[b], tg[b])[1] for b in range(B)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), q0=q0, targets=tg, q=q, success=ok, iters=it, resid=res,
                        e0=e0, J0=J0, params=np.array(params, dtype=np.float64))
    print("%s: B=%d converged=%d mean iters=%.2f" % (name, B, ok.sum(), it.mean()))


def topology_case(name, robot, ff):
    om = oracle_model(robot, ff)
    f = om.flat
    qs = W.sample_configurations(__import__("ik_b200").Model.builtin(robot, free_flyer=ff), 8, seed=3)
    frames = list(range(om.nframes))
    poses = np.stack([[om.frame_placement(q, fr) for fr in frames] for q in qs])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), joint_names=np.array(f["names"]), parent=f["parent"],
                        jtype=f["jtype"], idx_q=f["idx_q"], idx_v=f["idx_v"], placement=f["placement"],
                        lower=f["lower"], upper=f["upper"], frame_names=np.array(f["frame_names"]),
                        frame_parent=f["frame_parent"], frame_placement=f["frame_placement"], fk_q=qs, fk_poses=poses)
    print("%s: njoints=%d nframes=%d" % (name, om.njoints, om.nframes))


def main():
    os.makedirs(OUT, exist_ok=True)
    topology_case("topology_cassie", "cassie", True)
    topology_case("topology_ur5", "ur5", False)
    topology_case("topology_humanoid", "humanoid", True)
    topology_case("topology_manipulator", "manipulator", False)
    solve_case("cassie_defaults", "cassie", True, W.cassie_feet_pelvis_problem, 96, 2026, (100, 1.0, 1e-2, 1e-4),
               standing=W.CASSIE_STANDING)
    solve_case("cassie_demo", "cassie", True, W.cassie_feet_pelvis_problem, 48, 2027, (200, 1e-1, 1e-1, 1e-4),
               standing=W.CASSIE_STANDING)
    solve_case("manipulator_defaults", "manipulator", False, W.manipulator_problem, 96, 2028, (100, 1.0, 1e-2, 1e-4),
               start="near")
    solve_case("humanoid_defaults", "humanoid", True, W.humanoid_problem, 32, 2029, (100, 1.0, 1e-2, 1e-4), start="near")


if __name__ == "__main__":
    main()
