"""CPU ORACLE side (test infrastructure, NOT product code): the BASELINE.json headline workload built WITHOUT the product.

bench.py's `--impl reference` arm and its `cpu_baseline` leg time the restated reference CPU path; they must not load
libikb200.so.  This module rebuilds the Cassie feet+pelvis problem and its seeded
random reachable targets from oracle/urdf_flatten.py and the oracle's own FK only.  The sampling rule is a
restatement of ik_b200/workloads.py (counter-based SplitMix64 keyed by (seed, b, k)); tests/test_oracle_workload.py
checks that the two produce bit-identical q0 / targets.
"""
import os

import numpy as np

from . import oracle as O
from . import urdf_flatten as U

# cassie-description/srdf/cassie.srdf:22-39 (group_state "default"), in model joint order
CASSIE_STANDING = [0.0045, 0.0, 0.4973, -1.1997, 0.0, 1.4267, 0.0, -1.5968,
                   -0.0045, 0.0, 0.4973, -1.1997, 0.0, 1.4267, 0.0, -1.5968]
_URDF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ik_b200", "data", "cassie.urdf")  # data file only
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _uniform01(seed, b, k):
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) + np.uint64(0x632BE59BD9B4E019) * np.uint64(k + 1))
        x = _splitmix64(key ^ (b.astype(np.uint64) * np.uint64(0xD1342543DE82EF95)))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def sample_configurations(flat, B, seed=12345, b0=0):
    """q* [B, nq]: revolutes uniform inside the URDF limits, base position U[-0.2, 0.2]^3 m, base orientation exp3(u),
    u ~ U[-0.3, 0.3]^3 rad."""
    b = np.arange(b0, b0 + B, dtype=np.uint64)
    q = np.zeros((B, flat["nq"]))
    lo, hi = np.asarray(flat["lower"]), np.asarray(flat["upper"])
    k = 0
    for j in range(1, flat["njoints"]):
        iq = int(flat["idx_q"][j])
        if flat["jtype"][j] == U.J_FREEFLYER:
            for i in range(3):
                q[:, iq + i] = -0.2 + 0.4 * _uniform01(seed, b, k)
                k += 1
            u = np.stack([-0.3 + 0.6 * _uniform01(seed, b, k + i) for i in range(3)], axis=1)
            k += 3
            th = np.linalg.norm(u, axis=1)
            s = np.where(th > 1e-12, np.sin(th / 2) / np.maximum(th, 1e-300), 0.5)
            q[:, iq + 3:iq + 6] = u * s[:, None]
            q[:, iq + 6] = np.cos(th / 2)
        else:
            w = hi[iq] - lo[iq]
            q[:, iq] = lo[iq] + w * _uniform01(seed, b, k)
            k += 1
    return q


def cassie_feet_pelvis(B, seed=12345, b0=0):
    """(oracle problem, q0 [B, nq], targets [B, 36]): pelvis Full + LeftFootFront / RightFootFront Position in `universe`
    (BASELINE.json configs 1-3); targets = oracle FK of the seeded reachable configurations; q0 = SRDF standing pose."""
    with open(_URDF) as f:
        om = O.Model.from_urdf(f.read(), True)
    opb = O.Problem(om, 0)
    opb.add_frame_task("pelvis", O.FULL, "universe", 0)
    opb.add_frame_task("LeftFootFront", O.POSITION, "universe", 0)
    opb.add_frame_task("RightFootFront", O.POSITION, "universe", 0)
    qstar = sample_configurations(om.flat, B, seed, b0)
    ids = [om.frame_id(n) for n in ("pelvis", "LeftFootFront", "RightFootFront")]
    tg = np.zeros((B, 36))
    eye = np.eye(3).reshape(-1)
    for b in range(B):
        oMi = om.fk(qstar[b])
        for t, f in enumerate(ids):
            M = np.zeros(12)
            O.lib().iko_frame_placement(O.C.byref(om.c), O._pd(oMi), O.C.c_int(f), O._pd(M))
            tg[b, 12 * t:12 * t + 12] = M
            if t > 0:
                tg[b, 12 * t:12 * t + 9] = eye   # Position tasks: (Identity, p), as the reference's callers set them
    q0 = om.neutral()
    q0[om.nq - len(CASSIE_STANDING):] = CASSIE_STANDING
    return opb, np.tile(q0, (B, 1)), tg
