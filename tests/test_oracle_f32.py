"""The single-precision build of the oracle (ik_oracle.c -DIKO_F32: number_t = float, reference common.hpp:13) against the
FP64 build on BASELINE config 3 (Cassie feet+pelvis, 65,536 problems): how far ANY faithful FP32 evaluation of
ik::dls lands from the FP64 answer.  north_star asks for 1e-4 rad in FP32; this shows that bar is met by ~99.45 % of the
problems and is unreachable for the rest in single precision -- VERDICT r1 item 1c.  The GPU twin of this test
(tests/test_gpu_parity.py::test_f32_spread_is_inherent_to_single_precision) holds the FP32 kernel to the same
distribution."""
import os

import numpy as np

from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like

NT = os.cpu_count() or 1


def test_f32_oracle_against_f64_oracle_full_batch():
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 65536
    q0, tg, _ = make_workload(pb, om, B, standing=W.CASSIE_STANDING)
    q64, ok64, it64, res64 = O.dls_batch(opb, q0, tg, nthreads=NT)
    q32, ok32, it32, res32 = O.dls_batch_f32(opb, q0, tg, nthreads=NT)
    assert (ok32 == ok64).mean() > 0.9999           # measured: 1 flag of 65,536 differs
    both = ok32 & ok64 & (it32 == it64)
    assert both.mean() > 0.9735                     # = the converged fraction; step counts agree on 99.992 %
    err = np.abs(q32 - q64).max(axis=1)[both]
    med, p99, p999, mx = np.median(err), np.percentile(err, 99), np.percentile(err, 99.9), err.max()
    print("oracle f32 vs f64: median %.2e p99 %.2e p99.9 %.2e max %.2e within 1e-4: %.5f"
          % (med, p99, p999, mx, (err < 1e-4).mean()))
    # measured here: 2.05e-06 / 6.02e-05 / 3.85e-04 / 3.97e-02, 99.45 % within 1e-4 rad
    assert 1e-6 < med < 4e-6 and 3e-5 < p99 < 1e-4 and 1e-4 < p999 < 1e-3
    assert mx > 1e-3, "single precision cannot meet 1e-4 rad on every problem of this workload"
    assert 0.99 < (err < 1e-4).mean() < 0.999
    assert np.abs(res32 - res64)[both].max() < 1e-5
