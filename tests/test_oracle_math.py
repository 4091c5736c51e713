"""First-principles validation of the CPU oracle (parity is UNPINNED by the reference: it ships no golden vectors,
SURVEY.md 4 / 8c).  exp/log round trips, finite-difference Jacobians, LDLT vs numpy, integrate consistency."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.common import oracle_model

RNG = np.random.default_rng(20261018)


def rand_se3(scale=1.0):
    return O.exp6(np.concatenate([RNG.uniform(-1, 1, 3), RNG.uniform(-scale, scale, 3)]))


def test_exp3_is_rotation_and_log3_inverts_it():
    for _ in range(200):
        w = RNG.uniform(-1.5, 1.5, 3)
        R = O.exp3(w)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(R) - 1) < 1e-14
        w2, th = O.log3(R)
        assert np.allclose(w2, w, atol=1e-12) and abs(th - np.linalg.norm(w)) < 1e-12


def test_exp3_matches_matrix_exponential():
    from scipy.linalg import expm

    for _ in range(50):
        w = RNG.uniform(-2, 2, 3)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        assert np.allclose(O.exp3(w), expm(K), atol=1e-13)


def test_exp6_matches_matrix_exponential():
    from scipy.linalg import expm

    for _ in range(50):
        v = RNG.uniform(-1, 1, 6)
        w = v[3:]
        X = np.zeros((4, 4))
        X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        X[:3, 3] = v[:3]
        E = expm(X)
        M = O.exp6(v)
        assert np.allclose(M[:9].reshape(3, 3), E[:3, :3], atol=1e-13) and np.allclose(M[9:], E[:3, 3], atol=1e-13)


@pytest.mark.parametrize("scale", [1e-9, 1e-5, 1e-3, 0.5, 1.7])
def test_log6_exp6_round_trip(scale):
    for _ in range(100):
        v = np.concatenate([RNG.uniform(-1, 1, 3), RNG.uniform(-1, 1, 3) * scale])
        assert np.allclose(O.log6(O.exp6(v)), v, atol=1e-11)


def test_log3_near_pi_branch():
    for ang in (np.pi - 5e-3, np.pi - 1e-6, np.pi - 2e-2):
        axis = RNG.normal(size=3)
        axis /= np.linalg.norm(axis)
        w, th = O.log3(O.exp3(axis * ang))
        assert abs(th - ang) < 1e-7 and np.allclose(w, axis * ang, atol=1e-6)


@pytest.mark.parametrize("scale", [1e-6, 0.3, 2.0])
def test_Jlog6_against_central_differences(scale):
    """log6(M exp6(xi)) ~ log6(M) + Jlog6(M) xi (SURVEY 8c.4 convention)."""
    h = 1e-6
    for _ in range(20):
        M = rand_se3(scale)
        J = O.Jlog6(M)
        Jn = np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6)
            d[k] = h
            Jn[:, k] = (O.log6(O.se3_mul(M, O.exp6(d))) - O.log6(O.se3_mul(M, O.exp6(-d)))) / (2 * h)
        assert np.abs(J - Jn).max() < 5e-7


def test_se3_mul_actinv_consistency():
    for _ in range(50):
        A, B = rand_se3(), rand_se3()
        C = O.se3_mul(A, B)
        assert np.allclose(O.se3_actinv(A, C), B, atol=1e-13)


def test_quaternion_conversions():
    C = O.C
    for _ in range(200):
        R = O.exp3(RNG.uniform(-3.1, 3.1, 3))
        q = np.zeros(4)
        O.lib().iko_rot_to_quat(O._pd(np.ascontiguousarray(R.reshape(-1))), O._pd(q))
        assert abs(np.linalg.norm(q) - 1) < 1e-13
        R2 = np.zeros(9)
        O.lib().iko_quat_to_rot(O._pd(q), O._pd(R2))
        assert np.allclose(R2.reshape(3, 3), R, atol=1e-13)


def test_ldlt_solve_matches_numpy():
    for n in (1, 3, 6, 12, 30):
        for _ in range(20):
            J = RNG.normal(size=(n, n + 5))
            A = J @ J.T + 1e-4 * np.eye(n)
            b = RNG.normal(size=n)
            x = O.ldlt_solve(A, b)
            assert np.allclose(A @ x, b, atol=1e-8 * max(1, np.abs(b).max()))
            assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-6, atol=1e-8)


def test_ldlt_pivoting_handles_indefinite_ordering():
    A = np.diag([1e-8, 5.0, 2.0]) + 1e-3
    A = (A + A.T) / 2
    b = np.array([1.0, -2.0, 0.5])
    assert np.allclose(O.ldlt_solve(A, b), np.linalg.solve(A, b), rtol=1e-9)


@pytest.mark.parametrize("robot,ff", [("cassie", True), ("ur5", False), ("humanoid", True), ("manipulator", False)])
def test_integrate_is_consistent_with_fk(robot, ff):
    """Frame velocity check: FK(integrate(q, h v)) ~ FK(q) * exp6(h Jf_LOCAL v) for every frame-supporting column."""
    m = oracle_model(robot, ff)
    h = 1e-6
    q = m.neutral()
    q[m.nq - (m.nv - (6 if ff else 0)):] = RNG.uniform(-0.5, 0.5, m.nv - (6 if ff else 0))
    if ff:
        q = m.integrate(q, np.concatenate([RNG.uniform(-0.3, 0.3, 6), np.zeros(m.nv - 6)]))
    f = m.nframes - 1
    Jf = m.frame_jacobian_local(q, f)
    M0 = m.frame_placement(q, f)
    for c in range(m.nv):
        v = np.zeros(m.nv)
        v[c] = h
        M1 = m.frame_placement(m.integrate(q, v), f)
        xi = O.log6(O.se3_actinv(M0, M1)) / h
        assert np.abs(xi - Jf[:, c]).max() < 1e-5, (robot, c)


def test_freeflyer_integrate_keeps_unit_quaternion():
    m = oracle_model("cassie")
    q = m.neutral()
    for _ in range(200):
        v = np.zeros(m.nv)
        v[:6] = RNG.uniform(-0.5, 0.5, 6)
        q = m.integrate(q, v)
    assert abs(np.linalg.norm(q[3:7]) - 1) < 1e-9


def test_clip_applies_to_all_entries():
    m = oracle_model("cassie")
    q = np.full(m.nq, 10.0)
    qc = m.clip(q)
    assert np.array_equal(qc[:7], q[:7])  # free-flyer limits are +-DBL_MAX: no-op (SURVEY 8a notes)
    assert np.array_equal(qc[7:], m.flat["upper"][7:])
    assert np.array_equal(m.clip(-q)[7:], m.flat["lower"][7:])
