"""Structural checks of tools/gen_kernel.py's options for the role-distributed solve (humanoid specialisation): the
arithmetic of the generated code is checked against the oracle in test_device_code_cpu.py (CPU harness) and
test_gpu_parity.py; here the invariants the barrier scheme relies on -- every Gram entry and every coordinate of q has
exactly the owners the design says -- are checked on the generator's output itself."""
import copy
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_kernel as G  # noqa: E402

MODEL = os.path.join(ROOT, "build", "models", "humanoid_limbs.json")
SPEC = os.path.join(ROOT, "ik_b200", "specs", "humanoid_limbs.json")


@pytest.fixture(scope="module")
def inputs():
    if not os.path.exists(MODEL):  # written by `make gen` (conftest builds the library first)
        pytest.skip("flattened humanoid model not built")
    return json.load(open(MODEL)), json.load(open(SPEC))


def _emit(model, spec, **over):
    sp = copy.deepcopy(spec)
    sp.update(over)
    g = G.Generator(model, sp)
    return g, g.emit("SpecX", "x")


def _const(src, name):
    return int(re.search(r"\b%s = (-?\d+)" % name, src).group(1))


def _body(src, fn):
    i = src.index("static IKB_HD void %s(" % fn)
    j = src.index("\n    }\n", i)
    return src[i:j]


def test_committed_spec_uses_the_uniform_solve(inputs):
    model, spec = inputs
    assert spec.get("uniform_solve") and spec.get("parallel_solve")
    g, src = _emit(model, spec)
    M = _const(src, "M")
    nstrict = M * (M - 1) // 2
    assert _const(src, "NFACT") == nstrict + 3 * M      # L | d | rhs | Gram diagonal
    assert _const(src, "EOFF") == nstrict + M           # e is written straight into the rhs slots
    assert "DSTEP = true" in src and "PSOLVE = true" in src
    # one barrier after the Gram phase + one per block column except the last (solver role alone)
    assert _body(src, "psolve").count("sync();") == 1 + M // 5 - 1


def test_gram_phase_covers_every_entry_exactly_once(inputs):
    model, spec = inputs
    g, src = _emit(model, spec)
    M = _const(src, "M")
    nstrict = M * (M - 1) // 2
    seen = {}
    for role in range(5):
        for idx, i, j in re.findall(r"sL\.set\((\d+), g_(\d+)_(\d+)\);", _body(src, "ugram_w%d" % role)):
            i, j, idx = int(i), int(j), int(idx)
            assert (i, j) not in seen, "Gram entry (%d, %d) written by roles %d and %d" % (i, j, seen[(i, j)], role)
            seen[(i, j)] = role
            assert idx == (nstrict + 2 * M + i if i == j else i * (i - 1) // 2 + j)
    assert set(seen) == {(i, j) for i in range(M) for j in range(i + 1)}


@pytest.mark.parametrize("rolled", [True, False])
def test_rolled_gram_has_the_same_terms(inputs, rolled):
    """Rolled or not, every Gram entry accumulates one product per Jacobian column its two rows share."""
    model, spec = inputs
    g, src = _emit(model, spec, rolled_gram=rolled)
    assert ("J columns" in src) == rolled
    cols = {}
    for (r, c) in g.slots:
        cols.setdefault(r, set()).add(c)
    want = sum(len(cols[i] & cols[j]) for i in range(g.rows) for j in range(i + 1))
    assert sum(g.usolve_fma[0]) == want


def test_every_coordinate_of_q_is_stepped_and_stored(inputs):
    model, spec = inputs
    g, src = _emit(model, spec)
    nq = _const(src, "NQ")
    store = _body(src, "store_q")
    per_role = [set(int(k) for k in re.findall(r"dst\[(\d+) \* es\]", blk)) for blk in store.split("if (role ==")[1:]]
    assert len(per_role) == 5 and set().union(*per_role) == set(range(nq))
    assert set(range(7)) <= per_role[spec["solver_warp"]]              # the free-flyer belongs to the solver role
    step = _body(src, "step_role")
    common, *roles = step.split("if (role ==")
    clamp = lambda text: set(int(k) for k in re.findall(r"q\[(\d+)\] = min_\(", text))  # noqa: E731
    assert clamp(common) == set(range(7))                              # stepped by every role (redundantly)
    for k, text in enumerate(roles):
        assert clamp(text) | (clamp(common) if k == spec["solver_warp"] else set()) == per_role[k]


@pytest.mark.parametrize("rolled,extra", [(0, 0), (2, 1)])
def test_two_barrier_variant_still_generates(inputs, rolled, extra):
    """The earlier distributed solve (gen_solve_parallel: role-specific bodies, two barriers per block column)."""
    model, spec = inputs
    g, src = _emit(model, spec, uniform_solve=False, rolled_update=rolled)
    M = _const(src, "M")
    assert "DSTEP = false" in src and _const(src, "EOFF") == 0
    assert _const(src, "NFACT") == M * (M + 1) // 2 + extra * M        # + yp for the rolled left-looking loops
    assert ("minus column k of the factor" in src) == bool(rolled)
    assert all(" psolve_w%d(" % k in src for k in range(5))


# ---- arrow specs (round 2): staging slots of the one-trip-ahead refill, the SOLVER role off the shared-column system ----
ARROW = [("cassie_feet_pelvis_arrow", "cassie_feet_pelvis_arrow"), ("cassie_feet_pelvis_arrow_b", "cassie_feet_pelvis_arrow_b"),
         ("humanoid_limbs_arrow", "humanoid_limbs_arrow")]


def _arrow_inputs(name):
    model = os.path.join(ROOT, "build", "models", "%s.json" % name)
    spec = os.path.join(ROOT, "ik_b200", "specs", "%s.json" % name)
    if not os.path.exists(model):
        pytest.skip("flattened model of %s not built" % name)
    return json.load(open(model)), json.load(open(spec))


def _role_fn(src, name):
    m = re.search(r"static constexpr int %s\(int role\) \{ return (.*?); \}" % name, src)
    return {int(a): int(b) for a, b in re.findall(r"role == (\d+) \? (-?\d+)", m.group(1))}


@pytest.mark.parametrize("name,_", ARROW)
def test_staging_slots_belong_to_a_role_that_publishes_in_the_factor_strip(name, _):
    """dls_spec.cuh (kPrefetch) copies a converged slot's NEXT configuration into Spec::QSTAGE .. + NQ of the Jacobian strip
    in the middle of a trip.  Those slots must be rows of ONE non-solver role that does not publish its contribution there
    (a role without private columns does: the other roles read it while the copy is in flight)."""
    model, spec = _arrow_inputs(name)
    g, src = _emit(model, spec)
    q0, NQ, solver = _const(src, "QSTAGE"), _const(src, "NQ"), _const(src, "SOLVER")
    assert q0 >= 0
    first, count = _role_fn(src, "j_first"), _role_fn(src, "j_count")
    owners = [k for k in first if first[k] <= q0 and q0 + NQ <= first[k] + count[k]]
    assert len(owners) == 1 and owners[0] != solver
    assert g.arrow["pub"][owners[0]][0] == "L"
    # the option that keeps the strip out of shared memory or ends psolve() behind its own barrier switches the staging off
    assert _const(_emit(model, spec, arrow_cap_solo=True)[1], "QSTAGE") == -1


@pytest.mark.parametrize("name,_", ARROW)
def test_solver_role_skips_the_shared_system_only_when_the_spec_says_so(name, _):
    """`arrow_solver_skips_cap` (default on; off in the humanoid spec, where it measured slower): a SOLVER role without private
    columns leaves the shared-column system to the other roles, and the first of them stores s for everybody's step."""
    model, spec = _arrow_inputs(name)
    for on in (True, False):
        g, src = _emit(model, spec, arrow_solver_skips_cap=on)
        solver = _const(src, "SOLVER")
        body = _body(src, "psolve")
        assert ("if (role != %d) {" % solver in body) == on
        storer = re.search(r"if \(role == (\d+)\) \{\n#pragma unroll\n\s+for \(int i = 0; i < \d+; \+\+i\) sL\.set\(", body)
        assert storer and (int(storer.group(1)) != solver) == on
        assert body.count("sync();") == 1                      # still ONE barrier inside psolve()
    committed = json.load(open(os.path.join(ROOT, "ik_b200", "specs", "%s.json" % name)))
    assert committed.get("arrow_solver_skips_cap", True) == (not name.startswith("humanoid"))
