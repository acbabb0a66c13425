"""CPU: the diff_ik oracle (oracle/diffik_oracle.c, the reference's LP solved exactly) against the HiGHS-solved golden
vectors of the same LP (tests/golden/diff_ik_vectors.json, made by tests/golden/make_diff_ik_golden.py), plus the
reference's own property test (crates/optik/tests/test_ik.rs:184-209)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from conftest import ROOT

LINKS = {"ur3e": ("ur_base_link", "ur_ee_link"), "panda": ("panda_link0", "panda_link8")}


def chain(name):
    b, e = LINKS[name]
    return O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", name + ".urdf")).read(), b, e)


@pytest.fixture(scope="module")
def cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "diff_ik_vectors.json")))["cases"]


def test_oracle_matches_lp_golden(cases):
    chains = {n: chain(n) for n in LINKS}
    for c in cases:
        ch = chains[c["robot"]]
        alpha, v = ch.diff_ik(c["x0"], c["V_WE"], c["v_max"])
        assert abs(alpha - c["alpha"]) <= 1e-8, (c["robot"], alpha, c["alpha"])
        assert np.all(np.abs(v) <= np.array(c["v_max"]) * (1 + 1e-12))
        if c["v"] is not None:  # 6-DOF: the LP's v is unique
            assert np.abs(v - np.array(c["v"])).max() <= 1e-8


def test_oracle_solution_tracks_the_twist(cases):
    """J_W v = alpha V (the TODO of tests/test_ik.rs:207), checked with a central-difference world-frame Jacobian."""
    chains = {n: chain(n) for n in LINKS}
    for c in cases[::4]:
        ch = chains[c["robot"]]
        x0 = np.array(c["x0"])
        alpha, v = ch.diff_ik(x0, c["V_WE"], c["v_max"])
        h = 1e-6
        _, e1 = ch.fk(x0 + h * v)
        _, e0 = ch.fk(x0 - h * v)
        lin = (e1[4:7] - e0[4:7]) / (2 * h)
        assert np.abs(lin - alpha * np.array(c["V_WE"][:3])).max() < 1e-6


def test_reference_property_bounds():  # tests/test_ik.rs:184-209
    ch = chain("ur3e")
    rng = np.random.default_rng(42)
    for _ in range(20):
        x0 = rng.uniform(ch.lb, ch.ub)
        alpha, v = ch.diff_ik(x0, rng.random(6), np.ones(6))
        assert -1e-6 <= alpha <= 1 + 1e-6
        assert np.all(v >= -1 - 1e-6) and np.all(v <= 1 + 1e-6)


def test_singular_configuration_gives_a_zero_step():
    """The reference's LP is always feasible (alpha = 0, v = 0) and bounded: Clarabel reports Solved and diff_ik returns
    Some((alpha, v)) at singular configurations too (lib.rs:231-238); an unreachable twist gives alpha = 0."""
    ch = chain("ur3e")
    V = [0.3, 0.1, 0.2, 0.5, 0.4, 0.6]
    for x0 in (np.zeros(6), np.array([0.3, -1.0, 1.2, 0.4, 0.0, 0.2])):  # home pose; wrist singularity (q5 = 0)
        alpha, v = ch.diff_ik(x0, V, np.ones(6))
        assert alpha == 0.0 and np.all(v == 0.0)
    # a twist inside the range of the singular Jacobian is followed: joint 1 alone at the wrist singularity
    x0 = np.array([0.3, -1.0, 1.2, 0.4, 0.0, 0.2])
    from tests_golden_helpers import world_jacobian
    Jw = world_jacobian(ch, x0)
    Vr = Jw[:, 0] * 0.5
    alpha, v = ch.diff_ik(x0, Vr, np.ones(6))
    assert alpha == 1.0 and np.abs(Jw @ v - Vr).max() < 1e-12
