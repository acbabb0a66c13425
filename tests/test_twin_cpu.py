"""CPU: the solver twin's evaluator agrees with the golden-pinned oracle; twin solves satisfy the reference's
own behavioural pins (tests/test_ik.rs) when judged by the pinned oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from conftest import ROOT

ROBOTS = {"panda": ("panda_link0", "panda_link8"), "ur5": ("base_link", "ee_link"),
          "ur3e": ("ur_base_link", "ur_ee_link"), "snake20": ("seg0", "tip")}


def chain(name):
    b, e = ROBOTS[name]
    return O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", name + ".urdf")).read(), b, e)


@pytest.mark.parametrize("name", list(ROBOTS))
@pytest.mark.parametrize("weights", [((1, 1, 1), (1, 1, 1)), ((0.0, 5.0, 0.25), (0.005, 1.0, 0.99))])
def test_twin_evaluator_matches_pinned_oracle(name, weights):
    ch = chain(name)
    wl, wa = weights
    rng = np.random.default_rng(3)
    P = O.twin_params(wl=wl, wa=wa)
    for _ in range(50):
        q = rng.uniform(ch.lb, ch.ub)
        _, tgt = ch.fk(rng.uniform(ch.lb, ch.ub))
        ev = O.twin_eval(ch, q, tgt, P)
        _, ee = ch.fk(q)
        assert np.abs(ev["ee"][:7] - ee[:7]).max() < 1e-13
        f = ch.objective(q, tgt, wl, wa)
        assert abs(ev["f"] - f) <= 1e-12 * max(1.0, f)
        g = ch.objective_grad(q, tgt, wl, wa)
        assert np.abs(ev["grad"] - g).max() <= 1e-11 * max(1.0, np.abs(g).max())


def test_twin_evaluator_near_zero_error():
    ch = chain("ur3e")
    rng = np.random.default_rng(5)
    q = rng.uniform(ch.lb, ch.ub)
    _, tgt = ch.fk(q)
    for d in (0.0, 1e-9, 1e-5, 1e-3):
        ev = O.twin_eval(ch, q + d, tgt)
        assert np.isfinite(ev["f"]) and np.all(np.isfinite(ev["grad"]))
        if d > 0:
            g = ch.objective_grad(q + d, tgt)
            assert np.abs(ev["grad"] - g).max() < 1e-9


def test_forward_backward_ur3e():  # tests/test_ik.rs:91-130: tol_f = 1e-12, 25 restarts, seed zeros, FK(ik(T)) == T to 1e-6
    ch = chain("ur3e")
    rng = np.random.default_rng(42)
    P = O.twin_params(tol_f=1e-12)
    for _ in range(10):
        _, tgt = ch.fk(rng.random(6))
        res = O.twin_ik(ch, tgt, np.zeros(6), 0, 25, "speed", P)
        assert res["found"]
        _, ee = ch.fk(res["q"])
        assert np.abs(ee[4:7] - tgt[4:7]).max() < 1e-6
        assert min(np.abs(ee[:4] - tgt[:4]).max(), np.abs(ee[:4] + tgt[:4]).max()) < 1e-6
        assert np.all(res["q"] >= ch.lb) and np.all(res["q"] <= ch.ub)


def test_quality_not_farther_than_speed_ur3e():  # tests/test_ik.rs:132-182
    ch = chain("ur3e")
    rng = np.random.default_rng(42)
    for _ in range(20):
        _, tgt = ch.fk(rng.random(6))
        x0 = np.zeros(6)
        s = O.twin_ik(ch, tgt, x0, 0, 15, "speed")
        q = O.twin_ik(ch, tgt, x0, 0, 15, "quality")
        assert s["found"] and q["found"]
        assert np.linalg.norm(q["q"] - x0) <= np.linalg.norm(s["q"] - x0)


def test_twin_solutions_pass_reference_predicate():
    """Every attempt the twin calls converged satisfies f(q) < tol_f and lb <= q <= ub under the PINNED oracle."""
    for name in ("panda", "ur5", "snake20"):
        ch = chain(name)
        rng = np.random.default_rng(11)
        _, tgt = ch.fk(rng.uniform(ch.lb, ch.ub))
        q, f, st, ev = O.twin_attempts(ch, tgt, 0.5 * (ch.lb + ch.ub), 0, 200)
        ok = st == 1
        assert ok.sum() > 20
        for i in np.where(ok)[0]:
            assert ch.objective(q[i], tgt) < 1e-6
            assert np.all(q[i] >= ch.lb) and np.all(q[i] <= ch.ub)
        assert np.all(ev <= 32)


def test_unreachable_target_fails():  # tests/test_ik.rs:24-43 (impossible goal) -> no solution
    ch = chain("ur3e")
    tgt = O.pose8([0, 0, 0, 1], [100.0, 100.0, 100.0])
    res = O.twin_ik(ch, tgt, np.zeros(6), 0, 16, "speed")
    assert not res["found"]


def test_twin_vs_slsqp_standin_per_target():
    """CPU-side version of tests/test_gpu_parity.py::test_solve_against_slsqp_standin on a small sample: the LM twin (==
    the kernel, bit for bit) and the SLSQP stand-in agree on per-target success, and polished from the same start they
    land on the same solution (<= 1e-6 rad)."""
    import slsqp_standin as S
    ch = chain("ur3e")
    rng = np.random.default_rng(12)
    P = O.twin_params(layout=1)
    PP = O.twin_params(tol_f=1e-20, layout=1, max_evals=64)
    agree = n = 0
    worst = 0.0
    for t in range(120):
        tgt = ch.fk(rng.uniform(ch.lb, ch.ub))[1]
        x0 = rng.uniform(ch.lb, ch.ub)
        a = O.twin_ik(ch, tgt, x0, 0, 100, "speed", P)
        found, q_s, _, _ = S.slsqp_ik(ch, tgt, x0, 100)
        agree += a["found"] == found
        if a["found"] and found:
            qa, fa, sa, ea = O.twin_attempts(ch, tgt, a["q"], 0, 1, PP)
            okp, qp = S.slsqp_polish(ch, tgt, a["q"], 1e-20)
            if sa[0] == 1 and okp:
                n += 1
                worst = max(worst, np.abs(qa[0] - qp).max())
    assert agree == 120 and n >= 60 and worst <= 1e-6, (agree, n, worst)
