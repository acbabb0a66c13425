"""Golden vectors for Robot::diff_ik (crates/optik/src/lib.rs:101-239).

The reference solves   max alpha  s.t.  J_W v = alpha V_WE, |v_i| <= vmax_i, 0 <= alpha <= 1   with Clarabel
(un-vendored crate, not buildable here).  The LP's optimal alpha is unique, and for a 6-DOF arm with a regular
Jacobian so is v; this script builds exactly the reference's constraint set (alpha bounds lib.rs:134-151, velocity
box :155-174, J_W v - alpha V = 0 with the body Jacobian rotated into the world frame :178-197) from the
golden-pinned oracle's FK / Jacobian and solves it with scipy's HiGHS (an independent LP solver), so the fixture pins
the PROBLEM the reference poses rather than our closed form.  Inputs follow tests/test_ik.rs:184-209 (UR3e, x0 uniform
in the limits, V_WE uniform in [0,1)^6, vmax = 1) plus ragged vmax and a 7-DOF arm (Panda, an extension: the
reference's own assembly only accepts n = 6, lib.rs:194-195).
Writes tests/golden/diff_ik_vectors.json.      Run:  python tests/golden/make_diff_ik_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy.optimize import linprog

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

LINKS = {"ur3e": ("ur_base_link", "ur_ee_link"), "panda": ("panda_link0", "panda_link8")}


def quat_rot(q, v):
    u, w = np.array(q[:3]), q[3]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def world_jacobian(ch, x0):
    _, ee = ch.fk(x0)
    Jb = ch.joint_jacobian(x0)  # 6 x n, body frame, rows [lin; ang]
    Jw = np.zeros_like(Jb)
    for c in range(Jb.shape[1]):
        Jw[:3, c] = quat_rot(ee[:4], Jb[:3, c])
        Jw[3:, c] = quat_rot(ee[:4], Jb[3:, c])
    return Jw


def solve_lp(Jw, V, vmax):
    n = Jw.shape[1]
    c = np.zeros(n + 1)
    c[n] = -1.0
    A_eq = np.hstack([Jw, -V.reshape(6, 1)])
    res = linprog(c, A_eq=A_eq, b_eq=np.zeros(6), bounds=[(-m, m) for m in vmax] + [(0.0, 1.0)], method="highs",
                  options={"primal_feasibility_tolerance": 1e-10, "dual_feasibility_tolerance": 1e-10})
    assert res.status == 0, res.message
    return float(res.x[n]), res.x[:n]


def main():
    from oracle import oracle as O
    rng = np.random.default_rng(42)
    cases = []
    for name, count in (("ur3e", 24), ("panda", 24)):
        base, ee = LINKS[name]
        ch = O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", name + ".urdf")).read(), base, ee)
        for k in range(count):
            x0 = rng.uniform(ch.lb, ch.ub)
            V = rng.random(6) * (1.0 if k % 3 else 0.2)          # small twists give alpha = 1
            vmax = np.ones(ch.n) if k % 2 == 0 else rng.uniform(0.2, 2.0, ch.n)
            alpha, v = solve_lp(world_jacobian(ch, x0), V, vmax)
            cases.append({"robot": name, "x0": list(x0), "V_WE": list(V), "v_max": list(vmax), "alpha": alpha,
                          "v": list(v) if ch.n == 6 else None})
    # singular configurations (common home poses): the LP stays feasible and bounded, so the reference returns
    # Some((alpha, v)) there too (lib.rs:231-238); a generic twist is unreachable => alpha = 0.  Only alpha is unique.
    ur3e = O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", "ur3e.urdf")).read(), *LINKS["ur3e"])
    for x0 in (np.zeros(6), np.array([0.0, -1.0, 0.0, 0.3, 0.0, 0.0]), np.array([0.3, -1.0, 1.2, 0.4, 0.0, 0.2])):
        for V in (np.array([0.3, 0.1, 0.2, 0.5, 0.4, 0.6]), rng.random(6)):
            alpha, v = solve_lp(world_jacobian(ur3e, x0), V, np.ones(6))
            cases.append({"robot": "ur3e", "x0": list(x0), "V_WE": list(V), "v_max": [1.0] * 6, "alpha": alpha, "v": None,
                          "singular": True})
    out = {"source": "scipy.optimize.linprog(method='highs') on the LP of kylc/optik@355e463 crates/optik/src/lib.rs:123-239; "
                     "J_W from oracle/optik_oracle.c (pinned to the reference's FK goldens)",
           "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "diff_ik_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print(path, len(cases), "cases; alpha<1 in", sum(c["alpha"] < 1 - 1e-9 for c in cases))


if __name__ == "__main__":
    main()
