"""Calibration of the CPU baseline against an SLSQP-class inner solver.

The reference's inner solver is NLopt SLSQP (un-vendored, cannot be built here); `bench.py --impl reference` times the
reference's restart loop over OUR inner solver's CPU twin instead.  This script runs scipy's SLSQP (the same Kraft
algorithm family) over the golden-pinned oracle objective/gradient with the reference's stop rules -- stopval = tol_f
(lib.rs:345, emulated by raising from the callback), ftol_abs = 1e-3*tol_f (lib.rs:283-293), bounds (lib.rs:348-349),
success iff f < tol_f (lib.rs:376-377) -- on the bench workload's seeds, and prints evaluations per attempt, success
per attempt and evaluations per CONVERGED attempt for both solvers, so the port's solves/s can be scaled to an
SLSQP-class loop.     python tests/experiments/exp_slsqp_calibration.py [robot] [attempts]
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

LINKS = {"panda": ("panda_link0", "panda_link8"), "ur5": ("base_link", "ee_link"), "ur3e": ("ur_base_link", "ur_ee_link")}


class Reached(Exception):
    pass


def slsqp_attempt(ch, tgt, q0, tol_f=1e-6):
    evals = [0]
    best = [np.inf, None]

    def fun(q):
        evals[0] += 1
        f = ch.objective(q, tgt)
        if f < best[0]:
            best[0], best[1] = f, q.copy()
        if f < tol_f:
            raise Reached()
        return f

    def jac(q):
        return ch.objective_grad(q, tgt)

    try:
        minimize(fun, q0, jac=jac, method="SLSQP", bounds=list(zip(ch.lb, ch.ub)),
                 options={"ftol": 1e-3 * tol_f, "maxiter": 200})
    except Reached:
        return True, evals[0]
    return best[0] < tol_f, evals[0]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "panda"
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    b, e = LINKS[name]
    ch = O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", name + ".urdf")).read(), b, e)
    rng = np.random.default_rng(42)
    x0 = 0.5 * (ch.lb + ch.ub)
    targets = [ch.fk(rng.uniform(ch.lb, ch.ub))[1] for _ in range(8)]
    per = A // len(targets)
    s_ok = s_ev = l_ok = l_ev = 0
    t_eval = time.perf_counter()
    for _ in range(2000):
        ch.objective_grad(x0, targets[0])
    t_eval = (time.perf_counter() - t_eval) / 2000
    for tgt in targets:
        q, f, st, ev = O.twin_attempts(ch, tgt, x0, 1, 1 + per, O.twin_params(layout=1 if ch.n <= 8 else 0))
        l_ok += int((st == 1).sum())
        l_ev += int(ev.sum())
        for r in range(1, 1 + per):
            ok, n_ev = slsqp_attempt(ch, tgt, ch.restart_seed(r))
            s_ok += ok
            s_ev += n_ev
    n = per * len(targets)
    print(f"{name}: {n} attempts over {len(targets)} targets, same ChaCha8 seeds for both solvers")
    print(f"  scipy SLSQP (stand-in for NLopt SLSQP): success/attempt {s_ok / n:.3f}  evals/attempt {s_ev / n:.1f}  "
          f"evals/converged attempt {s_ev / max(s_ok, 1):.1f}")
    print(f"  LM twin (what --impl reference times):   success/attempt {l_ok / n:.3f}  evals/attempt {l_ev / n:.1f}  "
          f"evals/converged attempt {l_ev / max(l_ok, 1):.1f}")
    print(f"  => an SLSQP-class CPU loop needs {(s_ev / max(s_ok, 1)) / (l_ev / max(l_ok, 1)):.2f}x the objective evaluations per "
          f"solve of the port that bench.py times (plus SLSQP's own O(n^3) QP work per iteration)")


if __name__ == "__main__":
    main()
