"""SLSQP-class stand-in for the reference's inner solver -- TEST INFRASTRUCTURE ONLY.

The reference runs NLopt's SLSQP per restart (crates/optik/src/lib.rs:302-356, 372), an un-vendored dependency
(nlopt 0.8.1, kylc/rust-nlopt@8e731e3) that cannot be built in this image.  scipy's SLSQP is the same Kraft algorithm;
it is driven here over the golden-pinned oracle objective / gradient exactly as the reference configures NLopt:
    stopval  = tol_f          (lib.rs:345)   -> success the moment an evaluation has f < tol_f (raised from the callback)
    ftol_abs = 1e-3 * tol_f   (lib.rs:283-293, 346): a stall, which the default config counts as a FAILED attempt (:378)
    bounds   = joint limits   (lib.rs:348-349)
and the restart loop of lib.rs:360-413 in single-thread order (Speed: lowest-index converged restart).
"""
import numpy as np
from scipy.optimize import minimize


class _Reached(Exception):
    pass


def slsqp_attempt(ch, tgt, q0, tol_f=1e-6, ftol=None, maxiter=200):
    """One restart attempt.  -> (converged, q, evaluations); q = best point seen."""
    best = [np.inf, np.array(q0, dtype=float)]
    evals = [0]

    def fun(q):
        evals[0] += 1
        f = ch.objective(q, tgt)
        if f < best[0]:
            best[0], best[1] = f, q.copy()
        if f < tol_f:
            raise _Reached()
        return f

    try:
        minimize(fun, np.clip(q0, ch.lb, ch.ub), jac=lambda q: ch.objective_grad(q, tgt), method="SLSQP",
                 bounds=list(zip(ch.lb, ch.ub)), options={"ftol": 1e-3 * tol_f if ftol is None else ftol, "maxiter": maxiter})
    except _Reached:
        return True, best[1], evals[0]
    return False, best[1], evals[0]


def slsqp_ik(ch, tgt, x0, restarts, tol_f=1e-6):
    """Robot::ik, Speed mode, one thread: restart 0 = x0, i >= 1 = the reference's ChaCha8 seeds (lib.rs:360-370).
    -> (found, q, restart index, evaluations)."""
    total = 0
    for r in range(restarts):
        ok, q, ev = slsqp_attempt(ch, tgt, x0 if r == 0 else ch.restart_seed(r), tol_f)
        total += ev
        if ok:
            return True, q, r, total
    return False, None, restarts, total


def slsqp_polish(ch, tgt, q, tol):
    """Continue from q until f < tol (no stall rule).  -> (reached, q)."""
    ok, qq, _ = slsqp_attempt(ch, tgt, q, tol_f=tol, ftol=1e-40, maxiter=400)
    return ok, qq


def row_space_part(ch, q, dq):
    """Component of a joint-space difference that changes the pose to first order (J^+ J dq): for a redundant arm two
    answers may differ along the self-motion direction (null space of J) without being different solutions."""
    J = ch.joint_jacobian(q)
    return np.linalg.pinv(J) @ (J @ dq)
