"""ctypes front-end to oracle/_build/liboptik_oracle.so -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None
dp = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.oracle_objective.restype = C.c_double
        _lib.oracle_uniform.restype = C.c_double
        _lib.oracle_uniform.argtypes = [C.c_uint64, C.c_double, C.c_double]
        _lib.oracle_rng_u64.restype = C.c_uint64
        _lib.oracle_rng_u64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib.oracle_restart_seed.argtypes = [C.c_uint64, dp, dp, C.c_int, dp]
        _lib.oracle_chacha8_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    return _lib


IDENT = np.array([0, 0, 0, 1, 0, 0, 0, 0], dtype=np.float64)


def pose8(quat_xyzw, t):
    return np.array(list(quat_xyzw) + list(t) + [0.0], dtype=np.float64)


def pose8_from_matrix(M):
    """4x4 homogeneous (row-major numpy) -> pose8.  Shepperd's method."""
    M = np.asarray(M, dtype=np.float64)
    R = M[:3, :3]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    return pose8(q, M[:3, 3])


def pose8_to_matrix(p):
    x, y, z, w = p[:4]
    M = np.eye(4)
    M[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    M[:3, 3] = p[4:7]
    return M


class Chain:
    """Flat chain (njoints x 16) + the reference-style evaluator on it."""

    def __init__(self, arr):
        self.arr = _d(arr)
        self.njoints = self.arr.shape[0]
        self.n = int(np.sum(self.arr[:, 3] != 2))
        art = self.arr[self.arr[:, 3] != 2]
        self.lb = art[:, 12].copy()
        self.ub = art[:, 13].copy()

    @classmethod
    def from_urdf(cls, text, base, ee, **kw):
        from .urdf_chain import chain_from_urdf
        return cls(chain_from_urdf(text, base, ee, **kw))

    # crates/optik/src/lib.rs:93-99 + kinematics.rs:123-164
    def fk(self, q, ee_offset=None):
        q = _d(q)
        assert q.shape == (self.n,), "generalized position vector `q` is of incorrect length"
        tf = np.zeros((self.njoints, 8))
        ee = np.zeros(8)
        lib().oracle_fk(_p(self.arr), self.njoints, _p(q), _p(None if ee_offset is None else _d(ee_offset)), _p(tf), _p(ee))
        return tf, ee

    def joint_jacobian(self, q, ee_offset=None):
        tf, ee = self.fk(q, ee_offset)
        J = np.zeros((self.n, 6))  # column-major 6 x n == row-major n x 6
        lib().oracle_joint_jacobian(_p(self.arr), self.njoints, _p(tf), _p(ee), _p(J))
        return J.T.copy()

    def objective(self, q, target, wl=(1, 1, 1), wa=(1, 1, 1), ee_offset=None):
        return lib().oracle_objective(_p(self.arr), self.njoints, _p(_d(q)), _p(_d(target)),
                                      _p(None if ee_offset is None else _d(ee_offset)), _p(_d(wl)), _p(_d(wa)))

    def objective_grad(self, q, target, wl=(1, 1, 1), wa=(1, 1, 1), ee_offset=None):
        g = np.zeros(self.n)
        lib().oracle_objective_grad(_p(self.arr), self.njoints, _p(_d(q)), _p(_d(target)),
                                    _p(None if ee_offset is None else _d(ee_offset)), _p(_d(wl)), _p(_d(wa)), _p(g))
        return g

    # crates/optik/src/lib.rs:123-239 (the LP solved exactly, see diffik_oracle.c)
    def diff_ik(self, x0, V_WE, v_max, ee_offset=None):
        """-> (alpha, v).  The LP is always feasible: at a singular configuration an unreachable twist gives alpha = 0."""
        alpha = C.c_double(0.0)
        v = np.zeros(self.n)
        rc = lib().oracle_diff_ik(_p(self.arr), self.njoints, _p(None if ee_offset is None else _d(ee_offset)), _p(_d(x0)),
                                  _p(_d(V_WE)), _p(_d(v_max)), C.byref(alpha), _p(v))
        if rc < 0:
            raise ValueError(f"oracle_diff_ik: unsupported input ({rc})")
        return (alpha.value, v) if rc == 1 else None

    def restart_seed(self, restart):
        q = np.zeros(self.n)
        lib().oracle_restart_seed(int(restart), _p(self.lb), _p(self.ub), self.n, _p(q))
        return q


def so3_log(q):
    w = np.zeros(3)
    lib().oracle_so3_log(_p(_d(q)), _p(w))
    return w


def se3_log(p8):
    e = np.zeros(6)
    lib().oracle_se3_log(_p(_d(p8)), _p(e))
    return e


def so3_right_jacobian(w):
    J = np.zeros(9)
    lib().oracle_so3_right_jacobian(_p(_d(w)), _p(J))
    return J  # column-major


def se3_right_jacobian(p8):
    U = np.zeros(36)
    lib().oracle_se3_right_jacobian(_p(_d(p8)), _p(U))
    return U  # column-major


# ---------------------------------------------------------------- solver twin
ST_NAMES = {0: "none", 1: "stopval", 2: "ftol", 3: "xtol", 4: "itercap", 5: "stuck", 6: "nan", 7: "skipped"}


class TwinParams(C.Structure):
    _fields_ = [("tol_f", C.c_double), ("tol_df_eff", C.c_double), ("tol_df_user", C.c_double), ("tol_dx", C.c_double),
                ("wl", C.c_double * 3), ("wa", C.c_double * 3), ("weighted", C.c_int), ("max_evals", C.c_int),
                ("lambda0", C.c_double), ("lambda_dec", C.c_double), ("lambda_inc", C.c_double),
                ("lambda_min", C.c_double), ("lambda_max", C.c_double),
                ("stall_rel", C.c_double), ("stall_count", C.c_int), ("layout", C.c_int)]


# defaults of the kernel's LM iteration (DESIGN.md "Solver"); keep in sync with optik_b200/csrc/solver_params.h
LM_DEFAULTS = dict(max_evals=24, lambda0=1e-1, lambda_dec=0.3, lambda_inc=10.0, lambda_min=1e-9, lambda_max=1e6,
                   stall_rel=1e-1, stall_count=2, layout=0)


def twin_params(tol_f=1e-6, tol_df=-1.0, tol_dx=-1.0, wl=(1, 1, 1), wa=(1, 1, 1), **lm):
    """SolverConfig -> kernel parameters, as crates/optik/src/lib.rs:283-293 derives them."""
    d = dict(LM_DEFAULTS)
    d.update(lm)
    p = TwinParams()
    p.tol_f = tol_f
    p.tol_df_eff = tol_df if tol_df > 0.0 else 1e-3 * tol_f
    p.tol_df_user = tol_df
    p.tol_dx = tol_dx
    p.wl[:] = list(wl)
    p.wa[:] = list(wa)
    p.weighted = int(any(float(x) != 1.0 for x in list(wl) + list(wa)))
    for k, v in d.items():
        setattr(p, k, v)
    return p


def twin_eval(chain, q, target, params=None, ee_offset=None):
    params = params or twin_params()
    ee = np.zeros(8)
    f = C.c_double()
    r = np.zeros(6)
    Jr = np.zeros((chain.n, 6))
    g = np.zeros(chain.n)
    rc = lib().twin_eval_c(_p(chain.arr), chain.njoints, _p(None if ee_offset is None else _d(ee_offset)),
                           C.byref(params), _p(_d(target)), _p(_d(q)), _p(ee), C.byref(f), _p(r), _p(Jr), _p(g))
    assert rc == 0, rc
    return dict(ee=ee, f=f.value, r=r, Jr=Jr, grad=g)


def twin_attempts(chain, target, x0, r_begin, r_end, params=None, ee_offset=None):
    """Per-restart records for restarts [r_begin, r_end): q (R x n), f, status, evals."""
    params = params or twin_params()
    R = int(r_end - r_begin)
    q = np.zeros((R, chain.n))
    f = np.zeros(R)
    st = np.zeros(R, dtype=np.int32)
    ev = np.zeros(R, dtype=np.int32)
    lib().twin_attempts_c.argtypes = [dp, C.c_int, dp, C.POINTER(TwinParams), dp, dp, C.c_uint64, C.c_uint64, dp, dp,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = lib().twin_attempts_c(_p(chain.arr), chain.njoints, _p(None if ee_offset is None else _d(ee_offset)),
                               C.byref(params), _p(_d(target)), _p(_d(x0)), int(r_begin), int(r_end), _p(q), _p(f),
                               st.ctypes.data_as(C.POINTER(C.c_int)), ev.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0, rc
    return q, f, st, ev


def twin_ik(chain, target, x0, r_begin, r_end, mode="speed", params=None, ee_offset=None):
    """Reference selection semantics (lib.rs:397-413, single-threaded order) over the twin's attempts.

    Returns dict(found, q, f, status, restart).  Speed: lowest-index converged restart; Quality: arg-min
    ||q - x0||^2 (ties -> lower index) over all converged restarts in [r_begin, r_end)."""
    params = params or twin_params()
    n = chain.n
    q = np.zeros(n)
    f = C.c_double()
    st = C.c_int()
    rs = C.c_uint64()
    lib().twin_ik_c.argtypes = [dp, C.c_int, dp, C.POINTER(TwinParams), dp, dp, C.c_uint64, C.c_uint64, C.c_int, dp,
                                C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    found = lib().twin_ik_c(_p(chain.arr), chain.njoints, _p(None if ee_offset is None else _d(ee_offset)),
                            C.byref(params), _p(_d(target)), _p(_d(x0)), int(r_begin), int(r_end),
                            2 if mode == "speed" else 1, _p(q), C.byref(f), C.byref(st), C.byref(rs))
    assert found >= 0, found
    return dict(found=bool(found), q=q, f=f.value, status=st.value, restart=rs.value)


def ref_ik_threaded(chain, target, x0, r_begin, r_end, mode="quality", threads=1, params=None, ee_offset=None):
    """The reference's parallel restart driver (lib.rs:297-413) over the twin solver, `threads` pthread workers.
    Returns dict(found, q, f, restart, attempts, evals, converged)."""
    params = params or twin_params()
    q = np.zeros(chain.n)
    f = C.c_double()
    rs = C.c_uint64()
    stats = (C.c_uint64 * 3)()
    fn = lib().ref_ik_threaded
    fn.argtypes = [dp, C.c_int, dp, C.POINTER(TwinParams), dp, dp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, dp,
                   C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    found = fn(_p(chain.arr), chain.njoints, _p(None if ee_offset is None else _d(ee_offset)), C.byref(params),
               _p(_d(target)), _p(_d(x0)), int(r_begin), int(r_end), 2 if mode == "speed" else 1, int(threads), _p(q),
               C.byref(f), C.byref(rs), stats)
    assert found >= 0
    return dict(found=bool(found), q=q, f=f.value, restart=rs.value, attempts=stats[0], evals=stats[1], converged=stats[2])


def ref_batch_threaded(chain, targets, x0, restarts, mode="speed", threads=1, params=None):
    """A batch of independent targets, one target per worker thread (Robot::ik with set_parallelism(1) semantics per
    target).  Returns (q (T, n), f (T,), found (T,) bool)."""
    params = params or twin_params(layout=1 if chain.n <= 8 else 0)
    targets, x0 = _d(targets), _d(x0)
    T = targets.shape[0]
    q = np.zeros((T, chain.n))
    f = np.zeros(T)
    found = np.zeros(T, dtype=np.int32)
    fn = lib().ref_batch_threaded
    fn.argtypes = [dp, C.c_int, C.POINTER(TwinParams), dp, dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, dp, dp,
                   C.POINTER(C.c_int)]
    rc = fn(_p(chain.arr), chain.njoints, C.byref(params), _p(targets), _p(x0), chain.n, T, int(restarts),
            2 if mode == "speed" else 1, int(threads), _p(q), _p(f), found.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return q, f, found.astype(bool)


def eval_batch_threaded(chain, q, targets, threads=1, out=None):
    """Reference-style evaluator over a batch (FK + body Jacobian + objective + gradient per configuration, the
    reference's own call sequence) on `threads` pthread workers: bench.py's CPU number beside the evaluator kernel.
    Returns dict(ee (B,8), jac (B,n,6) = column-major 6 x n per configuration, f (B,), grad (B,n))."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    targets = np.ascontiguousarray(targets, dtype=np.float64)
    B, n = q.shape
    assert n == chain.n and targets.shape == (B, 8)
    if out is None:  # (pass the dict of a previous call to time the arithmetic without first-touch page faults)
        out = dict(ee=np.zeros((B, 8)), jac=np.zeros((B, n, 6)), f=np.zeros(B), grad=np.zeros((B, n)))
    ee, jac, f, grad = out["ee"], out["jac"], out["f"], out["grad"]
    fn = lib().oracle_eval_batch_threaded
    fn.argtypes = [dp, C.c_int, dp, dp, C.c_uint64, C.c_int, dp, dp, dp, dp]
    fn.restype = C.c_int
    rc = fn(_p(chain.arr), chain.njoints, _p(q), _p(targets), B, int(threads), _p(ee), _p(jac), _p(f), _p(grad))
    assert rc == 0, rc
    return out
