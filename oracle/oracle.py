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
