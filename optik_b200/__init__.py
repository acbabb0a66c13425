"""optik_b200 -- B200-native batched inverse kinematics behind OptIK's Robot / SolverConfig / ik() surface.

Host-side mirror of the reference's Python module (kylc/optik @ 355e463):
    optik.pyi:9-49, crates/optik-py/src/lib.rs:17-155   -> SolverConfig, Robot (same names, arguments, defaults,
                                                           4x4 row-major nested-list poses, None for "no solution")
implemented over the C ABI of ``lib/liboptik_b200.so`` (include/optik_b200.h) with ctypes.  All numerics run in the
CUDA kernels of ``csrc/``; there is no CPU fallback -- importing works without a GPU (model loading, limits), any
numerical call without a usable CUDA device raises ``OptikError``.

Additions over the reference surface: ``Robot.ik_batch`` / ``ik_attempts`` / ``eval_batch`` (numpy arrays = host path
with copies inside the call; torch CUDA tensors = zero-copy device path on the current torch stream).
"""
import ctypes as C
import os

import numpy as np

__all__ = ["Robot", "SolverConfig", "OptikError", "load_library", "data_path", "STATUS_NAMES"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liboptik_b200.so")
_lib = None

STATUS_NAMES = {0: "none", 1: "stopval", 2: "ftol", 3: "xtol", 4: "itercap", 5: "stuck", 6: "nan", 7: "skipped"}
U64_MAX = 2 ** 64 - 1


class OptikError(RuntimeError):
    pass


class _CSolverConfig(C.Structure):  # == CSolverConfig, crates/optik-cpp/src/lib.rs:10-20
    _fields_ = [("solution_mode", C.c_int32), ("max_time", C.c_double), ("max_restarts", C.c_ulong),
                ("tol_f", C.c_double), ("tol_df", C.c_double), ("tol_dx", C.c_double),
                ("linear_weight", C.c_double * 3), ("angular_weight", C.c_double * 3)]


class _BatchOpts(C.Structure):  # == optik_gpu_batch_opts
    _fields_ = [("struct_size", C.c_uint32), ("restarts", C.c_uint32), ("restart_begin", C.c_uint64),
                ("chunks", C.c_uint32), ("tile", C.c_uint32), ("max_evals", C.c_uint32), ("blocks", C.c_uint32),
                ("memory", C.c_int32), ("ee_offset", C.c_void_p), ("restart_out", C.c_void_p),
                ("evals_out", C.c_void_p), ("counters", C.c_void_p), ("best_record_out", C.c_void_p),
                ("flags", C.c_uint32), ("variant", C.c_uint32), ("push_peers", C.c_void_p), ("push_rank", C.c_uint32),
                ("push_world", C.c_uint32), ("push_seq", C.c_uint64)]


BATCH_ASYNC, BATCH_STATIC = 1, 8  # == OPTIK_BATCH_*
STATUS_CODE_MASK, STATUS_FLAG_SEED_CLAMPED = 0xff, 0x100  # == OPTIK_STATUS_*


RECORD_HEAD = 8  # candidate record: [found, score, restart, cost, status, 0, 0, 0, q...]


def data_path(name):
    """Path of a bundled kinematics-only URDF (panda, ur5, ur3e, snake20)."""
    return os.path.join(_HERE, "data", name if name.endswith(".urdf") else name + ".urdf")


ROBOT_LINKS = {"panda": ("panda_link0", "panda_link8"), "ur5": ("base_link", "ee_link"),
               "ur3e": ("ur_base_link", "ur_ee_link"), "snake20": ("seg0", "tip")}


def load_library():
    """dlopen liboptik_b200.so; fails loudly when it has not been built (python -m optik_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OptikError(f"{LIB_PATH} is missing: build it with `python -m optik_b200.build` "
                         "(there is no CPU fallback for the IK path)")
    lib = C.CDLL(LIB_PATH)
    vp, dp, cp = C.c_void_p, C.POINTER(C.c_double), C.c_char_p
    sig = {
        "optik_robot_from_urdf_file": (vp, [cp, cp, cp]),
        "optik_robot_from_urdf_str": (vp, [cp, cp, cp]),
        "optik_robot_try_from_urdf_str": (vp, [cp, cp, cp]),
        "optik_robot_from_chain": (vp, [dp, C.c_uint]),
        "optik_robot_free": (None, [vp]),
        "optik_robot_set_parallelism": (None, [vp, C.c_uint]),
        "optik_robot_num_positions": (C.c_uint, [vp]),
        "optik_robot_num_joints": (C.c_uint, [vp]),
        "optik_robot_chain": (C.c_int, [vp, dp]),
        "optik_robot_set_device": (C.c_int, [vp, C.c_int]),
        "optik_robot_joint_limits": (vp, [vp]),
        "optik_robot_random_configuration": (vp, [vp]),
        "optik_robot_joint_jacobian": (vp, [vp, dp]),
        "optik_robot_fk": (vp, [vp, dp]),
        "optik_robot_ik": (vp, [vp, C.POINTER(_CSolverConfig), dp, dp]),
        "optik_robot_ik_ex": (C.c_int, [vp, C.POINTER(_CSolverConfig), dp, dp, dp, dp, dp]),
        "optik_robot_diff_ik": (vp, [vp, dp, dp, dp]),
        "optik_last_error": (cp, []),
        "optik_set_urdf_correct_fold": (None, [C.c_int]),
        "optik_status_is_success": (C.c_int, [C.POINTER(_CSolverConfig), C.c_int]),
        "optik_gpu_ik_batch": (C.c_int, [vp, C.POINTER(_CSolverConfig), C.POINTER(_BatchOpts), vp, vp, C.c_uint64, vp, vp, vp, vp]),
        "optik_gpu_ik_attempts": (C.c_int, [vp, C.POINTER(_CSolverConfig), C.POINTER(_BatchOpts), vp, vp, vp, vp, vp, vp, vp]),
        "optik_gpu_eval_batch": (C.c_int, [vp, vp, vp, C.c_int, C.c_uint64, dp, dp, dp, C.c_int, vp, vp, vp, vp, vp]),
        "optik_gpu_select_records": (C.c_int, [vp, vp, C.c_uint32, vp, vp]),
        "optik_gpu_exchange_bytes": (C.c_uint64, [vp, C.c_uint32]),
        "optik_gpu_exchange_push": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp]),
        "optik_gpu_exchange_select": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64, vp, vp]),
        "optik_gpu_restart_seeds": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_int, vp, vp]),
        "optik_gpu_chacha8_block": (C.c_int, [vp, vp, C.c_uint64, vp]),
        "optik_gpu_diff_ik_batch": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int, C.c_uint64, vp, C.c_int, vp, vp, vp, vp]),
        "optik_robot_diff_ik_ex": (C.c_int, [vp, vp, vp, vp, vp, C.POINTER(C.c_double), vp]),
        "optik_gpu_stream_create": (C.c_int, [vp, C.POINTER(vp)]),
        "optik_gpu_stream_sync": (C.c_int, [vp]),
        "optik_gpu_stream_destroy": (None, [vp]),
        "optik_measure_fp64_peak": (C.c_double, [C.c_int, C.c_double]),
        "optik_measure_hbm_mix": (C.c_double, [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]),
        "optik_host_alloc": (vp, [C.c_uint64]),
        "optik_host_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.free = C.CDLL(None).free
    lib.free.argtypes = [vp]
    lib.free.restype = None
    _lib = lib
    return lib


def _err():
    return load_library().optik_last_error().decode("utf-8", "replace")


def _check(rc):
    if rc != 0:
        raise OptikError(f"optik_b200 error {rc}: {_err()}")


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _take(ptr, count):
    """Copy a malloc()ed double buffer returned by the C ABI and free() it (crates/optik-cpp/src/lib.cpp:71-73)."""
    lib = load_library()
    out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,)).copy()
    lib.free(ptr)
    return out


def _pose8_from_rows(m):
    """4x4 row-major nested list / array -> pose8 (crates/optik-py/src/lib.rs:8-15 parse_pose)."""
    M = np.asarray(m, dtype=np.float64)
    if M.shape != (4, 4):
        raise ValueError("invalid target transform specified")
    R = M[:3, :3]
    if not (np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.allclose(M[3], [0, 0, 0, 1], atol=1e-9)):
        raise ValueError("invalid target transform specified")  # nalgebra::try_convert::<Isometry3> fails
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
    q = np.asarray(q) / np.linalg.norm(q)
    return np.array([q[0], q[1], q[2], q[3], M[0, 3], M[1, 3], M[2, 3], 0.0])


def _rows_from_pose8(p):
    x, y, z, w = p[:4]
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), p[4]],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), p[5]],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), p[6]],
            [0.0, 0.0, 0.0, 1.0]]


class SolverConfig:
    """optik.pyi:9-20; defaults of crates/optik-py/src/lib.rs:24-31 (== config.rs:52-65)."""

    def __init__(self, solution_mode="speed", max_time=0.1, max_restarts=U64_MAX, tol_f=1e-6, tol_df=-1.0,
                 tol_dx=-1.0, linear_weight=(1.0, 1.0, 1.0), angular_weight=(1.0, 1.0, 1.0)):
        if solution_mode not in ("speed", "quality"):
            raise ValueError("invalid solution mode")  # crates/optik-py/src/lib.rs:43
        if max_time == 0.0 and max_restarts == 0:
            # crates/optik-py/src/lib.rs:45-47
            raise ValueError("no time or restart limit applied -- solver would run forever")
        self.solution_mode = solution_mode
        self.max_time = float(max_time)
        self.max_restarts = int(max_restarts)
        self.tol_f = float(tol_f)
        self.tol_df = float(tol_df)
        self.tol_dx = float(tol_dx)
        self.linear_weight = [float(x) for x in linear_weight]
        self.angular_weight = [float(x) for x in angular_weight]
        if len(self.linear_weight) != 3 or len(self.angular_weight) != 3:
            raise ValueError("weights must have 3 components")

    def _c(self):
        c = _CSolverConfig()
        c.solution_mode = 1 if self.solution_mode == "quality" else 2  # config.rs:5-8
        c.max_time = self.max_time
        c.max_restarts = 0 if self.max_restarts >= U64_MAX else self.max_restarts  # 0 == unlimited (lib.rs:273-277)
        c.tol_f, c.tol_df, c.tol_dx = self.tol_f, self.tol_df, self.tol_dx
        c.linear_weight[:] = self.linear_weight
        c.angular_weight[:] = self.angular_weight
        return c

    def is_success(self, status):
        """lib.rs:376-379 applied to a status code array."""
        st = np.asarray(status) & STATUS_CODE_MASK
        return ((self.tol_f >= 0) & (st == 1)) | ((self.tol_df >= 0) & (st == 2)) | ((self.tol_dx >= 0) & (st == 3))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Stream:
    """A CUDA stream owned by the library (optik_gpu_stream_*), for pipelined host-buffer calls:
    `robot.ik_attempts(..., stream=s, wait=False)` enqueues copies + kernels and returns; `s.synchronize()` waits.
    A torch.cuda.Stream can be passed to the same arguments instead."""

    def __init__(self, robot):
        h = C.c_void_p()
        _check(load_library().optik_gpu_stream_create(robot._h, C.byref(h)))
        self.handle = h.value

    def synchronize(self):
        _check(load_library().optik_gpu_stream_sync(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load_library().optik_gpu_stream_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _stream_handle(stream):
    if stream is None:
        return None
    if isinstance(stream, Stream):
        return stream.handle
    if hasattr(stream, "cuda_stream"):  # torch.cuda.Stream
        return stream.cuda_stream
    return int(stream)


class _PinnedBlock:
    """Owner of one optik_host_alloc block; numpy views keep it alive through their .base chain."""

    def __init__(self, nbytes):
        self._lib = load_library()
        self.ptr = self._lib.optik_host_alloc(max(int(nbytes), 1))
        if not self.ptr:
            raise OptikError(_err())
        self.__array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                self._lib.optik_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array over pinned host memory (optik_host_alloc): copies to/from it are truly asynchronous."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    count = int(np.prod(shape))
    nbytes = count * np.dtype(dtype).itemsize
    raw = np.asarray(_PinnedBlock(nbytes))
    return raw[:nbytes].view(dtype).reshape(shape)


class Robot:
    """optik.pyi:22-49 over the C ABI; plus batched entry points."""

    def __init__(self, handle):
        if not handle:
            raise OptikError(_err())
        self._h = handle
        self._n = load_library().optik_robot_num_positions(handle)
        lim = _take(load_library().optik_robot_joint_limits(handle), 2 * self._n)
        self._lb, self._ub = lim[:self._n].copy(), lim[self._n:].copy()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.optik_robot_free(h)

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def from_urdf_file(path, base_link, ee_link):
        with open(path, "r") as f:  # a missing file raises here instead of aborting the process
            text = f.read()
        return Robot.from_urdf_str(text, base_link, ee_link)

    @staticmethod
    def from_urdf_str(urdf, base_link, ee_link):
        lib = load_library()
        return Robot(lib.optik_robot_try_from_urdf_str(urdf.encode(), base_link.encode(), ee_link.encode()))

    @staticmethod
    def from_chain(chain):
        """Robot::new(KinematicChain) (lib.rs:42-47) from a flat (njoints, 16) chain array."""
        a = np.ascontiguousarray(chain, dtype=np.float64).reshape(-1, 16)
        return Robot(load_library().optik_robot_from_chain(_dptr(a), a.shape[0]))

    @staticmethod
    def named(name):
        base, ee = ROBOT_LINKS[name]
        return Robot.from_urdf_file(data_path(name), base, ee)

    # ---- reference surface --------------------------------------------------------------------
    def set_parallelism(self, n):
        load_library().optik_robot_set_parallelism(self._h, int(n))

    def set_device(self, device):
        _check(load_library().optik_robot_set_device(self._h, int(device)))

    def num_positions(self):
        return int(self._n)

    def joint_limits(self):
        return list(self._lb), list(self._ub)

    def chain(self):
        nj = load_library().optik_robot_num_joints(self._h)
        out = np.zeros((nj, 16))
        _check(load_library().optik_robot_chain(self._h, _dptr(out)))
        return out

    def random_configuration(self):
        return list(_take(load_library().optik_robot_random_configuration(self._h), self._n))

    def _x(self, x, what="x0"):
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        if x.shape[0] != self._n:
            raise ValueError(f"len({what}) != num_positions")  # crates/optik-py/src/lib.rs:94,106,127
        return x

    def fk(self, x, ee_offset=None):
        x = self._x(x, "x")
        out = self.eval_batch(x[None, :], ee_offset=ee_offset, want=("ee",))
        return _rows_from_pose8(out["ee"][0])

    def joint_jacobian(self, x, ee_offset=None):
        x = self._x(x, "x")
        out = self.eval_batch(x[None, :], ee_offset=ee_offset, want=("jac",))
        return [list(r) for r in out["jac"][0].reshape(self._n, 6).T]  # column-major 6 x n -> 6 row lists

    def ik(self, config, target, x0, ee_offset=None):
        """-> (q, cost) or None (crates/optik-py/src/lib.rs:117-132)."""
        x0 = self._x(x0)
        if np.any(x0 < self._lb) or np.any(x0 > self._ub):
            raise ValueError("seed joint position outside of joint limits")  # lib.rs:251-254
        tgt = _pose8_from_rows(target)
        eo = None if ee_offset is None else _pose8_from_rows(ee_offset)
        c = config._c()
        q = np.zeros(self._n)
        cost = C.c_double()
        rc = load_library().optik_robot_ik_ex(self._h, C.byref(c), _dptr(tgt), _dptr(x0), None if eo is None else _dptr(eo),
                                              _dptr(q), C.cast(C.byref(cost), C.POINTER(C.c_double)))
        if rc < 0:
            raise OptikError(f"optik_b200 error {-rc}: {_err()}")
        return (list(q), cost.value) if rc == 1 else None

    def diff_ik(self, x0, V_WE, v_max, ee_offset=None):
        """optik.pyi:39-49 / crates/optik-py/src/lib.rs:133-155: (alpha, v) or None."""
        x0 = self._x(x0)
        v_max = self._x(v_max, "v_max")
        V = np.ascontiguousarray(V_WE, dtype=np.float64).ravel()
        if V.shape[0] != 6:
            raise ValueError("V_WE must have 6 entries [linear; angular]")
        eo = None
        if ee_offset is not None:
            eo = np.ascontiguousarray(ee_offset, dtype=np.float64)
            eo = _pose8_from_rows(eo) if eo.shape == (4, 4) else eo.ravel()
        alpha = C.c_double(0.0)
        v = np.zeros(self._n)
        rc = load_library().optik_robot_diff_ik_ex(self._h, x0.ctypes.data, V.ctypes.data, v_max.ctypes.data,
                                                   None if eo is None else eo.ctypes.data, C.byref(alpha), v.ctypes.data)
        if rc < 0:
            raise OptikError(_err())
        return (alpha.value, [float(t) for t in v]) if rc == 1 else None

    def diff_ik_batch(self, x0, V_WE, v_max, ee_offset=None):
        """diff_ik over B configurations in one launch.  x0 (B, n); V_WE (B, 6) or (6,); v_max (B, n) or (n,).
        numpy -> host path, returns numpy (alpha (B,), v (B, n), status (B,) 1 = solved / 0 = none);
        torch CUDA tensors -> device path on the current stream."""
        lib = load_library()
        n = self._n
        eo = None
        if ee_offset is not None:
            eo = np.ascontiguousarray(ee_offset, dtype=np.float64)
            eo = _pose8_from_rows(eo) if eo.shape == (4, 4) else eo.ravel()
        eop = None if eo is None else eo.ctypes.data
        if _is_torch(x0):
            import torch
            B, dev = x0.shape[0], x0.device
            for t in (x0, V_WE, v_max):
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
            assert x0.shape == (B, n) and V_WE.shape in ((B, 6), (6,)) and v_max.shape in ((B, n), (n,))
            alpha = torch.empty((B,), dtype=torch.float64, device=dev)
            v = torch.empty((B, n), dtype=torch.float64, device=dev)
            st = torch.empty((B,), dtype=torch.int32, device=dev)
            _check(lib.optik_gpu_diff_ik_batch(self._h, x0.data_ptr(), V_WE.data_ptr(), int(V_WE.dim() == 1), v_max.data_ptr(),
                                               int(v_max.dim() == 1), B, eop, 1, alpha.data_ptr(), v.data_ptr(), st.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
            return alpha, v, st
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        V = np.ascontiguousarray(V_WE, dtype=np.float64)
        vm = np.ascontiguousarray(v_max, dtype=np.float64)
        B = x0.shape[0]
        if x0.shape != (B, n) or V.shape not in ((B, 6), (6,)) or vm.shape not in ((B, n), (n,)):
            raise ValueError("x0 must be (B, n); V_WE (B, 6) or (6,); v_max (B, n) or (n,)")
        alpha, v, st = np.zeros(B), np.zeros((B, n)), np.zeros(B, dtype=np.int32)
        _check(lib.optik_gpu_diff_ik_batch(self._h, x0.ctypes.data, V.ctypes.data, int(V.ndim == 1), vm.ctypes.data,
                                           int(vm.ndim == 1), B, eop, 0, alpha.ctypes.data, v.ctypes.data, st.ctypes.data, None))
        return alpha, v, st

    # ---- batched additions --------------------------------------------------------------------
    def _opts(self, memory, restarts=0, restart_begin=0, chunks=0, tile=0, max_evals=0, blocks=0, ee_offset=None,
              variant=0):
        o = _BatchOpts()
        o.variant = int(variant)
        o.struct_size = C.sizeof(_BatchOpts)
        o.restarts, o.restart_begin, o.chunks, o.tile = int(restarts), int(restart_begin), int(chunks), int(tile)
        o.max_evals, o.blocks, o.memory = int(max_evals), int(blocks), int(memory)
        keep = []
        if ee_offset is not None:
            eo = np.ascontiguousarray(ee_offset, dtype=np.float64)
            eo = _pose8_from_rows(eo) if eo.shape == (4, 4) else eo.ravel()
            keep.append(eo)
            o.ee_offset = eo.ctypes.data
        return o, keep

    def ik_batch(self, config, targets, x0, restarts=None, restart_begin=0, chunks=0, tile=0, max_evals=0, blocks=0,
                 ee_offset=None, stats=False, out=None, stream=None, wait=True, static=False, variant=0):
        """Robot::ik over T (target, x0) pairs in one launch.

        targets: (T, 8) pose8 rows {qx,qy,qz,qw,tx,ty,tz,0};  x0: (T, n).
        numpy in  -> host path (H2D/D2H inside the call), returns numpy (q, cost, status[, extra]).
        torch CUDA tensors in -> device path on torch's current stream, returns torch tensors (no sync).
        Host path with stream=<Stream> and wait=False: returns as soon as everything is enqueued; the outputs (give
        pinned `out` buffers, see pinned_empty) are valid after stream.synchronize().
        restarts=None: config.max_restarts; with the reference's default (no restart limit) restarts are drawn until
        config.max_time expires (lib.rs:260-277).
        static=True forces the static (target, chunk) schedule for Speed batches (OPTIK_BATCH_STATIC) instead of the
        dynamic chains -- same q / cost / status / winning restart, reproducible `evals`.
        """
        lib = load_library()
        c = config._c()
        if restarts is None:
            restarts = 0
        n = self._n
        if _is_torch(targets):
            import torch
            T = targets.shape[0]
            assert targets.is_cuda and x0.is_cuda and targets.dtype == torch.float64 and x0.dtype == torch.float64
            assert targets.is_contiguous() and x0.is_contiguous() and targets.shape == (T, 8) and x0.shape == (T, n)
            dev = targets.device
            if out is None:
                q = torch.empty((T, n), dtype=torch.float64, device=dev)
                f = torch.empty((T,), dtype=torch.float64, device=dev)
                st = torch.empty((T,), dtype=torch.int32, device=dev)
            else:
                q, f, st = out
            o, keep = self._opts(1, restarts, restart_begin, chunks, tile, max_evals, blocks, ee_offset, variant)
            if static:
                o.flags |= BATCH_STATIC
            extra = {}
            if stats:
                extra["restart"] = torch.empty((T,), dtype=torch.int64, device=dev)
                extra["evals"] = torch.empty((T,), dtype=torch.int32, device=dev)
                extra["counters"] = torch.zeros((3,), dtype=torch.int64, device=dev)
                o.restart_out, o.evals_out, o.counters = extra["restart"].data_ptr(), extra["evals"].data_ptr(), extra["counters"].data_ptr()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib.optik_gpu_ik_batch(self._h, C.byref(c), C.byref(o), targets.data_ptr(), x0.data_ptr(), T,
                                          q.data_ptr(), f.data_ptr(), st.data_ptr(), stream))
            return (q, f, st, extra) if stats else (q, f, st)
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        T = targets.shape[0]
        if targets.shape != (T, 8) or x0.shape != (T, n):
            raise ValueError("targets must be (T, 8) pose8 and x0 (T, num_positions)")
        if out is None:
            q, f, st = np.empty((T, n)), np.empty(T), np.empty(T, dtype=np.int32)
        else:
            q, f, st = out
        o, keep = self._opts(0, restarts, restart_begin, chunks, tile, max_evals, blocks, ee_offset, variant)
        sh = _stream_handle(stream)
        if not wait:
            if sh is None:
                raise ValueError("wait=False needs a stream")
            o.flags |= BATCH_ASYNC
        if static:
            o.flags |= BATCH_STATIC
        extra = {}
        if stats:
            extra["restart"] = np.zeros(T, dtype=np.uint64)
            extra["evals"] = np.zeros(T, dtype=np.int32)
            extra["counters"] = np.zeros(3, dtype=np.uint64)
            o.restart_out, o.evals_out, o.counters = extra["restart"].ctypes.data, extra["evals"].ctypes.data, extra["counters"].ctypes.data
        _check(lib.optik_gpu_ik_batch(self._h, C.byref(c), C.byref(o), targets.ctypes.data, x0.ctypes.data, T,
                                      q.ctypes.data, f.ctypes.data, st.ctypes.data, sh))
        return (q, f, st, extra) if stats else (q, f, st)

    def ik_attempts(self, config, target, x0, restarts, restart_begin=0, tile=0, max_evals=0, ee_offset=None,
                    best=False, out=None, counters=None, record=None, stream=None, wait=True, variant=0, push=None,
                    blocks=0):
        """Per-restart records for one target: (q_all (R,n), f_all, status_all, evals_all), every restart run to
        completion (no Speed-mode early exit) -- BASELINE config 2's output.  best=True also runs the selection
        pass (lib.rs:397-413) and appends the packed candidate record (RECORD_HEAD + n doubles:
        [found, score, restart, cost, status, 0,0,0, q...]).  numpy = host path (out = preallocated, e.g. pinned,
        buffers), torch CUDA tensors = device path on the current stream (no host sync).  Host path with
        stream=<Stream>, wait=False: enqueue only (target, x0 and the outputs must be pinned and stay alive);
        results are valid after stream.synchronize()."""
        lib = load_library()
        c = config._c()
        n, R = self._n, int(restarts)
        o, keep = self._opts(0, R, restart_begin, 0, tile, max_evals, blocks, ee_offset, variant)
        if _is_torch(target):
            import torch
            dev = target.device
            assert target.is_cuda and x0.is_cuda and target.dtype == torch.float64 and x0.dtype == torch.float64
            assert target.numel() == 8 and x0.numel() == n and target.is_contiguous() and x0.is_contiguous()
            o.memory = 1
            if out is None:
                out = (torch.empty((R, n), dtype=torch.float64, device=dev), torch.empty((R,), dtype=torch.float64, device=dev),
                       torch.empty((R,), dtype=torch.int32, device=dev), torch.empty((R,), dtype=torch.int32, device=dev))
            q, f, st, ev = out
            if best:
                if record is None:
                    record = torch.empty((RECORD_HEAD + n,), dtype=torch.float64, device=dev)
                o.best_record_out = record.data_ptr()
            if push is not None:  # (peers device array ptr, rank, world, seq): fused store into the peers' exchange buffers
                o.push_peers, o.push_rank, o.push_world, o.push_seq = int(push[0]), int(push[1]), int(push[2]), int(push[3])
            if counters is not None:
                o.counters = counters.data_ptr()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _check(lib.optik_gpu_ik_attempts(self._h, C.byref(c), C.byref(o), target.data_ptr(), x0.data_ptr(), q.data_ptr(),
                                             f.data_ptr(), st.data_ptr(), ev.data_ptr(), stream))
            return (q, f, st, ev, record) if best else (q, f, st, ev)
        sh = _stream_handle(stream)
        if not wait:
            if sh is None or out is None:
                raise ValueError("wait=False needs a stream and preallocated (pinned) out buffers")
            o.flags |= BATCH_ASYNC
            if not (isinstance(target, np.ndarray) and target.dtype == np.float64 and target.flags.c_contiguous
                    and isinstance(x0, np.ndarray) and x0.dtype == np.float64 and x0.flags.c_contiguous):
                raise ValueError("wait=False: target and x0 must be contiguous float64 arrays that outlive the call")
        else:
            x0 = self._x(x0)
        target = np.ascontiguousarray(target, dtype=np.float64).reshape(8)
        if out is None:
            out = (np.empty((R, n)), np.empty(R), np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32))
        q, f, st, ev = out
        if best:
            if record is None:
                record = np.zeros(RECORD_HEAD + n)
            o.best_record_out = record.ctypes.data
        if push is not None:
            o.push_peers, o.push_rank, o.push_world, o.push_seq = int(push[0]), int(push[1]), int(push[2]), int(push[3])
        if counters is not None:
            o.counters = counters.ctypes.data
        _check(lib.optik_gpu_ik_attempts(self._h, C.byref(c), C.byref(o), target.ctypes.data, x0.ctypes.data,
                                         q.ctypes.data, f.ctypes.data, st.ctypes.data, ev.ctypes.data, sh))
        return (q, f, st, ev, record) if best else (q, f, st, ev)

    def restart_seeds(self, restart_begin, count):
        """The solver's restart seeds (lib.rs:86-91, 360-370) for indices [restart_begin, restart_begin+count), (count, n)."""
        out = np.zeros((int(count), self._n))
        _check(load_library().optik_gpu_restart_seeds(self._h, int(restart_begin), int(count), 0, out.ctypes.data, None))
        return out

    def chacha8_block(self, key8, stream_id):
        """Known-answer hook: 16 output words of the kernels' ChaCha8 block function (block counter 0)."""
        key = np.ascontiguousarray(key8, dtype=np.uint32)
        assert key.shape == (8,)
        out = np.zeros(16, dtype=np.uint32)
        _check(load_library().optik_gpu_chacha8_block(self._h, key.ctypes.data, int(stream_id), out.ctypes.data))
        return out

    def select_records(self, records, out=None):
        """Best-pick (lib.rs:397-413) over gathered candidate records (W, RECORD_HEAD+n) CUDA tensor -> one record."""
        import torch
        assert records.is_cuda and records.dtype == torch.float64 and records.is_contiguous()
        if out is None:
            out = torch.empty((records.shape[1],), dtype=torch.float64, device=records.device)
        stream = torch.cuda.current_stream(records.device).cuda_stream
        _check(load_library().optik_gpu_select_records(self._h, records.data_ptr(), records.shape[0], out.data_ptr(), stream))
        return out

    def eval_batch(self, q, targets=None, linear_weight=None, angular_weight=None, ee_offset=None,
                   want=("ee", "jac", "f", "grad")):
        """Batched fk / joint_jacobian / objective / objective_grad (lib.rs:93-99, objective.rs:40-110).

        q: (B, n); targets: (B, 8) or (8,) shared.  Returns dict with ee (B,8), jac (B,6n) column-major 6 x n,
        f (B,), grad (B,n) -- those named in `want` (f/grad need targets).  numpy = host path, torch CUDA = device path.
        """
        lib = load_library()
        n = self._n
        want = set(want)
        if targets is None:
            want -= {"f", "grad"}
        wl = None if linear_weight is None else (C.c_double * 3)(*[float(x) for x in linear_weight])
        wa = None if angular_weight is None else (C.c_double * 3)(*[float(x) for x in angular_weight])
        eo = None
        if ee_offset is not None:
            e = np.ascontiguousarray(ee_offset, dtype=np.float64)
            eo = _pose8_from_rows(e) if e.shape == (4, 4) else e.ravel().copy()
        eop = None if eo is None else _dptr(eo)
        if _is_torch(q):
            import torch
            B = q.shape[0]
            assert q.is_cuda and q.dtype == torch.float64 and q.is_contiguous() and q.shape == (B, n)
            shared = 0
            tp = None
            if targets is not None:
                assert targets.is_cuda and targets.dtype == torch.float64 and targets.is_contiguous()
                shared = int(targets.numel() == 8 and B != 1)
                tp = targets.data_ptr()
            mk = lambda shape: torch.empty(shape, dtype=torch.float64, device=q.device)
            outs = {k: mk(s) for k, s in (("ee", (B, 8)), ("jac", (B, 6 * n)), ("f", (B,)), ("grad", (B, n))) if k in want}
            ptr = lambda k: outs[k].data_ptr() if k in outs else None
            stream = torch.cuda.current_stream(q.device).cuda_stream
            _check(lib.optik_gpu_eval_batch(self._h, q.data_ptr(), tp, shared, B, wl, wa, eop, 1, ptr("ee"), ptr("jac"),
                                            ptr("f"), ptr("grad"), stream))
            return outs
        q = np.ascontiguousarray(q, dtype=np.float64)
        B = q.shape[0]
        if q.shape != (B, n):
            raise ValueError("len(x) != num_positions")
        shared, tp = 0, None
        if targets is not None:
            targets = np.ascontiguousarray(targets, dtype=np.float64)
            shared = int(targets.size == 8 and B != 1)
            if not shared and targets.shape != (B, 8):
                raise ValueError("targets must be (B, 8) or (8,)")
            tp = targets.ctypes.data
        outs = {k: np.empty(s) for k, s in (("ee", (B, 8)), ("jac", (B, 6 * n)), ("f", (B,)), ("grad", (B, n))) if k in want}
        ptr = lambda k: outs[k].ctypes.data if k in outs else None
        _check(lib.optik_gpu_eval_batch(self._h, q.ctypes.data, tp, shared, B, wl, wa, eop, 0, ptr("ee"), ptr("jac"),
                                        ptr("f"), ptr("grad"), None))
        return outs
