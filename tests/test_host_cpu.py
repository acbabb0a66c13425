"""CPU: host logic of the product (URDF loader, config mirror, C ABI surface) -- no compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import optik_b200 as ob
from oracle import oracle as O
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "optik_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(optik_[a-z0-9_]+)\s*\(", hdr))
    assert {"optik_robot_ik", "optik_robot_fk", "optik_gpu_ik_batch", "optik_robot_from_urdf_str"} <= declared
    lib = C.CDLL(ob.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing


def test_reference_ffi_symbols_present():  # crates/optik-cpp/src/lib.cpp:5-31
    lib = C.CDLL(ob.LIB_PATH)
    for s in ["optik_robot_from_urdf_file", "optik_robot_from_urdf_str", "optik_robot_free", "optik_robot_set_parallelism",
              "optik_robot_num_positions", "optik_robot_joint_limits", "optik_robot_random_configuration",
              "optik_robot_joint_jacobian", "optik_robot_fk", "optik_robot_ik", "optik_robot_diff_ik"]:
        assert hasattr(lib, s), s


def test_solver_config_layout():  # CSolverConfig is 96 bytes on LP64 (crates/optik-cpp/src/lib.rs:10-20)
    assert C.sizeof(ob._CSolverConfig) == 96
    assert ob._CSolverConfig.max_time.offset == 8 and ob._CSolverConfig.linear_weight.offset == 48


def test_solver_config_defaults_and_validation():  # crates/optik-py/src/lib.rs:24-47
    c = ob.SolverConfig()
    assert (c.solution_mode, c.max_time, c.max_restarts, c.tol_f, c.tol_df, c.tol_dx) == ("speed", 0.1, 2**64 - 1, 1e-6, -1.0, -1.0)
    assert c._c().max_restarts == 0 and c._c().solution_mode == 2
    assert ob.SolverConfig(solution_mode="quality")._c().solution_mode == 1
    with pytest.raises(ValueError):
        ob.SolverConfig(solution_mode="fast")
    with pytest.raises(ValueError, match="run forever"):
        ob.SolverConfig(max_time=0.0, max_restarts=0)


@pytest.mark.parametrize("name", ["panda", "ur5", "ur3e", "snake20"])
def test_urdf_loader_matches_oracle_loader(name):
    base, ee = ob.ROBOT_LINKS[name]
    r = ob.Robot.named(name)
    ch = O.Chain.from_urdf(open(ob.data_path(name)).read(), base, ee)
    assert np.array_equal(r.chain(), ch.arr)
    assert r.num_positions() == ch.n
    lb, ub = r.joint_limits()
    assert np.array_equal(lb, ch.lb) and np.array_equal(ub, ch.ub)


def test_panda_fixture_sanity():  # flange at (0.088, 0, 0.926) for q = 0
    ch = O.Chain(ob.Robot.named("panda").chain())
    _, ee = ch.fk(np.zeros(7))
    assert np.allclose(ee[4:7], [0.088, 0.0, 0.926], atol=1e-12)


FOLD_URDF = """<robot name="f">
  <link name="a"/><link name="b"/><link name="c"/><link name="d"/><link name="e"/><link name="x"/>
  <joint name="j0" type="revolute"><parent link="a"/><child link="b"/><origin xyz="0 0 0.1" rpy="0 0 0"/><axis xyz="0 0 1"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>
  <joint name="f1" type="fixed"><parent link="b"/><child link="c"/><origin xyz="0.1 0 0" rpy="0.3 0 0"/></joint>
  <joint name="f2" type="fixed"><parent link="c"/><child link="d"/><origin xyz="0 0.2 0" rpy="0 0.4 0"/></joint>
  <joint name="j1" type="prismatic"><parent link="d"/><child link="e"/><origin xyz="0 0 0.3" rpy="0 0 0.5"/><axis xyz="0 2 0"/><limit lower="0" upper="0" effort="1" velocity="1"/></joint>
  <joint name="side" type="continuous"><parent link="a"/><child link="x"/></joint>
</robot>"""


def test_fold_order_and_defaults():
    """Non-commuting mid-chain fixed joints: reference order (kinematics.rs:70,77) vs URDF-correct order differ;
    both loaders agree in both modes.  Also: axis normalised, zero-width limits -> (-inf, inf) (kinematics.rs:299-303)."""
    urdf = FOLD_URDF.replace('<joint name="side" type="continuous"><parent link="a"/><child link="x"/></joint>', "")
    r = ob.Robot.from_urdf_str(urdf, "a", "e")
    ref_order = O.Chain.from_urdf(urdf, "a", "e").arr
    assert np.allclose(r.chain(), ref_order, atol=1e-15)
    assert np.allclose(r.chain()[1, 8:11], [0, 1, 0])
    assert r.joint_limits()[0][1] == -np.inf and r.joint_limits()[1][1] == np.inf
    correct = O.Chain.from_urdf(urdf, "a", "e", urdf_correct_fold=True).arr
    assert np.abs(correct[:, :11] - ref_order[:, :11]).max() > 1e-3
    ob.load_library().optik_set_urdf_correct_fold(1)
    try:
        r2 = ob.Robot.from_urdf_str(urdf, "a", "e")
        assert np.allclose(r2.chain(), correct, atol=1e-15)
    finally:
        ob.load_library().optik_set_urdf_correct_fold(0)


@pytest.mark.parametrize("urdf,base,ee,msg", [
    (FOLD_URDF, "a", "e", "joint type not supported"),                       # kinematics.rs:296
    ("<robot><link name='a'/><link name='b'/></robot>", "a", "zz", "EE link 'zz' does not exist"),
    ("<robot><link name='a'/><link name='b'/></robot>", "zz", "a", "base link 'zz' does not exist"),
    ("<robot><link name='a'/><link name='b'/></robot>", "a", "b", "no path from base to EE link"),
    ("<robot><link name='a'/><link name='b'/><joint name='j' type='fixed'><parent link='a'/><child link='b'/></joint></robot>",
     "a", "b", "kinematic chain is empty"),
    ("<robot><link name='a'/><joint name='j' type='fixed'><parent link='a'/><child link='q'/></joint></robot>",
     "a", "a", "joint child link 'q' does not exist"),
    ("<robot><link name='a'/><link name='b'/>"
     "<joint name='j' type='fixed'><parent link='a'/><child link='b'/></joint>"
     "<joint name='k' type='fixed'><parent link='b'/><child link='a'/></joint></robot>", "a", "b", "robot model contains loops"),
    ("<robot><link name='a'>", "a", "a", "error parsing URDF file!"),
])
def test_loader_errors(urdf, base, ee, msg):
    with pytest.raises(ob.OptikError, match=re.escape(msg)):
        ob.Robot.from_urdf_str(urdf, base, ee)
    with pytest.raises(ValueError):
        O.Chain.from_urdf(urdf, base, ee) if "parsing" not in msg else (_ for _ in ()).throw(ValueError())


def test_c_abi_panics_abort_like_the_reference():
    """optik_robot_ik with an out-of-limit seed aborts with the reference's message (lib.rs:251-254;
    tests/test_ik.rs:10-22), as a Rust panic across extern "C" does.  Checked before any CUDA use."""
    code = (
        "import ctypes as C, optik_b200 as ob\n"
        "lib = ob.load_library(); r = ob.Robot.named('ur3e'); cfg = ob.SolverConfig()._c()\n"
        "x0 = (C.c_double*6)(0,0,0,0,10.0,0); tgt = (C.c_double*16)(1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1)\n"
        "lib.optik_robot_ik(r._h, C.byref(cfg), tgt, x0)\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert p.returncode != 0
    assert "seed joint position outside of joint limits" in p.stderr


def test_python_surface_argument_checks():
    r = ob.Robot.named("ur3e")
    with pytest.raises(ValueError, match="num_positions"):
        r.ik(ob.SolverConfig(), np.eye(4), [0.0] * 5)
    with pytest.raises(ValueError, match="joint limits"):  # tests/test_ik.rs:10-22
        r.ik(ob.SolverConfig(), np.eye(4), [0, 0, 0, 0, 10.0, 0])
    with pytest.raises(ValueError, match="invalid target transform"):
        r.ik(ob.SolverConfig(), np.ones((4, 4)), [0.0] * 6)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = ob.Robot.named("ur3e")
    with pytest.raises(ob.OptikError):
        r.fk([0.0] * 6)
    with pytest.raises(ob.OptikError):
        r.ik(ob.SolverConfig(max_time=0.0, max_restarts=4), np.eye(4), [0.0] * 6)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "optik_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboptik_oracle" not in text and '#include "../../oracle' not in text, f
    out = subprocess.run(["ldd", ob.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_header_is_valid_c_and_links(tmp_path):
    exe = tmp_path / "c_abi_probe"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_probe.c"), "-o", str(exe),
                           ob.LIB_PATH, "-Wl,-rpath," + os.path.dirname(ob.LIB_PATH)])
    base, ee = ob.ROBOT_LINKS["panda"]
    p = subprocess.run([str(exe), ob.data_path("panda"), base, ee], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "n=7 joints=8 inside=1 lb0=-2.8973 ub0=2.8973" in p.stdout
    assert "EE link 'nope' does not exist" in p.stdout
    # the ctypes mirror of optik_gpu_batch_opts has the C compiler's layout
    assert f"sizeof(optik_gpu_batch_opts)={C.sizeof(ob._BatchOpts)}" in p.stdout


def test_cpp_consumer_with_its_own_declarations_links_and_runs(tmp_path):
    """A C++ wrapper that declares the 11 symbols ITSELF against its own opaque robot type and SolverConfig POD, the way
    crates/optik-cpp/src/lib.cpp:5-31 does, links against liboptik_b200.so and drives the non-compute calls."""
    exe = tmp_path / "cpp_wrapper_probe"
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp_wrapper_probe.cpp"),
                           "-o", str(exe), ob.LIB_PATH, "-Wl,-rpath," + os.path.dirname(ob.LIB_PATH)])
    base, ee = ob.ROBOT_LINKS["ur3e"]
    p = subprocess.run([str(exe), ob.data_path("ur3e"), base, ee, "cpu"], capture_output=True, text=True)
    assert p.returncode == 0 and "n=6 inside=1" in p.stdout, p.stdout + p.stderr


def test_new_entry_points_check_their_arguments_before_touching_the_gpu():
    """Shape / mode errors of the batched additions surface as Python exceptions on a machine without a GPU."""
    r = ob.Robot.named("ur3e")
    with pytest.raises(ValueError):
        r.diff_ik([0.0] * 5, [0.0] * 6, [1.0] * 6)            # len(x0) != num_positions (crates/optik-py/src/lib.rs:143)
    with pytest.raises(ValueError):
        r.diff_ik([0.0] * 6, [0.0] * 6, [1.0] * 5)            # len(v_max) != num_positions (:144-148)
    with pytest.raises(ValueError):
        r.diff_ik([0.0] * 6, [0.0] * 5, [1.0] * 6)
    with pytest.raises(ValueError):
        r.diff_ik_batch(np.zeros((3, 5)), np.zeros(6), np.ones(6))
    with pytest.raises(ValueError):
        r.diff_ik_batch(np.zeros((3, 6)), np.zeros((2, 6)), np.ones(6))
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=4)
    with pytest.raises(ValueError):
        r.ik_attempts(cfg, np.zeros(8), np.zeros(6), 4, wait=False)   # asynchronous call without a stream
    with pytest.raises(ValueError):
        r.ik_batch(cfg, np.zeros((2, 8)), np.zeros((2, 6)), restarts=4, wait=False)
    with pytest.raises(ValueError):
        r.ik_batch(cfg, np.zeros((2, 7)), np.zeros((2, 6)), restarts=4)


def test_batch_opts_flags_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "optik_b200.h")).read()
    for name, val in (("OPTIK_BATCH_ASYNC", ob.BATCH_ASYNC), ("OPTIK_BATCH_STATIC", ob.BATCH_STATIC)):
        assert f"#define {name} {val}u" in hdr
    assert f"#define OPTIK_RECORD_HEAD {ob.RECORD_HEAD}" in hdr
    assert f"#define OPTIK_STATUS_CODE_MASK {ob.STATUS_CODE_MASK:#x}" in hdr
    assert f"#define OPTIK_STATUS_FLAG_SEED_CLAMPED {ob.STATUS_FLAG_SEED_CLAMPED:#x}" in hdr
    # a clamped-seed flag never changes the success classification (lib.rs:376-379 on the status code)
    cfg = ob.SolverConfig(max_time=0.0, max_restarts=1)
    assert list(cfg.is_success(np.array([1, 1 | 0x100, 5 | 0x100, 2]))) == [True, True, False, False]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: oracle port on the host cores) prints ONE JSON line on stdout
    with the contract's keys, never touches a GPU and never maps the product library."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["gpu_launches"] == 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
