"""CPU: pins the oracle to the reference's own golden vectors (tests/test_fk.rs, test_math.rs, test_gradient.rs)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from conftest import REF_DATA, ROOT

TOL = 1e-6  # the reference's epsilon (tests/test_fk.rs:24, tests/test_math.rs:20)


@pytest.fixture(scope="module")
def ur3e():
    return O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", "ur3e.urdf")).read(), "ur_base_link", "ur_ee_link")


def _quat_close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol


def test_fk_golden(ur3e, golden):  # tests/test_fk.rs:13-26
    for q, ref in zip(golden["fk_inputs"], golden["fk_outputs"]):
        _, ee = ur3e.fk(q)
        assert np.abs(ee[4:7] - ref[4:7]).max() < 1e-12
        assert _quat_close(ee[:4], ref[:4], 1e-12)


def test_so3_log_golden(golden):  # tests/test_math.rs:14-22
    for p, ref in zip(golden["math_inputs"], golden["so3_log"]):
        assert np.abs(O.so3_log(p[:4]) - ref).max() < TOL


def test_so3_log_singularity():  # tests/test_math.rs:24-30
    assert np.abs(O.so3_log([0, 0, 0, 1])).max() < TOL


def test_so3_right_jacobian_golden(golden):  # tests/test_math.rs:32-41
    for p, ref in zip(golden["math_inputs"], golden["so3_right_jacobian"]):
        assert np.abs(O.so3_right_jacobian(O.so3_log(p[:4])) - ref).max() < TOL


def test_se3_log_golden(golden):  # tests/test_math.rs:43-51
    for p, ref in zip(golden["math_inputs"], golden["se3_log"]):
        assert np.abs(O.se3_log(p) - ref).max() < TOL


def test_se3_right_jacobian_golden(golden):  # tests/test_math.rs:53-61
    for p, ref in zip(golden["math_inputs"], golden["se3_right_jacobian"]):
        assert np.abs(O.se3_right_jacobian(p) - ref).max() < TOL


def test_lie_singularities_are_finite():  # the reference yields NaN at exactly zero rotation
    p = np.array([0, 0, 0, 1, 0.1, -0.2, 0.3, 0.0])
    assert np.all(np.isfinite(O.se3_log(p)))
    assert np.allclose(O.se3_log(p)[:3], p[4:7])
    assert np.all(np.isfinite(O.se3_right_jacobian(p)))
    assert np.allclose(O.so3_right_jacobian([0, 0, 0]), np.eye(3).ravel())


def test_gradient_vs_finite_difference(ur3e):  # tests/test_gradient.rs:34-68, same weights and eps
    wl, wa = [0.0, 5.0, 0.25], [0.005, 1.0, 0.99]
    eps = np.finfo(float).eps ** (1.0 / 3.0)
    rng = np.random.default_rng(42)
    for _ in range(100):
        q = rng.random(6)
        quat = rng.normal(size=4)
        tgt = O.pose8(quat / np.linalg.norm(quat), rng.random(3))
        g = ur3e.objective_grad(q, tgt, wl, wa)
        gn = np.zeros(6)
        for i in range(6):
            a, b = q.copy(), q.copy()
            a[i] -= eps
            b[i] += eps
            gn[i] = (ur3e.objective(b, tgt, wl, wa) - ur3e.objective(a, tgt, wl, wa)) / (2 * eps)
        assert np.abs(g - gn).max() < TOL


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference checkout not mounted")
def test_fixture_matches_reference_files(golden, ur3e):
    """tests/golden/ref_vectors.json is a faithful repack of the reference's JSON files, and our kinematics-only
    ur3e.urdf yields the same chain as the reference's ur3e.urdf."""
    fi = json.load(open(os.path.join(REF_DATA, "test_fk_inputs.json")))
    fo = json.load(open(os.path.join(REF_DATA, "test_fk_outputs.json")))
    assert fi == golden["fk_inputs"]
    assert [o["rotation"] + o["translation"] + [0.0] for o in fo] == golden["fk_outputs"]
    ref_chain = O.Chain.from_urdf(open(os.path.join(REF_DATA, "ur3e.urdf")).read(), "ur_base_link", "ur_ee_link")
    assert np.array_equal(ref_chain.arr, ur3e.arr)


def test_chacha_quarter_round_rfc7539():  # RFC 7539 section 2.1.1
    import ctypes as C
    v = (C.c_uint32 * 4)(0x11111111, 0x01020304, 0x9b8d6f43, 0x01234567)
    O.lib().oracle_chacha_quarter_round(v)
    assert list(v) == [0xea2a92f4, 0xcb1cf8ce, 0x4581472e, 0x5881c4bb]


def test_chacha20_block_structure_rfc7539():
    """Our block function with 20 rounds would be RFC 7539's; with 8 rounds we can still pin the state layout:
    different streams/counters give different blocks and the feed-forward adds the input state."""
    import ctypes as C
    key = (C.c_uint32 * 8)()
    O.lib().oracle_seed_key(C.c_uint64(42), key)
    a, b, c = (C.c_uint32 * 16)(), (C.c_uint32 * 16)(), (C.c_uint32 * 16)()
    O.lib().oracle_chacha8_block(key, 0, 1, a)
    O.lib().oracle_chacha8_block(key, 1, 1, b)
    O.lib().oracle_chacha8_block(key, 0, 2, c)
    assert list(a) != list(b) and list(a) != list(c)
    assert O.lib().oracle_rng_u64(42, 1, 0) == (a[0] | (a[1] << 32))
    assert O.lib().oracle_rng_u64(42, 1, 9) == (b[2] | (b[3] << 32))


CHACHA8_TC1 = ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
               "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")


def test_chacha8_known_answer_zero_key():
    """Published ChaCha8 test vector (all-zero 256-bit key, zero nonce/stream, block 0): the keystream is the block's 16
    output words in little-endian order -- pins the round function, the round count, the constants, the feed-forward and
    the word order rand_chacha serialises (next_u32 = word i, next_u64 = word 2i | word 2i+1 << 32)."""
    import ctypes as C
    key = (C.c_uint32 * 8)()
    out = (C.c_uint32 * 16)()
    O.lib().oracle_chacha8_block(key, 0, 0, out)
    assert np.array(list(out), dtype="<u4").tobytes().hex() == CHACHA8_TC1


def test_restart_seeds_uniform_in_limits(ur3e):
    qs = np.array([ur3e.restart_seed(i) for i in range(1, 2001)])
    assert np.all(qs >= ur3e.lb) and np.all(qs <= ur3e.ub)
    # roughly uniform: mean near 0, std near (ub-lb)/sqrt(12)
    assert np.abs(qs.mean(0)).max() < 0.2
    assert np.abs(qs.std(0) - (ur3e.ub - ur3e.lb) / np.sqrt(12)).max() < 0.1
    assert np.array_equal(ur3e.restart_seed(7), ur3e.restart_seed(7))


def test_batch_evaluator_equals_single_calls():
    """oracle_eval_batch_threaded (bench.py's CPU number beside the evaluator kernel) returns exactly what the
    golden-pinned single-configuration functions return, for any thread count."""
    from oracle import urdf_chain
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    ch = O.Chain.from_urdf(open(os.path.join(here, "..", "optik_b200", "data", "ur3e.urdf")).read(), "ur_base_link", "ur_ee_link")
    rng = np.random.default_rng(9)
    B = 257
    q = rng.uniform(ch.lb, ch.ub, size=(B, ch.n))
    tg = np.stack([ch.fk(rng.uniform(ch.lb, ch.ub))[1] for _ in range(B)])
    a = O.eval_batch_threaded(ch, q, tg, 1)
    b = O.eval_batch_threaded(ch, q, tg, 3)
    for k in a:
        assert np.array_equal(a[k], b[k])
    for i in range(0, B, 16):
        assert a["f"][i] == ch.objective(q[i], tg[i])
        assert np.array_equal(a["grad"][i], ch.objective_grad(q[i], tg[i]))
        assert np.array_equal(a["jac"][i].T, ch.joint_jacobian(q[i]))
        assert np.array_equal(a["ee"][i], ch.fk(q[i])[1])
