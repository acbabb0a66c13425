"""Packs the reference's own golden vectors into one fixture that travels to the GPU box.

Reads (read-only) crates/optik/tests/data/*.json of kylc/optik @ 355e463 -- the inputs/outputs of
tests/test_fk.rs:13-26 and tests/test_math.rs:14-61 -- and writes tests/golden/ref_vectors.json:
  fk_inputs   50 x 6 joint vectors (UR3e)           fk_outputs   50 x pose8 {qx,qy,qz,qw,tx,ty,tz,0}
  math_inputs 10 x pose8                             so3_log 10x3, se3_log 10x6,
  so3_right_jacobian 10x9 (column-major), se3_right_jacobian 10x36 (column-major)
Also records oracle-generated vectors (joint Jacobian, objective, gradient for the FK inputs against a fixed
target) produced by oracle/optik_oracle.c AFTER it reproduced the files above, so the GPU tests can be run
without /root/reference.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/crates/optik/tests/data"


def main():
    from oracle import oracle as O
    ld = lambda f: json.load(open(os.path.join(REF, f)))
    fi, fo = ld("test_fk_inputs.json"), ld("test_fk_outputs.json")
    mi = ld("test_math_inputs.json")
    out = {
        "source": "kylc/optik@355e463 crates/optik/tests/data (test_fk.rs:13-26, test_math.rs:14-61)",
        "fk_inputs": fi,
        "fk_outputs": [o["rotation"] + o["translation"] + [0.0] for o in fo],
        "math_inputs": [m["rotation"] + m["translation"] + [0.0] for m in mi],
    }
    for k in ("so3_log", "se3_log", "so3_right_jacobian", "se3_right_jacobian"):
        out[k] = [list(np.asarray(v, dtype=float).ravel()) for v in ld(f"test_math_outputs_{k}.json")]
    # oracle-generated (pinned oracle) vectors for quantities the reference has no golden for
    ch = O.Chain.from_urdf(open(os.path.join(ROOT, "optik_b200", "data", "ur3e.urdf")).read(), "ur_base_link", "ur_ee_link")
    tgt = np.array(out["math_inputs"][3])
    wl, wa = [0.0, 5.0, 0.25], [0.005, 1.0, 0.99]  # tests/test_gradient.rs:37-38
    out["oracle_target"] = list(tgt)
    out["oracle_weights"] = [wl, wa]
    out["oracle_jacobian"] = [list(ch.joint_jacobian(q).T.ravel()) for q in fi]  # column-major 6 x n
    out["oracle_objective"] = [ch.objective(q, tgt, wl, wa) for q in fi]
    out["oracle_gradient"] = [list(ch.objective_grad(q, tgt, wl, wa)) for q in fi]
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote ref_vectors.json", os.path.getsize(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
