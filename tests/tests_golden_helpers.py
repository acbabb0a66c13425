"""Shared helper: the world-frame Jacobian the reference's diff_ik uses (lib.rs:183-190), from the oracle."""
import numpy as np


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
