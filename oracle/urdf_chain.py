"""Oracle-side URDF -> kinematic chain, restating the reference's semantics.

TEST INFRASTRUCTURE ONLY.  Follows crates/optik/src/kinematics.rs:
  parse_urdf 269-319 (joint types, limit rule 299-303, origin 263-267),
  from_urdf 18-105 (acyclic check 21, base->EE path 24-45, fixed-joint folding
  64-86 with the reference's `joint.origin * collapsed_tfm` order, tip joint
  90-97, empty-chain assert 102).
urdf-rs defaults: origin xyz/rpy = 0, axis = (1,0,0), limit lower=upper=0.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

REVOLUTE, PRISMATIC, FIXED = 0, 1, 2


def _floats(s, n, default):
    if s is None:
        return list(default)
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return v


def quat_from_rpy(r, p, y):
    """UnitQuaternion::from_euler_angles(roll,pitch,yaw) = Rz(yaw) Ry(pitch) Rx(roll); xyzw."""
    sr, cr = math.sin(r / 2), math.cos(r / 2)
    sp, cp = math.sin(p / 2), math.cos(p / 2)
    sy, cy = math.sin(y / 2), math.cos(y / 2)
    return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy]


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return [aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz]


def _qrot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return list(R @ np.asarray(v))


def _tf_mul(a, b):
    """(t,q) composition a*b."""
    ta, qa = a
    tb, qb = b
    r = _qrot(qa, tb)
    return ([ta[0] + r[0], ta[1] + r[1], ta[2] + r[2]], _qmul(qa, qb))


IDENT = ([0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0])


def parse_urdf(text):
    root = ET.fromstring(text)
    links = [l.attrib["name"] for l in root.findall("link")]
    joints = []
    for j in root.findall("joint"):
        typ = j.attrib["type"]
        parent = j.find("parent").attrib["link"]
        child = j.find("child").attrib["link"]
        if parent not in links:
            raise ValueError(f"joint parent link '{parent}' does not exist")
        if child not in links:
            raise ValueError(f"joint child link '{child}' does not exist")
        o = j.find("origin")
        xyz = _floats(None if o is None else o.attrib.get("xyz"), 3, (0, 0, 0))
        rpy = _floats(None if o is None else o.attrib.get("rpy"), 3, (0, 0, 0))
        a = j.find("axis")
        axis = _floats(None if a is None else a.attrib.get("xyz"), 3, (1, 0, 0))
        lim = j.find("limit")
        lo = float(lim.attrib.get("lower", 0.0)) if lim is not None else 0.0
        hi = float(lim.attrib.get("upper", 0.0)) if lim is not None else 0.0
        if typ == "revolute":
            t = REVOLUTE
        elif typ == "prismatic":
            t = PRISMATIC
        elif typ == "fixed":
            t = FIXED
        else:
            raise ValueError(f"joint type not supported: {typ}")
        if t != FIXED:
            nrm = math.sqrt(sum(x * x for x in axis))
            axis = [x / nrm for x in axis]
        limits = (lo, hi) if hi - lo > 0.0 else (-math.inf, math.inf)
        joints.append(dict(name=j.attrib.get("name", ""), type=t, parent=parent, child=child,
                           origin=(xyz, quat_from_rpy(*rpy)), axis=axis, limits=limits))
    return links, joints


def chain_from_urdf(text, base_link, ee_link, urdf_correct_fold=False):
    """Returns an (njoints, 16) float64 array in the flat chain format of optik_oracle.c."""
    links, joints = parse_urdf(text)
    if base_link not in links:
        raise ValueError(f"base link '{base_link}' does not exist")
    if ee_link not in links:
        raise ValueError(f"EE link '{ee_link}' does not exist")
    out = {l: [] for l in links}
    for j in joints:
        out[j["parent"]].append(j)
    # acyclicity (kinematics.rs:21)
    state = {}

    def visit(u):
        state[u] = 1
        for j in out[u]:
            v = j["child"]
            if state.get(v) == 1:
                raise ValueError("robot model contains loops")
            if v not in state:
                visit(v)
        state[u] = 2

    for l in links:
        if l not in state:
            visit(l)
    # shortest path by hop count (A* with unit edge cost, zero heuristic)
    prev = {base_link: None}
    frontier = [base_link]
    while frontier and ee_link not in prev:
        nxt = []
        for u in frontier:
            for j in out[u]:
                if j["child"] not in prev:
                    prev[j["child"]] = j
                    nxt.append(j["child"])
        frontier = nxt
    if ee_link not in prev:
        raise ValueError("no path from base to EE link")
    path = []
    l = ee_link
    while prev[l] is not None:
        path.append(prev[l])
        l = prev[l]["parent"]
    path.reverse()
    # fold fixed joints (kinematics.rs:64-86)
    chain, collapsed = [], IDENT
    for j in path:
        if j["type"] == FIXED:
            collapsed = _tf_mul(collapsed, j["origin"]) if urdf_correct_fold else _tf_mul(j["origin"], collapsed)
        else:
            origin = _tf_mul(collapsed, j["origin"]) if urdf_correct_fold else _tf_mul(j["origin"], collapsed)
            chain.append(dict(j, origin=origin))
            collapsed = IDENT
    if collapsed != IDENT:
        chain.append(dict(name="", type=FIXED, origin=collapsed, axis=[0, 0, 0], limits=(0.0, 0.0)))
    if sum(1 for j in chain if j["type"] != FIXED) == 0:
        raise ValueError("kinematic chain is empty")
    arr = np.zeros((len(chain), 16))
    for i, j in enumerate(chain):
        t, q = j["origin"]
        arr[i, 0:3] = t
        arr[i, 3] = j["type"]
        arr[i, 4:8] = q
        arr[i, 8:11] = j["axis"]
        arr[i, 12], arr[i, 13] = j["limits"]
    return arr
