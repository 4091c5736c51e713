"""CPU ORACLE (test infrastructure, NOT product code): URDF -> flat kinematic tree.

Restates what the reference obtains from
``pinocchio::urdf::buildModelFromXML(xml, JointModelFreeFlyer(), model)``
(reference ik_ros/src/cassie.cpp:34-35); SURVEY.md 8c.1 lists the rules:

* urdfdom keeps joints in a ``std::map`` keyed by name, so a link's children
  are visited in byte-wise alphabetical order of JOINT name; Pinocchio walks
  the tree depth-first.
* movable joints get indices in visit order; ``fixed`` joints are merged into
  the supporting joint and become a FIXED_JOINT frame (joint name) plus a BODY
  frame (child link name) whose placement is accumulated from that joint.
* ``rpy`` goes through urdfdom's quaternion (setFromRPY + normalise) and Eigen's
  quaternion -> matrix; literals are parsed verbatim.
* axis (1,0,0)/(0,1,0)/(0,0,1) -> RX/RY/RZ, anything else -> unaligned (normalised).
* the free-flyer root joint ("root_joint", index 1) has limits +-DBL_MAX.

The product has its own, independently written C++ flattener
(ik_b200/csrc/urdf_model.cpp); tests require the two to agree exactly on the
topology and to 1e-15 on placements.  PARITY UNPINNED (see ik_oracle.h).
"""
import math
import sys
import xml.etree.ElementTree as ET

import numpy as np

J_UNIVERSE, J_FREEFLYER, J_RX, J_RY, J_RZ, J_REV_UNALIGNED, J_PX, J_PY, J_PZ, J_PRIS_UNALIGNED = range(10)
FRAME_OP, FRAME_JOINT, FRAME_FIXED_JOINT, FRAME_BODY = 0, 1, 2, 3


def _floats(s, n, default):
    if s is None:
        return list(default)
    vals = [float(t) for t in s.split()]
    if len(vals) != n:
        raise ValueError("expected %d numbers, got %r" % (n, s))
    return vals


def rpy_to_matrix(r, p, y):
    """urdfdom Rotation::setFromRPY followed by Eigen Quaternion::toRotationMatrix."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    qx = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    qy = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    qz = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    qw = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qx, qy, qz, qw = qx / s, qy / s, qz / s, qw / s
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    return [1 - (tyy + tzz), txy - twz, txz + twy,
            txy + twz, 1 - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, 1 - (txx + tyy)]


def se3_mul(a, b):
    ra, pa, rb, pb = a[:9], a[9:], b[:9], b[9:]
    r = [sum(ra[3 * i + k] * rb[3 * k + j] for k in range(3)) for i in range(3) for j in range(3)]
    p = [pa[i] + sum(ra[3 * i + k] * pb[k] for k in range(3)) for i in range(3)]
    return r + p


IDENTITY = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]


def flatten_urdf(xml_text, free_flyer=True):
    root = ET.fromstring(xml_text)
    links = [l.get("name") for l in root.findall("link")]
    # <inertial>: mass and centre of mass of every link in its own frame (the rotational inertia is irrelevant for IK)
    inertials = {}
    for l in root.findall("link"):
        ine = l.find("inertial")
        if ine is None or ine.find("mass") is None:
            continue
        origin = ine.find("origin")
        inertials[l.get("name")] = (float(ine.find("mass").get("value")),
                                    _floats(origin.get("xyz") if origin is not None else None, 3, (0, 0, 0)))
    joints = {}
    for j in root.findall("joint"):
        name = j.get("name")
        origin = j.find("origin")
        xyz = _floats(origin.get("xyz") if origin is not None else None, 3, (0, 0, 0))
        rpy = _floats(origin.get("rpy") if origin is not None else None, 3, (0, 0, 0))
        axis_el = j.find("axis")
        axis = _floats(axis_el.get("xyz") if axis_el is not None else None, 3, (1, 0, 0))
        lim = j.find("limit")
        lower = float(lim.get("lower", "0")) if lim is not None else 0.0
        upper = float(lim.get("upper", "0")) if lim is not None else 0.0
        joints[name] = dict(name=name, type=j.get("type"), parent=j.find("parent").get("link"),
                            child=j.find("child").get("link"), placement=rpy_to_matrix(*rpy) + xyz,
                            axis=axis, lower=lower, upper=upper)
    children = {l: [] for l in links}
    has_parent = set()
    # byte-wise order of joint names, like std::map<std::string, ...>
    for name in sorted(joints, key=lambda s: s.encode("utf-8")):
        jt = joints[name]
        children[jt["parent"]].append(jt)
        has_parent.add(jt["child"])
    roots = [l for l in links if l not in has_parent]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)

    m = dict(names=["universe"], parent=[0], jtype=[J_UNIVERSE], idx_q=[0], idx_v=[0], placement=[list(IDENTITY)],
             axis=[[0.0, 0.0, 0.0]], lower=[], upper=[],
             frame_names=["universe"], frame_parent=[0], frame_placement=[list(IDENTITY)], frame_type=[FRAME_OP])
    nq = nv = 0
    body_frame = {}

    def add_frame(name, parent_joint, placement, ftype):
        m["frame_names"].append(name)
        m["frame_parent"].append(parent_joint)
        m["frame_placement"].append(list(placement))
        m["frame_type"].append(ftype)
        return len(m["frame_names"]) - 1

    def add_joint(name, jtype, parent_joint, placement, axis, lower, upper):
        nonlocal nq, nv
        m["names"].append(name)
        m["parent"].append(parent_joint)
        m["jtype"].append(jtype)
        m["idx_q"].append(nq)
        m["idx_v"].append(nv)
        m["placement"].append(list(placement))
        m["axis"].append(list(axis))
        m["lower"].extend(lower)
        m["upper"].extend(upper)
        nq += len(lower)
        nv += 6 if jtype == J_FREEFLYER else 1
        return len(m["names"]) - 1

    root_link = roots[0]
    if free_flyer:
        big = sys.float_info.max
        jid = add_joint("root_joint", J_FREEFLYER, 0, IDENTITY, [0, 0, 0], [-big] * 7, [big] * 7)
        add_frame("root_joint", jid, IDENTITY, FRAME_JOINT)
        body_frame[root_link] = add_frame(root_link, jid, IDENTITY, FRAME_BODY)
    else:
        body_frame[root_link] = add_frame(root_link, 0, IDENTITY, FRAME_BODY)

    def classify(axis, aligned, unaligned):
        def approx(a, b):  # Eigen isApprox, precision 1e-12
            d2 = sum((x - y) ** 2 for x, y in zip(a, b))
            return d2 <= 1e-24 * min(sum(x * x for x in a), sum(y * y for y in b))
        for k, unit in enumerate(([1, 0, 0], [0, 1, 0], [0, 0, 1])):
            if approx(axis, unit):
                return aligned[k], [float(u) for u in unit]
        n = math.sqrt(sum(x * x for x in axis))
        return unaligned, [x / n for x in axis]

    def visit(jt):
        pframe = body_frame[jt["parent"]]
        support = m["frame_parent"][pframe]
        placement = se3_mul(m["frame_placement"][pframe], jt["placement"])
        if jt["type"] == "fixed":
            add_frame(jt["name"], support, placement, FRAME_FIXED_JOINT)
            body_frame[jt["child"]] = add_frame(jt["child"], support, placement, FRAME_BODY)
        elif jt["type"] in ("revolute", "prismatic"):
            if jt["type"] == "revolute":
                jtype, axis = classify(jt["axis"], (J_RX, J_RY, J_RZ), J_REV_UNALIGNED)
            else:
                jtype, axis = classify(jt["axis"], (J_PX, J_PY, J_PZ), J_PRIS_UNALIGNED)
            jid = add_joint(jt["name"], jtype, support, placement, axis, [jt["lower"]], [jt["upper"]])
            add_frame(jt["name"], jid, IDENTITY, FRAME_JOINT)
            body_frame[jt["child"]] = add_frame(jt["child"], jid, IDENTITY, FRAME_BODY)
        else:
            raise ValueError("unsupported joint type %r (joint %s)" % (jt["type"], jt["name"]))
        for ch in children[jt["child"]]:
            visit(ch)

    for jt in children[root_link]:
        visit(jt)

    # Pinocchio appends the inertia of every body to its supporting joint (bodies behind fixed joints included):
    # per joint the total mass and the centre of mass in the joint frame
    nj = len(m["names"])
    mass = np.zeros(nj)
    moment = np.zeros((nj, 3))
    for link, (ml, c) in inertials.items():
        if link not in body_frame:
            continue
        f = body_frame[link]
        pl = m["frame_placement"][f]
        cj = [pl[9 + i] + sum(pl[3 * i + k] * c[k] for k in range(3)) for i in range(3)]
        mass[m["frame_parent"][f]] += ml
        moment[m["frame_parent"][f]] += ml * np.array(cj)
    com = np.where(mass[:, None] > 0, moment / np.maximum(mass[:, None], 1e-300), 0.0)

    out = dict(
        mass=mass, com=com,
        names=m["names"], frame_names=m["frame_names"], njoints=len(m["names"]), nq=nq, nv=nv,
        parent=np.array(m["parent"], dtype=np.int32), jtype=np.array(m["jtype"], dtype=np.int32),
        idx_q=np.array(m["idx_q"], dtype=np.int32), idx_v=np.array(m["idx_v"], dtype=np.int32),
        placement=np.array(m["placement"], dtype=np.float64), axis=np.array(m["axis"], dtype=np.float64),
        lower=np.array(m["lower"], dtype=np.float64), upper=np.array(m["upper"], dtype=np.float64),
        nframes=len(m["frame_names"]), frame_parent=np.array(m["frame_parent"], dtype=np.int32),
        frame_placement=np.array(m["frame_placement"], dtype=np.float64),
        frame_type=np.array(m["frame_type"], dtype=np.int32),
    )
    return out


def frame_id(model, name):
    """model.getFrameId(name): first frame with that name, nframes when absent (reference common.hpp:50)."""
    try:
        return model["frame_names"].index(name)
    except ValueError:
        return model["nframes"]
