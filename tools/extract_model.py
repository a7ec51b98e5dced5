#!/usr/bin/env python3
"""Build-time model extraction: robot.urdf -> constant joint table.

Reads the reference robot description (src/frankaridgeback/model/robot.urdf,
joints :35-84,433-628,779-794, inertials :14-33,109-155,389-415,630-741, end
effector frame :795-801) and reduces it to the 12-joint tree the rollout
kernel integrates, merging fixed joints the way pinocchio::urdf::buildModel
does (pinocchio 2.7.1, the dependency pinned in vcpkg_overlays/pinocchio):

  * an actuated joint becomes a model joint whose placement is the product of
    all fixed-joint origins since the previous actuated joint;
  * the inertia of every link behind a fixed joint is appended to the body of
    the nearest actuated ancestor joint (Inertia::operator+, parallel axis);
  * fixed joints survive only as named frames (parent joint + placement).

Outputs (both committed, the URDF itself is never copied):
  assistedmanipulation_b200/csrc/robot_model.h   constants for CUDA + oracle
  tests/golden/robot_model.json                  same numbers + FK known-answers

The FK known-answer values are computed by an independent walk over the
*unmerged* URDF tree, so they also check the merging.

Usage: python tools/extract_model.py [/path/to/robot.urdf]
"""
import json
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEFAULT_URDF = "/root/reference/src/frankaridgeback/model/robot.urdf"

# Actuated joint order == FrankaRidgeback::State layout (state.hpp:113-194):
EXPECTED_ORDER = [
    "x_base_joint", "y_base_joint", "pivot_joint",
    "panda_joint1", "panda_joint2", "panda_joint3", "panda_joint4",
    "panda_joint5", "panda_joint6", "panda_joint7",
    "panda_finger_joint1", "panda_finger_joint2",
]
# Link enum -> body (dynamics.hpp:65-80); used for the true-FK self collision mode.
LINKS = ["omni_base_root_link", "x_slider", "y_slider", "pivot", "panda_link1",
         "panda_link2", "panda_link3", "panda_link4", "panda_link5",
         "panda_link6", "panda_link7", "panda_leftfinger", "panda_rightfinger"]
FRAMES_WANTED = ["panda_grasp_joint", "arm_mount_joint"]


def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def se3(R=None, p=None):
    return (np.eye(3) if R is None else np.array(R, float), np.zeros(3) if p is None else np.array(p, float))


def se3_mul(a, b):
    return (a[0] @ b[0], a[0] @ b[1] + a[1])


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


class Inertia:
    """mass, com, inertia about com (all in the owning joint frame)."""

    def __init__(self, m=0.0, c=None, I=None):
        self.m, self.c, self.I = float(m), np.zeros(3) if c is None else np.array(c, float), np.zeros((3, 3)) if I is None else np.array(I, float)

    def transformed(self, M):
        R, p = M
        return Inertia(self.m, R @ self.c + p, R @ self.I @ R.T)

    def __add__(self, o):
        m = self.m + o.m
        if m == 0.0:
            return Inertia()
        ab = self.c - o.c
        c = (self.m * self.c + o.m * o.c) / m
        S = skew(ab)
        I = self.I + o.I - (self.m * o.m / m) * (S @ S)
        return Inertia(m, c, I)


def parse(path):
    root = ET.parse(path).getroot()
    links, joints = {}, {}
    for l in root.findall("link"):
        i = l.find("inertial")
        if i is None:
            links[l.get("name")] = Inertia()
            continue
        o = i.find("origin")
        xyz = [float(x) for x in o.get("xyz", "0 0 0").split()] if o is not None else [0, 0, 0]
        rpy = [float(x) for x in o.get("rpy", "0 0 0").split()] if o is not None else [0, 0, 0]
        t = i.find("inertia").attrib
        I = np.array([[float(t["ixx"]), float(t["ixy"]), float(t["ixz"])],
                      [float(t["ixy"]), float(t["iyy"]), float(t["iyz"])],
                      [float(t["ixz"]), float(t["iyz"]), float(t["izz"])]])
        R = rpy_to_R(rpy)
        links[l.get("name")] = Inertia(float(i.find("mass").get("value")), xyz, R @ I @ R.T)
    for j in root.findall("joint"):
        o = j.find("origin")
        xyz = [float(x) for x in o.get("xyz", "0 0 0").split()] if o is not None else [0, 0, 0]
        rpy = [float(x) for x in o.get("rpy", "0 0 0").split()] if o is not None else [0, 0, 0]
        a = j.find("axis")
        axis = [float(x) for x in a.get("xyz").split()] if a is not None else [1, 0, 0]
        joints[j.get("name")] = dict(type=j.get("type"), parent=j.find("parent").get("link"),
                                     child=j.find("child").get("link"), M=se3(rpy_to_R(rpy), xyz), axis=np.array(axis, float))
    return links, joints


def build(links, joints):
    children = {}
    for name in sorted(joints):  # urdfdom fills child_joints from a std::map -> alphabetical
        children.setdefault(joints[name]["parent"], []).append(name)
    model = []          # actuated joints
    frames = {}         # fixed-joint (and actuated-joint) frames: name -> (parent joint idx, placement)
    link_home = {}      # link -> (joint idx, placement of link frame in joint frame)

    def visit(link, jidx, M_in_joint):
        link_home[link] = (jidx, M_in_joint)
        if jidx >= 0:
            model[jidx]["Y"] = model[jidx]["Y"] + links[link].transformed(M_in_joint)
        for jn in children.get(link, []):
            j = joints[jn]
            M = se3_mul(M_in_joint, j["M"])
            if j["type"] == "fixed":
                frames[jn] = (jidx, M)
                visit(j["child"], jidx, M)
            else:
                idx = len(model)
                model.append(dict(name=jn, parent=jidx, type=j["type"], axis=j["axis"], M=M, Y=Inertia()))
                frames[jn] = (idx, se3())
                visit(j["child"], idx, se3())

    visit("world", -1, se3())
    return model, frames, link_home


def joint_transform(j, q):
    if j["type"] == "revolute":
        a = j["axis"]
        assert np.allclose(a, [0, 0, 1])
        c, s = math.cos(q), math.sin(q)
        return se3([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    return se3(None, j["axis"] * q)


def fk_merged(model, frames, q):
    oMi = []
    for i, j in enumerate(model):
        li = se3_mul(j["M"], joint_transform(j, q[i]))
        oMi.append(li if j["parent"] < 0 else se3_mul(oMi[j["parent"]], li))
    out = {}
    for f in FRAMES_WANTED:
        p, M = frames[f]
        out[f] = se3_mul(oMi[p], M)[1]
    return out, oMi


def fk_unmerged(links, joints, q_by_name):
    """Independent chain walk over the raw URDF (no merging)."""
    children = {}
    for name, j in joints.items():
        children.setdefault(j["parent"], []).append(name)
    pos = {}

    def visit(link, M):
        for jn in children.get(link, []):
            j = joints[jn]
            Mj = se3_mul(M, j["M"])
            pos[jn] = Mj[1].copy()
            if j["type"] != "fixed":
                Mj = se3_mul(Mj, joint_transform(j, q_by_name[jn]))
            visit(j["child"], Mj)

    visit("world", se3())
    return pos


def fmt(x):
    return repr(float(x))


def carr(name, arr, dims):
    arr = np.asarray(arr, float)
    decl = "static const double %s%s = " % (name, "".join("[%d]" % d for d in dims))

    def rec(a):
        if a.ndim == 1:
            return "{" + ", ".join(fmt(v) for v in a) + "}"
        return "{\n  " + ",\n  ".join(rec(x) for x in a) + "}"
    return decl + rec(arr.reshape(dims)) + ";\n"


def main():
    urdf = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_URDF
    links, joints = parse(urdf)
    model, frames, link_home = build(links, joints)
    names = [j["name"] for j in model]
    assert names == EXPECTED_ORDER, names
    n = len(model)

    TYPE = {"PX": 0, "PY": 1, "RZ": 2, "PU": 3}
    jt = []
    for j in model:
        if j["type"] == "revolute":
            jt.append(TYPE["RZ"])
        elif np.allclose(j["axis"], [1, 0, 0]):
            jt.append(TYPE["PX"])
        elif np.allclose(j["axis"], [0, 1, 0]):
            jt.append(TYPE["PY"])
        else:
            jt.append(TYPE["PU"])

    presets = {  # state.cpp:5-49
        "ZERO": [0.0] * 12,
        "HUDDLED": [0.2, 0.2, math.pi / 4, 0.0, math.pi / 5, 0.0, -math.pi / 2, 0.0, 2, math.pi / 4, 0.025, 0.025],
        "REACH": [0.20, 0.20, math.pi / 4, 0.0, 1.5, 0.0, 0, 0, math.pi, math.pi / 4, 0.025, 0.025],
        "BEHIND": [0.20, 0.20, math.pi / 4, math.pi, 1.2, 0.0, -2, 0, math.pi / 2, math.pi / 4, 0.025, 0.025],
    }
    kats = {}
    for pname, q in presets.items():
        merged, oMi = fk_merged(model, frames, q)
        raw = fk_unmerged(links, joints, dict(zip(names, q)))
        for f in FRAMES_WANTED:
            assert np.allclose(merged[f], raw[f], atol=1e-14), (pname, f, merged[f], raw[f])
        # link COM positions in the world (true-FK self collision mode)
        coms = {}
        for ln in LINKS[3:11]:
            ji, _ = link_home[ln]
            R, p = oMi[ji]
            coms[ln] = (R @ model[ji]["Y"].c + p).tolist()
        kats[pname] = dict(q=q, frames={f: merged[f].tolist() for f in FRAMES_WANTED}, link_com=coms)

    # SURVEY Appendix B derived values (12 significant digits).
    assert np.allclose(kats["HUDDLED"]["frames"]["panda_grasp_joint"], [0.870297769478, 0.877368837289, 0.890776004667], atol=1e-11)
    assert np.allclose(kats["HUDDLED"]["frames"]["arm_mount_joint"], [0.405060966544, 0.412132034356, 0.725], atol=1e-11)
    assert np.allclose(kats["REACH"]["frames"]["panda_grasp_joint"], [1.036871909346, 1.043942977163, 1.209584514725], atol=1e-11)
    assert np.allclose(kats["ZERO"]["frames"]["panda_grasp_joint"], [0.383, 0.005, 1.556], atol=1e-10)

    link_joint = [link_home[l][0] for l in LINKS]
    ee_parent, ee_M = frames["panda_grasp_joint"]
    mt_parent, mt_M = frames["arm_mount_joint"]

    h = []
    h.append("/* GENERATED by tools/extract_model.py from the reference robot description\n"
             " * (src/frankaridgeback/model/robot.urdf). Do not edit.\n"
             " * 12 one-dof joints in FrankaRidgeback::State order; fixed joints merged the way\n"
             " * pinocchio::urdf::buildModel does. Rotations are row-major 3x3.\n"
             " * Joint types: 0 prismatic +x, 1 prismatic +y, 2 revolute +z, 3 prismatic along FR_AXIS.\n"
             " */\n#pragma once\n\n")
    h.append("#define FR_NJ %d\n" % n)
    h.append("#define FR_JT_PX 0\n#define FR_JT_PY 1\n#define FR_JT_RZ 2\n#define FR_JT_PU 3\n")
    h.append("#define FR_EE_PARENT %d\n#define FR_MOUNT_PARENT %d\n#define FR_NLINK %d\n\n" % (ee_parent, mt_parent, len(LINKS)))
    h.append("static const int FR_PARENT[%d] = {%s};\n" % (n, ", ".join(str(j["parent"]) for j in model)))
    h.append("static const int FR_JTYPE[%d] = {%s};\n" % (n, ", ".join(str(t) for t in jt)))
    h.append("/* Link enum (dynamics.hpp:65-80) -> owning joint (-1 = world) */\n")
    h.append("static const int FR_LINK_JOINT[%d] = {%s};\n" % (len(LINKS), ", ".join(str(x) for x in link_joint)))
    h.append(carr("FR_AXIS", [j["axis"] for j in model], (n, 3)))
    h.append(carr("FR_PLACE_R", [j["M"][0].reshape(9) for j in model], (n, 9)))
    h.append(carr("FR_PLACE_P", [j["M"][1] for j in model], (n, 3)))
    h.append(carr("FR_MASS", [j["Y"].m for j in model], (n,)))
    h.append(carr("FR_COM", [j["Y"].c for j in model], (n, 3)))
    h.append("/* rotational inertia about the COM, joint-frame axes: xx xy xz yy yz zz */\n")
    h.append(carr("FR_INERTIA", [[j["Y"].I[0, 0], j["Y"].I[0, 1], j["Y"].I[0, 2], j["Y"].I[1, 1], j["Y"].I[1, 2], j["Y"].I[2, 2]] for j in model], (n, 6)))
    h.append("/* end effector frame panda_grasp_joint (pinocchio_dynamics.hpp:58) in joint FR_EE_PARENT */\n")
    h.append(carr("FR_EE_R", ee_M[0].reshape(9), (9,)))
    h.append(carr("FR_EE_P", ee_M[1], (3,)))
    h.append("/* frame arm_mount_joint (assisted_manipulation.cpp:175) in joint FR_MOUNT_PARENT */\n")
    h.append(carr("FR_MOUNT_R", mt_M[0].reshape(9), (9,)))
    h.append(carr("FR_MOUNT_P", mt_M[1], (3,)))
    out_h = os.path.join(ROOT, "assistedmanipulation_b200", "csrc", "robot_model.h")
    os.makedirs(os.path.dirname(out_h), exist_ok=True)
    with open(out_h, "w") as f:
        f.write("".join(h))

    gold = dict(
        source="robot.urdf (reference src/frankaridgeback/model/), extracted by tools/extract_model.py",
        joints=[dict(name=j["name"], parent=j["parent"], type=jt[i], axis=j["axis"].tolist(),
                     R=j["M"][0].tolist(), p=j["M"][1].tolist(), mass=j["Y"].m, com=j["Y"].c.tolist(),
                     inertia=j["Y"].I.tolist()) for i, j in enumerate(model)],
        ee=dict(parent=ee_parent, R=ee_M[0].tolist(), p=ee_M[1].tolist()),
        mount=dict(parent=mt_parent, R=mt_M[0].tolist(), p=mt_M[1].tolist()),
        link_joint=link_joint,
        fk_kat=kats,
    )
    out_j = os.path.join(ROOT, "tests", "golden", "robot_model.json")
    os.makedirs(os.path.dirname(out_j), exist_ok=True)
    with open(out_j, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", out_h, "and", out_j)
    for i, j in enumerate(model):
        print(i, j["name"], "parent", j["parent"], "type", jt[i], "m=%.4f" % j["Y"].m, "com", np.round(j["Y"].c, 4))


if __name__ == "__main__":
    main()
