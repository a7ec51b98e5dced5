"""TEST INFRASTRUCTURE — a textbook derivation of the Franka+Ridgeback equations of motion that shares NO code and
no formulation with oracle/robot_oracle.hpp (spatial algebra: RNEA / ABA / CRBA): homogeneous transforms, per-body
geometric Jacobians, M(q) = sum_i m_i Jv_i^T Jv_i + Jw_i^T R_i I_i R_i^T Jw_i, gravity from the potential, Coriolis
forces from the Christoffel symbols of M (dM/dq by central differences). Input: the joint table extracted from the
URDF (tests/golden/robot_model.json). Used to cross-check the oracle where pinocchio itself is absent."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = json.load(open(os.path.join(ROOT, "tests", "golden", "robot_model.json")))
JOINTS = MODEL["joints"]
N = len(JOINTS)
GRAVITY = 9.81
REVOLUTE = 2


def _axis_rotation(axis, angle):
    a = np.asarray(axis, dtype=float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K   # Rodrigues


def forward(q):
    """World transform (4x4) of every joint frame after its own motion."""
    T = []
    for i, j in enumerate(JOINTS):
        fixed = np.eye(4)
        fixed[:3, :3], fixed[:3, 3] = np.array(j["R"]), np.array(j["p"])
        move = np.eye(4)
        if j["type"] == REVOLUTE:
            move[:3, :3] = _axis_rotation(j["axis"], q[i])
        else:
            move[:3, 3] = np.array(j["axis"]) * q[i]
        parent = T[j["parent"]] if j["parent"] >= 0 else np.eye(4)
        T.append(parent @ fixed @ move)
    return T


def ancestors(i):
    out = []
    while i >= 0:
        out.append(i)
        i = JOINTS[i]["parent"]
    return out


def mass_matrix_and_potential(q):
    T = forward(q)
    M = np.zeros((N, N))
    U = 0.0
    for i, body in enumerate(JOINTS):
        m = body["mass"]
        c = T[i][:3, :3] @ np.array(body["com"]) + T[i][:3, 3]
        Iw = T[i][:3, :3] @ np.array(body["inertia"]) @ T[i][:3, :3].T
        Jv, Jw = np.zeros((3, N)), np.zeros((3, N))
        for j in ancestors(i):
            a = T[j][:3, :3] @ np.array(JOINTS[j]["axis"])
            if JOINTS[j]["type"] == REVOLUTE:
                Jv[:, j] = np.cross(a, c - T[j][:3, 3])
                Jw[:, j] = a
            else:
                Jv[:, j] = a
        M += m * Jv.T @ Jv + Jw.T @ Iw @ Jw
        U += m * GRAVITY * c[2]
    return M, U


def nonlinear_effects(q, v, h=1e-6):
    """C(q, v) v + g(q) from the Lagrangian: d/dt(M v) - d/dq (1/2 v^T M v - U)."""
    q, v = np.asarray(q, dtype=float), np.asarray(v, dtype=float)
    dM = np.zeros((N, N, N))   # dM[k] = dM/dq_k
    dU = np.zeros(N)
    for k in range(N):
        e = np.zeros(N)
        e[k] = h
        Mp, Up = mass_matrix_and_potential(q + e)
        Mm, Um = mass_matrix_and_potential(q - e)
        dM[k] = (Mp - Mm) / (2 * h)
        dU[k] = (Up - Um) / (2 * h)
    Mdot = np.tensordot(v, dM, axes=(0, 0))
    return Mdot @ v - 0.5 * np.array([v @ dM[k] @ v for k in range(N)]) + dU


def end_effector_position(q):
    T = forward(q)
    ee = MODEL["ee"]
    return T[ee["parent"]][:3, :3] @ np.array(ee["p"]) + T[ee["parent"]][:3, 3]
