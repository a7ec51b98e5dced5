// TEST INFRASTRUCTURE — CPU oracle (see scalar.hpp header). FP64 restatement of the rigid-body
// arithmetic behind FrankaRidgeback::PinocchioDynamics (reference
// src/frankaridgeback/pinocchio_dynamics.cpp:153-260).
//
// PARITY UNPINNED at this boundary: the arithmetic lives in pinocchio 2.7.1
// (vcpkg_overlays/pinocchio/vcpkg.json:3), which is not in /root/reference and not installed
// here. What follows restates pinocchio's published algorithms and conventions:
//   * spatial vectors are [linear; angular] (Motion::toVector)
//   * nonLinearEffects = RNEA with zero acceleration and a_gf[0] = -gravity
//   * aba              = Featherstone's articulated body algorithm, local convention
//   * forwardKinematics(q,v,a) second order, updateFramePlacements oMf = oMi[parent]*placement
//   * computeFrameJacobian(..., WORLD): column j = oMi[j].act(S_j) for j supporting the frame
//   * getFrameVelocity/Acceleration(..., WORLD) = oMi[parent].act(v[parent] / a[parent])
// pinned only by self-consistency tests (tests/test_oracle_robot.py) and the URDF-derived
// FK known answers of SURVEY Appendix B.
#pragma once
#include "scalar.hpp"
#include "../assistedmanipulation_b200/csrc/robot_model.h"

namespace oracle {

template <class S> struct V3 {
    S x, y, z;
    V3() : x(0.0), y(0.0), z(0.0) {}
    V3(S a, S b, S c) : x(a), y(b), z(c) {}
};
template <class S> inline V3<S> operator+(const V3<S> &a, const V3<S> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> inline V3<S> operator-(const V3<S> &a, const V3<S> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> inline V3<S> operator*(const V3<S> &a, S s) { return {a.x * s, a.y * s, a.z * s}; }
template <class S> inline V3<S> cross(const V3<S> &a, const V3<S> &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class S> inline S dot(const V3<S> &a, const V3<S> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> inline S norm(const V3<S> &a) { return m_sqrt(dot(a, a)); }

// row-major 3x3
template <class S> struct M3 {
    S m[9];
    M3() { for (auto &e : m) e = S(0.0); }
    S &operator()(int r, int c) { return m[3 * r + c]; }
    const S &operator()(int r, int c) const { return m[3 * r + c]; }
    static M3 identity() { M3 r; r.m[0] = r.m[4] = r.m[8] = S(1.0); return r; }
};
template <class S> inline M3<S> operator*(const M3<S> &a, const M3<S> &b) {
    M3<S> r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}
template <class S> inline M3<S> operator+(const M3<S> &a, const M3<S> &b) { M3<S> r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
template <class S> inline M3<S> operator-(const M3<S> &a, const M3<S> &b) { M3<S> r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
template <class S> inline V3<S> operator*(const M3<S> &a, const V3<S> &v) {
    return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
template <class S> inline M3<S> transpose(const M3<S> &a) { M3<S> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = a(j, i); return r; }
template <class S> inline V3<S> tmul(const M3<S> &a, const V3<S> &v) {  // a^T v
    return {a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}
template <class S> inline M3<S> skew(const V3<S> &p) {
    M3<S> r;
    r(0, 1) = -p.z; r(0, 2) = p.y; r(1, 0) = p.z; r(1, 2) = -p.x; r(2, 0) = -p.y; r(2, 1) = p.x;
    return r;
}

template <class S> struct SE3 {
    M3<S> R; V3<S> p;
    SE3() : R(M3<S>::identity()) {}
};
template <class S> inline SE3<S> operator*(const SE3<S> &a, const SE3<S> &b) { SE3<S> r; r.R = a.R * b.R; r.p = a.R * b.p + a.p; return r; }

template <class S> struct Motion { V3<S> v, w; };   // [linear; angular]
template <class S> struct Force { V3<S> f, n; };    // [linear; angular]
template <class S> inline Motion<S> operator+(const Motion<S> &a, const Motion<S> &b) { return {a.v + b.v, a.w + b.w}; }
template <class S> inline Force<S> operator+(const Force<S> &a, const Force<S> &b) { return {a.f + b.f, a.n + b.n}; }
template <class S> inline Motion<S> operator*(const Motion<S> &a, S s) { return {a.v * s, a.w * s}; }

// SE3 actions (pinocchio SE3::act / actInv)
template <class S> inline Motion<S> act(const SE3<S> &M, const Motion<S> &m) { V3<S> Rw = M.R * m.w; return {M.R * m.v + cross(M.p, Rw), Rw}; }
template <class S> inline Motion<S> actInv(const SE3<S> &M, const Motion<S> &m) { return {tmul(M.R, m.v - cross(M.p, m.w)), tmul(M.R, m.w)}; }
template <class S> inline Force<S> act(const SE3<S> &M, const Force<S> &f) { V3<S> Rf = M.R * f.f; return {Rf, M.R * f.n + cross(M.p, Rf)}; }
// motion x motion, motion x* force
template <class S> inline Motion<S> mcross(const Motion<S> &a, const Motion<S> &b) { return {cross(a.w, b.v) + cross(a.v, b.w), cross(a.w, b.w)}; }
template <class S> inline Force<S> fcross(const Motion<S> &a, const Force<S> &f) { return {cross(a.w, f.f), cross(a.w, f.n) + cross(a.v, f.f)}; }
template <class S> inline S mdotf(const Motion<S> &m, const Force<S> &f) { return dot(m.v, f.f) + dot(m.w, f.n); }

// rigid body inertia (mass, com, rotational inertia about com)
template <class S> struct Inertia {
    S m; V3<S> c; M3<S> I;
    Force<S> operator*(const Motion<S> &a) const {  // pinocchio InertiaTpl::__mult__
        V3<S> f = (a.v - cross(c, a.w)) * m;
        return {f, I * a.w + cross(c, f)};
    }
};

// 6x6 articulated inertia in blocks: f = A v + B w ; n = B^T v + D w
template <class S> struct ArtInertia {
    M3<S> A, B, D;
    static ArtInertia from(const Inertia<S> &Y) {
        ArtInertia r;
        M3<S> C = skew(Y.c);
        r.A = M3<S>(); r.A(0, 0) = r.A(1, 1) = r.A(2, 2) = Y.m;
        M3<S> mC; for (int i = 0; i < 9; i++) mC.m[i] = C.m[i] * Y.m;
        r.B = M3<S>() - mC;              // -m [c]x
        r.D = Y.I - mC * C;              // I_c - m [c]x [c]x
        return r;
    }
    Force<S> operator*(const Motion<S> &a) const { return {A * a.v + B * a.w, tmul(B, a.v) + D * a.w}; }
    // child frame -> parent frame through M = (R,p): A' = RAR^T, B' = RBR^T - A'[p]x, D' = RDR^T - (RBR^T)^T[p]x + [p]x B'
    ArtInertia transformed(const SE3<S> &M) const {
        M3<S> Rt = transpose(M.R), P = skew(M.p);
        ArtInertia r;
        r.A = M.R * A * Rt;
        M3<S> Bb = M.R * B * Rt;
        r.B = Bb - r.A * P;
        r.D = M.R * D * Rt - transpose(Bb) * P + P * r.B;
        return r;
    }
};

struct FrModel {
    static constexpr int NJ = FR_NJ;
    template <class S> static SE3<S> placement(int i) {
        SE3<S> M;
        for (int k = 0; k < 9; k++) M.R.m[k] = S(FR_PLACE_R[i][k]);
        M.p = {S(FR_PLACE_P[i][0]), S(FR_PLACE_P[i][1]), S(FR_PLACE_P[i][2])};
        return M;
    }
    template <class S> static Inertia<S> inertia(int i) {
        Inertia<S> Y;
        Y.m = S(FR_MASS[i]);
        Y.c = {S(FR_COM[i][0]), S(FR_COM[i][1]), S(FR_COM[i][2])};
        const double *t = FR_INERTIA[i];
        Y.I(0, 0) = t[0]; Y.I(0, 1) = t[1]; Y.I(0, 2) = t[2];
        Y.I(1, 0) = t[1]; Y.I(1, 1) = t[3]; Y.I(1, 2) = t[4];
        Y.I(2, 0) = t[2]; Y.I(2, 1) = t[4]; Y.I(2, 2) = t[5];
        return Y;
    }
    template <class S> static Motion<S> subspace(int i) {  // S_i in the joint frame
        Motion<S> s;
        if (FR_JTYPE[i] == FR_JT_RZ) s.w = {S(0.0), S(0.0), S(1.0)};
        else s.v = {S(FR_AXIS[i][0]), S(FR_AXIS[i][1]), S(FR_AXIS[i][2])};
        return s;
    }
    template <class S> static SE3<S> joint_transform(int i, S q) {
        SE3<S> J;
        if (FR_JTYPE[i] == FR_JT_RZ) {
            S c = m_cos(q), s = m_sin(q);
            J.R(0, 0) = c; J.R(0, 1) = -s; J.R(1, 0) = s; J.R(1, 1) = c;
        } else {
            J.p = {S(FR_AXIS[i][0]) * q, S(FR_AXIS[i][1]) * q, S(FR_AXIS[i][2]) * q};
        }
        return J;
    }
};

// Workspace mirroring the pinocchio::Data fields the reference reads.
template <class S> struct RobotData {
    static constexpr int NJ = FR_NJ;
    SE3<S> liMi[NJ], oMi[NJ];
    Motion<S> v[NJ], a[NJ], a_gf[NJ], c[NJ];
    Force<S> f[NJ];
    ArtInertia<S> Yaba[NJ];
    Force<S> U[NJ], UDinv[NJ];
    S Dinv[NJ], u[NJ];
    S nle[NJ], ddq[NJ];
    SE3<S> oMf_ee, oMf_mount;
    S J[6][NJ];  // WORLD frame jacobian of the end effector frame, rows [linear; angular]
};

template <class S> inline void joint_placements(RobotData<S> &d, const S *q) {
    for (int i = 0; i < FR_NJ; i++) d.liMi[i] = FrModel::placement<S>(i) * FrModel::joint_transform<S>(i, q[i]);
}

// pinocchio::nonLinearEffects (rnea.hxx NLEForwardStep/NLEBackwardStep)
template <class S> inline void nonlinear_effects(RobotData<S> &d, const S *q, const S *qd) {
    const Motion<S> minus_g{{S(0.0), S(0.0), S(9.81)}, {}};
    joint_placements(d, q);
    for (int i = 0; i < FR_NJ; i++) {
        int p = FR_PARENT[i];
        Motion<S> Sj = FrModel::subspace<S>(i);
        Motion<S> vj = Sj * qd[i];
        d.v[i] = vj;
        if (p >= 0) d.v[i] = d.v[i] + actInv(d.liMi[i], d.v[p]);
        d.a_gf[i] = actInv(d.liMi[i], p >= 0 ? d.a_gf[p] : minus_g) + mcross(d.v[i], vj);
        Inertia<S> Y = FrModel::inertia<S>(i);
        d.f[i] = Y * d.a_gf[i] + fcross(d.v[i], Y * d.v[i]);
    }
    for (int i = FR_NJ - 1; i >= 0; i--) {
        int p = FR_PARENT[i];
        Motion<S> Sj = FrModel::subspace<S>(i);
        d.nle[i] = mdotf(Sj, d.f[i]);
        if (p >= 0) d.f[p] = d.f[p] + act(d.liMi[i], d.f[i]);
    }
}

// pinocchio::aba (aba.hxx AbaForwardStep1 / AbaBackwardStep / AbaForwardStep2), local convention
template <class S> inline void aba(RobotData<S> &d, const S *q, const S *qd, const S *tau) {
    const Motion<S> minus_g{{S(0.0), S(0.0), S(9.81)}, {}};
    joint_placements(d, q);
    for (int i = 0; i < FR_NJ; i++) {
        int p = FR_PARENT[i];
        Motion<S> vj = FrModel::subspace<S>(i) * qd[i];
        d.v[i] = vj;
        if (p >= 0) d.v[i] = d.v[i] + actInv(d.liMi[i], d.v[p]);
        d.c[i] = mcross(d.v[i], vj);
        Inertia<S> Y = FrModel::inertia<S>(i);
        d.Yaba[i] = ArtInertia<S>::from(Y);
        d.f[i] = fcross(d.v[i], Y * d.v[i]);
    }
    for (int i = FR_NJ - 1; i >= 0; i--) {
        int p = FR_PARENT[i];
        Motion<S> Sj = FrModel::subspace<S>(i);
        d.u[i] = tau[i] - mdotf(Sj, d.f[i]);
        d.U[i] = d.Yaba[i] * Sj;
        d.Dinv[i] = S(1.0) / mdotf(Sj, d.U[i]);
        d.UDinv[i] = {d.U[i].f * d.Dinv[i], d.U[i].n * d.Dinv[i]};
        if (p >= 0) {
            // Ia = Yaba - UDinv U^T
            ArtInertia<S> Ia = d.Yaba[i];
            const S ud[6] = {d.UDinv[i].f.x, d.UDinv[i].f.y, d.UDinv[i].f.z, d.UDinv[i].n.x, d.UDinv[i].n.y, d.UDinv[i].n.z};
            const S uu[6] = {d.U[i].f.x, d.U[i].f.y, d.U[i].f.z, d.U[i].n.x, d.U[i].n.y, d.U[i].n.z};
            for (int r = 0; r < 3; r++)
                for (int cc = 0; cc < 3; cc++) {
                    Ia.A(r, cc) = Ia.A(r, cc) - ud[r] * uu[cc];
                    Ia.B(r, cc) = Ia.B(r, cc) - ud[r] * uu[3 + cc];
                    Ia.D(r, cc) = Ia.D(r, cc) - ud[3 + r] * uu[3 + cc];
                }
            Force<S> pa = d.f[i] + Ia * d.c[i];
            pa = pa + Force<S>{d.UDinv[i].f * d.u[i], d.UDinv[i].n * d.u[i]};
            ArtInertia<S> Ip = Ia.transformed(d.liMi[i]);
            d.Yaba[p].A = d.Yaba[p].A + Ip.A;
            d.Yaba[p].B = d.Yaba[p].B + Ip.B;
            d.Yaba[p].D = d.Yaba[p].D + Ip.D;
            d.f[p] = d.f[p] + act(d.liMi[i], pa);
        }
    }
    for (int i = 0; i < FR_NJ; i++) {
        int p = FR_PARENT[i];
        Motion<S> Sj = FrModel::subspace<S>(i);
        d.a_gf[i] = actInv(d.liMi[i], p >= 0 ? d.a_gf[p] : minus_g) + d.c[i];
        // ddq = Dinv*u - UDinv^T a
        Motion<S> ag = d.a_gf[i];
        S proj = dot(d.UDinv[i].f, ag.v) + dot(d.UDinv[i].n, ag.w);
        d.ddq[i] = d.Dinv[i] * d.u[i] - proj;
        d.a_gf[i] = d.a_gf[i] + Sj * d.ddq[i];
    }
}

// Composite rigid body algorithm: joint space inertia matrix (for self-consistency tests only).
template <class S> inline void crba(RobotData<S> &d, const S *q, S *Mout /* NJ*NJ row-major */) {
    joint_placements(d, q);
    ArtInertia<S> Yc[FR_NJ];
    for (int i = 0; i < FR_NJ; i++) Yc[i] = ArtInertia<S>::from(FrModel::inertia<S>(i));
    for (int i = 0; i < FR_NJ * FR_NJ; i++) Mout[i] = S(0.0);
    for (int i = FR_NJ - 1; i >= 0; i--) {
        int p = FR_PARENT[i];
        if (p >= 0) {
            ArtInertia<S> t = Yc[i].transformed(d.liMi[i]);
            Yc[p].A = Yc[p].A + t.A; Yc[p].B = Yc[p].B + t.B; Yc[p].D = Yc[p].D + t.D;
        }
    }
    for (int i = 0; i < FR_NJ; i++) {
        Force<S> F = Yc[i] * FrModel::subspace<S>(i);
        Mout[i * FR_NJ + i] = mdotf(FrModel::subspace<S>(i), F);
        int j = i;
        while (FR_PARENT[j] >= 0) {
            F = act(d.liMi[j], F);
            j = FR_PARENT[j];
            Mout[i * FR_NJ + j] = Mout[j * FR_NJ + i] = mdotf(FrModel::subspace<S>(j), F);
        }
    }
}

// pinocchio::forwardKinematics(q, v, a) + updateFramePlacements + computeFrameJacobian(WORLD)
template <class S> inline void forward_kinematics2(RobotData<S> &d, const S *q, const S *qd, const S *qdd) {
    joint_placements(d, q);
    for (int i = 0; i < FR_NJ; i++) {
        int p = FR_PARENT[i];
        Motion<S> Sj = FrModel::subspace<S>(i);
        Motion<S> vj = Sj * qd[i];
        d.oMi[i] = p >= 0 ? d.oMi[p] * d.liMi[i] : d.liMi[i];
        d.v[i] = vj;
        if (p >= 0) d.v[i] = d.v[i] + actInv(d.liMi[i], d.v[p]);
        d.a[i] = Sj * qdd[i] + mcross(d.v[i], vj);
        if (p >= 0) d.a[i] = d.a[i] + actInv(d.liMi[i], d.a[p]);
    }
    SE3<S> ee, mt;
    for (int k = 0; k < 9; k++) { ee.R.m[k] = S(FR_EE_R[k]); mt.R.m[k] = S(FR_MOUNT_R[k]); }
    ee.p = {S(FR_EE_P[0]), S(FR_EE_P[1]), S(FR_EE_P[2])};
    mt.p = {S(FR_MOUNT_P[0]), S(FR_MOUNT_P[1]), S(FR_MOUNT_P[2])};
    d.oMf_ee = d.oMi[FR_EE_PARENT] * ee;
    d.oMf_mount = d.oMi[FR_MOUNT_PARENT] * mt;
}

template <class S> inline void frame_jacobian_world(RobotData<S> &d) {
    for (int r = 0; r < 6; r++) for (int j = 0; j < FR_NJ; j++) d.J[r][j] = S(0.0);
    for (int j = FR_EE_PARENT; j >= 0; j = FR_PARENT[j]) {
        Motion<S> col = act(d.oMi[j], FrModel::subspace<S>(j));
        d.J[0][j] = col.v.x; d.J[1][j] = col.v.y; d.J[2][j] = col.v.z;
        d.J[3][j] = col.w.x; d.J[4][j] = col.w.y; d.J[5][j] = col.w.z;
    }
}

}  // namespace oracle
