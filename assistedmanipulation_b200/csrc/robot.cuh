// Rigid-body step of the Franka Research 3 + Ridgeback model for ONE rollout, held by ONE thread.
// Replaces, per sample, the pinocchio calls of FrankaRidgeback::PinocchioDynamics::calculate
// (reference src/frankaridgeback/pinocchio_dynamics.cpp:153-224): nonLinearEffects (RNEA),
// aba, forwardKinematics, updateFramePlacements, computeFrameJacobian(WORLD),
// getFrameVelocity(WORLD).
//
// Written for the GPU, not translated: the 12-joint topology is fixed at compile time (every
// joint loop is a template recursion, so joint type and parent are constants in the SASS), all
// joints are axis aligned (x, y or z — panda_finger_joint2's -y axis is folded into a sign), the
// articulated inertia is carried as three 3x3 blocks with symmetric storage, and two evaluation
// modes exist (include/mppi_b200.h MPPI_B200_DYNAMICS_*):
//   FAITHFUL  tau = u + nle(q,v); a = ABA(q,v,tau)      — operation for operation what the reference asks of pinocchio
//   FUSED     a = ABA(q, 0, u) without gravity           — identical in exact arithmetic (nle cancels), ~40 % fewer flops;
//             nle is evaluated only when the energy tank needs tau^T v.
// Host/device so tests can compare against the oracle on the CPU.
#pragma once
#include "spatial.cuh"

namespace mppi_b200 {

constexpr int NJ = 12;
enum JointType { JT_PX = 0, JT_PY = 1, JT_RZ = 2 };

// Topology (checked against robot_model.h at engine creation).
template <int I> struct Joint {
    static constexpr int parent = (I == 11) ? 9 : I - 1;
    static constexpr int type = (I == 0) ? JT_PX : ((I == 1 || I >= 10) ? JT_PY : JT_RZ);
};

// Model constants in the arithmetic of the kernel. One instance per precision in __constant__ memory.
template <class R> struct RobotModel {
    R place_R[NJ][9];  // fixed placement of joint i in its parent joint frame
    R place_p[NJ][3];
    R sign[NJ];        // +1, or -1 for a joint whose axis points along the negative coordinate axis
    R mass[NJ];
    R mc[NJ][3];       // mass * com
    R Io[NJ][6];       // rotational inertia about the joint origin: xx xy xz yy yz zz
    R com[NJ][3];
    R ee_p[3];         // end effector frame origin in joint 9
    R mount_p[3];      // arm_mount_joint frame origin in joint 2
    R gravity;         // 9.81
};

// What the objective needs from the kinematics of one calculate() call. These are the values the
// reference's costs read "stale" at the next step.
template <class R> struct Kinematics {
    Vec3<R> ee_pos;      // data.oMf[ee].translation()
    Vec3<R> mount_pos;   // data.oMf[arm_mount_joint].translation()
    Vec3<R> ee_lin_vel;  // getFrameVelocity(ee, WORLD).linear()
    R manip_det;         // det(J_a J_a^T), J_a = rows 0-2 / columns 3-9 of the WORLD jacobian
    Vec3<R> link_com[8]; // world COM of Link::PIVOT, PANDA_LINK1..7
};

enum KinFlags { KIN_MOUNT = 1, KIN_VEL = 2, KIN_MANIP = 4, KIN_LINKS = 8 };

template <class R> MPPI_HD Art<R> body_inertia(const RobotModel<R> &M, int i) {
    Art<R> a;
    const R m = M.mass[i], cx = M.mc[i][0], cy = M.mc[i][1], cz = M.mc[i][2];
    a.A.xx = m; a.A.yy = m; a.A.zz = m; a.A.xy = R(0); a.A.xz = R(0); a.A.yz = R(0);
    // B = -m [c]x
    a.B.m[0] = R(0); a.B.m[1] = cz;   a.B.m[2] = -cy;
    a.B.m[3] = -cz;  a.B.m[4] = R(0); a.B.m[5] = cx;
    a.B.m[6] = cy;   a.B.m[7] = -cx;  a.B.m[8] = R(0);
    a.D.xx = M.Io[i][0]; a.D.xy = M.Io[i][1]; a.D.xz = M.Io[i][2]; a.D.yy = M.Io[i][3]; a.D.yz = M.Io[i][4]; a.D.zz = M.Io[i][5];
    return a;
}

// Y * motion for the rigid body of joint i: f = m v - mc x w ; n = mc x v + Io w
template <class R> MPPI_HD Frc<R> body_mul(const RobotModel<R> &M, int i, const Mot<R> &a) {
    const Vec3<R> mc = v3<R>(M.mc[i][0], M.mc[i][1], M.mc[i][2]);
    Sym3<R> Io; Io.xx = M.Io[i][0]; Io.xy = M.Io[i][1]; Io.xz = M.Io[i][2]; Io.yy = M.Io[i][3]; Io.yz = M.Io[i][4]; Io.zz = M.Io[i][5];
    Frc<R> f;
    f.f = a.v * M.mass[i] - cross(mc, a.w);
    f.n = cross(mc, a.v) + mul(Io, a.w);
    return f;
}

template <class R, int TYPE> MPPI_HD Mot<R> joint_motion(R qd) {  // S * qd
    Mot<R> m; m.v = v3<R>(R(0), R(0), R(0)); m.w = m.v;
    if (TYPE == JT_PX) m.v.x = qd; else if (TYPE == JT_PY) m.v.y = qd; else m.w.z = qd;
    return m;
}
// v x (S qd) specialised by joint type (skips the structurally zero products)
template <class R, int TYPE> MPPI_HD Mot<R> cross_joint(const Mot<R> &a, R qd) {
    Mot<R> o;
    if (TYPE == JT_RZ) {       // b = (0 ; 0,0,qd)
        o.v = v3<R>(a.v.y * qd, -(a.v.x * qd), R(0));
        o.w = v3<R>(a.w.y * qd, -(a.w.x * qd), R(0));
    } else if (TYPE == JT_PX) {  // b = (qd,0,0 ; 0): o.v = a.w x b.v
        o.v = v3<R>(R(0), a.w.z * qd, -(a.w.y * qd));
        o.w = v3<R>(R(0), R(0), R(0));
    } else {
        o.v = v3<R>(-(a.w.z * qd), R(0), a.w.x * qd);
        o.w = v3<R>(R(0), R(0), R(0));
    }
    return o;
}
template <class R, int TYPE> MPPI_HD R joint_dot(const Frc<R> &f) {  // S^T f
    return TYPE == JT_PX ? f.f.x : (TYPE == JT_PY ? f.f.y : f.n.z);
}
template <class R, int TYPE> MPPI_HD R joint_dot(const Mot<R> &m) {
    return TYPE == JT_PX ? m.v.x : (TYPE == JT_PY ? m.v.y : m.w.z);
}

// Per-step scratch. Arrays are indexed with compile-time constants only (template recursion), so
// ptxas keeps what fits in registers and parks the rest in thread-local memory.
template <class R> struct Scratch {
    Xf<R> li[NJ];      // parent <- joint transforms
    Mot<R> v[NJ];      // body velocities (joint frame)
    Mot<R> c[NJ];      // velocity-product accelerations
    Frc<R> f[NJ];      // RNEA forces
    Frc<R> pA[NJ];     // ABA bias forces
    Frc<R> U[NJ];      // Yaba * S
    R Dinv[NJ], u[NJ];
    Art<R> Ia;         // articulated inertia being propagated down the chain
    Art<R> Ia9;        // joint 9 collects two children
};

// cs / sn: the joint's cosine / sine when the caller already has them (FUSED mode shares one set per step), else null
template <class R, int I> MPPI_HD void joint_transform(const RobotModel<R> &M, R q, Xf<R> &X, const R *cs = nullptr, const R *sn = nullptr) {
    constexpr int T = Joint<I>::type;
    const R *P = M.place_R[I];
    if (T == JT_RZ) {
        R s, c;
        if (cs) { s = sn[I]; c = cs[I]; } else sincos_(q, &s, &c);
#pragma unroll
        for (int r = 0; r < 3; r++) {
            X.R_.m[3 * r + 0] = P[3 * r + 0] * c + P[3 * r + 1] * s;
            X.R_.m[3 * r + 1] = P[3 * r + 1] * c - P[3 * r + 0] * s;
            X.R_.m[3 * r + 2] = P[3 * r + 2];
        }
        X.p = v3<R>(M.place_p[I][0], M.place_p[I][1], M.place_p[I][2]);
    } else {
        constexpr int ax = (T == JT_PX) ? 0 : 1;
#pragma unroll
        for (int k = 0; k < 9; k++) X.R_.m[k] = P[k];
        X.p = v3<R>(M.place_p[I][0] + P[0 + ax] * q, M.place_p[I][1] + P[3 + ax] * q, M.place_p[I][2] + P[6 + ax] * q);
    }
}

// ---- PLANE: the transforms of joints 0..9 by their structure ---------------------------------------------------------
// FUSED mode has the step's joint sines / cosines at hand and (checked by fast_structure_matches at engine creation)
// joints 0, 1 translate along x / y without rotating, joint 2 is Rz(q) in place, and every arm joint is Rx(alpha) Rz(q)
// behind a fixed translation: motions and forces cross a joint by two plane rotations (8 operations per vector) and one
// cross product instead of two 3x3 products (30), and no transform matrix is formed. The fingers keep the generic path.
// Structural zeros of those offsets (bit k set = component k can be non-zero): the sliders move along one axis, the arm
// joints' fixed offsets are (x, y, z), 0, (0, y, 0), (x, 0, 0), (x, y, 0), 0, (x, 0, 0) for joints 3..9 — checked
// against the generated model by fast_structure_matches at engine creation. Products with the zeros are dropped
// (cross_add_m / cross_sub_m, spatial.cuh).
MPPI_HD constexpr unsigned offset_mask(int i) {
    return i == 0 ? 1u : i == 1 ? 2u : i == 3 ? 7u : i == 5 ? 2u : i == 6 ? 1u : i == 7 ? 3u : i == 9 ? 1u : (i >= 10 ? 7u : 0u);
}
// The first arm joint's fixed placement does not rotate (alpha = 0: the arm stands upright on its mount); the others
// turn by ~ +-90 degrees about x, with the URDF's rounded pi/2, so their cosine is 4.9e-12 and stays in the arithmetic.
MPPI_HD constexpr bool placement_is_flat(int i) { return i == 3; }
template <class R, int I> MPPI_HD Vec3<R> joint_offset(const RobotModel<R> &M, const R *q) {
    if (I == 0) return v3<R>(q[0] * M.sign[0], R(0), R(0));
    if (I == 1) return v3<R>(R(0), q[1] * M.sign[1], R(0));
    return v3<R>(M.place_p[I][0], M.place_p[I][1], M.place_p[I][2]);
}
template <class R, int I> MPPI_HD Vec3<R> rot_to_joint(const RobotModel<R> &M, const R *cs, const R *sn, const Vec3<R> &v) {   // E^T v
    if (I < 2) return v;
    if (I == 2 || placement_is_flat(I)) return rotz_t(cs[I], sn[I], v);
    return rotz_t(cs[I], sn[I], rotx_t(M.place_R[I][4], M.place_R[I][7], v));
}
template <class R, int I> MPPI_HD Vec3<R> rot_to_parent(const RobotModel<R> &M, const R *cs, const R *sn, const Vec3<R> &v) {   // E v
    if (I < 2) return v;
    if (I == 2 || placement_is_flat(I)) return rotz(cs[I], sn[I], v);
    return rotx(M.place_R[I][4], M.place_R[I][7], rotz(cs[I], sn[I], v));
}
template <class R, int I> MPPI_HD Mot<R> act_inv_joint(const RobotModel<R> &M, const R *q, const R *cs, const R *sn, const Mot<R> &m) {
    Mot<R> o;
    o.v = rot_to_joint<R, I>(M, cs, sn, (I == 2) ? m.v : cross_sub_m<offset_mask(I)>(joint_offset<R, I>(M, q), m.w, m.v));
    o.w = rot_to_joint<R, I>(M, cs, sn, m.w);
    return o;
}
template <class R, int I> MPPI_HD Frc<R> act_joint(const RobotModel<R> &M, const R *q, const R *cs, const R *sn, const Frc<R> &f) {
    Frc<R> o;
    o.f = rot_to_parent<R, I>(M, cs, sn, f.f);
    const Vec3<R> n = rot_to_parent<R, I>(M, cs, sn, f.n);
    o.n = (I == 2) ? n : cross_add_m<offset_mask(I)>(joint_offset<R, I>(M, q), o.f, n);
    return o;
}

// ---- pass 1: transforms, velocities, RNEA forces ----------------------------------------------
// STOP: last joint of this call's recursion (the rolled build runs joints 0..2, the arm loop, then 10..11)
template <class R, int I, bool VEL, bool NLE, bool BIAS, bool PLANE = false, int STOP = NJ - 1>
MPPI_HD void pass1(const RobotModel<R> &M, const R *q, const R *qd, Scratch<R> &S, Mot<R> *agf, const R *cs = nullptr, const R *sn = nullptr) {
    constexpr int P = Joint<I>::parent, T = Joint<I>::type;
    constexpr bool PL = PLANE && I <= 9;   // this joint's transform by its structure (no matrix formed)
    const R sg = M.sign[I];
    if (!PL) joint_transform<R, I>(M, q[I] * sg, S.li[I], cs, sn);
    if (PL && I < 2) {
        // The sliders translate without turning, and nothing below them turns either: their angular velocity and the
        // velocity-product terms vanish identically (v x* (m v) = 0), their spatial acceleration is gravity alone. Written
        // out, because the generic path multiplies through those zeros (~95 instructions per slider; the compiler may not
        // fold 0 * x).
        if (VEL) {
            const R w = qd[I] * sg;
            const Vec3<R> zero = v3<R>(R(0), R(0), R(0));
            S.v[I].v = (I == 0) ? v3<R>(w, R(0), R(0)) : v3<R>(S.v[0].v.x, w, R(0));
            S.v[I].w = zero;
            if (NLE || BIAS) {
                S.c[I].v = zero; S.c[I].w = zero;
                if (NLE) {
                    agf[I].v = v3<R>(R(0), R(0), M.gravity); agf[I].w = zero;
                    S.f[I].f = v3<R>(R(0), R(0), M.gravity * M.mass[I]);
                    S.f[I].n = v3<R>(M.mc[I][1] * M.gravity, -(M.mc[I][0] * M.gravity), R(0));
                }
                if (BIAS) { S.pA[I].f = zero; S.pA[I].n = zero; }
            }
        }
    } else
    if (VEL) {
        const R w = qd[I] * sg;
        Mot<R> vj = joint_motion<R, T>(w);
        if (P >= 0) {
            Mot<R> vp;
            if (PL) vp = act_inv_joint<R, I>(M, q, cs, sn, S.v[P < 0 ? 0 : P]); else vp = act_inv(S.li[I], S.v[P < 0 ? 0 : P]);
            S.v[I].v = vp.v + vj.v; S.v[I].w = vp.w + vj.w;
        }
        else S.v[I] = vj;
        if (NLE || BIAS) {
            S.c[I] = cross_joint<R, T>(S.v[I], w);
            Frc<R> h = body_mul(M, I, S.v[I]);
            Frc<R> vxh = fcross(S.v[I], h);
            if (NLE) {
                Mot<R> ap;
                if (P >= 0) { if (PL) ap = act_inv_joint<R, I>(M, q, cs, sn, agf[P < 0 ? 0 : P]); else ap = act_inv(S.li[I], agf[P < 0 ? 0 : P]); }
                else if (PL) { ap.v = v3<R>(R(0), R(0), M.gravity); ap.w = v3<R>(R(0), R(0), R(0)); }   // the root joint does not rotate
                else { ap.v = tmul(S.li[I].R_, v3<R>(R(0), R(0), M.gravity)); ap.w = v3<R>(R(0), R(0), R(0)); }
                agf[I].v = ap.v + S.c[I].v; agf[I].w = ap.w + S.c[I].w;
                Frc<R> ya = body_mul(M, I, agf[I]);
                S.f[I].f = ya.f + vxh.f; S.f[I].n = ya.n + vxh.n;
            }
            if (BIAS) S.pA[I] = vxh;
        }
    }
    if (I + 1 <= STOP) pass1<R, (I + 1 <= STOP ? I + 1 : I), VEL, NLE, BIAS, PLANE, STOP>(M, q, qd, S, agf, cs, sn);
}

// ---- RNEA backward: nle_i = S^T f_i ; f_parent += X f_i ----------------------------------------
// FIRST: lowest joint of this call's recursion
template <class R, int I, bool PLANE = false, int FIRST = 0> MPPI_HD void rnea_back(const RobotModel<R> &M, Scratch<R> &S, R *nle, const R *q = nullptr, const R *cs = nullptr, const R *sn = nullptr) {
    constexpr int P = Joint<I>::parent, T = Joint<I>::type;
    nle[I] = joint_dot<R, T>(S.f[I]) * M.sign[I];
    if (P >= 0) {
        Frc<R> fp;
        if (PLANE && I <= 9) fp = act_joint<R, I>(M, q, cs, sn, S.f[I]); else fp = act(S.li[I], S.f[I]);
        S.f[P < 0 ? 0 : P].f = S.f[P < 0 ? 0 : P].f + fp.f;
        S.f[P < 0 ? 0 : P].n = S.f[P < 0 ? 0 : P].n + fp.n;
    }
    if (I > FIRST) rnea_back<R, (I > FIRST ? I - 1 : FIRST), PLANE, FIRST>(M, S, nle, q, cs, sn);
}

// A P^ and P^ A helpers, P^ = [p]x
template <class R> MPPI_HD Mat3<R> mul_skew(const Mat3<R> &m, const Vec3<R> &p) {  // m * [p]x
    Mat3<R> o;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        o(i, 0) = m(i, 1) * p.z - m(i, 2) * p.y;
        o(i, 1) = m(i, 2) * p.x - m(i, 0) * p.z;
        o(i, 2) = m(i, 0) * p.y - m(i, 1) * p.x;
    }
    return o;
}
template <class R> MPPI_HD Mat3<R> sym_mul_skew(const Sym3<R> &s, const Vec3<R> &p) {
    Mat3<R> m;
    m.m[0] = s.xx; m.m[1] = s.xy; m.m[2] = s.xz; m.m[3] = s.xy; m.m[4] = s.yy; m.m[5] = s.yz; m.m[6] = s.xz; m.m[7] = s.yz; m.m[8] = s.zz;
    return mul_skew(m, p);
}

// Ia (joint frame of child) -> parent frame through X, accumulated into `dst`.
template <class R> MPPI_HD void art_transform_add(const Art<R> &Ia, const Xf<R> &X, Art<R> &dst) {
    const Sym3<R> A = conj(X.R_, Ia.A);
    const Mat3<R> Bb = conj(X.R_, Ia.B);
    const Sym3<R> Db = conj(X.R_, Ia.D);
    const Mat3<R> AP = sym_mul_skew(A, X.p);
    Mat3<R> Bn;
#pragma unroll
    for (int k = 0; k < 9; k++) Bn.m[k] = Bb.m[k] - AP.m[k];
    // D' = Db - Bb^T P^ + P^ B'   (only the upper triangle)
    const Vec3<R> p = X.p;
    // (Bb^T P^)(i,j): row i of Bb^T = column i of Bb
    auto btp = [&](int i, int j) -> R {
        const R b0 = Bb(0, i), b1 = Bb(1, i), b2 = Bb(2, i);
        return j == 0 ? (b1 * p.z - b2 * p.y) : (j == 1 ? (b2 * p.x - b0 * p.z) : (b0 * p.y - b1 * p.x));
    };
    auto pb = [&](int i, int j) -> R {  // (P^ B')(i,j)
        return i == 0 ? (p.y * Bn(2, j) - p.z * Bn(1, j)) : (i == 1 ? (p.z * Bn(0, j) - p.x * Bn(2, j)) : (p.x * Bn(1, j) - p.y * Bn(0, j)));
    };
    dst.A.xx += A.xx; dst.A.xy += A.xy; dst.A.xz += A.xz; dst.A.yy += A.yy; dst.A.yz += A.yz; dst.A.zz += A.zz;
#pragma unroll
    for (int k = 0; k < 9; k++) dst.B.m[k] += Bn.m[k];
    dst.D.xx += Db.xx - btp(0, 0) + pb(0, 0);
    dst.D.xy += Db.xy - btp(0, 1) + pb(0, 1);
    dst.D.xz += Db.xz - btp(0, 2) + pb(0, 2);
    dst.D.yy += Db.yy - btp(1, 1) + pb(1, 1);
    dst.D.yz += Db.yz - btp(1, 2) + pb(1, 2);
    dst.D.zz += Db.zz - btp(2, 2) + pb(2, 2);
}

template <class R> MPPI_HD Frc<R> art_mul(const Art<R> &I, const Mot<R> &a) {
    Frc<R> f;
    f.f = mul(I.A, a.v) + mul(I.B, a.w);
    f.n = tmul(I.B, a.v) + mul(I.D, a.w);
    return f;
}

// ---- ABA backward pass -------------------------------------------------------------------------
// `cur` holds the articulated inertia of joint I (its own body + whatever its children added).
template <class R, int I, bool BIAS>
MPPI_HD void aba_back(const RobotModel<R> &M, const R *tau, Scratch<R> &S, Art<R> &cur) {
    constexpr int P = Joint<I>::parent, T = Joint<I>::type;
    // U = Yaba * S : a column of the 6x6
    Frc<R> U;
    R D;
    if (T == JT_PX) { U.f = v3<R>(cur.A.xx, cur.A.xy, cur.A.xz); U.n = v3<R>(cur.B(0, 0), cur.B(0, 1), cur.B(0, 2)); D = cur.A.xx; }
    else if (T == JT_PY) { U.f = v3<R>(cur.A.xy, cur.A.yy, cur.A.yz); U.n = v3<R>(cur.B(1, 0), cur.B(1, 1), cur.B(1, 2)); D = cur.A.yy; }
    else { U.f = v3<R>(cur.B(0, 2), cur.B(1, 2), cur.B(2, 2)); U.n = v3<R>(cur.D.xz, cur.D.yz, cur.D.zz); D = cur.D.zz; }
    const R Dinv = R(1) / D;
    // Without bias terms pA is only what the children pushed down: nothing for the two leaves.
    constexpr bool HAS_PA = BIAS || I < 10;
    R u = tau[I] * M.sign[I];
    if (HAS_PA) u -= joint_dot<R, T>(S.pA[I]);
    S.U[I] = U; S.Dinv[I] = Dinv; S.u[I] = u;
    if (P >= 0) {
        Frc<R> UD; UD.f = U.f * Dinv; UD.n = U.n * Dinv;
        Art<R> Ia = cur;
        Ia.A.xx -= UD.f.x * U.f.x; Ia.A.xy -= UD.f.x * U.f.y; Ia.A.xz -= UD.f.x * U.f.z;
        Ia.A.yy -= UD.f.y * U.f.y; Ia.A.yz -= UD.f.y * U.f.z; Ia.A.zz -= UD.f.z * U.f.z;
        const R uf[3] = {UD.f.x, UD.f.y, UD.f.z}, un[3] = {U.n.x, U.n.y, U.n.z};
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) Ia.B(r, cc) -= uf[r] * un[cc];
        Ia.D.xx -= UD.n.x * U.n.x; Ia.D.xy -= UD.n.x * U.n.y; Ia.D.xz -= UD.n.x * U.n.z;
        Ia.D.yy -= UD.n.y * U.n.y; Ia.D.yz -= UD.n.y * U.n.z; Ia.D.zz -= UD.n.z * U.n.z;
        // pa = pA + Ia c + UDinv u
        Frc<R> pa;
        pa.f = UD.f * u; pa.n = UD.n * u;
        if (HAS_PA) { pa.f = pa.f + S.pA[I].f; pa.n = pa.n + S.pA[I].n; }
        if (BIAS) {
            Frc<R> ic = art_mul(Ia, S.c[I]);
            pa.f = pa.f + ic.f; pa.n = pa.n + ic.n;
        }
        Frc<R> pp = act(S.li[I], pa);
        constexpr int PP = P < 0 ? 0 : P;
        // with bias terms every pA starts as v x* (Y v); without, the first child to arrive assigns
        if (BIAS || I == 10) { S.pA[PP].f = S.pA[PP].f + pp.f; S.pA[PP].n = S.pA[PP].n + pp.n; }
        else S.pA[PP] = pp;
        // articulated inertia of the parent: its own body first (chain), or the collector (joint 9)
        if (I == 11) { S.Ia9 = body_inertia(M, 9); art_transform_add(Ia, S.li[I], S.Ia9); }
        else if (I == 10) { art_transform_add(Ia, S.li[I], S.Ia9); }
        else { Art<R> next = body_inertia(M, PP); art_transform_add(Ia, S.li[I], next); cur = next; }
    }
}

template <class R, int I, bool BIAS>
MPPI_HD void aba_back_all(const RobotModel<R> &M, const R *tau, Scratch<R> &S, Art<R> &cur) {
    if (I == 11 || I == 10) cur = body_inertia(M, I);
    if (I == 9) cur = S.Ia9;
    aba_back<R, I, BIAS>(M, tau, S, cur);
    if (I > 0) aba_back_all<R, (I > 0 ? I - 1 : 0), BIAS>(M, tau, S, cur);
}

// ---- ABA forward pass ---------------------------------------------------------------------------
template <class R, int I, bool BIAS>
MPPI_HD void aba_fwd(const RobotModel<R> &M, Scratch<R> &S, Mot<R> *a, R *qdd) {
    constexpr int P = Joint<I>::parent, T = Joint<I>::type;
    Mot<R> ap;
    if (P >= 0) ap = act_inv(S.li[I], a[P < 0 ? 0 : P]);
    else if (BIAS) { ap.v = tmul(S.li[I].R_, v3<R>(R(0), R(0), M.gravity)); ap.w = v3<R>(R(0), R(0), R(0)); }
    else { ap.v = v3<R>(R(0), R(0), R(0)); ap.w = ap.v; }
    if (BIAS) { ap.v = ap.v + S.c[I].v; ap.w = ap.w + S.c[I].w; }
    const R dd = S.Dinv[I] * (S.u[I] - (dot(S.U[I].f, ap.v) + dot(S.U[I].n, ap.w)));
    qdd[I] = dd * M.sign[I];
    a[I] = ap;
    if (T == JT_PX) a[I].v.x += dd; else if (T == JT_PY) a[I].v.y += dd; else a[I].w.z += dd;
    if (I + 1 < NJ) aba_fwd<R, (I + 1 < NJ ? I + 1 : I), BIAS>(M, S, a, qdd);
}

// ---- world kinematics for the objective ----------------------------------------------------------
template <class R, int I, int FLAGS, bool PLANE = false, int STOP = 9>
MPPI_HD void world_chain(const RobotModel<R> &M, const Scratch<R> &S, Xf<R> &oM, Kinematics<R> &K, R *Jl /* 3 x 7 */, const R *q = nullptr, const R *cs = nullptr, const R *sn = nullptr) {
    // oM enters as oMi[I-1], leaves as oMi[I]
    if (PLANE) {
        if (I == 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) oM.R_.m[k] = (k % 4 == 0) ? R(1) : R(0);
            oM.p = joint_offset<R, 0>(M, q);
        } else if (I == 1) {
            oM.p.y = q[1] * M.sign[1];
        } else if (I == 2) {   // Rz(q2) itself: the frame above does not rotate
            oM.R_.m[0] = cs[2]; oM.R_.m[1] = -sn[2]; oM.R_.m[3] = sn[2]; oM.R_.m[4] = cs[2];
        } else {
            oM.p = mul(oM.R_, joint_offset<R, I>(M, q)) + oM.p;
            // R <- R Rx(alpha) Rz(theta): two column rotations (24 operations; the 3x3 product is 45)
            const R ca = M.place_R[I][4], sa = M.place_R[I][7], c = cs[I], s = sn[I];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const R a = oM.R_(r, 0), b = oM.R_(r, 1), d = oM.R_(r, 2);
                const R b1 = ca * b + sa * d, d1 = ca * d - sa * b;
                oM.R_(r, 0) = c * a + s * b1; oM.R_(r, 1) = c * b1 - s * a; oM.R_(r, 2) = d1;
            }
        }
    } else if (I == 0) oM = S.li[0];
    else {
        Xf<R> n;
        n.R_ = matmul(oM.R_, S.li[I].R_);
        n.p = mul(oM.R_, S.li[I].p) + oM.p;
        oM = n;
    }
    if ((FLAGS & KIN_MOUNT) && I == 2) K.mount_pos = mul(oM.R_, v3<R>(M.mount_p[0], M.mount_p[1], M.mount_p[2])) + oM.p;
    if ((FLAGS & KIN_LINKS) && I >= 2) K.link_com[I - 2] = mul(oM.R_, v3<R>(M.com[I][0], M.com[I][1], M.com[I][2])) + oM.p;
    if ((FLAGS & KIN_MANIP) && I >= 3) {  // WORLD jacobian column of a revolute joint: linear = p x z
        const Vec3<R> z = v3<R>(oM.R_.m[2], oM.R_.m[5], oM.R_.m[8]);
        const Vec3<R> l = cross(oM.p, z);
        Jl[0 * 7 + (I - 3)] = l.x; Jl[1 * 7 + (I - 3)] = l.y; Jl[2 * 7 + (I - 3)] = l.z;
    }
    if (I == 9) {
        K.ee_pos = mul(oM.R_, v3<R>(M.ee_p[0], M.ee_p[1], M.ee_p[2])) + oM.p;
        if (FLAGS & KIN_VEL) K.ee_lin_vel = mul(oM.R_, S.v[9].v) + cross(oM.p, mul(oM.R_, S.v[9].w));
    }
    if (I < STOP) world_chain<R, (I < STOP ? I + 1 : I), FLAGS, PLANE, STOP>(M, S, oM, K, Jl, q, cs, sn);
}

// One PinocchioDynamics::calculate(): accelerations + the kinematics the objective will read.
//   u      : generalised forces commanded by the control, tau = [0,0,0,u3..u9,0,0] (pinocchio_dynamics.cpp:238-239)
//   FUSED  : a = M^-1 u
//   NLE    : also produce nle(q, qd) (needed for tau^T v of the energy tank, or in FAITHFUL mode)
//   PLANE  : cs / sn hold the joint sines / cosines and joints 0..9 are crossed by their structure (see act_inv_joint);
//            only with DO_ABA = false (the generic solver reads the transform matrices PLANE does not form)
template <class R, bool FAITHFUL, bool NLE, int FLAGS, bool DO_ABA = true, bool PLANE = false>
MPPI_HD void robot_calculate(const RobotModel<R> &M, const R *q, const R *qd, const R *u, R *qdd, R *nle, Kinematics<R> &K, const R *cs = nullptr, const R *sn = nullptr) {
    static_assert(!(PLANE && DO_ABA), "PLANE forms no transform matrices for the generic solver");
    Scratch<R> S;
    Mot<R> agf[NJ];
    constexpr bool NEED_NLE = NLE || FAITHFUL;
    constexpr bool VEL = NEED_NLE || (FLAGS & KIN_VEL);
    pass1<R, 0, VEL, NEED_NLE, FAITHFUL, PLANE>(M, q, qd, S, agf, cs, sn);
    {
        Xf<R> oM;
        R Jl[21];
        world_chain<R, 0, FLAGS, PLANE>(M, S, oM, K, Jl, q, cs, sn);
        if (FLAGS & KIN_MANIP) {
            R g[6];  // J J^T, symmetric: 00 01 02 11 12 22
            int n = 0;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = r; c < 3; c++) {
                    R s = R(0);
#pragma unroll
                    for (int j = 0; j < 7; j++) s += Jl[r * 7 + j] * Jl[c * 7 + j];
                    g[n++] = s;
                }
            // determinant laid out like Eigen's 3x3 cofactor expansion along the first row
            K.manip_det = g[0] * (g[3] * g[5] - g[4] * g[4]) - g[1] * (g[1] * g[5] - g[4] * g[2]) + g[2] * (g[1] * g[4] - g[3] * g[2]);
        }
    }
    R tau[NJ];
    if (NEED_NLE) rnea_back<R, NJ - 1, PLANE>(M, S, nle, q, cs, sn);
    if (!DO_ABA) return;  // the caller runs the structure-exploiting solver of robot_fast.cuh instead
#pragma unroll
    for (int i = 0; i < NJ; i++) tau[i] = FAITHFUL ? (u[i] + nle[i]) : u[i];
    Art<R> cur;
    aba_back_all<R, NJ - 1, FAITHFUL>(M, tau, S, cur);
    Mot<R> a[NJ];
    aba_fwd<R, 0, FAITHFUL>(M, S, a, qdd);
}

// ---- ROLLED: the seven arm joints as ONE loop body ------------------------------------------------------------------------
// The kernels that evaluate kinematics and the RNEA every step (assisted manipulation, full reach-to-pose) are bound by
// instruction fetch, not by issue: unrolled, these passes are ~2200 straight-line instructions per step, and a step body
// beyond the 32 KB instruction-cache level is streamed from L2 at ~5 bytes per cycle and SM. Joints
// 3..9 share one structure — revolute about z behind a fixed rotation about x and a fixed offset — so each pass runs them
// as a loop over the joint index, per-joint constants read from the model by index, per-joint results in thread-local
// arrays; the base joints (0..2) and the fingers (10, 11) keep their specialised code. The loop multiplies through the
// offsets' structural zeros that the unrolled build drops (0 * x + c = c for finite x): same values, ~30 % more executed
// operations, a third of the code. FUSED / PLANE only (cs, sn = the step's joint cosines / sines).
template <class R> MPPI_HD Vec3<R> arm_to_joint(R ca, R sa, R c, R s, const Vec3<R> &v) { return rotz_t(c, s, rotx_t(ca, sa, v)); }   // E^T v
template <class R> MPPI_HD Vec3<R> arm_to_parent(R ca, R sa, R c, R s, const Vec3<R> &v) { return rotx(ca, sa, rotz(c, s, v)); }     // E v

template <class R, bool NLE, int FLAGS>
MPPI_HD void robot_calculate_rolled(const RobotModel<R> &M, const R *q, const R *qd, R *nle, Kinematics<R> &K, const R *cs, const R *sn) {
    Scratch<R> S;
    Mot<R> agf[NJ];
    constexpr bool VEL = NLE || (FLAGS & KIN_VEL);
    pass1<R, 0, VEL, NLE, false, true, 2>(M, q, qd, S, agf, cs, sn);
    if (VEL) {
#pragma unroll 1
        for (int i = 3; i <= 9; i++) {
            const R ca = M.place_R[i][4], sa = M.place_R[i][7], c = cs[i], s = sn[i];
            const Vec3<R> off = v3<R>(M.place_p[i][0], M.place_p[i][1], M.place_p[i][2]);
            const R w = qd[i] * M.sign[i];
            const Mot<R> vpar = S.v[i - 1];
            Mot<R> v;
            v.v = arm_to_joint(ca, sa, c, s, cross_sub(off, vpar.w, vpar.v));
            v.w = arm_to_joint(ca, sa, c, s, vpar.w);
            v.w.z += w;                                     // + S qd
            S.v[i] = v;
            if (NLE) {
                const Mot<R> cv = cross_joint<R, JT_RZ>(v, w);
                const Frc<R> h = body_mul(M, i, v);
                const Frc<R> vxh = fcross(v, h);
                const Mot<R> apar = agf[i - 1];
                Mot<R> a;
                a.v = arm_to_joint(ca, sa, c, s, cross_sub(off, apar.w, apar.v)) + cv.v;
                a.w = arm_to_joint(ca, sa, c, s, apar.w) + cv.w;
                agf[i] = a;
                const Frc<R> ya = body_mul(M, i, a);
                S.f[i].f = ya.f + vxh.f; S.f[i].n = ya.n + vxh.n;
            }
        }
        pass1<R, 10, VEL, NLE, false, true, 11>(M, q, qd, S, agf, cs, sn);
    }
    {
        Xf<R> oM;
        R Jl[21];
        world_chain<R, 0, FLAGS, true, 2>(M, S, oM, K, Jl, q, cs, sn);
#pragma unroll 1
        for (int i = 3; i <= 9; i++) {
            const R ca = M.place_R[i][4], sa = M.place_R[i][7], c = cs[i], s = sn[i];
            oM.p = mul(oM.R_, v3<R>(M.place_p[i][0], M.place_p[i][1], M.place_p[i][2])) + oM.p;
#pragma unroll
            for (int r = 0; r < 3; r++) {   // R <- R Rx(alpha) Rz(theta): two column rotations
                const R a = oM.R_(r, 0), b = oM.R_(r, 1), d = oM.R_(r, 2);
                const R b1 = ca * b + sa * d, d1 = ca * d - sa * b;
                oM.R_(r, 0) = c * a + s * b1; oM.R_(r, 1) = c * b1 - s * a; oM.R_(r, 2) = d1;
            }
            if (FLAGS & KIN_LINKS) K.link_com[i - 2] = mul(oM.R_, v3<R>(M.com[i][0], M.com[i][1], M.com[i][2])) + oM.p;
            if (FLAGS & KIN_MANIP) {   // WORLD jacobian column of a revolute joint: linear = p x z
                const Vec3<R> l = cross(oM.p, v3<R>(oM.R_.m[2], oM.R_.m[5], oM.R_.m[8]));
                Jl[0 * 7 + (i - 3)] = l.x; Jl[1 * 7 + (i - 3)] = l.y; Jl[2 * 7 + (i - 3)] = l.z;
            }
        }
        K.ee_pos = mul(oM.R_, v3<R>(M.ee_p[0], M.ee_p[1], M.ee_p[2])) + oM.p;
        if (FLAGS & KIN_VEL) K.ee_lin_vel = mul(oM.R_, S.v[9].v) + cross(oM.p, mul(oM.R_, S.v[9].w));
        if (FLAGS & KIN_MANIP) {
            R g[6];  // J J^T, symmetric: 00 01 02 11 12 22
            int n = 0;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c2 = r; c2 < 3; c2++) {
                    R sum = R(0);
#pragma unroll
                    for (int j = 0; j < 7; j++) sum += Jl[r * 7 + j] * Jl[c2 * 7 + j];
                    g[n++] = sum;
                }
            K.manip_det = g[0] * (g[3] * g[5] - g[4] * g[4]) - g[1] * (g[1] * g[5] - g[4] * g[2]) + g[2] * (g[1] * g[4] - g[3] * g[2]);
        }
    }
    if (NLE) {
        rnea_back<R, NJ - 1, true, 10>(M, S, nle, q, cs, sn);    // the fingers add their forces to joint 9
#pragma unroll 1
        for (int i = 9; i >= 3; --i) {
            const R ca = M.place_R[i][4], sa = M.place_R[i][7], c = cs[i], s = sn[i];
            const Frc<R> f = S.f[i];
            nle[i] = f.n.z * M.sign[i];
            Frc<R> fp;
            fp.f = arm_to_parent(ca, sa, c, s, f.f);
            fp.n = cross_add(v3<R>(M.place_p[i][0], M.place_p[i][1], M.place_p[i][2]), fp.f, arm_to_parent(ca, sa, c, s, f.n));
            S.f[i - 1].f = S.f[i - 1].f + fp.f;
            S.f[i - 1].n = S.f[i - 1].n + fp.n;
        }
        rnea_back<R, 2, true, 0>(M, S, nle, q, cs, sn);
    }
}

}  // namespace mppi_b200
