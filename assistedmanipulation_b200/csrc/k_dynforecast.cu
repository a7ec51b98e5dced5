// FrankaRidgeback::DynamicsForecast::forecast on the device, batched: one thread per controller
// rolls the Pinocchio-backend dynamics forward under zero control for `steps` steps and records, BEFORE each
// step, what the reference records (src/frankaridgeback/dynamics.cpp:104-138): joint positions, the whole
// EndEffectorState of the last calculate() (pinocchio_dynamics.cpp:172-223), the powers (0 in this backend,
// pinocchio_dynamics.hpp:211-223), the tank energy and the forecast wrench. The dynamics are the FAITHFUL
// calculate() of robot.cuh (tau = tau_in + nle, full ABA); the end-effector state comes from a second-order
// forward pass along the chain to joint 9 written here (it is not on the rollout path: one thread, K = 1).
// apply_wrench adds tau += J_ee^T w (the line the reference leaves commented out, pinocchio_dynamics.cpp:240).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/mppi_b200.h"
#include "model_init.h"

namespace mppi_b200 {

constexpr int DF_RECORD = MPPI_B200_DYNAMICS_FORECAST_RECORD;

struct EeFrame { double R[9], p[3]; };
__constant__ RobotModel<double> c_model_df;
__constant__ EeFrame c_ee_df;

struct EeFull {
    Vec3<double> p, v, w, a, al;
    Mat3<double> R;
    double J[6][NJ];
};

__device__ inline void joint_local(const RobotModel<double> &M, int i, double q, Xf<double> &li) {
    const int type = (i == 0) ? JT_PX : ((i == 1) ? JT_PY : JT_RZ);   // chain 0..9 only
    Mat3<double> P;
    for (int k = 0; k < 9; k++) P.m[k] = M.place_R[i][k];
    const Vec3<double> pp = v3<double>(M.place_p[i][0], M.place_p[i][1], M.place_p[i][2]);
    if (type == JT_RZ) {
        double s, c;
        sincos(q, &s, &c);
        Mat3<double> Rz;
        Rz.m[0] = c; Rz.m[1] = -s; Rz.m[2] = 0; Rz.m[3] = s; Rz.m[4] = c; Rz.m[5] = 0; Rz.m[6] = 0; Rz.m[7] = 0; Rz.m[8] = 1;
        li.R_ = matmul(P, Rz);
        li.p = pp;
    } else {
        li.R_ = P;
        const Vec3<double> t = type == JT_PX ? v3<double>(q, 0.0, 0.0) : v3<double>(0.0, q, 0.0);
        li.p = pp + mul(P, t);
    }
}

// pinocchio::forwardKinematics(q, v, a) + updateFramePlacements + computeFrameJacobian(WORLD) +
// getFrameVelocity / getFrameAcceleration (WORLD) for the end-effector frame (pinocchio_dynamics.cpp:172-223)
__device__ void end_effector_full(const RobotModel<double> &M, const double *q, const double *qd, const double *qdd, EeFull &ee) {
    Xf<double> oM;
    Mot<double> v, a;
    for (int r = 0; r < 6; r++) for (int j = 0; j < NJ; j++) ee.J[r][j] = 0.0;
    for (int i = 0; i <= 9; i++) {
        Xf<double> li;
        joint_local(M, i, q[i] * M.sign[i], li);
        Mot<double> S;   // joint subspace column
        S.v = v3<double>(i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, 0.0);
        S.w = v3<double>(0.0, 0.0, i >= 2 ? 1.0 : 0.0);
        Mot<double> vj; vj.v = S.v * (qd[i] * M.sign[i]); vj.w = S.w * (qd[i] * M.sign[i]);
        Mot<double> aj; aj.v = S.v * (qdd[i] * M.sign[i]); aj.w = S.w * (qdd[i] * M.sign[i]);
        if (i == 0) {
            oM = li; v = vj; a = aj;   // + v x vj = 0
        } else {
            Xf<double> n;
            n.R_ = matmul(oM.R_, li.R_);
            n.p = oM.p + mul(oM.R_, li.p);
            oM = n;
            const Mot<double> vp = act_inv(li, v), ap = act_inv(li, a);
            v.v = vj.v + vp.v; v.w = vj.w + vp.w;
            const Mot<double> cx = mcross(v, vj);
            a.v = aj.v + cx.v + ap.v; a.w = aj.w + cx.w + ap.w;
        }
        const Mot<double> col = act(oM, S);
        ee.J[0][i] = col.v.x; ee.J[1][i] = col.v.y; ee.J[2][i] = col.v.z;
        ee.J[3][i] = col.w.x; ee.J[4][i] = col.w.y; ee.J[5][i] = col.w.z;
    }
    // base block relative to the arm (pinocchio_dynamics.cpp:194-200)
    double sy, cy;
    sincos(q[2], &sy, &cy);
    ee.J[0][0] = cy; ee.J[0][1] = -sy; ee.J[0][2] = 0.0;
    ee.J[1][0] = sy; ee.J[1][1] = cy;  ee.J[1][2] = 0.0;
    ee.J[2][0] = 0.0; ee.J[2][1] = 0.0; ee.J[2][2] = 1.0;
    const Mot<double> sv = act(oM, v), sa = act(oM, a);
    Mat3<double> F;
    for (int k = 0; k < 9; k++) F.m[k] = c_ee_df.R[k];
    ee.R = matmul(oM.R_, F);
    ee.p = oM.p + mul(oM.R_, v3<double>(c_ee_df.p[0], c_ee_df.p[1], c_ee_df.p[2]));
    ee.v = sv.v; ee.w = sv.w; ee.a = sa.v; ee.al = sa.w;
}

// Eigen::Quaterniond(Matrix3d), coefficient order x y z w
__device__ void rotation_to_quaternion(const Mat3<double> &m, double *xyzw) {
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        xyzw[3] = 0.5 * t;
        t = 0.5 / t;
        xyzw[0] = (m(2, 1) - m(1, 2)) * t;
        xyzw[1] = (m(0, 2) - m(2, 0)) * t;
        xyzw[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        xyzw[i] = 0.5 * t;
        t = 0.5 / t;
        xyzw[3] = (m(k, j) - m(j, k)) * t;
        xyzw[j] = (m(j, i) + m(i, j)) * t;
        xyzw[k] = (m(k, i) + m(i, k)) * t;
    }
}

// PinocchioDynamics::calculate(): tau += nle; a = aba(q, v, tau); kinematics
__device__ void calculate(const double *q, const double *v, double *tau, double *acc, EeFull &ee) {
    double nle[NJ];
    Kinematics<double> K;
    robot_calculate<double, true, true, 0, true>(c_model_df, q, v, tau, acc, nle, K);
    for (int i = 0; i < NJ; i++) tau[i] += nle[i];
    end_effector_full(c_model_df, q, v, acc, ee);
}

__global__ void __launch_bounds__(32) k_dynamics_forecast(int batch, int steps, double dt, int apply_wrench, const double *states /* batch x 31 */,
                                                          const double *wrench /* batch x steps x 6 or null */, double *tau_keep /* batch x 12 */,
                                                          double *record /* batch x steps x DF_RECORD */) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= batch) return;
    double q[NJ], v[NJ], tau[NJ], acc[NJ];
    const double *x = states + (size_t)c * 31;
    for (int i = 0; i < NJ; i++) { q[i] = x[i]; v[i] = x[NJ + i]; tau[i] = tau_keep[c * NJ + i]; }
    double energy = x[30];
    EeFull ee;
    calculate(q, v, tau, acc, ee);   // set_state (pinocchio_dynamics.cpp:142-151): the torque of the last step is still there
    for (int step = 0; step < steps; step++) {
        double *r = record + ((size_t)c * steps + step) * DF_RECORD;
        for (int i = 0; i < NJ; i++) r[i] = q[i];
        r[12] = ee.p.x; r[13] = ee.p.y; r[14] = ee.p.z;
        rotation_to_quaternion(ee.R, r + 15);
        r[19] = ee.v.x; r[20] = ee.v.y; r[21] = ee.v.z; r[22] = ee.w.x; r[23] = ee.w.y; r[24] = ee.w.z;
        r[25] = ee.a.x; r[26] = ee.a.y; r[27] = ee.a.z; r[28] = ee.al.x; r[29] = ee.al.y; r[30] = ee.al.z;
        r[31] = 0.0; r[32] = 0.0;
        r[33] = energy;
        double w[6];
        for (int k = 0; k < 6; k++) { w[k] = wrench ? wrench[((size_t)c * steps + step) * 6 + k] : 0.0; r[34 + k] = w[k]; }
        for (int a = 0; a < 6; a++) for (int j = 0; j < NJ; j++) r[40 + a * NJ + j] = ee.J[a][j];
        // step(zero control) (pinocchio_dynamics.cpp:226-260)
        v[0] = 0.0; v[1] = 0.0; v[2] = 0.0;   // Rotation2D(yaw) * 0, u[2]
        for (int i = 0; i < NJ; i++) tau[i] = 0.0;
        if (apply_wrench)
            for (int j = 0; j < NJ; j++) { double s = 0.0; for (int a = 0; a < 6; a++) s += ee.J[a][j] * w[a]; tau[j] += s; }
        calculate(q, v, tau, acc, ee);
        for (int i = 0; i < NJ; i++) v[i] += acc[i] * dt;
        for (int i = 0; i < NJ; i++) q[i] += v[i] * dt;
        double power = 0.0;
        for (int i = 0; i < NJ; i++) power += tau[i] * v[i];
        energy = std_max(0.0, energy + power * dt);   // energy.hpp:24-29
    }
    for (int i = 0; i < NJ; i++) tau_keep[c * NJ + i] = tau[i];
}

}  // namespace mppi_b200

using namespace mppi_b200;

struct mppi_b200_dynamics_forecast {
    mppi_b200_dynamics_forecast_config cfg{};
    mppi_b200_forecast *wrench = nullptr;   // not owned
    int steps = 0;
    cudaStream_t stream = nullptr;
    double *d_states = nullptr, *d_tau = nullptr, *d_record = nullptr, *h_states = nullptr;
    double last_forecast = std::numeric_limits<double>::min();   // dynamics.cpp:93
    std::string error;
};

namespace {
thread_local std::string g_df_error;
int dfail(mppi_b200_dynamics_forecast *f, int code, const std::string &why) { if (f) f->error = why; g_df_error = why; return code; }
#define DF_TRY(f, call) do { cudaError_t _c = (call); if (_c != cudaSuccess) return dfail((f), MPPI_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_c)); } while (0)
}  // namespace

extern "C" {

const char *mppi_b200_dynamics_forecast_last_error(const mppi_b200_dynamics_forecast *f) { return f ? f->error.c_str() : g_df_error.c_str(); }

void mppi_b200_dynamics_forecast_destroy(mppi_b200_dynamics_forecast *f) {
    if (!f) return;
    cudaSetDevice(f->cfg.device);
    if (f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
    cudaFree(f->d_states); cudaFree(f->d_tau); cudaFree(f->d_record);
    if (f->h_states) cudaFreeHost(f->h_states);
    delete f;
}

int mppi_b200_dynamics_forecast_create(const mppi_b200_dynamics_forecast_config *c, mppi_b200_forecast *wrench_forecast, mppi_b200_dynamics_forecast **out) {
    if (out) *out = nullptr;
    if (!c || !out) return dfail(nullptr, MPPI_B200_ERR_INVALID, "null argument");
    if (c->batch < 1) return dfail(nullptr, MPPI_B200_ERR_INVALID, "batch");
    const double steps = std::ceil(c->horison / c->time_step);
    if (!(steps > 0)) return dfail(nullptr, MPPI_B200_ERR_INVALID, "time horison is too small for time step");   // dynamics.cpp:71-75
    if (wrench_forecast && mppi_b200_forecast_batch(wrench_forecast) != c->batch) return dfail(nullptr, MPPI_B200_ERR_INVALID, "the wrench forecast holds a different number of forecasters");
    std::string why;
    if (!topology_matches(&why)) return dfail(nullptr, MPPI_B200_ERR_UNSUPPORTED, "robot model topology: " + why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return dfail(nullptr, MPPI_B200_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    auto *f = new mppi_b200_dynamics_forecast();
    f->cfg = *c; f->wrench = wrench_forecast; f->steps = (int)steps;
    auto bail = [&](const char *why_) { mppi_b200_dynamics_forecast_destroy(f); return dfail(nullptr, MPPI_B200_ERR_CUDA, why_); };
    if (cudaSetDevice(c->device) != cudaSuccess) return bail("cudaSetDevice");
    const RobotModel<double> m = make_robot_model<double>();
    EeFrame ee;
    for (int k = 0; k < 9; k++) ee.R[k] = FR_EE_R[k];
    for (int k = 0; k < 3; k++) ee.p[k] = FR_EE_P[k];
    if (cudaMemcpyToSymbol(c_model_df, &m, sizeof m) != cudaSuccess || cudaMemcpyToSymbol(c_ee_df, &ee, sizeof ee) != cudaSuccess) return bail("model upload");
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
    const size_t B = (size_t)c->batch;
    if (cudaMalloc(&f->d_states, B * 31 * 8) != cudaSuccess || cudaMalloc(&f->d_tau, B * NJ * 8) != cudaSuccess ||
        cudaMalloc(&f->d_record, B * f->steps * DF_RECORD * 8) != cudaSuccess || cudaMallocHost(&f->h_states, B * 31 * 8) != cudaSuccess) return bail("allocation failed");
    cudaMemset(f->d_tau, 0, B * NJ * 8);
    cudaMemset(f->d_record, 0, B * f->steps * DF_RECORD * 8);
    *out = f;
    return MPPI_B200_OK;
}

int mppi_b200_dynamics_forecast_steps(const mppi_b200_dynamics_forecast *f) { return f ? f->steps : 0; }

int mppi_b200_dynamics_forecast_run(mppi_b200_dynamics_forecast *f, const double *states, double time) {
    if (!f || !states) return MPPI_B200_ERR_INVALID;
    DF_TRY(f, cudaSetDevice(f->cfg.device));
    const double *table = nullptr;
    if (f->wrench) {
        const int rc = mppi_b200_forecast_table_device(f->wrench, time, f->cfg.time_step, f->steps, &table);   // forecast(time + step * time_step)
        if (rc != MPPI_B200_OK) return dfail(f, rc, mppi_b200_forecast_last_error(f->wrench));
    }
    DF_TRY(f, cudaStreamSynchronize(f->stream));
    std::memcpy(f->h_states, states, sizeof(double) * 31 * f->cfg.batch);
    DF_TRY(f, cudaMemcpyAsync(f->d_states, f->h_states, sizeof(double) * 31 * f->cfg.batch, cudaMemcpyHostToDevice, f->stream));
    k_dynamics_forecast<<<(f->cfg.batch + 31) / 32, 32, 0, f->stream>>>(f->cfg.batch, f->steps, f->cfg.time_step, f->cfg.apply_wrench, f->d_states, table, f->d_tau, f->d_record);
    DF_TRY(f, cudaGetLastError());
    DF_TRY(f, cudaStreamSynchronize(f->stream));   // the wrench table may be overwritten by the producer's next call
    f->last_forecast = time;
    return MPPI_B200_OK;
}

int mppi_b200_dynamics_forecast_read(mppi_b200_dynamics_forecast *f, double *records, size_t bytes) {
    if (!f || !records) return MPPI_B200_ERR_INVALID;
    const size_t need = sizeof(double) * (size_t)f->cfg.batch * f->steps * DF_RECORD;
    if (bytes != need) return dfail(f, MPPI_B200_ERR_INVALID, "read: wrong size");
    DF_TRY(f, cudaSetDevice(f->cfg.device));
    DF_TRY(f, cudaMemcpy(records, f->d_record, need, cudaMemcpyDeviceToHost));
    return MPPI_B200_OK;
}

int mppi_b200_dynamics_forecast_device_records(mppi_b200_dynamics_forecast *f, const double **records) {
    if (!f || !records) return MPPI_B200_ERR_INVALID;
    *records = f->d_record;
    return MPPI_B200_OK;
}

}  // extern "C"
