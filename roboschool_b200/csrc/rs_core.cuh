// rs_core.cuh -- per-env articulated-rigid-body step, fp32, one env per thread (lane = env).
//
// This is the device-side replacement for what the reference does inside
//   World::bullet_step -> b3InitStepSimulationCommand        (physics-bullet.cpp:358-412)
//   World::query_body_position                                (physics-bullet.cpp:440-495)
//   RoboschoolForwardWalker.calc_state / step                 (gym_forward_walker.py:45-133)
// for ONE env.  The kernels in rs_step.cu map 32 consecutive envs onto the 32 lanes of a warp;
// every per-env quantity is either a register, a thread-private scratch array or a
// [field][env] SoA slot, so all lanes execute the same instruction stream on coalesced data.
//
// The functions are __host__ __device__ so that tests/ can run the SAME fp32 code on the CPU
// (tests/host_emul) when no GPU is present; the product path only ever launches them on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/rs_model.h"

#ifdef __CUDACC__
#define RS_HD __host__ __device__ __forceinline__
#define RS_HDN __host__ __device__ __noinline__
#else
#define RS_HD inline
#define RS_HDN
#endif

namespace rs {

constexpr int ML = RS_MAX_LINKS;
constexpr int MG = RS_MAX_GEOMS;
constexpr int MP = RS_MAX_PAIRS;
constexpr int MA = RS_MAX_ACT;
constexpr int MF = RS_MAX_FEET;
constexpr int MPT = RS_MAX_PARTS;
constexpr float RS_FLT_EPS = 1.1920929e-07f;
constexpr int CDAT = 10;  // floats per cached contact point: localA3 localB3 normal3 dist

// ---------------------------------------------------------------------------------------------
// float, GPU-ready image of rs_model_desc (+ derived tables). Lives in __constant__ memory.
// ---------------------------------------------------------------------------------------------
struct DevModel {
    int n_links, n_dof, fixed_base, n_geoms, n_pairs, n_act, n_feet, n_parts;
    int obs_dim, has_floor, n_iter, frame_skip, alive_rule, body_link, max_episode_steps, max_contacts;
    float base_mass, base_inertia[3];
    int parent[ML], jtype[ML], dof_idx[ML], has_limit[ML], dof_link[ML];
    float Sw[ML][3], Sv[ML][3];       // joint motion subspace in link coords (axis ; axis x d)
    float R0[ML][9];                  // parent coords -> link coords at q=0
    float axis[ML][3], e_vec[ML][3], d_vec[ML][3];
    float mass[ML], inertia[ML][3];
    float lim_lo[ML], lim_hi[ML], jdamping[ML], max_vel[ML], armature[ML], stiffness[ML];
    float break_thresh[ML + 1];
    int g_link[MG], g_type[MG], g_floor[MG];
    float g_p0[MG][3], g_p1[MG][3], g_radius[MG], g_fric_floor[MG];
    float g_bound[MG];                // bounding-sphere radius about the segment midpoint (half length + radius)
    float g_rot[MG][9];               // box geoms: box coords -> link coords (identity otherwise)
    unsigned char pair_a[MP], pair_b[MP], pair_ss[MP];
    float pair_thresh[MP], pair_friction[MP];
    float gravity, timestep, lin_damping, ang_damping, max_coord_vel, erp, erp2, split_thresh, max_limit_impulse, use_gyro;
    float c_erp, c_erp2, c_split_thresh;   // contact rows (erp / erp2 / split_thresh: joint-limit rows)
    int act_dof[MA];
    float act_gain[MA];
    int foot_link[MF], part_link[MPT];
    float alive_z, alive_bonus, electricity_cost, stall_torque_cost, joints_at_limit_cost, foot_collision_cost, initial_z;
    double walk_target_x, walk_target_y, dt_env;
    float reset_joint_spread, reset_yaw_spread, reset_base_pos[3], reset_base_quat[4];
    int reset_set_base;
    int flag_task, flag_timeout_steps;
    double flag_halflen, flag_halfwidth;
    int task_kind, task_swingup, task_link, task_link2, n_reset;
    int reset_dof[MA], task_dof[4];
    float reset_center[MA], reset_spread[MA];
    // FlagrunHarder's thrown cube (rs_model_desc.has_cube): a free box next to the robot, body index n_links + 1
    int has_cube;
    float cube_half[3], cube_mass, cube_inertia[3], cube_thresh, cube_bound, cube_fric_floor, cube_start[3];
    float cube_fric_geom[MG];         // combined friction cube x robot geom
};

// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
RS_HD float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
RS_HD void cross3(float* o, const float* a, const float* b) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
RS_HD void mulv(float* o, const float* M, const float* v) {  // o = M v
    float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    float y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    float z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
RS_HD void tmulv(float* o, const float* M, const float* v) {  // o = M^T v
    float x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    float y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    float z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
RS_HD void mul33(float* o, const float* A, const float* B) {
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) o[i] = t[i];
}
RS_HD void quat_to_mat(float* R, float x, float y, float z, float w) {
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
// btPlaneSpace1
RS_HD void plane_space(const float* n, float* p, float* q) {
    if (fabsf(n[2]) > 0.70710678f) {
        float a = n[1] * n[1] + n[2] * n[2];
        float k = 1.0f / sqrtf(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        float a = n[0] * n[0] + n[1] * n[1];
        float k = 1.0f / sqrtf(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}
// spatial motion parent->child: w' = R w ; v' = R v - r x (R w)          (6-vectors: [w;v])
RS_HD void xf_motion(float* o, const float* R, const float* r, const float* in) {
    float w[3], v[3], c[3];
    mulv(w, R, in); mulv(v, R, in + 3); cross3(c, r, w);
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2]; o[3] = v[0] - c[0]; o[4] = v[1] - c[1]; o[5] = v[2] - c[2];
}
// spatial force child->parent: f' = R^T f ; t' = R^T (t + r x f)             (6-vectors: [t;f])
RS_HD void xf_force_inv_add(float* acc, const float* R, const float* r, const float* in) {
    float c[3], t[3], o[3];
    cross3(c, r, in + 3);
    t[0] = in[0] + c[0]; t[1] = in[1] + c[1]; t[2] = in[2] + c[2];
    tmulv(o, R, t); acc[0] += o[0]; acc[1] += o[1]; acc[2] += o[2];
    tmulv(o, R, in + 3); acc[3] += o[0]; acc[4] += o[1]; acc[5] += o[2];
}
RS_HD float dot6(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// articulated inertia as three 3x3 blocks: torque = A wd + B a ; force = B^T wd + Mm a
// storage: A sym (6: xx xy xz yy yz zz), B full (9), Mm sym (6)  -> 21 floats
RS_HD void sym_mulv(float* o, const float* S, const float* v) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}
RS_HD void inertia_mul(float* o, const float* I, const float* a) {  // o[6] = I * a[6]
    float t[3], u[3];
    sym_mulv(t, I, a); mulv(u, I + 6, a + 3);
    o[0] = t[0] + u[0]; o[1] = t[1] + u[1]; o[2] = t[2] + u[2];
    tmulv(t, I + 6, a); sym_mulv(u, I + 15, a + 3);
    o[3] = t[0] + u[0]; o[4] = t[1] + u[1]; o[5] = t[2] + u[2];
}
RS_HD void sym_to_full(float* F, const float* S) {
    F[0] = S[0]; F[1] = S[1]; F[2] = S[2]; F[3] = S[1]; F[4] = S[3]; F[5] = S[4]; F[6] = S[2]; F[7] = S[4]; F[8] = S[5];
}
// Ip += X^T Ic X  for X = motion transform parent->child (R, r)
RS_HD void inertia_to_parent_add(float* Ip, const float* Ic, const float* R, const float* r) {
    float A[9], B[9], M[9];
    sym_to_full(A, Ic); sym_to_full(M, Ic + 15);
#pragma unroll
    for (int i = 0; i < 9; i++) B[i] = Ic[6 + i];
    // shift to the parent origin (still child coords):  B' = B + r~ M ; A' = A - B r~ + r~ B'^T
    float rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    float rM[9], Bp[9], Brx[9], BpT[9], rBpT[9];
    mul33(rM, rx, M);
#pragma unroll
    for (int i = 0; i < 9; i++) Bp[i] = B[i] + rM[i];
    mul33(Brx, B, rx);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) BpT[i * 3 + j] = Bp[j * 3 + i];
    mul33(rBpT, rx, BpT);
    float Ap[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Ap[i] = A[i] - Brx[i] + rBpT[i];
    // rotate into parent coords: X_p = R^T X R
    float T[9], RT[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) RT[i * 3 + j] = R[j * 3 + i];
    float Ar[9], Br[9], Mr[9];
    mul33(T, Ap, R); mul33(Ar, RT, T);
    mul33(T, Bp, R); mul33(Br, RT, T);
    mul33(T, M, R);  mul33(Mr, RT, T);
    Ip[0] += Ar[0]; Ip[1] += 0.5f * (Ar[1] + Ar[3]); Ip[2] += 0.5f * (Ar[2] + Ar[6]);
    Ip[3] += Ar[4]; Ip[4] += 0.5f * (Ar[5] + Ar[7]); Ip[5] += Ar[8];
#pragma unroll
    for (int i = 0; i < 9; i++) Ip[6 + i] += Br[i];
    Ip[15] += Mr[0]; Ip[16] += 0.5f * (Mr[1] + Mr[3]); Ip[17] += 0.5f * (Mr[2] + Mr[6]);
    Ip[18] += Mr[4]; Ip[19] += 0.5f * (Mr[5] + Mr[7]); Ip[20] += Mr[8];
}
// Cholesky of the 6x6 SPD base articulated inertia; L packed lower-triangular row-major (21)
RS_HD void chol6(float* L, const float* I) {
    float a[6][6];
    float A[9], M[9];
    sym_to_full(A, I); sym_to_full(M, I + 15);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            a[i][j] = A[i * 3 + j]; a[i][j + 3] = I[6 + i * 3 + j]; a[i + 3][j] = I[6 + j * 3 + i]; a[i + 3][j + 3] = M[i * 3 + j];
        }
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float s = a[i][j];
#pragma unroll
            for (int p = 0; p < j; p++) s -= L[i * (i + 1) / 2 + p] * L[j * (j + 1) / 2 + p];
            if (i == j) L[k] = sqrtf(s); else L[k] = s / L[j * (j + 1) / 2 + j];
            k++;
        }
}
RS_HD void chol6_solve(float* x, const float* L, const float* b) {  // x = (L L^T)^-1 b
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float s = b[i];
#pragma unroll
        for (int p = 0; p < i; p++) s -= L[i * (i + 1) / 2 + p] * y[p];
        y[i] = s / L[i * (i + 1) / 2 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        float s = y[i];
#pragma unroll
        for (int p = i + 1; p < 6; p++) s -= L[p * (p + 1) / 2 + i] * x[p];
        x[i] = s / L[i * (i + 1) / 2 + i];
    }
}

// counter-based RNG, bit-identical with oracle/rs_oracle.c rs_hash
RS_HD uint32_t rs_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = 0x9E3779B9u ^ a;
    h = (h ^ (h >> 16)) * 0x85EBCA6Bu; h ^= b + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 13)) * 0xC2B2AE35u; h ^= c + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 16)) * 0x85EBCA6Bu; h ^= d + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 13)) * 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
RS_HD int rs_ctz(uint32_t x) {   // index of the lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
RS_HD int rs_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
RS_HD float rs_uniform(uint32_t seed, uint32_t env, uint32_t episode, uint32_t k) {
    return (float)(rs_hash(seed, env, episode, k) >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------
// per-thread working set.  NL/MC are capacities chosen per model family at launch.
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL = false>   // SL ("slim"): the static-topology path keeps the link cache in shared memory
struct Work {
    // generalized state
    float bpos[3], bquat[4], bw[3], bv[3];
    float q[NL], qd[NL], tau[NL];
    // kinematics (valid after kinematics())
    float Rpl[SL ? 1 : NL][9], rvec[SL ? 1 : NL][3];
    float Rwl[NL + 1][9], pos[NL + 1][3];
    // ABA cache for the unit-impulse response passes
    float h[SL ? 1 : NL][6], invD[SL ? 1 : NL];
    float Lb[21];
    float vel[SL ? 1 : NL + 1][6];
    // live contact points, double buffered (old list -> new list every sub-step)
    int ckey[2][MC];             // manifold id (floor: geom index ; self pair: n_geoms + pair index)
    float cdat[2][MC][CDAT];
    int ccount[2];
    int cur;
    int n_rows, n_contacts;      // of the last solve (reported)
    // joint motors (btMultiBodyJointMotor) of this env, or nullptr when the world has none: element (field f, dof d) at
    // motor[(f * n_dof + d) * motor_stride], fields kp, kd, target position, target velocity, impulse bound per sub-step
    const float* motor = nullptr;
    int motor_stride = 0;
    // FlagrunHarder's cube (only touched by kernels instantiated for a cube topology): world-frame state like the base,
    // cR = world -> cube coords (valid after the kinematics of the sub-step); body_xyz / body_vel: calc_state's body_xyz and
    // robot_body.speed(), read by the throw in the epilogue
    float cpos[3], cquat[4], cw[3], cv[3], cR[9];
    float body_xyz[3], body_vel[3];
};

// row workspace: one record per constraint row, strided so that lane-adjacent envs are address-adjacent
struct RowWS {
    float* p;      // base pointer for this thread
    int stride;    // distance (floats) between consecutive elements of one thread's record
    RS_HD float& at(int row, int rec, int k) const { return p[(size_t)(row * rec + k) * stride]; }
};
// record layout: [0,nd) J ; [nd,2nd) unit response M^-1 J^T   (row scalars stay thread-private)
enum { R_NSCALAR = 0 };

// ---------------------------------------------------------------------------------------------
// kinematics: rot_from_parent / cachedRVector / world transforms     (btMultiBody, forwardKinematics)
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
RS_HD void kinematics(const DevModel& M, Work<NL, MC, SL>& W) {
    float Rbw[9];
    quat_to_mat(Rbw, W.bquat[0], W.bquat[1], W.bquat[2], W.bquat[3]);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) W.Rwl[0][i * 3 + j] = Rbw[j * 3 + i];
    W.pos[0][0] = W.bpos[0]; W.pos[0][1] = W.bpos[1]; W.pos[0][2] = W.bpos[2];
    for (int i = 0; i < M.n_links; i++) {
        int p = M.parent[i];
        float R[9];
        if (M.jtype[i] == RS_JOINT_REVOLUTE) {
            // m_cachedRotParentToThis = btQuaternion(axis, -q) * m_zeroRotParentToThis
            float half = -0.5f * W.q[M.dof_idx[i]];
            float s = sinf(half), c = cosf(half), Ra[9];
            quat_to_mat(Ra, M.axis[i][0] * s, M.axis[i][1] * s, M.axis[i][2] * s, c);
            mul33(R, Ra, M.R0[i]);
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = M.R0[i][k];
        }
        float rv[3];
        mulv(rv, R, M.e_vec[i]);
        rv[0] += M.d_vec[i][0]; rv[1] += M.d_vec[i][1]; rv[2] += M.d_vec[i][2];
        if (M.jtype[i] == RS_JOINT_PRISMATIC) {
            float qq = W.q[M.dof_idx[i]];
            rv[0] += qq * M.axis[i][0]; rv[1] += qq * M.axis[i][1]; rv[2] += qq * M.axis[i][2];
        }
#pragma unroll
        for (int k = 0; k < 9; k++) W.Rpl[i][k] = R[k];
        W.rvec[i][0] = rv[0]; W.rvec[i][1] = rv[1]; W.rvec[i][2] = rv[2];
        mul33(W.Rwl[i + 1], R, W.Rwl[p + 1]);
        float rw[3];
        tmulv(rw, W.Rwl[i + 1], rv);
        W.pos[i + 1][0] = W.pos[p + 1][0] + rw[0]; W.pos[i + 1][1] = W.pos[p + 1][1] + rw[1]; W.pos[i + 1][2] = W.pos[p + 1][2] + rw[2];
    }
}

template <int NL, int MC, bool SL>
RS_HD void velocities(const DevModel& M, Work<NL, MC, SL>& W) {
    if (M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 6; k++) W.vel[0][k] = 0;
    } else { mulv(W.vel[0], W.Rwl[0], W.bw); mulv(W.vel[0] + 3, W.Rwl[0], W.bv); }
    for (int i = 0; i < M.n_links; i++) {
        int p = M.parent[i];
        xf_motion(W.vel[i + 1], W.Rpl[i], W.rvec[i], W.vel[p + 1]);
        int d = M.dof_idx[i];
        if (d >= 0) {
            float qd = W.qd[d];
#pragma unroll
            for (int k = 0; k < 3; k++) { W.vel[i + 1][k] += qd * M.Sw[i][k]; W.vel[i + 1][3 + k] += qd * M.Sv[i][k]; }
        }
    }
}

template <int NL, int MC, bool SL>
RS_HD void clamp_vel(const DevModel& M, Work<NL, MC, SL>& W) {
    float c = M.max_coord_vel;
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bw[k] = fminf(fmaxf(W.bw[k], -c), c); W.bv[k] = fminf(fmaxf(W.bv[k], -c), c); }
    }
    for (int d = 0; d < M.n_dof; d++) W.qd[d] = fminf(fmaxf(W.qd[d], -c), c);
}

// ---- FlagrunHarder's cube: a btMultiBody without links (oracle/rs_oracle.c cube_dynamics / cube_integrate) ----
template <int NL, int MC, bool SL>
RS_HD void cube_kinematics(Work<NL, MC, SL>& W) {
    float Rcw[9];
    quat_to_mat(Rcw, W.cquat[0], W.cquat[1], W.cquat[2], W.cquat[3]);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) W.cR[i * 3 + j] = Rcw[j * 3 + i];
}
template <int NL, int MC, bool SL>
RS_HD void cube_clamp(const DevModel& M, Work<NL, MC, SL>& W) {
    const float c = M.max_coord_vel;
#pragma unroll
    for (int k = 0; k < 3; k++) { W.cw[k] = fminf(fmaxf(W.cw[k], -c), c); W.cv[k] = fminf(fmaxf(W.cv[k], -c), c); }
}
// gravity, velocity damping k v (1 + |v|) and the gyroscopic torque (body coordinates), applied with applyDeltaVee(dt)
template <int NL, int MC, bool SL>
RS_HD void cube_dynamics(const DevModel& M, Work<NL, MC, SL>& W, float dt) {
    float wl[3], Iw[3], g1[3], al[3], aw[3];
    mulv(wl, W.cR, W.cw);
    const float wn = sqrtf(dot3(wl, wl)), vn = sqrtf(dot3(W.cv, W.cv));
    const float ka = M.ang_damping * (1.0f + wn), kl = M.lin_damping * (1.0f + vn);
#pragma unroll
    for (int k = 0; k < 3; k++) Iw[k] = M.cube_inertia[k] * wl[k];
    cross3(g1, wl, Iw);
#pragma unroll
    for (int k = 0; k < 3; k++) al[k] = -(Iw[k] * ka + M.use_gyro * g1[k]) / M.cube_inertia[k];
    tmulv(aw, W.cR, al);
#pragma unroll
    for (int k = 0; k < 3; k++) { W.cw[k] += dt * aw[k]; W.cv[k] += dt * (-kl * W.cv[k]); }
    W.cv[2] += dt * -M.gravity;
    cube_clamp(M, W);
}
template <int NL, int MC, bool SL>
RS_HD void cube_integrate(Work<NL, MC, SL>& W, float dt) {
#pragma unroll
    for (int k = 0; k < 3; k++) W.cpos[k] += dt * W.cv[k];
    float ang = sqrtf(dot3(W.cw, W.cw));
    if (ang * dt > 0.78539816f) ang = 0.5f * 1.57079633f / dt;
    float kk;
    if (ang < 0.001f) kk = 0.5f * dt - dt * dt * dt * 0.020833333333f * ang * ang;
    else kk = sinf(0.5f * ang * dt) / ang;
    float ax = W.cw[0] * kk, ay = W.cw[1] * kk, az = W.cw[2] * kk, aw = cosf(ang * dt * 0.5f);
    float bx = W.cquat[0], by = W.cquat[1], bz = W.cquat[2], bwq = W.cquat[3];
    float x = aw * bx + ax * bwq + ay * bz - az * by;
    float y = aw * by - ax * bz + ay * bwq + az * bx;
    float z = aw * bz + ax * by - ay * bx + az * bwq;
    float w = aw * bwq - ax * bx - ay * by - az * bz;
    float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
    W.cquat[0] = x * inv; W.cquat[1] = y * inv; W.cquat[2] = z * inv; W.cquat[3] = w * inv;
}

// ---------------------------------------------------------------------------------------------
// btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof + applyDeltaVeeMultiDof(dt)
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
RS_HD void aba(const DevModel& M, Work<NL, MC, SL>& W, float dt) {
    const int n = M.n_links;
    float IA[NL + 1][21];
    float Z[NL + 1][6], c[NL][6], Y[NL];
    velocities(M, W);
    for (int b = 0; b <= n; b++) {
        float mass = b == 0 ? M.base_mass : M.mass[b - 1];
        const float* I3 = b == 0 ? M.base_inertia : M.inertia[b - 1];
        const float* v = W.vel[b];
        float fw[3] = {0.f, 0.f, -M.gravity * mass}, fl[3];
        mulv(fl, W.Rwl[b], fw);
        float wn = sqrtf(dot3(v, v)), vn = sqrtf(dot3(v + 3, v + 3));
        float Iw[3] = {I3[0] * v[0], I3[1] * v[1], I3[2] * v[2]};
        float ka = M.ang_damping * (1.0f + wn), kl = M.lin_damping * (1.0f + vn);
        float g1[3], g2[3];
        cross3(g1, v, Iw); cross3(g2, v, v + 3);
        float gy = M.use_gyro;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            Z[b][k] = Iw[k] * ka + gy * g1[k];
            Z[b][3 + k] = -fl[k] + mass * v[3 + k] * kl + mass * g2[k];
        }
        if (b == 0 && M.fixed_base) {
#pragma unroll
            for (int k = 0; k < 6; k++) Z[0][k] = 0;
        }
#pragma unroll
        for (int k = 0; k < 21; k++) IA[b][k] = 0;
        IA[b][0] = I3[0]; IA[b][3] = I3[1]; IA[b][5] = I3[2];
        IA[b][15] = mass; IA[b][18] = mass; IA[b][20] = mass;
        if (b > 0) {
            int i = b - 1, d = M.dof_idx[i];
            if (d >= 0) {
                float qd = W.qd[d], jw[3], jv[3], t1[3], t2[3];
#pragma unroll
                for (int k = 0; k < 3; k++) { jw[k] = qd * M.Sw[i][k]; jv[k] = qd * M.Sv[i][k]; }
                cross3(c[i], v, jw);
                cross3(t1, v, jv); cross3(t2, v + 3, jw);
                c[i][3] = t1[0] + t2[0]; c[i][4] = t1[1] + t2[1]; c[i][5] = t1[2] + t2[2];
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) c[i][k] = 0;
            }
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        int p = M.parent[i], d = M.dof_idx[i];
        float Zt[6];
        if (d >= 0) {
            float S[6] = {M.Sw[i][0], M.Sw[i][1], M.Sw[i][2], M.Sv[i][0], M.Sv[i][1], M.Sv[i][2]};
            float h[6];
            inertia_mul(h, IA[i + 1], S);
            float D = dot6(S, h) + M.armature[i];
            float y = W.tau[d] - dot6(S, Z[i + 1]) - dot6(c[i], h);
            float invD = 1.0f / D;
            Y[i] = y; W.invD[i] = invD;
#pragma unroll
            for (int k = 0; k < 6; k++) W.h[i][k] = h[k];
            float Ic[6];
            inertia_mul(Ic, IA[i + 1], c[i]);
            float kk = invD * y;
#pragma unroll
            for (int k = 0; k < 6; k++) Zt[k] = Z[i + 1][k] + Ic[k] + h[k] * kk;
            // Ia = IA - h h^T / D   (in place: IA[i+1] is not needed afterwards)
            float* I = IA[i + 1];
            I[0] -= h[0] * h[0] * invD; I[1] -= h[0] * h[1] * invD; I[2] -= h[0] * h[2] * invD;
            I[3] -= h[1] * h[1] * invD; I[4] -= h[1] * h[2] * invD; I[5] -= h[2] * h[2] * invD;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b2 = 0; b2 < 3; b2++) I[6 + a * 3 + b2] -= h[a] * h[3 + b2] * invD;
            I[15] -= h[3] * h[3] * invD; I[16] -= h[3] * h[4] * invD; I[17] -= h[3] * h[5] * invD;
            I[18] -= h[4] * h[4] * invD; I[19] -= h[4] * h[5] * invD; I[20] -= h[5] * h[5] * invD;
        } else {
            Y[i] = 0; W.invD[i] = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) { W.h[i][k] = 0; Zt[k] = Z[i + 1][k]; }
        }
        inertia_to_parent_add(IA[p + 1], IA[i + 1], W.Rpl[i], W.rvec[i]);
        xf_force_inv_add(Z[p + 1], W.Rpl[i], W.rvec[i], Zt);
    }
    float acc[NL + 1][6];
    if (M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 6; k++) acc[0][k] = 0;
    } else {
        chol6(W.Lb, IA[0]);
        float r[6];
        chol6_solve(r, W.Lb, Z[0]);
#pragma unroll
        for (int k = 0; k < 6; k++) acc[0][k] = -r[k];
    }
    for (int i = 0; i < n; i++) {
        int p = M.parent[i], d = M.dof_idx[i];
        xf_motion(acc[i + 1], W.Rpl[i], W.rvec[i], acc[p + 1]);
        if (d >= 0) {
            float a = W.invD[i] * (Y[i] - dot6(acc[i + 1], W.h[i]));
#pragma unroll
            for (int k = 0; k < 3; k++) { acc[i + 1][k] += c[i][k] + a * M.Sw[i][k]; acc[i + 1][3 + k] += c[i][3 + k] + a * M.Sv[i][k]; }
            W.qd[d] += dt * a;
        }
    }
    if (!M.fixed_base) {
        float wd[3], vd[3], t[3], wxv[3];
        tmulv(wd, W.Rwl[0], acc[0]);
        cross3(wxv, W.vel[0], W.vel[0] + 3);
        t[0] = acc[0][3] + wxv[0]; t[1] = acc[0][4] + wxv[1]; t[2] = acc[0][5] + wxv[2];
        tmulv(vd, W.Rwl[0], t);
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bw[k] += dt * wd[k]; W.bv[k] += dt * vd[k]; }
    }
    clamp_vel(M, W);
}

// ---------------------------------------------------------------------------------------------
// unit-impulse response of a contact/limit direction, fused with the Jacobian fill:
//   btMultiBody::fillContactJacobianMultiDof + calcAccelerationDeltasMultiDof.
// A wrench (t,f) given in the coordinates of body b (about its COM) is walked to the root twice in
// one sweep: `f` unchanged (its projections S.f are the Jacobian entries) and `Z` with the
// articulated-body corrections; the second sweep distributes accelerations.  jtau: optional
// generalized joint force (limit rows).  J and u are ACCUMULATED into the row record.
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
RS_HD float row_response(const DevModel& M, const Work<NL, MC, SL>& W, int body, const float* wrench, int jdof, float jdir,
                         const RowWS& ws, int row, int rec, bool first, const float* vel, float* relvel) {
    const int n = M.n_links, nd = 6 + M.n_dof;
    float F[NL + 1][6], Z[NL + 1][6], Y[NL];
    for (int b = 0; b <= n; b++) {
#pragma unroll
        for (int k = 0; k < 6; k++) { F[b][k] = 0; Z[b][k] = 0; }
    }
    if (body >= 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) { F[body][k] = wrench[k]; Z[body][k] = -wrench[k]; }
    }
    float denom = 0.f;
    for (int i = n - 1; i >= 0; i--) {
        int p = M.parent[i], d = M.dof_idx[i];
        float Zt[6];
#pragma unroll
        for (int k = 0; k < 6; k++) Zt[k] = Z[i + 1][k];
        if (d >= 0) {
            float S[6] = {M.Sw[i][0], M.Sw[i][1], M.Sw[i][2], M.Sv[i][0], M.Sv[i][1], M.Sv[i][2]};
            float j = dot6(S, F[i + 1]) + (d == jdof ? jdir : 0.f);
            float y = (d == jdof ? jdir : 0.f) - dot6(S, Z[i + 1]);
            Y[i] = y;
            float kk = W.invD[i] * y;
#pragma unroll
            for (int k = 0; k < 6; k++) Zt[k] += W.h[i][k] * kk;
            float& J = ws.at(row, rec, 6 + d);
            J = first ? j : J + j;
            *relvel += j * vel[6 + d];
        } else Y[i] = 0;
        xf_force_inv_add(Z[p + 1], W.Rpl[i], W.rvec[i], Zt);
        xf_force_inv_add(F[p + 1], W.Rpl[i], W.rvec[i], F[i + 1]);
    }
    float acc[NL + 1][6];
    if (M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 6; k++) { acc[0][k] = 0; float& J = ws.at(row, rec, k); float& U = ws.at(row, rec, nd + k); if (first) { J = 0; U = 0; } }
    } else {
        // base Jacobian / response live in WORLD coordinates (Bullet's velocity vector convention)
        float jw[3], jv[3], r[6], uw[3], uv[3];
        tmulv(jw, W.Rwl[0], F[0]); tmulv(jv, W.Rwl[0], F[0] + 3);
        chol6_solve(r, W.Lb, Z[0]);
#pragma unroll
        for (int k = 0; k < 6; k++) acc[0][k] = -r[k];
        tmulv(uw, W.Rwl[0], acc[0]); tmulv(uv, W.Rwl[0], acc[0] + 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float& J0 = ws.at(row, rec, k); float& J1 = ws.at(row, rec, 3 + k);
            float& U0 = ws.at(row, rec, nd + k); float& U1 = ws.at(row, rec, nd + 3 + k);
            if (first) { J0 = jw[k]; J1 = jv[k]; U0 = uw[k]; U1 = uv[k]; }
            else { J0 += jw[k]; J1 += jv[k]; U0 += uw[k]; U1 += uv[k]; }
            denom += jw[k] * uw[k] + jv[k] * uv[k];
            *relvel += jw[k] * vel[k] + jv[k] * vel[3 + k];
        }
    }
    for (int i = 0; i < n; i++) {
        int p = M.parent[i], d = M.dof_idx[i];
        xf_motion(acc[i + 1], W.Rpl[i], W.rvec[i], acc[p + 1]);
        if (d >= 0) {
            float a = W.invD[i] * (Y[i] - dot6(acc[i + 1], W.h[i]));
#pragma unroll
            for (int k = 0; k < 3; k++) { acc[i + 1][k] += a * M.Sw[i][k]; acc[i + 1][3 + k] += a * M.Sv[i][k]; }
            float& U = ws.at(row, rec, nd + 6 + d);
            float S[6] = {M.Sw[i][0], M.Sw[i][1], M.Sw[i][2], M.Sv[i][0], M.Sv[i][1], M.Sv[i][2]};
            float j = dot6(S, F[i + 1]) + (d == jdof ? jdir : 0.f);
            U = first ? a : U + a;
            denom += j * a;
        }
    }
    return denom;   // j . u of THIS side (Bullet sums the two sides' denominators separately)
}

// wrench of a unit force along `dir` applied at world point `p`, in body b's coordinates about its COM
template <int NL, int MC, bool SL>
RS_HD void unit_wrench(const Work<NL, MC, SL>& W, int b, const float* p, const float* dir, float* out) {
    float rw[3] = {p[0] - W.pos[b][0], p[1] - W.pos[b][1], p[2] - W.pos[b][2]}, rl[3], nl[3];
    mulv(rl, W.Rwl[b], rw); mulv(nl, W.Rwl[b], dir);
    cross3(out, rl, nl);
    out[3] = nl[0]; out[4] = nl[1]; out[5] = nl[2];
}

// ---------------------------------------------------------------------------------------------
// narrowphase + persistent manifolds (btConvexPlaneCollisionAlgorithm, GJK-equivalent closest
// points for sphere/capsule pairs, btPersistentManifold add/replace/refresh)
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
RS_HD void to_world(const Work<NL, MC, SL>& W, int b, const float* l, float* w) {
    float t[3]; tmulv(t, W.Rwl[b], l);
    w[0] = W.pos[b][0] + t[0]; w[1] = W.pos[b][1] + t[1]; w[2] = W.pos[b][2] + t[2];
}
template <int NL, int MC, bool SL>
RS_HD void to_local(const Work<NL, MC, SL>& W, int b, const float* w, float* l) {
    float t[3] = {w[0] - W.pos[b][0], w[1] - W.pos[b][1], w[2] - W.pos[b][2]};
    mulv(l, W.Rwl[b], t);
}

RS_HD void seg_seg_closest(const float* p1, const float* q1, const float* p2, const float* q2, float* c1, float* c2) {
    float d1[3] = {q1[0] - p1[0], q1[1] - p1[1], q1[2] - p1[2]};
    float d2[3] = {q2[0] - p2[0], q2[1] - p2[1], q2[2] - p2[2]};
    float r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), s, t;
    const float EPS = 1e-12f;
    if (a <= EPS && e <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = fminf(fmaxf(f / e, 0.f), 1.f); }
    else {
        float c = dot3(d1, r);
        if (e <= EPS) { t = 0; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
        else {
            float b = dot3(d1, d2), denom = a * e - b * b;
            if (denom > EPS * a * e) s = fminf(fmaxf((b * f - c * e) / denom, 0.f), 1.f); else s = 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
            else if (t > 1) { t = 1; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
}

// Closest points between a solid box (half extents h, axis aligned in its own frame) and the segment [a,b] (box coords);
// same algorithm as oracle/rs_oracle.c box_seg_closest (GJK / EPA answer for btBoxShape against a sphere-swept segment).
// Returns the signed distance box <-> segment core; n: unit vector box -> segment, pb: point on the box, ps: on the segment.
RS_HD float box_seg_G(const float* h, const float* a, const float* d, float t) {
    float g = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float p = a[i] + t * d[i];
        const float ex = p > h[i] ? p - h[i] : (p < -h[i] ? p + h[i] : 0.f);
        g += ex * d[i];
    }
    return g;
}
RS_HD float box_seg_closest(const float* h, const float* a, const float* b, float* n, float* pb, float* ps) {
    const float d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    float t;
    const float G0 = box_seg_G(h, a, d, 0.f), G1 = box_seg_G(h, a, d, 1.f);
    if (G0 >= 0.f) t = 0.f;
    else if (G1 <= 0.f) t = 1.f;
    else {
        float tl = 0.f, gl = G0, th = 1.f, gh = G1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int sgn = 0; sgn < 2; sgn++) {
                float tc = 0.f;
                if (fabsf(d[i]) > 1e-12f) { const float tt = ((sgn ? -h[i] : h[i]) - a[i]) / d[i]; if (tt > 0.f && tt < 1.f) tc = tt; }
                const float gc = box_seg_G(h, a, d, tc);
                if (gc <= 0.f) { if (tc > tl) { tl = tc; gl = gc; } }
                else if (tc < th) { th = tc; gh = gc; }
            }
        }
        t = gh - gl > 0.f ? tl + (th - tl) * (-gl) / (gh - gl) : tl;
    }
    float df[3], dist2 = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        ps[i] = a[i] + t * d[i];
        pb[i] = ps[i] > h[i] ? h[i] : (ps[i] < -h[i] ? -h[i] : ps[i]);
        df[i] = ps[i] - pb[i]; dist2 += df[i] * df[i];
    }
    if (dist2 > 1e-18f) {
        const float dist = sqrtf(dist2), il = 1.0f / dist;
        n[0] = df[0] * il; n[1] = df[1] * il; n[2] = df[2] * il;
        return dist;
    }
    // core touches or intersects the box: minimum-translation axis over the face normals and the edge directions d x e_i
    const float c[3] = {0.5f * (a[0] + b[0]), 0.5f * (a[1] + b[1]), 0.5f * (a[2] + b[2])}, hd[3] = {0.5f * d[0], 0.5f * d[1], 0.5f * d[2]};
    float best = 1e30f, bn[3] = {0.f, 0.f, 1.f};
    bool best_edge = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float ov = h[i] + fabsf(hd[i]) - fabsf(c[i]);
        if (ov < best) { best = ov; bn[0] = bn[1] = bn[2] = 0.f; bn[i] = c[i] >= 0.f ? 1.f : -1.f; best_edge = false; }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float ei[3] = {0.f, 0.f, 0.f}, u[3];
        ei[i] = 1.f;
        cross3(u, d, ei);
        const float l2 = dot3(u, u);
        if (l2 < 1e-18f) continue;
        const float il = 1.0f / sqrtf(l2);
        u[0] *= il; u[1] *= il; u[2] *= il;
        const float cu = dot3(c, u);
        const float ov = h[0] * fabsf(u[0]) + h[1] * fabsf(u[1]) + h[2] * fabsf(u[2]) - fabsf(cu);
        if (ov < best) { best = ov; const float sg = cu >= 0.f ? 1.f : -1.f; bn[0] = sg * u[0]; bn[1] = sg * u[1]; bn[2] = sg * u[2]; best_edge = true; }
    }
    n[0] = bn[0]; n[1] = bn[1]; n[2] = bn[2];
    if (best_edge) { ps[0] = c[0]; ps[1] = c[1]; ps[2] = c[2]; }
    else {
        const bool useb = dot3(b, bn) < dot3(a, bn);
        ps[0] = useb ? b[0] : a[0]; ps[1] = useb ? b[1] : a[1]; ps[2] = useb ? b[2] : a[2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) pb[i] = ps[i] + bn[i] * best;
    return -best;
}

// one manifold (<=4 points) worth of cached points, processed in registers/scratch
struct Manifold4 {
    int n;
    float d[4][CDAT];
};

RS_HD int sort_cached_points(const Manifold4& mf, const float* np) {
    int maxi = -1; float maxp = np[9];
#pragma unroll
    for (int i = 0; i < 4; i++) if (mf.d[i][9] < maxp) { maxi = i; maxp = mf.d[i][9]; }
    float res[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < 4; w++) {
        if (maxi == w) continue;
        const int i0 = (w == 0) ? 1 : 0;
        const int i1 = (w == 3) ? 2 : 3;
        const int i2 = (w <= 1) ? 2 : 1;
        float a[3] = {np[0] - mf.d[i0][0], np[1] - mf.d[i0][1], np[2] - mf.d[i0][2]};
        float b[3] = {mf.d[i1][0] - mf.d[i2][0], mf.d[i1][1] - mf.d[i2][1], mf.d[i1][2] - mf.d[i2][2]};
        float c[3]; cross3(c, a, b);
        res[w] = dot3(c, c);
    }
    int best = -1; float mv = -1e30f;
#pragma unroll
    for (int i = 0; i < 4; i++) if (fabsf(res[i]) > mv) { mv = fabsf(res[i]); best = i; }
    return best;
}

// world transform of one body, fetched once per manifold (the per-body arrays are thread-private scratch: every
// dynamic-index access is a local-memory load)
struct BodyXf {
    float R[9], p[3];   // R: world -> body coords ; p: COM in world
};
template <int NL, int MC, bool SL>
RS_HD void load_xf(const Work<NL, MC, SL>& W, int b, BodyXf& X) {
#pragma unroll
    for (int k = 0; k < 9; k++) X.R[k] = W.Rwl[b][k];
    X.p[0] = W.pos[b][0]; X.p[1] = W.pos[b][1]; X.p[2] = W.pos[b][2];
}
template <int NL, int MC, bool SL>
RS_HD void load_xf_cube(const Work<NL, MC, SL>& W, BodyXf& X) {
#pragma unroll
    for (int k = 0; k < 9; k++) X.R[k] = W.cR[k];
    X.p[0] = W.cpos[0]; X.p[1] = W.cpos[1]; X.p[2] = W.cpos[2];
}
RS_HD void xf_to_world(const BodyXf& X, const float* l, float* w) {
    float t[3]; tmulv(t, X.R, l);
    w[0] = X.p[0] + t[0]; w[1] = X.p[1] + t[1]; w[2] = X.p[2] + t[2];
}
RS_HD void xf_to_local(const BodyXf& X, const float* w, float* l) {
    float t[3] = {w[0] - X.p[0], w[1] - X.p[1], w[2] - X.p[2]};
    mulv(l, X.R, t);
}

// bB < 0: the other body is the world (floor)
RS_HD void manifold_add(Manifold4& mf, const BodyXf& XA, const BodyXf& XB, int bB, const float* n, const float* pOnB, float depth, float thresh) {
    if (depth > thresh) return;
    float np[CDAT];
    float pA[3] = {pOnB[0] + n[0] * depth, pOnB[1] + n[1] * depth, pOnB[2] + n[2] * depth};
    xf_to_local(XA, pA, np);
    if (bB >= 0) xf_to_local(XB, pOnB, np + 3); else { np[3] = pOnB[0]; np[4] = pOnB[1]; np[5] = pOnB[2]; }
    np[6] = n[0]; np[7] = n[1]; np[8] = n[2]; np[9] = depth;
    float shortest = thresh * thresh; int nearest = -1;
#pragma unroll
    for (int i = 0; i < 4; i++) if (i < mf.n) {
        float d[3] = {mf.d[i][0] - np[0], mf.d[i][1] - np[1], mf.d[i][2] - np[2]};
        float dd = dot3(d, d);
        if (dd < shortest) { shortest = dd; nearest = i; }
    }
    int idx = nearest;
    if (idx < 0) { idx = mf.n; if (idx == 4) idx = sort_cached_points(mf, np); else mf.n++; }
#pragma unroll
    for (int i = 0; i < 4; i++) if (i == idx) {
#pragma unroll
        for (int k = 0; k < CDAT; k++) mf.d[i][k] = np[k];
    }
}

RS_HD void manifold_refresh(Manifold4& mf, const BodyXf& XA, const BodyXf& XB, int bB, float thresh) {
    bool rem[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        rem[i] = false;
        if (i < mf.n) {
            float pA[3], pB[3];
            xf_to_world(XA, mf.d[i], pA);
            if (bB >= 0) xf_to_world(XB, mf.d[i] + 3, pB); else { pB[0] = mf.d[i][3]; pB[1] = mf.d[i][4]; pB[2] = mf.d[i][5]; }
            float df[3] = {pA[0] - pB[0], pA[1] - pB[1], pA[2] - pB[2]};
            float dist = dot3(df, mf.d[i] + 6);
            mf.d[i][9] = dist;
            if (!(dist <= thresh)) rem[i] = true;
            else {
                float pd[3];
#pragma unroll
                for (int k = 0; k < 3; k++) pd[k] = pB[k] - (pA[k] - mf.d[i][6 + k] * dist);
                if (dot3(pd, pd) > thresh * thresh) rem[i] = true;
            }
        }
    }
    // removeContactPoint(i) for i = n-1 .. 0 : swap with the last live point
#pragma unroll
    for (int i = 3; i >= 0; i--) if (i < mf.n && rem[i]) {
        int last = mf.n - 1;
#pragma unroll
        for (int j = 0; j < 4; j++) if (j == last && j != i) {
#pragma unroll
            for (int k = 0; k < CDAT; k++) mf.d[i][k] = mf.d[j][k];
            rem[i] = rem[j];   // (already-visited higher slot: its flag is false by construction)
        }
        mf.n--;
    }
}

// world end points of geom g from its body's transform (sphere: both ends coincide)
RS_HD void geom_ends(const DevModel& M, int g, const BodyXf& X, float* e) {
    xf_to_world(X, M.g_p0[g], e);
    if (M.g_type[g] == RS_GEOM_CAPSULE) xf_to_world(X, M.g_p1[g], e + 3);
    else { e[3] = e[0]; e[4] = e[1]; e[5] = e[2]; }
}

// one manifold (floor geom k < n_geoms, or self pair k - n_geoms): cached points of the old list, narrowphase,
// refresh, append to the new list.  The geom end points are rebuilt from the body transform, which is fetched once
// (a manifold reaches this function only if the broadphase flagged it or it still holds points: a handful per env).
// oi / no: cursors into the old / new list.
// CUBE: manifolds n_geoms + n_pairs (cube-floor) and n_geoms + n_pairs + 1 + g (cube vs robot geom g) exist as well.
template <bool CUBE = false, int NL, int MC, bool SL>
RS_HD void collide_manifold(const DevModel& M, Work<NL, MC, SL>& W, int k, int old, int nw, int nold, int& oi, int& no) {
    Manifold4 mf; mf.n = 0;
    while (oi < nold && W.ckey[old][oi] == k) {
#pragma unroll
        for (int s = 0; s < 4; s++) if (s == mf.n) {
#pragma unroll
            for (int j = 0; j < CDAT; j++) mf.d[s][j] = W.cdat[old][oi][j];
        }
        mf.n++; oi++;
    }
    if (k < M.n_geoms) {
        if (!M.g_floor[k]) return;
        int b = M.g_link[k] + 1;
        float thr = M.break_thresh[b];
        BodyXf XA;
        load_xf(W, b, XA);
        float ge[6];
        geom_ends(M, k, XA, ge);
        // support vertex in -z: lower end sphere, first end wins ties
        const float* c = (M.g_type[k] == RS_GEOM_CAPSULE && ge[5] < ge[2]) ? ge + 3 : ge;
        float bv[3];
        if (M.g_type[k] == RS_GEOM_BOX) {
            // btBoxShape::localGetSupportingVertex(R^T (-n)): per box axis +half if the component is >= 0, else -half
            const float down[3] = {0.f, 0.f, -1.f};
            float dl[3], db[3], sv[3], t[3], vl[3];
            mulv(dl, XA.R, down); tmulv(db, M.g_rot[k], dl);
#pragma unroll
            for (int a = 0; a < 3; a++) sv[a] = db[a] >= 0.f ? M.g_p1[k][a] : -M.g_p1[k][a];
            mulv(t, M.g_rot[k], sv);
            vl[0] = M.g_p0[k][0] + t[0]; vl[1] = M.g_p0[k][1] + t[1]; vl[2] = M.g_p0[k][2] + t[2];
            xf_to_world(XA, vl, bv);
            c = bv;
        }
        float dist = c[2] - M.g_radius[k];
        if (dist < thr) {
            const float nz[3] = {0.f, 0.f, 1.f};
            float pOnB[3] = {c[0], c[1], 0.f};
            manifold_add(mf, XA, XA, -1, nz, pOnB, dist, thr);
        }
        if (mf.n) manifold_refresh(mf, XA, XA, -1, thr);
    } else if (CUBE && k >= M.n_geoms + M.n_pairs) {
        const float thr = M.cube_thresh;
        BodyXf XC;
        load_xf_cube(W, XC);
        if (k == M.n_geoms + M.n_pairs) {
            // cube vs floor: btConvexPlaneCollisionAlgorithm, one support vertex of the btBoxShape per sub-step
            const float down[3] = {0.f, 0.f, -1.f};
            float dl[3], sv[3], vtx[3];
            mulv(dl, XC.R, down);
#pragma unroll
            for (int a = 0; a < 3; a++) sv[a] = dl[a] >= 0.f ? M.cube_half[a] : -M.cube_half[a];
            xf_to_world(XC, sv, vtx);
            const float dist = vtx[2];
            if (dist < thr) {
                const float nz[3] = {0.f, 0.f, 1.f};
                float pOnB[3] = {vtx[0], vtx[1], 0.f};
                manifold_add(mf, XC, XC, -1, nz, pOnB, dist, thr);
            }
            if (mf.n) manifold_refresh(mf, XC, XC, -1, thr);
        } else {
            // robot geom g (body A) vs cube (body B): closest points of the box and the geom's sphere-swept segment
            const int g = k - M.n_geoms - M.n_pairs - 1;
            const int bA = M.g_link[g] + 1;
            BodyXf XA;
            load_xf(W, bA, XA);
            if (M.g_type[g] != RS_GEOM_BOX) {
                float ge[6], al[3], bl[3], n[3], pb[3], ps[3];
                geom_ends(M, g, XA, ge);
                xf_to_local(XC, ge, al); xf_to_local(XC, ge + 3, bl);
                const float dist = box_seg_closest(M.cube_half, al, bl, n, pb, ps) - M.g_radius[g];
                if (dist < thr) {
                    float nw3[3], pOnB[3];
                    tmulv(nw3, XC.R, n);
                    xf_to_world(XC, pb, pOnB);
                    manifold_add(mf, XA, XC, 1, nw3, pOnB, dist, thr);
                }
            }
            if (mf.n) manifold_refresh(mf, XA, XC, 1, thr);
        }
    } else {
        int pk = k - M.n_geoms;
        int ga = M.pair_a[pk], gb = M.pair_b[pk];
        int bA = M.g_link[ga] + 1, bB = M.g_link[gb] + 1;
        float thr = M.pair_thresh[pk];
        BodyXf XA, XB;
        load_xf(W, bA, XA); load_xf(W, bB, XB);
        float ea[6], eb[6];
        geom_ends(M, ga, XA, ea); geom_ends(M, gb, XB, eb);
        float cA[3], cB[3];
        seg_seg_closest(ea, ea + 3, eb, eb + 3, cA, cB);
        float d[3] = {cA[0] - cB[0], cA[1] - cB[1], cA[2] - cB[2]};
        float len = sqrtf(dot3(d, d)), rB = M.g_radius[gb];
        float dist = len - M.g_radius[ga] - rB;
        bool hit = M.pair_ss[pk] ? (dist <= 0.f) : (dist < thr);
        if (hit) {
            float nn[3] = {1.f, 0.f, 0.f};
            if (len > 1e-12f) { float il = 1.0f / len; nn[0] = d[0] * il; nn[1] = d[1] * il; nn[2] = d[2] * il; }
            float pOnB[3] = {cB[0] + nn[0] * rB, cB[1] + nn[1] * rB, cB[2] + nn[2] * rB};
            manifold_add(mf, XA, XB, bB, nn, pOnB, dist, thr);
        }
        if (mf.n) manifold_refresh(mf, XA, XB, bB, thr);
    }
#pragma unroll
    for (int s = 0; s < 4; s++) if (s < mf.n && no < MC && no < M.max_contacts) {
        W.ckey[nw][no] = k;
#pragma unroll
        for (int j = 0; j < CDAT; j++) W.cdat[nw][no][j] = mf.d[s][j];
        no++;
    }
}

template <int NL, int MC, bool SL>
RS_HD void collide(const DevModel& M, Work<NL, MC, SL>& W) {
    const int old = W.cur, nw = 1 - W.cur;
    int oi = 0, no = 0;
    const int nold = W.ccount[old];
    const int nman = M.n_geoms + M.n_pairs;
    // world end points of every geom, once per sub-step (each geom takes part in many pairs)
    float gw[MG][6];
    for (int g = 0; g < M.n_geoms; g++) {
        int b = M.g_link[g] + 1;
        to_world(W, b, M.g_p0[g], gw[g]);
        if (M.g_type[g] == RS_GEOM_CAPSULE) to_world(W, b, M.g_p1[g], gw[g] + 3);
        else { gw[g][3] = gw[g][0]; gw[g][4] = gw[g][1]; gw[g][5] = gw[g][2]; }
    }
    for (int k = 0; k < nman; k++) {
        const bool has_old = (oi < nold && W.ckey[old][oi] == k);
        // cheap rejects first: nothing cached for this manifold and the shapes are out of reach
        if (!has_old) {
            if (k < M.n_geoms) {
                if (!M.g_floor[k]) continue;
                float zmin = fminf(gw[k][2], gw[k][5]) - (M.g_type[k] == RS_GEOM_BOX ? M.g_bound[k] : M.g_radius[k]);
                if (!(zmin < M.break_thresh[M.g_link[k] + 1])) continue;
            } else {
                int pk = k - M.n_geoms, ga = M.pair_a[pk], gb = M.pair_b[pk];
                float dx = 0.5f * (gw[ga][0] + gw[ga][3] - gw[gb][0] - gw[gb][3]);
                float dy = 0.5f * (gw[ga][1] + gw[ga][4] - gw[gb][1] - gw[gb][4]);
                float dz = 0.5f * (gw[ga][2] + gw[ga][5] - gw[gb][2] - gw[gb][5]);
                float reach = M.g_bound[ga] + M.g_bound[gb] + M.pair_thresh[pk];
                if (dx * dx + dy * dy + dz * dz > reach * reach) continue;
            }
        }
        collide_manifold(M, W, k, old, nw, nold, oi, no);
    }
    W.ccount[nw] = no;
    W.cur = nw;
}

// ---------------------------------------------------------------------------------------------
// constraint rows + projected Gauss-Seidel (btMultiBodyConstraintSolver), Bullet's row order:
// per iteration: non-contact rows (alternating direction), normal rows, friction rows.
// ---------------------------------------------------------------------------------------------
// resolveSingleConstraintRowGeneric.  All 2*nd loads of the row are issued back to back (independent
// addresses) before the first use, so the workspace latency is paid once per row, not once per element.
template <int NDMAX>
RS_HD void resolve_row(const RowWS& ws, int row, int rec, int nd, float* dv, float rhs, float invd, float lo, float hi, float& applied) {
    float j[NDMAX], u[NDMAX];
#pragma unroll
    for (int k = 0; k < NDMAX; k++) { j[k] = 0.f; u[k] = 0.f; if (k < nd) { j[k] = ws.at(row, rec, k); u[k] = ws.at(row, rec, nd + k); } }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NDMAX; k++) dot += j[k] * dv[k];
    float dimp = rhs - dot * invd;
    float sum = applied + dimp;
    if (sum < lo) { dimp = lo - applied; applied = lo; }
    else if (sum > hi) { dimp = hi - applied; applied = hi; }
    else applied = sum;
#pragma unroll
    for (int k = 0; k < NDMAX; k++) dv[k] += u[k] * dimp;
}

// `resp(body, wrench, jdof, jdir, row, rec, first) -> denom` fills J / M^-1 J^T of a row (generic or static sweeps)
template <int NL, int MC, bool SL, class Resp>
RS_HD void solve_with(const DevModel& M, Work<NL, MC, SL>& W, float dt, const RowWS& ws, Resp&& resp) {
    constexpr int NDMAX = 6 + NL;
    const int nd = 6 + M.n_dof, rec = 2 * nd;
    // row-index bases of the three groups; worlds with joint motors reserve a second block of NL non-contact rows
    const int NCCAP = W.motor ? 2 * NL : NL;
    const int NC0 = 0, CN0 = NCCAP, CF0 = NCCAP + MC;
    int n_nc = 0, n_cn = 0, n_cf = 0;
    // row scalars stay thread-private; only J and M^-1 J^T (2*nd floats per row) go to the workspace
    float r_rhs[2 * NL + 3 * MC], r_invd[2 * NL + 3 * MC], r_applied[2 * NL + 3 * MC], r_hi[2 * NL], r_lo[2 * NL];
    float vel[NDMAX];
#pragma unroll
    for (int k = 0; k < NDMAX; k++) vel[k] = 0.f;
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) { vel[k] = W.bw[k]; vel[3 + k] = W.bv[k]; }
    }
#pragma unroll
    for (int d = 0; d < NL; d++) if (d < M.n_dof) vel[6 + d] = W.qd[d];
    const float inv_dt = 1.0f / dt;
    // joint limits: rows only when violated (btMultiBodyJointLimitConstraint::createConstraintRows)
    for (int i = 0; i < M.n_links; i++) {
        int d = M.dof_idx[i];
        if (!M.has_limit[i] || d < 0) continue;
        float q = W.q[d];
#pragma unroll
        for (int side = 0; side < 2; side++) {
            float pen = side == 0 ? q - M.lim_lo[i] : M.lim_hi[i] - q;
            if (pen > 0.f) continue;
            float dir = side == 0 ? 1.f : -1.f;
            int row = NC0 + n_nc++;
            float relvel = 0.f;
            float denom = resp(-1, (const float*)nullptr, d, dir, row, rec, true, (const float*)vel, &relvel);
            float invd = denom > RS_FLT_EPS ? 1.0f / denom : 0.f;
            float erp = pen > M.split_thresh ? M.erp : M.erp2;
            r_rhs[row] = pen > M.split_thresh ? (-pen * erp * inv_dt - relvel) * invd : (-relvel) * invd;
            r_invd[row] = invd; r_applied[row] = 0.f; r_hi[row] = M.max_limit_impulse; r_lo[row] = 0.f;
        }
    }
    // joint motors (btMultiBodyJointMotor::createConstraintRows; the motors join the world after the limit constraints, so
    // their rows follow the limit rows): velocity target kp (q* - q) / dt + kd (qd* - qd) relative to the current joint
    // velocity, impulse within -+ bound, no positional term
    if (W.motor) {
        for (int i = 0; i < M.n_links; i++) {
            int d = M.dof_idx[i];
            if (d < 0) continue;
            const float* mp = W.motor + (size_t)d * W.motor_stride;
            const size_t fs = (size_t)M.n_dof * W.motor_stride;
            float imp = mp[4 * fs];
            if (!(imp > 0.f)) continue;
            float kp = mp[0], kd = mp[fs], tp = mp[2 * fs], tv = mp[3 * fs];
            int row = NC0 + n_nc++;
            float relvel = 0.f;
            float denom = resp(-1, (const float*)nullptr, d, 1.f, row, rec, true, (const float*)vel, &relvel);
            float invd = denom > RS_FLT_EPS ? 1.0f / denom : 0.f;
            float desired = kp * (tp - W.q[d]) * inv_dt + W.qd[d] + kd * (tv - W.qd[d]);
            r_rhs[row] = (desired - relvel) * invd; r_invd[row] = invd; r_applied[row] = 0.f; r_hi[row] = imp; r_lo[row] = -imp;
        }
    }
    // contacts in pool order (manifold id, slot)
    const int cur = W.cur, nc = W.ccount[cur];
    for (int ci = 0; ci < nc; ci++) {
        int key = W.ckey[cur][ci];
        const float* cd = W.cdat[cur][ci];
        int bA, bB;
        if (key < M.n_geoms) { bA = M.g_link[key] + 1; bB = -1; }
        else { int pk = key - M.n_geoms; bA = M.g_link[M.pair_a[pk]] + 1; bB = M.g_link[M.pair_b[pk]] + 1; }
        float pA[3], pB[3];
        to_world(W, bA, cd, pA);
        if (bB >= 0) to_world(W, bB, cd + 3, pB);
        float dirs[3][3];
        dirs[0][0] = cd[6]; dirs[0][1] = cd[7]; dirs[0][2] = cd[8];
        plane_space(dirs[0], dirs[1], dirs[2]);
#pragma unroll
        for (int w = 0; w < 3; w++) {
            int row = (w == 0) ? CN0 + n_cn : CF0 + 2 * n_cn + (w - 1);
            float wr[6], relvel = 0.f;
            unit_wrench(W, bA, pA, dirs[w], wr);
            float denom = resp(bA, (const float*)wr, -1, 0.f, row, rec, true, (const float*)vel, &relvel);
            if (bB >= 0) {
                float nneg[3] = {-dirs[w][0], -dirs[w][1], -dirs[w][2]};
                unit_wrench(W, bB, pB, nneg, wr);
                denom += resp(bB, (const float*)wr, -1, 0.f, row, rec, false, (const float*)vel, &relvel);
            }
            float invd = denom > RS_FLT_EPS ? 1.0f / denom : 0.f;
            float rhs;
            if (w == 0) {
                float pen = cd[9];
                float erp = pen > M.c_split_thresh ? M.c_erp : M.c_erp2;
                float posErr = 0.f, velErr = -relvel;
                if (pen > 0.f) velErr -= pen * inv_dt; else posErr = -pen * erp * inv_dt;
                rhs = pen > M.c_split_thresh ? (posErr + velErr) * invd : velErr * invd;
            } else rhs = -relvel * invd;
            r_rhs[row] = rhs; r_invd[row] = invd; r_applied[row] = 0.f;
        }
        n_cn++;
    }
    n_cf = 2 * n_cn;
    W.n_rows = n_nc + n_cn + n_cf; W.n_contacts = n_cn;
    if (W.n_rows == 0) return;
    float dv[NDMAX];
#pragma unroll
    for (int k = 0; k < NDMAX; k++) dv[k] = 0.f;
    for (int it = 0; it < M.n_iter; it++) {
        for (int j = 0; j < n_nc; j++) {
            int row = NC0 + ((it & 1) ? j : n_nc - 1 - j);
            resolve_row<NDMAX>(ws, row, rec, nd, dv, r_rhs[row], r_invd[row], r_lo[row], r_hi[row], r_applied[row]);
        }
        for (int j = 0; j < n_cn; j++) {
            int row = CN0 + j;
            resolve_row<NDMAX>(ws, row, rec, nd, dv, r_rhs[row], r_invd[row], 0.f, 1e10f, r_applied[row]);
        }
        for (int j = 0; j < n_cf; j++) {
            int row = CF0 + j, ci = j >> 1;
            float tot = r_applied[CN0 + ci];
            if (tot > 0.f) {
                int key = W.ckey[cur][ci];
                float fr = key < M.n_geoms ? M.g_fric_floor[key] : M.pair_friction[key - M.n_geoms];
                resolve_row<NDMAX>(ws, row, rec, nd, dv, r_rhs[row], r_invd[row], -fr * tot, fr * tot, r_applied[row]);
            }
        }
    }
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bw[k] += dv[k]; W.bv[k] += dv[3 + k]; }
    }
#pragma unroll
    for (int d = 0; d < NL; d++) if (d < M.n_dof) W.qd[d] += dv[6 + d];
    clamp_vel(M, W);
}

template <int NL, int MC, bool SL>
RS_HD void solve(const DevModel& M, Work<NL, MC, SL>& W, float dt, const RowWS& ws) {
    solve_with(M, W, dt, ws, [&](int body, const float* wr, int jdof, float jdir, int row, int rec, bool first, const float* vel, float* relvel) {
        return row_response(M, W, body, wr, jdof, jdir, ws, row, rec, first, vel, relvel);
    });
}

// ---------------------------------------------------------------------------------------------
// btMultiBody::stepPositionsMultiDof
// ---------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
RS_HD void integrate(const DevModel& M, Work<NL, MC, SL>& W, float dt) {
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) W.bpos[k] += dt * W.bv[k];
        float ang = sqrtf(dot3(W.bw, W.bw));
        if (ang * dt > 0.78539816f) ang = 0.5f * 1.57079633f / dt;
        float kk;
        if (ang < 0.001f) kk = 0.5f * dt - dt * dt * dt * 0.020833333333f * ang * ang;
        else kk = sinf(0.5f * ang * dt) / ang;
        float ax = W.bw[0] * kk, ay = W.bw[1] * kk, az = W.bw[2] * kk, aw = cosf(ang * dt * 0.5f);
        float bx = W.bquat[0], by = W.bquat[1], bz = W.bquat[2], bwq = W.bquat[3];
        float x = aw * bx + ax * bwq + ay * bz - az * by;
        float y = aw * by - ax * bz + ay * bwq + az * bx;
        float z = aw * bz + ax * by - ay * bx + az * bwq;
        float w = aw * bwq - ax * bx - ay * by - az * bz;
        float inv = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
        W.bquat[0] = x * inv; W.bquat[1] = y * inv; W.bquat[2] = z * inv; W.bquat[3] = w * inv;
    }
    for (int d = 0; d < M.n_dof; d++) W.q[d] += dt * W.qd[d];
}

// optional per-phase cycle accounting (profiling build of the step kernel only)
struct PhaseClock {
    long long t[12];  // kinematics, collide, aba, solve, integrate, (epilogue) ; [6..10] inside solve: enumerate, limit tasks, contact tasks, rhs, PGS
};
#if defined(__CUDA_ARCH__)
#define RS_CLOCK() clock64()
#else
#define RS_CLOCK() 0LL
#endif

template <int NL, int MC, bool SL>
RS_HD void substep(const DevModel& M, Work<NL, MC, SL>& W, const RowWS& ws, PhaseClock* pc = nullptr) {
    float dt = M.timestep;
    long long c0 = 0, c1 = 0;
    if (pc) c0 = RS_CLOCK();
    kinematics(M, W);
    if (pc) { c1 = RS_CLOCK(); pc->t[0] += c1 - c0; c0 = c1; }
    collide(M, W);
    if (pc) { c1 = RS_CLOCK(); pc->t[1] += c1 - c0; c0 = c1; }
    aba(M, W, dt);
    if (pc) { c1 = RS_CLOCK(); pc->t[2] += c1 - c0; c0 = c1; }
    solve(M, W, dt, ws);
    if (pc) { c1 = RS_CLOCK(); pc->t[3] += c1 - c0; c0 = c1; }
    integrate(M, W, dt);
    if (pc) { c1 = RS_CLOCK(); pc->t[4] += c1 - c0; }
}

// ---------------------------------------------------------------------------------------------
// epilogue: query_body_position + calc_state + reward/done  (gym_forward_walker.py:45-133)
// ---------------------------------------------------------------------------------------------
RS_HD void mat_to_quat(const float* R, float* q) {  // R: local->world
    float t = R[0] + R[4] + R[8];
    if (t > 0) { float s = sqrtf(t + 1.0f) * 2; q[3] = 0.25f * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { float s = sqrtf(1.0f + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = 0.25f * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { float s = sqrtf(1.0f + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25f * s; q[2] = (R[5] + R[7]) / s; }
    else { float s = sqrtf(1.0f + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25f * s; }
}
template <int NL, int MC, bool SL>
RS_HD void body_pose(const Work<NL, MC, SL>& W, int b, float* pos, float* quat, float* linvel) {
    pos[0] = W.pos[b][0]; pos[1] = W.pos[b][1]; pos[2] = W.pos[b][2];
    if (b == 0) { quat[0] = W.bquat[0]; quat[1] = W.bquat[1]; quat[2] = W.bquat[2]; quat[3] = W.bquat[3]; }
    else {
        float Rlw[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rlw[i * 3 + j] = W.Rwl[b][j * 3 + i];
        mat_to_quat(Rlw, quat);
    }
    tmulv(linvel, W.Rwl[b], W.vel[b] + 3);
}

struct Episode {
    double potential;
    double target_x, target_y;   // walk target (constant 1 km ahead, or the Flagrun flag)
    float initial_z;
    float feet_contact[MF];
    int frame;
    uint32_t episode;
    int flag_timeout;            // Flagrun: env steps until the flag moves
    uint32_t flag_draws;         // Flagrun: RNG draws consumed by flag_reposition in this episode
    // FlagrunHarder (gym_humanoid_flagrun.py:73-169): task of this episode (0 walk, 1 stand up, 2 roll over), frames on the
    // ground, crawl bookkeeping of calc_potential, per-task history of the last <= 10 outcomes (bits 0-9, length in bits 16+)
    int task, on_ground, crawl_valid, done_latch;
    double crawl_start, crawl_ignored;
    uint32_t hist[3];
    float elec_cost;
};

struct StepOut {
    float reward;
    int done;
    float rewards5[5];
};

// joint_current_relative_position (physics-bullet.cpp:547-566)
RS_HD void rel_joint(const DevModel& M, int li, float q, float qd, float* pos, float* spd) {
    float l1 = M.lim_lo[li], l2 = M.lim_hi[li];
    if (l1 < l2) { float mid = 0.5f * (l1 + l2); q = 2.f * (q - mid) / (l2 - l1); }
    if (M.max_vel[li] > 0.f) qd /= M.max_vel[li];
    else qd *= (M.jtype[li] == RS_JOINT_REVOLUTE) ? 0.1f : 0.5f;
    *pos = q; *spd = qd;
}

// writes obs[k*ostride]; returns walk_target_dist (double) and pitch
template <int NL, int MC, bool SL>
RS_HD void calc_state_core(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, const float* bp, const float* bq, const float* bv,
                           float* obs, int ostride, double* dist, float* pitch, int* at_limit) {
    if (M.task_kind == RS_TASK_REACHER) {
        // gym_reacher.py calc_state (:38-55): relative joint positions (floats), target joint positions (floats),
        // fingertip - target from the part poses; the potential uses the 3-D distance
        const int d0 = M.task_dof[0], d1 = M.task_dof[1];
        float th, thd, ga, gad;
        rel_joint(M, M.dof_link[d0], W.q[d0], W.qd[d0], &th, &thd);
        rel_joint(M, M.dof_link[d1], W.q[d1], W.qd[d1], &ga, &gad);
        const float* pf = W.pos[M.task_link + 1];
        const float* pt = W.pos[M.task_link2 + 1];
        const double vx = (double)pf[0] - (double)pt[0], vy = (double)pf[1] - (double)pt[1], vz = (double)pf[2] - (double)pt[2];
        obs[0] = W.q[M.task_dof[2]]; obs[ostride] = W.q[M.task_dof[3]];
        obs[2 * ostride] = (float)vx; obs[3 * ostride] = (float)vy;
        obs[4 * ostride] = (float)cos((double)th); obs[5 * ostride] = (float)sin((double)th); obs[6 * ostride] = thd;
        obs[7 * ostride] = ga; obs[8 * ostride] = gad;
        *dist = sqrt(vx * vx + vy * vy + vz * vz); *pitch = 0.f; *at_limit = 0;
        return;
    }
    if (M.task_kind != RS_TASK_WALKER) {
        // gym_pendulums.py calc_state (:35-47, :97-105): Joint::current_position returns floats; cos/sin in double
        const int ds = M.task_dof[0], d1 = M.task_dof[1];
        int k = 0;
        obs[(k++) * ostride] = W.q[ds]; obs[(k++) * ostride] = W.qd[ds];
        if (M.task_kind == RS_TASK_DOUBLE_PENDULUM) obs[(k++) * ostride] = W.pos[M.task_link + 1][0];
        obs[(k++) * ostride] = (float)cos((double)W.q[d1]); obs[(k++) * ostride] = (float)sin((double)W.q[d1]); obs[(k++) * ostride] = W.qd[d1];
        if (M.task_kind == RS_TASK_DOUBLE_PENDULUM) {
            const int d2 = M.task_dof[2];
            obs[(k++) * ostride] = (float)cos((double)W.q[d2]); obs[(k++) * ostride] = (float)sin((double)W.q[d2]); obs[(k++) * ostride] = W.qd[d2];
        }
        *dist = 0.0; *pitch = 0.f; *at_limit = 0;
        return;
    }
    const int A = M.n_act;
    int nlim = 0;
    for (int k = 0; k < A; k++) {
        int d = M.act_dof[k], li = M.dof_link[d];
        float p, s;
        rel_joint(M, li, W.q[d], W.qd[d], &p, &s);
        if (fabsf(p) > 0.99f) nlim++;
        obs[(8 + 2 * k) * ostride] = fminf(fmaxf(p, -5.f), 5.f);
        obs[(8 + 2 * k + 1) * ostride] = fminf(fmaxf(s, -5.f), 5.f);
    }
    *at_limit = nlim;
    float mx = 0.f, my = 0.f;
    for (int k = 0; k < M.n_parts; k++) { mx += W.pos[M.part_link[k] + 1][0]; my += W.pos[M.part_link[k] + 1][1]; }
    mx /= (float)M.n_parts; my /= (float)M.n_parts;
    W.body_xyz[0] = mx; W.body_xyz[1] = my; W.body_xyz[2] = bp[2];
    W.body_vel[0] = bv[0]; W.body_vel[1] = bv[1]; W.body_vel[2] = bv[2];
    // Pose::rpy (python-binding.cpp:60-73)
    float qx = bq[0], qy = bq[1], qz = bq[2], qw = bq[3];
    float sqw = qw * qw, sqx = qx * qx, sqy = qy * qy, sqz = qz * qz;
    float t2 = -2.0f * (qx * qz - qy * qw) / (sqx + sqy + sqz + sqw);
    float yaw = atan2f(2.0f * (qx * qy + qz * qw), (sqx - sqy - sqz + sqw));
    float roll = atan2f(2.0f * (qy * qz + qx * qw), (-sqx - sqy + sqz + sqw));
    t2 = fminf(fmaxf(t2, -1.0f), 1.0f);
    float p = asinf(t2);
    float z = bp[2];
    if (E.initial_z < 0.f) E.initial_z = z;
    // target bearing / distance: the target is 1 km away, so these run in double (a handful of ops)
    double dy = E.target_y - (double)my, dx = E.target_x - (double)mx;
    *dist = sqrt(dy * dy + dx * dx);
    float ang = (float)(atan2(dy, dx) - (double)yaw);
    float cy = cosf(-yaw), sy = sinf(-yaw);
    float vx = cy * bv[0] - sy * bv[1], vy = sy * bv[0] + cy * bv[1], vz = bv[2];
    float more[8] = {z - E.initial_z, sinf(ang), cosf(ang), 0.3f * vx, 0.3f * vy, 0.3f * vz, roll, p};
#pragma unroll
    for (int k = 0; k < 8; k++) obs[k * ostride] = fminf(fmaxf(more[k], -5.f), 5.f);
    for (int k = 0; k < M.n_feet; k++) obs[(8 + 2 * A + k) * ostride] = E.feet_contact[k];
    *pitch = p;
}

// generic wrapper: world transforms + link velocities by the runtime-topology sweeps
template <int NL, int MC, bool SL>
RS_HD void calc_state(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, float* obs, int ostride, double* dist, float* pitch, int* at_limit) {
    kinematics(M, W); velocities(M, W);
    float bp[3], bq[4], bv[3];
    body_pose(W, M.body_link + 1, bp, bq, bv);
    calc_state_core(M, W, E, bp, bq, bv, obs, ostride, dist, pitch, at_limit);
}

// RoboschoolHumanoidFlagrun.flag_reposition (gym_humanoid_flagrun.py:21-35); draws come from the counter-based RNG
// in the reference's order: [randint(2) if first,] uniform x [, uniform y]
RS_HD void flag_reposition(const DevModel& M, Episode& E, uint32_t seed, uint32_t env_id, bool first) {
    auto draw = [&]() { return (double)rs_uniform(seed, env_id, E.episode, 2000u + E.flag_draws++); };
    bool straight = false;
    if (first) straight = draw() < 0.5;     // np_random.randint(2) == 0
    if (straight) {
        E.target_x = 0.5 * M.flag_halflen + (M.flag_halflen - 0.5 * M.flag_halflen) * draw();
        E.target_y = 0.0;
    } else {
        double x = -M.flag_halflen + (M.flag_halflen - -M.flag_halflen) * draw();
        double y = -M.flag_halfwidth + (M.flag_halfwidth - -M.flag_halfwidth) * draw();
        E.target_x = x * 0.5; E.target_y = y * 0.5;   // more_compact = 0.5
    }
    E.flag_timeout = M.flag_timeout_steps;
}

// ---- FlagrunHarder ----
// potential_leak (gym_humanoid_flagrun.py:141-147); z1 torso, z2 pelvis (world), valid after the kinematics of calc_state
RS_HD double harder_leak(float z1, float z2) {
    double z = 0.7 * (double)z1 + 0.3 * (double)z2;
    z = z < 0.0 ? 0.0 : (z > 0.7 ? 0.7 : z);
    return 0.5 + 1.5 * (z / 0.7);
}
// calc_potential (:149-169): progress towards the flag does not count while crawling (torso below 0.8)
RS_HD double harder_potential(const DevModel& M, Episode& E, double dist, float z1, float z2) {
    double frp = -dist / M.dt_env;
    if ((double)z1 < 0.8) {
        if (!E.crawl_valid) { E.crawl_start = frp - E.crawl_ignored; E.crawl_valid = 1; }
        E.crawl_ignored = frp - E.crawl_start;
        frp = E.crawl_start;
    } else {
        frp -= E.crawl_ignored;
        E.crawl_valid = 0;
    }
    return frp + harder_leak(z1, z2) * 100.0;
}
// alive_bonus (:101-139).  Every 30 frames the cube is thrown at the position the robot will have reached when it arrives:
// five np_random draws from the flag stream (angle, speed, 3 x velocity noise), then Robot::set_pose_and_speed
// (python-binding.cpp:289: pose in doubles, velocity through float, identity orientation, angular velocity zeroed).
// Worlds without the cube still consume the draws.
// CUBE: compile-time, true only in the kernels instantiated for a cube topology (the throw is double-precision trigonometry:
// ~1.5 k instructions that every other step kernel would otherwise carry in its epilogue).
template <bool CUBE, int NL, int MC, bool SL>
RS_HD float harder_alive(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, float z, float z1, float z2, uint32_t seed, uint32_t env_id) {
    if (E.frame % 30 == 0 && E.frame > 100 && E.on_ground == 0) {
        if constexpr (!CUBE) E.flag_draws += 5;
        else {
        double u[5];
#pragma unroll
        for (int k = 0; k < 5; k++) u[k] = (double)rs_uniform(seed, env_id, E.episode, 2000u + E.flag_draws++);
        if (M.has_cube) {
            double target[3] = {(double)W.body_xyz[0], (double)W.body_xyz[1], (double)W.body_xyz[2]}, cp[3], v[3];
            const double angle = -3.14 + (3.14 - -3.14) * u[0], from_dist = 4.0;
            const double attack_speed = 20.0 + (30.0 - 20.0) * u[1];
            const double ttt = from_dist / attack_speed;
#pragma unroll
            for (int k = 0; k < 3; k++) target[k] += (double)W.body_vel[k] * ttt;
            cp[0] = target[0] + from_dist * cos(angle); cp[1] = target[1] + from_dist * sin(angle); cp[2] = target[2] + 1.0;
#pragma unroll
            for (int k = 0; k < 3; k++) v[k] = target[k] - cp[k];
            const double sc = attack_speed / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                W.cpos[k] = (float)cp[k];
                W.cv[k] = (float)(v[k] * sc + (-1.0 + (1.0 - -1.0) * u[2 + k]));
                W.cw[k] = 0.f;
            }
            W.cquat[0] = W.cquat[1] = W.cquat[2] = 0.f; W.cquat[3] = 1.f;
        }
        }
    }
    if (z < 0.8f) {
        if (E.task == 0) E.on_ground = 10000;
        E.on_ground += 1;
        E.elec_cost = -2.0f / 5.0f;
    } else if (E.on_ground > 0) {
        E.on_ground -= 1;
        E.elec_cost = -2.0f / 2.0f;
    } else {
        E.elec_cost = -2.0f / 2.0f;
    }
    return E.on_ground < 170 ? (float)harder_leak(z1, z2) : -1.f;
}
// RepeatUnderlearnedTasks.decide_best_task / task_completed (:47-64)
RS_HD int harder_pick_task(const Episode& E, double u) {
    double prob[3], sum = 0.0, cdf[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint32_t bits = E.hist[i] & 0x3FFu;
        int nz = 0;
#pragma unroll
        for (int b = 0; b < 10; b++) nz += (bits >> b) & 1u;
        const int len = (int)(E.hist[i] >> 16);
        prob[i] = 1.0 - (double)nz / (1.0 + (double)len);
        sum += prob[i];
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) { acc += prob[i] / sum; cdf[i] = acc; }
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) if (cdf[i] / cdf[2] <= u) idx++;
    return idx > 2 ? 2 : idx;
}
RS_HD void harder_task_completed(Episode& E, bool completed) {
    const uint32_t h = E.hist[E.task];
    uint32_t len = h >> 16; if (len < 10u) len++;
    E.hist[E.task] = (((h << 1) | (completed ? 1u : 0u)) & 0x3FFu) | (len << 16);
}

// the potential the reward's progress term differentiates: distance to the target, or FlagrunHarder's / Reacher's variant.
// z1 / z2: world z of the robot body and of link M.task_link (FlagrunHarder: pelvis)
RS_HD double task_potential(const DevModel& M, Episode& E, double dist, float z1, float z2) {
    if (M.task_kind == RS_TASK_REACHER) return -100.0 * dist;
    if (M.flag_task == 2) return harder_potential(M, E, dist, z1, z2);
    return -dist / M.dt_env;
}

// RoboschoolHumanoidFlagrun.calc_state (gym_humanoid_flagrun.py:37-44) around the plain calc_state `cs`.
// Returns true if the flag moved (the caller must then not count a potential jump as progress).
template <int NL, int MC, bool SL, class CS>
RS_HD bool calc_state_flag(const DevModel& M, const Work<NL, MC, SL>& W, Episode& E, uint32_t seed, uint32_t env_id, float* obs, int ostride,
                           double* dist, float* pitch, int* at_limit, CS&& cs) {
    if (!M.flag_task) { cs(obs, ostride, dist, pitch, at_limit); return false; }
    E.flag_timeout -= 1;
    cs(obs, ostride, dist, pitch, at_limit);
    if (*dist < 1.0 || E.flag_timeout <= 0) {
        flag_reposition(M, E, seed, env_id, false);
        cs(obs, ostride, dist, pitch, at_limit);   // state again, against the new flag
        E.potential = task_potential(M, E, *dist, W.pos[M.body_link + 1][2], W.pos[(M.task_link >= 0 ? M.task_link : M.body_link) + 1][2]);   // avoid reward jump
        return true;
    }
    return false;
}

// `cs(obs, ostride, &dist, &pitch, &at_limit)` computes the observation (generic or static kinematics)
template <bool CUBE = false, int NL, int MC, bool SL, class CS>
RS_HD void epilogue_with(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, const float* act, int astride, float* obs, int ostride, StepOut& out, CS&& cs,
                         uint32_t seed = 0, uint32_t env_id = 0) {
    double dist; float pitch; int at_limit;
    calc_state_flag(M, W, E, seed, env_id, obs, ostride, &dist, &pitch, &at_limit, cs);
    if (M.task_kind == RS_TASK_REACHER) {
        // gym_reacher.py step (:60-79): potential = -100 |fingertip - target|; never done
        const double pot_old = E.potential;
        E.potential = -100.0 * dist;
        const float a0 = act[0], a1 = act[astride];
        const float thd = obs[6 * ostride], ga = obs[7 * ostride], gad = obs[8 * ostride];
        const float elec = -0.10f * (fabsf(a0 * thd) + fabsf(a1 * gad)) - 0.01f * (fabsf(a0) + fabsf(a1));
        const float stuck = fabsf(fabsf(ga) - 1.f) < 0.01f ? -0.1f : 0.f;
        out.rewards5[0] = (float)(E.potential - pot_old); out.rewards5[1] = elec; out.rewards5[2] = stuck; out.rewards5[3] = out.rewards5[4] = 0.f;
        out.reward = out.rewards5[0] + elec + stuck;
        out.done = 0;
        E.frame++;
        return;
    }
    if (M.task_kind != RS_TASK_WALKER) {
        // gym_pendulums.py step (:49-65, :107-123)
        float reward; int dn;
        out.rewards5[1] = out.rewards5[2] = out.rewards5[3] = out.rewards5[4] = 0.f;
        if (M.task_kind == RS_TASK_PENDULUM) {
            const float th = W.q[M.task_dof[1]];
            if (M.task_swingup) { reward = (float)cos((double)th); dn = 0; }
            else { reward = 1.0f; dn = fabsf(th) > 0.2f; }
            out.rewards5[0] = reward;
        } else {
            const double px = (double)W.pos[M.task_link + 1][0], py = (double)W.pos[M.task_link + 1][2];
            const double pen = 0.01 * px * px + (py + 0.3 - 2.0) * (py + 0.3 - 2.0);
            out.rewards5[0] = 10.f; out.rewards5[1] = (float)(-pen);
            reward = (float)(10.0 - pen); dn = (py + 0.3) <= 1.0;
        }
        for (int k = 0; k < M.obs_dim; k++) if (!isfinite(obs[k * ostride])) dn = 1;
        out.reward = reward; out.done = dn;
        E.frame++;
        return;
    }
    float z = obs[0] + E.initial_z;
    const float z_body = W.pos[M.body_link + 1][2], z_aux = W.pos[(M.task_link >= 0 ? M.task_link : M.body_link) + 1][2];
    float alive;
    if (M.flag_task == 2) alive = harder_alive<CUBE>(M, W, E, z, z_body, z_aux, seed, env_id);
    else if (M.alive_rule == RS_ALIVE_Z_AND_PITCH) alive = (z > M.alive_z && fabsf(pitch) < 1.0f) ? M.alive_bonus : -1.f;
    else if (M.alive_rule == RS_ALIVE_Z) alive = (z > M.alive_z) ? M.alive_bonus : -1.f;
    else if (M.alive_rule == RS_ALIVE_CHEETAH)   // feet_contact still holds the previous step's flags here
        alive = (fabsf(pitch) < 1.0f && E.feet_contact[1] == 0.f && E.feet_contact[2] == 0.f && E.feet_contact[4] == 0.f && E.feet_contact[5] == 0.f) ? M.alive_bonus : -1.f;
    else alive = M.alive_bonus;
    int done = alive < 0.f;
    for (int k = 0; k < M.obs_dim; k++) if (!isfinite(obs[k * ostride])) done = 1;
    double pot_old = E.potential;
    E.potential = task_potential(M, E, dist, z_body, z_aux);
    float progress = (float)(E.potential - pot_old);
    // feet: World::bullet_contact_list(foot) -> floor / other parts (gym_forward_walker.py:106-112)
    float feet_cost = 0.f;
    const int cur = W.cur, nc = W.ccount[cur];
    for (int f = 0; f < M.n_feet; f++) {
        int fb = M.foot_link[f] + 1, g = 0, o = 0;
        for (int ci = 0; ci < nc; ci++) {
            int key = W.ckey[cur][ci];
            if (key < M.n_geoms) { if (M.g_link[key] + 1 == fb) g = 1; }
            else if (key < M.n_geoms + M.n_pairs) { int pk = key - M.n_geoms; if (M.g_link[M.pair_a[pk]] + 1 == fb || M.g_link[M.pair_b[pk]] + 1 == fb) o = 1; }
            else if (key > M.n_geoms + M.n_pairs) { if (M.g_link[key - M.n_geoms - M.n_pairs - 1] + 1 == fb) o = 1; }   // the cube is not the floor
        }
        E.feet_contact[f] = g ? 1.f : 0.f;
        if (o) feet_cost += M.foot_collision_cost;
    }
    const int A = M.n_act;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < A; k++) {
        int d = M.act_dof[k], li = M.dof_link[d];
        float p, s, a = act[k * astride];
        rel_joint(M, li, W.q[d], W.qd[d], &p, &s);
        s1 += fabsf(a * s); s2 += a * a;
    }
    float elec = (M.flag_task == 2 ? E.elec_cost : M.electricity_cost) * (s1 / (float)A) + M.stall_torque_cost * (s2 / (float)A);
    float jl = M.joints_at_limit_cost * (float)at_limit;
    out.rewards5[0] = alive; out.rewards5[1] = progress; out.rewards5[2] = elec; out.rewards5[3] = jl; out.rewards5[4] = feet_cost;
    out.reward = alive + progress + elec + jl + feet_cost;
    out.done = done;
    E.frame++;
    if (M.flag_task == 2) {   // episode_over (gym_forward_walker.py:127-128) -> RepeatUnderlearnedTasks.task_completed
        if ((done && !E.done_latch) || E.frame == M.max_episode_steps) harder_task_completed(E, E.frame == M.max_episode_steps);
        E.done_latch |= done;
    }
}

template <int NL, int MC, bool SL>
RS_HD void epilogue(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, const float* act, int astride, float* obs, int ostride, StepOut& out,
                    uint32_t seed = 0, uint32_t env_id = 0) {
    epilogue_with(M, W, E, act, astride, obs, ostride, out, [&](float* o, int os, double* dist, float* pitch, int* al) {
        calc_state(M, W, E, o, os, dist, pitch, al);
    }, seed, env_id);
}

// reset to the reference's reset distribution (gym_forward_walker.py:24-30, gym_mujoco_walkers.py:106-128)
template <int NL, int MC, bool SL, class CS>
RS_HD void env_reset_with(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, uint32_t seed, uint32_t env_id, float* obs, int ostride, CS&& cs) {
    E.episode++;
    for (int d = 0; d < M.n_dof; d++) { W.q[d] = 0.f; W.qd[d] = 0.f; W.tau[d] = 0.f; }
    for (int k = 0; k < M.n_reset; k++) {
        float u = rs_uniform(seed, env_id, E.episode, (uint32_t)k);
        W.q[M.reset_dof[k]] = (float)((double)M.reset_center[k] + (double)M.reset_spread[k] * (2.0 * (double)u - 1.0));
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { W.bpos[k] = M.reset_base_pos[k]; W.bw[k] = 0.f; W.bv[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; k++) W.bquat[k] = M.reset_base_quat[k];
    E.task = 0;
    if (M.flag_task == 2) E.task = harder_pick_task(E, (double)rs_uniform(seed, env_id, E.episode, 1500u));   // humanoid_task()
    if (M.reset_set_base && M.reset_yaw_spread > 0.f) {
        float u = rs_uniform(seed, env_id, E.episode, 1000u);
        double yaw = (double)M.reset_yaw_spread * (2.0 * (double)u - 1.0);
        // set_initial_orientation (gym_mujoco_walkers.py:106-128) + Pose::set_rpy (python-binding.cpp:42-55), roll = 0
        double pitch = 0.0;
        if (E.task == 1) { pitch = 1.5707963267948966; W.bpos[2] = M.reset_base_pos[2] - 1.4f + 0.45f; }
        else if (E.task == 2) { pitch = 3.141592653589793 * 3 / 2 - 0.15; W.bpos[2] = M.reset_base_pos[2] - 1.4f + 0.22f; }
        const double t0 = cos(yaw * 0.5), t1 = sin(yaw * 0.5), t4 = cos(pitch * 0.5), t5 = sin(pitch * 0.5);
        W.bquat[3] = (float)(t0 * t4); W.bquat[0] = (float)(-t1 * t5); W.bquat[1] = (float)(t0 * t5); W.bquat[2] = (float)(t1 * t4);
    }
    if (M.has_cube) {   // load_urdf(cube.urdf, cpose) of robot_specific_reset: a fresh body at rest (gym_humanoid_flagrun.py:83-86)
#pragma unroll
        for (int k = 0; k < 3; k++) { W.cpos[k] = M.cube_start[k]; W.cw[k] = 0.f; W.cv[k] = 0.f; W.cquat[k] = 0.f; }
        W.cquat[3] = 1.f;
    }
    W.ccount[0] = W.ccount[1] = 0; W.cur = 0;
    for (int k = 0; k < M.n_feet; k++) E.feet_contact[k] = 0.f;
    E.initial_z = M.initial_z;
    E.frame = 0;
    E.target_x = M.walk_target_x; E.target_y = M.walk_target_y;
    E.flag_timeout = 0; E.flag_draws = 0;
    E.on_ground = 0; E.crawl_valid = 0; E.done_latch = 0; E.crawl_start = 0.0; E.crawl_ignored = 0.0; E.elec_cost = M.electricity_cost;
    if (M.flag_task) flag_reposition(M, E, seed, env_id, true);     // humanoid_task() (gym_humanoid_flagrun.py:17-19)
    double dist; float pitch; int al;
    calc_state_flag(M, W, E, seed, env_id, obs, ostride, &dist, &pitch, &al, cs);
    E.potential = task_potential(M, E, dist, W.pos[M.body_link + 1][2], W.pos[(M.task_link >= 0 ? M.task_link : M.body_link) + 1][2]);
}
template <int NL, int MC, bool SL>
RS_HD void env_reset(const DevModel& M, Work<NL, MC, SL>& W, Episode& E, uint32_t seed, uint32_t env_id, float* obs, int ostride) {
    env_reset_with(M, W, E, seed, env_id, obs, ostride, [&](float* o, int os, double* dist, float* pitch, int* al) {
        calc_state(M, W, E, o, os, dist, pitch, al);
    });
}

// apply_action: torque = float(power*coef*clip(a,-1,1))  (gym_forward_walker.py:40-43; set_motor_torque(float))
template <int NL, int MC, bool SL>
RS_HD void apply_action(const DevModel& M, Work<NL, MC, SL>& W, const float* act, int astride) {
    for (int d = 0; d < M.n_dof; d++) W.tau[d] = 0.f;
    for (int k = 0; k < M.n_act; k++) {
        float a = fminf(fmaxf(act[k * astride], -1.f), 1.f);
        W.tau[M.act_dof[k]] = M.act_gain[k] * a;
    }
    for (int l = 0; l < M.n_links; l++) {
        int d = M.dof_idx[l];
        if (d >= 0 && M.jdamping[l] != 0.f) W.tau[d] += -M.jdamping[l] * W.qd[d];
        if (d >= 0 && M.stiffness[l] != 0.f) W.tau[d] += -M.stiffness[l] * W.q[d];
    }
}

}  // namespace rs
