/*
 * rs_oracle.c -- CPU restatement (double precision, plain C) of the reference's
 * World::step hot path.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is
 * imported, linked or executed by the product path (roboschool_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED.  The arithmetic of the path lives in a third-party dependency
 * that is NOT under /root/reference: github.com/olegklimov/bullet3, branch
 * roboschool_self_collision (Bullet 2.87, built -DUSE_DOUBLE_PRECISION=1;
 * install_bullet.sh:8-11,27).  The reference has no golden vectors for it
 * (docker_test_wheel.sh:6 only checks gym.make does not crash).  This file
 * restates Bullet's published algorithm as recalled from upstream bullet3 2.87,
 * anchored on the reference's own call sites:
 *   - step:        World::bullet_step          physics-bullet.cpp:358-412
 *                  (gravity, 5 iterations, contactERP 0.9, dt*skip, skip sub-steps :368-376;
 *                   CONTROL_MODE_TORQUE re-sent each step :379-391; StepSimulation :396-397)
 *   - readback:    World::query_body_position  physics-bullet.cpp:440-495
 *   - rel. joints: Joint::joint_current_relative_position physics-bullet.cpp:547-566
 *   - contacts:    World::bullet_contact_list  physics-bullet.cpp:596-637
 *   - reset:       Joint::reset_current_position :568-577, World::robot_move :583-594
 *   - obs/reward:  RoboschoolForwardWalker.calc_state/step gym_forward_walker.py:45-133,
 *                  alive_bonus gym_mujoco_walkers.py:22-23,69-70,135-136,
 *                  Pose::rpy python-binding.cpp:60-73
 * Each function below names the upstream Bullet routine it restates.
 *
 * What IS pinned (tests/): the obs/reward epilogue against the reference's own
 * Python run in-container over a stub cpp_household (tests/golden/), the MLP
 * against the shipped .weights known answers, and closed-form physics
 * (free fall, pendulum period, energy/momentum conservation).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "../include/rs_model.h"

#define ML RS_MAX_LINKS
#define MAXROWS 256
#define MAXDOF (6 + RS_MAX_LINKS + 6)   /* base + joints + the free cube of FlagrunHarder */
#define MANIFOLD_CACHE_SIZE 4
#define WARMSTART 0 /* "disable warmstarting for btMultiBody, it has issues gaining energy" (btMultiBodyConstraintSolver.cpp, 2.87) */

/* Convention knobs for tools/convention_sweep.py: alternatives of recalled Bullet behaviour that are not
 * (yet) descriptor data.  The defaults are the conventions the CUDA path implements; a knob that wins the sweep is promoted
 * to a descriptor field and implemented on the CUDA side. */
static struct {
    int limit_always;        /* 0: limit rows only when violated ; 1: a row per limited joint side, gap handled through -gap/dt */
    int friction_dirs;       /* 2 (SOLVER_USE_2_FRICTION_DIRECTIONS) or 1 */
    int torque_first_only;   /* 1: joint torques act in the first sub-step only (forces cleared inside the sub-step loop) */
    double warmstart;        /* applied-impulse warm starting factor for contact rows (0: off) */
    int damping_mode;        /* joint damping: 0 explicit torque once per env step (applyJointDamping), 1 explicit per sub-step,
                                2 implicit per sub-step (D += dt*d) */
} K = {0, 2, 0, 0.0, 0};
void oracle_set_knob(const char *name, double v) {
    if (!strcmp(name, "limit_always")) K.limit_always = (int)v;
    else if (!strcmp(name, "friction_dirs")) K.friction_dirs = (int)v;
    else if (!strcmp(name, "torque_first_only")) K.torque_first_only = (int)v;
    else if (!strcmp(name, "warmstart")) K.warmstart = v;
    else if (!strcmp(name, "damping_mode")) K.damping_mode = (int)v;
    else fprintf(stderr, "oracle_set_knob: unknown knob %s\n", name);
}

typedef struct { double w[3], v[3]; } sv6;   /* motion: w=angular v=linear ; force: w=torque v=force */

typedef struct {
    double localA[3], localB[3];
    double normal[3];      /* world, on B, pointing B -> A */
    double posA[3], posB[3];
    double dist;
    double friction;
    double imp_n, imp_f1, imp_f2; /* applied impulses of the last solve (reported, not warm-started) */
    int lifetime;
} cpoint;

typedef struct {
    int n;
    cpoint p[MANIFOLD_CACHE_SIZE];
} manifold;

typedef struct {
    /* generalized state */
    double base_pos[3], base_quat[4]; /* quat x,y,z,w rotating base coords -> world */
    double base_w[3], base_v[3];      /* world frame */
    double q[ML], qd[ML];
    double tau[ML];                   /* joint torques, persist over the sub-steps of one env step */
    /* FlagrunHarder's cube: free body, world-frame velocities like the base (rs_model_desc.has_cube) */
    double cube_pos[3], cube_quat[4], cube_w[3], cube_v[3];
    double body_xyz[3], body_vel[3];  /* calc_state's body_xyz and robot_body.speed(), read by the cube throw */
    /* btMultiBodyJointMotor per dof (Joint::set_servo_target / set_target_speed, physics-bullet.cpp:510-537):
     * kp, kd, position / velocity targets and the impulse bound per sub-step; bound 0 = motor off (torque control) */
    double mot_kp[ML], mot_kd[ML], mot_pos[ML], mot_vel[ML], mot_imp[ML];
    manifold *man;                    /* [n_geoms floor manifolds][n_pairs pair manifolds] */
    /* episode bookkeeping (gym_forward_walker.py) */
    double potential;
    double initial_z;
    double feet_contact[RS_MAX_FEET];
    int frame;
    int done_latch;
    uint32_t episode;
    /* walk target: the model's constant, or the Flagrun flag (gym_humanoid_flagrun.py:21-45) */
    double target_x, target_y;
    int flag_timeout;
    uint32_t flag_draws;
    /* FlagrunHarder (gym_humanoid_flagrun.py:73-160): task of this episode, frames spent on the ground, crawl bookkeeping
     * of calc_potential, per-task history of the last <=10 episode outcomes (bits 0-9 newest first, length in bits 16+) */
    int task, on_ground;
    int crawl_valid;
    double crawl_start, crawl_ignored;
    uint32_t hist[3];
    double elec_cost;
    /* outputs of the last env step */
    double obs[8 + 2 * RS_MAX_ACT + RS_MAX_FEET];
    double reward, rewards5[5];
    int done;
    int n_rows_last, n_contacts_last;
    /* kinematics cache (valid after kinematics()) */
    double Rpl[ML][9];   /* rot parent -> link coords                     */
    double rvec[ML][3];  /* parent COM -> link COM in link coords          */
    double Rwl[ML + 1][9];  /* rot world -> link coords, [0]=base           */
    double pos[ML + 1][3];  /* world COM position, [0]=base                 */
    sv6 S[ML];           /* joint motion subspace in link coords            */
    /* ABA cache for constraint-response passes */
    sv6 h[ML];
    double invD[ML];
    double Ibase_inv[36];
    sv6 vel[ML + 1];     /* spatial velocity, link coords                   */
} oenv;

struct step_pool;
typedef struct oracle_world {
    rs_model_desc m;
    int n_envs;
    int nman;
    oenv *env;
    struct step_pool *pool;   /* persistent worker threads of oracle_step (created on first multi-threaded call) */
} oracle_world;

/* manifolds: [n_geoms geom-floor][n_pairs self pairs] and, with the cube, [cube-floor][cube-vs-geom g for every geom] */
static inline int nman_of(const rs_model_desc *m) { return m->n_geoms + m->n_pairs + (m->has_cube ? 1 + m->n_geoms : 0); }
static inline int nd_of(const rs_model_desc *m) { return 6 + m->n_dof + (m->has_cube ? 6 : 0); }
#define CUBE_BODY(m) ((m)->n_links + 1)   /* body index of the cube: transform cached in Rwl / pos like a link's */

/* ---------------- small vector helpers ---------------- */
static inline void v3set(double *a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static inline void v3cpy(double *a, const double *b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static inline double v3dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3cross(double *o, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void v3sub(double *o, const double *a, const double *b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void v3add(double *o, const double *a, const double *b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static inline void v3axpy(double *o, double s, const double *a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static inline double v3len(const double *a) { return sqrt(v3dot(a, a)); }
static inline void m3mulv(double *o, const double *M, const double *v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3tmulv(double *o, const double *M, const double *v) {
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3mul(double *o, const double *A, const double *B) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    memcpy(o, t, sizeof t);
}
/* rotation matrix of a unit quaternion (x,y,z,w): v_out = R v_in rotates by the quaternion */
static void quat_to_mat(double *R, const double *q) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
static void quat_mul(double *o, const double *a, const double *b) { /* o = a*b, (x,y,z,w) */
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* btPlaneSpace1 (LinearMath/btVector3.h) */
static void plane_space(const double *n, double *p, double *q) {
    if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
        double a = n[1] * n[1] + n[2] * n[2];
        double k = 1.0 / sqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        double a = n[0] * n[0] + n[1] * n[1];
        double k = 1.0 / sqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

/* ---------------- spatial transforms (btSpatialTransformationMatrix) ---------------- */
/* motion parent -> child: w' = R w ; v' = R v - r x (R w) */
static void xf_motion(sv6 *o, const double *R, const double *r, const sv6 *in) {
    double w[3], v[3], c[3];
    m3mulv(w, R, in->w);
    m3mulv(v, R, in->v);
    v3cross(c, r, w);
    v3sub(v, v, c);
    v3cpy(o->w, w); v3cpy(o->v, v);
}
/* force child -> parent (transformInverse): f' = R^T f ; t' = R^T (t + r x f) */
static void xf_force_inv(sv6 *o, const double *R, const double *r, const sv6 *in) {
    double t[3], c[3];
    v3cross(c, r, in->v);
    v3add(t, in->w, c);
    m3tmulv(o->w, R, t);
    m3tmulv(o->v, R, in->v);
}
static inline double sv_dot(const sv6 *m, const sv6 *f) { return v3dot(m->w, f->w) + v3dot(m->v, f->v); }
/* 6x6 matrix (row-major, order [w;v]) times motion -> force */
static void m6mulv(sv6 *o, const double *I, const sv6 *a) {
    double x[6] = {a->w[0], a->w[1], a->w[2], a->v[0], a->v[1], a->v[2]}, y[6];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += I[i * 6 + j] * x[j];
        y[i] = s;
    }
    v3set(o->w, y[0], y[1], y[2]); v3set(o->v, y[3], y[4], y[5]);
}
/* build the 6x6 motion transform X (parent->child) */
static void build_X(double *X, const double *R, const double *r) {
    /* [w';v'] = [[R,0],[-[r]x R, R]] [w;v] */
    double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0}, rxR[9];
    m3mul(rxR, rx, R);
    memset(X, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[i * 6 + j] = R[i * 3 + j];
            X[(i + 3) * 6 + j] = -rxR[i * 3 + j];
            X[(i + 3) * 6 + j + 3] = R[i * 3 + j];
        }
}
/* Ip += X^T Ic X */
static void inertia_to_parent_add(double *Ip, const double *Ic, const double *X) {
    double T[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += Ic[i * 6 + k] * X[k * 6 + j];
            T[i * 6 + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += X[k * 6 + i] * T[k * 6 + j];
            Ip[i * 6 + j] += s;
        }
}
/* dense symmetric-positive 6x6 inverse by Gauss-Jordan with partial pivoting
 * (btMultiBody::solveImatrix does the same job block-wise) */
static int inv6(double *out, const double *in) {
    double a[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { a[i][j] = in[i * 6 + j]; a[i][j + 6] = (i == j); }
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (fabs(a[p][c]) < 1e-300) return -1;
        if (p != c) for (int j = 0; j < 12; j++) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double inv = 1.0 / a[c][c];
        for (int j = 0; j < 12; j++) a[c][j] *= inv;
        for (int r = 0; r < 6; r++) if (r != c) {
            double f = a[r][c];
            if (f != 0) for (int j = 0; j < 12; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) out[i * 6 + j] = a[i][j + 6];
    return 0;
}

/* ---------------- kinematics (btMultiBody: rot_from_parent / cachedRVector / forwardKinematics) ---------------- */
static void kinematics(const rs_model_desc *m, oenv *e) {
    /* base: Bullet stores m_baseQuat as world->base; we store base->world and transpose */
    double Rbw[9];
    quat_to_mat(Rbw, e->base_quat);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) e->Rwl[0][i * 3 + j] = Rbw[j * 3 + i];
    v3cpy(e->pos[0], e->base_pos);
    for (int i = 0; i < m->n_links; i++) {
        int p = m->parent[i];
        double R0[9];
        quat_to_mat(R0, m->zero_rot[i]); /* parent coords -> link coords at q=0 */
        const double *ax = m->axis[i];
        if (m->jtype[i] == RS_JOINT_REVOLUTE) {
            /* m_cachedRotParentToThis = btQuaternion(axis, -q) * m_zeroRotParentToThis */
            double half = -0.5 * e->q[m->dof_idx[i]];
            double s = sin(half), qa[4] = {ax[0] * s, ax[1] * s, ax[2] * s, cos(half)}, Ra[9];
            quat_to_mat(Ra, qa);
            m3mul(e->Rpl[i], Ra, R0);
            /* m_cachedRVector = quatRotate(cachedRot, eVector) + dVector */
            m3mulv(e->rvec[i], e->Rpl[i], m->e_vec[i]);
            v3add(e->rvec[i], e->rvec[i], m->d_vec[i]);
            v3cpy(e->S[i].w, ax);
            v3cross(e->S[i].v, ax, m->d_vec[i]);   /* axis_bottom = axis x dVector */
        } else if (m->jtype[i] == RS_JOINT_PRISMATIC) {
            memcpy(e->Rpl[i], R0, sizeof R0);
            m3mulv(e->rvec[i], R0, m->e_vec[i]);
            v3add(e->rvec[i], e->rvec[i], m->d_vec[i]);
            v3axpy(e->rvec[i], e->q[m->dof_idx[i]], ax);
            v3set(e->S[i].w, 0, 0, 0);
            v3cpy(e->S[i].v, ax);
        } else {
            memcpy(e->Rpl[i], R0, sizeof R0);
            m3mulv(e->rvec[i], R0, m->e_vec[i]);
            v3add(e->rvec[i], e->rvec[i], m->d_vec[i]);
            v3set(e->S[i].w, 0, 0, 0); v3set(e->S[i].v, 0, 0, 0);
        }
        m3mul(e->Rwl[i + 1], e->Rpl[i], e->Rwl[p + 1]);
        double rw[3];
        m3tmulv(rw, e->Rwl[i + 1], e->rvec[i]);
        v3add(e->pos[i + 1], e->pos[p + 1], rw);
    }
    if (m->has_cube) {
        double Rcw[9];
        quat_to_mat(Rcw, e->cube_quat);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) e->Rwl[CUBE_BODY(m)][i * 3 + j] = Rcw[j * 3 + i];
        v3cpy(e->pos[CUBE_BODY(m)], e->cube_pos);
    }
}

/* spatial velocities of all links in link coords (first sweep of the ABA) */
static void velocities(const rs_model_desc *m, oenv *e) {
    if (m->fixed_base) { memset(&e->vel[0], 0, sizeof(sv6)); }
    else { m3mulv(e->vel[0].w, e->Rwl[0], e->base_w); m3mulv(e->vel[0].v, e->Rwl[0], e->base_v); }
    for (int i = 0; i < m->n_links; i++) {
        int p = m->parent[i];
        xf_motion(&e->vel[i + 1], e->Rpl[i], e->rvec[i], &e->vel[p + 1]);
        if (m->dof_idx[i] >= 0) {
            double qd = e->qd[m->dof_idx[i]];
            v3axpy(e->vel[i + 1].w, qd, e->S[i].w);
            v3axpy(e->vel[i + 1].v, qd, e->S[i].v);
        }
    }
}

/* ---------------- btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof ----------------
 * Gravity enters as an applied force m*g on every link (btMultiBodyDynamicsWorld::solveConstraints),
 * velocity damping k*v*(1+|v|) and the gyroscopic term are part of the zero-acceleration force.
 * Result: accelerations applied to the velocities (applyDeltaVeeMultiDof(output, dt), with the
 * m_maxCoordinateVelocity clamp). */
static void clamp_vel(const rs_model_desc *m, oenv *e) {
    double c = m->max_coord_vel;
    if (!m->fixed_base)
        for (int k = 0; k < 3; k++) {
            if (e->base_w[k] > c) e->base_w[k] = c; if (e->base_w[k] < -c) e->base_w[k] = -c;
            if (e->base_v[k] > c) e->base_v[k] = c; if (e->base_v[k] < -c) e->base_v[k] = -c;
        }
    for (int d = 0; d < m->n_dof; d++) { if (e->qd[d] > c) e->qd[d] = c; if (e->qd[d] < -c) e->qd[d] = -c; }
    if (m->has_cube)
        for (int k = 0; k < 3; k++) {
            if (e->cube_w[k] > c) e->cube_w[k] = c; if (e->cube_w[k] < -c) e->cube_w[k] = -c;
            if (e->cube_v[k] > c) e->cube_v[k] = c; if (e->cube_v[k] < -c) e->cube_v[k] = -c;
        }
}

static void aba(const rs_model_desc *m, oenv *e, double dt) {
    int n = m->n_links;
    static const double zero3[3] = {0, 0, 0};
    double IA[ML + 1][36];
    sv6 Z[ML + 1], c[ML], acc[ML + 1];
    double Y[ML];
    double gw[3] = {0, 0, -m->gravity};
    velocities(m, e);
    for (int b = 0; b <= n; b++) {
        double mass = b == 0 ? m->base_mass : m->mass[b - 1];
        const double *I3 = b == 0 ? m->base_inertia : m->inertia[b - 1];
        const sv6 *v = &e->vel[b];
        /* applied force: gravity (world) rotated into link coords; no applied torque */
        double fl[3], fw[3] = {gw[0] * mass, gw[1] * mass, gw[2] * mass};
        m3mulv(fl, e->Rwl[b], fw);
        double wn = v3len(v->w), vn = v3len(v->v);
        double Iw[3] = {I3[0] * v->w[0], I3[1] * v->w[1], I3[2] * v->w[2]};
        double ka = m->ang_damping * (1.0 + wn), kl = m->lin_damping * (1.0 + vn);
        for (int k = 0; k < 3; k++) {
            Z[b].w[k] = Iw[k] * ka;
            Z[b].v[k] = -fl[k] + mass * v->v[k] * kl;
        }
        if (m->use_gyro != 0.0) { double g[3]; v3cross(g, v->w, Iw); v3add(Z[b].w, Z[b].w, g); }
        { double g[3]; v3cross(g, v->w, v->v); v3axpy(Z[b].v, mass, g); }
        if (b == 0 && m->fixed_base) { memset(&Z[0], 0, sizeof(sv6)); }
        memset(IA[b], 0, sizeof IA[b]);
        IA[b][0] = I3[0]; IA[b][7] = I3[1]; IA[b][14] = I3[2];
        IA[b][21] = IA[b][28] = IA[b][35] = mass;
        if (b > 0) {
            int i = b - 1;
            if (m->dof_idx[i] >= 0) {
                /* spatCoriolisAcc = spatVel x spatJointVel */
                double qd = e->qd[m->dof_idx[i]];
                sv6 jv; for (int k = 0; k < 3; k++) { jv.w[k] = qd * e->S[i].w[k]; jv.v[k] = qd * e->S[i].v[k]; }
                double t1[3], t2[3];
                v3cross(c[i].w, v->w, jv.w);
                v3cross(t1, v->w, jv.v); v3cross(t2, v->v, jv.w);
                v3add(c[i].v, t1, t2);
            } else { v3cpy(c[i].w, zero3); v3cpy(c[i].v, zero3); }
        }
    }
    /* leaves -> base */
    for (int i = n - 1; i >= 0; i--) {
        int p = m->parent[i];
        double X[36];
        build_X(X, e->Rpl[i], e->rvec[i]);
        sv6 Zt;
        if (m->dof_idx[i] >= 0) {
            int d = m->dof_idx[i];
            m6mulv(&e->h[i], IA[i + 1], &e->S[i]);
            double D = sv_dot(&e->S[i], &e->h[i]) + m->armature[i];
            Y[i] = e->tau[d] - sv_dot(&e->S[i], &Z[i + 1]) - sv_dot(&c[i], &e->h[i]);
            if (K.damping_mode >= 1) Y[i] -= m->jdamping[i] * e->qd[d] + m->stiffness[i] * e->q[d];
            if (K.damping_mode == 2) D += dt * m->jdamping[i] + dt * dt * m->stiffness[i];
            e->invD[i] = 1.0 / D;
            double Ia[36], hh[6] = {e->h[i].w[0], e->h[i].w[1], e->h[i].w[2], e->h[i].v[0], e->h[i].v[1], e->h[i].v[2]};
            for (int a = 0; a < 6; a++) for (int b2 = 0; b2 < 6; b2++) Ia[a * 6 + b2] = IA[i + 1][a * 6 + b2] - hh[a] * hh[b2] * e->invD[i];
            inertia_to_parent_add(IA[p + 1], Ia, X);
            sv6 Ic; m6mulv(&Ic, IA[i + 1], &c[i]);
            double k = e->invD[i] * Y[i];
            for (int a = 0; a < 3; a++) { Zt.w[a] = Z[i + 1].w[a] + Ic.w[a] + e->h[i].w[a] * k; Zt.v[a] = Z[i + 1].v[a] + Ic.v[a] + e->h[i].v[a] * k; }
        } else {
            inertia_to_parent_add(IA[p + 1], IA[i + 1], X);
            Zt = Z[i + 1];
            e->invD[i] = 0; Y[i] = 0; memset(&e->h[i], 0, sizeof(sv6));
        }
        sv6 Zp; xf_force_inv(&Zp, e->Rpl[i], e->rvec[i], &Zt);
        v3add(Z[p + 1].w, Z[p + 1].w, Zp.w); v3add(Z[p + 1].v, Z[p + 1].v, Zp.v);
    }
    /* base acceleration */
    if (m->fixed_base) { memset(&acc[0], 0, sizeof(sv6)); memset(e->Ibase_inv, 0, sizeof e->Ibase_inv); }
    else {
        inv6(e->Ibase_inv, IA[0]);
        sv6 r; m6mulv(&r, e->Ibase_inv, &Z[0]);
        for (int k = 0; k < 3; k++) { acc[0].w[k] = -r.w[k]; acc[0].v[k] = -r.v[k]; }
    }
    double qdd[ML];
    for (int i = 0; i < n; i++) {
        int p = m->parent[i];
        xf_motion(&acc[i + 1], e->Rpl[i], e->rvec[i], &acc[p + 1]);
        if (m->dof_idx[i] >= 0) {
            double a = e->invD[i] * (Y[i] - sv_dot(&acc[i + 1], &e->h[i]));
            qdd[m->dof_idx[i]] = a;
            for (int k = 0; k < 3; k++) { acc[i + 1].w[k] += c[i].w[k] + a * e->S[i].w[k]; acc[i + 1].v[k] += c[i].v[k] + a * e->S[i].v[k]; }
        }
    }
    /* apply: base accelerations back in the world frame (+ omega x v for the linear part) */
    if (!m->fixed_base) {
        double wd[3], vd[3], t[3], wxv[3];
        m3tmulv(wd, e->Rwl[0], acc[0].w);
        v3cross(wxv, e->vel[0].w, e->vel[0].v);
        v3add(t, acc[0].v, wxv);
        m3tmulv(vd, e->Rwl[0], t);
        v3axpy(e->base_w, dt, wd); v3axpy(e->base_v, dt, vd);
    }
    for (int d = 0; d < m->n_dof; d++) e->qd[d] += dt * qdd[d];
    clamp_vel(m, e);
}

/* ---------------- btMultiBody::calcAccelerationDeltasMultiDof ----------------
 * force: [base torque(3, world) base force(3, world) joint forces(ndof)], out same layout. */
static void accel_deltas(const rs_model_desc *m, const oenv *e, const double *force, double *out) {
    int n = m->n_links;
    sv6 Z[ML + 1], acc[ML + 1];
    double Y[ML];
    memset(Z, 0, sizeof(sv6) * (n + 1));
    if (!m->fixed_base) {
        double t[3];
        m3mulv(t, e->Rwl[0], force);     Z[0].w[0] = -t[0]; Z[0].w[1] = -t[1]; Z[0].w[2] = -t[2];
        m3mulv(t, e->Rwl[0], force + 3); Z[0].v[0] = -t[0]; Z[0].v[1] = -t[1]; Z[0].v[2] = -t[2];
    }
    for (int i = n - 1; i >= 0; i--) {
        int p = m->parent[i];
        sv6 Zt = Z[i + 1];
        if (m->dof_idx[i] >= 0) {
            Y[i] = force[6 + m->dof_idx[i]] - sv_dot(&e->S[i], &Z[i + 1]);
            double k = e->invD[i] * Y[i];
            v3axpy(Zt.w, k, e->h[i].w); v3axpy(Zt.v, k, e->h[i].v);
        }
        sv6 Zp; xf_force_inv(&Zp, e->Rpl[i], e->rvec[i], &Zt);
        v3add(Z[p + 1].w, Z[p + 1].w, Zp.w); v3add(Z[p + 1].v, Z[p + 1].v, Zp.v);
    }
    if (m->fixed_base) memset(&acc[0], 0, sizeof(sv6));
    else {
        sv6 r; m6mulv(&r, e->Ibase_inv, &Z[0]);
        for (int k = 0; k < 3; k++) { acc[0].w[k] = -r.w[k]; acc[0].v[k] = -r.v[k]; }
    }
    for (int i = 0; i < n; i++) {
        int p = m->parent[i];
        xf_motion(&acc[i + 1], e->Rpl[i], e->rvec[i], &acc[p + 1]);
        if (m->dof_idx[i] >= 0) {
            double a = e->invD[i] * (Y[i] - sv_dot(&acc[i + 1], &e->h[i]));
            out[6 + m->dof_idx[i]] = a;
            v3axpy(acc[i + 1].w, a, e->S[i].w); v3axpy(acc[i + 1].v, a, e->S[i].v);
        }
    }
    if (m->fixed_base) { for (int k = 0; k < 6; k++) out[k] = 0; }
    else { m3tmulv(out, e->Rwl[0], acc[0].w); m3tmulv(out + 3, e->Rwl[0], acc[0].v); }
    if (m->has_cube) {   /* a btMultiBody without links: M^-1 = (R^T I^-1 R, 1/m) on the world-frame base velocities */
        const double *f = force + 6 + m->n_dof, *R = e->Rwl[CUBE_BODY(m)];
        double *o = out + 6 + m->n_dof, tl[3];
        m3mulv(tl, R, f);
        for (int k = 0; k < 3; k++) tl[k] /= m->cube_inertia[k];
        m3tmulv(o, R, tl);
        for (int k = 0; k < 3; k++) o[3 + k] = f[3 + k] / m->cube_mass;
    }
}

/* ---------------- btMultiBody::fillContactJacobianMultiDof ----------------
 * jac . [w_base v_base qd] = n . (velocity of world point p rigidly attached to body b)  (b: 0=base, i+1=link i) */
static void contact_jacobian(const rs_model_desc *m, const oenv *e, int b, const double *p, const double *n, double *jac) {
    memset(jac, 0, sizeof(double) * nd_of(m));
    if (m->has_cube && b == CUBE_BODY(m)) {
        double r[3]; v3sub(r, p, e->pos[b]);
        v3cross(jac + 6 + m->n_dof, r, n);
        v3cpy(jac + 6 + m->n_dof + 3, n);
        return;
    }
    if (!m->fixed_base) {
        double r[3]; v3sub(r, p, e->pos[0]);
        v3cross(jac, r, n);
        v3cpy(jac + 3, n);
    }
    int l = b - 1;
    while (l >= 0) {
        if (m->dof_idx[l] >= 0) {
            double aw[3];
            m3tmulv(aw, e->Rwl[l + 1], m->axis[l]);      /* axis in world */
            if (m->jtype[l] == RS_JOINT_REVOLUTE) {
                /* pivot = link COM - R^T d */
                double dw[3], piv[3], r[3], c[3];
                m3tmulv(dw, e->Rwl[l + 1], m->d_vec[l]);
                v3sub(piv, e->pos[l + 1], dw);
                v3sub(r, p, piv);
                v3cross(c, r, n);
                jac[6 + m->dof_idx[l]] = v3dot(aw, c);
            } else jac[6 + m->dof_idx[l]] = v3dot(aw, n);
        }
        l = m->parent[l];
    }
}

/* ---------------- narrowphase + persistent manifolds ---------------- */
static void link_to_world(const oenv *e, int b, const double *local, double *w) {
    double t[3]; m3tmulv(t, e->Rwl[b], local); v3add(w, e->pos[b], t);
}
static void world_to_link(const oenv *e, int b, const double *w, double *local) {
    double t[3]; v3sub(t, w, e->pos[b]); m3mulv(local, e->Rwl[b], t);
}

/* btPersistentManifold::sortCachedPoints (gContactCalcArea3Points = true) */
static int sort_cached_points(const manifold *mf, const cpoint *pt) {
    int maxPenetrationIndex = -1;
    double maxPenetration = pt->dist;
    for (int i = 0; i < 4; i++) if (mf->p[i].dist < maxPenetration) { maxPenetrationIndex = i; maxPenetration = mf->p[i].dist; }
    double res[4] = {0, 0, 0, 0}, a[3], b[3], c[3];
    if (maxPenetrationIndex != 0) { v3sub(a, pt->localA, mf->p[1].localA); v3sub(b, mf->p[3].localA, mf->p[2].localA); v3cross(c, a, b); res[0] = v3dot(c, c); }
    if (maxPenetrationIndex != 1) { v3sub(a, pt->localA, mf->p[0].localA); v3sub(b, mf->p[3].localA, mf->p[2].localA); v3cross(c, a, b); res[1] = v3dot(c, c); }
    if (maxPenetrationIndex != 2) { v3sub(a, pt->localA, mf->p[0].localA); v3sub(b, mf->p[3].localA, mf->p[1].localA); v3cross(c, a, b); res[2] = v3dot(c, c); }
    if (maxPenetrationIndex != 3) { v3sub(a, pt->localA, mf->p[0].localA); v3sub(b, mf->p[2].localA, mf->p[1].localA); v3cross(c, a, b); res[3] = v3dot(c, c); }
    /* btVector4::closestAxis4 -> absolute().maxAxis4(): first strict maximum */
    int best = -1; double mv = -1e300;
    for (int i = 0; i < 4; i++) if (fabs(res[i]) > mv) { mv = fabs(res[i]); best = i; }
    return best;
}

/* btManifoldResult::addContactPoint + btPersistentManifold::getCacheEntry/addManifoldPoint/replaceContactPoint */
static void manifold_add(manifold *mf, const oenv *e, int bA, int bB, const double *n, const double *pOnB, double depth,
                         double thresh, double friction) {
    if (depth > thresh) return;
    cpoint np; memset(&np, 0, sizeof np);
    double pA[3] = {pOnB[0] + n[0] * depth, pOnB[1] + n[1] * depth, pOnB[2] + n[2] * depth};
    world_to_link(e, bA, pA, np.localA);
    if (bB >= 0) world_to_link(e, bB, pOnB, np.localB); else v3cpy(np.localB, pOnB);
    v3cpy(np.normal, n); v3cpy(np.posA, pA); v3cpy(np.posB, pOnB);
    np.dist = depth; np.friction = friction;
    double shortest = thresh * thresh; int nearest = -1;
    for (int i = 0; i < mf->n; i++) {
        double d[3]; v3sub(d, mf->p[i].localA, np.localA);
        double dd = v3dot(d, d);
        if (dd < shortest) { shortest = dd; nearest = i; }
    }
    if (nearest >= 0) {
        /* replaceContactPoint: keeps lifetime and applied impulses */
        np.lifetime = mf->p[nearest].lifetime;
        np.imp_n = mf->p[nearest].imp_n; np.imp_f1 = mf->p[nearest].imp_f1; np.imp_f2 = mf->p[nearest].imp_f2;
        mf->p[nearest] = np;
    } else {
        int idx = mf->n;
        if (idx == MANIFOLD_CACHE_SIZE) idx = sort_cached_points(mf, &np); else mf->n++;
        mf->p[idx] = np;
    }
}

/* btPersistentManifold::refreshContactPoints */
static void manifold_refresh(manifold *mf, const oenv *e, int bA, int bB, double thresh) {
    for (int i = mf->n - 1; i >= 0; i--) {
        cpoint *p = &mf->p[i];
        link_to_world(e, bA, p->localA, p->posA);
        if (bB >= 0) link_to_world(e, bB, p->localB, p->posB); else v3cpy(p->posB, p->localB);
        double d[3]; v3sub(d, p->posA, p->posB);
        p->dist = v3dot(d, p->normal);
        p->lifetime++;
    }
    for (int i = mf->n - 1; i >= 0; i--) {
        cpoint *p = &mf->p[i];
        int remove = 0;
        if (!(p->dist <= thresh)) remove = 1;
        else {
            double proj[3], diff[3];
            for (int k = 0; k < 3; k++) proj[k] = p->posA[k] - p->normal[k] * p->dist;
            v3sub(diff, p->posB, proj);
            if (v3dot(diff, diff) > thresh * thresh) remove = 1;
        }
        if (remove) { mf->p[i] = mf->p[mf->n - 1]; mf->n--; }
    }
}

/* closest points between segments [p1,q1] and [p2,q2] (the analytic answer GJK converges to
 * for btMultiSphereShape capsules) */
static void seg_seg_closest(const double *p1, const double *q1, const double *p2, const double *q2, double *c1, double *c2) {
    double d1[3], d2[3], r[3];
    v3sub(d1, q1, p1); v3sub(d2, q2, p2); v3sub(r, p1, p2);
    double a = v3dot(d1, d1), e = v3dot(d2, d2), f = v3dot(d2, r), s, t;
    const double EPS = 1e-12;
    if (a <= EPS && e <= EPS) { s = t = 0; }
    else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
    else {
        double c = v3dot(d1, r);
        if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        else {
            double b = v3dot(d1, d2), denom = a * e - b * b;
            if (denom > EPS * a * e) { s = (b * f - c * e) / denom; s = s < 0 ? 0 : (s > 1 ? 1 : s); } else s = 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
            else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
        }
    }
    for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
}

/* Closest points between a solid box (half extents h, axis aligned in its own frame) and the segment [a,b] (box coords):
 * the analytic answer of GJK (separated) / of the minimum-translation search EPA performs (core inside the box) for
 * btBoxShape against the sphere-swept cores of sphere / capsule / btMultiSphereShape geoms.
 * Returns the signed distance box <-> segment core; n: unit vector from the box towards the segment, pb: point on the box,
 * ps: point on the segment.
 * Separated: dist^2(p(t), box) is convex and piecewise quadratic in t; its derivative G(t) = sum_i excess_i(t) d_i is
 * piecewise linear and non-decreasing with kinks where a coordinate crosses a face, so the root lies between the two
 * bracketing candidates {0, 1, kinks} and follows by linear interpolation.
 * Touching / penetrating: separating-axis search over the 3 face normals and the 3 edge directions d x e_i; the axis of
 * least overlap gives normal and depth, the witness point on the segment is its deeper end (its midpoint for an edge axis). */
static double box_seg_G(const double *h, const double *a, const double *d, double t) {
    double g = 0;
    for (int i = 0; i < 3; i++) {
        double p = a[i] + t * d[i];
        double ex = p > h[i] ? p - h[i] : (p < -h[i] ? p + h[i] : 0.0);
        g += ex * d[i];
    }
    return g;
}
static double box_seg_closest(const double *h, const double *a, const double *b, double *n, double *pb, double *ps) {
    double d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double cand[8]; int nc = 0;
    cand[nc++] = 0.0; cand[nc++] = 1.0;
    for (int i = 0; i < 3; i++) {
        if (fabs(d[i]) > 1e-12) {
            double t0 = (h[i] - a[i]) / d[i], t1 = (-h[i] - a[i]) / d[i];
            cand[nc++] = t0 > 0 && t0 < 1 ? t0 : 0.0;
            cand[nc++] = t1 > 0 && t1 < 1 ? t1 : 0.0;
        } else { cand[nc++] = 0.0; cand[nc++] = 0.0; }
    }
    double t;
    double G0 = box_seg_G(h, a, d, 0.0), G1 = box_seg_G(h, a, d, 1.0);
    if (G0 >= 0) t = 0.0;
    else if (G1 <= 0) t = 1.0;
    else {
        double tl = 0.0, gl = G0, th = 1.0, gh = G1;
        for (int k = 2; k < 8; k++) {
            double tc = cand[k], gc = box_seg_G(h, a, d, tc);
            if (gc <= 0) { if (tc > tl) { tl = tc; gl = gc; } }
            else if (tc < th) { th = tc; gh = gc; }
        }
        t = gh - gl > 0 ? tl + (th - tl) * (-gl) / (gh - gl) : tl;
    }
    double df[3], dist2 = 0;
    for (int i = 0; i < 3; i++) {
        ps[i] = a[i] + t * d[i];
        pb[i] = ps[i] > h[i] ? h[i] : (ps[i] < -h[i] ? -h[i] : ps[i]);
        df[i] = ps[i] - pb[i]; dist2 += df[i] * df[i];
    }
    if (dist2 > 1e-18) {
        double dist = sqrt(dist2);
        for (int i = 0; i < 3; i++) n[i] = df[i] / dist;
        return dist;
    }
    /* core touches or intersects the box: minimum-translation axis */
    double c[3] = {0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1]), 0.5 * (a[2] + b[2])}, hd[3] = {0.5 * d[0], 0.5 * d[1], 0.5 * d[2]};
    double best = 1e300, bn[3] = {0, 0, 1}; int best_edge = 0;
    for (int i = 0; i < 3; i++) {
        double ov = h[i] + fabs(hd[i]) - fabs(c[i]);
        if (ov < best) { best = ov; bn[0] = bn[1] = bn[2] = 0; bn[i] = c[i] >= 0 ? 1.0 : -1.0; best_edge = 0; }
    }
    for (int i = 0; i < 3; i++) {
        double u[3] = {0, 0, 0}, ei[3] = {0, 0, 0}; ei[i] = 1;
        v3cross(u, d, ei);
        double l2 = v3dot(u, u);
        if (l2 < 1e-18) continue;
        double il = 1.0 / sqrt(l2); u[0] *= il; u[1] *= il; u[2] *= il;
        double cu = v3dot(c, u);
        double ov = h[0] * fabs(u[0]) + h[1] * fabs(u[1]) + h[2] * fabs(u[2]) - fabs(cu);
        if (ov < best) { best = ov; double sg = cu >= 0 ? 1.0 : -1.0; bn[0] = sg * u[0]; bn[1] = sg * u[1]; bn[2] = sg * u[2]; best_edge = 1; }
    }
    v3cpy(n, bn);
    if (best_edge) v3cpy(ps, c);
    else { double da = v3dot(a, bn), db = v3dot(b, bn); if (db < da) v3cpy(ps, b); else v3cpy(ps, a); }
    for (int i = 0; i < 3; i++) pb[i] = ps[i] + bn[i] * best;
    return -best;
}

static void collide(const rs_model_desc *m, oenv *e) {
    /* floor: btConvexPlaneCollisionAlgorithm::collideSingleContact -- one support vertex per geom per
     * sub-step into that (geom, plane) manifold, then refreshContactPoints */
    const double nz[3] = {0, 0, 1};
    for (int g = 0; g < m->n_geoms; g++) {
        manifold *mf = &e->man[g];
        if (!m->g_floor[g]) { mf->n = 0; continue; }
        int b = m->g_link[g] + 1;
        double thr = m->break_thresh[b]; /* min(link compound, plane=huge) */
        double c0[3], c1[3];
        link_to_world(e, b, m->g_p0[g], c0);
        const double *c = c0;
        if (m->g_type[g] == RS_GEOM_CAPSULE) {
            link_to_world(e, b, m->g_p1[g], c1);
            if (c1[2] < c0[2]) c = c1;   /* support in -n: first sphere wins ties */
        }
        double vtx[3] = {c[0], c[1], c[2] - m->g_radius[g]};
        if (m->g_type[g] == RS_GEOM_BOX) {
            /* btBoxShape::localGetSupportingVertex(R^T (-n)): per box axis +half if the component is >= 0, else -half */
            double Rb[9], dl[3], db[3], sv[3], vl[3], t[3];
            const double down[3] = {0, 0, -1};
            quat_to_mat(Rb, m->g_quat[g]);                  /* box coords -> link coords */
            m3mulv(dl, e->Rwl[b], down);                    /* -n in link coords */
            m3tmulv(db, Rb, dl);                            /* -n in box coords */
            for (int k = 0; k < 3; k++) sv[k] = db[k] >= 0 ? m->g_p1[g][k] : -m->g_p1[g][k];
            m3mulv(t, Rb, sv);
            v3add(vl, m->g_p0[g], t);
            link_to_world(e, b, vl, vtx);
        }
        double dist = vtx[2];
        if (dist < thr) {
            double pOnB[3] = {vtx[0], vtx[1], 0.0};
            manifold_add(mf, e, b, -1, nz, pOnB, dist, thr, m->g_friction[g] * m->floor_friction);
        }
        if (mf->n) manifold_refresh(mf, e, b, -1, thr);
    }
    /* self collision: closest points between sphere/capsule cores */
    for (int k = 0; k < m->n_pairs; k++) {
        manifold *mf = &e->man[m->n_geoms + k];
        int ga = m->pair_a[k], gb = m->pair_b[k];
        int bA = m->g_link[ga] + 1, bB = m->g_link[gb] + 1;
        double thr = m->break_thresh[bA] < m->break_thresh[bB] ? m->break_thresh[bA] : m->break_thresh[bB];
        double a0[3], a1[3], b0[3], b1[3], cA[3], cB[3];
        link_to_world(e, bA, m->g_p0[ga], a0); link_to_world(e, bA, m->g_p1[ga], a1);
        link_to_world(e, bB, m->g_p0[gb], b0); link_to_world(e, bB, m->g_p1[gb], b1);
        seg_seg_closest(a0, a1, b0, b1, cA, cB);
        double d[3]; v3sub(d, cA, cB);
        double len = v3len(d), rA = m->g_radius[ga], rB = m->g_radius[gb];
        double dist = len - rA - rB;
        int both_spheres = m->g_type[ga] == RS_GEOM_SPHERE && m->g_type[gb] == RS_GEOM_SPHERE;
        /* btSphereSphereCollisionAlgorithm only reports penetrating pairs; GJK pairs report up to the breaking threshold */
        int hit = both_spheres ? (dist <= 0.0) : (dist < thr);
        if (hit) {
            double n[3] = {1, 0, 0};
            if (len > 1e-12) { n[0] = d[0] / len; n[1] = d[1] / len; n[2] = d[2] / len; }
            double pOnB[3] = {cB[0] + n[0] * rB, cB[1] + n[1] * rB, cB[2] + n[2] * rB};
            manifold_add(mf, e, bA, bB, n, pOnB, dist, thr, m->g_friction[ga] * m->g_friction[gb]);
        }
        if (mf->n) manifold_refresh(mf, e, bA, bB, thr);
    }
    if (m->has_cube) {
        /* the cube (btBoxShape): against the floor plane like any convex (one support vertex per sub-step), against every
         * robot geom through the closest points of box and sphere-swept segment.  Threshold: min of the two bodies' = the cube's */
        const int cb = CUBE_BODY(m), k0 = m->n_geoms + m->n_pairs;
        const double thr = m->cube_break_thresh;
        {
            manifold *mf = &e->man[k0];
            const double down[3] = {0, 0, -1};
            double dl[3], sv[3], vtx[3];
            m3mulv(dl, e->Rwl[cb], down);
            for (int k = 0; k < 3; k++) sv[k] = dl[k] >= 0 ? m->cube_half[k] : -m->cube_half[k];
            link_to_world(e, cb, sv, vtx);
            double dist = vtx[2];
            if (dist < thr) {
                double pOnB[3] = {vtx[0], vtx[1], 0.0};
                manifold_add(mf, e, cb, -1, nz, pOnB, dist, thr, m->cube_friction * m->floor_friction);
            }
            if (mf->n) manifold_refresh(mf, e, cb, -1, thr);
        }
        for (int g = 0; g < m->n_geoms; g++) {
            manifold *mf = &e->man[k0 + 1 + g];
            int bA = m->g_link[g] + 1;
            if (m->g_type[g] == RS_GEOM_BOX) { mf->n = 0; continue; }   /* box-box needs a polytope narrowphase: no shipped model has it */
            double a0[3], a1[3], al[3], bl[3], n[3], pb[3], ps[3];
            link_to_world(e, bA, m->g_p0[g], a0); link_to_world(e, bA, m->g_p1[g], a1);
            world_to_link(e, cb, a0, al); world_to_link(e, cb, a1, bl);
            double rA = m->g_radius[g];
            /* broadphase: bounding spheres */
            double mid[3] = {0.5 * (al[0] + bl[0]), 0.5 * (al[1] + bl[1]), 0.5 * (al[2] + bl[2])}, hl[3];
            v3sub(hl, bl, al);
            double reach = 0.5 * v3len(hl) + rA + v3len(m->cube_half) + thr;
            if (v3dot(mid, mid) <= reach * reach) {
                double dist = box_seg_closest(m->cube_half, al, bl, n, pb, ps) - rA;
                if (dist < thr) {
                    double nw[3], pOnB[3];
                    m3tmulv(nw, e->Rwl[cb], n);
                    link_to_world(e, cb, pb, pOnB);
                    manifold_add(mf, e, bA, cb, nw, pOnB, dist, thr, m->g_friction[g] * m->cube_friction);
                }
            }
            if (mf->n) manifold_refresh(mf, e, bA, cb, thr);
        }
    }
    /* bounded live-point pool shared with the CUDA path: points beyond max_contacts are dropped in manifold order */
    int total = 0;
    for (int k = 0; k < nman_of(m); k++) {
        manifold *mf = &e->man[k];
        if (total + mf->n > m->max_contacts) mf->n = m->max_contacts - total;
        total += mf->n;
    }
}

/* ---------------- constraint rows: btMultiBodyConstraintSolver ---------------- */
typedef struct {
    double jac[MAXDOF];     /* jacA + jacB (both act on the same multibody) */
    double unit[MAXDOF];    /* deltaVelocitiesUnitImpulse (A + B)           */
    double inv_denom, rhs, lo, hi, applied, friction;
    int friction_index;     /* friction rows: index of their normal row     */
    cpoint *cp; int which;  /* write-back target                            */
} crow;

static double row_finish(const rs_model_desc *m, const oenv *e, crow *r, const double *jA, const double *jB, int nd) {
    /* fillMultiBodyConstraint / setupMultiBodyContactConstraint: unit-impulse responses per body side,
     * denom = jA.uA + jB.uB (Bullet treats the two sides independently even on one multibody) */
    double uA[MAXDOF], uB[MAXDOF], denom = 0, relvel = 0;
    double vel[MAXDOF];
    for (int k = 0; k < 3; k++) { vel[k] = e->base_w[k]; vel[3 + k] = e->base_v[k]; }
    for (int d = 0; d < m->n_dof; d++) vel[6 + d] = e->qd[d];
    if (m->has_cube) for (int k = 0; k < 3; k++) { vel[6 + m->n_dof + k] = e->cube_w[k]; vel[6 + m->n_dof + 3 + k] = e->cube_v[k]; }
    accel_deltas(m, e, jA, uA);
    for (int k = 0; k < nd; k++) { denom += jA[k] * uA[k]; relvel += jA[k] * vel[k]; r->jac[k] = jA[k]; r->unit[k] = uA[k]; }
    if (jB) {
        accel_deltas(m, e, jB, uB);
        for (int k = 0; k < nd; k++) { denom += jB[k] * uB[k]; relvel += jB[k] * vel[k]; r->jac[k] += jB[k]; r->unit[k] += uB[k]; }
    }
    /* "disable the constraint row to handle singularity/redundant constraint" (d > SIMD_EPSILON);
     * FLT_EPSILON is used in both the oracle and the fp32 CUDA path (planar robots: exact 0) */
    r->inv_denom = denom > 1.1920929e-07 ? 1.0 / denom : 0.0;   /* relaxation 1, cfm 0 */
    r->applied = 0;
    return relvel;
}

static int setup_rows(const rs_model_desc *m, oenv *e, double dt, crow *nc, int *n_nc, crow *cn, int *n_cn, crow *cf, int *n_cf) {
    int nd = nd_of(m);
    *n_nc = *n_cn = *n_cf = 0;
    /* btMultiBodyJointLimitConstraint::createConstraintRows: row 0 lower, row 1 upper, only when violated */
    for (int i = 0; i < m->n_links; i++) {
        if (!m->has_limit[i] || m->dof_idx[i] < 0) continue;
        int d = m->dof_idx[i];
        for (int row = 0; row < 2; row++) {
            double pen = row == 0 ? e->q[d] - m->lim_lo[i] : m->lim_hi[i] - e->q[d];
            if (pen > 0 && !K.limit_always) continue;
            double jA[MAXDOF]; memset(jA, 0, sizeof jA);
            jA[6 + d] = row == 0 ? 1.0 : -1.0;
            crow *r = &nc[(*n_nc)++];
            double relvel = row_finish(m, e, r, jA, NULL, nd);
            double erp = m->erp2;
            if (pen > m->split_thresh) erp = m->erp;   /* m_splitImpulse is on by default */
            double posErr = -pen * erp / dt, velErr = -relvel;
            if (pen > 0) { posErr = 0; velErr = -pen / dt; }   /* (only reached with limit_always) "velocityError = -penetration / timeStep" */
            if (pen > m->split_thresh) r->rhs = (posErr + velErr) * r->inv_denom;
            else r->rhs = velErr * r->inv_denom;        /* m_rhsPenetration is never used for multibodies */
            r->lo = 0; r->hi = m->max_limit_impulse; r->cp = NULL; r->friction_index = -1;
        }
    }
    /* btMultiBodyJointMotor::createConstraintRows.  The motors are added to the world by createJointMotors() after the
     * import created the limit constraints, so their rows follow the limit rows.  desiredVelocity handed to
     * fillMultiBodyConstraint = kp * erp * (q* - q) / dt + qd + kd * (qd* - qd) with the motor's erp = 1; the row's
     * velocity error is that minus J.v (= qd), bounds -+ maxAppliedImpulse, no positional term. */
    for (int i = 0; i < m->n_links; i++) {
        int d = m->dof_idx[i];
        if (d < 0 || !(e->mot_imp[d] > 0)) continue;
        double jA[MAXDOF]; memset(jA, 0, sizeof jA);
        jA[6 + d] = 1.0;
        crow *r = &nc[(*n_nc)++];
        double relvel = row_finish(m, e, r, jA, NULL, nd);
        double desired = e->mot_kp[d] * (e->mot_pos[d] - e->q[d]) / dt + e->qd[d] + e->mot_kd[d] * (e->mot_vel[d] - e->qd[d]);
        r->rhs = (desired - relvel) * r->inv_denom;
        r->lo = -e->mot_imp[d]; r->hi = e->mot_imp[d]; r->cp = NULL; r->friction_index = -1;
    }
    /* contacts: btMultiBodyConstraintSolver::convertMultiBodyContact, manifold order, point order */
    int nman = nman_of(m), ngp = m->n_geoms + m->n_pairs;
    for (int k = 0; k < nman; k++) {
        manifold *mf = &e->man[k];
        int bA, bB;
        if (k < m->n_geoms) { bA = m->g_link[k] + 1; bB = -1; }
        else if (k < ngp) { bA = m->g_link[m->pair_a[k - m->n_geoms]] + 1; bB = m->g_link[m->pair_b[k - m->n_geoms]] + 1; }
        else if (k == ngp) { bA = CUBE_BODY(m); bB = -1; }
        else { bA = m->g_link[k - ngp - 1] + 1; bB = CUBE_BODY(m); }
        for (int pi = 0; pi < mf->n; pi++) {
            cpoint *cp = &mf->p[pi];
            if (*n_cn >= MAXROWS / 4) continue;
            double dirs[3][3];
            v3cpy(dirs[0], cp->normal);
            plane_space(cp->normal, dirs[1], dirs[2]);
            int normal_index = *n_cn;
            for (int w = 0; w < 1 + K.friction_dirs; w++) {
                double jA[MAXDOF], jB[MAXDOF], nneg[3] = {-dirs[w][0], -dirs[w][1], -dirs[w][2]};
                contact_jacobian(m, e, bA, cp->posA, dirs[w], jA);
                if (bB >= 0) contact_jacobian(m, e, bB, cp->posB, nneg, jB);
                crow *r = w == 0 ? &cn[(*n_cn)++] : &cf[(*n_cf)++];
                double relvel = row_finish(m, e, r, jA, bB >= 0 ? jB : NULL, nd);
                r->cp = cp; r->which = w; r->friction = cp->friction;
                if (w == 0) {
                    double pen = cp->dist;     /* + m_linearSlop (0) */
                    double erp = m->c_erp2;
                    if (pen > m->c_split_thresh) erp = m->c_erp;
                    double posErr = 0, velErr = -relvel;   /* restitution 0 */
                    if (pen > 0) velErr -= pen / dt; else posErr = -pen * erp / dt;
                    if (pen > m->c_split_thresh) r->rhs = (posErr + velErr) * r->inv_denom;
                    else r->rhs = velErr * r->inv_denom;
                    r->lo = 0; r->hi = 1e10; r->friction_index = -1;
                } else {
                    r->rhs = -relvel * r->inv_denom;
                    r->lo = -cp->friction; r->hi = cp->friction; r->friction_index = normal_index;
                }
            }
        }
    }
    return *n_nc + *n_cn + *n_cf;
}

/* resolveSingleConstraintRowGeneric */
static inline void resolve_row(crow *c, double *dv, int nd) {
    double dimp = c->rhs, dot = 0;        /* - applied*cfm (cfm 0) */
    for (int k = 0; k < nd; k++) dot += c->jac[k] * dv[k];
    dimp -= dot * c->inv_denom;
    double sum = c->applied + dimp;
    if (sum < c->lo) { dimp = c->lo - c->applied; c->applied = c->lo; }
    else if (sum > c->hi) { dimp = c->hi - c->applied; c->applied = c->hi; }
    else c->applied = sum;
    for (int k = 0; k < nd; k++) dv[k] += c->unit[k] * dimp;
}

static void solve(const rs_model_desc *m, oenv *e, double dt) {
    static __thread crow nc[3 * ML], cn[MAXROWS / 4], cf[MAXROWS / 2];
    int n_nc, n_cn, n_cf, nd = nd_of(m);
    int total = setup_rows(m, e, dt, nc, &n_nc, cn, &n_cn, cf, &n_cf);
    e->n_rows_last = total; e->n_contacts_last = n_cn;
    if (!total) return;
    double dv[MAXDOF]; memset(dv, 0, sizeof dv);
    if (K.warmstart > 0) {   /* sweep knob: applied impulses of the previous sub-step, scaled (m_warmstartingFactor) */
        for (int j = 0; j < n_cn; j++) { cn[j].applied = K.warmstart * cn[j].cp->imp_n; for (int k = 0; k < nd; k++) dv[k] += cn[j].unit[k] * cn[j].applied; }
        for (int j = 0; j < n_cf; j++) {
            cf[j].applied = K.warmstart * (cf[j].which == 1 ? cf[j].cp->imp_f1 : cf[j].cp->imp_f2);
            for (int k = 0; k < nd; k++) dv[k] += cf[j].unit[k] * cf[j].applied;
        }
    }
    /* btMultiBodyConstraintSolver::solveSingleIteration */
    for (int it = 0; it < m->n_iter; it++) {
        for (int j = 0; j < n_nc; j++) { int idx = (it & 1) ? j : n_nc - 1 - j; resolve_row(&nc[idx], dv, nd); }
        for (int j = 0; j < n_cn; j++) resolve_row(&cn[j], dv, nd);
        for (int j = 0; j < n_cf; j++) {
            double tot = cn[cf[j].friction_index].applied;
            if (tot > 0) { cf[j].lo = -cf[j].friction * tot; cf[j].hi = cf[j].friction * tot; resolve_row(&cf[j], dv, nd); }
        }
    }
    /* write back: velocities += deltaV ; applied impulses -> manifold points */
    if (!m->fixed_base) for (int k = 0; k < 3; k++) { e->base_w[k] += dv[k]; e->base_v[k] += dv[3 + k]; }
    for (int d = 0; d < m->n_dof; d++) e->qd[d] += dv[6 + d];
    if (m->has_cube) for (int k = 0; k < 3; k++) { e->cube_w[k] += dv[6 + m->n_dof + k]; e->cube_v[k] += dv[6 + m->n_dof + 3 + k]; }
    clamp_vel(m, e);
    for (int j = 0; j < n_cn; j++) cn[j].cp->imp_n = cn[j].applied;
    for (int j = 0; j < n_cf; j++) { if (cf[j].which == 1) cf[j].cp->imp_f1 = cf[j].applied; else cf[j].cp->imp_f2 = cf[j].applied; }
}

/* ---------------- btMultiBody::stepPositionsMultiDof ---------------- */
static void integrate(const rs_model_desc *m, oenv *e, double dt) {
    if (!m->fixed_base) {
        v3axpy(e->base_pos, dt, e->base_v);
        /* pQuatUpdateFun, baseBody branch, restated for a base->world quaternion: q <- dq(omega_world) * q */
        double ang = v3len(e->base_w), axis[3];
        if (ang * dt > 0.7853981633974483) ang = 0.5 * 1.5707963267948966 / dt; /* ANGULAR_MOTION_THRESHOLD */
        double k;
        if (ang < 0.001) k = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
        else k = sin(0.5 * ang * dt) / ang;
        for (int i = 0; i < 3; i++) axis[i] = e->base_w[i] * k;
        double dq[4] = {axis[0], axis[1], axis[2], cos(ang * dt * 0.5)}, nq[4];
        quat_mul(nq, dq, e->base_quat);
        double nn = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
        for (int i = 0; i < 4; i++) e->base_quat[i] = nq[i] / nn;
    }
    for (int d = 0; d < m->n_dof; d++) e->q[d] += dt * e->qd[d];
}

/* the cube is a btMultiBody without links: computeAccelerationsArticulatedBodyAlgorithmMultiDof reduces to the base terms --
 * gravity, velocity damping k v (1 + |v|) and the gyroscopic torque in body coordinates -- applied with applyDeltaVee(dt) */
static void cube_dynamics(const rs_model_desc *m, oenv *e, double dt) {
    const double *R = e->Rwl[CUBE_BODY(m)];
    double wl[3], Iw[3], g1[3], al[3], aw[3];
    m3mulv(wl, R, e->cube_w);
    double wn = v3len(wl), vn = v3len(e->cube_v);
    double ka = m->ang_damping * (1 + wn), kl = m->lin_damping * (1 + vn);
    for (int k = 0; k < 3; k++) Iw[k] = m->cube_inertia[k] * wl[k];
    v3cross(g1, wl, Iw);
    for (int k = 0; k < 3; k++) al[k] = -(Iw[k] * ka + m->use_gyro * g1[k]) / m->cube_inertia[k];
    m3tmulv(aw, R, al);
    for (int k = 0; k < 3; k++) { e->cube_w[k] += dt * aw[k]; e->cube_v[k] += dt * (-kl * e->cube_v[k]); }
    e->cube_v[2] += dt * -m->gravity;
}
static void cube_integrate(const rs_model_desc *m, oenv *e, double dt) {
    v3axpy(e->cube_pos, dt, e->cube_v);
    double ang = v3len(e->cube_w), axis[3];
    if (ang * dt > 0.7853981633974483) ang = 0.5 * 1.5707963267948966 / dt;
    double k;
    if (ang < 0.001) k = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
    else k = sin(0.5 * ang * dt) / ang;
    for (int i = 0; i < 3; i++) axis[i] = e->cube_w[i] * k;
    double dq[4] = {axis[0], axis[1], axis[2], cos(ang * dt * 0.5)}, nq[4];
    quat_mul(nq, dq, e->cube_quat);
    double nn = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
    for (int i = 0; i < 4; i++) e->cube_quat[i] = nq[i] / nn;
}

static void substep(const rs_model_desc *m, oenv *e) {
    double dt = m->timestep;
    kinematics(m, e);
    collide(m, e);      /* performDiscreteCollisionDetection */
    aba(m, e, dt);      /* solveConstraints: forward dynamics ... */
    if (m->has_cube) { cube_dynamics(m, e, dt); clamp_vel(m, e); }
    solve(m, e, dt);    /* ... then the multibody constraint solver */
    integrate(m, e, dt);
    if (m->has_cube) cube_integrate(m, e, dt);
}

/* ---------------- state readback: query_body_position (physics-bullet.cpp:440-495) ---------------- */
static void mat_to_quat(const double *R, double *q) { /* R rotates local->world ; (x,y,z,w) */
    double t = R[0] + R[4] + R[8];
    if (t > 0) { double s = sqrt(t + 1.0) * 2; q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s; }
    else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s; }
}
static void body_pose(const oenv *e, int b, double *pos, double *quat, double *linvel) {
    v3cpy(pos, e->pos[b]);
    double Rlw[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rlw[i * 3 + j] = e->Rwl[b][j * 3 + i];
    if (b == 0) memcpy(quat, e->base_quat, 4 * sizeof(double)); else mat_to_quat(Rlw, quat);
    m3tmulv(linvel, e->Rwl[b], e->vel[b].v);
}

/* Pose::rpy  (python-binding.cpp:60-73) */
static void quat_rpy(const double *q, double *r, double *p, double *y) {
    double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    double sqw = qw * qw, sqx = qx * qx, sqy = qy * qy, sqz = qz * qz;
    double t2 = -2.0 * (qx * qz - qy * qw) / (sqx + sqy + sqz + sqw);
    *y = atan2(2.0 * (qx * qy + qz * qw), (sqx - sqy - sqz + sqw));
    *r = atan2(2.0 * (qy * qz + qx * qw), (-sqx - sqy + sqz + sqw));
    t2 = t2 > 1.0 ? 1.0 : t2; t2 = t2 < -1.0 ? -1.0 : t2;
    *p = asin(t2);
}

/* foot contact flags: World::bullet_contact_list(foot) -> names (gym_forward_walker.py:106-112) */
static void foot_contacts(const rs_model_desc *m, const oenv *e, int foot, int *ground, int *other) {
    *ground = *other = 0;
    int fb = m->foot_link[foot] + 1;
    for (int g = 0; g < m->n_geoms; g++) if (m->g_link[g] + 1 == fb && e->man[g].n > 0) *ground = 1;
    for (int k = 0; k < m->n_pairs; k++) {
        int bA = m->g_link[m->pair_a[k]] + 1, bB = m->g_link[m->pair_b[k]] + 1;
        if ((bA == fb || bB == fb) && e->man[m->n_geoms + k].n > 0) *other = 1;
    }
    if (m->has_cube)   /* the cube is an object other than the floor (gym_forward_walker.py:110-113) */
        for (int g = 0; g < m->n_geoms; g++) if (m->g_link[g] + 1 == fb && e->man[m->n_geoms + m->n_pairs + 1 + g].n > 0) *other = 1;
}

/* calc_state (gym_forward_walker.py:45-76), float32 squeezes where the reference has them */
static void calc_state(const rs_model_desc *m, oenv *e, double *body_xy_dist) {
    kinematics(m, e); velocities(m, e);
    if (m->task_kind == RS_TASK_REACHER) {
        /* gym_reacher.py calc_state (:38-55) */
        int dd[2] = {m->task_dof[0], m->task_dof[1]};
        float rel[4];
        for (int k = 0; k < 2; k++) {
            int li = -1;
            for (int i = 0; i < m->n_links; i++) if (m->dof_idx[i] == dd[k]) li = i;
            float pos = (float)e->q[dd[k]], spd = (float)e->qd[dd[k]];
            float l1 = (float)m->lim_lo[li], l2 = (float)m->lim_hi[li];
            if (l1 < l2) { float mid = (float)(0.5 * (l1 + l2)); pos = 2 * (pos - mid) / (l2 - l1); }
            if (m->max_vel[li] > 0) spd /= (float)m->max_vel[li]; else spd *= (m->jtype[li] == RS_JOINT_REVOLUTE) ? 0.1f : 0.5f;
            rel[2 * k] = pos; rel[2 * k + 1] = spd;
        }
        const double *pf = e->pos[m->task_link + 1], *pt = e->pos[m->task_link2 + 1];
        double v[3] = {pf[0] - pt[0], pf[1] - pt[1], pf[2] - pt[2]};
        e->obs[0] = (float)e->q[m->task_dof[2]]; e->obs[1] = (float)e->q[m->task_dof[3]];
        e->obs[2] = v[0]; e->obs[3] = v[1];
        e->obs[4] = cos((double)rel[0]); e->obs[5] = sin((double)rel[0]); e->obs[6] = rel[1];
        e->obs[7] = rel[2]; e->obs[8] = rel[3];
        *body_xy_dist = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        return;
    }
    if (m->task_kind != RS_TASK_WALKER) {
        /* gym_pendulums.py calc_state (:35-47, :97-105): current_position() returns floats */
        int ds = m->task_dof[0], d1 = m->task_dof[1], k = 0;
        e->obs[k++] = (float)e->q[ds]; e->obs[k++] = (float)e->qd[ds];
        if (m->task_kind == RS_TASK_DOUBLE_PENDULUM) e->obs[k++] = e->pos[m->task_link + 1][0];
        e->obs[k++] = cos((double)(float)e->q[d1]); e->obs[k++] = sin((double)(float)e->q[d1]); e->obs[k++] = (float)e->qd[d1];
        if (m->task_kind == RS_TASK_DOUBLE_PENDULUM) {
            int d2 = m->task_dof[2];
            e->obs[k++] = cos((double)(float)e->q[d2]); e->obs[k++] = sin((double)(float)e->q[d2]); e->obs[k++] = (float)e->qd[d2];
        }
        *body_xy_dist = 0;
        return;
    }
    int A = m->n_act;
    float j[2 * RS_MAX_ACT];
    int at_limit = 0;
    for (int k = 0; k < A; k++) {
        int d = m->act_dof[k], li = -1;
        for (int i = 0; i < m->n_links; i++) if (m->dof_idx[i] == d) li = i;
        float pos = (float)e->q[d], spd = (float)e->qd[d];        /* Joint::joint_current_position is a float */
        float l1 = (float)m->lim_lo[li], l2 = (float)m->lim_hi[li];
        if (l1 < l2) { float mid = (float)(0.5 * (l1 + l2)); pos = 2 * (pos - mid) / (l2 - l1); }
        if (m->max_vel[li] > 0) spd /= (float)m->max_vel[li];
        else spd *= (m->jtype[li] == RS_JOINT_REVOLUTE) ? 0.1f : 0.5f;
        j[2 * k] = pos; j[2 * k + 1] = spd;
        if (fabsf(pos) > 0.99f) at_limit++;
    }
    e->rewards5[3] = m->joints_at_limit_cost * at_limit;
    double bp[3], bq[4], bv[3];
    body_pose(e, m->body_link + 1, bp, bq, bv);
    double mx = 0, my = 0;
    for (int k = 0; k < m->n_parts; k++) { mx += e->pos[m->part_link[k] + 1][0]; my += e->pos[m->part_link[k] + 1][1]; }
    mx /= m->n_parts; my /= m->n_parts;
    double z = bp[2], r, p, yaw;
    quat_rpy(bq, &r, &p, &yaw);
    e->body_xyz[0] = mx; e->body_xyz[1] = my; e->body_xyz[2] = z; v3cpy(e->body_vel, bv);
    if (e->initial_z < 0) e->initial_z = z;
    double theta = atan2(e->target_y - my, e->target_x - mx);
    double dy = e->target_y - my, dx = e->target_x - mx;
    *body_xy_dist = sqrt(dy * dy + dx * dx);
    double ang = theta - yaw;
    double cy = cos(-yaw), sy = sin(-yaw);
    double vx = cy * bv[0] - sy * bv[1], vy = sy * bv[0] + cy * bv[1], vz = bv[2];
    float more[8] = {(float)(z - e->initial_z), (float)sin(ang), (float)cos(ang), (float)(0.3 * vx), (float)(0.3 * vy), (float)(0.3 * vz), (float)r, (float)p};
    int o = 0;
    for (int k = 0; k < 8; k++) e->obs[o++] = more[k];
    for (int k = 0; k < 2 * A; k++) e->obs[o++] = j[k];
    for (int k = 0; k < m->n_feet; k++) e->obs[o++] = (float)e->feet_contact[k];
    for (int k = 0; k < o; k++) { if (e->obs[k] > 5) e->obs[k] = 5; if (e->obs[k] < -5) e->obs[k] = -5; }
    /* keep pitch for alive_bonus */
    e->rewards5[0] = p; /* temporarily */
}

static double rs_uniform(uint32_t seed, uint32_t env, uint32_t episode, uint32_t k);

/* RoboschoolHumanoidFlagrun.flag_reposition (gym_humanoid_flagrun.py:21-35).  The reference draws from
 * np_random in this order: [randint(2) if first,] uniform x [, uniform y]; here the draws are the counter-based
 * stream (seed, env, episode, 2000 + n). */
static void flag_reposition(const rs_model_desc *m, oenv *e, uint32_t seed, uint32_t env_id, int first) {
    int straight = 0;
    if (first) straight = rs_uniform(seed, env_id, e->episode, 2000u + e->flag_draws++) < 0.5;
    if (straight) {
        double u = rs_uniform(seed, env_id, e->episode, 2000u + e->flag_draws++);
        e->target_x = 0.5 * m->flag_halflen + (m->flag_halflen - 0.5 * m->flag_halflen) * u;
        e->target_y = 0.0;
    } else {
        double ux = rs_uniform(seed, env_id, e->episode, 2000u + e->flag_draws++);
        double uy = rs_uniform(seed, env_id, e->episode, 2000u + e->flag_draws++);
        double x = -m->flag_halflen + (m->flag_halflen - -m->flag_halflen) * ux;
        double y = -m->flag_halfwidth + (m->flag_halfwidth - -m->flag_halfwidth) * uy;
        e->target_x = x * 0.5; e->target_y = y * 0.5;
    }
    e->flag_timeout = m->flag_timeout_steps;
}

/* FlagrunHarder.potential_leak (gym_humanoid_flagrun.py:141-147); valid after calc_state (kinematics) */
static double harder_leak(const rs_model_desc *m, const oenv *e) {
    double z1 = e->pos[m->body_link + 1][2], z2 = e->pos[m->task_link + 1][2];
    double z = 0.7 * z1 + 0.3 * z2;
    z = z < 0 ? 0 : (z > 0.7 ? 0.7 : z);
    return 0.5 + 1.5 * (z / 0.7);
}
/* FlagrunHarder.calc_potential (gym_humanoid_flagrun.py:149-169): progress towards the flag does not count while crawling */
static double harder_potential(const rs_model_desc *m, oenv *e, double dist) {
    double frp = -dist / m->dt_env;
    if (e->pos[m->body_link + 1][2] < 0.8) {
        if (!e->crawl_valid) { e->crawl_start = frp - e->crawl_ignored; e->crawl_valid = 1; }
        e->crawl_ignored = frp - e->crawl_start;
        frp = e->crawl_start;
    } else {
        frp -= e->crawl_ignored;
        e->crawl_valid = 0;
    }
    return frp + harder_leak(m, e) * 100;
}
static double task_potential(const rs_model_desc *m, oenv *e, double dist) {
    if (m->task_kind == RS_TASK_REACHER) return -100.0 * dist;
    if (m->flag_task == 2) return harder_potential(m, e, dist);
    return -dist / m->dt_env;
}
/* FlagrunHarder.alive_bonus (gym_humanoid_flagrun.py:101-139).  Every 30 frames (after frame 100, robot not on the ground)
 * the cube is thrown at the position the robot will have reached when it arrives: five np_random draws (angle, speed,
 * 3 x velocity noise) from the flag stream, then Robot::set_pose_and_speed (python-binding.cpp:289 -> robot_move,
 * physics-bullet.cpp:579-591: pose in doubles, velocity through float, orientation identity, angular velocity zeroed
 * by the pose command).  Worlds without the cube (has_cube 0) still consume the draws. */
static double harder_alive(const rs_model_desc *m, oenv *e, double z, uint32_t seed, uint32_t env_id) {
    if (e->frame % 30 == 0 && e->frame > 100 && e->on_ground == 0) {
        double u[5];
        for (int k = 0; k < 5; k++) u[k] = rs_uniform(seed, env_id, e->episode, 2000u + e->flag_draws++);
        if (m->has_cube) {
            double target[3], cp[3], v[3];
            v3cpy(target, e->body_xyz);
            double angle = -3.14 + (3.14 - -3.14) * u[0], from_dist = 4.0;
            double attack_speed = 20.0 + (30.0 - 20.0) * u[1];
            double ttt = from_dist / attack_speed;
            v3axpy(target, ttt, e->body_vel);
            cp[0] = target[0] + from_dist * cos(angle); cp[1] = target[1] + from_dist * sin(angle); cp[2] = target[2] + 1.0;
            v3sub(v, target, cp);
            double sc = attack_speed / v3len(v);
            for (int k = 0; k < 3; k++) v[k] = v[k] * sc + (-1.0 + (1.0 - -1.0) * u[2 + k]);
            v3cpy(e->cube_pos, cp);
            e->cube_quat[0] = e->cube_quat[1] = e->cube_quat[2] = 0; e->cube_quat[3] = 1;
            for (int k = 0; k < 3; k++) { e->cube_v[k] = (double)(float)v[k]; e->cube_w[k] = 0; }
        }
    }
    if (z < 0.8) {
        if (e->task == 0) e->on_ground = 10000;      /* TASK_WALK: ends the episode immediately */
        e->on_ground += 1;
        e->elec_cost = -2.0 / 5.0;                   /* RoboschoolHumanoid.electricity_cost (class attribute, -2.0) / 5 */
    } else if (e->on_ground > 0) {
        e->on_ground -= 1;
        e->elec_cost = -2.0 / 2.0;
    } else {
        e->elec_cost = -2.0 / 2.0;
    }
    return e->on_ground < 170 ? harder_leak(m, e) : -1;
}
/* RepeatUnderlearnedTasks (gym_humanoid_flagrun.py:47-64) */
static int harder_pick_task(oenv *e, double u) {
    double prob[3], sum = 0, cdf[3];
    for (int i = 0; i < 3; i++) {
        int len = (int)(e->hist[i] >> 16), nz = __builtin_popcount(e->hist[i] & 0x3FFu);
        prob[i] = 1 - nz / (1.0 + len);
        sum += prob[i];
    }
    double acc = 0;
    for (int i = 0; i < 3; i++) { acc += prob[i] / sum; cdf[i] = acc; }
    int idx = 0;
    for (int i = 0; i < 3; i++) if (cdf[i] / cdf[2] <= u) idx++;
    return idx > 2 ? 2 : idx;
}
static void harder_task_completed(oenv *e, int completed) {
    uint32_t h = e->hist[e->task];
    uint32_t len = h >> 16; if (len < 10) len++;
    e->hist[e->task] = (((h << 1) | (completed ? 1u : 0u)) & 0x3FFu) | (len << 16);
}

/* RoboschoolHumanoidFlagrun.calc_state (gym_humanoid_flagrun.py:37-44) */
static void calc_state_flag(const rs_model_desc *m, oenv *e, double *dist, uint32_t seed, uint32_t env_id) {
    if (!m->flag_task) { calc_state(m, e, dist); return; }
    e->flag_timeout -= 1;
    calc_state(m, e, dist);
    if (*dist < 1.0 || e->flag_timeout <= 0) {
        flag_reposition(m, e, seed, env_id, 0);
        calc_state(m, e, dist);
        e->potential = task_potential(m, e, *dist);
    }
}

/* RoboschoolForwardWalker.step epilogue (gym_forward_walker.py:94-133) */
static void epilogue(const rs_model_desc *m, oenv *e, const double *a, uint32_t seed, uint32_t env_id) {
    double dist;
    calc_state_flag(m, e, &dist, seed, env_id);
    if (m->task_kind == RS_TASK_REACHER) {
        /* gym_reacher.py step (:60-79) */
        double pot_old = e->potential;
        e->potential = -100.0 * dist;
        double thd = e->obs[6], ga = e->obs[7], gad = e->obs[8];
        double elec = -0.10 * (fabs(a[0] * thd) + fabs(a[1] * gad)) - 0.01 * (fabs(a[0]) + fabs(a[1]));
        double stuck = fabs(fabs(ga) - 1) < 0.01 ? -0.1 : 0.0;
        memset(e->rewards5, 0, sizeof e->rewards5);
        e->rewards5[0] = e->potential - pot_old; e->rewards5[1] = elec; e->rewards5[2] = stuck;
        e->reward = e->rewards5[0] + elec + stuck; e->done = 0; e->frame++;
        return;
    }
    if (m->task_kind != RS_TASK_WALKER) {
        /* gym_pendulums.py step (:49-65, :107-123) */
        double reward; int dn;
        memset(e->rewards5, 0, sizeof e->rewards5);
        if (m->task_kind == RS_TASK_PENDULUM) {
            double th = (double)(float)e->q[m->task_dof[1]];
            if (m->task_swingup) { reward = cos(th); dn = 0; } else { reward = 1.0; dn = fabs(th) > 0.2; }
            e->rewards5[0] = reward;
        } else {
            double px = e->pos[m->task_link + 1][0], py = e->pos[m->task_link + 1][2];
            double pen = 0.01 * px * px + (py + 0.3 - 2.0) * (py + 0.3 - 2.0);
            e->rewards5[0] = 10.0; e->rewards5[1] = -pen;
            reward = 10.0 - pen; dn = (py + 0.3) <= 1.0;
        }
        for (int k = 0; k < m->obs_dim; k++) if (!isfinite(e->obs[k])) dn = 1;
        e->reward = reward; e->done = dn; e->frame++;
        return;
    }
    double pitch = e->rewards5[0];
    double z = e->obs[0] + e->initial_z;   /* state[0]+initial_z, state already float32 & clipped */
    double alive;
    if (m->alive_rule == RS_ALIVE_Z_AND_PITCH) alive = (z > m->alive_z && fabs(pitch) < 1.0) ? m->alive_bonus : -1;
    else if (m->flag_task == 2) alive = harder_alive(m, e, z, seed, env_id);
    else if (m->alive_rule == RS_ALIVE_Z) alive = (z > m->alive_z) ? m->alive_bonus : -1;
    else if (m->alive_rule == RS_ALIVE_CHEETAH)   /* gym_mujoco_walkers.py:50-52: stale feet_contact of shins and thighs */
        alive = (fabs(pitch) < 1.0 && !e->feet_contact[1] && !e->feet_contact[2] && !e->feet_contact[4] && !e->feet_contact[5]) ? m->alive_bonus : -1;
    else alive = m->alive_bonus;
    int done = alive < 0;
    for (int k = 0; k < m->obs_dim; k++) if (!isfinite(e->obs[k])) done = 1;
    double pot_old = e->potential;
    e->potential = task_potential(m, e, dist);
    double progress = e->potential - pot_old;
    double feet_cost = 0;
    for (int f = 0; f < m->n_feet; f++) {
        int g, o; foot_contacts(m, e, f, &g, &o);
        e->feet_contact[f] = g ? 1.0 : 0.0;
        if (o) feet_cost += m->foot_collision_cost;
    }
    int A = m->n_act;
    double s1 = 0, s2 = 0;
    /* joint_speeds = j[1::2] is taken BEFORE np.clip (gym_forward_walker.py:49,76) */
    for (int k = 0; k < A; k++) {
        s2 += a[k] * a[k];
        int d = m->act_dof[k], li = -1;
        for (int i = 0; i < m->n_links; i++) if (m->dof_idx[i] == d) li = i;
        float spd = (float)e->qd[d];
        if (m->max_vel[li] > 0) spd /= (float)m->max_vel[li]; else spd *= (m->jtype[li] == RS_JOINT_REVOLUTE) ? 0.1f : 0.5f;
        s1 += fabs(a[k] * (double)spd);
    }
    double elec = (m->flag_task == 2 ? e->elec_cost : m->electricity_cost) * (s1 / A) + m->stall_torque_cost * (s2 / A);
    double jl = e->rewards5[3];
    e->rewards5[0] = alive; e->rewards5[1] = progress; e->rewards5[2] = elec; e->rewards5[3] = jl; e->rewards5[4] = feet_cost;
    e->reward = alive + progress + elec + jl + feet_cost;
    e->done = done;
    e->frame++;
    if (m->flag_task == 2 && ((done && !e->done_latch) || e->frame == m->max_episode_steps))   /* episode_over (gym_forward_walker.py:127-128) */
        harder_task_completed(e, e->frame == m->max_episode_steps);
    e->done_latch |= done;
}

/* counter-based RNG shared bit-for-bit with the CUDA path (integer hash -> 24-bit uniform) */
static uint32_t rs_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = 0x9E3779B9u ^ a;
    h = (h ^ (h >> 16)) * 0x85EBCA6Bu; h ^= b + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 13)) * 0xC2B2AE35u; h ^= c + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 16)) * 0x85EBCA6Bu; h ^= d + 0x9E3779B9u + (h << 6) + (h >> 2);
    h = (h ^ (h >> 13)) * 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static double rs_uniform(uint32_t seed, uint32_t env, uint32_t episode, uint32_t k) {
    return (double)(rs_hash(seed, env, episode, k) >> 8) * (1.0 / 16777216.0);
}

/* reset: reference reset distribution (gym_forward_walker.py:24-30, gym_mujoco_walkers.py:106-128) */
static void env_reset(const rs_model_desc *m, oenv *e, int nman, uint32_t seed, uint32_t env_id) {
    e->episode++;
    memset(e->q, 0, sizeof e->q); memset(e->qd, 0, sizeof e->qd); memset(e->tau, 0, sizeof e->tau);
    for (int k = 0; k < m->n_reset; k++) {
        double u = rs_uniform(seed, env_id, e->episode, (uint32_t)k);
        e->q[m->reset_dof[k]] = (double)(float)(m->reset_center[k] + m->reset_spread[k] * (2.0 * u - 1.0)); /* reset_current_position(float,..) */
    }
    v3cpy(e->base_pos, m->reset_base_pos);
    memcpy(e->base_quat, m->reset_base_quat, sizeof e->base_quat);
    e->task = 0;
    if (m->flag_task == 2) e->task = harder_pick_task(e, rs_uniform(seed, env_id, e->episode, 1500u));   /* humanoid_task(): decide_best_task() */
    if (m->reset_set_base && m->reset_yaw_spread > 0) {
        double u = rs_uniform(seed, env_id, e->episode, 1000u);
        double yaw = m->reset_yaw_spread * (2.0 * u - 1.0);
        /* set_initial_orientation (gym_mujoco_walkers.py:106-128): TASK_WALK upright at z+1.4, TASK_STAND_UP lying face down
         * (pitch pi/2, z+0.45), TASK_ROLL_OVER on the back (pitch 3pi/2-0.15, z+0.22); Pose::set_rpy python-binding.cpp:42-55 */
        double pitch = 0, roll = 0;
        if (e->task == 1) { pitch = M_PI / 2; e->base_pos[2] = m->reset_base_pos[2] - 1.4 + 0.45; }
        else if (e->task == 2) { pitch = M_PI * 3 / 2 - 0.15; e->base_pos[2] = m->reset_base_pos[2] - 1.4 + 0.22; }
        double t0 = cos(yaw * 0.5), t1 = sin(yaw * 0.5), t2 = cos(roll * 0.5), t3 = sin(roll * 0.5), t4 = cos(pitch * 0.5), t5 = sin(pitch * 0.5);
        e->base_quat[3] = t0 * t2 * t4 + t1 * t3 * t5;
        e->base_quat[0] = t0 * t3 * t4 - t1 * t2 * t5;
        e->base_quat[1] = t0 * t2 * t5 + t1 * t3 * t4;
        e->base_quat[2] = t1 * t2 * t4 - t0 * t3 * t5;
    }
    v3set(e->base_w, 0, 0, 0); v3set(e->base_v, 0, 0, 0);
    if (m->has_cube) {   /* load_urdf(cube.urdf, cpose) of robot_specific_reset (gym_humanoid_flagrun.py:83-86): a fresh body at rest */
        v3cpy(e->cube_pos, m->cube_start);
        e->cube_quat[0] = e->cube_quat[1] = e->cube_quat[2] = 0; e->cube_quat[3] = 1;
        v3set(e->cube_w, 0, 0, 0); v3set(e->cube_v, 0, 0, 0);
    }
    for (int k = 0; k < nman; k++) e->man[k].n = 0;
    memset(e->feet_contact, 0, sizeof e->feet_contact);
    e->initial_z = m->initial_z;
    e->frame = 0; e->done = 0; e->done_latch = 0;
    e->target_x = m->walk_target_x; e->target_y = m->walk_target_y;
    e->flag_timeout = 0; e->flag_draws = 0;
    e->on_ground = 0; e->crawl_valid = 0; e->crawl_start = 0; e->crawl_ignored = 0; e->elec_cost = m->electricity_cost;
    if (m->flag_task) flag_reposition(m, e, seed, env_id, 1);
    double dist;
    calc_state_flag(m, e, &dist, seed, env_id);
    e->potential = task_potential(m, e, dist);
}

/* =========================== exported API =========================== */
static void pool_destroy(struct step_pool *p);
oracle_world *oracle_create(const rs_model_desc *m, int n_envs) {
    oracle_world *w = (oracle_world *)calloc(1, sizeof *w);
    w->m = *m; w->n_envs = n_envs; w->nman = nman_of(m);
    w->env = (oenv *)calloc((size_t)n_envs, sizeof(oenv));
    for (int i = 0; i < n_envs; i++) {
        w->env[i].man = (manifold *)calloc((size_t)(w->nman > 0 ? w->nman : 1), sizeof(manifold));
        w->env[i].base_quat[3] = 1; w->env[i].cube_quat[3] = 1;
        w->env[i].initial_z = m->initial_z;
        w->env[i].target_x = m->walk_target_x; w->env[i].target_y = m->walk_target_y;
    }
    return w;
}
void oracle_destroy(oracle_world *w) {
    pool_destroy(w->pool);
    for (int i = 0; i < w->n_envs; i++) free(w->env[i].man);
    free(w->env); free(w);
}
int oracle_state_dim(const oracle_world *w) { return (w->m.fixed_base ? 0 : 13) + 2 * w->m.n_dof + (w->m.has_cube ? 13 : 0); }
/* state layout: [pos3 quat4(xyzw) omega3 vel3] (floating base only) q[ndof] qd[ndof] */
void oracle_set_state(oracle_world *w, int env, const double *s, int clear_contacts) {
    oenv *e = &w->env[env]; const rs_model_desc *m = &w->m;
    if (!m->fixed_base) { v3cpy(e->base_pos, s); memcpy(e->base_quat, s + 3, 4 * sizeof(double)); v3cpy(e->base_w, s + 7); v3cpy(e->base_v, s + 10); s += 13; }
    for (int d = 0; d < m->n_dof; d++) { e->q[d] = s[d]; e->qd[d] = s[m->n_dof + d]; }
    if (m->has_cube) { const double *c = s + 2 * m->n_dof; v3cpy(e->cube_pos, c); memcpy(e->cube_quat, c + 3, 4 * sizeof(double)); v3cpy(e->cube_w, c + 7); v3cpy(e->cube_v, c + 10); }
    if (clear_contacts) for (int k = 0; k < w->nman; k++) e->man[k].n = 0;
    kinematics(m, e); velocities(m, e);
}
void oracle_get_state(const oracle_world *w, int env, double *s) {
    const oenv *e = &w->env[env]; const rs_model_desc *m = &w->m;
    if (!m->fixed_base) { v3cpy(s, e->base_pos); memcpy(s + 3, e->base_quat, 4 * sizeof(double)); v3cpy(s + 7, e->base_w); v3cpy(s + 10, e->base_v); s += 13; }
    for (int d = 0; d < m->n_dof; d++) { s[d] = e->q[d]; s[m->n_dof + d] = e->qd[d]; }
    if (m->has_cube) { double *c = s + 2 * m->n_dof; v3cpy(c, e->cube_pos); memcpy(c + 3, e->cube_quat, 4 * sizeof(double)); v3cpy(c + 7, e->cube_w); v3cpy(c + 10, e->cube_v); }
}
void oracle_set_episode(oracle_world *w, int env, double potential, double initial_z, const double *feet_contact, int frame) {
    oenv *e = &w->env[env];
    e->potential = potential; e->initial_z = initial_z; e->frame = frame;
    for (int k = 0; k < w->m.n_feet; k++) e->feet_contact[k] = feet_contact[k];
}
void oracle_get_episode(const oracle_world *w, int env, double *potential, double *initial_z, double *feet_contact, int *frame) {
    const oenv *e = &w->env[env];
    *potential = e->potential; *initial_z = e->initial_z; *frame = e->frame;
    for (int k = 0; k < w->m.n_feet; k++) feet_contact[k] = e->feet_contact[k];
}
/* Joint motor of one dof of one env (env < 0: every env).  max_impulse is per sub-step: the reference's control commands
 * set maxForce * physicsDeltaTime with physicsDeltaTime = timestep * frame_skip (physics-bullet.cpp:363,373); 0 = off. */
void oracle_set_joint_motor(oracle_world *w, int env, int dof, double kp, double kd, double target_pos, double target_vel, double max_impulse) {
    for (int k = (env < 0 ? 0 : env); k < (env < 0 ? w->n_envs : env + 1); k++) {
        oenv *e = &w->env[k];
        e->mot_kp[dof] = kp; e->mot_kd[dof] = kd; e->mot_pos[dof] = target_pos; e->mot_vel[dof] = target_vel; e->mot_imp[dof] = max_impulse;
    }
}
void oracle_set_flag(oracle_world *w, int env, double tx, double ty, int timeout, uint32_t draws) {
    oenv *e = &w->env[env];
    e->target_x = tx; e->target_y = ty; e->flag_timeout = timeout; e->flag_draws = draws;
}
void oracle_get_flag(const oracle_world *w, int env, double *tx, double *ty, int *timeout, uint32_t *draws) {
    const oenv *e = &w->env[env];
    *tx = e->target_x; *ty = e->target_y; *timeout = e->flag_timeout; *draws = e->flag_draws;
}
/* FlagrunHarder task state as 8 doubles: task, on_ground, crawl_valid, crawl_start, crawl_ignored, hist[0..2] */
void oracle_get_task_state(const oracle_world *w, int env, double *out) {
    const oenv *e = &w->env[env];
    out[0] = e->task; out[1] = e->on_ground; out[2] = e->crawl_valid; out[3] = e->crawl_start; out[4] = e->crawl_ignored;
    out[5] = e->hist[0]; out[6] = e->hist[1]; out[7] = e->hist[2];
}
void oracle_set_task_state(oracle_world *w, int env, const double *in) {
    oenv *e = &w->env[env];
    e->task = (int)in[0]; e->on_ground = (int)in[1]; e->crawl_valid = (int)in[2]; e->crawl_start = in[3]; e->crawl_ignored = in[4];
    e->hist[0] = (uint32_t)in[5]; e->hist[1] = (uint32_t)in[6]; e->hist[2] = (uint32_t)in[7];
}
void oracle_reset(oracle_world *w, int env, uint32_t seed, uint32_t env_id) { env_reset(&w->m, &w->env[env], w->nman, seed, env_id); }
void oracle_reset_all(oracle_world *w, uint32_t seed, uint32_t env_id0) {
    for (int i = 0; i < w->n_envs; i++) env_reset(&w->m, &w->env[i], w->nman, seed, env_id0 + (uint32_t)i);
}
/* raw torques (physics-only tests): tau[ndof] */
void oracle_set_tau(oracle_world *w, int env, const double *tau) { for (int d = 0; d < w->m.n_dof; d++) w->env[env].tau[d] = tau[d]; }
void oracle_substeps(oracle_world *w, int env, int n) {
    oenv *e = &w->env[env];
    for (int s = 0; s < n; s++) substep(&w->m, e);
    kinematics(&w->m, e); velocities(&w->m, e);
}
/* one full env step for all envs: apply_action -> frame_skip sub-steps -> calc_state/reward/done (-> auto reset) */
static void step_one(oracle_world *w, int i, const double *actions, int auto_reset, uint32_t seed, uint32_t env_id0) {
    const rs_model_desc *m = &w->m;
    oenv *e = &w->env[i];
    const double *a = actions + (size_t)i * m->n_act;
    for (int d = 0; d < m->n_dof; d++) e->tau[d] = 0;
    for (int k = 0; k < m->n_act; k++) {
        double c = a[k] > 1 ? 1 : (a[k] < -1 ? -1 : a[k]);
        e->tau[m->act_dof[k]] = (double)(float)(m->act_gain[k] * c);   /* set_motor_torque(float) */
    }
    /* applyJointDamping: once per stepSimulation, from the velocities at its start */
    if (K.damping_mode == 0) {
        for (int l = 0; l < m->n_links; l++) if (m->dof_idx[l] >= 0 && m->jdamping[l] != 0) e->tau[m->dof_idx[l]] += -m->jdamping[l] * e->qd[m->dof_idx[l]];
        for (int l = 0; l < m->n_links; l++) if (m->dof_idx[l] >= 0 && m->stiffness[l] != 0) e->tau[m->dof_idx[l]] += -m->stiffness[l] * e->q[m->dof_idx[l]];
    }
    for (int s = 0; s < m->frame_skip; s++) {
        substep(m, e);
        if (K.torque_first_only) for (int d = 0; d < m->n_dof; d++) e->tau[d] = 0;
    }
    epilogue(m, e, a, seed, env_id0 + (uint32_t)i);
    if (auto_reset && (e->done || e->frame >= m->max_episode_steps)) {
        /* obs returned for a finished env is the first obs of the next episode (vector-env convention) */
        double rew = e->reward;
        env_reset(m, e, w->nman, seed, env_id0 + (uint32_t)i);
        e->done = 1; e->reward = rew;
    }
}
/* Persistent worker pool: the threads are created once per world and parked on a condition variable between steps
 * (spawning and joining pthreads every step made the CPU baseline vary 5x between runs).  Envs are handed out in chunks
 * from an atomic counter, so a thread that drew cheap envs takes more of them; the calling thread works too. */
typedef struct step_pool {
    pthread_t th[256];
    int n_workers;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    unsigned long generation;
    int running, quit;
    oracle_world *w; const double *actions; int auto_reset; uint32_t seed, env_id0;
    int next, chunk;
} step_pool;
static void pool_drain(step_pool *p) {
    const int n = p->w->n_envs;
    for (;;) {
        int lo = __atomic_fetch_add(&p->next, p->chunk, __ATOMIC_RELAXED);
        if (lo >= n) break;
        int hi = lo + p->chunk < n ? lo + p->chunk : n;
        for (int i = lo; i < hi; i++) step_one(p->w, i, p->actions, p->auto_reset, p->seed, p->env_id0);
    }
}
static void *pool_worker(void *arg) {
    step_pool *p = (step_pool *)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->quit && p->generation == seen) pthread_cond_wait(&p->cv_start, &p->mu);
        if (p->quit) break;
        seen = p->generation;
        pthread_mutex_unlock(&p->mu);
        pool_drain(p);
        pthread_mutex_lock(&p->mu);
        if (--p->running == 0) pthread_cond_signal(&p->cv_done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}
static void pool_destroy(step_pool *p) {
    if (!p) return;
    pthread_mutex_lock(&p->mu); p->quit = 1; pthread_cond_broadcast(&p->cv_start); pthread_mutex_unlock(&p->mu);
    for (int t = 0; t < p->n_workers; t++) pthread_join(p->th[t], NULL);
    pthread_mutex_destroy(&p->mu); pthread_cond_destroy(&p->cv_start); pthread_cond_destroy(&p->cv_done);
    free(p);
}
static step_pool *pool_create(oracle_world *w, int n_workers) {
    step_pool *p = (step_pool *)calloc(1, sizeof *p);
    p->w = w; p->n_workers = n_workers;
    pthread_mutex_init(&p->mu, NULL); pthread_cond_init(&p->cv_start, NULL); pthread_cond_init(&p->cv_done, NULL);
    for (int t = 0; t < n_workers; t++) pthread_create(&p->th[t], NULL, pool_worker, p);
    return p;
}
void oracle_step(oracle_world *w, const double *actions, int auto_reset, uint32_t seed, uint32_t env_id0, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > w->n_envs) n_threads = w->n_envs;
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { for (int i = 0; i < w->n_envs; i++) step_one(w, i, actions, auto_reset, seed, env_id0); return; }
    if (w->pool && w->pool->n_workers != n_threads - 1) { pool_destroy(w->pool); w->pool = NULL; }
    if (!w->pool) w->pool = pool_create(w, n_threads - 1);
    step_pool *p = w->pool;
    pthread_mutex_lock(&p->mu);
    p->actions = actions; p->auto_reset = auto_reset; p->seed = seed; p->env_id0 = env_id0;
    p->next = 0;
    p->chunk = w->n_envs / (n_threads * 8); if (p->chunk < 1) p->chunk = 1;
    p->running = p->n_workers; p->generation++;
    pthread_cond_broadcast(&p->cv_start);
    pthread_mutex_unlock(&p->mu);
    pool_drain(p);
    pthread_mutex_lock(&p->mu);
    while (p->running > 0) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}
void oracle_get_obs(const oracle_world *w, double *obs, double *rew, int *done) {
    for (int i = 0; i < w->n_envs; i++) {
        for (int k = 0; k < w->m.obs_dim; k++) obs[(size_t)i * w->m.obs_dim + k] = w->env[i].obs[k];
        rew[i] = w->env[i].reward; done[i] = w->env[i].done;
    }
}
void oracle_get_rewards5(const oracle_world *w, int env, double *r5) { memcpy(r5, w->env[env].rewards5, 5 * sizeof(double)); }
/* (1+L) x {pos3, quat4, linvel3}: what query_body_position caches into Thingy */
void oracle_link_poses(oracle_world *w, int env, double *out) {
    oenv *e = &w->env[env];
    kinematics(&w->m, e); velocities(&w->m, e);
    for (int b = 0; b <= w->m.n_links; b++) body_pose(e, b, out + b * 10, out + b * 10 + 3, out + b * 10 + 7);
}
/* contact rows: per live manifold point {manifold_id, slot, dist, nx,ny,nz, posA xyz, imp_n, lifetime}; returns count */
int oracle_get_contacts(const oracle_world *w, int env, double *out, int max_pts) {
    const oenv *e = &w->env[env]; int c = 0;
    for (int k = 0; k < w->nman; k++)
        for (int p = 0; p < e->man[k].n && c < max_pts; p++) {
            const cpoint *cp = &e->man[k].p[p]; double *o = out + c * 12;
            o[0] = k; o[1] = p; o[2] = cp->dist; v3cpy(o + 3, cp->normal); v3cpy(o + 6, cp->posA); o[9] = cp->imp_n; o[10] = cp->lifetime; o[11] = cp->friction;
            c++;
        }
    return c;
}
/* persistent manifolds in solver order as the CUDA path stores them: keys[c] = manifold id, data[c][10] = localA3 localB3 normal3 dist */
int oracle_get_contact_cache(const oracle_world *w, int env, int *keys, double *data, int max_pts) {
    const oenv *e = &w->env[env]; int c = 0;
    for (int k = 0; k < w->nman; k++)
        for (int p = 0; p < e->man[k].n && c < max_pts; p++) {
            const cpoint *cp = &e->man[k].p[p]; double *o = data + c * 10;
            keys[c] = k; v3cpy(o, cp->localA); v3cpy(o + 3, cp->localB); v3cpy(o + 6, cp->normal); o[9] = cp->dist;
            c++;
        }
    return c;
}
/* btMultiBody::calcAccelerationDeltasMultiDof of the env's current configuration (after an ABA pass): out = M^-1 force,
 * both [base torque3 (world) base force3 (world) joint forces(ndof)] -- exported for the mechanics-identity tests */
void oracle_accel_deltas(oracle_world *w, int env, const double *force, double *out) {
    oenv *e = &w->env[env];
    oenv tmp = *e;                     /* the ABA pass fills the response cache (h, 1/D, base inverse); keep the state */
    kinematics(&w->m, &tmp);
    aba(&w->m, &tmp, 0.0);
    accel_deltas(&w->m, &tmp, force, out);
}
void oracle_counts(const oracle_world *w, int env, int *n_rows, int *n_contacts) { *n_rows = w->env[env].n_rows_last; *n_contacts = w->env[env].n_contacts_last; }
/* mechanical energy (kinetic + potential), for conservation tests */
double oracle_energy(oracle_world *w, int env) {
    const rs_model_desc *m = &w->m; oenv *e = &w->env[env];
    kinematics(m, e); velocities(m, e);
    double E = 0;
    for (int b = 0; b <= m->n_links; b++) {
        double mass = b == 0 ? m->base_mass : m->mass[b - 1];
        const double *I3 = b == 0 ? m->base_inertia : m->inertia[b - 1];
        if (b == 0 && m->fixed_base) continue;
        const sv6 *v = &e->vel[b];
        E += 0.5 * mass * v3dot(v->v, v->v) + 0.5 * (I3[0] * v->w[0] * v->w[0] + I3[1] * v->w[1] * v->w[1] + I3[2] * v->w[2] * v->w[2]);
        E += mass * m->gravity * e->pos[b][2];
    }
    return E;
}
void oracle_foot_flags(const oracle_world *w, int env, int *ground, int *other) {
    for (int f = 0; f < w->m.n_feet; f++) foot_contacts(&w->m, &w->env[env], f, &ground[f], &other[f]);
}
uint32_t oracle_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return rs_hash(a, b, c, d); }
