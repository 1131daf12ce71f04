// emul.cpp -- runs the SAME fp32 per-env code as the CUDA kernels (rs_core.cuh) on the CPU, one env
// at a time.  Test infrastructure for the GPU-less container: lets tests/ compare the kernel
// arithmetic with the double-precision oracle before spending GPU time.  Never used by the product.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../roboschool_b200/csrc/rs_devmodel.h"
#include "../../roboschool_b200/csrc/rs_pooled.cuh"

using namespace rs;
constexpr int NL = RS_MAX_LINKS, MC = 16;
typedef Work<NL, MC> W_t;

struct Emul {
    DevModel M;
    int n;
    std::vector<W_t> w;
    std::vector<Episode> ep;
    std::vector<float> ws, obs, rew;
    std::vector<int> done;
    std::vector<StepOut> out;
    int rec, maxr;
    int topo;                 // 0 generic, 1 humanoid, 2 ant, 3 hopper, 4 walker2d (static-topology sweeps)
    std::vector<float> lcbuf; // per-env link cache (shared memory on the GPU)
    int lcsize;
    std::vector<float> motor; // joint motors [n][5][n_dof], allocated by the first emul_set_joint_motor
};

template <class T>
static void step_env_static(Emul* e, int i, const float* a, int auto_reset, uint32_t seed, uint32_t env0) {
    const DevModel& M = e->M; W_t& W = e->w[i]; RowPool<pool_cap<T, MC, 1>()> pool{e->ws.data()};
    LinkCache<1> lc{e->lcbuf.data() + (size_t)i * e->lcsize};
    apply_action(M, W, a, 1);
    for (int s = 0; s < M.frame_skip; s++) substep_pooled<T>(WarpHost(), true, M, W, pool, lc, lc.p);
    float* obs = &e->obs[(size_t)i * M.obs_dim];
    auto cs = [&](float* o, int os, double* dist, float* pitch, int* al) { static_calc_state<T>(M, W, lc, e->ep[i], o, os, dist, pitch, al); };
    epilogue_with<T::CUBE != 0>(M, W, e->ep[i], a, 1, obs, 1, e->out[i], cs, seed, env0 + i);
    int done = e->out[i].done;
    if (auto_reset && (done || e->ep[i].frame >= M.max_episode_steps)) { env_reset_with(M, W, e->ep[i], seed, env0 + i, obs, 1, cs); done = 1; }
    e->rew[i] = e->out[i].reward; e->done[i] = done;
}
template <class T>
static void substeps_static(Emul* e, int env, int n) {
    RowPool<pool_cap<T, MC, 1>()> pool{e->ws.data()};
    LinkCache<1> lc{e->lcbuf.data() + (size_t)env * e->lcsize};
    for (int s = 0; s < n; s++) substep_pooled<T>(WarpHost(), true, e->M, e->w[env], pool, lc, lc.p);
}

extern "C" {
void* emul_create(const rs_model_desc* m, int n) {
    Emul* e = new Emul();
    const char* err = nullptr;
    if (build_dev_model(m, &e->M, &err)) { delete e; return nullptr; }
    e->n = n; e->w.resize(n); e->ep.resize(n); e->out.resize(n);
    for (int i = 0; i < n; i++) { memset(&e->w[i], 0, sizeof(W_t)); e->w[i].bquat[3] = 1; e->w[i].cquat[3] = 1; memset(&e->ep[i], 0, sizeof(Episode)); e->ep[i].initial_z = e->M.initial_z; e->ep[i].target_x = e->M.walk_target_x; e->ep[i].target_y = e->M.walk_target_y; }
    e->rec = 2 * (6 + e->M.n_dof + (e->M.has_cube ? 6 : 0)) + 1; e->maxr = 4 * NL + 4 * MC;   // rs_pooled.cuh pool_cap: rows + task directories
    e->ws.resize((size_t)e->rec * e->maxr);
    e->obs.resize((size_t)n * e->M.obs_dim); e->rew.resize(n); e->done.resize(n);
    e->topo = 0; e->lcsize = 9 * RS_MAX_LINKS + 30; e->lcbuf.assign((size_t)n * e->lcsize, 0.f);
    return e;
}
// select the static-topology sweeps if the descriptor matches a generated topology; returns the topology id
int emul_use_static(void* h, int on) {
    Emul* e = (Emul*)h;
    e->topo = 0;
    if (on) {
        if (topo_matches<TopoHumanoidCube>(e->M)) e->topo = 6;
        else if (topo_matches<TopoHumanoid>(e->M)) e->topo = 1;
        else if (topo_matches<TopoAnt>(e->M)) e->topo = 2;
        else if (topo_matches<TopoHopper>(e->M)) e->topo = 3;
        else if (topo_matches<TopoWalker2d>(e->M)) e->topo = 4;
        else if (topo_matches<TopoHalfCheetah>(e->M)) e->topo = 5;
    }
    return e->topo;
}
void emul_destroy(void* h) { delete (Emul*)h; }
int emul_state_dim(void* h) { Emul* e = (Emul*)h; return (e->M.fixed_base ? 0 : 13) + 2 * e->M.n_dof + (e->M.has_cube ? 13 : 0); }
void emul_set_state(void* h, int env, const double* s, int clear) {
    Emul* e = (Emul*)h; W_t& W = e->w[env];
    if (!e->M.fixed_base) { for (int k = 0; k < 3; k++) { W.bpos[k] = (float)s[k]; W.bw[k] = (float)s[7 + k]; W.bv[k] = (float)s[10 + k]; } for (int k = 0; k < 4; k++) W.bquat[k] = (float)s[3 + k]; s += 13; }
    for (int d = 0; d < e->M.n_dof; d++) { W.q[d] = (float)s[d]; W.qd[d] = (float)s[e->M.n_dof + d]; }
    if (e->M.has_cube) { const double* c = s + 2 * e->M.n_dof; for (int k = 0; k < 3; k++) { W.cpos[k] = (float)c[k]; W.cw[k] = (float)c[7 + k]; W.cv[k] = (float)c[10 + k]; } for (int k = 0; k < 4; k++) W.cquat[k] = (float)c[3 + k]; }
    if (clear) { W.ccount[0] = W.ccount[1] = 0; W.cur = 0; }
}
void emul_get_state(void* h, int env, double* s) {
    Emul* e = (Emul*)h; W_t& W = e->w[env];
    if (!e->M.fixed_base) { for (int k = 0; k < 3; k++) { s[k] = W.bpos[k]; s[7 + k] = W.bw[k]; s[10 + k] = W.bv[k]; } for (int k = 0; k < 4; k++) s[3 + k] = W.bquat[k]; s += 13; }
    for (int d = 0; d < e->M.n_dof; d++) { s[d] = W.q[d]; s[e->M.n_dof + d] = W.qd[d]; }
    if (e->M.has_cube) { double* c = s + 2 * e->M.n_dof; for (int k = 0; k < 3; k++) { c[k] = W.cpos[k]; c[7 + k] = W.cw[k]; c[10 + k] = W.cv[k]; } for (int k = 0; k < 4; k++) c[3 + k] = W.cquat[k]; }
}
void emul_set_tau(void* h, int env, const double* tau) { Emul* e = (Emul*)h; for (int d = 0; d < e->M.n_dof; d++) e->w[env].tau[d] = (float)tau[d]; }
// btMultiBodyJointMotor of one dof (rs_world_set_joint_motor); envs with motors stay on the static-topology sweeps, like the library
void emul_set_joint_motor(void* h, int env, int dof, double kp, double kd, double pos, double vel, double imp) {
    Emul* e = (Emul*)h; const int nd = e->M.n_dof;
    if (e->motor.empty()) {
        e->motor.assign((size_t)e->n * 5 * nd, 0.f);
        e->ws.resize((size_t)e->rec * e->maxr);
        for (int i = 0; i < e->n; i++) { e->w[i].motor = e->motor.data() + (size_t)i * 5 * nd; e->w[i].motor_stride = 1; }
    }
    const double v[5] = {kp, kd, pos, vel, imp > 0 ? imp : 0};
    for (int i = (env < 0 ? 0 : env); i < (env < 0 ? e->n : env + 1); i++)
        for (int f = 0; f < 5; f++) e->motor[(size_t)i * 5 * nd + f * nd + dof] = (float)v[f];
}
void emul_substeps(void* h, int env, int n) {
    Emul* e = (Emul*)h; RowWS ws{e->ws.data(), 1};
    switch (e->topo) {
        case 1: substeps_static<TopoHumanoid>(e, env, n); return;
        case 2: substeps_static<TopoAnt>(e, env, n); return;
        case 3: substeps_static<TopoHopper>(e, env, n); return;
        case 4: substeps_static<TopoWalker2d>(e, env, n); return;
        case 5: substeps_static<TopoHalfCheetah>(e, env, n); return;
        case 6: substeps_static<TopoHumanoidCube>(e, env, n); return;
    }
    for (int s = 0; s < n; s++) substep(e->M, e->w[env], ws);
}
void emul_reset_all(void* h, uint32_t seed, uint32_t env0) {
    Emul* e = (Emul*)h;
    for (int i = 0; i < e->n; i++) env_reset(e->M, e->w[i], e->ep[i], seed, env0 + i, &e->obs[(size_t)i * e->M.obs_dim], 1);
}
void emul_set_episode(void* h, int env, double potential, double initial_z, const double* fc, int frame) {
    Emul* e = (Emul*)h; Episode& E = e->ep[env];
    E.potential = potential; E.initial_z = (float)initial_z; E.frame = frame;
    for (int k = 0; k < e->M.n_feet; k++) E.feet_contact[k] = (float)fc[k];
}
void emul_set_flag(void* h, int env, double tx, double ty, int timeout, uint32_t draws) {
    Emul* e = (Emul*)h; Episode& E = e->ep[env]; E.target_x = tx; E.target_y = ty; E.flag_timeout = timeout; E.flag_draws = draws;
}
void emul_get_flag(void* h, int env, double* out) {
    Emul* e = (Emul*)h; Episode& E = e->ep[env]; out[0] = E.target_x; out[1] = E.target_y; out[2] = E.flag_timeout; out[3] = E.flag_draws;
}
void emul_set_task_state(void* h, int env, const double* s) {
    Emul* e = (Emul*)h; Episode& E = e->ep[env];
    E.task = (int)s[0]; E.on_ground = (int)s[1]; E.crawl_valid = (int)s[2]; E.crawl_start = s[3]; E.crawl_ignored = s[4];
    for (int k = 0; k < 3; k++) E.hist[k] = (uint32_t)s[5 + k];
}
void emul_get_task_state(void* h, int env, double* s) {
    Emul* e = (Emul*)h; Episode& E = e->ep[env];
    s[0] = E.task; s[1] = E.on_ground; s[2] = E.crawl_valid; s[3] = E.crawl_start; s[4] = E.crawl_ignored;
    for (int k = 0; k < 3; k++) s[5 + k] = E.hist[k];
}
void emul_step(void* h, const double* actions, int auto_reset, uint32_t seed, uint32_t env0) {
    Emul* e = (Emul*)h; RowWS ws{e->ws.data(), 1};
    const DevModel& M = e->M;
    for (int i = 0; i < e->n; i++) {
        float a[RS_MAX_ACT];
        for (int k = 0; k < M.n_act; k++) a[k] = (float)actions[(size_t)i * M.n_act + k];
        W_t& W = e->w[i];
        if (e->topo) {
            switch (e->topo) {
                case 1: step_env_static<TopoHumanoid>(e, i, a, auto_reset, seed, env0); break;
                case 2: step_env_static<TopoAnt>(e, i, a, auto_reset, seed, env0); break;
                case 3: step_env_static<TopoHopper>(e, i, a, auto_reset, seed, env0); break;
                case 4: step_env_static<TopoWalker2d>(e, i, a, auto_reset, seed, env0); break;
                case 6: step_env_static<TopoHumanoidCube>(e, i, a, auto_reset, seed, env0); break;
                default: step_env_static<TopoHalfCheetah>(e, i, a, auto_reset, seed, env0); break;
            }
            continue;
        }
        apply_action(M, W, a, 1);
        for (int s = 0; s < M.frame_skip; s++) substep(M, W, ws);
        float* obs = &e->obs[(size_t)i * M.obs_dim];
        epilogue(M, W, e->ep[i], a, 1, obs, 1, e->out[i], seed, env0 + i);
        int done = e->out[i].done;
        if (auto_reset && (done || e->ep[i].frame >= M.max_episode_steps)) { env_reset(M, W, e->ep[i], seed, env0 + i, obs, 1); done = 1; }
        e->rew[i] = e->out[i].reward; e->done[i] = done;
    }
}
void emul_get_obs(void* h, double* obs, double* rew, int* done) {
    Emul* e = (Emul*)h;
    for (size_t k = 0; k < e->obs.size(); k++) obs[k] = e->obs[k];
    for (int i = 0; i < e->n; i++) { rew[i] = e->rew[i]; done[i] = e->done[i]; }
}
void emul_get_rewards5(void* h, int env, double* r) { Emul* e = (Emul*)h; for (int k = 0; k < 5; k++) r[k] = e->out[env].rewards5[k]; }
int emul_get_contacts(void* h, int env, double* out, int max_pts) {
    Emul* e = (Emul*)h; W_t& W = e->w[env]; int c = 0, slot = 0, last = -1;
    for (int i = 0; i < W.ccount[W.cur] && c < max_pts; i++) {
        int key = W.ckey[W.cur][i];
        slot = key == last ? slot + 1 : 0; last = key;
        double* o = out + c * 12; const float* d = W.cdat[W.cur][i];
        o[0] = key; o[1] = slot; o[2] = d[9]; o[3] = d[6]; o[4] = d[7]; o[5] = d[8];
        const int ngp = e->M.n_geoms + e->M.n_pairs;
        float pA[3] = {0.f, 0.f, 0.f};
        if (key < ngp) to_world(W, key < e->M.n_geoms ? e->M.g_link[key] + 1 : e->M.g_link[e->M.pair_a[key - e->M.n_geoms]] + 1, d, pA);
        else if (key > ngp) to_world(W, e->M.g_link[key - ngp - 1] + 1, d, pA);
        o[6] = pA[0]; o[7] = pA[1]; o[8] = pA[2]; o[9] = 0; o[10] = 0; o[11] = 0;
        c++;
    }
    return c;
}
// overwrite one env's persistent contact manifolds (solver order), like rs_world_set_contacts
void emul_set_contacts(void* h, int env, int count, const int* keys, const double* data) {
    Emul* e = (Emul*)h; W_t& W = e->w[env];
    W.cur = 0; W.ccount[0] = count > MC ? MC : count; W.ccount[1] = 0;
    for (int i = 0; i < W.ccount[0]; i++) { W.ckey[0][i] = keys[i]; for (int k = 0; k < CDAT; k++) W.cdat[0][i][k] = (float)data[i * CDAT + k]; }
}
void emul_counts(void* h, int env, int* rows, int* contacts) { Emul* e = (Emul*)h; *rows = e->w[env].n_rows; *contacts = e->w[env].n_contacts; }
void emul_link_poses(void* h, int env, double* out) {
    Emul* e = (Emul*)h; W_t& W = e->w[env];
    kinematics(e->M, W); velocities(e->M, W);
    for (int b = 0; b <= e->M.n_links; b++) { float p[3], q[4], v[3]; body_pose(W, b, p, q, v); double* o = out + b * 10; for (int k = 0; k < 3; k++) { o[k] = p[k]; o[7 + k] = v[k]; } for (int k = 0; k < 4; k++) o[3 + k] = q[k]; }
}
}
