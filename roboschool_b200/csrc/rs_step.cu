// rs_step.cu -- kernels + C ABI (include/roboschool_b200.h) of the batched stepper.
//
// Data layout in HBM (all fp32 unless noted), SoA with the env index fastest so that the 32 lanes of
// a warp (= 32 consecutive envs) touch 128 contiguous bytes per field:
//   state   [S][N]            S = 13*(floating base) + 2*ndof          (read once, written once per env step)
//   ccount  [N] int32, ckey [MC][N] int32, cdat [MC][10][N]            (live contact points, persistent manifolds)
//   episode: potential [N] f64, initial_z [N], feet [F][N], frame [N] i32, episode [N] u32
//   rows    [MAXR*REC][N]     per-env PGS row workspace (J, M^-1 J^T, rhs, limits) -- scratch, L2 resident
// actions/obs are the caller's row-major [N][A] / [N][O] tensors.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/roboschool_b200.h"
#include "rs_devmodel.h"
#include "rs_kernels.cuh"

static thread_local std::string g_err;
static int fail(const std::string& s) { g_err = s; return 1; }
int rs_set_error(const char* s) { g_err = s; return 1; }   // shared with rs_mlp.cu
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" const char* rs_last_error(void) { return g_err.c_str(); }
extern "C" int rs_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// static-topology instantiations register themselves here (rs_static_inst.cu, one object file each)
static StaticLaunchFn g_static[8][2][17];   // [topology id][profiling][warps per CTA]
void rs_register_static(int topo, int prof, int wpc, StaticLaunchFn fn) {
    if (topo >= 0 && topo < 8 && wpc >= 0 && wpc < 17) g_static[topo][prof ? 1 : 0][wpc] = fn;
}
static SubstepsLaunchFn g_substeps[8];
void rs_register_substeps(int topo, SubstepsLaunchFn fn) { if (topo >= 0 && topo < 8) g_substeps[topo] = fn; }
// warps per CTA of the step kernel: RS_WPC (tuning knob) if that instantiation exists, else the per-topology default
static StaticLaunchFn pick_static(int topo, int prof, int def_wpc) {
    int wpc = def_wpc;
    if (const char* e = getenv("RS_WPC")) wpc = atoi(e);
    if (wpc < 0 || wpc > 16 || !g_static[topo][prof][wpc]) wpc = def_wpc;
    if (!g_static[topo][prof][wpc]) wpc = 1;
    return g_static[topo][prof][wpc];
}

#define RS_MAX_CHUNKS 8
#define RS_MAX_GRAPHS 4
struct rs_world {
    rs_model_desc desc;
    DevModel M;
    DevBuf B;
    int device, n, family;   // family: 0 (NL<=6), 1 (NL<=12), 2 (NL<=24)
    int wpc;                 // warps per CTA of the static-topology kernel
    int cluster, extra_sync; // cluster size / extra re-alignment points of the static-topology kernel (rs_kernels.cuh launch_static)
    int drop_rows;           // dead constraint rows dropped from the L2 after every sub-step (RowPool::discard)
    int topo;                // 0: runtime-topology kernel; 1..4: kernel specialised on a generated topology (rs_topo_gen.h)
    cudaStream_t stream;
    cudaStream_t chunk_stream[RS_MAX_CHUNKS];   // pipelined host entry point: one stream per env slice (created on first use)
    cudaEvent_t chunk_ev;
    cudaEvent_t chunk_done[RS_MAX_CHUNKS];      // join events of the slices (graph form of the pipelined entry point)
    cudaEvent_t caller_ev;   // recorded behind everything the world enqueues on a caller's stream
    int n_chunk_streams;
    // The pipelined host entry point enqueues ~40 operations per env step (6 slices x {wait, H2D, policy, step, 3 x D2H});
    // a caller that alternates between a few sets of page-locked buffers (ping-pong) replays them as ONE CUDA graph launch.
    // Key = everything the captured nodes hold by value: the caller's pointers, the policy, the slice count, the RNG keys.
    struct StepGraph { const float* in; float* out; float* rew; uint8_t* done; uint64_t pol; int n_chunks; uint32_t seed, env_id0; cudaGraphExec_t exec; };
    StepGraph graphs[RS_MAX_GRAPHS];
    int n_graphs, graph_next, graph_off;        // graph_off: capture failed once (or RS_E2E_GRAPH=0): direct launches from then on
    // pinned staging + device staging for the host-buffer entry point
    float *h_act, *h_obs, *h_rew; uint8_t* h_done;
    float *d_act, *d_obs, *d_rew; uint8_t* d_done;
    double* d_stats;
    uint32_t rollout_seed, env_id0;
};

// One env step per thread.  actions == nullptr: uniform random actions from the counter-based RNG.
template <int NL, int MC>
__global__ void __launch_bounds__(64) k_step(const __grid_constant__ DevModel M, const DevBuf B, const float* __restrict__ actions,
                                             float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
                                             int auto_reset, uint32_t seed, uint32_t env_id0, double* __restrict__ stats) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double st_rew = 0, st_ep = 0, st_len = 0, st_nf = 0, st_rows = 0, st_con = 0, st_steps = 0;
    if (e < B.N) {
        Work<NL, MC> W;
        Episode E;
        load_env(M, B, e, W, E);
        float a[MA];
        if (actions) { for (int k = 0; k < M.n_act; k++) a[k] = actions[(size_t)e * M.n_act + k]; }
        else { for (int k = 0; k < M.n_act; k++) a[k] = 2.f * rs_uniform(seed ^ 0xA5A5A5A5u, env_id0 + e, E.episode * 1024u + (uint32_t)E.frame, (uint32_t)k) - 1.f; }
        apply_action(M, W, a, 1);
        RowWS ws{B.rows + e, B.N};
        for (int s = 0; s < M.frame_skip; s++) substep(M, W, ws);
        float* o = obs + (size_t)e * M.obs_dim;
        StepOut out;
        epilogue(M, W, E, a, 1, o, 1, out, seed, env_id0 + e);
        int dn = out.done;
        bool finite = true;
        for (int k = 0; k < M.obs_dim; k++) finite = finite && isfinite(o[k]);
        int finished = dn || E.frame >= M.max_episode_steps;
        st_rew = out.reward; st_rows = W.n_rows; st_con = W.n_contacts; st_steps = 1; st_nf = finite ? 0 : 1;
        if (finished) { st_ep = 1; st_len = E.frame; }
        if (auto_reset && finished) { env_reset(M, W, E, seed, env_id0 + e, o, 1); dn = 1; }
        rew[e] = out.reward; done[e] = (uint8_t)dn;
        if (B.rewards5) {
#pragma unroll
            for (int k = 0; k < 5; k++) B.rewards5[(size_t)e * 5 + k] = out.rewards5[k];
        }
        store_env(M, B, e, W, E);
    }
    if (stats) {
        st_rew = warp_sum(st_rew); st_ep = warp_sum(st_ep); st_len = warp_sum(st_len); st_nf = warp_sum(st_nf);
        st_rows = warp_sum(st_rows); st_con = warp_sum(st_con); st_steps = warp_sum(st_steps);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(stats + 0, st_rew); atomicAdd(stats + 1, st_ep); atomicAdd(stats + 2, st_len); atomicAdd(stats + 3, st_nf);
            atomicAdd(stats + 4, st_rows); atomicAdd(stats + 5, st_con); atomicAdd(stats + 6, st_steps);
        }
    }
}

// Profiling twin of k_step: same work, plus clock64() deltas per phase summed over threads into cyc[8].
template <int NL, int MC>
__global__ void __launch_bounds__(64) k_step_prof(const __grid_constant__ DevModel M, const DevBuf B, const float* __restrict__ actions,
                                                  float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
                                                  uint32_t seed, uint32_t env_id0, unsigned long long* __restrict__ cyc) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.N) return;
    Work<NL, MC> W;
    Episode E;
    long long t0 = clock64();
    load_env(M, B, e, W, E);
    float a[MA];
    for (int k = 0; k < M.n_act; k++) a[k] = actions[(size_t)e * M.n_act + k];
    apply_action(M, W, a, 1);
    RowWS ws{B.rows + e, B.N};
    PhaseClock pc;
    for (int k = 0; k < 6; k++) pc.t[k] = 0;
    long long t1 = clock64();
    for (int s = 0; s < M.frame_skip; s++) substep(M, W, ws, &pc);
    long long t2 = clock64();
    float* o = obs + (size_t)e * M.obs_dim;
    StepOut out;
    epilogue(M, W, E, a, 1, o, 1, out, seed, env_id0 + e);
    int dn = out.done;
    if (dn || E.frame >= M.max_episode_steps) { env_reset(M, W, E, seed, env_id0 + e, o, 1); dn = 1; }
    rew[e] = out.reward; done[e] = (uint8_t)dn;
    store_env(M, B, e, W, E);
    long long t3 = clock64();
    if ((threadIdx.x & 31) == 0) {   // one lane per warp: warp-level cycles
        for (int k = 0; k < 5; k++) atomicAdd(cyc + k, (unsigned long long)pc.t[k]);
        atomicAdd(cyc + 5, (unsigned long long)(t3 - t2));
        atomicAdd(cyc + 6, (unsigned long long)(t1 - t0));
        atomicAdd(cyc + 7, (unsigned long long)(t3 - t0));
    }
}

template <int NL, int MC>
__global__ void __launch_bounds__(64) k_reset(const __grid_constant__ DevModel M, const DevBuf B, const uint8_t* __restrict__ mask,
                                              float* __restrict__ obs, uint32_t seed, uint32_t env_id0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.N) return;
    if (mask && !mask[e]) return;
    Work<NL, MC> W;
    Episode E;
    load_env(M, B, e, W, E);
    float tmp[8 + 2 * MA + MF];
    env_reset(M, W, E, seed, env_id0 + e, tmp, 1);
    if (obs) for (int k = 0; k < M.obs_dim; k++) obs[(size_t)e * M.obs_dim + k] = tmp[k];
    store_env(M, B, e, W, E);
}

template <int NL, int MC>
__global__ void __launch_bounds__(64) k_substeps(const __grid_constant__ DevModel M, const DevBuf B, int nsub) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.N) return;
    Work<NL, MC> W;
    Episode E;
    load_env(M, B, e, W, E);
    for (int d = 0; d < M.n_dof; d++) W.tau[d] = B.tau[(size_t)d * B.N + e];
    RowWS ws{B.rows + e, B.N};
    for (int s = 0; s < nsub; s++) substep(M, W, ws);
    store_env(M, B, e, W, E);
}

template <int NL, int MC>
__global__ void __launch_bounds__(64) k_link_poses(const __grid_constant__ DevModel M, const DevBuf B, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B.N) return;
    Work<NL, MC> W;
    Episode E;
    load_env(M, B, e, W, E);
    kinematics(M, W); velocities(M, W);
    for (int b = 0; b <= M.n_links; b++) {
        float p[3], q[4], v[3];
        body_pose(W, b, p, q, v);
        float* o = out + ((size_t)e * (M.n_links + 1) + b) * 10;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = q[0]; o[4] = q[1]; o[5] = q[2]; o[6] = q[3]; o[7] = v[0]; o[8] = v[1]; o[9] = v[2];
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
#define FAMILY_DISPATCH(w, CALL)                                   \
    switch ((w)->family) {                                         \
        case 0: { constexpr int NL = 6, MC = 8; CALL; } break;     \
        case 1: { constexpr int NL = 12, MC = 16; CALL; } break;   \
        default: { constexpr int NL = 20, MC = 16; CALL; } break;  \
    }

static int family_of(const rs_model_desc* m, int* NL, int* MC) {
    if (m->n_links <= 6 && m->max_contacts <= 8) { *NL = 6; *MC = 8; return 0; }
    if (m->n_links <= 12 && m->max_contacts <= 16) { *NL = 12; *MC = 16; return 1; }
    if (m->n_links <= 20 && m->max_contacts <= 16) { *NL = 20; *MC = 16; return 2; }
    return -1;
}

extern "C" int rs_world_create(const rs_model_desc* model, int n_envs, int device, rs_world** out) {
    if (!model || !out || n_envs <= 0) return fail("rs_world_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("rs_world_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("rs_world_create: device index out of range");
    CK(cudaSetDevice(device));
    rs_world* w = new rs_world();
    memset(w, 0, sizeof *w);
    w->desc = *model;
    const char* err = nullptr;
    if (build_dev_model(model, &w->M, &err)) { delete w; return fail(std::string("rs_world_create: ") + err); }
    int NL, MC;
    w->family = family_of(model, &NL, &MC);
    if (w->family < 0) { delete w; return fail("rs_world_create: model exceeds kernel capacity (links<=20, contacts<=16)"); }
    w->topo = 0;     // 0: the model is none of the generated topologies -> runtime-topology kernel (models loaded through load_mjcf)
    {
        if (topo_matches<TopoHumanoidCube>(w->M) && model->max_contacts <= 16) w->topo = 6;
        else if (topo_matches<TopoHumanoid>(w->M) && model->max_contacts <= 16) w->topo = 1;
        else if (topo_matches<TopoAnt>(w->M) && model->max_contacts <= 16) w->topo = 2;
        else if (topo_matches<TopoHopper>(w->M) && model->max_contacts <= 8) w->topo = 3;
        else if (topo_matches<TopoWalker2d>(w->M) && model->max_contacts <= 16) w->topo = 4;
        else if (topo_matches<TopoHalfCheetah>(w->M) && model->max_contacts <= 16) w->topo = 5;
    }
    if (model->has_cube && !w->topo) { delete w; return fail("rs_world_create: a world with the cube needs its static-topology kernel (model does not match TopoHumanoidCube)"); }
    w->device = device; w->n = n_envs;
    // warps per CTA (measured on B200): 7-warp CTAs that re-align at the phase barriers once the batch
    // gives every SM several warps (Humanoid 32768: +13 %, Ant 65536: +16 %, HalfCheetah 65536: +67 %); single-warp CTAs for
    // small batches, which would otherwise occupy only a few SMs
    // measured crossover of 7-warp aligned CTAs over 1-warp CTAs (a 7-warp CTA occupies a whole SM: below ~148 CTAs part of the GPU
    // idles): 24576 envs for Humanoid, Ant and the planar models (16384-20480: 1-warp CTAs +5..25 %), 8192 for the Flagrun tasks
    w->wpc = n_envs >= (model->flag_task ? 8192 : 24576) ? 7 : 1;
    if (w->topo && !model->flag_task) {
        // One wave of aligned CTAs: the narrowest instantiation whose CTAs (one per SM -- 255 registers fill it) hold the whole
        // batch, so that the warps of an SM walk the instruction stream together and no SM runs a second wave.  Measured
        // (profiles/r02_v4_cta_width.txt): Humanoid 16384 envs 1.15 ms (1-warp CTAs) -> 0.65 (4-warp, TPC pairs), 8192 envs 0.75 ->
        // 0.58 (2-warp); Ant 16384 0.58 -> 0.44, 8192 0.39 -> 0.36; Ant 65536 in 14-warp CTAs at 128 registers 1.48 -> 1.35 ms (at
        // 32768 envs they lose to 7-warp CTAs, 0.85 vs 0.70 ms: spills); Humanoid 37888 envs in 8-warp CTAs 0.92 ms against 1.49 ms
        // in two waves (9-warp CTAs spill 2.2 KB at 168 registers and lose).  Batches beyond one wave of the widest instantiation
        // run as several waves of 7-warp CTAs.
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        const int need = (n_envs + 32 * n_sm - 1) / (32 * n_sm);
        int pick = 0;
        for (int k = need; k <= 16 && !pick; k++) if (g_static[w->topo][0][k]) pick = k;
        w->wpc = pick ? pick : 7;
    }
    // TPC-pair clusters pay on the Humanoid-family kernels (largest instruction stream); measured neutral to -2 % on the Ant.
    // RS_CLUSTER / RS_EXTRA_SYNC override (tuning knobs, like RS_WPC)
    w->cluster = (w->topo == 1 || w->topo == 6) && w->wpc > 1 ? 2 : 0;
    w->extra_sync = model->flag_task ? 1 : 0;      // Flagrun 32768: 1.23 -> 1.19 ms, FlagrunHarder 3.29 -> 2.83 ms; Humanoid 0.82 -> 0.85 ms (off)
    if (const char* e = getenv("RS_CLUSTER")) w->cluster = atoi(e);
    if (const char* e = getenv("RS_EXTRA_SYNC")) w->extra_sync = atoi(e) != 0;
    // measured: Humanoid 32768 0.85 -> 0.81 ms, Walker2d 65536 0.99 -> 0.95 ms, Hopper 4096 (everything fits the L2) 0.103 -> 0.106 ms
    w->drop_rows = n_envs >= 16384 ? 1 : 0;
    if (const char* e = getenv("RS_DROP_ROWS")) w->drop_rows = atoi(e) != 0;
    DevBuf& B = w->B;
    B.N = n_envs; B.S = (model->fixed_base ? 0 : 13) + 2 * model->n_dof + (model->has_cube ? 13 : 0); B.MCcap = MC;
    B.rec = 2 * (6 + model->n_dof + (model->has_cube ? 6 : 0)) + 1; B.maxr = 4 * NL + 4 * MC;   // +1: pooled rows keep their denominator in the record; slots per env: a limit and a motor row per joint, 3 rows per contact, the two task directories (rs_pooled.cuh pool_cap)
    size_t N = (size_t)n_envs;
    CK(cudaMalloc(&B.state, sizeof(float) * N * (B.S > 0 ? B.S : 1)));
    CK(cudaMalloc(&B.ccount, sizeof(int) * N));
    CK(cudaMalloc(&B.ckey, sizeof(int) * N * MC));
    CK(cudaMalloc(&B.cdat, sizeof(float) * N * MC * CDAT));
    CK(cudaMalloc(&B.potential, sizeof(double) * N));
    CK(cudaMalloc(&B.initial_z, sizeof(float) * N));
    CK(cudaMalloc(&B.feet, sizeof(float) * N * MF));
    CK(cudaMalloc(&B.frame, sizeof(int) * N));
    CK(cudaMalloc(&B.episode, sizeof(uint32_t) * N));
    CK(cudaMalloc(&B.target, sizeof(double) * N * 2));
    CK(cudaMalloc(&B.flag_timeout, sizeof(int) * N));
    CK(cudaMalloc(&B.flag_draws, sizeof(uint32_t) * N));
    CK(cudaMemset(B.target, 0, sizeof(double) * N * 2));
    CK(cudaMemset(B.flag_timeout, 0, sizeof(int) * N));
    CK(cudaMemset(B.flag_draws, 0, sizeof(uint32_t) * N));
    CK(cudaMalloc(&B.hx_i, sizeof(int) * N * 7));
    CK(cudaMalloc(&B.hx_d, sizeof(double) * N * 2));
    CK(cudaMemset(B.hx_i, 0, sizeof(int) * N * 7));
    CK(cudaMemset(B.hx_d, 0, sizeof(double) * N * 2));
    CK(cudaMalloc(&B.rows, sizeof(float) * (size_t)((n_envs + 511) / 512 * 512 + 512) * B.rec * B.maxr));   // one pool per warp of 32 envs, incl. the spare warps of the last CTA (up to 16 warps per CTA)
    CK(cudaMalloc(&B.tau, sizeof(float) * N * (model->n_dof > 0 ? model->n_dof : 1)));
    CK(cudaMalloc(&B.rewards5, sizeof(float) * N * 5));
    CK(cudaMalloc(&B.nrows, sizeof(int) * N));
    CK(cudaMalloc(&B.ncontacts, sizeof(int) * N));
    CK(cudaMemset(B.state, 0, sizeof(float) * N * (B.S > 0 ? B.S : 1)));
    CK(cudaMemset(B.ccount, 0, sizeof(int) * N));
    CK(cudaMemset(B.potential, 0, sizeof(double) * N));
    CK(cudaMemset(B.feet, 0, sizeof(float) * N * MF));
    CK(cudaMemset(B.frame, 0, sizeof(int) * N));
    CK(cudaMemset(B.episode, 0, sizeof(uint32_t) * N));
    CK(cudaMemset(B.tau, 0, sizeof(float) * N * (model->n_dof > 0 ? model->n_dof : 1)));
    CK(cudaMemset(B.rewards5, 0, sizeof(float) * N * 5));
    CK(cudaMemset(B.nrows, 0, sizeof(int) * N));
    CK(cudaMemset(B.ncontacts, 0, sizeof(int) * N));
    {   // identity base quaternion, initial_z from the model
        std::vector<float> s((size_t)N * (B.S > 0 ? B.S : 1), 0.f), iz(N, (float)model->initial_z);
        if (!model->fixed_base) for (size_t e = 0; e < N; e++) s[6 * N + e] = 1.f;
        if (model->has_cube) for (size_t e = 0; e < N; e++) s[(size_t)(B.S - 13 + 6) * N + e] = 1.f;
        CK(cudaMemcpy(B.state, s.data(), sizeof(float) * s.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B.initial_z, iz.data(), sizeof(float) * N, cudaMemcpyHostToDevice));
    }
    CK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&w->caller_ev, cudaEventDisableTiming));
    int A = model->n_act > 0 ? model->n_act : 1, O = model->obs_dim > 0 ? model->obs_dim : 1;
    CK(cudaMallocHost(&w->h_act, sizeof(float) * N * A));
    CK(cudaMallocHost(&w->h_obs, sizeof(float) * N * O));
    CK(cudaMallocHost(&w->h_rew, sizeof(float) * N));
    CK(cudaMallocHost(&w->h_done, N));
    CK(cudaMalloc(&w->d_act, sizeof(float) * N * A));
    CK(cudaMalloc(&w->d_obs, sizeof(float) * N * O));
    CK(cudaMalloc(&w->d_rew, sizeof(float) * N));
    CK(cudaMalloc(&w->d_done, N));
    CK(cudaMalloc(&w->d_stats, sizeof(double) * 8));
    *out = w;
    return 0;
}

// the captured step graphs hold DevModel / DevBuf by value: dropped whenever either changes
static void drop_graphs(rs_world* w) {
    for (int k = 0; k < w->n_graphs; k++) { if (w->graphs[k].exec) cudaGraphExecDestroy(w->graphs[k].exec); w->graphs[k].exec = nullptr; }
    w->n_graphs = 0; w->graph_next = 0;
}

extern "C" int rs_world_destroy(rs_world* w) {
    if (!w) return 0;
    cudaSetDevice(w->device);
    DevBuf& B = w->B;
    cudaFree(B.state); cudaFree(B.ccount); cudaFree(B.ckey); cudaFree(B.cdat); cudaFree(B.potential); cudaFree(B.initial_z);
    cudaFree(B.feet); cudaFree(B.frame); cudaFree(B.episode); cudaFree(B.rows); cudaFree(B.tau); cudaFree(B.rewards5);
    cudaFree(B.nrows); cudaFree(B.ncontacts); cudaFree(B.target); cudaFree(B.flag_timeout); cudaFree(B.flag_draws); cudaFree(B.hx_i); cudaFree(B.hx_d); cudaFree(B.motor);
    cudaFreeHost(w->h_act); cudaFreeHost(w->h_obs); cudaFreeHost(w->h_rew); cudaFreeHost(w->h_done);
    cudaFree(w->d_act); cudaFree(w->d_obs); cudaFree(w->d_rew); cudaFree(w->d_done); cudaFree(w->d_stats);
    cudaStreamDestroy(w->stream);
    cudaEventDestroy(w->caller_ev);
    drop_graphs(w);
    for (int k = 0; k < w->n_chunk_streams; k++) { cudaStreamDestroy(w->chunk_stream[k]); cudaEventDestroy(w->chunk_done[k]); }
    if (w->n_chunk_streams) cudaEventDestroy(w->chunk_ev);
    delete w;
    return 0;
}

extern "C" int rs_world_n_envs(const rs_world* w) { return w->n; }
extern "C" int rs_world_kernel_variant(const rs_world* w) { return w->topo; }
extern "C" int rs_world_cta_warps(const rs_world* w) {
    if (!w->topo) return 2;                                    // runtime-topology kernels: 64-thread CTAs
    int per = 0;
    StaticLaunch q{&w->M, &w->B, w->n, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, nullptr, nullptr, nullptr};
    q.per_out = &per;
    StaticLaunchFn fn = pick_static(w->topo, 0, w->wpc);
    if (fn) fn(q);
    return per / 32;
}
extern "C" int rs_world_state_dim(const rs_world* w) { return w->B.S; }
extern "C" int rs_world_obs_dim(const rs_world* w) { return w->desc.obs_dim; }
extern "C" int rs_world_act_dim(const rs_world* w) { return w->desc.n_act; }
extern "C" int rs_world_state_dev(rs_world* w, float** p) { *p = w->B.state; return 0; }

static inline int nblocks(int n) { return (n + 63) / 64; }

// Ordering between caller streams and the world's private stream: every entry point that enqueues on a caller's
// stream records caller_ev behind its work; every entry point that works on the private stream first makes that stream wait
// for the event.  (The private stream is synchronised before those entry points return, so the other direction is ordered.)
static inline cudaError_t mark_caller(rs_world* w, cudaStream_t st) { return cudaEventRecord(w->caller_ev, st); }
static inline cudaError_t wait_caller(rs_world* w) { return cudaStreamWaitEvent(w->stream, w->caller_ev, 0); }
// host getters / setters: everything enqueued so far (caller streams through the event, then the private stream) has finished
static inline cudaError_t quiesce(rs_world* w) {
    cudaError_t e = wait_caller(w);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(w->stream);
}

extern "C" int rs_world_reset(rs_world* w, uint32_t seed, uint32_t env_id0, const uint8_t* mask_dev, float* obs_dev, void* stream) {
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    FAMILY_DISPATCH(w, (k_reset<NL, MC><<<nblocks(w->n), 64, 0, st>>>(w->M, w->B, mask_dev, obs_dev, seed, env_id0)));
    CK(cudaGetLastError());
    CK(mark_caller(w, st));
    w->rollout_seed = seed; w->env_id0 = env_id0;
    return 0;
}

static int launch_step(rs_world* w, const float* actions, float* obs, float* rew, uint8_t* done, int auto_reset, uint32_t seed,
                       uint32_t env_id0, double* stats, cudaStream_t st, int cta0 = 0, int ncta = 0) {
    if (w->topo) {
        StaticLaunch a{&w->M, &w->B, w->n, actions, obs, rew, done, auto_reset, seed, env_id0, stats, nullptr, st};
        a.cta0 = cta0; a.ncta = ncta; a.cluster = w->cluster; a.extra_sync = w->extra_sync; a.drop_rows = w->drop_rows;
        StaticLaunchFn fn = pick_static(w->topo, 0, w->wpc);
        if (!fn) return fail("static-topology kernel not linked");
        fn(a);
        CK(cudaGetLastError());
        return 0;
    }
    FAMILY_DISPATCH(w, (k_step<NL, MC><<<nblocks(w->n), 64, 0, st>>>(w->M, w->B, actions, obs, rew, done, auto_reset, seed, env_id0, stats)));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int rs_world_step(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                             int auto_reset, void* stream) {
    if (!actions_dev || !obs_dev || !reward_dev || !done_dev) return fail("rs_world_step: null buffer");
    CK(cudaSetDevice(w->device));
    if (launch_step(w, actions_dev, obs_dev, reward_dev, done_dev, auto_reset, w->rollout_seed, w->env_id0, nullptr, (cudaStream_t)stream)) return 1;
    CK(mark_caller(w, (cudaStream_t)stream));
    return 0;
}

extern "C" int rs_world_step_stats(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                                   int auto_reset, double* stats_dev, void* stream) {
    if (!actions_dev || !obs_dev || !reward_dev || !done_dev) return fail("rs_world_step_stats: null buffer");
    CK(cudaSetDevice(w->device));
    if (launch_step(w, actions_dev, obs_dev, reward_dev, done_dev, auto_reset, w->rollout_seed, w->env_id0, stats_dev, (cudaStream_t)stream)) return 1;
    CK(mark_caller(w, (cudaStream_t)stream));
    return 0;
}

// profiling: one step with per-phase warp-cycle totals -> cycles_host[16]
// [kinematics, collide, aba, solve, integrate, epilogue+store, load+apply_action, total,
//  inside solve (static-topology kernels): enumerate rows, limit-row tasks, contact tasks, right-hand sides, PGS, 0, 0, 0]
extern "C" int rs_world_step_profile(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                                     unsigned long long* cycles_host, void* stream) {
    CK(cudaSetDevice(w->device));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 16 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d, 0, 16 * sizeof(unsigned long long), (cudaStream_t)stream));
    cudaStream_t st = (cudaStream_t)stream;
    StaticLaunch sl{&w->M, &w->B, w->n, actions_dev, obs_dev, reward_dev, done_dev, 1, w->rollout_seed, w->env_id0, nullptr, d, st};
    switch (w->topo) {
        case 1: case 2: case 3: case 4: case 5: case 6: {
            StaticLaunchFn fn = pick_static(w->topo, 1, 1);
            if (!fn) { cudaFree(d); return fail("profiling kernel not linked"); }
            fn(sl);
        } break;
        default:
            FAMILY_DISPATCH(w, (k_step_prof<NL, MC><<<nblocks(w->n), 64, 0, st>>>(w->M, w->B, actions_dev, obs_dev, reward_dev, done_dev,
                                                                                  w->rollout_seed, w->env_id0, d)));
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(std::string("rs_world_step_profile: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int rs_world_step_host(rs_world* w, const float* actions_host, float* obs_host, float* reward_host, uint8_t* done_host,
                                  int auto_reset) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n, A = (size_t)w->desc.n_act, O = (size_t)w->desc.obs_dim;
    memcpy(w->h_act, actions_host, sizeof(float) * N * A);
    CK(wait_caller(w));
    CK(cudaMemcpyAsync(w->d_act, w->h_act, sizeof(float) * N * A, cudaMemcpyHostToDevice, w->stream));
    if (launch_step(w, w->d_act, w->d_obs, w->d_rew, w->d_done, auto_reset, w->rollout_seed, w->env_id0, nullptr, w->stream)) return 1;
    CK(cudaMemcpyAsync(w->h_obs, w->d_obs, sizeof(float) * N * O, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(w->h_rew, w->d_rew, sizeof(float) * N, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(w->h_done, w->d_done, N, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    memcpy(obs_host, w->h_obs, sizeof(float) * N * O);
    memcpy(reward_host, w->h_rew, sizeof(float) * N);
    memcpy(done_host, w->h_done, N);
    return 0;
}

// Joint::set_servo_target / set_target_speed (physics-bullet.cpp:510-537) -> btMultiBodyJointMotor of one dof.
extern "C" int rs_world_set_joint_motor(rs_world* w, int env, int dof, float kp, float kd, float target_pos, float target_vel, float max_impulse) {
    if (!w) return fail("rs_world_set_joint_motor: null world");
    const int nd = w->desc.n_dof, N = w->n;
    if (dof < 0 || dof >= nd || env >= N) return fail("rs_world_set_joint_motor: dof or env out of range");
    CK(cudaSetDevice(w->device));
    DevBuf& B = w->B;
    CK(quiesce(w));
    if (!B.motor) {
        if (!(max_impulse > 0.f)) return 0;                  // switching off a motor of a world that never had one
        // first motor of this world: motor table (all off); the row workspace already has room for one motor row per dof
        CK(cudaMalloc(&B.motor, sizeof(float) * 5 * (size_t)nd * N));
        CK(cudaMemset(B.motor, 0, sizeof(float) * 5 * (size_t)nd * N));
        drop_graphs(w);                                      // DevBuf changed
    }
    const float v[5] = {kp, kd, target_pos, target_vel, max_impulse > 0.f ? max_impulse : 0.f};
    const int e0 = env < 0 ? 0 : env, ne = env < 0 ? N : 1;
    std::vector<float> row((size_t)ne);
    for (int f = 0; f < 5; f++) {
        for (int k = 0; k < ne; k++) row[k] = v[f];
        CK(cudaMemcpy(B.motor + ((size_t)f * nd + dof) * N + e0, row.data(), sizeof(float) * ne, cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" int rs_world_substeps(rs_world* w, int n_substeps, const float* tau_host) {
    if (w->desc.has_cube && !(w->topo > 0 && w->topo < 8 && g_substeps[w->topo])) return fail("rs_world_substeps: the physics-only kernel of this cube topology is not linked");
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n; int nd = w->desc.n_dof;
    CK(quiesce(w));
    if (tau_host && nd > 0) {
        std::vector<float> t(N * nd);
        for (size_t e = 0; e < N; e++) for (int d = 0; d < nd; d++) t[(size_t)d * N + e] = tau_host[e * nd + d];
        CK(cudaMemcpy(w->B.tau, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
    }
    if (w->desc.has_cube) {
        // the cube is stepped by the static-topology code only (rs_kernels.cuh k_substeps_static)
        SubstepsLaunch a{&w->M, &w->B, w->n, n_substeps, w->stream};
        g_substeps[w->topo](a);
    } else {
        FAMILY_DISPATCH(w, (k_substeps<NL, MC><<<nblocks(w->n), 64, 0, w->stream>>>(w->M, w->B, n_substeps)));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

extern "C" int rs_world_set_state(rs_world* w, const float* state_host, int clear_contacts) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n; int S = w->B.S;
    std::vector<float> s(N * S);
    for (size_t e = 0; e < N; e++) for (int f = 0; f < S; f++) s[(size_t)f * N + e] = state_host[e * S + f];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(w->B.state, s.data(), sizeof(float) * s.size(), cudaMemcpyHostToDevice));
    if (clear_contacts) CK(cudaMemset(w->B.ccount, 0, sizeof(int) * N));
    return 0;
}

extern "C" int rs_world_get_state(rs_world* w, float* state_host) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n; int S = w->B.S;
    std::vector<float> s(N * S);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(s.data(), w->B.state, sizeof(float) * s.size(), cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < N; e++) for (int f = 0; f < S; f++) state_host[e * S + f] = s[(size_t)f * N + e];
    return 0;
}

extern "C" int rs_world_set_episode(rs_world* w, const double* potential, const float* initial_z, const float* feet_contact, const int* frame) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n; int F = w->desc.n_feet;
    CK(cudaDeviceSynchronize());
    if (potential) CK(cudaMemcpy(w->B.potential, potential, sizeof(double) * N, cudaMemcpyHostToDevice));
    if (initial_z) CK(cudaMemcpy(w->B.initial_z, initial_z, sizeof(float) * N, cudaMemcpyHostToDevice));
    if (frame) CK(cudaMemcpy(w->B.frame, frame, sizeof(int) * N, cudaMemcpyHostToDevice));
    if (feet_contact && F > 0) {
        std::vector<float> f(N * F);
        for (size_t e = 0; e < N; e++) for (int k = 0; k < F; k++) f[(size_t)k * N + e] = feet_contact[e * F + k];
        CK(cudaMemcpy(w->B.feet, f.data(), sizeof(float) * f.size(), cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" int rs_world_get_episode(rs_world* w, double* potential, float* initial_z, float* feet_contact, int* frame) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n; int F = w->desc.n_feet;
    CK(cudaDeviceSynchronize());
    if (potential) CK(cudaMemcpy(potential, w->B.potential, sizeof(double) * N, cudaMemcpyDeviceToHost));
    if (initial_z) CK(cudaMemcpy(initial_z, w->B.initial_z, sizeof(float) * N, cudaMemcpyDeviceToHost));
    if (frame) CK(cudaMemcpy(frame, w->B.frame, sizeof(int) * N, cudaMemcpyDeviceToHost));
    if (feet_contact && F > 0) {
        std::vector<float> f(N * F);
        CK(cudaMemcpy(f.data(), w->B.feet, sizeof(float) * f.size(), cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < N; e++) for (int k = 0; k < F; k++) feet_contact[e * F + k] = f[(size_t)k * N + e];
    }
    return 0;
}

extern "C" int rs_world_set_flags(rs_world* w, const double* target_xy, const int* timeout, const uint32_t* draws) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n;
    CK(cudaDeviceSynchronize());
    if (target_xy) {
        std::vector<double> t(2 * N);
        for (size_t e = 0; e < N; e++) { t[e] = target_xy[2 * e]; t[N + e] = target_xy[2 * e + 1]; }
        CK(cudaMemcpy(w->B.target, t.data(), sizeof(double) * 2 * N, cudaMemcpyHostToDevice));
    }
    if (timeout) CK(cudaMemcpy(w->B.flag_timeout, timeout, sizeof(int) * N, cudaMemcpyHostToDevice));
    if (draws) CK(cudaMemcpy(w->B.flag_draws, draws, sizeof(uint32_t) * N, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int rs_world_get_flags(rs_world* w, double* target_xy, int* timeout, uint32_t* draws) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n;
    CK(cudaDeviceSynchronize());
    if (target_xy) {
        std::vector<double> t(2 * N);
        CK(cudaMemcpy(t.data(), w->B.target, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < N; e++) { target_xy[2 * e] = t[e]; target_xy[2 * e + 1] = t[N + e]; }
    }
    if (timeout) CK(cudaMemcpy(timeout, w->B.flag_timeout, sizeof(int) * N, cudaMemcpyDeviceToHost));
    if (draws) CK(cudaMemcpy(draws, w->B.flag_draws, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    return 0;
}

// FlagrunHarder task state, 8 doubles per env: task, on_ground, crawl_valid, crawl_start, crawl_ignored, hist[0..2]
extern "C" int rs_world_set_task_state(rs_world* w, const double* in_host) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n;
    std::vector<int> hi(7 * N); std::vector<double> hd(2 * N);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hi.data(), w->B.hx_i, sizeof(int) * 7 * N, cudaMemcpyDeviceToHost));   // keeps done_latch
    for (size_t e = 0; e < N; e++) {
        const double* s = in_host + 8 * e;
        hi[e] = (int)s[0]; hi[N + e] = (int)s[1]; hi[2 * N + e] = (int)s[2];
        for (int k = 0; k < 3; k++) hi[(4 + k) * N + e] = (int)(uint32_t)s[5 + k];
        hd[e] = s[3]; hd[N + e] = s[4];
    }
    CK(cudaMemcpy(w->B.hx_i, hi.data(), sizeof(int) * 7 * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->B.hx_d, hd.data(), sizeof(double) * 2 * N, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int rs_world_get_task_state(rs_world* w, double* out_host) {
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n;
    std::vector<int> hi(7 * N); std::vector<double> hd(2 * N);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hi.data(), w->B.hx_i, sizeof(int) * 7 * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hd.data(), w->B.hx_d, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < N; e++) {
        double* s = out_host + 8 * e;
        s[0] = hi[e]; s[1] = hi[N + e]; s[2] = hi[2 * N + e]; s[3] = hd[e]; s[4] = hd[N + e];
        for (int k = 0; k < 3; k++) s[5 + k] = (double)(uint32_t)hi[(4 + k) * N + e];
    }
    return 0;
}

extern "C" int rs_world_link_poses(rs_world* w, float* out_host) {
    CK(cudaSetDevice(w->device));
    size_t cnt = (size_t)w->n * (w->desc.n_links + 1) * 10;
    float* d = nullptr;
    CK(cudaMalloc(&d, sizeof(float) * cnt));
    CK(wait_caller(w));
    FAMILY_DISPATCH(w, (k_link_poses<NL, MC><<<nblocks(w->n), 64, 0, w->stream>>>(w->M, w->B, d)));
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d, sizeof(float) * cnt, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(std::string("rs_world_link_poses: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int rs_world_get_contacts(rs_world* w, int env, int* keys_host, float* data_host, int max_pts, int* count) {
    CK(cudaSetDevice(w->device));
    if (env < 0 || env >= w->n) return fail("rs_world_get_contacts: env out of range");
    CK(cudaDeviceSynchronize());
    int nc = 0;
    CK(cudaMemcpy(&nc, w->B.ccount + env, sizeof(int), cudaMemcpyDeviceToHost));
    if (nc > max_pts) nc = max_pts;
    size_t N = (size_t)w->n;
    for (int i = 0; i < nc; i++) {
        CK(cudaMemcpy(keys_host + i, w->B.ckey + (size_t)i * N + env, sizeof(int), cudaMemcpyDeviceToHost));
        for (int k = 0; k < CDAT; k++)
            CK(cudaMemcpy(data_host + i * CDAT + k, w->B.cdat + ((size_t)i * CDAT + k) * N + env, sizeof(float), cudaMemcpyDeviceToHost));
    }
    *count = nc;
    return 0;
}

extern "C" int rs_world_set_contacts(rs_world* w, const int* counts_host, const int* keys_host, const float* data_host, int max_pts) {
    CK(cudaSetDevice(w->device));
    if (!counts_host || max_pts < 0) return fail("rs_world_set_contacts: bad arguments");
    const size_t N = (size_t)w->n; const int MC = w->B.MCcap;
    std::vector<int> cnt(N), key(N * MC, 0);
    std::vector<float> dat(N * MC * CDAT, 0.f);
    for (size_t e = 0; e < N; e++) {
        int c = counts_host[e];
        if (c < 0 || c > max_pts) return fail("rs_world_set_contacts: count out of range");
        if (c > MC) c = MC;
        if (c > w->desc.max_contacts) c = w->desc.max_contacts;
        cnt[e] = c;
        for (int i = 0; i < c; i++) {
            key[(size_t)i * N + e] = keys_host[e * max_pts + i];
            for (int k = 0; k < CDAT; k++) dat[((size_t)i * CDAT + k) * N + e] = data_host[(e * max_pts + i) * CDAT + k];
        }
    }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(w->B.ccount, cnt.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->B.ckey, key.data(), sizeof(int) * N * MC, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->B.cdat, dat.data(), sizeof(float) * N * MC * CDAT, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int rs_world_get_counts(rs_world* w, int* rows_host, int* contacts_host) {
    CK(cudaSetDevice(w->device));
    CK(cudaDeviceSynchronize());
    if (rows_host) CK(cudaMemcpy(rows_host, w->B.nrows, sizeof(int) * w->n, cudaMemcpyDeviceToHost));
    if (contacts_host) CK(cudaMemcpy(contacts_host, w->B.ncontacts, sizeof(int) * w->n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rs_world_get_rewards5(rs_world* w, float* out_host) {
    CK(cudaSetDevice(w->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, w->B.rewards5, sizeof(float) * w->n * 5, cudaMemcpyDeviceToHost));
    return 0;
}

// implemented in rs_mlp.cu
extern "C" int rs_policy_forward(rs_policy* p, const float* obs_dev, float* act_dev, int n, void* stream);

extern "C" int rs_world_rollout(rs_world* w, rs_policy* p, int n_steps, double* stats_dev, void* stream) {
    const uint32_t seed = w->rollout_seed, env_id0 = w->env_id0;   // as stored by the last reset
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    for (int t = 0; t < n_steps; t++) {
        const float* act = nullptr;
        if (p) {
            // obs of the previous step (or of reset) lives in d_obs
            if (rs_policy_forward(p, w->d_obs, w->d_act, w->n, stream)) return 1;
            act = w->d_act;
        }
        if (launch_step(w, act, w->d_obs, w->d_rew, w->d_done, 1, seed, env_id0, stats_dev, st)) return 1;
    }
    CK(mark_caller(w, st));
    return 0;
}

// Rollout recorder (the caller side of step for PPO-style trainers): T policy+env steps written straight
// into the caller's time-major tensors -- obs [T+1][N][O] (obs[0] = current observations, filled here from the world's
// buffer, or from obs0_dev if given), act [T][N][A], rew [T][N], done [T][N].  No staging copies: the policy reads obs[t] and writes act[t]; the
// step kernel reads act[t] and writes obs[t+1], rew[t], done[t]; finished envs auto-reset (done[t]=1, obs[t+1] is the
// first observation of the new episode).  p == NULL: act[t] must already hold the actions to replay.
extern "C" int rs_world_rollout_record(rs_world* w, rs_policy* p, int n_steps, const float* obs0_dev, float* obs_T, float* act_T, float* rew_T,
                                       uint8_t* done_T, double* stats_dev, void* stream) {
    if (!obs_T || !act_T || !rew_T || !done_T || n_steps <= 0) return fail("rs_world_rollout_record: bad arguments");
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)w->n, O = (size_t)w->desc.obs_dim, A = (size_t)w->desc.n_act;
    CK(cudaMemcpyAsync(obs_T, obs0_dev ? obs0_dev : w->d_obs, sizeof(float) * N * O, cudaMemcpyDeviceToDevice, st));
    for (int t = 0; t < n_steps; t++) {
        float* a = act_T + (size_t)t * N * A;
        if (p && rs_policy_forward(p, obs_T + (size_t)t * N * O, a, w->n, stream)) return 1;
        if (launch_step(w, a, obs_T + (size_t)(t + 1) * N * O, rew_T + (size_t)t * N, done_T + (size_t)t * N, 1, w->rollout_seed, w->env_id0,
                        stats_dev, st)) return 1;
    }
    // keep the world's own observation buffer current so that rollouts can be chained
    CK(cudaMemcpyAsync(w->d_obs, obs_T + (size_t)n_steps * N * O, sizeof(float) * N * O, cudaMemcpyDeviceToDevice, st));
    CK(mark_caller(w, st));
    return 0;
}

// End-to-end call for callers whose observations live in HOST memory: H2D obs -> policy on the tensor cores ->
// env step (auto-reset) -> D2H obs/reward/done, all on the world's stream, one synchronisation at the end.
// page-locked (cudaHostAlloc / cudaHostRegister) host memory can be the source / target of an async copy directly;
// pageable memory goes through the world's pinned staging buffers
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Pipelined form of the host entry point for worlds on a static-topology kernel and page-locked caller buffers: the envs are cut
// into slices of whole CTAs, each slice runs H2D -> policy -> step -> D2H on its own stream, so the copies of one slice overlap
// the kernels of the others (one copy engine per direction keeps the copies themselves in issue order).  Slices touch disjoint
// envs, constraint-row pools and output rows, so the result is bit-identical to the single-stream form.
uint64_t rs_policy_serial(const rs_policy* p);   // rs_mlp.cu: unique per policy object (a freed policy's address may be reused)

// the slices of one env step: H2D -> policy -> step -> D2H per slice, each on its own stream behind chunk_ev
static int enqueue_slices(rs_world* w, rs_policy* p, const float* obs_in_host, float* obs_out_host, float* reward_host,
                          uint8_t* done_host, int n_chunks, int per, int ncta, bool join) {
    const size_t O = (size_t)w->desc.obs_dim, A = (size_t)w->desc.n_act;
    // slice boundaries in CTAs.  Only the first slice's H2D and the last slice's D2H are exposed, so with 4 or more slices
    // the two outer ones are half as large as the inner ones (RS_E2E_EVEN=1: equal slices).  With cluster launches a boundary
    // is a multiple of the cluster size: a slice with an odd number of CTAs is padded with a CTA that owns no env but occupies
    // an SM to the end, and six such slices asked for 152 SMs (a second wave) where 148 exist.
    int bound[RS_MAX_CHUNKS + 1];
    {
        const char* ev = getenv("RS_E2E_EVEN");
        const bool even = n_chunks < 4 || (ev && atoi(ev) != 0);
        const double units = even ? (double)n_chunks : (double)(n_chunks - 1);     // outer slices weigh 1/2
        const int cl = (w->cluster > 1 && w->wpc > 1) ? w->cluster : 1;
        double acc = 0.0;
        bound[0] = 0;
        for (int c = 0; c < n_chunks; c++) {
            acc += (even || (c > 0 && c < n_chunks - 1)) ? 1.0 : 0.5;
            bound[c + 1] = (int)(acc / units * ncta / cl + 0.5) * cl;
            if (bound[c + 1] > ncta) bound[c + 1] = ncta;
        }
        bound[n_chunks] = ncta;
    }
    for (int c = 0; c < n_chunks; c++) {
        const int c0 = bound[c], c1 = bound[c + 1];
        if (c0 >= c1) continue;
        const size_t e0 = (size_t)c0 * per, e1 = ((size_t)c1 * per < (size_t)w->n) ? (size_t)c1 * per : (size_t)w->n, ne = e1 - e0;
        cudaStream_t st = w->chunk_stream[c];
        CK(cudaStreamWaitEvent(st, w->chunk_ev, 0));
        if (obs_in_host) CK(cudaMemcpyAsync(w->d_obs + e0 * O, obs_in_host + e0 * O, sizeof(float) * ne * O, cudaMemcpyHostToDevice, st));
        if (rs_policy_forward(p, w->d_obs + e0 * O, w->d_act + e0 * A, (int)ne, (void*)st)) return 1;
        if (launch_step(w, w->d_act, w->d_obs, w->d_rew, w->d_done, 1, w->rollout_seed, w->env_id0, nullptr, st, c0, c1 - c0)) return 1;
        CK(cudaMemcpyAsync(obs_out_host + e0 * O, w->d_obs + e0 * O, sizeof(float) * ne * O, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(reward_host + e0, w->d_rew + e0, sizeof(float) * ne, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(done_host + e0, w->d_done + e0, ne, cudaMemcpyDeviceToHost, st));
        if (join) {     // graph form: the slices join the origin stream again
            CK(cudaEventRecord(w->chunk_done[c], st));
            CK(cudaStreamWaitEvent(w->stream, w->chunk_done[c], 0));
        }
    }
    return 0;
}

static int policy_step_host_pipelined(rs_world* w, rs_policy* p, const float* obs_in_host, float* obs_out_host, float* reward_host,
                                      uint8_t* done_host, int n_chunks) {
    int per = 0;
    {
        StaticLaunch q{&w->M, &w->B, w->n, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, nullptr, nullptr, nullptr};
        q.per_out = &per;
        StaticLaunchFn fn = pick_static(w->topo, 0, w->wpc);
        if (!fn) return fail("static-topology kernel not linked");
        fn(q);
    }
    const int ncta = (w->n + per - 1) / per;
    if (n_chunks > ncta) n_chunks = ncta;
    while (w->n_chunk_streams < n_chunks) {
        if (!w->n_chunk_streams) CK(cudaEventCreateWithFlags(&w->chunk_ev, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&w->chunk_stream[w->n_chunk_streams], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&w->chunk_done[w->n_chunk_streams], cudaEventDisableTiming));
        w->n_chunk_streams++;
    }
    CK(wait_caller(w));
    if (const char* e = getenv("RS_E2E_GRAPH")) { if (atoi(e) == 0) w->graph_off = 1; }
    if (!w->graph_off) {
        // replay the step as one graph launch when this set of buffers has been seen before; capture it on its second use (a
        // caller that passes fresh buffers every step never pays for a capture)
        const uint64_t pol = rs_policy_serial(p);
        rs_world::StepGraph* g = nullptr;
        for (int k = 0; k < w->n_graphs; k++) {
            rs_world::StepGraph& q = w->graphs[k];
            if (q.in == obs_in_host && q.out == obs_out_host && q.rew == reward_host && q.done == done_host && q.pol == pol &&
                q.n_chunks == n_chunks && q.seed == w->rollout_seed && q.env_id0 == w->env_id0) { g = &q; break; }
        }
        if (!g) {
            // first sight: remember the key (exec == nullptr), run this step with direct launches
            int k;
            if (w->n_graphs < RS_MAX_GRAPHS) k = w->n_graphs++;
            else { k = w->graph_next++ % RS_MAX_GRAPHS; if (w->graphs[k].exec) cudaGraphExecDestroy(w->graphs[k].exec); }
            w->graphs[k] = rs_world::StepGraph{obs_in_host, obs_out_host, reward_host, done_host, pol, n_chunks, w->rollout_seed, w->env_id0, nullptr};
        } else {
            if (!g->exec) {
                cudaGraph_t graph = nullptr;
                bool ok = cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
                if (ok) {
                    ok = cudaEventRecord(w->chunk_ev, w->stream) == cudaSuccess &&
                         enqueue_slices(w, p, obs_in_host, obs_out_host, reward_host, done_host, n_chunks, per, ncta, true) == 0;
                    cudaError_t ce = cudaStreamEndCapture(w->stream, &graph);
                    ok = ok && ce == cudaSuccess && graph != nullptr;
                }
                if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                if (!ok) {
                    cudaGetLastError();
                    g->exec = nullptr; w->graph_off = 1;
                    fprintf(stderr, "roboschool_b200: CUDA graph capture of the pipelined host step failed (%s); using direct launches\n", g_err.c_str());
                }
            }
            if (g->exec) {
                CK(cudaGraphLaunch(g->exec, w->stream));
                CK(cudaStreamSynchronize(w->stream));
                return 0;
            }
        }
    }
    CK(cudaEventRecord(w->chunk_ev, w->stream));            // everything queued on the world's / the callers' streams so far comes first
    if (enqueue_slices(w, p, obs_in_host, obs_out_host, reward_host, done_host, n_chunks, per, ncta, false)) return 1;
    for (int c = 0; c < n_chunks; c++) CK(cudaStreamSynchronize(w->chunk_stream[c]));
    return 0;
}

extern "C" int rs_world_policy_step_host(rs_world* w, rs_policy* p, const float* obs_in_host, float* obs_out_host, float* reward_host,
                                         uint8_t* done_host) {
    if (!p || !obs_out_host || !reward_host || !done_host) return fail("rs_world_policy_step_host: null argument");
    CK(cudaSetDevice(w->device));
    size_t N = (size_t)w->n, O = (size_t)w->desc.obs_dim;
    const bool direct = host_is_pinned(obs_out_host) && host_is_pinned(reward_host) && host_is_pinned(done_host);
    {
        // measured (32768 Humanoid envs: +11 %, 24576: +9 %, 16384: none, 1-warp CTAs below that: slower): slices from 24576 envs up.
        // RS_E2E_CHUNKS=k forces k slices for any batch size (1 = the single-stream form).
        int n_chunks = w->n >= 24576 ? 6 : 1;
        if (const char* e = getenv("RS_E2E_CHUNKS")) n_chunks = atoi(e);
        if (n_chunks > RS_MAX_CHUNKS) n_chunks = RS_MAX_CHUNKS;
        if (n_chunks > 1 && w->topo && direct && w->n >= 1024 && (!obs_in_host || host_is_pinned(obs_in_host)))
            return policy_step_host_pipelined(w, p, obs_in_host, obs_out_host, reward_host, done_host, n_chunks);
    }
    CK(wait_caller(w));
    if (obs_in_host) {   // NULL: continue from the observations already on the device (after rs_world_rollout_reset / rollout)
        const float* src = obs_in_host;
        if (!host_is_pinned(obs_in_host)) { memcpy(w->h_obs, obs_in_host, sizeof(float) * N * O); src = w->h_obs; }
        CK(cudaMemcpyAsync(w->d_obs, src, sizeof(float) * N * O, cudaMemcpyHostToDevice, w->stream));
    }
    if (rs_policy_forward(p, w->d_obs, w->d_act, w->n, (void*)w->stream)) return 1;
    if (launch_step(w, w->d_act, w->d_obs, w->d_rew, w->d_done, 1, w->rollout_seed, w->env_id0, nullptr, w->stream)) return 1;
    CK(cudaMemcpyAsync(direct ? obs_out_host : w->h_obs, w->d_obs, sizeof(float) * N * O, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(direct ? reward_host : w->h_rew, w->d_rew, sizeof(float) * N, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(direct ? done_host : w->h_done, w->d_done, N, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    if (!direct) {
        memcpy(obs_out_host, w->h_obs, sizeof(float) * N * O);
        memcpy(reward_host, w->h_rew, sizeof(float) * N);
        memcpy(done_host, w->h_done, N);
    }
    return 0;
}

// obs buffer owned by the world (the rollout's policy input); reset writes the first observation here
extern "C" int rs_world_rollout_reset(rs_world* w, uint32_t seed, uint32_t env_id0, void* stream) {
    return rs_world_reset(w, seed, env_id0, nullptr, w->d_obs, stream);
}
extern "C" int rs_world_rollout_buffers(rs_world* w, float** obs_dev, float** act_dev, float** rew_dev, uint8_t** done_dev) {
    if (obs_dev) *obs_dev = w->d_obs;
    if (act_dev) *act_dev = w->d_act;
    if (rew_dev) *rew_dev = w->d_rew;
    if (done_dev) *done_dev = w->d_done;
    return 0;
}
