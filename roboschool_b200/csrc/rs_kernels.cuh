// rs_kernels.cuh -- device buffers, env load/store and the static-topology step kernel template.
// Included by rs_step.cu (ABI + runtime-topology kernels) and by rs_static_inst.cu, which is compiled once per
// (topology, profiling) pair so that the instantiations build in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include "rs_pooled.cuh"

using namespace rs;

struct DevBuf {
    float* state; int* ccount; int* ckey; float* cdat;
    double* potential; float* initial_z; float* feet; int* frame; uint32_t* episode;
    double* target; int* flag_timeout; uint32_t* flag_draws;   // Flagrun: target [2][N], steps left, RNG draws used
    int* hx_i; double* hx_d;   // FlagrunHarder: [7][N] task, on_ground, crawl_valid, done_latch, hist[3] ; [2][N] crawl_start, crawl_ignored
    float* rows; float* tau; float* rewards5; int* nrows; int* ncontacts;
    float* motor;              // joint motors [5][n_dof][N] (kp, kd, target pos, target vel, impulse bound), nullptr until one is set
    int N, S, MCcap, rec, maxr;
};

// ----------------------------------------------------------------------------------------------
template <int NL, int MC, bool SL>
__device__ __forceinline__ void load_env(const DevModel& M, const DevBuf& B, int e, Work<NL, MC, SL>& W, Episode& E) {
    const int N = B.N;
    const float* s = B.state + e;
    int f = 0;
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) W.bpos[k] = s[(size_t)(f++) * N];
#pragma unroll
        for (int k = 0; k < 4; k++) W.bquat[k] = s[(size_t)(f++) * N];
#pragma unroll
        for (int k = 0; k < 3; k++) W.bw[k] = s[(size_t)(f++) * N];
#pragma unroll
        for (int k = 0; k < 3; k++) W.bv[k] = s[(size_t)(f++) * N];
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bpos[k] = 0; W.bw[k] = 0; W.bv[k] = 0; W.bquat[k] = 0; }
        W.bquat[3] = 1.f;
    }
    for (int d = 0; d < M.n_dof; d++) W.q[d] = s[(size_t)(f + d) * N];
    for (int d = 0; d < M.n_dof; d++) W.qd[d] = s[(size_t)(f + M.n_dof + d) * N];
    if (M.has_cube) {   // 13 more state rows: pos3 quat4 omega3 vel3 of the cube
        const int c0 = f + 2 * M.n_dof;
#pragma unroll
        for (int k = 0; k < 3; k++) { W.cpos[k] = s[(size_t)(c0 + k) * N]; W.cw[k] = s[(size_t)(c0 + 7 + k) * N]; W.cv[k] = s[(size_t)(c0 + 10 + k) * N]; }
#pragma unroll
        for (int k = 0; k < 4; k++) W.cquat[k] = s[(size_t)(c0 + 3 + k) * N];
    }
    int nc = B.ccount[e];
    W.cur = 0; W.ccount[0] = nc; W.ccount[1] = 0;
    for (int i = 0; i < nc; i++) {
        W.ckey[0][i] = B.ckey[(size_t)i * N + e];
#pragma unroll
        for (int k = 0; k < CDAT; k++) W.cdat[0][i][k] = B.cdat[((size_t)i * CDAT + k) * N + e];
    }
    E.potential = B.potential[e]; E.initial_z = B.initial_z[e]; E.frame = B.frame[e]; E.episode = B.episode[e];
    for (int k = 0; k < M.n_feet; k++) E.feet_contact[k] = B.feet[(size_t)k * N + e];
    if (M.flag_task) { E.target_x = B.target[e]; E.target_y = B.target[(size_t)N + e]; E.flag_timeout = B.flag_timeout[e]; E.flag_draws = B.flag_draws[e]; }
    else { E.target_x = M.walk_target_x; E.target_y = M.walk_target_y; E.flag_timeout = 0; E.flag_draws = 0; }
    if (M.flag_task == 2) {
        E.task = B.hx_i[e]; E.on_ground = B.hx_i[(size_t)N + e]; E.crawl_valid = B.hx_i[(size_t)2 * N + e]; E.done_latch = B.hx_i[(size_t)3 * N + e];
#pragma unroll
        for (int k = 0; k < 3; k++) E.hist[k] = (uint32_t)B.hx_i[(size_t)(4 + k) * N + e];
        E.crawl_start = B.hx_d[e]; E.crawl_ignored = B.hx_d[(size_t)N + e];
    } else {
        E.task = 0; E.on_ground = 0; E.crawl_valid = 0; E.done_latch = 0; E.hist[0] = E.hist[1] = E.hist[2] = 0u; E.crawl_start = 0.0; E.crawl_ignored = 0.0;
    }
    E.elec_cost = M.electricity_cost;
    W.n_rows = 0; W.n_contacts = 0;
    W.motor = B.motor ? B.motor + e : nullptr; W.motor_stride = N;
}

template <int NL, int MC, bool SL>
__device__ __forceinline__ void store_env(const DevModel& M, const DevBuf& B, int e, const Work<NL, MC, SL>& W, const Episode& E) {
    const int N = B.N;
    float* s = B.state + e;
    int f = 0;
    if (!M.fixed_base) {
#pragma unroll
        for (int k = 0; k < 3; k++) s[(size_t)(f++) * N] = W.bpos[k];
#pragma unroll
        for (int k = 0; k < 4; k++) s[(size_t)(f++) * N] = W.bquat[k];
#pragma unroll
        for (int k = 0; k < 3; k++) s[(size_t)(f++) * N] = W.bw[k];
#pragma unroll
        for (int k = 0; k < 3; k++) s[(size_t)(f++) * N] = W.bv[k];
    }
    for (int d = 0; d < M.n_dof; d++) s[(size_t)(f + d) * N] = W.q[d];
    for (int d = 0; d < M.n_dof; d++) s[(size_t)(f + M.n_dof + d) * N] = W.qd[d];
    if (M.has_cube) {
        const int c0 = f + 2 * M.n_dof;
#pragma unroll
        for (int k = 0; k < 3; k++) { s[(size_t)(c0 + k) * N] = W.cpos[k]; s[(size_t)(c0 + 7 + k) * N] = W.cw[k]; s[(size_t)(c0 + 10 + k) * N] = W.cv[k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) s[(size_t)(c0 + 3 + k) * N] = W.cquat[k];
    }
    const int cur = W.cur, nc = W.ccount[cur];
    B.ccount[e] = nc;
    for (int i = 0; i < nc; i++) {
        B.ckey[(size_t)i * N + e] = W.ckey[cur][i];
#pragma unroll
        for (int k = 0; k < CDAT; k++) B.cdat[((size_t)i * CDAT + k) * N + e] = W.cdat[cur][i][k];
    }
    B.potential[e] = E.potential; B.initial_z[e] = E.initial_z; B.frame[e] = E.frame; B.episode[e] = E.episode;
    for (int k = 0; k < M.n_feet; k++) B.feet[(size_t)k * N + e] = E.feet_contact[k];
    if (M.flag_task) { B.target[e] = E.target_x; B.target[(size_t)N + e] = E.target_y; B.flag_timeout[e] = E.flag_timeout; B.flag_draws[e] = E.flag_draws; }
    if (M.flag_task == 2) {
        B.hx_i[e] = E.task; B.hx_i[(size_t)N + e] = E.on_ground; B.hx_i[(size_t)2 * N + e] = E.crawl_valid; B.hx_i[(size_t)3 * N + e] = E.done_latch;
#pragma unroll
        for (int k = 0; k < 3; k++) B.hx_i[(size_t)(4 + k) * N + e] = (int)E.hist[k];
        B.hx_d[e] = E.crawl_start; B.hx_d[(size_t)N + e] = E.crawl_ignored;
    }
    B.nrows[e] = W.n_rows; B.ncontacts[e] = W.n_contacts;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Step kernel specialised on a compile-time topology: one warp per CTA (32 envs), link cache in shared memory,
// constraint rows built by the whole warp (rs_pooled.cuh).  Every lane runs the kernel to the end: the spare lanes of
// the tail warp own no env (valid == false) but still help with the warp's row tasks.
// PROF: also accumulate clock64() deltas per phase into cyc[16] (profiling twin).
template <class T, int MC, bool PROF, int WPC>
__global__ void __launch_bounds__(32 * WPC) k_step_static(const __grid_constant__ DevModel M, const DevBuf B, const float* __restrict__ actions,
                                                    float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
                                                    int auto_reset, uint32_t seed, uint32_t env_id0, double* __restrict__ stats,
                                                    unsigned long long* __restrict__ cyc, int cta0, int cluster, int cta_end) {
    extern __shared__ float s_lc[];
    const int cta = blockIdx.x + cta0;             // a launch may cover a slice of the CTAs (pipelined host entry point)
    const int e = cta * (32 * WPC) + threadIdx.x;
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp in CTA, lane
    const bool valid = e < B.N && cta < cta_end;     // cta_end: CTAs padding a cluster launch own no env
    double st_rew = 0, st_ep = 0, st_len = 0, st_nf = 0, st_rows = 0, st_con = 0, st_steps = 0;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (PROF) t0 = clock64();
    Work<T::NL, MC, true> W;
    Episode E;
    float* const lc0 = s_lc + wic * (32 * lc_size<T>());   // this warp's link caches, [slot][lane]
    LinkCache<32> lc{lc0 + lane};
    float a[MA];
    if (valid) {
        load_env(M, B, e, W, E);
        if (actions) { for (int k = 0; k < M.n_act; k++) a[k] = actions[(size_t)e * M.n_act + k]; }
        else { for (int k = 0; k < M.n_act; k++) a[k] = 2.f * rs_uniform(seed ^ 0xA5A5A5A5u, env_id0 + e, E.episode * 1024u + (uint32_t)E.frame, (uint32_t)k) - 1.f; }
        apply_action(M, W, a, 1);
    }
    constexpr int CAP = pool_cap<T, MC, 32>();      // <= 32 * B.maxr: the allocation covers it
    const RowPool<CAP, pool_layout_rec<T>()> pool{B.rows + (size_t)(cta * WPC + wic) * ((size_t)CAP * pool_rec<T>())};
    WarpDev<WPC> warp;
    warp.cluster = (cluster & 1) != 0; warp.extra = (cluster & 2) != 0; warp.drop = (cluster & 4) != 0;
    PhaseClock pc;
    if (PROF) { for (int k = 0; k < 12; k++) pc.t[k] = 0; t1 = clock64(); }
    for (int s = 0; s < M.frame_skip; s++) substep_pooled<T>(warp, valid, M, W, pool, lc, lc0, PROF ? &pc : nullptr);
    if (PROF) t2 = clock64();
    if (valid) {
        float* o = obs + (size_t)e * M.obs_dim;
        StepOut out;
        auto cs = [&](float* oo, int os, double* dist, float* pitch, int* al) { static_calc_state<T>(M, W, lc, E, oo, os, dist, pitch, al); };
        epilogue_with<T::CUBE != 0>(M, W, E, a, 1, o, 1, out, cs, seed, env_id0 + e);
        int dn = out.done;
        bool finite = true;
        for (int k = 0; k < M.obs_dim; k++) finite = finite && isfinite(o[k]);
        int finished = dn || E.frame >= M.max_episode_steps;
        st_rew = out.reward; st_rows = W.n_rows; st_con = W.n_contacts; st_steps = 1; st_nf = finite ? 0 : 1;
        if (finished) { st_ep = 1; st_len = E.frame; }
        if (auto_reset && finished) {
            // (a reset runs in one lane after the rest of its CTA has finished: measured, moving it into a separate small kernel
            // does not pay -- what a steady-state batch costs over a synchronised one is the spread of per-CTA times, not this code)
            env_reset_with(M, W, E, seed, env_id0 + e, o, 1, cs);
            dn = 1;
        }
        rew[e] = out.reward; done[e] = (uint8_t)dn;
        if (B.rewards5) {
#pragma unroll
            for (int k = 0; k < 5; k++) B.rewards5[(size_t)e * 5 + k] = out.rewards5[k];
        }
        store_env(M, B, e, W, E);
    }
    if (PROF && lane == 0) {
        long long t3 = clock64();
        for (int k = 0; k < 5; k++) atomicAdd(cyc + k, (unsigned long long)pc.t[k]);
        for (int k = 0; k < 5; k++) atomicAdd(cyc + 8 + k, (unsigned long long)pc.t[6 + k]);   // inside the solve phase
        atomicAdd(cyc + 5, (unsigned long long)(t3 - t2));
        atomicAdd(cyc + 6, (unsigned long long)(t1 - t0));
        atomicAdd(cyc + 7, (unsigned long long)(t3 - t0));
    }
    if (stats) {
        st_rew = warp_sum(st_rew); st_ep = warp_sum(st_ep); st_len = warp_sum(st_len); st_nf = warp_sum(st_nf);
        st_rows = warp_sum(st_rows); st_con = warp_sum(st_con); st_steps = warp_sum(st_steps);
        // one set of atomics per CTA: 1024 warps adding to the same seven doubles serialise in the L2 (measured: 3 % of the step)
        double* sv = stats;
        if constexpr (WPC > 1) {
            __shared__ double s_st[WPC][7];
            if (lane == 0) {
                s_st[wic][0] = st_rew; s_st[wic][1] = st_ep; s_st[wic][2] = st_len; s_st[wic][3] = st_nf;
                s_st[wic][4] = st_rows; s_st[wic][5] = st_con; s_st[wic][6] = st_steps;
            }
            __syncthreads();
            if (threadIdx.x < 7) {
                double v = 0;
#pragma unroll
                for (int k = 0; k < WPC; k++) v += s_st[k][threadIdx.x];
                atomicAdd(stats + threadIdx.x, v);
            }
            sv = nullptr;
        }
        if (sv && lane == 0) {
            atomicAdd(stats + 0, st_rew); atomicAdd(stats + 1, st_ep); atomicAdd(stats + 2, st_len); atomicAdd(stats + 3, st_nf);
            atomicAdd(stats + 4, st_rows); atomicAdd(stats + 5, st_con); atomicAdd(stats + 6, st_steps);
        }
    }
}

// Physics-only sub-steps on a compile-time topology (World::step(n) of the drop-in shim, where the env's own Python computes
// observation and reward): the generalized forces come from the world's torque table instead of an action, no epilogue, no reset.
// One warp per CTA.  Instantiated for the topologies the runtime-topology kernel cannot step (FlagrunHarder's cube).
template <class T, int MC>
__global__ void __launch_bounds__(32) k_substeps_static(const __grid_constant__ DevModel M, const DevBuf B, int nsub) {
    extern __shared__ float s_lc[];
    const int e = blockIdx.x * 32 + threadIdx.x, lane = threadIdx.x & 31;
    const bool valid = e < B.N;
    Work<T::NL, MC, true> W;
    Episode E;
    LinkCache<32> lc{s_lc + lane};
    if (valid) {
        load_env(M, B, e, W, E);
        for (int d = 0; d < M.n_dof; d++) W.tau[d] = B.tau[(size_t)d * B.N + e];
    }
    constexpr int CAP = pool_cap<T, MC, 32>();
    const RowPool<CAP, pool_layout_rec<T>()> pool{B.rows + (size_t)blockIdx.x * ((size_t)CAP * pool_rec<T>())};
    WarpDev<1> warp;
    for (int s = 0; s < nsub; s++) substep_pooled<T>(warp, valid, M, W, pool, lc, s_lc, nullptr);
    if (valid) store_env(M, B, e, W, E);
}
struct SubstepsLaunch { const DevModel* M; const DevBuf* B; int n, nsub; cudaStream_t st; };
template <class T, int MC>
static int launch_substeps_static(const SubstepsLaunch& a) {
    const size_t smem = sizeof(float) * 32 * lc_size<T>();
    cudaFuncSetAttribute(k_substeps_static<T, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_substeps_static<T, MC><<<(a.n + 31) / 32, 32, smem, a.st>>>(*a.M, *a.B, a.nsub);
    return 0;
}
typedef int (*SubstepsLaunchFn)(const SubstepsLaunch&);
void rs_register_substeps(int topo, SubstepsLaunchFn fn);

// arguments of one static-topology step launch (plain struct: the instantiations live in other translation units)
struct StaticLaunch {
    const DevModel* M; const DevBuf* B; int n;
    const float* actions; float* obs; float* rew; uint8_t* done;
    int auto_reset; uint32_t seed, env_id0; double* stats; unsigned long long* cyc; cudaStream_t st;
    int cta0 = 0, ncta = 0;      // slice of the grid to launch (ncta == 0: all CTAs)
    int cluster = 0;             // > 1: launch as thread-block clusters of that many CTAs, phase barriers span the cluster
    int extra_sync = 0;          // 1: two more re-alignment points inside the solve phase (heterogeneous batches)
    int drop_rows = 0;           // 1: dead constraint rows are dropped from the L2 after every sub-step (large batches)
    int* per_out = nullptr;      // query: receives envs per CTA, nothing is launched
};

template <class T, int MC, bool PROF, int WPC>
static int launch_static(const StaticLaunch& a) {
    constexpr int per = 32 * WPC;
    if (a.per_out) { *a.per_out = per; return 0; }
    const size_t smem = sizeof(float) * per * lc_size<T>();
    // the opt-in to > 48 KB of dynamic shared memory is per device: remember which devices have it
    static unsigned attr_devices = 0;
    static int regs_per_thread = 0, n_sm = 0, last_carveout = -2;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_devices >> (dev & 31)) & 1u)) {
        cudaFuncSetAttribute(k_step_static<T, MC, PROF, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, k_step_static<T, MC, PROF, WPC>) == cudaSuccess) regs_per_thread = fa.numRegs;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        attr_devices |= 1u << (dev & 31);
        last_carveout = -2;
    }
    {
        // Shared-memory carve-out: just enough for the CTAs that can be resident (registers) and that this grid needs per SM.
        // The default heuristic sizes it for ONE small CTA (1 warp, 10-26 KB of link cache) although registers allow eight, so
        // a multi-wave grid of small CTAs runs at 1/8 of the occupancy; asking for the maximum instead shrinks the L1 that
        // holds the spilled per-env scratch.  RS_CARVEOUT=<percent> overrides, RS_CARVEOUT=-1 keeps the driver's default.
        const int grid_all = (a.n + per - 1) / per;       // the whole world: slices of one step run concurrently
        int by_regs = regs_per_thread > 0 ? 65536 / (((regs_per_thread + 7) / 8 * 8) * per) : 1;
        if (by_regs < 1) by_regs = 1;
        int need = n_sm > 0 ? (grid_all + n_sm - 1) / n_sm : 1;
        if (need > by_regs) need = by_regs;
        int pct = (int)(((size_t)need * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        if (const char* ev = getenv("RS_CARVEOUT")) pct = atoi(ev);
        if (pct != last_carveout && pct >= 0) {
            cudaFuncSetAttribute(k_step_static<T, MC, PROF, WPC>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            last_carveout = pct;
        }
    }
    int grid = a.ncta > 0 ? a.ncta : (a.n + per - 1) / per;
    const int cta_end = a.cta0 + grid;
    // Thread-block clusters of 2 CTAs = the two SMs of a TPC, which share an instruction cache: with cluster-wide phase
    // barriers both SMs walk the same lines of the (instruction-fetch bound) stream.  Measured on a steady-state batch of 32768
    // Humanoid envs: 0.996 -> 0.933 ms per step; clusters of 3..8 CTAs cannot all be resident (one CTA fills an SM) and run
    // 1.4 ms.  The grid is padded with CTAs that own no env (cta >= cta_end) up to a multiple of the cluster size.
    const int extra = (a.extra_sync ? 2 : 0) | (a.drop_rows ? 4 : 0);
    if (a.cluster > 1 && WPC > 1) {
        grid = (grid + a.cluster - 1) / a.cluster * a.cluster;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(per); cfg.dynamicSmemBytes = smem; cfg.stream = a.st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = a.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_step_static<T, MC, PROF, WPC>, *a.M, *a.B, a.actions, a.obs, a.rew, a.done, a.auto_reset, a.seed, a.env_id0, a.stats, a.cyc, a.cta0, 1 | extra, cta_end);
        return 0;
    }
    k_step_static<T, MC, PROF, WPC><<<grid, per, smem, a.st>>>(*a.M, *a.B, a.actions, a.obs, a.rew, a.done, a.auto_reset, a.seed, a.env_id0, a.stats, a.cyc, a.cta0, extra, cta_end);
    return 0;
}

// registry of the static-topology instantiations (each lives in its own translation unit, rs_static_inst.cu)
typedef int (*StaticLaunchFn)(const StaticLaunch&);
void rs_register_static(int topo, int prof, int wpc, StaticLaunchFn fn);
