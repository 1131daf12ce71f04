// rs_pooled.cuh -- constraint-row setup balanced across the lanes of a warp ("row pooling").
//
// The step kernel maps one env to one lane.  Envs differ in how many constraint rows they have in a
// sub-step (Humanoid: mean 3-4, max 8-12 inside one warp), and building a row -- its Jacobian and its
// unit-impulse response M^-1 J^T, i.e. btMultiBody::fillContactJacobianMultiDof +
// calcAccelerationDeltasMultiDof -- is the most expensive part of the sub-step.  Here the warp pools the
// row-building TASKS of its 32 envs and deals them out evenly:
//
//   owner lanes   enumerate their limit rows / contact points, write one task descriptor each into the
//                 warp's pool (slots assigned by a warp prefix sum = ballot-style compaction);
//   all lanes     take tasks t = lane, lane+32, ...: one task = one limit row (1 right-hand side) or one
//                 contact point (normal + 2 friction rows = 3 right-hand sides through ONE pair of tree
//                 sweeps).  The lane reads the OWNER env's link cache (cos/sin, h, 1/D, base Cholesky,
//                 base rotation) from shared memory -- laid out [slot][lane], so any lane can read any
//                 env of the warp without bank conflicts -- and writes J, M^-1 J^T and the row's
//                 denominator into the pool, [word][slot], slot-contiguous => coalesced;
//   owner lanes   run Bullet's projected Gauss-Seidel over their own rows exactly as before.
//
// The leaves->root sweep only visits the links between the loaded body and the root (the wrench is zero
// elsewhere); the root->leaves sweep visits every link.  Arithmetic per row is the same as in
// the single-row sweeps of round 1; the denominator J.M^-1 J^T is taken as wrench . (spatial response of the body),
// which is the same sum in a different association.
//
// The functions are __host__ __device__: tests/host_emul runs them with a 1-lane "warp" (WarpHost).
#pragma once
#include "rs_static.cuh"

namespace rs {

// pool of row records shared by the lanes of a warp: word k of slot s at p[k * cap + s]
// CAP (slots per pool) is a compile-time constant so that the word index k becomes an immediate offset
// Pool words are written by one lane and read by another within the same sub-step and are dead after it: on the device they
// are accessed with the cache-streaming (evict-first) policy so that they do not push the per-thread local frames and the
// instruction lines out of the L2 (RS_POOL_CS=0 builds the plain accessors for A/B runs).
#ifndef RS_POOL_CS
#define RS_POOL_CS 0
#endif
struct PoolRef {
    float* p;
    RS_HD float get() const {
#if defined(__CUDA_ARCH__) && RS_POOL_CS
        return __ldcs(p);
#else
        return *p;
#endif
    }
    RS_HD void put(float v) const {
#if defined(__CUDA_ARCH__) && RS_POOL_CS
        __stcs(p, v);
#else
        *p = v;
#endif
    }
    RS_HD operator float() const { return get(); }
    RS_HD const PoolRef& operator=(float v) const { put(v); return *this; }
    RS_HD const PoolRef& operator=(const PoolRef& o) const { put(o.get()); return *this; }
    RS_HD const PoolRef& operator+=(float v) const { put(get() + v); return *this; }
};
// The rows of a sub-step are dead once its Gauss-Seidel sweeps are through, but they sit in the L2 as dirty lines and would be
// written back to DRAM when the next sub-step's rows (or another warp's local frame) evict them: ~100 MB per step of writes
// nobody reads.  discard.global.L2 drops a line without the write-back and makes it the preferred victim (RS_POOL_DISCARD=0: off).
#ifndef RS_POOL_DISCARD
#define RS_POOL_DISCARD 1
#endif
// REC > 0: record-major layout -- the REC words of a slot are contiguous (word k of slot s at p[s * REC + k]): a Gauss-Seidel sweep
// then reads a row as 6 full sectors instead of one word out of each of 2 nd + 1 sectors (the builders' stores lose their
// coalescing across lanes instead).  REC == 0: word-major, word k of slot s at p[k * CAP + s] (slot-contiguous, coalesced stores).
template <int CAP, int REC = 0>
struct RowPool {
    float* p;
    RS_HD PoolRef at(int slot, int k) const { return PoolRef{REC > 0 ? p + (slot * REC + k) : p + (k * CAP + slot)}; }
    // drop the 128-byte lines that hold slots [slot0, slot0 + nslots) of words [0, nwords): one line per lane and step (CAP is a multiple of 32
    // floats and the pool of a warp starts on a 128-byte boundary, so no line is shared with live data of anybody else)
    RS_HD void discard(int lane, int width, int slot0, int nslots, int nwords) const {      // slot0: a multiple of 32
#if defined(__CUDA_ARCH__) && RS_POOL_DISCARD
        if (REC > 0) {
            const int total = (nslots * REC + 31) >> 5;       // lines of the contiguous record range (rounded up: the tail line holds dead slots only)
            for (int i = lane; i < total; i += width) {
                const float* a = p + ((size_t)slot0 * REC + 32 * i);
                asm volatile("discard.global.L2 [%0], 128;" :: "l"(a) : "memory");
            }
            return;
        }
        const int lpw = (nslots + 31) >> 5;                 // lines per word
        const int total = lpw * nwords;
        for (int i = lane; i < total; i += width) {
            const int k = i / lpw, l = i - k * lpw;
            const float* a = p + ((size_t)k * CAP + slot0 + 32 * l);
            asm volatile("discard.global.L2 [%0], 128;" :: "l"(a) : "memory");
        }
#endif
    }
};
#ifndef RS_POOL_ROWMAJOR
#define RS_POOL_ROWMAJOR 0
#endif
#ifndef RS_PGS_PAIR
#define RS_PGS_PAIR 0
#endif
template <class T> constexpr int pool_layout_rec() { return RS_POOL_ROWMAJOR ? 2 * (6 + T::ND + (T::CUBE ? 6 : 0)) + 1 : 0; }
// Slot map of a warp's pool (WIDTH = lanes = envs of the warp), two forms chosen per topology (pool_columns<T>):
//   dense    rows of the warp packed in owner order: non-contact row j of an env at lbase + j, contact ci direction w at
//            LT + 3 (cbase + ci) + w (lbase / cbase: warp prefix sums).  Touches the fewest cache lines when the envs of a warp
//            differ in row count (Humanoid family: 0.85 vs 0.96 ms per step against the column form; FlagrunHarder 2.95 vs 3.47).
//   columns  every env owns a column: non-contact row j of env o at j * WIDTH + o, contact ci direction w at
//            (NNC + 3 ci + w) * WIDTH + o, so owner lanes that walk their rows in step -- the Gauss-Seidel sweeps -- read adjacent
//            words.  Pays when every env of a warp has about the same rows (Ant, four feet down: 0.79 -> 0.70 ms per step).
//   both     directory of limit-row tasks at slots [ROWS * WIDTH, ...), of contact tasks at [(ROWS + NNC) * WIDTH, ...): word 0 =
//            the task's first row slot      (NNC = 2 NL non-contact rows, ROWS = NNC + 3 MC rows per env)
#ifdef RS_POOL_COLUMNS
template <class T> constexpr bool pool_columns() { return RS_POOL_COLUMNS != 0; }      // A/B builds
#else
template <class T> constexpr bool pool_columns() { return T::NL <= 12 && !T::CUBE; }
#endif
template <class T> constexpr int pool_nnc() { return T::NL > 0 ? 2 * T::NL : 1; }
template <class T, int MC> constexpr int pool_rows() { return pool_nnc<T>() + 3 * MC; }
template <class T, int MC, int WIDTH> constexpr int pool_cap() { return WIDTH * (pool_rows<T, MC>() + pool_nnc<T>() + MC); }
// words of a row record: [0,nd) J ; [nd,2nd) M^-1 J^T ; [2nd] denominator.  A task descriptor lives in
// words [0,25) of the task's first row slot until the task is executed (needs 2*nd+1 >= 25).
constexpr int POOL_DESC_WORDS = 25;
// solver dofs of a topology: base 6 + joints, + the 6 world-frame dofs of FlagrunHarder's cube (T::CUBE) at the end
template <class T> constexpr int nd_of() { return 6 + T::ND + (T::CUBE ? 6 : 0); }
template <class T> constexpr int pool_rec() { return 2 * nd_of<T>() + 1; }

struct WarpHost {
    static constexpr int WIDTH = 1;
    RS_HD int lane() const { return 0; }
    RS_HD int excl_scan(int v, int& total) const { total = v; return 0; }
    RS_HD int max_all(int v) const { return v; }
    RS_HD void sync() const {}
    RS_HD void cta_sync() const {}
    RS_HD void extra_sync() const {}
    RS_HD bool drop_rows() const { return false; }
    RS_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
    RS_HD float shfl(float v, int) const { return v; }
    RS_HD int shfl(int v, int) const { return v; }
};
#ifdef __CUDACC__
// WPC: warps per CTA.  With WPC > 1 the warps of a CTA are re-aligned at the phase boundaries of the sub-step
// (cta_sync), so that they walk the (large, fully unrolled) instruction stream together and share instruction fetches.
template <int WPC>
struct WarpDev {
    static constexpr int WIDTH = 32;
    bool cluster = false;      // launched as a thread-block cluster: the phase barriers span the CTAs of the cluster
    bool extra = false;        // two more re-alignment points inside the solve phase (before the row tasks, before the PGS): pays
                               // when the envs of a batch differ a lot in row count (FlagrunHarder: 6.0 -> 5.0 ms per step), costs ~5 %
                               // on homogeneous batches
    bool drop = false;         // drop the dead rows of a sub-step from the L2 (RowPool::discard): pays once the batch's working set
                               // competes for the L2 (Humanoid 32768: -4 %), costs 2 % on batches that fit it anyway (Hopper 4096)
    RS_HD bool drop_rows() const { return drop; }
    RS_HD void extra_sync() const { if (extra) cta_sync(); }
    RS_HD void cta_sync() const {
#ifdef __CUDA_ARCH__
        if (WPC > 1) __syncthreads();
        // Cluster-wide re-alignment: the CTAs of a cluster sit on SMs of one GPC and walk the same (instruction-fetch bound)
        // stream; keeping them on the same lines lets the GPC-level instruction cache serve all of them with one L2 fetch.
        if (cluster) asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
    }
    RS_HD int lane() const {
#ifdef __CUDA_ARCH__
        return threadIdx.x & 31;
#else
        return 0;
#endif
    }
    RS_HD int excl_scan(int v, int& total) const {
#ifdef __CUDA_ARCH__
        int x = v;
        const int l = threadIdx.x & 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (l >= o) x += y; }
        total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
#else
        total = v; return 0;
#endif
    }
    RS_HD int max_all(int v) const {
#ifdef __CUDA_ARCH__
        return __reduce_max_sync(0xffffffffu, v);
#else
        return v;
#endif
    }
    RS_HD void sync() const {
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
    }
    RS_HD unsigned ballot(bool p) const {
#ifdef __CUDA_ARCH__
        return __ballot_sync(0xffffffffu, p);
#else
        return p ? 1u : 0u;
#endif
    }
    RS_HD float shfl(float v, int src) const {
#ifdef __CUDA_ARCH__
        return __shfl_sync(0xffffffffu, v, src);
#else
        return v;
#endif
    }
    RS_HD int shfl(int v, int src) const {
#ifdef __CUDA_ARCH__
        return __shfl_sync(0xffffffffu, v, src);
#else
        return v;
#endif
    }
};
#endif

RS_HD float i2f(int v) { union { int i; float f; } u; u.i = v; return u.f; }
RS_HD int f2i(float v) { union { int i; float f; } u; u.f = v; return u.i; }

// ---------------------------------------------------------------------------------------------
// NR simultaneous unit-impulse responses that share the loaded body.
//   HASW: the right-hand sides are wrenches wr[rr] (torque;force) on body `body` (0 = base, i+1 = link i), in that
//         body's coordinates about its COM.  !HASW: one unit generalized force jdir on the joint of link `body-1`.
//   Rows go to pool slots slot0, slot0 + rstr, ...  first: overwrite; else accumulate (second body of a pair).
//   denom[rr] = J . M^-1 J^T of THIS side.
// ---------------------------------------------------------------------------------------------
template <class T, int NR, bool HASW, bool first, int STRIDE, class Pool>
RS_HD void pooled_response(const LinkCache<STRIDE>& lc, int body, const float (*wr)[6], float jdir,
                           const Pool& pool, int slot0, float* denom, int rstr = 1) {   // rstr: slots between the NR rows
    constexpr int n = T::NL, nd = nd_of<T>();
    float F[HASW ? NR : 1][6], Z[NR][6], Y[NR][n];
#pragma unroll
    for (int rr = 0; rr < NR; rr++) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if constexpr (HASW) { F[rr][k] = wr[rr][k]; Z[rr][k] = -wr[rr][k]; }
            else Z[rr][k] = 0.f;
        }
        denom[rr] = 0.f;
    }
    int cur = body;   // body whose frame F/Z are expressed in; walks to the root
    static_rfor<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i], d = T::dof[i];
        const bool on = (cur == i + 1);
        if (on) {
            float R[9], r[3];
            cached_xform<T, i>(lc, R, r);
            if constexpr (d >= 0) {
                float h[6];
#pragma unroll
                for (int k = 0; k < 6; k++) h[k] = lc.at(lc_h<T>() + 6 * LcSlot<T, i>::v + k);
                const float invD = lc.at(lc_invd<T>() + LcSlot<T, i>::v);
                const float jj = (!HASW && body == i + 1) ? jdir : 0.f;
#pragma unroll
                for (int rr = 0; rr < NR; rr++) {
                    float j = jj;
                    if constexpr (HASW) j += sdot<T, i>(F[rr]);
                    const float y = jj - sdot<T, i>(Z[rr]);
                    Y[rr][i] = y;
                    const float kk = invD * y;
#pragma unroll
                    for (int k = 0; k < 6; k++) Z[rr][k] += h[k] * kk;
                    const PoolRef J = pool.at(slot0 + rr * rstr, 6 + d);
                    if constexpr (first) J = j; else J += j;
                }
            }
#pragma unroll
            for (int rr = 0; rr < NR; rr++) {
                float t[6];
                xf_force_inv_k<rot_kind<T, i>()>(t, R, r, Z[rr]);
#pragma unroll
                for (int k = 0; k < 6; k++) Z[rr][k] = t[k];
                if constexpr (HASW) {
                    xf_force_inv_k<rot_kind<T, i>()>(t, R, r, F[rr]);
#pragma unroll
                    for (int k = 0; k < 6; k++) F[rr][k] = t[k];
                }
            }
            cur = p + 1;
        } else {
            if constexpr (d >= 0) {
#pragma unroll
                for (int rr = 0; rr < NR; rr++) { Y[rr][i] = 0.f; if constexpr (first) pool.at(slot0 + rr * rstr, 6 + d) = 0.f; }
            }
        }
    });
    float acc[NR][n + 1][6];
    if constexpr (T::FIXED_BASE) {
#pragma unroll
        for (int rr = 0; rr < NR; rr++) {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                acc[rr][0][k] = 0.f;
                if constexpr (first) { pool.at(slot0 + rr * rstr, k) = 0.f; pool.at(slot0 + rr * rstr, nd + k) = 0.f; }
            }
        }
    } else {
        float Lb[21], Rb[9];
#pragma unroll
        for (int k = 0; k < 21; k++) Lb[k] = lc.at(lc_lb<T>() + k);
#pragma unroll
        for (int k = 0; k < 9; k++) Rb[k] = lc.at(lc_rb<T>() + k);
#pragma unroll
        for (int rr = 0; rr < NR; rr++) {
            float jw[3] = {0.f, 0.f, 0.f}, jv[3] = {0.f, 0.f, 0.f}, sol[6], uw[3], uv[3];
            if constexpr (HASW) { tmulv(jw, Rb, F[rr]); tmulv(jv, Rb, F[rr] + 3); }
            chol6_solve(sol, Lb, Z[rr]);
#pragma unroll
            for (int k = 0; k < 6; k++) acc[rr][0][k] = -sol[k];
            tmulv(uw, Rb, acc[rr][0]); tmulv(uv, Rb, acc[rr][0] + 3);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const PoolRef J0 = pool.at(slot0 + rr * rstr, k), J1 = pool.at(slot0 + rr * rstr, 3 + k);
                const PoolRef U0 = pool.at(slot0 + rr * rstr, nd + k), U1 = pool.at(slot0 + rr * rstr, nd + 3 + k);
                if constexpr (first) { J0 = jw[k]; J1 = jv[k]; U0 = uw[k]; U1 = uv[k]; }
                else { J0 += jw[k]; J1 += jv[k]; U0 += uw[k]; U1 += uv[k]; }
            }
            if constexpr (HASW) { if (body == 0) denom[rr] = dot6(wr[rr], acc[rr][0]); }
        }
    }
    static_for<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i], d = T::dof[i];
        float R[9], r[3];
        cached_xform<T, i>(lc, R, r);
        float h[6], invD = 0.f;
        if constexpr (d >= 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) h[k] = lc.at(lc_h<T>() + 6 * LcSlot<T, i>::v + k);
            invD = lc.at(lc_invd<T>() + LcSlot<T, i>::v);
        }
#pragma unroll
        for (int rr = 0; rr < NR; rr++) {
            xf_motion_k<rot_kind<T, i>()>(acc[rr][i + 1], R, r, acc[rr][p + 1]);
            if constexpr (d >= 0) {
                const float a = invD * (Y[rr][i] - dot6(acc[rr][i + 1], h));
                saxpy<T, i>(acc[rr][i + 1], a);
                const PoolRef U = pool.at(slot0 + rr * rstr, nd + 6 + d);
                if constexpr (first) U = a; else U += a;
                if constexpr (!HASW) { if (body == i + 1) denom[rr] = a * jdir; }
            }
            if constexpr (HASW) { if (body == i + 1) denom[rr] = dot6(wr[rr], acc[rr][i + 1]); }
        }
    });
}

// second body of a two-body contact: accumulate variant, kept out of line (rare; keeps the hot instruction stream short)
template <class T, int STRIDE, class Pool>
RS_HDN void pooled_response_second(const LinkCache<STRIDE>& lc, int body, const float (*wr)[6], const Pool& pool, int slot0, float* denom, int rstr) {
    pooled_response<T, 3, true, false>(lc, body, wr, 0.f, pool, slot0, denom, rstr);
}

// resolveSingleConstraintRowGeneric on a pooled row record, over the dofs [K0, K1) the row acts on: a row between the robot
// and the floor (or itself) has exact zeros on the cube's dofs and vice versa, so leaving those words out changes nothing
// (0 * dv adds an exact zero to the dot product, u = 0 leaves dv alone) -- and they are then never written either.
template <int K0, int K1, int ND, class Pool>
RS_HD void resolve_pooled(const Pool& pool, int slot, float* dv, float rhs, float invd, float lo, float hi, float& applied) {
    float j[K1 - K0], u[K1 - K0];
#pragma unroll
    for (int k = K0; k < K1; k++) { j[k - K0] = pool.at(slot, k); u[k - K0] = pool.at(slot, ND + k); }
    float dot = 0.f;
#pragma unroll
    for (int k = K0; k < K1; k++) dot += j[k - K0] * dv[k];
    float dimp = rhs - dot * invd;
    float sum = applied + dimp;
    if (sum < lo) { dimp = lo - applied; applied = lo; }
    else if (sum > hi) { dimp = hi - applied; applied = hi; }
    else applied = sum;
#pragma unroll
    for (int k = K0; k < K1; k++) dv[k] += u[k - K0] * dimp;
}
// The two friction rows of a contact point, resolved one after the other like resolve_pooled does, but with the loads of BOTH rows
// issued before the first is resolved: a row's J / M^-1 J^T words do not depend on dv, so the second row's L2 round trip runs
// under the first row's arithmetic (the Gauss-Seidel sweeps are bound by that latency).  Same operations in the same order per row.
template <int K0, int K1, int ND, class Pool>
RS_HD void resolve_pair_pooled(const Pool& pool, int slotA, int slotB, float* dv, const float* rhs, const float* invd, float lim, float* applied) {
    float jA[K1 - K0], uA[K1 - K0], jB[K1 - K0], uB[K1 - K0];
#pragma unroll
    for (int k = K0; k < K1; k++) { jA[k - K0] = pool.at(slotA, k); uA[k - K0] = pool.at(slotA, ND + k); }
#pragma unroll
    for (int k = K0; k < K1; k++) { jB[k - K0] = pool.at(slotB, k); uB[k - K0] = pool.at(slotB, ND + k); }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const float* j = r ? jB : jA;
        const float* u = r ? uB : uA;
        float dot = 0.f;
#pragma unroll
        for (int k = K0; k < K1; k++) dot += j[k - K0] * dv[k];
        float dimp = rhs[r] - dot * invd[r];
        float sum = applied[r] + dimp;
        if (sum < -lim) { dimp = -lim - applied[r]; applied[r] = -lim; }
        else if (sum > lim) { dimp = lim - applied[r]; applied[r] = lim; }
        else applied[r] = sum;
#pragma unroll
        for (int k = K0; k < K1; k++) dv[k] += u[k - K0] * dimp;
    }
}
template <class T, class Pool>
RS_HD void resolve_pair_span(int span, const Pool& pool, int slotA, int slotB, float* dv, const float* rhs, const float* invd, float lim, float* applied) {
    constexpr int nd = nd_of<T>(), ndr = 6 + T::ND;
    if constexpr (T::CUBE != 0) {
        if (span == 1) { resolve_pair_pooled<ndr, nd, nd>(pool, slotA, slotB, dv, rhs, invd, lim, applied); return; }
        if (span == 2) { resolve_pair_pooled<0, nd, nd>(pool, slotA, slotB, dv, rhs, invd, lim, applied); return; }
    }
    resolve_pair_pooled<0, ndr, nd>(pool, slotA, slotB, dv, rhs, invd, lim, applied);
}
// dofs of a contact row by manifold id: 0 robot only, 1 cube only (cube-floor), 2 both (robot geom vs cube)
template <class T>
RS_HD int row_span(const DevModel& M, int key) {
    if (T::CUBE == 0) return 0;
    const int ngp = M.n_geoms + M.n_pairs;
    return key < ngp ? 0 : (key == ngp ? 1 : 2);
}
template <class T, class Pool>
RS_HD void resolve_span(int span, const Pool& pool, int slot, float* dv, float rhs, float invd, float lo, float hi, float& applied) {
    constexpr int nd = nd_of<T>(), ndr = 6 + T::ND;
    if constexpr (T::CUBE != 0) {
        if (span == 1) { resolve_pooled<ndr, nd, nd>(pool, slot, dv, rhs, invd, lo, hi, applied); return; }
        if (span == 2) { resolve_pooled<0, nd, nd>(pool, slot, dv, rhs, invd, lo, hi, applied); return; }
    }
    resolve_pooled<0, ndr, nd>(pool, slot, dv, rhs, invd, lo, hi, applied);
}

// ---------------------------------------------------------------------------------------------
// btMultiBodyConstraintSolver for one sub-step, rows built by the whole warp.
//   valid   : this lane owns a live env (the tail warp's spare lanes only help)
//   lc_lane0: link cache of lane 0 of this warp; the cache of lane o is lc_lane0 + o (host: o == 0)
// ---------------------------------------------------------------------------------------------
template <class T, class Warp, int NL, int MC, bool SL, int STRIDE, class Pool>
RS_HD void solve_pooled(const Warp& warp, bool valid, const DevModel& M, Work<NL, MC, SL>& W, float dt, const Pool& pool, float* lc_lane0, int* used, PhaseClock* pc = nullptr) {
    long long q0 = 0, q1 = 0;
    if (pc) q0 = RS_CLOCK();
    static_assert(pool_rec<T>() >= POOL_DESC_WORDS, "row record too small to hold a task descriptor");
    constexpr int nd = nd_of<T>(), n = T::NL, ndr = 6 + T::ND;   // ndr: the robot's dofs; [ndr, nd): the cube's
    constexpr int CB = n + 1;                                     // body index of the cube
    const int lane = warp.lane();
    // ---- owner: enumerate rows ----
    // non-contact rows: joint limits, then joint motors (the motors join the world after the limit constraints)
    constexpr int NNC = pool_nnc<T>(), Wd = Warp::WIDTH;
    constexpr int LDIR = pool_rows<T, MC>() * Wd, CDIR = LDIR + NNC * Wd;     // task directories behind the rows (slot map above)
    constexpr bool COL = pool_columns<T>();
    constexpr int RST = COL ? Wd : 1;                                         // slots between the three rows of a contact
    int n_nc = 0, nc = 0;
    int lim_code[NNC];
    float lim_pen[NNC];     // limit rows: penetration ; motor rows: desired joint velocity
    float mot_imp[NNC];     // motor rows: impulse bound
    const float inv_dt = 1.0f / dt;
    if (valid) {
        static_for<n>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int d = T::dof[i];
            if constexpr (d >= 0) {
                if (M.has_limit[i]) {
                    const float q = W.q[d];
                    const float plo = q - M.lim_lo[i], phi = M.lim_hi[i] - q;
                    // btMultiBodyJointLimitConstraint: a row per violated side (lower first)
                    if (!(plo > 0.f)) { lim_code[n_nc] = lane | ((i + 1) << 8); lim_pen[n_nc] = plo; n_nc++; }
                    else if (!(phi > 0.f)) { lim_code[n_nc] = lane | ((i + 1) << 8) | (1 << 16); lim_pen[n_nc] = phi; n_nc++; }
                }
            }
        });
        if (W.motor) {
            // btMultiBodyJointMotor::createConstraintRows: velocity target kp (q* - q) / dt + qd + kd (qd* - qd), impulse within
            // -+ bound, no positional term (physics-bullet.cpp:510-537 -> rs_world_set_joint_motor)
            static_for<n>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int d = T::dof[i];
                if constexpr (d >= 0) {
                    const float* mp = W.motor + (size_t)d * W.motor_stride;
                    const size_t fs = (size_t)T::ND * W.motor_stride;
                    const float imp = mp[4 * fs];
                    if (imp > 0.f) {
                        const float kp = mp[0], kd = mp[fs], tp = mp[2 * fs], tv = mp[3 * fs];
                        lim_code[n_nc] = lane | ((i + 1) << 8) | (1 << 17);
                        lim_pen[n_nc] = kp * (tp - W.q[d]) * inv_dt + W.qd[d] + kd * (tv - W.qd[d]);
                        mot_imp[n_nc] = imp; n_nc++;
                    }
                }
            });
        }
        nc = W.ccount[W.cur];
    }
    int LT, CT;
    const int lbase = warp.excl_scan(n_nc, LT);
    const int cbase = warp.excl_scan(nc, CT);
    auto slot_nc = [&](int j) { return COL ? j * Wd + lane : lbase + j; };                       // this lane's rows
    auto slot_c = [&](int ci) { return COL ? (NNC + 3 * ci) * Wd + lane : LT + 3 * (cbase + ci); };  // direction w: + w * RST
    W.n_rows = n_nc + 3 * nc; W.n_contacts = nc;
    // row slots this warp touches in this sub-step (warp-uniform; the caller drops their lines once every lane is through)
    if constexpr (COL) {
        used[0] = 0; used[1] = warp.max_all(n_nc) * Wd; used[2] = NNC * Wd; used[3] = 3 * warp.max_all(nc) * Wd;
    } else {
        used[0] = 0; used[1] = LT + 3 * CT; used[2] = 0; used[3] = 0;
    }
    // The helpers' task lists are compact directories filled by the owners: one entry per limit / motor row, one per contact
    // that has a robot side (a cube-floor point needs no helper: the owner writes its rows, and a resting cube has four of them
    // in every env).
    int TT = 0, n_rt = nc;
    if constexpr (T::CUBE != 0) {
        n_rt = 0;
        if (valid) for (int ci = 0; ci < nc; ci++) n_rt += W.ckey[W.cur][ci] != M.n_geoms + M.n_pairs;
    }
    const int tbase = warp.excl_scan(n_rt, TT);
    int rt_k = 0;
    if (LT + CT == 0) { warp.extra_sync(); warp.extra_sync(); return; }   // warp-uniform (the CTA-wide syncs below must still be matched)
    const int cur = W.cur;
    for (int j = 0; j < n_nc; j++) { pool.at(slot_nc(j), 0) = i2f(lim_code[j]); pool.at(LDIR + lbase + j, 0) = i2f(slot_nc(j)); }
    for (int ci = 0; ci < nc; ci++) {
        const int key = W.ckey[cur][ci];
        const float* cd = W.cdat[cur][ci];
        int bA, bB;
        if (key < M.n_geoms) { bA = M.g_link[key] + 1; bB = -1; }
        else if (!T::CUBE || key < M.n_geoms + M.n_pairs) { int pk = key - M.n_geoms; bA = M.g_link[M.pair_a[pk]] + 1; bB = M.g_link[M.pair_b[pk]] + 1; }
        else if (key == M.n_geoms + M.n_pairs) { bA = CB; bB = -1; }
        else { bA = M.g_link[key - M.n_geoms - M.n_pairs - 1] + 1; bB = CB; }
        float dirs[3][3];
        dirs[0][0] = cd[6]; dirs[0][1] = cd[7]; dirs[0][2] = cd[8];
        plane_space(dirs[0], dirs[1], dirs[2]);
        const int slot = slot_c(ci);
        if constexpr (T::CUBE != 0) {
            if (bA == CB || bB == CB) {
                // The cube's side of the three rows is written by the owner lane itself (its transform is not in the shared link
                // cache): J = (r x n, n) and M^-1 J^T = (R^T I^-1 R (r x n), n / m) on the cube's world-frame dofs, plus this side's
                // J.M^-1 J^T.  A cube-floor point has no robot side: the task is marked done.
                const float* lp = bA == CB ? cd : cd + 3;           // contact point in cube coords
                const float sgn = bA == CB ? 1.f : -1.f;
                float rw[3], cden[3];
                tmulv(rw, W.cR, lp);
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    const float nn[3] = {sgn * dirs[w][0], sgn * dirs[w][1], sgn * dirs[w][2]};
                    float jw[3], tl[3], uw[3];
                    cross3(jw, rw, nn);
                    mulv(tl, W.cR, jw);
#pragma unroll
                    for (int k = 0; k < 3; k++) tl[k] /= M.cube_inertia[k];
                    tmulv(uw, W.cR, tl);
                    const float im = 1.0f / M.cube_mass;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        pool.at(slot + w * RST, ndr + k) = jw[k]; pool.at(slot + w * RST, ndr + 3 + k) = nn[k];
                        pool.at(slot + w * RST, nd + ndr + k) = uw[k]; pool.at(slot + w * RST, nd + ndr + 3 + k) = nn[k] * im;
                    }
                    cden[w] = dot3(jw, uw) + dot3(nn, nn) * im;
                }
                if (bA == CB) {     // (the robot words of these rows are never read: row_span)
#pragma unroll
                    for (int w = 0; w < 3; w++) pool.at(slot + w * RST, 2 * nd) = cden[w];
                    continue;                                        // no helper task (not in the directory)
                }
                // robot geom vs cube: the helpers build the robot side like a floor contact and add the cube's denominators
                pool.at(CDIR + tbase + rt_k, 0) = i2f(slot); rt_k++;
                pool.at(slot, 0) = i2f(lane | (bA << 8) | (1 << 24));
#pragma unroll
                for (int k = 0; k < 3; k++) { pool.at(slot, 1 + k) = cd[k]; pool.at(slot, 13 + k) = cden[k]; }
                float Rb[9];
#pragma unroll
                for (int k = 0; k < 9; k++) Rb[k] = W.Rwl[bA][k];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    float dl[3];
                    mulv(dl, Rb, dirs[w]);
#pragma unroll
                    for (int k = 0; k < 3; k++) pool.at(slot, 4 + 3 * w + k) = dl[k];
                }
                continue;
            }
        }
        pool.at(CDIR + tbase + rt_k, 0) = i2f(slot); rt_k++;
        pool.at(slot, 0) = i2f(lane | (bA << 8) | ((bB + 1) << 16));
#pragma unroll
        for (int k = 0; k < 3; k++) pool.at(slot, 1 + k) = cd[k];
        float Rb[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Rb[k] = W.Rwl[bA][k];
#pragma unroll
        for (int w = 0; w < 3; w++) {
            float dl[3];
            mulv(dl, Rb, dirs[w]);
#pragma unroll
            for (int k = 0; k < 3; k++) pool.at(slot, 4 + 3 * w + k) = dl[k];
        }
        if (bB >= 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) Rb[k] = W.Rwl[bB][k];
#pragma unroll
            for (int k = 0; k < 3; k++) pool.at(slot, 13 + k) = cd[3 + k];
#pragma unroll
            for (int w = 0; w < 3; w++) {
                float dl[3];
                mulv(dl, Rb, dirs[w]);
#pragma unroll
                for (int k = 0; k < 3; k++) pool.at(slot, 16 + 3 * w + k) = -dl[k];
            }
        }
    }
    warp.sync();
    warp.extra_sync();
    if (pc) { q1 = RS_CLOCK(); pc->t[6] += q1 - q0; q0 = q1; }
    // ---- all lanes: limit-row tasks ----
    for (int t0 = 0; t0 < LT; t0 += Warp::WIDTH) {
        const int t = t0 + lane;
        if (t < LT) {
            const int slot = f2i(pool.at(LDIR + t, 0));
            const int code = f2i(pool.at(slot, 0));
            const int o = code & 255, body = (code >> 8) & 255;
            const float dir = ((code >> 16) & 1) ? -1.f : 1.f;
            LinkCache<STRIDE> lco{lc_lane0 + o};
            float dn[1];
            pooled_response<T, 1, false, true>(lco, body, (const float (*)[6]) nullptr, dir, pool, slot, dn);
            pool.at(slot, 2 * nd) = dn[0];
        }
    }
    if (pc) { q1 = RS_CLOCK(); pc->t[7] += q1 - q0; q0 = q1; }
    // ---- all lanes: contact-point tasks (normal + 2 friction rows each) ----
    // A contact between two robot bodies needs a second pair of sweeps (the wrench on the second body, accumulated into the same
    // rows).  Run inside the task it would idle the other lanes of the round for a whole response, in every round that holds one
    // such contact (measured on FlagrunHarder, whose robots lie on the floor with arms against legs: half of the contact-task
    // time).  So second bodies are tasks of their own: their descriptors are dealt to the lanes of the warp (one pending task per
    // lane, handed over with shuffles) and executed together when 32 are pending or the contact tasks are through.  Per row the
    // arithmetic and its order are unchanged (first body written, second body added).
    bool have2 = false;
    int slot2 = 0, own2 = 0, body2 = 0, npend = 0;     // npend: warp-uniform
    float d2[12];                                       // contact point and the three directions in the second body's coords
#pragma unroll
    for (int k = 0; k < 12; k++) d2[k] = 0.f;
    auto flush_second = [&]() {
        warp.sync();                                    // the first bodies' rows are complete and visible
        if (have2) {
            float wr[3][6], dn[3];
#pragma unroll
            for (int w = 0; w < 3; w++) {
                cross3(wr[w], d2, d2 + 3 + 3 * w);
                wr[w][3] = d2[3 + 3 * w]; wr[w][4] = d2[4 + 3 * w]; wr[w][5] = d2[5 + 3 * w];
            }
            LinkCache<STRIDE> lco{lc_lane0 + own2};
            pooled_response_second<T>(lco, body2, (const float (*)[6]) wr, pool, slot2, dn, RST);
#pragma unroll
            for (int w = 0; w < 3; w++) pool.at(slot2 + w * RST, 2 * nd) += dn[w];
            have2 = false;
        }
        npend = 0;
    };
    for (int t0 = 0; t0 < TT; t0 += Warp::WIDTH) {
        const int t = t0 + lane;
        int slot = 0, o = 0, bB = -1;
        float desc[POOL_DESC_WORDS];
#pragma unroll
        for (int k = 13; k < POOL_DESC_WORDS; k++) desc[k] = 0.f;
        if (t < TT) {
            slot = f2i(pool.at(CDIR + t, 0));
            const int code = f2i(pool.at(slot, 0));
            o = code & 255; bB = ((code >> 16) & 255) - 1;
            const int bA = (code >> 8) & 255;
            const bool cube_side = T::CUBE != 0 && ((code >> 24) & 1);
#pragma unroll
            for (int k = 1; k < 13; k++) desc[k] = pool.at(slot, k);
            if (cube_side) {
#pragma unroll
                for (int k = 13; k < 16; k++) desc[k] = pool.at(slot, k);
            }
            if (bB >= 0) {
#pragma unroll
                for (int k = 13; k < POOL_DESC_WORDS; k++) desc[k] = pool.at(slot, k);
            }
            LinkCache<STRIDE> lco{lc_lane0 + o};
            float dsum[3];
            {
                float wr[3][6];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    cross3(wr[w], desc + 1, desc + 4 + 3 * w);
                    wr[w][3] = desc[4 + 3 * w]; wr[w][4] = desc[5 + 3 * w]; wr[w][5] = desc[6 + 3 * w];
                }
                pooled_response<T, 3, true, true>(lco, bA, (const float (*)[6]) wr, 0.f, pool, slot, dsum, RST);
            }
            if (cube_side) { dsum[0] += desc[13]; dsum[1] += desc[14]; dsum[2] += desc[15]; }
#pragma unroll
            for (int w = 0; w < 3; w++) pool.at(slot + w * RST, 2 * nd) = dsum[w];
        }
        // hand this round's second bodies to the lanes without a pending one
        unsigned m2 = warp.ballot(bB >= 0);
        if (m2) {
            if (npend + rs_popc(m2) > Warp::WIDTH) flush_second();
            while (m2) {
                const int src = rs_ctz(m2);
                m2 &= m2 - 1u;
                const int dst = npend++;
                const int s_slot = warp.shfl(slot, src), s_o = warp.shfl(o, src), s_b = warp.shfl(bB, src);
                float v[12];
#pragma unroll
                for (int k = 0; k < 12; k++) v[k] = warp.shfl(desc[13 + k], src);
                if (lane == dst) {
                    have2 = true; slot2 = s_slot; own2 = s_o; body2 = s_b;
#pragma unroll
                    for (int k = 0; k < 12; k++) d2[k] = v[k];
                }
            }
        }
    }
    if (npend > 0) flush_second();
    warp.sync();
    warp.extra_sync();
    if (pc) { q1 = RS_CLOCK(); pc->t[8] += q1 - q0; q0 = q1; }
    if (!valid || W.n_rows == 0) return;
    // ---- owner: right-hand sides, then PGS in Bullet's order ----
    float vel[nd];
#pragma unroll
    for (int k = 0; k < nd; k++) vel[k] = 0.f;
    if constexpr (!T::FIXED_BASE) {
#pragma unroll
        for (int k = 0; k < 3; k++) { vel[k] = W.bw[k]; vel[3 + k] = W.bv[k]; }
    }
#pragma unroll
    for (int d = 0; d < T::ND; d++) vel[6 + d] = W.qd[d];
    if constexpr (T::CUBE != 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { vel[ndr + k] = W.cw[k]; vel[ndr + 3 + k] = W.cv[k]; }
    }
    float r_rhs[NNC + 3 * MC], r_invd[NNC + 3 * MC], r_applied[NNC + 3 * MC];
    auto row_relvel = [&](int slot, int span) {     // J . v over the dofs the row acts on (row_span)
        float s = 0.f;
        if (span != 1) {
            float j[ndr];
#pragma unroll
            for (int k = 0; k < ndr; k++) j[k] = pool.at(slot, k);
#pragma unroll
            for (int k = 0; k < ndr; k++) s += j[k] * vel[k];
        }
        if constexpr (T::CUBE != 0) {
            if (span != 0) {
#pragma unroll
                for (int k = ndr; k < nd; k++) s += pool.at(slot, k) * vel[k];
            }
        }
        return s;
    };
    // local row index: [0,n_nc) limits ; n_nc + 3*ci + w  contact ci, direction w
    for (int j = 0; j < n_nc; j++) {
        const int slot = slot_nc(j);
        const float relvel = row_relvel(slot, 0), denom = pool.at(slot, 2 * nd), pen = lim_pen[j];
        const float invd = denom > RS_FLT_EPS ? 1.0f / denom : 0.f;
        const float erp = pen > M.split_thresh ? M.erp : M.erp2;
        if (lim_code[j] & (1 << 17)) r_rhs[j] = (pen - relvel) * invd;      // motor row: pen holds the desired velocity
        else { r_rhs[j] = pen > M.split_thresh ? (-pen * erp * inv_dt - relvel) * invd : (-relvel) * invd; mot_imp[j] = -1.f; }
        r_invd[j] = invd; r_applied[j] = 0.f;
    }
    for (int ci = 0; ci < nc; ci++) {
        const float pen = W.cdat[cur][ci][9];
        const int span = row_span<T>(M, W.ckey[cur][ci]);
#pragma unroll
        for (int w = 0; w < 3; w++) {
            const int slot = slot_c(ci) + w * RST, row = n_nc + 3 * ci + w;
            const float relvel = row_relvel(slot, span), denom = pool.at(slot, 2 * nd);
            const float invd = denom > RS_FLT_EPS ? 1.0f / denom : 0.f;
            float rhs;
            if (w == 0) {
                const float erp = pen > M.c_split_thresh ? M.c_erp : M.c_erp2;
                float posErr = 0.f, velErr = -relvel;
                if (pen > 0.f) velErr -= pen * inv_dt; else posErr = -pen * erp * inv_dt;
                rhs = pen > M.c_split_thresh ? (posErr + velErr) * invd : velErr * invd;
            } else rhs = -relvel * invd;
            r_rhs[row] = rhs; r_invd[row] = invd; r_applied[row] = 0.f;
        }
    }
    if (pc) { q1 = RS_CLOCK(); pc->t[9] += q1 - q0; q0 = q1; }
    float dv[nd];
#pragma unroll
    for (int k = 0; k < nd; k++) dv[k] = 0.f;
    const float lim_hi_imp = M.max_limit_impulse;
    for (int it = 0; it < M.n_iter; it++) {
        for (int j = 0; j < n_nc; j++) {
            const int row = (it & 1) ? j : n_nc - 1 - j;
            const float mi = mot_imp[row];      // < 0: limit row (unilateral) ; > 0: motor row (+- impulse bound)
            resolve_span<T>(0, pool, slot_nc(row), dv, r_rhs[row], r_invd[row], mi > 0.f ? -mi : 0.f, mi > 0.f ? mi : lim_hi_imp, r_applied[row]);
        }
        for (int ci = 0; ci < nc; ci++) {
            const int row = n_nc + 3 * ci;
            resolve_span<T>(row_span<T>(M, W.ckey[cur][ci]), pool, slot_c(ci), dv, r_rhs[row], r_invd[row], 0.f, 1e10f, r_applied[row]);
        }
        for (int ci = 0; ci < nc; ci++) {
            const float tot = r_applied[n_nc + 3 * ci];
            if (tot > 0.f) {
                const int key = W.ckey[cur][ci];
                float fr;
                if (key < M.n_geoms) fr = M.g_fric_floor[key];
                else if (!T::CUBE || key < M.n_geoms + M.n_pairs) fr = M.pair_friction[key - M.n_geoms];
                else fr = key == M.n_geoms + M.n_pairs ? M.cube_fric_floor : M.cube_fric_geom[key - M.n_geoms - M.n_pairs - 1];
                // measured (profiles/r02_v5_pgs_pair_ab.txt): loading both friction rows before resolving the first pays on the
                // row-heavy cube topology (FlagrunHarder 2.83 -> 2.77 ms per step), nothing on the others, +2 % on Hopper 4096
                if constexpr (RS_PGS_PAIR != 0 || T::CUBE != 0) {
                    const int row = n_nc + 3 * ci + 1;      // rows + 1 and + 2 are adjacent in r_rhs / r_invd / r_applied
                    resolve_pair_span<T>(row_span<T>(M, key), pool, slot_c(ci) + RST, slot_c(ci) + 2 * RST, dv, r_rhs + row, r_invd + row, fr * tot, r_applied + row);
                } else {
#pragma unroll
                    for (int w = 1; w < 3; w++) {
                        const int row = n_nc + 3 * ci + w;
                        resolve_span<T>(row_span<T>(M, key), pool, slot_c(ci) + w * RST, dv, r_rhs[row], r_invd[row], -fr * tot, fr * tot, r_applied[row]);
                    }
                }
            }
        }
    }
    if (pc) { q1 = RS_CLOCK(); pc->t[10] += q1 - q0; }
    if constexpr (!T::FIXED_BASE) {
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bw[k] += dv[k]; W.bv[k] += dv[3 + k]; }
    }
#pragma unroll
    for (int d = 0; d < T::ND; d++) W.qd[d] += dv[6 + d];
    clamp_vel(M, W);
    if constexpr (T::CUBE != 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { W.cw[k] += dv[ndr + k]; W.cv[k] += dv[ndr + 3 + k]; }
        cube_clamp(M, W);
    }
}

// one sub-step with the pooled solver; every lane of the warp must call it (valid == false: helper only)
template <class T, class Warp, int NL, int MC, bool SL, int STRIDE, class Pool>
RS_HD void substep_pooled(const Warp& warp, bool valid, const DevModel& M, Work<NL, MC, SL>& W, const Pool& pool,
                          const LinkCache<STRIDE>& lc, float* lc_lane0, PhaseClock* pc = nullptr) {
    const float dt = M.timestep;
    long long c0 = 0, c1 = 0;
    warp.cta_sync();
    if (pc) c0 = RS_CLOCK();
    if (valid) {
        float gwr[T::NG][6];
        static_kinematics<T, true>(W, lc, gwr);
        if (pc) { c1 = RS_CLOCK(); pc->t[0] += c1 - c0; c0 = c1; }
        static_collide<T>(M, W, gwr);
    }
    if (pc) { c1 = RS_CLOCK(); pc->t[1] += c1 - c0; c0 = c1; }
    warp.cta_sync();
    if (valid) {
        static_aba<T>(M, W, lc, dt);
        if constexpr (T::CUBE != 0) cube_dynamics(M, W, dt);
    }
    if (pc) { c1 = RS_CLOCK(); pc->t[2] += c1 - c0; c0 = c1; }
    warp.cta_sync();
    warp.sync();   // link caches complete before other lanes read them
    int used[4];
    solve_pooled<T, Warp, NL, MC, SL, STRIDE, Pool>(warp, valid, M, W, dt, pool, lc_lane0, used, pc);
    warp.sync();   // all helpers done with this sub-step's caches before kinematics overwrites them
    // the rows are dead: drop their L2 lines instead of letting them be written back (RowPool::discard)
    if (warp.drop_rows()) {
        if (used[1] > 0) pool.discard(warp.lane(), Warp::WIDTH, used[0], used[1], pool_rec<T>());
        if (used[3] > 0) pool.discard(warp.lane(), Warp::WIDTH, used[2], used[3], pool_rec<T>());
    }
    if (pc) { c1 = RS_CLOCK(); pc->t[3] += c1 - c0; c0 = c1; }
    if (valid) {
        integrate(M, W, dt);
        if constexpr (T::CUBE != 0) cube_integrate(W, dt);
    }
    if (pc) { c1 = RS_CLOCK(); pc->t[4] += c1 - c0; }
}

}  // namespace rs
