// rs_static.cuh -- tree sweeps specialised on a compile-time topology (rs_topo_gen.h).
//
// Same arithmetic as the generic sweeps in rs_core.cuh (kinematics / ABA / unit-impulse response), but
// every loop over links is unrolled against constexpr parent/dof tables: accumulators of the
// leaves->root and root->leaves sweeps live in registers, joint-subspace dot products collapse to the
// non-zero terms, and the per-link cache that the response passes re-read for every constraint row
// (cos q, sin q, h = I^A S, 1/D, Cholesky factor of the base inertia, base rotation: 9*ND+30 floats per env, slots by dof index) is staged
// in SHARED memory, laid out [slot][lane] so a warp's accesses are conflict-free.
#pragma once
#include <utility>
#include "rs_core.cuh"
#include "rs_topo_gen.h"

namespace rs {

template <class F, int... Is>
RS_HD void static_for_impl(F&& f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F>
RS_HD void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }
template <int N, class F, int... Is>
RS_HD void static_rfor_impl(F&& f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, N - 1 - Is>{}), ...); }
template <int N, class F>
RS_HD void static_rfor(F&& f) { static_rfor_impl<N>(f, std::make_integer_sequence<int, N>{}); }

// multiply by a compile-time constant, folding 0 / +-1 (IEEE forbids the compiler to do it on its own)
#define RS_CM(cst, x) ((cst) == 0.0f ? 0.0f : ((cst) == 1.0f ? (x) : ((cst) == -1.0f ? -(x) : (cst) * (x))))

// per-env link cache in shared memory: element k of this lane lives at p[k * STRIDE]
template <int STRIDE>
struct LinkCache {
    float* p;
    RS_HD float& at(int k) const { return p[k * STRIDE]; }
};
// slot of link I's joint quantities (cos / sin, h, 1/D): its dof index -- links without a dof (fixed joints) take no room
template <class T, int I> struct LcSlot { static constexpr int v = T::dof[I]; };
template <class T> constexpr int lc_cs() { return 0; }
template <class T> constexpr int lc_h() { return 2 * T::ND; }
template <class T> constexpr int lc_invd() { return 8 * T::ND; }
template <class T> constexpr int lc_lb() { return 9 * T::ND; }
template <class T> constexpr int lc_rb() { return 9 * T::ND + 21; }     // base rotation (world -> base coords), 9 floats
template <class T> constexpr int lc_size() { return 9 * T::ND + 30; }

// S . x for link I with the constant joint subspace
template <class T, int I>
RS_HD float sdot(const float* x) {
    constexpr float s0 = T::Sw[3 * I], s1 = T::Sw[3 * I + 1], s2 = T::Sw[3 * I + 2], s3 = T::Sv[3 * I], s4 = T::Sv[3 * I + 1], s5 = T::Sv[3 * I + 2];
    return RS_CM(s0, x[0]) + RS_CM(s1, x[1]) + RS_CM(s2, x[2]) + RS_CM(s3, x[3]) + RS_CM(s4, x[4]) + RS_CM(s5, x[5]);
}
// x += a * S
template <class T, int I>
RS_HD void saxpy(float* x, float a) {
    constexpr float s0 = T::Sw[3 * I], s1 = T::Sw[3 * I + 1], s2 = T::Sw[3 * I + 2], s3 = T::Sv[3 * I], s4 = T::Sv[3 * I + 1], s5 = T::Sv[3 * I + 2];
    x[0] += RS_CM(s0, a); x[1] += RS_CM(s1, a); x[2] += RS_CM(s2, a);
    x[3] += RS_CM(s3, a); x[4] += RS_CM(s4, a); x[5] += RS_CM(s5, a);
}

// sin and cos of a joint angle: Cody-Waite reduction by pi/2 (three constants, exact product for |x| < ~1e5 rad -- joint
// angles are a few rad) and the minimax polynomials on [-pi/4, pi/4]; ~1 ulp.  Straight-line: unlike sincosf() there is
// no large-argument slow path to branch around (17 joints x 4 sub-steps of taken branches in the instruction stream).
RS_HD void rs_sincos(float x, float* sn, float* cs) {
    const float j = rintf(x * 0.636619772f);
    const int q = (int)j;
    float r = fmaf(j, -1.57079601e+00f, x);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float s = r * r;
    float ps = fmaf(2.86567956e-6f, s, -1.98559923e-4f);
    ps = fmaf(ps, s, 8.33338592e-3f);
    ps = fmaf(ps, s, -1.66666672e-1f);
    const float sr = fmaf(r * s, ps, r);
    float pc = fmaf(2.44677067e-5f, s, -1.38877297e-3f);
    pc = fmaf(pc, s, 4.16666567e-2f);
    pc = fmaf(pc, s, -0.5f);
    const float cr = fmaf(s, pc, 1.0f);
    const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;     // sin, cos before the sign of the quadrant
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// rotation parent->link coords and COM offset of link I from (cos q, sin q, q)
template <class T, int I>
RS_HD void link_xform(float c, float s, float q, float* R, float* r) {
    constexpr float R0_0 = T::R0[9 * I], R0_1 = T::R0[9 * I + 1], R0_2 = T::R0[9 * I + 2], R0_3 = T::R0[9 * I + 3], R0_4 = T::R0[9 * I + 4],
                    R0_5 = T::R0[9 * I + 5], R0_6 = T::R0[9 * I + 6], R0_7 = T::R0[9 * I + 7], R0_8 = T::R0[9 * I + 8];
    constexpr float ax = T::axis[3 * I], ay = T::axis[3 * I + 1], az = T::axis[3 * I + 2];
    constexpr bool r0_identity = R0_0 == 1.0f && R0_4 == 1.0f && R0_8 == 1.0f && R0_1 == 0.0f && R0_2 == 0.0f && R0_3 == 0.0f &&
                                 R0_5 == 0.0f && R0_6 == 0.0f && R0_7 == 0.0f;
    constexpr int jt = T::jtype[I];
    if constexpr (jt == RS_JOINT_REVOLUTE) {
        // rotation by -q about the axis:  c I + (1-c) a a^T - s [a]x
        const float k = 1.0f - c;
        float Ra[9];
        Ra[0] = c + RS_CM(ax * ax, k);          Ra[1] = RS_CM(ax * ay, k) + RS_CM(az, s); Ra[2] = RS_CM(ax * az, k) - RS_CM(ay, s);
        Ra[3] = RS_CM(ay * ax, k) - RS_CM(az, s); Ra[4] = c + RS_CM(ay * ay, k);          Ra[5] = RS_CM(ay * az, k) + RS_CM(ax, s);
        Ra[6] = RS_CM(az * ax, k) + RS_CM(ay, s); Ra[7] = RS_CM(az * ay, k) - RS_CM(ax, s); Ra[8] = c + RS_CM(az * az, k);
        if constexpr (r0_identity) {
#pragma unroll
            for (int j = 0; j < 9; j++) R[j] = Ra[j];
        } else {
            const float R0v[9] = {R0_0, R0_1, R0_2, R0_3, R0_4, R0_5, R0_6, R0_7, R0_8};
            mul33(R, Ra, R0v);
        }
    } else {
        R[0] = R0_0; R[1] = R0_1; R[2] = R0_2; R[3] = R0_3; R[4] = R0_4; R[5] = R0_5; R[6] = R0_6; R[7] = R0_7; R[8] = R0_8;
    }
    constexpr float ex = T::e[3 * I], ey = T::e[3 * I + 1], ez = T::e[3 * I + 2];
    constexpr float dx = T::d[3 * I], dy = T::d[3 * I + 1], dz = T::d[3 * I + 2];
    r[0] = RS_CM(ex, R[0]) + RS_CM(ey, R[1]) + RS_CM(ez, R[2]) + dx;
    r[1] = RS_CM(ex, R[3]) + RS_CM(ey, R[4]) + RS_CM(ez, R[5]) + dy;
    r[2] = RS_CM(ex, R[6]) + RS_CM(ey, R[7]) + RS_CM(ez, R[8]) + dz;
    if constexpr (jt == RS_JOINT_PRISMATIC) { r[0] += RS_CM(ax, q); r[1] += RS_CM(ay, q); r[2] += RS_CM(az, q); }
}

// ---------------------------------------------------------------------------------------------
// Structure of the parent->link rotation, known at compile time.  For a revolute joint about a principal axis with an
// identity zero-rotation, R is a planar rotation (4 non-trivial entries); for a fixed joint with identity zero-rotation it
// is the identity.  IEEE arithmetic forbids the compiler to drop 0*x on its own, so the sweeps call these helpers with
// the kind as a template argument: a planar R costs 4 instead of 9 multiply-adds per vector, 24 instead of 54 per
// similarity transform.  (k, a, b) cyclic; rows: R[k] = e_k, R[a] = c e_a + s e_b, R[b] = -s e_a + c e_b with
// c = R[a][a], s = R[a][b] read from the matrix itself (s carries the sign of the axis).
// ---------------------------------------------------------------------------------------------
enum { ROT_X = 0, ROT_Y = 1, ROT_Z = 2, ROT_GEN = 3, ROT_ID = 4 };
template <class T, int I>
constexpr int rot_kind() {
    const bool r0_identity = T::R0[9 * I] == 1.0f && T::R0[9 * I + 4] == 1.0f && T::R0[9 * I + 8] == 1.0f && T::R0[9 * I + 1] == 0.0f &&
                             T::R0[9 * I + 2] == 0.0f && T::R0[9 * I + 3] == 0.0f && T::R0[9 * I + 5] == 0.0f && T::R0[9 * I + 6] == 0.0f &&
                             T::R0[9 * I + 7] == 0.0f;
    if (!r0_identity) return ROT_GEN;
    if (T::jtype[I] != RS_JOINT_REVOLUTE) return ROT_ID;
    const float ax = T::axis[3 * I], ay = T::axis[3 * I + 1], az = T::axis[3 * I + 2];
    if ((ax == 1.0f || ax == -1.0f) && ay == 0.0f && az == 0.0f) return ROT_X;
    if ((ay == 1.0f || ay == -1.0f) && ax == 0.0f && az == 0.0f) return ROT_Y;
    if ((az == 1.0f || az == -1.0f) && ax == 0.0f && ay == 0.0f) return ROT_Z;
    return ROT_GEN;
}
template <int K> RS_HD void rot_mulv(float* o, const float* R, const float* v) {   // o = R v   (o must not alias v)
    if constexpr (K == ROT_GEN) mulv(o, R, v);
    else if constexpr (K == ROT_ID) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
    else {
        constexpr int k = K, a = (K + 1) % 3, b = (K + 2) % 3;
        const float c = R[a * 3 + a], s = R[a * 3 + b];
        o[k] = v[k]; o[a] = c * v[a] + s * v[b]; o[b] = c * v[b] - s * v[a];
    }
}
template <int K> RS_HD void rot_tmulv(float* o, const float* R, const float* v) {  // o = R^T v
    if constexpr (K == ROT_GEN) tmulv(o, R, v);
    else if constexpr (K == ROT_ID) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
    else {
        constexpr int k = K, a = (K + 1) % 3, b = (K + 2) % 3;
        const float c = R[a * 3 + a], s = R[a * 3 + b];
        o[k] = v[k]; o[a] = c * v[a] - s * v[b]; o[b] = s * v[a] + c * v[b];
    }
}
template <int K> RS_HD void rot_mul33(float* O, const float* R, const float* A) {  // O = R A  (rows of A mixed; O must not alias A)
    if constexpr (K == ROT_GEN) mul33(O, R, A);
    else if constexpr (K == ROT_ID) {
#pragma unroll
        for (int i = 0; i < 9; i++) O[i] = A[i];
    } else {
        constexpr int k = K, a = (K + 1) % 3, b = (K + 2) % 3;
        const float c = R[a * 3 + a], s = R[a * 3 + b];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            O[k * 3 + j] = A[k * 3 + j];
            O[a * 3 + j] = c * A[a * 3 + j] + s * A[b * 3 + j];
            O[b * 3 + j] = c * A[b * 3 + j] - s * A[a * 3 + j];
        }
    }
}
// O = R^T X R for a full 3x3 X (O must not alias X)
template <int K> RS_HD void rot_sim(float* O, const float* R, const float* X) {
    if constexpr (K == ROT_GEN) {
        float T[9], RT[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) RT[i * 3 + j] = R[j * 3 + i];
        mul33(T, X, R); mul33(O, RT, T);
    } else if constexpr (K == ROT_ID) {
#pragma unroll
        for (int i = 0; i < 9; i++) O[i] = X[i];
    } else {
        constexpr int k = K, a = (K + 1) % 3, b = (K + 2) % 3;
        const float c = R[a * 3 + a], s = R[a * 3 + b];
        float Y[9];   // Y = R^T X : rows a, b of X mixed
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Y[k * 3 + j] = X[k * 3 + j];
            Y[a * 3 + j] = c * X[a * 3 + j] - s * X[b * 3 + j];
            Y[b * 3 + j] = s * X[a * 3 + j] + c * X[b * 3 + j];
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {   // O = Y R : columns a, b of Y mixed
            O[i * 3 + k] = Y[i * 3 + k];
            O[i * 3 + a] = c * Y[i * 3 + a] - s * Y[i * 3 + b];
            O[i * 3 + b] = s * Y[i * 3 + a] + c * Y[i * 3 + b];
        }
    }
}
// spatial motion parent->child: w' = R w ; v' = R v - r x (R w)
template <int K> RS_HD void xf_motion_k(float* o, const float* R, const float* r, const float* in) {
    float w[3], v[3], c[3];
    rot_mulv<K>(w, R, in); rot_mulv<K>(v, R, in + 3); cross3(c, r, w);
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2]; o[3] = v[0] - c[0]; o[4] = v[1] - c[1]; o[5] = v[2] - c[2];
}
// spatial force child->parent: f' = R^T f ; t' = R^T (t + r x f)
template <int K> RS_HD void xf_force_inv_k(float* out, const float* R, const float* r, const float* in) {
    float c[3], t[3];
    cross3(c, r, in + 3);
    t[0] = in[0] + c[0]; t[1] = in[1] + c[1]; t[2] = in[2] + c[2];
    rot_tmulv<K>(out, R, t);
    rot_tmulv<K>(out + 3, R, in + 3);
}
template <int K> RS_HD void xf_force_inv_add_k(float* acc, const float* R, const float* r, const float* in) {
    float o[6];
    xf_force_inv_k<K>(o, R, r, in);
#pragma unroll
    for (int k = 0; k < 6; k++) acc[k] += o[k];
}
// Ip += X^T Ic X for X = motion transform parent->child (R, r).  Same algebra as inertia_to_parent_add(), with the skew
// products written out (r~ has six non-zeros) and the rotations specialised on K.
template <int K> RS_HD void inertia_to_parent_add_k(float* Ip, const float* Ic, const float* R, const float* r) {
    float A[9], M[9];
    sym_to_full(A, Ic); sym_to_full(M, Ic + 15);
    const float* B = Ic + 6;
    const float x = r[0], y = r[1], z = r[2];
    // r~ M : row0 = -z M1 + y M2 ; row1 = z M0 - x M2 ; row2 = -y M0 + x M1     (Mi = row i of M)
    float Bp[9], Ap[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Bp[j] = B[j] + (y * M[6 + j] - z * M[3 + j]);
        Bp[3 + j] = B[3 + j] + (z * M[j] - x * M[6 + j]);
        Bp[6 + j] = B[6 + j] + (x * M[3 + j] - y * M[j]);
    }
    // A' = A - B r~ + r~ Bp^T.   (B r~)[i] = ( z Bi1 - y Bi2 , -z Bi0 + x Bi2 , y Bi0 - x Bi1 )
    //                            (r~ Bp^T)[0][j] = -z Bp[j][1] + y Bp[j][2] ; [1][j] = z Bp[j][0] - x Bp[j][2] ; [2][j] = -y Bp[j][0] + x Bp[j][1]
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Ap[i * 3 + 0] = A[i * 3 + 0] - (z * B[i * 3 + 1] - y * B[i * 3 + 2]);
        Ap[i * 3 + 1] = A[i * 3 + 1] - (x * B[i * 3 + 2] - z * B[i * 3 + 0]);
        Ap[i * 3 + 2] = A[i * 3 + 2] - (y * B[i * 3 + 0] - x * B[i * 3 + 1]);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Ap[j] += y * Bp[j * 3 + 2] - z * Bp[j * 3 + 1];
        Ap[3 + j] += z * Bp[j * 3 + 0] - x * Bp[j * 3 + 2];
        Ap[6 + j] += x * Bp[j * 3 + 1] - y * Bp[j * 3 + 0];
    }
    float Ar[9], Br[9], Mr[9];
    rot_sim<K>(Ar, R, Ap); rot_sim<K>(Br, R, Bp); rot_sim<K>(Mr, R, M);
    Ip[0] += Ar[0]; Ip[1] += 0.5f * (Ar[1] + Ar[3]); Ip[2] += 0.5f * (Ar[2] + Ar[6]);
    Ip[3] += Ar[4]; Ip[4] += 0.5f * (Ar[5] + Ar[7]); Ip[5] += Ar[8];
#pragma unroll
    for (int i = 0; i < 9; i++) Ip[6 + i] += Br[i];
    Ip[15] += Mr[0]; Ip[16] += 0.5f * (Mr[1] + Mr[3]); Ip[17] += 0.5f * (Mr[2] + Mr[6]);
    Ip[18] += Mr[4]; Ip[19] += 0.5f * (Mr[5] + Mr[7]); Ip[20] += Mr[8];
}
// h = I^A S with the constant joint subspace of link I (zero entries of S dropped)
template <class T, int I>
RS_HD void inertia_mul_S(float* h, const float* IA) {
    constexpr float s0 = T::Sw[3 * I], s1 = T::Sw[3 * I + 1], s2 = T::Sw[3 * I + 2], s3 = T::Sv[3 * I], s4 = T::Sv[3 * I + 1], s5 = T::Sv[3 * I + 2];
    const float* A = IA; const float* B = IA + 6; const float* M = IA + 15;   // A sym(6) B full(9) M sym(6)
    h[0] = RS_CM(s0, A[0]) + RS_CM(s1, A[1]) + RS_CM(s2, A[2]) + RS_CM(s3, B[0]) + RS_CM(s4, B[1]) + RS_CM(s5, B[2]);
    h[1] = RS_CM(s0, A[1]) + RS_CM(s1, A[3]) + RS_CM(s2, A[4]) + RS_CM(s3, B[3]) + RS_CM(s4, B[4]) + RS_CM(s5, B[5]);
    h[2] = RS_CM(s0, A[2]) + RS_CM(s1, A[4]) + RS_CM(s2, A[5]) + RS_CM(s3, B[6]) + RS_CM(s4, B[7]) + RS_CM(s5, B[8]);
    h[3] = RS_CM(s0, B[0]) + RS_CM(s1, B[3]) + RS_CM(s2, B[6]) + RS_CM(s3, M[0]) + RS_CM(s4, M[1]) + RS_CM(s5, M[2]);
    h[4] = RS_CM(s0, B[1]) + RS_CM(s1, B[4]) + RS_CM(s2, B[7]) + RS_CM(s3, M[1]) + RS_CM(s4, M[3]) + RS_CM(s5, M[4]);
    h[5] = RS_CM(s0, B[2]) + RS_CM(s1, B[5]) + RS_CM(s2, B[8]) + RS_CM(s3, M[2]) + RS_CM(s4, M[4]) + RS_CM(s5, M[5]);
}

// (prismatic joints keep q in the cos slot, so that any lane of the warp can rebuild any env's transforms)
template <class T, int I, int STRIDE>
RS_HD void cached_xform(const LinkCache<STRIDE>& lc, float* R, float* r) {
    float c = 1.0f, s = 0.0f, q = 0.0f;
    constexpr int jt = T::jtype[I];
    if constexpr (jt == RS_JOINT_REVOLUTE) { c = lc.at(lc_cs<T>() + 2 * LcSlot<T, I>::v); s = lc.at(lc_cs<T>() + 2 * LcSlot<T, I>::v + 1); }
    if constexpr (jt == RS_JOINT_PRISMATIC) q = lc.at(lc_cs<T>() + 2 * LcSlot<T, I>::v);
    link_xform<T, I>(c, s, q, R, r);
}

// ---------------------------------------------------------------------------------------------
// kinematics: fills the (cos, sin) cache and the world transforms W.Rwl / W.pos (thread-private,
// read by the narrowphase / contact rows / epilogue by dynamic body index)
// ---------------------------------------------------------------------------------------------
// world end points of the geoms attached to body B (0 = base), while that body's transform is still in registers
template <class T, int B, int NL, int MC, bool SL>
RS_HD void static_body_geoms(const Work<NL, MC, SL>& W, float (*gwr)[6]) {
    static_for<T::NG>([&](auto gc) {
        constexpr int g = decltype(gc)::value;
        if constexpr (T::g_link[g] + 1 == B) {
            const float* R = W.Rwl[B];
            const float* p = W.pos[B];
            constexpr float a0 = T::g_p0[3 * g], a1 = T::g_p0[3 * g + 1], a2 = T::g_p0[3 * g + 2];
            gwr[g][0] = p[0] + (RS_CM(a0, R[0]) + RS_CM(a1, R[3]) + RS_CM(a2, R[6]));
            gwr[g][1] = p[1] + (RS_CM(a0, R[1]) + RS_CM(a1, R[4]) + RS_CM(a2, R[7]));
            gwr[g][2] = p[2] + (RS_CM(a0, R[2]) + RS_CM(a1, R[5]) + RS_CM(a2, R[8]));
            if constexpr (T::g_type[g] == RS_GEOM_CAPSULE) {
                constexpr float b0 = T::g_p1[3 * g], b1 = T::g_p1[3 * g + 1], b2 = T::g_p1[3 * g + 2];
                gwr[g][3] = p[0] + (RS_CM(b0, R[0]) + RS_CM(b1, R[3]) + RS_CM(b2, R[6]));
                gwr[g][4] = p[1] + (RS_CM(b0, R[1]) + RS_CM(b1, R[4]) + RS_CM(b2, R[7]));
                gwr[g][5] = p[2] + (RS_CM(b0, R[2]) + RS_CM(b1, R[5]) + RS_CM(b2, R[8]));
            } else { gwr[g][3] = gwr[g][0]; gwr[g][4] = gwr[g][1]; gwr[g][5] = gwr[g][2]; }
        }
    });
}

// GEOMS: also produce the world end points of every collision geom into gwr[NG][6] (statically indexed => registers)
template <class T, bool GEOMS = false, int NL, int MC, bool SL, int STRIDE>
RS_HD void static_kinematics(Work<NL, MC, SL>& W, const LinkCache<STRIDE>& lc, float (*gwr)[6] = nullptr) {
    float Rbw[9];
    quat_to_mat(Rbw, W.bquat[0], W.bquat[1], W.bquat[2], W.bquat[3]);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) W.Rwl[0][i * 3 + j] = Rbw[j * 3 + i];
    W.pos[0][0] = W.bpos[0]; W.pos[0][1] = W.bpos[1]; W.pos[0][2] = W.bpos[2];
#pragma unroll
    for (int k = 0; k < 9; k++) lc.at(lc_rb<T>() + k) = W.Rwl[0][k];
    if constexpr (GEOMS) static_body_geoms<T, 0>(W, gwr);
    if constexpr (T::CUBE != 0) cube_kinematics(W);
    static_for<T::NL>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i], jt = T::jtype[i], dI = T::dof[i];
        float c = 1.0f, s = 0.0f, q = 0.0f;
        if constexpr (jt == RS_JOINT_REVOLUTE) {
            rs_sincos(W.q[dI], &s, &c);
            lc.at(lc_cs<T>() + 2 * LcSlot<T, i>::v) = c; lc.at(lc_cs<T>() + 2 * LcSlot<T, i>::v + 1) = s;
        }
        if constexpr (jt == RS_JOINT_PRISMATIC) { q = W.q[dI]; lc.at(lc_cs<T>() + 2 * LcSlot<T, i>::v) = q; }
        float R[9], r[3];
        link_xform<T, i>(c, s, q, R, r);
        rot_mul33<rot_kind<T, i>()>(W.Rwl[i + 1], R, W.Rwl[p + 1]);
        float rw[3];
        tmulv(rw, W.Rwl[i + 1], r);
        W.pos[i + 1][0] = W.pos[p + 1][0] + rw[0]; W.pos[i + 1][1] = W.pos[p + 1][1] + rw[1]; W.pos[i + 1][2] = W.pos[p + 1][2] + rw[2];
        if constexpr (GEOMS) static_body_geoms<T, i + 1>(W, gwr);
    });
}

// ---------------------------------------------------------------------------------------------
// narrowphase driver with a compile-time broadphase: the floor test of every geom and the bounding-sphere test of
// every self-collision pair run on registers (gwr, from static_kinematics<T, true>) and set one bit per manifold;
// only manifolds whose bit is set -- or that still hold cached points -- reach collide_manifold().
// Same rejects as collide() (they are conservative: a manifold skipped here has no point within its threshold).
// ---------------------------------------------------------------------------------------------
template <class T, int NL, int MC, bool SL>
RS_HD void static_collide(const DevModel& M, Work<NL, MC, SL>& W, const float (*gwr)[6]) {
    constexpr int NG = T::NG, NP = T::NP, NMAN = NG + NP + (T::CUBE ? 1 + NG : 0), NW = (NMAN + 31) / 32;
    uint32_t cand[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) cand[w] = 0u;
    static_for<NG>([&](auto gc) {
        constexpr int g = decltype(gc)::value;
        if constexpr (T::g_floor[g] != 0) {
            constexpr float rad = T::g_radius[g], thr = T::g_thresh[g];
            const float zmin = fminf(gwr[g][2], gwr[g][5]) - rad;
            if (zmin < thr) cand[g >> 5] |= 1u << (g & 31);
        }
    });
    if constexpr (NP > 0) {
        float sm[NG][3];   // end-point sums (2 x midpoint)
        static_for<NG>([&](auto gc) {
            constexpr int g = decltype(gc)::value;
            sm[g][0] = gwr[g][0] + gwr[g][3]; sm[g][1] = gwr[g][1] + gwr[g][4]; sm[g][2] = gwr[g][2] + gwr[g][5];
        });
        static_for<NP>([&](auto pc) {
            constexpr int pk = decltype(pc)::value;
            constexpr int ga = T::pair_a[pk], gb = T::pair_b[pk], k = NG + pk;
            constexpr float reach = T::pair_reach[pk];
            const float dx = 0.5f * (sm[ga][0] - sm[gb][0]), dy = 0.5f * (sm[ga][1] - sm[gb][1]), dz = 0.5f * (sm[ga][2] - sm[gb][2]);
            if (!(dx * dx + dy * dy + dz * dz > reach * reach)) cand[k >> 5] |= 1u << (k & 31);
        });
    }
    if constexpr (T::CUBE != 0) {
        // the cube: against the floor when its bounding sphere reaches the threshold, against a geom when the bounding spheres do
        constexpr int kf = NG + NP;
        if (W.cpos[2] - M.cube_bound < M.cube_thresh) cand[kf >> 5] |= 1u << (kf & 31);
        static_for<NG>([&](auto gc) {
            constexpr int g = decltype(gc)::value;
            constexpr int k = NG + NP + 1 + g;
            constexpr float hl0 = 0.5f * (T::g_p1[3 * g] - T::g_p0[3 * g]), hl1 = 0.5f * (T::g_p1[3 * g + 1] - T::g_p0[3 * g + 1]), hl2 = 0.5f * (T::g_p1[3 * g + 2] - T::g_p0[3 * g + 2]);
            const float bound = sqrtf(hl0 * hl0 + hl1 * hl1 + hl2 * hl2) + T::g_radius[g];
            const float dx = 0.5f * (gwr[g][0] + gwr[g][3]) - W.cpos[0], dy = 0.5f * (gwr[g][1] + gwr[g][4]) - W.cpos[1], dz = 0.5f * (gwr[g][2] + gwr[g][5]) - W.cpos[2];
            const float reach = bound * 1.0001f + 1e-6f + M.cube_bound + M.cube_thresh;
            if (!(dx * dx + dy * dy + dz * dz > reach * reach)) cand[k >> 5] |= 1u << (k & 31);
        });
    }
    const int old = W.cur, nw = 1 - W.cur;
    const int nold = W.ccount[old];
    for (int i = 0; i < nold; i++) {
        const int k = W.ckey[old][i];
#pragma unroll
        for (int w = 0; w < NW; w++) if ((k >> 5) == w) cand[w] |= 1u << (k & 31);
    }
    int oi = 0, no = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        uint32_t m = cand[w];
        while (m) {
            const int k = w * 32 + rs_ctz(m);
            m &= m - 1u;
            collide_manifold<T::CUBE != 0>(M, W, k, old, nw, nold, oi, no);
        }
    }
    W.ccount[nw] = no;
    W.cur = nw;
}

// ---------------------------------------------------------------------------------------------
// ABA (computeAccelerationsArticulatedBodyAlgorithmMultiDof + applyDeltaVeeMultiDof)
// ---------------------------------------------------------------------------------------------
template <class T, int NL, int MC, bool SL, int STRIDE>
RS_HD void static_aba(const DevModel& M, Work<NL, MC, SL>& W, const LinkCache<STRIDE>& lc, float dt) {
    constexpr int n = T::NL;
    float vel[n + 1][6], Z[n + 1][6], c[n][6], Y[n];
    // ---- root -> leaves: velocities, bias forces, Coriolis ----
    if constexpr (T::FIXED_BASE) {
#pragma unroll
        for (int k = 0; k < 6; k++) { vel[0][k] = 0.f; Z[0][k] = 0.f; }
    } else {
        mulv(vel[0], W.Rwl[0], W.bw); mulv(vel[0] + 3, W.Rwl[0], W.bv);
    }
    auto bias = [&](const float* v, float mass, const float* I3, const float* Rwl, float* Zb) {
        // gravity as an applied force rotated into link coords: only the z column of Rwl is needed
        float g = -M.gravity * mass;
        float fl[3] = {Rwl[2] * g, Rwl[5] * g, Rwl[8] * g};
        float wn = sqrtf(dot3(v, v)), vn = sqrtf(dot3(v + 3, v + 3));
        float Iw[3] = {I3[0] * v[0], I3[1] * v[1], I3[2] * v[2]};
        float ka = M.ang_damping * (1.0f + wn), kl = M.lin_damping * (1.0f + vn);
        float g1[3], g2[3];
        cross3(g1, v, Iw); cross3(g2, v, v + 3);
        float gy = M.use_gyro;
#pragma unroll
        for (int k = 0; k < 3; k++) { Zb[k] = Iw[k] * ka + gy * g1[k]; Zb[3 + k] = -fl[k] + mass * v[3 + k] * kl + mass * g2[k]; }
    };
    constexpr float bI0 = T::base_inertia[0], bI1 = T::base_inertia[1], bI2 = T::base_inertia[2], bM = T::base_mass;
    if constexpr (!T::FIXED_BASE) {
        const float I3[3] = {bI0, bI1, bI2};
        bias(vel[0], bM, I3, W.Rwl[0], Z[0]);
    }
    static_for<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i];
        constexpr int dI = T::dof[i];
        constexpr float mI = T::mass[i], i0 = T::inertia[3 * i], i1 = T::inertia[3 * i + 1], i2 = T::inertia[3 * i + 2];
        float R[9], r[3];
        cached_xform<T, i>(lc, R, r);
        xf_motion_k<rot_kind<T, i>()>(vel[i + 1], R, r, vel[p + 1]);
        if constexpr (dI >= 0) {
            float qd = W.qd[dI];
            saxpy<T, i>(vel[i + 1], qd);
            // c = v x (S qd)
            float jv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            saxpy<T, i>(jv, qd);
            const float* v = vel[i + 1];
            float t1[3], t2[3];
            cross3(c[i], v, jv);
            cross3(t1, v, jv + 3); cross3(t2, v + 3, jv);
            c[i][3] = t1[0] + t2[0]; c[i][4] = t1[1] + t2[1]; c[i][5] = t1[2] + t2[2];
        }
        if constexpr (mI == 0.0f && i0 == 0.0f && i1 == 0.0f && i2 == 0.0f) {
            // massless link (importer-generated dummy of a multi-joint body): no gravity, damping or gyroscopic force
#pragma unroll
            for (int k = 0; k < 6; k++) Z[i + 1][k] = 0.f;
        } else {
            const float I3[3] = {i0, i1, i2};
            bias(vel[i + 1], mI, I3, W.Rwl[i + 1], Z[i + 1]);
        }
    });
    // ---- leaves -> root: articulated inertias, h, D, Y ----
    float IA[n + 1][21];
    auto init_rigid = [&](float* I, float mass, float ix, float iy, float iz) {
#pragma unroll
        for (int k = 0; k < 21; k++) I[k] = 0.f;
        I[0] = ix; I[3] = iy; I[5] = iz; I[15] = mass; I[18] = mass; I[20] = mass;
    };
    static_rfor<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i], dI = T::dof[i], pp = p >= 0 ? p : 0;
        constexpr bool is_leaf = T::leaf[i] != 0, is_first = T::first_fold[i] != 0;
        constexpr float mI = T::mass[i], i0 = T::inertia[3 * i], i1 = T::inertia[3 * i + 1], i2 = T::inertia[3 * i + 2];
        constexpr float mP = T::mass[pp], p0 = T::inertia[3 * pp], p1 = T::inertia[3 * pp + 1], p2 = T::inertia[3 * pp + 2];
        if constexpr (is_leaf) init_rigid(IA[i + 1], mI, i0, i1, i2);
        float R[9], r[3];
        cached_xform<T, i>(lc, R, r);
        float Zt[6];
        if constexpr (dI >= 0) {
            float h[6];
            inertia_mul_S<T, i>(h, IA[i + 1]);
            float D = sdot<T, i>(h) + M.armature[i];
            float y = W.tau[dI] - sdot<T, i>(Z[i + 1]) - dot6(c[i], h);
            float invD = 1.0f / D;
            Y[i] = y;
            lc.at(lc_invd<T>() + LcSlot<T, i>::v) = invD;
#pragma unroll
            for (int k = 0; k < 6; k++) lc.at(lc_h<T>() + 6 * LcSlot<T, i>::v + k) = h[k];
            float Ic[6];
            inertia_mul(Ic, IA[i + 1], c[i]);
            float kk = invD * y;
#pragma unroll
            for (int k = 0; k < 6; k++) Zt[k] = Z[i + 1][k] + Ic[k] + h[k] * kk;
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
#pragma unroll
            for (int k = 0; k < 6; k++) Zt[k] = Z[i + 1][k];
        }
        // first child to fold initialises the parent's accumulator with the parent's rigid-body inertia
        if constexpr (is_first) {
            if constexpr (p >= 0) init_rigid(IA[p + 1], mP, p0, p1, p2);
            else init_rigid(IA[0], bM, bI0, bI1, bI2);
        }
        inertia_to_parent_add_k<rot_kind<T, i>()>(IA[p + 1], IA[i + 1], R, r);
        xf_force_inv_add_k<rot_kind<T, i>()>(Z[p + 1], R, r, Zt);
    });
    // ---- base acceleration + root -> leaves joint accelerations ----
    float acc[n + 1][6];
    if constexpr (T::FIXED_BASE) {
#pragma unroll
        for (int k = 0; k < 6; k++) acc[0][k] = 0.f;
    } else {
        if constexpr (T::base_leaf != 0) init_rigid(IA[0], bM, bI0, bI1, bI2);
        float Lb[21], rr[6];
        chol6(Lb, IA[0]);
#pragma unroll
        for (int k = 0; k < 21; k++) lc.at(lc_lb<T>() + k) = Lb[k];
        chol6_solve(rr, Lb, Z[0]);
#pragma unroll
        for (int k = 0; k < 6; k++) acc[0][k] = -rr[k];
    }
    static_for<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i];
        float R[9], r[3];
        cached_xform<T, i>(lc, R, r);
        xf_motion_k<rot_kind<T, i>()>(acc[i + 1], R, r, acc[p + 1]);
        constexpr int dI = T::dof[i];
        if constexpr (dI >= 0) {
            float h[6];
#pragma unroll
            for (int k = 0; k < 6; k++) h[k] = lc.at(lc_h<T>() + 6 * LcSlot<T, i>::v + k);
            float a = lc.at(lc_invd<T>() + LcSlot<T, i>::v) * (Y[i] - dot6(acc[i + 1], h));
#pragma unroll
            for (int k = 0; k < 6; k++) acc[i + 1][k] += c[i][k];
            saxpy<T, i>(acc[i + 1], a);
            W.qd[dI] += dt * a;
        }
    });
    if constexpr (!T::FIXED_BASE) {
        float wd[3], vd[3], t[3], wxv[3];
        tmulv(wd, W.Rwl[0], acc[0]);
        cross3(wxv, vel[0], vel[0] + 3);
        t[0] = acc[0][3] + wxv[0]; t[1] = acc[0][4] + wxv[1]; t[2] = acc[0][5] + wxv[2];
        tmulv(vd, W.Rwl[0], t);
#pragma unroll
        for (int k = 0; k < 3; k++) { W.bw[k] += dt * wd[k]; W.bv[k] += dt * vd[k]; }
    }
    clamp_vel(M, W);
}

// ---------------------------------------------------------------------------------------------
// observation with the static sweeps (the constraint-row responses live in rs_pooled.cuh)
// ---------------------------------------------------------------------------------------------
template <class T, int NL, int MC, bool SL, int STRIDE>
RS_HD void static_calc_state(const DevModel& M, Work<NL, MC, SL>& W, const LinkCache<STRIDE>& lc, Episode& E, float* obs, int ostride,
                             double* dist, float* pitch, int* at_limit) {
    constexpr int n = T::NL;
    static_kinematics<T>(W, lc);
    float vel[n + 1][6];
    if constexpr (T::FIXED_BASE) {
#pragma unroll
        for (int k = 0; k < 6; k++) vel[0][k] = 0.f;
    } else { mulv(vel[0], W.Rwl[0], W.bw); mulv(vel[0] + 3, W.Rwl[0], W.bv); }
    const int body = M.body_link + 1;
    float vl[3] = {vel[0][3], vel[0][4], vel[0][5]};
    static_for<n>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int p = T::parent[i];
        float R[9], r[3];
        cached_xform<T, i>(lc, R, r);
        xf_motion_k<rot_kind<T, i>()>(vel[i + 1], R, r, vel[p + 1]);
        constexpr int dI = T::dof[i];
        if constexpr (dI >= 0) saxpy<T, i>(vel[i + 1], W.qd[dI]);
        if (body == i + 1) { vl[0] = vel[i + 1][3]; vl[1] = vel[i + 1][4]; vl[2] = vel[i + 1][5]; }
    });
    float bp[3] = {W.pos[body][0], W.pos[body][1], W.pos[body][2]}, bq[4], bv[3];
    if (body == 0) { bq[0] = W.bquat[0]; bq[1] = W.bquat[1]; bq[2] = W.bquat[2]; bq[3] = W.bquat[3]; }
    else {
        float Rlw[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rlw[i * 3 + j] = W.Rwl[body][j * 3 + i];
        mat_to_quat(Rlw, bq);
    }
    tmulv(bv, W.Rwl[body], vl);
    calc_state_core(M, W, E, bp, bq, bv, obs, ostride, dist, pitch, at_limit);
}

// true iff the run-time descriptor is exactly the model this topology was generated from
template <class T>
inline bool topo_matches(const DevModel& D) {
    if (D.n_links != T::NL || D.n_dof != T::ND || D.fixed_base != T::FIXED_BASE || (D.has_cube != 0) != (T::CUBE != 0)) return false;
    auto eq = [](float a, float b) { return fabsf(a - b) <= 1e-6f * fmaxf(1.0f, fabsf(b)); };
    for (int i = 0; i < T::NL; i++) {
        if (D.parent[i] != T::parent[i] || D.jtype[i] != T::jtype[i] || D.dof_idx[i] != T::dof[i]) return false;
        if (!eq(D.mass[i], T::mass[i])) return false;
        for (int k = 0; k < 3; k++) {
            if (!eq(D.axis[i][k], T::axis[3 * i + k]) || !eq(D.e_vec[i][k], T::e[3 * i + k]) || !eq(D.d_vec[i][k], T::d[3 * i + k]) ||
                !eq(D.Sw[i][k], T::Sw[3 * i + k]) || !eq(D.Sv[i][k], T::Sv[3 * i + k]) || !eq(D.inertia[i][k], T::inertia[3 * i + k])) return false;
        }
        for (int k = 0; k < 9; k++) if (!eq(D.R0[i][k], T::R0[9 * i + k])) return false;
    }
    if (D.n_geoms != T::NG || D.n_pairs != T::NP) return false;
    for (int g = 0; g < T::NG; g++) {
        if (D.g_link[g] != T::g_link[g] || D.g_type[g] != T::g_type[g] || (D.g_floor[g] != 0) != (T::g_floor[g] != 0)) return false;
        if (!eq(D.g_radius[g], T::g_radius[g]) || !eq(D.break_thresh[D.g_link[g] + 1], T::g_thresh[g])) return false;
        for (int k = 0; k < 3; k++) if (!eq(D.g_p0[g][k], T::g_p0[3 * g + k]) || !eq(D.g_p1[g][k], T::g_p1[3 * g + k])) return false;
    }
    for (int k = 0; k < T::NP; k++) {
        if (D.pair_a[k] != T::pair_a[k] || D.pair_b[k] != T::pair_b[k]) return false;
        if (!eq(D.g_bound[D.pair_a[k]] + D.g_bound[D.pair_b[k]] + D.pair_thresh[k], T::pair_reach[k])) return false;
    }
    if (!T::FIXED_BASE) {
        if (!eq(D.base_mass, T::base_mass)) return false;
        for (int k = 0; k < 3; k++) if (!eq(D.base_inertia[k], T::base_inertia[k])) return false;
    }
    return true;
}

}  // namespace rs
