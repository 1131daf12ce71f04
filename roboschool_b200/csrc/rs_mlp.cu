// rs_mlp.cu -- agent_zoo v1 policy forward: relu(relu(x W1 + b1) W2 + b2) W3 + b3
// (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34,63-68), batched over envs on the tensor cores.
//
// The default kernel is the tcgen05 / TMEM one in rs_mlp_tc5.cu.  This file holds the policy object, the C ABI and the
// round-1 mma.sync kernel, kept selectable (RS_MLP_TC5=0) as the measured baseline the tcgen05 kernel is compared with.
//
// mma.sync kernel:
// Design: persistent CTAs (one per SM).  The whole weight set is staged ONCE per CTA in shared memory as
// fp16, transposed (N-major) so a B fragment is one conflict-free 32-bit load.  Each warp owns 16 envs and
// runs them through the three layers with mma.sync.m16n8k16 (fp16 in, fp32 accumulate).  The accumulator
// fragment of layer L (row g, cols 2t,2t+1) is exactly the A-fragment layout of layer L+1, so hidden
// activations never leave the register file: no shared-memory round trip, no HBM traffic beyond
// obs in (4*O B/env) and actions out (4*A B/env).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/roboschool_b200.h"
#include "rs_mlp_tc5.h"
#include <stdlib.h>

int rs_set_error(const char* s);
static int mfail(const std::string& s) { rs_set_error(s.c_str()); return 1; }
#define MCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return mfail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct rs_policy {
    int in_dim, h1, h2, out_dim, device;
    int in_p, out_p, variant;
    void* w;                 // packed fp16 transposed weights + fp32 biases (device)
    size_t bytes;
    // tcgen05 / TMEM kernel (rs_mlp_tc5.cu): operand-layout weight image
    int tc5_variant, n_sm, use_tc5;
    void* tc5_image;
    size_t tc5_smem;
};

__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// One layer on a 16-row warp tile.  A: K/16 k-slabs x 4 regs (fp16x2).  Wt: smem fp16 [N][K+8].
// OUT_FRAG: write relu(result) packed as the next layer's A fragments; else raw fp32 accumulators.
template <int K, int N, bool RELU>
__device__ __forceinline__ void layer_frag(const uint32_t (&A)[K / 16][4], const __half* Wt, const float* bias,
                                           uint32_t (&Aout)[N / 16][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int LDK = K + 8;
#pragma unroll
    for (int n0 = 0; n0 < N; n0 += 8) {
        float c[4];
        c[0] = c[2] = bias[n0 + 2 * t];
        c[1] = c[3] = bias[n0 + 2 * t + 1];
        const uint32_t* wrow = reinterpret_cast<const uint32_t*>(Wt + (size_t)(n0 + g) * LDK);
#pragma unroll
        for (int ks = 0; ks < K / 16; ks++) {
            uint32_t b0 = wrow[ks * 8 + t], b1 = wrow[ks * 8 + 4 + t];
            mma_f16(c, A[ks], b0, b1);
        }
        if (RELU) { c[0] = fmaxf(c[0], 0.f); c[1] = fmaxf(c[1], 0.f); c[2] = fmaxf(c[2], 0.f); c[3] = fmaxf(c[3], 0.f); }
        // C tile n0 (8 cols) -> half of k-slab n0/16 of the next layer: even tile -> regs 0,1 ; odd tile -> regs 2,3
        Aout[n0 / 16][(n0 / 8) % 2 * 2 + 0] = pack_h2(c[0], c[1]);
        Aout[n0 / 16][(n0 / 8) % 2 * 2 + 1] = pack_h2(c[2], c[3]);
    }
}

template <int IN_P, int H1, int H2, int OUT_P>
struct MlpLayout {
    static constexpr size_t w1 = 0;                                   // halves
    static constexpr size_t w2 = w1 + (size_t)H1 * (IN_P + 8);
    static constexpr size_t w3 = w2 + (size_t)H2 * (H1 + 8);
    static constexpr size_t n_half = w3 + (size_t)OUT_P * (H2 + 8);
    static constexpr size_t bias_off = (n_half * 2 + 15) / 16 * 16;   // bytes
    static constexpr size_t bytes = bias_off + sizeof(float) * (H1 + H2 + OUT_P);
};

constexpr int MLP_WARPS = 8;

template <int IN_P, int H1, int H2, int OUT_P>
__global__ void __launch_bounds__(32 * MLP_WARPS) k_mlp(const unsigned char* __restrict__ wpack, const float* __restrict__ obs,
                                                        float* __restrict__ act, int n, int in_dim, int out_dim) {
    using L = MlpLayout<IN_P, H1, H2, OUT_P>;
    extern __shared__ __align__(16) unsigned char smem[];
    {   // stage weights: 16-byte vector copies, coalesced
        const uint4* src = reinterpret_cast<const uint4*>(wpack);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (size_t i = threadIdx.x; i < L::bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const __half* sW = reinterpret_cast<const __half*>(smem);
    const float* sB = reinterpret_cast<const float*>(smem + L::bias_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n_tiles = (n + 15) / 16;
    for (int tile = blockIdx.x * MLP_WARPS + warp; tile < n_tiles; tile += gridDim.x * MLP_WARPS) {
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        uint32_t X[IN_P / 16][4];
#pragma unroll
        for (int ks = 0; ks < IN_P / 16; ks++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int c = ks * 16 + h * 8 + 2 * t;
                float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
                if (r0 < n) { if (c < in_dim) v00 = obs[(size_t)r0 * in_dim + c]; if (c + 1 < in_dim) v01 = obs[(size_t)r0 * in_dim + c + 1]; }
                if (r1 < n) { if (c < in_dim) v10 = obs[(size_t)r1 * in_dim + c]; if (c + 1 < in_dim) v11 = obs[(size_t)r1 * in_dim + c + 1]; }
                X[ks][h * 2 + 0] = pack_h2(v00, v01);
                X[ks][h * 2 + 1] = pack_h2(v10, v11);
            }
        }
        uint32_t Hd1[H1 / 16][4], Hd2[H2 / 16][4];
        layer_frag<IN_P, H1, true>(X, sW + L::w1, sB, Hd1, lane);
        layer_frag<H1, H2, true>(Hd1, sW + L::w2, sB + H1, Hd2, lane);
        // last layer: fp32 accumulators straight to global
        constexpr int LDK = H2 + 8;
#pragma unroll
        for (int n0 = 0; n0 < OUT_P; n0 += 8) {
            float c[4];
            c[0] = c[2] = sB[H1 + H2 + n0 + 2 * t];
            c[1] = c[3] = sB[H1 + H2 + n0 + 2 * t + 1];
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(sW + L::w3 + (size_t)(n0 + g) * LDK);
#pragma unroll
            for (int ks = 0; ks < H2 / 16; ks++) mma_f16(c, Hd2[ks], wrow[ks * 8 + t], wrow[ks * 8 + 4 + t]);
            int col = n0 + 2 * t;
            if (r0 < n) { if (col < out_dim) act[(size_t)r0 * out_dim + col] = c[0]; if (col + 1 < out_dim) act[(size_t)r0 * out_dim + col + 1] = c[1]; }
            if (r1 < n) { if (col < out_dim) act[(size_t)r1 * out_dim + col] = c[2]; if (col + 1 < out_dim) act[(size_t)r1 * out_dim + col + 1] = c[3]; }
        }
    }
}

static inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

template <int IN_P, int H1, int H2, int OUT_P>
static int create_variant(rs_policy* p, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3) {
    using L = MlpLayout<IN_P, H1, H2, OUT_P>;
    std::vector<unsigned char> pack(L::bytes, 0);
    __half* hw = reinterpret_cast<__half*>(pack.data());
    float* hb = reinterpret_cast<float*>(pack.data() + L::bias_off);
    // transposed: Wt[n][k]
    for (int n = 0; n < p->h1; n++) for (int k = 0; k < p->in_dim; k++) hw[L::w1 + (size_t)n * (IN_P + 8) + k] = __float2half_rn(w1[(size_t)k * p->h1 + n]);
    for (int n = 0; n < p->h2; n++) for (int k = 0; k < p->h1; k++) hw[L::w2 + (size_t)n * (H1 + 8) + k] = __float2half_rn(w2[(size_t)k * p->h2 + n]);
    for (int n = 0; n < p->out_dim; n++) for (int k = 0; k < p->h2; k++) hw[L::w3 + (size_t)n * (H2 + 8) + k] = __float2half_rn(w3[(size_t)k * p->out_dim + n]);
    for (int j = 0; j < p->h1; j++) hb[j] = b1[j];
    for (int j = 0; j < p->h2; j++) hb[H1 + j] = b2[j];
    for (int j = 0; j < p->out_dim; j++) hb[H1 + H2 + j] = b3[j];
    MCK(cudaMalloc(&p->w, L::bytes));
    MCK(cudaMemcpy(p->w, pack.data(), L::bytes, cudaMemcpyHostToDevice));
    p->bytes = L::bytes;
    MCK(cudaFuncSetAttribute(k_mlp<IN_P, H1, H2, OUT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
    return 0;
}

extern "C" int rs_policy_create(int in_dim, int h1, int h2, int out_dim, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, int device, rs_policy** out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return mfail("rs_policy_create: no CUDA device (there is no CPU fallback)");
    MCK(cudaSetDevice(device));
    // Narrower hidden layers (the May-2017 pendulum policies are 64-32, agent_zoo/RoboschoolInvertedPendulum_v0_2017may.py:8-10)
    // run on the next wider kernel shape with zero-padded weights: relu(0) = 0 and the extra products are exact zeros, so
    // the result is unchanged.
    std::vector<float> pw1, pb1, pw2, pb2, pw3;
    {
        int H1 = 0, H2 = 0;
        if (h1 <= 128 && h2 <= 64) { H1 = 128; H2 = 64; } else if (h1 <= 256 && h2 <= 128) { H1 = 256; H2 = 128; }
        if (H1 && (H1 != h1 || H2 != h2) && h1 > 0 && h2 > 0) {
            pw1.assign((size_t)in_dim * H1, 0.f); pb1.assign(H1, 0.f); pw2.assign((size_t)H1 * H2, 0.f); pb2.assign(H2, 0.f);
            pw3.assign((size_t)H2 * out_dim, 0.f);
            for (int i = 0; i < in_dim; i++) for (int j = 0; j < h1; j++) pw1[(size_t)i * H1 + j] = w1[(size_t)i * h1 + j];
            for (int j = 0; j < h1; j++) pb1[j] = b1[j];
            for (int i = 0; i < h1; i++) for (int j = 0; j < h2; j++) pw2[(size_t)i * H2 + j] = w2[(size_t)i * h2 + j];
            for (int j = 0; j < h2; j++) pb2[j] = b2[j];
            for (int i = 0; i < h2; i++) for (int j = 0; j < out_dim; j++) pw3[(size_t)i * out_dim + j] = w3[(size_t)i * out_dim + j];
            w1 = pw1.data(); b1 = pb1.data(); w2 = pw2.data(); b2 = pb2.data(); w3 = pw3.data();
            h1 = H1; h2 = H2;
        }
    }
    rs_policy* p = new rs_policy();
    p->in_dim = in_dim; p->h1 = h1; p->h2 = h2; p->out_dim = out_dim; p->device = device; p->w = nullptr;
    p->in_p = pad_to(in_dim, 16); p->out_p = pad_to(out_dim, 8);
    int rc = 1;
    if (h1 == 256 && h2 == 128 && p->in_p == 48 && p->out_p == 24) { p->variant = 0; rc = create_variant<48, 256, 128, 24>(p, w1, b1, w2, b2, w3, b3); }
    else if (h1 == 128 && h2 == 64 && p->in_p == 16 && p->out_p == 8) { p->variant = 1; rc = create_variant<16, 128, 64, 8>(p, w1, b1, w2, b2, w3, b3); }
    else if (h1 == 128 && h2 == 64 && p->in_p == 32 && p->out_p == 8) { p->variant = 2; rc = create_variant<32, 128, 64, 8>(p, w1, b1, w2, b2, w3, b3); }
    else { delete p; return mfail("rs_policy_create: unsupported layer sizes (agent_zoo v1 shapes only: 44-256-128-17, {15,22,26,28}-128-64-{3,6,8})"); }
    if (rc) { delete p; return 1; }
    p->tc5_variant = rs_tc5_supported(in_dim, h1, h2, out_dim);
    p->tc5_image = nullptr; p->use_tc5 = 0;
    if (p->tc5_variant) {
        cudaDeviceProp prop;
        MCK(cudaGetDeviceProperties(&prop, device));
        p->n_sm = prop.multiProcessorCount;
        if (rs_tc5_pack(p->tc5_variant, in_dim, out_dim, w1, b1, w2, b2, w3, b3, &p->tc5_image, &p->tc5_smem)) { cudaGetLastError(); p->tc5_variant = 0; }
        const char* e = getenv("RS_MLP_TC5");       // RS_MLP_TC5=0 selects the mma.sync kernel below (A/B measurements)
        p->use_tc5 = p->tc5_variant && !(e && atoi(e) == 0);
    }
    *out = p;
    return 0;
}

extern "C" int rs_policy_destroy(rs_policy* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaFree(p->w);
    if (p->tc5_image) cudaFree(p->tc5_image);
    delete p;
    return 0;
}

extern "C" int rs_policy_forward(rs_policy* p, const float* obs_dev, float* act_dev, int n, void* stream) {
    if (!p || !obs_dev || !act_dev) return mfail("rs_policy_forward: null argument");
    MCK(cudaSetDevice(p->device));
    int tiles = (n + 16 * MLP_WARPS - 1) / (16 * MLP_WARPS);
    int grid = tiles < 148 ? tiles : 148;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->use_tc5) {
        if (rs_tc5_forward(p->tc5_variant, p->tc5_image, p->tc5_smem, obs_dev, act_dev, n, p->in_dim, p->out_dim, p->n_sm, st))
            return mfail("rs_policy_forward: tcgen05 kernel launch failed");
        return 0;
    }
    const unsigned char* w = (const unsigned char*)p->w;
    switch (p->variant) {
        case 0: k_mlp<48, 256, 128, 24><<<grid, 32 * MLP_WARPS, p->bytes, st>>>(w, obs_dev, act_dev, n, p->in_dim, p->out_dim); break;
        case 1: k_mlp<16, 128, 64, 8><<<grid, 32 * MLP_WARPS, p->bytes, st>>>(w, obs_dev, act_dev, n, p->in_dim, p->out_dim); break;
        default: k_mlp<32, 128, 64, 8><<<grid, 32 * MLP_WARPS, p->bytes, st>>>(w, obs_dev, act_dev, n, p->in_dim, p->out_dim); break;
    }
    MCK(cudaGetLastError());
    return 0;
}
