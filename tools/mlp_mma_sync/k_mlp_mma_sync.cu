// k_mlp_mma_sync.cu -- the round-1 policy kernel (mma.sync.m16n8k16, fp16 in / fp32 accumulate), kept OUT of the shipped
// library as the measured baseline the tcgen05 kernel (roboschool_b200/csrc/rs_mlp_tc5.cu) replaced: 18.5 us against
// 16.9 us for 32768 Humanoid envs (profiles/r01_v6_mlp_tc5_summary.md).  Stand-alone: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o bench_mma k_mlp_mma_sync.cu && ./bench_mma
//
// Design: persistent CTAs (one per SM).  The whole weight set is staged ONCE per CTA in shared memory as fp16, transposed
// (N-major) so a B fragment is one conflict-free 32-bit load.  Each warp owns 16 envs and runs them through the three layers;
// the accumulator fragment of layer L (row g, cols 2t,2t+1) is exactly the A-fragment layout of layer L+1, so hidden
// activations never leave the register file.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

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


int main() {
    constexpr int IN_P = 48, H1 = 256, H2 = 128, OUT_P = 24, in_dim = 44, out_dim = 17, n = 32768;
    using L = MlpLayout<IN_P, H1, H2, OUT_P>;
    std::vector<unsigned char> pack(L::bytes, 0);
    __half* hw = reinterpret_cast<__half*>(pack.data());
    for (size_t k = 0; k < L::bias_off / sizeof(__half); k++) hw[k] = __float2half_rn(0.01f * (float)((k * 37) % 19 - 9));
    void* w; float *obs, *act;
    cudaMalloc(&w, L::bytes); cudaMemcpy(w, pack.data(), L::bytes, cudaMemcpyHostToDevice);
    cudaMalloc(&obs, sizeof(float) * n * in_dim); cudaMemset(obs, 0, sizeof(float) * n * in_dim);
    cudaMalloc(&act, sizeof(float) * n * out_dim);
    cudaFuncSetAttribute(k_mlp<IN_P, H1, H2, OUT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
    int tiles = (n + 16 * MLP_WARPS - 1) / (16 * MLP_WARPS), grid = tiles < 148 ? tiles : 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 20; it++) k_mlp<IN_P, H1, H2, OUT_P><<<grid, 32 * MLP_WARPS, L::bytes>>>((const unsigned char*)w, obs, act, n, in_dim, out_dim);
    cudaEventRecord(e0);
    for (int it = 0; it < 100; it++) k_mlp<IN_P, H1, H2, OUT_P><<<grid, 32 * MLP_WARPS, L::bytes>>>((const unsigned char*)w, obs, act, n, in_dim, out_dim);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("k_mlp (mma.sync) %d envs: %.2f us per launch (%s)\n", n, 10.f * ms, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
