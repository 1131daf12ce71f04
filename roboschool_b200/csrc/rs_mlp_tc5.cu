// rs_mlp_tc5.cu -- agent_zoo v1 policy forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   act = relu(relu(obs W1 + b1) W2 + b2) W3 + b3        (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34,63-68)
//
// One persistent CTA per SM, 128 threads; a tile is 128 envs (UMMA M = 128, one env per TMEM lane / per thread).
// The three layers are chained on chip:
//   obs tile -> fp16 -> smem A1 [128 x 64]            (K-major, 128-byte swizzle: the canonical UMMA operand layout)
//   tcgen05.mma  D1[128 x H1] (TMEM)  = A1 . W1^T     (one elected thread issues K/16 instructions)
//   tcgen05.ld   D1 -> registers, + b1, relu, -> fp16 -> smem A2 [128 x H1]
//   tcgen05.mma  D2[128 x H2]         = A2 . W2^T ;  epilogue -> smem A3 [128 x H2]
//   tcgen05.mma  D3[128 x OUTP]       = A3 . W3^T ;  epilogue + b3 -> global act
// Hidden activations never leave the SM; the weights are staged once per CTA, already in operand layout
// (the host packs them swizzled, so staging is one linear TMA bulk copy, cp.async.bulk, overlapped with the first tile's load).  Completion of each MMA group is signalled with
// tcgen05.commit on an mbarrier; generic-proxy writes of the next A operand are ordered before the async-proxy reads
// with fence.proxy.async.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "rs_mlp_tc5.h"

namespace tc5 {

constexpr int TILE_M = 128;       // envs per tile = UMMA M
constexpr int K1P = 64;           // padded input width (one 128-byte swizzle atom of fp16)
constexpr int ATOM_K = 64;        // fp16 elements per 128-byte row

// byte offset of element (row r, column c) in a K-major SWIZZLE_128B operand of `rows` rows: 64-column atoms of
// rows x 128 B; inside an atom the 16-byte chunk index is XORed with (r mod 8)        (cute Swizzle<3,4,3>)
__host__ __device__ inline uint32_t sw128_offset(int rows, int r, int c) {
    const int atom = c / ATOM_K, cc = (c % ATOM_K) >> 3, e = c & 7;
    return (uint32_t)(atom * rows * 128 + r * 128 + ((cc ^ (r & 7)) << 4) + e * 2);
}

template <int H1, int H2, int OUTP>
struct Layout {
    // weights first (one contiguous image copied from global), then the activation operands
    static constexpr uint32_t b1 = 0;                                   // W1^T : H1 rows x K1P
    static constexpr uint32_t b2 = b1 + H1 * 128 * (K1P / ATOM_K);      // W2^T : H2 rows x H1
    static constexpr uint32_t b3 = b2 + H2 * 128 * (H1 / ATOM_K);       // W3^T : OUTP rows x H2
    static constexpr uint32_t bias = b3 + OUTP * 128 * (H2 / ATOM_K);   // fp32 [H1 + H2 + OUTP]
    static constexpr uint32_t image_bytes = (bias + 4 * (H1 + H2 + OUTP) + 1023) / 1024 * 1024;
    static constexpr uint32_t a1 = image_bytes;                         // 128 x K1P
    static constexpr uint32_t a2 = a1 + TILE_M * 128 * (K1P / ATOM_K);  // 128 x H1
    static constexpr uint32_t a3 = a2 + TILE_M * 128 * (H1 / ATOM_K);   // 128 x H2
    static constexpr uint32_t bar = a3 + TILE_M * 128 * (H2 / ATOM_K);  // MMA mbarrier (8 B), weight mbarrier (8 B), TMEM base (4 B)
    static constexpr uint32_t bytes = bar + 32;
    static constexpr int tmem_cols_used = H1 + H2 + OUTP;
    static constexpr int tmem_cols = tmem_cols_used <= 32 ? 32 : tmem_cols_used <= 64 ? 64 : tmem_cols_used <= 128 ? 128 : tmem_cols_used <= 256 ? 256 : 512;
    static_assert(H1 % 64 == 0 && H2 % 64 == 0 && OUTP % 16 == 0 && OUTP <= 32, "layer widths");
    static_assert(tmem_cols_used <= 512, "TMEM columns");
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: A,B fp16 K-major, D fp32, M x N               (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n"
                 :: "r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// D[128 x N] (TMEM column tmem_d) = A[128 x K] . B[N x K]^T, both K-major SW128 in shared memory; issued by one thread
template <int N, int K>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t bar) {
    const uint32_t idesc = umma_idesc(TILE_M, N);
#pragma unroll
    for (int k = 0; k < K / 16; k++) {
        const int atom = (k * 16) / ATOM_K, kin = (k * 16) % ATOM_K;
        const uint64_t ad = umma_desc(a_base + atom * (TILE_M * 128) + kin * 2);
        const uint64_t bd = umma_desc(b_base + atom * (N * 128) + kin * 2);
        umma_f16(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar);
}

// accumulators of this thread's row -> + bias, relu -> fp16 -> next A operand (row `row`, columns 0..N-1)
template <int N>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_row, const float* bias, unsigned char* a_next, int row) {
#pragma unroll 2
    for (int j = 0; j < N / 32; j++) {
        float v[32];
        tmem_ld32(tmem_row + j * 32, v);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 w;
            const int c0 = j * 32 + q * 8;
            w.x = pack_h2(fmaxf(v[q * 8 + 0] + bias[c0 + 0], 0.f), fmaxf(v[q * 8 + 1] + bias[c0 + 1], 0.f));
            w.y = pack_h2(fmaxf(v[q * 8 + 2] + bias[c0 + 2], 0.f), fmaxf(v[q * 8 + 3] + bias[c0 + 3], 0.f));
            w.z = pack_h2(fmaxf(v[q * 8 + 4] + bias[c0 + 4], 0.f), fmaxf(v[q * 8 + 5] + bias[c0 + 5], 0.f));
            w.w = pack_h2(fmaxf(v[q * 8 + 6] + bias[c0 + 6], 0.f), fmaxf(v[q * 8 + 7] + bias[c0 + 7], 0.f));
            *reinterpret_cast<uint4*>(a_next + sw128_offset(TILE_M, row, c0)) = w;
        }
    }
}
#endif  // __CUDA_ARCH__

template <int H1, int H2, int OUTP>
__global__ void __launch_bounds__(TILE_M, 1) k_mlp_tc5(const unsigned char* __restrict__ image, const float* __restrict__ obs,
                                                       float* __restrict__ act, int n, int in_dim, int out_dim) {
#if defined(__CUDA_ARCH__)
    using L = Layout<H1, H2, OUTP>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(smem + L::bar), wbar = bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::bar + 16);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // stage the weight image (already in operand layout) + biases with the TMA bulk engine; it lands while the
        // first observation tile is being converted
        mbar_expect_tx(wbar, L::image_bytes);
        constexpr uint32_t CH = 32768;
#pragma unroll
        for (uint32_t off = 0; off < L::image_bytes; off += CH)
            bulk_g2s(smem_u32(smem) + off, image + off, (L::image_bytes - off) < CH ? (L::image_bytes - off) : CH, wbar);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)L::tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);     // this warp's 32 lanes
    const float* sbias = reinterpret_cast<const float*>(smem + L::bias);
    unsigned char* sA1 = smem + L::a1;
    unsigned char* sA2 = smem + L::a2;
    unsigned char* sA3 = smem + L::a3;
    const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2), a3 = smem_u32(sA3);
    const uint32_t b1 = smem_u32(smem + L::b1), b2 = smem_u32(smem + L::b2), b3 = smem_u32(smem + L::b3);
    uint32_t phase = 0;
    bool first_tile = true;
    const int n_tiles = (n + TILE_M - 1) / TILE_M;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = tid, env = tile * TILE_M + row;
        // ---- obs row -> fp16 -> A1 (one row per thread; columns >= in_dim are zero).  Rows of a tile are adjacent in
        // global memory, so the 128-bit loads of neighbouring threads share cache lines ----
        {
            const float* orow = obs + (size_t)env * in_dim;
            const bool vec = (in_dim & 3) == 0;       // rows 16-byte aligned
#pragma unroll
            for (int ch = 0; ch < K1P / 8; ch++) {
                float x[8];
                if (vec) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int c = ch * 8 + h * 4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (env < n && c < in_dim) v = *reinterpret_cast<const float4*>(orow + c);
                        x[h * 4 + 0] = v.x; x[h * 4 + 1] = v.y; x[h * 4 + 2] = v.z; x[h * 4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int c = ch * 8 + e;
                        x[e] = (env < n && c < in_dim) ? orow[c] : 0.f;
                    }
                }
                uint4 w;
                w.x = pack_h2(x[0], x[1]); w.y = pack_h2(x[2], x[3]); w.z = pack_h2(x[4], x[5]); w.w = pack_h2(x[6], x[7]);
                *reinterpret_cast<uint4*>(sA1 + sw128_offset(TILE_M, row, ch * 8)) = w;
            }
        }
        if (first_tile) { mbar_wait(wbar, 0); first_tile = false; }   // weights + biases have landed
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        // ---- layer 1 ----
        if (tid == 0) { tc_fence_after(); issue_layer<H1, K1P>(tmem_base, a1, b1, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_hidden<H1>(tmem_row, sbias, sA2, row);
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        // ---- layer 2 ----
        if (tid == 0) { tc_fence_after(); issue_layer<H2, H1>(tmem_base + H1, a2, b2, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_hidden<H2>(tmem_row + H1, sbias + H1, sA3, row);
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        // ---- layer 3 ----
        if (tid == 0) { tc_fence_after(); issue_layer<OUTP, H2>(tmem_base + H1 + H2, a3, b3, bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32];
            if (OUTP == 32) tmem_ld32(tmem_row + H1 + H2, v); else tmem_ld16(tmem_row + H1 + H2, v);
            if (env < n) {
#pragma unroll
                for (int c = 0; c < OUTP; c++) if (c < out_dim) act[(size_t)env * out_dim + c] = v[c] + sbias[H1 + H2 + c];
            }
        }
        tc_fence_before();
        __syncthreads();      // all accumulators consumed before the next tile's MMAs overwrite them
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"((uint32_t)L::tmem_cols) : "memory");
    }
#endif
}

// host: weights (row-major [K][N], as in the .weights files) -> operand image
template <int H1, int H2, int OUTP>
static void pack_image(std::vector<unsigned char>& img, int in_dim, int out_dim, const float* w1, const float* b1, const float* w2,
                       const float* b2, const float* w3, const float* b3) {
    using L = Layout<H1, H2, OUTP>;
    img.assign(L::image_bytes, 0);
    auto put = [&](uint32_t base, int rows, int r, int c, float v) {
        __half h = __float2half_rn(v);
        memcpy(img.data() + base + sw128_offset(rows, r, c), &h, 2);
    };
    for (int nn = 0; nn < H1; nn++) for (int k = 0; k < in_dim; k++) put(L::b1, H1, nn, k, w1[(size_t)k * H1 + nn]);
    for (int nn = 0; nn < H2; nn++) for (int k = 0; k < H1; k++) put(L::b2, H2, nn, k, w2[(size_t)k * H2 + nn]);
    for (int nn = 0; nn < out_dim; nn++) for (int k = 0; k < H2; k++) put(L::b3, OUTP, nn, k, w3[(size_t)k * out_dim + nn]);
    float* hb = reinterpret_cast<float*>(img.data() + L::bias);
    for (int j = 0; j < H1; j++) hb[j] = b1[j];
    for (int j = 0; j < H2; j++) hb[H1 + j] = b2[j];
    for (int j = 0; j < out_dim; j++) hb[H1 + H2 + j] = b3[j];
}

}  // namespace tc5

// ---- entry points used by rs_mlp.cu ----
int rs_tc5_supported(int in_dim, int h1, int h2, int out_dim) {
    if (in_dim > tc5::K1P) return 0;
    if (h1 == 256 && h2 == 128 && out_dim <= 32) return 1;
    if (h1 == 128 && h2 == 64 && out_dim <= 16) return 2;
    return 0;
}

int rs_tc5_pack(int variant, int in_dim, int out_dim, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                const float* b3, void** image_dev, size_t* smem_bytes) {
    std::vector<unsigned char> img;
    cudaError_t e;
    if (variant == 1) {
        tc5::pack_image<256, 128, 32>(img, in_dim, out_dim, w1, b1, w2, b2, w3, b3);
        *smem_bytes = tc5::Layout<256, 128, 32>::bytes;
        e = cudaFuncSetAttribute(tc5::k_mlp_tc5<256, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem_bytes);
    } else {
        tc5::pack_image<128, 64, 16>(img, in_dim, out_dim, w1, b1, w2, b2, w3, b3);
        *smem_bytes = tc5::Layout<128, 64, 16>::bytes;
        e = cudaFuncSetAttribute(tc5::k_mlp_tc5<128, 64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem_bytes);
    }
    if (e != cudaSuccess) return 1;
    if (cudaMalloc(image_dev, img.size()) != cudaSuccess) return 1;
    if (cudaMemcpy(*image_dev, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
    return 0;
}

int rs_tc5_forward(int variant, const void* image_dev, size_t smem_bytes, const float* obs, float* act, int n, int in_dim, int out_dim,
                   int n_sm, cudaStream_t st) {
    const int tiles = (n + tc5::TILE_M - 1) / tc5::TILE_M;
    const int grid = tiles < n_sm ? tiles : n_sm;
    const unsigned char* img = (const unsigned char*)image_dev;
    if (variant == 1) tc5::k_mlp_tc5<256, 128, 32><<<grid, tc5::TILE_M, smem_bytes, st>>>(img, obs, act, n, in_dim, out_dim);
    else tc5::k_mlp_tc5<128, 64, 16><<<grid, tc5::TILE_M, smem_bytes, st>>>(img, obs, act, n, in_dim, out_dim);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
