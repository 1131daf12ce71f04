// rs_mlp_tc5.cu -- agent_zoo v1 policy forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   act = relu(relu(obs W1 + b1) W2 + b2) W3 + b3        (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34,63-68)
//
// One persistent CTA per SM, 256 threads = two tile slots of 128 threads; a tile is 128 envs (UMMA M = 128, one env per TMEM
// lane / per thread of the slot).  The slots share the weight image and the tensor pipe: one slot's epilogue overlaps the other
// slot's MMAs, and 2 x 148 slots hold all 256 tiles of a 32768-env batch at once.  Per tile, chained on chip:
//   obs tile -> fp16 -> smem A1 [128 x 64]            (K-major, 128-byte swizzle: the canonical UMMA operand layout)
//   for each chunk c of 128 hidden units:
//     tcgen05.mma  X[128 x 128] (TMEM) = A1 . W1^T[chunk c]       (one elected thread per slot issues K/16 instructions)
//     tcgen05.ld   X -> registers, + b1, relu, -> fp16 -> smem A2 [128 x 128]
//     tcgen05.mma  D2[128 x H2] (+)= A2 . W2^T[:, chunk c]        (issued together with chunk c+1's layer-1 MMA)
//   epilogue D2 -> smem A3 [128 x H2] (in A2's place) ;  tcgen05.mma X[128 x OUTP] = A3 . W3^T ;  + b3 -> global act
// Hidden activations never leave the SM; the weights are staged once per CTA, already in operand layout
// (the host packs them swizzled, so staging is one linear TMA bulk copy, cp.async.bulk, overlapped with the first tiles' loads).
// Completion of each MMA group is signalled with tcgen05.commit on the slot's mbarrier; generic-proxy writes of the next A
// operand are ordered before the async-proxy reads with fence.proxy.async.
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

constexpr int NCHUNK = 128;      // hidden-layer-1 columns per pass (layer 1 is produced and consumed in N-chunks of 128)
constexpr int TMEM_PER_TILE = 256;   // TMEM columns of one tile slot: X[128] (layer-1 chunk, later the output) | D2[<=128]

template <int H1, int H2, int OUTP>
struct Layout {
    // weights first (one contiguous image copied from global), then the activation operands of the two tile slots
    static constexpr uint32_t b1 = 0;                                   // W1^T : H1 rows x K1P
    static constexpr uint32_t b2 = b1 + H1 * 128 * (K1P / ATOM_K);      // W2^T : H2 rows x H1
    static constexpr uint32_t b3 = b2 + H2 * 128 * (H1 / ATOM_K);       // W3^T : OUTP rows x H2
    static constexpr uint32_t bias = b3 + OUTP * 128 * (H2 / ATOM_K);   // fp32 [H1 + H2 + OUTP]
    static constexpr uint32_t image_bytes = (bias + 4 * (H1 + H2 + OUTP) + 1023) / 1024 * 1024;
    // per tile slot: A1 [128 x K1P] and A2 [128 x NCHUNK] (one chunk of the hidden layer; the layer-3 operand A3 [128 x H2]
    // reuses it once layer 2 has consumed the last chunk)
    static constexpr uint32_t a1_bytes = TILE_M * 128 * (K1P / ATOM_K);
    static constexpr uint32_t a2_bytes = TILE_M * 128 * (NCHUNK / ATOM_K);
    static constexpr uint32_t slot_bytes = a1_bytes + a2_bytes;
    static constexpr uint32_t slots = image_bytes;                      // slot s at slots + s * slot_bytes : A1 | A2
    static constexpr uint32_t bar = slots + 2 * slot_bytes;             // MMA mbarrier per slot (2 x 8 B), two weight mbarriers (2 x 8 B), TMEM base (4 B)
    static constexpr uint32_t bytes = bar + 48;
    static constexpr int tmem_cols = 2 * TMEM_PER_TILE;
    static_assert(H1 % NCHUNK == 0 && H2 % 64 == 0 && H2 <= 128 && OUTP % 16 == 0 && OUTP <= 32, "layer widths");
    static_assert(bytes <= 227 * 1024, "shared memory");
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

// D[128 x N] (TMEM column tmem_d) (+)= A[128 x K] . B[N x K]^T over K-atoms [atom0, atom0 + K/64) of B (rows_b rows per atom);
// A's atoms start at 0.  Both K-major SW128 in shared memory; issued by one thread, no commit.
template <int N, int K>
__device__ __forceinline__ void issue_mma(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, int rows_b, int b_atom0, bool accumulate) {
    const uint32_t idesc = umma_idesc(TILE_M, N);
#pragma unroll
    for (int k = 0; k < K / 16; k++) {
        const int atom = (k * 16) / ATOM_K, kin = (k * 16) % ATOM_K;
        const uint64_t ad = umma_desc(a_base + atom * (TILE_M * 128) + kin * 2);
        const uint64_t bd = umma_desc(b_base + (b_atom0 + atom) * (rows_b * 128) + kin * 2);
        umma_f16(tmem_d, ad, bd, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
}

// accumulators of this thread's row (N columns from tmem_row) -> + bias, relu -> fp16 -> next A operand (row `row`, columns 0..N-1)
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
// barrier over the 128 threads of one tile slot (named barriers 1 and 2; barrier 0 is __syncthreads)
__device__ __forceinline__ void slot_sync(int slot) { asm volatile("bar.sync %0, 128;\n" :: "r"(slot + 1) : "memory"); }
#endif  // __CUDA_ARCH__

// Two tile slots per CTA (one per warpgroup of 128 threads) share the weight image and the tensor pipe: while one slot is in
// an epilogue (tcgen05.ld, bias, relu, fp16, swizzled store) the other slot's MMAs run.  With one CTA per SM that is 296 slots
// for the 256 tiles of 32768 envs: every tile is in flight from the start (one pass instead of 1.73 waves).  A slot owns 256
// TMEM columns: X[128] | D2[H2].  Layer 1 is produced in chunks of 128 hidden units into X; each chunk goes through its
// epilogue into A2 and is consumed at once by a partial layer-2 MMA (K = 128, accumulating into D2) that is issued together
// with the next chunk's layer-1 MMA; the layer-3 result lands in X again.
template <int H1, int H2, int OUTP>
__global__ void __launch_bounds__(2 * TILE_M, 1) k_mlp_tc5(const unsigned char* __restrict__ image, const float* __restrict__ obs,
                                                           float* __restrict__ act, int n, int in_dim, int out_dim) {
#if defined(__CUDA_ARCH__)
    using L = Layout<H1, H2, OUTP>;
    constexpr int NC = H1 / NCHUNK;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, slot = tid >> 7, row = tid & (TILE_M - 1);
    const uint32_t bar = smem_u32(smem + L::bar) + 8 * slot, wbar1 = smem_u32(smem + L::bar) + 16, wbar2 = wbar1 + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::bar + 32);
    if (tid == 0) {
        mbar_init(smem_u32(smem + L::bar), 1);
        mbar_init(smem_u32(smem + L::bar) + 8, 1);
        mbar_init(wbar1, 1);
        mbar_init(wbar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // stage the weight image (already in operand layout) + biases with the TMA bulk engine, in the order the layers need
        // them: W1 and the biases complete on wbar1 (layer 1 starts as soon as those 34 KB are in), W2 and W3 on wbar2 (they
        // land while the first observation tiles are converted and layer 1 runs)
        constexpr uint32_t CH = 32768;
        auto stage = [&](uint32_t lo, uint32_t hi, uint32_t wb) {
            for (uint32_t off = lo; off < hi; off += CH) bulk_g2s(smem_u32(smem) + off, image + off, (hi - off) < CH ? (hi - off) : CH, wb);
        };
        mbar_expect_tx(wbar1, (L::b2 - L::b1) + (L::image_bytes - L::bias));
        mbar_expect_tx(wbar2, L::bias - L::b2);
        stage(L::b1, L::b2, wbar1);
        stage(L::bias, L::image_bytes, wbar1);
        stage(L::b2, L::bias, wbar2);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)L::tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + (uint32_t)(slot * TMEM_PER_TILE);  // this slot's columns
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16); // this warp's 32 lanes (a warp reaches lanes 32 (warp % 4) ..)
    const uint32_t tmX = tmem_base, tmD2 = tmem_base + NCHUNK;
    const float* sbias = reinterpret_cast<const float*>(smem + L::bias);
    unsigned char* sA1 = smem + L::slots + slot * L::slot_bytes;
    unsigned char* sA2 = sA1 + L::a1_bytes;
    const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2);
    const uint32_t b1 = smem_u32(smem + L::b1), b2 = smem_u32(smem + L::b2), b3 = smem_u32(smem + L::b3);
    const bool issuer = row == 0;
    uint32_t phase = 0;
    bool first_tile = true;
    const int n_tiles = (n + TILE_M - 1) / TILE_M;
    for (int tile = slot * gridDim.x + blockIdx.x; tile < n_tiles; tile += 2 * gridDim.x) {
        const int env = tile * TILE_M + row;
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
        if (first_tile) mbar_wait(wbar1, 0);   // W1 + biases have landed
        proxy_fence();
        tc_fence_before();
        slot_sync(slot);
        // ---- layer 1, chunk 0 -> X ----
        if (issuer) { tc_fence_after(); issue_mma<NCHUNK, K1P>(tmX, a1, b1, NCHUNK, 0, false); umma_commit(bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NC; c++) {
            // chunk c of the hidden layer: X -> + b1, relu, fp16 -> A2 (the partial layer-2 MMA that read A2 before has completed:
            // the commit waited on above covered it)
            epilogue_hidden<NCHUNK>(tmem_row, sbias + c * NCHUNK, sA2, row);
            if (first_tile && c == 0) { mbar_wait(wbar2, 0); first_tile = false; }   // W2 + W3 have landed
            proxy_fence();
            tc_fence_before();
            slot_sync(slot);
            if (issuer) {
                tc_fence_after();
                issue_mma<H2, NCHUNK>(tmD2, a2, b2, H2, c * (NCHUNK / ATOM_K), c > 0);                          // D2 (+)= A2 . W2^T[:, chunk c]
                if (c + 1 < NC) issue_mma<NCHUNK, K1P>(tmX, a1, b1 + (c + 1) * (NCHUNK * 128), NCHUNK, 0, false); // next chunk of layer 1 -> X
                umma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            tc_fence_after();
        }
        // ---- layer 2 epilogue -> A3 (in A2's place), layer 3 -> X[0, OUTP) ----
        epilogue_hidden<H2>(tmem_row + NCHUNK, sbias + H1, sA2, row);
        proxy_fence();
        tc_fence_before();
        slot_sync(slot);
        if (issuer) { tc_fence_after(); issue_mma<OUTP, H2>(tmX, a2, b3, OUTP, 0, false); umma_commit(bar); }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32];
            if (OUTP == 32) tmem_ld32(tmem_row, v); else tmem_ld16(tmem_row, v);
            if (env < n) {
#pragma unroll
                for (int c = 0; c < OUTP; c++) if (c < out_dim) act[(size_t)env * out_dim + c] = v[c] + sbias[H1 + H2 + c];
            }
        }
        tc_fence_before();
        slot_sync(slot);      // all accumulators consumed before the slot's next tile overwrites them
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(*tmem_slot), "r"((uint32_t)L::tmem_cols) : "memory");
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
    if (variant == 1) tc5::k_mlp_tc5<256, 128, 32><<<grid, 2 * tc5::TILE_M, smem_bytes, st>>>(img, obs, act, n, in_dim, out_dim);
    else tc5::k_mlp_tc5<128, 64, 16><<<grid, 2 * tc5::TILE_M, smem_bytes, st>>>(img, obs, act, n, in_dim, out_dim);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
