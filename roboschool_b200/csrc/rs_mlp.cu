// rs_mlp.cu -- agent_zoo v1 policy forward: relu(relu(x W1 + b1) W2 + b2) W3 + b3
// (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34,63-68), batched over envs on the tensor cores.
//
// The kernel is the tcgen05 / TMEM one in rs_mlp_tc5.cu; this file holds the policy object and its C ABI.  There is one
// compute path: a shape the tcgen05 kernel does not cover is an error, not a fallback.  (The round-1 mma.sync kernel lives in
// tools/mlp_mma_sync/ as a stand-alone baseline; it is not part of the library.)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/roboschool_b200.h"
#include "rs_mlp_tc5.h"

int rs_set_error(const char* s);
static int mfail(const std::string& s) { rs_set_error(s.c_str()); return 1; }
#define MCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return mfail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct rs_policy {
    int in_dim, h1, h2, out_dim, device;
    // tcgen05 / TMEM kernel (rs_mlp_tc5.cu): operand-layout weight image
    int tc5_variant, n_sm;
    void* tc5_image;
    size_t tc5_smem;
    uint64_t serial;     // unique per object: keys the step graphs that hold this policy's launch arguments (rs_step.cu)
};
uint64_t rs_policy_serial(const rs_policy* p) { return p ? p->serial : 0; }

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
    { static uint64_t next_serial = 0; p->serial = ++next_serial; }
    p->in_dim = in_dim; p->h1 = h1; p->h2 = h2; p->out_dim = out_dim; p->device = device;
    p->tc5_variant = rs_tc5_supported(in_dim, h1, h2, out_dim);
    p->tc5_image = nullptr;
    if (!p->tc5_variant) { delete p; return mfail("rs_policy_create: unsupported layer sizes (agent_zoo shapes only: 44-256-128-17, {5..28}-128-64-{1..8}, narrower hidden layers zero-padded)"); }
    cudaDeviceProp prop;
    MCK(cudaGetDeviceProperties(&prop, device));
    p->n_sm = prop.multiProcessorCount;
    if (rs_tc5_pack(p->tc5_variant, in_dim, out_dim, w1, b1, w2, b2, w3, b3, &p->tc5_image, &p->tc5_smem)) {
        cudaError_t e = cudaGetLastError();
        delete p;
        return mfail(std::string("rs_policy_create: packing the weight image failed: ") + cudaGetErrorString(e));
    }
    *out = p;
    return 0;
}

extern "C" int rs_policy_destroy(rs_policy* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    if (p->tc5_image) cudaFree(p->tc5_image);
    delete p;
    return 0;
}

extern "C" int rs_policy_forward(rs_policy* p, const float* obs_dev, float* act_dev, int n, void* stream) {
    if (!p || !obs_dev || !act_dev) return mfail("rs_policy_forward: null argument");
    MCK(cudaSetDevice(p->device));
    if (rs_tc5_forward(p->tc5_variant, p->tc5_image, p->tc5_smem, obs_dev, act_dev, n, p->in_dim, p->out_dim, p->n_sm, (cudaStream_t)stream))
        return mfail("rs_policy_forward: tcgen05 kernel launch failed");
    return 0;
}
