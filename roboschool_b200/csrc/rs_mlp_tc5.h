// rs_mlp_tc5.h -- internal interface of the tcgen05 policy kernel (rs_mlp_tc5.cu), used by rs_mlp.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

// 0: shape not covered; 1: 256-128 hidden (Humanoid family); 2: 128-64 hidden (the other agent_zoo v1 policies)
int rs_tc5_supported(int in_dim, int h1, int h2, int out_dim);
// builds the operand-layout weight image on the device; returns 0 on success
int rs_tc5_pack(int variant, int in_dim, int out_dim, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                const float* b3, void** image_dev, size_t* smem_bytes);
int rs_tc5_forward(int variant, const void* image_dev, size_t smem_bytes, const float* obs, float* act, int n, int in_dim, int out_dim,
                   int n_sm, cudaStream_t st);
