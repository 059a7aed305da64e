// bn_cluster.cuh -- BatchNorm2d(+residual)(+ReLU) of the conv shell: argument block shared by shell_ops.cu (one CTA per channel)
// and bn_cluster.cu (a thread-block cluster per channel).
#pragma once
#include <cuda_runtime.h>

struct BnActArgs {
  int N, C, HW, G, training, relu;
  float eps, momentum;
};

// 1 = launched, 0 = not handled (caller uses the one-CTA-per-channel kernels), < 0 = CUDA error.  Both need HW % 4 == 0 and
// 16-byte aligned tensors (the caller checks).
int bn_cluster_forward(const BnActArgs& a, const float* x, const float* res, const float* gamma, const float* beta, float* rm,
                       float* rv, float* y, float* save_mean, float* save_invstd, cudaStream_t st);
int bn_cluster_backward(const BnActArgs& a, const float* x, const float* y, const float* dy, const float* gamma,
                        const float* save_mean, const float* save_invstd, float* dx, float* dres, float* dgamma, float* dbeta,
                        cudaStream_t st);
