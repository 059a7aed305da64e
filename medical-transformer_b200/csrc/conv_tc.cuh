// conv_tc.cuh -- interface of the tcgen05 (5th-generation tensor core) implicit-GEMM convolution kernels (conv_tc.cu).
#pragma once
#include <cuda_runtime.h>

// One problem = one split-K GEMM  D[rowsA][rowsB] (+)= A[rowsA][kdim] . B[rowsB][kdim]^T  whose operands are gathered on the fly:
//   MODE 0 (forward / dgrad)  rowsA = positions (n, oh, ow), rowsB = output channels, kdim = taps (c, kh, kw)
//   MODE 1 (wgrad)            rowsA = taps (ci, kh, kw),     rowsB = output channels, kdim = positions
struct TcConv {
  int N, Cimg, Crow, Hi, Wi, Ho, Wo, stride, pad, KH, KW, KS, K;   // Cimg: channels of the gathered image, K = Cimg * KS taps
  int dgrad, Wm;                                                    // MODE 0 weight indexing: flipped taps, dim-1 size of w
  int lgWo, lgHWo;                                                  // MODE 1: log2 of Wo and Ho * Wo (powers of two)
  long P;                                                           // positions N * Ho * Wo
  int rowsA, rowsB, kdim;
  int kb_per;                                                       // k blocks (of 32) per gridDim.z slice
  int atomic;                                                       // epilogue accumulates with fp32 atomics (split-K)
  int dbg;                                                          // bit 0: swap LBO / SBO in the smem descriptors (bring-up)
  int R, tiles_per_img, nbuf;                                       // slab kernel (3x3 / stride 1): slab rows, tiles per image, smem buffers
};

// 1 = launched, 0 = shape not eligible (caller uses the FMA kernels), < 0 = CUDA error.  When the launcher splits the reduction over
// gridDim.z it zeroes `out` itself (cudaMemsetAsync on `st`) and the slices accumulate with fp32 atomics.
int tc_conv_fwd_or_dgrad(const TcConv& geom, const float* img, const float* w, float* out, cudaStream_t st);
int tc_conv_wgrad(const TcConv& geom, const float* x, const float* dy, float* dw, cudaStream_t st);
int tc_conv_mode();              // MT_CONV_TC / mt_set_conv_tc: 0 off, 1 on (default), 2 on + debug bit 0
void tc_conv_set_mode(int m);
