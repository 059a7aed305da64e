// head_ce.cu -- the tail of one training step as two kernels: `adjust(relu(x))` (the 1x1 conv with bias that produces the two-class
// logits, reference axialnet.py:440,502 / :588,708) fused with the loss `F.cross_entropy(logits, y)` (== LogNLLLoss, metrics.py:17-20,
// train.py:156), forward and backward.  In the ATen form this tail is ~10 launches of 3 - 27 us (relu, sgemm + bias, log_softmax,
// nll_loss2d forward / backward and their fills, conv dgrad / wgrad, a 27 us single-block bias reduction, relu backward): ~0.11 ms
// of a 3.5 ms step for 65 k pixels of 16 channels.  Used by medt_b200.TrainStep (which owns the loss); model(x) keeps the plain path.
//   forward : logits[n][k][p] = b[k] + sum_c w[k][c] relu(x[n][c][p]);  loss = mean_p (logsumexp_k logits - logits[y])
//   backward: dl_k = (softmax_k - [y == k]) * dloss / P;  dx = (x > 0) * sum_k w[k][c] dl_k;  dw[k][c] = sum_p dl_k relu(x);  db[k] = sum_p dl_k
#include "common.cuh"
extern thread_local cudaError_t g_last_cuda;
int mt_fail_msg(int code, const char* msg);

#define HC_MAXC 64

__device__ __forceinline__ float hc_block_sum(float v, float* red) {
  v = seg_sum(v, 32);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int k = 0; k < nw; ++k) s += red[k];
  return s;
}

__global__ void __launch_bounds__(256)
k_head_ce_fwd(int N, int C, int HW, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
              const long long* __restrict__ y, float* __restrict__ logits, float* __restrict__ loss) {
  __shared__ float ws[2 * HC_MAXC + 2], red[8];
  for (int k = threadIdx.x; k < 2 * C + 2; k += blockDim.x) ws[k] = k < 2 * C ? w[k] : b[k - 2 * C];
  __syncthreads();
  const long P = (long)N * HW;
  float part = 0.f;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
    const int n = (int)(p / HW), q = (int)(p - (long)n * HW);
    const float* xp = x + (size_t)n * C * HW + q;
    float l0 = ws[2 * C], l1 = ws[2 * C + 1];
#pragma unroll 8
    for (int c = 0; c < C; ++c) {
      const float v = fmaxf(__ldg(xp + (size_t)c * HW), 0.f);
      l0 = fmaf(ws[c], v, l0); l1 = fmaf(ws[C + c], v, l1);
    }
    logits[((size_t)n * 2) * HW + q] = l0;
    logits[((size_t)n * 2 + 1) * HW + q] = l1;
    const float m = fmaxf(l0, l1), lse = m + logf(expf(l0 - m) + expf(l1 - m));
    part += lse - (y[p] == 0 ? l0 : l1);
  }
  const float s = hc_block_sum(part, red);
  if (threadIdx.x == 0) atomicAdd(loss, s / (float)P);
}

__global__ void __launch_bounds__(256)
k_head_ce_bwd(int N, int C, int HW, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ logits,
              const long long* __restrict__ y, const float* __restrict__ dloss, float* __restrict__ dx, float* __restrict__ dw,
              float* __restrict__ db) {
  __shared__ float ws[2 * HC_MAXC], acc[2 * HC_MAXC + 2];
  for (int k = threadIdx.x; k < 2 * C; k += blockDim.x) ws[k] = w[k];
  for (int k = threadIdx.x; k < 2 * C + 2; k += blockDim.x) acc[k] = 0.f;
  __syncthreads();
  const long P = (long)N * HW;
  const float scale = __ldg(dloss) / (float)P;
  float b0 = 0.f, b1 = 0.f;
  // channels in the outer loop of the accumulation so that a thread keeps only two running sums per channel pass
  for (long p0 = (long)blockIdx.x * blockDim.x; p0 < P; p0 += (long)gridDim.x * blockDim.x) {
    const long p = p0 + threadIdx.x;
    const bool ok = p < P;
    int n = 0, q = 0;
    float d0 = 0.f, d1 = 0.f;
    if (ok) {
      n = (int)(p / HW); q = (int)(p - (long)n * HW);
      const float l0 = logits[((size_t)n * 2) * HW + q], l1 = logits[((size_t)n * 2 + 1) * HW + q];
      const float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), inv = 1.f / (e0 + e1);
      const bool cls1 = y[p] != 0;
      d0 = (e0 * inv - (cls1 ? 0.f : 1.f)) * scale;
      d1 = (e1 * inv - (cls1 ? 1.f : 0.f)) * scale;
      b0 += d0; b1 += d1;
    }
    const float* xp = x + (size_t)n * C * HW + q;
    float* dxp = dx ? dx + (size_t)n * C * HW + q : nullptr;
    for (int c = 0; c < C; ++c) {
      const float xv = ok ? __ldg(xp + (size_t)c * HW) : 0.f;
      const bool pos = xv > 0.f;
      if (ok && dxp) dxp[(size_t)c * HW] = pos ? fmaf(ws[c], d0, ws[C + c] * d1) : 0.f;
      float g0 = pos ? d0 * xv : 0.f, g1 = pos ? d1 * xv : 0.f;
      g0 = seg_sum(g0, 32); g1 = seg_sum(g1, 32);
      if ((threadIdx.x & 31) == 0) { atomicAdd(acc + c, g0); atomicAdd(acc + C + c, g1); }
    }
  }
  b0 = seg_sum(b0, 32); b1 = seg_sum(b1, 32);
  if ((threadIdx.x & 31) == 0) { atomicAdd(acc + 2 * C, b0); atomicAdd(acc + 2 * C + 1, b1); }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * C + 2; k += blockDim.x) atomicAdd(k < 2 * C ? dw + k : db + (k - 2 * C), acc[k]);
}

extern "C" {

int mt_head_ce_forward(int N, int C, int HW, const float* x, const float* w, const float* b, const long long* y, float* logits,
                       float* loss, void* stream) {
  if (N <= 0 || C <= 0 || C > HC_MAXC || HW <= 0 || !x || !w || !b || !y || !logits || !loss)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_head_ce_forward: bad argument (C <= 64, two classes)");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return mt_fail_msg(MT_ERR_CUDA, "mt_head_ce_forward: memset failed");
  const long P = (long)N * HW;
  int grid = (int)((P + 255) / 256);
  if (grid > 592) grid = 592;
  k_head_ce_fwd<<<grid, 256, 0, st>>>(N, C, HW, x, w, b, y, logits, loss);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_head_ce_forward: launch failed");
}

int mt_head_ce_backward(int N, int C, int HW, const float* x, const float* w, const float* logits, const long long* y,
                        const float* dloss, float* dx, float* dw, float* db, void* stream) {
  if (N <= 0 || C <= 0 || C > HC_MAXC || HW <= 0 || !x || !w || !logits || !y || !dloss || !dw || !db)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_head_ce_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dw, 0, sizeof(float) * 2 * C, st) != cudaSuccess || cudaMemsetAsync(db, 0, sizeof(float) * 2, st) != cudaSuccess)
    return mt_fail_msg(MT_ERR_CUDA, "mt_head_ce_backward: memset failed");
  const long P = (long)N * HW;
  int grid = (int)((P + 255) / 256);
  if (grid > 296) grid = 296;
  k_head_ce_bwd<<<grid, 256, 0, st>>>(N, C, HW, x, w, logits, y, dloss, dx, dw, db);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_head_ce_backward: launch failed");
}

}  // extern "C"
