// conv1x1.cu -- the bias-free 1x1 convolutions of the axial blocks (conv_down, conv_up, downsample;
// reference axialnet.py:14-16, 272, 276, 450-453) as the same 64x64x16 register-tiled fp32 SGEMM the
// projection uses.  NCHW in, NCHW out, optional spatial stride (the downsample path of stride-2 blocks).
//   k_c1_fwd   Y[n][co][q] = sum_ci W[co][ci] X[n][ci][src(q)]
//   k_c1_dx    dX[n][ci][src(q)] = sum_co W[co][ci] dY[n][co][q]          (other dX entries stay zero)
//   k_c1_dw    dW[co][ci] += sum_{n,q} dY[n][co][q] X[n][ci][src(q)]       (split over position chunks)
// relu_in = 1 folds the ReLU in front of conv_up (axialnet.py:291-293, :333-335, :380-382) into the operand loads / the dX store.
// cuDNN's fp32 (non-TF32) kernels for these shapes take 17-29 us each (profiles/r1_notes.md); TF32 is not
// an option.
#include "common.cuh"
#include "pdl.cuh"
extern thread_local cudaError_t g_last_cuda;
int mt_fail_msg(int code, const char* msg);
cudaStream_t mt_side_fork(cudaStream_t main_stream);
void mt_side_join(cudaStream_t main_stream, cudaStream_t forked);

#define C1_LD 68
struct C1Args { int N, Ci, Co, H, W, stride, Ho, Wo, relu_in; };   // relu_in: the conv reads relu(x) (conv_up of the blocks)

__device__ __forceinline__ void c1_mma(const float* As, const float* Bs, int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(As + k * C1_LD + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bs + k * C1_LD + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// output position q (global over N*Ho*Wo) -> image n, offset of the source pixel inside an input plane,
// offset inside an output plane
__device__ __forceinline__ bool c1_pos(const C1Args& a, long q, int& n, int& src, int& dst) {
  const int HWo = a.Ho * a.Wo;
  if (q >= (long)a.N * HWo) return false;
  n = (int)(q / HWo);
  dst = (int)(q - (long)n * HWo);
  if (a.stride == 1) { src = dst; return true; }
  const int ho = dst / a.Wo, wo = dst - ho * a.Wo;
  src = ho * a.stride * a.W + wo * a.stride;
  return true;
}

// grid (ceil(P/64), ceil(Co/64)); M = co, N = positions
__global__ void __launch_bounds__(256)
k_c1_fwd(C1Args a, const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y) {
  pdl_entry();
  __shared__ __align__(16) float As[16 * C1_LD], Bs[16 * C1_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int co0 = blockIdx.y * 64;
  const long q0 = (long)blockIdx.x * 64;
  const int HW = a.H * a.W, HWo = a.Ho * a.Wo;
  int pn, psrc, pdst;
  const bool pok = c1_pos(a, q0 + (tid & 63), pn, psrc, pdst);
  const float* xb = x + (pok ? (size_t)pn * a.Ci * HW + psrc : 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // the K loop is a chain of L2-latency-bound chunk loads (a few CTAs per launch, nothing else to hide them): the loads of
  // chunk k+1 are issued into registers before the FMAs of chunk k
  float va[4], vb[4];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {                // A: W[co0+mm][k0+kk], thread -> kk = tid&15, mm = tid>>4 + 16 it
      const int kk = tid & 15, mm = (tid >> 4) + 16 * it;
      va[it] = (co0 + mm < a.Co && k0 + kk < a.Ci) ? __ldg(w + (size_t)(co0 + mm) * a.Ci + k0 + kk) : 0.f;
      const int kb = (tid >> 6) + 4 * it;           // B: X[k0+kb][pos], pos = tid & 63
      vb[it] = (pok && k0 + kb < a.Ci) ? __ldg(xb + (size_t)(k0 + kb) * HW) : 0.f;
      if (a.relu_in) vb[it] = fmaxf(vb[it], 0.f);
    }
  };
  load_chunk(0);
  for (int k0 = 0; k0 < a.Ci; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      As[(tid & 15) * C1_LD + (tid >> 4) + 16 * it] = va[it];
      Bs[((tid >> 6) + 4 * it) * C1_LD + (tid & 63)] = vb[it];
    }
    __syncthreads();
    if (k0 + 16 < a.Ci) load_chunk(k0 + 16);
    c1_mma(As, Bs, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int n, src, dst;
    if (!c1_pos(a, q0 + tx * 4 + j, n, src, dst)) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + ty * 4 + i;
      if (co < a.Co) y[((size_t)n * a.Co + co) * HWo + dst] = acc[i][j];
    }
  }
}

// grid (ceil(P/64), ceil(Ci/64)); M = ci, N = positions, K = co
__global__ void __launch_bounds__(256)
k_c1_dx(C1Args a, const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ x,
        float* __restrict__ dx) {
  pdl_entry();
  __shared__ __align__(16) float As[16 * C1_LD], Bs[16 * C1_LD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ci0 = blockIdx.y * 64;
  const long q0 = (long)blockIdx.x * 64;
  const int HW = a.H * a.W, HWo = a.Ho * a.Wo;
  int pn, psrc, pdst;
  const bool pok = c1_pos(a, q0 + (tid & 63), pn, psrc, pdst);
  const float* yb = dy + (pok ? (size_t)pn * a.Co * HWo + pdst : 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float va[4], vb[4];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {                // A: W^T: As[kk][mm] = W[k0+kk][ci0+mm]; thread -> mm = tid&63, kk = tid>>6 + 4 it
      const int mm = tid & 63, kk = (tid >> 6) + 4 * it;
      va[it] = (ci0 + mm < a.Ci && k0 + kk < a.Co) ? __ldg(w + (size_t)(k0 + kk) * a.Ci + ci0 + mm) : 0.f;
      vb[it] = (pok && k0 + kk < a.Co) ? __ldg(yb + (size_t)(k0 + kk) * HWo) : 0.f;
    }
  };
  load_chunk(0);
  for (int k0 = 0; k0 < a.Co; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      As[((tid >> 6) + 4 * it) * C1_LD + (tid & 63)] = va[it];
      Bs[((tid >> 6) + 4 * it) * C1_LD + (tid & 63)] = vb[it];
    }
    __syncthreads();
    if (k0 + 16 < a.Co) load_chunk(k0 + 16);        // next chunk in flight during this chunk's FMAs
    c1_mma(As, Bs, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int n, src, dst;
    if (!c1_pos(a, q0 + tx * 4 + j, n, src, dst)) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ci = ci0 + ty * 4 + i;
      if (ci < a.Ci) {
        const size_t o = ((size_t)n * a.Ci + ci) * HW + src;
        dx[o] = (a.relu_in && !(__ldg(x + o) > 0.f)) ? 0.f : acc[i][j];      // d relu: the gradient passes where x > 0
      }
    }
  }
}

// grid (ceil(Co/64), ceil(Ci/64), ksplit); M = co, N = ci, K = positions
__global__ void __launch_bounds__(256)
k_c1_dw(C1Args a, const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, int ntiles) {
  extern __shared__ __align__(16) float smw[];
  float* Am = smw;                 // [64 pos][C1_LD] : Am[k][m] = dY[co0+m][pos k]
  float* Bm = Am + 64 * C1_LD;     // [64 pos][C1_LD] : Bm[k][n] = X[ci0+n][pos k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
  const int HW = a.H * a.W, HWo = a.Ho * a.Wo;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int tile = blockIdx.z; tile < ntiles; tile += gridDim.z) {
    int pn, psrc, pdst;
    const int r = tid & 63;
    const bool pok = c1_pos(a, (long)tile * 64 + r, pn, psrc, pdst);
    const float* yb = dy + (pok ? (size_t)pn * a.Co * HWo + pdst : 0);
    const float* xb = x + (pok ? (size_t)pn * a.Ci * HW + psrc : 0);
    for (int m0 = tid >> 6; m0 < 64; m0 += 16) {
      float vy[4], vx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mm = m0 + 4 * u;
        vy[u] = (pok && co0 + mm < a.Co) ? __ldg(yb + (size_t)(co0 + mm) * HWo) : 0.f;
        vx[u] = (pok && ci0 + mm < a.Ci) ? __ldg(xb + (size_t)(ci0 + mm) * HW) : 0.f;
        if (a.relu_in) vx[u] = fmaxf(vx[u], 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { Am[r * C1_LD + m0 + 4 * u] = vy[u]; Bm[r * C1_LD + m0 + 4 * u] = vx[u]; }
    }
    __syncthreads();
#pragma unroll
    for (int kb = 0; kb < 64; kb += 16) c1_mma(Am + kb * C1_LD, Bm + kb * C1_LD, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty * 4 + i, ci = ci0 + tx * 4 + j;
      if (co < a.Co && ci < a.Ci) atomicAdd(dw + (size_t)co * a.Ci + ci, acc[i][j]);
    }
}

extern "C" {

int mt_conv1x1_forward(int N, int Ci, int Co, int H, int W, int stride, int relu_in, const float* x, const float* w,
                       float* y, void* stream) {
  if (N <= 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || stride < 1 || !x || !w || !y)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_conv1x1_forward: bad argument");
  C1Args a{N, Ci, Co, H, W, stride, (H - 1) / stride + 1, (W - 1) / stride + 1, relu_in ? 1 : 0};
  const long P = (long)N * a.Ho * a.Wo;
#if defined(__CUDACC__)
  if (mt_pdl_enabled()) {
    if (mt_launch_pdl(k_c1_fwd, dim3((unsigned)((P + 63) / 64), (Co + 63) / 64), dim3(256), 0, (cudaStream_t)stream, 1, a, x, w, y) != cudaSuccess)
      return mt_fail_msg(MT_ERR_CUDA, "mt_conv1x1_forward: launch failed");
  } else
#endif
  k_c1_fwd<<<dim3((unsigned)((P + 63) / 64), (Co + 63) / 64), 256, 0, (cudaStream_t)stream>>>(a, x, w, y);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_conv1x1_forward: launch failed");
}

int mt_conv1x1_backward(int N, int Ci, int Co, int H, int W, int stride, int relu_in, const float* x, const float* w,
                        const float* dy, float* dx, float* dw, void* stream) {
  if (N <= 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || stride < 1 || !x || !w || !dy)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_conv1x1_backward: bad argument");
  C1Args a{N, Ci, Co, H, W, stride, (H - 1) / stride + 1, (W - 1) / stride + 1, relu_in ? 1 : 0};
  const long P = (long)N * a.Ho * a.Wo;
  cudaStream_t st = (cudaStream_t)stream;
  cudaStream_t sd = (dw && dx) ? mt_side_fork(st) : st;          // dW next to dX, not behind it
  if (dw) {
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * Ci, sd);
    const int ntiles = (int)((P + 63) / 64), mn = ((Co + 63) / 64) * ((Ci + 63) / 64);
    int ksplit = (148 + mn - 1) / mn;
    if (ksplit > ntiles) ksplit = ntiles;
    const size_t smem = sizeof(float) * 2 * 64 * C1_LD;
    static bool once = (cudaFuncSetAttribute(k_c1_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), true);
    (void)once;
    k_c1_dw<<<dim3((Co + 63) / 64, (Ci + 63) / 64, ksplit), 256, smem, sd>>>(a, dy, x, dw, ntiles);
  }
  if (dx) {
    if (stride > 1) cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * Ci * H * W, st);
#if defined(__CUDACC__)
    if (mt_pdl_enabled()) mt_launch_pdl(k_c1_dx, dim3((unsigned)((P + 63) / 64), (Ci + 63) / 64), dim3(256), 0, st, 1, a, dy, w, x, dx);
    else
#endif
    k_c1_dx<<<dim3((unsigned)((P + 63) / 64), (Ci + 63) / 64), 256, 0, st>>>(a, dy, w, x, dx);
  }
  mt_side_join(st, sd);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_conv1x1_backward: launch failed");
}

}  // extern "C"
