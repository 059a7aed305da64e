// conv_gemm.cu -- the KxK convolutions of the conv shell (stem conv2 / conv3, the decoder convs; reference axialnet.py:424-440,
// :475-502) as implicit-GEMM fp32 kernels on the FMA pipes, split-K over the whole GPU.
//
// Why own kernels: cuDNN's fp32 (non-TF32: DESIGN.md section 3) algorithms for these shapes are 0.4 - 8 TFLOP/s -- ncu of one
// 128-px step: `implicit_convolve_sgemm` on 16 CTAs for decoder1 (256->256, 8x8 -> 4x4): 172 us; FFT / Winograd pipelines of
// four launches for decoder2-4; 18 - 74 us for the stem -- 0.4 ms forward, 0.17 ms dgrad, 0.19 ms wgrad of a 3.7 ms step.
// Every one of them is a small GEMM (75 - 300 MFLOP) whose K = Ci*9 is long and whose M x P is small, i.e. a split-K problem:
//
//   k_conv_gemm   C[M][P] (+)= A[M][K] . B[K][P]   B gathered on the fly from the NCHW image (im2col never materialised)
//       forward           M = Co, K = (ci, kh, kw), A = W (contiguous in k), image = x
//       dgrad (stride 1)  M = Ci, K = (co, kh', kw'), A[ci][(co,r)] = W[co][ci][KS-1-r] (flipped taps), image = dy, pad' = KH-1-pad
//   k_conv_wgrad  dW[Co][K] += dY[Co][P] . Xcol[K][P]^T   reduction over positions, split over position slices
//
// CTA = 256 threads, thread tile 4 x 4, CTA tile (4*TY) x (4*TX) with TY*TX = 256: TY = 16 (64 x 64) down to TY = 2 (8 channels x
// 512 positions) so that an 8-channel output (conv3) does not waste the M dimension.  K chunks of 16 through shared memory, the
// next chunk's global loads issued before the FMAs of the current one; gridDim.z slices of K accumulate with fp32 atomics into the
// zero-initialised output (last-bit run-to-run differences, like cuDNN's own split-K algorithms).
#include "common.cuh"
#include "conv_tc.cuh"
extern thread_local cudaError_t g_last_cuda;
int mt_fail_msg(int code, const char* msg);
cudaStream_t mt_side_fork(cudaStream_t main_stream);
void mt_side_join(cudaStream_t main_stream, cudaStream_t forked);
int coop_sm_count();

struct CgArgs {
  int N, Kimg, M, Hi, Wi, Ho, Wo, stride, pad, KH, KW, KS, K, dgrad, Wm;   // Wm: the weight tensor's dim-1 size when dgrad (== M)
  long P;
};

__device__ __forceinline__ void cg_mma(const float* As, int lda, const float* Bs, int ldb, int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(As + k * lda + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bs + k * ldb + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// position p (over N*Ho*Wo) -> image base offset (channel 0) and the top-left input coordinate of its window
struct CgPos { int ok, n, ih0, iw0, dst; };
__device__ __forceinline__ CgPos cg_pos(const CgArgs& a, long p) {
  CgPos r; r.ok = p < a.P;
  const long pp = r.ok ? p : 0;
  const int HWo = a.Ho * a.Wo;
  r.n = (int)(pp / HWo);
  r.dst = (int)(pp - (long)r.n * HWo);
  const int oh = r.dst / a.Wo, ow = r.dst - oh * a.Wo;
  r.ih0 = oh * a.stride - a.pad; r.iw0 = ow * a.stride - a.pad;
  return r;
}
__device__ __forceinline__ float cg_gather(const CgArgs& a, const float* __restrict__ img, const CgPos& q, int k) {
  if (!q.ok || k >= a.K) return 0.f;
  const int kc = a.KS == 9 ? k / 9 : k / a.KS, r = k - kc * a.KS;
  const int kh = a.KW == 3 ? r / 3 : r / a.KW, kw = r - kh * a.KW;
  const int ih = q.ih0 + kh, iw = q.iw0 + kw;
  if ((unsigned)ih >= (unsigned)a.Hi || (unsigned)iw >= (unsigned)a.Wi) return 0.f;
  return __ldg(img + (((size_t)q.n * a.Kimg + kc) * a.Hi + ih) * a.Wi + iw);
}
__device__ __forceinline__ float cg_weight(const CgArgs& a, const float* __restrict__ w, int m, int k) {
  if (m >= a.M || k >= a.K) return 0.f;
  if (!a.dgrad) return __ldg(w + (size_t)m * a.K + k);
  const int kc = a.KS == 9 ? k / 9 : k / a.KS, r = k - kc * a.KS;
  return __ldg(w + ((size_t)kc * a.Wm + m) * a.KS + (a.KS - 1 - r));
}

// grid (ceil(P / BN), ceil(M / BM), ksplit)
template <int TY>
__global__ void __launch_bounds__(256)
k_conv_gemm(CgArgs a, const float* __restrict__ img, const float* __restrict__ w, float* __restrict__ out, int kper) {
  constexpr int TX = 256 / TY, BM = 4 * TY, BN = 4 * TX, LDA = BM + 4, LDB = BN + 4;
  constexpr int NA = (16 * BM + 255) / 256, NB = 16 * BN / 256;      // elements per thread and chunk
  constexpr int PPT = BN > 256 ? BN / 256 : 1;                       // positions per thread in the B loads
  constexpr int TPC = BN < 256 ? 256 / BN : 1;                       // threads sharing a position column
  __shared__ __align__(16) float As[16 * LDA], Bs[16 * LDB];
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM;
  const long p0 = (long)blockIdx.x * BN;
  const int kbeg = blockIdx.z * kper, kend = a.K < kbeg + kper ? a.K : kbeg + kper;
  CgPos pos[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) pos[j] = cg_pos(a, p0 + (tid % BN) + 256 * j);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float va[NA], vb[NB];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int it = 0; it < NA; ++it) {                 // A[mm][kk]: kk = tid & 15 fastest (contiguous weights in the forward case)
      const int e = tid + 256 * it, kk = e & 15, mm = e >> 4;
      va[it] = (mm < BM && k0 + kk < kend) ? cg_weight(a, w, m0 + mm, k0 + kk) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) {                 // B[kb][pos]: positions fastest
      const int j = PPT > 1 ? it % PPT : 0;
      const int kb = PPT > 1 ? it / PPT : (tid / BN) + TPC * it;
      vb[it] = (k0 + kb < kend) ? cg_gather(a, img, pos[j], k0 + kb) : 0.f;
    }
  };
  load_chunk(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int it = 0; it < NA; ++it) {
      const int e = tid + 256 * it, kk = e & 15, mm = e >> 4;
      if (mm < BM) As[kk * LDA + mm] = va[it];
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) {
      const int j = PPT > 1 ? it % PPT : 0;
      const int kb = PPT > 1 ? it / PPT : (tid / BN) + TPC * it;
      Bs[kb * LDB + (tid % BN) + 256 * j] = vb[it];
    }
    __syncthreads();
    if (k0 + 16 < kend) load_chunk(k0 + 16);
    cg_mma(As, LDA, Bs, LDB, ty, tx, acc);
    __syncthreads();
  }
  const bool atomic = gridDim.z > 1;
  const int HWo = a.Ho * a.Wo;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long p = p0 + tx * 4 + j;
    if (p >= a.P) continue;
    const int n = (int)(p / HWo), dst = (int)(p - (long)n * HWo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mrow = m0 + ty * 4 + i;
      if (mrow < a.M) {
        float* o = out + ((size_t)n * a.M + mrow) * HWo + dst;
        if (atomic) atomicAdd(o, acc[i][j]); else *o = acc[i][j];
      }
    }
  }
}

// dW[co][k] += sum_p dY[co][p] Xcol[k][p].  grid (ceil(K / 64), ceil(Co / BM), psplit); K-dimension of the GEMM = positions.
template <int TY>
__global__ void __launch_bounds__(TY * 16)
k_conv_wgrad(CgArgs a, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, long pper) {
  constexpr int NT = TY * 16, BM = 4 * TY, BN = 64, LDA = BM + 4, LDB = BN + 4;
  constexpr int NA = 16 * BM / NT, NB = 16 * BN / NT;              // = 4 and 64 / TY
  __shared__ __align__(16) float As[16 * LDA], Bs[16 * LDB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const long pbeg = (long)blockIdx.z * pper, pend = pbeg + pper < a.P ? pbeg + pper : a.P;
  const int HWo = a.Ho * a.Wo;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // the thread's position inside a chunk is fixed (pp = tid & 15); its columns are col = (tid >> 4) + TY * it
  const int pp = tid & 15;
  float va[NA], vb[NB];
  auto load_chunk = [&](long pc) {
    const long p = pc + pp;
    CgArgs b = a;                                    // pend bounds the slice
    b.P = pend;
    const CgPos q = cg_pos(b, p);
#pragma unroll
    for (int it = 0; it < NA; ++it) {                // A[mm][pp] = dY[n][m0 + mm][dst]
      const int mm = ty + TY * it;
      va[it] = (q.ok && m0 + mm < a.M) ? __ldg(dy + ((size_t)q.n * a.M + m0 + mm) * HWo + q.dst) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < NB; ++it) vb[it] = cg_gather(a, x, q, c0 + ty + TY * it);
  };
  load_chunk(pbeg);
  for (long pc = pbeg; pc < pend; pc += 16) {
#pragma unroll
    for (int it = 0; it < NA; ++it) As[pp * LDA + ty + TY * it] = va[it];
#pragma unroll
    for (int it = 0; it < NB; ++it) Bs[pp * LDB + ty + TY * it] = vb[it];
    __syncthreads();
    if (pc + 16 < pend) load_chunk(pc + 16);
    cg_mma(As, LDA, Bs, LDB, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mrow = m0 + ty * 4 + i, col = c0 + tx * 4 + j;
      if (mrow < a.M && col < a.K) atomicAdd(dw + (size_t)mrow * a.K + col, acc[i][j]);
    }
}

static int cg_launch_gemm(const CgArgs& a, const float* img, const float* w, float* out, cudaStream_t st) {
  const int nsm = coop_sm_count();
  int ty = 16;
  while (ty > 2 && 4 * (ty / 2) >= a.M) ty /= 2;                  // smallest M tile that still covers M (8 channels -> TY = 2)
  const int BM = 4 * ty, BN = 4 * (256 / ty);
  const long gx = (a.P + BN - 1) / BN;
  const int gy = (a.M + BM - 1) / BM;
  const int chunks = (a.K + 15) / 16;
  int ks = (int)((2L * nsm + gx * gy - 1) / (gx * gy));           // ~2 CTAs per SM in flight
  if (ks > chunks) ks = chunks;
  if (ks < 1) ks = 1;
  const int kper = ((chunks + ks - 1) / ks) * 16;
  ks = (a.K + kper - 1) / kper;
  if (ks > 1 && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)a.P * a.M, st) != cudaSuccess) return -1;
  const dim3 grid((unsigned)gx, gy, ks);
  switch (ty) {
    case 16: k_conv_gemm<16><<<grid, 256, 0, st>>>(a, img, w, out, kper); break;
    case 8:  k_conv_gemm<8><<<grid, 256, 0, st>>>(a, img, w, out, kper); break;
    case 4:  k_conv_gemm<4><<<grid, 256, 0, st>>>(a, img, w, out, kper); break;
    default: k_conv_gemm<2><<<grid, 256, 0, st>>>(a, img, w, out, kper); break;
  }
  return 0;
}

static TcConv cg_tc_geom(const CgArgs& a) {
  TcConv t{};
  t.N = a.N; t.Cimg = a.Kimg; t.Crow = a.M; t.Hi = a.Hi; t.Wi = a.Wi; t.Ho = a.Ho; t.Wo = a.Wo; t.stride = a.stride; t.pad = a.pad;
  t.KH = a.KH; t.KW = a.KW; t.KS = a.KS; t.K = a.K; t.dgrad = a.dgrad; t.Wm = a.Wm; t.P = a.P;
  return t;
}

extern "C" {

int mt_set_conv_tc(int mode) { tc_conv_set_mode(mode); return MT_OK; }

int mt_conv2d_forward(int N, int Ci, int Co, int H, int W, int KH, int KW, int stride, int pad, const float* x, const float* w,
                      float* y, void* stream) {
  if (N <= 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || stride < 1 || pad < 0 || !x || !w || !y)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_conv2d_forward: bad argument");
  CgArgs a{};
  a.N = N; a.Kimg = Ci; a.M = Co; a.Hi = H; a.Wi = W; a.Ho = (H + 2 * pad - KH) / stride + 1; a.Wo = (W + 2 * pad - KW) / stride + 1;
  a.stride = stride; a.pad = pad; a.KH = KH; a.KW = KW; a.KS = KH * KW; a.K = Ci * a.KS; a.dgrad = 0; a.Wm = Co;
  a.P = (long)N * a.Ho * a.Wo;
  if (a.Ho <= 0 || a.Wo <= 0) return mt_fail_msg(MT_ERR_BAD_ARG, "mt_conv2d_forward: empty output");
  {                                                               // tensor-core path (conv_tc.cu) where the shape is eligible
    const int rc = tc_conv_fwd_or_dgrad(cg_tc_geom(a), x, w, y, (cudaStream_t)stream);
    if (rc < 0) return mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_forward: tcgen05 launch failed");
    if (rc > 0) { g_last_cuda = cudaGetLastError(); return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_forward: launch failed"); }
  }
  if (cg_launch_gemm(a, x, w, y, (cudaStream_t)stream) != 0) return mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_forward: memset failed");
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_forward: launch failed");
}

int mt_conv2d_backward(int N, int Ci, int Co, int H, int W, int KH, int KW, int stride, int pad, const float* x, const float* w,
                       const float* dy, float* dx, float* dw, void* stream) {
  if (N <= 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || stride < 1 || pad < 0 || !x || !w || !dy)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_conv2d_backward: bad argument");
  if (dx && stride != 1) return mt_fail_msg(MT_ERR_UNSUPPORTED, "mt_conv2d_backward: dx needs stride 1 (pass dx = NULL)");
  if (dx && (KH - 1 - pad < 0 || KW - 1 - pad < 0 || KH != KW))
    return mt_fail_msg(MT_ERR_UNSUPPORTED, "mt_conv2d_backward: dx needs a square kernel with pad <= K - 1");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  cudaStream_t sd = (dw && dx) ? mt_side_fork(st) : st;          // dW next to dX
  if (dw) {
    CgArgs a{};
    a.N = N; a.Kimg = Ci; a.M = Co; a.Hi = H; a.Wi = W; a.Ho = Ho; a.Wo = Wo; a.stride = stride; a.pad = pad; a.KH = KH; a.KW = KW;
    a.KS = KH * KW; a.K = Ci * a.KS; a.dgrad = 0; a.Wm = Co; a.P = (long)N * Ho * Wo;
    const int rc = tc_conv_wgrad(cg_tc_geom(a), x, dy, dw, sd);
    if (rc < 0) { mt_side_join(st, sd); return mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_backward: tcgen05 wgrad launch failed"); }
    if (rc == 0) {
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Co * a.K, sd);
    const int ty = Co <= 16 ? 4 : 16, BM = 4 * ty;
    const int gx = (a.K + 63) / 64, gy = (Co + BM - 1) / BM;
    const long chunks = (a.P + 15) / 16;
    long ps = (2L * coop_sm_count() * (ty == 4 ? 4 : 1) + gx * gy - 1) / (gx * gy);
    if (ps > chunks) ps = chunks;
    if (ps < 1) ps = 1;
    const long pper = ((chunks + ps - 1) / ps) * 16;
    ps = (a.P + pper - 1) / pper;
    const dim3 grid(gx, gy, (unsigned)ps);
    if (ty == 4) k_conv_wgrad<4><<<grid, 64, 0, sd>>>(a, x, dy, dw, pper);
    else k_conv_wgrad<16><<<grid, 256, 0, sd>>>(a, x, dy, dw, pper);
    }
  }
  if (dx) {
    CgArgs a{};                                                   // dX = conv(dY, flipped W^T), pad' = K - 1 - pad, stride 1
    a.N = N; a.Kimg = Co; a.M = Ci; a.Hi = Ho; a.Wi = Wo; a.Ho = H; a.Wo = W; a.stride = 1; a.pad = KH - 1 - pad; a.KH = KH; a.KW = KW;
    a.KS = KH * KW; a.K = Co * a.KS; a.dgrad = 1; a.Wm = Ci; a.P = (long)N * H * W;
    const int rc = tc_conv_fwd_or_dgrad(cg_tc_geom(a), dy, w, dx, st);
    if (rc < 0) { mt_side_join(st, sd); return mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_backward: tcgen05 dgrad launch failed"); }
    if (rc == 0 && cg_launch_gemm(a, dy, w, dx, st) != 0) { mt_side_join(st, sd); return mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_backward: memset failed"); }
  }
  mt_side_join(st, sd);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_conv2d_backward: launch failed");
}

}  // extern "C"
