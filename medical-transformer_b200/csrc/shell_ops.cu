// shell_ops.cu -- the glue of the conv shell around the axial ops, as fused sm_100a kernels:
//
//   mt_up2x_relu_add_*  relu(F.interpolate(x, scale 2, bilinear, align_corners=False)) + skip
//                       (one decoder step, reference axialnet.py:493-501, 650-652, 690-698)
//   mt_bn_act_*         BatchNorm2d with batch statistics (+ residual) (+ ReLU), per statistic
//                       group: AxialBlock*.forward :285-300 / :327-342 / :373-389, the stems
//                       :475-483, and the 16 per-patch BN calls of the LoGo loop :661-700 batched
//                       as 16 groups with the reference's sequential running-stat updates.
// All fp32, NCHW.  HBM-bound elementwise / reduction work: coalesced float loads along W,
// one CTA per (group, channel) for the reductions.
#include <stdio.h>
#include "common.cuh"
#include "bn_cluster.cuh"
#include "pdl.cuh"
#include <cstdlib>

// ---------------------------------------------------------------------------------------------
// bilinear x2, align_corners=False:  src = dst/2 - 0.25 clamped at 0
__device__ __forceinline__ void up_taps(int dst, int n_in, int& i0, int& i1, float& lam) {
  float src = 0.5f * (float)dst - 0.25f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  lam = src - (float)i0;
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
}

// one thread per output pixel pair along W (float2 store); mask byte = (pre-activation > 0)
__global__ void __launch_bounds__(256)
k_up2x_fwd(int planes, int C, int H, int W, const float* __restrict__ in, const float* __restrict__ bias,
           const float* __restrict__ skip, float* __restrict__ out, unsigned char* __restrict__ mask) {
  pdl_entry();
  const int Wo = 2 * W, Ho = 2 * H;
  const long total = (long)planes * Ho * W;                 // W pairs per output row
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wp = (int)(idx % W);
  const long t = idx / W;
  const int ho = (int)(t % Ho);
  const long pl = t / Ho;
  int h0, h1; float lh;
  up_taps(ho, H, h0, h1, lh);
  const float* r0 = in + (pl * H + h0) * W;
  const float* r1 = in + (pl * H + h1) * W;
  const int wm = wp > 0 ? wp - 1 : 0, wpl = wp + 1 < W ? wp + 1 : W - 1;
  // columns: out 2wp uses (wp-1, wp) with lam 0.75 (or clamp), out 2wp+1 uses (wp, wp+1) with lam 0.25
  const float a_m = (1.f - lh) * __ldg(r0 + wm) + lh * __ldg(r1 + wm);
  const float a_c = (1.f - lh) * __ldg(r0 + wp) + lh * __ldg(r1 + wp);
  const float a_p = (1.f - lh) * __ldg(r0 + wpl) + lh * __ldg(r1 + wpl);
  float v0 = wp > 0 ? 0.25f * a_m + 0.75f * a_c : a_c;
  float v1 = 0.75f * a_c + 0.25f * a_p;                     // at wp == W-1: a_p == a_c -> a_c
  if (bias) {                                               // the taps sum to 1: conv bias commutes with the upsample
    const float bb = __ldg(bias + (int)(pl % C));
    v0 += bb; v1 += bb;
  }
  const long o = (pl * Ho + ho) * Wo + 2 * wp;
  unsigned char m0 = v0 > 0.f, m1 = v1 > 0.f;
  v0 = m0 ? v0 : 0.f; v1 = m1 ? v1 : 0.f;
  if (skip) { const float2 s = __ldg(reinterpret_cast<const float2*>(skip + o)); v0 += s.x; v1 += s.y; }
  *reinterpret_cast<float2*>(out + o) = make_float2(v0, v1);
  if (mask) { mask[o] = m0; mask[o + 1] = m1; }
}

// gather form: one thread per input pixel sums its (up to) 4x4 masked output gradients; with `dbias` the same
// pass also reduces d bias[c] = sum over (n, h, w) of din (warp shuffle -> CTA -> one atomic per channel run)
__global__ void __launch_bounds__(256)
k_up2x_bwd(int planes, int C, int H, int W, const float* __restrict__ dout, const unsigned char* __restrict__ mask,
           float* __restrict__ din, float* __restrict__ dbias) {
  pdl_entry();
  const long total = (long)planes * H * W;
  const long gidx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = gidx < total;
  const long idx = act ? gidx : total - 1;
  const int w = (int)(idx % W);
  const long t = idx / W;
  const int h = (int)(t % H);
  const long pl = t / H;
  const int Wo = 2 * W, Ho = 2 * H;
  float wr[4], wc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 2 * h - 1 + k, cidx = 2 * w - 1 + k;
    int i0, i1; float lam;
    wr[k] = 0.f; wc[k] = 0.f;
    if (r >= 0 && r < Ho) { up_taps(r, H, i0, i1, lam); wr[k] = (i0 == h ? 1.f - lam : 0.f) + (i1 == h ? lam : 0.f); }
    if (cidx >= 0 && cidx < Wo) { up_taps(cidx, W, i0, i1, lam); wc[k] = (i0 == w ? 1.f - lam : 0.f) + (i1 == w ? lam : 0.f); }
  }
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wr[a] == 0.f) continue;
    const int r = 2 * h - 1 + a;
    const long base = (pl * Ho + r) * Wo;
    float rowacc = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wc[b] == 0.f) continue;
      const long o = base + 2 * w - 1 + b;
      const float g = mask[o] ? __ldg(dout + o) : 0.f;
      rowacc = fmaf(wc[b], g, rowacc);
    }
    acc = fmaf(wr[a], rowacc, acc);
  }
  if (act) din[idx] = acc;
  if (!dbias) return;                                       // uniform over the grid
  __shared__ float wsum[8];
  __shared__ int wch[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ch = act ? (int)(pl % C) : -1;
  const int ch0 = __shfl_sync(0xffffffffu, ch, 0);
  const bool uni = __all_sync(0xffffffffu, ch == ch0);      // whole warp inside one channel plane (the usual case)
  float sacc = uni ? acc : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
  if (!uni && act) atomicAdd(dbias + ch, acc);
  if (lane == 0) { wsum[wid] = sacc; wch[wid] = uni ? ch0 : -1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    int cur = -1; float run = 0.f;                          // merge runs of warps that share a channel
    for (int k = 0; k < nw; ++k) {
      if (wch[k] != cur) {
        if (cur >= 0) atomicAdd(dbias + cur, run);
        cur = wch[k]; run = 0.f;
      }
      run += wsum[k];
    }
    if (cur >= 0) atomicAdd(dbias + cur, run);
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d (+residual)(+ReLU), one CTA per (channel, statistic group)

__device__ __forceinline__ float2 block_sum2(float a, float b) {
  __shared__ float ra[32], rb[32];
  a = seg_sum(a, 32); b = seg_sum(b, 32);
  const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { ra[wid] = a; rb[wid] = b; }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
  for (int k = 0; k < nw; ++k) { sa += ra[k]; sb += rb[k]; }      // fixed order: deterministic
  return make_float2(sa, sb);
}

// element i4 (float4 index inside the (group, channel) slab) -> global float4 offset
__device__ __forceinline__ size_t bn_off4(const BnActArgs& a, int g, int c, int npg, int hw4, int i4) {
  const int n = i4 / hw4, r = i4 - n * hw4;
  return ((size_t)(g * npg + n) * a.C + c) * hw4 + r;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
k_bn_act_fwd(BnActArgs a, const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
             const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv, float* __restrict__ y,
             float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  constexpr int U = 4, W = VEC ? 4 : 1;
  const int c = blockIdx.x, g = blockIdx.y;
  const int npg = a.N / a.G, hw4 = a.HW / W, cnt4 = npg * hw4, cnt = npg * a.HW;
  float mean, invstd;
  if (a.training) {
    // pass 1: shifted sums (shift = first element) keep the fp32 variance well conditioned
    const float shift = __ldg(x + ((size_t)(g * npg) * a.C + c) * a.HW);
    float s1 = 0.f, s2 = 0.f;
    for (int base = 0; base < cnt4; base += U * blockDim.x) {
      float v[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i4 = base + u * blockDim.x + threadIdx.x;
        if (i4 < cnt4) {
          if (VEC) { const float4 t = __ldg(reinterpret_cast<const float4*>(x) + bn_off4(a, g, c, npg, hw4, i4));
                     v[u][0] = t.x; v[u][W > 1 ? 1 : 0] = t.y; v[u][W > 2 ? 2 : 0] = t.z; v[u][W > 3 ? 3 : 0] = t.w; }
          else v[u][0] = __ldg(x + bn_off4(a, g, c, npg, hw4, i4));
        } else {
#pragma unroll
          for (int k = 0; k < W; ++k) v[u][k] = shift;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < W; ++k) { const float d = v[u][k] - shift; s1 += d; s2 = fmaf(d, d, s2); }
    }
    const float2 s = block_sum2(s1, s2);
    const float inv_n = 1.f / (float)cnt;
    const float ms = s.x * inv_n;
    float var = s.y * inv_n - ms * ms;
    var = var > 0.f ? var : 0.f;
    mean = ms + shift;
    invstd = rsqrt_nr(var + a.eps);
    if (threadIdx.x == 0) {
      save_mean[g * a.C + c] = mean; save_invstd[g * a.C + c] = invstd;
      save_mean[(a.G + g) * a.C + c] = var;                 // biased variance, for the grouped EMA kernel
      if (a.G == 1) {                                       // single statistic group: EMA right here, no extra launch
        const float unb = cnt > 1 ? (float)cnt / (float)(cnt - 1) : 1.f;
        rm[c] = (1.f - a.momentum) * rm[c] + a.momentum * mean;
        rv[c] = (1.f - a.momentum) * rv[c] + a.momentum * var * unb;
      }
    }
  } else {
    mean = rm[c]; invstd = rsqrt_nr(rv[c] + a.eps);
    if (threadIdx.x == 0) { save_mean[g * a.C + c] = mean; save_invstd[g * a.C + c] = invstd; }
  }
  const float ga = gamma[c] * invstd, be = beta[c] - mean * ga;
  for (int base = 0; base < cnt4; base += U * blockDim.x) {
    float v[U][W], rr[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
#pragma unroll
      for (int k = 0; k < W; ++k) { v[u][k] = 0.f; rr[u][k] = 0.f; }
      if (i4 < cnt4) {
        const size_t o = bn_off4(a, g, c, npg, hw4, i4);
        if (VEC) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(x) + o);
          v[u][0] = t.x; v[u][W > 1 ? 1 : 0] = t.y; v[u][W > 2 ? 2 : 0] = t.z; v[u][W > 3 ? 3 : 0] = t.w;
          if (res) { const float4 q = __ldg(reinterpret_cast<const float4*>(res) + o);
                     rr[u][0] = q.x; rr[u][W > 1 ? 1 : 0] = q.y; rr[u][W > 2 ? 2 : 0] = q.z; rr[u][W > 3 ? 3 : 0] = q.w; }
        } else {
          v[u][0] = __ldg(x + o);
          if (res) rr[u][0] = __ldg(res + o);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
      if (i4 < cnt4) {
        float o4[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
          float t = fmaf(ga, v[u][k], be) + rr[u][k];
          o4[k] = (a.relu && !(t > 0.f)) ? 0.f : t;
        }
        const size_t o = bn_off4(a, g, c, npg, hw4, i4);
        if (VEC) reinterpret_cast<float4*>(y)[o] = make_float4(o4[0], o4[W > 1 ? 1 : 0], o4[W > 2 ? 2 : 0], o4[W > 3 ? 3 : 0]);
        else y[o] = o4[0];
      }
    }
  }
}

// sequential running-stat EMA over the statistic groups (one BN call per group in the reference)
__global__ void k_bn_ema(BnActArgs a, const float* __restrict__ save_mean, float* __restrict__ rm, float* __restrict__ rv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float cnt = (float)(a.N / a.G) * a.HW;
  const float unb = cnt > 1.f ? cnt / (cnt - 1.f) : 1.f;
  float m_ = rm[c], v_ = rv[c];
  for (int g = 0; g < a.G; ++g) {
    m_ = (1.f - a.momentum) * m_ + a.momentum * save_mean[g * a.C + c];
    v_ = (1.f - a.momentum) * v_ + a.momentum * save_mean[(a.G + g) * a.C + c] * unb;
  }
  rm[c] = m_; rv[c] = v_;
}

// dx (and d residual = masked dy), dgamma / dbeta accumulated over groups
template <bool VEC>
__global__ void __launch_bounds__(256)
k_bn_act_bwd(BnActArgs a, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
             const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
             float* __restrict__ dx, float* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int U = 4, W = VEC ? 4 : 1;
  const int c = blockIdx.x, g = blockIdx.y;
  const int npg = a.N / a.G, hw4 = a.HW / W, cnt4 = npg * hw4, cnt = npg * a.HW;
  const float mean = save_mean[g * a.C + c], invstd = save_invstd[g * a.C + c];
  float s1 = 0.f, s2 = 0.f;
  for (int base = 0; base < cnt4; base += U * blockDim.x) {
    float vx[U][W], vy[U][W], vd[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
#pragma unroll
      for (int k = 0; k < W; ++k) { vx[u][k] = mean; vy[u][k] = 1.f; vd[u][k] = 0.f; }
      if (i4 < cnt4) {
        const size_t o = bn_off4(a, g, c, npg, hw4, i4);
        if (VEC) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(x) + o), d = __ldg(reinterpret_cast<const float4*>(dy) + o);
          vx[u][0] = t.x; vx[u][W > 1 ? 1 : 0] = t.y; vx[u][W > 2 ? 2 : 0] = t.z; vx[u][W > 3 ? 3 : 0] = t.w;
          vd[u][0] = d.x; vd[u][W > 1 ? 1 : 0] = d.y; vd[u][W > 2 ? 2 : 0] = d.z; vd[u][W > 3 ? 3 : 0] = d.w;
          if (a.relu) { const float4 q = __ldg(reinterpret_cast<const float4*>(y) + o);
                        vy[u][0] = q.x; vy[u][W > 1 ? 1 : 0] = q.y; vy[u][W > 2 ? 2 : 0] = q.z; vy[u][W > 3 ? 3 : 0] = q.w; }
        } else {
          vx[u][0] = __ldg(x + o); vd[u][0] = __ldg(dy + o);
          if (a.relu) vy[u][0] = __ldg(y + o);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float d = (a.relu && !(vy[u][k] > 0.f)) ? 0.f : vd[u][k];
        s1 += d; s2 = fmaf(d, (vx[u][k] - mean) * invstd, s2);
      }
  }
  const float2 s = block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    if (a.G == 1) { dgamma[c] = s.y; dbeta[c] = s.x; }
    else { atomicAdd(dgamma + c, s.y); atomicAdd(dbeta + c, s.x); }
  }
  const float ga = gamma[c] * invstd, inv_n = 1.f / (float)cnt;
  const float m1 = a.training ? s.x * inv_n : 0.f, m2 = a.training ? s.y * inv_n : 0.f;
  for (int base = 0; base < cnt4; base += U * blockDim.x) {
    float vx[U][W], vy[U][W], vd[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
#pragma unroll
      for (int k = 0; k < W; ++k) { vx[u][k] = mean; vy[u][k] = 1.f; vd[u][k] = 0.f; }
      if (i4 < cnt4) {
        const size_t o = bn_off4(a, g, c, npg, hw4, i4);
        if (VEC) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(x) + o), d = __ldg(reinterpret_cast<const float4*>(dy) + o);
          vx[u][0] = t.x; vx[u][W > 1 ? 1 : 0] = t.y; vx[u][W > 2 ? 2 : 0] = t.z; vx[u][W > 3 ? 3 : 0] = t.w;
          vd[u][0] = d.x; vd[u][W > 1 ? 1 : 0] = d.y; vd[u][W > 2 ? 2 : 0] = d.z; vd[u][W > 3 ? 3 : 0] = d.w;
          if (a.relu) { const float4 q = __ldg(reinterpret_cast<const float4*>(y) + o);
                        vy[u][0] = q.x; vy[u][W > 1 ? 1 : 0] = q.y; vy[u][W > 2 ? 2 : 0] = q.z; vy[u][W > 3 ? 3 : 0] = q.w; }
        } else {
          vx[u][0] = __ldg(x + o); vd[u][0] = __ldg(dy + o);
          if (a.relu) vy[u][0] = __ldg(y + o);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
      if (i4 < cnt4) {
        float od[W], ox[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
          od[k] = (a.relu && !(vy[u][k] > 0.f)) ? 0.f : vd[u][k];
          ox[k] = ga * (od[k] - m1 - (vx[u][k] - mean) * invstd * m2);
        }
        const size_t o = bn_off4(a, g, c, npg, hw4, i4);
        if (VEC) {
          reinterpret_cast<float4*>(dx)[o] = make_float4(ox[0], ox[W > 1 ? 1 : 0], ox[W > 2 ? 2 : 0], ox[W > 3 ? 3 : 0]);
          if (dres) reinterpret_cast<float4*>(dres)[o] = make_float4(od[0], od[W > 1 ? 1 : 0], od[W > 2 ? 2 : 0], od[W > 3 ? 3 : 0]);
        } else {
          dx[o] = ox[0];
          if (dres) dres[o] = od[0];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
extern thread_local cudaError_t g_last_cuda;
int mt_fail_msg(int code, const char* msg);

// MT_BN_CLUSTER=0 keeps the one-CTA-per-channel kernels (A/B runs)
static int bn_cluster_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("MT_BN_CLUSTER"); mode = e ? (atoi(e) != 0) : 1; }
  return mode;
}

extern "C" {

int mt_up2x_relu_add_forward(int planes, int C, int H, int W, const float* in, const float* bias, const float* skip,
                             float* out, unsigned char* mask, void* stream) {
  if (planes <= 0 || C <= 0 || planes % C || H <= 0 || W <= 0 || !in || !out)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_up2x_relu_add_forward: bad argument");
  if (!mt_al16(out) || !mt_al16(skip)) return mt_fail_msg(MT_ERR_BAD_ARG, "mt_up2x_relu_add_forward: out / skip must be 16-byte aligned");
  const long total = (long)planes * 2 * H * W;
#if defined(__CUDACC__)
  if (mt_pdl_enabled()) mt_launch_pdl(k_up2x_fwd, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1, planes, C, H, W, in, bias, skip, out, mask);
  else
#endif
  k_up2x_fwd<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(planes, C, H, W, in, bias, skip, out, mask);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_up2x_relu_add_forward: launch failed");
}

int mt_up2x_relu_add_backward(int planes, int C, int H, int W, const float* dout, const unsigned char* mask, float* din,
                              float* dbias, void* stream) {
  if (planes <= 0 || C <= 0 || planes % C || H <= 0 || W <= 0 || !dout || !mask || !din)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_up2x_relu_add_backward: bad argument");
  const long total = (long)planes * H * W;
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * C, (cudaStream_t)stream);
  k_up2x_bwd<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(planes, C, H, W, dout, mask, din, dbias);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_up2x_relu_add_backward: launch failed");
}

int mt_bn_act_forward(int N, int C, int HW, int groups, int training, int relu, float eps, float momentum,
                      const float* x, const float* residual, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                      void* stream) {
  if (N <= 0 || C <= 0 || HW <= 0 || groups <= 0 || N % groups || !x || !gamma || !beta || !running_mean || !running_var ||
      !y || !save_mean || !save_invstd)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_bn_act_forward: bad argument");
  if (training && (long)(N / groups) * HW < 2) return mt_fail_msg(MT_ERR_BAD_ARG, "mt_bn_act_forward: training needs more than one value per channel");
  BnActArgs a{N, C, HW, groups, training, relu, eps, momentum};
  const bool vec = HW % 4 == 0 && mt_al16(x) && mt_al16(y) && mt_al16(residual);   // else the scalar kernels
  const int cnt = (N / groups) * HW / (vec ? 4 : 1);
  const int bs = cnt >= 1024 ? 256 : (cnt >= 256 ? 128 : 64);
  int clustered = 0;
  if (vec && bn_cluster_mode()) {                        // a thread-block cluster per channel (bn_cluster.cu)
    clustered = bn_cluster_forward(a, x, residual, gamma, beta, running_mean, running_var, y, save_mean, save_invstd, (cudaStream_t)stream);
    if (clustered < 0) return mt_fail_msg(MT_ERR_CUDA, "mt_bn_act_forward: cluster launch failed");
  }
  if (clustered) {
  } else if (vec)
    k_bn_act_fwd<true><<<dim3(C, groups), bs, 0, (cudaStream_t)stream>>>(a, x, residual, gamma, beta, running_mean,
                                                                           running_var, y, save_mean, save_invstd);
  else
    k_bn_act_fwd<false><<<dim3(C, groups), bs, 0, (cudaStream_t)stream>>>(a, x, residual, gamma, beta, running_mean,
                                                                            running_var, y, save_mean, save_invstd);
  if (training && groups > 1)   // sequential EMA over the statistic groups (one BN call per patch in the reference)
    k_bn_ema<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, save_mean, running_mean, running_var);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_bn_act_forward: launch failed");
}

int mt_bn_act_backward(int N, int C, int HW, int groups, int training, int relu, const float* x, const float* y,
                       const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                       float* dx, float* dresidual, float* dgamma, float* dbeta, void* stream) {
  if (N <= 0 || C <= 0 || HW <= 0 || groups <= 0 || N % groups || !x || !y || !dy || !gamma || !save_mean || !save_invstd ||
      !dx || !dgamma || !dbeta)
    return mt_fail_msg(MT_ERR_BAD_ARG, "mt_bn_act_backward: bad argument");
  BnActArgs a{N, C, HW, groups, training, relu, 0.f, 0.f};
  if (groups > 1) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * C, (cudaStream_t)stream);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * C, (cudaStream_t)stream);
  }
  const bool vec = HW % 4 == 0 && mt_al16(x) && mt_al16(y) && mt_al16(dy) && mt_al16(dx) && mt_al16(dresidual);
  const int cnt = (N / groups) * HW / (vec ? 4 : 1);
  const int bs = cnt >= 1024 ? 256 : (cnt >= 256 ? 128 : 64);
  int clustered = 0;
  if (vec && bn_cluster_mode()) {
    clustered = bn_cluster_backward(a, x, y, dy, gamma, save_mean, save_invstd, dx, dresidual, dgamma, dbeta, (cudaStream_t)stream);
    if (clustered < 0) return mt_fail_msg(MT_ERR_CUDA, "mt_bn_act_backward: cluster launch failed");
  }
  if (clustered) {
  } else if (vec)
    k_bn_act_bwd<true><<<dim3(C, groups), bs, 0, (cudaStream_t)stream>>>(a, x, y, dy, gamma, save_mean, save_invstd, dx,
                                                                           dresidual, dgamma, dbeta);
  else
    k_bn_act_bwd<false><<<dim3(C, groups), bs, 0, (cudaStream_t)stream>>>(a, x, y, dy, gamma, save_mean, save_invstd, dx,
                                                                            dresidual, dgamma, dbeta);
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? MT_OK : mt_fail_msg(MT_ERR_CUDA, "mt_bn_act_backward: launch failed");
}

}  // extern "C"
