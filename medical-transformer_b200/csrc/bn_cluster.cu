// bn_cluster.cu -- BatchNorm2d(+residual)(+ReLU) forward / backward with a THREAD-BLOCK CLUSTER per (channel, statistic group).
//
// The round-1 kernels (shell_ops.cu) run one CTA per channel: 8 - 128 CTAs on 148 SMs, and every CTA reads its channel twice
// (statistics, then normalise).  Here a cluster of CS <= 8 CTAs shares one channel: each CTA loads its slice ONCE and keeps it in
// registers (<= 4 float4 per thread), the CS partial sums are exchanged through distributed shared memory (one cluster barrier),
// and every CTA normalises its own registers.  Same arithmetic as shell_ops.cu (shifted fp32 sums, Newton-refined rsqrt), partial
// sums added in rank order (deterministic).  Slices too large for the registers (512-px inputs) are re-read, still CS-fold wider.
#include "common.cuh"
#include "bn_cluster.cuh"
#include "pdl.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int BC_KEEP = 4;       // float4 per thread held between the two passes

__device__ __forceinline__ float2 bc_block_sum2(float a, float b) {
  __shared__ float ra[8], rb[8];
  a = seg_sum(a, 32); b = seg_sum(b, 32);
  const int wid = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { ra[wid] = a; rb[wid] = b; }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sa += ra[k]; sb += rb[k]; }
  return make_float2(sa, sb);
}

// the cluster's total of (a, b): own partial -> shared memory, cluster barrier, rank-ordered sum over the CS partials (DSMEM reads),
// second barrier so that no CTA exits (or reuses `part`) while a peer still reads it
__device__ __forceinline__ float2 bc_cluster_sum2(cg::cluster_group& cl, int CS, float a, float b) {
  __shared__ float2 part;
  const float2 s = bc_block_sum2(a, b);
  if (CS == 1) return s;
  if (threadIdx.x == 0) part = s;
  cl.sync();
  float2 t = make_float2(0.f, 0.f);
  for (int r = 0; r < CS; ++r) {
    const float2 p = *cl.map_shared_rank(&part, r);
    t.x += p.x; t.y += p.y;
  }
  cl.sync();
  return t;
}

__device__ __forceinline__ unsigned bc_off4(const BnActArgs& a, int g, int c, int npg, int hw4, int i4) {
  const int n = i4 / hw4, r = i4 - n * hw4;
  return (unsigned)(((g * npg + n) * a.C + c) * hw4 + r);
}

// grid (C * CS, G), cluster (CS, 1, 1), 256 threads
template <bool HOLD>
__global__ void __launch_bounds__(256)
k_bn_act_fwd_cl(BnActArgs a, int CS, const float4* __restrict__ x, const float4* __restrict__ res, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv, float4* __restrict__ y,
                float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  pdl_entry();
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int c = blockIdx.x / CS, g = blockIdx.y, tid = threadIdx.x;
  const int npg = a.N / a.G, hw4 = a.HW / 4, cnt4 = npg * hw4, cnt = npg * a.HW;
  const int per = (cnt4 + CS - 1) / CS, lo = rank * per, hi = min(cnt4, lo + per);
  float4 v[BC_KEEP];
  unsigned o[BC_KEEP];
  if (HOLD) {
#pragma unroll
    for (int k = 0; k < BC_KEEP; ++k) {
      const int i4 = lo + k * 256 + tid;
      o[k] = i4 < hi ? bc_off4(a, g, c, npg, hw4, i4) : 0xffffffffu;
      if (i4 < hi) v[k] = __ldg(x + o[k]);
    }
  }
  float mean, invstd;
  if (a.training) {
    const float shift = __ldg(reinterpret_cast<const float*>(x) + (size_t)((g * npg) * a.C + c) * a.HW);
    float s1 = 0.f, s2 = 0.f;
    if (HOLD) {
#pragma unroll
      for (int k = 0; k < BC_KEEP; ++k)
        if (o[k] != 0xffffffffu) {
          const float d0 = v[k].x - shift, d1 = v[k].y - shift, d2 = v[k].z - shift, d3 = v[k].w - shift;
          s1 += (d0 + d1) + (d2 + d3);
          s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
        }
    } else {
      for (int i4 = lo + tid; i4 < hi; i4 += 256) {
        const float4 t = __ldg(x + bc_off4(a, g, c, npg, hw4, i4));
        const float d0 = t.x - shift, d1 = t.y - shift, d2 = t.z - shift, d3 = t.w - shift;
        s1 += (d0 + d1) + (d2 + d3);
        s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
      }
    }
    const float2 s = bc_cluster_sum2(cl, CS, s1, s2);
    const float inv_n = 1.f / (float)cnt;
    const float ms = s.x * inv_n;
    float var = s.y * inv_n - ms * ms;
    var = var > 0.f ? var : 0.f;
    mean = ms + shift;
    invstd = rsqrt_nr(var + a.eps);
    if (tid == 0 && rank == 0) {
      save_mean[g * a.C + c] = mean; save_invstd[g * a.C + c] = invstd;
      save_mean[(a.G + g) * a.C + c] = var;                 // biased variance, for the grouped EMA kernel
      if (a.G == 1) {
        const float unb = cnt > 1 ? (float)cnt / (float)(cnt - 1) : 1.f;
        rm[c] = (1.f - a.momentum) * rm[c] + a.momentum * mean;
        rv[c] = (1.f - a.momentum) * rv[c] + a.momentum * var * unb;
      }
    }
  } else {
    mean = rm[c]; invstd = rsqrt_nr(rv[c] + a.eps);
    if (tid == 0 && rank == 0) { save_mean[g * a.C + c] = mean; save_invstd[g * a.C + c] = invstd; }
  }
  const float ga = gamma[c] * invstd, be = beta[c] - mean * ga;
  auto emit = [&](unsigned off, const float4& t) {
    float4 q = res ? __ldg(res + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    q.x += fmaf(ga, t.x, be); q.y += fmaf(ga, t.y, be); q.z += fmaf(ga, t.z, be); q.w += fmaf(ga, t.w, be);
    if (a.relu) {
      q.x = q.x > 0.f ? q.x : 0.f; q.y = q.y > 0.f ? q.y : 0.f; q.z = q.z > 0.f ? q.z : 0.f; q.w = q.w > 0.f ? q.w : 0.f;
    }
    y[off] = q;
  };
  if (HOLD) {
#pragma unroll
    for (int k = 0; k < BC_KEEP; ++k)
      if (o[k] != 0xffffffffu) emit(o[k], v[k]);
  } else {
    for (int i4 = lo + tid; i4 < hi; i4 += 256) {
      const unsigned off = bc_off4(a, g, c, npg, hw4, i4);
      emit(off, __ldg(x + off));
    }
  }
}

template <bool HOLD>
__global__ void __launch_bounds__(256)
k_bn_act_bwd_cl(BnActArgs a, int CS, const float4* __restrict__ x, const float4* __restrict__ y, const float4* __restrict__ dy,
                const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                float4* __restrict__ dx, float4* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_entry();
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int c = blockIdx.x / CS, g = blockIdx.y, tid = threadIdx.x;
  const int npg = a.N / a.G, hw4 = a.HW / 4, cnt4 = npg * hw4, cnt = npg * a.HW;
  const int per = (cnt4 + CS - 1) / CS, lo = rank * per, hi = min(cnt4, lo + per);
  const float mean = save_mean[g * a.C + c], invstd = save_invstd[g * a.C + c];
  // masked dy and normalised x of one float4
  auto fetch = [&](unsigned off, float4& d, float4& xh) {
    const float4 t = __ldg(x + off);
    d = __ldg(dy + off);
    if (a.relu) {
      const float4 q = __ldg(y + off);
      d.x = q.x > 0.f ? d.x : 0.f; d.y = q.y > 0.f ? d.y : 0.f; d.z = q.z > 0.f ? d.z : 0.f; d.w = q.w > 0.f ? d.w : 0.f;
    }
    xh.x = (t.x - mean) * invstd; xh.y = (t.y - mean) * invstd; xh.z = (t.z - mean) * invstd; xh.w = (t.w - mean) * invstd;
  };
  float4 vd[BC_KEEP], vx[BC_KEEP];
  unsigned o[BC_KEEP];
  float s1 = 0.f, s2 = 0.f;
  if (HOLD) {
#pragma unroll
    for (int k = 0; k < BC_KEEP; ++k) {
      const int i4 = lo + k * 256 + tid;
      o[k] = i4 < hi ? bc_off4(a, g, c, npg, hw4, i4) : 0xffffffffu;
      if (i4 < hi) {
        fetch(o[k], vd[k], vx[k]);
        s1 += (vd[k].x + vd[k].y) + (vd[k].z + vd[k].w);
        s2 = fmaf(vd[k].x, vx[k].x, s2); s2 = fmaf(vd[k].y, vx[k].y, s2); s2 = fmaf(vd[k].z, vx[k].z, s2); s2 = fmaf(vd[k].w, vx[k].w, s2);
      }
    }
  } else {
    for (int i4 = lo + tid; i4 < hi; i4 += 256) {
      float4 d, xh;
      fetch(bc_off4(a, g, c, npg, hw4, i4), d, xh);
      s1 += (d.x + d.y) + (d.z + d.w);
      s2 = fmaf(d.x, xh.x, s2); s2 = fmaf(d.y, xh.y, s2); s2 = fmaf(d.z, xh.z, s2); s2 = fmaf(d.w, xh.w, s2);
    }
  }
  const float2 s = bc_cluster_sum2(cl, CS, s1, s2);
  if (tid == 0 && rank == 0) {
    if (a.G == 1) { dgamma[c] = s.y; dbeta[c] = s.x; }
    else { atomicAdd(dgamma + c, s.y); atomicAdd(dbeta + c, s.x); }
  }
  const float ga = gamma[c] * invstd, inv_n = 1.f / (float)cnt;
  const float m1 = a.training ? s.x * inv_n : 0.f, m2 = a.training ? s.y * inv_n : 0.f;
  auto emit = [&](unsigned off, const float4& d, const float4& xh) {
    dx[off] = make_float4(ga * (d.x - m1 - xh.x * m2), ga * (d.y - m1 - xh.y * m2), ga * (d.z - m1 - xh.z * m2),
                          ga * (d.w - m1 - xh.w * m2));
    if (dres) dres[off] = d;
  };
  if (HOLD) {
#pragma unroll
    for (int k = 0; k < BC_KEEP; ++k)
      if (o[k] != 0xffffffffu) emit(o[k], vd[k], vx[k]);
  } else {
    for (int i4 = lo + tid; i4 < hi; i4 += 256) {
      const unsigned off = bc_off4(a, g, c, npg, hw4, i4);
      float4 d, xh;
      fetch(off, d, xh);
      emit(off, d, xh);
    }
  }
}

// cluster size: enough CTAs to fill the GPU, slices small enough for the register copy where possible
void bc_plan(const BnActArgs& a, int& cs, bool& hold) {
  const int cnt4 = (a.N / a.G) * (a.HW / 4);
  cs = 1;
  while (cs < 8 && ((cnt4 + cs - 1) / cs > 256 * BC_KEEP || (long)a.C * a.G * cs < 96) && cnt4 / (2 * cs) >= 32) cs *= 2;
  hold = (cnt4 + cs - 1) / cs <= 256 * BC_KEEP;
}

template <typename K, typename... Args>
int bc_launch(K kern, const BnActArgs& a, int cs, cudaStream_t st, Args... args) {
  return mt_launch_pdl(kern, dim3((unsigned)(a.C * cs), (unsigned)a.G, 1), dim3(256, 1, 1), 0, st, cs, args...) == cudaSuccess ? 1 : -1;
}

}  // namespace

int bn_cluster_forward(const BnActArgs& a, const float* x, const float* res, const float* gamma, const float* beta, float* rm,
                       float* rv, float* y, float* save_mean, float* save_invstd, cudaStream_t st) {
  if ((long)a.N * a.C * (a.HW / 4) > 0x7fffffffL || a.G > 65535) return 0;
  int cs; bool hold;
  bc_plan(a, cs, hold);
  const float4 *x4 = reinterpret_cast<const float4*>(x), *r4 = reinterpret_cast<const float4*>(res);
  float4* y4 = reinterpret_cast<float4*>(y);
  return hold ? bc_launch(k_bn_act_fwd_cl<true>, a, cs, st, a, cs, x4, r4, gamma, beta, rm, rv, y4, save_mean, save_invstd)
              : bc_launch(k_bn_act_fwd_cl<false>, a, cs, st, a, cs, x4, r4, gamma, beta, rm, rv, y4, save_mean, save_invstd);
}

int bn_cluster_backward(const BnActArgs& a, const float* x, const float* y, const float* dy, const float* gamma,
                        const float* save_mean, const float* save_invstd, float* dx, float* dres, float* dgamma, float* dbeta,
                        cudaStream_t st) {
  if ((long)a.N * a.C * (a.HW / 4) > 0x7fffffffL || a.G > 65535) return 0;
  int cs; bool hold;
  bc_plan(a, cs, hold);
  const float4 *x4 = reinterpret_cast<const float4*>(x), *y4 = reinterpret_cast<const float4*>(y), *d4 = reinterpret_cast<const float4*>(dy);
  float4 *dx4 = reinterpret_cast<float4*>(dx), *dr4 = reinterpret_cast<float4*>(dres);
  return hold ? bc_launch(k_bn_act_bwd_cl<true>, a, cs, st, a, cs, x4, y4, d4, gamma, save_mean, save_invstd, dx4, dr4, dgamma, dbeta)
              : bc_launch(k_bn_act_bwd_cl<false>, a, cs, st, a, cs, x4, y4, d4, gamma, save_mean, save_invstd, dx4, dr4, dgamma, dbeta);
}
