// common.cuh -- shared geometry, statistic layout and device helpers for the axial kernels.
// sm_100a only.  All arithmetic fp32 with fp64 statistic accumulators (see DESIGN.md:
// bf16/fp16 operands break the reference's 1e-2 train-mode parity budget).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/medt_axial.h"

#define MT_G 8              // heads: fixed by the reference (axialnet.py:400)
#define MT_LOG2E 1.4426950408889634f
#define MT_NREP 8           // replicated fp64 accumulator slots (CTA picks blockIdx & 7): cuts same-address atomic serialisation

struct Dims {
  int N, C, H, W, axis, stride, variant, training, ngroups;
  int dbg;                 // MT_DEBUG_FLAGS env (timing experiments only): bit0 = skip global statistic atomics
  int L, other, B, Bg, gp, d, nt, no, LP, pos;
  float eps, momentum;
  // fp64 statistic accumulator layout (per statistic group)
  int f_qkv_sum, f_qkv_sq, f_sim_sum, f_sim_sq, f_out_sum, f_out_sq, nstat_f;
  int b_out_s1, b_out_s2, b_sim_t, b_qkv_s1, b_qkv_s2, nstat_b;
  // saved-buffer layout (byte offsets)
  size_t sv_stats, sv_yraw, sv_oraw, sv_lse, sv_bytes;
  size_t ws_bstats, ws_gs, ws_dyh, ws_bytes;
};

static inline size_t mt_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static inline bool make_dims(const MtAxialDesc* d, Dims& m) {
  m.N = d->N; m.C = d->C; m.H = d->H; m.W = d->W; m.axis = d->axis; m.stride = d->stride;
  m.variant = d->variant; m.training = d->training; m.ngroups = d->ngroups;
  m.eps = d->eps; m.momentum = d->momentum;
  { static int flags = -1; if (flags < 0) { const char* e = getenv("MT_DEBUG_FLAGS"); flags = e ? atoi(e) : 0; } m.dbg = flags; }
  if (m.N <= 0 || m.C <= 0 || m.H <= 0 || m.W <= 0 || m.ngroups <= 0) return false;
  if (m.axis != 0 && m.axis != 1) return false;
  if (m.stride < 1 || m.H % m.stride || m.W % m.stride) return false;
  if (m.variant < 0 || m.variant > 2) return false;
  if (m.N % m.ngroups) return false;
  if (m.C % (2 * MT_G)) return false;
  m.L = m.axis == 0 ? m.H : m.W;
  m.other = m.axis == 0 ? m.W : m.H;
  m.B = m.N * m.other;
  m.Bg = m.B / m.ngroups;
  m.gp = m.C / MT_G;
  m.d = m.gp / 2;
  m.pos = m.variant != MT_VARIANT_NOPOS;
  m.nt = m.pos ? 3 : 1;
  m.no = m.pos ? 2 : 1;
  m.LP = (m.L % 32 == 0) ? m.L : m.L + 1;      // odd-ish pitch kills bank conflicts when a warp spans heads
  int C = m.C;
  m.f_qkv_sum = 0; m.f_qkv_sq = 2 * C;
  m.f_sim_sum = 4 * C; m.f_sim_sq = 4 * C + m.nt * MT_G;
  m.f_out_sum = 4 * C + 2 * m.nt * MT_G; m.f_out_sq = m.f_out_sum + m.no * C;
  m.nstat_f = m.f_out_sq + m.no * C;
  m.b_out_s1 = 0; m.b_out_s2 = C; m.b_sim_t = C + m.no * C;
  m.b_qkv_s1 = m.b_sim_t + m.nt * MT_G; m.b_qkv_s2 = m.b_qkv_s1 + 2 * C;
  m.nstat_b = m.b_qkv_s2 + 2 * C;
  size_t off = 0;
  m.sv_stats = off; off = mt_align(off + sizeof(double) * (size_t)MT_NREP * m.ngroups * m.nstat_f);
  m.sv_yraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * 2 * C * m.L);
  m.sv_oraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * m.no * C * m.L);
  m.sv_lse = off;   off = mt_align(off + sizeof(float) * (size_t)m.B * MT_G * m.L);
  m.sv_bytes = off;
  off = 0;
  m.ws_bstats = off; off = mt_align(off + sizeof(double) * (size_t)MT_NREP * m.ngroups * m.nstat_b);
  m.ws_gs = off;     off = mt_align(off + sizeof(float) * (size_t)m.B * C * m.L);
  m.ws_dyh = off;    off = mt_align(off + sizeof(float) * (size_t)m.B * 2 * C * m.L);
  m.ws_bytes = off;
  return true;
}

// ----------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sum over aligned lane segments of `width` lanes (width = power of two <= 32)
__device__ __forceinline__ float seg_sum(float v, int width) {
  for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NCHW address of element l of strip b, channel c
__device__ __forceinline__ size_t nchw_addr(const Dims& m, int b, int c, int l) {
  int n = b / m.other, o = b - n * m.other;
  return m.axis == 0 ? (((size_t)n * m.C + c) * m.H + l) * m.W + o
                     : (((size_t)n * m.C + c) * m.H + o) * m.W + l;
}

// BN as a per-channel affine map of the RAW (un-gated) value x:  y = a*x + b.
//   training : mean/var from the fp64 accumulators (sum, sumsq over n elements)
//   eval     : running stats
// `f` is the gate that multiplies x BEFORE the BN (1 when there is none).
struct BnCoef { float a, b, mean, var; };
__device__ __forceinline__ BnCoef bn_coef(int training, double sum, double sq, double n, float f,
                                          float gamma, float beta, float rm, float rv, float eps) {
  BnCoef c;
  if (training) {
    double mu = sum / n;
    double var = sq / n - mu * mu;
    if (var < 0.0) var = 0.0;
    double su2 = (double)f * f * var + (double)eps;      // variance of the gated value + eps
    double a = (double)gamma * f / sqrt(su2);
    c.a = (float)a; c.b = (float)((double)beta - a * mu);
    c.mean = (float)mu; c.var = (float)var;
  } else {
    double inv = 1.0 / sqrt((double)rv + (double)eps);
    c.a = (float)((double)gamma * f * inv);
    c.b = (float)((double)beta - (double)gamma * rm * inv);
    c.mean = rm; c.var = rv;
  }
  return c;
}

// BN backward as an affine map:  dx_raw = A*dy + Bc + Cc*x_raw   (x_raw un-gated).
//   s1 = sum dy, s2 = sum dy*x_raw over the channel's n elements (training only).
struct BnBwdCoef { float A, Bc, Cc; };
__device__ __forceinline__ BnBwdCoef bn_bwd_coef(int training, double sum, double sq, double n, float f,
                                                 float gamma, float rv, float eps, double s1, double s2) {
  BnBwdCoef r;
  if (training) {
    double mu = sum / n;
    double var = sq / n - mu * mu;
    if (var < 0.0) var = 0.0;
    double su2 = (double)f * f * var + (double)eps;
    double a = (double)gamma * f / sqrt(su2);
    double cc = -a * (double)f * f * (s2 - mu * s1) / (n * su2);
    r.A = (float)a; r.Cc = (float)cc; r.Bc = (float)(-a * s1 / n - cc * mu);
  } else {
    r.A = (float)((double)gamma * f / sqrt((double)rv + (double)eps)); r.Bc = 0.f; r.Cc = 0.f;
  }
  return r;
}

// replicated accumulators: layout [MT_NREP][ngroups][nstat]
__device__ __forceinline__ double* stat_slot(const Dims& m, double* base, int grp, int nstat, int cta) {
  return base + ((size_t)(cta & (MT_NREP - 1)) * m.ngroups + grp) * nstat;
}
__device__ __forceinline__ double stat_get(const Dims& m, const double* base, int grp, int nstat, int off) {
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < MT_NREP; ++r) s += base[((size_t)r * m.ngroups + grp) * nstat + off];
  return s;
}

// sequential EMA of the running statistics over the statistic groups (one BN call each)
__device__ __forceinline__ void ema_update(const Dims& m, float* rm, float* rv, int ch, const double* stats, int nstat,
                                           int off_sum, int off_sq, double n, float f) {
  float m_ = rm[ch], v_ = rv[ch];
  const float mom = m.momentum;
  for (int g = 0; g < m.ngroups; ++g) {
    double sum = stat_get(m, stats, g, nstat, off_sum + ch), sq = stat_get(m, stats, g, nstat, off_sq + ch);
    double mu = sum / n, var = sq / n - mu * mu;
    if (var < 0.0) var = 0.0;
    double mu_u = f * mu, var_u = (double)f * f * var * (n / (n - 1.0));
    m_ = (1.f - mom) * m_ + mom * (float)mu_u;
    v_ = (1.f - mom) * v_ + mom * (float)var_u;
  }
  rm[ch] = m_; rv[ch] = v_;
}

// global (cross-CTA) statistic accumulation; skipped only by the MT_DEBUG_FLAGS timing experiment
#define MT_GADD(m, ptr, val) do { if (!((m).dbg & 1)) atomicAdd((ptr), (val)); } while (0)
