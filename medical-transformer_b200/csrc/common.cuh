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
  size_t sv_stats, sv_cnt, sv_coef, sv_yraw, sv_oraw, sv_lse, sv_bytes, sv_zero_bytes;
  size_t ws_bstats, ws_cnt, ws_coef, ws_gs, ws_dyh, ws_bytes, ws_zero_bytes;
  // float coefficient tables written once by the last CTA of each statistics-producing kernel
  int cf_q, cf_s, cf_o, ncoef_f;      // fwd, per group: q [2C][2] (a,b) | s [nt*G][2] (a*log2e, a) | o [no*C][2] (a,b)
  int cb_o, cb_s, cb_q, ncoef_b;      // bwd, per group: o [no*C][3] | s [nt*G][3] | q [2C][3]   (A, Bc, Cc)
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
  m.cf_q = 0; m.cf_s = 4 * C; m.cf_o = m.cf_s + 2 * m.nt * MT_G; m.ncoef_f = m.cf_o + 2 * m.no * C;
  m.cb_o = 0; m.cb_s = 3 * m.no * C; m.cb_q = m.cb_s + 3 * m.nt * MT_G; m.ncoef_b = m.cb_q + 6 * C;
  size_t off = 0;
  m.sv_stats = off; off += sizeof(double) * (size_t)MT_NREP * m.ngroups * m.nstat_f;
  m.sv_cnt = off;   off = mt_align(off + 64);          // CTA-done tickets, zeroed together with the accumulators
  m.sv_zero_bytes = off;
  m.sv_coef = off;  off = mt_align(off + sizeof(float) * (size_t)m.ngroups * m.ncoef_f);
  m.sv_yraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * 2 * C * m.L);
  m.sv_oraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * m.no * C * m.L);
  m.sv_lse = off;   off = mt_align(off + sizeof(float) * (size_t)m.B * MT_G * m.L);
  m.sv_bytes = off;
  off = 0;
  m.ws_bstats = off; off += sizeof(double) * (size_t)MT_NREP * m.ngroups * m.nstat_b;
  m.ws_cnt = off;    off = mt_align(off + 64);
  m.ws_zero_bytes = off;
  m.ws_coef = off;   off = mt_align(off + sizeof(float) * (size_t)m.ngroups * m.ncoef_b);
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
  for (int r = 0; r < MT_NREP; ++r) s += __ldcg(base + ((size_t)r * m.ngroups + grp) * nstat + off);
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

// "last CTA done": after a CTA has issued its global accumulations it takes a ticket; the CTA
// that draws the last one sees every accumulation (fence + L2 atomics) and finalises.
__device__ __forceinline__ bool last_cta_done(unsigned* counter, unsigned total) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == total - 1u);
  __syncthreads();
  return is_last != 0;
}

__device__ __forceinline__ float gate_val(const float* p) { return p ? __ldg(p) : 1.f; }

// ---- forward coefficient finalisers (all threads of ONE CTA) ---------------------------------
__device__ __forceinline__ void fin_qkv(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = (double)m.Bg * m.L;
  for (int k = threadIdx.x; k < m.ngroups * 2 * m.C; k += blockDim.x) {
    const int g = k / (2 * m.C), ch = k - g * 2 * m.C;
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch) : 0.0, n, 1.f, p.bn_qkv_w[ch],
                       p.bn_qkv_b[ch], p.bn_qkv_rm[ch], p.bn_qkv_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_q + 2 * ch;
    o[0] = c.a; o[1] = c.b;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < 2 * m.C; ch += blockDim.x)
      ema_update(m, p.bn_qkv_rm, p.bn_qkv_rv, ch, stats, m.nstat_f, m.f_qkv_sum, m.f_qkv_sq, n, 1.f);
}
__device__ __forceinline__ void fin_sim(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = (double)m.Bg * m.L * m.L;
  const int nch = m.nt * MT_G;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch, t = ch / MT_G;
    const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch) : 0.0, n, f, p.bn_sim_w[ch],
                       p.bn_sim_b[ch], p.bn_sim_rm[ch], p.bn_sim_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_s + 2 * ch;
    o[0] = c.a * MT_LOG2E; o[1] = c.a;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
      const int t = ch / MT_G;
      const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
      ema_update(m, p.bn_sim_rm, p.bn_sim_rv, ch, stats, m.nstat_f, m.f_sim_sum, m.f_sim_sq, n, f);
    }
}
__device__ __forceinline__ void fin_out(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = (double)m.Bg * m.L;
  const int nch = m.no * m.C;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch;
    const float f = (m.no == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch) : 0.0, n, f, p.bn_out_w[ch],
                       p.bn_out_b[ch], p.bn_out_rm[ch], p.bn_out_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_o + 2 * ch;
    o[0] = c.a; o[1] = c.b;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
      const float f = (m.no == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
      ema_update(m, p.bn_out_rm, p.bn_out_rv, ch, stats, m.nstat_f, m.f_out_sum, m.f_out_sq, n, f);
    }
}

// ---- backward coefficient finalisers ------------------------------------------------------------
__device__ __forceinline__ void fin_bwd_out(const Dims& m, const MtAxialParams& p, const double* stats,
                                            const double* bstats, float* coef) {
  const double n = (double)m.Bg * m.L;
  const int nch = m.no * m.C;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch;
    const float f = (m.no == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
    BnBwdCoef c = bn_bwd_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch) : 0.0,
                              m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch) : 0.0, n, f, p.bn_out_w[ch],
                              p.bn_out_rv[ch], m.eps, m.training ? stat_get(m, bstats, g, m.nstat_b, m.b_out_s1 + ch / m.no) : 0.0,
                              m.training ? stat_get(m, bstats, g, m.nstat_b, m.b_out_s2 + ch) : 0.0);
    float* o = coef + (size_t)g * m.ncoef_b + m.cb_o + 3 * ch;
    o[0] = c.A; o[1] = c.Bc; o[2] = c.Cc;
  }
}
__device__ __forceinline__ void fin_bwd_sim(const Dims& m, const MtAxialParams& p, const double* stats,
                                            const double* bstats, float* coef) {
  const double n = (double)m.Bg * m.L * m.L;
  const int nch = m.nt * MT_G;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch, t = ch / MT_G;
    const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
    // dy = dZ for every term; sum dZ == 0 exactly (softmax rows), so s1 = 0, s2 = T
    BnBwdCoef c = bn_bwd_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch) : 0.0,
                              m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch) : 0.0, n, f, p.bn_sim_w[ch],
                              p.bn_sim_rv[ch], m.eps, 0.0, stat_get(m, bstats, g, m.nstat_b, m.b_sim_t + ch));
    float* o = coef + (size_t)g * m.ncoef_b + m.cb_s + 3 * ch;
    o[0] = c.A; o[1] = c.Bc; o[2] = c.Cc;
  }
}
__device__ __forceinline__ void fin_bwd_qkv(const Dims& m, const MtAxialParams& p, const double* stats,
                                            const double* bstats, float* coef) {
  const double n = (double)m.Bg * m.L;
  for (int k = threadIdx.x; k < m.ngroups * 2 * m.C; k += blockDim.x) {
    const int g = k / (2 * m.C), ch = k - g * 2 * m.C;
    BnBwdCoef c = bn_bwd_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch) : 0.0,
                              m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch) : 0.0, n, 1.f, p.bn_qkv_w[ch],
                              p.bn_qkv_rv[ch], m.eps, m.training ? stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s1 + ch) : 0.0,
                              m.training ? stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s2 + ch) : 0.0);
    float* o = coef + (size_t)g * m.ncoef_b + m.cb_q + 3 * ch;
    o[0] = c.A; o[1] = c.Bc; o[2] = c.Cc;
  }
}

// ============================================================================ parameter gradients
__device__ __forceinline__ void bwd_param_grads(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr,
                                const double* stats, const double* bstats) {
  const int tid = threadIdx.x, nthr = blockDim.x, C = m.C, NO = m.no, NOC = NO * C, NT = m.nt;
  for (int ch = tid; ch < 2 * C; ch += nthr) {
    double dg = 0.0, db = 0.0;
    for (int g = 0; g < m.ngroups; ++g) {
      const double n = (double)m.Bg * m.L;
      const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s1 + ch), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s2 + ch);
      if (m.training) {
        double mu = stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch) / n;
        double var = stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch) / n - mu * mu;
        if (var < 0.0) var = 0.0;
        dg += (s2 - mu * s1) / sqrt(var + (double)m.eps);
      } else {
        dg += (s2 - (double)p.bn_qkv_rm[ch] * s1) / sqrt((double)p.bn_qkv_rv[ch] + (double)m.eps);
      }
      db += s1;
    }
    if (gr.dbn_qkv_w) gr.dbn_qkv_w[ch] = (float)dg;
    if (gr.dbn_qkv_b) gr.dbn_qkv_b[ch] = (float)db;
  }
  double gpart[4] = {0.0, 0.0, 0.0, 0.0};
  for (int ch = tid; ch < NOC; ch += nthr) {
    const int sidx = NO == 2 ? (ch & 1) : 0;
    const float f = sidx ? gate_val(p.f_sve) : gate_val(p.f_sv);
    const double gam = p.bn_out_w[ch];
    double dg = 0.0, db = 0.0;
    for (int g = 0; g < m.ngroups; ++g) {
      const double n = (double)m.Bg * m.L;
      const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s1 + ch / NO), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s2 + ch);
      if (m.training) {
        double mu = stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch) / n;
        double var = stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch) / n - mu * mu;
        if (var < 0.0) var = 0.0;
        const double su2 = (double)f * f * var + (double)m.eps, su = sqrt(su2);
        dg += (double)f * (s2 - mu * s1) / su;
        gpart[2 + sidx] += gam * (s2 - mu * s1) * (double)m.eps / (su2 * su);
      } else {
        const double inv = 1.0 / sqrt((double)p.bn_out_rv[ch] + (double)m.eps);
        dg += ((double)f * s2 - (double)p.bn_out_rm[ch] * s1) * inv;
        gpart[2 + sidx] += gam * inv * s2;
      }
      db += s1;
    }
    if (gr.dbn_out_w) gr.dbn_out_w[ch] = (float)dg;
    if (gr.dbn_out_b) gr.dbn_out_b[ch] = (float)db;
  }
  if (tid < NT * MT_G) {
    const int ch = tid, t = ch / MT_G;
    const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
    const double gam = p.bn_sim_w[ch];
    double dg = 0.0;
    for (int g = 0; g < m.ngroups; ++g) {
      const double n = (double)m.Bg * m.L * m.L, T = stat_get(m, bstats, g, m.nstat_b, m.b_sim_t + ch);
      if (m.training) {
        double mu = stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch) / n;
        double var = stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch) / n - mu * mu;
        if (var < 0.0) var = 0.0;
        const double su2 = (double)f * f * var + (double)m.eps, su = sqrt(su2);
        dg += (double)f * T / su;
        if (t > 0) gpart[t - 1] += gam * T * (double)m.eps / (su2 * su);
      } else {
        const double inv = 1.0 / sqrt((double)p.bn_sim_rv[ch] + (double)m.eps);
        dg += (double)f * T * inv;
        if (t > 0) gpart[t - 1] += gam * inv * T;
      }
    }
    if (gr.dbn_sim_w) gr.dbn_sim_w[ch] = (float)dg;
    if (gr.dbn_sim_b) gr.dbn_sim_b[ch] = 0.f;        // sum dZ == 0: softmax is shift invariant
  }
  __shared__ double gred[4][256];                      // fixed-order block reduce -> deterministic
  for (int k = 0; k < 4; ++k) gred[k][tid] = gpart[k];
  __syncthreads();
  if (tid < 4 && gr.dgates) {
    double t = 0.0;
    for (int k = 0; k < nthr && k < 256; ++k) t += gred[tid][k];
    gr.dgates[tid] = (float)t;
  }
}

