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
#define MT_NREP 8           // max replicated fp64 accumulator slots (CTA picks blockIdx % nrep): cuts same-address atomic
                            // serialisation; used only when there are few statistic groups (m.nrep)

struct Dims {
  int N, C, H, W, axis, stride, variant, training, ngroups;
  int nrep;                // accumulator replicas in use: 8 when ngroups < 4, else 1 (groups already spread the atomics)
  int dbg;                 // MT_DEBUG_FLAGS env (timing experiments only): bit0 = skip global statistic atomics, bit1 = force the v1 attention kernels, bit2 = skip the attention main loops, bit3 = tiled wrapped-diagonal d-relative pass, bit4 = rows / columns of k_bwd2 in separate CTAs (bits 3-4: unmeasured experiments, see axial_attn2_bwd.cu), bit5 = round-1 launch chain instead of the persistent cooperative kernels
  int L, other, B, Bg, gp, d, nt, no, LP, pos;
  float eps, momentum;
  double inv_n1, inv_n2;   // 1/(Bg*L), 1/(Bg*L*L): element counts of the BN1d / BN2d statistics
  float unb1, unb2;        // n/(n-1) for the running (unbiased) variance
  // fp64 statistic accumulator layout (per statistic group)
  int f_qkv_sum, f_qkv_sq, f_sim_sum, f_sim_sq, f_out_sum, f_out_sq, nstat_f;
  int b_out_s1, b_out_s2, b_sim_t, b_qkv_s1, b_qkv_s2, nstat_b;
  // saved-buffer layout (byte offsets)
  size_t sv_stats, sv_cnt, sv_coef, sv_yraw, sv_oraw, sv_lse, sv_tab, sv_bytes, sv_zero_bytes;
  size_t ws_bstats, ws_cnt, ws_coef, ws_gs, ws_dyh, ws_bytes, ws_zero_bytes;
  // float coefficient tables written once by the last CTA of each statistics-producing kernel
  int cf_q, cf_s, cf_o, ncoef_f;      // fwd, per group: q [2C][2] (a,b) | s [nt*G][2] (a*log2e, a) | o [no*C][2] (a,b)
  int cb_o, cb_s, cb_q, ncoef_b;      // bwd, per group: o [no*C][3] | s [nt*G][3] | q [2C][3]   (A, Bc, Cc)
};

static inline size_t mt_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
// activation / scratch buffers are read with 128-bit loads and TMA bulk copies: the ABI requires 16-byte alignment
static inline bool mt_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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
  m.nrep = m.ngroups < 4 ? MT_NREP : 1;
  m.L = m.axis == 0 ? m.H : m.W;
  m.other = m.axis == 0 ? m.W : m.H;
  m.B = m.N * m.other;
  m.Bg = m.B / m.ngroups;
  m.gp = m.C / MT_G;
  m.d = m.gp / 2;
  m.pos = m.variant != MT_VARIANT_NOPOS;
  m.nt = m.pos ? 3 : 1;
  m.no = m.pos ? 2 : 1;
  { double n1 = (double)m.Bg * m.L, n2 = n1 * m.L; m.inv_n1 = 1.0 / n1; m.inv_n2 = 1.0 / n2;
    m.unb1 = n1 > 1 ? (float)(n1 / (n1 - 1.0)) : 1.f; m.unb2 = n2 > 1 ? (float)(n2 / (n2 - 1.0)) : 1.f; }
  m.LP = (m.L % 32 == 0) ? m.L : m.L + 4;      // padded pitch (16-byte aligned rows) spreads heads over banks when a warp spans heads
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
  m.sv_stats = off; off += sizeof(double) * (size_t)m.nrep * m.ngroups * m.nstat_f;
  m.sv_cnt = off;   off = mt_align(off + 64);          // CTA-done tickets, zeroed together with the accumulators
  m.sv_zero_bytes = off;
  m.sv_coef = off;  off = mt_align(off + sizeof(float) * (size_t)m.ngroups * m.ncoef_f);
  m.sv_yraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * 2 * C * m.L);
  m.sv_oraw = off;  off = mt_align(off + sizeof(float) * (size_t)m.B * m.no * C * m.L);
  m.sv_lse = off;   off = mt_align(off + sizeof(float) * (size_t)m.B * MT_G * m.L);
  // window sums of the Toeplitz tables for the closed-form similarity statistics (axial_coop_fwd.cu), built grid-wide
  m.sv_tab = off;   off = mt_align(off + sizeof(float) * (m.pos ? (size_t)2 * (m.d + m.d * (m.d + 1) / 2) * m.L : 0));
  m.sv_bytes = off;
  off = 0;
  m.ws_bstats = off; off += sizeof(double) * (size_t)m.nrep * m.ngroups * m.nstat_b;
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

// fp64 is used ONLY where cancellation demands it (E[x^2] - E[x]^2, s2 - mu*s1): B200's
// non-tensor fp64 div / sqrt are ~30-instruction dependent chains on a narrow pipe (measured:
// a serial tail of 30+ us when the coefficient math was all-double), so everything else is
// fp32 with a Newton-refined rsqrt.
__device__ __forceinline__ float rsqrt_nr(float x) {
  float r = rsqrtf(x);
  return r * (1.5f - 0.5f * x * r * r);
}
struct Moments { float mu, var; double mu_d; };
__device__ __forceinline__ Moments moments(double sum, double sq, double inv_n) {
  Moments r;
  r.mu_d = sum * inv_n;
  double v = fma(-r.mu_d, r.mu_d, sq * inv_n);
  r.mu = (float)r.mu_d; r.var = v > 0.0 ? (float)v : 0.f;
  return r;
}

// BN as a per-channel affine map of the RAW (un-gated) value x:  y = a*x + b.
//   training : mean/var from the fp64 accumulators (sum, sumsq; inv_n = 1/count)
//   eval     : running stats
// `f` is the gate that multiplies x BEFORE the BN (1 when there is none).
struct BnCoef { float a, b, mean, var; };
__device__ __forceinline__ BnCoef bn_coef(int training, double sum, double sq, double inv_n, float f,
                                          float gamma, float beta, float rm, float rv, float eps) {
  BnCoef c;
  if (training) {
    Moments mo = moments(sum, sq, inv_n);
    const float inv = rsqrt_nr(f * f * mo.var + eps);        // 1/sqrt(variance of the gated value + eps)
    c.a = gamma * f * inv; c.b = beta - c.a * mo.mu;
    c.mean = mo.mu; c.var = mo.var;
  } else {
    const float inv = rsqrt_nr(rv + eps);
    c.a = gamma * f * inv; c.b = beta - gamma * rm * inv;
    c.mean = rm; c.var = rv;
  }
  return c;
}

// BN backward as an affine map:  dx_raw = A*dy + Bc + Cc*x_raw   (x_raw un-gated).
//   s1 = sum dy, s2 = sum dy*x_raw over the channel's elements (training only).
struct BnBwdCoef { float A, Bc, Cc; };
__device__ __forceinline__ BnBwdCoef bn_bwd_coef(int training, double sum, double sq, double inv_n, float f,
                                                 float gamma, float rv, float eps, double s1, double s2) {
  BnBwdCoef r;
  if (training) {
    Moments mo = moments(sum, sq, inv_n);
    const float inv = rsqrt_nr(f * f * mo.var + eps);
    const float a = gamma * f * inv;
    const float d = (float)(fma(-mo.mu_d, s1, s2) * inv_n);                   // (s2 - mu*s1)/n
    const float cc = -a * f * f * d * inv * inv;
    r.A = a; r.Cc = cc; r.Bc = -a * (float)(s1 * inv_n) - cc * mo.mu;
  } else {
    r.A = gamma * f * rsqrt_nr(rv + eps); r.Bc = 0.f; r.Cc = 0.f;
  }
  return r;
}

// replicated accumulators: layout [MT_NREP][ngroups][nstat]
__device__ __forceinline__ double* stat_slot(const Dims& m, double* base, int grp, int nstat, int cta) {
  return base + ((size_t)(cta & (m.nrep - 1)) * m.ngroups + grp) * nstat;
}
// ONE out-of-line copy: inlined, each of the ~40 call sites of a cooperative kernel carried ~100 SASS lines (8 address
// computations, 8 loads, 7 fp64 adds) -- 25 % of a 260 KB kernel whose warps execute almost every instruction exactly once, so
// that instruction fetch is its top stall (ncu: stall_no_inst 41 % of the samples of k_axial_fwd<4,1,2>, profiles/r3_notes.md)
static __device__ __noinline__ double stat_get_rep(const double* p, size_t rep_stride) {
  double v[MT_NREP], s = 0.0;
#pragma unroll
  for (int r = 0; r < MT_NREP; ++r) v[r] = __ldcg(p + (size_t)r * rep_stride);   // 8 loads in flight
#pragma unroll
  for (int r = 0; r < MT_NREP; ++r) s += v[r];
  return s;
}
__device__ __forceinline__ double stat_get(const Dims& m, const double* base, int grp, int nstat, int off) {
  if (m.nrep == 1) return __ldcg(base + (size_t)grp * nstat + off);
  return stat_get_rep(base + (size_t)grp * nstat + off, (size_t)m.ngroups * nstat);
}

// sequential EMA of the running statistics over the statistic groups (one BN call each)
__device__ __forceinline__ void ema_update(const Dims& m, float* rm, float* rv, int ch, const double* stats, int nstat,
                                           int off_sum, int off_sq, double inv_n, float unbias, float f) {
  float m_ = rm[ch], v_ = rv[ch];
  const float mom = m.momentum;
  for (int g = 0; g < m.ngroups; ++g) {
    Moments mo = moments(stat_get(m, stats, g, nstat, off_sum + ch), stat_get(m, stats, g, nstat, off_sq + ch), inv_n);
    m_ = (1.f - mom) * m_ + mom * (f * mo.mu);
    v_ = (1.f - mom) * v_ + mom * (f * f * mo.var * unbias);
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
// (training: the running statistics are NOT read here -- with ngroups > 1 another thread's ema_update() below writes the
//  same channel without a barrier in between; tests/emu/run_tsan.sh found that read, which bn_coef ignored anyway)
__device__ __forceinline__ void fin_qkv(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = m.inv_n1;
  for (int k = threadIdx.x; k < m.ngroups * 2 * m.C; k += blockDim.x) {
    const int g = k / (2 * m.C), ch = k - g * 2 * m.C;
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch) : 0.0, n, 1.f, p.bn_qkv_w[ch],
                       p.bn_qkv_b[ch], m.training ? 0.f : p.bn_qkv_rm[ch], m.training ? 1.f : p.bn_qkv_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_q + 2 * ch;
    o[0] = c.a; o[1] = c.b;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < 2 * m.C; ch += blockDim.x)
      ema_update(m, p.bn_qkv_rm, p.bn_qkv_rv, ch, stats, m.nstat_f, m.f_qkv_sum, m.f_qkv_sq, n, m.unb1, 1.f);
}
__device__ __forceinline__ void fin_sim(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = m.inv_n2;
  const int nch = m.nt * MT_G;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch, t = ch / MT_G;
    const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch) : 0.0, n, f, p.bn_sim_w[ch],
                       p.bn_sim_b[ch], m.training ? 0.f : p.bn_sim_rm[ch], m.training ? 1.f : p.bn_sim_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_s + 2 * ch;
    o[0] = c.a * MT_LOG2E; o[1] = c.a;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
      const int t = ch / MT_G;
      const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
      ema_update(m, p.bn_sim_rm, p.bn_sim_rv, ch, stats, m.nstat_f, m.f_sim_sum, m.f_sim_sq, n, m.unb2, f);
    }
}
__device__ __forceinline__ void fin_out(const Dims& m, const MtAxialParams& p, const double* stats, float* coef) {
  const double n = m.inv_n1;
  const int nch = m.no * m.C;
  for (int k = threadIdx.x; k < m.ngroups * nch; k += blockDim.x) {
    const int g = k / nch, ch = k - g * nch;
    const float f = (m.no == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
    BnCoef c = bn_coef(m.training, m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch) : 0.0,
                       m.training ? stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch) : 0.0, n, f, p.bn_out_w[ch],
                       p.bn_out_b[ch], m.training ? 0.f : p.bn_out_rm[ch], m.training ? 1.f : p.bn_out_rv[ch], m.eps);
    float* o = coef + (size_t)g * m.ncoef_f + m.cf_o + 2 * ch;
    o[0] = c.a; o[1] = c.b;
  }
  if (m.training)
    for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
      const float f = (m.no == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
      ema_update(m, p.bn_out_rm, p.bn_out_rv, ch, stats, m.nstat_f, m.f_out_sum, m.f_out_sq, n, m.unb1, f);
    }
}

// ---- backward coefficient finalisers ------------------------------------------------------------
__device__ __forceinline__ void fin_bwd_out(const Dims& m, const MtAxialParams& p, const double* stats,
                                            const double* bstats, float* coef) {
  const double n = m.inv_n1;
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
  const double n = m.inv_n2;
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
  const double n = m.inv_n1;
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
  // d gamma = sum dy*xhat, d beta = sum dy, gate grads in closed form (eps/sigma^3 identity);
  // only (s2 - mu*s1) is formed in fp64.
  const int tid = threadIdx.x, nthr = blockDim.x, C = m.C, NO = m.no, NOC = NO * C, NT = m.nt;
  for (int ch = tid; ch < 2 * C; ch += nthr) {
    float dg = 0.f, db = 0.f;
    for (int g = 0; g < m.ngroups; ++g) {
      const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s1 + ch), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s2 + ch);
      if (m.training) {
        Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch), m.inv_n1);
        dg += (float)fma(-mo.mu_d, s1, s2) * rsqrt_nr(mo.var + m.eps);
      } else {
        dg += (float)fma(-(double)p.bn_qkv_rm[ch], s1, s2) * rsqrt_nr(p.bn_qkv_rv[ch] + m.eps);
      }
      db += (float)s1;
    }
    if (gr.dbn_qkv_w) gr.dbn_qkv_w[ch] = dg;
    if (gr.dbn_qkv_b) gr.dbn_qkv_b[ch] = db;
  }
  float gpart[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ch = tid; ch < NOC; ch += nthr) {
    const int sidx = NO == 2 ? (ch & 1) : 0;
    const float f = sidx ? gate_val(p.f_sve) : gate_val(p.f_sv);
    const float gam = p.bn_out_w[ch];
    float dg = 0.f, db = 0.f;
    for (int g = 0; g < m.ngroups; ++g) {
      const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s1 + ch / NO), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s2 + ch);
      if (m.training) {
        Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch), m.inv_n1);
        const float inv = rsqrt_nr(f * f * mo.var + m.eps), d = (float)fma(-mo.mu_d, s1, s2);
        dg += f * d * inv;
        gpart[2 + sidx] += gam * d * m.eps * inv * inv * inv;
      } else {
        const float inv = rsqrt_nr(p.bn_out_rv[ch] + m.eps);
        dg += (f * (float)s2 - p.bn_out_rm[ch] * (float)s1) * inv;
        gpart[2 + sidx] += gam * inv * (float)s2;
      }
      db += (float)s1;
    }
    if (gr.dbn_out_w) gr.dbn_out_w[ch] = dg;
    if (gr.dbn_out_b) gr.dbn_out_b[ch] = db;
  }
  if (tid < NT * MT_G) {
    const int ch = tid, t = ch / MT_G;
    const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
    const float gam = p.bn_sim_w[ch];
    float dg = 0.f;
    for (int g = 0; g < m.ngroups; ++g) {
      const float T = (float)stat_get(m, bstats, g, m.nstat_b, m.b_sim_t + ch);
      if (m.training) {
        Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch), m.inv_n2);
        const float inv = rsqrt_nr(f * f * mo.var + m.eps);
        dg += f * T * inv;
        if (t > 0) gpart[t - 1] += gam * T * m.eps * inv * inv * inv;
      } else {
        const float inv = rsqrt_nr(p.bn_sim_rv[ch] + m.eps);
        dg += f * T * inv;
        if (t > 0) gpart[t - 1] += gam * inv * T;
      }
    }
    if (gr.dbn_sim_w) gr.dbn_sim_w[ch] = dg;
    if (gr.dbn_sim_b) gr.dbn_sim_b[ch] = 0.f;        // sum dZ == 0: softmax is shift invariant
  }
  // block reduce of the four gate partials: warp shuffles, then a fixed-order sum over warps
  __shared__ float gred[4][32];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = seg_sum(gpart[k], 32);
    if ((tid & 31) == 0) gred[k][tid >> 5] = v;
  }
  __syncthreads();
  if (tid < 4 && gr.dgates) {
    float t = 0.f;
    for (int k = 0; k < (nthr + 31) / 32; ++k) t += gred[tid][k];
    gr.dgates[tid] = t;
  }
}

// Grid-distributed form of bwd_param_grads for the cooperative backward kernel: item k of [0, 2C + NO*C + NT*8) is one
// BN channel; thread `first`, `first + stride`, ... of the whole grid take the items, the four gate gradients are summed
// with atomics into gr.dgates (zeroed by the kernel before its first barrier).
__device__ __forceinline__ void bwd_param_grads_dist(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr,
                                                     const double* stats, const double* bstats, int first, int stride) {
  const int C = m.C, NO = m.no, NOC = NO * C, NT = m.nt, nitems = 2 * C + NOC + NT * MT_G;
  for (int item = first; item < nitems; item += stride) {
    if (item < 2 * C) {
      const int ch = item;
      float dg = 0.f, db = 0.f;
      for (int g = 0; g < m.ngroups; ++g) {
        const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s1 + ch), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_qkv_s2 + ch);
        if (m.training) {
          Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_qkv_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_qkv_sq + ch), m.inv_n1);
          dg += (float)fma(-mo.mu_d, s1, s2) * rsqrt_nr(mo.var + m.eps);
        } else {
          dg += (float)fma(-(double)p.bn_qkv_rm[ch], s1, s2) * rsqrt_nr(p.bn_qkv_rv[ch] + m.eps);
        }
        db += (float)s1;
      }
      if (gr.dbn_qkv_w) gr.dbn_qkv_w[ch] = dg;
      if (gr.dbn_qkv_b) gr.dbn_qkv_b[ch] = db;
    } else if (item < 2 * C + NOC) {
      const int ch = item - 2 * C, sidx = NO == 2 ? (ch & 1) : 0;
      const float f = sidx ? gate_val(p.f_sve) : gate_val(p.f_sv);
      const float gam = p.bn_out_w[ch];
      float dg = 0.f, db = 0.f, gp_ = 0.f;
      for (int g = 0; g < m.ngroups; ++g) {
        const double s1 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s1 + ch / NO), s2 = stat_get(m, bstats, g, m.nstat_b, m.b_out_s2 + ch);
        if (m.training) {
          Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_out_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_out_sq + ch), m.inv_n1);
          const float inv = rsqrt_nr(f * f * mo.var + m.eps), d = (float)fma(-mo.mu_d, s1, s2);
          dg += f * d * inv;
          gp_ += gam * d * m.eps * inv * inv * inv;
        } else {
          const float inv = rsqrt_nr(p.bn_out_rv[ch] + m.eps);
          dg += (f * (float)s2 - p.bn_out_rm[ch] * (float)s1) * inv;
          gp_ += gam * inv * (float)s2;
        }
        db += (float)s1;
      }
      if (gr.dbn_out_w) gr.dbn_out_w[ch] = dg;
      if (gr.dbn_out_b) gr.dbn_out_b[ch] = db;
      if (gr.dgates) atomicAdd(gr.dgates + 2 + sidx, gp_);
    } else {
      const int ch = item - 2 * C - NOC, t = ch / MT_G;
      const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
      const float gam = p.bn_sim_w[ch];
      float dg = 0.f, gp_ = 0.f;
      for (int g = 0; g < m.ngroups; ++g) {
        const float T = (float)stat_get(m, bstats, g, m.nstat_b, m.b_sim_t + ch);
        if (m.training) {
          Moments mo = moments(stat_get(m, stats, g, m.nstat_f, m.f_sim_sum + ch), stat_get(m, stats, g, m.nstat_f, m.f_sim_sq + ch), m.inv_n2);
          const float inv = rsqrt_nr(f * f * mo.var + m.eps);
          dg += f * T * inv;
          gp_ += gam * T * m.eps * inv * inv * inv;
        } else {
          const float inv = rsqrt_nr(p.bn_sim_rv[ch] + m.eps);
          dg += f * T * inv;
          gp_ += gam * inv * T;
        }
      }
      if (gr.dbn_sim_w) gr.dbn_sim_w[ch] = dg;
      if (gr.dbn_sim_b) gr.dbn_sim_b[ch] = 0.f;        // sum dZ == 0: softmax is shift invariant
      if (t > 0 && gr.dgates) atomicAdd(gr.dgates + (t - 1), gp_);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched staging.  ncu (profiles/r1_notes.md): the first kernels spent most of their cycles in
// "long scoreboard" stalls because global->shared staging loops did load -> use -> store one
// element at a time (16-64 serialised L2 round trips per thread).  These helpers issue U
// independent 16-byte loads per thread before touching any of them.
//
// rows [nrows][L] contiguous in global (L % 4 == 0, 16-byte aligned) -> dst[row][LP], with the
// per-row affine a[row % ncoef] * x + b[row % ncoef].
template <int U>
__device__ __forceinline__ void stage_rows_affine(const float* __restrict__ src, int nrows, int L, int LP,
                                                  const float* a, const float* b, int ncoef, float* dst) {
  const int L4 = L >> 2, total4 = nrows * L4;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int base = 0; base < total4; base += U * blockDim.x) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
      v[u] = i4 < total4 ? __ldg(s4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i4 = base + u * blockDim.x + threadIdx.x;
      if (i4 < total4) {
        const int row = i4 / L4, l = (i4 - row * L4) << 2, rc = row % ncoef;
        const float ka = a[rc], kb = b[rc];
        float4 o = make_float4(fmaf(ka, v[u].x, kb), fmaf(ka, v[u].y, kb), fmaf(ka, v[u].z, kb), fmaf(ka, v[u].w, kb));
        *reinterpret_cast<float4*>(dst + (size_t)row * LP + l) = o;
      }
    }
  }
}
// scalar fallback (L % 4 != 0)
__device__ __forceinline__ void stage_rows_affine_scalar(const float* __restrict__ src, int nrows, int L, int LP,
                                                         const float* a, const float* b, int ncoef, float* dst) {
  for (int idx = threadIdx.x; idx < nrows * L; idx += blockDim.x) {
    const int row = idx / L, l = idx - row * L, rc = row % ncoef;
    dst[(size_t)row * LP + l] = fmaf(a[rc], __ldg(src + idx), b[rc]);
  }
}
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int nrows, int L, int LP, const float* a,
                                           const float* b, int ncoef, float* dst) {
  if ((L & 3) == 0 && (LP & 3) == 0) stage_rows_affine<4>(src, nrows, L, LP, a, b, ncoef, dst);
  else stage_rows_affine_scalar(src, nrows, L, LP, a, b, ncoef, dst);
}
