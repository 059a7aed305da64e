// axial_fwd.cu -- forward of one axial-attention op (reference: lib/models/axialnet.py
// :142-189 gated, :52-92 pos, :222-253 nopos), written for sm_100a.
//
// Dataflow (train mode; the three batch-statistic BatchNorms force three grid-wide
// dependencies, SURVEY 7 hard part 1):
//   k_proj       y_raw = W_qkv . x strips        + fp64 sum/sumsq per qkv channel
//   k_sim_stats  sum/sumsq of qk, qr, kr over (B,L,L) per (term, head)  [never materialised]
//   k_attn       softmax_j( sum_t w_t S_t ) -> o_raw = [P v | P Rv], lse   + sum/sumsq per output channel
//   k_out        BN(bn_output) + pair sum + NCHW un-permute + AvgPool(stride)
// Eval mode skips k_sim_stats and every statistic accumulation (running stats are used).
// The (B, 3G, L, L) similarity tensor of the reference (100 MB per call at 128 px, 6.4 GB
// at 512 px) never exists: scores live in registers, one query row per thread.
#include "common.cuh"
extern thread_local cudaError_t g_last_cuda;

// ============================================================================ shared loaders
// qkv BN coefficients (y_hat = a*y_raw + b) of statistic group `grp`, from the table k_proj's last CTA wrote
__device__ __forceinline__ void load_qkv_coef(const Dims& m, const float* __restrict__ coef, int grp, float* ca, float* cb) {
  const float* tb = coef + (size_t)grp * m.ncoef_f + m.cf_q;
  for (int ch = threadIdx.x; ch < 2 * m.C; ch += blockDim.x) { ca[ch] = tb[2 * ch]; cb[ch] = tb[2 * ch + 1]; }
}

// normalised q,k,v strips of `spb` consecutive strips starting at b0 -> Yn[s][ch][LP]
__device__ __forceinline__ void load_strips_norm(const Dims& m, const float* __restrict__ yraw, int b0, int spb,
                                                 const float* ca, const float* cb, float* Yn) {
  stage_rows(yraw + (size_t)b0 * 2 * m.C * m.L, spb * 2 * m.C, m.L, m.LP, ca, cb, 2 * m.C, Yn);
}

__device__ __forceinline__ void load_rel(const Dims& m, const float* __restrict__ rel, float* R) {
  const int nrel = 2 * m.gp * (2 * m.L - 1);
  for (int idx = threadIdx.x; idx < nrel; idx += blockDim.x) R[idx] = __ldg(rel + idx);
}

// ============================================================================ k_sim_stats
template <int D, bool POS>
__global__ void __launch_bounds__(256)
k_sim_stats(Dims m, MtAxialParams p, const float* __restrict__ yraw, double* __restrict__ stats,
            float* __restrict__ coef, unsigned* __restrict__ cnt, int SPB, int segw) {
  extern __shared__ float sm[];
  const int C = m.C, L = m.L, LP = m.LP, RL = 2 * L - 1;
  float* ca = sm;
  float* cb = ca + 2 * C;
  float* red = cb + 2 * C;                 // [3][G][2]
  float* Yn = red + 48;
  float* R = Yn + (size_t)SPB * 2 * C * LP;
  const int b0 = blockIdx.x * SPB;
  const int grp = b0 / m.Bg;
  load_qkv_coef(m, coef, grp, ca, cb);
  if (threadIdx.x < 48) red[threadIdx.x] = 0.f;
  if (POS) load_rel(m, p.relative, R);
  __syncthreads();
  load_strips_norm(m, yraw, b0, SPB, ca, cb, Yn);
  __syncthreads();
  const int nitems = SPB * MT_G * L;
  for (int base = 0; base < nitems; base += blockDim.x) {
    const int item = base + threadIdx.x;
    const bool act = item < nitems;
    const int it = act ? item : 0;
    const int i = it % L, g = (it / L) % MT_G, s = it / (L * MT_G);
    const float* Yb = Yn + ((size_t)s * 2 * C + g * 4 * D) * LP;
    float q[D];
#pragma unroll
    for (int c = 0; c < D; ++c) q[c] = Yb[c * LP + i];
    const float* Kp = Yb + D * LP;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    for (int j = 0; j < L; ++j) {
      float sqk = 0.f, sqr = 0.f, skr = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        float kc = Kp[c * LP + j];
        sqk = fmaf(q[c], kc, sqk);
        if (POS) {
          sqr = fmaf(q[c], R[c * RL + (i - j + L - 1)], sqr);
          skr = fmaf(kc, R[(D + c) * RL + (j - i + L - 1)], skr);
        }
      }
      a0 += sqk; a1 = fmaf(sqk, sqk, a1);
      if (POS) { a2 += sqr; a3 = fmaf(sqr, sqr, a3); a4 += skr; a5 = fmaf(skr, skr, a5); }
    }
    if (!act) { a0 = a1 = a2 = a3 = a4 = a5 = 0.f; }
    a0 = seg_sum(a0, segw); a1 = seg_sum(a1, segw);
    if (POS) { a2 = seg_sum(a2, segw); a3 = seg_sum(a3, segw); a4 = seg_sum(a4, segw); a5 = seg_sum(a5, segw); }
    if (act && (threadIdx.x & (segw - 1)) == 0) {
      atomicAdd(red + (0 * MT_G + g) * 2 + 0, a0); atomicAdd(red + (0 * MT_G + g) * 2 + 1, a1);
      if (POS) {
        atomicAdd(red + (1 * MT_G + g) * 2 + 0, a2); atomicAdd(red + (1 * MT_G + g) * 2 + 1, a3);
        atomicAdd(red + (2 * MT_G + g) * 2 + 0, a4); atomicAdd(red + (2 * MT_G + g) * 2 + 1, a5);
      }
    }
  }
  __syncthreads();
  double* st = stat_slot(m, stats, grp, m.nstat_f, blockIdx.x);
  if (threadIdx.x < m.nt * MT_G) {
    MT_GADD(m, st + m.f_sim_sum + threadIdx.x, (double)red[threadIdx.x * 2 + 0]);
    MT_GADD(m, st + m.f_sim_sq + threadIdx.x, (double)red[threadIdx.x * 2 + 1]);
  }
  if (last_cta_done(cnt + 1, gridDim.x)) fin_sim(m, p, stats, coef);   // softmax weights + bn_similarity running stats
}

// ============================================================================ k_attn
template <int D, bool POS>
__global__ void __launch_bounds__(256)
k_attn(Dims m, MtAxialParams p, const float* __restrict__ yraw, double* __restrict__ stats, float* __restrict__ coef,
       unsigned* __restrict__ cnt, float* __restrict__ oraw, float* __restrict__ lse, int SPB, int segw) {
  constexpr int GP = 2 * D;
  extern __shared__ float sm[];
  const int C = m.C, L = m.L, LP = m.LP, RL = 2 * L - 1;
  const int NO = POS ? 2 : 1;
  float* ca = sm;
  float* cb = ca + 2 * C;
  float* wsim = cb + 2 * C;                // [3][G] score weights (log2e folded in)
  float* red = wsim + 24;                  // [NO*C][2]
  float* Yn = red + 2 * NO * C;
  float* R = Yn + (size_t)SPB * 2 * C * LP;
  const int b0 = blockIdx.x * SPB;
  const int grp = b0 / m.Bg;
  load_qkv_coef(m, coef, grp, ca, cb);
  if (threadIdx.x < m.nt * MT_G) wsim[threadIdx.x] = coef[(size_t)grp * m.ncoef_f + m.cf_s + 2 * threadIdx.x];
  for (int k = threadIdx.x; k < 2 * NO * C; k += blockDim.x) red[k] = 0.f;
  if (POS) load_rel(m, p.relative, R);
  __syncthreads();
  load_strips_norm(m, yraw, b0, SPB, ca, cb, Yn);
  __syncthreads();
  const int nitems = SPB * MT_G * L;
  for (int base = 0; base < nitems; base += blockDim.x) {
    const int item = base + threadIdx.x;
    const bool act = item < nitems;
    const int it = act ? item : 0;
    const int i = it % L, g = (it / L) % MT_G, s = it / (L * MT_G);
    const float* Yb = Yn + ((size_t)s * 2 * C + g * 4 * D) * LP;
    float q[D], qr_[D];
    const float wqk = wsim[g], wqr = POS ? wsim[MT_G + g] : 0.f, wkr = POS ? wsim[2 * MT_G + g] : 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) { q[c] = Yb[c * LP + i] * wqk; qr_[c] = Yb[c * LP + i] * wqr; }
    const float* Kp = Yb + D * LP;
    const float* Vp = Yb + GP * LP;
    float mx = -INFINITY, lsum = 0.f;
    float asv[GP], asve[GP];
#pragma unroll
    for (int c = 0; c < GP; ++c) { asv[c] = 0.f; asve[c] = 0.f; }
    for (int j = 0; j < L; ++j) {
      float z = 0.f, zk = 0.f;
      const int mi = i - j + L - 1, mj = j - i + L - 1;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        float kc = Kp[c * LP + j];
        z = fmaf(q[c], kc, z);
        if (POS) {
          z = fmaf(qr_[c], R[c * RL + mi], z);
          zk = fmaf(kc, R[(D + c) * RL + mj], zk);
        }
      }
      if (POS) z = fmaf(wkr, zk, z);
      if (z > mx) {                          // lazy rescale (rare once the running max settles)
        float sc = ex2(mx - z);
        mx = z; lsum *= sc;
#pragma unroll
        for (int c = 0; c < GP; ++c) { asv[c] *= sc; if (POS) asve[c] *= sc; }
      }
      float pj = ex2(z - mx);
      lsum += pj;
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        asv[c] = fmaf(pj, Vp[c * LP + j], asv[c]);
        if (POS) asve[c] = fmaf(pj, R[(GP + c) * RL + mi], asve[c]);
      }
    }
    const float inv = 1.f / lsum;
    const int b = b0 + s;
    if (act) lse[((size_t)b * MT_G + g) * L + i] = mx + lg2(lsum);
#pragma unroll
    for (int c = 0; c < GP; ++c) {
      const int ch = (g * GP + c) * NO;
      float o0 = asv[c] * inv, o1 = POS ? asve[c] * inv : 0.f;
      if (act) {
        oraw[((size_t)b * NO * C + ch) * L + i] = o0;
        if (POS) oraw[((size_t)b * NO * C + ch + 1) * L + i] = o1;
      }
      if (m.training) {
        if (!act) { o0 = 0.f; o1 = 0.f; }
        float s0 = seg_sum(o0, segw), q0 = seg_sum(o0 * o0, segw);
        float s1 = 0.f, q1 = 0.f;
        if (POS) { s1 = seg_sum(o1, segw); q1 = seg_sum(o1 * o1, segw); }
        if (act && (threadIdx.x & (segw - 1)) == 0) {
          atomicAdd(red + ch * 2, s0); atomicAdd(red + ch * 2 + 1, q0);
          if (POS) { atomicAdd(red + (ch + 1) * 2, s1); atomicAdd(red + (ch + 1) * 2 + 1, q1); }
        }
      }
    }
  }
  if (m.training) {
    __syncthreads();
    double* st = stat_slot(m, stats, grp, m.nstat_f, blockIdx.x);
    for (int ch = threadIdx.x; ch < NO * C; ch += blockDim.x) {
      MT_GADD(m, st + m.f_out_sum + ch, (double)red[ch * 2]);
      MT_GADD(m, st + m.f_out_sq + ch, (double)red[ch * 2 + 1]);
    }
  }
  if (last_cta_done(cnt + 2, gridDim.x)) fin_out(m, p, stats, coef);      // bn_output coefficients + running stats
}

// ============================================================================ host launchers
int launch_proj(const Dims& m, const MtAxialParams& p, const float* x, float* yraw, double* stats, float* coef,
                unsigned* cnt, cudaStream_t st);
int launch_out(const Dims& m, const float* coef, const float* oraw, float* out, cudaStream_t st);
int launch_coop_fwd(const Dims& m, const MtAxialParams& p, const float* x, float* out, double* stats, float* coef,
                    unsigned* cnt, float* yraw, float* oraw, float* lse, float* tabg, cudaStream_t st);

template <int MODE>
int launch_attn2(const Dims& m, const MtAxialParams& p, const float* yraw, double* stats, float* coef, unsigned* cnt,
                 float* oraw, float* lse, cudaStream_t st);

struct LaunchPlan { int SPB, segw; size_t smem_attn; };

static int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

// strips per CTA: enough (head,row) items to fill 256 threads, dividing the group
static LaunchPlan plan_attn(const Dims& m, int extra_floats_per_strip, int extra_floats) {
  LaunchPlan pl;
  int spb = 1;
  while (spb * MT_G * m.L < 256 && spb * 2 <= m.Bg && m.Bg % (spb * 2) == 0) spb *= 2;
  pl.SPB = spb;
  const bool p2 = (m.L & (m.L - 1)) == 0;
  pl.segw = p2 ? (m.L < 32 ? m.L : 32) : 1;
  size_t fl = 4 * (size_t)m.C + (size_t)spb * 2 * m.C * m.LP + (m.pos ? 2 * (size_t)m.gp * (2 * m.L - 1) : 0) +
              (size_t)spb * extra_floats_per_strip + extra_floats;
  pl.smem_attn = fl * sizeof(float);
  return pl;
}

template <typename K>
static cudaError_t set_smem(K kern, size_t bytes) {
  // static shared (reduction scratch, tickets) counts against the 48 KB no-opt-in limit too
  if (bytes > 32 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

#define MT_DISPATCH_D(m, CALL)                                                     \
  switch ((m).d) {                                                                 \
    case 1: if ((m).pos) { CALL(1, true); } else { CALL(1, false); } break;        \
    case 2: if ((m).pos) { CALL(2, true); } else { CALL(2, false); } break;        \
    case 4: if ((m).pos) { CALL(4, true); } else { CALL(4, false); } break;        \
    case 8: if ((m).pos) { CALL(8, true); } else { CALL(8, false); } break;        \
    default: return -2;                                                            \
  }

// returns 0 ok, -1 cuda error (cudaGetLastError holds it), -2 unsupported
int axial_forward_launch(const Dims& m, const MtAxialParams& p, const float* x, float* out, char* saved,
                         cudaStream_t st, int* nlaunch, int mask) {
  double* stats = reinterpret_cast<double*>(saved + m.sv_stats);
  unsigned* cnt = reinterpret_cast<unsigned*>(saved + m.sv_cnt);
  float* coef = reinterpret_cast<float*>(saved + m.sv_coef);
  float* yraw = reinterpret_cast<float*>(saved + m.sv_yraw);
  float* oraw = reinterpret_cast<float*>(saved + m.sv_oraw);
  float* lse = reinterpret_cast<float*>(saved + m.sv_lse);
  int nl = 0;
  if (mask == 0x1f) {      // accumulators + CTA-done tickets
    if (cudaMemsetAsync(saved + m.sv_stats, 0, m.sv_zero_bytes, st) != cudaSuccess) return -1;
  }
  // default: the whole op as ONE persistent cooperative kernel (axial_coop_fwd.cu).  MT_DEBUG_FLAGS bit5 (and the
  // per-stage profiling entry point) keep the round-1 launch chain below, which also serves shapes it does not cover.
  if (mask == 0x1f && !(m.dbg & 32)) {
    const int rc = launch_coop_fwd(m, p, x, out, stats, coef, cnt, yraw, oraw, lse, reinterpret_cast<float*>(saved + m.sv_tab), st);
    if (rc < 0) return rc;
    if (rc > 0) { *nlaunch = rc; return 0; }
  }
  if (mask & 1) nl += launch_proj(m, p, x, yraw, stats, coef, cnt, st);
  const bool v1 = (m.dbg & 2) != 0;        // MT_DEBUG_FLAGS bit1: force the v1 (one row per thread) kernels
  int r2 = 0;
  if (m.training && (mask & 2) && !v1) {
    r2 = launch_attn2<0>(m, p, yraw, stats, coef, cnt, oraw, lse, st);
    if (r2 < 0) return r2;
    nl += r2;
  }
  if (m.training && (mask & 2) && r2 == 0) {
    LaunchPlan pl = plan_attn(m, 0, 48);
    if (pl.smem_attn > 227 * 1024) return -2;
#define CALL_SIM(DD, PP)                                                                       \
    if (set_smem(k_sim_stats<DD, PP>, pl.smem_attn) != cudaSuccess) return -1;                   \
    k_sim_stats<DD, PP><<<m.B / pl.SPB, 256, pl.smem_attn, st>>>(m, p, yraw, stats, coef, cnt, pl.SPB, pl.segw)
    MT_DISPATCH_D(m, CALL_SIM)
#undef CALL_SIM
    ++nl;
  }
  r2 = 0;
  if ((mask & 4) && !v1) {
    r2 = launch_attn2<1>(m, p, yraw, stats, coef, cnt, oraw, lse, st);
    if (r2 < 0) return r2;
    nl += r2;
  }
  if ((mask & 4) && r2 == 0) {
    LaunchPlan pl = plan_attn(m, 0, 24 + 2 * m.no * m.C);
    if (pl.smem_attn > 227 * 1024) return -2;
#define CALL_ATTN(DD, PP)                                                                      \
    if (set_smem(k_attn<DD, PP>, pl.smem_attn) != cudaSuccess) return -1;                        \
    k_attn<DD, PP><<<m.B / pl.SPB, 256, pl.smem_attn, st>>>(m, p, yraw, stats, coef, cnt, oraw, lse, pl.SPB, pl.segw)
    MT_DISPATCH_D(m, CALL_ATTN)
#undef CALL_ATTN
    ++nl;
  }
  if (mask & 8) nl += launch_out(m, coef, oraw, out, st);
  *nlaunch = nl;
  g_last_cuda = cudaGetLastError();
  return g_last_cuda == cudaSuccess ? 0 : -1;
}
