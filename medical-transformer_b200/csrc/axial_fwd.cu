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

// ============================================================================ k_proj
// Tile = 64 positions (TS strips x TL consecutive l) x 32 output channels.
// smem X tile is position-major in strip order (l fastest) so that y_raw stores are
// contiguous; the NCHW gather for axis 0 is transposed while staging.
__global__ void __launch_bounds__(256)
k_proj(Dims m, const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ yraw,
       double* __restrict__ stats, int TS, int TL, int tiles_per_group, int nTL) {
  extern __shared__ float Xs[];                       // [C][64]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, grp = tile / tiles_per_group, tin = tile - grp * tiles_per_group;
  const int ts_idx = tin / nTL, tl_idx = tin - ts_idx * nTL;
  const int sb0 = ts_idx * TS, l0 = tl_idx * TL, npos = TS * TL;
  const int C = m.C;
  for (int idx = tid; idx < C * 64; idx += 256) {
    int c = idx >> 6, r = idx & 63;
    if (r >= npos) continue;
    int s, l;
    if (m.axis == 0) { s = r % TS; l = r / TS; } else { l = r % TL; s = r / TL; }
    float v = 0.f;
    if (sb0 + s < m.Bg && l0 + l < m.L) v = __ldg(x + nchw_addr(m, grp * m.Bg + sb0 + s, c, l0 + l));
    Xs[c * 64 + s * TL + l] = v;
  }
  __syncthreads();
  const int pos = tid & 63, ocl = tid >> 6;
  const int s = pos / TL, l = pos - s * TL;
  const bool valid = pos < npos && sb0 + s < m.Bg && l0 + l < m.L;
  const int oc0 = blockIdx.y * 32 + ocl * 8;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  if (pos < npos) {
    for (int c = 0; c < C; c += 4) {
      float x0 = Xs[(c + 0) * 64 + pos], x1 = Xs[(c + 1) * 64 + pos];
      float x2 = Xs[(c + 2) * 64 + pos], x3 = Xs[(c + 3) * 64 + pos];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(oc0 + o) * C + c));
        acc[o] = fmaf(wv.x, x0, fmaf(wv.y, x1, fmaf(wv.z, x2, fmaf(wv.w, x3, acc[o]))));
      }
    }
  }
  const int b = grp * m.Bg + sb0 + s;
  if (valid) {
#pragma unroll
    for (int o = 0; o < 8; ++o) yraw[((size_t)b * 2 * C + oc0 + o) * m.L + l0 + l] = acc[o];
  }
  if (m.training) {
    double* st = stats + (size_t)grp * m.nstat_f;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float v = valid ? acc[o] : 0.f;
      float s1 = seg_sum(v, 32), s2 = seg_sum(v * v, 32);
      if ((tid & 31) == 0) {
        atomicAdd(st + m.f_qkv_sum + oc0 + o, (double)s1);
        atomicAdd(st + m.f_qkv_sq + oc0 + o, (double)s2);
      }
    }
  }
}

// ============================================================================ shared loaders
// qkv BN coefficients (y_hat = a*y_raw + b) for all 2C channels of statistic group `grp`
__device__ __forceinline__ void load_qkv_coef(const Dims& m, const MtAxialParams& p, const double* stats, int grp,
                                              float* ca, float* cb) {
  const double* st = stats + (size_t)grp * m.nstat_f;
  const double n = (double)m.Bg * m.L;
  for (int ch = threadIdx.x; ch < 2 * m.C; ch += blockDim.x) {
    BnCoef c = bn_coef(m.training, m.training ? st[m.f_qkv_sum + ch] : 0.0, m.training ? st[m.f_qkv_sq + ch] : 0.0, n,
                       1.f, p.bn_qkv_w[ch], p.bn_qkv_b[ch], p.bn_qkv_rm[ch], p.bn_qkv_rv[ch], m.eps);
    ca[ch] = c.a; cb[ch] = c.b;
  }
}

// normalised q,k,v strips of `spb` consecutive strips starting at b0 -> Yn[s][ch][LP]
__device__ __forceinline__ void load_strips_norm(const Dims& m, const float* __restrict__ yraw, int b0, int spb,
                                                 const float* ca, const float* cb, float* Yn) {
  const int per = 2 * m.C * m.L;
  const float* src = yraw + (size_t)b0 * per;
  for (int idx = threadIdx.x; idx < spb * per; idx += blockDim.x) {
    int s = idx / per, rem = idx - s * per;
    int ch = rem / m.L, l = rem - ch * m.L;
    Yn[((size_t)s * 2 * m.C + ch) * m.LP + l] = fmaf(ca[ch], __ldg(src + idx), cb[ch]);
  }
}

__device__ __forceinline__ void load_rel(const Dims& m, const float* __restrict__ rel, float* R) {
  const int nrel = 2 * m.gp * (2 * m.L - 1);
  for (int idx = threadIdx.x; idx < nrel; idx += blockDim.x) R[idx] = __ldg(rel + idx);
}

__device__ __forceinline__ float gate_of(const float* p) { return p ? __ldg(p) : 1.f; }

// softmax-side weight of similarity term t, head g:  a_t = gamma*f/sqrt(f^2 var + eps)
__device__ __forceinline__ BnCoef sim_coef(const Dims& m, const MtAxialParams& p, const double* stats, int grp, int t,
                                           int g) {
  const double* st = stats + (size_t)grp * m.nstat_f;
  const int ch = t * MT_G + g;
  const float f = t == 0 ? 1.f : (t == 1 ? gate_of(p.f_qr) : gate_of(p.f_kr));
  const double n = (double)m.Bg * m.L * m.L;
  return bn_coef(m.training, m.training ? st[m.f_sim_sum + ch] : 0.0, m.training ? st[m.f_sim_sq + ch] : 0.0, n, f,
                 p.bn_sim_w[ch], p.bn_sim_b[ch], p.bn_sim_rm[ch], p.bn_sim_rv[ch], m.eps);
}

// ============================================================================ k_sim_stats
template <int D, bool POS>
__global__ void __launch_bounds__(256)
k_sim_stats(Dims m, MtAxialParams p, const float* __restrict__ yraw, double* __restrict__ stats, int SPB, int segw) {
  extern __shared__ float sm[];
  const int C = m.C, L = m.L, LP = m.LP, RL = 2 * L - 1;
  float* ca = sm;
  float* cb = ca + 2 * C;
  float* red = cb + 2 * C;                 // [3][G][2]
  float* Yn = red + 48;
  float* R = Yn + (size_t)SPB * 2 * C * LP;
  const int b0 = blockIdx.x * SPB;
  const int grp = b0 / m.Bg;
  load_qkv_coef(m, p, stats, grp, ca, cb);
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
  double* st = stats + (size_t)grp * m.nstat_f;
  if (threadIdx.x < m.nt * MT_G) {
    atomicAdd(st + m.f_sim_sum + threadIdx.x, (double)red[threadIdx.x * 2 + 0]);
    atomicAdd(st + m.f_sim_sq + threadIdx.x, (double)red[threadIdx.x * 2 + 1]);
  }
  // running statistics of bn_qkv: one CTA, sequential EMA over the statistic groups
  if (blockIdx.x == 0) {
    const double n = (double)m.Bg * L;
    for (int ch = threadIdx.x; ch < 2 * C; ch += blockDim.x)
      ema_update(p.bn_qkv_rm, p.bn_qkv_rv, ch, stats, m.nstat_f, m.f_qkv_sum, m.f_qkv_sq, m.ngroups, n, 1.f, m.momentum);
  }
}

// ============================================================================ k_attn
template <int D, bool POS>
__global__ void __launch_bounds__(256)
k_attn(Dims m, MtAxialParams p, const float* __restrict__ yraw, double* __restrict__ stats, float* __restrict__ oraw,
       float* __restrict__ lse, int SPB, int segw) {
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
  load_qkv_coef(m, p, stats, grp, ca, cb);
  if (threadIdx.x < m.nt * MT_G) {
    BnCoef c = sim_coef(m, p, stats, grp, threadIdx.x / MT_G, threadIdx.x % MT_G);
    wsim[threadIdx.x] = c.a * MT_LOG2E;
  }
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
    double* st = stats + (size_t)grp * m.nstat_f;
    for (int ch = threadIdx.x; ch < NO * C; ch += blockDim.x) {
      atomicAdd(st + m.f_out_sum + ch, (double)red[ch * 2]);
      atomicAdd(st + m.f_out_sq + ch, (double)red[ch * 2 + 1]);
    }
    if (blockIdx.x == 0) {                 // running statistics of bn_similarity
      const double n = (double)m.Bg * L * L;
      for (int ch = threadIdx.x; ch < m.nt * MT_G; ch += blockDim.x) {
        const int t = ch / MT_G;
        const float f = t == 0 ? 1.f : (t == 1 ? gate_of(p.f_qr) : gate_of(p.f_kr));
        ema_update(p.bn_sim_rm, p.bn_sim_rv, ch, stats, m.nstat_f, m.f_sim_sum, m.f_sim_sq, m.ngroups, n, f, m.momentum);
      }
    }
  }
}

// ============================================================================ k_out
// One CTA per (n, c) plane: BN(bn_output) as a per-channel affine map of the raw sv/sve,
// pair sum, un-permute to NCHW, AvgPool2d(stride).
__global__ void __launch_bounds__(256)
k_out(Dims m, MtAxialParams p, const double* __restrict__ stats, const float* __restrict__ oraw,
      float* __restrict__ out) {
  __shared__ float sa[2], sb[2];
  const int n = blockIdx.x / m.C, c = blockIdx.x - n * m.C;
  const int grp = n / (m.N / m.ngroups);
  const int NO = m.no, L = m.L;
  if (threadIdx.x < NO) {
    const int ch = c * NO + threadIdx.x;
    const double* st = stats + (size_t)grp * m.nstat_f;
    const float f = threadIdx.x == 0 ? gate_of(p.f_sv) : gate_of(p.f_sve);
    BnCoef k = bn_coef(m.training, m.training ? st[m.f_out_sum + ch] : 0.0, m.training ? st[m.f_out_sq + ch] : 0.0,
                       (double)m.Bg * L, f, p.bn_out_w[ch], p.bn_out_b[ch], p.bn_out_rm[ch], p.bn_out_rv[ch], m.eps);
    sa[threadIdx.x] = k.a; sb[threadIdx.x] = k.b;
  }
  __syncthreads();
  const float a0 = sa[0], a1 = NO == 2 ? sa[1] : 0.f, bsum = sb[0] + (NO == 2 ? sb[1] : 0.f);
  const int st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_;
  const float inv = 1.f / (float)(st_ * st_);
  for (int idx = threadIdx.x; idx < Ho * Wo; idx += blockDim.x) {
    const int ho = idx / Wo, wo = idx - ho * Wo;
    float acc = 0.f;
    for (int dh = 0; dh < st_; ++dh)
      for (int dw = 0; dw < st_; ++dw) {
        const int h = ho * st_ + dh, w = wo * st_ + dw;
        const int b = m.axis == 0 ? n * m.W + w : n * m.H + h;
        const int l = m.axis == 0 ? h : w;
        const float* o = oraw + ((size_t)b * NO * m.C + c * NO) * L + l;
        acc = fmaf(a0, __ldg(o), acc);
        if (NO == 2) acc = fmaf(a1, __ldg(o + L), acc);
      }
    out[((size_t)blockIdx.x * Ho + ho) * Wo + wo] = fmaf(acc, inv, bsum);
  }
  if (m.training && blockIdx.x == 0) {     // running statistics of bn_output
    const double cnt = (double)m.Bg * L;
    for (int ch = threadIdx.x; ch < NO * m.C; ch += blockDim.x) {
      const float f = (NO == 2 && (ch & 1)) ? gate_of(p.f_sve) : gate_of(p.f_sv);
      ema_update(p.bn_out_rm, p.bn_out_rv, ch, stats, m.nstat_f, m.f_out_sum, m.f_out_sq, m.ngroups, cnt, f, m.momentum);
    }
  }
}

// ============================================================================ host launchers
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
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
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
  float* yraw = reinterpret_cast<float*>(saved + m.sv_yraw);
  float* oraw = reinterpret_cast<float*>(saved + m.sv_oraw);
  float* lse = reinterpret_cast<float*>(saved + m.sv_lse);
  int nl = 0;
  if (m.training && mask == 0x1f) {
    if (cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)m.ngroups * m.nstat_f, st) != cudaSuccess) return -1;
  }
  if (mask & 1) {  // k_proj
    int TS, TL;
    if (m.axis == 1) { TL = m.L < 64 ? m.L : 64; TS = 64 / TL; if (TS > m.Bg) TS = m.Bg; }
    else { TS = m.Bg < 8 ? m.Bg : 8; TL = 64 / TS; if (TL > m.L) TL = m.L; }
    if (TS < 1) TS = 1;
    const int nTL = (m.L + TL - 1) / TL, nTS = (m.Bg + TS - 1) / TS;
    dim3 grid(m.ngroups * nTS * nTL, 2 * m.C / 32);
    const size_t smem = (size_t)m.C * 64 * sizeof(float);
    if (set_smem(k_proj, smem) != cudaSuccess) return -1;
    k_proj<<<grid, 256, smem, st>>>(m, x, p.w_qkv, yraw, stats, TS, TL, nTS * nTL, nTL);
    ++nl;
  }
  if (m.training && (mask & 2)) {
    LaunchPlan pl = plan_attn(m, 0, 48);
    if (pl.smem_attn > 227 * 1024) return -2;
#define CALL_SIM(DD, PP)                                                                       \
    if (set_smem(k_sim_stats<DD, PP>, pl.smem_attn) != cudaSuccess) return -1;                   \
    k_sim_stats<DD, PP><<<m.B / pl.SPB, 256, pl.smem_attn, st>>>(m, p, yraw, stats, pl.SPB, pl.segw)
    MT_DISPATCH_D(m, CALL_SIM)
#undef CALL_SIM
    ++nl;
  }
  if (mask & 4) {
    LaunchPlan pl = plan_attn(m, 0, 24 + 2 * m.no * m.C);
    if (pl.smem_attn > 227 * 1024) return -2;
#define CALL_ATTN(DD, PP)                                                                      \
    if (set_smem(k_attn<DD, PP>, pl.smem_attn) != cudaSuccess) return -1;                        \
    k_attn<DD, PP><<<m.B / pl.SPB, 256, pl.smem_attn, st>>>(m, p, yraw, stats, oraw, lse, pl.SPB, pl.segw)
    MT_DISPATCH_D(m, CALL_ATTN)
#undef CALL_ATTN
    ++nl;
  }
  if (mask & 8) {
    k_out<<<m.N * m.C, 256, 0, st>>>(m, p, stats, oraw, out);
    ++nl;
  }
  *nlaunch = nl;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
