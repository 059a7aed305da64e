// axial_coop_fwd.cu -- one AxialAttention* forward (reference axialnet.py:142-189 / :52-92 / :222-253) as ONE persistent
// cooperative kernel.  Phases (a grid barrier between them in training mode; none in eval mode):
//
//   P1  async-gather the unit's x strips (NCHW) and the CTA's slice of W_qkv, Y = W X (register-tiled fp32), keep Y in shared
//       memory, write y_raw for the backward, sum / sumsq per projection channel          -> bn_qkv statistics
//   P2  normalise Y in place; similarity statistics in CLOSED FORM (no L^2 pass):
//         sum_ij qk   = sum_c (sum_i q_ci)(sum_j k_cj)           sum_ij qk^2 = <Q Q^T, K K^T>_F   (d x d Gram matrices)
//         sum_ij qr   = sum_i q_i . PR_i                          sum_ij qr^2 = sum_i q_i^T M_i q_i
//       with PR_i / M_i the length-L window sums of R and R R^T (Toeplitz: relative[c, i-j+L-1], axialnet.py:155); the
//       window sums are built ONCE per op, spread over the whole grid before the first barrier; the transposed kr term has
//       the same form in k                                                                 -> bn_similarity statistics
//   P3  softmax_j(sum_t w_t S_t) -> [P v | P Rv] with the scores in registers (sliding Toeplitz windows); a row block's keys
//       are split over KS adjacent lanes and merged with shuffles; o_raw, log-sum-exp, sum / sumsq  -> bn_output statistics
//   P4  BN(bn_output) affine + pair sum + NCHW un-permute (+ AvgPool) straight from the resident o_raw
//
// Every CTA derives the BN coefficients it needs from the fp64 accumulators itself.  The CTA that owns the first unit of a
// statistic group also writes them to the coefficient table the backward reads, and the owner of group 0 applies the
// running-statistics EMA -- no serial finaliser at the end of the kernel.
// Index arithmetic is incremental / shift based (DivS): ncu of the first version showed 16 % of all issued instructions
// in runtime integer divisions of the staging loops.
#include "axial_coop.cuh"
extern thread_local cudaError_t g_last_cuda;

namespace {

__device__ __forceinline__ float warp_sum(float v) { return seg_sum(v, 32); }

// running-statistics EMA of channel `ch`.  One statistic group (every op outside the batched LoGo branch): straight from the
// moments the caller just derived (c.mean / c.var of the RAW value; f = gate in front of the BN).  Several groups: the
// sequential chain of `ngroups` BatchNorm calls, with the groups' moments loaded eight at a time (one L2 round trip per
// batch instead of one per group).
__device__ __forceinline__ void ema_apply(const Dims& m, float* rm, float* rv, int ch, const BnCoef& c, const double* stats,
                                       int off_sum, int off_sq, double inv_n, float unbias, float f) {
  const float mom = m.momentum;
  float m_ = rm[ch], v_ = rv[ch];
  if (m.ngroups == 1) {
    m_ = (1.f - mom) * m_ + mom * (f * c.mean);
    v_ = (1.f - mom) * v_ + mom * (f * f * c.var * unbias);
  } else {
    for (int g0 = 0; g0 < m.ngroups; g0 += 8) {
      double s[8], q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int g = g0 + k < m.ngroups ? g0 + k : m.ngroups - 1;
        s[k] = stat_get(m, stats, g, m.nstat_f, off_sum + ch); q[k] = stat_get(m, stats, g, m.nstat_f, off_sq + ch);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (g0 + k < m.ngroups) {
          Moments mo = moments(s[k], q[k], inv_n);
          m_ = (1.f - mom) * m_ + mom * (f * mo.mu);
          v_ = (1.f - mom) * v_ + mom * (f * f * mo.var * unbias);
        }
    }
  }
  rm[ch] = m_; rv[ch] = v_;
}

// ---- projection micro-GEMM of one unit: Y[RU x XP] = Wt^T[RU x C] . Xs[C x XP], thread tile TM x TN --------------
template <int TM, int TN>
__device__ __forceinline__ void unit_gemm(const float* __restrict__ Wt, int WTP, const float* __restrict__ Xs, int XP, int C,
                                          int RU, int L, float* __restrict__ Ys, float* __restrict__ yraw_u, int ystride) {
  const int ntc = XP / TN, ntiles = (RU / TM) * ntc;
  for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
    const int rt_ = tile / ntc, ct = tile - rt_ * ntc, r0 = rt_ * TM, c0 = ct * TN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < C; ++k) {
      float a[TM], b[TN];
      ldv<TM>(a, Wt + k * WTP + r0);
      ldv<TN>(b, Xs + k * XP + c0);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // TN divides L whenever TN > 1 (host plan), so a tile never straddles two strips
    const int s = c0 / L, l = c0 - s * L;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      typedef typename VecT<TN>::T V;
      *reinterpret_cast<V*>(Ys + ((size_t)s * RU + r0 + i) * L + l) = *reinterpret_cast<V*>(acc[i]);
      *reinterpret_cast<V*>(yraw_u + (size_t)s * ystride + (size_t)(r0 + i) * L + l) = *reinterpret_cast<V*>(acc[i]);
    }
  }
}

// per-row sum / sumsq of a [nstrips][nrows][L] shared-memory block, ADDED to red[2*row], red[2*row+1].  TPR (a power of two
// <= 32) adjacent lanes share a row, each takes float4 chunks, then a TPR-wide shuffle tree; a row belongs to exactly one
// lane group, so there are no atomics and the summation order is fixed.
__device__ __forceinline__ void rows_stats(const float* __restrict__ blk, int nrows, int L, int nstrips, int sstride,
                                           float* __restrict__ red, int tpr) {
  const int sub = threadIdx.x & (tpr - 1), rpp = blockDim.x / tpr;          // rows per pass
  const bool v4 = (L & 3) == 0;
  for (int r0 = 0; r0 < nrows; r0 += rpp) {
    const int r = r0 + (int)threadIdx.x / tpr;
    float s1 = 0.f, s2 = 0.f;
    if (r < nrows) {
      for (int s = 0; s < nstrips; ++s) {
        const float* row = blk + (size_t)s * sstride + (size_t)r * L;
        if (v4) {
          for (int l = 4 * sub; l < L; l += 4 * tpr) {
            const float4 v = *reinterpret_cast<const float4*>(row + l);
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
          }
        } else {
          for (int l = sub; l < L; l += tpr) { const float v = row[l]; s1 += v; s2 = fmaf(v, v, s2); }
        }
      }
    }
    s1 = seg_sum(s1, tpr); s2 = seg_sum(s2, tpr);
    if (sub == 0 && r < nrows) { red[2 * r] += s1; red[2 * r + 1] += s2; }
  }
}

// y = a[row] * y + b[row] in place on a [nstrips][RU][L] block
__device__ __forceinline__ void normalise_rows(float* __restrict__ Ys, int total, int L, const DivS dL4, const DivS dRU,
                                               const float* __restrict__ ca, const float* __restrict__ cb) {
  if ((L & 3) == 0) {
    float4* y4 = reinterpret_cast<float4*>(Ys);
    for (int i4 = threadIdx.x; i4 < (total >> 2); i4 += blockDim.x) {
      const int r = md(dv(i4, dL4), dRU);
      const float ka = ca[r], kb = cb[r];
      float4 v = y4[i4];
      v.x = fmaf(ka, v.x, kb); v.y = fmaf(ka, v.y, kb); v.z = fmaf(ka, v.z, kb); v.w = fmaf(ka, v.w, kb);
      y4[i4] = v;
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int r = (i / L) % dRU.d;
      Ys[i] = fmaf(ca[r], Ys[i], cb[r]);
    }
  }
}

}  // namespace

template <int D, bool POS, int RT>
__global__ void __launch_bounds__(MT_COOP_THREADS)
k_axial_fwd(Dims m, MtAxialParams p, CoopPlan pl, const float* __restrict__ x, float* __restrict__ out,
            double* __restrict__ stats, float* __restrict__ coef, unsigned* __restrict__ cnt, float* __restrict__ yraw,
            float* __restrict__ oraw, float* __restrict__ lse, float* __restrict__ tabg) {
  constexpr int GP = 2 * D, NO = POS ? 2 : 1, NT = POS ? 3 : 1, DD = D * (D + 1) / 2;
  extern __shared__ __align__(16) float sm[];
  const int C = m.C, L = m.L, RP = rel_pitch(L);
  const int G = pl.G, SPU = pl.SPU, HS = pl.HS, HPU = pl.HPU, RU = pl.RU, NOCU = pl.NOCU, CU = pl.CU, KS = pl.KS,
            LB = pl.LB, XP = pl.XP;
  float* ca = sm + pl.o_ca;        // [RU]    bn_qkv affine of the current statistic group
  float* cb = sm + pl.o_cb;
  float* wsim = sm + pl.o_wsim;    // [3][8]  a_t * log2(e) per (term, head)
  float* oa = sm + pl.o_oa;        // [NOCU]  bn_output affine
  float* ob = sm + pl.o_ob;
  float* redq = sm + pl.o_redq;    // [RU][2]   CTA partial sums of the current group
  float* reds = sm + pl.o_reds;    // [3][8][2]
  float* redo = sm + pl.o_redo;    // [NOCU][2]
  float* tab = sm + pl.o_tab;      // PRq [D][L] | PRk [D][L] | Mq [DD][L] | Mk [DD][L]
  float* Rp = sm + pl.o_rp;        // plain / shifted Toeplitz tables (axial_attn2.cuh)
  float* Rs = sm + pl.o_rs;
  float* Wt = sm + pl.o_wt;        // [C][WTP] transposed slice of W_qkv: rows of this CTA's head group
  float* Xs = sm + pl.o_x;         // [C][XP]
  float* slots = sm + pl.o_slot;   // nslot x { Ys [SPU][RU][L] | Os [SPU][NOCU][L] }
  const int tid = threadIdx.x, cta = blockIdx.x;
  const int hg = cta % HS, row0 = hg * RU, och0 = hg * NOCU;      // G % HS == 0: the head group is fixed per CTA
  unsigned epoch = 0;
  unsigned* bar = cnt + 8;
  const bool train = m.training != 0;
  const size_t plane = (size_t)m.H * m.W;
  phase_stamp(m, 0);

  // async gather of one unit's strips: x (NCHW) -> Xs[c][s*L + l]
  auto issue_x = [&](int b0) {
    const int n = b0 / m.other, o0 = b0 - n * m.other;
    if (m.axis == 1) {
      const float* xb = x + ((size_t)n * C * m.H + o0) * m.W;                 // + c*plane + s*W + l
      if ((L & 3) == 0) {
        const int total4 = (C * XP) >> 2;
        for (int i4 = tid; i4 < total4; i4 += blockDim.x) {
          const int row = dv(i4, pl.dL4), l4 = i4 - row * pl.dL4.d, c = dv(row, pl.dSPU), s = row - c * SPU;
          cp_async16(Xs + c * XP + s * L + 4 * l4, xb + (size_t)c * plane + (size_t)s * m.W + 4 * l4);
        }
      } else {
        for (int i = tid; i < C * XP; i += blockDim.x) {
          const int row = i / L, l = i - row * L, c = row / SPU, s = row - c * SPU;
          cp_async4(Xs + c * XP + s * L + l, xb + (size_t)c * plane + (size_t)s * m.W + l);
        }
      }
    } else {
      const float* xb = x + (size_t)n * C * plane + o0;                        // + c*plane + l*W + s
      for (int i = tid; i < C * XP; i += blockDim.x) {                         // s fastest: neighbouring strips are neighbouring addresses
        const int rem = dv(i, pl.dSPU), s = i - rem * SPU, c = dv(rem, pl.dL), l = rem - c * L;
        cp_async4(Xs + c * XP + s * L + l, xb + (size_t)c * plane + (size_t)l * m.W + s);
      }
    }
  };

  // ---------------------------------------------------------------- one-time staging
  { const int u0 = pl.contig ? cta * pl.upc : cta;                // first unit's gather is in flight during everything below
    if (u0 < pl.nunits) issue_x((u0 / HS) * SPU); }
  for (int idx = tid; idx < RU * C; idx += blockDim.x) {
    const int r = idx / C, k = idx - r * C;
    cp_async4(Wt + k * pl.WTP + r, p.w_qkv + (size_t)(row0 + r) * C + k);
  }
  if (POS) load_rel2(m, p.relative, Rp, Rs);
  __syncthreads();
  if (POS && train) {
    // window sums over mm in [i, i+L-1] of R[c][mm] (PR) and w * R[c][mm] R[c2][mm] (M; w = 2 off the diagonal): every
    // entry is computed by exactly one thread of the grid and published through `tabg` (read back after barrier 1)
    const int per = D + DD, nent = 2 * per * L;
    for (int e = cta * blockDim.x + tid; e < nent; e += G * blockDim.x) {
      const int i = e % L, k = (e / L) % per, which = e / (L * per);      // which: 0 = q rows, 1 = k rows of `relative`
      const float* R0 = Rp + (which * D) * RP;
      float acc = 0.f;
      int dst;
      if (k < D) {
        const float* r = R0 + k * RP + i;
        for (int mm = 0; mm < L; ++mm) acc += r[mm];
        dst = (which * D + k) * L + i;
      } else {
        int pc = k - D, c1 = 0;
        while (pc >= D - c1) { pc -= D - c1; ++c1; }
        const int c2 = c1 + pc;
        const float* r1 = R0 + c1 * RP + i;
        const float* r2 = R0 + c2 * RP + i;
        for (int mm = 0; mm < L; ++mm) acc = fmaf(r1[mm], r2[mm], acc);
        if (c1 != c2) acc *= 2.f;
        dst = (2 * D + which * DD + (k - D)) * L + i;
      }
      tabg[dst] = acc;
    }
  }
  phase_stamp(m, 1);

  // =================================================================================== P1: projection + bn_qkv sums
  {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Ys = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      if (k > 0) { __syncthreads(); issue_x(b0); }        // Xs (and a streamed slot) are reused: previous readers are done
      if (grp != cur) {
        __syncthreads();
        if (cur >= 0 && train) {
          double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
          for (int r = tid; r < RU; r += blockDim.x) {
            MT_GADD(m, st + m.f_qkv_sum + row0 + r, (double)redq[2 * r]);
            MT_GADD(m, st + m.f_qkv_sq + row0 + r, (double)redq[2 * r + 1]);
          }
        }
        __syncthreads();
        for (int r = tid; r < 2 * RU; r += blockDim.x) redq[r] = 0.f;
        cur = grp;
      }
      cp_async_wait_all();
      __syncthreads();
      float* yu = yraw + ((size_t)b0 * 2 * C + row0) * L;
      const int ystride = 2 * C * L;
      if (pl.tm == 4 && pl.tn == 4) unit_gemm<4, 4>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else if (pl.tm == 2 && pl.tn == 4) unit_gemm<2, 4>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else if (pl.tm == 2 && pl.tn == 2) unit_gemm<2, 2>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else unit_gemm<2, 1>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      __syncthreads();
      if (train) rows_stats(Ys, RU, L, SPU, RU * L, redq, pl.tpr_q);
    }
    if (cur >= 0 && train) {
      __syncthreads();
      double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
      for (int r = tid; r < RU; r += blockDim.x) {
        MT_GADD(m, st + m.f_qkv_sum + row0 + r, (double)redq[2 * r]);
        MT_GADD(m, st + m.f_qkv_sq + row0 + r, (double)redq[2 * r + 1]);
      }
    }
  }
  phase_stamp(m, 2);
  if (train) grid_barrier(bar, G, epoch);
  phase_stamp(m, 3);

  // bn_qkv coefficients of statistic group `grp` for this CTA's rows -> ca / cb; the owner of the group's first unit also
  // publishes them (coefficient table of the backward) and the owner of group 0 applies the running-statistics EMA
  auto qkv_coefs = [&](int grp, bool owner) {
    for (int r = tid; r < RU; r += blockDim.x) {
      const int ch = row0 + r;
      BnCoef c = coef_from_stats(m, stats, grp, m.f_qkv_sum, m.f_qkv_sq, ch, m.inv_n1, 1.f, p.bn_qkv_w, p.bn_qkv_b, p.bn_qkv_rm, p.bn_qkv_rv);
      ca[r] = c.a; cb[r] = c.b;
      if (owner) {
        float* o = coef + (size_t)grp * m.ncoef_f + m.cf_q + 2 * ch;
        o[0] = c.a; o[1] = c.b;
        if (train && grp == 0) ema_apply(m, p.bn_qkv_rm, p.bn_qkv_rv, ch, c, stats, m.f_qkv_sum, m.f_qkv_sq, m.inv_n1, m.unb1, 1.f);
      }
    }
  };

  // =================================================================================== P2: normalise + similarity sums
  // (eval mode: the normalisation happens at the top of P3 and there is nothing to accumulate)
  if (train) {
    if (POS) for (int i = tid; i < pl.tab_fl; i += blockDim.x) tab[i] = __ldcg(tabg + i);
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Ys = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      if (!pl.resident) __syncthreads();      // the single slot is reused: the previous unit's readers must be done
      if (grp != cur) {
        __syncthreads();
        if (cur >= 0) {
          double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
          if (tid < NT * HPU) {
            const int t = tid / HPU, g = hg * HPU + tid % HPU;
            MT_GADD(m, st + m.f_sim_sum + t * MT_G + g, (double)reds[(t * MT_G + g) * 2]);
            MT_GADD(m, st + m.f_sim_sq + t * MT_G + g, (double)reds[(t * MT_G + g) * 2 + 1]);
          }
          __syncthreads();
        }
        qkv_coefs(grp, b0 == grp * m.Bg);
        if (tid < 48) reds[tid] = 0.f;
        cur = grp;
        __syncthreads();
      }
      if (!pl.resident) {          // streamed: this CTA's own y_raw back from L2
        for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
          const int l = i % L, r = (i / L) % RU, s = i / (L * RU);
          Ys[i] = __ldcg(yraw + ((size_t)(b0 + s) * 2 * C + row0 + r) * L + l);
        }
        __syncthreads();
      }
      normalise_rows(Ys, SPU * RU * L, L, pl.dL4, pl.dRU, ca, cb);
      __syncthreads();
      // closed-form statistics on the normalised q, k: one warp per (strip, head)
      {
        const int w = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
        for (int pr = w; pr < SPU * HPU; pr += nw) {
          const int s = dv(pr, pl.dHPU), hl = pr - s * HPU, g = hg * HPU + hl;
          const float* Yb = Ys + ((size_t)s * RU + hl * 4 * D) * L;
          float sq[D], sk[D], gq[DD], gk[DD], aqr = 0.f, aqr2 = 0.f, akr = 0.f, akr2 = 0.f;
#pragma unroll
          for (int c = 0; c < D; ++c) { sq[c] = 0.f; sk[c] = 0.f; }
#pragma unroll
          for (int c = 0; c < DD; ++c) { gq[c] = 0.f; gk[c] = 0.f; }
          for (int i = lane; i < L; i += 32) {
            float q[D], kk[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { q[c] = Yb[c * L + i]; kk[c] = Yb[(D + c) * L + i]; }
            int pc = 0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
              sq[c] += q[c]; sk[c] += kk[c];
              if (POS) { aqr = fmaf(q[c], tab[c * L + i], aqr); akr = fmaf(kk[c], tab[(D + c) * L + i], akr); }
#pragma unroll
              for (int c2 = c; c2 < D; ++c2) {
                const float qq = q[c] * q[c2], k2 = kk[c] * kk[c2];
                gq[pc] += qq; gk[pc] += k2;
                if (POS) {
                  aqr2 = fmaf(qq, tab[(2 * D + pc) * L + i], aqr2);
                  akr2 = fmaf(k2, tab[(2 * D + DD + pc) * L + i], akr2);
                }
                ++pc;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < D; ++c) { sq[c] = warp_sum(sq[c]); sk[c] = warp_sum(sk[c]); }
#pragma unroll
          for (int c = 0; c < DD; ++c) { gq[c] = warp_sum(gq[c]); gk[c] = warp_sum(gk[c]); }
          if (POS) { aqr = warp_sum(aqr); aqr2 = warp_sum(aqr2); akr = warp_sum(akr); akr2 = warp_sum(akr2); }
          if (lane == 0) {
            double a0 = 0.0, a1 = 0.0;
            int pc = 0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
              a0 += (double)sq[c] * (double)sk[c];
#pragma unroll
              for (int c2 = c; c2 < D; ++c2) { a1 += (c2 == c ? 1.0 : 2.0) * (double)gq[pc] * (double)gk[pc]; ++pc; }
            }
            atomicAdd(reds + (0 * MT_G + g) * 2, (float)a0); atomicAdd(reds + (0 * MT_G + g) * 2 + 1, (float)a1);
            if (POS) {
              atomicAdd(reds + (1 * MT_G + g) * 2, aqr); atomicAdd(reds + (1 * MT_G + g) * 2 + 1, aqr2);
              atomicAdd(reds + (2 * MT_G + g) * 2, akr); atomicAdd(reds + (2 * MT_G + g) * 2 + 1, akr2);
            }
          }
        }
      }
    }
    if (cur >= 0) {
      __syncthreads();
      double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
      if (tid < NT * HPU) {
        const int t = tid / HPU, g = hg * HPU + tid % HPU;
        MT_GADD(m, st + m.f_sim_sum + t * MT_G + g, (double)reds[(t * MT_G + g) * 2]);
        MT_GADD(m, st + m.f_sim_sq + t * MT_G + g, (double)reds[(t * MT_G + g) * 2 + 1]);
      }
    }
    phase_stamp(m, 4);
    grid_barrier(bar, G, epoch);
  }
  phase_stamp(m, 5);

  // =================================================================================== P3: softmax, P v, P Rv
  {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Ys = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      if (!pl.resident) __syncthreads();      // the single slot is reused: the previous unit's readers must be done
      float* Os = Ys + pl.ys_fl;
      if (grp != cur) {
        __syncthreads();
        if (cur >= 0 && train) {
          double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
          for (int r = tid; r < NOCU; r += blockDim.x) {
            MT_GADD(m, st + m.f_out_sum + och0 + r, (double)redo[2 * r]);
            MT_GADD(m, st + m.f_out_sq + och0 + r, (double)redo[2 * r + 1]);
          }
          __syncthreads();
        }
        const bool owner = b0 == grp * m.Bg;
        if (tid < NT * HPU) {
          const int t = tid / HPU, g = hg * HPU + tid % HPU, ch = t * MT_G + g;
          const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
          BnCoef c = coef_from_stats(m, stats, grp, m.f_sim_sum, m.f_sim_sq, ch, m.inv_n2, f, p.bn_sim_w, p.bn_sim_b, p.bn_sim_rm, p.bn_sim_rv);
          wsim[ch] = c.a * MT_LOG2E;
          if (owner) {
            float* o = coef + (size_t)grp * m.ncoef_f + m.cf_s + 2 * ch;
            o[0] = c.a * MT_LOG2E; o[1] = c.a;
            if (train && grp == 0) ema_apply(m, p.bn_sim_rm, p.bn_sim_rv, ch, c, stats, m.f_sim_sum, m.f_sim_sq, m.inv_n2, m.unb2, f);
          }
        }
        if (!train || !pl.resident) qkv_coefs(grp, owner && !train);      // training: P2 published them already
        for (int r = tid; r < 2 * NOCU; r += blockDim.x) redo[r] = 0.f;
        cur = grp;
        __syncthreads();
      }
      if (!pl.resident) {            // streamed slot: this unit's raw projection back from L2 (the CTA's own y_raw)
        for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
          const int l = i % L, r = (i / L) % RU, s = i / (L * RU);
          Ys[i] = __ldcg(yraw + ((size_t)(b0 + s) * 2 * C + row0 + r) * L + l);
        }
        __syncthreads();
      }
      if (!train || !pl.resident) {  // training + resident: P2 already normalised the slot in place
        normalise_rows(Ys, SPU * RU * L, L, pl.dL4, pl.dRU, ca, cb);
        __syncthreads();
      }

      const int nitems = (m.dbg & 4) ? 0 : SPU * HPU * LB * KS;
      const int kpb = LB / KS;                                   // key blocks per lane
      for (int base = 0; base < nitems; base += blockDim.x) {
        const int item = base + tid;
        const bool act = item < nitems;
        const int it = act ? item : 0;
        const int i1 = dv(it, pl.dKS), ks = it - i1 * KS, i2 = dv(i1, pl.dLB), t = i1 - i2 * LB, s = dv(i2, pl.dHPU), hl = i2 - s * HPU;
        const int g = hg * HPU + hl;
        const float* Yb = Ys + ((size_t)s * RU + hl * 4 * D) * L;
        const int wb0 = ks * kpb, wb1 = wb0 + kpb;
        float q[D][RT];
#pragma unroll
        for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * L + RT * t);
        WinA<RT> wq[D]; WinB<RT> wk[D]; WinA<RT> wv[GP];
        const int KA0 = t + LB - wb0, KB0 = LB - t + wb0;
        if (POS) {
#pragma unroll
          for (int c = 0; c < D; ++c) { wq[c].init(Rp + c * RP, KA0); wk[c].init(Rs + (D + c) * RP, KB0); }
#pragma unroll
          for (int c = 0; c < GP; ++c) wv[c].init(Rp + (GP + c) * RP, KA0);
        }
        const float wqk = wsim[g], wqr = POS ? wsim[MT_G + g] : 1.f, wkr = POS ? wsim[2 * MT_G + g] : 1.f;
        float mx[RT], lsum[RT], asv[RT][GP], asve[RT][GP];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          mx[r] = -INFINITY; lsum[r] = 0.f;
#pragma unroll
          for (int c = 0; c < GP; ++c) { asv[r][c] = 0.f; asve[r][c] = 0.f; }
        }
        for (int wb = wb0; wb < wb1; ++wb) {
          float kk[D][RT], vv[GP][RT];
#pragma unroll
          for (int c = 0; c < D; ++c) ldv<RT>(kk[c], Yb + (D + c) * L + RT * wb);
#pragma unroll
          for (int c = 0; c < GP; ++c) ldv<RT>(vv[c], Yb + (GP + c) * L + RT * wb);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            float z[RT];
#pragma unroll
            for (int w = 0; w < RT; ++w) {
              float sqk = 0.f, sqr = 0.f, skr = 0.f;
#pragma unroll
              for (int c = 0; c < D; ++c) {
                sqk = fmaf(q[c][r], kk[c][w], sqk);
                if (POS) { sqr = fmaf(q[c][r], wq[c].get(r, w), sqr); skr = fmaf(kk[c][w], wk[c].get(r, w), skr); }
              }
              z[w] = wqk * sqk;
              if (POS) z[w] = fmaf(wqr, sqr, fmaf(wkr, skr, z[w]));
            }
            float bm = z[0];
#pragma unroll
            for (int w = 1; w < RT; ++w) bm = fmaxf(bm, z[w]);
            if (bm > mx[r]) {                     // lazy rescale, once per RT keys
              const float sc = ex2(mx[r] - bm);
              mx[r] = bm; lsum[r] *= sc;
#pragma unroll
              for (int c = 0; c < GP; ++c) { asv[r][c] *= sc; if (POS) asve[r][c] *= sc; }
            }
#pragma unroll
            for (int w = 0; w < RT; ++w) {
              const float pj = ex2(z[w] - mx[r]);
              lsum[r] += pj;
#pragma unroll
              for (int c = 0; c < GP; ++c) {
                asv[r][c] = fmaf(pj, vv[c][w], asv[r][c]);
                if (POS) asve[r][c] = fmaf(pj, wv[c].get(r, w), asve[r][c]);
              }
            }
          }
          if (POS && wb + 1 < wb1) {
#pragma unroll
            for (int c = 0; c < D; ++c) { wq[c].advance(Rp + c * RP, KA0 - (wb - wb0) - 1); wk[c].advance(Rs + (D + c) * RP, KB0 + (wb - wb0) + 1); }
#pragma unroll
            for (int c = 0; c < GP; ++c) wv[c].advance(Rp + (GP + c) * RP, KA0 - (wb - wb0) - 1);
          }
        }
        // merge the KS key ranges of a row block (adjacent lanes, ks fastest): every lane ends with the totals
        for (int msk = 1; msk < KS; msk <<= 1) {
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const float mo = __shfl_xor_sync(0xffffffffu, mx[r], msk), lo = __shfl_xor_sync(0xffffffffu, lsum[r], msk);
            const float nm = fmaxf(mx[r], mo), s1 = ex2(mx[r] - nm), s2 = ex2(mo - nm);
            lsum[r] = lsum[r] * s1 + lo * s2;
#pragma unroll
            for (int c = 0; c < GP; ++c) {
              const float ao = __shfl_xor_sync(0xffffffffu, asv[r][c], msk);
              asv[r][c] = asv[r][c] * s1 + ao * s2;
              if (POS) {
                const float eo = __shfl_xor_sync(0xffffffffu, asve[r][c], msk);
                asve[r][c] = asve[r][c] * s1 + eo * s2;
              }
            }
            mx[r] = nm;
          }
        }
        if (act) {
          typedef typename VecT<RT>::T V;
          const int b = b0 + s;
          float inv[RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) inv[r] = 1.f / lsum[r];
          if (ks == 0) {
            float lv[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) lv[r] = mx[r] + lg2(lsum[r]);
            *reinterpret_cast<V*>(lse + ((size_t)b * MT_G + g) * L + RT * t) = *reinterpret_cast<V*>(lv);
          }
#pragma unroll
          for (int c = 0; c < GP; ++c) {
            if ((c & (KS - 1)) != ks) continue;        // the KS lanes share the stores
            const int chl = (hl * GP + c) * NO, ch = (g * GP + c) * NO;
            float o0[RT], o1[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) { o0[r] = asv[r][c] * inv[r]; o1[r] = POS ? asve[r][c] * inv[r] : 0.f; }
            *reinterpret_cast<V*>(Os + ((size_t)s * NOCU + chl) * L + RT * t) = *reinterpret_cast<V*>(o0);
            *reinterpret_cast<V*>(oraw + ((size_t)b * NO * C + ch) * L + RT * t) = *reinterpret_cast<V*>(o0);
            if (POS) {
              *reinterpret_cast<V*>(Os + ((size_t)s * NOCU + chl + 1) * L + RT * t) = *reinterpret_cast<V*>(o1);
              *reinterpret_cast<V*>(oraw + ((size_t)b * NO * C + ch + 1) * L + RT * t) = *reinterpret_cast<V*>(o1);
            }
          }
        }
      }
      __syncthreads();
      if (train) rows_stats(Os, NOCU, L, SPU, NOCU * L, redo, pl.tpr_o);
    }
    if (cur >= 0 && train) {
      __syncthreads();
      double* st = stat_slot(m, stats, cur, m.nstat_f, cta);
      for (int r = tid; r < NOCU; r += blockDim.x) {
        MT_GADD(m, st + m.f_out_sum + och0 + r, (double)redo[2 * r]);
        MT_GADD(m, st + m.f_out_sq + och0 + r, (double)redo[2 * r + 1]);
      }
    }
  }
  phase_stamp(m, 6);
  if (train) grid_barrier(bar, G, epoch);
  pdl_launch_dependents();          // last phase: a PDL-launched successor (pdl.cuh) may take its SM slots now and wait there
  phase_stamp(m, 7);

  // =================================================================================== P4: bn_output + un-permute (+ pool)
  {
    const int st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_, SO = SPU / st_, LO = L / st_, PP = SO * LO;
    const size_t oplane = (size_t)Ho * Wo;
    const float invp = 1.f / (float)(st_ * st_);
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Os = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl + pl.ys_fl;
      if (!pl.resident) __syncthreads();
      if (grp != cur) {
        __syncthreads();
        const bool owner = b0 == grp * m.Bg;
        for (int r = tid; r < NOCU; r += blockDim.x) {
          const int ch = och0 + r;
          const float f = (NO == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
          BnCoef c = coef_from_stats(m, stats, grp, m.f_out_sum, m.f_out_sq, ch, m.inv_n1, f, p.bn_out_w, p.bn_out_b, p.bn_out_rm, p.bn_out_rv);
          oa[r] = c.a; ob[r] = c.b;
          if (owner) {
            float* o = coef + (size_t)grp * m.ncoef_f + m.cf_o + 2 * ch;
            o[0] = c.a; o[1] = c.b;
            if (train && grp == 0) ema_apply(m, p.bn_out_rm, p.bn_out_rv, ch, c, stats, m.f_out_sum, m.f_out_sq, m.inv_n1, m.unb1, f);
          }
        }
        cur = grp;
        __syncthreads();
      }
      if (!pl.resident && pl.upc > 1) {
        __syncthreads();
        for (int i = tid; i < SPU * NOCU * L; i += blockDim.x) {
          const int l = i % L, r = (i / L) % NOCU, s = i / (L * NOCU);
          Os[i] = __ldcg(oraw + ((size_t)(b0 + s) * NO * C + och0 + r) * L + l);
        }
        __syncthreads();
      }
      const int n = b0 / m.other, o0 = b0 - n * m.other;       // SPU divides `other`: one image per unit
      // a thread keeps one output position (so, lo) and walks the unit's channels: no index division in the loop
      const int tpp = PP < (int)blockDim.x ? PP : (int)blockDim.x, cpp = blockDim.x / tpp;
      const int cl0 = tid / tpp;
      if (cl0 < cpp)
        for (int pos = tid - cl0 * tpp; pos < PP; pos += tpp) {
          int so, lo;
          if (m.axis == 0) { lo = pos / SO; so = pos - lo * SO; } else { so = pos / LO; lo = pos - so * LO; }
          const int oo = o0 / st_ + so;
          const size_t a0_ = m.axis == 0 ? (((size_t)n * C + hg * CU) * Ho + lo) * Wo + oo : (((size_t)n * C + hg * CU) * Ho + oo) * Wo + lo;
          for (int cl = cl0; cl < CU; cl += cpp) {
            const float a0 = oa[cl * NO], a1 = NO == 2 ? oa[cl * NO + 1] : 0.f, bsum = ob[cl * NO] + (NO == 2 ? ob[cl * NO + 1] : 0.f);
            float acc = 0.f;
            for (int ds = 0; ds < st_; ++ds)
              for (int dl = 0; dl < st_; ++dl) {
                const float* o = Os + ((size_t)(so * st_ + ds) * NOCU + cl * NO) * L + lo * st_ + dl;
                acc = fmaf(a0, o[0], acc);
                if (NO == 2) acc = fmaf(a1, o[L], acc);
              }
            out[a0_ + (size_t)cl * oplane] = fmaf(acc, invp, bsum);
          }
        }
    }
  }
  phase_stamp(m, 8);
}

// developer hook: copies the phase stamps (MT_PH_MAX x MT_PH_CTAS, ns) of the last cooperative forward launch to the host
extern "C" int mt_debug_read_phases(unsigned long long* host, int n) {
  if (!host || n <= 0) return 1;
  const size_t cnt = (size_t)n < (size_t)MT_PH_MAX * MT_PH_CTAS ? (size_t)n : (size_t)MT_PH_MAX * MT_PH_CTAS;
  return cudaMemcpyFromSymbol(host, g_phase_ns, cnt * sizeof(unsigned long long)) == cudaSuccess ? 0 : 3;
}

// =============================================================================================== host side
static int g_sm_count[32] = {0};
int coop_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return 148;
  if (!g_sm_count[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}
// MT_COOP_MAX_CTAS (developer / emulation tier): cap on the cooperative grid; MT_COOP_STREAM=1 forces the streamed mode
int coop_max_ctas_env() {
  static const int v = getenv("MT_COOP_MAX_CTAS") ? atoi(getenv("MT_COOP_MAX_CTAS")) : 0;
  return v;
}
int coop_force_stream() {
  static const int v = getenv("MT_COOP_STREAM") ? atoi(getenv("MT_COOP_STREAM")) : 0;
  return v;
}
int coop_env_int(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

// RT of the cooperative kernels: register budget shrinks with d
int coop_rt(const Dims& m) {
  int rt = m.d == 1 ? 4 : (m.d <= 4 ? 2 : 1);
  static const int force = coop_env_int("MT_FORCE_RT");
  if (force == 1 || force == 2 || force == 4) rt = force < rt ? force : rt;
  while (rt > 1 && (m.L % rt)) rt >>= 1;
  return rt;
}

static int pow2_le(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }
// lanes per row of rows_stats: enough to give every thread a row share, at most one float4 chunk set per lane
int coop_rows_tpr(int nrows, int L) {
  int t = pow2_le(MT_COOP_THREADS / (nrows > 0 ? nrows : 1));
  if (t < 1) t = 1;
  if (t > 32) t = 32;
  const int chunks = (L % 4 == 0) ? L / 4 : L;
  while (t > 1 && t > chunks) t >>= 1;
  return t;
}

// fills `pl` for a grid of at most `maxctas` CTAs; returns false when one unit does not fit in shared memory
static bool coop_plan_fwd_hs(const Dims& m, int maxctas, size_t smem_cap, CoopPlan& pl, int hs_min) {
  const int C = m.C, L = m.L, D = m.d, GP = 2 * D, NO = m.no;
  pl.RT = coop_rt(m);
  pl.LB = L / pl.RT;
  pl.SPU = m.stride;                                    // pooling partners live in one unit
  if (m.other % pl.SPU) return false;
  const int nsb = m.B / pl.SPU;
  int hs = 1;
  while (hs < MT_G && nsb * hs * 2 <= maxctas) hs *= 2;  // split heads while that still adds CTAs
  if (hs < hs_min) hs = hs_min;                          // smaller units when one does not fit in shared memory
  static const int force_hs = coop_env_int("MT_COOP_HS");
  if (force_hs == 1 || force_hs == 2 || force_hs == 4 || force_hs == 8) hs = force_hs;
  pl.HS = hs; pl.HPU = MT_G / hs;
  pl.RU = pl.HPU * 4 * D; pl.CU = pl.HPU * GP; pl.NOCU = pl.CU * NO;
  pl.XP = pl.SPU * L;
  pl.nunits = nsb * hs;
  int G = pl.nunits < maxctas ? pl.nunits : maxctas;
  G -= G % hs;
  if (G < hs) return false;
  pl.upc = (pl.nunits + G - 1) / G;
  pl.contig = (hs == 1 && pl.upc > 1) ? 1 : 0;          // consecutive strips share a statistic group: fewer coefficient reloads
  if (pl.contig) G = (pl.nunits + pl.upc - 1) / pl.upc;
  pl.G = G;
  // key split: aim at >= 256 (row block, key range) items per unit, at most 8 lanes per row block
  int ks = 1;
  const int items0 = pl.SPU * pl.HPU * pl.LB;
  while (ks < 8 && items0 * ks < MT_COOP_THREADS && (pl.LB % (ks * 2)) == 0) ks *= 2;
  static const int force_ks = coop_env_int("MT_COOP_KS");
  if (force_ks > 0 && (pl.LB % force_ks) == 0 && (force_ks & (force_ks - 1)) == 0 && force_ks <= 32) ks = force_ks;
  pl.KS = ks;
  // GEMM tile: TN | L (a tile stays inside one strip), prefer >= 256 tiles
  const int tn_max = (L % 4 == 0) ? 4 : (L % 2 == 0 ? 2 : 1);
  if (tn_max == 4 && (pl.RU / 4) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 4; pl.tn = 4; }
  else if (tn_max == 4 && (pl.RU / 2) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 2; pl.tn = 4; }
  else if (tn_max >= 2) { pl.tm = 2; pl.tn = 2; }
  else { pl.tm = 2; pl.tn = 1; }
  pl.WTP = pl.RU + 4;
  pl.tpr_q = coop_rows_tpr(pl.RU, L);
  pl.tpr_o = coop_rows_tpr(pl.NOCU, L);
  pl.dL = make_divs(L); pl.dL4 = make_divs(L % 4 == 0 ? L / 4 : 1); pl.dLB = make_divs(pl.LB); pl.dKS = make_divs(pl.KS);
  pl.dHPU = make_divs(pl.HPU); pl.dSPU = make_divs(pl.SPU); pl.dRU = make_divs(pl.RU); pl.dNOCU = make_divs(pl.NOCU);
  pl.dPP = make_divs(1);
  // shared-memory carve
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
  pl.o_ca = take(pl.RU); pl.o_cb = take(pl.RU); pl.o_wsim = take(24);
  pl.o_oa = take(pl.NOCU); pl.o_ob = take(pl.NOCU);
  pl.o_redq = take(2 * pl.RU); pl.o_reds = take(48); pl.o_redo = take(2 * pl.NOCU);
  pl.tab_fl = m.pos ? 2 * (D + D * (D + 1) / 2) * L : 0;
  pl.o_tab = take(pl.tab_fl);
  pl.o_rp = take(m.pos ? 2 * m.gp * rel_pitch(L) : 0);
  pl.o_rs = take(m.pos ? 2 * m.gp * rel_pitch(L) : 0);
  pl.o_wt = take(C * pl.WTP);
  pl.o_x = take(C * pl.XP);
  pl.o_slot = off;
  pl.ys_fl = pl.SPU * pl.RU * L;
  pl.slot_fl = (pl.ys_fl + pl.SPU * pl.NOCU * L + 3) & ~3;
  const size_t fixed = (size_t)off * sizeof(float), slot = (size_t)pl.slot_fl * sizeof(float);
  if (fixed + slot > smem_cap) return false;
  int nslot = (int)((smem_cap - fixed) / slot);
  if (nslot > pl.upc) nslot = pl.upc;
  pl.resident = nslot >= pl.upc && !coop_force_stream();
  pl.nslot = pl.resident ? pl.upc : 1;
  pl.smem_bytes = fixed + (size_t)pl.nslot * slot;
  return true;
}

static bool coop_plan_fwd(const Dims& m, int maxctas, size_t smem_cap, CoopPlan& pl) {
  for (int hs = 1; hs <= MT_G; hs *= 2)
    if (coop_plan_fwd_hs(m, maxctas, smem_cap, pl, hs)) return true;
  return false;
}

template <int D, bool POS, int RT>
static int coop_launch_fwd(const Dims& m, const MtAxialParams& p, const float* x, float* out, double* stats, float* coef,
                           unsigned* cnt, float* yraw, float* oraw, float* lse, float* tabg, cudaStream_t st) {
  auto kern = k_axial_fwd<D, POS, RT>;
  static bool attr_set[64] = {};          // per device: function attributes live in a context (nn.DataParallel: one process, two GPUs)
  int adev = 0;
  if (cudaGetDevice(&adev) != cudaSuccess || adev < 0 || adev >= 64) return -1;
  if (!attr_set[adev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
    attr_set[adev] = true;
  }
  const int nsm = coop_sm_count();
  const int cap = coop_max_ctas_env();
  CoopPlan pl;
  // first choice: two CTAs per SM (more warps to hide latency) when the units are small enough; else one per SM
  bool ok = false;
  for (int occ = 2; occ >= 1 && !ok; --occ) {
    int maxctas = nsm * occ;
    if (cap > 0 && maxctas > cap) maxctas = cap;
    const size_t smem_cap = occ == 2 ? 110 * 1024 : 225 * 1024;
    if (!coop_plan_fwd(m, maxctas, smem_cap, pl)) continue;
    if (occ == 2) {
      if (pl.G <= nsm) { ok = true; break; }              // fits one per SM anyway
      static thread_local size_t q_smem[8] = {0};                // tiny cache: the query is a driver call, the launch is hot
      static thread_local int q_nb[8] = {0}, q_n = 0;
      int nb = -1;
      for (int i = 0; i < q_n; ++i) if (q_smem[i] == pl.smem_bytes) nb = q_nb[i];
      if (nb < 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, MT_COOP_THREADS, pl.smem_bytes) != cudaSuccess) return -1;
        if (q_n < 8) { q_smem[q_n] = pl.smem_bytes; q_nb[q_n] = nb; ++q_n; }
      }
      if (nb >= 2) ok = true;
    } else ok = true;
  }
  if (!ok) return 0;              // not covered (one unit does not fit): the caller falls back to the launch chain
  Dims mm = m; MtAxialParams pp = p;
  MT_COOP_LAUNCH((kern), pl.G, MT_COOP_THREADS, pl.smem_bytes, st, mm, pp, pl, x, out, stats, coef, cnt, yraw, oraw, lse, tabg);
  return g_last_cuda == cudaSuccess ? 1 : -1;
}

// returns launches done (1), 0 when the shape is not covered, < 0 on error
int launch_coop_fwd(const Dims& m, const MtAxialParams& p, const float* x, float* out, double* stats, float* coef,
                    unsigned* cnt, float* yraw, float* oraw, float* lse, float* tabg, cudaStream_t st) {
  const int rt = coop_rt(m);
#define GO(DD, RR)                                                                                              \
  return m.pos ? coop_launch_fwd<DD, true, RR>(m, p, x, out, stats, coef, cnt, yraw, oraw, lse, tabg, st)         \
               : coop_launch_fwd<DD, false, RR>(m, p, x, out, stats, coef, cnt, yraw, oraw, lse, tabg, st)
  if (m.d == 1) { if (rt == 4) { GO(1, 4); } if (rt == 2) { GO(1, 2); } GO(1, 1); }
  if (m.d == 2) { if (rt == 2) { GO(2, 2); } GO(2, 1); }
  if (m.d == 4) { if (rt == 2) { GO(4, 2); } GO(4, 1); }
  if (m.d == 8) { GO(8, 1); }
#undef GO
  return 0;
}
