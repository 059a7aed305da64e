// axial_coop_fwd.cu -- one AxialAttention* forward (reference axialnet.py:142-189 / :52-92 / :222-253) as ONE persistent
// cooperative kernel.  Phases (a grid barrier between them in training mode; none in eval mode):
//
//   P1  stage the unit's x strips (NCHW gather) and its slice of W_qkv, Y = W X (register-tiled fp32), keep Y in shared
//       memory, write y_raw for the backward, sum / sumsq per projection channel          -> bn_qkv statistics
//   P2  normalise Y in place; similarity statistics in CLOSED FORM (no L^2 pass):
//         sum_ij qk   = sum_c (sum_i q_ci)(sum_j k_cj)           sum_ij qk^2 = <Q Q^T, K K^T>_F   (d x d Gram matrices)
//         sum_ij qr   = sum_i q_i . PR_i                          sum_ij qr^2 = sum_i q_i^T M_i q_i
//       with PR_i / M_i the length-L window sums of R and R R^T (Toeplitz: relative[c, i-j+L-1], axialnet.py:155), built
//       once per CTA; the transposed kr term has the same form in k                         -> bn_similarity statistics
//   P3  softmax_j(sum_t w_t S_t) -> [P v | P Rv] with the scores in registers (sliding Toeplitz windows); a row block's keys
//       are split over KS adjacent lanes and merged with shuffles; o_raw, log-sum-exp, sum / sumsq  -> bn_output statistics
//   P4  BN(bn_output) affine + pair sum + NCHW un-permute (+ AvgPool) straight from the resident o_raw
//
// Every CTA derives the BN coefficients it needs from the fp64 accumulators itself; the last CTA writes the coefficient
// table for the backward and applies the running-statistics EMA after P4, off everybody's critical path.
#include "axial_coop.cuh"
extern thread_local cudaError_t g_last_cuda;

namespace {

__device__ __forceinline__ float warp_sum(float v) { return seg_sum(v, 32); }

// ---- projection micro-GEMM of one unit: Y[RU x XP] = Wt^T[RU x C] . Xs[C x XP], thread tile TM x TN --------------
template <int TM, int TN>
__device__ __forceinline__ void unit_gemm(const float* __restrict__ Wt, int WTP, const float* __restrict__ Xs, int XP, int C,
                                          int RU, int L, float* __restrict__ Ys, float* __restrict__ yraw_u, int ystride) {
  const int ntc = XP / TN, ntiles = (RU / TM) * ntc;
  for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
    const int ct = tile % ntc, r0 = (tile / ntc) * TM, c0 = ct * TN;
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

// per-row sum / sumsq of a [nrows][len] shared-memory block (row pitch `pitch`, rows of strip s at s*sstride): one warp
// per row, fixed summation order, result ADDED to red[2*row], red[2*row+1] (a row belongs to exactly one warp: no atomics)
__device__ __forceinline__ void rows_stats(const float* __restrict__ blk, int nrows, int L, int nstrips, int sstride,
                                           float* __restrict__ red) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = w; r < nrows; r += nw) {
    float s1 = 0.f, s2 = 0.f;
    for (int s = 0; s < nstrips; ++s) {
      const float* row = blk + (size_t)s * sstride + (size_t)r * L;
      for (int l = lane; l < L; l += 32) { const float v = row[l]; s1 += v; s2 = fmaf(v, v, s2); }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { red[2 * r] += s1; red[2 * r + 1] += s2; }
  }
}

}  // namespace

template <int D, bool POS, int RT>
__global__ void __launch_bounds__(MT_COOP_THREADS)
k_axial_fwd(Dims m, MtAxialParams p, CoopPlan pl, const float* __restrict__ x, float* __restrict__ out,
            double* __restrict__ stats, float* __restrict__ coef, unsigned* __restrict__ cnt, float* __restrict__ yraw,
            float* __restrict__ oraw, float* __restrict__ lse) {
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

  // ---------------------------------------------------------------- one-time staging: W slice, Toeplitz tables
  for (int idx = tid; idx < RU * C; idx += blockDim.x) {
    const int r = idx / C, k = idx - r * C;
    Wt[k * pl.WTP + r] = __ldg(p.w_qkv + (size_t)(row0 + r) * C + k);
  }
  if (POS) load_rel2(m, p.relative, Rp, Rs);
  __syncthreads();
  if (POS && train) {
    // window sums over mm in [i, i+L-1] of R[c][mm] (PR) and w * R[c][mm] R[c2][mm] (M; w = 2 off the diagonal)
    const int per = D + DD;
    for (int e = tid; e < 2 * per * L; e += blockDim.x) {
      const int i = e % L, k = (e / L) % per, which = e / (L * per);      // which: 0 = q rows, 1 = k rows of `relative`
      const float* R0 = Rp + (which * D) * RP;
      float acc = 0.f;
      float* dst;
      if (k < D) {
        const float* r = R0 + k * RP + i;
        for (int mm = 0; mm < L; ++mm) acc += r[mm];
        dst = tab + (which * D + k) * L + i;
      } else {
        int pc = k - D, c1 = 0;
        while (pc >= D - c1) { pc -= D - c1; ++c1; }
        const int c2 = c1 + pc;
        const float* r1 = R0 + c1 * RP + i;
        const float* r2 = R0 + c2 * RP + i;
        for (int mm = 0; mm < L; ++mm) acc = fmaf(r1[mm], r2[mm], acc);
        if (c1 != c2) acc *= 2.f;
        dst = tab + (2 * D + which * DD + (k - D)) * L + i;
      }
      *dst = acc;
    }
  }
  __syncthreads();

  // =================================================================================== P1: projection + bn_qkv sums
  {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Ys = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      if (!pl.resident) __syncthreads();      // the single slot is reused: the previous unit's readers must be done
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
      // x strips -> Xs[c][s*L + l]
      if (m.axis == 1 && (L & 3) == 0) {
        const int L4 = L >> 2, total = C * SPU * L4;
        for (int base = 0; base < total; base += 4 * blockDim.x) {
          float4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i4 = base + q * blockDim.x + tid;
            if (i4 < total) {
              const int l4 = i4 % L4, s = (i4 / L4) % SPU, c = i4 / (L4 * SPU);
              v[q] = __ldg(reinterpret_cast<const float4*>(x + nchw_addr(m, b0 + s, c, 0)) + l4);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i4 = base + q * blockDim.x + tid;
            if (i4 < total) {
              const int l4 = i4 % L4, s = (i4 / L4) % SPU, c = i4 / (L4 * SPU);
              *reinterpret_cast<float4*>(Xs + c * XP + s * L + 4 * l4) = v[q];
            }
          }
        }
      } else {
        const int total = C * XP;
        const bool sfast = m.axis == 0;        // axis 0: neighbouring strips are neighbouring addresses
        for (int base = 0; base < total; base += 4 * blockDim.x) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = base + q * blockDim.x + tid;
            if (i < total) {
              int s, l, c;
              if (sfast) { s = i % SPU; l = (i / SPU) % L; c = i / XP; } else { l = i % L; s = (i / L) % SPU; c = i / XP; }
              v[q] = __ldg(x + nchw_addr(m, b0 + s, c, l));
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = base + q * blockDim.x + tid;
            if (i < total) {
              int s, l, c;
              if (sfast) { s = i % SPU; l = (i / SPU) % L; c = i / XP; } else { l = i % L; s = (i / L) % SPU; c = i / XP; }
              Xs[c * XP + s * L + l] = v[q];
            }
          }
        }
      }
      __syncthreads();
      float* yu = yraw + ((size_t)b0 * 2 * C + row0) * L;
      const int ystride = 2 * C * L;
      if (pl.tm == 4 && pl.tn == 4) unit_gemm<4, 4>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else if (pl.tm == 2 && pl.tn == 4) unit_gemm<2, 4>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else if (pl.tm == 2 && pl.tn == 2) unit_gemm<2, 2>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      else unit_gemm<2, 1>(Wt, pl.WTP, Xs, XP, C, RU, L, Ys, yu, ystride);
      __syncthreads();
      if (train) rows_stats(Ys, RU, L, SPU, RU * L, redq);
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
  if (train) grid_barrier(bar, G, epoch);

  // =================================================================================== P2: normalise + similarity sums
  // (eval mode: the normalisation happens at the top of P3 and there is nothing to accumulate)
  if (train) {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = cta + k * G;
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
        for (int r = tid; r < RU; r += blockDim.x) {
          const int ch = row0 + r;
          BnCoef c = bn_coef(1, stat_get(m, stats, grp, m.nstat_f, m.f_qkv_sum + ch),
                             stat_get(m, stats, grp, m.nstat_f, m.f_qkv_sq + ch), m.inv_n1, 1.f, p.bn_qkv_w[ch],
                             p.bn_qkv_b[ch], 0.f, 1.f, m.eps);
          ca[r] = c.a; cb[r] = c.b;
        }
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
      // closed-form statistics on the NORMALISED q, k (the in-place normalisation of resident slots happens here too)
      for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
        const int r = (i / L) % RU;
        Ys[i] = fmaf(ca[r], Ys[i], cb[r]);
      }
      __syncthreads();
      {
        const int w = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
        for (int pr = w; pr < SPU * HPU; pr += nw) {
          const int hl = pr % HPU, s = pr / HPU, g = hg * HPU + hl;
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
    grid_barrier(bar, G, epoch);
  }

  // =================================================================================== P3: softmax, P v, P Rv
  {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = cta + k * G;
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
        if (tid < NT * HPU) {
          const int t = tid / HPU, g = hg * HPU + tid % HPU, ch = t * MT_G + g;
          const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
          BnCoef c = bn_coef(m.training, train ? stat_get(m, stats, grp, m.nstat_f, m.f_sim_sum + ch) : 0.0,
                             train ? stat_get(m, stats, grp, m.nstat_f, m.f_sim_sq + ch) : 0.0, m.inv_n2, f, p.bn_sim_w[ch],
                             p.bn_sim_b[ch], train ? 0.f : p.bn_sim_rm[ch], train ? 1.f : p.bn_sim_rv[ch], m.eps);
          wsim[ch] = c.a * MT_LOG2E;
        }
        if (!train)
          for (int r = tid; r < RU; r += blockDim.x) {
            const int ch = row0 + r;
            BnCoef c = bn_coef(0, 0.0, 0.0, m.inv_n1, 1.f, p.bn_qkv_w[ch], p.bn_qkv_b[ch], p.bn_qkv_rm[ch], p.bn_qkv_rv[ch], m.eps);
            ca[r] = c.a; cb[r] = c.b;
          }
        if (train && !pl.resident)
          for (int r = tid; r < RU; r += blockDim.x) {
            const int ch = row0 + r;
            BnCoef c = bn_coef(1, stat_get(m, stats, grp, m.nstat_f, m.f_qkv_sum + ch),
                               stat_get(m, stats, grp, m.nstat_f, m.f_qkv_sq + ch), m.inv_n1, 1.f, p.bn_qkv_w[ch],
                               p.bn_qkv_b[ch], 0.f, 1.f, m.eps);
            ca[r] = c.a; cb[r] = c.b;
          }
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
        for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
          const int r = (i / L) % RU;
          Ys[i] = fmaf(ca[r], Ys[i], cb[r]);
        }
        __syncthreads();
      }

      const int nitems = (m.dbg & 4) ? 0 : SPU * HPU * LB * KS;
      const int kpb = LB / KS;                                   // key blocks per lane
      for (int base = 0; base < nitems; base += blockDim.x) {
        const int item = base + tid;
        const bool act = item < nitems;
        const int it = act ? item : 0;
        const int ks = it % KS, t = (it / KS) % LB, hl = (it / (KS * LB)) % HPU, s = it / (KS * LB * HPU);
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
            if ((c % KS) != ks) continue;              // the KS lanes share the stores
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
      if (train) rows_stats(Os, NOCU, L, SPU, NOCU * L, redo);
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
  if (train) grid_barrier(bar, G, epoch);

  // =================================================================================== P4: bn_output + un-permute (+ pool)
  {
    const int st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_, SO = SPU / st_, LO = L / st_;
    const float invp = 1.f / (float)(st_ * st_);
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* Os = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl + pl.ys_fl;
      if (!pl.resident) __syncthreads();
      if (grp != cur) {
        __syncthreads();
        for (int r = tid; r < NOCU; r += blockDim.x) {
          const int ch = och0 + r;
          const float f = (NO == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
          BnCoef c = bn_coef(m.training, train ? stat_get(m, stats, grp, m.nstat_f, m.f_out_sum + ch) : 0.0,
                             train ? stat_get(m, stats, grp, m.nstat_f, m.f_out_sq + ch) : 0.0, m.inv_n1, f, p.bn_out_w[ch],
                             p.bn_out_b[ch], train ? 0.f : p.bn_out_rm[ch], train ? 1.f : p.bn_out_rv[ch], m.eps);
          oa[r] = c.a; ob[r] = c.b;
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
      const int total = CU * SO * LO;
      for (int i = tid; i < total; i += blockDim.x) {
        int so, lo, cl;
        if (m.axis == 0) { so = i % SO; lo = (i / SO) % LO; cl = i / (SO * LO); } else { lo = i % LO; so = (i / LO) % SO; cl = i / (SO * LO); }
        const float a0 = oa[cl * NO], a1 = NO == 2 ? oa[cl * NO + 1] : 0.f, bsum = ob[cl * NO] + (NO == 2 ? ob[cl * NO + 1] : 0.f);
        float acc = 0.f;
        for (int ds = 0; ds < st_; ++ds)
          for (int dl = 0; dl < st_; ++dl) {
            const float* o = Os + ((size_t)(so * st_ + ds) * NOCU + cl * NO) * L + lo * st_ + dl;
            acc = fmaf(a0, o[0], acc);
            if (NO == 2) acc = fmaf(a1, o[L], acc);
          }
        const int c = hg * CU + cl, oo = o0 / st_ + so;
        const size_t addr = m.axis == 0 ? (((size_t)n * C + c) * Ho + lo) * Wo + oo : (((size_t)n * C + c) * Ho + oo) * Wo + lo;
        out[addr] = fmaf(acc, invp, bsum);
      }
    }
  }

  // ---- coefficient tables for the backward + running statistics: last CTA, after its own work (nobody waits on it)
  if (cta == G - 1) {
    __syncthreads();
    fin_qkv(m, p, stats, coef);
    fin_sim(m, p, stats, coef);
    fin_out(m, p, stats, coef);
  }
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
static int coop_force_stream() {
  static const int v = getenv("MT_COOP_STREAM") ? atoi(getenv("MT_COOP_STREAM")) : 0;
  return v;
}
static int env_int(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

static int pow2_floor_i(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

// RT of the cooperative kernels: register budget shrinks with d
static int coop_rt(const Dims& m) {
  int rt = m.d == 1 ? 4 : (m.d <= 4 ? 2 : 1);
  static const int force = env_int("MT_FORCE_RT");
  if (force == 1 || force == 2 || force == 4) rt = force < rt ? force : rt;
  while (rt > 1 && (m.L % rt)) rt >>= 1;
  return rt;
}

// fills `pl` for a grid of at most `maxctas` CTAs; returns false when one unit does not fit in shared memory
static bool coop_plan_fwd(const Dims& m, int maxctas, size_t smem_cap, CoopPlan& pl) {
  const int C = m.C, L = m.L, D = m.d, GP = 2 * D, NO = m.no;
  pl.RT = coop_rt(m);
  pl.LB = L / pl.RT;
  pl.SPU = m.stride;                                    // pooling partners live in one unit
  if (m.other % pl.SPU) return false;
  const int nsb = m.B / pl.SPU;
  int hs = 1;
  while (hs < MT_G && nsb * hs * 2 <= maxctas) hs *= 2;  // split heads while that still adds CTAs
  static const int force_hs = env_int("MT_COOP_HS");
  if (force_hs == 1 || force_hs == 2 || force_hs == 4 || force_hs == 8) hs = force_hs;
  pl.HS = hs; pl.HPU = MT_G / hs;
  pl.RU = pl.HPU * 4 * D; pl.CU = pl.HPU * GP; pl.NOCU = pl.CU * NO;
  pl.XP = pl.SPU * L;
  pl.nunits = nsb * hs;
  int G = pl.nunits < maxctas ? pl.nunits : maxctas;
  G -= G % hs;
  if (G < hs) return false;
  pl.G = G;
  pl.upc = (pl.nunits + G - 1) / G;
  // key split: aim at >= 256 (row block, key range) items per unit, at most 8 lanes per row block
  int ks = 1;
  const int items0 = pl.SPU * pl.HPU * pl.LB;
  while (ks < 8 && items0 * ks < MT_COOP_THREADS && (pl.LB % (ks * 2)) == 0) ks *= 2;
  static const int force_ks = env_int("MT_COOP_KS");
  if (force_ks > 0 && (pl.LB % force_ks) == 0 && (force_ks & (force_ks - 1)) == 0 && force_ks <= 32) ks = force_ks;
  pl.KS = ks;
  // GEMM tile: TN | L (a tile stays inside one strip), prefer >= 256 tiles
  const int tn_max = (L % 4 == 0) ? 4 : (L % 2 == 0 ? 2 : 1);
  if (tn_max == 4 && (pl.RU / 4) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 4; pl.tn = 4; }
  else if (tn_max == 4 && (pl.RU / 2) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 2; pl.tn = 4; }
  else if (tn_max >= 2) { pl.tm = 2; pl.tn = 2; }
  else { pl.tm = 2; pl.tn = 1; }
  pl.WTP = pl.RU + 4;
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

template <int D, bool POS, int RT>
static int coop_launch_fwd(const Dims& m, const MtAxialParams& p, const float* x, float* out, double* stats, float* coef,
                           unsigned* cnt, float* yraw, float* oraw, float* lse, cudaStream_t st) {
  auto kern = k_axial_fwd<D, POS, RT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
    attr_set = true;
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
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, MT_COOP_THREADS, pl.smem_bytes) != cudaSuccess) return -1;
      if (nb >= 2) ok = true;
    } else ok = true;
  }
  if (!ok) return -2;
  Dims mm = m; MtAxialParams pp = p;
  MT_COOP_LAUNCH((kern), pl.G, MT_COOP_THREADS, pl.smem_bytes, st, mm, pp, pl, x, out, stats, coef, cnt, yraw, oraw, lse);
  return g_last_cuda == cudaSuccess ? 1 : -1;
}

// returns launches done (1), 0 when the shape is not covered, < 0 on error
int launch_coop_fwd(const Dims& m, const MtAxialParams& p, const float* x, float* out, double* stats, float* coef,
                    unsigned* cnt, float* yraw, float* oraw, float* lse, cudaStream_t st) {
  const int rt = coop_rt(m);
#define GO(DD, RR)                                                                                              \
  return m.pos ? coop_launch_fwd<DD, true, RR>(m, p, x, out, stats, coef, cnt, yraw, oraw, lse, st)               \
               : coop_launch_fwd<DD, false, RR>(m, p, x, out, stats, coef, cnt, yraw, oraw, lse, st)
  if (m.d == 1) { if (rt == 4) { GO(1, 4); } if (rt == 2) { GO(1, 2); } GO(1, 1); }
  if (m.d == 2) { if (rt == 2) { GO(2, 2); } GO(2, 1); }
  if (m.d == 4) { if (rt == 2) { GO(4, 2); } GO(4, 1); }
  if (m.d == 8) { GO(8, 1); }
#undef GO
  return 0;
}
