// axial_coop_bwd.cu -- backward of one axial-attention op as ONE persistent cooperative kernel (+ the weight-gradient GEMM
// on the forked side stream).  The reference has no hand-written backward (train.py:160 -> torch.autograd over
// axialnet.py:142-189); the closed forms are the ones pinned in oracle/axial_np.py::axial_backward.
//
// Every BatchNorm backward is the per-channel affine map  dx = A*dy + Bc + Cc*x  whose coefficients need two grid-wide sums
// (sum dy, sum dy*x); the three BNs of the op are three grid barriers:
//
//   B1  gather dout (NCHW, through AvgPool) into strips g, load o_raw; sum g, sum g*o_raw             -> bn_output backward
//   B2  dO = A g + Bc + Cc o_raw (in place), D_i = sum_c dO.O, stage + normalise q,k,v, log-sum-exp;
//       T[t,head] = sum dZ * S_t  (row-owned sweep; P recomputed from the saved log-sum-exp)           -> bn_similarity backward
//   B3  rows: dq ; columns: dk, dv ; diagonals: d relative -- all on the SAME resident staging;
//       dy (w.r.t. the normalised projection) -> shared memory + workspace; sum dy, sum dy*y_raw       -> bn_qkv backward
//   B4  dy_raw = A dy + Bc + Cc y_raw, dX = W^T dy_raw per (strip block, input-channel group), NCHW scatter;
//       the owner CTAs publish the bn_qkv backward coefficients for k_bwd_dw; the last CTA writes the BN / gate gradients
//
// Unit = SPU strips x one head group (as in the forward); in B4 the same CTA index serves an INPUT-channel group instead
// (dX sums over all heads), reading the other head groups' dy from the workspace.
#include "axial_coop.cuh"
#include "axial_diag_tile.cuh"
#include "axial_fused_bwd.cuh"
extern thread_local cudaError_t g_last_cuda;

struct CoopPlanB {
  DivS dL, dL4, dLB, dKS, dHPU, dSPU, dRU, dNOCU, dCU, dRL, dK2;
  int G, SPU, HS, HPU, RU, CU, NOCU, RT, KS, LB, XP, nunits, upc, nslot, resident, contig;
  int CU4;               // input channels per CTA in B4 (C / HS)
  int tm, tn, WTP;       // dX GEMM thread tile, pitch of the W slice
  int tpr_o, tpr_q;
  // shared-memory carve (float offsets)
  int o_ca, o_cb, o_wsim, o_sA, o_sB, o_sC, o_oA, o_oB, o_oC, o_redo, o_reds, o_redq, o_rp, o_rs, o_drel, o_slot;
  int f_yn, f_do, f_gs, f_dl, f_ls, f_dy, slot_fl;          // offsets inside a slot
  int o_qA, o_qB, o_qC, o_w4, o_dyr, o_dxs;                 // B4 staging (aliases the slots)
  // single-sweep path (axial_fused_bwd.cuh): lanes per (strip, head) group, segments of L/RT diagonal blocks, extra carve
  int fused, NLg, NSEG, VP, o_wtab, o_gram, o_vacc;
  DivS dNL;
  size_t smem_bytes;
};

namespace {

// n contiguous floats global -> shared, 16 bytes at a time when both sides allow it
__device__ __forceinline__ void copy_async(float* dst, const float* src, int n) {
  if (((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((smem_u32(dst) & 15) == 0)) {
    for (int i = 4 * threadIdx.x; i < n; i += 4 * blockDim.x) cp_async16(dst + i, src + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(dst + i, src + i);
  }
}

// dX micro-GEMM: out[RU x XP] = Wt^T[RU x K] . Xs[K x XP] into shared memory (row pitch XP)
template <int TM, int TN>
__device__ __forceinline__ void gemm_smem(const float* __restrict__ Wt, int WTP, const float* __restrict__ Xs, int XP, int K,
                                          int RU, float* __restrict__ out) {
  const int ntc = XP / TN, ntiles = (RU / TM) * ntc;
  for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
    const int rt_ = tile / ntc, ct = tile - rt_ * ntc, r0 = rt_ * TM, c0 = ct * TN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float a[TM], b[TN];
      ldv<TM>(a, Wt + k * WTP + r0);
      ldv<TN>(b, Xs + k * XP + c0);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      typedef typename VecT<TN>::T V;
      *reinterpret_cast<V*>(out + (size_t)(r0 + i) * XP + c0) = *reinterpret_cast<V*>(acc[i]);
    }
  }
}

__device__ __forceinline__ float warp_sum32(float v) { return seg_sum(v, 32); }

// BN backward coefficients of one channel from the forward / backward fp64 accumulators (one out-of-line copy)
__device__ __forceinline__ BnBwdCoef bwd_coef_from_stats(const Dims& m, const double* stats, const double* bstats, int grp,
                                                             int f_sum, int f_sq, int fch, double inv_n, float f, float gamma,
                                                             float rv, int b_s1, int b_s2, bool s1_zero) {
  const bool train = m.training != 0;
  return bn_bwd_coef(m.training, train ? stat_get(m, stats, grp, m.nstat_f, f_sum + fch) : 0.0,
                     train ? stat_get(m, stats, grp, m.nstat_f, f_sq + fch) : 0.0, inv_n, f, gamma, rv, m.eps,
                     (train && !s1_zero) ? stat_get(m, bstats, grp, m.nstat_b, b_s1) : 0.0,
                     stat_get(m, bstats, grp, m.nstat_b, b_s2));
}

}  // namespace

// symmetric pair index of (c, c') in a D x D matrix stored as its upper triangle
__device__ __forceinline__ int sym_idx(int c, int e, int D_) { if (c > e) { const int t = c; c = e; e = t; } return c * D_ - (c * (c - 1)) / 2 + (e - c); }

template <int D, bool POS, int RT, bool FUSED>
__global__ void __launch_bounds__(MT_COOP_THREADS, ((D <= 2 && !FUSED) || (FUSED && D * RT <= 2) ? 2 : 1))
k_axial_bwd(Dims m, MtAxialParams p, MtAxialGrads gr, CoopPlanB pl, const float* __restrict__ x, const float* __restrict__ dout,
            const double* __restrict__ stats, const float* __restrict__ coef, const float* __restrict__ yraw,
            const float* __restrict__ oraw, const float* __restrict__ lse, double* __restrict__ bstats,
            float* __restrict__ bcoef, unsigned* __restrict__ bcnt, float* __restrict__ dyh) {
  constexpr int GP = 2 * D, NO = POS ? 2 : 1, NT = POS ? 3 : 1;
  extern __shared__ __align__(16) float sm[];
  const int C = m.C, L = m.L, RP = rel_pitch(L), RL = 2 * L - 1;
  const int G = pl.G, SPU = pl.SPU, HS = pl.HS, HPU = pl.HPU, RU = pl.RU, NOCU = pl.NOCU, CU = pl.CU, KS = pl.KS,
            LB = pl.LB, XP = pl.XP;
  float* ca = sm + pl.o_ca;       float* cb = sm + pl.o_cb;        // [RU] bn_qkv forward affine
  float* wsim = sm + pl.o_wsim;   float* sA = sm + pl.o_sA;        // [3][8]
  float* sB = sm + pl.o_sB;       float* sC = sm + pl.o_sC;
  float* oA = sm + pl.o_oA;       float* oB = sm + pl.o_oB;   float* oC = sm + pl.o_oC;     // [NOCU] bn_output backward affine
  float* redo = sm + pl.o_redo;   // [CU] s1 | [NOCU] s2
  float* reds = sm + pl.o_reds;   // [3][8] T
  float* redq = sm + pl.o_redq;   // [RU][2]
  float* Rp = sm + pl.o_rp;       float* Rs = sm + pl.o_rs;
  float* dRs = sm + pl.o_drel;    // [2GP][2L-1] CTA partial of d relative
  float* slots = sm + pl.o_slot;
  const int tid = threadIdx.x, cta = blockIdx.x;
  const int hg = cta % HS, row0 = hg * RU, och0 = hg * NOCU, c0u = hg * CU;
  unsigned epoch = 0;
  unsigned* bar = bcnt + 8;
  const bool train = m.training != 0;
  const int st_ = m.stride, Ho = m.H / st_, Wo = m.W / st_;
  const float invp = 1.f / (float)(st_ * st_);
  const int nrel = POS ? 2 * GP * RL : 0;
  phase_stamp(m, 0);

  if (POS) load_rel2(m, p.relative, Rp, Rs);
  for (int k = tid; k < nrel; k += blockDim.x) dRs[k] = 0.f;
  if (POS && cta == 0) for (int k = tid; k < nrel; k += blockDim.x) gr.drelative[k] = 0.f;     // accumulated after barrier 2
  if (cta == 0 && tid < 4 && gr.dgates) gr.dgates[tid] = 0.f;                                  // accumulated after barrier 3
  for (int k = tid; k < 24; k += blockDim.x) reds[k] = 0.f;
  constexpr int DD = D * (D + 1) / 2;
  float* wtab = sm + pl.o_wtab;    // FUSED: PRq [D][L] | PRk [D][L] | Mq [DD][L] | Mk [DD][L]  (window sums of the Toeplitz tables)
  float* gram = sm + pl.o_gram;    // FUSED: per (strip, head) of the unit: K K^T [DD] | sum k [D] | Q Q^T [DD] | sum q [D]
  float* vacc = sm + pl.o_vacc;    // FUSED: Vq [DD][VP] | Bq [D][VP] | Vk [DD][VP] | Bk [D][VP]: S-parts of d relative, summed over the CTA's units
  if constexpr (FUSED) {
    if (POS && train) {
      __syncthreads();                                        // Rp staged
      const int ntab = 2 * (D + DD);
      for (int it = tid; it < ntab * L; it += blockDim.x) {
        const int tb = dv(it, pl.dL), i = it - tb * L;
        int c = 0, e = 0;
        const float* base = Rp;
        if (tb < 2 * D) { c = tb < D ? tb : tb - D; e = -1; base = Rp + (tb < D ? 0 : D) * RP; }
        else {
          int pi = tb - 2 * D;
          if (pi >= DD) { pi -= DD; base = Rp + D * RP; }
          for (c = 0; c < D; ++c) { if (pi < D - c) { e = c + pi; break; } pi -= D - c; }
        }
        const float* ra = base + c * RP + i;
        float acc = 0.f;
        if (e < 0) { for (int j = 0; j < L; ++j) acc += ra[j]; }
        else { const float* rb = base + e * RP + i; for (int j = 0; j < L; ++j) acc = fmaf(ra[j], rb[j], acc); }
        wtab[it] = acc;
      }
      for (int k = tid; k < 2 * (DD + D) * pl.VP; k += blockDim.x) vacc[k] = 0.f;
    }
  }

  // ================================================================== B1: g strips, o_raw, bn_output backward sums
  {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* slot = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      float* Gs = slot + pl.f_gs;      // [SPU][CU][L]
      float* Os = slot + pl.f_do;      // [SPU][NOCU][L]
      __syncthreads();
      if (grp != cur) {
        if (cur >= 0) {
          double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
          for (int r = tid; r < CU; r += blockDim.x) MT_GADD(m, bs + m.b_out_s1 + c0u + r, (double)redo[r]);
          for (int r = tid; r < NOCU; r += blockDim.x) MT_GADD(m, bs + m.b_out_s2 + och0 + r, (double)redo[CU + r]);
          __syncthreads();
        }
        for (int r = tid; r < CU + NOCU; r += blockDim.x) redo[r] = 0.f;
        cur = grp;
      }
      // o_raw rows of the unit (contiguous per strip)
      for (int s = 0; s < SPU; ++s) copy_async(Os + (size_t)s * NOCU * L, oraw + ((size_t)(b0 + s) * NO * C + och0) * L, NOCU * L);
      // g[s][cl][l] = dout[n, c, ..] / stride^2 (AvgPool backward)
      {
        const int n = b0 / m.other, o0 = b0 - n * m.other;
        for (int i = tid; i < SPU * CU * L; i += blockDim.x) {
          int s, l, cl;
          if (m.axis == 0) { const int rem = dv(i, pl.dSPU); s = i - rem * SPU; cl = dv(rem, pl.dL); l = rem - cl * L; }
          else { const int rem = dv(i, pl.dL); l = i - rem * L; cl = dv(rem, pl.dSPU); s = rem - cl * SPU; }
          const int oo = (o0 + s) / st_, lo = l / st_, c = c0u + cl;
          const size_t a = m.axis == 0 ? (((size_t)n * C + c) * Ho + lo) * Wo + oo : (((size_t)n * C + c) * Ho + oo) * Wo + lo;
          Gs[((size_t)s * CU + cl) * L + l] = __ldg(dout + a) * invp;
        }
      }
      cp_async_wait_all();
      __syncthreads();
      // per-channel sums: s1[c] = sum g, s2[ch] = sum g * o_raw   (tpr lanes per channel)
      {
        const int tpr = pl.tpr_o, sub = tid & (tpr - 1), rpp = blockDim.x / tpr;
        for (int r0 = 0; r0 < CU; r0 += rpp) {
          const int cl = r0 + tid / tpr;
          float a1 = 0.f, a2 = 0.f, a3 = 0.f;
          if (cl < CU)
            for (int s = 0; s < SPU; ++s) {
              const float* g = Gs + ((size_t)s * CU + cl) * L;
              const float* o = Os + ((size_t)s * NOCU + cl * NO) * L;
              for (int l = sub; l < L; l += tpr) {
                const float gv = g[l];
                a1 += gv; a2 = fmaf(gv, o[l], a2);
                if (NO == 2) a3 = fmaf(gv, o[L + l], a3);
              }
            }
          a1 = seg_sum(a1, tpr); a2 = seg_sum(a2, tpr);
          if (NO == 2) a3 = seg_sum(a3, tpr);
          if (sub == 0 && cl < CU) {
            redo[cl] += a1; redo[CU + cl * NO] += a2;
            if (NO == 2) redo[CU + cl * NO + 1] += a3;
          }
        }
      }
    }
    if (cur >= 0) {
      __syncthreads();
      double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
      for (int r = tid; r < CU; r += blockDim.x) MT_GADD(m, bs + m.b_out_s1 + c0u + r, (double)redo[r]);
      for (int r = tid; r < NOCU; r += blockDim.x) MT_GADD(m, bs + m.b_out_s2 + och0 + r, (double)redo[CU + r]);
    }
  }
  phase_stamp(m, 1);
  grid_barrier(bar, G, epoch);
  phase_stamp(m, 2);

  // per-group staging of the forward coefficients (written by the forward kernel) and the bn_output backward affine
  auto group_coefs = [&](int grp) {
    const float* tq = coef + (size_t)grp * m.ncoef_f + m.cf_q;
    for (int r = tid; r < RU; r += blockDim.x) { ca[r] = tq[2 * (row0 + r)]; cb[r] = tq[2 * (row0 + r) + 1]; }
    if (tid < NT * MT_G) wsim[tid] = coef[(size_t)grp * m.ncoef_f + m.cf_s + 2 * tid];
    for (int r = tid; r < NOCU; r += blockDim.x) {
      const int ch = och0 + r;
      const float f = (NO == 2 && (ch & 1)) ? gate_val(p.f_sve) : gate_val(p.f_sv);
      BnBwdCoef c = bwd_coef_from_stats(m, stats, bstats, grp, m.f_out_sum, m.f_out_sq, ch, m.inv_n1, f, p.bn_out_w[ch],
                                        p.bn_out_rv[ch], m.b_out_s1 + ch / NO, m.b_out_s2 + ch, false);
      oA[r] = c.A; oB[r] = c.Bc; oC[r] = c.Cc;
    }
  };

  // dO, D, normalised q/k/v and log-sum-exp of one unit (B2 in resident mode; again in B3 when the slot is streamed)
  auto stage_unit = [&](float* slot, int b0, bool have_g) {
    float* Yn = slot + pl.f_yn; float* Os = slot + pl.f_do; float* Gs = slot + pl.f_gs;
    float* Dl = slot + pl.f_dl; float* Ls = slot + pl.f_ls;
    for (int s = 0; s < SPU; ++s) {
      copy_async(Yn + (size_t)s * RU * L, yraw + ((size_t)(b0 + s) * 2 * C + row0) * L, RU * L);
      copy_async(Ls + (size_t)s * HPU * L, lse + ((size_t)(b0 + s) * MT_G + hg * HPU) * L, HPU * L);
    }
    if (!have_g) {
      for (int s = 0; s < SPU; ++s) copy_async(Os + (size_t)s * NOCU * L, oraw + ((size_t)(b0 + s) * NO * C + och0) * L, NOCU * L);
      const int n = b0 / m.other, o0 = b0 - n * m.other;
      for (int i = tid; i < SPU * CU * L; i += blockDim.x) {
        int s, l, cl;
        if (m.axis == 0) { const int rem = dv(i, pl.dSPU); s = i - rem * SPU; cl = dv(rem, pl.dL); l = rem - cl * L; }
        else { const int rem = dv(i, pl.dL); l = i - rem * L; cl = dv(rem, pl.dSPU); s = rem - cl * SPU; }
        const int oo = (o0 + s) / st_, lo = l / st_, c = c0u + cl;
        const size_t a = m.axis == 0 ? (((size_t)n * C + c) * Ho + lo) * Wo + oo : (((size_t)n * C + c) * Ho + oo) * Wo + lo;
        Gs[((size_t)s * CU + cl) * L + l] = __ldg(dout + a) * invp;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // dO = A g + Bc + Cc o (in place over o_raw) and D[s][hl][i] = sum over the head's channels of dO * O_raw
    for (int i = tid; i < SPU * HPU * L; i += blockDim.x) {
      const int r1 = dv(i, pl.dL), l = i - r1 * L, s = dv(r1, pl.dHPU), hl = r1 - s * HPU;
      float* o = Os + ((size_t)s * NOCU + hl * GP * NO) * L + l;
      const float* g = Gs + ((size_t)s * CU + hl * GP) * L + l;
      float acc = 0.f;
#pragma unroll 4
      for (int k = 0; k < GP * NO; ++k) {
        const int chl = hl * GP * NO + k;
        const float ov = o[(size_t)k * L], d = fmaf(oA[chl], g[(size_t)(k / NO) * L], fmaf(oC[chl], ov, oB[chl]));
        acc = fmaf(d, ov, acc);
        o[(size_t)k * L] = d;
      }
      Dl[i] = acc;
    }
    // y -> normalised q, k, v
    if ((L & 3) == 0) {
      float4* y4 = reinterpret_cast<float4*>(Yn);
      for (int i4 = tid; i4 < (SPU * RU * L) >> 2; i4 += blockDim.x) {
        const int r = md(dv(i4, pl.dL4), pl.dRU);
        const float ka = ca[r], kb = cb[r];
        float4 v = y4[i4];
        v.x = fmaf(ka, v.x, kb); v.y = fmaf(ka, v.y, kb); v.z = fmaf(ka, v.z, kb); v.w = fmaf(ka, v.w, kb);
        y4[i4] = v;
      }
    } else {
      for (int i = tid; i < SPU * RU * L; i += blockDim.x) { const int r = (i / L) % RU; Yn[i] = fmaf(ca[r], Yn[i], cb[r]); }
    }
    __syncthreads();
  };

  // ================================================================== B2 (single-sweep path): every L x L quantity in ONE pass
  if constexpr (FUSED) {
    const int NL = pl.NLg, NSEG = pl.NSEG;
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* slot = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      __syncthreads();
      if (grp != cur) {
        if (cur >= 0) {
          double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
          if (tid < NT * HPU) { const int t = tid / HPU, g = hg * HPU + tid % HPU; MT_GADD(m, bs + m.b_sim_t + t * MT_G + g, (double)reds[t * MT_G + g]); }
          __syncthreads();
          if (tid < 24) reds[tid] = 0.f;
        }
        group_coefs(grp);
        if (tid < NT * HPU) {           // A_t of the bn_similarity backward does not depend on T: known before the barrier
          const int t = tid / HPU, g = hg * HPU + tid % HPU, ch = t * MT_G + g;
          const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
          BnBwdCoef c = bwd_coef_from_stats(m, stats, bstats, grp, m.f_sim_sum, m.f_sim_sq, ch, m.inv_n2, f, p.bn_sim_w[ch],
                                            p.bn_sim_rv[ch], 0, m.b_sim_t + ch, true);
          sA[ch] = c.A;
        }
        cur = grp;
        __syncthreads();
      }
      stage_unit(slot, b0, pl.resident != 0);
      const float* Yn = slot + pl.f_yn; const float* dOs = slot + pl.f_do; const float* Dl = slot + pl.f_dl; const float* Ls = slot + pl.f_ls;
      float* dY = slot + pl.f_dy;
      for (int i = tid; i < SPU * RU * L; i += blockDim.x) dY[i] = 0.f;
      __syncthreads();
      const int nit = SPU * HPU * NL;
      for (int base = 0; base < nit; base += blockDim.x) {
        if (base + (tid & ~31) >= nit) break;                  // whole warp idle
        const int item = base + tid;
        const bool act = item < nit && !(m.dbg & 4);
        const int it = item < nit ? item : 0;
        const int gi = dv(it, pl.dNL), lig = tid & (NL - 1), s = dv(gi, pl.dHPU), hl = gi - s * HPU, g = hg * HPU + hl;
        mtf::Coefs cf;
        cf.wqk = wsim[g]; cf.wqr = POS ? wsim[MT_G + g] : 0.f; cf.wkr = POS ? wsim[2 * MT_G + g] : 0.f;
        cf.A1 = sA[g]; cf.A2 = POS ? sA[MT_G + g] : 0.f; cf.A3 = POS ? sA[2 * MT_G + g] : 0.f;
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int sg = 0; sg < NSEG; ++sg)
          mtf::fused_item<D, POS, RT>(L, RP, sg * NL + lig, lig, NL, NSEG == 1, act, Yn + ((size_t)s * RU + hl * 4 * D) * L,
                                      dOs + ((size_t)s * NOCU + hl * GP * NO) * L, Ls + (s * HPU + hl) * L, Dl + (s * HPU + hl) * L,
                                      Rp, cf, dRs, dY + ((size_t)s * RU + hl * 4 * D) * L, t0, t1, t2);
        t0 = seg_sum(t0, NL);
        if (POS) { t1 = seg_sum(t1, NL); t2 = seg_sum(t2, NL); }
        if (act && lig == 0) {
          atomicAdd(reds + g, t0);
          if (POS) { atomicAdd(reds + MT_G + g, t1); atomicAdd(reds + 2 * MT_G + g, t2); }
        }
      }
      if (!pl.resident) {                                      // streamed unit: the dZ parts wait in the workspace for the barrier
        __syncthreads();
        for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
          const int row = dv(i, pl.dL), l = i - row * L, s = dv(row, pl.dRU), r = row - s * RU;
          dyh[((size_t)(b0 + s) * 2 * C + row0 + r) * L + l] = dY[i];
        }
      }
    }
    if (cur >= 0) {
      __syncthreads();
      double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
      if (tid < NT * HPU) { const int t = tid / HPU, g = hg * HPU + tid % HPU; MT_GADD(m, bs + m.b_sim_t + t * MT_G + g, (double)reds[t * MT_G + g]); }
    }
  }
  // ================================================================== B2: T[t,head] = sum dZ * S_t
  if constexpr (!FUSED) {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* slot = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      __syncthreads();
      if (grp != cur) {
        if (cur >= 0) {
          double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
          if (tid < NT * HPU) { const int t = tid / HPU, g = hg * HPU + tid % HPU; MT_GADD(m, bs + m.b_sim_t + t * MT_G + g, (double)reds[t * MT_G + g]); }
          __syncthreads();
          if (tid < 24) reds[tid] = 0.f;
        }
        group_coefs(grp);
        cur = grp;
        __syncthreads();
      }
      stage_unit(slot, b0, pl.resident != 0);
      const float* Yn = slot + pl.f_yn; const float* dOs = slot + pl.f_do; const float* Dl = slot + pl.f_dl; const float* Ls = slot + pl.f_ls;
      const int nitems = (m.dbg & 4) ? 0 : SPU * HPU * LB * KS, kpb = LB / KS;
      for (int base = 0; base < nitems; base += blockDim.x) {
        const int item = base + tid;
        const bool act = item < nitems;
        const int it = act ? item : 0;
        const int i1 = dv(it, pl.dKS), ks = it - i1 * KS, i2 = dv(i1, pl.dLB), t = i1 - i2 * LB, s = dv(i2, pl.dHPU), hl = i2 - s * HPU;
        const int g = hg * HPU + hl;
        const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
        const float* dOb = dOs + ((size_t)s * NOCU + hl * GP * NO) * L;
        const float* lsp = Ls + (s * HPU + hl) * L;
        const float* Dp = Dl + (s * HPU + hl) * L;
        const float wqk = wsim[g], wqr = POS ? wsim[MT_G + g] : 0.f, wkr = POS ? wsim[2 * MT_G + g] : 0.f;
        const int wb0 = ks * kpb, wb1 = wb0 + kpb, KA0 = t + LB - wb0, KB0 = LB - t + wb0;
        float q[D][RT], dsv[GP][RT], dse[GP][RT], lsr[RT], Dr[RT];
#pragma unroll
        for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * L + RT * t);
#pragma unroll
        for (int c = 0; c < GP; ++c) {
          ldv<RT>(dsv[c], dOb + (c * NO) * L + RT * t);
          if (POS) ldv<RT>(dse[c], dOb + (c * NO + 1) * L + RT * t);
        }
        ldv<RT>(lsr, lsp + RT * t); ldv<RT>(Dr, Dp + RT * t);
        WinA<RT> wq[D]; WinB<RT> wk[D]; WinA<RT> wv[GP];
        if (POS) {
#pragma unroll
          for (int c = 0; c < D; ++c) { wq[c].init(Rp + c * RP, KA0); wk[c].init(Rs + (D + c) * RP, KB0); }
#pragma unroll
          for (int c = 0; c < GP; ++c) wv[c].init(Rp + (GP + c) * RP, KA0);
        }
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int wb = wb0; wb < wb1; ++wb) {
          float kk[D][RT], vv[GP][RT];
#pragma unroll
          for (int c = 0; c < D; ++c) ldv<RT>(kk[c], Yb + (D + c) * L + RT * wb);
#pragma unroll
          for (int c = 0; c < GP; ++c) ldv<RT>(vv[c], Yb + (GP + c) * L + RT * wb);
#pragma unroll
          for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int w = 0; w < RT; ++w) {
              float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
#pragma unroll
              for (int c = 0; c < D; ++c) {
                sqk = fmaf(q[c][r], kk[c][w], sqk);
                if (POS) { sqr = fmaf(q[c][r], wq[c].get(r, w), sqr); skr = fmaf(kk[c][w], wk[c].get(r, w), skr); }
              }
              float z = wqk * sqk;
              if (POS) z = fmaf(wqr, sqr, fmaf(wkr, skr, z));
              const float P = ex2(z - lsr[r]);
#pragma unroll
              for (int c = 0; c < GP; ++c) {
                dP = fmaf(dsv[c][r], vv[c][w], dP);
                if (POS) dP = fmaf(dse[c][r], wv[c].get(r, w), dP);
              }
              const float dZ = P * (dP - Dr[r]);
              t0 = fmaf(dZ, sqk, t0);
              if (POS) { t1 = fmaf(dZ, sqr, t1); t2 = fmaf(dZ, skr, t2); }
            }
          if (POS && wb + 1 < wb1) {
#pragma unroll
            for (int c = 0; c < D; ++c) { wq[c].advance(Rp + c * RP, KA0 - (wb - wb0) - 1); wk[c].advance(Rs + (D + c) * RP, KB0 + (wb - wb0) + 1); }
#pragma unroll
            for (int c = 0; c < GP; ++c) wv[c].advance(Rp + (GP + c) * RP, KA0 - (wb - wb0) - 1);
          }
        }
        if (!act) { t0 = t1 = t2 = 0.f; }
        // lanes of one head are contiguous: segmented sum over LB*KS lanes (or the warp, when a head spans several)
        const int segw = LB * KS < 32 ? LB * KS : 32;
        t0 = seg_sum(t0, segw);
        if (POS) { t1 = seg_sum(t1, segw); t2 = seg_sum(t2, segw); }
        if (act && (tid & (segw - 1)) == 0) {
          atomicAdd(reds + g, t0);
          if (POS) { atomicAdd(reds + MT_G + g, t1); atomicAdd(reds + 2 * MT_G + g, t2); }
        }
      }
    }
    if (cur >= 0) {
      __syncthreads();
      double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
      if (tid < NT * HPU) { const int t = tid / HPU, g = hg * HPU + tid % HPU; MT_GADD(m, bs + m.b_sim_t + t * MT_G + g, (double)reds[t * MT_G + g]); }
    }
  }
  phase_stamp(m, 3);
  grid_barrier(bar, G, epoch);
  phase_stamp(m, 4);

  // ================================================================== B3 (single-sweep path): T is complete -> add the closed-form
  // S_t / constant parts of dS_t = A_t dZ + Cc_t S_t + Bc_t  (no pair sweep), then the bn_qkv backward sums
  if constexpr (FUSED) {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* slot = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      __syncthreads();
      if (grp != cur) {
        if (cur >= 0) {
          double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
          for (int r = tid; r < RU; r += blockDim.x) {
            MT_GADD(m, bs + m.b_qkv_s1 + row0 + r, (double)redq[2 * r]);
            MT_GADD(m, bs + m.b_qkv_s2 + row0 + r, (double)redq[2 * r + 1]);
          }
          __syncthreads();
        }
        if (!pl.resident) group_coefs(grp);
        else {                                               // bn_qkv forward affine of THIS group (the stats below invert it)
          const float* tq = coef + (size_t)grp * m.ncoef_f + m.cf_q;
          for (int r = tid; r < RU; r += blockDim.x) { ca[r] = tq[2 * (row0 + r)]; cb[r] = tq[2 * (row0 + r) + 1]; }
        }
        if (tid < NT * HPU) {
          const int t = tid / HPU, g = hg * HPU + tid % HPU, ch = t * MT_G + g;
          const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
          BnBwdCoef c = bwd_coef_from_stats(m, stats, bstats, grp, m.f_sim_sum, m.f_sim_sq, ch, m.inv_n2, f, p.bn_sim_w[ch],
                                            p.bn_sim_rv[ch], 0, m.b_sim_t + ch, true);
          sA[ch] = c.A; sB[ch] = c.Bc; sC[ch] = c.Cc;
        }
        for (int r = tid; r < 2 * RU; r += blockDim.x) redq[r] = 0.f;
        cur = grp;
        __syncthreads();
      }
      float* Yn = slot + pl.f_yn;
      float* dY = slot + pl.f_dy;
      if (!pl.resident) {
        stage_unit(slot, b0, false);
        for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
          const int row = dv(i, pl.dL), l = i - row * L, s = dv(row, pl.dRU), r = row - s * RU;
          dY[i] = __ldcg(dyh + ((size_t)(b0 + s) * 2 * C + row0 + r) * L + l);
        }
        __syncthreads();
      }
      if (train) {
        // Gram matrices and sums of the unit's heads: one warp per (strip, head)
        const int w = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
        for (int gi = w; gi < SPU * HPU; gi += nw) {
          const int s = dv(gi, pl.dHPU), hl = gi - s * HPU;
          const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
          float kk[DD], qq[DD], ks[D], qs[D];
#pragma unroll
          for (int e = 0; e < DD; ++e) { kk[e] = 0.f; qq[e] = 0.f; }
#pragma unroll
          for (int c = 0; c < D; ++c) { ks[c] = 0.f; qs[c] = 0.f; }
          for (int j = lane; j < L; j += 32) {
            float qv[D], kv[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { qv[c] = Yb[c * L + j]; kv[c] = Yb[(D + c) * L + j]; qs[c] += qv[c]; ks[c] += kv[c]; }
            int e = 0;
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
              for (int f = c; f < D; ++f) { kk[e] = fmaf(kv[c], kv[f], kk[e]); qq[e] = fmaf(qv[c], qv[f], qq[e]); ++e; }
          }
          float* gm = gram + (size_t)gi * (2 * DD + 2 * D);
#pragma unroll
          for (int e = 0; e < DD; ++e) { kk[e] = warp_sum32(kk[e]); qq[e] = warp_sum32(qq[e]); }
#pragma unroll
          for (int c = 0; c < D; ++c) { ks[c] = warp_sum32(ks[c]); qs[c] = warp_sum32(qs[c]); }
          if (lane == 0) {
#pragma unroll
            for (int e = 0; e < DD; ++e) { gm[e] = kk[e]; gm[DD + D + e] = qq[e]; }
#pragma unroll
            for (int c = 0; c < D; ++c) { gm[DD + c] = ks[c]; gm[2 * DD + D + c] = qs[c]; }
          }
        }
        __syncthreads();
        const float* PRq = wtab; const float* PRk = wtab + D * L; const float* Mq = wtab + 2 * D * L; const float* Mk = Mq + DD * L;
        for (int it = tid; it < SPU * HPU * L; it += blockDim.x) {
          const int gi = dv(it, pl.dL), i = it - gi * L, s = dv(gi, pl.dHPU), hl = gi - s * HPU, g = hg * HPU + hl;
          const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
          float* dYb = dY + ((size_t)s * RU + hl * 4 * D) * L;
          const float* gm = gram + (size_t)gi * (2 * DD + 2 * D);
          const float C1 = sC[g], B1 = sB[g];
          const float C2 = POS ? sC[MT_G + g] : 0.f, B2 = POS ? sB[MT_G + g] : 0.f, C3 = POS ? sC[2 * MT_G + g] : 0.f, B3 = POS ? sB[2 * MT_G + g] : 0.f;
          float qv[D], kv[D];
#pragma unroll
          for (int c = 0; c < D; ++c) { qv[c] = Yb[c * L + i]; kv[c] = Yb[(D + c) * L + i]; }
#pragma unroll
          for (int c = 0; c < D; ++c) {
            float kkq = 0.f, qqk = 0.f, mq = 0.f, mk = 0.f;
#pragma unroll
            for (int e = 0; e < D; ++e) {
              const int si = sym_idx(c, e, D);
              kkq = fmaf(gm[si], qv[e], kkq); qqk = fmaf(gm[DD + D + si], kv[e], qqk);
              if (POS) { mq = fmaf(Mq[si * L + i], qv[e], mq); mk = fmaf(Mk[si * L + i], kv[e], mk); }
            }
            float aq = fmaf(C1, kkq, B1 * gm[DD + c]), ak = fmaf(C1, qqk, B1 * gm[2 * DD + D + c]);
            if (POS) { aq += fmaf(C2, mq, B2 * PRq[c * L + i]); ak += fmaf(C3, mk, B3 * PRk[c * L + i]); }
            dYb[c * L + i] += aq; dYb[(D + c) * L + i] += ak;
          }
        }
        if (POS) {
          // S-parts of d relative: per position i, summed over the unit's (strip, head)s with their coefficients
          const int VP = pl.VP;
          for (int i = tid; i < L; i += blockDim.x) {
            float vq[DD], vk[DD], bq[D], bk[D];
#pragma unroll
            for (int e = 0; e < DD; ++e) { vq[e] = 0.f; vk[e] = 0.f; }
#pragma unroll
            for (int c = 0; c < D; ++c) { bq[c] = 0.f; bk[c] = 0.f; }
            for (int gi = 0; gi < SPU * HPU; ++gi) {
              const int s = dv(gi, pl.dHPU), hl = gi - s * HPU, g = hg * HPU + hl;
              const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
              const float C2 = sC[MT_G + g], B2 = sB[MT_G + g], C3 = sC[2 * MT_G + g], B3 = sB[2 * MT_G + g];
              float qv[D], kv[D];
#pragma unroll
              for (int c = 0; c < D; ++c) { qv[c] = Yb[c * L + i]; kv[c] = Yb[(D + c) * L + i]; bq[c] = fmaf(B2, qv[c], bq[c]); bk[c] = fmaf(B3, kv[c], bk[c]); }
              int e = 0;
#pragma unroll
              for (int c = 0; c < D; ++c)
#pragma unroll
                for (int f = c; f < D; ++f) { vq[e] = fmaf(C2 * qv[c], qv[f], vq[e]); vk[e] = fmaf(C3 * kv[c], kv[f], vk[e]); ++e; }
            }
#pragma unroll
            for (int e = 0; e < DD; ++e) { vacc[e * VP + i] += vq[e]; vacc[(DD + D + e) * VP + i] += vk[e]; }
#pragma unroll
            for (int c = 0; c < D; ++c) { vacc[(DD + c) * VP + i] += bq[c]; vacc[(2 * DD + D + c) * VP + i] += bk[c]; }
          }
        }
        __syncthreads();
      }
      // dy (w.r.t. the normalised projection) -> workspace for B4 / k_bwd_dw
      for (int i = tid; i < SPU * RU * L; i += blockDim.x) {
        const int row = dv(i, pl.dL), l = i - row * L, s = dv(row, pl.dRU), r = row - s * RU;
        dyh[((size_t)(b0 + s) * 2 * C + row0 + r) * L + l] = dY[i];
      }
      // bn_qkv backward sums: s1 = sum dy, s2 = sum dy * y_raw.  The staged rows hold yn = ca*y_raw + cb, so
      // s2 = (sum dy*yn - cb*s1) / ca straight from shared memory (no L2 round trip); a zero scale falls back to y_raw
      {
        const int tpr = pl.tpr_q, sub = tid & (tpr - 1), rpp = blockDim.x / tpr;
        for (int r0 = 0; r0 < RU; r0 += rpp) {
          const int r = r0 + tid / tpr;
          float a1 = 0.f, a2 = 0.f;
          const float car = r < RU ? ca[r] : 1.f;
          if (r < RU)
            for (int s = 0; s < SPU; ++s) {
              const float* dy = dY + ((size_t)s * RU + r) * L;
              if (car != 0.f) {
                const float* yn = Yn + ((size_t)s * RU + r) * L;
                for (int l = sub; l < L; l += tpr) { const float d = dy[l]; a1 += d; a2 = fmaf(d, yn[l], a2); }
              } else {
                const float* yr = yraw + ((size_t)(b0 + s) * 2 * C + row0 + r) * L;
                for (int l = sub; l < L; l += tpr) { const float d = dy[l]; a1 += d; a2 = fmaf(d, __ldg(yr + l), a2); }
              }
            }
          a1 = seg_sum(a1, tpr); a2 = seg_sum(a2, tpr);
          if (sub == 0 && r < RU) {
            redq[2 * r] += a1;
            redq[2 * r + 1] += car != 0.f ? (a2 - cb[r] * a1) / car : a2;
          }
        }
      }
    }
    __syncthreads();
    if (cur >= 0) {
      double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
      for (int r = tid; r < RU; r += blockDim.x) {
        MT_GADD(m, bs + m.b_qkv_s1 + row0 + r, (double)redq[2 * r]);
        MT_GADD(m, bs + m.b_qkv_s2 + row0 + r, (double)redq[2 * r + 1]);
      }
    }
    if (POS && train) {
      // d relative[c, m] += sum_e R[e, m] * RS(V[c,e]) + RS(B[c]),  RS = sum over the rows (columns) i of diagonal m: direct sums
      // (O(L) each, once per CTA; no prefix differences -> no cancellation)
      const int VP = pl.VP;
      for (int it = tid; it < 2 * D * RL; it += blockDim.x) {
        const int row = dv(it, pl.dRL), mm = it - row * RL;        // row in [0, 2D): q-table rows, then k-table rows
        const bool isk = row >= D;
        const int c = isk ? row - D : row;
        const int dl = mm - (L - 1), lo = dl > 0 ? dl : 0, hi = dl < 0 ? L - 1 + dl : L - 1;
        const float* V = vacc + (size_t)(isk ? DD + D : 0) * VP;
        const float* Bv = V + (size_t)DD * VP + c * VP;
        float acc = 0.f;
        for (int i = lo; i <= hi; ++i) acc += Bv[i];
#pragma unroll
        for (int e = 0; e < D; ++e) {
          const float* Vr = V + (size_t)sym_idx(c, e, D) * VP;
          float rs = 0.f;
          for (int i = lo; i <= hi; ++i) rs += Vr[i];
          acc = fmaf(Rp[(size_t)(isk ? D + e : e) * RP + mm], rs, acc);
        }
        dRs[row * RL + mm] += acc;
      }
      __syncthreads();
    }
    if (POS) for (int k = tid; k < nrel; k += blockDim.x) MT_GADD(m, gr.drelative + k, dRs[k]);
  }
  // ================================================================== B3: dq / dk, dv / d relative, bn_qkv backward sums
  if constexpr (!FUSED) {
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      float* slot = slots + (size_t)(pl.resident ? k : 0) * pl.slot_fl;
      __syncthreads();
      if (grp != cur) {
        if (cur >= 0) {
          double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
          for (int r = tid; r < RU; r += blockDim.x) {
            MT_GADD(m, bs + m.b_qkv_s1 + row0 + r, (double)redq[2 * r]);
            MT_GADD(m, bs + m.b_qkv_s2 + row0 + r, (double)redq[2 * r + 1]);
          }
          __syncthreads();
        }
        group_coefs(grp);
        if (tid < NT * HPU) {
          const int t = tid / HPU, g = hg * HPU + tid % HPU, ch = t * MT_G + g;
          const float f = t == 0 ? 1.f : (t == 1 ? gate_val(p.f_qr) : gate_val(p.f_kr));
          // dy = dZ for every term; sum dZ == 0 exactly (softmax rows), so s1 = 0, s2 = T
          BnBwdCoef c = bwd_coef_from_stats(m, stats, bstats, grp, m.f_sim_sum, m.f_sim_sq, ch, m.inv_n2, f, p.bn_sim_w[ch],
                                            p.bn_sim_rv[ch], 0, m.b_sim_t + ch, true);
          sA[ch] = c.A; sB[ch] = c.Bc; sC[ch] = c.Cc;
        }
        for (int r = tid; r < 2 * RU; r += blockDim.x) redq[r] = 0.f;
        cur = grp;
        __syncthreads();
      }
      if (!pl.resident) stage_unit(slot, b0, false);
      const float* Yn = slot + pl.f_yn; const float* dOs = slot + pl.f_do; const float* Dl = slot + pl.f_dl; const float* Ls = slot + pl.f_ls;
      float* dY = slot + pl.f_dy;
      const int nitems = (m.dbg & 4) ? 0 : SPU * HPU * LB * KS, kpb = LB / KS;
      for (int phase = 0; phase < 2; ++phase)
      for (int base = 0; base < nitems; base += blockDim.x) {
        const int item = base + tid;
        const bool act = item < nitems;
        const int it = act ? item : 0;
        const int i1 = dv(it, pl.dKS), ks = it - i1 * KS, i2 = dv(i1, pl.dLB), t = i1 - i2 * LB, s = dv(i2, pl.dHPU), hl = i2 - s * HPU;
        const int g = hg * HPU + hl;
        const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
        const float* dOb = dOs + ((size_t)s * NOCU + hl * GP * NO) * L;
        const float* lsp = Ls + (s * HPU + hl) * L;
        const float* Dp = Dl + (s * HPU + hl) * L;
        float* dYb = dY + ((size_t)s * RU + hl * 4 * D) * L;
        float* dyg = dyh + ((size_t)(b0 + s) * 2 * C + row0 + hl * 4 * D) * L;
        const float wqk = wsim[g], wqr = POS ? wsim[MT_G + g] : 0.f, wkr = POS ? wsim[2 * MT_G + g] : 0.f;
        const float Aqk = sA[g], Bqk = sB[g], Cqk = sC[g];
        const float Aqr = POS ? sA[MT_G + g] : 0.f, Bqr = POS ? sB[MT_G + g] : 0.f, Cqr = POS ? sC[MT_G + g] : 0.f;
        const float Akr = POS ? sA[2 * MT_G + g] : 0.f, Bkr = POS ? sB[2 * MT_G + g] : 0.f, Ckr = POS ? sC[2 * MT_G + g] : 0.f;
        const int wb0 = ks * kpb, wb1 = wb0 + kpb, KA0 = t + LB - wb0, KB0 = LB - t + wb0;
        typedef typename VecT<RT>::T V;
        if (phase == 0) {
          // ------------------------------------------------------------ thread owns rows i = RT*t + r
          float q[D][RT], dsv[GP][RT], dse[GP][RT], lsr[RT], Dr[RT];
#pragma unroll
          for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * L + RT * t);
#pragma unroll
          for (int c = 0; c < GP; ++c) {
            ldv<RT>(dsv[c], dOb + (c * NO) * L + RT * t);
            if (POS) ldv<RT>(dse[c], dOb + (c * NO + 1) * L + RT * t);
          }
          ldv<RT>(lsr, lsp + RT * t); ldv<RT>(Dr, Dp + RT * t);
          WinA<RT> wq[D]; WinB<RT> wk[D]; WinA<RT> wv[GP];
          if (POS) {
#pragma unroll
            for (int c = 0; c < D; ++c) { wq[c].init(Rp + c * RP, KA0); wk[c].init(Rs + (D + c) * RP, KB0); }
#pragma unroll
            for (int c = 0; c < GP; ++c) wv[c].init(Rp + (GP + c) * RP, KA0);
          }
          float dq[D][RT];
#pragma unroll
          for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < RT; ++r) dq[c][r] = 0.f;
          for (int wb = wb0; wb < wb1; ++wb) {
            float kk[D][RT], vv[GP][RT];
#pragma unroll
            for (int c = 0; c < D; ++c) ldv<RT>(kk[c], Yb + (D + c) * L + RT * wb);
#pragma unroll
            for (int c = 0; c < GP; ++c) ldv<RT>(vv[c], Yb + (GP + c) * L + RT * wb);
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
              for (int w = 0; w < RT; ++w) {
                float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                  sqk = fmaf(q[c][r], kk[c][w], sqk);
                  if (POS) { sqr = fmaf(q[c][r], wq[c].get(r, w), sqr); skr = fmaf(kk[c][w], wk[c].get(r, w), skr); }
                }
                float z = wqk * sqk;
                if (POS) z = fmaf(wqr, sqr, fmaf(wkr, skr, z));
                const float P = ex2(z - lsr[r]);
#pragma unroll
                for (int c = 0; c < GP; ++c) {
                  dP = fmaf(dsv[c][r], vv[c][w], dP);
                  if (POS) dP = fmaf(dse[c][r], wv[c].get(r, w), dP);
                }
                const float dZ = P * (dP - Dr[r]);
                const float dSqk = fmaf(Aqk, dZ, fmaf(Cqk, sqk, Bqk));
                const float dSqr = POS ? fmaf(Aqr, dZ, fmaf(Cqr, sqr, Bqr)) : 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                  dq[c][r] = fmaf(dSqk, kk[c][w], dq[c][r]);
                  if (POS) dq[c][r] = fmaf(dSqr, wq[c].get(r, w), dq[c][r]);
                }
              }
            if (POS && wb + 1 < wb1) {
#pragma unroll
              for (int c = 0; c < D; ++c) { wq[c].advance(Rp + c * RP, KA0 - (wb - wb0) - 1); wk[c].advance(Rs + (D + c) * RP, KB0 + (wb - wb0) + 1); }
#pragma unroll
              for (int c = 0; c < GP; ++c) wv[c].advance(Rp + (GP + c) * RP, KA0 - (wb - wb0) - 1);
            }
          }
          for (int msk = 1; msk < KS; msk <<= 1)
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
              for (int r = 0; r < RT; ++r) dq[c][r] += __shfl_xor_sync(0xffffffffu, dq[c][r], msk);
          if (act)
#pragma unroll
            for (int c = 0; c < D; ++c)
              if ((c & (KS - 1)) == ks) {
                *reinterpret_cast<V*>(dYb + c * L + RT * t) = *reinterpret_cast<V*>(dq[c]);
                *reinterpret_cast<V*>(dyg + (size_t)c * L + RT * t) = *reinterpret_cast<V*>(dq[c]);
              }
        } else {
          // ------------------------------------------------------------ thread owns columns j = RT*t + r
          float kc_[D][RT], vc_[GP][RT];
#pragma unroll
          for (int c = 0; c < D; ++c) ldv<RT>(kc_[c], Yb + (D + c) * L + RT * t);
#pragma unroll
          for (int c = 0; c < GP; ++c) ldv<RT>(vc_[c], Yb + (GP + c) * L + RT * t);
          // index i - j + L - 1 = walk - own : WinB on the shifted copy (Rq, Rv); j - i + L - 1 = own - walk : WinA (Rk)
          WinB<RT> wq[D]; WinA<RT> wk[D]; WinB<RT> wv[GP];
          if (POS) {
#pragma unroll
            for (int c = 0; c < D; ++c) { wq[c].init(Rs + c * RP, KB0); wk[c].init(Rp + (D + c) * RP, KA0); }
#pragma unroll
            for (int c = 0; c < GP; ++c) wv[c].init(Rs + (GP + c) * RP, KB0);
          }
          float dk[D][RT], dvv[GP][RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) {
#pragma unroll
            for (int c = 0; c < D; ++c) dk[c][r] = 0.f;
#pragma unroll
            for (int c = 0; c < GP; ++c) dvv[c][r] = 0.f;
          }
          for (int wb = wb0; wb < wb1; ++wb) {
            float qq[D][RT], osv[GP][RT], ose[GP][RT], lsw[RT], Dw[RT];
#pragma unroll
            for (int c = 0; c < D; ++c) ldv<RT>(qq[c], Yb + c * L + RT * wb);
#pragma unroll
            for (int c = 0; c < GP; ++c) {
              ldv<RT>(osv[c], dOb + (c * NO) * L + RT * wb);
              if (POS) ldv<RT>(ose[c], dOb + (c * NO + 1) * L + RT * wb);
            }
            ldv<RT>(lsw, lsp + RT * wb); ldv<RT>(Dw, Dp + RT * wb);
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
              for (int w = 0; w < RT; ++w) {
                float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                  sqk = fmaf(qq[c][w], kc_[c][r], sqk);
                  if (POS) { sqr = fmaf(qq[c][w], wq[c].get(r, w), sqr); skr = fmaf(kc_[c][r], wk[c].get(r, w), skr); }
                }
                float z = wqk * sqk;
                if (POS) z = fmaf(wqr, sqr, fmaf(wkr, skr, z));
                const float P = ex2(z - lsw[w]);
#pragma unroll
                for (int c = 0; c < GP; ++c) {
                  dP = fmaf(osv[c][w], vc_[c][r], dP);
                  if (POS) dP = fmaf(ose[c][w], wv[c].get(r, w), dP);
                }
                const float dZ = P * (dP - Dw[w]);
                const float dSqk = fmaf(Aqk, dZ, fmaf(Cqk, sqk, Bqk));
                const float dSkr = POS ? fmaf(Akr, dZ, fmaf(Ckr, skr, Bkr)) : 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                  dk[c][r] = fmaf(dSqk, qq[c][w], dk[c][r]);
                  if (POS) dk[c][r] = fmaf(dSkr, wk[c].get(r, w), dk[c][r]);
                }
#pragma unroll
                for (int c = 0; c < GP; ++c) dvv[c][r] = fmaf(P, osv[c][w], dvv[c][r]);
              }
            if (POS && wb + 1 < wb1) {
#pragma unroll
              for (int c = 0; c < D; ++c) { wq[c].advance(Rs + c * RP, KB0 + (wb - wb0) + 1); wk[c].advance(Rp + (D + c) * RP, KA0 - (wb - wb0) - 1); }
#pragma unroll
              for (int c = 0; c < GP; ++c) wv[c].advance(Rs + (GP + c) * RP, KB0 + (wb - wb0) + 1);
            }
          }
          for (int msk = 1; msk < KS; msk <<= 1) {
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
              for (int r = 0; r < RT; ++r) dk[c][r] += __shfl_xor_sync(0xffffffffu, dk[c][r], msk);
#pragma unroll
            for (int c = 0; c < GP; ++c)
#pragma unroll
              for (int r = 0; r < RT; ++r) dvv[c][r] += __shfl_xor_sync(0xffffffffu, dvv[c][r], msk);
          }
          if (act) {
#pragma unroll
            for (int c = 0; c < D; ++c)
              if ((c & (KS - 1)) == ks) {
                *reinterpret_cast<V*>(dYb + (D + c) * L + RT * t) = *reinterpret_cast<V*>(dk[c]);
                *reinterpret_cast<V*>(dyg + (size_t)(D + c) * L + RT * t) = *reinterpret_cast<V*>(dk[c]);
              }
#pragma unroll
            for (int c = 0; c < GP; ++c)
              if ((c & (KS - 1)) == ks) {
                *reinterpret_cast<V*>(dYb + (GP + c) * L + RT * t) = *reinterpret_cast<V*>(dvv[c]);
                *reinterpret_cast<V*>(dyg + (size_t)(GP + c) * L + RT * t) = *reinterpret_cast<V*>(dvv[c]);
              }
          }
        }
      }
      // ------------------------------------------------------------ diagonals (d relative): one thread per diagonal
      if (POS && D <= 2 && !(m.dbg & 4) && !(m.dbg & 128)) {
        // register-tiled pass over WRAPPED diagonals (axial_diag_tile.cuh): balanced items, ~5 shared-memory loads per
        // RT x RT pairs instead of ~20 scalar loads per pair (ncu: the scalar walk below was 38 % of this kernel's instructions)
        const int nit = SPU * HPU * LB;
        for (int item = tid; item < nit; item += blockDim.x) {
          const int i2 = dv(item, pl.dLB), t = item - i2 * LB, s = dv(i2, pl.dHPU), hl = i2 - s * HPU, g = hg * HPU + hl;
          mtdiag::Coefs cf;
          cf.wqk = wsim[g]; cf.wqr = wsim[MT_G + g]; cf.wkr = wsim[2 * MT_G + g];
          cf.Aqr = sA[MT_G + g]; cf.Bqr = sB[MT_G + g]; cf.Cqr = sC[MT_G + g];
          cf.Akr = sA[2 * MT_G + g]; cf.Bkr = sB[2 * MT_G + g]; cf.Ckr = sC[2 * MT_G + g];
          mtdiag::diag_tile_item<(D <= 2 ? D : 1), RT>(L, L, RP, t, Yn + ((size_t)s * RU + hl * 4 * D) * L,
                                                       dOs + ((size_t)s * NOCU + hl * GP * NO) * L, Ls + (s * HPU + hl) * L,
                                                       Dl + (s * HPU + hl) * L, Rp, cf, dRs);
        }
      } else if (POS && !(m.dbg & 4)) {
        const int nd = SPU * HPU * RL;
        for (int item = tid; item < nd; item += blockDim.x) {
          const int i2 = dv(item, pl.dRL), mm = item - i2 * RL, s = dv(i2, pl.dHPU), hl = i2 - s * HPU, g = hg * HPU + hl;
          const float* Yb = Yn + ((size_t)s * RU + hl * 4 * D) * L;
          const float* dOb = dOs + ((size_t)s * NOCU + hl * GP * NO) * L;
          const float wqk = wsim[g], wqr = wsim[MT_G + g], wkr = wsim[2 * MT_G + g];
          const float Aqr = sA[MT_G + g], Bqr = sB[MT_G + g], Cqr = sC[MT_G + g];
          const float Akr = sA[2 * MT_G + g], Bkr = sB[2 * MT_G + g], Ckr = sC[2 * MT_G + g];
          const float* lsp = Ls + (s * HPU + hl) * L;
          const float* Dp = Dl + (s * HPU + hl) * L;
          float rq[D], rk[D], rv[GP], aq[D], ak[D], av[GP];
#pragma unroll
          for (int c = 0; c < D; ++c) { rq[c] = Rp[c * RP + mm]; rk[c] = Rp[(D + c) * RP + (2 * (L - 1) - mm)]; aq[c] = 0.f; ak[c] = 0.f; }
#pragma unroll
          for (int c = 0; c < GP; ++c) { rv[c] = Rp[(GP + c) * RP + mm]; av[c] = 0.f; }
          const int ilo = mm - (L - 1) > 0 ? mm - (L - 1) : 0, ihi = mm < L - 1 ? mm : L - 1;
          for (int i = ilo; i <= ihi; ++i) {
            const int j = i - mm + L - 1;
            float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
              const float qc = Yb[c * L + i], kc = Yb[(D + c) * L + j];
              sqk = fmaf(qc, kc, sqk); sqr = fmaf(qc, rq[c], sqr); skr = fmaf(kc, rk[c], skr);
            }
            const float P = ex2(fmaf(wqr, sqr, fmaf(wkr, skr, wqk * sqk)) - lsp[i]);
#pragma unroll
            for (int c = 0; c < GP; ++c) dP = fmaf(dOb[(2 * c) * L + i], Yb[(GP + c) * L + j], fmaf(dOb[(2 * c + 1) * L + i], rv[c], dP));
            const float dZ = P * (dP - Dp[i]);
            const float dSqr = fmaf(Aqr, dZ, fmaf(Cqr, sqr, Bqr)), dSkr = fmaf(Akr, dZ, fmaf(Ckr, skr, Bkr));
#pragma unroll
            for (int c = 0; c < D; ++c) { aq[c] = fmaf(dSqr, Yb[c * L + i], aq[c]); ak[c] = fmaf(dSkr, Yb[(D + c) * L + j], ak[c]); }
#pragma unroll
            for (int c = 0; c < GP; ++c) av[c] = fmaf(P, dOb[(2 * c + 1) * L + i], av[c]);
          }
#pragma unroll
          for (int c = 0; c < D; ++c) { atomicAdd(dRs + c * RL + mm, aq[c]); atomicAdd(dRs + (D + c) * RL + (2 * (L - 1) - mm), ak[c]); }
#pragma unroll
          for (int c = 0; c < GP; ++c) atomicAdd(dRs + (GP + c) * RL + mm, av[c]);
        }
      }
      __syncthreads();
      // bn_qkv backward sums: s1 = sum dy, s2 = sum dy * y_raw (the raw projection comes back from L2)
      {
        const int tpr = pl.tpr_q, sub = tid & (tpr - 1), rpp = blockDim.x / tpr;
        for (int r0 = 0; r0 < RU; r0 += rpp) {
          const int r = r0 + tid / tpr;
          float a1 = 0.f, a2 = 0.f;
          if (r < RU)
            for (int s = 0; s < SPU; ++s) {
              const float* dy = dY + ((size_t)s * RU + r) * L;
              const float* yr = yraw + ((size_t)(b0 + s) * 2 * C + row0 + r) * L;
              for (int l = sub; l < L; l += tpr) { const float d = dy[l]; a1 += d; a2 = fmaf(d, __ldg(yr + l), a2); }
            }
          a1 = seg_sum(a1, tpr); a2 = seg_sum(a2, tpr);
          if (sub == 0 && r < RU) { redq[2 * r] += a1; redq[2 * r + 1] += a2; }
        }
      }
    }
    __syncthreads();
    if (cur >= 0) {
      double* bs = stat_slot(m, bstats, cur, m.nstat_b, cta);
      for (int r = tid; r < RU; r += blockDim.x) {
        MT_GADD(m, bs + m.b_qkv_s1 + row0 + r, (double)redq[2 * r]);
        MT_GADD(m, bs + m.b_qkv_s2 + row0 + r, (double)redq[2 * r + 1]);
      }
    }
    if (POS) for (int k = tid; k < nrel; k += blockDim.x) MT_GADD(m, gr.drelative + k, dRs[k]);
  }
  phase_stamp(m, 5);
  grid_barrier(bar, G, epoch);
  pdl_launch_dependents();          // last phase: a PDL-launched successor (pdl.cuh) may take its SM slots now and wait there
  phase_stamp(m, 6);

  // ================================================================== B4: bn_qkv backward + dX = W^T dy_raw
  {
    float* qA = sm + pl.o_qA; float* qB = sm + pl.o_qB; float* qC = sm + pl.o_qC;      // [2C]
    float* W4 = sm + pl.o_w4;        // [2C][WTP]: W[o][c4 .. c4+CU4)
    float* dyR = sm + pl.o_dyr;      // [2C][XP]
    float* dXs = sm + pl.o_dxs;      // [CU4][XP]
    const int CU4 = pl.CU4, c4 = hg * CU4, K = 2 * C;
    __syncthreads();
    for (int idx = tid; idx < K * CU4; idx += blockDim.x) {
      const int o = idx / CU4, c = idx - o * CU4;
      cp_async4(W4 + o * pl.WTP + c, p.w_qkv + (size_t)o * C + c4 + c);
    }
    int cur = -1;
    for (int k = 0; k < pl.upc; ++k) {
      const int u = pl.contig ? cta * pl.upc + k : cta + k * G;
      if (u >= pl.nunits) break;
      const int b0 = (u / HS) * SPU, grp = b0 / m.Bg;
      __syncthreads();
      if (grp != cur) {
        const bool owner = (b0 == grp * m.Bg) && hg == 0;
        for (int ch = tid; ch < K; ch += blockDim.x) {
          BnBwdCoef c = bwd_coef_from_stats(m, stats, bstats, grp, m.f_qkv_sum, m.f_qkv_sq, ch, m.inv_n1, 1.f, p.bn_qkv_w[ch],
                                            p.bn_qkv_rv[ch], m.b_qkv_s1 + ch, m.b_qkv_s2 + ch, false);
          qA[ch] = c.A; qB[ch] = c.Bc; qC[ch] = c.Cc;
          if (owner) { float* o = bcoef + (size_t)grp * m.ncoef_b + m.cb_q + 3 * ch; o[0] = c.A; o[1] = c.Bc; o[2] = c.Cc; }
        }
        cur = grp;
        __syncthreads();
      }
      // dy_raw[o][s*L + l] = A dy + Bc + Cc y_raw for ALL 2C rows of the unit's strips
      if ((L & 3) == 0) {
        const int L4 = L >> 2;
        for (int i4 = tid; i4 < (SPU * K * L) >> 2; i4 += blockDim.x) {
          const int row = dv(i4, pl.dL4), l4 = i4 - row * L4, s = dv(row, pl.dK2), o = row - s * K;
          const size_t a = ((size_t)(b0 + s) * K + o) * L + 4 * l4;
          const float4 d = __ldcg(reinterpret_cast<const float4*>(dyh + a)), y = __ldg(reinterpret_cast<const float4*>(yraw + a));
          const float A = qA[o], Bc = qB[o], Cc = qC[o];
          float4 r;
          r.x = fmaf(A, d.x, fmaf(Cc, y.x, Bc)); r.y = fmaf(A, d.y, fmaf(Cc, y.y, Bc));
          r.z = fmaf(A, d.z, fmaf(Cc, y.z, Bc)); r.w = fmaf(A, d.w, fmaf(Cc, y.w, Bc));
          *reinterpret_cast<float4*>(dyR + (size_t)o * XP + s * L + 4 * l4) = r;
        }
      } else {
        for (int i = tid; i < SPU * K * L; i += blockDim.x) {
          const int row = i / L, l = i - row * L, s = row / K, o = row - s * K;
          const size_t a = ((size_t)(b0 + s) * K + o) * L + l;
          dyR[(size_t)o * XP + s * L + l] = fmaf(qA[o], __ldcg(dyh + a), fmaf(qC[o], __ldg(yraw + a), qB[o]));
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (pl.tm == 4 && pl.tn == 4) gemm_smem<4, 4>(W4, pl.WTP, dyR, XP, K, CU4, dXs);
      else if (pl.tm == 2 && pl.tn == 4) gemm_smem<2, 4>(W4, pl.WTP, dyR, XP, K, CU4, dXs);
      else if (pl.tm == 2 && pl.tn == 2) gemm_smem<2, 2>(W4, pl.WTP, dyR, XP, K, CU4, dXs);
      else gemm_smem<2, 1>(W4, pl.WTP, dyR, XP, K, CU4, dXs);
      __syncthreads();
      if (gr.dx) {
        const int n = b0 / m.other, o0 = b0 - n * m.other;
        const size_t plane = (size_t)m.H * m.W;
        if (m.axis == 1) {
          float* xb = gr.dx + (((size_t)n * C + c4) * m.H + o0) * m.W;                  // + c*plane + s*W + l
          if ((L & 3) == 0) {
            for (int i4 = tid; i4 < (CU4 * XP) >> 2; i4 += blockDim.x) {
              const int row = dv(i4, pl.dL4), l4 = i4 - row * pl.dL4.d, c = dv(row, pl.dSPU), s = row - c * SPU;
              *reinterpret_cast<float4*>(xb + (size_t)c * plane + (size_t)s * m.W + 4 * l4) =
                  *reinterpret_cast<const float4*>(dXs + (size_t)c * XP + s * L + 4 * l4);
            }
          } else {
            for (int i = tid; i < CU4 * XP; i += blockDim.x) {
              const int row = i / L, l = i - row * L, c = row / SPU, s = row - c * SPU;
              xb[(size_t)c * plane + (size_t)s * m.W + l] = dXs[(size_t)c * XP + s * L + l];
            }
          }
        } else {
          float* xb = gr.dx + ((size_t)n * C + c4) * plane + o0;                        // + c*plane + l*W + s
          for (int i = tid; i < CU4 * XP; i += blockDim.x) {
            const int rem = dv(i, pl.dSPU), s = i - rem * SPU, c = dv(rem, pl.dL), l = rem - c * L;
            xb[(size_t)c * plane + (size_t)l * m.W + s] = dXs[(size_t)c * XP + s * L + l];
          }
        }
      }
    }
  }
  phase_stamp(m, 7);
  // every BN / gate parameter gradient, one channel per thread across the WHOLE grid (all sums are final since barrier 3);
  // the CTAs take the items from the back so that the owners of the first units are not the ones delayed
  bwd_param_grads_dist(m, p, gr, stats, bstats, (G - 1 - cta) * blockDim.x + tid, G * blockDim.x);
  phase_stamp(m, 8);
}

// developer hook: phase stamps of the last cooperative backward launch
extern "C" int mt_debug_read_phases_bwd(unsigned long long* host, int n) {
  if (!host || n <= 0) return 1;
  const size_t cnt = (size_t)n < (size_t)MT_PH_MAX * MT_PH_CTAS ? (size_t)n : (size_t)MT_PH_MAX * MT_PH_CTAS;
  return cudaMemcpyFromSymbol(host, g_phase_ns, cnt * sizeof(unsigned long long)) == cudaSuccess ? 0 : 3;
}

// =============================================================================================== host side
int coop_rt(const Dims& m);
// RT of the backward kernel: d = 1 runs RT = 2 (113 registers, two CTAs per SM; measured 13 % faster than RT = 4 at one
// CTA per SM on the L = 64 layers); MT_BWD_RT=4 restores the forward's choice for A/B runs
static int coop_rt_bwd(const Dims& m) {
  int rt = coop_rt(m);
  static const int force = getenv("MT_BWD_RT") ? atoi(getenv("MT_BWD_RT")) : 0;
  if (m.d == 1 && rt == 4 && force != 4) rt = 2;
  return rt;
}
int coop_force_stream();
int coop_env_int(const char* name);
int coop_rows_tpr(int nrows, int L);

// single-sweep path: RT of the wrapped-diagonal tile.  The NL = min(L/RT, 32) lanes of a (strip, head) must be a power of two and
// at least D*RT (transposing row reduction); the largest RT that still fills the CTA's 256 threads wins.
static int coop_rt_fused(const Dims& m, int groups_per_cta) {
  static const int off = coop_env_int("MT_BWD_NOFUSE"), force = coop_env_int("MT_BWD_FUSED_RT"), all = coop_env_int("MT_BWD_FUSED_ALL");
  if (off > 0 || m.d > 4 || (m.L & (m.L - 1)) != 0) return 0;
  // measured (profiles/r2_notes.md): the single sweep wins where the L^2 work dominates -- d = 1 from L = 64, d = 2 at L = 64;
  // short strips (the combine phase costs more than the sweeps it removes), d = 4 (1 x 1 tiles) and d = 2 at L >= 128 (250
  // registers, one CTA per SM, four lanes groups per CTA) stay on the multi-sweep path.  MT_BWD_FUSED_ALL=1 lifts the policy.
  if (!all && !force && !((m.d == 1 && m.L >= 64) || (m.d == 2 && m.L == 64))) return 0;
  const int cand[3] = {4, 2, 1};
  int best = 0;
  bool filled = false;
  for (int k = 0; k < 3; ++k) {
    const int rt = cand[k];
    if (rt * m.d > 4 || m.L % rt) continue;
    const int lb = m.L / rt, nl = lb < 32 ? lb : 32;
    if (nl < m.d * rt || nl < 2) continue;
    if (force == rt) return rt;
    if (filled) continue;
    best = rt;                                       // smaller RT = more lanes
    if (groups_per_cta * nl >= MT_COOP_THREADS) filled = true;
  }
  return best;
}

static bool coop_plan_bwd_hs(const Dims& m, int maxctas, size_t smem_cap, CoopPlanB& pl, int hs_min, int want_fused) {
  const int C = m.C, L = m.L, D = m.d, GP = 2 * D, NO = m.no;
  pl.RT = coop_rt_bwd(m);
  pl.LB = L / pl.RT;
  pl.SPU = m.stride;
  if (m.other % pl.SPU) return false;
  const int nsb = m.B / pl.SPU;
  int hs = 1;
  while (hs < MT_G && nsb * hs * 2 <= maxctas) hs *= 2;
  if (hs < hs_min) hs = hs_min;
  static const int force_hs = coop_env_int("MT_COOP_HS");
  if (force_hs == 1 || force_hs == 2 || force_hs == 4 || force_hs == 8) hs = force_hs;
  pl.HS = hs; pl.HPU = MT_G / hs;
  pl.RU = pl.HPU * 4 * D; pl.CU = pl.HPU * GP; pl.NOCU = pl.CU * NO; pl.CU4 = C / hs;
  pl.XP = pl.SPU * L;
  pl.nunits = nsb * hs;
  int G = pl.nunits < maxctas ? pl.nunits : maxctas;
  G -= G % hs;
  if (G < hs) return false;
  pl.upc = (pl.nunits + G - 1) / G;
  pl.contig = (hs == 1 && pl.upc > 1) ? 1 : 0;          // consecutive strips share a statistic group: fewer coefficient reloads
  if (pl.contig) G = (pl.nunits + pl.upc - 1) / pl.upc;
  pl.G = G;
  int ks = 1;
  const int items0 = pl.SPU * pl.HPU * pl.LB;
  while (ks < 8 && items0 * ks < MT_COOP_THREADS && (pl.LB % (ks * 2)) == 0) ks *= 2;
  static const int force_ks = coop_env_int("MT_COOP_KS");
  if (force_ks > 0 && (pl.LB % force_ks) == 0 && (force_ks & (force_ks - 1)) == 0 && force_ks <= 32) ks = force_ks;
  // the T pass sums lanes of one head with a segmented shuffle: LB*KS lanes must be a power of two (or >= 32 and a multiple)
  { const int seg = pl.LB * ks; if (seg < 32 ? (seg & (seg - 1)) != 0 : (seg % 32) != 0) return false; }
  pl.KS = ks;
  const int tn_max = (L % 4 == 0) ? 4 : (L % 2 == 0 ? 2 : 1);
  if (tn_max == 4 && pl.CU4 % 4 == 0 && (pl.CU4 / 4) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 4; pl.tn = 4; }
  else if (tn_max == 4 && (pl.CU4 / 2) * (pl.XP / 4) >= MT_COOP_THREADS) { pl.tm = 2; pl.tn = 4; }
  else if (tn_max >= 2) { pl.tm = 2; pl.tn = 2; }
  else { pl.tm = 2; pl.tn = 1; }
  pl.WTP = pl.CU4 + 4;
  pl.tpr_o = coop_rows_tpr(pl.CU, L);
  pl.tpr_q = coop_rows_tpr(pl.RU, L);
  { int t = pl.tpr_o; while (t > 1 && t > L) t >>= 1; pl.tpr_o = t; t = pl.tpr_q; while (t > 1 && t > L) t >>= 1; pl.tpr_q = t; }
  pl.dL = make_divs(L); pl.dL4 = make_divs(L % 4 == 0 ? L / 4 : 1); pl.dLB = make_divs(pl.LB); pl.dKS = make_divs(pl.KS);
  pl.dHPU = make_divs(pl.HPU); pl.dSPU = make_divs(pl.SPU); pl.dRU = make_divs(pl.RU); pl.dNOCU = make_divs(pl.NOCU);
  pl.dCU = make_divs(pl.CU); pl.dRL = make_divs(2 * L - 1); pl.dK2 = make_divs(2 * C);
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
  pl.o_ca = take(pl.RU); pl.o_cb = take(pl.RU);
  pl.o_wsim = take(24); pl.o_sA = take(24); pl.o_sB = take(24); pl.o_sC = take(24);
  pl.o_oA = take(pl.NOCU); pl.o_oB = take(pl.NOCU); pl.o_oC = take(pl.NOCU);
  pl.o_redo = take(pl.CU + pl.NOCU); pl.o_reds = take(24); pl.o_redq = take(2 * pl.RU);
  pl.o_rp = take(m.pos ? 2 * m.gp * rel_pitch(L) : 0);
  pl.o_rs = take(m.pos ? 2 * m.gp * rel_pitch(L) : 0);
  pl.o_drel = take(m.pos ? 2 * m.gp * (2 * L - 1) : 0);
  pl.fused = 0; pl.NLg = 1; pl.NSEG = 1; pl.VP = L + 1; pl.o_wtab = pl.o_gram = pl.o_vacc = off; pl.dNL = make_divs(1);
  if (want_fused) {
    const int rt = coop_rt_fused(m, pl.SPU * pl.HPU);
    if (!rt) return false;
    const int DD = D * (D + 1) / 2;
    pl.fused = 1; pl.RT = rt; pl.LB = L / rt;
    pl.NLg = pl.LB < 32 ? pl.LB : 32; pl.NSEG = pl.LB / pl.NLg; pl.dNL = make_divs(pl.NLg); pl.dLB = make_divs(pl.LB);
    pl.o_wtab = take((m.pos && m.training) ? 2 * (D + DD) * L : 0);
    pl.o_gram = take(pl.SPU * pl.HPU * (2 * DD + 2 * D));
    pl.o_vacc = take((m.pos && m.training) ? 2 * (DD + D) * pl.VP : 0);
  }
  pl.o_slot = off;
  int so = 0;
  auto stake = [&](int n) { int o = so; so += (n + 3) & ~3; return o; };
  pl.f_yn = stake(pl.SPU * pl.RU * L); pl.f_do = stake(pl.SPU * pl.NOCU * L); pl.f_gs = stake(pl.SPU * pl.CU * L);
  pl.f_dl = stake(pl.SPU * pl.HPU * L); pl.f_ls = stake(pl.SPU * pl.HPU * L); pl.f_dy = stake(pl.SPU * pl.RU * L);
  pl.slot_fl = so;
  // B4 staging aliases the slots (dead after barrier 3)
  int bo = pl.o_slot;
  auto btake = [&](int n) { int o = bo; bo += (n + 3) & ~3; return o; };
  pl.o_qA = btake(2 * C); pl.o_qB = btake(2 * C); pl.o_qC = btake(2 * C);
  pl.o_w4 = btake(2 * C * pl.WTP); pl.o_dyr = btake(2 * C * pl.XP); pl.o_dxs = btake(pl.CU4 * pl.XP);
  const size_t fixed = (size_t)pl.o_slot * sizeof(float), slot = (size_t)pl.slot_fl * sizeof(float);
  const size_t b4 = (size_t)(bo - pl.o_slot) * sizeof(float);
  if (fixed + (slot > b4 ? slot : b4) > smem_cap) return false;
  int nslot = (int)((smem_cap - fixed) / slot);
  if (nslot > pl.upc) nslot = pl.upc;
  pl.resident = nslot >= pl.upc && !coop_force_stream();
  pl.nslot = pl.resident ? pl.upc : 1;
  const size_t slots_b = (size_t)pl.nslot * slot;
  pl.smem_bytes = fixed + (slots_b > b4 ? slots_b : b4);
  return true;
}

static bool coop_plan_bwd(const Dims& m, int maxctas, size_t smem_cap, CoopPlanB& pl, int want_fused) {
  for (int hs = 1; hs <= MT_G; hs *= 2)
    if (coop_plan_bwd_hs(m, maxctas, smem_cap, pl, hs, want_fused)) return true;
  return false;
}

template <int D, bool POS, int RT, bool FUSED>
static int coop_launch_bwd(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const float* x, const float* dout,
                           const double* stats, const float* coef, const float* yraw, const float* oraw, const float* lse,
                           double* bstats, float* bcoef, unsigned* bcnt, float* dyh, cudaStream_t st) {
  auto kern = k_axial_bwd<D, POS, RT, FUSED>;
  static bool attr_set[64] = {};          // per device: function attributes live in a context (nn.DataParallel: one process, two GPUs)
  int adev = 0;
  if (cudaGetDevice(&adev) != cudaSuccess || adev < 0 || adev >= 64) return -1;
  if (!attr_set[adev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) return -1;
    attr_set[adev] = true;
  }
  const int nsm = coop_sm_count();
  const int cap = coop_max_ctas_env();
  CoopPlanB pl;
  bool ok = false;
  for (int occ = 2; occ >= 1 && !ok; --occ) {
    int maxctas = nsm * occ;
    if (cap > 0 && maxctas > cap) maxctas = cap;
    const size_t smem_cap = occ == 2 ? 110 * 1024 : 224 * 1024;
    if (!coop_plan_bwd(m, maxctas, smem_cap, pl, FUSED ? 1 : 0)) continue;
    if (FUSED && pl.RT != RT) continue;
    if (occ == 2) {
      if (pl.G <= nsm) { ok = true; break; }
      static thread_local size_t q_smem[8] = {0};
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
  if (!ok) return 0;              // not covered (e.g. one unit does not fit): the caller falls back to the launch chain
  Dims mm = m; MtAxialParams pp = p; MtAxialGrads gg = gr;
  MT_COOP_LAUNCH((kern), pl.G, MT_COOP_THREADS, pl.smem_bytes, st, mm, pp, gg, pl, x, dout, stats, coef, yraw, oraw, lse, bstats,
                 bcoef, bcnt, dyh);
  return g_last_cuda == cudaSuccess ? 1 : -1;
}

// returns launches done (1), 0 when the shape is not covered, < 0 on error
int launch_coop_bwd(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const float* x, const float* dout,
                    const double* stats, const float* coef, const float* yraw, const float* oraw, const float* lse,
                    double* bstats, float* bcoef, unsigned* bcnt, float* dyh, cudaStream_t st) {
#define GOF(DD, RR)                                                                                                          \
  { const int rc_ = m.pos ? coop_launch_bwd<DD, true, RR, true>(m, p, gr, x, dout, stats, coef, yraw, oraw, lse, bstats, bcoef, bcnt, dyh, st)   \
                          : coop_launch_bwd<DD, false, RR, true>(m, p, gr, x, dout, stats, coef, yraw, oraw, lse, bstats, bcoef, bcnt, dyh, st); \
    if (rc_ != 0) return rc_; }
  // single-sweep path first (every RT the planner may pick for this d); the multi-sweep kernel below covers the rest
  if (m.d == 1) { GOF(1, 4); GOF(1, 2); GOF(1, 1); }
  if (m.d == 2) { GOF(2, 2); GOF(2, 1); }
  if (m.d == 4) { GOF(4, 1); }
#undef GOF
  const int rt = coop_rt_bwd(m);
#define GO(DD, RR)                                                                                                          \
  return m.pos ? coop_launch_bwd<DD, true, RR, false>(m, p, gr, x, dout, stats, coef, yraw, oraw, lse, bstats, bcoef, bcnt, dyh, st)   \
               : coop_launch_bwd<DD, false, RR, false>(m, p, gr, x, dout, stats, coef, yraw, oraw, lse, bstats, bcoef, bcnt, dyh, st)
  if (m.d == 1) { if (rt == 4) { GO(1, 4); } if (rt == 2) { GO(1, 2); } GO(1, 1); }
  if (m.d == 2) { if (rt == 2) { GO(2, 2); } GO(2, 1); }
  if (m.d == 4) { if (rt == 2) { GO(4, 2); } GO(4, 1); }
  if (m.d == 8) { GO(8, 1); }
#undef GO
  return 0;
}
