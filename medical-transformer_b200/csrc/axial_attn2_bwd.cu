// axial_attn2_bwd.cu -- register-tiled backward attention passes (layout of axial_attn2.cuh):
//   MODE 0  T[t,g] = sum dZ * S_t                       (bn_similarity backward needs it first)
//   MODE 1  gradients; blockIdx.y picks the pass, all CTAs of one strip group run concurrently:
//     y=0    rows:    dq[c,i] = sum_j dS_qk k[c,j] + dS_qr Rq[c,i-j]                                   (thread owns RT rows)
//            then, on the same staging,
//            columns: dk[c,j] = sum_i dS_qk q[c,i] + dS_kr Rk[c,j-i] ; dv[c,j] = sum_i P dO_sv[c,i]     (thread owns RT columns)
//     y=1,2  diagonals: d relative[c, i-j+L-1], one thread per diagonal (scalar walk), half of the diagonals each
//   (experimental layouts behind MT_DEBUG_FLAGS bits 3 / 4: see the XP template parameter of k_bwd2)
// P is recomputed from the saved log-sum-exp, dZ = P (dP - D).  Closed forms: oracle/axial_np.py.
#include "axial_attn2.cuh"
#include "axial_diag_tile.cuh"
extern thread_local cudaError_t g_last_cuda;

struct Bwd2Smem {
  float *ca, *cb, *wsim, *sA, *sB, *sC, *oA, *oB, *oC, *red, *Yn, *Rp, *Rs, *dOs, *Dl, *lses;
};
__device__ __forceinline__ Bwd2Smem carve2(const Dims& m, float* sm, int SPB, int nred) {
  Bwd2Smem s;
  const int C = m.C, NOC = m.no * m.C, RP = rel_pitch(m.L);
  s.ca = sm; s.cb = s.ca + 2 * C;
  s.wsim = s.cb + 2 * C; s.sA = s.wsim + 24; s.sB = s.sA + 24; s.sC = s.sB + 24;
  s.oA = s.sC + 24; s.oB = s.oA + NOC; s.oC = s.oB + NOC;
  s.red = s.oC + NOC;
  s.Yn = s.red + ((nred + 3) & ~3);
  s.Rp = s.Yn + (size_t)SPB * 2 * C * m.LP;
  s.Rs = s.Rp + (m.pos ? 2 * m.gp * RP : 0);
  s.dOs = s.Rs + (m.pos ? 2 * m.gp * RP : 0);
  s.Dl = s.dOs + (size_t)SPB * NOC * m.LP;
  s.lses = s.Dl + (size_t)SPB * MT_G * m.L;
  return s;
}
static size_t bwd2_smem_floats(const Dims& m, int SPB, int nred) {
  const size_t NOC = (size_t)m.no * m.C;
  return 4 * (size_t)m.C + 96 + 3 * NOC + ((nred + 3) & ~3) + (size_t)SPB * 2 * m.C * m.LP +
         (m.pos ? 4 * (size_t)m.gp * rel_pitch(m.L) : 0) + (size_t)SPB * NOC * m.LP + 2 * (size_t)SPB * MT_G * m.L;
}

// XP: experimental variants of MODE 1, selected at run time by MT_DEBUG_FLAGS bits 3 / 4 (XP = 0 is the measured default
// and compiles to exactly the kernel profiled in profiles/):  bit 0 = register-tiled wrapped-diagonal pass for d relative
// (axial_diag_tile.cuh, one diagonal CTA instead of two), bit 1 = rows and columns in separate CTAs (y = 0 / y = 1).
template <int D, bool POS, int RT, int MODE, int XP = 0>
__global__ void __launch_bounds__(128)
k_bwd2(Dims m, MtAxialParams p, MtAxialGrads gr, const double* __restrict__ stats, double* __restrict__ bstats,
       const float* __restrict__ coef, float* __restrict__ bcoef, unsigned* __restrict__ bcnt,
       const float* __restrict__ yraw, const float* __restrict__ oraw, const float* __restrict__ lse,
       const float* __restrict__ gs, float* __restrict__ dyh, int SPB, int segw) {
  constexpr int GP = 2 * D, NO = POS ? 2 : 1;
  // d relative: concurrent diagonal CTAs (blockIdx.y >= NMAIN).  Accumulating it inside the row pass instead (mirror window
  // accumulators flushed with shared-memory atomics) was measured 2.4 % slower at 128 px and 10 % at 512 px and removed
  // (profiles/r1_notes.md item 10).
  extern __shared__ __align__(16) float sm[];
  const int C = m.C, L = m.L, LP = m.LP, NOC = NO * C, RP = rel_pitch(L), LB = L / RT;
  const int b0 = blockIdx.x * SPB, grp = b0 / m.Bg;
  // MODE 1: CTA y = 0 does the row pass then the column pass on ONE staging of the strips (the kernel is
  // L2-staging bound, profiles/r1_notes.md); y >= 1 are the diagonal-walk CTAs (d relative), split nds ways
  constexpr bool TDIAG = (XP & 1) && POS && MODE == 1 && D <= 2;
  constexpr int NMAIN = ((XP & 2) && MODE == 1) ? 2 : 1;
  const bool diag_cta = MODE == 1 && blockIdx.y >= NMAIN;
  const int dsplit = (int)blockIdx.y - NMAIN, nds = (int)gridDim.y - NMAIN;
  const int nrel = POS ? 2 * GP * (2 * L - 1) : 0;
  const int nred = MODE == 0 ? 24 : 4 * C + nrel;
  Bwd2Smem s = carve2(m, sm, SPB, nred);
  float* dRs = s.red + 4 * C;
  __shared__ __align__(8) uint64_t tma_bar;
  const bool use_tma = (LP == L);
  if (use_tma && threadIdx.x == 0)
    tma_strips_issue(&tma_bar, s.Yn, yraw + (size_t)b0 * 2 * C * L, (uint32_t)(SPB * 2 * C * L * sizeof(float)));
  for (int k = threadIdx.x; k < nred; k += blockDim.x) s.red[k] = 0.f;
  // ---- staging (coefficient tables -> strips, dO, lse -> D), batched loads
  {
    const float* tq = coef + (size_t)grp * m.ncoef_f + m.cf_q;
    for (int ch = threadIdx.x; ch < 2 * C; ch += blockDim.x) { s.ca[ch] = tq[2 * ch]; s.cb[ch] = tq[2 * ch + 1]; }
    const float* to = bcoef + (size_t)grp * m.ncoef_b + m.cb_o;
    for (int ch = threadIdx.x; ch < NOC; ch += blockDim.x) { s.oA[ch] = to[3 * ch]; s.oB[ch] = to[3 * ch + 1]; s.oC[ch] = to[3 * ch + 2]; }
    if (threadIdx.x < m.nt * MT_G) {
      const int ch = threadIdx.x;
      s.wsim[ch] = coef[(size_t)grp * m.ncoef_f + m.cf_s + 2 * ch];
      if (MODE != 0) {
        const float* ts = bcoef + (size_t)grp * m.ncoef_b + m.cb_s + 3 * ch;
        s.sA[ch] = ts[0]; s.sB[ch] = ts[1]; s.sC[ch] = ts[2];
      }
    }
    if (POS) load_rel2(m, p.relative, s.Rp, s.Rs);
  }
  __syncthreads();
  if (use_tma) { tma_strips_wait(&tma_bar); affine_rows_inplace(s.Yn, SPB * 2 * C, L, s.ca, s.cb, 2 * C); }
  else stage_rows(yraw + (size_t)b0 * 2 * C * L, SPB * 2 * C, L, LP, s.ca, s.cb, 2 * C, s.Yn);
  {
    const int per4 = NOC * L / 4, L4 = L / 4;
    const float4* so = reinterpret_cast<const float4*>(oraw + (size_t)b0 * NOC * L);
    for (int base = 0; base < SPB * per4; base += 2 * blockDim.x) {
      float4 vo[2], vg[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i4 = base + u * blockDim.x + threadIdx.x;
        if (i4 < SPB * per4) {
          const int ss = i4 / per4, rem = i4 - ss * per4, ch = rem / L4, l4 = rem - ch * L4;
          vo[u] = __ldg(so + i4);
          vg[u] = __ldg(reinterpret_cast<const float4*>(gs + ((size_t)(b0 + ss) * C + ch / NO) * L) + l4);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i4 = base + u * blockDim.x + threadIdx.x;
        if (i4 < SPB * per4) {
          const int ss = i4 / per4, rem = i4 - ss * per4, ch = rem / L4, l4 = rem - ch * L4;
          const float a = s.oA[ch], b = s.oB[ch], c = s.oC[ch];
          float4 o = make_float4(fmaf(a, vg[u].x, fmaf(c, vo[u].x, b)), fmaf(a, vg[u].y, fmaf(c, vo[u].y, b)),
                                 fmaf(a, vg[u].z, fmaf(c, vo[u].z, b)), fmaf(a, vg[u].w, fmaf(c, vo[u].w, b)));
          *reinterpret_cast<float4*>(s.dOs + ((size_t)ss * NOC + ch) * LP + 4 * l4) = o;
        }
      }
    }
    for (int idx = threadIdx.x; idx < SPB * MT_G * L; idx += blockDim.x) s.lses[idx] = __ldg(lse + (size_t)b0 * MT_G * L + idx);
  }
  __syncthreads();
  {  // D[s][g][i] = sum over the head's output channels of dO * O_raw
    const int per = NOC * L, gpn = GP * NO;
    for (int idx = threadIdx.x; idx < SPB * MT_G * L; idx += blockDim.x) {
      const int i = idx % L, g = (idx / L) % MT_G, ss = idx / (L * MT_G);
      const float* o = oraw + (size_t)(b0 + ss) * per + (size_t)g * gpn * L + i;
      const float* dO = s.dOs + ((size_t)ss * NOC + g * gpn) * LP + i;
      float ov[gpn], acc = 0.f;
#pragma unroll
      for (int k = 0; k < gpn; ++k) ov[k] = __ldg(o + (size_t)k * L);
#pragma unroll
      for (int k = 0; k < gpn; ++k) acc = fmaf(dO[(size_t)k * LP], ov[k], acc);
      s.Dl[idx] = acc;
    }
  }
  __syncthreads();

  const int nitems = (diag_cta || (m.dbg & 4)) ? 0 : SPB * MT_G * LB;   // dbg bit2: fixed-cost experiment
  const int ph_lo = NMAIN == 2 ? (int)blockIdx.y : 0;                   // split variant: CTA y runs pass y only
  const int ph_hi = MODE == 0 ? 1 : (NMAIN == 2 ? ph_lo + 1 : 2);
  for (int phase = ph_lo; phase < ph_hi; ++phase)
  for (int base = 0; base < nitems; base += blockDim.x) {
    const int item = base + threadIdx.x;
    const bool act = item < nitems;
    const int it = act ? item : 0;
    const int t = it % LB, g = (it / LB) % MT_G, ss = it / (LB * MT_G);
    const float* Yb = s.Yn + ((size_t)ss * 2 * C + g * 4 * D) * LP;
    const float* dOb = s.dOs + ((size_t)ss * NOC + g * GP * NO) * LP;
    const float* lsp = s.lses + (ss * MT_G + g) * L;
    const float* Dp = s.Dl + (ss * MT_G + g) * L;
    const float wqk = s.wsim[g], wqr = POS ? s.wsim[MT_G + g] : 0.f, wkr = POS ? s.wsim[2 * MT_G + g] : 0.f;
    const float Aqk = MODE ? s.sA[g] : 0.f, Bqk = MODE ? s.sB[g] : 0.f, Cqk = MODE ? s.sC[g] : 0.f;
    const float Aqr = (MODE && POS) ? s.sA[MT_G + g] : 0.f, Bqr = (MODE && POS) ? s.sB[MT_G + g] : 0.f,
                Cqr = (MODE && POS) ? s.sC[MT_G + g] : 0.f;
    const float Akr = (MODE && POS) ? s.sA[2 * MT_G + g] : 0.f, Bkr = (MODE && POS) ? s.sB[2 * MT_G + g] : 0.f,
                Ckr = (MODE && POS) ? s.sC[2 * MT_G + g] : 0.f;
    const int KA0 = t + LB, KB0 = LB - t;

    if (MODE == 0 || phase == 0) {
      // ------------------------------------------------------------ thread owns rows i = RT*t + r
      float q[D][RT], dsv[GP][RT], dse[GP][RT], lsr[RT], Dr[RT];
#pragma unroll
      for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * LP + RT * t);
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        ldv<RT>(dsv[c], dOb + (c * NO) * LP + RT * t);
        if (POS) ldv<RT>(dse[c], dOb + (c * NO + 1) * LP + RT * t);
      }
      ldv<RT>(lsr, lsp + RT * t); ldv<RT>(Dr, Dp + RT * t);
      WinA<RT> wq[D]; WinB<RT> wk[D]; WinA<RT> wv[GP];
      if (POS) {
#pragma unroll
        for (int c = 0; c < D; ++c) { wq[c].init(s.Rp + c * RP, KA0); wk[c].init(s.Rs + (D + c) * RP, KB0); }
#pragma unroll
        for (int c = 0; c < GP; ++c) wv[c].init(s.Rp + (GP + c) * RP, KA0);
      }
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, dq[D][RT];
#pragma unroll
      for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < RT; ++r) dq[c][r] = 0.f;
      for (int wb = 0; wb < LB; ++wb) {
        float kk[D][RT], vv[GP][RT];
#pragma unroll
        for (int c = 0; c < D; ++c) ldv<RT>(kk[c], Yb + (D + c) * LP + RT * wb);
#pragma unroll
        for (int c = 0; c < GP; ++c) ldv<RT>(vv[c], Yb + (GP + c) * LP + RT * wb);
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
            if (MODE == 0) {
              t0 = fmaf(dZ, sqk, t0);
              if (POS) { t1 = fmaf(dZ, sqr, t1); t2 = fmaf(dZ, skr, t2); }
            } else {
              const float dSqk = fmaf(Aqk, dZ, fmaf(Cqk, sqk, Bqk));
              const float dSqr = POS ? fmaf(Aqr, dZ, fmaf(Cqr, sqr, Bqr)) : 0.f;
#pragma unroll
              for (int c = 0; c < D; ++c) {
                dq[c][r] = fmaf(dSqk, kk[c][w], dq[c][r]);
                if (POS) dq[c][r] = fmaf(dSqr, wq[c].get(r, w), dq[c][r]);
              }
            }
          }
        if (POS && wb + 1 < LB) {
#pragma unroll
          for (int c = 0; c < D; ++c) { wq[c].advance(s.Rp + c * RP, KA0 - wb - 1); wk[c].advance(s.Rs + (D + c) * RP, KB0 + wb + 1); }
#pragma unroll
          for (int c = 0; c < GP; ++c) wv[c].advance(s.Rp + (GP + c) * RP, KA0 - wb - 1);
        }
      }
      if (MODE == 0) {
        if (!act) { t0 = t1 = t2 = 0.f; }
        t0 = seg_sum(t0, segw);
        if (POS) { t1 = seg_sum(t1, segw); t2 = seg_sum(t2, segw); }
        if (act && (threadIdx.x & (segw - 1)) == 0) {
          atomicAdd(s.red + g, t0);
          if (POS) { atomicAdd(s.red + MT_G + g, t1); atomicAdd(s.red + 2 * MT_G + g, t2); }
        }
      } else {
        const int b = b0 + ss;
        typedef typename VecT<RT>::T V;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const int ch = g * 2 * GP + c;
          const size_t o = ((size_t)b * 2 * C + ch) * L + RT * t;
          float yr[RT], s1 = 0.f, s2 = 0.f;
          if (act) {
            const V yv = __ldg(reinterpret_cast<const V*>(yraw + o));
            const float* f = reinterpret_cast<const float*>(&yv);
#pragma unroll
            for (int r = 0; r < RT; ++r) yr[r] = f[r];
            *reinterpret_cast<V*>(dyh + o) = *reinterpret_cast<V*>(dq[c]);
#pragma unroll
            for (int r = 0; r < RT; ++r) { s1 += dq[c][r]; s2 = fmaf(dq[c][r], yr[r], s2); }
          }
          s1 = seg_sum(s1, segw); s2 = seg_sum(s2, segw);
          if (act && (threadIdx.x & (segw - 1)) == 0) { atomicAdd(s.red + ch * 2, s1); atomicAdd(s.red + ch * 2 + 1, s2); }
        }
      }
    } else {
      // ------------------------------------------------------------ thread owns columns j = RT*t + r
      float kc_[D][RT], vc_[GP][RT];
#pragma unroll
      for (int c = 0; c < D; ++c) ldv<RT>(kc_[c], Yb + (D + c) * LP + RT * t);
#pragma unroll
      for (int c = 0; c < GP; ++c) ldv<RT>(vc_[c], Yb + (GP + c) * LP + RT * t);
      // index i - j + L - 1 = walk - own : WinB on the shifted copy (Rq, Rv); j - i + L - 1 = own - walk : WinA (Rk)
      WinB<RT> wq[D]; WinA<RT> wk[D]; WinB<RT> wv[GP];
      if (POS) {
#pragma unroll
        for (int c = 0; c < D; ++c) { wq[c].init(s.Rs + c * RP, KB0); wk[c].init(s.Rp + (D + c) * RP, KA0); }
#pragma unroll
        for (int c = 0; c < GP; ++c) wv[c].init(s.Rs + (GP + c) * RP, KB0);
      }
      float dk[D][RT], dv[GP][RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
#pragma unroll
        for (int c = 0; c < D; ++c) dk[c][r] = 0.f;
#pragma unroll
        for (int c = 0; c < GP; ++c) dv[c][r] = 0.f;
      }
      for (int wb = 0; wb < LB; ++wb) {
        float qq[D][RT], osv[GP][RT], ose[GP][RT], lsw[RT], Dw[RT];
#pragma unroll
        for (int c = 0; c < D; ++c) ldv<RT>(qq[c], Yb + c * LP + RT * wb);
#pragma unroll
        for (int c = 0; c < GP; ++c) {
          ldv<RT>(osv[c], dOb + (c * NO) * LP + RT * wb);
          if (POS) ldv<RT>(ose[c], dOb + (c * NO + 1) * LP + RT * wb);
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
            for (int c = 0; c < GP; ++c) dv[c][r] = fmaf(P, osv[c][w], dv[c][r]);
          }
        if (POS && wb + 1 < LB) {
#pragma unroll
          for (int c = 0; c < D; ++c) { wq[c].advance(s.Rs + c * RP, KB0 + wb + 1); wk[c].advance(s.Rp + (D + c) * RP, KA0 - wb - 1); }
#pragma unroll
          for (int c = 0; c < GP; ++c) wv[c].advance(s.Rs + (GP + c) * RP, KB0 + wb + 1);
        }
      }
      const int b = b0 + ss;
      typedef typename VecT<RT>::T V;
#pragma unroll
      for (int c = 0; c < D + GP; ++c) {
        const int ch = g * 2 * GP + D + c;                     // k rows [D,2D) then v rows [2D,4D)
        const size_t o = ((size_t)b * 2 * C + ch) * L + RT * t;
        float val[RT], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int r = 0; r < RT; ++r) val[r] = c < D ? dk[c < D ? c : 0][r] : dv[c >= D ? c - D : 0][r];
        if (act) {
          const V yv = __ldg(reinterpret_cast<const V*>(yraw + o));
          const float* f = reinterpret_cast<const float*>(&yv);
          *reinterpret_cast<V*>(dyh + o) = *reinterpret_cast<V*>(val);
#pragma unroll
          for (int r = 0; r < RT; ++r) { s1 += val[r]; s2 = fmaf(val[r], f[r], s2); }
        }
        s1 = seg_sum(s1, segw); s2 = seg_sum(s2, segw);
        if (act && (threadIdx.x & (segw - 1)) == 0) { atomicAdd(s.red + ch * 2, s1); atomicAdd(s.red + ch * 2 + 1, s2); }
      }
    }
  }
  // ------------------------------------------------------------ diagonals (d relative)
  if (TDIAG && diag_cta) {          // experimental: register-tiled wrapped diagonals, as many balanced items as the row pass
    const int nit = SPB * MT_G * LB;
    const int per = (nit + nds - 1) / nds, i0 = dsplit * per, i1 = i0 + per < nit ? i0 + per : nit;
    for (int item = i0 + threadIdx.x; item < i1; item += blockDim.x) {
      const int t = item % LB, g = (item / LB) % MT_G, ss = item / (LB * MT_G);
      mtdiag::Coefs cf;
      cf.wqk = s.wsim[g]; cf.wqr = s.wsim[MT_G + g]; cf.wkr = s.wsim[2 * MT_G + g];
      cf.Aqr = s.sA[MT_G + g]; cf.Bqr = s.sB[MT_G + g]; cf.Cqr = s.sC[MT_G + g];
      cf.Akr = s.sA[2 * MT_G + g]; cf.Bkr = s.sB[2 * MT_G + g]; cf.Ckr = s.sC[2 * MT_G + g];
      mtdiag::diag_tile_item<(D <= 2 ? D : 1), RT>(L, LP, RP, t, s.Yn + ((size_t)ss * 2 * C + g * 4 * D) * LP,
                                                   s.dOs + ((size_t)ss * NOC + g * GP * NO) * LP, s.lses + (ss * MT_G + g) * L,
                                                   s.Dl + (ss * MT_G + g) * L, s.Rp, cf, dRs);
    }
  } else if (POS && diag_cta) {                                   // default: one thread per diagonal, scalar walk
    const int RL = 2 * L - 1, nd = SPB * MT_G * RL;
    const int per = (nd + nds - 1) / nds, i0 = dsplit * per, i1 = i0 + per < nd ? i0 + per : nd;
    for (int item = i0 + threadIdx.x; item < i1; item += blockDim.x) {
      const int mm = item % RL, g = (item / RL) % MT_G, ss = item / (RL * MT_G);
      const float* Yb = s.Yn + ((size_t)ss * 2 * C + g * 4 * D) * LP;
      const float* dOb = s.dOs + ((size_t)ss * NOC + g * GP * NO) * LP;
      const float wqk = s.wsim[g], wqr = s.wsim[MT_G + g], wkr = s.wsim[2 * MT_G + g];
      const float Aqr = s.sA[MT_G + g], Bqr = s.sB[MT_G + g], Cqr = s.sC[MT_G + g];
      const float Akr = s.sA[2 * MT_G + g], Bkr = s.sB[2 * MT_G + g], Ckr = s.sC[2 * MT_G + g];
      const float* lsp = s.lses + (ss * MT_G + g) * L;
      const float* Dp = s.Dl + (ss * MT_G + g) * L;
      float rq[D], rk[D], rv[GP], aq[D], ak[D], av[GP];
#pragma unroll
      for (int c = 0; c < D; ++c) { rq[c] = s.Rp[c * RP + mm]; rk[c] = s.Rp[(D + c) * RP + (2 * (L - 1) - mm)]; aq[c] = 0.f; ak[c] = 0.f; }
#pragma unroll
      for (int c = 0; c < GP; ++c) { rv[c] = s.Rp[(GP + c) * RP + mm]; av[c] = 0.f; }
      const int ilo = mm - (L - 1) > 0 ? mm - (L - 1) : 0, ihi = mm < L - 1 ? mm : L - 1;
      for (int i = ilo; i <= ihi; ++i) {
        const int j = i - mm + L - 1;
        float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float qc = Yb[c * LP + i], kc = Yb[(D + c) * LP + j];
          sqk = fmaf(qc, kc, sqk); sqr = fmaf(qc, rq[c], sqr); skr = fmaf(kc, rk[c], skr);
        }
        const float P = ex2(fmaf(wqr, sqr, fmaf(wkr, skr, wqk * sqk)) - lsp[i]);
#pragma unroll
        for (int c = 0; c < GP; ++c) dP = fmaf(dOb[(2 * c) * LP + i], Yb[(GP + c) * LP + j], fmaf(dOb[(2 * c + 1) * LP + i], rv[c], dP));
        const float dZ = P * (dP - Dp[i]);
        const float dSqr = fmaf(Aqr, dZ, fmaf(Cqr, sqr, Bqr)), dSkr = fmaf(Akr, dZ, fmaf(Ckr, skr, Bkr));
#pragma unroll
        for (int c = 0; c < D; ++c) { aq[c] = fmaf(dSqr, Yb[c * LP + i], aq[c]); ak[c] = fmaf(dSkr, Yb[(D + c) * LP + j], ak[c]); }
#pragma unroll
        for (int c = 0; c < GP; ++c) av[c] = fmaf(P, dOb[(2 * c + 1) * LP + i], av[c]);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) { atomicAdd(dRs + c * RL + mm, aq[c]); atomicAdd(dRs + (D + c) * RL + (2 * (L - 1) - mm), ak[c]); }
#pragma unroll
      for (int c = 0; c < GP; ++c) atomicAdd(dRs + (GP + c) * RL + mm, av[c]);
    }
  }
  __syncthreads();
  double* bs = stat_slot(m, bstats, grp, m.nstat_b, blockIdx.x);
  if (MODE == 0) {
    if (threadIdx.x < m.nt * MT_G) MT_GADD(m, bs + m.b_sim_t + threadIdx.x, (double)s.red[threadIdx.x]);
    if (last_cta_done(bcnt + 1, gridDim.x)) fin_bwd_sim(m, p, stats, bstats, bcoef);
  } else {
    if (!diag_cta)
      for (int ch = threadIdx.x; ch < 2 * C; ch += blockDim.x) {
        MT_GADD(m, bs + m.b_qkv_s1 + ch, (double)s.red[ch * 2]);
        MT_GADD(m, bs + m.b_qkv_s2 + ch, (double)s.red[ch * 2 + 1]);
      }
    if (POS && diag_cta)
      for (int k = threadIdx.x; k < nrel; k += blockDim.x) MT_GADD(m, gr.drelative + k, dRs[k]);
    // last CTA of all passes: bn_qkv backward coefficients for the projection GEMMs + every BN / gate gradient
    if (last_cta_done(bcnt + 2, gridDim.x * gridDim.y)) {
      fin_bwd_qkv(m, p, stats, bstats, bcoef);
      bwd_param_grads(m, p, gr, stats, bstats);
    }
  }
}

// ---------------------------------------------------------------------------- host side
int attn2_rt(const Dims& m);

template <int D, bool POS, int RT, int MODE, int XP>
static int launch_bx(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const double* stats, double* bstats,
                     const float* coef, float* bcoef, unsigned* bcnt, const float* yraw, const float* oraw, const float* lse,
                     const float* gs, float* dyh, cudaStream_t st) {
  const int LB = m.L / RT;
  int spb = 1;
  while (spb * MT_G * LB < 128 && spb * 2 <= m.Bg && m.Bg % (spb * 2) == 0) spb *= 2;
  const int segw = LB < 32 ? LB : 32;
  const int nrel = m.pos ? 2 * m.gp * (2 * m.L - 1) : 0;
  const size_t smem = bwd2_smem_floats(m, spb, MODE == 0 ? 24 : 4 * m.C + nrel) * sizeof(float);
  if (smem > 200 * 1024) return 0;             // fall back to the v1 kernel
  if (smem > 32 * 1024 &&
      cudaFuncSetAttribute(k_bwd2<D, POS, RT, MODE, XP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -1;
  constexpr bool TDIAG = (XP & 1) && POS && MODE == 1 && D <= 2;
  constexpr int NMAIN = ((XP & 2) && MODE == 1) ? 2 : 1;
  const int ndiag = (MODE == 1 && POS) ? (TDIAG ? 1 : 2) : 0;         // default: y = 0 rows then columns; y = 1, 2 diagonal walk
  const dim3 grid(m.B / spb, NMAIN + ndiag);
  k_bwd2<D, POS, RT, MODE, XP><<<grid, 128, smem, st>>>(m, p, gr, stats, bstats, coef, bcoef, bcnt, yraw, oraw, lse, gs, dyh,
                                                        spb, segw);
  return 1;
}

template <int D, bool POS, int RT, int MODE>
static int launch_b(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const double* stats, double* bstats,
                    const float* coef, float* bcoef, unsigned* bcnt, const float* yraw, const float* oraw, const float* lse,
                    const float* gs, float* dyh, cudaStream_t st) {
#define GX(XX) return launch_bx<D, POS, RT, MODE, XX>(m, p, gr, stats, bstats, coef, bcoef, bcnt, yraw, oraw, lse, gs, dyh, st)
  if constexpr (MODE == 1) {                      // MT_DEBUG_FLAGS bits 3, 4 -> experimental variants (see k_bwd2)
    const int xp = (m.dbg >> 3) & 3;
    if constexpr (POS && D <= 2) {
      if (xp == 1) { GX(1); }
      if (xp == 3) { GX(3); }
    }
    if (xp & 2) { GX(2); }
  }
  GX(0);
#undef GX
}

// returns launches done (0 = shape not covered -> caller falls back), <0 on error
template <int MODE>
int launch_bwd2(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const double* stats, double* bstats,
                const float* coef, float* bcoef, unsigned* bcnt, const float* yraw, const float* oraw, const float* lse,
                const float* gs, float* dyh, cudaStream_t st) {
  const int rt = attn2_rt(m);
  if ((m.L & 3) || (m.LP & 3)) return 0;       // float4 staging of dO needs L % 4 == 0
#define GO(DD, RR)                                                                                                    \
  return m.pos ? launch_b<DD, true, RR, MODE>(m, p, gr, stats, bstats, coef, bcoef, bcnt, yraw, oraw, lse, gs, dyh, st)   \
               : launch_b<DD, false, RR, MODE>(m, p, gr, stats, bstats, coef, bcoef, bcnt, yraw, oraw, lse, gs, dyh, st)
  if (m.d == 1 && rt == 4) { GO(1, 4); }
  if (m.d == 1 && rt == 2) { GO(1, 2); }
  if (m.d == 2 && rt == 2) { GO(2, 2); }
  if (m.d == 4 && rt == 2) { GO(4, 2); }
#undef GO
  return 0;
}
template int launch_bwd2<0>(const Dims&, const MtAxialParams&, const MtAxialGrads&, const double*, double*, const float*, float*, unsigned*, const float*, const float*, const float*, const float*, float*, cudaStream_t);
template int launch_bwd2<1>(const Dims&, const MtAxialParams&, const MtAxialGrads&, const double*, double*, const float*, float*, unsigned*, const float*, const float*, const float*, const float*, float*, cudaStream_t);
