// axial_attn2_fwd.cu -- register-tiled forward attention passes (see axial_attn2.cuh):
//   MODE 0  similarity statistics: sum / sumsq of qk, qr, kr per (term, head)   [train only]
//   MODE 1  softmax_j(sum_t w_t S_t) -> o_raw = [P v | P Rv], log-sum-exp, bn_output sums
// One thread owns RT consecutive query rows of one head; reference: axialnet.py:155-172.
#include "axial_attn2.cuh"
extern thread_local cudaError_t g_last_cuda;

template <int D, bool POS, int RT, int MODE>
__global__ void __launch_bounds__(128)
k_attn2(Dims m, MtAxialParams p, const float* __restrict__ yraw, double* __restrict__ stats, float* __restrict__ coef,
        unsigned* __restrict__ cnt, float* __restrict__ oraw, float* __restrict__ lse, int SPB, int segw) {
  constexpr int GP = 2 * D;
  constexpr int NO = POS ? 2 : 1;
  extern __shared__ __align__(16) float sm[];
  const int C = m.C, L = m.L, LP = m.LP, RP = rel_pitch(L), LB = L / RT;
  float* ca = sm;
  float* cb = ca + 2 * C;
  float* wsim = cb + 2 * C;                         // [3][G]
  float* red = wsim + 24;                           // MODE 0: [3][G][2]; MODE 1: [NO*C][2]
  const int nred = MODE == 0 ? 48 : 2 * NO * C;
  float* Yn = red + ((nred + 3) & ~3);
  float* Rp = Yn + (size_t)SPB * 2 * C * LP;        // plain tables
  float* Rs = Rp + 2 * GP * RP;                     // shifted tables
  const int b0 = blockIdx.x * SPB, grp = b0 / m.Bg;
  __shared__ __align__(8) uint64_t tma_bar;
  const bool use_tma = (LP == L) && ((L & 3) == 0);
  if (use_tma && threadIdx.x == 0)
    tma_strips_issue(&tma_bar, Yn, yraw + (size_t)b0 * 2 * C * L, (uint32_t)(SPB * 2 * C * L * sizeof(float)));
  {
    const float* tb = coef + (size_t)grp * m.ncoef_f + m.cf_q;
    for (int ch = threadIdx.x; ch < 2 * C; ch += blockDim.x) { ca[ch] = tb[2 * ch]; cb[ch] = tb[2 * ch + 1]; }
  }
  if (MODE == 1 && threadIdx.x < m.nt * MT_G) wsim[threadIdx.x] = coef[(size_t)grp * m.ncoef_f + m.cf_s + 2 * threadIdx.x];
  for (int k = threadIdx.x; k < nred; k += blockDim.x) red[k] = 0.f;
  if (POS) load_rel2(m, p.relative, Rp, Rs);
  __syncthreads();
  if (use_tma) { tma_strips_wait(&tma_bar); affine_rows_inplace(Yn, SPB * 2 * C, L, ca, cb, 2 * C); }
  else stage_rows(yraw + (size_t)b0 * 2 * C * L, SPB * 2 * C, L, LP, ca, cb, 2 * C, Yn);
  __syncthreads();

  const int nitems = (m.dbg & 4) ? 0 : SPB * MT_G * LB;      // MT_DEBUG_FLAGS bit2: skip the main loop (fixed-cost experiment)
  for (int base = 0; base < nitems; base += blockDim.x) {
    const int item = base + threadIdx.x;
    const bool act = item < nitems;
    const int it = act ? item : 0;
    const int t = it % LB, g = (it / LB) % MT_G, s = it / (LB * MT_G);
    const float* Yb = Yn + ((size_t)s * 2 * C + g * 4 * D) * LP;
    float q[D][RT];
#pragma unroll
    for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * LP + RT * t);
    WinA<RT> wq[D]; WinB<RT> wk[D]; WinA<RT> wv[GP];
    const int KA0 = t + LB, KB0 = LB - t;
    if (POS) {
#pragma unroll
      for (int c = 0; c < D; ++c) { wq[c].init(Rp + c * RP, KA0); wk[c].init(Rs + (D + c) * RP, KB0); }
      if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < GP; ++c) wv[c].init(Rp + (GP + c) * RP, KA0);
      }
    }
    // MODE 0 accumulators
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    // MODE 1 state
    const float wqk = MODE == 1 ? wsim[g] : 1.f, wqr = (MODE == 1 && POS) ? wsim[MT_G + g] : 1.f,
                wkr = (MODE == 1 && POS) ? wsim[2 * MT_G + g] : 1.f;
    float mx[RT], lsum[RT], asv[RT][GP], asve[RT][GP];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      mx[r] = -INFINITY; lsum[r] = 0.f;
#pragma unroll
      for (int c = 0; c < GP; ++c) { asv[r][c] = 0.f; asve[r][c] = 0.f; }
    }
    for (int wb = 0; wb < LB; ++wb) {
      float kk[D][RT], vv[GP][RT];
#pragma unroll
      for (int c = 0; c < D; ++c) ldv<RT>(kk[c], Yb + (D + c) * LP + RT * wb);
      if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < GP; ++c) ldv<RT>(vv[c], Yb + (GP + c) * LP + RT * wb);
      }
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
          if (MODE == 0) {
            a0 += sqk; a1 = fmaf(sqk, sqk, a1);
            if (POS) { a2 += sqr; a3 = fmaf(sqr, sqr, a3); a4 += skr; a5 = fmaf(skr, skr, a5); }
          } else {
            z[w] = wqk * sqk;
            if (POS) z[w] = fmaf(wqr, sqr, fmaf(wkr, skr, z[w]));
          }
        }
        if (MODE == 1) {
          float bm = z[0];
#pragma unroll
          for (int w = 1; w < RT; ++w) bm = fmaxf(bm, z[w]);
          if (bm > mx[r]) {                       // lazy rescale, once per RT columns
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
      }
      if (POS && wb + 1 < LB) {
#pragma unroll
        for (int c = 0; c < D; ++c) { wq[c].advance(Rp + c * RP, KA0 - wb - 1); wk[c].advance(Rs + (D + c) * RP, KB0 + wb + 1); }
        if (MODE == 1) {
#pragma unroll
          for (int c = 0; c < GP; ++c) wv[c].advance(Rp + (GP + c) * RP, KA0 - wb - 1);
        }
      }
    }
    if (MODE == 0) {
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
    } else {
      const int b = b0 + s;
      float inv[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) inv[r] = 1.f / lsum[r];
      if (act) {
        typedef typename VecT<RT>::T V;
        float lv[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) lv[r] = mx[r] + lg2(lsum[r]);
        *reinterpret_cast<V*>(lse + ((size_t)b * MT_G + g) * L + RT * t) = *reinterpret_cast<V*>(lv);
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        const int ch = (g * GP + c) * NO;
        float o0[RT], o1[RT];
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          o0[r] = asv[r][c] * inv[r]; o1[r] = POS ? asve[r][c] * inv[r] : 0.f;
          s0 += o0[r]; q0 = fmaf(o0[r], o0[r], q0); s1 += o1[r]; q1 = fmaf(o1[r], o1[r], q1);
        }
        if (act) {
          typedef typename VecT<RT>::T V;
          *reinterpret_cast<V*>(oraw + ((size_t)b * NO * C + ch) * L + RT * t) = *reinterpret_cast<V*>(o0);
          if (POS) *reinterpret_cast<V*>(oraw + ((size_t)b * NO * C + ch + 1) * L + RT * t) = *reinterpret_cast<V*>(o1);
        }
        if (m.training) {
          if (!act) { s0 = q0 = s1 = q1 = 0.f; }
          s0 = seg_sum(s0, segw); q0 = seg_sum(q0, segw);
          if (POS) { s1 = seg_sum(s1, segw); q1 = seg_sum(q1, segw); }
          if (act && (threadIdx.x & (segw - 1)) == 0) {
            atomicAdd(red + ch * 2, s0); atomicAdd(red + ch * 2 + 1, q0);
            if (POS) { atomicAdd(red + (ch + 1) * 2, s1); atomicAdd(red + (ch + 1) * 2 + 1, q1); }
          }
        }
      }
    }
  }
  if (MODE == 0) {
    __syncthreads();
    double* st = stat_slot(m, stats, grp, m.nstat_f, blockIdx.x);
    if (threadIdx.x < m.nt * MT_G) {
      MT_GADD(m, st + m.f_sim_sum + threadIdx.x, (double)red[threadIdx.x * 2 + 0]);
      MT_GADD(m, st + m.f_sim_sq + threadIdx.x, (double)red[threadIdx.x * 2 + 1]);
    }
    if (last_cta_done(cnt + 1, gridDim.x)) fin_sim(m, p, stats, coef);
  } else {
    if (m.training) {
      __syncthreads();
      double* st = stat_slot(m, stats, grp, m.nstat_f, blockIdx.x);
      for (int ch = threadIdx.x; ch < NO * C; ch += blockDim.x) {
        MT_GADD(m, st + m.f_out_sum + ch, (double)red[ch * 2]);
        MT_GADD(m, st + m.f_out_sq + ch, (double)red[ch * 2 + 1]);
      }
    }
    if (last_cta_done(cnt + 2, gridDim.x)) fin_out(m, p, stats, coef);
  }
}

// ---------------------------------------------------------------------------- host side
static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// RT for this shape, or 0 when the v2 layout does not apply (falls back to the v1 kernels)
int attn2_rt(const Dims& m) {
  int rt = m.d == 1 ? 4 : (m.d == 2 || m.d == 4 ? 2 : 0);
  // MT_FORCE_RT (developer): RT for the d = 1 layers.  Measured: RT 2 vs 4 is a wash at 128 px (profiles/r1_notes.md)
  static const int force = getenv("MT_FORCE_RT") ? atoi(getenv("MT_FORCE_RT")) : 0;
  if ((force == 2 || force == 4) && m.d == 1) rt = force;
  if (rt == 4 && (m.L % 4 || !is_pow2(m.L / 4))) rt = 2;
  if (rt == 2 && (m.L % 2 || !is_pow2(m.L / 2))) rt = 0;
  if (rt && m.L / rt > 0 && (m.LP % rt)) rt = 0;
  return rt;
}

template <int D, bool POS, int RT, int MODE>
static int launch_one(const Dims& m, const MtAxialParams& p, const float* yraw, double* stats, float* coef, unsigned* cnt,
                      float* oraw, float* lse, cudaStream_t st) {
  const int LB = m.L / RT;
  int spb = 1;
  while (spb * MT_G * LB < 128 && spb * 2 <= m.Bg && m.Bg % (spb * 2) == 0) spb *= 2;
  const int segw = LB < 32 ? LB : 32;
  const int nred = MODE == 0 ? 48 : 2 * m.no * m.C;
  const size_t fl = 4 * (size_t)m.C + 24 + ((nred + 3) & ~3) + (size_t)spb * 2 * m.C * m.LP + (POS ? 4 * (size_t)m.gp * rel_pitch(m.L) : 0);
  const size_t smem = fl * sizeof(float);
  if (smem > 200 * 1024) return -2;
  if (smem > 32 * 1024 &&
      cudaFuncSetAttribute(k_attn2<D, POS, RT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return -1;
  k_attn2<D, POS, RT, MODE><<<m.B / spb, 128, smem, st>>>(m, p, yraw, stats, coef, cnt, oraw, lse, spb, segw);
  return 1;
}

template <int MODE>
int launch_attn2(const Dims& m, const MtAxialParams& p, const float* yraw, double* stats, float* coef, unsigned* cnt,
                 float* oraw, float* lse, cudaStream_t st) {
  const int rt = attn2_rt(m);
#define GO(DD, RR)                                                                                              \
  return m.pos ? launch_one<DD, true, RR, MODE>(m, p, yraw, stats, coef, cnt, oraw, lse, st)                     \
               : launch_one<DD, false, RR, MODE>(m, p, yraw, stats, coef, cnt, oraw, lse, st)
  if (m.d == 1 && rt == 4) { GO(1, 4); }
  if (m.d == 1 && rt == 2) { GO(1, 2); }
  if (m.d == 2 && rt == 2) { GO(2, 2); }
  if (m.d == 4 && rt == 2) { GO(4, 2); }
#undef GO
  return 0;    // not covered: caller uses the v1 kernel
}
template int launch_attn2<0>(const Dims&, const MtAxialParams&, const float*, double*, float*, unsigned*, float*, float*, cudaStream_t);
template int launch_attn2<1>(const Dims&, const MtAxialParams&, const float*, double*, float*, unsigned*, float*, float*, cudaStream_t);
