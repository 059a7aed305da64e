// axial_diag_tile.cuh -- register-tiled pass for d relative (EXPERIMENTAL: MT_DEBUG_FLAGS bit 3, off by default).
//
// d relative[c, m] sums, over all pairs (i, j) of one diagonal m = i - j + L - 1, terms that need the full per-pair
// state (score, P, dZ).  The default kernel walks each diagonal with one thread and ~20 scalar shared-memory loads per
// pair, and diagonals have 1..L pairs (a triangle of work).  Here a thread owns RT consecutive WRAPPED diagonals
// r = (i - j) mod L -- every wrapped diagonal has exactly L pairs, so the items are balanced and there are as many of
// them as in the row pass -- and walks i in blocks of RT:
//   * q / dO / lse / D of the RT rows: one vector load each per block (as in the row pass);
//   * k / v of the 2RT-1 columns the RT x RT pair block touches: SLIDING REGISTER WINDOWS over the k / v rows (the
//     dual of the row pass, where the tables slide and k / v are block-loaded), advancing one block with wrap-around;
//   * the table values are per-diagonal register CONSTANTS.  A wrapped diagonal r is two real ones: m1 = r + L - 1
//     (pairs with i >= r, no wrap) and m2 = r - 1 (pairs with i < r): blocks before the thread's own block use set 2,
//     blocks after it set 1, and inside its own block the choice (a >= b) is static after unrolling;
//   * accumulators stay in registers; one atomicAdd per (table row, diagonal) at the end.
// Arithmetic per pair is identical to the scalar walk in k_bwd2 (same operations, same order inside a pair).
//
// The routine is __host__ __device__ so that tests/test_diag_tile_host.py can run it on the CPU against a plain
// double loop over (i, j) (oracle-style definition) -- the index logic is what can go wrong, not the arithmetic.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MT_HD __host__ __device__ __forceinline__
#else
#define MT_HD inline
#endif

namespace mtdiag {

struct Coefs { float wqk, wqr, wkr, Aqr, Bqr, Cqr, Akr, Bkr, Ckr; };

MT_HD float exp2_hd(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));      // == ex2() of common.cuh
  return y;
#else
  return exp2f(x);
#endif
}
MT_HD void add_hd(float* p, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
template <int RT>
MT_HD void ldblk(float (&d)[RT], const float* p) {            // RT consecutive floats, RT*4-byte aligned
#if defined(__CUDA_ARCH__)
  if (RT == 4) { const float4 v = *reinterpret_cast<const float4*>(p); d[0] = v.x; d[1] = v.y; d[RT > 2 ? 2 : 0] = v.z; d[RT > 3 ? 3 : 0] = v.w; return; }
  if (RT == 2) { const float2 v = *reinterpret_cast<const float2*>(p); d[0] = v.x; d[1] = v.y; return; }
#endif
#pragma unroll
  for (int k = 0; k < RT; ++k) d[k] = p[k];
}

// one RT x RT pair block.  MODE 0: every pair belongs to set 2 (i < r); 2: set 1; 1: the thread's own block (a >= b -> set 1)
template <int D, int RT, int MODE>
MT_HD void pair_block(const Coefs& cf, const float (&q)[D][RT], const float (&dsv)[2 * D][RT], const float (&dse)[2 * D][RT],
                      const float (&ls)[RT], const float (&Dd)[RT],
                      const float (&khi)[D][RT], const float (&klo)[D][RT], const float (&vhi)[2 * D][RT], const float (&vlo)[2 * D][RT],
                      const float (&rq1)[D][RT], const float (&rk1)[D][RT], const float (&rv1)[2 * D][RT],
                      const float (&rq2)[D][RT], const float (&rk2)[D][RT], const float (&rv2)[2 * D][RT],
                      float (&aq1)[D][RT], float (&ak1)[D][RT], float (&av1)[2 * D][RT],
                      float (&aq2)[D][RT], float (&ak2)[D][RT], float (&av2)[2 * D][RT]) {
  constexpr int GP = 2 * D;
#pragma unroll
  for (int a = 0; a < RT; ++a)
#pragma unroll
    for (int b = 0; b < RT; ++b) {
      const bool s1 = MODE == 2 || (MODE == 1 && a >= b);
      const int d = a - b;                                    // column = window slot: hi[d] (d >= 0) or lo[RT + d]
      float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f, kc[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        kc[c] = d >= 0 ? khi[c][d >= 0 ? d : 0] : klo[c][d < 0 ? RT + d : 0];
        const float rq = s1 ? rq1[c][b] : rq2[c][b], rk = s1 ? rk1[c][b] : rk2[c][b];
        sqk = fmaf(q[c][a], kc[c], sqk); sqr = fmaf(q[c][a], rq, sqr); skr = fmaf(kc[c], rk, skr);
      }
      const float P = exp2_hd(fmaf(cf.wqr, sqr, fmaf(cf.wkr, skr, cf.wqk * sqk)) - ls[a]);
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        const float vc = d >= 0 ? vhi[c][d >= 0 ? d : 0] : vlo[c][d < 0 ? RT + d : 0];
        const float rv = s1 ? rv1[c][b] : rv2[c][b];
        dP = fmaf(dsv[c][a], vc, fmaf(dse[c][a], rv, dP));
      }
      const float dZ = P * (dP - Dd[a]);
      const float dSqr = fmaf(cf.Aqr, dZ, fmaf(cf.Cqr, sqr, cf.Bqr)), dSkr = fmaf(cf.Akr, dZ, fmaf(cf.Ckr, skr, cf.Bkr));
#pragma unroll
      for (int c = 0; c < D; ++c) {
        if (s1) { aq1[c][b] = fmaf(dSqr, q[c][a], aq1[c][b]); ak1[c][b] = fmaf(dSkr, kc[c], ak1[c][b]); }
        else    { aq2[c][b] = fmaf(dSqr, q[c][a], aq2[c][b]); ak2[c][b] = fmaf(dSkr, kc[c], ak2[c][b]); }
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        if (s1) av1[c][b] = fmaf(P, dse[c][a], av1[c][b]);
        else    av2[c][b] = fmaf(P, dse[c][a], av2[c][b]);
      }
    }
}

// All pairs of the wrapped diagonals r = RT*rt .. RT*rt + RT - 1 of ONE (strip, head).
//   Yb : head base of the staged strip, pitch LP: rows [0,D) q, [D,2D) k, [2D,4D) v      (BN-qkv already applied)
//   dOb: head base of the staged output gradients, pitch LP: row 2c = d sv[c], 2c+1 = d sve[c]
//   lsp, Dp: log2-sum-exp and D = sum dO*O of the head's rows
//   Rp : plain Toeplitz tables T[row][m], pitch RP: rows [0,D) q-table, [D,2D) k-table, [2D,4D) v-table
//   dRs: accumulator [4D][2L-1], same row order
template <int D, int RT>
MT_HD void diag_tile_item(int L, int LP, int RP, int rt, const float* Yb, const float* dOb, const float* lsp, const float* Dp,
                          const float* Rp, const Coefs& cf, float* dRs) {
  constexpr int GP = 2 * D;
  const int LB = L / RT, RL = 2 * L - 1, r0 = RT * rt;
  float rq1[D][RT], rk1[D][RT], rv1[GP][RT], rq2[D][RT], rk2[D][RT], rv2[GP][RT];
  float aq1[D][RT], ak1[D][RT], av1[GP][RT], aq2[D][RT], ak2[D][RT], av2[GP][RT];
#pragma unroll
  for (int b = 0; b < RT; ++b) {
    const int m1 = r0 + b + L - 1;                            // in [L-1, 2L-2]
    const int m2 = r0 + b - 1 > 0 ? r0 + b - 1 : 0;           // r - 1; r == 0 has no set-2 pairs (clamped, never flushed)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      rq1[c][b] = Rp[c * RP + m1]; rk1[c][b] = Rp[(D + c) * RP + (2 * (L - 1) - m1)];
      rq2[c][b] = Rp[c * RP + m2]; rk2[c][b] = Rp[(D + c) * RP + (2 * (L - 1) - m2)];
      aq1[c][b] = 0.f; ak1[c][b] = 0.f; aq2[c][b] = 0.f; ak2[c][b] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < GP; ++c) { rv1[c][b] = Rp[(GP + c) * RP + m1]; rv2[c][b] = Rp[(GP + c) * RP + m2]; av1[c][b] = 0.f; av2[c][b] = 0.f; }
  }
  // column windows: block cb = (ib - rt) mod LB in hi, block cb - 1 (mod LB) in lo
  float khi[D][RT], klo[D][RT], vhi[GP][RT], vlo[GP][RT];
  int cb = (LB - rt) % LB;
  {
    const int pb = (cb + LB - 1) % LB;
#pragma unroll
    for (int c = 0; c < D; ++c) { ldblk<RT>(khi[c], Yb + (D + c) * LP + RT * cb); ldblk<RT>(klo[c], Yb + (D + c) * LP + RT * pb); }
#pragma unroll
    for (int c = 0; c < GP; ++c) { ldblk<RT>(vhi[c], Yb + (GP + c) * LP + RT * cb); ldblk<RT>(vlo[c], Yb + (GP + c) * LP + RT * pb); }
  }
  for (int ib = 0; ib < LB; ++ib) {
    float q[D][RT], dsv[GP][RT], dse[GP][RT], ls[RT], Dd[RT];
#pragma unroll
    for (int c = 0; c < D; ++c) ldblk<RT>(q[c], Yb + c * LP + RT * ib);
#pragma unroll
    for (int c = 0; c < GP; ++c) { ldblk<RT>(dsv[c], dOb + (2 * c) * LP + RT * ib); ldblk<RT>(dse[c], dOb + (2 * c + 1) * LP + RT * ib); }
    ldblk<RT>(ls, lsp + RT * ib); ldblk<RT>(Dd, Dp + RT * ib);
    if (ib < rt)
      pair_block<D, RT, 0>(cf, q, dsv, dse, ls, Dd, khi, klo, vhi, vlo, rq1, rk1, rv1, rq2, rk2, rv2, aq1, ak1, av1, aq2, ak2, av2);
    else if (ib == rt)
      pair_block<D, RT, 1>(cf, q, dsv, dse, ls, Dd, khi, klo, vhi, vlo, rq1, rk1, rv1, rq2, rk2, rv2, aq1, ak1, av1, aq2, ak2, av2);
    else
      pair_block<D, RT, 2>(cf, q, dsv, dse, ls, Dd, khi, klo, vhi, vlo, rq1, rk1, rv1, rq2, rk2, rv2, aq1, ak1, av1, aq2, ak2, av2);
    if (ib + 1 < LB) {                                        // slide the column windows one block forward (wrapping)
      cb = cb + 1 == LB ? 0 : cb + 1;
#pragma unroll
      for (int c = 0; c < D; ++c) {
#pragma unroll
        for (int k = 0; k < RT; ++k) klo[c][k] = khi[c][k];
        ldblk<RT>(khi[c], Yb + (D + c) * LP + RT * cb);
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) {
#pragma unroll
        for (int k = 0; k < RT; ++k) vlo[c][k] = vhi[c][k];
        ldblk<RT>(vhi[c], Yb + (GP + c) * LP + RT * cb);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < RT; ++b) {
    const int r = r0 + b, m1 = r + L - 1, m2 = r - 1;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      add_hd(dRs + c * RL + m1, aq1[c][b]); add_hd(dRs + (D + c) * RL + (2 * (L - 1) - m1), ak1[c][b]);
      if (m2 >= 0) { add_hd(dRs + c * RL + m2, aq2[c][b]); add_hd(dRs + (D + c) * RL + (2 * (L - 1) - m2), ak2[c][b]); }
    }
#pragma unroll
    for (int c = 0; c < GP; ++c) {
      add_hd(dRs + (GP + c) * RL + m1, av1[c][b]);
      if (m2 >= 0) add_hd(dRs + (GP + c) * RL + m2, av2[c][b]);
    }
  }
}

// the plain definition (what k_bwd2's scalar diagonal walk computes): every pair (i, j), m = i - j + L - 1
template <int D>
MT_HD void diag_reference(int L, int LP, int RP, const float* Yb, const float* dOb, const float* lsp, const float* Dp,
                          const float* Rp, const Coefs& cf, float* dRs) {
  constexpr int GP = 2 * D;
  const int RL = 2 * L - 1;
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const int mm = i - j + L - 1;
      float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f;
      for (int c = 0; c < D; ++c) {
        const float qc = Yb[c * LP + i], kc = Yb[(D + c) * LP + j];
        sqk = fmaf(qc, kc, sqk); sqr = fmaf(qc, Rp[c * RP + mm], sqr); skr = fmaf(kc, Rp[(D + c) * RP + (2 * (L - 1) - mm)], skr);
      }
      const float P = exp2_hd(fmaf(cf.wqr, sqr, fmaf(cf.wkr, skr, cf.wqk * sqk)) - lsp[i]);
      for (int c = 0; c < GP; ++c)
        dP = fmaf(dOb[(2 * c) * LP + i], Yb[(GP + c) * LP + j], fmaf(dOb[(2 * c + 1) * LP + i], Rp[(GP + c) * RP + mm], dP));
      const float dZ = P * (dP - Dp[i]);
      const float dSqr = fmaf(cf.Aqr, dZ, fmaf(cf.Cqr, sqr, cf.Bqr)), dSkr = fmaf(cf.Akr, dZ, fmaf(cf.Ckr, skr, cf.Bkr));
      for (int c = 0; c < D; ++c) {
        dRs[c * RL + mm] += dSqr * Yb[c * LP + i];
        dRs[(D + c) * RL + (2 * (L - 1) - mm)] += dSkr * Yb[(D + c) * LP + j];
      }
      for (int c = 0; c < GP; ++c) dRs[(GP + c) * RL + mm] += P * dOb[(2 * c + 1) * LP + i];
    }
}

}  // namespace mtdiag
