// axial_fused_bwd.cuh -- ONE sweep over the L x L pairs of a (strip, head) for the whole attention backward.
//
// Round 1/2a swept the pairs four times (T = sum dZ*S_t; rows: dq; columns: dk, dv; diagonals: d relative), recomputing the
// scores, exp and dP each time (ncu: 54 M warp instructions for 8.4 M pairs at C16 L64 B256).  Two facts remove three sweeps:
//
//  (1) The BatchNorm(similarity) backward is  dS_t = A_t dZ + Cc_t S_t + Bc_t  with  Cc_t, Bc_t  proportional to the grid-wide
//      sum T_t = sum dZ S_t, but A_t known from the forward statistics.  Every gradient is LINEAR in dS, and the S / constant
//      parts have closed forms that need no pair sweep (S_qk = q.k, S_qr = q.Rq[m], S_kr = k.Rk[m'] are bilinear):
//          sum_j S_qk[i,j] k_j = (K K^T) q_i        sum_j S_qr[i,j] Rq[m] = Mq_i q_i      (Mq_i, PRq_i: window sums of the table)
//      so the sweep only accumulates the dZ parts (and T itself); k_axial_bwd adds the closed-form parts after the barrier
//      that completes T.  (DESIGN.md section 5 derives all terms.)
//  (2) With i - j = r there are three index families (rows i, columns j, diagonals r) and a thread tile can be aligned with two.
//      A lane OWNS RT wrapped diagonals r = (i - j) mod L (table values are register constants, d relative accumulates in
//      registers, every wrapped diagonal has exactly L pairs), all lanes of a (strip, head) WALK the same RT rows per step
//      (row quantities: one transposing shuffle reduction per step), and the columns are skewed by exactly one lane per
//      step: the column accumulators (dk, dv) travel lane -> lane + 1 with one shuffle per value and step, around the ring
//      of the L / RT lanes -- after L / RT steps every accumulator has met every lane.  No shared-memory atomics in the loop.
//
// Per pair (d = 1): 3 score FMAs, 3 for z, ex2, 4 for dP, dZ, 3 for T, 2 row, 4 column, 4 diagonal FMAs = ~30 instructions
// once, against ~40 in each of four sweeps.
#pragma once
#include "axial_attn2.cuh"

namespace mtf {

struct Coefs { float wqk, wqr, wkr, A1, A2, A3; };

// Transposing reduction over the NL lanes of a group (NL a power of two, groups aligned, NL >= NV): on return every lane
// holds in v[0] the group total of value `idx` (lanes with the same low log2(NV) bits hold the same one).
template <int NV>
__device__ __forceinline__ int reduce_scatter(float (&v)[NV], int lig, int NL) {
  int idx = 0;
#pragma unroll
  for (int k = 0, n = NV; n > 1; ++k, n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lig >> k) & 1;
#pragma unroll
    for (int e = 0; e < half; ++e) {
      const float send = upper ? v[e] : v[e + half];
      const float keep = upper ? v[e + half] : v[e];
      v[e] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << k);
    }
    if (upper) idx += half;
  }
  for (int mask = NV; mask < NL; mask <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
  return idx;
}

// All pairs of the wrapped diagonals r = RT*rt .. RT*rt+RT-1 of ONE (strip, head); the NL lanes of the group call it together
// (lig = lane in group; rt = sg*NL + lig for segment sg when L/RT > NL).
//   Yb  : staged head, pitch L: rows [0,D) q, [D,2D) k, [2D,4D) v (bn_qkv applied)
//   dOb : staged dO, pitch L: row c*NO = d sv[c], c*NO+1 = d sve[c]
//   dYb : OUT (+=, zeroed by the caller) pitch L: rows [0,D) A1 sum_j dZ k + A2 sum_j dZ Rq ; [D,2D) A1 sum_i dZ q + A3 sum_i dZ Rk ;
//         [2D,4D) sum_i P dsv
//   dRs : OUT (atomicAdd) [4D][2L-1]: A2 sum dZ q, A3 sum dZ k, sum P dsve along the diagonals
//   t0..t2 : OUT (+=) this lane's share of T_t = sum dZ S_t
// ring: the group covers all L/RT diagonal blocks (column accumulators wrap from the last lane to lane 0); otherwise the chain
// is open: the last lane adds what leaves it to dYb, lane 0 starts from zero (the caller runs the segments one after another).
template <int D, bool POS, int RT>
__device__ __forceinline__ void fused_item(int L, int RP, int rt, int lig, int NL, bool ring, bool act, const float* Yb,
                                           const float* dOb, const float* lsp, const float* Dp, const float* Rp, const Coefs& cf,
                                           float* dRs, float* dYb, float& t0, float& t1, float& t2) {
  constexpr int GP = 2 * D, NO = POS ? 2 : 1, QC = D + GP, NV = D * RT;
  const int LB = L / RT, RL = 2 * L - 1, r0 = RT * rt;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int src = (lane & ~(NL - 1)) | ((lig - 1) & (NL - 1));
  float rq1[D][RT], rk1[D][RT], rv1[GP][RT], rq2[D][RT], rk2[D][RT], rv2[GP][RT];
  float aq1[D][RT], ak1[D][RT], av1[GP][RT], aq2[D][RT], ak2[D][RT], av2[GP][RT];
  if (POS) {
#pragma unroll
    for (int b = 0; b < RT; ++b) {
      const int m1 = r0 + b + L - 1;                            // pairs with i >= r
      const int m2 = r0 + b - 1 > 0 ? r0 + b - 1 : 0;           // pairs with i <  r (r == 0 has none: clamped, never flushed)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        rq1[c][b] = Rp[c * RP + m1]; rk1[c][b] = Rp[(D + c) * RP + (2 * (L - 1) - m1)];
        rq2[c][b] = Rp[c * RP + m2]; rk2[c][b] = Rp[(D + c) * RP + (2 * (L - 1) - m2)];
        aq1[c][b] = 0.f; ak1[c][b] = 0.f; aq2[c][b] = 0.f; ak2[c][b] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) { rv1[c][b] = Rp[(GP + c) * RP + m1]; rv2[c][b] = Rp[(GP + c) * RP + m2]; av1[c][b] = 0.f; av2[c][b] = 0.f; }
    }
  }
  // column windows: block cb = (ib - rt) mod LB in hi, block cb - 1 in lo; the accumulators chi / clo ride with them
  float khi[D][RT], klo[D][RT], vhi[GP][RT], vlo[GP][RT], chi[QC][RT], clo[QC][RT];
  int cb = (LB - rt) % LB;
  {
    const int pb = (cb + LB - 1) % LB;
#pragma unroll
    for (int c = 0; c < D; ++c) { ldv<RT>(khi[c], Yb + (D + c) * L + RT * cb); ldv<RT>(klo[c], Yb + (D + c) * L + RT * pb); }
#pragma unroll
    for (int c = 0; c < GP; ++c) { ldv<RT>(vhi[c], Yb + (GP + c) * L + RT * cb); ldv<RT>(vlo[c], Yb + (GP + c) * L + RT * pb); }
  }
#pragma unroll
  for (int c = 0; c < QC; ++c)
#pragma unroll
    for (int k = 0; k < RT; ++k) { chi[c][k] = 0.f; clo[c][k] = 0.f; }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;

  // column accumulators of block `blk` -> dYb rows [D, 4D)   (hi: slots 0..RT-1; lo: slots 1..RT-1)
  auto flush_cols = [&](const float (&cacc)[QC][RT], int blk, int first) {
#pragma unroll
    for (int c = 0; c < QC; ++c)
#pragma unroll
      for (int k = 0; k < RT; ++k)
        if (k >= first) dYb[(D + c) * L + RT * blk + k] += cacc[c][k];
  };

  for (int ib = 0; ib < LB; ++ib) {
    float q[D][RT], dsv[GP][RT], dse[GP][RT], ls[RT], Dd[RT], u[NV];
#pragma unroll
    for (int c = 0; c < D; ++c) ldv<RT>(q[c], Yb + c * L + RT * ib);
#pragma unroll
    for (int c = 0; c < GP; ++c) {
      ldv<RT>(dsv[c], dOb + (c * NO) * L + RT * ib);
      if (POS) ldv<RT>(dse[c], dOb + (c * NO + 1) * L + RT * ib);
    }
    ldv<RT>(ls, lsp + RT * ib); ldv<RT>(Dd, Dp + RT * ib);
#pragma unroll
    for (int e = 0; e < NV; ++e) u[e] = 0.f;
    const int mode = ib < rt ? 0 : (ib == rt ? 1 : 2);
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
      for (int b = 0; b < RT; ++b) {
        const bool s1 = mode == 2 || (mode == 1 && a >= b);
        const int d = a - b;
        float sqk = 0.f, sqr = 0.f, skr = 0.f, dP = 0.f, kc[D], rq[D], rk[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          kc[c] = d >= 0 ? khi[c][d >= 0 ? d : 0] : klo[c][d < 0 ? RT + d : 0];
          sqk = fmaf(q[c][a], kc[c], sqk);
          if (POS) {
            rq[c] = s1 ? rq1[c][b] : rq2[c][b]; rk[c] = s1 ? rk1[c][b] : rk2[c][b];
            sqr = fmaf(q[c][a], rq[c], sqr); skr = fmaf(kc[c], rk[c], skr);
          }
        }
        float z = cf.wqk * sqk;
        if (POS) z = fmaf(cf.wqr, sqr, fmaf(cf.wkr, skr, z));
        const float P = ex2(z - ls[a]);
#pragma unroll
        for (int c = 0; c < GP; ++c) {
          const float vc = d >= 0 ? vhi[c][d >= 0 ? d : 0] : vlo[c][d < 0 ? RT + d : 0];
          dP = fmaf(dsv[c][a], vc, dP);
          if (POS) dP = fmaf(dse[c][a], s1 ? rv1[c][b] : rv2[c][b], dP);
        }
        const float dZ = P * (dP - Dd[a]);
        a0 = fmaf(dZ, sqk, a0);
        if (POS) { a1 = fmaf(dZ, sqr, a1); a2 = fmaf(dZ, skr, a2); }
        const float dZ1 = dZ * cf.A1, dZ2 = dZ * cf.A2, dZ3 = dZ * cf.A3;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          // row i = RT*ib + a:  sum_j dS_qk k + dS_qr Rq  (dZ parts)
          u[c * RT + a] = fmaf(dZ1, kc[c], u[c * RT + a]);
          if (POS) u[c * RT + a] = fmaf(dZ2, rq[c], u[c * RT + a]);
          // column j:  sum_i dS_qk q + dS_kr Rk
          float ck = dZ1 * q[c][a];
          if (POS) ck = fmaf(dZ3, rk[c], ck);
          if (d >= 0) chi[c][d >= 0 ? d : 0] += ck; else clo[c][d < 0 ? RT + d : 0] += ck;
          if (POS) {
            if (s1) { aq1[c][b] = fmaf(dZ, q[c][a], aq1[c][b]); ak1[c][b] = fmaf(dZ, kc[c], ak1[c][b]); }
            else    { aq2[c][b] = fmaf(dZ, q[c][a], aq2[c][b]); ak2[c][b] = fmaf(dZ, kc[c], ak2[c][b]); }
          }
        }
#pragma unroll
        for (int c = 0; c < GP; ++c) {
          if (d >= 0) chi[D + c][d >= 0 ? d : 0] = fmaf(P, dsv[c][a], chi[D + c][d >= 0 ? d : 0]);
          else        clo[D + c][d < 0 ? RT + d : 0] = fmaf(P, dsv[c][a], clo[D + c][d < 0 ? RT + d : 0]);
          if (POS) {
            if (s1) av1[c][b] = fmaf(P, dse[c][a], av1[c][b]);
            else    av2[c][b] = fmaf(P, dse[c][a], av2[c][b]);
          }
        }
      }
    // ---- rows: total over the group's lanes (all diagonals = all columns of these RT rows)
    {
      const int idx = reduce_scatter<NV>(u, lig, NL);
      if (act && lig < NV) dYb[(idx / RT) * L + RT * ib + (idx % RT)] += u[0];
    }
    if (ib + 1 < LB) {
      // ---- columns: what this lane gathered for blocks cb / cb-1 moves to the lane that meets them next
      if (!ring) {
        __syncwarp();
        if (act && lig == NL - 1) { flush_cols(chi, cb, 0); if (RT > 1) flush_cols(clo, (cb + LB - 1) % LB, 1); }
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < QC; ++c)
#pragma unroll
        for (int k = 0; k < RT; ++k) {
          chi[c][k] = __shfl_sync(full, chi[c][k], src);
          if (RT > 1 && k >= 1) clo[c][k] = __shfl_sync(full, clo[c][k], src);
          if (!ring && lig == 0) { chi[c][k] = 0.f; clo[c][k] = 0.f; }
        }
      cb = cb + 1 == LB ? 0 : cb + 1;
#pragma unroll
      for (int c = 0; c < D; ++c) {
#pragma unroll
        for (int k = 0; k < RT; ++k) klo[c][k] = khi[c][k];
        ldv<RT>(khi[c], Yb + (D + c) * L + RT * cb);
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) {
#pragma unroll
        for (int k = 0; k < RT; ++k) vlo[c][k] = vhi[c][k];
        ldv<RT>(vhi[c], Yb + (GP + c) * L + RT * cb);
      }
    }
  }
  // ---- flush: column accumulators in flight (hi blocks are distinct per lane, lo blocks too; hi and lo of two lanes overlap)
  __syncwarp();
  if (act) flush_cols(chi, cb, 0);
  __syncwarp();
  if (act && RT > 1) flush_cols(clo, (cb + LB - 1) % LB, 1);
  __syncwarp();
  if (POS && act) {
#pragma unroll
    for (int b = 0; b < RT; ++b) {
      const int r = r0 + b, m1 = r + L - 1, m2 = r - 1;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        atomicAdd(dRs + c * RL + m1, cf.A2 * aq1[c][b]); atomicAdd(dRs + (D + c) * RL + (2 * (L - 1) - m1), cf.A3 * ak1[c][b]);
        if (m2 >= 0) { atomicAdd(dRs + c * RL + m2, cf.A2 * aq2[c][b]); atomicAdd(dRs + (D + c) * RL + (2 * (L - 1) - m2), cf.A3 * ak2[c][b]); }
      }
#pragma unroll
      for (int c = 0; c < GP; ++c) {
        atomicAdd(dRs + (GP + c) * RL + m1, av1[c][b]);
        if (m2 >= 0) atomicAdd(dRs + (GP + c) * RL + m2, av2[c][b]);
      }
    }
  }
  if (act) { t0 += a0; t1 += a1; t2 += a2; }
}

}  // namespace mtf
