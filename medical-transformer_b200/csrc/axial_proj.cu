// axial_proj.cu -- the three projection GEMMs of the axial op as one register-tiled fp32
// SGEMM core (64 x 64 output tile, 256 threads, 4 x 4 per thread, K chunks of 16):
//
//   k_proj       Y[2C x P]  = W[2C x C] . X[C x P]        (+ fp64 sum / sumsq per output channel)
//   k_bwd_dx     dX[C x P]  = W^T[C x 2C] . dY[2C x P]     (bn_qkv backward applied while loading dY)
//   k_bwd_dw     dW[2C x C] += dY[2C x Pk] . X^T[Pk x C]   (split over position chunks, atomics)
//
// P = B*L positions.  X lives in NCHW (reference conv1d on permuted strips, axialnet.py:143-151),
// Y / dY in strip-major [B][2C][L]; a position tile is TS strips x TL consecutive l so that both
// sides move in >= 32-byte runs.  fp32 FFMA on purpose: tf32/bf16 operands break the 1e-2
// train-mode parity budget (DESIGN.md "precision"); a 3xTF32 tcgen05 version is future work.
#include "common.cuh"

namespace {

struct PosTile {           // decode of a 64-position tile
  int grp, sb0, l0, TS, TL, npos;
};

__device__ __forceinline__ PosTile pos_tile(const Dims& m, int tile, int TS, int TL, int tiles_per_group, int nTL) {
  PosTile t;
  t.grp = tile / tiles_per_group;
  const int tin = tile - t.grp * tiles_per_group;
  const int ts_idx = tin / nTL, tl_idx = tin - ts_idx * nTL;
  t.sb0 = ts_idx * TS; t.l0 = tl_idx * TL; t.TS = TS; t.TL = TL; t.npos = TS * TL;
  return t;
}

// position slot r (0..63) -> (strip-in-tile, l-in-tile); "x order" is NCHW friendly for axis 0
__device__ __forceinline__ void slot_x(const Dims& m, const PosTile& t, int r, int& s, int& l) {
  if (m.axis == 0) { s = r % t.TS; l = r / t.TS; } else { l = r % t.TL; s = r / t.TL; }
}
__device__ __forceinline__ bool slot_ok(const Dims& m, const PosTile& t, int r, int s, int l) {
  return r < t.npos && t.sb0 + s < m.Bg && t.l0 + l < m.L;
}

// 64x64x16 micro-kernel step on smem tiles As[16][64+pad] (k-major), Bs[16][64+pad]
#define TILE_LD 68
__device__ __forceinline__ void mma_chunk(const float* As, const float* Bs, int ty, int tx, float acc[4][4], int kc) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (k >= kc) break;
    const float4 a = *reinterpret_cast<const float4*>(As + k * TILE_LD + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bs + k * TILE_LD + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

}  // namespace

// ============================================================================ k_proj
// grid (position tiles, ceil(2C/64)).  M = oc (ty), N = position (tx).
__global__ void __launch_bounds__(256)
k_proj(Dims m, MtAxialParams p, const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ yraw,
       double* __restrict__ stats, float* __restrict__ coef, unsigned* __restrict__ cnt, int TS, int TL,
       int tiles_per_group, int nTL) {
  __shared__ __align__(16) float As[16 * TILE_LD];
  __shared__ __align__(16) float Bs[16 * TILE_LD];
  __shared__ float red[64][2];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int C = m.C, M = 2 * C;
  const PosTile t = pos_tile(m, blockIdx.x, TS, TL, tiles_per_group, nTL);
  const int oc0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (tid < 64) { red[tid][0] = 0.f; red[tid][1] = 0.f; }
  const size_t plane = (size_t)m.H * m.W;
  int xslot; bool xok; const float* xb;
  {
    const int r = tid & 63;
    int s, l;
    slot_x(m, t, r, s, l);
    xok = slot_ok(m, t, r, s, l);
    xslot = r < t.npos ? s * t.TL + l : r;
    xb = x + (xok ? nchw_addr(m, t.grp * m.Bg + t.sb0 + s, 0, t.l0 + l) : 0);
  }
  for (int k0 = 0; k0 < C; k0 += 16) {
    // A chunk: W[oc0 + mm][k0 + kk] -> As[kk][mm]   (each thread: one float4 of 4 consecutive k)
    {
      const int mm = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oc0 + mm < M && k0 + kq < C) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(oc0 + mm) * C + k0 + kq));
      As[(kq + 0) * TILE_LD + mm] = v.x; As[(kq + 1) * TILE_LD + mm] = v.y;
      As[(kq + 2) * TILE_LD + mm] = v.z; As[(kq + 3) * TILE_LD + mm] = v.w;
    }
    // B chunk: X[k0 + kk][pos r] gathered from NCHW in x order, stored in strip order (l fastest).
    // r = tid & 63 is the same for all four loads of this thread: address math hoisted (xb, xok, xslot)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int kk = (tid >> 6) + 4 * it;
      float v = 0.f;
      if (xok && k0 + kk < C) v = __ldg(xb + (size_t)(k0 + kk) * plane);
      Bs[kk * TILE_LD + xslot] = v;
    }
    __syncthreads();
    mma_chunk(As, Bs, ty, tx, acc, C - k0 < 16 ? C - k0 : 16);
    __syncthreads();
  }
  // epilogue: thread owns oc = oc0 + ty*4 + i, positions tx*4 + j (strip order)
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = tx * 4 + j, s = pos / t.TL, l = pos - s * t.TL;
    const bool ok = pos < t.npos && t.sb0 + s < m.Bg && t.l0 + l < m.L;
    if (!ok) continue;
    const int b = t.grp * m.Bg + t.sb0 + s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oc = oc0 + ty * 4 + i;
      if (oc < M) {
        yraw[((size_t)b * M + oc) * m.L + t.l0 + l] = acc[i][j];
        s1[i] += acc[i][j]; s2[i] = fmaf(acc[i][j], acc[i][j], s2[i]);
      }
    }
  }
  if (m.training) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // reduce over the 16 tx lanes that share this oc
      float a = seg_sum(s1[i], 16), b2 = seg_sum(s2[i], 16);
      if (tx == 0) { red[ty * 4 + i][0] = a; red[ty * 4 + i][1] = b2; }
    }
    __syncthreads();
    if (tid < 64 && oc0 + tid < M) {
      double* st = stat_slot(m, stats, t.grp, m.nstat_f, blockIdx.x);
      MT_GADD(m, st + m.f_qkv_sum + oc0 + tid, (double)red[tid][0]);
      MT_GADD(m, st + m.f_qkv_sq + oc0 + tid, (double)red[tid][1]);
    }
  }
  // last CTA: bn_qkv coefficients (+ running stats); in eval mode also the similarity weights,
  // because k_sim_stats is not launched then
  if (last_cta_done(cnt + 0, gridDim.x * gridDim.y)) {
    fin_qkv(m, p, stats, coef);
    if (!m.training) fin_sim(m, p, stats, coef);
  }
}

// ============================================================================ k_bwd_dx
// grid (position tiles, ceil(C/64)).  M = c (ty), N = position (tx), K = 2C.
__global__ void __launch_bounds__(256)
k_bwd_dx(Dims m, MtAxialParams p, MtAxialGrads gr, const float* __restrict__ bcoef,
         const float* __restrict__ yraw, const float* __restrict__ dyh, int TS, int TL, int tiles_per_group, int nTL) {
  extern __shared__ __align__(16) float smx[];
  float* As = smx;                          // [16][TILE_LD]
  float* Bs = As + 16 * TILE_LD;            // [16][TILE_LD]
  float* qA = Bs + 16 * TILE_LD;            // [2C] x 3 bn_qkv backward coefficients
  float* qB = qA + 2 * m.C;
  float* qC = qB + 2 * m.C;
  float* Ts = qC + 2 * m.C;                 // [64][65] output staging (transposed store)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int C = m.C, K = 2 * C;
  const PosTile t = pos_tile(m, blockIdx.x, TS, TL, tiles_per_group, nTL);
  const int c0 = blockIdx.y * 64;
  {
    const float* tb = bcoef + (size_t)t.grp * m.ncoef_b + m.cb_q;
    for (int ch = tid; ch < K; ch += 256) { qA[ch] = tb[3 * ch]; qB[ch] = tb[3 * ch + 1]; qC[ch] = tb[3 * ch + 2]; }
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool yok; size_t ybase;
  {
    const int r = tid & 63, s = r / t.TL, l = r - s * t.TL;
    yok = r < t.npos && t.sb0 + s < m.Bg && t.l0 + l < m.L;
    ybase = yok ? (size_t)(t.grp * m.Bg + t.sb0 + s) * K * m.L + t.l0 + l : 0;
  }
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A chunk: W^T: As[kk][mm] = W[k0+kk][c0+mm]  (row of W is contiguous in c)
    {
      const int kk = tid >> 4, mq = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < K && c0 + mq < C) v = __ldg(reinterpret_cast<const float4*>(p.w_qkv + (size_t)(k0 + kk) * C + c0 + mq));
      *reinterpret_cast<float4*>(As + kk * TILE_LD + mq) = v;
    }
    // B chunk: dy_raw[k0+kk][pos] = A*dyh + Bc + Cc*yraw, strip order (contiguous TL-runs); r = tid & 63 fixed
    {
      float vd[4], vy[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int ch = k0 + (tid >> 6) + 4 * it;
        const bool ok = yok && ch < K;
        vd[it] = ok ? __ldg(dyh + ybase + (size_t)ch * m.L) : 0.f;
        vy[it] = ok ? __ldg(yraw + ybase + (size_t)ch * m.L) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int kk = (tid >> 6) + 4 * it, ch = k0 + kk;
        Bs[kk * TILE_LD + (tid & 63)] = (yok && ch < K) ? fmaf(qA[ch], vd[it], fmaf(qC[ch], vy[it], qB[ch])) : 0.f;
      }
    }
    __syncthreads();
    mma_chunk(As, Bs, ty, tx, acc, K - k0 < 16 ? K - k0 : 16);
    __syncthreads();
  }
  // stage (c, pos) tile and store to NCHW with x-order threads (coalesced along the contiguous axis)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ts[(ty * 4 + i) * 65 + tx * 4 + j] = acc[i][j];
  __syncthreads();
  if (gr.dx) {
    const int r = tid & 63;
    int s, l;
    slot_x(m, t, r, s, l);
    if (slot_ok(m, t, r, s, l)) {
      float* dst = gr.dx + nchw_addr(m, t.grp * m.Bg + t.sb0 + s, 0, t.l0 + l);
      const size_t plane = (size_t)m.H * m.W;
      const int slot = s * t.TL + l;
#pragma unroll 4
      for (int cc = tid >> 6; cc < 64; cc += 4)
        if (c0 + cc < C) dst[(size_t)(c0 + cc) * plane] = Ts[cc * 65 + slot];
    }
  }
}

// ============================================================================ k_bwd_dw
// grid (ceil(2C/64), ceil(C/64), KSPLIT).  M = oc (ty), N = c (tx), K = positions of this split.
__global__ void __launch_bounds__(256)
k_bwd_dw(Dims m, MtAxialGrads gr, const float* __restrict__ x, const float* __restrict__ bcoef,
         const float* __restrict__ yraw, const float* __restrict__ dyh, int TS, int TL, int tiles_per_group, int nTL,
         int ntiles) {
  extern __shared__ __align__(16) float smw[];
  float* Am = smw;                          // [64 pos][TILE_LD]  (k-major: As[k][m] with k = pos)
  float* Bm = Am + 64 * TILE_LD;            // [64 pos][TILE_LD]
  float* qA = Bm + 64 * TILE_LD;
  float* qB = qA + 2 * m.C;
  float* qC = qB + 2 * m.C;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int C = m.C, M = 2 * C;
  const int oc0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int cur_grp = -1;
  for (int tile = blockIdx.z; tile < ntiles; tile += gridDim.z) {
    const PosTile t = pos_tile(m, tile, TS, TL, tiles_per_group, nTL);
    if (t.grp != cur_grp) {                 // bn_qkv backward coefficients of this statistic group
      __syncthreads();
      {
        const float* tb = bcoef + (size_t)t.grp * m.ncoef_b + m.cb_q;
        for (int ch = tid; ch < M; ch += 256) { qA[ch] = tb[3 * ch]; qB[ch] = tb[3 * ch + 1]; qC[ch] = tb[3 * ch + 2]; }
      }
      cur_grp = t.grp;
      __syncthreads();
    }
    // A: dy_raw[oc0+mm][pos r] -> Am[r][mm]; B: X[c0+nn][pos r] -> Bm[slot][nn].  r = tid & 63 is fixed
    // per thread, so all address math is done once per tile and 4 + 4 loads are issued back to back.
    {
      const int r = tid & 63, sy = r / t.TL, ly = r - sy * t.TL;
      const bool yok = r < t.npos && t.sb0 + sy < m.Bg && t.l0 + ly < m.L;
      const size_t ybase = yok ? (size_t)(t.grp * m.Bg + t.sb0 + sy) * M * m.L + t.l0 + ly : 0;
      int sx, lx;
      slot_x(m, t, r, sx, lx);
      const bool xok = slot_ok(m, t, r, sx, lx);
      const float* xb = x + (xok ? nchw_addr(m, t.grp * m.Bg + t.sb0 + sx, 0, t.l0 + lx) : 0);
      const size_t plane = (size_t)m.H * m.W;
      const int xslot = r < t.npos ? sx * t.TL + lx : r;
      for (int m0 = tid >> 6; m0 < 64; m0 += 16) {
        float vd[4], vy[4], vx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int mm = m0 + 4 * u, ch = oc0 + mm, cc = c0 + mm;
          const bool oky = yok && ch < M, okx = xok && cc < C;
          vd[u] = oky ? __ldg(dyh + ybase + (size_t)ch * m.L) : 0.f;
          vy[u] = oky ? __ldg(yraw + ybase + (size_t)ch * m.L) : 0.f;
          vx[u] = okx ? __ldg(xb + (size_t)cc * plane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int mm = m0 + 4 * u, ch = oc0 + mm;
          Am[r * TILE_LD + mm] = (yok && ch < M) ? fmaf(qA[ch], vd[u], fmaf(qC[ch], vy[u], qB[ch])) : 0.f;
          Bm[xslot * TILE_LD + mm] = vx[u];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kb = 0; kb < 64; kb += 16) mma_chunk(Am + kb * TILE_LD, Bm + kb * TILE_LD, ty, tx, acc, 16);
    __syncthreads();
  }
  if (gr.dw_qkv) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int oc = oc0 + ty * 4 + i, c = c0 + tx * 4 + j;
        if (oc < M && c < C) MT_GADD(m, gr.dw_qkv + (size_t)oc * C + c, acc[i][j]);
      }
  }
}

// ============================================================================ host launchers
static void proj_tiling(const Dims& m, int& TS, int& TL, int& nTL, int& nTS) {
  if (m.axis == 1) { TL = m.L < 64 ? m.L : 64; TS = 64 / TL; if (TS > m.Bg) TS = m.Bg; }
  else { TS = m.Bg < 8 ? m.Bg : 8; TL = 64 / TS; if (TL > m.L) TL = m.L; }
  if (TS < 1) TS = 1;
  nTL = (m.L + TL - 1) / TL; nTS = (m.Bg + TS - 1) / TS;
}

int launch_proj(const Dims& m, const MtAxialParams& p, const float* x, float* yraw, double* stats, float* coef,
                unsigned* cnt, cudaStream_t st) {
  int TS, TL, nTL, nTS;
  proj_tiling(m, TS, TL, nTL, nTS);
  dim3 grid(m.ngroups * nTS * nTL, (2 * m.C + 63) / 64);
  k_proj<<<grid, 256, 0, st>>>(m, p, x, p.w_qkv, yraw, stats, coef, cnt, TS, TL, nTS * nTL, nTL);
  return 1;
}

cudaStream_t mt_side_fork(cudaStream_t main_stream);
void mt_side_join(cudaStream_t main_stream, cudaStream_t forked);

int launch_bwd_proj(const Dims& m, const MtAxialParams& p, const MtAxialGrads& gr, const float* x, const float* bcoef,
                    const float* yraw, const float* dyh, cudaStream_t st) {
  int TS, TL, nTL, nTS;
  proj_tiling(m, TS, TL, nTL, nTS);
  const int ntiles = m.ngroups * nTS * nTL;
  cudaStream_t sd = st;
  {
    // dW only depends on what k_bwd2 left behind, exactly like dX: run it on the forked side stream
    sd = gr.dw_qkv ? mt_side_fork(st) : st;          // fork BEFORE dX is enqueued so the two overlap
    const int mn = ((2 * m.C + 63) / 64) * ((m.C + 63) / 64);
    int ksplit = (148 + mn - 1) / mn;                 // fill the 148 SMs: CTAs = Mtiles * Ntiles * ksplit
    if (ksplit > ntiles) ksplit = ntiles;
    dim3 grid((2 * m.C + 63) / 64, (m.C + 63) / 64, ksplit);
    const size_t smem = sizeof(float) * (2 * 64 * TILE_LD + 6 * (size_t)m.C);
    static bool once = (cudaFuncSetAttribute(k_bwd_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), true);
    (void)once;
    if (gr.dw_qkv) cudaMemsetAsync(gr.dw_qkv, 0, sizeof(float) * 2 * (size_t)m.C * m.C, sd);
    k_bwd_dw<<<grid, 256, smem, sd>>>(m, gr, x, bcoef, yraw, dyh, TS, TL, nTS * nTL, nTL, ntiles);
  }
  {
    dim3 grid(ntiles, (m.C + 63) / 64);
    const size_t smem = sizeof(float) * (2 * 16 * TILE_LD + 6 * (size_t)m.C + 64 * 65);
    k_bwd_dx<<<grid, 256, smem, st>>>(m, p, gr, bcoef, yraw, dyh, TS, TL, nTS * nTL, nTL);
  }
  mt_side_join(st, sd);
  return 2;
}

// weight gradient only (the cooperative backward kernel produces dX itself): dW = dY_raw X^T on the forked side stream
int launch_bwd_dw(const Dims& m, const MtAxialGrads& gr, const float* x, const float* bcoef, const float* yraw,
                  const float* dyh, cudaStream_t st) {
  if (!gr.dw_qkv) return 0;
  int TS, TL, nTL, nTS;
  proj_tiling(m, TS, TL, nTL, nTS);
  const int ntiles = m.ngroups * nTS * nTL;
  cudaStream_t sd = mt_side_fork(st);
  const int mn = ((2 * m.C + 63) / 64) * ((m.C + 63) / 64);
  int ksplit = (148 + mn - 1) / mn;
  if (ksplit > ntiles) ksplit = ntiles;
  dim3 grid((2 * m.C + 63) / 64, (m.C + 63) / 64, ksplit);
  const size_t smem = sizeof(float) * (2 * 64 * TILE_LD + 6 * (size_t)m.C);
  static bool once = (cudaFuncSetAttribute(k_bwd_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), true);
  (void)once;
  cudaMemsetAsync(gr.dw_qkv, 0, sizeof(float) * 2 * (size_t)m.C * m.C, sd);
  k_bwd_dw<<<grid, 256, smem, sd>>>(m, gr, x, bcoef, yraw, dyh, TS, TL, nTS * nTL, nTL, ntiles);
  mt_side_join(st, sd);
  return 1;
}
