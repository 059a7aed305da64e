// axial_attn2.cuh -- register-tiled building blocks of the attention passes ("v2").
//
// A thread owns RT consecutive rows (or columns) and walks the other index in blocks of RT, so
//   * q/k/v/dO values of the walked index are one 32*RT-bit shared-memory load per RT pairs;
//   * the Toeplitz tables relative[c, i-j+L-1] are read through SLIDING REGISTER WINDOWS: the
//     RT x RT pair block touches only 2*RT-1 distinct diagonals, of which RT are new per step
//     (one aligned vector load), instead of one scalar load per pair and per table row.
// ncu on the v1 kernels (profiles/r1_stage_kernels_v1.md): 45 issued warp instructions per 32
// scores, issue-bound at 21 % warps active; this layout needs ~13-20.
#pragma once
#include "common.cuh"

template <int RT> struct VecT;
template <> struct VecT<4> { typedef float4 T; };
template <> struct VecT<2> { typedef float2 T; };

template <int RT>
__device__ __forceinline__ void ldv(float (&d)[RT], const float* p) {
  typedef typename VecT<RT>::T V;
  const V v = *reinterpret_cast<const V*>(p);
  const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
  for (int k = 0; k < RT; ++k) d[k] = f[k];
}

// window over a table indexed by  own - walk + L - 1  (own = RT*t + r, walk = RT*wb + w):
//   value(r, w) = T[RT*K - 1 + (r - w)],  K = t + L/RT - wb;  walking forward decrements K.
template <int RT>
struct WinA {
  float hi[RT], lo[RT];
  __device__ __forceinline__ void init(const float* row, int K) { ldv<RT>(hi, row + RT * K); ldv<RT>(lo, row + RT * (K - 1)); }
  __device__ __forceinline__ void advance(const float* row, int Knew) {
#pragma unroll
    for (int k = 0; k < RT; ++k) hi[k] = lo[k];
    ldv<RT>(lo, row + RT * (Knew - 1));
  }
  __device__ __forceinline__ float get(int r, int w) const { const int d = r - w; return d >= 1 ? hi[d - 1] : lo[RT - 1 + d]; }
};

// window over a table indexed by  walk - own + L - 1, stored SHIFTED by one (Ts[x] = T[x-1]):
//   value(r, w) = Ts[RT*K + (w - r)],  K = wb - t + L/RT;  walking forward increments K.
template <int RT>
struct WinB {
  float lo[RT], hi[RT];
  __device__ __forceinline__ void init(const float* row, int K) { ldv<RT>(hi, row + RT * K); ldv<RT>(lo, row + RT * (K - 1)); }
  __device__ __forceinline__ void advance(const float* row, int Knew) {
#pragma unroll
    for (int k = 0; k < RT; ++k) lo[k] = hi[k];
    ldv<RT>(hi, row + RT * Knew);
  }
  __device__ __forceinline__ float get(int r, int w) const { const int d = w - r; return d >= 0 ? hi[d] : lo[RT + d]; }
};

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) staging of the q/k/v strips.  A CTA's strips are one contiguous block of
// SPB*2C*L floats in y_raw (strip-major layout) and, when LP == L, also in shared memory, so ONE
// cp.async.bulk (SASS: UBLKCP) issued by thread 0 at kernel start brings them in while the other
// threads load coefficient tables and Toeplitz tables; completion is signalled on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_strips_issue(uint64_t* bar, float* dst, const float* src, uint32_t bytes) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_strips_wait(uint64_t* bar) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();          // never hang the GPU on a lost copy
  }
}
// in-place BN affine on the staged block: row r uses (a, b)[r % ncoef]
__device__ __forceinline__ void affine_rows_inplace(float* Yn, int nrows, int L, const float* a, const float* b, int ncoef) {
  const int L4 = L >> 2, total4 = nrows * L4;
  float4* y4 = reinterpret_cast<float4*>(Yn);
  for (int i4 = threadIdx.x; i4 < total4; i4 += blockDim.x) {
    const int rc = (i4 / L4) % ncoef;
    const float ka = a[rc], kb = b[rc];
    float4 v = y4[i4];
    v.x = fmaf(ka, v.x, kb); v.y = fmaf(ka, v.y, kb); v.z = fmaf(ka, v.z, kb); v.w = fmaf(ka, v.w, kb);
    y4[i4] = v;
  }
}

// shared-memory pitch of one Toeplitz table row in the v2 layout (plain copy; the shifted copy
// follows at +RP): 2L entries used, padded so the last window pre-load stays in bounds
__host__ __device__ __forceinline__ int rel_pitch(int L) { return 2 * L + 8; }

// stage `relative` (2gp rows x (2L-1)) into smem twice: plain T[m] and shifted Ts[m+1]
__device__ __forceinline__ void load_rel2(const Dims& m, const float* __restrict__ rel, float* Rp, float* Rs) {
  const int RP = rel_pitch(m.L), RL = 2 * m.L - 1, rows = 2 * m.gp;
  for (int x = threadIdx.x; x < RP; x += blockDim.x) {
    const bool in0 = x < RL, in1 = x >= 1 && x <= RL;
    for (int c0 = 0; c0 < rows; c0 += 8) {          // 8 rows per batch: 16 independent loads in flight
      float a[8], b[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        a[k] = (in0 && c < rows) ? __ldg(rel + c * RL + x) : 0.f;
        b[k] = (in1 && c < rows) ? __ldg(rel + c * RL + x - 1) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        if (c < rows) { Rp[c * RP + x] = a[k]; Rs[c * RP + x] = b[k]; }
      }
    }
  }
}
