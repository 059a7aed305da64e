// axial_coop.cuh -- persistent cooperative ("one launch per op direction") kernels: shared plan, grid barrier, launch.
//
// Round 1 ran one axial op as 4 forward + 5 backward launches; each of the three batch-statistics BatchNorms inside the
// op (reference axialnet.py:151,167,179) was a kernel boundary, and y_raw / o_raw / g_s / dy went through L2/HBM between
// them (ncu: 6.1x the algorithmic DRAM bytes, ~10 us fixed cost per launch).  Here ONE cooperative launch walks all
// phases; the grid-wide dependency of each BatchNorm is a grid barrier (~0.5 us: one L2 atomic + an acquire spin), and a
// CTA keeps the q/k/v rows of its strips RESIDENT in shared memory across the phases whenever its share of the op fits
// (always at 128 px; at 512 px the kernel streams units through one slot and re-reads its own y_raw from L2).
//
// Work unit = SPU consecutive strips x one head group (8/HS heads).  Heads are independent from the projection onwards
// (output channel o = g*2gp + r belongs to head g, axialnet.py:152), so small-batch layers (B = 64 strips at L = 16) are
// split over heads to fill the 148 SMs.  Unit u = strip_block*HS + head_group runs on CTA u % G; G is a multiple of HS,
// so a CTA always serves the same head group and stages its slice of W_qkv once.
#pragma once
#include "axial_attn2.cuh"
#include "pdl.cuh"

#define MT_COOP_THREADS 256

template <> struct VecT<1> { typedef float T; };

// divisor known on the host: x / d as ONE multiply-high, branch free:  q = umulhi(x, mul) + (x & one)
//   d = 1: mul = 0, one = ~0;   d = 2^k: mul = 2^(32-k) (exact for every x);   else mul = floor(2^32 / d) + 1, exact while x * d < 2^32
// (x is an index into a shared-memory slot or a unit count: < 2^18, d <= L * rows <= 2^12).  The earlier form -- shift when d is a
// power of two, else a division -- cost a compare, a branch and a call sequence at each of the ~80 sites of a cooperative kernel: 6 %
// of the code of kernels whose top stall is instruction fetch.
struct DivS { int d, lg; unsigned mul, one; };
static inline DivS make_divs(int d) {
  DivS k; k.d = d > 0 ? d : 1; k.lg = -1;
  for (int b = 0; b < 31; ++b) if ((1 << b) == k.d) k.lg = b;
  if (k.d == 1) { k.mul = 0u; k.one = 0xffffffffu; }
  else { k.mul = (unsigned)((1ull << 32) / (unsigned)k.d) + (k.lg >= 0 ? 0u : 1u); k.one = 0u; }
  return k;
}
__device__ __forceinline__ int dv(int x, const DivS k) { return (int)(__umulhi((unsigned)x, k.mul) + ((unsigned)x & k.one)); }
__device__ __forceinline__ int md(int x, const DivS k) { return x - dv(x, k) * k.d; }

// BN coefficients of one channel straight from the fp64 accumulators (training) or the running statistics (eval): ONE
// out-of-line copy shared by every phase of the cooperative kernels
__device__ __forceinline__ BnCoef coef_from_stats(const Dims& m, const double* stats, int grp, int off_sum, int off_sq, int ch,
                                                      double inv_n, float f, const float* gamma, const float* beta,
                                                      const float* rm, const float* rv) {
  const bool train = m.training != 0;
  return bn_coef(m.training, train ? stat_get(m, stats, grp, m.nstat_f, off_sum + ch) : 0.0,
                 train ? stat_get(m, stats, grp, m.nstat_f, off_sq + ch) : 0.0, inv_n, f, gamma[ch], beta[ch],
                 train ? 0.f : rm[ch], train ? 1.f : rv[ch], m.eps);
}

struct CoopPlan {
  DivS dL, dL4, dLB, dKS, dHPU, dSPU, dRU, dNOCU, dPP;
  int G;                 // CTAs (all co-resident: cooperative launch)
  int SPU, HS, HPU;      // strips per unit, head split, heads per unit
  int RU, CU, NOCU;      // rows of y (q|k|v of the unit's heads), output channels, o_raw channels per unit
  int RT, KS, LB;        // rows per thread, key split (lanes sharing a row block), L / RT
  int XP;                // SPU * L positions per unit
  int nunits, upc, nslot, resident, contig;
  int tm, tn;            // projection GEMM thread tile
  int tpr_q, tpr_o;      // lanes per row of the per-channel statistics passes
  int WTP;               // pitch of the transposed W slice
  // shared-memory carve (float offsets)
  int o_ca, o_cb, o_wsim, o_oa, o_ob, o_redq, o_reds, o_redo, o_tab, o_rp, o_rs, o_wt, o_x, o_slot;
  int ys_fl, slot_fl, tab_fl;
  size_t smem_bytes;
};

// ---------------------------------------------------------------------------------------------------------------
// Grid barrier of a cooperative launch.  `bar` is a zero-initialised counter in the op's own buffer (zeroed with the
// statistic accumulators); epoch e is complete when the counter reaches e * nctas.  Arrive = fence + relaxed atomic,
// wait = acquire load spin by one thread; a lost CTA traps instead of hanging the device.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nctas, unsigned& epoch) {
  ++epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = epoch * nctas;
    unsigned v = 0;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > (1ll << 31)) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- Ampere-style async copies (SASS LDGSTS): global -> shared without a register round trip, so the gathers of a unit's
// strips overlap the table / coefficient staging that follows their issue
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// cooperative launch; arguments are passed as lvalues of exactly the kernel's parameter types.
// (the CPU emulation tier of tests/emu pre-defines this macro with its own co-scheduled launch)
#ifndef MT_COOP_LAUNCH
template <typename... A>
static inline cudaError_t coop_launch_(const void* k, dim3 g, dim3 b, size_t smem, cudaStream_t st, A&... a) {
  void* args[] = {(void*)&a...};
  return cudaLaunchCooperativeKernel(k, g, b, args, smem, st);
}
#define MT_COOP_LAUNCH(kern, grid, block, smem, stream, ...) \
  (g_last_cuda = coop_launch_((const void*)(kern), dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__))
#endif

// ---- developer timeline (MT_DEBUG_FLAGS bit6): thread 0 of every CTA stamps %globaltimer at the phase boundaries of the
// cooperative kernels into a device array that mt_debug_read_phases() copies out (tools/time_ops.py --phases)
#define MT_PH_MAX 16
#define MT_PH_CTAS 512
static __device__ unsigned long long g_phase_ns[MT_PH_MAX * MT_PH_CTAS];      // one copy per translation unit
__device__ __forceinline__ void phase_stamp(const Dims& m, int ph) {
  if ((m.dbg & 64) && threadIdx.x == 0 && blockIdx.x < MT_PH_CTAS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_phase_ns[ph * MT_PH_CTAS + blockIdx.x] = t;
  }
}

int coop_sm_count();
int coop_max_ctas_env();
