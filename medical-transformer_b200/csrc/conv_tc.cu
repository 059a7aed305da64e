// conv_tc.cu -- the KxK convolutions of the conv shell (stem conv2 / conv3, decoder convs; reference axialnet.py:424-440, :475-502)
// on the 5th-generation tensor cores: implicit GEMM, tcgen05.mma kind::tf32 with the 3xTF32 operand split, fp32 accumulators in
// tensor memory (TMEM), split-K over the GPU.
//
// Why 3xTF32 and not plain TF32: the network's 16 stacked batch-statistics BatchNorms amplify operand rounding -- cuDNN's TF32
// convolutions give 1e-1 relative logits error against the fp64 reference (bar 1e-2).  Each fp32 operand x is
// split into hi = rn_tf32(x) and lo = rn_tf32(x - hi); D += A_hi.B_lo + A_lo.B_hi + A_hi.B_hi drops only the lo.lo term
// (<= 2^-22 relative per product), i.e. fp32-grade products on the tensor pipe at 1/3 of its TF32 rate.
//
// Two kernels: the GENERIC one described next (any K x K / stride; used for decoder1's strided forward and odd shapes) and the
// 3x3 / stride-1 SLAB kernel further down (k_conv_tc3: every other stem / decoder convolution, forward and dX) -- see its header.
//
// Generic kernel: one CTA = one 128 x BN tile of D and one slice of the reduction:
//   * 256 threads.  Warps 0-3: one A row each (128 rows), warps 4-7: the BN rows of B.  Per k block of 32 they gather their
//     elements straight from the NCHW tensors (im2col is never materialised; zero padding and tails are predicated), split them
//     into hi / lo and store them as the UMMA canonical K-major SWIZZLE_NONE layout (core matrix = 8 rows x 16 B, chunk-major:
//     [k / 4][row][16 B], so that consecutive rows are consecutive 16-byte words: conflict-free 128-bit stores).
//   * two smem stages; the loads of block i+1 are in flight while the tensor core consumes block i.  Thread 0 issues the 12
//     tcgen05.mma (4 k steps of 8 x 3 split terms) of a block and commits them to the stage's mbarrier, which the loaders wait on
//     before they overwrite the stage (generic-proxy stores are published with fence.proxy.async).
//   * epilogue: all 8 warps read their TMEM lane quarter with tcgen05.ld (lane = D row), stores coalesce over the rows.
//     Split-K slices accumulate with fp32 atomics into the zeroed output (last-bit run-to-run differences, as in conv_gemm.cu).
#include "common.cuh"
#include "conv_tc.cuh"
#include "pdl.cuh"
#include <cstdlib>

namespace {

constexpr int TC_KB = 32;                 // reduction elements per smem stage
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = 128 * TC_KB * 4;          // one hi or lo tile of A

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error traps (the launch fails loudly) instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  if (tc_mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!tc_mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}

__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start address, leading-dimension byte
// offset (core matrix -> next core matrix along K), stride byte offset (8-row group -> next 8-row group), descriptor version 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Accuracy: the tensor core adds the products of one MMA to the fp32 accumulator with truncated alignment (measured: the error of
// one long accumulation chain grows like sqrt(#MMAs) * 1e-7 of the result, 5 - 20x the FMA pipes' round-to-nearest chain).  The
// MMAs are therefore dealt round-robin to NACC independent TMEM accumulators (all 512 columns: 4 at BN = 128 ... 16 at BN <= 32)
// which the epilogue adds in fp32 with round-to-nearest: each chain is NACC times shorter and sqrt(NACC) times smaller.
__host__ __device__ constexpr int tc_nacc(int bn) { return 512 / bn > 16 ? 16 : 512 / bn; }

__device__ __forceinline__ void tc_ld16_sum(uint32_t taddr, int bn, int nacc, float (&acc)[16]) {
  tc_ld16(taddr, acc);
  for (int j = 1; j < nacc; ++j) {
    float t[16];
    tc_ld16(taddr + (uint32_t)(j * bn), t);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] += t[i];
  }
}

// MODE 0: forward / dgrad (A = im2col rows, B = weights).  MODE 1: wgrad (A = im2col^T rows = taps, B = dy rows = channels).
template <int MODE, int BN>
__global__ void __launch_bounds__(256, 1)
k_conv_tc(const TcConv a, const float* __restrict__ img, const float* __restrict__ mat, float* __restrict__ out) {
  constexpr int B_BYTES = BN * TC_KB * 4;                    // one hi or lo tile of B
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  constexpr int TPR = 128 / BN;                              // B-loader threads per row
  constexpr int KPT = TC_KB / TPR;                           // reduction elements per B-loader thread and block (BN / 4)
  constexpr int NACC = tc_nacc(BN);                          // independent TMEM accumulators, summed in fp32 by the epilogue
  constexpr uint32_t TMEM_COLS = NACC * BN;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(KPT % 4 == 0 && KPT <= 32, "B tile split");

  extern __shared__ __align__(128) unsigned char tc_dyn[];
  __shared__ __align__(8) unsigned long long s_bar[TC_STAGES];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long rowA0 = (long)blockIdx.x * 128;
  const int rowB0 = blockIdx.y * BN;
  const int nkb = (a.kdim + TC_KB - 1) / TC_KB;
  const int kb0 = blockIdx.z * a.kb_per;
  const int kb1 = min(nkb, kb0 + a.kb_per);

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) tc_mbar_init(tc_smem(&s_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  const int KS = a.KS, KW = a.KW, KH = a.KH, Hi = a.Hi, Wi = a.Wi;
  const int HWo = a.Ho * a.Wo;
  const bool loaderA = tid < 128;
  const int u = tid - 128;                                   // B-loader index
  const int brow = loaderA ? 0 : u % BN, bpart = loaderA ? 0 : u / BN;

  // ---- per-thread row state -------------------------------------------------------------------------------------------------
  int r_base = 0, r_ok = 0;           // A loader MODE 0: image offset of the window's top-left tap; MODE 1: tap (ci, kh, kw)
  uint32_t r_vh = 0, r_vw = 0;        // MODE 0: valid kh / kw bit masks
  int r_ci = 0, r_kh = 0, r_kw = 0;   // MODE 1
  if (loaderA) {
    if (MODE == 0) {
      const long p = rowA0 + tid;
      r_ok = p < a.P;
      const long pp = r_ok ? p : 0;
      const int n = (int)(pp / HWo), dst = (int)(pp - (long)n * HWo);
      const int oh = dst / a.Wo, ow = dst - oh * a.Wo;
      const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
      r_base = (n * a.Cimg * Hi + ih0) * Wi + iw0;
      if (r_ok) {
        for (int k = 0; k < KH; ++k) if ((unsigned)(ih0 + k) < (unsigned)Hi) r_vh |= 1u << k;
        for (int k = 0; k < KW; ++k) if ((unsigned)(iw0 + k) < (unsigned)Wi) r_vw |= 1u << k;
      }
    } else {
      const int kk = (int)rowA0 + tid;
      r_ok = kk < a.K;
      const int kq = r_ok ? kk : 0;
      r_ci = kq / KS;
      const int r = kq - r_ci * KS;
      r_kh = r / KW; r_kw = r - r_kh * KW;
    }
  } else {
    r_ok = (rowB0 + brow) < a.rowsB;
  }

  float v[32], vn[32];

  auto load_block = [&](int kb, float (&d)[32]) {
    if (loaderA) {
      if (MODE == 0) {
        const int k = kb * TC_KB;
        int kc = k / KS;
        const int r = k - kc * KS;
        int kh = r / KW, kw = r - kh * KW;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool valid = (kc < a.Cimg) && ((r_vh >> kh) & 1u) && ((r_vw >> kw) & 1u);
          const int off = r_base + (kc * Hi + kh) * Wi + kw;
          d[j] = valid ? __ldg(img + off) : 0.f;
          if (++kw == KW) { kw = 0; if (++kh == KH) { kh = 0; ++kc; } }
        }
      } else {
        const long p0 = (long)kb * TC_KB;
        const int n = (int)(p0 >> a.lgHWo), dst0 = (int)(p0 & (HWo - 1));
        const bool okb = r_ok && p0 < a.P;
        const float* base = img + ((size_t)n * a.Cimg + r_ci) * Hi * Wi;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int dst = dst0 + j;
          const int oh = dst >> a.lgWo, ow = dst & (a.Wo - 1);
          const int ih = oh * a.stride - a.pad + r_kh, iw = ow * a.stride - a.pad + r_kw;
          const bool valid = okb && (unsigned)ih < (unsigned)Hi && (unsigned)iw < (unsigned)Wi;
          d[j] = valid ? __ldg(base + ih * Wi + iw) : 0.f;
        }
      }
    } else {
      if (MODE == 0) {
        const int m = rowB0 + brow;
        const int k = kb * TC_KB + bpart * KPT;
        int kc = k / KS, r = k - kc * KS;
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
          const bool valid = r_ok && kc < a.Cimg;
          const size_t off = a.dgrad ? ((size_t)kc * a.Wm + m) * KS + (KS - 1 - r) : (size_t)m * a.K + (size_t)(kc * KS + r);
          d[j] = valid ? __ldg(mat + off) : 0.f;
          if (++r == KS) { r = 0; ++kc; }
        }
      } else {
        const int co = rowB0 + brow;
        const long p0 = (long)kb * TC_KB + bpart * KPT;
        const int n = (int)(p0 >> a.lgHWo), dst0 = (int)(p0 & (HWo - 1));
        const bool okb = r_ok && p0 < a.P;
        const float* base = mat + ((size_t)n * a.Crow + (r_ok ? co : 0)) * HWo + dst0;
#pragma unroll
        for (int j = 0; j < KPT; ++j) d[j] = okb ? __ldg(base + j) : 0.f;
      }
    }
  };

  auto store_block = [&](int stage, const float (&d)[32]) {
    unsigned char* st = tc_dyn + (size_t)stage * STAGE_BYTES;
    if (loaderA) {
      float4* hi = reinterpret_cast<float4*>(st) + tid;                              // chunk c at + c * 128 float4
      float4* lo = reinterpret_cast<float4*>(st + TC_A_BYTES) + tid;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 h, l;
        tc_split(d[4 * c + 0], h.x, l.x); tc_split(d[4 * c + 1], h.y, l.y);
        tc_split(d[4 * c + 2], h.z, l.z); tc_split(d[4 * c + 3], h.w, l.w);
        hi[c * 128] = h; lo[c * 128] = l;
      }
    } else {
      float4* hi = reinterpret_cast<float4*>(st + 2 * TC_A_BYTES) + brow;            // chunk c at + c * BN float4
      float4* lo = reinterpret_cast<float4*>(st + 2 * TC_A_BYTES + B_BYTES) + brow;
#pragma unroll
      for (int cc = 0; cc < KPT / 4; ++cc) {
        const int c = bpart * (KPT / 4) + cc;
        float4 h, l;
        tc_split(d[4 * cc + 0], h.x, l.x); tc_split(d[4 * cc + 1], h.y, l.y);
        tc_split(d[4 * cc + 2], h.z, l.z); tc_split(d[4 * cc + 3], h.w, l.w);
        hi[c * BN] = h; lo[c * BN] = l;
      }
    }
  };

  // ---- main loop --------------------------------------------------------------------------------------------------------------
  const uint32_t lboA = (a.dbg & 1) ? 128u : 128u * 16u, sboA = (a.dbg & 1) ? 128u * 16u : 128u;
  const uint32_t lboB = (a.dbg & 1) ? 128u : (uint32_t)BN * 16u, sboB = (a.dbg & 1) ? (uint32_t)BN * 16u : 128u;
  if (kb0 < kb1) load_block(kb0, v);
  for (int kb = kb0; kb < kb1; ++kb) {
    const int it = kb - kb0, stage = it & 1;
    if (kb + 1 < kb1) load_block(kb + 1, vn);
    if (it >= 2) tc_mbar_wait(tc_smem(&s_bar[stage]), (uint32_t)(((it >> 1) - 1) & 1));     // the MMAs that read this stage are done
    store_block(stage, v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sA = tc_smem(tc_dyn + (size_t)stage * STAGE_BYTES), sB = sA + 2 * TC_A_BYTES;
      const uint64_t ahi0 = tc_desc(sA, lboA, sboA), alo0 = tc_desc(sA + TC_A_BYTES, lboA, sboA);
      const uint64_t bhi0 = tc_desc(sB, lboB, sboB), blo0 = tc_desc(sB + B_BYTES, lboB, sboB);
#pragma unroll
      for (int ks = 0; ks < TC_KB / 8; ++ks) {                  // the start-address field counts 16-byte units
        const uint64_t ahi = ahi0 + (uint64_t)(ks * 2 * 128), alo = alo0 + (uint64_t)(ks * 2 * 128);
        const uint64_t bhi = bhi0 + (uint64_t)(ks * 2 * BN), blo = blo0 + (uint64_t)(ks * 2 * BN);
        const int step = it * (TC_KB / 8) + ks;
        const uint32_t d = tmem + (uint32_t)((step % NACC) * BN);
        tc_mma_tf32(d, ahi, blo, IDESC, step >= NACC ? 1u : 0u);
        tc_mma_tf32(d, alo, bhi, IDESC, 1u);
        tc_mma_tf32(d, ahi, bhi, IDESC, 1u);
      }
      tc_commit(tc_smem(&s_bar[stage]));
    }
    if (kb + 1 < kb1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = vn[j];
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------------------
  const int nit = kb1 - kb0;
  if (nit > 0) {
    const int last = nit - 1;
    tc_mbar_wait(tc_smem(&s_bar[last & 1]), (uint32_t)((last >> 1) & 1));          // a commit covers every earlier MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                                                          // TMEM lane quarter this warp may read
    const long row = rowA0 + q * 32 + lane;
    size_t obase = 0; bool okr;
    if (MODE == 0) {
      okr = row < a.P;
      const long pp = okr ? row : 0;
      const int n = (int)(pp / HWo), dst = (int)(pp - (long)n * HWo);
      obase = (size_t)n * a.Crow * HWo + dst;                                        // + m * HWo
    } else {
      okr = row < a.K;
      obase = (size_t)(okr ? row : 0);                                               // + co * K
    }
    const size_t cstride = MODE == 0 ? (size_t)HWo : (size_t)a.K;
    const int nacc = min(NACC, nit * (TC_KB / 8));
    for (int c = (warp >> 2) * 16; c < BN; c += 32) {
      float acc[16];
      tc_ld16_sum(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, BN, nacc, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int m = rowB0 + c + i;
        if (okr && m < a.rowsB) {
          float* dst = out + obase + (size_t)m * cstride;
          if (a.atomic) atomicAdd(dst, acc[i]); else *dst = acc[i];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 (every stem and decoder convolution but decoder1 and the 7x7 stem): the "slab" kernel.
//
// Output positions are numbered on the ZERO-PADDED grid of the whole batch, q = (n (H + 2) + oh) * Wp + ow with Wp = W + 2 (the two
// extra columns per row and two extra rows per image are garbage rows of the tiles, dropped by the epilogue: H W / ((H + 2)(W + 2))
// of the MMA rows are useful, and tiles run across image boundaries -- the 64 small patches of MedT's local branch pack densely).  Then tap (kh, kw) of position q reads
// the padded input at q + kh * Wp + kw -- a CONSTANT row shift.  One CTA stages, per group of 8 input channels, the padded input
// slab of its 128 positions plus halo ([channel / 4][R = 128 + 2 Wp + 2 rows][4 channels x 4 B], i.e. the K-major SWIZZLE_NONE
// layout with 16-byte rows) ONCE, split into hi / lo, and issues the nine taps as nine MMAs whose A descriptors differ only in the
// start address (+ (kh * Wp + kw) * 16 B): every input element is loaded, split and stored once and used by nine taps -- 9x less
// gather work per MAC than the generic kernel above.  B = the 8 x 9 weights of the stage for the tile's BN output channels
// ([tap][channel / 4][BN rows][16 B]).  Per stage: 27 tcgen05.mma (9 taps x 3 split terms) of 128 x BN x 8.
// Register double buffering: the global loads of stage i + 2 are issued right after stage i's registers were stored, so a load has
// two stages of MMA time to land.  dgrad = the same kernel on dy with the taps flipped and the weight matrix transposed.
template <int BN>
__global__ void __launch_bounds__(288, 1)
k_conv_tc3(const TcConv a, const float* __restrict__ img, const float* __restrict__ wgt, float* __restrict__ out) {
  pdl_entry();
  constexpr int NACC = tc_nacc(BN);                           // independent TMEM accumulators, summed in fp32 by the epilogue
  constexpr uint32_t TMEM_COLS = NACC * BN;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int B_HALF = 9 * 2 * BN * 16;                     // hi (or lo) weights of one stage
  extern __shared__ __align__(128) unsigned char tc_dyn[];
  __shared__ __align__(8) unsigned long long s_full[2], s_empty[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.Hi, W = a.Wi, Wp = W + 2, HW = H * W, R = a.R;
  const int PI = (H + 2) * Wp;                                // padded rows of one image (its own zero halo rows separate the images)
  const int q0 = blockIdx.x * 128;                            // the tile's first row on the padded grid of the WHOLE batch
  const int rowB0 = blockIdx.y * BN;
  const int nst = a.Cimg >> 3;
  const int st0 = blockIdx.z * a.kb_per, st1 = min(nst, st0 + a.kb_per);
  const int A_HALF = 2 * R * 16;                              // hi (or lo) slab of one stage
  const int STAGE = 2 * A_HALF + 2 * B_HALF;

  if (tid == 0) {
    tc_mbar_init(tc_smem(&s_full[0]), 256);                  // every loader thread arrives
    tc_mbar_init(tc_smem(&s_full[1]), 256);
    tc_mbar_init(tc_smem(&s_empty[0]), 1);                   // tcgen05.commit arrives
    tc_mbar_init(tc_smem(&s_empty[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  // slab rows of this thread: padded index g = q0 + r -> input pixel (g / Wp - 1, g % Wp - 1), or -1 (zero padding / outside)
  int off[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = tid + 256 * i;
    const int g = q0 + r, n = g / PI, rem = g - n * PI, pa = rem / Wp, pb = rem - pa * Wp;
    const int ih = pa - 1, iw = pb - 1;
    off[i] = (r < R && n < a.N && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) ? n * a.Cimg * HW + ih * W + iw : -1;
  }
  const float* imgn = img;
  // weight task of this thread: output-channel row `brow`, channel quad `bc` of the stage
  const bool hasB = tid < 2 * BN;
  const int brow = tid % BN, bc = tid / BN;
  const int bm = rowB0 + brow;
  const bool okB = hasB && bm < a.Crow;

  auto load_stage = [&](int s, float (&va)[24], float (&vb)[36]) {
    const int c0 = s * 8;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float* src = imgn + (size_t)c0 * HW + (off[i] >= 0 ? off[i] : 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) va[i * 8 + j] = off[i] >= 0 ? __ldg(src + (size_t)j * HW) : 0.f;
    }
    if (okB) {
      if (!a.dgrad) {                                         // W[co = bm][c0 + 4 bc .. + 4][9]: 36 contiguous floats, 16-byte aligned
        const float4* src = reinterpret_cast<const float4*>(wgt + ((size_t)bm * a.Cimg + c0 + 4 * bc) * 9);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float4 t = __ldg(src + i);
          vb[4 * i] = t.x; vb[4 * i + 1] = t.y; vb[4 * i + 2] = t.z; vb[4 * i + 3] = t.w;
        }
      } else {                                                // W[co = c0 + 4 bc + e][ci = bm][9], taps flipped at store time
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* src = wgt + ((size_t)(c0 + 4 * bc + e) * a.Wm + bm) * 9;
#pragma unroll
          for (int t = 0; t < 9; ++t) vb[e * 9 + t] = __ldg(src + t);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 36; ++i) vb[i] = 0.f;
    }
  };

  auto store_stage = [&](int buf, const float (&va)[24], const float (&vb)[36]) {
    unsigned char* st = tc_dyn + (size_t)buf * STAGE;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int r = tid + 256 * i;
      if (r < R) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float4 h, l;
          tc_split(va[i * 8 + 4 * c + 0], h.x, l.x); tc_split(va[i * 8 + 4 * c + 1], h.y, l.y);
          tc_split(va[i * 8 + 4 * c + 2], h.z, l.z); tc_split(va[i * 8 + 4 * c + 3], h.w, l.w);
          reinterpret_cast<float4*>(st)[c * R + r] = h;
          reinterpret_cast<float4*>(st + A_HALF)[c * R + r] = l;
        }
      }
    }
    if (hasB) {
      float4* bh = reinterpret_cast<float4*>(st + 2 * A_HALF);
      float4* bl = reinterpret_cast<float4*>(st + 2 * A_HALF + B_HALF);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float4 h, l;
        float x0, x1, x2, x3;
        if (a.dgrad) { x0 = vb[8 - t]; x1 = vb[9 + 8 - t]; x2 = vb[18 + 8 - t]; x3 = vb[27 + 8 - t]; }
        else { x0 = vb[t]; x1 = vb[9 + t]; x2 = vb[18 + t]; x3 = vb[27 + t]; }
        tc_split(x0, h.x, l.x); tc_split(x1, h.y, l.y); tc_split(x2, h.z, l.z); tc_split(x3, h.w, l.w);
        bh[(t * 2 + bc) * BN + brow] = h;
        bl[(t * 2 + bc) * BN + brow] = l;
      }
    }
  };

  // loader side of one stage: wait until the MMAs that read this buffer have retired, store, publish to the async proxy, arrive
  auto run_stage = [&](int it, const float (&va)[24], const float (&vb)[36]) {
    const int buf = a.nbuf == 2 ? (it & 1) : 0;
    const int use = a.nbuf == 2 ? (it >> 1) : it;             // how often this buffer was filled before
    if (use >= 1) tc_mbar_wait(tc_smem(&s_empty[buf]), (uint32_t)((use - 1) & 1));
    store_stage(buf, va, vb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_mbar_arrive(tc_smem(&s_full[buf]));
  };

  const int nit = st1 - st0;
  if (warp < 8) {
    // ===== loader warps: registers double-buffered two stages ahead, shared memory double-buffered against the tensor core
    float va0[24], vb0[36], va1[24], vb1[36];
    if (nit > 0) load_stage(st0, va0, vb0);
    if (nit > 1) load_stage(st0 + 1, va1, vb1);
    for (int it = 0; it < nit; it += 2) {
      run_stage(it, va0, vb0);
      if (it + 2 < nit) load_stage(st0 + it + 2, va0, vb0);
      if (it + 1 < nit) {
        run_stage(it + 1, va1, vb1);
        if (it + 3 < nit) load_stage(st0 + it + 3, va1, vb1);
      }
    }
  } else if (lane == 0) {
    // ===== MMA issuer (one thread): 27 tcgen05.mma per stage, descriptors = base + constant row / tap offsets
    const uint32_t lboA = (uint32_t)R * 16u, lboB = (uint32_t)BN * 16u;
    for (int it = 0; it < nit; ++it) {
      const int buf = a.nbuf == 2 ? (it & 1) : 0;
      const int use = a.nbuf == 2 ? (it >> 1) : it;
      tc_mbar_wait(tc_smem(&s_full[buf]), (uint32_t)(use & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sA = tc_smem(tc_dyn + (size_t)buf * STAGE), sB = sA + 2 * A_HALF;
      const uint64_t ahi0 = tc_desc(sA, lboA, 128u), alo0 = tc_desc(sA + A_HALF, lboA, 128u);
      const uint64_t bhi0 = tc_desc(sB, lboB, 128u), blo0 = tc_desc(sB + B_HALF, lboB, 128u);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int kh = t / 3, kw = t - kh * 3;
        const uint64_t sh = (uint64_t)(uint32_t)(kh * Wp + kw);           // start-address field counts 16-byte units = slab rows
        const uint64_t tb = (uint64_t)(uint32_t)(t * 2 * BN);
        const int step = it * 9 + t;
        const uint32_t d = tmem + (uint32_t)((step % NACC) * BN);
        tc_mma_tf32(d, ahi0 + sh, blo0 + tb, IDESC, step >= NACC ? 1u : 0u);
        tc_mma_tf32(d, alo0 + sh, bhi0 + tb, IDESC, 1u);
        tc_mma_tf32(d, ahi0 + sh, bhi0 + tb, IDESC, 1u);
      }
      tc_commit(tc_smem(&s_empty[buf]));                      // frees the buffer; the last one also tells the epilogue
    }
  }

  if (nit > 0 && warp < 8) {
    const int last = nit - 1;
    const int buf = a.nbuf == 2 ? (last & 1) : 0, use = a.nbuf == 2 ? (last >> 1) : last;
    tc_mbar_wait(tc_smem(&s_empty[buf]), (uint32_t)(use & 1));               // a commit covers every earlier MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int g = q0 + q * 32 + lane, n = g / PI, rem = g - n * PI, oh = rem / Wp, ow = rem - oh * Wp;
    const bool okr = n < a.N && ow < W && oh < H;
    float* obase = out + (okr ? (size_t)n * a.Crow * HW + oh * W + ow : 0);
    const int nacc = min(NACC, nit * 9);
    for (int c = (warp >> 2) * 16; c < BN; c += 32) {
      float acc[16];
      tc_ld16_sum(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, BN, nacc, acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int m = rowB0 + c + i;
        if (okr && m < a.Crow) {
          float* dst = obase + (size_t)m * HW;
          if (a.atomic) atomicAdd(dst, acc[i]); else *dst = acc[i];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <int BN>
int tc3_launch_bn(const TcConv& a, dim3 grid, int smem, const float* img, const float* w, float* out, cudaStream_t st) {
  static int attr_smem[64] = {};                                   // per device: function attributes belong to a context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (smem > attr_smem[dev]) {
    if (cudaFuncSetAttribute(k_conv_tc3<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    attr_smem[dev] = smem;
  }
  if (mt_pdl_enabled()) return mt_launch_pdl(k_conv_tc3<BN>, grid, dim3(288), (size_t)smem, st, 1, a, img, w, out) == cudaSuccess ? 1 : -1;
  k_conv_tc3<BN><<<grid, 288, smem, st>>>(a, img, w, out);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int g_tc_mode = -1;

template <int MODE, int BN>
int tc_launch_bn(const TcConv& a, dim3 grid, const float* img, const float* mat, float* out, cudaStream_t st) {
  constexpr int smem = TC_STAGES * (2 * TC_A_BYTES + 2 * BN * TC_KB * 4);
  static int attr_smem[64] = {};                                   // per device: function attributes belong to a context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (smem > attr_smem[dev]) {
    if (cudaFuncSetAttribute(k_conv_tc<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    attr_smem[dev] = smem;
  }
  k_conv_tc<MODE, BN><<<grid, 256, smem, st>>>(a, img, mat, out);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int MODE>
int tc_launch(TcConv a, const float* img, const float* mat, float* out, size_t out_elems, cudaStream_t st) {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int bn = a.rowsB <= 16 ? 16 : a.rowsB <= 32 ? 32 : a.rowsB <= 64 ? 64 : 128;
  const long gx = ((long)a.rowsA + 127) / 128;
  const int gy = (a.rowsB + bn - 1) / bn;
  const int nkb = (a.kdim + TC_KB - 1) / TC_KB;
  // CTAs in flight: one per SM at BN = 128 (128 KB of shared memory), two or three for the narrow tiles
  const int per_sm = bn == 128 ? 1 : bn == 64 ? 2 : 3;
  long ks = ((long)nsm * per_sm + gx * gy - 1) / (gx * gy);
  const long ks_max = (nkb + 3) / 4;                              // at least 4 k blocks per slice: the epilogue's atomics are not free
  if (ks > ks_max) ks = ks_max;
  if (ks < 1) ks = 1;
  a.kb_per = (int)((nkb + ks - 1) / ks);
  ks = (nkb + a.kb_per - 1) / a.kb_per;
  a.atomic = ks > 1;
  a.dbg = g_tc_mode == 2 ? 1 : 0;
  if (gx > 0x7fffffffL || gy > 65535 || ks > 65535) return 0;
  if (a.atomic && cudaMemsetAsync(out, 0, sizeof(float) * out_elems, st) != cudaSuccess) return -1;
  const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)ks);
  switch (bn) {
    case 16: return tc_launch_bn<MODE, 16>(a, grid, img, mat, out, st);
    case 32: return tc_launch_bn<MODE, 32>(a, grid, img, mat, out, st);
    case 64: return tc_launch_bn<MODE, 64>(a, grid, img, mat, out, st);
    default: return tc_launch_bn<MODE, 128>(a, grid, img, mat, out, st);
  }
}

int tc3_launch(TcConv a, const float* img, const float* w, float* out, cudaStream_t st) {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int Wp = a.Wi + 2;
  a.R = 128 + 2 * Wp + 2;
  a.tiles_per_img = 0;
  const long prow = (long)a.N * (a.Hi + 2) * Wp;                   // rows of the batch's padded grid
  if (prow + a.R > 0x7fffffffL) return 0;
  int bn = a.Crow <= 16 ? 16 : a.Crow <= 32 ? 32 : a.Crow <= 64 ? 64 : 128;
  {
    static int min_bn = -1;                                        // MT_TC_MIN_BN: bring-up knob (N = 16 vs 32 accumulation accuracy)
    if (min_bn < 0) { const char* e = getenv("MT_TC_MIN_BN"); min_bn = e ? atoi(e) : 16; }
    if (bn < min_bn) bn = min_bn <= 32 ? 32 : min_bn <= 64 ? 64 : 128;
  }
  const int nst = a.Cimg / 8;
  auto stage_bytes = [&](int b) { return 64 * a.R + 576 * b; };
  a.nbuf = nst > 1 ? 2 : 1;
  if (a.nbuf * stage_bytes(bn) > 220 * 1024 && bn == 128) bn = 64;           // wide rows: narrower channel tile instead of one buffer
  if (a.nbuf * stage_bytes(bn) > 220 * 1024) a.nbuf = 1;
  if (stage_bytes(bn) > 220 * 1024) return 0;
  const long gx = (prow + 127) / 128;
  const int gy = (a.Crow + bn - 1) / bn;
  long ks = (long)nsm / (gx * gy);                                 // one CTA per SM (registers): stay inside ONE wave
  const long ks_max = (nst + 1) / 2;                               // at least two stages per slice
  if (ks > ks_max) ks = ks_max;
  if (ks < 1) ks = 1;
  a.kb_per = (int)((nst + ks - 1) / ks);
  ks = (nst + a.kb_per - 1) / a.kb_per;
  a.atomic = ks > 1;
  if (gx > 0x7fffffffL || gy > 65535 || ks > 65535) return 0;
  if (a.atomic && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)a.P * a.Crow, st) != cudaSuccess) return -1;
  const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)ks);
  const int smem = a.nbuf * stage_bytes(bn);
  switch (bn) {
    case 16: return tc3_launch_bn<16>(a, grid, smem, img, w, out, st);
    case 32: return tc3_launch_bn<32>(a, grid, smem, img, w, out, st);
    case 64: return tc3_launch_bn<64>(a, grid, smem, img, w, out, st);
    default: return tc3_launch_bn<128>(a, grid, smem, img, w, out, st);
  }
}

int ilog2_exact(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

int tc_conv_mode() {
  if (g_tc_mode < 0) {
    const char* e = getenv("MT_CONV_TC");
    g_tc_mode = e ? atoi(e) : 1;
    if (g_tc_mode < 0 || g_tc_mode > 2) g_tc_mode = 1;
  }
  return g_tc_mode;
}
void tc_conv_set_mode(int m) { g_tc_mode = (m < 0 || m > 2) ? 1 : m; }

int tc_conv_fwd_or_dgrad(const TcConv& geom, const float* img, const float* w, float* out, cudaStream_t st) {
  if (tc_conv_mode() == 0) return 0;
  TcConv a = geom;
  if (a.KH > 8 || a.KW > 8 || a.P <= 0 || a.P > 0x7fffffffL) return 0;
  if ((long)a.N * a.Cimg * a.Hi * a.Wi > 0x7fffffffL) return 0;              // 32-bit gather offsets
  a.rowsA = (int)a.P; a.rowsB = a.Crow; a.kdim = a.K; a.lgWo = a.lgHWo = 0;
  if (a.KH == 3 && a.KW == 3 && a.stride == 1 && a.pad == 1 && a.Cimg % 8 == 0 && a.Wi <= 256 && a.Ho == a.Hi && a.Wo == a.Wi &&
      ((uintptr_t)w & 15) == 0) {
    const int rc = tc3_launch(a, img, w, out, st);
    if (rc != 0) return rc;
  }
  return tc_launch<0>(a, img, w, out, (size_t)a.P * a.Crow, st);
}

int tc_conv_wgrad(const TcConv& geom, const float* x, const float* dy, float* dw, cudaStream_t st) {
  if (tc_conv_mode() == 0) return 0;
  TcConv a = geom;
  a.lgWo = ilog2_exact(a.Wo);
  a.lgHWo = ilog2_exact(a.Ho * a.Wo);
  if (a.lgWo < 0 || a.lgHWo < 5 || a.P <= 0 || a.P > 0x7fffffffL) return 0;   // a k block of 32 positions stays inside one image
  a.rowsA = a.K; a.rowsB = a.Crow; a.kdim = (int)a.P;
  return tc_launch<1>(a, x, dy, dw, (size_t)a.Crow * a.K, st);
}
