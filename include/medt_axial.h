/* medt_axial.h -- C ABI of libmedt_b200.so: the gated axial-attention hot path of
 * Medical-Transformer as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary.  Every entry point replaces a piece of
 * the reference's pure-PyTorch path in /root/reference/lib/models/axialnet.py:
 *
 *   mt_axial_forward   <- AxialAttention_dynamic.forward  axialnet.py:142-189  (variant 0)
 *                         AxialAttention.forward          axialnet.py:52-92    (variant 1)
 *                         AxialAttention_wopos.forward    axialnet.py:222-253  (variant 2)
 *   mt_axial_backward  <- torch.autograd over the same lines (the reference has no
 *                         hand-written backward; train.py:160 calls loss.backward())
 *   mt_bn_act_*        <- BatchNorm2d(+ReLU)(+residual) glue of the blocks / stems / LoGo loop
 *   mt_up2x_relu_add_* <- relu(bilinear x2) + skip of the decoders
 *   mt_conv1x1_*       <- the bias-free 1x1 convolutions of the blocks
 *   mt_conv2d_*        <- the KxK convolutions of the stems and decoders
 *   mt_head_ce_*       <- adjust(relu(x)) + F.cross_entropy, the tail of one training step (train.py:140,156)
 *
 * Conventions
 *   - plain C, no torch types; all pointers are DEVICE pointers unless named host_*;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the device, never allocates, and is CUDA-graph capturable;
 *   - the caller owns every buffer (activations, `saved`, `workspace`); sizes come from
 *     the mt_*_bytes() queries;
 *   - return value: 0 = ok, nonzero = error (message via mt_last_error(), thread-local);
 *     nothing throws across the ABI, nothing calls exit();
 *   - re-entrant: the device is whatever is current on the calling thread (torch's device
 *     guard), so nn.DataParallel's one-thread-per-GPU model works.
 *   - all tensors fp32, activations NCHW (the reference's layout and dtype).
 *   - activation, `saved`, `workspace` and dx buffers and w_qkv must be dense and 16-byte aligned (128-bit
 *     loads, TMA bulk copies); a misaligned pointer is MT_ERR_BAD_ARG, never a fault.  mt_bn_act_* instead
 *     falls back to its scalar kernels; mt_conv1x1_* has no alignment requirement.
 */
#ifndef MEDT_AXIAL_H_
#define MEDT_AXIAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_ABI_VERSION 2

enum MtStatus {
  MT_OK = 0,
  MT_ERR_BAD_ARG = 1,      /* null pointer, non-positive size, stride not in {1,2,...}    */
  MT_ERR_UNSUPPORTED = 2,  /* channel / length combination the kernels do not cover       */
  MT_ERR_CUDA = 3          /* a CUDA runtime call failed (launch error, bad stream, ...)  */
};

enum MtVariant {
  MT_VARIANT_GATED = 0,    /* AxialAttention_dynamic: rel-pos terms + 4 gates             */
  MT_VARIANT_POS = 1,      /* AxialAttention: rel-pos terms, gates == 1                   */
  MT_VARIANT_NOPOS = 2     /* AxialAttention_wopos: qk only, BN2d(G), BN1d(C)             */
};

/* Geometry of one axial-attention op.  groups is fixed at 8 (axialnet.py:400,715-728). */
typedef struct MtAxialDesc {
  int32_t N, C, H, W;      /* input x is N x C x H x W                                    */
  int32_t axis;            /* 0: attend along H (reference width=False); 1: along W       */
  int32_t stride;          /* AvgPool2d(stride) on the op's output (axialnet.py:186-187)  */
  int32_t variant;         /* enum MtVariant                                              */
  int32_t training;        /* 1: batch statistics + running-stat EMA; 0: running stats    */
  int32_t ngroups;         /* statistic groups: N is split into `ngroups` consecutive
                              blocks with SEPARATE batch statistics and one sequential EMA
                              update each (16 = the batched LoGo patches); 1 = reference  */
  float eps, momentum;     /* BatchNorm eps / momentum (1e-5 / 0.1 in the reference)      */
} MtAxialDesc;

/* Parameters and BN buffers of one AxialAttention* module (same tensors as its
 * state_dict: qkv_transform.weight, relative, f_*, bn_qkv.*, bn_similarity.*, bn_output.*). */
typedef struct MtAxialParams {
  const float* w_qkv;      /* (2C, C)   qkv_transform.weight[:, :, 0]                     */
  const float* relative;   /* (2*C/8, 2L-1) or NULL for MT_VARIANT_NOPOS                  */
  const float* f_qr;       /* 0-dim gate tensors (device); NULL unless MT_VARIANT_GATED   */
  const float* f_kr;
  const float* f_sv;
  const float* f_sve;
  const float* bn_qkv_w;   const float* bn_qkv_b;   float* bn_qkv_rm;   float* bn_qkv_rv;   /* (2C)      */
  const float* bn_sim_w;   const float* bn_sim_b;   float* bn_sim_rm;   float* bn_sim_rv;   /* (24) / (8) */
  const float* bn_out_w;   const float* bn_out_b;   float* bn_out_rm;   float* bn_out_rv;   /* (2C) / (C) */
} MtAxialParams;

/* Gradients written by mt_axial_backward (each buffer fully overwritten; NULL = skip). */
typedef struct MtAxialGrads {
  float* dx;               /* (N, C, H, W)                                                */
  float* dw_qkv;           /* (2C, C)                                                     */
  float* drelative;        /* (2*C/8, 2L-1)                                               */
  float* dgates;           /* (4): f_qr, f_kr, f_sv, f_sve                                */
  float* dbn_qkv_w;        float* dbn_qkv_b;
  float* dbn_sim_w;        float* dbn_sim_b;
  float* dbn_out_w;        float* dbn_out_b;
} MtAxialGrads;

int mt_version(void);
const char* mt_last_error(void);

/* Bytes of the `saved` buffer (forward -> backward state: projection, un-normalised
 * outputs, log-sum-exp, fp64 statistic accumulators) and of the backward workspace. */
size_t mt_axial_saved_bytes(const MtAxialDesc* d);
size_t mt_axial_bwd_workspace_bytes(const MtAxialDesc* d);

/* out: (N, C, H/stride, W/stride).  Updates the running statistics in `p` when
 * d->training (num_batches_tracked is the caller's: += ngroups). */
int mt_axial_forward(const MtAxialDesc* d, const MtAxialParams* p, const float* x, float* out,
                     void* saved, void* stream);

/* dout: (N, C, H/stride, W/stride).  `saved` is what mt_axial_forward filled for the same
 * desc / x / params. */
int mt_axial_backward(const MtAxialDesc* d, const MtAxialParams* p, const float* x, const float* dout,
                      const void* saved, void* workspace, const MtAxialGrads* g, void* stream);

/* ---- fused glue of the conv shell (all NCHW fp32) ---------------------------------------------
 * relu(F.interpolate(in, scale_factor=2, mode='bilinear', align_corners=False)) + skip
 *   <- one decoder step, axialnet.py:493-501 / :650-652 / :690-698.
 * in: (planes, H, W) with planes = N*C; out / skip / mask: (planes, 2H, 2W); skip may be NULL;
 * mask (1 byte per output, = pre-activation > 0) may be NULL when no backward will follow.
 * bias (C floats, nullable): the bias of the decoder conv that produced `in`, added here instead of by a separate
 * elementwise pass -- relu(up(in + b)) == relu(up(in) + b) because the bilinear taps sum to 1. */
int mt_up2x_relu_add_forward(int planes, int C, int H, int W, const float* in, const float* bias, const float* skip,
                             float* out, unsigned char* mask, void* stream);
/* din = upsample^T(dout * mask); the gradient of `skip` is dout itself (not written).
 * dbias (C floats, nullable, overwritten) = sum over (n, h, w) of din: the conv's bias gradient, reduced in the same
 * pass (fp32 atomics: last-bit run-to-run differences). */
int mt_up2x_relu_add_backward(int planes, int C, int H, int W, const float* dout, const unsigned char* mask, float* din,
                              float* dbias, void* stream);

/* y = [relu]( BatchNorm2d(x) [+ residual] ) with batch statistics per statistic group
 *   <- bn+relu(+identity) of AxialBlock*.forward axialnet.py:285-300 / :327-342 / :373-389, the
 *      stems :475-483, and the 16 per-patch calls of the LoGo loop :661-700 (groups = 16:
 *      per-group statistics + `groups` sequential running-stat EMA steps, as 16 BN calls do).
 * x, residual (nullable), y: (N, C, HW).  save_mean: (2*groups, C) [means, then biased
 * variances]; save_invstd: (groups, C).  training = 0 uses / leaves the running statistics. */
int mt_bn_act_forward(int N, int C, int HW, int groups, int training, int relu, float eps, float momentum,
                      const float* x, const float* residual, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                      void* stream);
/* dx, dresidual (nullable; = dy masked by the ReLU), dgamma, dbeta (C each, overwritten). */
int mt_bn_act_backward(int N, int C, int HW, int groups, int training, int relu, const float* x, const float* y,
                       const float* dy, const float* gamma, const float* save_mean, const float* save_invstd,
                       float* dx, float* dresidual, float* dgamma, float* dbeta, void* stream);

/* Bias-free 1x1 convolution (conv_down / conv_up / downsample of the blocks: `conv1x1`, axialnet.py:14-16,
 * :272-276, :450-453) as a tiled fp32 SGEMM.  x: (N,Ci,H,W), w: (Co,Ci), y / dy: (N,Co,Ho,Wo) with
 * Ho = (H-1)/stride+1.  backward overwrites dx (N,Ci,H,W) and dw (Co,Ci); either may be NULL.
 * relu_in != 0: the conv reads relu(x) (`conv_up(self.relu(output))`, axialnet.py:291-293 / :333-335 / :380-382) -- the ReLU is
 * applied while loading the operand and its backward while storing dx; pass the SAME (pre-ReLU) x to backward. */
int mt_conv1x1_forward(int N, int Ci, int Co, int H, int W, int stride, int relu_in, const float* x, const float* w,
                       float* y, void* stream);
int mt_conv1x1_backward(int N, int Ci, int Co, int H, int W, int stride, int relu_in, const float* x, const float* w,
                        const float* dy, float* dx, float* dw, void* stream);

/* KxK convolution of the conv shell (stem conv2 / conv3 `axialnet.py:424-428,478-483`, the decoder convs `:431-440,493-502`,
 * LoGo's `decoderf` `:585`) as an implicit-GEMM fp32 kernel with split-K over the whole GPU (cuDNN's fp32 algorithms for these
 * small-M x long-K problems run at 0.4 - 8 TFLOP/s).  x: (N,Ci,H,W), w: (Co,Ci,KH,KW), y / dy: (N,Co,Ho,Wo) with
 * Ho = (H + 2 pad - KH) / stride + 1; no bias (the decoders' bias is added by mt_up2x_relu_add_forward), no dilation / groups.
 * backward overwrites dx (N,Ci,H,W; stride 1 only: pass NULL otherwise -> MT_ERR_UNSUPPORTED) and dw (Co,Ci,KH,KW); either may be
 * NULL.  Split-K slices accumulate with fp32 atomics: last-bit run-to-run differences. */
int mt_conv2d_forward(int N, int Ci, int Co, int H, int W, int KH, int KW, int stride, int pad, const float* x, const float* w,
                      float* y, void* stream);
int mt_conv2d_backward(int N, int Ci, int Co, int H, int W, int KH, int KW, int stride, int pad, const float* x, const float* w,
                       const float* dy, float* dx, float* dw, void* stream);

/* Routing of mt_conv2d_*: 1 (default; env MT_CONV_TC) = the tcgen05 tensor-core kernels (csrc/conv_tc.cu: kind::tf32 MMAs with the
 * 3xTF32 operand split, fp32-grade results, accumulators in TMEM) wherever the shape is eligible, 0 = the FMA-pipe kernels only.
 * Process-wide; not thread-safe against concurrent mt_conv2d_* calls. */
int mt_set_conv_tc(int mode);     /* returns MT_OK */

/* The tail of a training step: logits = adjust(relu(x)) (1x1 conv with bias to TWO classes, axialnet.py:440,502 / :588,708) fused
 * with loss = F.cross_entropy(logits, y) (metrics.py:17-20, train.py:156; mean over all pixels, labels in {0, 1}).
 * x: (N,C,HW) pre-ReLU features, w: (2,C), b: (2), y: (N,HW) int64, logits: (N,2,HW) out, loss: 1 float, overwritten.
 * backward: dloss (1 float, device) -> dx (N,C,HW; nullable), dw (2,C), db (2), all overwritten.  C <= 64. */
int mt_head_ce_forward(int N, int C, int HW, const float* x, const float* w, const float* b, const long long* y, float* logits,
                       float* loss, void* stream);
int mt_head_ce_backward(int N, int C, int HW, const float* x, const float* w, const float* logits, const long long* y,
                        const float* dloss, float* dx, float* dw, float* db, void* stream);

/* Profiling hooks (bench.py roofline, ncu): launch only the stages selected by `stage_mask`
 * of a forward (bit 0 k_proj, 1 k_sim_stats, 2 k_attn, 3 k_out) or backward (bit 0
 * k_bwd_gather, 1 k_bwd_simstats, 2 k_bwd_attn, 3 k_bwd_proj) on buffers a full call has
 * already filled.  Statistic accumulators are NOT re-zeroed, so outputs are meaningless;
 * durations are those of the real launches.  Mask 15 == the full launch CHAIN minus the memsets (the round-1 path, kept for
 * shapes the cooperative kernels do not cover).  Mask 0x3f == the DEFAULT path: the persistent cooperative kernel of the op
 * (k_axial_fwd / k_axial_bwd, one launch) with its accumulator memset and, for backward, WITHOUT the weight-gradient GEMM
 * that a full call forks onto the side stream -- results are valid, this is what bench.py's roofline times. */
int mt_axial_forward_stages(const MtAxialDesc* d, const MtAxialParams* p, const float* x, float* out,
                            void* saved, void* stream, int stage_mask);
int mt_axial_backward_stages(const MtAxialDesc* d, const MtAxialParams* p, const float* x, const float* dout,
                             const void* saved, void* workspace, const MtAxialGrads* g, void* stream,
                             int stage_mask);

/* Number of kernels the last forward / backward call on this thread launched (used by
 * bench.py's gpu_launches accounting). */
int mt_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MEDT_AXIAL_H_ */
