// medt_abi.cu -- the extern "C" surface declared in include/medt_axial.h.
// No torch types, no exceptions, no allocation, no device synchronisation.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

int axial_forward_launch(const Dims& m, const MtAxialParams& p, const float* x, float* out, char* saved,
                         cudaStream_t st, int* nlaunch, int mask);
int axial_backward_launch(const Dims& m, const MtAxialParams& p, const float* x, const float* dout, const char* saved,
                          char* ws, const MtAxialGrads& gr, cudaStream_t st, int* nlaunch, int mask);

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
thread_local cudaError_t g_last_cuda = cudaSuccess;   // set by the launchers when a runtime call fails

static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int mt_fail_msg(int code, const char* msg) { return fail(code, msg); }
// ---- fork/join side stream (per thread and device): independent kernels of one call (the weight-gradient
// GEMMs next to the input-gradient GEMMs) overlap instead of queueing behind each other.  Event fork/join
// is capturable, so the overlap survives inside the training-step CUDA graph.
struct SideStream { cudaStream_t s; cudaEvent_t fork, join; bool ok; };
static thread_local SideStream g_side[32] = {};
static SideStream* side_for_current_device() {
  static const bool disabled = getenv("MT_NO_SIDE_STREAM") != nullptr;
  if (disabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  SideStream& ss = g_side[dev];
  if (!ss.ok) {
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ss.ok = true;
  }
  return &ss;
}
// returns the stream to launch the forked work on (== main when no side stream is available)
cudaStream_t mt_side_fork(cudaStream_t main_stream) {
  SideStream* ss = side_for_current_device();
  if (!ss) return main_stream;
  if (cudaEventRecord(ss->fork, main_stream) != cudaSuccess) return main_stream;
  if (cudaStreamWaitEvent(ss->s, ss->fork, 0) != cudaSuccess) return main_stream;
  return ss->s;
}
void mt_side_join(cudaStream_t main_stream, cudaStream_t forked) {
  if (forked == main_stream) return;
  SideStream* ss = side_for_current_device();
  if (!ss) return;
  cudaEventRecord(ss->join, forked);
  cudaStreamWaitEvent(main_stream, ss->join, 0);
}

static int fail_cuda(const char* where) {
  cudaError_t e = g_last_cuda != cudaSuccess ? g_last_cuda : cudaGetLastError();
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
  return MT_ERR_CUDA;
}

static int check_common(const MtAxialDesc* d, const MtAxialParams* p, Dims& m) {
  if (!d || !p) return fail(MT_ERR_BAD_ARG, "null descriptor or params");
  if (!make_dims(d, m))
    return fail(MT_ERR_BAD_ARG, "bad MtAxialDesc (need N,C,H,W>0; C % 16 == 0; H,W % stride == 0; N % ngroups == 0)");
  if (m.d != 1 && m.d != 2 && m.d != 4 && m.d != 8)
    return fail(MT_ERR_UNSUPPORTED, "C/16 must be 1, 2, 4 or 8 (C in {16,32,64,128})");
  if (m.training && (double)m.Bg * m.L < 2.0)
    return fail(MT_ERR_BAD_ARG, "training needs more than one value per channel (as torch BatchNorm)");
  if (!p->w_qkv || !p->bn_qkv_w || !p->bn_qkv_b || !p->bn_qkv_rm || !p->bn_qkv_rv || !p->bn_sim_w || !p->bn_sim_b ||
      !p->bn_sim_rm || !p->bn_sim_rv || !p->bn_out_w || !p->bn_out_b || !p->bn_out_rm || !p->bn_out_rv)
    return fail(MT_ERR_BAD_ARG, "null parameter pointer");
  if (m.pos && !p->relative) return fail(MT_ERR_BAD_ARG, "relative is required for the position variants");
  if (!mt_al16(p->w_qkv)) return fail(MT_ERR_BAD_ARG, "w_qkv must be 16-byte aligned");
  return MT_OK;
}

extern "C" {

int mt_version(void) { return MT_ABI_VERSION; }
const char* mt_last_error(void) { return g_err; }
int mt_last_launch_count(void) { return g_launches; }

size_t mt_axial_saved_bytes(const MtAxialDesc* d) {
  Dims m;
  if (!d || !make_dims(d, m)) return 0;
  return m.sv_bytes;
}
size_t mt_axial_bwd_workspace_bytes(const MtAxialDesc* d) {
  Dims m;
  if (!d || !make_dims(d, m)) return 0;
  return m.ws_bytes;
}

static int forward_impl(const MtAxialDesc* d, const MtAxialParams* p, const float* x, float* out, void* saved,
                        void* stream, int mask) {
  Dims m;
  int rc = check_common(d, p, m);
  if (rc) return rc;
  if (!x || !out || !saved) return fail(MT_ERR_BAD_ARG, "null x / out / saved");
  if (!mt_al16(x) || !mt_al16(out) || !mt_al16(saved)) return fail(MT_ERR_BAD_ARG, "x / out / saved must be 16-byte aligned");
  g_launches = 0;
  int r = axial_forward_launch(m, *p, x, out, static_cast<char*>(saved), static_cast<cudaStream_t>(stream), &g_launches, mask);
  if (r == -2) return fail(MT_ERR_UNSUPPORTED, "shape needs more shared memory than one CTA has (C*L too large)");
  if (r != 0) return fail_cuda("mt_axial_forward");
  return MT_OK;
}
int mt_axial_forward(const MtAxialDesc* d, const MtAxialParams* p, const float* x, float* out, void* saved,
                     void* stream) {
  return forward_impl(d, p, x, out, saved, stream, 0x1f);
}
int mt_axial_forward_stages(const MtAxialDesc* d, const MtAxialParams* p, const float* x, float* out, void* saved,
                            void* stream, int stage_mask) {
  return forward_impl(d, p, x, out, saved, stream, stage_mask == 0x3f ? 0x1f : (stage_mask & 0xf));
}

static int backward_impl(const MtAxialDesc* d, const MtAxialParams* p, const float* x, const float* dout,
                         const void* saved, void* workspace, const MtAxialGrads* g, void* stream, int mask) {
  Dims m;
  int rc = check_common(d, p, m);
  if (rc) return rc;
  if (!x || !dout || !saved || !workspace || !g) return fail(MT_ERR_BAD_ARG, "null x / dout / saved / workspace / grads");
  if (!mt_al16(x) || !mt_al16(dout) || !mt_al16(saved) || !mt_al16(workspace) || !mt_al16(g->dx))
    return fail(MT_ERR_BAD_ARG, "x / dout / saved / workspace / dx must be 16-byte aligned");
  g_launches = 0;
  int r = axial_backward_launch(m, *p, x, dout, static_cast<const char*>(saved), static_cast<char*>(workspace), *g,
                                static_cast<cudaStream_t>(stream), &g_launches, mask);
  if (r == -2) return fail(MT_ERR_UNSUPPORTED, "shape needs more shared memory than one CTA has (C*L too large)");
  if (r == -3) return fail(MT_ERR_BAD_ARG, "drelative is required for the position variants");
  if (r != 0) return fail_cuda("mt_axial_backward");
  return MT_OK;
}
int mt_axial_backward(const MtAxialDesc* d, const MtAxialParams* p, const float* x, const float* dout,
                      const void* saved, void* workspace, const MtAxialGrads* g, void* stream) {
  return backward_impl(d, p, x, dout, saved, workspace, g, stream, 0x1f);
}
int mt_axial_backward_stages(const MtAxialDesc* d, const MtAxialParams* p, const float* x, const float* dout,
                             const void* saved, void* workspace, const MtAxialGrads* g, void* stream,
                             int stage_mask) {
  return backward_impl(d, p, x, dout, saved, workspace, g, stream, stage_mask == 0x3f ? 0x3f : (stage_mask & 0xf));
}

}  // extern "C"
