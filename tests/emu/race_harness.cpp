// TEST INFRASTRUCTURE -- drives the emulated library (built with -fsanitize=thread -DEMU_TSAN) through the C ABI on random
// data; ThreadSanitizer reports every pair of conflicting accesses by different GPU threads that no __syncthreads / warp
// collective / atomic / kernel boundary orders.  No result checking here (tests/test_emu_kernels.py does that).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/medt_axial.h"

static unsigned g_seed = 12345u;
static float frand() { g_seed = g_seed * 1664525u + 1013904223u; return ((g_seed >> 8) & 0xffff) / 32768.0f - 1.0f; }
static float* buf(size_t n, float scale = 1.f, float shift = 0.f) {
  float* p = (float*)aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64);
  for (size_t i = 0; i < n; ++i) p[i] = scale * frand() + shift;
  return p;
}

static int axial(int variant, int C, int N, int H, int W, int axis, int stride, int training, int ngroups) {
  MtAxialDesc d = {N, C, H, W, axis, stride, variant, training, ngroups, 1e-5f, 0.1f};
  const int L = axis == 0 ? H : W, gp = C / 8, nsim = variant == 2 ? 8 : 24, nout = variant == 2 ? C : 2 * C;
  MtAxialParams p;
  memset(&p, 0, sizeof(p));
  p.w_qkv = buf((size_t)2 * C * C, 0.3f);
  if (variant != 2) p.relative = buf((size_t)2 * gp * (2 * L - 1), 0.5f);
  if (variant == 0) { p.f_qr = buf(1, 0.f, 0.3f); p.f_kr = buf(1, 0.f, -0.2f); p.f_sv = buf(1, 0.f, 0.9f); p.f_sve = buf(1, 0.f, 0.15f); }
  p.bn_qkv_w = buf(2 * C, 0.2f, 1.f); p.bn_qkv_b = buf(2 * C, 0.2f); p.bn_qkv_rm = buf(2 * C, 0.1f); p.bn_qkv_rv = buf(2 * C, 0.2f, 1.f);
  p.bn_sim_w = buf(nsim, 0.2f, 1.f); p.bn_sim_b = buf(nsim, 0.2f); p.bn_sim_rm = buf(nsim, 0.1f); p.bn_sim_rv = buf(nsim, 0.2f, 1.f);
  p.bn_out_w = buf(nout, 0.2f, 1.f); p.bn_out_b = buf(nout, 0.2f); p.bn_out_rm = buf(nout, 0.1f); p.bn_out_rv = buf(nout, 0.2f, 1.f);
  float* x = buf((size_t)N * C * H * W);
  float* out = buf((size_t)N * C * (H / stride) * (W / stride));
  void* saved = aligned_alloc(256, ((mt_axial_saved_bytes(&d) + 255) / 256) * 256);
  void* ws = aligned_alloc(256, ((mt_axial_bwd_workspace_bytes(&d) + 255) / 256) * 256);
  int rc = mt_axial_forward(&d, &p, x, out, saved, nullptr);
  if (rc) { printf("forward rc=%d %s\n", rc, mt_last_error()); return 1; }
  float* dout = buf((size_t)N * C * (H / stride) * (W / stride));
  MtAxialGrads g;
  memset(&g, 0, sizeof(g));
  g.dx = buf((size_t)N * C * H * W); g.dw_qkv = buf((size_t)2 * C * C); g.dgates = buf(4);
  if (variant != 2) g.drelative = buf((size_t)2 * gp * (2 * L - 1));
  g.dbn_qkv_w = buf(2 * C); g.dbn_qkv_b = buf(2 * C); g.dbn_sim_w = buf(nsim); g.dbn_sim_b = buf(nsim); g.dbn_out_w = buf(nout); g.dbn_out_b = buf(nout);
  rc = mt_axial_backward(&d, &p, x, dout, saved, ws, &g, nullptr);
  if (rc) { printf("backward rc=%d %s\n", rc, mt_last_error()); return 1; }
  printf("axial variant=%d C=%d N=%d H=%d W=%d axis=%d stride=%d train=%d groups=%d done\n", variant, C, N, H, W, axis, stride, training, ngroups);
  return 0;
}

static int shell() {
  {  // bn_act: vector and scalar kernels, grouped
    const int cfg[3][5] = {{2, 8, 64, 1, 1}, {6, 5, 21, 3, 1}, {4, 16, 16, 4, 0}};      // N, C, HW, groups, relu
    for (auto& c : cfg) {
      const int N = c[0], C = c[1], HW = c[2], G = c[3], relu = c[4];
      float *x = buf((size_t)N * C * HW), *res = buf((size_t)N * C * HW), *ga = buf(C, 0.2f, 1.f), *be = buf(C, 0.2f), *rm = buf(C, 0.1f), *rv = buf(C, 0.2f, 1.f);
      float *y = buf((size_t)N * C * HW), *sm = buf((size_t)2 * G * C), *si = buf((size_t)G * C);
      if (mt_bn_act_forward(N, C, HW, G, 1, relu, 1e-5f, 0.1f, x, res, ga, be, rm, rv, y, sm, si, nullptr)) return 1;
      float *dy = buf((size_t)N * C * HW), *dx = buf((size_t)N * C * HW), *dr = buf((size_t)N * C * HW), *dg = buf(C), *db = buf(C);
      if (mt_bn_act_backward(N, C, HW, G, 1, relu, x, y, dy, ga, sm, si, dx, dr, dg, db, nullptr)) return 1;
    }
    printf("bn_act done\n");
  }
  {  // up2x with bias and bias gradient
    const int N = 2, C = 6, H = 5, W = 9, planes = N * C;
    float *in = buf((size_t)planes * H * W), *bias = buf(C), *skip = buf((size_t)planes * 4 * H * W), *out = buf((size_t)planes * 4 * H * W);
    unsigned char* mask = (unsigned char*)aligned_alloc(64, (((size_t)planes * 4 * H * W + 63) / 64) * 64);
    if (mt_up2x_relu_add_forward(planes, C, H, W, in, bias, skip, out, mask, nullptr)) return 1;
    float *dout = buf((size_t)planes * 4 * H * W), *din = buf((size_t)planes * H * W), *dbias = buf(C);
    if (mt_up2x_relu_add_backward(planes, C, H, W, dout, mask, din, dbias, nullptr)) return 1;
    printf("up2x done\n");
  }
  {  // 1x1 conv, stride 2, relu_in
    const int N = 2, Ci = 20, Co = 70, H = 6, W = 10, st = 2, Ho = (H - 1) / st + 1, Wo = (W - 1) / st + 1;
    float *x = buf((size_t)N * Ci * H * W), *w = buf((size_t)Co * Ci, 0.2f), *y = buf((size_t)N * Co * Ho * Wo);
    if (mt_conv1x1_forward(N, Ci, Co, H, W, st, 1, x, w, y, nullptr)) return 1;
    float *dy = buf((size_t)N * Co * Ho * Wo), *dx = buf((size_t)N * Ci * H * W), *dw = buf((size_t)Co * Ci);
    if (mt_conv1x1_backward(N, Ci, Co, H, W, st, 1, x, w, dy, dx, dw, nullptr)) return 1;
    printf("conv1x1 done\n");
  }
  return 0;
}

int main() {
  int bad = 0;
  bad += axial(0, 16, 2, 16, 8, 0, 1, 1, 1);     // tiled kernels, d=1 RT=4, padded pitch
  bad += axial(0, 16, 1, 64, 2, 0, 1, 1, 1);     // L=64: TMA-staged layout
  bad += axial(0, 32, 2, 4, 32, 1, 2, 1, 1);     // d=2, stride 2
  bad += axial(1, 16, 4, 8, 8, 1, 1, 1, 2);      // grouped statistics
  bad += axial(2, 16, 4, 8, 8, 0, 2, 1, 4);      // nopos, 4 groups (no replicas)
  bad += axial(0, 64, 1, 8, 8, 0, 1, 1, 1);      // d=4
  bad += axial(0, 128, 1, 8, 4, 1, 1, 1, 1);     // d=8: one-row-per-thread kernels
  bad += axial(0, 16, 2, 6, 10, 0, 1, 1, 1);     // ragged length
  bad += axial(0, 16, 2, 16, 8, 0, 1, 0, 1);     // eval
  bad += axial(0, 16, 1, 32, 32, 0, 1, 1, 1);    // transposed plane kernels (H, W >= 32)
  bad += shell();
  printf(bad ? "FAILED\n" : "ALL DONE\n");
  return bad;
}
