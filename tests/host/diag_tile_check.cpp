// Host-side check of medical-transformer_b200/csrc/axial_diag_tile.cuh (compiled with g++, no CUDA): the register-tiled
// wrapped-diagonal pass must reproduce the plain double loop over (i, j) for d relative.  Prints "OK <maxrel>" per case.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../medical-transformer_b200/csrc/axial_diag_tile.cuh"

static float frand(unsigned& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; }

template <int D, int RT>
static int run(int L, int pad) {
  const int GP = 2 * D, LP = L + pad, RP = 2 * L + 8, RL = 2 * L - 1;
  unsigned seed = 1234u + 17u * L + 3u * D + RT;
  std::vector<float> Y((size_t)4 * D * LP), dO((size_t)2 * GP * LP), ls(L), Dp(L), R((size_t)2 * GP * RP), a((size_t)2 * GP * RL, 0.f),
      b((size_t)2 * GP * RL, 0.f);
  float* Yal = (float*)aligned_alloc(64, sizeof(float) * Y.size());
  float* dOal = (float*)aligned_alloc(64, sizeof(float) * dO.size());
  for (size_t k = 0; k < Y.size(); ++k) Yal[k] = frand(seed);
  for (size_t k = 0; k < dO.size(); ++k) dOal[k] = frand(seed);
  for (auto& v : ls) v = 2.0f + frand(seed);
  for (auto& v : Dp) v = 0.3f * frand(seed);
  for (auto& v : R) v = 0.5f * frand(seed);
  mtdiag::Coefs cf{0.7f, -0.4f, 0.55f, 1.3f, 0.02f, -0.15f, 0.9f, -0.03f, 0.2f};
  mtdiag::diag_reference<D>(L, LP, RP, Yal, dOal, ls.data(), Dp.data(), R.data(), cf, a.data());
  for (int rt = 0; rt < L / RT; ++rt)
    mtdiag::diag_tile_item<D, RT>(L, LP, RP, rt, Yal, dOal, ls.data(), Dp.data(), R.data(), cf, b.data());
  double mx = 0, err = 0;
  for (size_t k = 0; k < a.size(); ++k) { mx = fmax(mx, fabs(a[k])); err = fmax(err, fabs((double)a[k] - b[k])); }
  free(Yal); free(dOal);
  const double rel = err / (mx > 0 ? mx : 1);
  printf("%s D=%d RT=%d L=%d LP=%d maxrel=%.3e\n", rel < 2e-5 ? "OK" : "FAIL", D, RT, L, LP, rel);
  return rel < 2e-5 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<1, 4>(64, 0);  bad += run<1, 4>(16, 4);  bad += run<1, 4>(4, 4);   bad += run<1, 4>(256, 0);
  bad += run<1, 2>(64, 0);  bad += run<1, 2>(6, 4);   bad += run<1, 2>(2, 4);
  bad += run<2, 2>(64, 0);  bad += run<2, 2>(32, 0);  bad += run<2, 2>(8, 4);   bad += run<2, 2>(256, 0);
  bad += run<4, 2>(16, 4);
  return bad;
}
