// TEST INFRASTRUCTURE -- fiber scheduler of the CUDA-on-CPU shim (see include/cuda_runtime.h).
#include <ucontext.h>
#include <vector>
#include "cuda_runtime.h"

namespace emu {
enum { READY = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int lin, state;
  unsigned long long slot;
};
Fiber* cur = nullptr;
uint3 g_bid = {0, 0, 0};
dim3 g_bdim, g_gdim;
unsigned char* g_dyn_smem = nullptr;

static ucontext_t g_sched;
static std::vector<Fiber> g_fibers;
static std::vector<char*> g_stacks;
static const std::function<void()>* g_body = nullptr;
static const size_t STACK = 512 * 1024;

const uint3& tid() { return cur->tid; }
int lin() { return cur->lin; }
int nthreads() { return (int)g_fibers.size(); }
unsigned long long* wslot(int l) { return &g_fibers[l].slot; }
static void yield_as(int st) { cur->state = st; swapcontext(&cur->ctx, &g_sched); }
void barrier_cta() { yield_as(WAIT_CTA); }
void barrier_warp() { yield_as(WAIT_WARP); }

static void fiber_main() {
  (*g_body)();
  cur->state = DONE;
  swapcontext(&cur->ctx, &g_sched);
}

static void run_cta() {
  const int n = (int)g_fibers.size();
  for (;;) {
    int alive = 0;
    for (int i = 0; i < n; ++i) {
      Fiber& f = g_fibers[i];
      if (f.state == READY) { cur = &f; swapcontext(&g_sched, &f.ctx); }
      if (f.state != DONE) ++alive;
    }
    if (!alive) return;
    bool released = false;
    // warp barriers first: a warp whose alive lanes all wait at a warp collective continues
    for (int w = 0; w < n; w += 32) {
      int a = 0, ww = 0;
      for (int l = w; l < w + 32 && l < n; ++l) { if (g_fibers[l].state != DONE) ++a; if (g_fibers[l].state == WAIT_WARP) ++ww; }
      if (a && ww == a) { for (int l = w; l < w + 32 && l < n; ++l) if (g_fibers[l].state == WAIT_WARP) g_fibers[l].state = READY; released = true; }
    }
    if (released) continue;
    int wc = 0;
    for (int i = 0; i < n; ++i) if (g_fibers[i].state == WAIT_CTA) ++wc;
    if (wc == alive) { for (int i = 0; i < n; ++i) if (g_fibers[i].state == WAIT_CTA) g_fibers[i].state = READY; continue; }
    fprintf(stderr, "emu: deadlock in CTA (%u,%u,%u): %d alive, %d at __syncthreads, rest inside a divergent warp collective\n",
            g_bid.x, g_bid.y, g_bid.z, alive, wc);
    abort();
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int n = (int)(block.x * block.y * block.z);
  g_bdim = block; g_gdim = grid; g_body = &body;
  while ((int)g_stacks.size() < n) g_stacks.push_back((char*)aligned_alloc(64, STACK));
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  g_fibers.assign(n, Fiber());
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_bid = {bx, by, bz};
        memset(g_dyn_smem, 0xCD, smem);           // poison: shared memory starts undefined
        for (int i = 0; i < n; ++i) {
          Fiber& f = g_fibers[i];
          f.lin = i; f.state = READY; f.slot = 0;
          f.tid = {(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = g_stacks[i]; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &g_sched;
          makecontext(&f.ctx, fiber_main, 0);
        }
        run_cta();
      }
  cur = nullptr; g_body = nullptr; g_dyn_smem = nullptr;
}
}  // namespace emu
