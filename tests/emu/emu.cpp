// TEST INFRASTRUCTURE -- fiber scheduler of the CUDA-on-CPU shim (see include/cuda_runtime.h).
#include <ucontext.h>
#include <map>
#include <utility>
#include <vector>
#include "cuda_runtime.h"

// (compile THIS file without -fsanitize=thread: the scheduler's own bookkeeping is not part of the race model)
// -DEMU_TSAN (+ -fsanitize=thread on the kernel sources): every GPU thread becomes a ThreadSanitizer fiber; __syncthreads / warp collectives /
// kernel boundaries are the only happens-before edges (plus the builtin atomics), so a missing barrier between a
// shared- or global-memory write and another thread's access is reported as a data race with both source lines.
#ifdef EMU_TSAN
#include <sanitizer/tsan_interface.h>
#define TSAN_ONLY(x) x
#else
#define TSAN_ONLY(x)
#endif

namespace emu {
enum { READY = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3, WAIT_GRID = 4 };
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int lin, state;
  int cta;                  // index into g_ctas (cooperative launches co-schedule several CTAs)
  unsigned long long slot;
  void* ts;                 // ThreadSanitizer fiber context (EMU_TSAN)
};
struct Cta {
  uint3 bid;
  unsigned char* dyn;
  std::map<const void*, void*> blocks;      // static __shared__ storage of this CTA, by declaration
  int first;                                // index of its first fiber
};
Fiber* cur = nullptr;
uint3 g_bid = {0, 0, 0};
dim3 g_bdim, g_gdim;
unsigned char* g_dyn_smem = nullptr;

static ucontext_t g_sched;
static std::vector<Fiber> g_fibers;
static std::vector<Cta> g_ctas;
static int g_tpc = 0;                                           // threads per CTA
static std::vector<char*> g_stacks;
static const std::function<void()>* g_body = nullptr;
static const size_t STACK = 512 * 1024;

const uint3& tid() { return cur->tid; }
int lin() { return cur->lin; }
int nthreads() { return g_tpc; }
static void* g_sched_ts = nullptr;
static char g_sync_cta, g_sync_warp[64], g_sync_kernel[2], g_sync_cta_end, g_sync_grid;
static int g_kidx = 0;

static void* fresh_block(size_t n) {                            // a NEW heap block: undefined contents, no race history
  void* p = aligned_alloc(64, ((n ? n : 1) + 63) / 64 * 64);
  volatile unsigned char* q = (volatile unsigned char*)p;      // plain loop: memset() would be intercepted by the sanitizers
  for (size_t i = 0; i < n; ++i) q[i] = 0xCD;                   // and attributed to whichever thread got here first
  return p;
}
void* cta_shared(const void* key, size_t bytes) {
  std::map<const void*, void*>& g_cta_blocks = g_ctas[cur->cta].blocks;
  auto it = g_cta_blocks.find(key);
  if (it != g_cta_blocks.end()) {
    TSAN_ONLY(__tsan_acquire(const_cast<void*>(key));)          // the allocation (by the first thread to get here) precedes every use
    return it->second;
  }
  void* p = fresh_block(bytes);
  g_cta_blocks[key] = p;
  TSAN_ONLY(__tsan_release(const_cast<void*>(key));)
  return p;
}
unsigned long long exchange(unsigned long long raw, int src_lane_in_warp) {
  const int f0 = g_ctas[cur->cta].first, me = cur->lin, base = me & ~31;
  g_fibers[f0 + me].slot = raw;
  barrier_warp();
  const int src = base + (src_lane_in_warp & 31);
  const unsigned long long r = src < g_tpc ? g_fibers[f0 + src].slot : raw;
  barrier_warp();
  return r;
}
int all(int pred) {
  const int f0 = g_ctas[cur->cta].first, me = cur->lin, base = me & ~31, n = g_tpc;
  g_fibers[f0 + me].slot = pred ? 1ull : 0ull;
  barrier_warp();
  int a = 1;
  for (int l = base; l < base + 32 && l < n; ++l) a &= (int)g_fibers[f0 + l].slot;
  barrier_warp();
  return a;
}
static void to_sched() {
  TSAN_ONLY(__tsan_switch_to_fiber(g_sched_ts, __tsan_switch_to_fiber_no_sync);)
  swapcontext(&cur->ctx, &g_sched);
}
static void yield_as(int st) { cur->state = st; to_sched(); }
void barrier_cta() {
  TSAN_ONLY(__tsan_release(&g_sync_cta);)
  yield_as(WAIT_CTA);
  TSAN_ONLY(__tsan_acquire(&g_sync_cta);)
}
void barrier_grid() {                       // called by ONE thread per CTA (the cooperative kernels' grid barrier)
  TSAN_ONLY(__tsan_release(&g_sync_grid);)
  yield_as(WAIT_GRID);
  TSAN_ONLY(__tsan_acquire(&g_sync_grid);)
}
void barrier_warp() {
  char* o = &g_sync_warp[(cur->lin >> 5) & 63];
  (void)o;
  TSAN_ONLY(__tsan_release(o);)
  yield_as(WAIT_WARP);
  TSAN_ONLY(__tsan_acquire(o);)
}

static void fiber_main() {
  TSAN_ONLY(__tsan_acquire(&g_sync_kernel[g_kidx ^ 1]);)      // after everything before this launch ...
  (*g_body)();
  TSAN_ONLY(__tsan_release(&g_sync_kernel[g_kidx]);)          // ... and before everything after it; CTAs of one launch stay unordered
  TSAN_ONLY(__tsan_release(&g_sync_cta_end);)                 // the scheduler frees this CTA's shared memory after its last thread
  cur->state = DONE;
  to_sched();
}

static void switch_in(Fiber& f) {
  cur = &f;
  g_bid = g_ctas[f.cta].bid;
  g_dyn_smem = g_ctas[f.cta].dyn;
  TSAN_ONLY(__tsan_switch_to_fiber(f.ts, __tsan_switch_to_fiber_no_sync);)
  swapcontext(&g_sched, &f.ctx);
}

// runs the co-scheduled CTAs [0, g_ctas.size()) to completion
static void run_ctas() {
  const int n = g_tpc, nc = (int)g_ctas.size(), nf = (int)g_fibers.size();
  for (;;) {
    int alive = 0;
    for (int i = 0; i < nf; ++i) {
      Fiber& f = g_fibers[i];
      if (f.state == READY) switch_in(f);
      if (f.state != DONE) ++alive;
    }
    if (!alive) return;
    bool released = false;
    // warp barriers first: a warp whose alive lanes all wait at a warp collective continues
    for (int c = 0; c < nc; ++c)
      for (int w = 0; w < n; w += 32) {
        const int f0 = g_ctas[c].first;
        int a = 0, ww = 0;
        for (int l = w; l < w + 32 && l < n; ++l) { if (g_fibers[f0 + l].state != DONE) ++a; if (g_fibers[f0 + l].state == WAIT_WARP) ++ww; }
        if (a && ww == a) { for (int l = w; l < w + 32 && l < n; ++l) if (g_fibers[f0 + l].state == WAIT_WARP) g_fibers[f0 + l].state = READY; released = true; }
      }
    if (released) continue;
    for (int c = 0; c < nc; ++c) {
      const int f0 = g_ctas[c].first;
      int a = 0, wc = 0;
      for (int l = 0; l < n; ++l) { if (g_fibers[f0 + l].state != DONE) ++a; if (g_fibers[f0 + l].state == WAIT_CTA) ++wc; }
      if (a && wc == a) { for (int l = 0; l < n; ++l) g_fibers[f0 + l].state = READY; released = true; }
    }
    if (released) continue;
    // grid barrier: exactly one waiting thread in EVERY CTA of the launch, everybody else parked at a CTA barrier
    int wg = 0, wcta = 0;
    for (int i = 0; i < nf; ++i) { if (g_fibers[i].state == WAIT_GRID) ++wg; if (g_fibers[i].state == WAIT_CTA) ++wcta; }
    if (wg == nc && wg + wcta == alive) {
      for (int i = 0; i < nf; ++i) if (g_fibers[i].state == WAIT_GRID) g_fibers[i].state = READY;
      continue;
    }
    fprintf(stderr, "emu: deadlock (first CTA (%u,%u,%u) of %d co-scheduled): %d alive, %d at __syncthreads, %d at the grid barrier, "
            "rest inside a divergent warp collective or an exited CTA is missing from a grid barrier\n",
            g_ctas[0].bid.x, g_ctas[0].bid.y, g_ctas[0].bid.z, nc, alive, wcta, wg);
    abort();
  }
}

static void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, bool coop) {
  const int n = (int)(block.x * block.y * block.z);
  const int total = (int)(grid.x * grid.y * grid.z), group = coop ? total : 1;
  g_bdim = block; g_gdim = grid; g_body = &body; g_tpc = n;
  while ((int)g_stacks.size() < n * group) g_stacks.push_back((char*)aligned_alloc(64, STACK));
  g_kidx ^= 1;
  TSAN_ONLY(g_sched_ts = __tsan_get_current_fiber(); __tsan_release(&g_sync_kernel[g_kidx ^ 1]);)
  for (int c0 = 0; c0 < total; c0 += group) {
    g_ctas.assign(group, Cta());
    g_fibers.assign((size_t)n * group, Fiber());
    for (int c = 0; c < group; ++c) {
      const unsigned lin = (unsigned)(c0 + c);
      Cta& ct = g_ctas[c];
      ct.bid = {lin % grid.x, (lin / grid.x) % grid.y, lin / (grid.x * grid.y)};
      ct.dyn = (unsigned char*)fresh_block(smem);           // exact size: overruns are visible to AddressSanitizer
      ct.first = c * n;
      for (int i = 0; i < n; ++i) {
        Fiber& f = g_fibers[c * n + i];
        f.lin = i; f.state = READY; f.slot = 0; f.cta = c;
        f.tid = {(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = g_stacks[c * n + i]; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, fiber_main, 0);
        TSAN_ONLY(f.ts = __tsan_create_fiber(0);)
      }
    }
    run_ctas();
    TSAN_ONLY(for (size_t i = 0; i < g_fibers.size(); ++i) __tsan_destroy_fiber(g_fibers[i].ts);)
    TSAN_ONLY(__tsan_acquire(&g_sync_cta_end);)
#ifdef EMU_TSAN
    for (char*& st : g_stacks) { free(st); st = (char*)aligned_alloc(64, STACK); }   // stacks are reused by unrelated threads: renew
#endif
    for (Cta& ct : g_ctas) {
      free(ct.dyn);
      for (auto& kv : ct.blocks) free(kv.second);
    }
    g_ctas.clear();
  }
  TSAN_ONLY(__tsan_acquire(&g_sync_kernel[g_kidx]);)
  cur = nullptr; g_body = nullptr; g_dyn_smem = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) { launch_impl(grid, block, smem, body, false); }
// cooperative launch: ALL CTAs of the grid are co-scheduled (fibers of one OS thread), so grid barriers can complete
void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) { launch_impl(grid, block, smem, body, true); }
}  // namespace emu
