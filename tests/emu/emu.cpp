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
enum { READY = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int lin, state;
  unsigned long long slot;
  void* ts;                 // ThreadSanitizer fiber context (EMU_TSAN)
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
static std::map<const void*, void*> g_cta_blocks;            // static __shared__ storage of the running CTA, by declaration
static void* g_sched_ts = nullptr;
static char g_sync_cta, g_sync_warp[64], g_sync_kernel[2], g_sync_cta_end;
static int g_kidx = 0;

static void* fresh_block(size_t n) {                            // a NEW heap block: undefined contents, no race history
  void* p = aligned_alloc(64, ((n ? n : 1) + 63) / 64 * 64);
  volatile unsigned char* q = (volatile unsigned char*)p;      // plain loop: memset() would be intercepted by the sanitizers
  for (size_t i = 0; i < n; ++i) q[i] = 0xCD;                   // and attributed to whichever thread got here first
  return p;
}
void* cta_shared(const void* key, size_t bytes) {
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
  const int me = cur->lin, base = me & ~31;
  g_fibers[me].slot = raw;
  barrier_warp();
  const int src = base + (src_lane_in_warp & 31);
  const unsigned long long r = src < (int)g_fibers.size() ? g_fibers[src].slot : raw;
  barrier_warp();
  return r;
}
int all(int pred) {
  const int me = cur->lin, base = me & ~31, n = (int)g_fibers.size();
  g_fibers[me].slot = pred ? 1ull : 0ull;
  barrier_warp();
  int a = 1;
  for (int l = base; l < base + 32 && l < n; ++l) a &= (int)g_fibers[l].slot;
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

static void run_cta() {
  const int n = (int)g_fibers.size();
  for (;;) {
    int alive = 0;
    for (int i = 0; i < n; ++i) {
      Fiber& f = g_fibers[i];
      if (f.state == READY) {
        cur = &f;
        TSAN_ONLY(__tsan_switch_to_fiber(f.ts, __tsan_switch_to_fiber_no_sync);)
        swapcontext(&g_sched, &f.ctx);
      }
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
  g_fibers.assign(n, Fiber());
  g_kidx ^= 1;
  TSAN_ONLY(g_sched_ts = __tsan_get_current_fiber(); __tsan_release(&g_sync_kernel[g_kidx ^ 1]);)
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_bid = {bx, by, bz};
        g_dyn_smem = (unsigned char*)fresh_block(smem);      // exact size: overruns are visible to AddressSanitizer
        for (int i = 0; i < n; ++i) {
          Fiber& f = g_fibers[i];
          f.lin = i; f.state = READY; f.slot = 0;
          f.tid = {(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = g_stacks[i]; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &g_sched;
          makecontext(&f.ctx, fiber_main, 0);
          TSAN_ONLY(f.ts = __tsan_create_fiber(0);)
        }
        run_cta();
        TSAN_ONLY(for (int i = 0; i < n; ++i) __tsan_destroy_fiber(g_fibers[i].ts);)
        TSAN_ONLY(__tsan_acquire(&g_sync_cta_end);)
#ifdef EMU_TSAN
        for (char*& st : g_stacks) { free(st); st = (char*)aligned_alloc(64, STACK); }   // stacks are reused by unrelated threads: renew
#endif
        free(g_dyn_smem);
        for (auto& kv : g_cta_blocks) free(kv.second);
        g_cta_blocks.clear();
      }
  TSAN_ONLY(__tsan_acquire(&g_sync_kernel[g_kidx]);)
  cur = nullptr; g_body = nullptr; g_dyn_smem = nullptr;
}
}  // namespace emu
