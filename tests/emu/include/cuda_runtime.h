// TEST INFRASTRUCTURE -- a minimal CUDA-on-CPU shim so that the product's kernels (medical-transformer_b200/csrc, rewritten
// by tests/emu/prep.py: launches -> emu::launch, dynamic shared memory, the inline-PTX helpers) compile with g++ and run on
// the host for the "-m 'not gpu'" tier.  One CTA at a time; the threads of a CTA are ucontext fibers on ONE OS thread
// (deterministic; __syncthreads / warp collectives are scheduler barriers; atomics are plain read-modify-writes).
// It checks kernel LOGIC (indexing, barriers, reductions, the C ABI plumbing) against the oracle -- not performance, not
// PTX.  Never loaded by the product (medt_b200/_cabi.py only knows libmedt_b200.so).
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>

#define MT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static     /* only for forms prep.py does not rewrite; static __shared__ storage normally becomes emu::cta_shared() */

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }      // a 4-"SM" device: small co-scheduled grids
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
// cooperative launches of the product (axial_coop.cuh) become co-scheduled emulator launches
#define MT_COOP_LAUNCH(kern, grid, block, smem, stream, ...) \
  (emu::launch_coop(dim3(grid), dim3(block), (size_t)(smem), [&]() { (kern)(__VA_ARGS__); }), g_last_cuda = cudaSuccess)
static inline long long clock64() { return 0; }
template <class T> static inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& sym, size_t n) { memcpy(dst, &sym, n); return cudaSuccess; }
#define __noinline__
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

namespace emu {
struct Fiber;
extern Fiber* cur;
extern uint3 g_bid;
extern dim3 g_bdim, g_gdim;
extern unsigned char* g_dyn_smem;
const uint3& tid();
int lin();                                  // linear thread index inside the CTA
int nthreads();
void barrier_cta();
void barrier_warp();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);   // all CTAs co-scheduled (grid barriers)
void barrier_grid();                        // one thread per CTA
void* cta_shared(const void* key, size_t bytes);   // static __shared__ storage: one fresh, poisoned heap block per CTA and declaration
unsigned long long exchange(unsigned long long raw, int src_lane_in_warp);   // warp shuffle (all alive lanes of the warp call it)
int all(int pred);
}  // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::g_bid)
#define blockDim (emu::g_bdim)
#define gridDim (emu::g_gdim)

static inline void __syncthreads() { emu::barrier_cta(); }
static inline void __threadfence() {}
static inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
#define __cvta_generic_to_shared(p) ((size_t)(p))

template <class T> static inline T emu_exchange(T v, int src_lane_in_warp) {
  static_assert(sizeof(T) <= 8, "emu shuffle: 8-byte values at most");
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  raw = emu::exchange(raw, src_lane_in_warp);
  T r;
  memcpy(&r, &raw, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_exchange(v, (emu::lin() & 31) ^ lane_mask); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src_lane) { return emu_exchange(v, src_lane); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu_exchange(0, (int)(emu::lin() & 31)); }     // warp collective == scheduler barrier
static inline int __all_sync(unsigned, int pred) { return emu::all(pred); }
// atomics go through the compiler builtins so that ThreadSanitizer (tests/emu/run_tsan.sh) sees them as atomics
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
  float f;
  do { memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  memcpy(&f, &old, 4);
  return f;
}
static inline double atomicAdd(double* p, double v) {
  uint64_t* u = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(u, __ATOMIC_SEQ_CST), nw;
  double f;
  do { memcpy(&f, &old, 8); f += v; memcpy(&nw, &f, 8); } while (!__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  memcpy(&f, &old, 8);
  return f;
}

template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
