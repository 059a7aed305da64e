// pdl.cuh -- programmatic dependent launch (PDL) for the chains of small shell kernels (conv1x1 -> BatchNorm -> ... ).
//
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start as soon as every CTA of its predecessor in
// the stream has executed griddepcontrol.launch_dependents (or exited); it must execute griddepcontrol.wait before it touches
// memory -- that returns when the predecessor grid has completed and flushed.  Both are the FIRST instructions of the kernels here,
// so the semantics equal a normal launch and only the launch latency / CTA scheduling of kernel i+1 overlaps the tail of kernel i.
// Kernels that were not launched with the attribute execute both instructions as no-ops.  Opt-in: MT_PDL=1 (A/B in
// profiles/r3_notes.md).  Inside a CUDA-graph capture the attribute becomes a programmatic dependency edge.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_entry() { pdl_wait(); pdl_launch_dependents(); }

inline int mt_pdl_enabled() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("MT_PDL"); mode = e ? (atoi(e) != 0) : 0; }
  return mode;
}

#if defined(__CUDACC__)
// launch with the PDL attribute; `cluster` > 1 adds a cluster dimension (x)
template <typename K, typename... Args>
inline cudaError_t mt_launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (mt_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif
