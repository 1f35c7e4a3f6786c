// Shared helpers for libpcc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <stdio.h>

#include "../../include/pcc.h"

namespace pcc {

void set_error(const std::string& msg);
void note_launch(int kernels);  // launch accounting (pcc_launch_count)

// CUDA-event bracket around one of the big fused kernels (pcc_prof_*); no-op unless enabled.
// Not usable while the stream is being captured into a CUDA graph.
struct ProfScope {
  int slot;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(int slot, cudaStream_t st);
  ~ProfScope();
};

// kernel launch with accounting: PCC_K(kernel<...>)<<<grid, block, smem, stream>>>(args)
#define PCC_K(...) (pcc::note_launch(1), (__VA_ARGS__))

// Programmatic dependent launch for the kernels of the DeepSets train step (14 back-to-back launches whose
// hand-over gaps are ~6 % of the step).  A kernel launched with launch_dep may be SCHEDULED while its predecessor
// in the stream is still running; it calls pdl_enter() as its first statement, which blocks until the predecessor
// has completed and flushed (griddepcontrol.wait) and then lets its own successor be scheduled
// (griddepcontrol.launch_dependents).  Because every kernel waits before its first global access, completion is
// transitive (a kernel cannot finish before its predecessor has) and stream order is preserved for every buffer,
// including the caching allocator's reuse.  Both instructions are no-ops in a kernel launched without the
// attribute.  PCC_PDL=0 launches without it.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KA, typename... A>
inline void launch_dep(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  note_launch(1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);  // errors surface through check_launch()
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    ok = (cudaSetDevice(device) == cudaSuccess);
    if (!ok) { cudaGetLastError(); }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline int fail(const char* where, const std::string& msg) {
  set_error(std::string(where) + ": " + msg);
  return -1;
}

inline int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
  return 0;
}

#define PCC_ENTER(device)                                                                  \
  pcc::DeviceGuard _guard(device);                                                         \
  if (!_guard.ok) return pcc::fail(__func__, "cannot select CUDA device (no CPU fallback exists)")

#define PCC_REQUIRE(cond, msg)                       \
  do {                                               \
    if (!(cond)) return pcc::fail(__func__, msg);    \
  } while (0)

#define PCC_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t _e = (call);                                                \
    if (_e != cudaSuccess) return pcc::fail(__func__, cudaGetErrorString(_e)); \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- activations (fp32)
__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case PCC_ACT_RELU: return z > 0.f ? z : 0.f;
    case PCC_ACT_GELU: return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
    case PCC_ACT_SILU: return z / (1.f + expf(-z));
    case PCC_ACT_TANH: return tanhf(z);
    default: return z;
  }
}

// derivative of act at pre-activation z
__device__ __forceinline__ float act_grad(int act, float z) {
  switch (act) {
    case PCC_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case PCC_ACT_GELU: {
      float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
      float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case PCC_ACT_SILU: {
      float s = 1.f / (1.f + expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    case PCC_ACT_TANH: {
      float t = tanhf(z);
      return 1.f - t * t;
    }
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pcc
