// Gradient all-reduce over NVLink / NVSwitch peer memory (one process per GPU, one node).
//   data-parallel training averages ONE flat fp32 gradient bucket per step (0.8 MB for
//   the yaml DeepSets, 0.27 MB GraphNet): latency bound.  An NCCL call costs a host launch outside the
//   captured train step plus ~20 us on the device; this kernel is a plain CUDA kernel, so it is captured
//   into the same CUDA graph as forward + backward, and does a ONE-SHOT all-reduce:
//     A  every rank copies its bucket into its own staging buffer (CUDA IPC memory mapped by all peers),
//     B  system-scope release of a sequence number into every peer's flag slot; wait for all peers' flags,
//     C  every rank sums the staging buffers of all ranks (in rank order: bitwise identical results on
//        every rank) straight over NVLink loads and writes the average back into its own bucket.
//   Two staging buffers alternate from step to step (sequence parity): a rank may run ahead and refill
//   buffer (s+1)&1 while a slow peer still reads buffer s&1; it cannot reach step s+2 before that peer has
//   signalled step s+1, i.e. finished reading step s.
//   The grid is at most one CTA per SM (all CTAs co-resident: phase B spins inside the kernel).
#include "pcc_common.cuh"

namespace pcc {

constexpr int kPeerMax = 16;
constexpr int kPeerThreads = 512;

struct PeerParams {
  float* bucket;              // local gradient bucket, reduced in place
  int64_t n;                  // floats
  uint8_t* stage[kPeerMax];   // staging region of every rank (peer mappings; [rank] is the local one)
  int64_t buf_bytes;          // bytes of ONE staging buffer (two per region, after the 1 KB flag block)
  int rank, world;
  float scale;                // 1 / world for the average
  unsigned int* counters;     // local scratch: [0] = sequence number, [1] = CTA arrival counter
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cv_f4(const float4* p) {  // never served from a stale cache line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kPeerThreads, 1) peer_allreduce_kernel(const PeerParams p) {
  const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(p.counters) + 1;  // this step (1, 2, ...)
  const int64_t n4 = p.n / 4;  // the bucket is padded to a multiple of 4 floats by the caller
  uint8_t* mine = p.stage[p.rank] + 1024 + (seq & 1) * p.buf_bytes;

  // ---- A: bucket -> own staging buffer
  const float4* src = reinterpret_cast<const float4*>(p.bucket);
  float4* dst = reinterpret_cast<float4*>(mine);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
  __syncthreads();
  // ---- B: the last CTA to finish A publishes the sequence number to every peer; everyone waits for all peers
  //      (one system-scope fence per CTA, after the barrier: it is cumulative over the CTA's stores)
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int arrived = atomicAdd(p.counters + 1, 1u) + 1;
    if (arrived == gridDim.x * seq) {  // the arrival counter is never reset: seq-th multiple of the grid size
      // ONE system-scope fence, then relaxed flag stores: a release per peer serialised `world` fences (43 us at 8
      // ranks against 23 us at 2)
      __threadfence_system();
      for (int r = 0; r < p.world; ++r) st_relaxed_sys(reinterpret_cast<unsigned int*>(p.stage[r]) + p.rank, seq);
    }
  }
  if (threadIdx.x < p.world) {
    const unsigned int* flag = reinterpret_cast<const unsigned int*>(p.stage[p.rank]) + threadIdx.x;
    // bounded: a peer that died or never launched this step makes the kernel TRAP (a CUDA error on every healthy
    // rank) after ~10 s instead of spinning forever and hanging the node
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
      if (clock64() - t0 > 20000000000ll) __trap();
    }
  }
  __syncthreads();
  // ---- C: sum over ranks in rank order, scale, write back
  float4* out = reinterpret_cast<float4*>(p.bucket);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // all peers' loads of a batch are issued before the first add: one NVLink round trip, not `world` of them
#pragma unroll 1
    for (int rb = 0; rb < p.world; rb += 8) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = rb + j;
        v[j] = (r < p.world) ? ld_cv_f4(reinterpret_cast<const float4*>(p.stage[r] + 1024 + (seq & 1) * p.buf_bytes) + i)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    out[i] = make_float4(acc.x * p.scale, acc.y * p.scale, acc.z * p.scale, acc.w * p.scale);
  }
  // ---- the sequence number advances once per launch (last CTA out)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int left = atomicAdd(p.counters + 2, 1u) + 1;
    if (left == gridDim.x * seq) *reinterpret_cast<volatile unsigned int*>(p.counters) = seq;
  }
}

}  // namespace pcc

using namespace pcc;

// Staging memory must be exportable through CUDA IPC, so the library allocates it (cudaMalloc) — the one
// exception to "the caller owns all memory".  Layout: 1 KB of flags + 2 staging buffers of buf_bytes.
extern "C" int pcc_peer_alloc(int64_t buf_bytes, void** region, void* ipc_handle_64, int device) {
  PCC_ENTER(device);
  PCC_REQUIRE(region && ipc_handle_64 && buf_bytes > 0, "bad arguments");
  void* ptr = nullptr;
  const size_t total = 1024 + 2 * (size_t)buf_bytes;
  PCC_CUDA(cudaMalloc(&ptr, total));
  PCC_CUDA(cudaMemset(ptr, 0, total));
  PCC_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    cudaFree(ptr);
    return fail(__func__, cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_64, &h, 64);
  *region = ptr;
  return 0;
}

extern "C" int pcc_peer_open(const void* ipc_handle_64, void** region, int device) {
  PCC_ENTER(device);
  PCC_REQUIRE(region && ipc_handle_64, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  void* ptr = nullptr;
  PCC_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  *region = ptr;
  return 0;
}

extern "C" int pcc_peer_close(void* region, int device) {
  PCC_ENTER(device);
  if (region) PCC_CUDA(cudaIpcCloseMemHandle(region));
  return 0;
}

extern "C" int pcc_peer_free(void* region, int device) {
  PCC_ENTER(device);
  if (region) PCC_CUDA(cudaFree(region));
  return 0;
}

// bucket[n] (n % 4 == 0) <- scale * sum over ranks.  regions[world]: staging regions (regions[rank] local, the
// others opened with pcc_peer_open), counters: 16 zero-initialised bytes of local device memory that persist
// across calls.  Collective: every rank must call it the same number of times.  Capturable in a CUDA graph.
extern "C" int pcc_peer_allreduce(float* bucket, int64_t n, void* const* regions, int64_t buf_bytes, int rank, int world,
                                  float scale, void* counters, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world, "bad rank / world");
  PCC_REQUIRE(n % 4 == 0 && n * 4 <= buf_bytes, "bucket must be a multiple of 4 floats and fit the staging buffer");
  PCC_REQUIRE(((uintptr_t)bucket & 15) == 0, "bucket must be 16-byte aligned");
  if (n == 0) return 0;
  PeerParams p{};
  p.bucket = bucket; p.n = n; p.buf_bytes = buf_bytes; p.rank = rank; p.world = world; p.scale = scale;
  p.counters = (unsigned int*)counters;
  for (int r = 0; r < world; ++r) p.stage[r] = (uint8_t*)regions[r];
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int64_t grid = cdiv(n / 4, kPeerThreads);
  if (grid > sms) grid = sms;
  PCC_K(peer_allreduce_kernel)<<<(unsigned)grid, kPeerThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch(__func__);
}
