// Library-level entry points: error string, version, device check.
#include "pcc_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace pcc {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

static std::atomic<long long> g_launches{0};
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("PCC_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
void note_launch(int kernels) { g_launches.fetch_add(kernels, std::memory_order_relaxed); }

static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfRec { int slot; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(int slot_, cudaStream_t st_) : slot(slot_), st(st_) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; cudaGetLastError(); return; }
  cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEventRecord(e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back({slot, e0, e1});
}
}  // namespace pcc

extern "C" int64_t pcc_launch_count(int reset) {
  return reset ? pcc::g_launches.exchange(0) : pcc::g_launches.load();
}

extern "C" int pcc_prof_enable(int on) {
  pcc::g_prof_on.store(on ? 1 : 0);
  return 0;
}

// sums the recorded brackets of `slot` (caller synchronises first); clears them
extern "C" int pcc_prof_read(int slot, double* ms_total, int64_t* count) {
  std::lock_guard<std::mutex> lk(pcc::g_prof_mu);
  double tot = 0.0;
  int64_t n = 0;
  std::vector<pcc::ProfRec> keep;
  for (auto& r : pcc::g_prof) {
    if (r.slot != slot) { keep.push_back(r); continue; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { tot += ms; ++n; } else { cudaGetLastError(); }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  pcc::g_prof.swap(keep);
  if (ms_total) *ms_total = tot;
  if (count) *count = n;
  return 0;
}

extern "C" const char* pcc_last_error(void) { return pcc::g_last_error.c_str(); }

extern "C" int pcc_debug_set_pdl(int on) {
  pcc::g_pdl = on ? 1 : 0;
  return 0;
}

extern "C" int pcc_version(void) { return 100; }

extern "C" int pcc_check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return pcc::fail(__func__, "no CUDA device visible; libpcc has no CPU fallback");
  }
  if (device < 0 || device >= count) return pcc::fail(__func__, "device index out of range");
  cudaDeviceProp prop;
  PCC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return pcc::fail(__func__, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                   "; libpcc is built for sm_100a (B200) only");
  }
  return 0;
}
