// placeholder until the tcgen05 kernel lands (replaced below in the same round)
#include "pcc_common.cuh"
extern "C" int pcc_phi_fused_supported(const pcc_phi_desc* d) { (void)d; return pcc::fail(__func__, "fused path not built"); }
extern "C" int64_t pcc_phi_fused_workspace_bytes(const pcc_phi_desc* d, int64_t n, int64_t B) { (void)d; (void)n; (void)B; return 0; }
extern "C" int pcc_deepsets_phi_pool_fwd(const pcc_phi_desc*, const float*, const int64_t*, int64_t, int64_t, float*, int32_t*, void*, int, void*) { return pcc::fail(__func__, "fused path not built"); }
extern "C" int pcc_deepsets_phi_pool_bwd(const pcc_phi_desc*, const float*, const int64_t*, int64_t, int64_t, const float*, const int32_t*, float* const*, float* const*, void*, int, void*) { return pcc::fail(__func__, "fused path not built"); }
