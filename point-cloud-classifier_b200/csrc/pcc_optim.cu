// Fused multi-tensor Adam / AdamW step.
//   reference: /root/reference/models/wrapper.py:30-33 constructs torch.optim.Adam / torch.optim.AdamW with default
//   hyper-parameters and :70 calls optimizer.step() — ~10 foreach launches per step for the yaml model, a visible
//   fraction of a 0.24 ms training step.  Here every parameter tensor is updated by ONE launch driven by a pointer
//   table; the moments live in two flat fp32 buffers; the step counter lives on the device, so the update can be
//   captured in the CUDA graph of the train step.  Operation order follows torch's single-tensor implementation
//   (torch/optim/adam.py, adamw.py): decay, exp_avg.lerp, exp_avg_sq.mul.addcmul, bias corrections in double,
//   denom = sqrt(v) / sqrt(bc2) + eps, param -= (lr / bc1) * m / denom.
#include "pcc_common.cuh"

namespace pcc {

struct AdamArgs {
  const int64_t* table;  // device [n][4]: param pointer, grad pointer, numel, offset into the flat moment buffers
  int n;
  int64_t total;
  float* m;
  float* v;
  const int64_t* step;   // device scalar: number of steps INCLUDING this one (incremented by the launch before)
  float lr, beta1, beta2, eps, wd;
  int decoupled;
};

__global__ void adam_bump_step_kernel(int64_t* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  __shared__ double s_bc1, s_bc2s;
  if (threadIdx.x == 0) {
    const double t = (double)*a.step;
    s_bc1 = 1.0 - pow((double)a.beta1, t);
    s_bc2s = sqrt(1.0 - pow((double)a.beta2, t));
  }
  __syncthreads();
  const float step_size = (float)((double)a.lr / s_bc1);
  const float bc2s = (float)s_bc2s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = a.n;  // tensor of flat index i: last t with offset[t] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(a.table + 4 * mid + 3) <= i) lo = mid; else hi = mid;
    }
    const int64_t off = __ldg(a.table + 4 * lo + 3);
    const int64_t j = i - off;
    if (j >= __ldg(a.table + 4 * lo + 2)) continue;  // padding between tensors
    float* p = reinterpret_cast<float*>(__ldg(a.table + 4 * lo + 0)) + j;
    const float* gp = reinterpret_cast<const float*>(__ldg(a.table + 4 * lo + 1));
    if (gp == nullptr) continue;                      // parameter without a gradient this step: untouched, like torch
    float g = gp[j];
    float w = *p;
    if (a.wd != 0.f) {
      if (a.decoupled) w = w * (1.f - a.lr * a.wd);   // AdamW: param.mul_(1 - lr * weight_decay)
      else g = g + a.wd * w;                          // Adam: grad = grad.add(param, alpha=weight_decay)
    }
    float m = a.m[i], v = a.v[i];
    m = m + (g - m) * (1.f - a.beta1);                // exp_avg.lerp_(grad, 1 - beta1)
    v = v * a.beta2 + (1.f - a.beta2) * g * g;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / bc2s + a.eps;
    w = w - step_size * (m / denom);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
    a.m[i] = m;
    a.v[i] = v;
    *p = w;
  }
}

}  // namespace pcc

using namespace pcc;

extern "C" int pcc_adam_step(const int64_t* table, int n_tensors, int64_t total, float* exp_avg, float* exp_avg_sq,
                             int64_t* step, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int decoupled, int device, void* stream) {
  PCC_ENTER(device);
  PCC_REQUIRE(n_tensors >= 0 && total >= 0, "negative size");
  PCC_REQUIRE(table && exp_avg && exp_avg_sq && step, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  PCC_K(adam_bump_step_kernel)<<<1, 1, 0, st>>>(step);
  if (n_tensors == 0 || total == 0) return check_launch(__func__);
  AdamArgs a{table, n_tensors, total, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay, decoupled};
  int64_t blocks = cdiv(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  PCC_K(adam_step_kernel)<<<(unsigned)blocks, 256, 0, st>>>(a);
  return check_launch(__func__);
}
