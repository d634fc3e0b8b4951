// Fused optimiser tail on a FLAT parameter group:
//   global-norm clip(max_norm) + AdamW(weight decay) [+ target EMA] [+ gradient zeroing] in three launches, no host sync.
// Replaces  nn.utils.clip_grad_norm_ + torch.optim.AdamW.step  (WorldModel.py:195-200, Agent.py:141-151) and
// Agent.soft_update_target (Agent.py:90-94).  HBM-bound streaming kernels: grids are multiples of the SM count, float4 accesses.
#include "internal.h"

namespace drm {
namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_BLOCKS = 148 * 4;   // partial sums per norm pass; also the cap of the apply grid

// stage 1: per-block partial sums of squares (fixed block -> element assignment: deterministic)
__global__ void __launch_bounds__(OPT_THREADS) sqnorm_partial_kernel(const float* __restrict__ g, long n, double* __restrict__ partials) {
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long i = (long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long)gridDim.x * OPT_THREADS) {
    const float4 v = __ldg(g4 + i);
    acc[0] = fmaf(v.x, v.x, acc[0]); acc[1] = fmaf(v.y, v.y, acc[1]);
    acc[2] = fmaf(v.z, v.z, acc[2]); acc[3] = fmaf(v.w, v.w, acc[3]);
  }
  double s = (double)acc[0] + (double)acc[1] + (double)acc[2] + (double)acc[3];
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s += (double)v * (double)v;
  }
  __shared__ double red[OPT_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// stage 2 (one block): total norm, clip coefficient, step counter, bias corrections -> state
//   state[0] step count (float, advanced only when the step is taken)   state[1] gradient norm of this call
//   state[2] clip coefficient applied                                    state[3] 1.0 when the step was skipped (non-finite)
//   state[4] step_size = lr / (1 - beta1^t)                              state[5] 1 / sqrt(1 - beta2^t)
__global__ void __launch_bounds__(OPT_THREADS) adamw_prepare_kernel(const double* __restrict__ partials, int n_partials,
                                                                    float* __restrict__ state, float lr, float beta1, float beta2,
                                                                    float max_norm, float* __restrict__ norm_only) {
  __shared__ double red[OPT_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) s += partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    if (norm_only) {   // drm_grad_norm: report, do not touch the optimiser state
      *norm_only = norm;
      return;
    }
    const bool ok = isfinite(norm);
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.0f;   // torch.nn.utils.clip_grad_norm_
    coef = fminf(coef, 1.0f);
    const float t = state[0] + (ok ? 1.0f : 0.0f);
    state[0] = t;
    state[1] = norm;
    state[2] = ok ? coef : 0.f;
    state[3] = ok ? 0.f : 1.f;
    const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
    state[4] = (float)((double)lr / bc1);
    state[5] = (float)(1.0 / sqrt(bc2));
  }
}

struct AdamwArgs {
  float lr, beta1, beta2, eps, weight_decay, tau;
  int zero_grad;
};

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamwArgs& a, float coef, float step_size,
                                           float inv_sqrt_bc2) {
  g *= coef;
  p *= 1.0f - a.lr * a.weight_decay;
  m = m + (1.0f - a.beta1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(1.0f - a.beta2, g * g, v * a.beta2);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
  p -= step_size * (m / denom);
}

// stage 3: the update.  A skipped step (non-finite gradient norm) leaves p, m, v untouched but still zeroes the gradient.
__global__ void __launch_bounds__(OPT_THREADS) adamw_apply_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                                  float* __restrict__ v, long n, const float* __restrict__ state,
                                                                  float* __restrict__ ema, AdamwArgs a) {
  const float coef = state[2], step_size = state[4], inv_sqrt_bc2 = state[5];
  const bool skip = state[3] != 0.f;
  const long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* e4 = reinterpret_cast<float4*>(ema);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long i = (long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long)gridDim.x * OPT_THREADS) {
    if (!skip) {
      float4 P = p4[i], M = m4[i], V = v4[i];
      const float4 G = g4[i];
      adamw_elem(P.x, G.x, M.x, V.x, a, coef, step_size, inv_sqrt_bc2);
      adamw_elem(P.y, G.y, M.y, V.y, a, coef, step_size, inv_sqrt_bc2);
      adamw_elem(P.z, G.z, M.z, V.z, a, coef, step_size, inv_sqrt_bc2);
      adamw_elem(P.w, G.w, M.w, V.w, a, coef, step_size, inv_sqrt_bc2);
      p4[i] = P; m4[i] = M; v4[i] = V;
      if (ema) {   // Agent.soft_update_target: target = (1 - tau) * target + tau * current
        float4 E = e4[i];
        E.x = E.x * (1.0f - a.tau) + a.tau * P.x; E.y = E.y * (1.0f - a.tau) + a.tau * P.y;
        E.z = E.z * (1.0f - a.tau) + a.tau * P.z; E.w = E.w * (1.0f - a.tau) + a.tau * P.w;
        e4[i] = E;
      }
    }
    if (a.zero_grad) g4[i] = zero;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (n4 << 2) + threadIdx.x;
    if (!skip) {
      float P = p[i], M = m[i], V = v[i];
      adamw_elem(P, g[i], M, V, a, coef, step_size, inv_sqrt_bc2);
      p[i] = P; m[i] = M; v[i] = V;
      if (ema) ema[i] = ema[i] * (1.0f - a.tau) + a.tau * P;
    }
    if (a.zero_grad) g[i] = 0.f;
  }
}

inline int grid_for4(long n) {
  const long want = ((n >> 2) + OPT_THREADS - 1) / OPT_THREADS;
  if (want <= 148) return (int)(want < 1 ? 1 : want);
  long g = (want / 148) * 148;
  return (int)(g > OPT_BLOCKS ? OPT_BLOCKS : g);
}

}  // namespace
}  // namespace drm

using namespace drm;

#define DRM_RC(x)         \
  do {                    \
    int rc__ = (x);       \
    if (rc__) return rc__; \
  } while (0)

extern "C" int64_t drm_adamw_scratch_bytes(void) { return (int64_t)OPT_BLOCKS * sizeof(double); }

extern "C" int drm_grad_norm(const float* grad, int64_t n, void* scratch, float* norm_out, void* stream) {
  DRM_RC(check_arch());
  DRM_REQUIRE(n >= 0, DRM_ERR_SHAPE, "drm_grad_norm: negative size");
  DRM_REQUIRE(scratch && norm_out, DRM_ERR_ARG, "drm_grad_norm: NULL scratch / output");
  DRM_REQUIRE(n == 0 || grad, DRM_ERR_ARG, "drm_grad_norm: NULL gradient");
  DRM_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0, DRM_ERR_ALIGN,
              "drm_grad_norm: gradient must be 16-byte aligned, scratch 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = grid_for4(n);
  sqnorm_partial_kernel<<<nb, OPT_THREADS, 0, st>>>(grad, n, static_cast<double*>(scratch));
  DRM_LAUNCH_CHECK();
  adamw_prepare_kernel<<<1, OPT_THREADS, 0, st>>>(static_cast<const double*>(scratch), nb, nullptr, 0.f, 0.f, 0.f, 0.f, norm_out);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state, void* scratch,
                              float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, float* ema_target,
                              float tau, int32_t zero_grad, void* stream) {
  DRM_RC(check_arch());
  DRM_REQUIRE(n >= 0, DRM_ERR_SHAPE, "drm_adamw_step: negative size");
  if (n == 0) return DRM_OK;
  DRM_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && scratch, DRM_ERR_ARG, "drm_adamw_step: NULL argument");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  DRM_REQUIRE(al16(param) && al16(grad) && al16(exp_avg) && al16(exp_avg_sq) && (!ema_target || al16(ema_target)) &&
                  (reinterpret_cast<uintptr_t>(scratch) & 7) == 0,
              DRM_ERR_ALIGN, "drm_adamw_step: flat buffers must be 16-byte aligned");
  DRM_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f, DRM_ERR_ARG,
              "drm_adamw_step: lr / betas / eps out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = grid_for4(n);
  sqnorm_partial_kernel<<<nb, OPT_THREADS, 0, st>>>(grad, n, static_cast<double*>(scratch));
  DRM_LAUNCH_CHECK();
  adamw_prepare_kernel<<<1, OPT_THREADS, 0, st>>>(static_cast<const double*>(scratch), nb, state, lr, beta1, beta2, max_norm, nullptr);
  DRM_LAUNCH_CHECK();
  AdamwArgs a{lr, beta1, beta2, eps, weight_decay, tau, zero_grad};
  adamw_apply_kernel<<<nb, OPT_THREADS, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, state, ema_target, a);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
