// Fused training-step tail (reference experiments/base_experiment.py:573-597):
// global gradient-norm clipping, the AdamW update and the refresh of the bf16 tensor-core operand copy of
// every GEMM weight in ONE multi-tensor pass: each parameter element is read once (p, g, m, v) and
// written once (p, m, v, + bf16 copy) instead of the ~4 passes of clip_grad_norm_ + AdamW + recast.
// Arithmetic follows torch.optim.AdamW (decoupled weight decay, bias correction, eps outside the sqrt)
// and torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))).
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  const int64_t n4 = n / 4;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) acc += x[i] * x[i];
  } else {
    for (int64_t i = tid; i < n; i += stride) acc += x[i] * x[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

struct AdamArgs {
  const float* norm_sq;  // device scalar: sum of squared gradients (null: no clipping)
  float max_norm;
  float lr, beta1, beta2, eps, weight_decay;
  float bc1, bc2_sqrt;   // 1 - beta1^t, sqrt(1 - beta2^t) from the host step count ...
  const int* step_dev;   // ... or, when non-null, computed from this device step counter (CUDA-graph replay)
  const float* lr_dev;   // optional device learning rate overriding `lr` (schedulers under graph replay)
  // exponential moving average of the parameters (torch_ema arithmetic, reference experiments/base_experiment.py:
  // 127-134, :594): shadow -= (1 - d) (shadow - p), d = min(decay, (1 + n) / (10 + n)), n = updates so far
  float ema_decay;           // <= 0: no EMA
  int ema_updates;           // n from the host ...
  const int* ema_updates_dev;  // ... or from this device counter (CUDA-graph replay)
};

__device__ __forceinline__ float ema_one_minus_decay(const AdamArgs& a) {
  const float n = (float)(a.ema_updates_dev ? *a.ema_updates_dev : a.ema_updates);
  const float d = fminf(a.ema_decay, (1.f + n) / (10.f + n));
  return 1.f - d;
}

__global__ void counter_increment_kernel(int* c) {
  pdl_wait(); *c += 1; }

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float coef) {
  g *= coef;
  p *= 1.f - a.lr * a.weight_decay;
  m = m + (g - m) * (1.f - a.beta1);           // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + (1.f - a.beta2) * g * g;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  return p;
}

// grid (chunks, jobs): one parameter tensor per blockIdx.y
__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const v4h_adamw_job* __restrict__ jobs, AdamArgs a) {
  pdl_wait();
  const v4h_adamw_job j = jobs[blockIdx.y];
  if (a.step_dev) {
    const float t = (float)*a.step_dev;
    a.bc1 = 1.f - powf(a.beta1, t);
    a.bc2_sqrt = sqrtf(1.f - powf(a.beta2, t));
  }
  if (a.lr_dev) a.lr = *a.lr_dev;
  float coef = 1.f;
  if (a.norm_sq) {
    const float c = a.max_norm / (sqrtf(*a.norm_sq) + 1e-6f);
    coef = c < 1.f ? c : 1.f;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bf16* dst = reinterpret_cast<bf16*>(j.bf16_dst);
  const bool vec = ((reinterpret_cast<uintptr_t>(j.p) | reinterpret_cast<uintptr_t>(j.g) | reinterpret_cast<uintptr_t>(j.m) |
                     reinterpret_cast<uintptr_t>(j.v) | reinterpret_cast<uintptr_t>(j.f32_dst)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  const bool ema = a.ema_decay > 0.f && j.ema != nullptr;
  const float omd = ema ? ema_one_minus_decay(a) : 0.f;
  const int64_t n4 = (vec && (reinterpret_cast<uintptr_t>(j.ema) & 15) == 0) ? j.n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(j.p)[i];
    const float4 g = reinterpret_cast<const float4*>(j.g)[i];
    float4 m = reinterpret_cast<float4*>(j.m)[i];
    float4 v = reinterpret_cast<float4*>(j.v)[i];
    adam_one(p.x, g.x, m.x, v.x, a, coef);
    adam_one(p.y, g.y, m.y, v.y, a, coef);
    adam_one(p.z, g.z, m.z, v.z, a, coef);
    adam_one(p.w, g.w, m.w, v.w, a, coef);
    reinterpret_cast<float4*>(j.p)[i] = p;
    reinterpret_cast<float4*>(j.m)[i] = m;
    reinterpret_cast<float4*>(j.v)[i] = v;
    if (j.f32_dst) reinterpret_cast<float4*>(j.f32_dst)[i] = p;
    if (ema) {
      float4 e = reinterpret_cast<float4*>(j.ema)[i];
      e.x -= omd * (e.x - p.x); e.y -= omd * (e.y - p.y); e.z -= omd * (e.z - p.z); e.w -= omd * (e.w - p.w);
      reinterpret_cast<float4*>(j.ema)[i] = e;
    }
    if (dst) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      reinterpret_cast<uint2*>(dst)[i] =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  for (int64_t i = n4 * 4 + tid; i < j.n; i += stride) {
    float p = j.p[i], m = j.m[i], v = j.v[i];
    adam_one(p, j.g[i], m, v, a, coef);
    j.p[i] = p; j.m[i] = m; j.v[i] = v;
    if (dst) dst[i] = __float2bfloat16_rn(p);
    if (j.f32_dst) j.f32_dst[i] = p;
    if (ema) { const float e = j.ema[i]; j.ema[i] = e - omd * (e - p); }
  }
}

// stand-alone EMA update (ExponentialMovingAverage.update() when it is not fused into the optimizer pass)
__global__ void __launch_bounds__(OPT_THREADS) ema_kernel(const v4h_adamw_job* __restrict__ jobs, AdamArgs a) {
  pdl_wait();
  const v4h_adamw_job j = jobs[blockIdx.y];
  if (j.ema == nullptr) return;
  const float omd = ema_one_minus_decay(a);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(j.p) | reinterpret_cast<uintptr_t>(j.ema)) & 15) == 0;
  const int64_t n4 = vec ? j.n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 p = reinterpret_cast<const float4*>(j.p)[i];
    float4 e = reinterpret_cast<float4*>(j.ema)[i];
    e.x -= omd * (e.x - p.x); e.y -= omd * (e.y - p.y); e.z -= omd * (e.z - p.z); e.w -= omd * (e.w - p.w);
    reinterpret_cast<float4*>(j.ema)[i] = e;
  }
  for (int64_t i = n4 * 4 + tid; i < j.n; i += stride) { const float e = j.ema[i]; j.ema[i] = e - omd * (e - j.p[i]); }
}

}  // namespace

int grad_norm_sq(const float* flat, int64_t n, float* out, cudaStream_t s) {
  V4H_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  int64_t blocks = ceil_div(n, (int64_t)OPT_THREADS * 16);
  if (blocks < 1) blocks = 1;
  if (blocks > 1184) blocks = 1184;
  V4H_CUDA(launch_pdl(sumsq_kernel, dim3((unsigned)blocks), dim3(OPT_THREADS), 0, s, flat, n, out));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int counter_increment(int* counter, cudaStream_t s) {
  V4H_CUDA(launch_pdl(counter_increment_kernel, dim3(1), dim3(1), 0, s, counter));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int ema_update(const v4h_adamw_job* jobs_dev, int njobs, int64_t max_n, float decay, int num_updates,
               const int* num_updates_dev, cudaStream_t s) {
  AdamArgs a;
  memset(&a, 0, sizeof(a));
  a.ema_decay = decay; a.ema_updates = num_updates; a.ema_updates_dev = num_updates_dev;
  int64_t gx = ceil_div(max_n, (int64_t)OPT_THREADS * 4 * 4);
  if (gx < 1) gx = 1;
  if (gx > 128) gx = 128;
  V4H_CUDA(launch_pdl(ema_kernel, dim3(dim3((unsigned)gx, (unsigned)njobs)), dim3(OPT_THREADS), 0, s, jobs_dev, a));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int adamw_step(const v4h_adamw_job* jobs_dev, int njobs, int64_t max_n, const float* norm_sq, float max_norm, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, const int* step_dev,
               const float* lr_dev, float ema_decay, int ema_updates, const int* ema_updates_dev, cudaStream_t s) {
  AdamArgs a;
  a.ema_decay = ema_decay; a.ema_updates = ema_updates; a.ema_updates_dev = ema_updates_dev;
  a.step_dev = step_dev; a.lr_dev = lr_dev;
  a.norm_sq = norm_sq; a.max_norm = max_norm;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, (float)step);
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  int64_t gx = ceil_div(max_n, (int64_t)OPT_THREADS * 4 * 4);
  if (gx < 1) gx = 1;
  if (gx > 128) gx = 128;
  V4H_CUDA(launch_pdl(adamw_kernel, dim3(dim3((unsigned)gx, (unsigned)njobs)), dim3(OPT_THREADS), 0, s, jobs_dev, a));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace v4h
