// Pre-processing of raw calorimeter showers on the GPU (the data feed): the FORWARD pass
// of the CaloChallenge ds2 / ds3 shape-model transform chain that the reference's dataset applies on the CPU when
// it is constructed (reference experiments/calochallenge/datasets.py:44-47, chain of
// configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28, classes in experiments/calochallenge/transforms.py):
//   NormalizeByElayer, ScaleTotalEnergy, CutValues (identity forwards), ExclusiveLogitTransform(rescale),
//   GlobalStandardizeFromFile, LogEnergy, ScaleEnergy, AddFeaturesToCond, Reshape
// One warp per (shower, layer) keeps the layer in registers: it sums the voxels, writes logit(voxel / layer energy)
// and leaves the layer energy for a one-thread-per-shower kernel that forms the u features (layers longer than 1024
// voxels: a generic one-CTA-per-shower kernel with two sweeps).  GlobalStandardizeFromFile either applies a known
// (mean, std) in the same sweep, or — its `written == False` branch, transforms.py:55-63 — the sweep accumulates
// count / sum / sum of squares of the non-saturated features in fp64, a one-thread kernel turns them into
// (mean, unbiased std) on the device, and a second elementwise kernel standardises in place.
// HBM-bound: 4 B read + 4 B written per voxel, + 8 B per voxel for the second kernel.
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int PRE_THREADS = 256;
constexpr int PRE_MAX_LAYERS = 128;

struct PreArgs {
  const float* showers;  // (n, voxels) raw energies per voxel
  const float* e_inc;    // (n) incident energies
  int n, voxels, n_layers;
  const int* bounds;     // (n_layers + 1) voxel offsets of the layers (device)
  float eps;             // NormalizeByElayer
  float factor;          // ScaleTotalEnergy (u_0 only)
  float delta, one_minus_2delta;  // ExclusiveLogitTransform(rescale=True)
  float alpha, e_min, e_scale;    // LogEnergy, ScaleEnergy (e_max - e_min)
  const float* mean_std; // device [2] or nullptr: standardise in this sweep
  double* stats;         // device [3] (count, sum, sum of squares) or nullptr
  float* x;              // (n, voxels)
  float* cond;           // (n, n_layers + 1): u features, then the scaled log incident energy
};

// ExclusiveLogitTransform forwards (transforms.py:11-17): z = x * (1 - 2 delta) + delta, then log(z / (1 - z)) as
// separate IEEE operations like the reference's tensor ops
__device__ __forceinline__ float logit_rescaled(float v, float one_minus_2delta, float delta) {
  const float z = __fadd_rn(__fmul_rn(v, one_minus_2delta), delta);
  return logf(__fdiv_rn(z, __fsub_rn(1.f, z)));
}

// the same with the division as a reciprocal multiply (one more rounding of the ratio: 1e-7 absolute on the logit)
__device__ __forceinline__ float logit_fast(float v, float one_minus_2delta, float delta) {
  const float z = __fadd_rn(__fmul_rn(v, one_minus_2delta), delta);
  return logf(z * __frcp_rn(__fsub_rn(1.f, z)));
}
// v / d from r = 1 / d with one residual correction: the correctly rounded quotient (a voxel that holds all of its
// layer's energy must give exactly 1: one ulp below moves its logit by 0.06)
__device__ __forceinline__ float div_by(float v, float d, float r) {
  const float q = v * r;
  return fmaf(fmaf(-q, d, v), r, q);
}

__global__ void __launch_bounds__(PRE_THREADS) preprocess_kernel(PreArgs a) {
  pdl_wait();
  __shared__ float lsum[PRE_MAX_LAYERS];
  __shared__ float us[PRE_MAX_LAYERS];
  __shared__ int lb[PRE_MAX_LAYERS + 1];
  __shared__ double red[3][PRE_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = PRE_THREADS / 32;
  for (int i = threadIdx.x; i <= a.n_layers; i += PRE_THREADS) lb[i] = a.bounds[i];
  float mean = 0.f, std = 1.f;
  if (a.mean_std) { mean = a.mean_std[0]; std = a.mean_std[1]; }
  // GlobalStandardizeFromFile keeps x in (logit(eps), -logit(eps)) with eps = delta's default 1e-6 — a zero voxel
  // maps exactly onto the lower edge and is excluded (transforms.py:37,56)
  const float sat = logf(__fdiv_rn(1.0e-6f, __fsub_rn(1.f, 1.0e-6f)));
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  __syncthreads();
  for (int s = blockIdx.x; s < a.n; s += gridDim.x) {
    const float* in = a.showers + (size_t)s * a.voxels;
    float* x = a.x + (size_t)s * a.voxels;
    float* c = a.cond + (size_t)s * (a.n_layers + 1);
    // sweep 1: layer energies, one warp per layer
    for (int l = warp; l < a.n_layers; l += nwarps) {
      float acc = 0.f;
      for (int v = lb[l] + lane; v < lb[l + 1]; v += 32) acc += in[v];
      acc = warp_sum(acc);
      if (lane == 0) lsum[l] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // u_0 = E_tot / E_inc (scaled), u_{l+1} = E_l / (sum_{j >= l} E_j + eps)     transforms.py:388-394, :197-199
      float rem = 0.f;
      for (int l = a.n_layers - 1; l >= 0; --l) {
        rem = __fadd_rn(rem, lsum[l]);
        if (l + 1 < a.n_layers) us[l + 1] = __fdiv_rn(lsum[l], __fadd_rn(rem, a.eps));
      }
      us[0] = __fmul_rn(__fdiv_rn(rem, a.e_inc[s]), a.factor);
      // LogEnergy, ScaleEnergy                                                        transforms.py:162-163, :222-223
      c[a.n_layers] = __fdiv_rn(__fsub_rn(logf(__fadd_rn(a.e_inc[s], a.alpha)), a.e_min), a.e_scale);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_layers; i += PRE_THREADS) {
      float t = logit_rescaled(us[i], a.one_minus_2delta, a.delta);
      if (a.stats && t > sat && t < -sat) { cnt += 1.0; sum += (double)t; sq += (double)t * (double)t; }
      if (a.mean_std) t = __fdiv_rn(__fsub_rn(t, mean), std);
      c[i] = t;
    }
    // sweep 2: voxels normalised to unit layer sum, logit, (standardise)
    for (int l = warp; l < a.n_layers; l += nwarps) {
      const float denom = __fadd_rn(lsum[l], a.eps);
      for (int v = lb[l] + lane; v < lb[l + 1]; v += 32) {
        float t = logit_rescaled(__fdiv_rn(in[v], denom), a.one_minus_2delta, a.delta);
        if (a.stats && t > sat && t < -sat) { cnt += 1.0; sum += (double)t; sq += (double)t * (double)t; }
        if (a.mean_std) t = __fdiv_rn(__fsub_rn(t, mean), std);
        x[v] = t;
      }
    }
    __syncthreads();
  }
  if (a.stats) {
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) { red[0][warp] = cnt; red[1][warp] = sum; red[2][warp] = sq; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int w = 0; w < nwarps; ++w) t += red[threadIdx.x][w];
      atomicAdd(a.stats + threadIdx.x, t);
    }
  }
}

// ---- layers of at most 32 * CAP voxels (ds2: 144, ds3: 900): ONE WARP per (shower, layer) keeps the layer in
// registers, so every voxel is read once (all loads of the layer in flight together) and written once, and nothing
// in the kernel waits for anything else: no CTA barrier, no serial section (the first version ran one CTA per shower
// with a barrier and a one-warp tail per shower).  Small layers (G = 2 at CAP = 8) are taken two ADJACENT ones at a
// time: ten loads per lane in flight over 1.1 KB of contiguous memory.  The layer sums are parked in the shower's row
// of `cond`; a second, tiny kernel (one THREAD per shower) turns them into the u features in place.  The generic
// kernel above re-reads the voxels in a second sweep and chains one load per lane.
__device__ __forceinline__ void stats_reduce(double cnt, double sum, double sq, double* stats, double (*red)[PRE_THREADS / 32]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) { red[0][warp] = cnt; red[1][warp] = sum; red[2][warp] = sq; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < PRE_THREADS / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(stats + threadIdx.x, t);
  }
}

template <int CAP, int G>
__global__ void __launch_bounds__(PRE_THREADS) preprocess_layer_kernel(PreArgs a) {
  pdl_wait();
  __shared__ int lb[PRE_MAX_LAYERS + 1];
  __shared__ double red[3][PRE_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = PRE_THREADS / 32;
  for (int i = threadIdx.x; i <= a.n_layers; i += PRE_THREADS) lb[i] = a.bounds[i];
  float mean = 0.f, std = 1.f;
  if (a.mean_std) { mean = a.mean_std[0]; std = a.mean_std[1]; }
  const float inv_std = __frcp_rn(std);
  // a zero voxel must land exactly on the lower saturation edge: same function as the voxels go through
  const float sat = logit_fast(0.f, 1.f, 1.0e-6f);
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  __syncthreads();
  const int64_t ntasks = (int64_t)a.n * a.n_layers, ngroups = (ntasks + G - 1) / G, stride = (int64_t)gridDim.x * nwarps;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += stride) {
    float v[G][CAP];
    int base[G], len[G];
    int64_t sh[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int64_t task = grp * G + j;
      sh[j] = task / a.n_layers;
      const int l = (int)(task - sh[j] * a.n_layers);
      base[j] = lb[l];
      len[j] = task < ntasks ? lb[l + 1] - base[j] : 0;
      const float* in = a.showers + (size_t)sh[j] * a.voxels + base[j];
#pragma unroll
      for (int k = 0; k < CAP; ++k) v[j][k] = lane + 32 * k < len[j] ? __ldg(in + lane + 32 * k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (grp * G + j >= ntasks) break;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < CAP; ++k) acc += v[j][k];
      acc = warp_sum(acc);
      const int l = (int)(grp * G + j - sh[j] * a.n_layers);
      if (lane == 0) a.cond[(size_t)sh[j] * (a.n_layers + 1) + l] = acc;  // layer energy, for preprocess_cond_kernel
      // the three divisions per voxel of the reference's tensor ops through reciprocals (with IEEE divisions the
      // kernel was ALU-bound at 0.2 of the HBM rate)
      const float denom = __fadd_rn(acc, a.eps), inv_denom = __frcp_rn(denom);
      float* x = a.x + (size_t)sh[j] * a.voxels + base[j];
#pragma unroll
      for (int k = 0; k < CAP; ++k) {
        if (lane + 32 * k < len[j]) {
          float t = logit_fast(div_by(v[j][k], denom, inv_denom), a.one_minus_2delta, a.delta);
          if (a.stats && t > sat && t < -sat) { cnt += 1.0; sum += (double)t; sq += (double)t * (double)t; }
          if (a.mean_std) t = (t - mean) * inv_std;
          x[lane + 32 * k] = t;
        }
      }
    }
  }
  if (a.stats) stats_reduce(cnt, sum, sq, a.stats, red);
}

// one thread per shower: layer energies (parked in cond[s][0 .. L-1]) -> u features, in place from the last layer down;
// u_0 = E_tot / E_inc (scaled), u_{l+1} = E_l / (sum_{j >= l} E_j + eps)          transforms.py:388-394, :197-199
__global__ void __launch_bounds__(PRE_THREADS) preprocess_cond_kernel(PreArgs a) {
  pdl_wait();
  __shared__ double red[3][PRE_THREADS / 32];
  float mean = 0.f, std = 1.f;
  if (a.mean_std) { mean = a.mean_std[0]; std = a.mean_std[1]; }
  const float inv_std = __frcp_rn(std);
  const float sat = logit_fast(0.f, 1.f, 1.0e-6f);
  double cnt = 0.0, sum = 0.0, sq = 0.0;
  const int L = a.n_layers;
  for (int64_t s = (int64_t)blockIdx.x * PRE_THREADS + threadIdx.x; s < a.n; s += (int64_t)gridDim.x * PRE_THREADS) {
    float* c = a.cond + (size_t)s * (L + 1);
    const float e_inc = a.e_inc[s];
    auto emit = [&](int i, float u) {
      float t = logit_fast(u, a.one_minus_2delta, a.delta);
      if (a.stats && t > sat && t < -sat) { cnt += 1.0; sum += (double)t; sq += (double)t * (double)t; }
      if (a.mean_std) t = (t - mean) * inv_std;
      c[i] = t;
    };
    float r = c[L - 1];                    // E_{L-1}: part of every suffix sum, gives no u of its own
    for (int l = L - 2; l >= 0; --l) {
      const float e_l = c[l];              // read before slot l + 1 (already consumed) is overwritten
      r = __fadd_rn(r, e_l);
      emit(l + 1, __fdiv_rn(e_l, __fadd_rn(r, a.eps)));
    }
    emit(0, __fmul_rn(__fdiv_rn(r, e_inc), a.factor));
    // LogEnergy, ScaleEnergy                                                        transforms.py:162-163, :222-223
    c[L] = __fdiv_rn(__fsub_rn(logf(__fadd_rn(e_inc, a.alpha)), a.e_min), a.e_scale);
  }
  if (a.stats) stats_reduce(cnt, sum, sq, a.stats, red);
}

// (count, sum, sum of squares) -> (mean, unbiased std) like Tensor.mean() / Tensor.std() (transforms.py:59-60)
__global__ void preprocess_stats_kernel(const double* stats, float* mean_std) {
  pdl_wait();
  const double n = stats[0], m = stats[1] / n;
  const double var = (stats[2] - n * m * m) / (n - 1.0);
  mean_std[0] = (float)m;
  mean_std[1] = (float)sqrt(var > 0.0 ? var : 0.0);
}

// (x - mean) / std over the voxels and the u features (not the energy column of cond)
__global__ void __launch_bounds__(PRE_THREADS) preprocess_standardize_kernel(float* x, int64_t nx, float* cond, int64_t n,
                                                                             int n_layers, const float* mean_std) {
  pdl_wait();
  const float mean = mean_std[0], std = mean_std[1], inv_std = __frcp_rn(std);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 ? nx / 4 : 0;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 v = x4[i];
    v.x = (v.x - mean) * inv_std; v.y = (v.y - mean) * inv_std; v.z = (v.z - mean) * inv_std; v.w = (v.w - mean) * inv_std;
    x4[i] = v;
  }
  for (int64_t i = 4 * n4 + tid; i < nx; i += stride) x[i] = __fdiv_rn(__fsub_rn(x[i], mean), std);
  const int64_t nc = n * n_layers;
  for (int64_t i = tid; i < nc; i += stride) {
    float* p = cond + (i / n_layers) * (n_layers + 1) + (i % n_layers);
    *p = __fdiv_rn(__fsub_rn(*p, mean), std);
  }
}

}  // namespace

int preprocess_showers(const float* showers, const float* e_inc, int64_t n, int voxels, int n_layers,
                       const int32_t* bounds_dev, int max_layer, float eps, float factor, float delta, float alpha,
                       float e_min, float e_max, float* mean_std_dev, int compute_stats, double* stats_dev, float* x,
                       float* cond, cudaStream_t s) {
  V4H_REQUIRE(n_layers >= 1 && n_layers <= PRE_MAX_LAYERS, "preprocess: 1 <= n_layers <= %d", PRE_MAX_LAYERS);
  PreArgs a;
  a.showers = showers; a.e_inc = e_inc; a.n = (int)n; a.voxels = voxels; a.n_layers = n_layers; a.bounds = bounds_dev;
  a.eps = eps; a.factor = factor; a.delta = delta; a.one_minus_2delta = (float)(1.0 - 2.0 * (double)delta);
  a.alpha = alpha; a.e_min = e_min; a.e_scale = (float)((double)e_max - (double)e_min);
  a.mean_std = compute_stats ? nullptr : mean_std_dev;
  a.stats = compute_stats ? stats_dev : nullptr;
  a.x = x; a.cond = cond;
  if (compute_stats) V4H_CUDA(cudaMemsetAsync(stats_dev, 0, 3 * sizeof(double), s));
  if (max_layer <= 32 * 32) {
    // G = 2 (two adjacent small layers per warp) measured slower at ds2: 0.60 / 0.57 ms against 0.57 / 0.49
    const int G = 1;
    const int64_t groups = (n * n_layers + G - 1) / G, want = (groups + PRE_THREADS / 32 - 1) / (PRE_THREADS / 32);
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    if (max_layer <= 32 * 8) V4H_CUDA(launch_pdl(preprocess_layer_kernel<8, 1>, dim3(grid), dim3(PRE_THREADS), 0, s, a));
    else V4H_CUDA(launch_pdl(preprocess_layer_kernel<32, 1>, dim3(grid), dim3(PRE_THREADS), 0, s, a));
    V4H_LAUNCH_CHECK();
    const unsigned gridc = (unsigned)((n + PRE_THREADS - 1) / PRE_THREADS);
    V4H_CUDA(launch_pdl(preprocess_cond_kernel, dim3(gridc), dim3(PRE_THREADS), 0, s, a));
  } else {
    const int64_t grid = n < 148 * 8 ? n : 148 * 8;
    V4H_CUDA(launch_pdl(preprocess_kernel, dim3((unsigned)grid), dim3(PRE_THREADS), 0, s, a));
  }
  V4H_LAUNCH_CHECK();
  if (compute_stats) {
    V4H_CUDA(launch_pdl(preprocess_stats_kernel, dim3(1), dim3(1), 0, s, (const double*)stats_dev, mean_std_dev));
    V4H_LAUNCH_CHECK();
    const int64_t nx = n * voxels;
    V4H_CUDA(launch_pdl(preprocess_standardize_kernel, dim3(148 * 8), dim3(PRE_THREADS), 0, s, x, nx, cond, n, n_layers,
                        (const float*)mean_std_dev));
    V4H_LAUNCH_CHECK();
  }
  return V4H_OK;
}

}  // namespace v4h
