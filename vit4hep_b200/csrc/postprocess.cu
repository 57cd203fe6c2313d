// Post-processing of sampled showers on the GPU: the REVERSE pass of the
// CaloChallenge ds2 / ds3 shape-model transform chain (reference configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28,
// applied back to front by experiments/calochallenge/experiment.py:286-289):
//   Reshape, AddFeaturesToCond, ScaleEnergy, LogEnergy, GlobalStandardizeFromFile, ExclusiveLogitTransform(rescale),
//   CutValues, ScaleTotalEnergy, NormalizeByElayer        (reference experiments/calochallenge/transforms.py)
// The reference runs these as ~10 elementwise passes plus two 45-iteration Python loops on the CPU after
// `.cpu()`; here a one-thread-per-shower kernel derives the layer energies from the conditions (the u-recursion of
// NormalizeByElayer) and one warp per (shower, layer) does the rest with the layer in registers: un-standardise,
// un-logit, cut, sum, normalise to unit layer sum, scale to the layer energy (layers longer than 1024 voxels: a
// generic one-CTA-per-shower kernel with two sweeps).  HBM-bound: 4 B read + 4 B written per voxel.
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int PP_THREADS = 256;
constexpr int PP_MAX_LAYERS = 128;

struct PostArgs {
  const float* x;     // (n, voxels) sampled showers in the network's normalised space
  const float* cond;  // (n, n_layers + 1): the n_layers u features, then the scaled log incident energy
  int n, voxels, n_layers;
  const int* bounds;  // (n_layers + 1) voxel offsets of the layers (device)
  float mean, std;    // GlobalStandardizeFromFile
  float delta, one_minus_2delta;  // ExclusiveLogitTransform(rescale=True)
  float cut;          // CutValues (voxels only)
  float factor;       // ScaleTotalEnergy (u_0 only)
  float e_scale, e_min, alpha;  // ScaleEnergy (e_max - e_min, e_min), LogEnergy
  float eps, norm_cut;          // NormalizeByElayer
  float* out;         // (n, voxels) energies per voxel
  float* e_out;       // (n) incident energies
};

// GlobalStandardize rev -> ExclusiveLogit rev: sigmoid(v * std + mean), rescaled from [delta, 1 - delta] to [0, 1].
// Separate multiply / add and a true division, like the reference's tensor ops (no fused multiply-add).
__device__ __forceinline__ float unlogit(float v, const PostArgs& a) {
  const float t = __fadd_rn(__fmul_rn(v, a.std), a.mean);
  const float z = 1.f / (1.f + expf(-t));
  return __fdiv_rn(__fsub_rn(z, a.delta), a.one_minus_2delta);
}
__device__ __forceinline__ float voxel_value(float v, const PostArgs& a) {
  const float z = unlogit(v, a);
  return (a.cut != 0.f && z <= a.cut) ? 0.f : z;  // CutValues rev (skipped entirely when cut == 0)
}

// the same with the two divisions per voxel as reciprocal multiplies (one more rounding each, 1e-7 relative; the IEEE
// divisions made the kernel ALU-bound at 0.2 of the HBM rate)
__device__ __forceinline__ float voxel_value_fast(float v, const PostArgs& a, float inv_scale) {
  const float t = fmaf(v, a.std, a.mean);
  const float z = (__frcp_rn(1.f + expf(-t)) - a.delta) * inv_scale;
  return (a.cut != 0.f && z <= a.cut) ? 0.f : z;
}

__global__ void __launch_bounds__(PP_THREADS) postprocess_kernel(PostArgs a) {
  pdl_wait();
  __shared__ float lsum[PP_MAX_LAYERS];
  __shared__ float layer_e[PP_MAX_LAYERS];
  __shared__ float us[PP_MAX_LAYERS];
  __shared__ int lb[PP_MAX_LAYERS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = PP_THREADS / 32;
  for (int i = threadIdx.x; i <= a.n_layers; i += PP_THREADS) lb[i] = a.bounds[i];
  __syncthreads();
  for (int s = blockIdx.x; s < a.n; s += gridDim.x) {
    const float* x = a.x + (size_t)s * a.voxels;
    const float* c = a.cond + (size_t)s * (a.n_layers + 1);
    // u features: same standardisation / logit chain (they carry `u_transform`), then ScaleTotalEnergy rev on
    // u_0 and the clip of NormalizeByElayer rev on u_{i>0}
    for (int i = threadIdx.x; i < a.n_layers; i += PP_THREADS) {
      float u = unlogit(c[i], a);
      if (i == 0) u = __fdiv_rn(u, a.factor);
      else u = fminf(fmaxf(u, 0.f), 1.f);
      us[i] = u;
    }
    // sweep 1: layer sums, one warp per layer
    for (int l = warp; l < a.n_layers; l += nwarps) {
      float acc = 0.f;
      for (int v = lb[l] + lane; v < lb[l + 1]; v += 32) acc += voxel_value(x[v], a);
      acc = warp_sum(acc);
      if (lane == 0) lsum[l] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // ScaleEnergy rev, LogEnergy rev; then the layer energies from the u's (reference transforms.py:363-371)
      const float e_inc = __fsub_rn(expf(__fadd_rn(__fmul_rn(c[a.n_layers], a.e_scale), a.e_min)), a.alpha);
      a.e_out[s] = e_inc;
      const float total = __fmul_rn(e_inc, us[0]);
      float cum = 0.f;
      for (int i = 0; i + 1 < a.n_layers; ++i) {
        const float le = __fmul_rn(__fsub_rn(total, cum), us[i + 1]);
        layer_e[i] = le;
        cum = __fadd_rn(cum, le);
      }
      layer_e[a.n_layers - 1] = __fsub_rn(total, cum);
    }
    __syncthreads();
    // sweep 2: normalise each layer to unit sum, apply the normalised cut, scale to the layer energy
    float* out = a.out + (size_t)s * a.voxels;
    for (int l = warp; l < a.n_layers; l += nwarps) {
      const float denom = __fadd_rn(lsum[l], a.eps), le = layer_e[l];
      for (int v = lb[l] + lane; v < lb[l + 1]; v += 32) {
        float z = __fdiv_rn(voxel_value(x[v], a), denom);
        if (z <= a.norm_cut) z = 0.f;
        out[v] = __fmul_rn(z, le);
      }
    }
    __syncthreads();
  }
}

// ---- layers of at most 32 * CAP voxels: the layer energies only depend on the conditions, so a first, tiny kernel
// (one THREAD per shower) works them out and parks each in the output slot of its layer's first voxel; then ONE WARP
// per (shower, layer) — two adjacent layers at a time when they are small — pulls the layer into registers (all loads
// in flight together), sums, normalises, scales and stores: every voxel is read once and written once, no CTA barrier,
// no serial section in the sweep.
__global__ void __launch_bounds__(PP_THREADS) postprocess_energy_kernel(PostArgs a) {
  pdl_wait();
  __shared__ int lb[PP_MAX_LAYERS + 1];
  for (int i = threadIdx.x; i <= a.n_layers; i += PP_THREADS) lb[i] = a.bounds[i];
  __syncthreads();
  const int L = a.n_layers;
  for (int64_t s = (int64_t)blockIdx.x * PP_THREADS + threadIdx.x; s < a.n; s += (int64_t)gridDim.x * PP_THREADS) {
    const float* c = a.cond + (size_t)s * (L + 1);
    float* out = a.out + (size_t)s * a.voxels;
    // ScaleEnergy rev, LogEnergy rev; then the layer energies from the u's (reference transforms.py:363-371)
    const float e_inc = __fsub_rn(expf(__fadd_rn(__fmul_rn(c[L], a.e_scale), a.e_min)), a.alpha);
    a.e_out[s] = e_inc;
    const float total = __fmul_rn(e_inc, __fdiv_rn(unlogit(c[0], a), a.factor));
    float cum = 0.f;
    for (int i = 0; i + 1 < L; ++i) {
      const float u = fminf(fmaxf(unlogit(c[i + 1], a), 0.f), 1.f);
      const float le = __fmul_rn(__fsub_rn(total, cum), u);
      out[lb[i]] = le;
      cum = __fadd_rn(cum, le);
    }
    out[lb[L - 1]] = __fsub_rn(total, cum);
  }
}

template <int CAP, int G>
__global__ void __launch_bounds__(PP_THREADS) postprocess_layer_kernel(PostArgs a) {
  pdl_wait();
  __shared__ int lb[PP_MAX_LAYERS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = PP_THREADS / 32;
  for (int i = threadIdx.x; i <= a.n_layers; i += PP_THREADS) lb[i] = a.bounds[i];
  const float inv_scale = __frcp_rn(a.one_minus_2delta);
  __syncthreads();
  const int64_t ntasks = (int64_t)a.n * a.n_layers, ngroups = (ntasks + G - 1) / G, stride = (int64_t)gridDim.x * nwarps;
  for (int64_t grp = (int64_t)blockIdx.x * nwarps + warp; grp < ngroups; grp += stride) {
    float v[G][CAP], le[G];
    int len[G];
    float* out[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int64_t task = grp * G + j, s = task / a.n_layers;
      const int l = (int)(task - s * a.n_layers);
      const int base = lb[l];
      len[j] = task < ntasks ? lb[l + 1] - base : 0;
      const float* x = a.x + (size_t)s * a.voxels + base;
      out[j] = a.out + (size_t)s * a.voxels + base;
#pragma unroll
      for (int k = 0; k < CAP; ++k) v[j][k] = lane + 32 * k < len[j] ? __ldg(x + lane + 32 * k) : 0.f;
      // the layer energy parked by postprocess_energy_kernel
      le[j] = __shfl_sync(0xffffffffu, (lane == 0 && len[j] > 0) ? out[j][0] : 0.f, 0);
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (len[j] == 0) break;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < CAP; ++k) {
        v[j][k] = lane + 32 * k < len[j] ? voxel_value_fast(v[j][k], a, inv_scale) : 0.f;
        acc += v[j][k];
      }
      acc = warp_sum(acc);
      const float inv_denom = __frcp_rn(__fadd_rn(acc, a.eps));
#pragma unroll
      for (int k = 0; k < CAP; ++k) {
        if (lane + 32 * k < len[j]) {
          float z = v[j][k] * inv_denom;
          if (z <= a.norm_cut) z = 0.f;
          out[j][lane + 32 * k] = __fmul_rn(z, le[j]);
        }
      }
    }
  }
}

}  // namespace

int postprocess_showers(const float* x, const float* cond, int64_t n, int voxels, int n_layers, const int32_t* bounds_dev,
                        int max_layer, float mean, float std, float delta, float cut, float factor, float e_min, float e_max, float alpha,
                        float eps, float norm_cut, float* out, float* e_out, cudaStream_t s) {
  V4H_REQUIRE(n_layers >= 1 && n_layers <= PP_MAX_LAYERS, "postprocess: 1 <= n_layers <= %d", PP_MAX_LAYERS);
  PostArgs a;
  a.x = x; a.cond = cond; a.n = (int)n; a.voxels = voxels; a.n_layers = n_layers; a.bounds = bounds_dev;
  a.mean = mean; a.std = std; a.delta = delta; a.one_minus_2delta = (float)(1.0 - 2.0 * (double)delta);
  a.cut = cut; a.factor = factor; a.e_scale = (float)((double)e_max - (double)e_min); a.e_min = e_min; a.alpha = alpha;
  a.eps = eps; a.norm_cut = norm_cut; a.out = out; a.e_out = e_out;
  const int64_t grid = n < 148 * 8 ? n : 148 * 8;
  if (max_layer <= 32 * 32) {
    V4H_CUDA(launch_pdl(postprocess_energy_kernel, dim3((unsigned)((n + PP_THREADS - 1) / PP_THREADS)), dim3(PP_THREADS), 0, s, a));
    V4H_LAUNCH_CHECK();
    // G = 2 (two adjacent small layers per warp) measured slower at ds2: 0.60 / 0.57 ms against 0.57 / 0.49
    const int G = 1;
    const int64_t groups = (n * n_layers + G - 1) / G, want = (groups + PP_THREADS / 32 - 1) / (PP_THREADS / 32);
    const unsigned grid2 = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    if (max_layer <= 32 * 8) V4H_CUDA(launch_pdl(postprocess_layer_kernel<8, 1>, dim3(grid2), dim3(PP_THREADS), 0, s, a));
    else V4H_CUDA(launch_pdl(postprocess_layer_kernel<32, 1>, dim3(grid2), dim3(PP_THREADS), 0, s, a));
  } else {
    V4H_CUDA(launch_pdl(postprocess_kernel, dim3((unsigned)grid), dim3(PP_THREADS), 0, s, a));
  }
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace v4h
