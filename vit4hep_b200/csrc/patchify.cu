// 3D patchify / unpatchify over (layer, angular, radial) voxel grids as a block-diagonal
// permutation.  Replaces einops "b c (l p1) (a p2) (r p3) -> b (l a r) (p1 p2 p3 c)" and its
// inverse (reference experiments/calochallenge/calochallenge_cfm/model.py:40-60) and the segmented
// split / rearrange / cat variants (reference experiments/calogan/model.py:60-87).
//
// Observation that makes it coalesced both ways: for C == 1 a slab of P1 layers of one segment is a
// contiguous range of the input AND maps onto a contiguous range of tokens of identical extent, so the
// whole map is a permutation inside each slab.  A CTA stages a group of whole slabs (a "chunk")
// through shared memory: 128-bit coalesced loads of the source range, table-driven shared-memory
// reads, coalesced stores of the same range of the destination.  HBM traffic = 2 x 4 bytes / voxel.
#include <vector>

#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int PP_THREADS = 256;

// in-chunk permutation through shared memory; chunk_bounds[c] .. chunk_bounds[c+1] is both the source
// and the destination range (per sample)
__global__ void __launch_bounds__(PP_THREADS) patch_permute_smem_kernel(
    const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ table,
    const int32_t* __restrict__ chunk_bounds, int per_sample) {
  pdl_wait();
  extern __shared__ float stage[];
  const int c0 = chunk_bounds[blockIdx.x], c1 = chunk_bounds[blockIdx.x + 1];
  const int n = c1 - c0;
  const size_t base = (size_t)blockIdx.y * per_sample + c0;
  const float* s = src + base;
  float* d = dst + base;
  if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += PP_THREADS)
      reinterpret_cast<float4*>(stage)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += PP_THREADS) stage[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < n; i += PP_THREADS) stage[i] = s[i];
  }
  __syncthreads();
  const int32_t* tb = table + c0;
  if ((reinterpret_cast<uintptr_t>(d) & 15) == 0 && (reinterpret_cast<uintptr_t>(tb) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += PP_THREADS) {
      const int4 t = __ldg(reinterpret_cast<const int4*>(tb) + i);
      float4 v;
      v.x = stage[t.x - c0]; v.y = stage[t.y - c0]; v.z = stage[t.z - c0]; v.w = stage[t.w - c0];
      reinterpret_cast<float4*>(d)[i] = v;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += PP_THREADS) d[i] = stage[tb[i] - c0];
  } else {
    for (int i = threadIdx.x; i < n; i += PP_THREADS) d[i] = stage[tb[i] - c0];
  }
}

// generic fallback (multi-channel geometries, or a slab larger than shared memory): plain gather
__global__ void patch_permute_gather_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                            const int32_t* __restrict__ table, int per_sample) {
  pdl_wait();
  const size_t base = (size_t)blockIdx.y * per_sample;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < per_sample; j += gridDim.x * blockDim.x)
    dst[base + j] = src[base + table[j]];
}

}  // namespace

int patch_permute(const float* src, float* dst, const int32_t* table, const int32_t* chunk_bounds,
                  int num_chunks, int max_chunk, int64_t B, int per_sample, cudaStream_t s) {
  if (B == 0) return V4H_OK;
  V4H_REQUIRE(B <= 65535, "patch_permute: batch %lld exceeds 65535 per call", (long long)B);
  if (num_chunks > 0) {
    V4H_CUDA(launch_pdl(patch_permute_smem_kernel, dim3(dim3((unsigned)num_chunks, (unsigned)B)), dim3(PP_THREADS), (size_t)max_chunk * sizeof(float), s, src, dst, table, chunk_bounds, per_sample));
  } else {
    unsigned gx = (unsigned)ceil_div(per_sample, 256 * 4);
    V4H_CUDA(launch_pdl(patch_permute_gather_kernel, dim3(dim3(gx, (unsigned)B)), dim3(256), 0, s, src, dst, table, per_sample));
  }
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace v4h
