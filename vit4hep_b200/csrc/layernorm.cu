// LayerNorm (no affine, eps 1e-6, biased variance) fused with adaLN modulation, and its
// backward fused with the gate backward of the next branch in the chain.
// Restates: reference nn/vit.py:309-311 (norm1/norm2), :457-458 (modulate), :331-332 (gated residual).
// One warp per token row; the per-sample reductions (d shift, d scale, d gate) are done per CTA in
// registers/shared memory and published with one atomicAdd per (CTA, column).
#include <algorithm>
#include <cstdlib>
#include <initializer_list>

#include "kernels.cuh"
#include "umma.cuh"

namespace v4h {

namespace {

constexpr float LN_EPS = 1e-6f;
constexpr int MAXV = 16;  // columns per lane -> D <= 512
constexpr int WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) ln_mod_fwd_kernel(
    const float* __restrict__ h, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, T* __restrict__ a, int ld_a, float2* __restrict__ stats, int M, int D, int rows_per_sample) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* hr = h + (size_t)row * D;
  float v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int d = lane + i * 32;
    v[i] = d < D ? hr[d] : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int d = lane + i * 32;
    float c = d < D ? v[i] - mean : 0.f;
    sq += c * c;
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + LN_EPS);
  if (lane == 0 && stats) stats[row] = make_float2(mean, rstd);
  const int b = row / rows_per_sample;
  const float* sh = shift + (size_t)b * mod_stride;
  const float* sc = scale + (size_t)b * mod_stride;
  T* ar = a + (size_t)row * ld_a;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int d = lane + i * 32;
    if (d < D) ar[d] = from_f<T>((v[i] - mean) * rstd * (1.f + sc[d]) + sh[d]);
  }
  if (lane < ld_a - D) ar[D + lane] = from_f<T>(lane == 0 ? 1.f : 0.f);  // "ones" column, see ln_modulate_fwd
}

// grid (chunks, B); each CTA owns rows [chunk*rows_per_cta, ...) of ONE sample.
template <typename T, bool HAS_LN, bool HAS_GATE>
__global__ void __launch_bounds__(WARPS * 32) ln_mod_bwd_kernel(
    const T* __restrict__ da, const float* __restrict__ h, const float2* __restrict__ stats,
    const float* __restrict__ scale, int mod_stride, float* __restrict__ dh, bool dh_accumulate,
    float* __restrict__ dshift, float* __restrict__ dscale, int dmod_stride, const T* __restrict__ y,
    const float* __restrict__ gate, T* __restrict__ dy, float* __restrict__ dgate,
    float* __restrict__ dbias, int D, int rows_per_sample, int rows_per_cta) {
  pdl_wait();
  __shared__ float red[WARPS][MAXV * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows_per_sample, r_begin + rows_per_cta);

  float sc1[MAXV], gt[MAXV];
  float acc_shift[MAXV], acc_scale[MAXV], acc_gate[MAXV], acc_bias[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int d = lane + i * 32;
    sc1[i] = (HAS_LN && d < D) ? 1.f + scale[(size_t)b * mod_stride + d] : 0.f;
    gt[i] = (HAS_GATE && d < D) ? gate[(size_t)b * mod_stride + d] : 0.f;
    acc_shift[i] = acc_scale[i] = acc_gate[i] = acc_bias[i] = 0.f;
  }

  for (int r = r_begin + warp; r < r_end; r += WARPS) {
    const size_t row = (size_t)b * rows_per_sample + r;
    float dx[MAXV];
    if (HAS_LN) {
      const float2 st = stats[row];
      float g[MAXV], xh[MAXV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        int d = lane + i * 32;
        if (d < D) {
          float dav = to_f(da[row * D + d]);
          xh[i] = (h[row * D + d] - st.x) * st.y;
          g[i] = dav * sc1[i];
          acc_shift[i] += dav;
          acc_scale[i] += dav * xh[i];
          s1 += g[i];
          s2 += g[i] * xh[i];
        } else {
          xh[i] = 0.f; g[i] = 0.f;
        }
      }
      s1 = warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) dx[i] = st.y * (g[i] - s1 - xh[i] * s2);
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int d = lane + i * 32;
      if (d < D) {
        float dhn;
        if (HAS_LN) {
          dhn = dx[i] + (dh_accumulate ? dh[row * D + d] : 0.f);
          dh[row * D + d] = dhn;
        } else {
          dhn = dh[row * D + d];
        }
        if (HAS_GATE) {
          float dyv = gt[i] * dhn;
          dy[row * D + d] = from_f<T>(dyv);
          acc_gate[i] += dhn * to_f(y[row * D + d]);
          acc_bias[i] += dyv;
        }
      }
    }
  }

  // cross-warp reduction, one quantity at a time
  auto publish = [&](float (&acc)[MAXV], float* dst) {
    if (dst == nullptr) return;  // uniform across the CTA
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXV; ++i) red[warp][lane + i * 32] = acc[i];
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += red[w][d];
      atomicAdd(dst + d, t);
    }
  };
  if (HAS_LN) {
    publish(acc_shift, dshift ? dshift + (size_t)b * dmod_stride : nullptr);
    publish(acc_scale, dscale ? dscale + (size_t)b * dmod_stride : nullptr);
  }
  if (HAS_GATE) {
    publish(acc_gate, dgate ? dgate + (size_t)b * dmod_stride : nullptr);
    publish(acc_bias, dbias);
  }
}

// ---- vectorised variant (D % 4 == 0): thread = 4 consecutive columns, CTA = one slab of rows of ONE
// sample, RB rows in flight per iteration; the two LayerNorm row sums cross the warps through a
// double-buffered shared-memory exchange (one __syncthreads per RB rows).  Per-thread state is
// 4 columns x 4 accumulators, so many CTAs fit per SM and the loads of RB rows overlap.
constexpr int RB = 2;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <typename T, bool HAS_LN, bool HAS_GATE>
__global__ void __launch_bounds__(128, 6) ln_mod_bwd_vec_kernel(
    const T* __restrict__ da, const float* __restrict__ h, const float2* __restrict__ stats,
    const float* __restrict__ scale, int mod_stride, float* __restrict__ dh, bool dh_accumulate,
    float* __restrict__ dshift, float* __restrict__ dscale, int dmod_stride, const T* __restrict__ y,
    const float* __restrict__ gate, T* __restrict__ dy, float* __restrict__ dgate,
    float* __restrict__ dbias, int D, int rows_per_sample, int rows_per_cta, int DBG_SKIP) {
  pdl_wait();
  __shared__ float2 part[2][RB][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows_per_sample, r_begin + rows_per_cta);
  const int col = threadIdx.x * 4;
  const bool act = col < D;
  const float inv_d = 1.f / (float)D;

  float4 sc1 = make_float4(0.f, 0.f, 0.f, 0.f), gt = sc1;
  if (act) {
    if (HAS_LN) {
      sc1 = ld4(scale + (size_t)b * mod_stride + col);
      sc1.x += 1.f; sc1.y += 1.f; sc1.z += 1.f; sc1.w += 1.f;
    }
    if (HAS_GATE) gt = ld4(gate + (size_t)b * mod_stride + col);
  }
  float4 a_shift = make_float4(0.f, 0.f, 0.f, 0.f), a_scale = a_shift, a_gate = a_shift, a_bias = a_shift;

  int buf = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += RB, buf ^= 1) {
    float4 g[RB], xh[RB], dold[RB], yv[RB];
    float rstd[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const bool ok = act && r0 + k < r_end;
      const size_t off = ((size_t)b * rows_per_sample + min(r0 + k, r_end - 1)) * D + col;
      g[k] = xh[k] = dold[k] = yv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      rstd[k] = 0.f;
      if (ok) {
        if (HAS_LN) {
          g[k] = ld4(da + off);  // holds da until the accumulators have seen it
          xh[k] = ld4(h + off);
          if (dh_accumulate) dold[k] = ld4(dh + off);
        } else {
          dold[k] = ld4(dh + off);
        }
        if (HAS_GATE) yv[k] = ld4(y + off);
      }
    }
    if (HAS_LN) {
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        const float2 st = stats[(size_t)b * rows_per_sample + min(r0 + k, r_end - 1)];
        rstd[k] = st.y;
        const bool ok = act && r0 + k < r_end;
        float4 x = xh[k], d4 = g[k];
        x.x = (x.x - st.x) * st.y; x.y = (x.y - st.x) * st.y; x.z = (x.z - st.x) * st.y; x.w = (x.w - st.x) * st.y;
        if (!ok) x = make_float4(0.f, 0.f, 0.f, 0.f);
        a_shift.x += d4.x; a_shift.y += d4.y; a_shift.z += d4.z; a_shift.w += d4.w;
        a_scale.x += d4.x * x.x; a_scale.y += d4.y * x.y; a_scale.z += d4.z * x.z; a_scale.w += d4.w * x.w;
        d4.x *= sc1.x; d4.y *= sc1.y; d4.z *= sc1.z; d4.w *= sc1.w;
        xh[k] = x; g[k] = d4;
        float s1 = d4.x + d4.y + d4.z + d4.w;
        float s2 = d4.x * x.x + d4.y * x.y + d4.z * x.z + d4.w * x.w;
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) part[buf][k][warp] = make_float2(s1, s2);
      }
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const bool ok = act && r0 + k < r_end;
      float4 dn = dold[k];
      if (HAS_LN) {
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < nwarps; ++w) { const float2 p2 = part[buf][k][w]; s1 += p2.x; s2 += p2.y; }
        s1 *= inv_d; s2 *= inv_d;
        dn.x += rstd[k] * (g[k].x - s1 - xh[k].x * s2);
        dn.y += rstd[k] * (g[k].y - s1 - xh[k].y * s2);
        dn.z += rstd[k] * (g[k].z - s1 - xh[k].z * s2);
        dn.w += rstd[k] * (g[k].w - s1 - xh[k].w * s2);
      }
      if (ok) {
        const size_t off = ((size_t)b * rows_per_sample + r0 + k) * D + col;
        if (HAS_LN) st4(dh + off, dn);
        if (HAS_GATE) {
          const float4 dyv = make_float4(gt.x * dn.x, gt.y * dn.y, gt.z * dn.z, gt.w * dn.w);
          st4(dy + off, dyv);
          a_gate.x += dn.x * yv[k].x; a_gate.y += dn.y * yv[k].y; a_gate.z += dn.z * yv[k].z; a_gate.w += dn.w * yv[k].w;
          a_bias.x += dyv.x; a_bias.y += dyv.y; a_bias.z += dyv.z; a_bias.w += dyv.w;
        }
      }
    }
  }
  if (act) {
    if (HAS_LN) {
      if (dshift && !(DBG_SKIP & 2)) red4(dshift + (size_t)b * dmod_stride + col, a_shift);
      if (dscale && !(DBG_SKIP & 2)) red4(dscale + (size_t)b * dmod_stride + col, a_scale);
    }
    if (HAS_GATE) {
      if (dgate && !(DBG_SKIP & 2)) red4(dgate + (size_t)b * dmod_stride + col, a_gate);
      if (dbias && !(DBG_SKIP & 1)) red4(dbias + col, a_bias);
    }
  }
}


// ---- streaming backward: the rows of a slab arrive in shared memory through 1-d bulk copies
// (cp.async.bulk, one per operand and stage of SR rows) on a full / empty mbarrier
// ring, SS - 1 stages ahead of the arithmetic, and every consumer WARP owns one row of a stage: the two
// LayerNorm row sums are plain warp reductions (no CTA barrier in the loop), a lane carries 16 independent
// columns (ILP instead of occupancy), and the per-sample column sums stay in registers until the slab ends.
// The vector kernel above keeps its loads in registers and pays a memory round trip plus a CTA barrier per
// pair of rows (about 40 % of the HBM rate, 2.4x the instructions per row).
constexpr int SW = 8;   // consumer warps = rows per stage
constexpr int SR = SW;
constexpr int SS = 3;   // stages
constexpr int SV = 4;   // float4 per lane and row -> D <= 512

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sm100::smem_u32(smem_dst)),
               "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SW * 32) : "memory"); }

template <typename T, bool HAS_LN, bool HAS_GATE>
__global__ void __launch_bounds__(SW * 32, 1) ln_mod_bwd_stream_kernel(
    const T* __restrict__ da, const float* __restrict__ h, const float2* __restrict__ stats,
    const float* __restrict__ scale, int mod_stride, float* __restrict__ dh, bool dh_accumulate,
    float* __restrict__ dshift, float* __restrict__ dscale, int dmod_stride, const T* __restrict__ y,
    const float* __restrict__ gate, T* __restrict__ dy, float* __restrict__ dgate,
    float* __restrict__ dbias, int D, int rows_per_sample, int rows_per_cta) {
  using namespace sm100;
  extern __shared__ uint8_t ln_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [SS] producer -> consumers
  uint64_t* empty = full + SS;                         // [SS] consumers -> producer (SW arrivals)
  uint8_t* ring = smem + 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows_per_sample, r_begin + rows_per_cta);
  const int nv = D >> 2;  // float4 per row
  const float inv_d = 1.f / (float)D;
  const bool need_dold = !HAS_LN || dh_accumulate;
  // stage layout: [da SR x D (T)] [h SR x D (f32)] [dh SR x D (f32)] [y SR x D (T)], absent operands take no room
  const uint32_t row_t = (uint32_t)D * sizeof(T), row_f = (uint32_t)D * 4u;
  const uint32_t o_da = 0, o_h = o_da + (HAS_LN ? SR * row_t : 0u), o_dh = o_h + (HAS_LN ? SR * row_f : 0u),
                 o_y = o_dh + (need_dold ? SR * row_f : 0u), stage_bytes = o_y + (HAS_GATE ? SR * row_t : 0u);
  const int nchunks = (r_end - r_begin + SR - 1) / SR;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], SW); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();

  // lane 0 of warp 0 doubles as the producer (a ninth warp would put three warps on one scheduler partition and
  // cap the registers at 168): chunk c + SS - 1 is issued at the top of iteration c, into the stage of chunk
  // c - 1, once every warp has pulled its row of that chunk into registers (empty barrier)
  auto issue = [&](int c) {
    const int stage = c % SS;
    if (c >= SS) mbar_wait(&empty[stage], (uint32_t)((c / SS) - 1) & 1u);
    const int r0 = r_begin + c * SR, rows = min(SR, r_end - r0);
    uint8_t* st = ring + (size_t)stage * stage_bytes;
    const size_t off = ((size_t)b * rows_per_sample + r0) * D;
    mbar_expect_tx(&full[stage],
                   (uint32_t)rows * ((HAS_LN ? row_t + row_f : 0u) + (need_dold ? row_f : 0u) + (HAS_GATE ? row_t : 0u)));
    if (HAS_LN) {
      bulk_load(st + o_da, da + off, (uint32_t)rows * row_t, &full[stage]);
      bulk_load(st + o_h, h + off, (uint32_t)rows * row_f, &full[stage]);
    }
    if (need_dold) bulk_load(st + o_dh, dh + off, (uint32_t)rows * row_f, &full[stage]);
    if (HAS_GATE) bulk_load(st + o_y, y + off, (uint32_t)rows * row_t, &full[stage]);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < SS - 1 && c < nchunks; ++c) issue(c);

  // -------------------------------------------------------------------- consumer warps: lane owns the float4
  // columns lane + 32 i
  float4 sc1[SV], gt[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) {
    const int cv = lane + 32 * i;
    sc1[i] = gt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cv < nv) {
      if (HAS_LN) {
        sc1[i] = ld4(scale + (size_t)b * mod_stride + 4 * cv);
        sc1[i].x += 1.f; sc1[i].y += 1.f; sc1[i].z += 1.f; sc1[i].w += 1.f;
      }
      if (HAS_GATE) gt[i] = ld4(gate + (size_t)b * mod_stride + 4 * cv);
    }
  }
  float4 a_shift[SV], a_scale[SV], a_gate[SV], a_bias[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) a_shift[i] = a_scale[i] = a_gate[i] = a_bias[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 st_next = make_float2(0.f, 0.f);  // row statistics, fetched one chunk ahead
  if (HAS_LN && r_begin + warp < r_end) st_next = stats[(size_t)b * rows_per_sample + r_begin + warp];

  for (int c = 0; c < nchunks; ++c) {
    const int stage = c % SS;
    if (threadIdx.x == 0 && c + SS - 1 < nchunks) issue(c + SS - 1);
    __syncwarp();
    const int row = r_begin + c * SR + warp;  // this warp's row of the chunk
    const bool row_ok = row < r_end;
    const float2 st2 = st_next;
    if (HAS_LN && row + SR < r_end) st_next = stats[(size_t)b * rows_per_sample + row + SR];
    mbar_wait(&full[stage], (uint32_t)(c / SS) & 1u);
    const uint8_t* st = ring + (size_t)stage * stage_bytes;
    float4 g[SV], xh[SV], dn[SV], yv[SV];
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      const int cv = lane + 32 * i;
      g[i] = xh[i] = dn[i] = yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && cv < nv) {
        if (HAS_LN) {
          g[i] = ld4(reinterpret_cast<const T*>(st + o_da) + warp * D + 4 * cv);
          xh[i] = ld4(reinterpret_cast<const float*>(st + o_h) + warp * D + 4 * cv);
        }
        if (need_dold) dn[i] = ld4(reinterpret_cast<const float*>(st + o_dh) + warp * D + 4 * cv);
        if (HAS_GATE) yv[i] = ld4(reinterpret_cast<const T*>(st + o_y) + warp * D + 4 * cv);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // this warp holds its row in registers
    if (!row_ok) continue;
    if (HAS_LN) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < SV; ++i) {
        float4 x = xh[i], d4 = g[i];
        x.x = (x.x - st2.x) * st2.y; x.y = (x.y - st2.x) * st2.y; x.z = (x.z - st2.x) * st2.y; x.w = (x.w - st2.x) * st2.y;
        if (lane + 32 * i >= nv) x = make_float4(0.f, 0.f, 0.f, 0.f);
        a_shift[i].x += d4.x; a_shift[i].y += d4.y; a_shift[i].z += d4.z; a_shift[i].w += d4.w;
        a_scale[i].x += d4.x * x.x; a_scale[i].y += d4.y * x.y; a_scale[i].z += d4.z * x.z; a_scale[i].w += d4.w * x.w;
        d4.x *= sc1[i].x; d4.y *= sc1[i].y; d4.z *= sc1[i].z; d4.w *= sc1[i].w;
        xh[i] = x; g[i] = d4;
        s1 += (d4.x + d4.y) + (d4.z + d4.w);
        s2 += (d4.x * x.x + d4.y * x.y) + (d4.z * x.z + d4.w * x.w);
      }
      s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
      const float rstd = st2.y;
#pragma unroll
      for (int i = 0; i < SV; ++i) {
        dn[i].x += rstd * (g[i].x - s1 - xh[i].x * s2);
        dn[i].y += rstd * (g[i].y - s1 - xh[i].y * s2);
        dn[i].z += rstd * (g[i].z - s1 - xh[i].z * s2);
        dn[i].w += rstd * (g[i].w - s1 - xh[i].w * s2);
      }
    }
    const size_t off = ((size_t)b * rows_per_sample + row) * D;
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      const int cv = lane + 32 * i;
      if (cv < nv) {
        if (HAS_LN) st4(dh + off + 4 * cv, dn[i]);
        if (HAS_GATE) {
          const float4 dyv = make_float4(gt[i].x * dn[i].x, gt[i].y * dn[i].y, gt[i].z * dn[i].z, gt[i].w * dn[i].w);
          st4(dy + off + 4 * cv, dyv);
          a_gate[i].x += dn[i].x * yv[i].x; a_gate[i].y += dn[i].y * yv[i].y;
          a_gate[i].z += dn[i].z * yv[i].z; a_gate[i].w += dn[i].w * yv[i].w;
          a_bias[i].x += dyv.x; a_bias[i].y += dyv.y; a_bias[i].z += dyv.z; a_bias[i].w += dyv.w;
        }
      }
    }
  }

  // ---- column sums: warps -> shared memory (the ring is drained: every stage was waited for) -> one
  // red.global per (CTA, column)
  float4* red = reinterpret_cast<float4*>(ring);  // [4 kinds][SW][nv]
  consumer_bar_sync();                            // every warp is past its last stage read
#pragma unroll
  for (int i = 0; i < SV; ++i) {
    const int cv = lane + 32 * i;
    if (cv < nv) {
      if (HAS_LN) {
        red[(0 * SW + warp) * nv + cv] = a_shift[i];
        red[(1 * SW + warp) * nv + cv] = a_scale[i];
      }
      if (HAS_GATE) {
        red[(2 * SW + warp) * nv + cv] = a_gate[i];
        red[(3 * SW + warp) * nv + cv] = a_bias[i];
      }
    }
  }
  consumer_bar_sync();
  for (int idx = threadIdx.x; idx < 4 * nv; idx += SW * 32) {
    const int kind = idx / nv, cv = idx - kind * nv;
    if ((kind < 2 && !HAS_LN) || (kind >= 2 && !HAS_GATE)) continue;
    float* dst = kind == 0 ? dshift : kind == 1 ? dscale : kind == 2 ? dgate : dbias;
    if (!dst) continue;
    float4 acc = red[(kind * SW) * nv + cv];
#pragma unroll
    for (int w = 1; w < SW; ++w) {
      const float4 t = red[(kind * SW + w) * nv + cv];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    red4(dst + (kind == 3 ? (size_t)0 : (size_t)b * dmod_stride) + 4 * cv, acc);
  }
}


// vectorised forward (D % 4 == 0, D <= 512): one warp per row, each lane owns up to four float4 at
// columns 4 * (lane + 32 i); two rows in flight per warp
template <typename T>
__global__ void __launch_bounds__(256) ln_mod_fwd_vec_kernel(
    const float* __restrict__ h, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, T* __restrict__ a, int ld_a, float2* __restrict__ stats, int M, int D, int rows_per_sample) {
  pdl_wait();
  constexpr int RW = 4;  // rows per warp, all loads issued before the first reduction
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nv = D >> 2;  // float4 per row
  const float inv_d = 1.f / (float)D;
  const int row0 = warp * RW;
  if (row0 >= M) return;
  float4 v[RW][4];
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    const int row = min(row0 + rr, M - 1);
    const float4* hr = reinterpret_cast<const float4*>(h + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      v[rr][i] = c < nv ? hr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    const int row = row0 + rr;
    if (row >= M) break;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) sum += v[rr][i].x + v[rr][i].y + v[rr][i].z + v[rr][i].w;
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (lane + 32 * i < nv) {
        const float a0 = v[rr][i].x - mean, a1 = v[rr][i].y - mean, a2 = v[rr][i].z - mean, a3 = v[rr][i].w - mean;
        sq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + LN_EPS);
    if (lane == 0 && stats) stats[row] = make_float2(mean, rstd);
    const int b = row / rows_per_sample;
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
    T* ar = a + (size_t)row * ld_a;
    if (lane < ld_a - D) ar[D + lane] = from_f<T>(lane == 0 ? 1.f : 0.f);  // "ones" column
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 s4 = __ldg(sc + c), t4 = __ldg(sh + c);
        float4 o;
        o.x = (v[rr][i].x - mean) * rstd * (1.f + s4.x) + t4.x;
        o.y = (v[rr][i].y - mean) * rstd * (1.f + s4.y) + t4.y;
        o.z = (v[rr][i].z - mean) * rstd * (1.f + s4.z) + t4.z;
        o.w = (v[rr][i].w - mean) * rstd * (1.f + s4.w) + t4.w;
        st4(ar + 4 * c, o);
      }
    }
  }
}

// streaming forward: same ring as the streaming backward (rows of h bulk-copied SF stages ahead, warp per row);
// (1 + scale) and shift of the current sample stay in registers and are reloaded when the sample changes
constexpr int SF = 4;  // stages
template <typename T>
__global__ void __launch_bounds__(SW * 32, 3) ln_mod_fwd_stream_kernel(
    const float* __restrict__ h, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, T* __restrict__ a, int ld_a, float2* __restrict__ stats, int M, int D, int rows_per_sample,
    int rows_per_cta) {
  using namespace sm100;
  extern __shared__ uint8_t ln_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ln_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [SF]
  uint64_t* empty = full + SF;                         // [SF], SW arrivals
  uint8_t* ring = smem + 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(M, r_begin + rows_per_cta);
  const int nv = D >> 2;
  const float inv_d = 1.f / (float)D;
  const uint32_t row_f = (uint32_t)D * 4u, stage_bytes = SR * row_f;
  const int nchunks = (r_end - r_begin + SR - 1) / SR;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SF; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], SW); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (nchunks <= 0) return;

  auto issue = [&](int c) {
    const int stage = c % SF;
    if (c >= SF) mbar_wait(&empty[stage], (uint32_t)((c / SF) - 1) & 1u);
    const int r0 = r_begin + c * SR, rows = min(SR, r_end - r0);
    mbar_expect_tx(&full[stage], (uint32_t)rows * row_f);
    bulk_load(ring + (size_t)stage * stage_bytes, h + (size_t)r0 * D, (uint32_t)rows * row_f, &full[stage]);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < SF - 1 && c < nchunks; ++c) issue(c);

  float4 sc1[SV], sh[SV];
  int cur_b = -1;
  for (int c = 0; c < nchunks; ++c) {
    const int stage = c % SF;
    if (threadIdx.x == 0 && c + SF - 1 < nchunks) issue(c + SF - 1);
    __syncwarp();
    const int row = r_begin + c * SR + warp;
    const bool row_ok = row < r_end;
    if (row_ok) {
      const int b = row / rows_per_sample;
      if (b != cur_b) {  // warp-uniform
        cur_b = b;
#pragma unroll
        for (int i = 0; i < SV; ++i) {
          const int cv = lane + 32 * i;
          sc1[i] = sh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cv < nv) {
            sc1[i] = ld4(scale + (size_t)b * mod_stride + 4 * cv);
            sc1[i].x += 1.f; sc1[i].y += 1.f; sc1[i].z += 1.f; sc1[i].w += 1.f;
            sh[i] = ld4(shift + (size_t)b * mod_stride + 4 * cv);
          }
        }
      }
    }
    mbar_wait(&full[stage], (uint32_t)(c / SF) & 1u);
    float4 v[SV];
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      const int cv = lane + 32 * i;
      v[i] = (row_ok && cv < nv) ? ld4(reinterpret_cast<const float*>(ring + (size_t)stage * stage_bytes) + warp * D + 4 * cv)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (!row_ok) continue;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < SV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      if (lane + 32 * i < nv) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        sq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + LN_EPS);
    if (lane == 0 && stats) stats[row] = make_float2(mean, rstd);
    T* ar = a + (size_t)row * ld_a;
    if (lane < ld_a - D) ar[D + lane] = from_f<T>(lane == 0 ? 1.f : 0.f);  // "ones" column
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      const int cv = lane + 32 * i;
      if (cv < nv) {
        float4 o;
        o.x = (v[i].x - mean) * rstd * sc1[i].x + sh[i].x;
        o.y = (v[i].y - mean) * rstd * sc1[i].y + sh[i].y;
        o.z = (v[i].z - mean) * rstd * sc1[i].z + sh[i].z;
        o.w = (v[i].w - mean) * rstd * sc1[i].w + sh[i].w;
        st4(ar + 4 * cv, o);
      }
    }
  }
}


// every pointer the vector kernel touches with 16-byte (fp32) / 8-byte (bf16) accesses
inline bool ln_vec_ok(int D, int mod_stride, int dmod_stride, std::initializer_list<const void*> ptrs) {
  if (D % 4 || D > 512 || mod_stride % 4 || dmod_stride % 4) return false;
  for (const void* q : ptrs)
    if (reinterpret_cast<uintptr_t>(q) & 15) return false;
  return true;
}

// streaming backward: launch geometry shared by ln_modulate_bwd and gate_bwd
inline bool ln_stream_enabled() {
  static const int on = [] { const char* e = getenv("V4H_LN_STREAM"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}
inline int sm_count_ln() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 148;
    return v;
  }();
  return n;
}
template <typename T, bool HAS_LN, bool HAS_GATE, typename... Args>
int launch_ln_stream(int B, int D, int rows_per_sample, bool need_dold, cudaStream_t s, Args... args) {
  const size_t stage = (size_t)SR * D * ((HAS_LN ? sizeof(T) + 4 : 0) + (need_dold ? 4 : 0) + (HAS_GATE ? sizeof(T) : 0));
  const size_t smem = 256 + std::max(SS * stage, (size_t)4 * SW * D * 4);
  // one CTA per SM (shared memory), all of them in ONE wave; slabs a multiple of the stage rows
  const int want = sm_count_ln() / B;
  const int slabs = std::max(1, std::min(want, (int)ceil_div(rows_per_sample, SR)));
  const int rows_per_cta = (int)ceil_div(ceil_div(rows_per_sample, slabs), SR) * SR;
  static size_t configured = 0;
  if (smem > configured) {
    V4H_CUDA(cudaFuncSetAttribute(ln_mod_bwd_stream_kernel<T, HAS_LN, HAS_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((unsigned)ceil_div(rows_per_sample, rows_per_cta), (unsigned)B);
  V4H_CUDA(launch_pdl(ln_mod_bwd_stream_kernel<T, HAS_LN, HAS_GATE>, grid, dim3(SW * 32), smem, s, args..., D, rows_per_sample, rows_per_cta));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace

template <typename T>
int ln_modulate_fwd(const float* h, const float* shift, const float* scale, int mod_stride, T* a, int ld_a,
                    float2* stats, int M, int D, int rows_per_sample, cudaStream_t s) {
  if (D > MAXV * 32) return fail(V4H_ERR_UNSUPPORTED, "ln_modulate: hidden_dim %d > %d", D, MAXV * 32);
  if (ld_a < D || ld_a - D > 32) return fail(V4H_ERR_INVALID, "ln_modulate: bad output pitch %d for %d columns", ld_a, D);
  if (ld_a % 4 == 0 && ln_vec_ok(D, mod_stride, 0, {h, shift, scale, a}) && ln_stream_enabled()) {
    const size_t smem = 256 + (size_t)SF * SR * D * 4;
    static size_t configured = 0;
    if (smem > configured) {
      V4H_CUDA(cudaFuncSetAttribute(ln_mod_fwd_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    // three CTAs per SM (shared memory) in one wave, slabs a multiple of the stage rows
    const int nslab = std::max(1, std::min((int)ceil_div(M, SR), 3 * sm_count_ln()));
    const int rows_per_cta = (int)ceil_div(ceil_div(M, nslab), SR) * SR;
    V4H_CUDA(launch_pdl(ln_mod_fwd_stream_kernel<T>, dim3((unsigned)ceil_div(M, rows_per_cta)), dim3(SW * 32), smem, s, h, shift,
                        scale, mod_stride, a, ld_a, stats, M, D, rows_per_sample, rows_per_cta));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  if (ld_a % 4 == 0 && ln_vec_ok(D, mod_stride, 0, {h, shift, scale, a})) {
    V4H_CUDA(launch_pdl(ln_mod_fwd_vec_kernel<T>, dim3((unsigned)ceil_div(M, 32)), dim3(256), 0, s, h, shift, scale, mod_stride, a, ld_a, stats, M, D, rows_per_sample));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  V4H_CUDA(launch_pdl(ln_mod_fwd_kernel<T>, dim3((unsigned)ceil_div(M, WARPS)), dim3(WARPS * 32), 0, s, h, shift, scale, mod_stride, a, ld_a, stats, M, D, rows_per_sample));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

template <typename T>
int ln_modulate_bwd(const T* da, const float* h, const float2* stats, const float* scale, int mod_stride,
                    float* dh, bool dh_accumulate, float* dshift, float* dscale, int dmod_stride,
                    const T* y, const float* gate, T* dy, float* dgate, float* dbias, int M, int D,
                    int rows_per_sample, cudaStream_t s) {
  if (D > MAXV * 32) return fail(V4H_ERR_UNSUPPORTED, "ln_modulate: hidden_dim %d > %d", D, MAXV * 32);
  const int B = M / rows_per_sample;
  if (ln_vec_ok(D, mod_stride, dmod_stride, {da, h, scale, dh, dshift, dscale, y, gate, dy, dgate, dbias})) {
    if (ln_stream_enabled() && (D * sizeof(T)) % 16 == 0) {
      if (gate != nullptr)
        return launch_ln_stream<T, true, true>(B, D, rows_per_sample, dh_accumulate, s, da, h, stats, scale, mod_stride, dh,
                                               dh_accumulate, dshift, dscale, dmod_stride, y, gate, dy, dgate, dbias);
      return launch_ln_stream<T, true, false>(B, D, rows_per_sample, dh_accumulate, s, da, h, stats, scale, mod_stride, dh,
                                              dh_accumulate, dshift, dscale, dmod_stride, (const T*)nullptr,
                                              (const float*)nullptr, (T*)nullptr, (float*)nullptr, (float*)nullptr);
    }
    const int rows_per_cta = 16, threads = (int)ceil_div(D / 4, 32) * 32;
    dim3 vgrid((unsigned)ceil_div(rows_per_sample, rows_per_cta), (unsigned)B);
    static const int dbg_skip = [] { const char* e = getenv("V4H_LN_DBG_SKIP"); return e ? atoi(e) : 0; }();
    if (gate != nullptr)
      V4H_CUDA(launch_pdl(ln_mod_bwd_vec_kernel<T, true, true>, dim3(vgrid), dim3(threads), 0, s,  da, h, stats, scale, mod_stride, dh, dh_accumulate, dshift, dscale, dmod_stride, y, gate, dy, dgate, dbias, D, rows_per_sample, rows_per_cta, dbg_skip));
    else
      V4H_CUDA(launch_pdl(ln_mod_bwd_vec_kernel<T, true, false>, dim3(vgrid), dim3(threads), 0, s,  da, h, stats, scale, mod_stride, dh, dh_accumulate, dshift, dscale, dmod_stride, nullptr, nullptr, nullptr, nullptr, nullptr, D, rows_per_sample, rows_per_cta, dbg_skip));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  const int rows_per_cta = 32;
  dim3 grid((unsigned)ceil_div(rows_per_sample, rows_per_cta), (unsigned)B);
  if (gate != nullptr) {
    V4H_CUDA(launch_pdl(ln_mod_bwd_kernel<T, true, true>, dim3(grid), dim3(WARPS * 32), 0, s,  da, h, stats, scale, mod_stride, dh, dh_accumulate, dshift, dscale, dmod_stride, y, gate, dy, dgate, dbias, D, rows_per_sample, rows_per_cta));
  } else {
    V4H_CUDA(launch_pdl(ln_mod_bwd_kernel<T, true, false>, dim3(grid), dim3(WARPS * 32), 0, s,  da, h, stats, scale, mod_stride, dh, dh_accumulate, dshift, dscale, dmod_stride, nullptr, nullptr, nullptr, nullptr, nullptr, D, rows_per_sample, rows_per_cta));
  }
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

template <typename T>
int gate_bwd(const float* dh, const T* y, const float* gate, int mod_stride, T* dy, float* dgate,
             int dmod_stride, float* dbias, int M, int D, int rows_per_sample, cudaStream_t s) {
  if (D > MAXV * 32) return fail(V4H_ERR_UNSUPPORTED, "gate_bwd: hidden_dim %d > %d", D, MAXV * 32);
  const int B = M / rows_per_sample;
  if (ln_vec_ok(D, mod_stride, dmod_stride, {dh, y, gate, dy, dgate, dbias})) {
    if (ln_stream_enabled() && (D * sizeof(T)) % 16 == 0)
      return launch_ln_stream<T, false, true>(B, D, rows_per_sample, true, s, (const T*)nullptr, (const float*)nullptr,
                                              (const float2*)nullptr, (const float*)nullptr, mod_stride,
                                              const_cast<float*>(dh), false, (float*)nullptr, (float*)nullptr, dmod_stride, y,
                                              gate, dy, dgate, dbias);
    const int rows_per_cta = 16, threads = (int)ceil_div(D / 4, 32) * 32;
    dim3 vgrid((unsigned)ceil_div(rows_per_sample, rows_per_cta), (unsigned)B);
    V4H_CUDA(launch_pdl(ln_mod_bwd_vec_kernel<T, false, true>, dim3(vgrid), dim3(threads), 0, s,  nullptr, nullptr, nullptr, nullptr, mod_stride, const_cast<float*>(dh), false, nullptr, nullptr, dmod_stride, y, gate, dy, dgate, dbias, D, rows_per_sample, rows_per_cta, 0));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  const int rows_per_cta = 32;
  dim3 grid((unsigned)ceil_div(rows_per_sample, rows_per_cta), (unsigned)B);
  V4H_CUDA(launch_pdl(ln_mod_bwd_kernel<T, false, true>, dim3(grid), dim3(WARPS * 32), 0, s,  nullptr, nullptr, nullptr, nullptr, mod_stride, const_cast<float*>(dh), false, nullptr, nullptr, dmod_stride, y, gate, dy, dgate, dbias, D, rows_per_sample, rows_per_cta));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

#define INST(T)                                                                                          \
  template int ln_modulate_fwd<T>(const float*, const float*, const float*, int, T*, int, float2*, int, int, \
                                  int, cudaStream_t);                                                    \
  template int ln_modulate_bwd<T>(const T*, const float*, const float2*, const float*, int, float*,    \
                                  bool, float*, float*, int, const T*, const float*, T*, float*,        \
                                  float*, int, int, int, cudaStream_t);                                  \
  template int gate_bwd<T>(const float*, const T*, const float*, int, T*, float*, int, float*, int, int, \
                           int, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace v4h
