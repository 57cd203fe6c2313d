// fp32 SIMT attention forward / backward: softmax(q k^T / sqrt(dh)) v per (sample, head), no mask,
// no dropout (reference nn/vit.py:425-451 with the xformers / SDPA call it makes).
// This is the arithmetic of the fp32 precision mode; the bf16 mode uses the tcgen05 kernels in
// attention_umma.cu.  Layout: qkv (B, T, 3, H, dh) exactly as the qkv Linear writes it
// (reference nn/vit.py:427 reshape), o (B, T, H, dh), lse (B, H, T) = log sum exp of scaled scores.
//
// 4 threads share one query (or key) row, each owning DPT consecutive head dims; 32 rows per CTA;
// keys/values (or queries/dO) stream through shared memory in tiles of 32 rows.
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int ROWS = 32;    // rows per CTA
constexpr int TILE = 32;    // streamed rows per shared-memory tile
constexpr int ATHREADS = ROWS * 4;

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// cooperative load of `rows` rows x dh columns (row stride `ld` elements) into a padded smem tile
template <typename T, int DPT>
__device__ __forceinline__ void load_tile(float (*dst)[4 * (DPT + 1)], const T* __restrict__ src, size_t ld,
                                          int row0, int nrows_total, int dh) {
  for (int idx = threadIdx.x; idx < TILE * 4 * DPT; idx += ATHREADS) {
    const int j = idx / (4 * DPT), d = idx % (4 * DPT);
    float v = 0.f;
    if (row0 + j < nrows_total && d < dh) v = to_f(src[(size_t)(row0 + j) * ld + d]);
    dst[j][(d / DPT) * (DPT + 1) + (d % DPT)] = v;
  }
}

template <typename T, int DPT>
__global__ void __launch_bounds__(ATHREADS) attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o,
                                                            float* __restrict__ lse, int Tn, int H, int dh,
                                                            float scale) {
  pdl_wait();
  __shared__ float Ks[TILE][4 * (DPT + 1)];
  __shared__ float Vs[TILE][4 * (DPT + 1)];
  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int r = blockIdx.x * ROWS + threadIdx.x / 4, quarter = threadIdx.x & 3;
  const size_t ld = (size_t)3 * H * dh;
  const T* qbase = qkv + (size_t)b * Tn * ld + (size_t)hd * dh;
  const T* kbase = qbase + (size_t)H * dh;
  const T* vbase = qbase + (size_t)2 * H * dh;
  const bool row_ok = r < Tn;

  float q[DPT], acc[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = quarter * DPT + i;
    q[i] = (row_ok && d < dh) ? to_f(qbase[(size_t)r * ld + d]) * scale : 0.f;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < Tn; j0 += TILE) {
    __syncthreads();
    load_tile<T, DPT>(Ks, kbase, ld, j0, Tn, dh);
    load_tile<T, DPT>(Vs, vbase, ld, j0, Tn, dh);
    __syncthreads();
    const int nj = min(TILE, Tn - j0);
    float sc[TILE];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < TILE; ++j) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < DPT; ++i) p = fmaf(q[i], Ks[j][quarter * (DPT + 1) + i], p);
      p = quad_sum(p);
      sc[j] = j < nj ? p : -INFINITY;
      tmax = fmaxf(tmax, sc[j]);
    }
    const float mnew = fmaxf(m, tmax);
    const float corr = __expf(m - mnew);
    l *= corr;
#pragma unroll
    for (int i = 0; i < DPT; ++i) acc[i] *= corr;
#pragma unroll
    for (int j = 0; j < TILE; ++j) {
      const float p = __expf(sc[j] - mnew);  // exp(-inf) = 0 for padded keys
      l += p;
#pragma unroll
      for (int i = 0; i < DPT; ++i) acc[i] = fmaf(p, Vs[j][quarter * (DPT + 1) + i], acc[i]);
    }
    m = mnew;
  }
  if (row_ok) {
    const float inv = 1.f / l;
    T* orow = o + ((size_t)b * Tn + r) * H * dh + (size_t)hd * dh;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = quarter * DPT + i;
      if (d < dh) orow[d] = from_f<T>(acc[i] * inv);
    }
    if (quarter == 0) lse[(size_t)bh * Tn + r] = m + __logf(l);
  }
}

// dQ: one CTA per (query tile, sample-head); streams K and V.
//   delta_i = dO_i . O_i ; P_ij = exp(s_ij - lse_i) ; dS_ij = P_ij (dO_i . V_j - delta_i) ;
//   dQ_i = scale * sum_j dS_ij K_j
template <typename T, int DPT>
__global__ void __launch_bounds__(ATHREADS) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                               const float* __restrict__ lse,
                                                               const T* __restrict__ d_o, T* __restrict__ dqkv,
                                                               int Tn, int H, int dh, float scale) {
  pdl_wait();
  __shared__ float Ks[TILE][4 * (DPT + 1)];
  __shared__ float Vs[TILE][4 * (DPT + 1)];
  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int r = blockIdx.x * ROWS + threadIdx.x / 4, quarter = threadIdx.x & 3;
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const T* qbase = qkv + (size_t)b * Tn * ld + (size_t)hd * dh;
  const T* kbase = qbase + (size_t)H * dh;
  const T* vbase = qbase + (size_t)2 * H * dh;
  const bool row_ok = r < Tn;

  float q[DPT], dov[DPT], dq[DPT];
  float delta = 0.f;
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = quarter * DPT + i;
    const bool ok = row_ok && d < dh;
    q[i] = ok ? to_f(qbase[(size_t)r * ld + d]) * scale : 0.f;
    dov[i] = ok ? to_f(d_o[((size_t)b * Tn + r) * ldo + (size_t)hd * dh + d]) : 0.f;
    const float ov = ok ? to_f(o[((size_t)b * Tn + r) * ldo + (size_t)hd * dh + d]) : 0.f;
    delta = fmaf(dov[i], ov, delta);
    dq[i] = 0.f;
  }
  delta = quad_sum(delta);
  const float lse_r = row_ok ? lse[(size_t)bh * Tn + r] : 0.f;

  for (int j0 = 0; j0 < Tn; j0 += TILE) {
    __syncthreads();
    load_tile<T, DPT>(Ks, kbase, ld, j0, Tn, dh);
    load_tile<T, DPT>(Vs, vbase, ld, j0, Tn, dh);
    __syncthreads();
    const int nj = min(TILE, Tn - j0);
#pragma unroll 4
    for (int j = 0; j < TILE; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        s = fmaf(q[i], Ks[j][quarter * (DPT + 1) + i], s);
        dp = fmaf(dov[i], Vs[j][quarter * (DPT + 1) + i], dp);
      }
      s = quad_sum(s);
      dp = quad_sum(dp);
      const float p = j < nj ? __expf(s - lse_r) : 0.f;
      const float ds = p * (dp - delta);
#pragma unroll
      for (int i = 0; i < DPT; ++i) dq[i] = fmaf(ds, Ks[j][quarter * (DPT + 1) + i], dq[i]);
    }
  }
  if (row_ok) {
    T* out = dqkv + ((size_t)b * Tn + r) * ld + (size_t)hd * dh;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = quarter * DPT + i;
      if (d < dh) out[d] = from_f<T>(dq[i] * scale);
    }
  }
}

// dK, dV: one CTA per (key tile, sample-head); streams Q and dO (and per-row lse, delta).
//   dV_j = sum_i P_ij dO_i ;  dK_j = scale * sum_i dS_ij Q_i
template <typename T, int DPT>
__global__ void __launch_bounds__(ATHREADS) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                                const float* __restrict__ lse,
                                                                const T* __restrict__ d_o, T* __restrict__ dqkv,
                                                                int Tn, int H, int dh, float scale) {
  pdl_wait();
  __shared__ float Qs[TILE][4 * (DPT + 1)];
  __shared__ float Ds[TILE][4 * (DPT + 1)];
  __shared__ float lse_s[TILE], delta_s[TILE];
  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int r = blockIdx.x * ROWS + threadIdx.x / 4, quarter = threadIdx.x & 3;
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const T* qbase = qkv + (size_t)b * Tn * ld + (size_t)hd * dh;
  const T* kbase = qbase + (size_t)H * dh;
  const T* vbase = qbase + (size_t)2 * H * dh;
  const T* dobase = d_o + (size_t)b * Tn * ldo + (size_t)hd * dh;
  const T* obase = o + (size_t)b * Tn * ldo + (size_t)hd * dh;
  const bool row_ok = r < Tn;

  float k[DPT], v[DPT], dk[DPT], dv[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = quarter * DPT + i;
    const bool ok = row_ok && d < dh;
    k[i] = ok ? to_f(kbase[(size_t)r * ld + d]) * scale : 0.f;
    v[i] = ok ? to_f(vbase[(size_t)r * ld + d]) : 0.f;
    dk[i] = dv[i] = 0.f;
  }

  for (int i0 = 0; i0 < Tn; i0 += TILE) {
    __syncthreads();
    load_tile<T, DPT>(Qs, qbase, ld, i0, Tn, dh);
    load_tile<T, DPT>(Ds, dobase, ldo, i0, Tn, dh);
    // delta and lse of the streamed query rows: 4 threads per row as everywhere else
    {
      const int i = threadIdx.x / 4;
      float part = 0.f;
      if (i0 + i < Tn)
        for (int d = quarter; d < dh; d += 4)
          part = fmaf(to_f(dobase[(size_t)(i0 + i) * ldo + d]), to_f(obase[(size_t)(i0 + i) * ldo + d]), part);
      part = quad_sum(part);
      if (quarter == 0) {
        delta_s[i] = part;
        lse_s[i] = i0 + i < Tn ? lse[(size_t)bh * Tn + i0 + i] : 0.f;
      }
    }
    __syncthreads();
    const int ni = min(TILE, Tn - i0);
#pragma unroll 4
    for (int i = 0; i < TILE; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int e = 0; e < DPT; ++e) {
        s = fmaf(k[e], Qs[i][quarter * (DPT + 1) + e], s);
        dp = fmaf(v[e], Ds[i][quarter * (DPT + 1) + e], dp);
      }
      s = quad_sum(s);
      dp = quad_sum(dp);
      const float p = i < ni ? __expf(s - lse_s[i]) : 0.f;
      const float ds = p * (dp - delta_s[i]);
#pragma unroll
      for (int e = 0; e < DPT; ++e) {
        dv[e] = fmaf(p, Ds[i][quarter * (DPT + 1) + e], dv[e]);
        dk[e] = fmaf(ds, Qs[i][quarter * (DPT + 1) + e], dk[e]);
      }
    }
  }
  if (row_ok) {
    T* dkout = dqkv + ((size_t)b * Tn + r) * ld + (size_t)H * dh + (size_t)hd * dh;
    T* dvout = dkout + (size_t)H * dh;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = quarter * DPT + i;
      if (d < dh) {
        dkout[d] = from_f<T>(dk[i] * scale);
        dvout[d] = from_f<T>(dv[i]);
      }
    }
  }
}

template <typename T, int DPT>
int launch_fwd(const T* qkv, T* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Tn, ROWS), (unsigned)(B * H));
  V4H_CUDA(launch_pdl(attn_fwd_kernel<T, DPT>, dim3(grid), dim3(ATHREADS), 0, s, qkv, o, lse, Tn, H, dh, 1.f / sqrtf((float)dh)));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}
template <typename T, int DPT>
int launch_bwd(const T* qkv, const T* o, const float* lse, const T* d_o, T* dqkv, int B, int Tn, int H, int dh,
               cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(Tn, ROWS), (unsigned)(B * H));
  const float scale = 1.f / sqrtf((float)dh);
  V4H_CUDA(launch_pdl(attn_bwd_dq_kernel<T, DPT>, dim3(grid), dim3(ATHREADS), 0, s, qkv, o, lse, d_o, dqkv, Tn, H, dh, scale));
  V4H_LAUNCH_CHECK();
  V4H_CUDA(launch_pdl(attn_bwd_dkv_kernel<T, DPT>, dim3(grid), dim3(ATHREADS), 0, s, qkv, o, lse, d_o, dqkv, Tn, H, dh, scale));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace

// Note on scaling: q (or k) is pre-multiplied by `scale`, so s = scale * q.k directly; the chain rule
// factor `scale` on dQ / dK is applied once at the store.
template <typename T>
int attention_fwd_simt(const T* qkv, T* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s) {
  V4H_REQUIRE(B * H <= 65535, "attention: batch*heads %d exceeds 65535", B * H);
  if (dh <= 32) return launch_fwd<T, 8>(qkv, o, lse, B, Tn, H, dh, s);
  if (dh <= 64) return launch_fwd<T, 16>(qkv, o, lse, B, Tn, H, dh, s);
  if (dh <= 80) return launch_fwd<T, 20>(qkv, o, lse, B, Tn, H, dh, s);
  if (dh <= 128) return launch_fwd<T, 32>(qkv, o, lse, B, Tn, H, dh, s);
  return fail(V4H_ERR_UNSUPPORTED, "attention: head_dim %d > 128", dh);
}
template <typename T>
int attention_bwd_simt(const T* qkv, const T* o, const float* lse, const T* d_o, T* dqkv, int B, int Tn, int H,
                       int dh, cudaStream_t s) {
  V4H_REQUIRE(B * H <= 65535, "attention: batch*heads %d exceeds 65535", B * H);
  if (dh <= 32) return launch_bwd<T, 8>(qkv, o, lse, d_o, dqkv, B, Tn, H, dh, s);
  if (dh <= 64) return launch_bwd<T, 16>(qkv, o, lse, d_o, dqkv, B, Tn, H, dh, s);
  if (dh <= 80) return launch_bwd<T, 20>(qkv, o, lse, d_o, dqkv, B, Tn, H, dh, s);
  if (dh <= 128) return launch_bwd<T, 32>(qkv, o, lse, d_o, dqkv, B, Tn, H, dh, s);
  return fail(V4H_ERR_UNSUPPORTED, "attention: head_dim %d > 128", dh);
}

template int attention_fwd_simt<float>(const float*, float*, float*, int, int, int, int, cudaStream_t);
template int attention_fwd_simt<bf16>(const bf16*, bf16*, float*, int, int, int, int, cudaStream_t);
template int attention_bwd_simt<float>(const float*, const float*, const float*, const float*, float*, int,
                                       int, int, int, cudaStream_t);
template int attention_bwd_simt<bf16>(const bf16*, const bf16*, const float*, const bf16*, bf16*, int, int,
                                      int, int, cudaStream_t);

}  // namespace v4h
