// Memory-bound helper kernels: CFM trajectory/loss, RK stage combination, embeddings, casts.
// All are streaming kernels sized at a multiple of the SM count with 128-bit accesses where the
// layout allows.
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int EW_THREADS = 256;
inline unsigned ew_grid(int64_t n, int per_thread = 4) {
  int64_t blocks = ceil_div(n, (int64_t)EW_THREADS * per_thread);
  const int64_t cap = 148 * 8;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ---------------------------------------------------------------- column sums (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int ld, float* __restrict__ out, int M, int N,
                              int rows_per_cta) {
  pdl_wait();
  // blockDim = (32, 8): 32 consecutive columns x 8 row lanes
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc = 0.f;
  if (col < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f(x[(size_t)r * ld + col]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + col, t);
  }
}


// vectorised variant: thread = 8 consecutive columns (16-byte loads of bf16, 2 x 16 bytes of fp32),
// blockDim (32, 8) = 256 columns x 8 row lanes
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, int ld, float* __restrict__ out,
                                                         int M, int N, int rows_per_cta) {
  pdl_wait();
  __shared__ float red[8][32][9];
  const int col = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    int r = r0 + threadIdx.y;
    for (; r + 8 < r1; r += 16) {  // two rows in flight
      float a[8], b[8];
      load8(x + (size_t)r * ld + col, a);
      load8(x + (size_t)(r + 8) * ld + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i] + b[i];
    }
    for (; r < r1; r += 8) {
      float a[8];
      load8(x + (size_t)r * ld + col, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += a[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  // 256 threads -> 256 columns of this CTA
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int c = blockIdx.x * 256 + t;
  if (c < N) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i][t >> 3][t & 7];
    atomicAdd(out + c, sum);
  }
}

// ---------------------------------------------------------------- timestep embedding
// reference nn/vit.py:368-389: cat(cos(t f), sin(t f)), f_i = exp(-ln(1e4) i / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int shared_t, float* __restrict__ out,
                                          bf16* __restrict__ out_bf, int B, int dim) {
  pdl_wait();
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx % half;
  const float tv = t[shared_t ? 0 : b];
  // match torch: exp(-log(10000) * arange / half) in fp32
  const float f = expf(-9.210340371976184f * (float)i / (float)half);
  const float arg = tv * f;
  const float cv = cosf(arg), sv = sinf(arg);
  out[(size_t)b * dim + i] = cv;
  out[(size_t)b * dim + half + i] = sv;
  if ((dim & 1) && i == 0) out[(size_t)b * dim + dim - 1] = 0.f;
  if (out_bf) {  // bf16 copy: the A operand of the first t_embedder Linear on the tensor cores
    out_bf[(size_t)b * dim + i] = __float2bfloat16_rn(cv);
    out_bf[(size_t)b * dim + half + i] = __float2bfloat16_rn(sv);
    if ((dim & 1) && i == 0) out_bf[(size_t)b * dim + dim - 1] = __float2bfloat16_rn(0.f);
  }
}

// ---------------------------------------------------------------- learnable positional embedding
// reference nn/vit.py:156-162: w = 2 pi freqs; pe = cat(sin(x w), cos(x w), sin(y w), cos(y w), sin(z w), cos(z w))
__global__ void pos_embed_fwd_kernel(const float* __restrict__ freqs, const float* __restrict__ pz,
                                     const float* __restrict__ py, const float* __restrict__ px,
                                     float* __restrict__ pe, int Tn, int F) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Tn * F) return;
  const int t = idx / F, f = idx % F;
  const float w = freqs[f] * 2.f * 3.14159265358979323846f;
  float* row = pe + (size_t)t * 6 * F;
  const float ax = px[t] * w, ay = py[t] * w, az = pz[t] * w;
  row[f] = sinf(ax);          row[F + f] = cosf(ax);
  row[2 * F + f] = sinf(ay);  row[3 * F + f] = cosf(ay);
  row[4 * F + f] = sinf(az);  row[5 * F + f] = cosf(az);
}

// grid = Tn CTAs, blockDim = 128: thread loops over the 6F columns, sums dh over the batch, applies
// d pe / d freq and accumulates into dfreqs with one atomic per (CTA, column).
__global__ void pos_embed_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ freqs,
                                     const float* __restrict__ pz, const float* __restrict__ py,
                                     const float* __restrict__ px, float* __restrict__ dfreqs, int B,
                                     int Tn, int F) {
  pdl_wait();
  const int t = blockIdx.x;
  const int D = 6 * F;
  const float two_pi = 2.f * 3.14159265358979323846f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float sum = 0.f;
    for (int b = 0; b < B; ++b) sum += dh[((size_t)b * Tn + t) * D + d];
    const int part = d / F, f = d % F;
    const float pos = part < 2 ? px[t] : (part < 4 ? py[t] : pz[t]);
    const float arg = pos * freqs[f] * two_pi;
    // d/dfreq sin(pos 2pi freq) = cos(arg) pos 2pi ; d/dfreq cos = -sin(arg) pos 2pi
    const float deriv = ((part & 1) ? -sinf(arg) : cosf(arg)) * pos * two_pi;
    atomicAdd(dfreqs + f, sum * deriv);
  }
}

// ---------------------------------------------------------------- casts
__global__ void silu_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, int64_t n) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(silu_f(x[i]));
}

__device__ __forceinline__ void cast_span(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n,
                                          int64_t start, int64_t stride) {
  // 8 elements per thread-iteration when both pointers are 16/32-byte aligned
  if ((((uintptr_t)src & 31) == 0) && (((uintptr_t)dst & 15) == 0)) {
    const int64_t n8 = n / 8;
    for (int64_t i = start; i < n8; i += stride) {
      float4 a = reinterpret_cast<const float4*>(src)[2 * i];
      float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
      reinterpret_cast<uint4*>(dst)[i] = o;
    }
    for (int64_t i = n8 * 8 + start; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
  } else {
    for (int64_t i = start; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

__global__ void cast_kernel(const float* __restrict__ x, bf16* __restrict__ out, int64_t n) {
  pdl_wait();
  cast_span(x, out, n, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// one launch for every weight tensor: blockIdx.y = job
__global__ void cast_many_kernel(const CastJob* __restrict__ jobs) {
  pdl_wait();
  const CastJob j = jobs[blockIdx.y];
  cast_span(j.src, j.dst, j.n, blockIdx.x * (int64_t)blockDim.x + threadIdx.x,
            (int64_t)gridDim.x * blockDim.x);
}

// out = x * silu'(pre)   (backward through the SiLU in front of the adaLN Linears)
__global__ void dsilu_mul_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                                 float* __restrict__ out, bf16* __restrict__ out_bf, int64_t n) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i] * dsilu_f(pre[i]);
    out[i] = v;
    if (out_bf) out_bf[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- CFM
// reference models/trajectories.py:5-8 + to_patches: one pass over the voxels of x1 and x0 emits
// x_t and (x1 - x0) directly in token order via the gather table.
__global__ void cfm_prepare_kernel(const float* __restrict__ x1, const float* __restrict__ x0,
                                   const float* __restrict__ t, const int32_t* __restrict__ table,
                                   float* __restrict__ xt, float* __restrict__ target, int per_sample) {
  pdl_wait();
  const int b = blockIdx.y;
  const float tv = t[b];
  const size_t base = (size_t)b * per_sample;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < per_sample; j += gridDim.x * blockDim.x) {
    const int src = table[j];
    const float a = x0[base + src], d = x1[base + src];
    xt[base + j] = (1.f - tv) * a + tv * d;
    target[base + j] = d - a;
  }
}

// reference models/base_model.py:217-218: mean((v - target)^2); also d loss / d v
__global__ void cfm_loss_kernel(const float* __restrict__ v, const float* __restrict__ target, int64_t n,
                                float inv_n, float grad_scale, float* __restrict__ loss_out,
                                float* __restrict__ dv) {
  pdl_wait();
  __shared__ float red[EW_THREADS / 32];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = v[i] - target[i];
    acc += d * d;
    if (dv) dv[i] = 2.f * d * inv_n * grad_scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < EW_THREADS / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(loss_out, t * inv_n);
  }
}

// out = y + a0 k0 + a1 k1 + a2 k2 + a3 k3  (RK stage states and the final 3/8-rule combination)
// out may alias y (the final RK combination updates the state in place): neither is __restrict__
__global__ void axpy4_kernel(float* out, const float* y, const float* __restrict__ k0,
                             float a0, const float* __restrict__ k1, float a1, const float* __restrict__ k2,
                             float a2, const float* __restrict__ k3, float a3, int64_t n) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float r = y[i];
    if (k0) r = fmaf(a0, k0[i], r);
    if (k1) r = fmaf(a1, k1[i], r);
    if (k2) r = fmaf(a2, k2[i], r);
    if (k3) r = fmaf(a3, k3[i], r);
    out[i] = r;
  }
}

}  // namespace

template <typename T>
int colsum_add(const T* x, int ld, float* out, int M, int N, cudaStream_t s) {
  if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int rows = M >= 4096 ? 128 : 64;
    dim3 vgrid((unsigned)ceil_div(N, 256), (unsigned)ceil_div(M, rows));
    V4H_CUDA(launch_pdl(colsum_vec_kernel<T>, dim3(vgrid), dim3(dim3(32, 8)), 0, s, x, ld, out, M, N, rows));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  const int rows_per_cta = 256;
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, rows_per_cta));
  V4H_CUDA(launch_pdl(colsum_kernel<T>, dim3(grid), dim3(dim3(32, 8)), 0, s, x, ld, out, M, N, rows_per_cta));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}
template int colsum_add<float>(const float*, int, float*, int, int, cudaStream_t);
template int colsum_add<bf16>(const bf16*, int, float*, int, int, cudaStream_t);

int timestep_embedding(const float* t, int shared_t, float* out, bf16* out_bf, int B, int dim, cudaStream_t s) {
  const int n = B * (dim / 2);
  V4H_CUDA(launch_pdl(timestep_embedding_kernel, dim3((unsigned)ceil_div(n, 128)), dim3(128), 0, s, t, shared_t, out, out_bf, B, dim));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int pos_embedding_fwd(const float* freqs, const float* pz, const float* py, const float* px, float* pe,
                      int Tn, int F, cudaStream_t s) {
  V4H_CUDA(launch_pdl(pos_embed_fwd_kernel, dim3((unsigned)ceil_div((int64_t)Tn * F, 128)), dim3(128), 0, s, freqs, pz, py, px, pe, Tn, F));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int pos_embedding_bwd(const float* dh, const float* freqs, const float* pz, const float* py, const float* px,
                      float* dfreqs, int B, int Tn, int F, cudaStream_t s) {
  V4H_CUDA(launch_pdl(pos_embed_bwd_kernel, dim3((unsigned)Tn), dim3(128), 0, s, dh, freqs, pz, py, px, dfreqs, B, Tn, F));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int silu_to_bf16(const float* x, bf16* out, int64_t n, cudaStream_t s) {
  V4H_CUDA(launch_pdl(silu_to_bf16_kernel, dim3(ew_grid(n, 1)), dim3(EW_THREADS), 0, s, x, out, n));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int dsilu_mul(const float* x, const float* pre, float* out, bf16* out_bf, int64_t n, cudaStream_t s) {
  V4H_CUDA(launch_pdl(dsilu_mul_kernel, dim3(ew_grid(n, 1)), dim3(EW_THREADS), 0, s, x, pre, out, out_bf, n));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int cast_f32_to_bf16(const float* x, bf16* out, int64_t n, cudaStream_t s) {
  V4H_CUDA(launch_pdl(cast_kernel, dim3(ew_grid(n, 8)), dim3(EW_THREADS), 0, s, x, out, n));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int cast_many_f32_to_bf16(const CastJob* jobs_dev, int njobs, int64_t max_n, cudaStream_t s) {
  unsigned gx = (unsigned)ceil_div(max_n, (int64_t)EW_THREADS * 8 * 4);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  V4H_CUDA(launch_pdl(cast_many_kernel, dim3(dim3(gx, (unsigned)njobs)), dim3(EW_THREADS), 0, s, jobs_dev));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int cfm_prepare(const float* x1, const float* x0, const float* t, const int32_t* table, float* xt_tok,
                float* target_tok, int64_t B, int per_sample, cudaStream_t s) {
  unsigned gx = (unsigned)ceil_div(per_sample, EW_THREADS * 4);
  V4H_CUDA(launch_pdl(cfm_prepare_kernel, dim3(dim3(gx, (unsigned)B)), dim3(EW_THREADS), 0, s, x1, x0, t, table, xt_tok, target_tok, per_sample));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int cfm_loss(const float* v, const float* target, int64_t n, float grad_scale, float* loss_out, float* dv,
             cudaStream_t s) {
  V4H_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), s));
  V4H_CUDA(launch_pdl(cfm_loss_kernel, dim3(ew_grid(n, 4)), dim3(EW_THREADS), 0, s, v, target, n, 1.f / (float)n, grad_scale, loss_out, dv));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int axpy4(float* out, const float* y, const float* k0, float a0, const float* k1, float a1, const float* k2,
          float a2, const float* k3, float a3, int64_t n, cudaStream_t s) {
  V4H_CUDA(launch_pdl(axpy4_kernel, dim3(ew_grid(n, 4)), dim3(EW_THREADS), 0, s, out, y, k0, a0, k1, a1, k2, a2, k3, a3, n));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace v4h
