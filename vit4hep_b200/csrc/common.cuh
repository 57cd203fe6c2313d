// Shared helpers for the vit4hep_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/vit4hep_b200.h"

namespace v4h {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error reporting
char* last_error_buffer();  // thread-local, 512 bytes (defined in api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define V4H_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::v4h::fail(V4H_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
                         cudaGetErrorString(_e));                                        \
  } while (0)

// every kernel launch in the library is followed by this: counts the launch (v4h_launch_count) and
// surfaces launch errors
void count_launch();  // api.cu
#define V4H_LAUNCH_CHECK()       \
  do {                           \
    ::v4h::count_launch();       \
    V4H_CUDA(cudaGetLastError()); \
  } while (0)

// Optional per-kernel-class timing (v4h_profile_begin / v4h_profile_end): when enabled, a ProfScope
// brackets the launches made during its lifetime with CUDA events on the launching stream and books
// their algorithmic flops / bytes under `tag`.  Disabled: a single branch.
bool profiling_enabled();
void profile_open(const char* tag, double flops, double bytes, cudaStream_t s, int* slot);
void profile_close(int slot, cudaStream_t s);
struct ProfScope {
  int slot = -1;
  cudaStream_t stream;
  ProfScope(const char* tag, double flops, double bytes, cudaStream_t s) : stream(s) {
    if (profiling_enabled()) profile_open(tag, flops, bytes, s, &slot);
  }
  ~ProfScope() {
    if (slot >= 0) profile_close(slot, stream);
  }
};

#define V4H_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return ::v4h::fail(V4H_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define V4H_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != V4H_OK) return _rc; \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------- element access
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- activations
enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU_TANH = 2 };

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu_f(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}
// tanh-approximated GELU (torch nn.GELU(approximate="tanh"), reference nn/vit.py:314-315)
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}
__device__ __forceinline__ float dgelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float th = tanhf(k0 * (x + k1 * x * x2));
  float dinner = k0 * (1.f + 3.f * k1 * x2);
  return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * dinner;
}
template <int ACT> __device__ __forceinline__ float act_f(float x) {
  if (ACT == ACT_SILU) return silu_f(x);
  if (ACT == ACT_GELU_TANH) return gelu_tanh_f(x);
  return x;
}
template <int ACT> __device__ __forceinline__ float dact_f(float x) {
  if (ACT == ACT_SILU) return dsilu_f(x);
  if (ACT == ACT_GELU_TANH) return dgelu_tanh_f(x);
  return 1.f;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- GEMM epilogues
// Shared by the SIMT fp32 GEMM (gemm_simt.cu) and the tcgen05 GEMM (gemm_umma.cu): both call
// epilogue_run<...>() on a run of NV consecutive columns of one output row.
enum EpiKind {
  EPI_BIAS_ACT = 0,  // pre = acc + bias[n] + addend[row % addend_rows, n]; out2 = pre; out = act(pre)
  EPI_GATE_RES = 1,  // y = acc + bias[n]; out2 = y; res_out = res_in + gate[row / T][n] * y
  EPI_DACT = 2,      // out = acc * act'(aux[row, n])
  EPI_ATOMIC = 3     // atomicAdd(out_f32[row, n], acc)    (split-K weight gradients)
};

struct EpiParams {
  const float* bias = nullptr;
  void* out = nullptr;   // TOut (M, ldo)
  void* out2 = nullptr;  // TOut (M, ldo) optional
  int ldo = 0;
  const float* addend = nullptr;  // fp32 (addend_rows, ld_addend)
  int addend_rows = 0;            // 0: one row per output row
  int ld_addend = 0;
  const float* gate = nullptr;  // fp32, gate[b * mod_stride + n]
  int mod_stride = 0;
  int rows_per_sample = 1;
  const float* res_in = nullptr;  // fp32 (M, N)
  float* res_out = nullptr;       // fp32 (M, N)
  const void* aux = nullptr;      // TOut (M, ld_aux)
  int ld_aux = 0;
};

template <int EPI, int ACT, typename TOut, int NV>
__device__ __forceinline__ void epilogue_run(const EpiParams& p, int row, int col0, int ncols_valid,
                                             const float (&acc)[NV]) {
  // ncols_valid: number of leading entries of acc that are inside the matrix
  if (EPI == EPI_BIAS_ACT) {
    TOut* o = reinterpret_cast<TOut*>(p.out) + (size_t)row * p.ldo + col0;
    TOut* o2 = p.out2 ? reinterpret_cast<TOut*>(p.out2) + (size_t)row * p.ldo + col0 : nullptr;
    const float* ad = nullptr;
    if (p.addend) {
      int ar = p.addend_rows > 0 ? row % p.addend_rows : row;
      ad = p.addend + (size_t)ar * p.ld_addend + col0;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i < ncols_valid) {
        float pre = acc[i];
        if (p.bias) pre += p.bias[col0 + i];
        if (ad) pre += ad[i];
        if (o2) o2[i] = from_f<TOut>(pre);
        o[i] = from_f<TOut>(act_f<ACT>(pre));
      }
    }
  } else if (EPI == EPI_GATE_RES) {
    TOut* o2 = p.out2 ? reinterpret_cast<TOut*>(p.out2) + (size_t)row * p.ldo + col0 : nullptr;
    const float* gate = p.gate + (size_t)(row / p.rows_per_sample) * p.mod_stride + col0;
    const float* rin = p.res_in + (size_t)row * p.ldo + col0;
    float* rout = p.res_out + (size_t)row * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i < ncols_valid) {
        float y = acc[i] + (p.bias ? p.bias[col0 + i] : 0.f);
        if (o2) o2[i] = from_f<TOut>(y);
        rout[i] = rin[i] + gate[i] * y;
      }
    }
  } else if (EPI == EPI_DACT) {
    TOut* o = reinterpret_cast<TOut*>(p.out) + (size_t)row * p.ldo + col0;
    const TOut* aux = reinterpret_cast<const TOut*>(p.aux) + (size_t)row * p.ld_aux + col0;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < ncols_valid) o[i] = from_f<TOut>(acc[i] * dact_f<ACT>(to_f(aux[i])));
  } else {  // EPI_ATOMIC
    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < ncols_valid) atomicAdd(o + i, acc[i]);
  }
}

}  // namespace v4h
