// Shared helpers for the vit4hep_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>

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
// V4H_LAUNCH_SYNC=1 (debug): synchronise after every launch so that a device fault is reported at the
// launch that caused it
bool launch_sync_enabled();  // api.cu
#define V4H_LAUNCH_CHECK()                                              \
  do {                                                                  \
    ::v4h::count_launch();                                              \
    V4H_CUDA(cudaGetLastError());                                       \
    if (::v4h::launch_sync_enabled()) V4H_CUDA(cudaDeviceSynchronize()); \
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

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library starts with pdl_wait() (griddepcontrol.wait: blocks until the preceding
// kernel of the stream has completed and its writes are visible) and is launched with the programmatic
// stream-serialization attribute, so the launch latency and the prologue of kernel N+1 overlap the tail of
// kernel N instead of adding ~2-3 us per launch to a step of ~150 launches.  V4H_PDL=0 turns it off.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();  // api.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
// ACT_GELU_TANH_FAST: same function with the hardware tanh.approx (abs. error ~5e-4, below bf16
// resolution); used by the bf16 tensor-core epilogues only
enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU_TANH = 2, ACT_GELU_TANH_FAST = 3, ACT_RELU = 4 };
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
// written as explicit multiply-adds: 3 FMUL + 2 FFMA + 1 MUFU per element (the epilogues that call these are
// issue-bound; the straightforward expression compiled to 9 instructions)
__device__ __forceinline__ float gelu_tanh_fast_f(float x) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float th = tanh_fast(fmaf(x2, k0k1, k0) * x);  // tanh(k0 (x + k1 x^3))
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
__device__ __forceinline__ float dgelu_tanh_fast_f(float x) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float th = tanh_fast(fmaf(x2, k0k1, k0) * x);
  const float dinner = fmaf(x2, 3.f * k0k1, k0);  // k0 (1 + 3 k1 x^2)
  const float q = (0.5f * x) * fmaf(-th, th, 1.f);  // 0.5 x sech^2
  return fmaf(q, dinner, fmaf(th, 0.5f, 0.5f));
}
template <int ACT> __device__ __forceinline__ float act_f(float x) {
  if (ACT == ACT_GELU_TANH_FAST) return gelu_tanh_fast_f(x);
  if (ACT == ACT_SILU) return silu_f(x);
  if (ACT == ACT_GELU_TANH) return gelu_tanh_f(x);
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  return x;
}
template <int ACT> __device__ __forceinline__ float dact_f(float x) {
  if (ACT == ACT_GELU_TANH_FAST) return dgelu_tanh_fast_f(x);
  if (ACT == ACT_SILU) return dsilu_f(x);
  if (ACT == ACT_GELU_TANH) return dgelu_tanh_f(x);
  if (ACT == ACT_RELU) return x > 0.f ? 1.f : 0.f;
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
  // EPI_ATOMIC only: output column `extra_col` (an appended "ones" column of the B operand, i.e. the
  // column sums of A^T) is accumulated into extra_out[row] instead of out[row, extra_col]
  float* extra_out = nullptr;
  int extra_col = -1;
  bool vec_ok = false;  // set by the launcher (epilogue_vec_ok)
  // EPI_GATE_RES only, tcgen05 engine (gemm_gate_res_ln): the LayerNorm + adaLN modulation that consumes res_out,
  // fused into this epilogue: ln_out[row, :] = LN(res_out[row, :]) * (1 + ln_scale[b, :]) + ln_shift[b, :] in bf16
  // (row pitch ld_ln >= N; the column N of a wider pitch receives the "ones" column of layernorm.cu), statistics
  // (mean, rstd) per row into ln_stats when non-null.  Shift / scale rows use mod_stride like the gate.
  const float* ln_shift = nullptr;
  const float* ln_scale = nullptr;
  void* ln_out = nullptr;
  int ld_ln = 0;
  float2* ln_stats = nullptr;
  float ln_eps = 1e-6f;
};

// ---- vector access helpers: NV consecutive values of one row, 16-byte transactions when `vec`
template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[NV], bool vec) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = p[i];
  }
}
template <int NV>
__device__ __forceinline__ void load_row(const bf16* __restrict__ p, float (&v)[NV], bool vec) {
  if (vec) {
    if (NV % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NV / 8; ++i) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + 8 * i);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
          v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const uint2 t = *reinterpret_cast<const uint2*>(p + 4 * i);
        const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
        const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
        v[4 * i] = f0.x; v[4 * i + 1] = f0.y; v[4 * i + 2] = f1.x; v[4 * i + 3] = f1.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __bfloat162float(p[i]);
  }
}
template <int NV>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[NV], bool vec, int nvalid) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) p[i] = v[i];
  }
}
template <int NV>
__device__ __forceinline__ void store_row(bf16* __restrict__ p, const float (&v)[NV], bool vec, int nvalid) {
  if (vec) {
    if (NV % 8 == 0) {
#pragma unroll
      for (int i = 0; i < NV / 8; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
          w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p + 8 * i) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[4 * i], v[4 * i + 1]);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(v[4 * i + 2], v[4 * i + 3]);
        *reinterpret_cast<uint2*>(p + 4 * i) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
  }
}

// One run of NV consecutive columns [col0, col0 + NV) of output row `row`; the first `ncols_valid`
// are inside the matrix.  When the run is complete and p.vec_ok says every row pitch / base pointer
// involved keeps it 16-byte aligned, all accesses are 128-bit.
template <int EPI, int ACT, typename TOut, int NV>
__device__ __forceinline__ void epilogue_run(const EpiParams& p, int row, int col0, int ncols_valid,
                                             const float (&acc)[NV]) {
  const bool vec = p.vec_ok && ncols_valid == NV;
  if (EPI == EPI_BIAS_ACT) {
    float pre[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pre[i] = acc[i];
    if (p.bias) {
      float b[NV];
      if (vec) load_row<NV>(p.bias + col0, b, true);
      else {
#pragma unroll
        for (int i = 0; i < NV; ++i) b[i] = i < ncols_valid ? p.bias[col0 + i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) pre[i] += b[i];
    }
    if (p.addend) {
      const int ar = p.addend_rows > 0 ? row % p.addend_rows : row;
      const float* ad = p.addend + (size_t)ar * p.ld_addend + col0;
      float a[NV];
      if (vec) load_row<NV>(ad, a, true);
      else {
#pragma unroll
        for (int i = 0; i < NV; ++i) a[i] = i < ncols_valid ? ad[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) pre[i] += a[i];
    }
    if (p.out2) store_row<NV>(reinterpret_cast<TOut*>(p.out2) + (size_t)row * p.ldo + col0, pre, vec, ncols_valid);
    if (ACT != ACT_NONE) {
#pragma unroll
      for (int i = 0; i < NV; ++i) pre[i] = act_f<ACT>(pre[i]);
    }
    store_row<NV>(reinterpret_cast<TOut*>(p.out) + (size_t)row * p.ldo + col0, pre, vec, ncols_valid);
  } else if (EPI == EPI_GATE_RES) {
    const float* gate = p.gate + (size_t)(row / p.rows_per_sample) * p.mod_stride + col0;
    const float* rin = p.res_in + (size_t)row * p.ldo + col0;
    float y[NV], gt[NV], r[NV];
    if (vec) {
      load_row<NV>(gate, gt, true);
      load_row<NV>(rin, r, true);
      if (p.bias) load_row<NV>(p.bias + col0, y, true);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool ok = i < ncols_valid;
        gt[i] = ok ? gate[i] : 0.f; r[i] = ok ? rin[i] : 0.f; y[i] = (ok && p.bias) ? p.bias[col0 + i] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      y[i] = acc[i] + (p.bias ? y[i] : 0.f);
      r[i] = r[i] + gt[i] * y[i];
    }
    if (p.out2) store_row<NV>(reinterpret_cast<TOut*>(p.out2) + (size_t)row * p.ldo + col0, y, vec, ncols_valid);
    store_row<NV>(p.res_out + (size_t)row * p.ldo + col0, r, vec, ncols_valid);
  } else if (EPI == EPI_DACT) {
    const TOut* aux = reinterpret_cast<const TOut*>(p.aux) + (size_t)row * p.ld_aux + col0;
    float u[NV];
    if (vec) load_row<NV>(aux, u, true);
    else {
#pragma unroll
      for (int i = 0; i < NV; ++i) u[i] = i < ncols_valid ? to_f(aux[i]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) u[i] = acc[i] * dact_f<ACT>(u[i]);
    store_row<NV>(reinterpret_cast<TOut*>(p.out) + (size_t)row * p.ldo + col0, u, vec, ncols_valid);
  } else {  // EPI_ATOMIC
    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
    if (vec) {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(acc[4 * i]),
                     "f"(acc[4 * i + 1]), "f"(acc[4 * i + 2]), "f"(acc[4 * i + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i < ncols_valid) {
          if (p.extra_out != nullptr && col0 + i == p.extra_col) atomicAdd(p.extra_out + row, acc[i]);
          else atomicAdd(o + i, acc[i]);
        }
    }
  }
}

// host side: can every row run of this epilogue use 128-bit accesses?  (col0 is always a multiple of NV)
inline bool epilogue_vec_ok(const EpiParams& p, int epi, bool out_bf16) {
  auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  const int unit = out_bf16 ? 8 : 4;  // elements per 16 bytes of TOut
  if (p.ldo % unit) return false;
  if (!al(p.out) || !al(p.out2) || !al(p.bias) || !al(p.addend) || !al(p.gate) || !al(p.res_in) || !al(p.res_out) ||
      !al(p.aux))
    return false;
  if (p.addend && p.ld_addend % 4) return false;
  if (p.gate && p.mod_stride % 4) return false;
  if (epi == EPI_GATE_RES && p.ldo % 4) return false;
  if (p.aux && p.ld_aux % unit) return false;
  return true;
}

}  // namespace v4h
