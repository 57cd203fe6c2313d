// tcgen05 / TMEM attention forward and backward for sm_100a (bf16 operands, fp32 accumulation and
// softmax): softmax(q k^T / sqrt(dh)) v per (sample, head), no mask, no dropout
// (reference nn/vit.py:425-451 with the xformers memory_efficient_attention call it makes).
//
// Layouts are the ones the qkv Linear produces (reference nn/vit.py:427): qkv (B, T, 3, H, dh),
// o / d_o (B, T, H, dh), lse and delta (B, H, T) fp32, lse = log sum_j exp(s_ij) of the scaled scores.
//
// Every operand tile lives in shared memory in ONE physical format ("G8"): element (row, col) of a
// [ROWS x COLS] bf16 tile sits at  (col / 8) * gstride + row * 16 + (col % 8) * 2  bytes, i.e. 8x8 core
// matrices of 128 contiguous bytes, the non-swizzled canonical UMMA layout.  The same bytes are read
//   - as a K-major operand  (rows = M/N index, cols = K index):  LBO = gstride, SBO = 128
//   - as an MN-major operand (rows = K index, cols = M/N index): LBO = 128,     SBO = gstride
// so Q, K, V, dO are staged once and serve both Q K^T-like and P V-like products.  gstride carries one
// 16-byte pad so that the 16-byte cp.async / st.shared writes of a warp spread over all banks.
// Head dims that are not multiples of 64 (dh = 80 here) cost nothing: K steps are 16 wide.
//
// Three kernels, 128 threads each (thread = one TMEM lane = one row of the 128-row M tile):
//   fwd  : CTA per (128 queries, sample-head); loops over key blocks with an online softmax
//          S = Q K^T -> TMEM, p = exp2(..) -> bf16 P in smem, O += P V through TMEM
//   dq   : CTA per (128 queries, sample-head); S and dP = dO V^T in TMEM, dS -> smem, dQ += dS K
//          accumulated in TMEM over the key blocks; also emits delta = rowsum(dO * O)
//   dkv  : CTA per (128 keys, sample-head); S^T = K Q^T and dP^T = V dO^T in TMEM, P^T / dS^T -> smem,
//          dV += P^T dO, dK += dS^T Q accumulated in TMEM over the query blocks
// Rows / keys beyond T are zero-filled on load and masked in the softmax.
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.cuh"
#include "umma.cuh"

namespace v4h {

using namespace sm100;

namespace {

constexpr int ATT_THREADS = 128;
constexpr int MT = 128;  // rows of the M tile = TMEM lanes

__device__ __forceinline__ uint64_t desc_ns(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout type bits [61,64) = 0: no swizzle
  return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_result, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// G8 tile geometry
__host__ __device__ constexpr uint32_t g8_stride(int rows) { return (uint32_t)(rows + 1) * 16u; }
__host__ __device__ constexpr uint32_t g8_bytes(int rows, int cols) { return (uint32_t)(cols / 8) * g8_stride(rows); }
// allocation of a tile that TMA may write: 128-byte aligned extents
__host__ __device__ constexpr uint32_t g8_alloc(int rows, int cols) { return (g8_bytes(rows, cols) + 127u) & ~127u; }

// Stage `rows` x DHP (bf16) from global (row pitch `ld` elements) into a G8 tile; rows >= rows_valid
// and column groups >= dh are zero-filled.
// Only the first `load_rows` rows are touched (the rest of the tile keeps whatever it held: rows that no
// stored result depends on).
template <int DHP, int NTHREADS = ATT_THREADS>
__device__ __forceinline__ void stage_tile(uint32_t dst, int rows, const bf16* __restrict__ src, size_t ld,
                                           int rows_valid, int dh, int load_rows = -1) {
  constexpr int NCG = DHP / 8;        // 16-byte column groups per row
  constexpr int RPI = NTHREADS / NCG;  // rows per sweep: every thread keeps one column group
  if ((int)threadIdx.x >= RPI * NCG) return;
  const int r0 = (int)threadIdx.x / NCG, cg = (int)threadIdx.x - r0 * NCG;
  const bool col_ok = cg * 8 < dh;
  const bf16* p = src + (size_t)r0 * ld + cg * 8;
  uint32_t d = dst + cg * g8_stride(rows) + r0 * 16;
  const int nload = load_rows < 0 ? rows : load_rows;
  for (int row = r0; row < nload; row += RPI, p += (size_t)RPI * ld, d += RPI * 16) {
    const bool ok = col_ok && row < rows_valid;
    cp_async16(d, ok ? (const void*)p : (const void*)src, ok);
  }
}

// D[tmem 128 x n] (+)= A[128 x 16*ksteps] B, A always K-major G8 (128 rows); B either K-major G8
// (rows = N index) or MN-major G8 (rows = K index).  One thread.
__device__ __forceinline__ void issue_mma(uint32_t d_tmem, uint32_t a_base, uint32_t a_gs, uint32_t b_base,
                                          uint32_t b_gs, bool b_mn, int n, int ksteps, bool accumulate_first) {
  const uint32_t idesc = make_idesc_bf16(MT, n, false, b_mn);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t ad = desc_ns(a_base + 2 * ks * a_gs, a_gs, 128);
    const uint64_t bd = b_mn ? desc_ns(b_base + ks * 256, 128, b_gs) : desc_ns(b_base + 2 * ks * b_gs, b_gs, 128);
    umma_bf16(d_tmem, ad, bd, idesc, (accumulate_first || ks > 0) ? 1u : 0u);
  }
}

struct AttnSmem {
  uint64_t bar;      // MMA completion
  uint64_t ld_bar;   // TMA operand loads
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint8_t* align128(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127));
}

// common prologue: barrier init + TMEM allocation; returns the TMEM base address
__device__ __forceinline__ uint32_t attn_prologue(AttnSmem* ctl, uint32_t tmem_cols) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&ctl->bar, 1);
    mbar_init(&ctl->ld_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_dyn(&ctl->tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // only now may dependents start: this CTA already holds its TMEM columns, so a dependent CTA that lands on
  // the same SM and allocates TMEM in its own prologue can never starve it
  pdl_wait();
  return ctl->tmem_slot;
}
__device__ __forceinline__ void attn_epilogue(uint32_t tmem_base, uint32_t tmem_cols) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, tmem_cols);
  }
}
// all generic-proxy smem writes of this thread are done -> visible to the tensor core; whole CTA
__device__ __forceinline__ void publish_smem_and_sync() {
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

struct AttnArgs {
  const bf16* qkv;
  bf16* o;          // fwd: output; bwd: forward output (read)
  float* lse;
  const bf16* d_o;
  float* delta;
  bf16* dqkv;
  int T, H, dh;
  int BN;           // key (fwd, dq) or query (dkv) block, multiple of 16
  int nblocks;
  uint32_t tmem_cols;
  float scale;       // dh^-0.5
  float scale_log2;  // scale * log2(e)
  long long* dbg;    // optional cycle counters per phase (v4h_debug_attention_counters)
  int use_tma;       // operand tiles staged by TMA (5-d tensor maps writing the G8 layout) instead of cp.async
};

struct ALap {  // cycle accounting of thread 0 of each CTA
  long long* dbg; long long t; long long acc[10];
  __device__ __forceinline__ explicit ALap(long long* d) : dbg(threadIdx.x == 0 ? d : nullptr), t(0) {
    if (dbg) { t = clock64(); for (int i = 0; i < 10; ++i) acc[i] = 0; }
  }
  __device__ __forceinline__ void lap(int i) { if (dbg) { const long long n = clock64(); acc[i] += n - t; t = n; } }
  __device__ __forceinline__ void flush() {
    if (dbg) for (int i = 0; i < 10; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(dbg + i), (unsigned long long)acc[i]);
  }
};

// ------------------------------------------------------------------------------------------ forward
// tmQ / tmKV: 5-d views (8 elements, T rows, dh/8 column groups, 3H, B) of qkv with boxes of 128 / BN rows: one
// TMA load writes a whole operand tile in the G8 layout (group stride rows * 16, no pad)
template <int DHP>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmKV,
                                                                    const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align128(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  const int BN = a.BN, T = a.T, H = a.H, dh = a.dh;
  const uint32_t sQ = smem_u32(smem + 128);
  const uint32_t sK = sQ + g8_alloc(MT, DHP);
  const uint32_t sV = sK + g8_alloc(BN, DHP);
  const uint32_t sP = sV + g8_alloc(BN, DHP);
  uint8_t* sP_ptr = smem + 128 + g8_alloc(MT, DHP) + 2 * g8_alloc(BN, DHP);
  const bool tma = a.use_tma != 0;
  const uint32_t gsQ = tma ? MT * 16u : g8_stride(MT), gsKV = tma ? (uint32_t)BN * 16u : g8_stride(BN), gsP = g8_stride(MT);
  uint32_t ld_phase = 0;

  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int q0 = blockIdx.x * MT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ld = (size_t)3 * H * dh;
  const bf16* qbase = a.qkv + (size_t)b * T * ld + (size_t)hd * dh;
  const bf16* kbase = qbase + (size_t)H * dh;
  const bf16* vbase = qbase + (size_t)2 * H * dh;

  ALap L(a.dbg);
  const uint32_t tmem = attn_prologue(ctl, a.tmem_cols);
  L.lap(0);
  const uint32_t tS = tmem, tO = tmem + (uint32_t)BN;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  if (!tma) stage_tile<DHP>(sQ, MT, qbase + (size_t)q0 * ld, ld, T - q0, dh, min(MT, (T - q0 + 15) / 16 * 16));

  float o_acc[DHP];
#pragma unroll
  for (int i = 0; i < DHP; ++i) o_acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const bool warp_active = q0 + warp * 32 < T;  // any valid query row in this warp

  for (int blk = 0; blk < a.nblocks; ++blk) {
    const int n0 = blk * BN;
    const int nvalid = min(BN, T - n0);
    if (tma) {
      if (tid == 0) {
        const uint32_t kv_bytes = (uint32_t)BN * DHP * 2;
        mbar_expect_tx(&ctl->ld_bar, 2 * kv_bytes + (blk == 0 ? (uint32_t)MT * DHP * 2 : 0u));
        if (blk == 0) tma_load_5d(smem + 128, &tmQ, &ctl->ld_bar, 0, q0, 0, hd, b);
        tma_load_5d(smem + 128 + g8_alloc(MT, DHP), &tmKV, &ctl->ld_bar, 0, n0, 0, H + hd, b);
        tma_load_5d(smem + 128 + g8_alloc(MT, DHP) + g8_alloc(BN, DHP), &tmKV, &ctl->ld_bar, 0, n0, 0, 2 * H + hd, b);
      }
      L.lap(1);
      mbar_wait(&ctl->ld_bar, ld_phase); ld_phase ^= 1;
    } else {
      stage_tile<DHP>(sK, BN, kbase + (size_t)n0 * ld, ld, nvalid, dh);
      stage_tile<DHP>(sV, BN, vbase + (size_t)n0 * ld, ld, nvalid, dh);
      L.lap(1);
      cp_async_wait_all();
    }
    L.lap(2);
    publish_smem_and_sync();
    L.lap(3);
    if (tid == 0) {
      issue_mma(tS, sQ, gsQ, sK, gsKV, false, BN, DHP / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(4);

    // ---- online softmax on this thread's row (warps whose 32 rows are all beyond T skip the work; their
    // P rows stay garbage, which only reaches their own, never stored, O rows): pass 1 = row maximum
    float mx = -INFINITY, corr = 0.f, lsum = 0.f;
    if (warp_active) {
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld16(tS + lane_off + c0, v);
        tmem_ld_wait();
        if (c0 + 16 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, c0 + i < nvalid ? v[i] : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_run, mx * a.scale_log2);
      corr = exp2_fast(m_run - m_new);  // first block: exp2(-inf) = 0
      m_run = m_new;
      // pass 2: p = exp2(s * scale_log2 - m), written as the bf16 A operand of P V
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld16(tS + lane_off + c0, v);
        tmem_ld_wait();
        uint32_t w[8];
        if (c0 + 16 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = exp2_fast(fmaf(v[i], a.scale_log2, -m_new));
            const float p1 = exp2_fast(fmaf(v[i + 1], a.scale_log2, -m_new));
            lsum += p0 + p1;
            w[i / 2] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = c0 + i < nvalid ? exp2_fast(fmaf(v[i], a.scale_log2, -m_new)) : 0.f;
            const float p1 = c0 + i + 1 < nvalid ? exp2_fast(fmaf(v[i + 1], a.scale_log2, -m_new)) : 0.f;
            lsum += p0 + p1;
            w[i / 2] = pack_bf16(p0, p1);
          }
        }
        uint8_t* dst = sP_ptr + (size_t)(c0 / 8) * gsP + tid * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + gsP) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    l_run = l_run * corr + lsum;
    L.lap(5);
    publish_smem_and_sync();
    L.lap(6);
    if (tid == 0) {
      issue_mma(tO, sP, gsP, sV, gsKV, true, DHP, BN / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(7);
    if (warp_active) {
#pragma unroll
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        float v[16];
        tmem_ld16(tO + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], corr, v[i]);
      }
    }
    // the next iteration overwrites sK / sV / sP and the S / O accumulators: both MMAs have completed
    // (waited above); order this thread's TMEM reads before the next MMA issue
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  const int q = q0 + tid;
  if (q < T) {
    const float inv = 1.f / l_run;
    bf16* orow = a.o + ((size_t)b * T + q) * H * dh + (size_t)hd * dh;
#pragma unroll
    for (int c = 0; c < DHP; c += 8) {
      if (c < dh) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = pack_bf16(o_acc[c + 2 * i] * inv, o_acc[c + 2 * i + 1] * inv);
        *reinterpret_cast<uint4*>(orow + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    a.lse[(size_t)bh * T + q] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  }
  L.lap(8);
  attn_epilogue(tmem, a.tmem_cols);
  L.lap(9);
  L.flush();
}

// ------------------------------------------------------------------------------------------ dQ (+ delta)
template <int DHP>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_umma_kernel(const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align128(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  const int BN = a.BN, T = a.T, H = a.H, dh = a.dh;
  const uint32_t sQ = smem_u32(smem + 128);
  const uint32_t sdO = sQ + g8_bytes(MT, DHP);
  const uint32_t sK = sdO + g8_bytes(MT, DHP);
  const uint32_t sV = sK + g8_bytes(BN, DHP);
  const uint32_t sdS = sV + g8_bytes(BN, DHP);
  uint8_t* sdS_ptr = smem + 128 + 2 * g8_bytes(MT, DHP) + 2 * g8_bytes(BN, DHP);
  const uint32_t gsQ = g8_stride(MT), gsKV = g8_stride(BN), gsS = g8_stride(MT);

  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int q0 = blockIdx.x * MT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const bf16* qbase = a.qkv + (size_t)b * T * ld + (size_t)hd * dh;
  const bf16* kbase = qbase + (size_t)H * dh;
  const bf16* vbase = qbase + (size_t)2 * H * dh;
  const bf16* dobase = a.d_o + (size_t)b * T * ldo + (size_t)hd * dh;
  const bf16* obase = a.o + (size_t)b * T * ldo + (size_t)hd * dh;

  const uint32_t tmem = attn_prologue(ctl, a.tmem_cols);
  const uint32_t tS = tmem, tdP = tmem + (uint32_t)BN, tdQ = tmem + 2u * (uint32_t)BN;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  stage_tile<DHP>(sQ, MT, qbase + (size_t)q0 * ld, ld, T - q0, dh);
  stage_tile<DHP>(sdO, MT, dobase + (size_t)q0 * ldo, ldo, T - q0, dh);

  // this thread's row statistics: delta = sum_d dO * O, lse (in log2 units)
  const int q = q0 + tid;
  float delta = 0.f, lse2 = 0.f;
  if (q < T) {
    const bf16* dor = dobase + (size_t)q * ldo;
    const bf16* orow = obase + (size_t)q * ldo;
    uint4 xv[DHP / 8], yv[DHP / 8];  // all loads of the row in flight before the first use
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      const bool in = c * 8 < dh;
      xv[c] = in ? *reinterpret_cast<const uint4*>(dor + c * 8) : make_uint4(0, 0, 0, 0);
      yv[c] = in ? *reinterpret_cast<const uint4*>(orow + c * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < DHP / 8; ++c) {
      const uint32_t xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, ys[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[i]));
        const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[i]));
        delta = fmaf(fx.x, fy.x, delta);
        delta = fmaf(fx.y, fy.y, delta);
      }
    }
    a.delta[(size_t)bh * T + q] = delta;
    lse2 = a.lse[(size_t)bh * T + q] * 1.4426950408889634f;
  }

  for (int blk = 0; blk < a.nblocks; ++blk) {
    const int n0 = blk * BN;
    const int nvalid = min(BN, T - n0);
    stage_tile<DHP>(sK, BN, kbase + (size_t)n0 * ld, ld, nvalid, dh);
    stage_tile<DHP>(sV, BN, vbase + (size_t)n0 * ld, ld, nvalid, dh);
    cp_async_wait_all();
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tS, sQ, gsQ, sK, gsKV, false, BN, DHP / 16, false);
      issue_mma(tdP, sdO, gsQ, sV, gsKV, false, BN, DHP / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float s[16], dp[16];
      tmem_ld16(tS + lane_off + c0, s);
      tmem_ld16(tdP + lane_off + c0, dp);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = c0 + i < nvalid ? exp2_fast(fmaf(s[i], a.scale_log2, -lse2)) : 0.f;
        const float p1 = c0 + i + 1 < nvalid ? exp2_fast(fmaf(s[i + 1], a.scale_log2, -lse2)) : 0.f;
        w[i / 2] = pack_bf16(p0 * (dp[i] - delta), p1 * (dp[i + 1] - delta));
      }
      uint8_t* dst = sdS_ptr + (size_t)(c0 / 8) * gsS + tid * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + gsS) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tdQ, sdS, gsS, sK, gsKV, true, DHP, BN / 16, blk > 0);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // tcgen05.ld is .sync.aligned over the whole warp: every thread loads, valid rows store
  bf16* out = a.dqkv + ((size_t)b * T + min(q, T - 1)) * ld + (size_t)hd * dh;
#pragma unroll
  for (int c0 = 0; c0 < DHP; c0 += 16) {
    float v[16];
    tmem_ld16(tdQ + lane_off + c0, v);
    tmem_ld_wait();
    if (q < T) {
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (c0 + 8 * h8 < dh) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16(v[8 * h8 + 2 * i] * a.scale, v[8 * h8 + 2 * i + 1] * a.scale);
          *reinterpret_cast<uint4*>(out + c0 + 8 * h8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  attn_epilogue(tmem, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------ dK, dV
template <int DHP>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_umma_kernel(const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align128(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  const int BQ = a.BN, T = a.T, H = a.H, dh = a.dh;
  float* s_lse = reinterpret_cast<float*>(smem + 128);       // [BQ] lse * log2(e), +inf for padded queries
  float* s_delta = s_lse + 256;                              // [BQ]
  uint8_t* tiles = smem + 128 + 2048;
  const uint32_t sK = smem_u32(tiles);
  const uint32_t sV = sK + g8_bytes(MT, DHP);
  const uint32_t sQ = sV + g8_bytes(MT, DHP);
  const uint32_t sdO = sQ + g8_bytes(BQ, DHP);
  const uint32_t sPT = sdO + g8_bytes(BQ, DHP);
  const uint32_t sdST = sPT + g8_bytes(MT, BQ);
  uint8_t* sPT_ptr = tiles + 2 * g8_bytes(MT, DHP) + 2 * g8_bytes(BQ, DHP);
  uint8_t* sdST_ptr = sPT_ptr + g8_bytes(MT, BQ);
  const uint32_t gsKV = g8_stride(MT), gsQ = g8_stride(BQ), gsS = g8_stride(MT);

  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int k0 = blockIdx.x * MT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const bf16* qbase = a.qkv + (size_t)b * T * ld + (size_t)hd * dh;
  const bf16* kbase = qbase + (size_t)H * dh;
  const bf16* vbase = qbase + (size_t)2 * H * dh;
  const bf16* dobase = a.d_o + (size_t)b * T * ldo + (size_t)hd * dh;

  const uint32_t tmem = attn_prologue(ctl, a.tmem_cols);
  const uint32_t tS = tmem, tdP = tmem + (uint32_t)BQ, tdV = tmem + 2u * (uint32_t)BQ, tdK = tdV + DHP;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  stage_tile<DHP>(sK, MT, kbase + (size_t)k0 * ld, ld, T - k0, dh);
  stage_tile<DHP>(sV, MT, vbase + (size_t)k0 * ld, ld, T - k0, dh);

  for (int blk = 0; blk < a.nblocks; ++blk) {
    const int i0 = blk * BQ;
    const int nvalid = min(BQ, T - i0);
    stage_tile<DHP>(sQ, BQ, qbase + (size_t)i0 * ld, ld, nvalid, dh);
    stage_tile<DHP>(sdO, BQ, dobase + (size_t)i0 * ldo, ldo, nvalid, dh);
    for (int i = tid; i < BQ; i += ATT_THREADS) {
      const bool ok = i < nvalid;
      s_lse[i] = ok ? a.lse[(size_t)bh * T + i0 + i] * 1.4426950408889634f : INFINITY;
      s_delta[i] = ok ? a.delta[(size_t)bh * T + i0 + i] : 0.f;
    }
    cp_async_wait_all();
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tS, sK, gsKV, sQ, gsQ, false, BQ, DHP / 16, false);
      issue_mma(tdP, sV, gsKV, sdO, gsQ, false, BQ, DHP / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c0 = 0; c0 < BQ; c0 += 16) {
      float s[16], dp[16];
      tmem_ld16(tS + lane_off + c0, s);
      tmem_ld16(tdP + lane_off + c0, dp);
      tmem_ld_wait();
      uint32_t wp[8], wd[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = exp2_fast(fmaf(s[i], a.scale_log2, -s_lse[c0 + i]));       // padded query: exp2(-inf) = 0
        const float p1 = exp2_fast(fmaf(s[i + 1], a.scale_log2, -s_lse[c0 + i + 1]));
        wp[i / 2] = pack_bf16(p0, p1);
        wd[i / 2] = pack_bf16(p0 * (dp[i] - s_delta[c0 + i]), p1 * (dp[i + 1] - s_delta[c0 + i + 1]));
      }
      const size_t off = (size_t)(c0 / 8) * gsS + tid * 16;
      *reinterpret_cast<uint4*>(sPT_ptr + off) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
      *reinterpret_cast<uint4*>(sPT_ptr + off + gsS) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
      *reinterpret_cast<uint4*>(sdST_ptr + off) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      *reinterpret_cast<uint4*>(sdST_ptr + off + gsS) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    }
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tdV, sPT, gsS, sdO, gsQ, true, DHP, BQ / 16, blk > 0);
      issue_mma(tdK, sdST, gsS, sQ, gsQ, true, DHP, BQ / 16, blk > 0);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  const int key = k0 + tid;
  bf16* dkout = a.dqkv + ((size_t)b * T + min(key, T - 1)) * ld + (size_t)H * dh + (size_t)hd * dh;
  bf16* dvout = dkout + (size_t)H * dh;
#pragma unroll
  for (int c0 = 0; c0 < DHP; c0 += 16) {
    float vk[16], vv[16];
    tmem_ld16(tdK + lane_off + c0, vk);
    tmem_ld16(tdV + lane_off + c0, vv);
    tmem_ld_wait();
    if (key < T) {
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (c0 + 8 * h8 < dh) {
          uint32_t wk[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            wk[i] = pack_bf16(vk[8 * h8 + 2 * i] * a.scale, vk[8 * h8 + 2 * i + 1] * a.scale);
            wv[i] = pack_bf16(vv[8 * h8 + 2 * i], vv[8 * h8 + 2 * i + 1]);
          }
          *reinterpret_cast<uint4*>(dkout + c0 + 8 * h8) = make_uint4(wk[0], wk[1], wk[2], wk[3]);
          *reinterpret_cast<uint4*>(dvout + c0 + 8 * h8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
    }
  }
  attn_epilogue(tmem, a.tmem_cols);
}


// ------------------------------------------------------------------------------------------ fused backward
// Short sequences (T <= 160: every key fits one MMA N extent): ONE CTA per (sample, head) computes dQ,
// dK and dV with S and P evaluated once.  256 threads: warps w and w + 4 own the same 32 rows (TMEM lane
// quadrant w % 4) and split the key columns.  Per 128-query tile:
//   S = Q K^T -> TMEM          p = exp2(s c - lse) -> bf16 P in smem [q][key]
//   dP = dO V^T -> TMEM (over S)   dV^T += dO^T P      (A = dO read MN-major, B = P read MN-major)
//   dS = p (dP - delta) -> over P in smem
//   dQ = dS K -> TMEM (over dP)    dK^T += Q^T dS      (A = Q read MN-major, B = dS read MN-major)
// dK^T / dV^T live in TMEM as [d (lane)][key (column)] across the query tiles and are written at the end;
// the transposed products are what lets one [q][key] tile of P / dS feed both contractions.
constexpr int FUSED_THREADS = 256;

template <int DHP>
__global__ void __launch_bounds__(FUSED_THREADS) attn_bwd_fused_umma_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                          const __grid_constant__ CUtensorMap tmKV,
                                                                          const __grid_constant__ CUtensorMap tmdO,
                                                                          const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align128(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  float* s_delta = reinterpret_cast<float*>(smem + 128);  // [128]
  float* s_lse = s_delta + MT;                            // [128], log2 units
  const int NK = a.BN, T = a.T, H = a.H, dh = a.dh;
  uint8_t* tiles = smem + 128 + 1024;
  const uint32_t sQ = smem_u32(tiles);
  const uint32_t sdO = sQ + g8_alloc(MT, DHP);
  const uint32_t sK = sdO + g8_alloc(MT, DHP);
  const uint32_t sV = sK + g8_alloc(NK, DHP);
  const uint32_t sP = sV + g8_alloc(NK, DHP);
  uint8_t* sP_ptr = tiles + 2 * g8_alloc(MT, DHP) + 2 * g8_alloc(NK, DHP);
  const bool tma = a.use_tma != 0;
  // TMA writes tiles densely (group stride rows * 16); the cp.async path pads the stride by 16 bytes
  const uint32_t gsKV = tma ? (uint32_t)NK * 16u : g8_stride(NK), gsP = g8_stride(MT);
  uint32_t ld_phase = 0;

  const int bh = blockIdx.x, b = bh / H, hd = bh % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int r = quad * 32 + lane;  // row of the M tile = TMEM lane
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const bf16* qbase = a.qkv + (size_t)b * T * ld + (size_t)hd * dh;
  const bf16* kbase = qbase + (size_t)H * dh;
  const bf16* vbase = qbase + (size_t)2 * H * dh;
  const bf16* dobase = a.d_o + (size_t)b * T * ldo + (size_t)hd * dh;
  const bf16* obase = a.o + (size_t)b * T * ldo + (size_t)hd * dh;

  // prologue (256 threads): barrier + TMEM
  if (tid == 0) { mbar_init(&ctl->bar, 1); mbar_init(&ctl->ld_bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_dyn(&ctl->tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctl->tmem_slot;
  const int R = NK > DHP ? NK : DHP;  // S / dP / dQ share the first R columns
  const uint32_t tS = tmem, tdV = tmem + (uint32_t)R, tdK = tdV + (uint32_t)NK;
  const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
  uint32_t phase = 0;

  // key-column chunks (16 wide) of this warp half
  const int nchunks = NK / 16;
  const int ch0 = half == 0 ? 0 : (nchunks + 1) / 2;
  const int ch1 = half == 0 ? (nchunks + 1) / 2 : nchunks;

  ALap L(a.dbg);
  if (!tma) {
    stage_tile<DHP, FUSED_THREADS>(sK, NK, kbase, ld, T, dh);
    stage_tile<DHP, FUSED_THREADS>(sV, NK, vbase, ld, T, dh);
  }

  const int ntiles = (T + MT - 1) / MT;
  for (int qt = 0; qt < ntiles; ++qt) {
    const int q0 = qt * MT;
    const int nq = min(MT, T - q0);
    const int kq = (nq + 15) / 16;            // K steps of the contractions over the query rows
    const bool warp_rows = quad * 32 < kq * 16;  // this warp's rows take part in those contractions
    // full tiles (and K / V, once) arrive by TMA; a short tail tile is staged by cp.async, only the rows of
    // its K steps
    const bool tile_tma = tma && nq == MT;
    const uint32_t gsQ = tile_tma ? MT * 16u : g8_stride(MT);
    if (tid == 0 && (tile_tma || (tma && qt == 0))) {
      const uint32_t q_bytes = (uint32_t)MT * DHP * 2, kv_bytes = (uint32_t)NK * DHP * 2;
      mbar_expect_tx(&ctl->ld_bar, (tile_tma ? 2 * q_bytes : 0u) + (qt == 0 ? 2 * kv_bytes : 0u));
      if (qt == 0) {
        tma_load_5d(tiles + 2 * g8_alloc(MT, DHP), &tmKV, &ctl->ld_bar, 0, 0, 0, H + hd, b);
        tma_load_5d(tiles + 2 * g8_alloc(MT, DHP) + g8_alloc(NK, DHP), &tmKV, &ctl->ld_bar, 0, 0, 0, 2 * H + hd, b);
      }
      if (tile_tma) {
        tma_load_5d(tiles, &tmQ, &ctl->ld_bar, 0, q0, 0, hd, b);
        tma_load_5d(tiles + g8_alloc(MT, DHP), &tmdO, &ctl->ld_bar, 0, q0, 0, hd, b);
      }
    }
    if (!tile_tma) {
      stage_tile<DHP, FUSED_THREADS>(sQ, MT, qbase + (size_t)q0 * ld, ld, nq, dh, kq * 16);
      stage_tile<DHP, FUSED_THREADS>(sdO, MT, dobase + (size_t)q0 * ldo, ldo, nq, dh, kq * 16);
    }
    if (half == 0) {  // row statistics: delta = sum_d dO * O, lse in log2 units
      float delta = 0.f, lse2 = INFINITY;  // padded query row: p = exp2(-inf) = 0
      const int q = q0 + r;
      if (q < T) {
        const bf16* dor = dobase + (size_t)q * ldo;
        const bf16* orow = obase + (size_t)q * ldo;
        // every load of the row issued before the first use (the loop would otherwise serialise 2 * dh / 8
        // L2 round trips)
        uint4 xv[DHP / 8], yv[DHP / 8];
#pragma unroll
        for (int c = 0; c < DHP / 8; ++c) {
          const bool in = c * 8 < dh;
          xv[c] = in ? *reinterpret_cast<const uint4*>(dor + c * 8) : make_uint4(0, 0, 0, 0);
          yv[c] = in ? *reinterpret_cast<const uint4*>(orow + c * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < DHP / 8; ++c) {
          const uint32_t xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, ys[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[i]));
            const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[i]));
            delta = fmaf(fx.x, fy.x, delta);
            delta = fmaf(fx.y, fy.y, delta);
          }
        }
        lse2 = a.lse[(size_t)bh * T + q] * 1.4426950408889634f;
      }
      s_delta[r] = delta;
      s_lse[r] = lse2;
    }
    L.lap(0);
    cp_async_wait_all();
    if (tile_tma || (tma && qt == 0)) { mbar_wait(&ctl->ld_bar, ld_phase); ld_phase ^= 1; }
    publish_smem_and_sync();
    L.lap(1);
    if (tid == 0) {
      issue_mma(tS, sQ, gsQ, sK, gsKV, false, NK, DHP / 16, false);
      umma_commit(&ctl->bar);
    }
    const float delta = s_delta[r], lse2 = s_lse[r];
    const bool row_ok = q0 + r < T;
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(2);
    // ---- P = exp2(s c - lse) (zero for padded rows / keys) -> smem [q][key]
    if (warp_rows) {
      for (int ch = ch0; ch < ch1; ++ch) {
        const int c0 = ch * 16;
        float sv[16];
        tmem_ld16(tS + lane_off + c0, sv);
        tmem_ld_wait();
        uint32_t w[8];
        if (c0 + 16 <= T) {  // no padded key in this chunk; a padded row has lse2 = +inf -> p = 0
#pragma unroll
          for (int i = 0; i < 16; i += 2)
            w[i / 2] = pack_bf16(exp2_fast(fmaf(sv[i], a.scale_log2, -lse2)), exp2_fast(fmaf(sv[i + 1], a.scale_log2, -lse2)));
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = c0 + i < T ? exp2_fast(fmaf(sv[i], a.scale_log2, -lse2)) : 0.f;
            const float p1 = c0 + i + 1 < T ? exp2_fast(fmaf(sv[i + 1], a.scale_log2, -lse2)) : 0.f;
            w[i / 2] = pack_bf16(p0, p1);
          }
        }
        uint8_t* dst = sP_ptr + (size_t)(c0 / 8) * gsP + r * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + gsP) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    L.lap(3);
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tS, sdO, gsQ, sV, gsKV, false, NK, DHP / 16, false);  // dP over S
      // dV^T[d][key] += sum_q dO[q][d] P[q][key]: both operands MN-major (rows = contraction index q)
      const uint32_t idesc = make_idesc_bf16(MT, NK, true, true);
      for (int ks = 0; ks < kq; ++ks)
        umma_bf16(tdV, desc_ns(sdO + ks * 256, 128, gsQ), desc_ns(sP + ks * 256, 128, gsP), idesc,
                  (qt > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(4);
    // ---- dS = p (dP - delta), in place over P
    if (warp_rows) {
      for (int ch = ch0; ch < ch1; ++ch) {
        const int c0 = ch * 16;
        float dp[16];
        tmem_ld16(tS + lane_off + c0, dp);
        tmem_ld_wait();
        uint8_t* dst = sP_ptr + (size_t)(c0 / 8) * gsP + r * 16;
        const uint4 pa = *reinterpret_cast<const uint4*>(dst), pb = *reinterpret_cast<const uint4*>(dst + gsP);
        const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[i]));
          w[i] = pack_bf16(pf.x * (dp[2 * i] - delta), pf.y * (dp[2 * i + 1] - delta));
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(dst + gsP) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    } else {
      tmem_ld_wait();
    }
    L.lap(5);
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma(tS, sP, gsP, sK, gsKV, true, DHP, NK / 16, false);  // dQ over dP
      const uint32_t idesc = make_idesc_bf16(MT, NK, true, true);
      for (int ks = 0; ks < kq; ++ks)  // dK^T[d][key] += sum_q Q[q][d] dS[q][key]
        umma_bf16(tdK, desc_ns(sQ + ks * 256, 128, gsQ), desc_ns(sP + ks * 256, 128, gsP), idesc,
                  (qt > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(6);
    // ---- dQ rows out: the two warp halves split the head dimension in 16-column chunks
    {
      constexpr int NDC = DHP / 16;
      const int d0 = half == 0 ? 0 : (NDC + 1) / 2, d1 = half == 0 ? (NDC + 1) / 2 : NDC;
      bf16* out = a.dqkv + ((size_t)b * T + min(q0 + r, T - 1)) * ld + (size_t)hd * dh;
      for (int dc = d0; dc < d1; ++dc) {
        float v[16];
        tmem_ld16(tS + lane_off + dc * 16, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (dc * 16 + 8 * h8 < dh) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                w[i] = pack_bf16(v[8 * h8 + 2 * i] * a.scale, v[8 * h8 + 2 * i + 1] * a.scale);
              *reinterpret_cast<uint4*>(out + dc * 16 + 8 * h8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
    }
    // the next tile overwrites sQ / sdO / sP and the S region: all MMAs have completed (waited above)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    L.lap(7);
  }

  // ---- dK^T, dV^T out: thread = head dim d (TMEM lane), columns = keys; a warp writes 32 consecutive d
  // of one key (64 contiguous bytes)
  if (quad * 32 < dh) {
    // lane pairs exchange one value per key pair so that every lane stores two consecutive head dims of ONE
    // key as a 4-byte word: even lane -> (d, d+1) of the even key, odd lane -> (d-1, d) of the odd key
    const int dd = r;
    const bool odd = (lane & 1) != 0;
    const int dcol = odd ? dd - 1 : dd;
    bf16* base = a.dqkv + (size_t)b * T * ld + (size_t)hd * dh + dcol;
    for (int ch = ch0; ch < ch1; ++ch) {
      const int c0 = ch * 16;
      float vk[16], vv[16];
      tmem_ld16(tdK + lane_off + c0, vk);
      tmem_ld16(tdV + lane_off + c0, vv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float rk = __shfl_xor_sync(0xffffffffu, odd ? vk[i] : vk[i + 1], 1);
        const float rv = __shfl_xor_sync(0xffffffffu, odd ? vv[i] : vv[i + 1], 1);
        const int key = c0 + i + (odd ? 1 : 0);
        if (dcol + 1 < dh + 1 && dcol < dh && key < T) {
          const float k_lo = odd ? rk : vk[i], k_hi = odd ? vk[i + 1] : rk;
          const float v_lo = odd ? rv : vv[i], v_hi = odd ? vv[i + 1] : rv;
          bf16* dst = base + (size_t)key * ld + (size_t)H * dh;
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16(k_lo * a.scale, k_hi * a.scale);
          *reinterpret_cast<uint32_t*>(dst + (size_t)H * dh) = pack_bf16(v_lo, v_hi);
        }
      }
    }
  }
  L.lap(8);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem, a.tmem_cols);
  }
  L.lap(9);
  L.flush();
}

// ------------------------------------------------------------------------------------------ host side
int pick_dhp(int dh) {
  if (dh <= 32) return 32;
  if (dh <= 64) return 64;
  if (dh <= 80) return 80;
  if (dh <= 128) return 128;
  return 0;
}
uint32_t pow2_cols(int cols) {
  uint32_t c = 32;
  while ((int)c < cols) c <<= 1;
  return c;
}
// block length: the fewest blocks of at most `cap` rows, rows rounded up to a multiple of 16
void pick_block(int T, int cap, int* bn, int* nblocks) {
  const int nb = (int)ceil_div(T, cap);
  *nblocks = nb;
  *bn = (int)ceil_div(ceil_div(T, nb), 16) * 16;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  V4H_REQUIRE(bytes <= 227 * 1024, "attention: %zu bytes of shared memory exceed the 227 KB per CTA", bytes);
  V4H_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return V4H_OK;
}

// 5-d tensor map over a (B, T, nh, dh) bf16 tensor with row pitch `ld` elements between tokens: dims
// (8, T, dh/8, nh, B), box (8, rows, dh/8, 1, 1): the box lands in shared memory as [dh/8][rows][8] = G8.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_g8_map(const void* base, int B, int T, int nh, int dh, size_t ld, int rows, CUtensorMap* out) {
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  if (!encode) return fail(V4H_ERR_CUDA, "attention: cuTensorMapEncodeTiled is not available from the driver");
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int, int, int, size_t, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(base, B, T, nh, dh, ld, rows);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return V4H_OK; }
  cuuint64_t dims[5] = {8, (cuuint64_t)T, (cuuint64_t)(dh / 8), (cuuint64_t)nh, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, 16, (cuuint64_t)dh * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)rows, (cuuint32_t)(dh / 8), 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap m;
  const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(V4H_ERR_CUDA, "attention: cuTensorMapEncodeTiled (5-d G8 view) failed (%d)", (int)r);
  if (cache.size() > 1024) cache.clear();
  cache[key] = m;
  *out = m;
  return V4H_OK;
}
bool attn_tma_enabled() {
  static const int on = [] { const char* e = getenv("V4H_ATTN_TMA"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <int DHP>
int fwd_launch(AttnArgs a, int B, cudaStream_t s) {
  const int cap = std::min(160, 512 - DHP) / 16 * 16;
  pick_block(a.T, cap, &a.BN, &a.nblocks);
  a.tmem_cols = pow2_cols(a.BN + DHP);
  const size_t smem = 256 + g8_alloc(MT, DHP) + 2 * g8_alloc(a.BN, DHP) + g8_alloc(MT, a.BN);
  static size_t configured = 0;
  if (smem > configured) { V4H_TRY(set_smem(attn_fwd_umma_kernel<DHP>, smem)); configured = smem; }
  CUtensorMap mq, mkv;
  memset(&mq, 0, sizeof(mq)); memset(&mkv, 0, sizeof(mkv));
  a.use_tma = (attn_tma_enabled() && a.dh == DHP) ? 1 : 0;  // the box must cover the padded head dim exactly
  if (a.use_tma) {
    const size_t ld = (size_t)3 * a.H * a.dh;
    V4H_TRY(make_g8_map(a.qkv, B, a.T, 3 * a.H, a.dh, ld, MT, &mq));
    V4H_TRY(make_g8_map(a.qkv, B, a.T, 3 * a.H, a.dh, ld, a.BN, &mkv));
  }
  dim3 grid((unsigned)ceil_div(a.T, MT), (unsigned)(B * a.H));
  V4H_CUDA(launch_pdl(attn_fwd_umma_kernel<DHP>, dim3(grid), dim3(ATT_THREADS), smem, s, mq, mkv, a));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

bool fused_bwd_enabled() {
  static const int on = [] { const char* e = getenv("V4H_ATTN_FUSED_BWD"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <int DHP>
int bwd_launch(AttnArgs a, int B, cudaStream_t s) {
  if (a.T <= 160 && fused_bwd_enabled()) {  // one CTA per (sample, head): S and P evaluated once
    AttnArgs f = a;
    f.BN = (int)ceil_div(a.T, 16) * 16;
    f.nblocks = 1;
    const int R = f.BN > DHP ? f.BN : DHP;
    f.tmem_cols = pow2_cols(R + 2 * f.BN);
    // the MN-major reads of Q / dO span 128 "M" columns = 16 column groups even when DHP < 128: the
    // groups past DHP land in the tiles that follow (garbage rows of dK^T / dV^T that are never stored)
    const size_t smem = 256 + 1024 + 2 * g8_alloc(MT, DHP) + 2 * g8_alloc(f.BN, DHP) + g8_alloc(MT, f.BN) +
                        (DHP < 128 ? 16 * g8_stride(MT) : 0);
    static size_t configured = 0;
    if (smem > configured) { V4H_TRY(set_smem(attn_bwd_fused_umma_kernel<DHP>, smem)); configured = smem; }
    CUtensorMap mq, mkv, mdo;
    memset(&mq, 0, sizeof(mq)); memset(&mkv, 0, sizeof(mkv)); memset(&mdo, 0, sizeof(mdo));
    f.use_tma = (attn_tma_enabled() && a.dh == DHP) ? 1 : 0;
    if (f.use_tma) {
      const size_t ldq = (size_t)3 * a.H * a.dh;
      V4H_TRY(make_g8_map(a.qkv, B, a.T, 3 * a.H, a.dh, ldq, MT, &mq));
      V4H_TRY(make_g8_map(a.qkv, B, a.T, 3 * a.H, a.dh, ldq, f.BN, &mkv));
      V4H_TRY(make_g8_map(a.d_o, B, a.T, a.H, a.dh, (size_t)a.H * a.dh, MT, &mdo));
    }
    V4H_CUDA(launch_pdl(attn_bwd_fused_umma_kernel<DHP>, dim3((unsigned)(B * a.H)), dim3(FUSED_THREADS), smem, s, mq, mkv, mdo, f));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  dim3 grid((unsigned)ceil_div(a.T, MT), (unsigned)(B * a.H));
  {  // dQ + delta
    AttnArgs q = a;
    const int cap = std::min(160, (512 - DHP) / 2) / 16 * 16;
    pick_block(a.T, cap, &q.BN, &q.nblocks);
    q.tmem_cols = pow2_cols(2 * q.BN + DHP);
    const size_t smem = 256 + 2 * g8_bytes(MT, DHP) + 2 * g8_bytes(q.BN, DHP) + g8_bytes(MT, q.BN);
    static size_t configured = 0;
    if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dq_umma_kernel<DHP>, smem)); configured = smem; }
    V4H_CUDA(launch_pdl(attn_bwd_dq_umma_kernel<DHP>, dim3(grid), dim3(ATT_THREADS), smem, s, q));
    V4H_LAUNCH_CHECK();
  }
  {  // dK, dV
    AttnArgs k = a;
    const int cap = std::min(160, (512 - 2 * DHP) / 2) / 16 * 16;
    pick_block(a.T, cap, &k.BN, &k.nblocks);
    k.tmem_cols = pow2_cols(2 * k.BN + 2 * DHP);
    const size_t smem = 256 + 2048 + 2 * g8_bytes(MT, DHP) + 2 * g8_bytes(k.BN, DHP) + 2 * g8_bytes(MT, k.BN);
    static size_t configured = 0;
    if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dkv_umma_kernel<DHP>, smem)); configured = smem; }
    V4H_CUDA(launch_pdl(attn_bwd_dkv_umma_kernel<DHP>, dim3(grid), dim3(ATT_THREADS), smem, s, k));
    V4H_LAUNCH_CHECK();
  }
  return V4H_OK;
}

int check_args(const AttnArgs& a, int B) {
  V4H_REQUIRE(B > 0 && a.T > 0 && a.H > 0, "attention: empty problem");
  V4H_REQUIRE(B * a.H <= 65535, "attention: batch*heads %d exceeds 65535", B * a.H);
  V4H_REQUIRE(attention_umma_supported(a.dh), "attention: head_dim %d is not supported by the tcgen05 kernels", a.dh);
  V4H_REQUIRE((reinterpret_cast<uintptr_t>(a.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0,
              "attention: qkv / o must be 16-byte aligned");
  return V4H_OK;
}

}  // namespace

static long long* g_attn_dbg = nullptr;
void attention_debug_counters(long long* dev_counters) { g_attn_dbg = dev_counters; }

bool attention_umma_supported(int dh) { return dh >= 8 && dh % 8 == 0 && dh <= 128; }

int attention_fwd_umma(const bf16* qkv, bf16* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s) {
  AttnArgs a{};
  a.qkv = qkv; a.o = o; a.lse = lse; a.T = Tn; a.H = H; a.dh = dh;
  a.dbg = g_attn_dbg;
  a.scale = 1.f / sqrtf((float)dh);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  V4H_TRY(check_args(a, B));
  switch (pick_dhp(dh)) {
    case 32: return fwd_launch<32>(a, B, s);
    case 64: return fwd_launch<64>(a, B, s);
    case 80: return fwd_launch<80>(a, B, s);
    default: return fwd_launch<128>(a, B, s);
  }
}

int attention_bwd_umma(const bf16* qkv, const bf16* o, const float* lse, const bf16* d_o, float* delta, bf16* dqkv,
                       int B, int Tn, int H, int dh, cudaStream_t s) {
  AttnArgs a{};
  a.qkv = qkv; a.o = const_cast<bf16*>(o); a.lse = const_cast<float*>(lse); a.d_o = d_o; a.delta = delta;
  a.dqkv = dqkv; a.T = Tn; a.H = H; a.dh = dh;
  a.dbg = g_attn_dbg;
  a.scale = 1.f / sqrtf((float)dh);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  V4H_TRY(check_args(a, B));
  V4H_REQUIRE((reinterpret_cast<uintptr_t>(d_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 && delta,
              "attention: d_o / dqkv must be 16-byte aligned and delta non-null");
  switch (pick_dhp(dh)) {
    case 32: return bwd_launch<32>(a, B, s);
    case 64: return bwd_launch<64>(a, B, s);
    case 80: return bwd_launch<80>(a, B, s);
    default: return bwd_launch<128>(a, B, s);
  }
}

}  // namespace v4h
