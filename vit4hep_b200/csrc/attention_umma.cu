// tcgen05 / TMEM attention forward and backward for sm_100a (bf16 operands, fp32 accumulation and
// softmax): softmax(q k^T / sqrt(dh)) v per (sample, head), no mask, no dropout
// (reference nn/vit.py:425-451 with the xformers memory_efficient_attention call it makes).
//
// Layouts are the ones the qkv Linear produces (reference nn/vit.py:427): qkv (B, T, 3, H, dh),
// o / d_o (B, T, H, dh), lse and delta (B, H, T) fp32, lse = log sum_j exp(s_ij) of the scaled scores.
//
// Two shared-memory tile formats:
//  - "W" tiles for everything TMA loads (Q, K, V, dO, O): ceil(dh / 64) boxes of [rows][64 columns], 128-byte
//    swizzle (see w_boxes / desc_kmajor / desc_mnmajor below).
//  - "G8" tiles for what threads write (P, dS, P^T, dS^T): element
//    (row, col) of a [ROWS x COLS] bf16 tile sits at  (col / 8) * gstride + row * 16 + (col % 8) * 2  bytes,
//    i.e. 8x8 core matrices of 128 contiguous bytes, the non-swizzled canonical UMMA layout; gstride carries
//    one 16-byte pad so that the 16-byte st.shared writes of a warp spread over all banks.
// Either format is read
//   - as a K-major operand  (rows = M/N index, cols = K index)
//   - as an MN-major operand (rows = K index, cols = M/N index)
// so Q, K, V, dO are staged once and serve both Q K^T-like and P V-like products, and one [q][key] tile of
// P / dS feeds dV^T = dO^T P and dK^T = Q^T dS without a transposed copy.
// Head dims that are not multiples of 64 (dh = 80 here) cost nothing: K steps are 16 wide.
//
// Kernels (thread = one TMEM lane = one row of the 128-row M tile):
//   fwd   : CTA per (128 queries, sample-head), 128 threads; loops over key blocks with an online softmax
//           (one block, registers only, for T <= 160): S = Q K^T -> TMEM, p = exp2(..) -> bf16 P in smem,
//           O += P V through TMEM
//   fused : T <= 160: ONE CTA (256 threads) per (sample, head) computes dQ, dK, dV in one pass
//   dq    : CTA per (128 queries, sample-head); S and dP = dO V^T in TMEM, dS -> smem, dQ += dS K
//           accumulated in TMEM over the key blocks; also emits delta = rowsum(dO * O)
//   dkv   : CTA per (128 keys, sample-head); S^T = K Q^T and dP^T = V dO^T in TMEM, P^T / dS^T -> smem,
//           dV += P^T dO, dK += dS^T Q accumulated in TMEM over the query blocks
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

struct AttnSmem {
  uint64_t bar;      // MMA completion
  uint64_t ld_bar;   // TMA operand loads
  uint64_t v_bar;    // forward: the V tile (second TMA barrier)
  uint32_t tmem_slot;
};

// ---- "W" tiles: what TMA writes with the 128-byte swizzle.  A ROWS x DHP bf16 tile is ceil(DHP / 64) boxes
// of [ROWS][64] (128-byte rows; the 16-byte chunk c of row r sits at chunk (c ^ (r & 7))), ROWS * 128 bytes
// apart, 1024-byte aligned.  One box row is one TMA "piece", so a tile moves in 2 pieces per row where the
// G8 view needs dh / 8 -- the 16-byte pieces of the G8 loads ran at 12-14 B/clk/SM and dominated the forward
// and the fused backward.  The same bytes serve as
//   - K-major operand  (rows = M/N index, K step ks = columns [16 ks, 16 ks + 16)): box ks / 4, +32 bytes per step
//   - MN-major operand (rows = K index, 16 per step; columns = M/N index in 64-wide chunks): LBO = box stride
// The tiles P / dS that threads write stay G8 (row per thread, conflict-free 16-byte stores).
__host__ __device__ constexpr int w_boxes(int dhp) { return (dhp + 63) / 64; }
__host__ __device__ constexpr uint32_t w_box(int rows) { return (uint32_t)rows * 128u; }
__host__ __device__ constexpr uint32_t w_bytes(int rows, int dhp) { return (uint32_t)w_boxes(dhp) * w_box(rows); }

struct Opnd {       // one shared-memory operand tile
  uint32_t base;    // shared address
  uint32_t stride;  // W: box stride (rows * 128); G8: group stride
  bool w;
};
__device__ __forceinline__ Opnd opnd_w(uint32_t base, int rows) { return Opnd{base, w_box(rows), true}; }
__device__ __forceinline__ Opnd opnd_g8(uint32_t base, uint32_t gstride) { return Opnd{base, gstride, false}; }
// descriptor of K step `ks` when the tile's COLUMNS are the contraction index
__device__ __forceinline__ uint64_t desc_kmajor(const Opnd& o, int ks) {
  return o.w ? make_smem_desc(o.base + (uint32_t)(ks >> 2) * o.stride + (uint32_t)(ks & 3) * 32u, 0, 1024)
             : desc_ns(o.base + 2u * ks * o.stride, o.stride, 128);
}
// descriptor of K step `ks` when the tile's ROWS are the contraction index
__device__ __forceinline__ uint64_t desc_mnmajor(const Opnd& o, int ks) {
  return o.w ? make_smem_desc(o.base + (uint32_t)ks * 2048u, o.stride, 1024) : desc_ns(o.base + (uint32_t)ks * 256u, 128, o.stride);
}
// D[tmem 128 x n] (+)= A B over `ksteps` steps of 16; *_mn: that operand is read MN-major.  One thread.
__device__ __forceinline__ void issue_mma_x(uint32_t d_tmem, const Opnd& A, bool a_mn, const Opnd& B, bool b_mn, int n,
                                            int ksteps, bool accumulate_first) {
  const uint32_t idesc = make_idesc_bf16(MT, n, a_mn, b_mn);
  for (int ks = 0; ks < ksteps; ++ks)
    umma_bf16(d_tmem, a_mn ? desc_mnmajor(A, ks) : desc_kmajor(A, ks), b_mn ? desc_mnmajor(B, ks) : desc_kmajor(B, ks),
              idesc, (accumulate_first || ks > 0) ? 1u : 0u);
}

// explicit shared-window accesses: the tile pointers come out of an integer round trip (alignment), so plain
// dereferences compile to generic ST.E / LD.E with 64-bit address arithmetic per access
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("{ .reg .b16 h; cvt.u16.u32 h, %1; st.shared.b16 [%0], h; }" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// cp.async fallback of the TMA load of a W tile (head dims the box cannot cover, TMA switched off): rows >=
// rows_valid and columns >= dh are zero-filled; only the first `load_rows` rows are touched
template <int DHP, int NTHREADS>
__device__ __forceinline__ void stage_tile_w(uint32_t dst, int rows, const bf16* __restrict__ src, size_t ld,
                                             int rows_valid, int dh, int load_rows) {
  constexpr int NCG = DHP / 8;
  constexpr int RPI = NTHREADS / NCG;
  if ((int)threadIdx.x >= RPI * NCG) return;
  const int r0 = (int)threadIdx.x / NCG, cg = (int)threadIdx.x - r0 * NCG;
  const bool col_ok = cg * 8 < dh;
  const bf16* p = src + (size_t)r0 * ld + cg * 8;
  const uint32_t d = dst + (uint32_t)(cg >> 3) * w_box(rows);
  for (int row = r0; row < load_rows; row += RPI, p += (size_t)RPI * ld) {
    const bool ok = col_ok && row < rows_valid;
    cp_async16(d + (uint32_t)row * 128u + (uint32_t)(((cg & 7) ^ (row & 7)) << 4), ok ? (const void*)p : (const void*)src, ok);
  }
}
// TMA load of a W tile: box b covers columns [col0 + 64 b, +64) of rows [row0, row0 + rows) of sample `batch`
template <int DHP>
__device__ __forceinline__ void tma_load_w(uint8_t* dst, int rows, const CUtensorMap* map, uint64_t* bar, int col0,
                                           int row0, int batch) {
#pragma unroll
  for (int b = 0; b < w_boxes(DHP); ++b) tma_load_3d(dst + b * w_box(rows), map, bar, col0 + 64 * b, row0, batch);
}
// the same boxes, only as far as L2
template <int DHP>
__device__ __forceinline__ void tma_prefetch_w(const CUtensorMap* map, int col0, int row0, int batch) {
#pragma unroll
  for (int b = 0; b < w_boxes(DHP); ++b)
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(col0 + 64 * b), "r"(row0), "r"(batch)
                 : "memory");
}
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}


__device__ __forceinline__ uint8_t* align128(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127));
}

// common prologue: barrier init + TMEM allocation; returns the TMEM base address
__device__ __forceinline__ uint32_t attn_prologue(AttnSmem* ctl, uint32_t tmem_cols) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&ctl->bar, 1);
    mbar_init(&ctl->ld_bar, 1);
    mbar_init(&ctl->v_bar, 1);
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
  int use_tma;       // operand tiles staged by TMA instead of cp.async
  int p_off;         // forward: byte offset of the P tile from the first operand tile
  // fused backward: byte offsets of the tiles (Q, dO, K, V, P, tail Q, tail dO, tail P) and the TMEM column of
  // the tail tile's dQ accumulator
  int off[8];
  uint32_t dq2_col;
  int pf_stride;     // CTAs resident at a time (L2 prefetch distance); 0 = no prefetch
  int stages;        // pipelined long-T kernels: shared-memory stages of the streamed operand pair (2 or 3)
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
// tmQ / tmKV: 3-d views (3 H dh columns, T rows, B) of qkv with 128-byte-swizzled boxes of [128 | BN rows][64
// columns].  ONE = every key fits one block (T <= 160): the score row stays in registers between the
// maximum and the exponentials (one TMEM read), O is read once at the end and P may overlay Q / K.
template <int DHP, bool ONE>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmKV,
                                                                    const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  uint8_t* tiles = smem + 1024;
  const int BN = a.BN, T = a.T, H = a.H, dh = a.dh;
  uint8_t* pQ = tiles;
  uint8_t* pK = pQ + w_bytes(MT, DHP);
  uint8_t* pV = pK + w_bytes(BN, DHP);
  uint8_t* sP_ptr = tiles + a.p_off;
  const uint32_t gsP = g8_stride(MT);
  const Opnd oQ = opnd_w(smem_u32(pQ), MT), oK = opnd_w(smem_u32(pK), BN), oV = opnd_w(smem_u32(pV), BN);
  const Opnd oP = opnd_g8(smem_u32(sP_ptr), gsP);
  const bool tma = a.use_tma != 0;
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
  const int q_rows = min(MT, (T - q0 + 15) / 16 * 16);  // rows of the Q tile any stored result depends on

  float o_acc[ONE ? 1 : DHP];
#pragma unroll
  for (int i = 0; i < (ONE ? 1 : DHP); ++i) o_acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const bool warp_active = q0 + warp * 32 < T;  // any valid query row in this warp

  for (int blk = 0; blk < a.nblocks; ++blk) {
    const int n0 = blk * BN;
    const int nvalid = min(BN, T - n0);
    // Q and K arrive on one barrier, V on a second one: S = Q K^T and the softmax run under the V load
    if (tma) {
      if (tid == 0) {
        mbar_expect_tx(&ctl->ld_bar, w_bytes(BN, DHP) + (blk == 0 ? w_bytes(MT, DHP) : 0u));
        if (blk == 0) tma_load_w<DHP>(pQ, MT, &tmQ, &ctl->ld_bar, hd * dh, q0, b);
        tma_load_w<DHP>(pK, BN, &tmKV, &ctl->ld_bar, (H + hd) * dh, n0, b);
        mbar_expect_tx(&ctl->v_bar, w_bytes(BN, DHP));
        tma_load_w<DHP>(pV, BN, &tmKV, &ctl->v_bar, (2 * H + hd) * dh, n0, b);
      }
      L.lap(1);
      mbar_wait(&ctl->ld_bar, ld_phase);
    } else {
      if (blk == 0) stage_tile_w<DHP, ATT_THREADS>(oQ.base, MT, qbase + (size_t)q0 * ld, ld, T - q0, dh, q_rows);
      stage_tile_w<DHP, ATT_THREADS>(oK.base, BN, kbase + (size_t)n0 * ld, ld, nvalid, dh, BN);
      cp_async_commit();
      stage_tile_w<DHP, ATT_THREADS>(oV.base, BN, vbase + (size_t)n0 * ld, ld, nvalid, dh, BN);
      cp_async_commit();
      L.lap(1);
      cp_async_wait_group<1>();
    }
    L.lap(2);
    publish_smem_and_sync();
    L.lap(3);
    if (tid == 0) {
      issue_mma_x(tS, oQ, false, oK, false, BN, DHP / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(4);

    // ---- softmax on this thread's row (warps whose 32 rows are all beyond T skip the work; their P rows
    // stay garbage, which only reaches their own, never stored, O rows)
    float corr = 0.f, lsum = 0.f;
    if (warp_active) {
      if (ONE) {
        constexpr int MAXCH = 10;  // BN <= 160
        float sv[MAXCH][16];
        const int nch = BN / 16;
#pragma unroll
        for (int c = 0; c < MAXCH; ++c)
          if (c < nch) tmem_ld16(tS + lane_off + c * 16, sv[c]);
        tmem_ld_wait();
        // padded keys (only in the last chunk) are pushed to -inf once: exp2 then gives their p = 0
        const int lastc = nch - 1;
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
          if (c == lastc) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sv[c][i] = c * 16 + i < nvalid ? sv[c][i] : -INFINITY;
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
          if (c < nch) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], sv[c][i]);
          }
        }
        const float m_new = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * a.scale_log2;
        m_run = m_new;
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
          if (c < nch) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = exp2_fast(fmaf(sv[c][i], a.scale_log2, -m_new));
              const float p1 = exp2_fast(fmaf(sv[c][i + 1], a.scale_log2, -m_new));
              ls4[(i >> 1) & 3] += p0 + p1;
              w[i / 2] = pack_bf16(p0, p1);
            }
            const uint32_t dst = oP.base + (uint32_t)(2 * c) * gsP + (uint32_t)tid * 16u;
            sts128(dst, w[0], w[1], w[2], w[3]);
            sts128(dst + gsP, w[4], w[5], w[6], w[7]);
          }
        }
        lsum = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      } else {
        // online softmax, pass 1 = row maximum
        float mx = -INFINITY;
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
          const uint32_t dst = oP.base + (uint32_t)(c0 / 8) * gsP + (uint32_t)tid * 16u;
          sts128(dst, w[0], w[1], w[2], w[3]);
          sts128(dst + gsP, w[4], w[5], w[6], w[7]);
        }
      }
    }
    l_run = l_run * corr + lsum;
    L.lap(5);
    if (tma) { mbar_wait(&ctl->v_bar, ld_phase); ld_phase ^= 1; }
    else cp_async_wait_group<0>();
    publish_smem_and_sync();
    L.lap(6);
    if (tid == 0) {
      issue_mma_x(tO, oP, false, oV, true, DHP, BN / 16, false);
      umma_commit(&ctl->bar);
    }
    mbar_wait(&ctl->bar, phase); phase ^= 1;
    tc_fence_after();
    L.lap(7);
    if (!ONE) {
      if (warp_active) {
#pragma unroll
        for (int c0 = 0; c0 < DHP; c0 += 16) {
          float v[16];
          tmem_ld16(tO + lane_off + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o_acc[ONE ? 0 : c0 + i] = fmaf(o_acc[ONE ? 0 : c0 + i], corr, v[i]);
        }
      }
      // the next iteration overwrites sK / sV / sP and the S / O accumulators: both MMAs have completed
      // (waited above); order this thread's TMEM reads before the next MMA issue
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  const int q = q0 + tid;
  if (ONE) {
    if (warp_active) {
      const float inv = 1.f / l_run;
      bf16* orow = a.o + ((size_t)b * T + min(q, T - 1)) * H * dh + (size_t)hd * dh;
      float v[DHP / 16][16];
#pragma unroll
      for (int c = 0; c < DHP / 16; ++c) tmem_ld16(tO + lane_off + c * 16, v[c]);
      tmem_ld_wait();
      if (q < T) {
#pragma unroll
        for (int c = 0; c < DHP / 8; ++c) {
          if (c * 8 < dh) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              w[i] = pack_bf16(v[c / 2][(c & 1) * 8 + 2 * i] * inv, v[c / 2][(c & 1) * 8 + 2 * i + 1] * inv);
            *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  } else if (q < T) {
    const float inv = 1.f / l_run;
    bf16* orow = a.o + ((size_t)b * T + q) * H * dh + (size_t)hd * dh;
#pragma unroll
    for (int c = 0; c < DHP; c += 8) {
      if (c < dh) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack_bf16(o_acc[ONE ? 0 : c + 2 * i] * inv, o_acc[ONE ? 0 : c + 2 * i + 1] * inv);
        *reinterpret_cast<uint4*>(orow + c) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  if (q < T) a.lse[(size_t)bh * T + q] = (m_run + log2f(l_run)) * 0.6931471805599453f;
  L.lap(8);
  attn_epilogue(tmem, a.tmem_cols);
  L.lap(9);
  L.flush();
}

// ------------------------------------------------------------------------------------------ dQ (+ delta)
template <int DHP>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_umma_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                       const __grid_constant__ CUtensorMap tmKV,
                                                                       const __grid_constant__ CUtensorMap tmdO,
                                                                       const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  uint8_t* tiles = smem + 1024;
  const int BN = a.BN, T = a.T, H = a.H, dh = a.dh;
  uint8_t* pQ = tiles;
  uint8_t* pdO = pQ + w_bytes(MT, DHP);
  uint8_t* pK = pdO + w_bytes(MT, DHP);
  uint8_t* pV = pK + w_bytes(BN, DHP);
  uint8_t* sdS_ptr = pV + w_bytes(BN, DHP);
  const uint32_t gsS = g8_stride(MT);
  const Opnd oQ = opnd_w(smem_u32(pQ), MT), odO = opnd_w(smem_u32(pdO), MT), oK = opnd_w(smem_u32(pK), BN),
             oV = opnd_w(smem_u32(pV), BN), odS = opnd_g8(smem_u32(sdS_ptr), gsS);
  const bool tma = a.use_tma != 0;
  uint32_t ld_phase = 0;

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

  // Q and dO of this tile and the first K / V block are requested before the row statistics are read
  auto load_kv = [&](int blk, bool with_q) {
    const int n0 = blk * BN;
    if (tma) {
      if (tid == 0) {
        mbar_expect_tx(&ctl->ld_bar, 2 * w_bytes(BN, DHP) + (with_q ? 2 * w_bytes(MT, DHP) : 0u));
        if (with_q) {
          tma_load_w<DHP>(pQ, MT, &tmQ, &ctl->ld_bar, hd * dh, q0, b);
          tma_load_w<DHP>(pdO, MT, &tmdO, &ctl->ld_bar, hd * dh, q0, b);
        }
        tma_load_w<DHP>(pK, BN, &tmKV, &ctl->ld_bar, (H + hd) * dh, n0, b);
        tma_load_w<DHP>(pV, BN, &tmKV, &ctl->ld_bar, (2 * H + hd) * dh, n0, b);
      }
    } else {
      const int nvalid = min(BN, T - n0);
      if (with_q) {
        stage_tile_w<DHP, ATT_THREADS>(oQ.base, MT, qbase + (size_t)q0 * ld, ld, T - q0, dh, MT);
        stage_tile_w<DHP, ATT_THREADS>(odO.base, MT, dobase + (size_t)q0 * ldo, ldo, T - q0, dh, MT);
      }
      stage_tile_w<DHP, ATT_THREADS>(oK.base, BN, kbase + (size_t)n0 * ld, ld, nvalid, dh, BN);
      stage_tile_w<DHP, ATT_THREADS>(oV.base, BN, vbase + (size_t)n0 * ld, ld, nvalid, dh, BN);
    }
  };
  load_kv(0, true);

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
    if (blk > 0) load_kv(blk, false);
    if (tma) { mbar_wait(&ctl->ld_bar, ld_phase); ld_phase ^= 1; }
    else cp_async_wait_all();
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma_x(tS, oQ, false, oK, false, BN, DHP / 16, false);
      issue_mma_x(tdP, odO, false, oV, false, BN, DHP / 16, false);
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
      const uint32_t dst = odS.base + (uint32_t)(c0 / 8) * gsS + (uint32_t)tid * 16u;
      sts128(dst, w[0], w[1], w[2], w[3]);
      sts128(dst + gsS, w[4], w[5], w[6], w[7]);
    }
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma_x(tdQ, odS, false, oK, true, DHP, BN / 16, blk > 0);
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
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_umma_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                        const __grid_constant__ CUtensorMap tmQ,
                                                                        const __grid_constant__ CUtensorMap tmdO,
                                                                        const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  const int BQ = a.BN, T = a.T, H = a.H, dh = a.dh;
  float* s_lse = reinterpret_cast<float*>(smem + 128);       // [BQ] lse * log2(e), +inf for padded queries
  float* s_delta = s_lse + 256;                              // [BQ]
  uint8_t* tiles = smem + 3072;
  uint8_t* pK = tiles;
  uint8_t* pV = pK + w_bytes(MT, DHP);
  uint8_t* pQ = pV + w_bytes(MT, DHP);
  uint8_t* pdO = pQ + w_bytes(BQ, DHP);
  uint8_t* sPT_ptr = pdO + w_bytes(BQ, DHP);
  uint8_t* sdST_ptr = sPT_ptr + g8_bytes(MT, BQ);
  const uint32_t gsS = g8_stride(MT);
  const Opnd oK = opnd_w(smem_u32(pK), MT), oV = opnd_w(smem_u32(pV), MT), oQ = opnd_w(smem_u32(pQ), BQ),
             odO = opnd_w(smem_u32(pdO), BQ), oPT = opnd_g8(smem_u32(sPT_ptr), gsS), odST = opnd_g8(smem_u32(sdST_ptr), gsS);
  const bool tma = a.use_tma != 0;
  uint32_t ld_phase = 0;

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

  for (int blk = 0; blk < a.nblocks; ++blk) {
    const int i0 = blk * BQ;
    const int nvalid = min(BQ, T - i0);
    if (tma) {
      if (tid == 0) {
        mbar_expect_tx(&ctl->ld_bar, 2 * w_bytes(BQ, DHP) + (blk == 0 ? 2 * w_bytes(MT, DHP) : 0u));
        if (blk == 0) {
          tma_load_w<DHP>(pK, MT, &tmK, &ctl->ld_bar, (H + hd) * dh, k0, b);
          tma_load_w<DHP>(pV, MT, &tmK, &ctl->ld_bar, (2 * H + hd) * dh, k0, b);
        }
        tma_load_w<DHP>(pQ, BQ, &tmQ, &ctl->ld_bar, hd * dh, i0, b);
        tma_load_w<DHP>(pdO, BQ, &tmdO, &ctl->ld_bar, hd * dh, i0, b);
      }
    } else {
      if (blk == 0) {
        stage_tile_w<DHP, ATT_THREADS>(oK.base, MT, kbase + (size_t)k0 * ld, ld, T - k0, dh, MT);
        stage_tile_w<DHP, ATT_THREADS>(oV.base, MT, vbase + (size_t)k0 * ld, ld, T - k0, dh, MT);
      }
      stage_tile_w<DHP, ATT_THREADS>(oQ.base, BQ, qbase + (size_t)i0 * ld, ld, nvalid, dh, BQ);
      stage_tile_w<DHP, ATT_THREADS>(odO.base, BQ, dobase + (size_t)i0 * ldo, ldo, nvalid, dh, BQ);
    }
    for (int i = tid; i < BQ; i += ATT_THREADS) {
      const bool ok = i < nvalid;
      s_lse[i] = ok ? a.lse[(size_t)bh * T + i0 + i] * 1.4426950408889634f : INFINITY;
      s_delta[i] = ok ? a.delta[(size_t)bh * T + i0 + i] : 0.f;
    }
    if (tma) { mbar_wait(&ctl->ld_bar, ld_phase); ld_phase ^= 1; }
    else cp_async_wait_all();
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma_x(tS, oK, false, oQ, false, BQ, DHP / 16, false);
      issue_mma_x(tdP, oV, false, odO, false, BQ, DHP / 16, false);
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
      const uint32_t off = (uint32_t)(c0 / 8) * gsS + (uint32_t)tid * 16u;
      sts128(oPT.base + off, wp[0], wp[1], wp[2], wp[3]);
      sts128(oPT.base + off + gsS, wp[4], wp[5], wp[6], wp[7]);
      sts128(odST.base + off, wd[0], wd[1], wd[2], wd[3]);
      sts128(odST.base + off + gsS, wd[4], wd[5], wd[6], wd[7]);
    }
    publish_smem_and_sync();
    if (tid == 0) {
      issue_mma_x(tdV, oPT, false, odO, true, DHP, BQ / 16, blk > 0);
      issue_mma_x(tdK, odST, false, oQ, true, DHP, BQ / 16, blk > 0);
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


// ------------------------------------------------------------------------------------------ pipelined long-T backward
// The two kernels above run load -> MMA -> exponentials -> MMA strictly one after the other, one CTA per SM
// (ds3, T = 450: 154 TFLOP/s), and the thread that issues the MMAs is also one of the 128 softmax workers: the
// products of a block are short on flops but long on operand fetches (M = 128, N = 80: 6.5 KB of shared memory per
// 16-deep step), the tensor queue pushes back, and ~800 cycles of issue per block land on the critical path.
// The *_pipe variants keep the arithmetic and the tile formats and reorganise the block loop (TMA operands only):
//   - warp 8 is the ISSUER: one lane requests the TMA loads and issues every MMA; warps 0-7 are WORKERS (warps w
//     and w + 4 own the same 32 rows / TMEM lane quadrant and split the columns of every tile).  The two sides
//     meet on mbarriers only (no CTA barrier in the loop);
//   - the streamed operand pair (K, V for dQ; Q, dO for dK / dV) sits in NS = 2 or 3 shared-memory stages, a stage
//     is refilled as soon as the accumulating MMAs that read it have completed;
//   - the score products S / dP alternate between two TMEM buffers: the tensor core works on block j + 1 while
//     the workers turn block j into dS (and P^T).
constexpr int PIPE_WORKERS = 256;
constexpr int PIPE_THREADS = PIPE_WORKERS + 32;
struct PipeCtl {
  uint64_t ld[3];     // TMA: stage s holds its operand pair (dQ: its K block)            (issuer -> issuer)
  uint64_t ldv[2];    // dQ kernel: V stage s
  uint64_t sbar[2];   // score products of a block are in TMEM buffer u                   (tensor core -> workers)
  uint64_t sfree[2];  // every worker warp has read TMEM buffer u                         (workers -> issuer)
  uint64_t dsfull;    // dS (and P^T) of a block are in shared memory                     (workers -> issuer)
  uint64_t acc;       // accumulating products of a block done: dS / P^T and its stage are free (tensor core -> all)
  uint32_t tmem_slot;
};
__device__ __forceinline__ uint32_t pipe_prologue(PipeCtl* ctl, uint32_t tmem_cols) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&ctl->ld[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&ctl->ldv[i], 1); mbar_init(&ctl->sbar[i], 1); mbar_init(&ctl->sfree[i], PIPE_WORKERS / 32); }
    mbar_init(&ctl->dsfull, PIPE_WORKERS / 32);
    mbar_init(&ctl->acc, 1);
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 1) tmem_alloc_dyn(&ctl->tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  return ctl->tmem_slot;
}
// worker side: this warp has pulled its part of the block's score buffer into registers ...
__device__ __forceinline__ void pipe_scores_read(PipeCtl* ctl, int blk) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&ctl->sfree[blk & 1]);
}
// ... and has written its part of dS (and P^T) to shared memory
__device__ __forceinline__ void pipe_worker_done(PipeCtl* ctl) {
  fence_proxy_async();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&ctl->dsfull);
}

template <int DHP>
__global__ void __launch_bounds__(PIPE_THREADS) attn_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                        const __grid_constant__ CUtensorMap tmKV,
                                                                        const __grid_constant__ CUtensorMap tmdO,
                                                                        const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem);
  uint8_t* tiles = smem + 1024;
  const int BN = a.BN, T = a.T, H = a.H, dh = a.dh, nb = a.nblocks;
  // K (read by S = Q K^T and by dQ += dS K) lives until the block's accumulation is done: 3 stages.  V (read by
  // dP = dO V^T only) is free as soon as the block's scores are: 2 stages.  Every load is then requested two
  // blocks before its first use.
  constexpr int NSK = 3, NSV = 2;
  float* s_part = reinterpret_cast<float*>(smem + 256);     // [128] half-row partial sums of delta
  uint8_t* pQ = tiles;
  uint8_t* pdO = pQ + w_bytes(MT, DHP);
  uint8_t* pKs = pdO + w_bytes(MT, DHP);
  uint8_t* pVs = pKs + (size_t)NSK * w_bytes(BN, DHP);
  uint8_t* sdS_ptr = pVs + (size_t)NSV * w_bytes(BN, DHP);
  const uint32_t gsS = g8_stride(MT), kv_tile = w_bytes(BN, DHP);
  const Opnd oQ = opnd_w(smem_u32(pQ), MT), odO = opnd_w(smem_u32(pdO), MT), odS = opnd_g8(smem_u32(sdS_ptr), gsS);

  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int q0 = blockIdx.x * MT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;

  ALap L(a.dbg);
  const uint32_t tmem = pipe_prologue(ctl, a.tmem_cols);
  const uint32_t tdQ = tmem + 4u * (uint32_t)BN;            // buffer u: S at tmem + 2 u BN, dP at + BN
  L.lap(0);

  if (warp == PIPE_WORKERS / 32) {
    // ------------------------------------------------------------------------------------ issuer
    if ((tid & 31) == 0) {
      auto load_k = [&](int j) {
        const int st = j % NSK;
        mbar_expect_tx(&ctl->ld[st], kv_tile + (j == 0 ? 2 * w_bytes(MT, DHP) : 0u));
        if (j == 0) {
          tma_load_w<DHP>(pQ, MT, &tmQ, &ctl->ld[st], hd * dh, q0, b);
          tma_load_w<DHP>(pdO, MT, &tmdO, &ctl->ld[st], hd * dh, q0, b);
        }
        tma_load_w<DHP>(pKs + (size_t)st * kv_tile, BN, &tmKV, &ctl->ld[st], (H + hd) * dh, j * BN, b);
      };
      auto load_v = [&](int j) {
        const int st = j % NSV;
        mbar_expect_tx(&ctl->ldv[st], kv_tile);
        tma_load_w<DHP>(pVs + (size_t)st * kv_tile, BN, &tmKV, &ctl->ldv[st], (2 * H + hd) * dh, j * BN, b);
      };
      auto issue_scores = [&](int j) {                      // S = Q K^T, dP = dO V^T of block j -> buffer j & 1
        if (j >= 2) mbar_wait(&ctl->sfree[j & 1], (uint32_t)((j >> 1) - 1) & 1u);
        mbar_wait(&ctl->ld[j % NSK], (uint32_t)(j / NSK) & 1u);
        mbar_wait(&ctl->ldv[j % NSV], (uint32_t)(j / NSV) & 1u);
        tc_fence_after();
        const uint32_t tS = tmem + (uint32_t)(j & 1) * 2u * (uint32_t)BN;
        issue_mma_x(tS, oQ, false, opnd_w(smem_u32(pKs + (size_t)(j % NSK) * kv_tile), BN), false, BN, DHP / 16, false);
        issue_mma_x(tS + (uint32_t)BN, odO, false, opnd_w(smem_u32(pVs + (size_t)(j % NSV) * kv_tile), BN), false, BN,
                    DHP / 16, false);
        umma_commit(&ctl->sbar[j & 1]);
      };
      for (int j = 0; j < NSK && j < nb; ++j) {  // in the order of first use
        load_k(j);
        if (j < NSV) load_v(j);
      }
      issue_scores(0);
      for (int blk = 0; blk < nb; ++blk) {
        if (blk > 0) {  // dQ += dS K of block blk - 1 has completed: its K stage takes block blk - 1 + NSK
          mbar_wait(&ctl->acc, (uint32_t)(blk - 1) & 1u);
          if (blk - 1 + NSK < nb) load_k(blk - 1 + NSK);
        }
        if (blk + 1 < nb) issue_scores(blk + 1);
        if (blk + NSV < nb) {  // dP of this block is complete: its V stage takes block blk + NSV
          mbar_wait(&ctl->sbar[blk & 1], (uint32_t)(blk >> 1) & 1u);
          load_v(blk + NSV);
        }
        mbar_wait(&ctl->dsfull, (uint32_t)blk & 1u);
        tc_fence_after();
        issue_mma_x(tdQ, odS, false, opnd_w(smem_u32(pKs + (size_t)(blk % NSK) * kv_tile), BN), true, DHP, BN / 16, blk > 0);
        umma_commit(&ctl->acc);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------ workers
    const int half = warp >> 2, row = (warp & 3) * 32 + (tid & 31);
    const int nch = BN / 16, ch0 = half ? (nch + 1) / 2 : 0, ch1 = half ? nch : (nch + 1) / 2;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const bf16* dobase = a.d_o + (size_t)b * T * ldo + (size_t)hd * dh;
    const bf16* obase = a.o + (size_t)b * T * ldo + (size_t)hd * dh;
    // row statistics: delta = sum_d dO * O (the two warps of a row take half of the head dims each and meet in
    // shared memory), lse in log2 units
    const int q = q0 + row;
    float delta = 0.f, lse2 = 0.f;
    {
      constexpr int NC = DHP / 8, C0 = (NC + 1) / 2;
      const int cb = half ? C0 : 0, ce = half ? NC : C0;
      float part = 0.f;
      if (q < T) {
        const bf16* dor = dobase + (size_t)q * ldo;
        const bf16* orow = obase + (size_t)q * ldo;
        uint4 xv[C0], yv[C0];
#pragma unroll
        for (int c = 0; c < C0; ++c) {
          const bool in = cb + c < ce && (cb + c) * 8 < dh;
          xv[c] = in ? *reinterpret_cast<const uint4*>(dor + (cb + c) * 8) : make_uint4(0, 0, 0, 0);
          yv[c] = in ? *reinterpret_cast<const uint4*>(orow + (cb + c) * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < C0; ++c) {
          const uint32_t xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, ys[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[i]));
            const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[i]));
            part = fmaf(fx.x, fy.x, part);
            part = fmaf(fx.y, fy.y, part);
          }
        }
        lse2 = a.lse[(size_t)bh * T + q] * 1.4426950408889634f;
      }
      if (half) s_part[row] = part;
      asm volatile("bar.sync 1, %0;" ::"n"(PIPE_WORKERS) : "memory");
      if (!half) s_part[row] = part = part + s_part[row];
      asm volatile("bar.sync 1, %0;" ::"n"(PIPE_WORKERS) : "memory");
      delta = s_part[row];
      if (!half && q < T) a.delta[(size_t)bh * T + q] = delta;
    }
    L.lap(1);
    for (int blk = 0; blk < nb; ++blk) {
      const int nvalid = min(BN, T - blk * BN);
      mbar_wait(&ctl->sbar[blk & 1], (uint32_t)(blk >> 1) & 1u);
      tc_fence_after();
      L.lap(2);
      L.lap(3);
      const uint32_t tS = tmem + (uint32_t)(blk & 1) * 2u * (uint32_t)BN, tdP = tS + (uint32_t)BN;
      for (int c0 = ch0 * 16; c0 < ch1 * 16; c0 += 16) {
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
        // dS of the previous block must have been consumed before the first store of this one
        if (c0 == ch0 * 16 && blk > 0) mbar_wait(&ctl->acc, (uint32_t)(blk - 1) & 1u);
        const uint32_t dst = odS.base + (uint32_t)(c0 / 8) * gsS + (uint32_t)row * 16u;
        sts128(dst, w[0], w[1], w[2], w[3]);
        sts128(dst + gsS, w[4], w[5], w[6], w[7]);
      }
      pipe_scores_read(ctl, blk);
      L.lap(4);
      pipe_worker_done(ctl);
      L.lap(5);
    }
    mbar_wait(&ctl->acc, (uint32_t)(nb - 1) & 1u);
    tc_fence_after();
    L.lap(6);
    bf16* out = a.dqkv + ((size_t)b * T + min(q, T - 1)) * ld + (size_t)hd * dh;
    constexpr int OCH = DHP / 16;
    for (int c0 = (half ? (OCH + 1) / 2 : 0) * 16; c0 < (half ? OCH : (OCH + 1) / 2) * 16; c0 += 16) {
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
    L.lap(7);
  }
  attn_epilogue(tmem, a.tmem_cols);
  L.lap(8);
  L.flush();
}

template <int DHP>
__global__ void __launch_bounds__(PIPE_THREADS) attn_bwd_dkv_pipe_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                         const __grid_constant__ CUtensorMap tmQ,
                                                                         const __grid_constant__ CUtensorMap tmdO,
                                                                         const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem);
  const int BQ = a.BN, T = a.T, H = a.H, dh = a.dh, NS = a.stages, nb = a.nblocks;
  uint8_t* tiles = smem + 3072;
  uint8_t* pK = tiles;
  uint8_t* pV = pK + w_bytes(MT, DHP);
  uint8_t* pQdO = pV + w_bytes(MT, DHP);                     // stage s: Q at + s * 2 * w_bytes(BQ), dO right after
  uint8_t* sPT_ptr = pQdO + (size_t)NS * 2 * w_bytes(BQ, DHP);
  uint8_t* sdST_ptr = sPT_ptr + g8_bytes(MT, BQ);
  float* s_lse = reinterpret_cast<float*>(sdST_ptr + g8_alloc(MT, BQ));  // [nb * BQ] lse * log2(e), +inf for padded queries
  float* s_delta = s_lse + nb * BQ;                                      // [nb * BQ]
  const uint32_t gsS = g8_stride(MT), q_stage = 2 * w_bytes(BQ, DHP);
  const Opnd oK = opnd_w(smem_u32(pK), MT), oV = opnd_w(smem_u32(pV), MT), oPT = opnd_g8(smem_u32(sPT_ptr), gsS),
             odST = opnd_g8(smem_u32(sdST_ptr), gsS);

  const int bh = blockIdx.y, b = bh / H, hd = bh % H;
  const int k0 = blockIdx.x * MT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t ld = (size_t)3 * H * dh;

  ALap L(a.dbg);
  const uint32_t tmem = pipe_prologue(ctl, a.tmem_cols);
  L.lap(0);
  const uint32_t tdV = tmem + 4u * (uint32_t)BQ, tdK = tdV + DHP;  // buffer u: S^T at tmem + 2 u BQ, dP^T at + BQ

  if (warp == PIPE_WORKERS / 32) {
    // ------------------------------------------------------------------------------------ issuer
    if ((tid & 31) == 0) {
      auto load = [&](int j) {
        const int st = j % NS;
        uint8_t* pQ = pQdO + (size_t)st * q_stage;
        mbar_expect_tx(&ctl->ld[st], q_stage + (j == 0 ? 2 * w_bytes(MT, DHP) : 0u));
        if (j == 0) {
          tma_load_w<DHP>(pK, MT, &tmK, &ctl->ld[st], (H + hd) * dh, k0, b);
          tma_load_w<DHP>(pV, MT, &tmK, &ctl->ld[st], (2 * H + hd) * dh, k0, b);
        }
        tma_load_w<DHP>(pQ, BQ, &tmQ, &ctl->ld[st], hd * dh, j * BQ, b);
        tma_load_w<DHP>(pQ + w_bytes(BQ, DHP), BQ, &tmdO, &ctl->ld[st], hd * dh, j * BQ, b);
      };
      auto issue_scores = [&](int j) {                       // S^T = K Q^T, dP^T = V dO^T of block j
        const int st = j % NS;
        if (j >= 2) mbar_wait(&ctl->sfree[j & 1], (uint32_t)((j >> 1) - 1) & 1u);
        mbar_wait(&ctl->ld[st], (uint32_t)(j / NS) & 1u);
        tc_fence_after();
        const uint32_t qb = smem_u32(pQdO + (size_t)st * q_stage);
        const uint32_t tS = tmem + (uint32_t)(j & 1) * 2u * (uint32_t)BQ;
        issue_mma_x(tS, oK, false, opnd_w(qb, BQ), false, BQ, DHP / 16, false);
        issue_mma_x(tS + (uint32_t)BQ, oV, false, opnd_w(qb + w_bytes(BQ, DHP), BQ), false, BQ, DHP / 16, false);
        umma_commit(&ctl->sbar[j & 1]);
      };
      for (int j = 0; j < NS && j < nb; ++j) load(j);
      issue_scores(0);
      for (int blk = 0; blk < nb; ++blk) {
        if (blk > 0) {  // dV += P^T dO and dK += dS^T Q of block blk - 1 have completed
          mbar_wait(&ctl->acc, (uint32_t)(blk - 1) & 1u);
          if (blk - 1 + NS < nb) load(blk - 1 + NS);
        }
        if (blk + 1 < nb) issue_scores(blk + 1);
        mbar_wait(&ctl->dsfull, (uint32_t)blk & 1u);
        tc_fence_after();
        const uint32_t qb = smem_u32(pQdO + (size_t)(blk % NS) * q_stage);
        issue_mma_x(tdV, oPT, false, opnd_w(qb + w_bytes(BQ, DHP), BQ), true, DHP, BQ / 16, blk > 0);
        issue_mma_x(tdK, odST, false, opnd_w(qb, BQ), true, DHP, BQ / 16, blk > 0);
        umma_commit(&ctl->acc);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------ workers
    const int half = warp >> 2, row = (warp & 3) * 32 + (tid & 31);
    const int nch = BQ / 16, ch0 = half ? (nch + 1) / 2 : 0, ch1 = half ? nch : (nch + 1) / 2;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    // lse / delta of every query of this (sample, head), once
    for (int i = tid; i < nb * BQ; i += PIPE_WORKERS) {
      const bool ok = i < T;
      s_lse[i] = ok ? a.lse[(size_t)bh * T + i] * 1.4426950408889634f : INFINITY;
      s_delta[i] = ok ? a.delta[(size_t)bh * T + i] : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PIPE_WORKERS) : "memory");
    L.lap(1);
    for (int blk = 0; blk < nb; ++blk) {
      mbar_wait(&ctl->sbar[blk & 1], (uint32_t)(blk >> 1) & 1u);
      tc_fence_after();
      L.lap(2);
      const uint32_t tS = tmem + (uint32_t)(blk & 1) * 2u * (uint32_t)BQ, tdP = tS + (uint32_t)BQ;
      const float4* l4 = reinterpret_cast<const float4*>(s_lse + blk * BQ);
      const float4* d4 = reinterpret_cast<const float4*>(s_delta + blk * BQ);
      for (int c0 = ch0 * 16; c0 < ch1 * 16; c0 += 16) {
        float s[16], dp[16];
        tmem_ld16(tS + lane_off + c0, s);
        tmem_ld16(tdP + lane_off + c0, dp);
        tmem_ld_wait();
        float l[16], d[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 lv = l4[c0 / 4 + i], dv = d4[c0 / 4 + i];
          l[4 * i] = lv.x; l[4 * i + 1] = lv.y; l[4 * i + 2] = lv.z; l[4 * i + 3] = lv.w;
          d[4 * i] = dv.x; d[4 * i + 1] = dv.y; d[4 * i + 2] = dv.z; d[4 * i + 3] = dv.w;
        }
        uint32_t wp[8], wd[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float p0 = exp2_fast(fmaf(s[i], a.scale_log2, -l[i]));       // padded query: exp2(-inf) = 0
          const float p1 = exp2_fast(fmaf(s[i + 1], a.scale_log2, -l[i + 1]));
          wp[i / 2] = pack_bf16(p0, p1);
          wd[i / 2] = pack_bf16(p0 * (dp[i] - d[i]), p1 * (dp[i + 1] - d[i + 1]));
        }
        // P^T / dS^T of the previous block must have been consumed before the first store of this one
        if (c0 == ch0 * 16 && blk > 0) mbar_wait(&ctl->acc, (uint32_t)(blk - 1) & 1u);
        const uint32_t off = (uint32_t)(c0 / 8) * gsS + (uint32_t)row * 16u;
        sts128(oPT.base + off, wp[0], wp[1], wp[2], wp[3]);
        sts128(oPT.base + off + gsS, wp[4], wp[5], wp[6], wp[7]);
        sts128(odST.base + off, wd[0], wd[1], wd[2], wd[3]);
        sts128(odST.base + off + gsS, wd[4], wd[5], wd[6], wd[7]);
      }
      pipe_scores_read(ctl, blk);
      L.lap(4);
      pipe_worker_done(ctl);
      L.lap(5);
    }
    mbar_wait(&ctl->acc, (uint32_t)(nb - 1) & 1u);
    tc_fence_after();
    L.lap(6);

    const int key = k0 + row;
    bf16* dkout = a.dqkv + ((size_t)b * T + min(key, T - 1)) * ld + (size_t)H * dh + (size_t)hd * dh;
    bf16* dvout = dkout + (size_t)H * dh;
    constexpr int OCH = DHP / 16;
    for (int c0 = (half ? (OCH + 1) / 2 : 0) * 16; c0 < (half ? OCH : (OCH + 1) / 2) * 16; c0 += 16) {
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
    L.lap(7);
  }
  attn_epilogue(tmem, a.tmem_cols);
  L.lap(8);
  L.flush();
}


// ------------------------------------------------------------------------------------------ fused backward
// Short sequences (T <= 160: every key fits one MMA N extent): ONE CTA per (sample, head) computes dQ,
// dK and dV with S and P evaluated once, in ONE pass over the query rows: rows [0, 128) are the main M
// tile, rows [128, T) a 32-row tail tile whose products are issued in the same MMA batches into their own
// TMEM columns (the tail would otherwise repeat the whole dependent chain for a handful of rows).
// 256 threads: warps w and w + 4 own the same 32 rows (TMEM lane quadrant w % 4) and split the key columns;
// the tail rows belong to the two quadrant-0 warps.
//   S = Q K^T -> TMEM          p = exp2(s c - lse) -> bf16 P in smem [q][key]
//   dP = dO V^T -> TMEM (over S)   dV^T += dO^T P      (A = dO read MN-major, B = P read MN-major)
//   dS = p (dP - delta) -> over P in smem
//   dQ = dS K -> TMEM (over dP)    dK^T += Q^T dS      (A = Q read MN-major, B = dS read MN-major)
// dK^T / dV^T live in TMEM as [d (lane)][key (column)] and leave through a [key][d] staging tile in shared
// memory (coalesced 16-byte rows); the transposed products are what lets one [q][key] tile of P / dS feed
// both contractions.  TMEM columns: [0, R) S / dP / dQ, [R, R + NK) dV^T, [R + NK, R + 2 NK) dK^T -- which
// first holds the tail's S / dP -- and the tail's dQ at dq2_col.  delta = rowsum(dO * O) is taken from the
// staged dO and O tiles (O lands where P goes later).
constexpr int FUSED_THREADS = 256;
constexpr int TAILR = 32;  // rows of the tail query tile
enum { OQ = 0, ODO, OK_, OV, OP, OQ2, ODO2, OP2 };

// sum_d x[r][d] * y[r][d] over the valid head dims of row r of two W tiles
__device__ __forceinline__ float row_dot_w(uint32_t x, uint32_t y, int rows, int r, int dh) {
  float acc = 0.f;
  for (int c = 0; c < dh / 8; ++c) {
    const uint32_t off = (uint32_t)(c >> 3) * w_box(rows) + (uint32_t)r * 128u + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
    const uint4 xv = lds128(x + off), yv = lds128(y + off);
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[i]));
      const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[i]));
      acc = fmaf(fx.x, fy.x, acc);
      acc = fmaf(fx.y, fy.y, acc);
    }
  }
  return acc;
}

struct FusedMaps { CUtensorMap q, kv, q2, d_o, d_o2, o, o2; };

template <int DHP>
__global__ void __launch_bounds__(FUSED_THREADS) attn_bwd_fused_umma_kernel(const __grid_constant__ FusedMaps tm,
                                                                          const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  AttnSmem* ctl = reinterpret_cast<AttnSmem*>(smem);
  float* s_delta = reinterpret_cast<float*>(smem + 128);  // [128 + 32]: main rows, tail rows
  float* s_lse = s_delta + MT + TAILR;                    // [128 + 32], log2 units
  const int NK = a.BN, T = a.T, H = a.H, dh = a.dh;
  uint8_t* tiles = smem + 2048;
  uint8_t* pQ = tiles + a.off[OQ];
  uint8_t* pdO = tiles + a.off[ODO];
  uint8_t* pK = tiles + a.off[OK_];
  uint8_t* pV = tiles + a.off[OV];
  uint8_t* pP = tiles + a.off[OP];
  uint8_t* pQ2 = tiles + a.off[OQ2];
  uint8_t* pdO2 = tiles + a.off[ODO2];
  uint8_t* pP2 = tiles + a.off[OP2];
  const uint32_t gsP = g8_stride(MT), gsP2 = g8_stride(TAILR);
  const Opnd oQ = opnd_w(smem_u32(pQ), MT), odO = opnd_w(smem_u32(pdO), MT), oK = opnd_w(smem_u32(pK), NK),
             oV = opnd_w(smem_u32(pV), NK), oP = opnd_g8(smem_u32(pP), gsP);
  const Opnd oQ2 = opnd_w(smem_u32(pQ2), TAILR), odO2 = opnd_w(smem_u32(pdO2), TAILR), oP2 = opnd_g8(smem_u32(pP2), gsP2);
  const bool tma = a.use_tma != 0;

  const int bh = blockIdx.x, b = bh / H, hd = bh % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int r = quad * 32 + lane;  // row of the main tile = TMEM lane
  const size_t ld = (size_t)3 * H * dh, ldo = (size_t)H * dh;
  const bf16* qbase = a.qkv + (size_t)b * T * ld + (size_t)hd * dh;
  const bf16* kbase = qbase + (size_t)H * dh;
  const bf16* vbase = qbase + (size_t)2 * H * dh;
  const bf16* dobase = a.d_o + (size_t)b * T * ldo + (size_t)hd * dh;
  const bf16* obase = a.o + (size_t)b * T * ldo + (size_t)hd * dh;

  const bool tail = T > MT;
  const int n1 = min(T, MT);                     // valid rows of the main tile
  const int kq1 = (n1 + 15) / 16;                // K steps of the contractions over its rows
  const int kq2 = tail ? (T - MT + 15) / 16 : 0;  // ... over the tail rows
  const bool warp_rows = quad * 32 < kq1 * 16;   // this warp's main rows take part in those contractions
  const bool tail_warp = tail && quad == 0;      // warps 0 and 4: tail row = lane

  // prologue (256 threads): barriers + TMEM
  if (tid == 0) { mbar_init(&ctl->bar, 1); mbar_init(&ctl->ld_bar, 1); mbar_init(&ctl->v_bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_dyn(&ctl->tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = ctl->tmem_slot;
  const int R = NK > DHP ? NK : DHP;  // S / dP / dQ share the first R columns
  const uint32_t tS = tmem, tdV = tmem + (uint32_t)R, tdK = tdV + (uint32_t)NK, tdQ2 = tmem + a.dq2_col;
  const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
  uint32_t phase = 0;

  // key-column chunks (16 wide) of this warp half
  const int nchunks = NK / 16;
  const int ch0 = half == 0 ? 0 : (nchunks + 1) / 2;
  const int ch1 = half == 0 ? (nchunks + 1) / 2 : nchunks;

  ALap L(a.dbg);
  // ---- operand tiles: Q, K (and the tail Q) on the first barrier -- S starts as soon as they are in -- dO, O, V
  // on the second
  if (tma) {
    if (tid == 0) {
      mbar_expect_tx(&ctl->ld_bar, w_bytes(MT, DHP) + w_bytes(NK, DHP) + (tail ? w_bytes(TAILR, DHP) : 0u));
      tma_load_w<DHP>(pK, NK, &tm.kv, &ctl->ld_bar, (H + hd) * dh, 0, b);
      tma_load_w<DHP>(pQ, MT, &tm.q, &ctl->ld_bar, hd * dh, 0, b);
      if (tail) tma_load_w<DHP>(pQ2, TAILR, &tm.q2, &ctl->ld_bar, hd * dh, MT, b);
      mbar_expect_tx(&ctl->v_bar, 2 * w_bytes(MT, DHP) + w_bytes(NK, DHP) + (tail ? 2 * w_bytes(TAILR, DHP) : 0u));
      tma_load_w<DHP>(pdO, MT, &tm.d_o, &ctl->v_bar, hd * dh, 0, b);
      tma_load_w<DHP>(pP, MT, &tm.o, &ctl->v_bar, hd * dh, 0, b);
      if (tail) {
        tma_load_w<DHP>(pdO2, TAILR, &tm.d_o2, &ctl->v_bar, hd * dh, MT, b);
        tma_load_w<DHP>(pP2, TAILR, &tm.o2, &ctl->v_bar, hd * dh, MT, b);
      }
      tma_load_w<DHP>(pV, NK, &tm.kv, &ctl->v_bar, (2 * H + hd) * dh, 0, b);
    }
    // one CTA per SM and nothing to overlap its loads with: pull the tiles of the CTA that will follow on this
    // SM (same index + number of SMs) into L2 while this one computes
    const int nbh = bh + a.pf_stride;
    if (warp == 2 && lane == 0 && a.pf_stride > 0 && nbh < (int)gridDim.x) {
      const int nb = nbh / H, nh = nbh % H;
      tma_prefetch_w<DHP>(&tm.kv, (H + nh) * dh, 0, nb);
      tma_prefetch_w<DHP>(&tm.q, nh * dh, 0, nb);
      tma_prefetch_w<DHP>(&tm.d_o, nh * dh, 0, nb);
      tma_prefetch_w<DHP>(&tm.o, nh * dh, 0, nb);
      tma_prefetch_w<DHP>(&tm.kv, (2 * H + nh) * dh, 0, nb);
      if (tail) {
        tma_prefetch_w<DHP>(&tm.q2, nh * dh, MT, nb);
        tma_prefetch_w<DHP>(&tm.d_o2, nh * dh, MT, nb);
        tma_prefetch_w<DHP>(&tm.o2, nh * dh, MT, nb);
      }
    }
  } else {
    stage_tile_w<DHP, FUSED_THREADS>(oK.base, NK, kbase, ld, T, dh, NK);
    stage_tile_w<DHP, FUSED_THREADS>(oQ.base, MT, qbase, ld, T, dh, kq1 * 16);
    stage_tile_w<DHP, FUSED_THREADS>(odO.base, MT, dobase, ldo, T, dh, kq1 * 16);
    stage_tile_w<DHP, FUSED_THREADS>(smem_u32(pP), MT, obase, ldo, T, dh, kq1 * 16);
    stage_tile_w<DHP, FUSED_THREADS>(oV.base, NK, vbase, ld, T, dh, NK);
    if (tail) {
      stage_tile_w<DHP, FUSED_THREADS>(oQ2.base, TAILR, qbase + (size_t)MT * ld, ld, T - MT, dh, TAILR);
      stage_tile_w<DHP, FUSED_THREADS>(odO2.base, TAILR, dobase + (size_t)MT * ldo, ldo, T - MT, dh, TAILR);
      stage_tile_w<DHP, FUSED_THREADS>(smem_u32(pP2), TAILR, obase + (size_t)MT * ldo, ldo, T - MT, dh, TAILR);
    }
  }
  // lse of the rows whose statistics this thread produces: half 0 -> main row r, warp 4 -> tail row `lane`
  const bool stat_main = half == 0, stat_tail = tail && warp == 4;
  float my_lse2 = INFINITY;  // padded query row: p = exp2(-inf) = 0
  {
    const int q = stat_main ? r : MT + lane;
    if ((stat_main || stat_tail) && q < T) my_lse2 = a.lse[(size_t)bh * T + q] * 1.4426950408889634f;
  }
  auto issue_S = [&]() {
    issue_mma_x(tS, oQ, false, oK, false, NK, DHP / 16, false);
    if (tail) issue_mma_x(tdK, oQ2, false, oK, false, NK, DHP / 16, false);
    umma_commit(&ctl->bar);
  };
  L.lap(0);
  if (tma) {
    if (tid == 0) {
      mbar_wait(&ctl->ld_bar, 0);
      tc_fence_after();
      issue_S();
    }
    __syncwarp();
    mbar_wait(&ctl->v_bar, 0);
  } else {
    cp_async_wait_all();
    publish_smem_and_sync();
    if (tid == 0) issue_S();
  }
  L.lap(1);
  // ---- row statistics from the staged tiles: delta = sum_d dO * O
  if (stat_main) {
    s_delta[r] = r < T ? row_dot_w(odO.base, oP.base, MT, r, dh) : 0.f;
    s_lse[r] = my_lse2;
  } else if (stat_tail) {
    s_delta[MT + lane] = MT + lane < T ? row_dot_w(odO2.base, oP2.base, TAILR, lane, dh) : 0.f;
    s_lse[MT + lane] = my_lse2;
  }
  __syncthreads();  // statistics visible; every read of the O tiles is done before P overwrites them
  const float delta = s_delta[r], lse2 = s_lse[r];
  const float delta_t = tail_warp ? s_delta[MT + lane] : 0.f, lse2_t = tail_warp ? s_lse[MT + lane] : INFINITY;
  mbar_wait(&ctl->bar, phase); phase ^= 1;
  tc_fence_after();
  L.lap(2);

  // ---- P = exp2(s c - lse) (zero for padded rows / keys) -> smem [q][key].  Row set 0 = main rows, 1 = tail
  // rows; the loops stay rolled: the kernel runs once per CTA, straight-line code would only cost
  // instruction fetches
  const int nsets = tail ? 2 : 1;
#pragma unroll 1
  for (int set = 0; set < nsets; ++set) {
    if (!(set == 0 ? warp_rows : tail_warp)) continue;  // warp-uniform
    const uint32_t t_src = set == 0 ? tS + lane_off : tdK;
    const uint32_t gs = set == 0 ? gsP : gsP2;
    const uint32_t tile = (set == 0 ? oP.base : oP2.base) + (uint32_t)(set == 0 ? r : lane) * 16u;
    const float lse = set == 0 ? lse2 : lse2_t;
#pragma unroll 1
    for (int ch = ch0; ch < ch1; ++ch) {
      const int c0 = ch * 16;
      float sv[16];
      tmem_ld16(t_src + (uint32_t)c0, sv);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        // a padded row has lse = +inf -> p = 0; padded keys (last chunk only) are zeroed below
        w[i / 2] = pack_bf16(exp2_fast(fmaf(sv[i], a.scale_log2, -lse)), exp2_fast(fmaf(sv[i + 1], a.scale_log2, -lse)));
      }
      if (c0 + 16 > T) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (c0 + 2 * i >= T) w[i] = 0u;
          else if (c0 + 2 * i + 1 >= T) w[i] &= 0xFFFFu;
        }
      }
      const uint32_t dst = tile + (uint32_t)(c0 / 8) * gs;
      sts128(dst, w[0], w[1], w[2], w[3]);
      sts128(dst + gs, w[4], w[5], w[6], w[7]);
    }
  }
  L.lap(3);
  publish_smem_and_sync();
  if (tid == 0) {
    issue_mma_x(tS, odO, false, oV, false, NK, DHP / 16, false);  // dP over S
    if (tail) issue_mma_x(tdK, odO2, false, oV, false, NK, DHP / 16, false);
    // dV^T[d][key] = sum_q dO[q][d] P[q][key]: both operands MN-major (rows = contraction index q)
    issue_mma_x(tdV, odO, true, oP, true, NK, kq1, false);
    if (tail) issue_mma_x(tdV, odO2, true, oP2, true, NK, kq2, true);
    umma_commit(&ctl->bar);
  }
  mbar_wait(&ctl->bar, phase); phase ^= 1;
  tc_fence_after();
  L.lap(4);

  // ---- dS = p (dP - delta), in place over P
#pragma unroll 1
  for (int set = 0; set < nsets; ++set) {
    if (!(set == 0 ? warp_rows : tail_warp)) continue;
    const uint32_t t_src = set == 0 ? tS + lane_off : tdK;
    const uint32_t gs = set == 0 ? gsP : gsP2;
    const uint32_t tile = (set == 0 ? oP.base : oP2.base) + (uint32_t)(set == 0 ? r : lane) * 16u;
    const float dl = set == 0 ? delta : delta_t;
#pragma unroll 1
    for (int ch = ch0; ch < ch1; ++ch) {
      float dp[16];
      tmem_ld16(t_src + (uint32_t)ch * 16u, dp);
      const uint32_t dst = tile + (uint32_t)(ch * 2) * gs;
      const uint4 pa = lds128(dst), pb = lds128(dst + gs);
      tmem_ld_wait();
      const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[i]));
        w[i] = pack_bf16(pf.x * (dp[2 * i] - dl), pf.y * (dp[2 * i + 1] - dl));
      }
      sts128(dst, w[0], w[1], w[2], w[3]);
      sts128(dst + gs, w[4], w[5], w[6], w[7]);
    }
  }
  L.lap(5);
  publish_smem_and_sync();
  if (tid == 0) {
    issue_mma_x(tS, oP, false, oK, true, DHP, NK / 16, false);  // dQ over dP
    if (tail) issue_mma_x(tdQ2, oP2, false, oK, true, DHP, NK / 16, false);
    // dK^T[d][key] = sum_q Q[q][d] dS[q][key], over the tail's (consumed) dP
    issue_mma_x(tdK, oQ, true, oP, true, NK, kq1, false);
    if (tail) issue_mma_x(tdK, oQ2, true, oP2, true, NK, kq2, true);
    umma_commit(&ctl->bar);
  }
  mbar_wait(&ctl->bar, phase); phase ^= 1;
  tc_fence_after();
  L.lap(6);

  // ---- dQ rows out: the two warp halves split the head dimension in 16-column chunks
  {
    constexpr int NDC = DHP / 16;
    const int d0 = half == 0 ? 0 : (NDC + 1) / 2, d1 = half == 0 ? (NDC + 1) / 2 : NDC;
#pragma unroll 1
    for (int set = 0; set < nsets; ++set) {
      if (set == 1 && !tail_warp) continue;
      const uint32_t t_src = set == 0 ? tS + lane_off : tdQ2;
      const int q = set == 0 ? r : MT + lane;
      bf16* out = a.dqkv + ((size_t)b * T + min(q, T - 1)) * ld + (size_t)hd * dh;
#pragma unroll 1
      for (int dc = d0; dc < d1; ++dc) {
        float v[16];
        tmem_ld16(t_src + (uint32_t)dc * 16u, v);
        tmem_ld_wait();
        if (q < T) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (dc * 16 + 8 * h8 < dh) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) w[i] = pack_bf16(v[8 * h8 + 2 * i] * a.scale, v[8 * h8 + 2 * i + 1] * a.scale);
              *reinterpret_cast<uint4*>(out + dc * 16 + 8 * h8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
    }
  }
  L.lap(7);

  // ---- dK^T, dV^T out: thread = head dim d (TMEM lane), columns = keys.  The values go through [key][d]
  // staging tiles (the dead Q / dO and K / V regions; every MMA has completed) and leave as 16-byte pieces
  // of contiguous rows
  const uint32_t pitch = (uint32_t)dh * 2u + 16u;
  const uint32_t stK = oQ.base, stV = oK.base;  // rows [0, NK): padded keys are staged too, never copied out
  if (quad * 32 < dh) {  // warp-uniform: tcgen05.ld is a whole-warp instruction
#pragma unroll 1
    for (int ch = ch0; ch < ch1; ++ch) {
      {
        const int c0 = ch * 16;
        float vk[16], vv[16];
        tmem_ld16(tdK + lane_off + c0, vk);
        tmem_ld16(tdV + lane_off + c0, vv);
        tmem_ld_wait();
        if (r < dh) {
          uint32_t ak = stK + (uint32_t)c0 * pitch + (uint32_t)r * 2u, av = stV + (uint32_t)c0 * pitch + (uint32_t)r * 2u;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint32_t wk = pack_bf16(vk[i] * a.scale, vk[i + 1] * a.scale), wv = pack_bf16(vv[i], vv[i + 1]);
            sts16(ak, wk); sts16(ak + pitch, wk >> 16);
            sts16(av, wv); sts16(av + pitch, wv >> 16);
            ak += 2 * pitch; av += 2 * pitch;
          }
        }
      }
    }
  }
  __syncthreads();
  {
    const int ppr = dh / 8;  // 16-byte pieces per row
    bf16* kout = a.dqkv + (size_t)b * T * ld + (size_t)(H + hd) * dh;
    bf16* vout = kout + (size_t)H * dh;
    for (int idx = tid; idx < T * ppr; idx += FUSED_THREADS) {
      const int key = idx / ppr, c = idx - key * ppr;
      const uint32_t so = (uint32_t)key * pitch + (uint32_t)c * 16u;
      *reinterpret_cast<uint4*>(kout + (size_t)key * ld + c * 8) = lds128(stK + so);
      *reinterpret_cast<uint4*>(vout + (size_t)key * ld + c * 8) = lds128(stV + so);
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
// A / B measurements: V4H_ATTN_{FWD,DQ,DKV}_CAP shrink the key / query block of the multi-block kernels (smaller
// tiles -> more CTAs per SM)
int env_cap(const char* name, int T, int cap) {
  const char* e = getenv(name);
  if (!e || T <= 160) return cap;
  const int v = atoi(e) / 16 * 16;
  return v >= 16 && v < cap ? v : cap;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool attn_tma_enabled() {
  static const int on = [] { const char* e = getenv("V4H_ATTN_TMA"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

// 3-d tensor map over a (B, T, cols) bf16 tensor with row pitch `ld` elements: box = [rows][64 columns],
// 128-byte swizzle -- one box of a W tile; columns past `cols` and rows past T read as zero
int make_w_map(const void* base, int B, int T, int cols, size_t ld, int rows, CUtensorMap* out) {
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
  static std::map<std::tuple<const void*, int, int, int, size_t, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(base, B, T, cols, ld, rows);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return V4H_OK; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(V4H_ERR_CUDA, "attention: cuTensorMapEncodeTiled (swizzled tile view) failed (%d)", (int)r);
  if (cache.size() > 1024) cache.clear();
  cache[key] = m;
  *out = m;
  return V4H_OK;
}

template <int DHP, bool ONE>
int fwd_launch_one(AttnArgs a, int B, cudaStream_t s) {
  const uint32_t qkv_bytes = w_bytes(MT, DHP) + 2 * w_bytes(a.BN, DHP), p_bytes = g8_bytes(MT, a.BN) + 16;
  // one key block: P may overlay Q and K (both dead once S is complete)
  const bool overlay = ONE && p_bytes <= w_bytes(MT, DHP) + w_bytes(a.BN, DHP);
  a.p_off = overlay ? 0 : (int)qkv_bytes;
  const size_t smem = 2048 + qkv_bytes + (overlay ? 0 : p_bytes);
  static size_t configured = 0;
  if (smem > configured) { V4H_TRY((set_smem(attn_fwd_umma_kernel<DHP, ONE>, smem))); configured = smem; }
  CUtensorMap mq, mkv;
  memset(&mq, 0, sizeof(mq)); memset(&mkv, 0, sizeof(mkv));
  a.use_tma = (attn_tma_enabled() && a.dh == DHP) ? 1 : 0;  // columns past dh must read as zero otherwise
  if (a.use_tma) {
    const size_t ld = (size_t)3 * a.H * a.dh;
    V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ld, MT, &mq));
    V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ld, a.BN, &mkv));
  }
  dim3 grid((unsigned)ceil_div(a.T, MT), (unsigned)(B * a.H));
  V4H_CUDA(launch_pdl(attn_fwd_umma_kernel<DHP, ONE>, dim3(grid), dim3(ATT_THREADS), smem, s, mq, mkv, a));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}
template <int DHP>
int fwd_launch(AttnArgs a, int B, cudaStream_t s) {
  int cap = std::min(160, 512 - DHP) / 16 * 16;
  if (a.T > cap) {
    // several key blocks: the block loop is one dependent chain per CTA, so the block is sized for TWO CTAs per SM
    // (ds3, T = 450: 3 blocks of 160 keys, one CTA per SM, 170 us; 5 blocks of 96, two per SM, 113 us)
    for (int bn = cap; bn >= 32; bn -= 16) {
      cap = bn;
      if (2048 + w_bytes(MT, DHP) + 2 * w_bytes(bn, DHP) + g8_bytes(MT, bn) + 16 <= 113 * 1024) break;
    }
  }
  cap = env_cap("V4H_ATTN_FWD_CAP", a.T, cap);
  pick_block(a.T, cap, &a.BN, &a.nblocks);
  a.tmem_cols = pow2_cols(a.BN + DHP);
  return a.nblocks == 1 ? fwd_launch_one<DHP, true>(a, B, s) : fwd_launch_one<DHP, false>(a, B, s);
}

int sm_count() {
  static const int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 0;
    return v;
  }();
  return n;
}
bool fused_bwd_enabled() {
  static const int on = [] { const char* e = getenv("V4H_ATTN_FUSED_BWD"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

// fused backward: tile offsets, TMEM columns; false when the shape does not fit one CTA
template <int DHP>
bool fused_plan(AttnArgs& f, size_t* smem_bytes) {
  auto up = [](uint32_t x) { return (x + 1023u) & ~1023u; };
  f.BN = (int)ceil_div(f.T, 16) * 16;
  f.nblocks = 1;
  const int NK = f.BN;
  const bool tail = f.T > MT;
  if (f.T > MT + TAILR) return false;
  const int R = NK > DHP ? NK : DHP;
  int cols = R + 2 * NK;
  f.dq2_col = 0;
  if (tail) {
    if (2 * DHP <= R) f.dq2_col = (uint32_t)DHP;  // next to the main dQ inside the S region
    else if (cols + DHP <= 512) { f.dq2_col = (uint32_t)cols; cols += DHP; }
    else return false;
  }
  if (cols > 512) return false;
  f.tmem_cols = pow2_cols(cols);
  const uint32_t W1 = w_bytes(MT, DHP), WK = w_bytes(NK, DHP), W2 = w_bytes(TAILR, DHP);
  // the P tiles first receive the O tiles (row statistics)
  const uint32_t P1 = up(std::max(g8_bytes(MT, NK) + 16u, W1)), P2 = up(std::max(g8_bytes(TAILR, NK) + 16u, W2));
  f.off[OQ] = 0; f.off[ODO] = (int)W1; f.off[OK_] = (int)(2 * W1); f.off[OV] = (int)(2 * W1 + WK);
  f.off[OP] = (int)(2 * W1 + 2 * WK);
  uint32_t end = (uint32_t)f.off[OP] + P1;
  f.off[OQ2] = f.off[ODO2] = f.off[OP2] = (int)end;
  if (tail) {
    f.off[OQ2] = (int)end; f.off[ODO2] = (int)(end + W2); f.off[OP2] = (int)(end + 2 * W2);
    end += 2 * W2 + P2;
    // K-major A reads span 128 rows: the 32-row tail tiles are over-read into what follows (garbage rows of
    // the accumulators that are never stored); keep those reads inside the allocation
    end = std::max(end, (uint32_t)f.off[ODO2] + W2 + 128u * 128u);
    end = std::max(end, (uint32_t)f.off[OP2] + (uint32_t)(NK / 8) * g8_stride(TAILR) + 128u * 16u);
  }
  // MN-major A reads span two 64-column chunks even when DHP <= 64; the dK / dV staging tiles ([key][dh + 8])
  // reuse Q + dO and K + V
  end = std::max(end, (uint32_t)f.off[OP] + 2 * w_box(MT));
  if ((uint32_t)f.T * ((uint32_t)f.dh * 2u + 16u) > 2 * std::min(W1, WK)) return false;
  *smem_bytes = 3072 + end;
  return *smem_bytes <= 227 * 1024;
}

bool pipe_enabled() {
  static const int on = [] { const char* e = getenv("V4H_ATTN_PIPE"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}
int pipe_stages() {
  static const int n = [] { const char* e = getenv("V4H_ATTN_PIPE_STAGES"); return (e && e[0] == '3') ? 3 : 2; }();
  return n;
}
// Block length / stage count of a pipelined long-T kernel: `acc_cols` TMEM columns of accumulators next to two
// buffers of two score tiles (4 BN columns); shared memory = `fixed` + NS stages of BN rows of `row_stage` bytes +
// `ntiles` thread-written [128][BN] tiles.  The wanted stage count falls back to 2 when 3 leaves less than 64 rows.
bool pipe_plan(int T, int want_stages, int acc_cols, size_t fixed, size_t row_stage, int ntiles, AttnArgs* out,
               size_t* smem_bytes) {
  for (int ns = want_stages; ns >= (want_stages == 1 ? 1 : 2); --ns) {
    int cap = 0;
    for (int bn = 128; bn >= 16; bn -= 16) {
      const size_t smem = fixed + (size_t)ns * bn * row_stage + (size_t)ntiles * g8_bytes(MT, bn) + 16;
      if (4 * bn + acc_cols <= 512 && smem <= 227 * 1024) { cap = bn; break; }
    }
    if (cap == 0 || (ns == 3 && cap < 64)) continue;
    pick_block(T, cap, &out->BN, &out->nblocks);
    out->stages = ns;
    out->tmem_cols = pow2_cols(4 * out->BN + acc_cols);
    *smem_bytes = fixed + (size_t)ns * out->BN * row_stage + (size_t)ntiles * g8_bytes(MT, out->BN) + 16;
    return true;
  }
  return false;
}

template <int DHP>
int bwd_launch(AttnArgs a, int B, cudaStream_t s) {
  AttnArgs f = a;
  size_t fsmem = 0;
  if (a.T <= 160 && fused_bwd_enabled() && fused_plan<DHP>(f, &fsmem)) {  // one CTA per (sample, head)
    static size_t configured = 0;
    if (fsmem > configured) { V4H_TRY(set_smem(attn_bwd_fused_umma_kernel<DHP>, fsmem)); configured = fsmem; }
    FusedMaps tm;
    memset(&tm, 0, sizeof(tm));
    f.use_tma = (attn_tma_enabled() && a.dh == DHP) ? 1 : 0;
    f.pf_stride = sm_count();
    if (f.use_tma) {
      const size_t ldq = (size_t)3 * a.H * a.dh, ldo = (size_t)a.H * a.dh;
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, MT, &tm.q));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, f.BN, &tm.kv));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, TAILR, &tm.q2));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, MT, &tm.d_o));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, TAILR, &tm.d_o2));
      V4H_TRY(make_w_map(a.o, B, a.T, a.H * a.dh, ldo, MT, &tm.o));
      V4H_TRY(make_w_map(a.o, B, a.T, a.H * a.dh, ldo, TAILR, &tm.o2));
    }
    V4H_CUDA(launch_pdl(attn_bwd_fused_umma_kernel<DHP>, dim3((unsigned)(B * a.H)), dim3(FUSED_THREADS), fsmem, s, tm, f));
    V4H_LAUNCH_CHECK();
    return V4H_OK;
  }
  dim3 grid((unsigned)ceil_div(a.T, MT), (unsigned)(B * a.H));
  const int use_tma = (attn_tma_enabled() && a.dh == DHP) ? 1 : 0;
  const size_t ldq = (size_t)3 * a.H * a.dh, ldo = (size_t)a.H * a.dh;
  if (use_tma && a.T > 160 && pipe_enabled()) {  // software-pipelined dQ and dK / dV
    const int want = pipe_stages();
    const char* skip = getenv("V4H_ATTN_SKIP");  // timing of one of the two kernels alone (results are then incomplete)
    if (!(skip && skip[0] == 'q')) {
      AttnArgs q = a;
      size_t smem = 0;
      // dQ: Q, dO tiles + NS x (K, V) blocks + dS; TMEM 2 x (S, dP) + dQ
      V4H_REQUIRE(pipe_plan(a.T, 1, DHP, 2048 + 2 * (size_t)w_bytes(MT, DHP), 5 * (size_t)w_bytes(1, DHP), 1, &q, &smem),
                  "attention: no pipelined dQ plan for T = %d", a.T);
      q.use_tma = 1;
      static size_t configured = 0;
      if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dq_pipe_kernel<DHP>, smem)); configured = smem; }
      CUtensorMap mq, mkv, mdo;
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, MT, &mq));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, q.BN, &mkv));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, MT, &mdo));
      V4H_CUDA(launch_pdl(attn_bwd_dq_pipe_kernel<DHP>, dim3(grid), dim3(PIPE_THREADS), smem, s, mq, mkv, mdo, q));
      V4H_LAUNCH_CHECK();
    }
    if (!(skip && skip[0] == 'k')) {
      AttnArgs k = a;
      size_t smem = 0;
      // dK / dV: K, V tiles + NS x (Q, dO) blocks + P^T, dS^T; TMEM 2 x (S^T, dP^T) + dV + dK
      V4H_REQUIRE(pipe_plan(a.T, want, 2 * DHP, 4096 + 2 * (size_t)w_bytes(MT, DHP), 2 * (size_t)w_bytes(1, DHP), 2, &k, &smem),
                  "attention: no pipelined dK / dV plan for T = %d", a.T);
      k.use_tma = 1;
      smem += 128 + (size_t)k.nblocks * k.BN * 8;  // lse and delta of all queries
      V4H_REQUIRE(smem <= 227 * 1024, "attention: T = %d is too long for the pipelined dK / dV kernel", a.T);
      static size_t configured = 0;
      if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dkv_pipe_kernel<DHP>, smem)); configured = smem; }
      CUtensorMap mk, mq, mdo;
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, MT, &mk));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, k.BN, &mq));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, k.BN, &mdo));
      V4H_CUDA(launch_pdl(attn_bwd_dkv_pipe_kernel<DHP>, dim3(grid), dim3(PIPE_THREADS), smem, s, mk, mq, mdo, k));
      V4H_LAUNCH_CHECK();
    }
    return V4H_OK;
  }
  {  // dQ + delta
    AttnArgs q = a;
    const int cap = env_cap("V4H_ATTN_DQ_CAP", a.T, std::min(160, (512 - DHP) / 2) / 16 * 16);
    pick_block(a.T, cap, &q.BN, &q.nblocks);
    q.tmem_cols = pow2_cols(2 * q.BN + DHP);
    q.use_tma = use_tma;
    const size_t smem = 2048 + 2 * w_bytes(MT, DHP) + 2 * w_bytes(q.BN, DHP) + g8_bytes(MT, q.BN) + 16;
    static size_t configured = 0;
    if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dq_umma_kernel<DHP>, smem)); configured = smem; }
    CUtensorMap mq, mkv, mdo;
    memset(&mq, 0, sizeof(mq)); memset(&mkv, 0, sizeof(mkv)); memset(&mdo, 0, sizeof(mdo));
    if (use_tma) {
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, MT, &mq));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, q.BN, &mkv));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, MT, &mdo));
    }
    V4H_CUDA(launch_pdl(attn_bwd_dq_umma_kernel<DHP>, dim3(grid), dim3(ATT_THREADS), smem, s, mq, mkv, mdo, q));
    V4H_LAUNCH_CHECK();
  }
  {  // dK, dV: K, V tiles of 128 keys; Q / dO stream through in blocks of at most 128 queries (the two [key][q]
     // tiles of P^T and dS^T plus four operand tiles have to fit the 227 KB)
    AttnArgs k = a;
    const int cap = env_cap("V4H_ATTN_DKV_CAP", a.T, std::min(128, (512 - 2 * DHP) / 2) / 16 * 16);
    pick_block(a.T, cap, &k.BN, &k.nblocks);
    k.tmem_cols = pow2_cols(2 * k.BN + 2 * DHP);
    k.use_tma = use_tma;
    const size_t smem = 4096 + 2 * w_bytes(MT, DHP) + 2 * w_bytes(k.BN, DHP) + 2 * g8_bytes(MT, k.BN) + 16;
    static size_t configured = 0;
    if (smem > configured) { V4H_TRY(set_smem(attn_bwd_dkv_umma_kernel<DHP>, smem)); configured = smem; }
    CUtensorMap mk, mq, mdo;
    memset(&mk, 0, sizeof(mk)); memset(&mq, 0, sizeof(mq)); memset(&mdo, 0, sizeof(mdo));
    if (use_tma) {
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, MT, &mk));
      V4H_TRY(make_w_map(a.qkv, B, a.T, 3 * a.H * a.dh, ldq, k.BN, &mq));
      V4H_TRY(make_w_map(a.d_o, B, a.T, a.H * a.dh, ldo, k.BN, &mdo));
    }
    V4H_CUDA(launch_pdl(attn_bwd_dkv_umma_kernel<DHP>, dim3(grid), dim3(ATT_THREADS), smem, s, mk, mq, mdo, k));
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
