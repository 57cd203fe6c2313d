// tcgen05 / TMEM / TMA GEMM for sm_100a: bf16 operands, fp32 accumulation in tensor memory, the
// shared epilogues of common.cuh applied straight out of TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer   (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier ring of 3-4 stages)
//   warp 1    MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M = 128, N = bn <= 256, K = 16)
//   warp 2    TMEM allocator (512 columns = two accumulator stages of up to 256 columns)
//   warp 3    epilogue-input producer: TMA loads of the tile-shaped epilogue operand (the fp32 residual
//             stream of EPI_GATE_RES, the saved pre-activation of EPI_DACT) into a ring of 32-column
//             slabs, running ahead of the epilogue
//   warps 4-7 epilogue       (tcgen05.ld 32x32b, one output row per thread, overlaps the next tile's MMAs)
//
// Epilogue data movement ("slab" mode): results are produced 32 columns at a time into swizzled
// shared-memory boxes of [128 rows][32 columns] and written with TMA stores (cp.async.bulk.tensor
// shared -> global), so every global write is a full row segment and ragged M / N edges are clipped
// by the tensor map; the residual / pre-activation inputs arrive the same way and are transformed in
// place.  A 16-column tail (bn % 32 == 16), and every case the slab mode does not cover (split-K
// atomics, addends, misaligned buffers), use the direct per-thread path (epilogue_run).
//
// All three layouts of kernels.cuh run on the same kernel: an operand is either K-major (row = M/N
// index, K contiguous: forward activations and weights) or MN-major (row = K index, M/N contiguous:
// the second operand of dgrad, both operands of wgrad), selected per operand in the instruction
// descriptor and in how the tile is fetched (one [rows x 64k] box vs. several [64k x 64mn] boxes).
// Ragged edges in M, N and K rely on TMA zero fill; stores are bounds-checked.
// Split-K work items (EPI_ATOMIC) cover disjoint K ranges of one output tile.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "kernels.cuh"
#include "umma.cuh"

namespace v4h {

using namespace sm100;

namespace {

constexpr int BM = 128;        // UMMA M
constexpr int BK = 64;         // bf16 elements per stage along K = one 128-byte swizzle row
constexpr int MAX_BN = 256;    // UMMA N limit
constexpr int MAX_STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = MAX_BN * BK * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 48 KB
constexpr int CHUNK_BYTES = 64 * BK * 2;       // one [64 k][64 mn] box of an MN-major operand (8 KB)
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr int SLAB = 32;                       // epilogue slab width in columns
constexpr int MAX_SLOTS = 6;                   // epilogue-input ring
constexpr int BAR_BYTES = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

struct UmmaArgs {
  int M, N, K;
  int bn;                 // UMMA N of this launch (multiple of 16)
  int tiles_m, tiles_n, splits;
  int kblocks;            // ceil(K / BK)
  int kblocks_per_split;
  int a_mn, b_mn;         // 1 = MN-major operand
  uint32_t idesc;
  int stages;             // smem ring depth of the mainloop
  int nslots;             // epilogue-input ring depth (slab mode with a tile-shaped input)
  int has_out2;
};

// byte offset of 16-byte chunk c of row r inside a [128][32]-element box written / read by TMA
__device__ __forceinline__ uint32_t box_off(int r, int c, int esize) {
  return esize == 4 ? (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4))          // 128-byte rows, SWIZZLE_128B
                    : (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));   // 64-byte rows, SWIZZLE_64B
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float4 lds4(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(uint8_t* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void unpack8(uint4 t, float* v) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}
// 32 values of row r -> box (TOut = float: 8 chunks of 4; bf16: 4 chunks of 8)
template <typename TOut>
__device__ __forceinline__ void box_write(uint8_t* box, int r, const float (&v)[SLAB]) {
  if (sizeof(TOut) == 4) {
#pragma unroll
    for (int c = 0; c < 8; ++c) sts4(box + box_off(r, c, 4), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(box + box_off(r, c, 2)) = pack8(&v[8 * c]);
  }
}

// 32 fp32 values from global (bias, gate); guarded scalar path at the N edge / when not 16-byte aligned
__device__ __forceinline__ void load32(const float* __restrict__ p, int nvalid, bool vec, float (&v)[SLAB]) {
  if (vec && nvalid >= SLAB) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c);
      v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SLAB; ++i) v[i] = i < nvalid ? p[i] : 0.f;
  }
}

template <int EPI, int ACT, typename TOut, bool SLABMODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmOut2, const UmmaArgs g, const EpiParams ep) {
  constexpr bool HAS_IN = SLABMODE && (EPI == EPI_GATE_RES || EPI == EPI_DACT);
  constexpr int IN_ESIZE = EPI == EPI_GATE_RES ? 4 : (int)sizeof(TOut);
  constexpr int IN_BOX = BM * SLAB * IN_ESIZE;
  constexpr int OUT_BOX = BM * SLAB * (int)sizeof(TOut);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* pool = smem + g.stages * STAGE_BYTES;  // epilogue boxes (1024-byte aligned: STAGE_BYTES is)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (SMEM_LIMIT - BAR_BYTES));  // fixed place at the end
  uint64_t* full = bars;                        // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty = full + MAX_STAGES;          // [MAX_STAGES]  MMA -> TMA
  uint64_t* acc_full = empty + MAX_STAGES;      // [2]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;           // [2]  epilogue -> MMA
  uint64_t* in_full = acc_empty + 2;            // [MAX_SLOTS]  epilogue-input TMA -> epilogue
  uint64_t* in_empty = in_full + MAX_SLOTS;     // [MAX_SLOTS]  epilogue -> epilogue-input TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_empty + MAX_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int STAGES = g.stages;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (SLABMODE) {
      prefetch_tensormap(&tmOut);
      if (HAS_IN) prefetch_tensormap(&tmIn);
      if (g.has_out2) prefetch_tensormap(&tmOut2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = g.tiles_m * g.tiles_n * g.splits;
  const uint32_t b_tile_bytes = g.b_mn ? (uint32_t)((g.bn + 63) / 64) * CHUNK_BYTES : (uint32_t)g.bn * BK * 2;
  const int nslab = SLABMODE ? g.bn / SLAB : 0;  // full slabs per tile

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int split = work % g.splits;
        const int tile = work / g.splits;
        const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * g.bn;
        const int kb0 = split * g.kblocks_per_split;
        const int kb1 = min(g.kblocks, kb0 + g.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full[stage], A_BYTES + b_tile_bytes);
          if (!g.a_mn) {
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
          } else {
            tma_load_2d(sa, &tmA, &full[stage], m0, kb * BK);
            tma_load_2d(sa + CHUNK_BYTES, &tmA, &full[stage], m0 + 64, kb * BK);
          }
          if (!g.b_mn) {
            tma_load_2d(sb, &tmB, &full[stage], kb * BK, n0);
          } else {
            for (int j = 0; j * 64 < g.bn; ++j)
              tma_load_2d(sb + j * CHUNK_BYTES, &tmB, &full[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int split = work % g.splits;
      const int kb0 = split * g.kblocks_per_split;
      const int kb1 = min(g.kblocks, kb0 + g.kblocks_per_split);
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * MAX_BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          // number of K=16 steps with any in-range data in this block (TMA zero-fills the rest)
          const int ksteps = min(BK / 16, (g.K - kb * BK + 15) / 16);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc = g.a_mn ? make_smem_desc(sa + k * 16 * 128, BK * 128, 1024)
                                          : make_smem_desc(sa + k * 32, 0, 1024);
            const uint64_t bdesc = g.b_mn ? make_smem_desc(sb + k * 16 * 128, BK * 128, 1024)
                                          : make_smem_desc(sb + k * 32, 0, 1024);
            umma_bf16(d_tmem, adesc, bdesc, g.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                    // frees the smem slot when the MMAs have read it
          if (kb == kb1 - 1) umma_commit(&acc_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == 3) {
    // ===================================================================== epilogue-input producer
    if (HAS_IN && lane == 0) {
      int slot = 0; uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int tile = work / g.splits;
        const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * g.bn;
        for (int j = 0; j < nslab; ++j) {
          mbar_wait(&in_empty[slot], phase ^ 1);
          mbar_expect_tx(&in_full[slot], IN_BOX);
          tma_load_2d(pool + slot * IN_BOX, &tmIn, &in_full[slot], n0 + j * SLAB, m0);
          if (++slot == g.nslots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;       // row inside the tile = TMEM lane
    const bool leader = threadIdx.x == 128;
    int acc = 0; uint32_t acc_phase = 0;
    int slot = 0; uint32_t in_phase = 0;   // epilogue-input ring position
    int prev_slot = -1;                    // ring slot whose TMA store was issued last
    uint32_t it = 0;                       // slab counter: staging buffer = it & 1
    // pool layout: [ring: nslots x IN_BOX] [staging out: 2 x OUT_BOX] [staging out2: 2 x OUT_BOX]
    uint8_t* ring = pool;
    uint8_t* st_out = pool + (HAS_IN ? g.nslots * IN_BOX : 0);
    uint8_t* st_out2 = st_out + ((EPI == EPI_BIAS_ACT) ? 2 * OUT_BOX : 0);
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int tile = work / g.splits;
      const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * g.bn;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + r;
      const uint32_t t_row = tmem_base + (uint32_t)acc * MAX_BN + ((uint32_t)(q * 32) << 16);
      if (SLABMODE) {
        for (int j = 0; j < nslab; ++j, ++it) {
          const int col0 = n0 + j * SLAB;
          const int nvalid = g.N - col0;  // > 0 for every launched tile column... may be < SLAB at the N edge
          float b32[SLAB], gt[SLAB];
          if (EPI != EPI_DACT) {
            if (ep.bias) load32(ep.bias + col0, nvalid, ep.vec_ok, b32);
            else {
#pragma unroll
              for (int i = 0; i < SLAB; ++i) b32[i] = 0.f;
            }
          }
          if (EPI == EPI_GATE_RES) {
            const int rr = min(row, g.M - 1);
            load32(ep.gate + (size_t)(rr / ep.rows_per_sample) * ep.mod_stride + col0, nvalid, ep.vec_ok, gt);
          }
          float v[SLAB];
          {
            float lo[16], hi[16];
            tmem_ld16(t_row + j * SLAB, lo);
            tmem_ld16(t_row + j * SLAB + 16, hi);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = hi[i]; }
          }
          if (HAS_IN) mbar_wait(&in_full[slot], in_phase);
          uint8_t* so = st_out + (it & 1) * OUT_BOX;
          if (EPI == EPI_BIAS_ACT) {
#pragma unroll
            for (int i = 0; i < SLAB; ++i) v[i] += b32[i];
            if (g.has_out2) box_write<TOut>(st_out2 + (it & 1) * OUT_BOX, r, v);
            if (ACT != ACT_NONE) {
#pragma unroll
              for (int i = 0; i < SLAB; ++i) v[i] = act_f<ACT>(v[i]);
            }
            box_write<TOut>(so, r, v);
          } else if (EPI == EPI_GATE_RES) {
            uint8_t* box = ring + slot * IN_BOX;
#pragma unroll
            for (int i = 0; i < SLAB; ++i) v[i] += b32[i];                 // y
            if (g.has_out2) box_write<TOut>(so, r, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint8_t* p = box + box_off(r, c, 4);
              float4 x = lds4(p);
              x.x += gt[4 * c] * v[4 * c]; x.y += gt[4 * c + 1] * v[4 * c + 1];
              x.z += gt[4 * c + 2] * v[4 * c + 2]; x.w += gt[4 * c + 3] * v[4 * c + 3];
              sts4(p, x);
            }
          } else {  // EPI_DACT, in place on the pre-activation box
            uint8_t* box = ring + slot * IN_BOX;
            if (sizeof(TOut) == 2) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint4* p = reinterpret_cast<uint4*>(box + box_off(r, c, 2));
                float u[8];
                unpack8(*p, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = v[8 * c + i] * dact_f<ACT>(u[i]);
                *p = pack8(u);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint8_t* p = box + box_off(r, c, 4);
                float4 x = lds4(p);
                x.x = v[4 * c] * dact_f<ACT>(x.x); x.y = v[4 * c + 1] * dact_f<ACT>(x.y);
                x.z = v[4 * c + 2] * dact_f<ACT>(x.z); x.w = v[4 * c + 3] * dact_f<ACT>(x.w);
                sts4(p, x);
              }
            }
          }
          fence_proxy_async();  // this thread's smem writes -> visible to the TMA store
          if (leader) {
            // the stores of the previous slab have finished READING shared memory: its staging buffers
            // (reused by the next slab) and its ring slot are free after the barrier below
            tma_store_wait_read0();
            if (HAS_IN && prev_slot >= 0) mbar_arrive(&in_empty[prev_slot]);
          }
          epi_bar_sync();
          if (leader) {
            if (EPI == EPI_BIAS_ACT) {
              tma_store_2d(&tmOut, so, col0, m0);
              if (g.has_out2) tma_store_2d(&tmOut2, st_out2 + (it & 1) * OUT_BOX, col0, m0);
            } else if (EPI == EPI_GATE_RES) {
              tma_store_2d(&tmOut, ring + slot * IN_BOX, col0, m0);   // residual stream out
              if (g.has_out2) tma_store_2d(&tmOut2, so, col0, m0);    // y
            } else {
              tma_store_2d(&tmOut, ring + slot * IN_BOX, col0, m0);
            }
            tma_store_commit();
            prev_slot = slot;
          }
          if (HAS_IN && ++slot == g.nslots) { slot = 0; in_phase ^= 1; }
        }
      }
      // direct path: everything (no slab mode) or the 16-column tail
      for (int c0 = nslab * SLAB; c0 < g.bn; c0 += 16) {
        float v[16];
        tmem_ld16(t_row + c0, v);
        tmem_ld_wait();
        const int col0 = n0 + c0;
        if (row < g.M && col0 < g.N) epilogue_run<EPI, ACT, TOut, 16>(ep, row, col0, min(16, g.N - col0), v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (SLABMODE && leader) tma_store_wait_read0();  // shared memory must outlive the last store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap a, b, in, out, out2; };

template <int EPI, int ACT, typename TOut, bool SLABMODE>
int launch(const Maps& m, const UmmaArgs& g, const EpiParams& ep, int grid, cudaStream_t s) {
  static bool configured = false;  // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    V4H_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<EPI, ACT, TOut, SLABMODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  gemm_umma_kernel<EPI, ACT, TOut, SLABMODE><<<grid, THREADS, SMEM_LIMIT, s>>>(m.a, m.b, m.in, m.out, m.out2, g, ep);
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}
template <int EPI, int ACT, typename TOut>
int launch2(bool slab, const Maps& m, const UmmaArgs& g, const EpiParams& ep, int grid, cudaStream_t s) {
  return slab ? launch<EPI, ACT, TOut, true>(m, g, ep, grid, s) : launch<EPI, ACT, TOut, false>(m, g, ep, grid, s);
}

}  // namespace

struct UmmaContext {
  EncodeTiledFn encode = nullptr;
  int num_sms = 148;
  std::mutex mu;
  // (base pointer, inner extent, outer extent, row pitch in elements, box inner, box outer, element bytes) -> map
  std::map<std::tuple<const void*, int, int, int, int, int, int>, CUtensorMap> cache;
};

UmmaContext* umma_context_create() {
  UmmaContext* c = new UmmaContext();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) c->num_sms = n;
  }
  return c;
}

void umma_context_destroy(UmmaContext* c) { delete c; }

bool gemm_umma_supported(const GemmDesc& g) {
  if (g.a_dtype != DT_BF16 || g.b_dtype != DT_BF16) return false;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return false;
  if ((g.lda % 8) || (g.ldb % 8)) return false;  // TMA: row pitch must be a multiple of 16 bytes
  if (g.epi == EPI_ATOMIC ? g.out_dtype != DT_F32 : false) return false;
  return true;
}

// esize 2: bf16 operand / epilogue boxes; esize 4: fp32 epilogue boxes.  The swizzle follows the box
// row length: 128-byte rows -> SWIZZLE_128B, 64-byte rows -> SWIZZLE_64B.
static int get_map(UmmaContext* ctx, const void* base, int inner, int outer, int pitch, int box_inner, int box_outer,
                   int esize, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto key = std::make_tuple(base, inner, outer, pitch, box_inner, box_outer, esize);
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) { *out = it->second; return V4H_OK; }
  if (!ctx->encode) return fail(V4H_ERR_CUDA, "gemm_umma: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const int row_bytes = box_inner * esize;
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap m;
  CUresult r = ctx->encode(&m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(V4H_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for base %p dims (%d, %d) pitch %d box (%d, %d)", (int)r,
                base, inner, outer, pitch, box_inner, box_outer);
  if (ctx->cache.size() > 4096) ctx->cache.clear();
  ctx->cache[key] = m;
  *out = m;
  return V4H_OK;
}

// largest UMMA N (multiple of 16, <= 256) that wastes the fewest columns; ties go to the wider tile
static int choose_bn(int N) {
  if (N <= MAX_BN) return (int)ceil_div(N, 16) * 16;
  int best = MAX_BN; long best_waste = -1;
  for (int bn = MAX_BN; bn >= 128; bn -= 16) {
    const long waste = ceil_div(N, bn) * bn - N;
    if (best_waste < 0 || waste < best_waste) { best = bn; best_waste = waste; }
  }
  return best;
}

// can the epilogue move its tile-shaped streams with TMA boxes?
static bool slab_ok(const GemmDesc& d, int bn) {
  if (d.epi == EPI_ATOMIC || bn < SLAB) return false;
  const EpiParams& p = d.ep;
  const int esz = d.out_dtype == DT_BF16 ? 2 : 4;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (p.addend) return false;
  if ((p.ldo * esz) % 16) return false;
  switch (d.epi) {
    case EPI_BIAS_ACT:
      return p.out && al(p.out) && al(p.out2);
    case EPI_GATE_RES:
      return d.out_dtype == DT_BF16 && p.res_in && p.res_out && p.gate && al(p.res_in) && al(p.res_out) &&
             al(p.out2) && (p.ldo * 4) % 16 == 0;
    case EPI_DACT:
      return p.out && p.aux && al(p.out) && al(p.aux) && (p.ld_aux * esz) % 16 == 0;
  }
  return false;
}

int gemm_umma(UmmaContext* ctx, const GemmDesc& d, cudaStream_t s) {
  V4H_REQUIRE(ctx != nullptr, "gemm_umma: no context");
  V4H_REQUIRE(gemm_umma_supported(d), "gemm_umma: unsupported operands (M=%d N=%d K=%d lda=%d ldb=%d)", d.M, d.N, d.K,
              d.lda, d.ldb);
  UmmaArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.a_mn = d.layout == GEMM_TN ? 1 : 0;
  g.b_mn = d.layout == GEMM_NT ? 0 : 1;
  g.bn = choose_bn(d.N);
  g.tiles_m = (int)ceil_div(d.M, BM);
  g.tiles_n = (int)ceil_div(d.N, g.bn);
  g.kblocks = (int)ceil_div(d.K, BK);
  int splits = d.epi == EPI_ATOMIC ? d.splitk : 1;
  if (d.epi == EPI_ATOMIC && splits <= 0) {
    // auto: the split count whose work items fill whole rounds of the persistent grid.  Cost model in
    // k-block units: every round of items costs the k-blocks of one item plus a fixed epilogue term
    // (the fp32 atomics of one output tile), so 150 items on 148 SMs (two rounds) lose to 120 items.
    const int tiles = g.tiles_m * g.tiles_n;
    const int epilogue_cost = 6;
    long best_cost = -1;
    splits = 1;
    for (int sp = 1; sp <= g.kblocks && (long)tiles * sp <= 4L * ctx->num_sms; ++sp) {
      const int kper = (int)ceil_div(g.kblocks, sp);
      const int eff = (int)ceil_div(g.kblocks, kper);
      const long rounds = ceil_div((long)tiles * eff, ctx->num_sms);
      const long cost = rounds * (kper + epilogue_cost);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
    }
  }
  if (splits < 1) splits = 1;
  if (splits > g.kblocks) splits = g.kblocks;
  g.kblocks_per_split = (int)ceil_div(g.kblocks, splits);
  g.splits = (int)ceil_div(g.kblocks, g.kblocks_per_split);
  g.idesc = make_idesc_bf16(BM, g.bn, g.a_mn != 0, g.b_mn != 0);

  Maps m;
  memset(&m, 0, sizeof(m));
  if (!g.a_mn) V4H_TRY(get_map(ctx, d.A, d.K, d.M, d.lda, BK, BM, 2, &m.a));      // A (M, K): box 64 k x 128 rows
  else         V4H_TRY(get_map(ctx, d.A, d.M, d.K, d.lda, 64, BK, 2, &m.a));      // A (K, M): box 64 m x 64 k rows
  if (!g.b_mn) V4H_TRY(get_map(ctx, d.B, d.K, d.N, d.ldb, BK, g.bn, 2, &m.b));    // B (N, K): box 64 k x bn rows
  else         V4H_TRY(get_map(ctx, d.B, d.N, d.K, d.ldb, 64, BK, 2, &m.b));      // B (K, N): box 64 n x 64 k rows

  const bool obf = d.out_dtype == DT_BF16;
  const int osz = obf ? 2 : 4;
  EpiParams ep = d.ep;
  ep.vec_ok = epilogue_vec_ok(ep, d.epi, d.epi == EPI_ATOMIC ? false : obf);
  const bool slab = slab_ok(d, g.bn);
  // shared-memory budget: [mainloop stages][epilogue pool]; barriers sit in the last BAR_BYTES
  int pool_bytes = 0;
  g.nslots = 0;
  g.has_out2 = ep.out2 != nullptr;
  if (slab) {
    const int out_box = BM * SLAB * osz;
    switch (d.epi) {
      case EPI_BIAS_ACT:
        pool_bytes = (g.has_out2 ? 4 : 2) * out_box;
        V4H_TRY(get_map(ctx, ep.out, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out));
        if (g.has_out2) V4H_TRY(get_map(ctx, ep.out2, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out2));
        break;
      case EPI_GATE_RES:
        g.nslots = 4;
        pool_bytes = g.nslots * BM * SLAB * 4 + 2 * out_box;
        V4H_TRY(get_map(ctx, ep.res_in, d.N, d.M, ep.ldo, SLAB, BM, 4, &m.in));
        V4H_TRY(get_map(ctx, ep.res_out, d.N, d.M, ep.ldo, SLAB, BM, 4, &m.out));
        if (g.has_out2) V4H_TRY(get_map(ctx, ep.out2, d.N, d.M, ep.ldo, SLAB, BM, 2, &m.out2));
        break;
      case EPI_DACT:
        g.nslots = MAX_SLOTS;
        pool_bytes = g.nslots * out_box;
        V4H_TRY(get_map(ctx, ep.aux, d.N, d.M, ep.ld_aux, SLAB, BM, osz, &m.in));
        V4H_TRY(get_map(ctx, ep.out, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out));
        break;
    }
  }
  g.stages = (SMEM_LIMIT - 1024 - BAR_BYTES - pool_bytes) / STAGE_BYTES;
  if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
  V4H_REQUIRE(g.stages >= 2, "gemm_umma: internal shared-memory budget error");

  const int total = g.tiles_m * g.tiles_n * g.splits;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  switch (d.epi) {
    case EPI_BIAS_ACT:
      if (d.act == ACT_NONE)
        return obf ? launch2<EPI_BIAS_ACT, ACT_NONE, bf16>(slab, m, g, ep, grid, s)
                   : launch2<EPI_BIAS_ACT, ACT_NONE, float>(slab, m, g, ep, grid, s);
      if (d.act == ACT_GELU_TANH && obf) return launch2<EPI_BIAS_ACT, ACT_GELU_TANH_FAST, bf16>(slab, m, g, ep, grid, s);
      if (d.act == ACT_SILU)
        return obf ? launch2<EPI_BIAS_ACT, ACT_SILU, bf16>(slab, m, g, ep, grid, s)
                   : launch2<EPI_BIAS_ACT, ACT_SILU, float>(slab, m, g, ep, grid, s);
      break;
    case EPI_GATE_RES:
      if (obf) return launch2<EPI_GATE_RES, ACT_NONE, bf16>(slab, m, g, ep, grid, s);
      break;
    case EPI_DACT:
      if (d.act == ACT_GELU_TANH && obf) return launch2<EPI_DACT, ACT_GELU_TANH_FAST, bf16>(slab, m, g, ep, grid, s);
      if (d.act == ACT_SILU && obf) return launch2<EPI_DACT, ACT_SILU, bf16>(slab, m, g, ep, grid, s);
      break;
    case EPI_ATOMIC:
      return launch<EPI_ATOMIC, ACT_NONE, float, false>(m, g, ep, grid, s);
  }
  return fail(V4H_ERR_UNSUPPORTED, "gemm_umma: epilogue %d / activation %d / output dtype %d is not instantiated", d.epi,
              d.act, d.out_dtype);
}

}  // namespace v4h
