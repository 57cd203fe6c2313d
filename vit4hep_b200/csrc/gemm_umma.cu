// tcgen05 / TMEM / TMA GEMM for sm_100a: bf16 operands, fp32 accumulation in tensor memory, the
// shared epilogues of common.cuh applied straight out of TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer of A (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier ring of 3-4 stages)
//   warp 1    MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M = 128, N = bn <= 256, K = 16)
//   warp 2    TMEM allocator (512 columns = two accumulator stages of up to 256 columns), then TMA producer of B
//   warp 3    epilogue-input producer: TMA loads of the tile-shaped epilogue operand (the fp32 residual
//             stream of EPI_GATE_RES, the saved pre-activation of EPI_DACT) into a ring of 32-column
//             slabs, running ahead of the epilogue
//   warps 4-7 epilogue       (tcgen05.ld 32x32b, one output row per thread, overlaps the next tile's MMAs)
//
// Epilogue data movement ("slab" mode): results are produced 32 columns at a time into swizzled
// shared-memory boxes of [128 rows][32 columns] (row-per-thread writes, conflict-free thanks to the
// TMA swizzle pattern) and copied out cooperatively, consecutive lanes taking consecutive 16-byte chunks
// of a row, so every global write is a full 64 / 128-byte row segment; the residual / pre-activation
// inputs arrive in the same boxes by TMA and are transformed in place.  (TMA stores were tried first:
// waiting for their shared-memory reads to retire serialised the slabs at about 1 us each.)  A 16-column tail (bn % 32 == 16), and every case the slab mode does not cover (split-K
// atomics, addends, misaligned buffers), use the direct per-thread path (epilogue_run).
//
// CTA pairs (CTAS = 2, used whenever M > 128): the two CTAs of a cluster share one 256-row tile with
// tcgen05.mma cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the B tile, so the bytes
// every SM pulls from L2 per flop drop by a third (this GEMM family is bound by the L2 -> shared-memory
// feed, not by the tensor pipe).  The leader CTA (cluster rank 0) owns the `full` barriers (both
// producers' TMA transactions are counted there) and issues the MMAs; tcgen05.commit multicasts the
// `empty` / `acc_full` arrivals to both CTAs; both epilogues arrive on the leader's `acc_empty`.
//
// All three layouts of kernels.cuh run on the same kernel: an operand is either K-major (row = M/N
// index, K contiguous: forward activations and weights) or MN-major (row = K index, M/N contiguous:
// the second operand of dgrad, both operands of wgrad), selected per operand in the instruction
// descriptor and in how the tile is fetched (one [rows x 64k] box vs. several [64k x 64mn] boxes).
// Ragged edges in M, N and K rely on TMA zero fill; stores are bounds-checked.
// Split-K work items (EPI_ATOMIC) cover disjoint K ranges of one output tile.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
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
constexpr int MAX_STAGES = 6;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = MAX_BN * BK * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 48 KB
constexpr int CHUNK_BYTES = 64 * BK * 2;       // one [64 k][64 mn] box of an MN-major operand (8 KB)
constexpr int EG = 2;                          // epilogue groups of 4 warps
constexpr int THREADS = 128 + 128 * EG;
constexpr int TMEM_COLS = 512;
constexpr int SLAB = 32;                       // epilogue slab width in columns
constexpr int MAX_SLOTS = 8;                   // epilogue-input ring
constexpr int BAR_BYTES = 512;
constexpr int SMEM_LIMIT = 227 * 1024;
// per-tile epilogue vectors staged in shared memory by each epilogue group (slab mode): the tile's bias
// [MAX_BN] and the gate rows of the (at most two) samples its 128 rows belong to [2][MAX_BN], fp32
constexpr int VEC_FLOATS = 3 * MAX_BN;
constexpr int VEC_BYTES = EG * VEC_FLOATS * 4;

struct UmmaArgs {
  int M, N, K;
  int bn;                 // UMMA N of the full-width column tiles (multiple of 16)
  int n_pad;              // N rounded up to the tile granularity: the last column tile is n_pad - (tiles_n-1)*bn wide
  int b_box_rows;         // rows of the K-major B box (TMA always transfers the whole box)
  int tiles_m, tiles_n, splits;
  int kblocks;            // ceil(K / BK)
  int kblocks_per_split;
  int a_mn, b_mn;         // 1 = MN-major operand
  uint32_t idesc;
  int stage_bytes;        // A tile + this CTA's part of the B tile, 1024-byte multiple
  int stages;             // smem ring depth of the mainloop
  int nslots;             // epilogue-input ring depth (slab mode with a tile-shaped input)
  int has_out2;
  int tma_store;          // slab mode: results leave through TMA stores (bias / activation / act' epilogues)
  int nbuf;               // staging buffers (or in-flight ring slots) per epilogue group in that mode
  long long* dbg;         // optional device counters (cycles per role / phase), see v4h_debug_gemm
  int stage_vecs;         // slab mode: bias / gate vectors of a tile staged in shared memory (V4H_GEMM_STAGE_VECS=0: per-slab global loads)
  int dbg_skip;           // V4H_GEMM_DBG_SKIP (experiments only): 1 = no TMEM reads, 2 = no staging / stores, 4 = no activation
};

// cycle accounting for v4h_debug_gemm: T.lap(slot) books the cycles since the previous lap
struct Lap {
  long long* dbg;
  long long t;
  long long acc[8];
  __device__ __forceinline__ explicit Lap(long long* d) : dbg(d), t(0) {
    if (dbg) { t = clock64(); for (int i = 0; i < 8; ++i) acc[i] = 0; }
  }
  __device__ __forceinline__ void lap(int i) {
    if (dbg) { const long long n = clock64(); acc[i] += n - t; t = n; }
  }
  __device__ __forceinline__ void flush(int base, int n) {
    if (dbg) for (int i = 0; i < n; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(dbg + base + i), (unsigned long long)acc[i]);
  }
};

// ---- cluster / cta_group::2 primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose mbarrier may live in the peer CTA of the pair (bar = shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

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
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void group_bar_sync(int grp) {  // the 128 threads of one epilogue group
  asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}

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

// Cooperative copy of one staged box to global memory by the 128 epilogue threads: consecutive lanes
// take consecutive 16-byte chunks of a row, so every row leaves as one contiguous 64 / 128-byte segment.
template <int ESIZE>
__device__ __forceinline__ void box_copy_out(const uint8_t* box, void* gbase, int ld, int m0, int col0, int M, int N,
                                             int t) {
  constexpr int CPR = SLAB * ESIZE / 16;  // chunks per row: 4 (bf16) or 8 (fp32)
  constexpr int EPC = 16 / ESIZE;         // elements per chunk
  uint8_t* g = reinterpret_cast<uint8_t*>(gbase);
#pragma unroll
  for (int i = 0; i < CPR; ++i) {
    const int idx = t + i * 128;
    const int row = idx / CPR, c = idx % CPR;
    if (m0 + row < M && col0 + c * EPC < N) {
      const uint4 v = *reinterpret_cast<const uint4*>(box + box_off(row, c, ESIZE));
      *reinterpret_cast<uint4*>(g + ((size_t)(m0 + row) * ld + col0 + c * EPC) * ESIZE) = v;
    }
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

template <int EPI, int ACT, typename TOut, bool SLABMODE, int CTAS>
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
  uint8_t* pool = smem + g.stages * g.stage_bytes;  // epilogue boxes (1024-byte aligned: stage_bytes is)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (SMEM_LIMIT - BAR_BYTES));  // fixed place at the end
  float* vecs = reinterpret_cast<float*>(smem_raw + (SMEM_LIMIT - BAR_BYTES - VEC_BYTES));  // just below them
  uint64_t* full = bars;                        // [MAX_STAGES]  TMA -> MMA
  uint64_t* empty = full + MAX_STAGES;          // [MAX_STAGES]  MMA -> TMA
  uint64_t* acc_full = empty + MAX_STAGES;      // [2]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;           // [2]  epilogue -> MMA
  uint64_t* in_full = acc_empty + 2;            // [MAX_SLOTS]  epilogue-input TMA -> epilogue
  uint64_t* in_empty = in_full + MAX_SLOTS;     // [MAX_SLOTS]  epilogue -> epilogue-input TMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_empty + MAX_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int STAGES = g.stages;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;  // CTA of the pair; 0 = leader
  const int unit = (int)blockIdx.x / CTAS, nunits = (int)gridDim.x / CTAS;  // persistent work index / stride

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (HAS_IN) prefetch_tensormap(&tmIn);
    if (SLABMODE && g.tma_store) {
      prefetch_tensormap(&tmOut);
      if (g.has_out2 && EPI == EPI_BIAS_ACT) prefetch_tensormap(&tmOut2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * EG * CTAS); }
    for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], g.tma_store ? 1 : 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTAS == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // barriers, TMEM and tensor-map prefetch above overlap the previous kernel's tail

  const int total_work = g.tiles_m * g.tiles_n * g.splits;
  // tile (tm, tn): rows [tm * BM * CTAS, ..), columns [tn * bn, tn * bn + width); the last column tile is
  // narrower (width a multiple of 16, of 32 in slab mode)
  auto tile_width = [&](int tn) { return min(g.bn, g.n_pad - tn * g.bn); };

  if (warp == 0 || warp == 2) {
    // ===================================================================== TMA producers
    // warp 0 stages A (and arms the stage's transaction count), warp 2 -- idle once TMEM is allocated --
    // stages B: issuing a TMA box costs ~100 cycles, and an MN-major B tile is up to four boxes per k-block
    const bool load_a = warp == 0;
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      Lap T(load_a ? g.dbg : nullptr);
      for (int work = unit; work < total_work; work += nunits) {
        const int split = work % g.splits;
        const int tile = work / g.splits;
        const int tn = tile / g.tiles_m;
        const int bnh = tile_width(tn) / CTAS;  // B rows / columns staged by this CTA
        const uint32_t b_part_bytes = g.b_mn ? (uint32_t)((bnh + 63) / 64) * CHUNK_BYTES : (uint32_t)g.b_box_rows * BK * 2;
        const int m0 = (tile % g.tiles_m) * (BM * CTAS) + (int)rank * BM;
        const int n0 = tn * g.bn + (int)rank * bnh;
        const int kb0 = split * g.kblocks_per_split;
        const int kb1 = min(g.kblocks, kb0 + g.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          T.lap(0);
          uint8_t* sa = smem + stage * g.stage_bytes;
          uint8_t* sb = sa + A_BYTES;
          if (CTAS == 1) {
            if (load_a) {
              mbar_expect_tx(&full[stage], A_BYTES + b_part_bytes);
              if (!g.a_mn) {
                tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
              } else {
                tma_load_2d(sa, &tmA, &full[stage], m0, kb * BK);
                tma_load_2d(sa + CHUNK_BYTES, &tmA, &full[stage], m0 + 64, kb * BK);
              }
            } else if (!g.b_mn) {
              tma_load_2d(sb, &tmB, &full[stage], kb * BK, n0);
            } else {
              for (int j = 0; j * 64 < bnh; ++j)
                tma_load_2d(sb + j * CHUNK_BYTES, &tmB, &full[stage], n0 + j * 64, kb * BK);
            }
          } else {
            // both CTAs' bytes are counted on the LEADER's full barrier
            const uint32_t bar = map_to_cta(smem_u32(&full[stage]), 0);
            if (load_a) {
              if (rank == 0) mbar_expect_tx(&full[stage], 2 * (A_BYTES + b_part_bytes));
              if (!g.a_mn) {
                tma_load_2d_pair(sa, &tmA, bar, kb * BK, m0);
              } else {
                tma_load_2d_pair(sa, &tmA, bar, m0, kb * BK);
                tma_load_2d_pair(sa + CHUNK_BYTES, &tmA, bar, m0 + 64, kb * BK);
              }
            } else if (!g.b_mn) {
              tma_load_2d_pair(sb, &tmB, bar, kb * BK, n0);
            } else {
              for (int j = 0; j * 64 < bnh; ++j)
                tma_load_2d_pair(sb + j * CHUNK_BYTES, &tmB, bar, n0 + j * 64, kb * BK);
            }
          }
          T.lap(1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      T.flush(0, 2);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // ONE thread runs the whole loop (pairs: of the leader CTA, for both CTAs).  Its instruction stream is a
    // dependent chain that shares a scheduler with two epilogue warps, so every instruction per k-block
    // counts: the shared-memory descriptors are linear in the address, base descriptors are built once and
    // a k-block costs a few 32-bit adds on their low words (the address field, bits 0-13, in 16-byte units)
    if (rank == 0 && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t sa0 = smem_u32(smem);
      const uint64_t adesc0 = g.a_mn ? make_smem_desc(sa0, BK * 128, 1024) : make_smem_desc(sa0, 0, 1024);
      const uint64_t bdesc0 = g.b_mn ? make_smem_desc(sa0 + A_BYTES, BK * 128, 1024) : make_smem_desc(sa0 + A_BYTES, 0, 1024);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t a_kstep = (g.a_mn ? 16u * 128u : 32u) >> 4, b_kstep = (g.b_mn ? 16u * 128u : 32u) >> 4;
      const uint32_t stage_step = (uint32_t)g.stage_bytes >> 4;
      uint32_t stage_off = 0;  // stage * stage_step
      Lap T(g.dbg);
      for (int work = unit; work < total_work; work += nunits) {
        const int split = work % g.splits;
        const int tile = work / g.splits;
        const uint32_t idesc = make_idesc_bf16(BM * CTAS, tile_width(tile / g.tiles_m), g.a_mn != 0, g.b_mn != 0);
        const int kb0 = split * g.kblocks_per_split;
        const int kb1 = min(g.kblocks, kb0 + g.kblocks_per_split);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        T.lap(0);
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * MAX_BN;
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          T.lap(1);
          const uint32_t a_lo = a_lo0 + stage_off, b_lo = b_lo0 + stage_off;
          // K = 16 steps with any in-range data in this block (TMA zero-fills the rest): 4 except at the K edge
          const int ksteps = min(BK / 16, (g.K - kb * BK + 15) / 16);
          if (ksteps == BK / 16) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + k * a_kstep), bd = ((uint64_t)b_hi << 32) | (b_lo + k * b_kstep);
              if (CTAS == 2) umma_bf16_pair(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
              else umma_bf16(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + k * a_kstep), bd = ((uint64_t)b_hi << 32) | (b_lo + k * b_kstep);
              if (CTAS == 2) umma_bf16_pair(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
              else umma_bf16(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
            }
          }
          accumulate = 1;
          if (CTAS == 2) {
            umma_commit_pair(&empty[stage]);                    // frees the smem slot in both CTAs
            if (kb == kb1 - 1) umma_commit_pair(&acc_full[acc]);  // accumulator complete, both epilogues
          } else {
            umma_commit(&empty[stage]);                    // frees the smem slot when the MMAs have read it
            if (kb == kb1 - 1) umma_commit(&acc_full[acc]);  // accumulator complete
          }
          T.lap(2);
          stage_off += stage_step;
          if (++stage == STAGES) { stage = 0; stage_off = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      T.flush(2, 3);
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================================== epilogue-input producer
    if (HAS_IN && lane == 0) {
      int slot = 0; uint32_t phase = 0;
      for (int work = unit; work < total_work; work += nunits) {
        const int tile = work / g.splits;
        const int tn = tile / g.tiles_m;
        const int m0 = (tile % g.tiles_m) * (BM * CTAS) + (int)rank * BM, n0 = tn * g.bn;
        const int nslab = tile_width(tn) / SLAB;
        for (int j = 0; j < nslab; ++j) {
          mbar_wait(&in_empty[slot], phase ^ 1);
          mbar_expect_tx(&in_full[slot], IN_BOX);
          tma_load_2d(pool + slot * IN_BOX, &tmIn, &in_full[slot], n0 + j * SLAB, m0);
          if (++slot == g.nslots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (EG groups of 4 warps)
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;   // epilogue group: takes the slabs j = grp (mod EG) of every tile
    const int r = q * 32 + lane;       // row inside the tile = TMEM lane
    int acc = 0; uint32_t acc_phase = 0;
    int slab_base = 0;                 // slabs of the tiles this CTA has finished (ring position)
    uint32_t it = 0;                   // slabs done by this group: staging buffer = it & 1
    // pool layout: [ring: nslots x IN_BOX] then per group [out: 2 x OUT_BOX] [out2: 2 x OUT_BOX]
    uint8_t* ring = pool;
    const int NB = g.tma_store ? g.nbuf : 2;  // staging buffers per output and group
    const int group_staging =
        (EPI == EPI_BIAS_ACT ? (g.has_out2 ? 2 : 1) * NB : (EPI == EPI_GATE_RES && !g.tma_store ? 2 : 0)) * OUT_BOX;
    uint8_t* st_out = pool + (HAS_IN ? g.nslots * IN_BOX : 0) + grp * group_staging;
    uint8_t* st_out2 = st_out + NB * OUT_BOX;
    const bool leader = (threadIdx.x & 127) == 0;  // first thread of the group: issues its TMA stores
    int hist[2] = {-1, -1};                        // ring slots of the group's previous two slabs
    const uint32_t acc_empty_leader[2] = {CTAS == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : 0u,
                                          CTAS == 2 ? map_to_cta(smem_u32(&acc_empty[1]), 0) : 0u};
    Lap T(threadIdx.x == 128 ? g.dbg : nullptr);
    for (int work = unit; work < total_work; work += nunits) {
      const int tile = work / g.splits;
      const int tn = tile / g.tiles_m;
      const int width = tile_width(tn);
      const int m0 = (tile % g.tiles_m) * (BM * CTAS) + (int)rank * BM, n0 = tn * g.bn;
      const int row = m0 + r;
      // Per-tile vectors -> shared memory, BEFORE the accumulator is waited for: the global-load latency of the
      // bias / gate values (the first thing every slab consumed: 8 % of the kernel's stall samples sat on it) is
      // paid once per tile, under the mainloop, and the slabs read them back as broadcast LDS.128
      float* bias_s = vecs + grp * VEC_FLOATS;
      float* gate_s = bias_s + MAX_BN;
      bool gate_staged = false;
      int b_lo = 0;
      if (SLABMODE && EPI != EPI_DACT && g.stage_vecs) {
        const int t128 = (int)threadIdx.x & 127;
        for (int i = t128; i < width; i += 128) bias_s[i] = (ep.bias != nullptr && n0 + i < g.N) ? __ldg(ep.bias + n0 + i) : 0.f;
        if (EPI == EPI_GATE_RES) {
          b_lo = min(m0, g.M - 1) / ep.rows_per_sample;
          const int b_hi = min(m0 + BM - 1, g.M - 1) / ep.rows_per_sample;
          gate_staged = b_hi - b_lo <= 1;
          if (gate_staged) {
            for (int i = t128; i < 2 * width; i += 128) {
              const int srow = i >= width ? 1 : 0, c = i - srow * width;
              gate_s[srow * MAX_BN + c] =
                  n0 + c < g.N ? __ldg(ep.gate + (size_t)min(b_lo + srow, b_hi) * ep.mod_stride + n0 + c) : 0.f;
            }
          }
        }
        group_bar_sync(grp);
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      T.lap(0);
      const uint32_t t_row = tmem_base + (uint32_t)acc * MAX_BN + ((uint32_t)(q * 32) << 16);
      if (SLABMODE) {
        const int nslab = width / SLAB;
        for (int j = grp; j < nslab; j += EG, ++it) {
          const int col0 = n0 + j * SLAB;
          const int nvalid = g.N - col0;  // may be < SLAB (even <= 0) at the N edge
          const int slot = HAS_IN ? (slab_base + j) % g.nslots : 0;
          const uint32_t in_phase = HAS_IN ? (uint32_t)(((slab_base + j) / g.nslots) & 1) : 0u;
          float b32[SLAB], gt[SLAB];
          if (EPI != EPI_DACT) {
            if (g.stage_vecs) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 t4 = *reinterpret_cast<const float4*>(bias_s + j * SLAB + 4 * c);
                b32[4 * c] = t4.x; b32[4 * c + 1] = t4.y; b32[4 * c + 2] = t4.z; b32[4 * c + 3] = t4.w;
              }
            } else if (ep.bias) {
              load32(ep.bias + col0, nvalid, ep.vec_ok, b32);
            } else {
#pragma unroll
              for (int i = 0; i < SLAB; ++i) b32[i] = 0.f;
            }
          }
          if (EPI == EPI_GATE_RES) {
            const int rr = min(row, g.M - 1);
            if (gate_staged) {
              const float* gp = gate_s + (rr / ep.rows_per_sample - b_lo) * MAX_BN + j * SLAB;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 t4 = *reinterpret_cast<const float4*>(gp + 4 * c);
                gt[4 * c] = t4.x; gt[4 * c + 1] = t4.y; gt[4 * c + 2] = t4.z; gt[4 * c + 3] = t4.w;
              }
            } else {  // more than two samples under one 128-row tile (rows_per_sample < 127)
              load32(ep.gate + (size_t)(rr / ep.rows_per_sample) * ep.mod_stride + col0, nvalid, ep.vec_ok, gt);
            }
          }
          float v[SLAB];
          if (!(g.dbg_skip & 1)) {
            float lo[16], hi[16];
            tmem_ld16(t_row + j * SLAB, lo);
            tmem_ld16(t_row + j * SLAB + 16, hi);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = hi[i]; }
          } else {
#pragma unroll
            for (int i = 0; i < SLAB; ++i) v[i] = (float)(i + r);
          }
          T.lap(1);
          if (HAS_IN) mbar_wait(&in_full[slot], in_phase);
          T.lap(2);
          const int sbuf = (int)(it % (uint32_t)NB);
          uint8_t* so = st_out + sbuf * OUT_BOX;
          if (EPI == EPI_BIAS_ACT) {
#pragma unroll
            for (int i = 0; i < SLAB; ++i) v[i] += b32[i];
            if (g.has_out2 && !(g.dbg_skip & 2)) box_write<TOut>(st_out2 + sbuf * OUT_BOX, r, v);
            if (ACT != ACT_NONE && !(g.dbg_skip & 4)) {
#pragma unroll
              for (int i = 0; i < SLAB; ++i) v[i] = act_f<ACT>(v[i]);
            }
            if (!(g.dbg_skip & 2)) box_write<TOut>(so, r, v);
            else if (v[0] == 123.456f) box_write<TOut>(so, r, v);  // keeps v alive
          } else if (EPI == EPI_GATE_RES) {
            uint8_t* box = ring + slot * IN_BOX;
#pragma unroll
            for (int i = 0; i < SLAB; ++i) v[i] += b32[i];                 // y
            if (g.has_out2) {
              if (!g.tma_store) {
                box_write<TOut>(so, r, v);
              } else if (row < g.M) {
                // TMA-store mode keeps no staging for y: the thread writes its row's 32 values directly (whole
                // 32-byte sectors; y is only read again by the backward)
                TOut* yrow = reinterpret_cast<TOut*>(ep.out2) + (size_t)row * ep.ldo + col0;
                if (nvalid >= SLAB && ep.vec_ok) {
                  if (sizeof(TOut) == 2) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) reinterpret_cast<uint4*>(yrow)[c] = pack8(&v[8 * c]);
                  } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                      reinterpret_cast<float4*>(yrow)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                  }
                } else {
                  for (int i = 0; i < SLAB; ++i)
                    if (i < nvalid) yrow[i] = from_f<TOut>(v[i]);
                }
              }
            }
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
          T.lap(3);
          if (g.tma_store) {
            // staged boxes complete -> TMA stores, asynchronous: the group goes on with the next slab while
            // they drain.  Before the barrier the leader makes sure the stores of NB - 1 slabs ago have
            // finished reading shared memory: their staging buffer is the one the NEXT slab writes, and
            // their ring slot (act' epilogue, transformed in place) goes back to its producer.
            fence_proxy_async();  // this thread's smem writes -> visible to the TMA store
            if (leader) {
              if (NB >= 3) tma_store_wait_read1(); else tma_store_wait_read0();
              const int done = NB >= 3 ? hist[1] : hist[0];
              if (HAS_IN && done >= 0) mbar_arrive(&in_empty[done]);
            }
            group_bar_sync(grp);
            T.lap(4);
            if (leader && !(g.dbg_skip & 2)) {
              if (EPI == EPI_BIAS_ACT) {
                tma_store_2d(&tmOut, so, col0, m0);
                if (g.has_out2) tma_store_2d(&tmOut2, st_out2 + sbuf * OUT_BOX, col0, m0);
              } else {  // gate+residual / act': the transformed input box itself
                tma_store_2d(&tmOut, ring + slot * IN_BOX, col0, m0);
              }
              tma_store_commit();
            }
            hist[1] = hist[0]; hist[0] = slot;
          } else {
            // staged boxes complete -> cooperative copy out.  Two staging buffers alternate per group; a
            // thread reaches the group's next barrier only after its reads of this one, so one barrier per
            // slab orders every reuse.
            group_bar_sync(grp);
            T.lap(4);
            if (EPI == EPI_BIAS_ACT) {
              box_copy_out<(int)sizeof(TOut)>(so, ep.out, ep.ldo, m0, col0, g.M, g.N, r);
              if (g.has_out2) box_copy_out<(int)sizeof(TOut)>(st_out2 + sbuf * OUT_BOX, ep.out2, ep.ldo, m0, col0, g.M, g.N, r);
            } else if (EPI == EPI_GATE_RES) {
              box_copy_out<4>(ring + slot * IN_BOX, ep.res_out, ep.ldo, m0, col0, g.M, g.N, r);
              if (g.has_out2) box_copy_out<(int)sizeof(TOut)>(so, ep.out2, ep.ldo, m0, col0, g.M, g.N, r);
            } else {
              box_copy_out<(int)sizeof(TOut)>(ring + slot * IN_BOX, ep.out, ep.ldo, m0, col0, g.M, g.N, r);
            }
            if (HAS_IN) {  // this warp has read its part of the ring slot: 4 arrivals hand it back to the producer
              __syncwarp();
              if (lane == 0) {
                fence_proxy_async();  // generic reads of the slot are ordered before its next TMA write
                mbar_arrive(&in_empty[slot]);
              }
            }
          }
          T.lap(5);
        }
        slab_base += nslab;
      } else {
        // direct path (split-K atomics, addends, misaligned buffers): the groups alternate 16-column chunks
        for (int c0 = grp * 16; c0 < width; c0 += 16 * EG) {
          float v[16];
          tmem_ld16(t_row + c0, v);
          tmem_ld_wait();
          const int col0 = n0 + c0;
          if (row < g.M && col0 < g.N) epilogue_run<EPI, ACT, TOut, 16>(ep, row, col0, min(16, g.N - col0), v);
        }
      }
      T.lap(6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(acc_empty_leader[acc]);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (SLABMODE && g.tma_store && leader) tma_store_wait_read0();  // smem must outlive the last stores' reads
    T.flush(5, 7);
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();  // the peer may still signal this CTA's barriers / read its tiles
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------
// Gated-residual GEMM with the following LayerNorm + adaLN modulation in its epilogue
// (reference nn/vit.py:331-332: x = x + gate * branch(modulate(norm(x), shift, scale)); the norm + modulate of the
// NEXT branch is applied here, where the new residual row is produced).
//
// One CTA per 128-row tile owns whole output rows -- CN = 1: all N columns in two accumulators (N = 480: 2 x 240
// TMEM columns); CN = 2: a 2-CTA cluster splits the columns (256 + 224) and exchanges the per-row partial sums
// through distributed shared memory -- so the row statistics never leave the SM.  Not persistent: after the
// mainloop the operand stages are free and hold the epilogue's boxes.
//   pass 1 (per 32-column chunk): y = acc + bias; x = res_in + gate * y (in place on the TMA-loaded residual box,
//           stored to res_out by TMA); x also goes BACK INTO TMEM over the accumulator (tcgen05.st) and into the
//           thread's running sum / sum of squares (a thread owns one row);
//   row statistics: two epilogue groups (+ the peer CTA) combine their partial sums;
//   pass 2: x from TMEM -> (x - mean) * rstd * (1 + scale) + shift -> bf16 box -> TMA store to ln_out.
// Warp roles as in gemm_umma_kernel: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator + residual-box producer,
// 4-11 two epilogue groups that alternate chunks.
// ------------------------------------------------------------------------------------------
constexpr int LN_MAX_SAMPLES = 3;   // samples a 128-row tile may span (rows_per_sample >= 64)
constexpr int LN_MAX_SLOTS = 16;
constexpr int LN_NBUF = 3;          // bf16 staging boxes per epilogue group (pass 2)

struct LnArgs {
  int M, N, K, kblocks;
  int col0[2], wc[2];   // per cluster rank: first column, width (multiple of 32)
  int a0w[2], a1w[2];   // accumulator widths (a1w = 0: one accumulator); a0w + a1w = wc
  int stage_bytes, stages;
  int nslots_pre, nslots_post;  // residual ring: dedicated slots / slots inside the freed operand stages
  int vec_w;            // floats per staged vector (max wc)
  int tmem_cols;
  int has_y, ld_ln;
  long long* dbg;       // optional cycle counters of the first epilogue thread (v4h_debug_gemm_ln)
};

__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void epi_bar_sync_all() {  // the 256 threads of both epilogue groups
  asm volatile("bar.sync 3, 256;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_n(uint32_t* slot, int cols) {
  if (cols <= COLS) tmem_alloc<COLS>(slot);
}

template <int CN, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
gemm_gate_res_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                        const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmIn,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmLn,
                        const LnArgs g, const EpiParams ep) {
  constexpr int IN_BOX = BM * SLAB * 4;   // fp32 residual box
  constexpr int LN_BOX = BM * SLAB * 2;   // bf16 output box
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_area = g.stages * g.stage_bytes;
  uint8_t* ring_pre = smem + stage_area;                       // dedicated residual slots
  uint8_t* ring_post = smem;                                   // residual slots inside the freed stages
  uint8_t* ln_stage = smem + stage_area - EG * LN_NBUF * LN_BOX;  // pass-2 staging, at the top of the stage area
  float* vecs = reinterpret_cast<float*>(ring_pre + g.nslots_pre * IN_BOX);
  float* bias_s = vecs;                                         // [vec_w]
  float* gate_s = bias_s + g.vec_w;                             // [LN_MAX_SAMPLES][vec_w]
  float* shift_s = gate_s + LN_MAX_SAMPLES * g.vec_w;           // [LN_MAX_SAMPLES][vec_w]
  float* scale_s = shift_s + LN_MAX_SAMPLES * g.vec_w;          // [LN_MAX_SAMPLES][vec_w]  holds 1 + scale
  float* part = scale_s + LN_MAX_SAMPLES * g.vec_w;             // [EG][BM][2] partial (sum, sum of squares)
  float* peer_part = part + EG * BM * 2;                        // [BM][2] written by the peer CTA (CN = 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (SMEM_LIMIT - BAR_BYTES));
  uint64_t* full = bars;                        // [MAX_STAGES]
  uint64_t* empty = full + MAX_STAGES;          // [MAX_STAGES]
  uint64_t* acc_full = empty + MAX_STAGES;      // [1]
  uint64_t* peer_bar = acc_full + 1;            // [1]
  uint64_t* in_full = peer_bar + 1;             // [LN_MAX_SLOTS]
  uint64_t* in_empty = in_full + LN_MAX_SLOTS;  // [LN_MAX_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_empty + LN_MAX_SLOTS);

  // CN = 2: the cluster splits the COLUMNS (rank = column half).  PAIR: the cluster is a cta_group::2 pair over
  // 256 ROWS -- each CTA stages its own 128 rows of A and HALF of every W tile, the leader issues M = 256 MMAs and
  // every CTA ends up with its 128 full rows in its own TMEM: 46 KB of operands per k-block and SM for 128 x 480
  // outputs instead of 76 KB (the mainloop is bound by the SM's L2 ingest)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CN == 2 || PAIR) ? cluster_ctarank() : 0u;
  const int tile = PAIR ? ((int)blockIdx.x / 2) * 2 + (int)rank : (int)blockIdx.x / CN;
  const int m0 = tile * BM;
  const uint32_t crank = PAIR ? 0u : rank;  // which column range this CTA owns
  const int col0 = g.col0[crank], wc = g.wc[crank], a0w = g.a0w[crank], a1w = g.a1w[crank];
  const int nch = wc / SLAB;
  const int nslots = g.nslots_pre + g.nslots_post;
  auto slot_of = [&](int j) { return j < g.nslots_pre ? j : g.nslots_pre + (j - g.nslots_pre) % g.nslots_post; };
  auto use_of = [&](int j) { return j < g.nslots_pre ? 0 : (j - g.nslots_pre) / g.nslots_post; };  // n-th use of its slot
  auto slot_ptr = [&](int sl) { return sl < g.nslots_pre ? ring_pre + sl * IN_BOX : ring_post + (sl - g.nslots_pre) * IN_BOX; };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA); prefetch_tensormap(&tmB0); prefetch_tensormap(&tmIn);
    prefetch_tensormap(&tmOut); prefetch_tensormap(&tmLn);
    if (CN == 2) prefetch_tensormap(&tmB1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(peer_bar, BM);
    for (int i = 0; i < LN_MAX_SLOTS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair<512>(tmem_slot);
    else if (g.tmem_cols > 256) tmem_alloc<512>(tmem_slot);
    else if (g.tmem_cols > 128) tmem_alloc<256>(tmem_slot);
    else if (g.tmem_cols > 64) tmem_alloc<128>(tmem_slot);
    else if (g.tmem_cols > 32) tmem_alloc<64>(tmem_slot);
    else tmem_alloc<32>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CN == 2 || PAIR) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer (A and this CTA's rows of W)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const CUtensorMap* tmB = (CN == 2 && rank == 1) ? &tmB1 : &tmB0;
      for (int kb = 0; kb < g.kblocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * g.stage_bytes;
        uint8_t* sb = sa + A_BYTES;
        if (PAIR) {
          // both CTAs' bytes are counted on the LEADER's full barrier; each stages its half of the W rows of
          // every accumulator
          const uint32_t bar = map_to_cta(smem_u32(&full[stage]), 0);
          const int h0 = a0w / 2, h1 = a1w / 2;
          if (rank == 0) mbar_expect_tx(&full[stage], 2u * (A_BYTES + (uint32_t)(h0 + h1) * BK * 2));
          tma_load_2d_pair(sa, &tmA, bar, kb * BK, m0);
          tma_load_2d_pair(sb, tmB, bar, kb * BK, (int)rank * h0);
          if (a1w) tma_load_2d_pair(sb + h0 * BK * 2, tmB, bar, kb * BK, a0w + (int)rank * h1);
        } else {
          mbar_expect_tx(&full[stage], A_BYTES + (uint32_t)wc * BK * 2);
          tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
          tma_load_2d(sb, tmB, &full[stage], kb * BK, col0);
          if (a1w) tma_load_2d(sb + a0w * BK * 2, tmB, &full[stage], kb * BK, col0 + a0w);
        }
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (PAIR: the leader, for both CTAs)
    if (lane == 0 && (!PAIR || rank == 0)) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t sa0 = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc(sa0, 0, 1024), bdesc0 = make_smem_desc(sa0 + A_BYTES, 0, 1024);
      const uint32_t a_hi = (uint32_t)(adesc0 >> 32), b_hi = (uint32_t)(bdesc0 >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0;
      const uint32_t b1_off = (uint32_t)((PAIR ? a0w / 2 : a0w) * BK * 2) >> 4;  // W rows of accumulator 0 held by this CTA
      const uint32_t idesc0 = make_idesc_bf16(PAIR ? 2 * BM : BM, a0w, false, false);
      const uint32_t idesc1 = make_idesc_bf16(PAIR ? 2 * BM : BM, a1w ? a1w : 16, false, false);
      const uint32_t stage_step = (uint32_t)g.stage_bytes >> 4;
      uint32_t stage_off = 0, accumulate = 0;
      for (int kb = 0; kb < g.kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const int ksteps = min(BK / 16, (g.K - kb * BK + 15) / 16);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo0 + stage_off + k * 2u);
          const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + stage_off + k * 2u);
          if (PAIR) {
            umma_bf16_pair(tmem_base, ad, bd, idesc0, k == 0 ? accumulate : 1u);
            if (a1w) umma_bf16_pair(tmem_base + 256u, ad, bd + b1_off, idesc1, k == 0 ? accumulate : 1u);
          } else {
            umma_bf16(tmem_base, ad, bd, idesc0, k == 0 ? accumulate : 1u);
            if (a1w) umma_bf16(tmem_base + 256u, ad, bd + b1_off, idesc1, k == 0 ? accumulate : 1u);
          }
        }
        accumulate = 1;
        if (PAIR) {
          umma_commit_pair(&empty[stage]);                       // frees the stage in both CTAs
          if (kb == g.kblocks - 1) umma_commit_pair(acc_full);  // both epilogues
        } else {
          umma_commit(&empty[stage]);
          if (kb == g.kblocks - 1) umma_commit(acc_full);
        }
        stage_off += stage_step;
        if (++stage == g.stages) { stage = 0; stage_off = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================================================== residual-box producer
    if (lane == 0) {
      bool waited_acc = false;
      for (int j = 0; j < nch; ++j) {
        const int sl = slot_of(j), use = use_of(j);
        if (j >= g.nslots_pre && !waited_acc) {  // the slot lives in the operand stages: the mainloop must be over
          mbar_wait(acc_full, 0);
          waited_acc = true;
        }
        if (use > 0) mbar_wait(&in_empty[sl], (uint32_t)(use - 1) & 1u);
        mbar_expect_tx(&in_full[sl], IN_BOX);
        tma_load_2d(slot_ptr(sl), &tmIn, &in_full[sl], col0 + j * SLAB, m0);
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int row = m0 + r;
    const int et = (int)threadIdx.x - 128;         // 0..255 over both groups
    const bool leader = (threadIdx.x & 127) == 0;
    // per-tile vectors -> shared memory while the mainloop runs
    const int rps = ep.rows_per_sample;
    const int b_lo = min(m0, g.M - 1) / rps, b_hi = min(m0 + BM - 1, g.M - 1) / rps;
    for (int i = et; i < wc; i += 256) bias_s[i] = ep.bias ? __ldg(ep.bias + col0 + i) : 0.f;
    for (int i = et; i < LN_MAX_SAMPLES * wc; i += 256) {
      const int sidx = i / wc, c = i - sidx * wc;
      const size_t off = (size_t)min(b_lo + sidx, b_hi) * ep.mod_stride + col0 + c;
      gate_s[sidx * g.vec_w + c] = __ldg(ep.gate + off);
      shift_s[sidx * g.vec_w + c] = __ldg(ep.ln_shift + off);
      scale_s[sidx * g.vec_w + c] = 1.f + __ldg(ep.ln_scale + off);
    }
    epi_bar_sync_all();
    const int sidx = min(row, g.M - 1) / rps - b_lo;   // which staged sample this thread's row belongs to
    const float* gate_r = gate_s + sidx * g.vec_w;
    const float* shift_r = shift_s + sidx * g.vec_w;
    const float* scale_r = scale_s + sidx * g.vec_w;
    Lap T(threadIdx.x == 128 ? g.dbg : nullptr);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    T.lap(0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    // TMEM column of output column c of this CTA (second accumulator starts at column 256)
    auto tcol = [&](int c) { return (uint32_t)(c < a0w ? c : 256 + (c - a0w)); };

    // ---------------- pass 1
    float s1 = 0.f, s2 = 0.f;
    int hist[2] = {-1, -1};
    for (int j = grp; j < nch; j += EG) {
      const int sl = slot_of(j);
      const int c0 = j * SLAB;
      float v[SLAB];
      {
        float lo[16], hi[16];
        tmem_ld16(t_row + tcol(c0), lo);
        tmem_ld16(t_row + tcol(c0 + 16), hi);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = hi[i]; }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * c);
        v[4 * c] += b4.x; v[4 * c + 1] += b4.y; v[4 * c + 2] += b4.z; v[4 * c + 3] += b4.w;
      }
      T.lap(1);
      if (g.has_y && row < g.M) {  // y = branch output, kept for the backward (d gate = sum dh * y)
        bf16* yrow = reinterpret_cast<bf16*>(ep.out2) + (size_t)row * ep.ldo + col0 + c0;
#pragma unroll
        for (int c = 0; c < 4; ++c) reinterpret_cast<uint4*>(yrow)[c] = pack8(&v[8 * c]);
      }
      mbar_wait(&in_full[sl], (uint32_t)use_of(j) & 1u);
      T.lap(2);
      uint8_t* box = slot_ptr(sl);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint8_t* pbox = box + box_off(r, c, 4);
        float4 x = lds4(pbox);
        const float4 g4 = *reinterpret_cast<const float4*>(gate_r + c0 + 4 * c);
        x.x += g4.x * v[4 * c]; x.y += g4.y * v[4 * c + 1]; x.z += g4.z * v[4 * c + 2]; x.w += g4.w * v[4 * c + 3];
        sts4(pbox, x);
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        s1 += (x.x + x.y) + (x.z + x.w);
        s2 += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
      }
      {  // the new residual row goes back into TMEM: pass 2 reads it from there
        float lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = v[i]; hi[i] = v[16 + i]; }
        tmem_st16(t_row + tcol(c0), lo);
        tmem_st16(t_row + tcol(c0 + 16), hi);
      }
      fence_proxy_async();
      T.lap(3);
      if (leader) {
        tma_store_wait_read1();                 // stores of two chunks ago have read their slot
        if (hist[1] >= 0) mbar_arrive(&in_empty[hist[1]]);
      }
      group_bar_sync(grp);
      if (leader) {
        tma_store_2d(&tmOut, box, col0 + c0, m0);
        tma_store_commit();
      }
      hist[1] = hist[0]; hist[0] = sl;
      T.lap(4);
    }
    tmem_st_wait();
    if (leader) {
      tma_store_wait_read0();
      if (hist[1] >= 0) mbar_arrive(&in_empty[hist[1]]);
      if (hist[0] >= 0) mbar_arrive(&in_empty[hist[0]]);
    }
    // ---------------- row statistics
    part[(grp * BM + r) * 2] = s1;
    part[(grp * BM + r) * 2 + 1] = s2;
    epi_bar_sync_all();
    float t1 = part[r * 2] + part[(BM + r) * 2], t2 = part[r * 2 + 1] + part[(BM + r) * 2 + 1];
    if (CN == 2) {
      if (grp == 0) {
        st_cluster_f2(map_to_cta(smem_u32(peer_part + r * 2), rank ^ 1u), t1, t2);
        mbar_arrive_cluster(map_to_cta(smem_u32(peer_bar), rank ^ 1u));  // release.cluster: orders the store above
      }
      {  // bounded like mbar_wait: a protocol bug traps instead of hanging the GPU
        uint32_t spins = 0;
        uint64_t t0 = 0;
        while (!mbar_try_wait_cluster(peer_bar, 0)) {
          if ((++spins & 4095u) == 0) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
          }
        }
      }
      t1 += peer_part[r * 2];
      t2 += peer_part[r * 2 + 1];
    }
    T.lap(5);
    const float inv_n = 1.f / (float)g.N;
    const float mean = t1 * inv_n;
    const float rstd = rsqrtf(fmaxf(t2 * inv_n - mean * mean, 0.f) + ep.ln_eps);
    if (ep.ln_stats != nullptr && grp == 0 && (PAIR || rank == 0) && row < g.M) ep.ln_stats[row] = make_float2(mean, rstd);
    // ---------------- pass 2
    uint8_t* stg = ln_stage + grp * LN_NBUF * LN_BOX;
    uint32_t it = 0;
    for (int j = grp; j < nch; j += EG, ++it) {
      const int c0 = j * SLAB;
      float v[SLAB];
      {
        float lo[16], hi[16];
        tmem_ld16(t_row + tcol(c0), lo);
        tmem_ld16(t_row + tcol(c0 + 16), hi);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = hi[i]; }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 sc4 = *reinterpret_cast<const float4*>(scale_r + c0 + 4 * c);
        const float4 sh4 = *reinterpret_cast<const float4*>(shift_r + c0 + 4 * c);
        v[4 * c] = fmaf((v[4 * c] - mean) * rstd, sc4.x, sh4.x);
        v[4 * c + 1] = fmaf((v[4 * c + 1] - mean) * rstd, sc4.y, sh4.y);
        v[4 * c + 2] = fmaf((v[4 * c + 2] - mean) * rstd, sc4.z, sh4.z);
        v[4 * c + 3] = fmaf((v[4 * c + 3] - mean) * rstd, sc4.w, sh4.w);
      }
      uint8_t* so = stg + (it % LN_NBUF) * LN_BOX;
      box_write<bf16>(so, r, v);
      fence_proxy_async();
      T.lap(6);
      if (leader) tma_store_wait_read1();
      group_bar_sync(grp);
      if (leader) {
        tma_store_2d(&tmLn, so, col0 + c0, m0);
        tma_store_commit();
      }
      T.lap(7);
    }
    // the "ones" column of a wider pitch (layernorm.cu): turns the weight-gradient GEMM into [dW | bias gradient]
    if (g.ld_ln > g.N && (PAIR || rank == CN - 1) && grp == 0 && row < g.M) {
      bf16* orow = reinterpret_cast<bf16*>(ep.ln_out) + (size_t)row * g.ld_ln + g.N;
      for (int i = 0; i < g.ld_ln - g.N; ++i) orow[i] = __float2bfloat16_rn(i == 0 ? 1.f : 0.f);
    }
    if (leader) tma_store_wait_read0();
    T.flush(0, 8);
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (CN == 2 || PAIR) cluster_sync_all();  // the peer may still write this CTA's partial sums / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<512>(tmem_base);
    else if (g.tmem_cols > 256) tmem_dealloc<512>(tmem_base);
    else if (g.tmem_cols > 128) tmem_dealloc<256>(tmem_base);
    else if (g.tmem_cols > 64) tmem_dealloc<128>(tmem_base);
    else if (g.tmem_cols > 32) tmem_dealloc<64>(tmem_base);
    else tmem_dealloc<32>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap a, b, in, out, out2; };

template <int EPI, int ACT, typename TOut, bool SLABMODE, int CTAS>
int launch_k(const Maps& m, const UmmaArgs& g, const EpiParams& ep, int grid, cudaStream_t s) {
  auto kernel = gemm_umma_kernel<EPI, ACT, TOut, SLABMODE, CTAS>;
  static bool configured = false;  // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    V4H_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  if (CTAS == 1) {
    V4H_CUDA(launch_pdl(kernel, dim3(grid), dim3(THREADS), SMEM_LIMIT, s, m.a, m.b, m.in, m.out, m.out2, g, ep));
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_LIMIT;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    V4H_CUDA(cudaLaunchKernelEx(&cfg, kernel, m.a, m.b, m.in, m.out, m.out2, g, ep));
  }
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}
template <int EPI, int ACT, typename TOut, bool SLABMODE>
int launch(const Maps& m, const UmmaArgs& g, const EpiParams& ep, int ctas, int grid, cudaStream_t s) {
  return ctas == 2 ? launch_k<EPI, ACT, TOut, SLABMODE, 2>(m, g, ep, grid, s)
                   : launch_k<EPI, ACT, TOut, SLABMODE, 1>(m, g, ep, grid, s);
}
template <int EPI, int ACT, typename TOut>
int launch2(bool slab, const Maps& m, const UmmaArgs& g, const EpiParams& ep, int ctas, int grid, cudaStream_t s) {
  return slab ? launch<EPI, ACT, TOut, true>(m, g, ep, ctas, grid, s)
              : launch<EPI, ACT, TOut, false>(m, g, ep, ctas, grid, s);
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
  // V4H_GEMM_SMS: persistent CTAs to launch (default: every SM).  A persistent CTA needs a whole SM; when another
  // kernel (an NCCL all-reduce under the backward) holds some SMs, the CTAs that do not fit start a second round.
  if (const char* e = getenv("V4H_GEMM_SMS")) {
    const int v = atoi(e);
    if (v >= 8 && v <= c->num_sms) c->num_sms = v;
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

// Column tiling: full tiles of 256 columns plus one narrower last tile, everything a multiple of `gran`
// (32 in slab mode: whole epilogue slabs; 16 otherwise: the UMMA N granularity at M = 128).
static void choose_tiling(int N, int gran, int max_bn, int* bn, int* n_pad, int* tiles_n) {
  *n_pad = (int)ceil_div(N, gran) * gran;
  *bn = *n_pad < max_bn ? *n_pad : max_bn;
  *tiles_n = (int)ceil_div(*n_pad, *bn);
}

// can the epilogue move its tile-shaped streams with TMA boxes?
static bool slab_ok(const GemmDesc& d) {
  if (d.epi == EPI_ATOMIC) return false;
  const EpiParams& p = d.ep;
  const int esz = d.out_dtype == DT_BF16 ? 2 : 4;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (p.addend) return false;
  if ((p.ldo * esz) % 16 || (d.N * esz) % 16) return false;  // rows leave as whole 16-byte chunks
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

// ------------------------------------------------------------------------------------------
// TMA feed probe (measurement only, v4h_debug_tma_probe): what a persistent CTA can pull through an
// mbarrier ring of `stages` stages of `boxes` [128 x 64] bf16 boxes (16 KB each, 128-byte swizzle) when
// nothing consumes the data -- the ceiling of the GEMM mainloop's operand feed for a given ring depth,
// number of CTAs and working set (L2-resident or streamed from HBM).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(160, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int stages, int boxes,
                                                           int box_rows, int producers, int iters, int tiles_cols,
                                                           int tiles_total, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // full[16], empty[16]
  uint64_t* full = bars;
  uint64_t* empty = bars + 16;
  uint8_t* ring = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
    prefetch_tensormap(&tm);
  }
  __syncthreads();
  const uint32_t box_bytes = (uint32_t)box_rows * 128u, stage_bytes = (uint32_t)boxes * box_bytes;
  if (warp < producers && lane == 0) {
    // producer warp w issues the boxes b = w (mod producers) of every stage; warp 0 arms the stage's byte count
    for (int it = 0; it < iters; ++it) {
      const int st = it % stages;
      if (it >= stages) mbar_wait(&empty[st], (uint32_t)((it / stages) - 1) & 1u);
      if (warp == 0) mbar_expect_tx(&full[st], stage_bytes);
      for (int b = warp; b < boxes; b += producers) {
        const long long tile = ((long long)blockIdx.x * iters + it) * boxes + b;
        const int t = (int)(tile % tiles_total);
        tma_load_2d(ring + (size_t)st * stage_bytes + (size_t)b * box_bytes, &tm, &full[st], (t % tiles_cols) * 64,
                    (t / tiles_cols) * box_rows);
      }
    }
  } else if (warp == 4 && lane == 0) {  // consumer: hands every stage straight back
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it % stages;
      mbar_wait(&full[st], (uint32_t)(it / stages) & 1u);
      mbar_arrive(&empty[st]);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace

int tma_probe(UmmaContext* ctx, const void* buf, int rows, int cols, int stages, int boxes, int box_rows, int producers,
              int iters, int ctas, long long* cycles, cudaStream_t s) {
  V4H_REQUIRE(ctx && buf && cycles, "tma_probe: null argument");
  V4H_REQUIRE((box_rows == 64 || box_rows == 128 || box_rows == 256) && rows % box_rows == 0 && cols % 64 == 0 &&
                  stages >= 1 && stages <= 16 && boxes >= 1 && producers >= 1 && producers <= 4 && iters >= 1 &&
                  ctas >= 1 && (size_t)stages * boxes * box_rows * 128 + 2048 <= (size_t)SMEM_LIMIT,
              "tma_probe: bad configuration");
  CUtensorMap tm;
  V4H_TRY(get_map(ctx, buf, cols, rows, cols, 64, box_rows, 2, &tm));
  const size_t smem = (size_t)stages * boxes * box_rows * 128 + 2048;
  V4H_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  // empty[] is armed with ONE arrival (the consumer's); every producer warp waits on it, none arrives
  tma_probe_kernel<<<ctas, 160, smem, s>>>(tm, stages, boxes, box_rows, producers, iters, cols / 64,
                                           (rows / box_rows) * (cols / 64), cycles);
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

int gemm_umma(UmmaContext* ctx, const GemmDesc& d, cudaStream_t s) {
  V4H_REQUIRE(ctx != nullptr, "gemm_umma: no context");
  V4H_REQUIRE(gemm_umma_supported(d), "gemm_umma: unsupported operands (M=%d N=%d K=%d lda=%d ldb=%d)", d.M, d.N, d.K,
              d.lda, d.ldb);
  UmmaArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.a_mn = d.layout == GEMM_TN ? 1 : 0;
  g.b_mn = d.layout == GEMM_NT ? 0 : 1;
  const bool slab = slab_ok(d);
  choose_tiling(d.N, slab ? SLAB : 16, MAX_BN, &g.bn, &g.n_pad, &g.tiles_n);
  // (Measured: halving the tile width of the one-round N = 480 GEMMs so that epilogues overlap the next
  // tile's MMAs is slower -- fc2 40 us vs 29 us -- the extra A traffic and N = 128 MMAs cost more.)
  // CTA pairs (cta_group::2, V4H_GEMM_PAIRS=1) when there is more than one 128-row tile of M.  Measured on
  // the ds2 shapes they are correct but 2-3 % slower than single CTAs (the epilogue, not the L2 -> smem
  // feed, bounds these K = 480 GEMMs), so they are opt-in.
  static const int pairs_enabled = [] { const char* e = getenv("V4H_GEMM_PAIRS"); return (e && e[0] == '1') ? 1 : 0; }();
  const int ctas = (pairs_enabled && d.M > BM && ctx->num_sms % 2 == 0) ? 2 : 1;
  const int units = ctx->num_sms / ctas;  // persistent CTAs (or CTA pairs)
  g.tiles_m = (int)ceil_div(d.M, BM * ctas);
  g.kblocks = (int)ceil_div(d.K, BK);
  int splits = d.epi == EPI_ATOMIC ? d.splitk : 1;
  if (d.epi == EPI_ATOMIC && splits <= 0) {
    // auto: the split count whose work items fill whole rounds of the persistent grid.  Cost model in
    // k-block units: every round of items costs the k-blocks of one item plus a fixed epilogue term
    // (the fp32 atomics of one output tile), so 150 items on 148 SMs (two rounds) lose to 120 items.
    // (Measured: two rounds of half-length items with overlapped epilogues are NOT faster, the atomics
    // of the extra items cost what the overlap saves.)
    const int tiles = g.tiles_m * g.tiles_n;
    const int epilogue_cost = 6;
    long best_cost = -1;
    splits = 1;
    for (int sp = 1; sp <= g.kblocks && (long)tiles * sp <= 4L * units; ++sp) {
      const int kper = (int)ceil_div(g.kblocks, sp);
      const int eff = (int)ceil_div(g.kblocks, kper);
      const long rounds = ceil_div((long)tiles * eff, units);
      const long cost = rounds * (kper + epilogue_cost);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
    }
  }
  if (splits < 1) splits = 1;
  if (splits > g.kblocks) splits = g.kblocks;
  g.kblocks_per_split = (int)ceil_div(g.kblocks, splits);
  g.splits = (int)ceil_div(g.kblocks, g.kblocks_per_split);
  g.idesc = 0;  // built per tile in the kernel (the last column tile is narrower)
  const int bnh = g.bn / ctas;
  g.b_box_rows = bnh;
  const int b_part_bytes = g.b_mn ? (int)ceil_div(bnh, 64) * CHUNK_BYTES : bnh * BK * 2;
  g.stage_bytes = (int)align_up((size_t)A_BYTES + b_part_bytes, 1024);

  Maps m;
  memset(&m, 0, sizeof(m));
  if (!g.a_mn) V4H_TRY(get_map(ctx, d.A, d.K, d.M, d.lda, BK, BM, 2, &m.a));      // A (M, K): box 64 k x 128 rows
  else         V4H_TRY(get_map(ctx, d.A, d.M, d.K, d.lda, 64, BK, 2, &m.a));      // A (K, M): box 64 m x 64 k rows
  if (!g.b_mn) V4H_TRY(get_map(ctx, d.B, d.K, d.N, d.ldb, BK, bnh, 2, &m.b));     // B (N, K): box 64 k x bn (/2) rows
  else         V4H_TRY(get_map(ctx, d.B, d.N, d.K, d.ldb, 64, BK, 2, &m.b));      // B (K, N): box 64 n x 64 k rows

  const bool obf = d.out_dtype == DT_BF16;
  const int osz = obf ? 2 : 4;
  EpiParams ep = d.ep;
  ep.vec_ok = epilogue_vec_ok(ep, d.epi, d.epi == EPI_ATOMIC ? false : obf);
  // shared-memory budget: [mainloop stages][epilogue pool]; barriers sit in the last BAR_BYTES
  int pool_bytes = 0;
  g.nslots = 0;
  g.tma_store = 0;
  g.nbuf = 2;
  static const int tma_store_enabled = [] { const char* e = getenv("V4H_GEMM_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();
  g.has_out2 = ep.out2 != nullptr;
  g.dbg = d.dbg;
  static const int stage_vecs = [] { const char* e = getenv("V4H_GEMM_STAGE_VECS"); return (e && e[0] == '0') ? 0 : 1; }();
  g.stage_vecs = stage_vecs;
  static const int dbg_skip = [] { const char* e = getenv("V4H_GEMM_DBG_SKIP"); return e ? atoi(e) : 0; }();
  g.dbg_skip = dbg_skip;
  static const int gate_res_tma_store = [] { const char* e = getenv("V4H_GEMM_GATE_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();
  static const int pool_small = [] { const char* e = getenv("V4H_GEMM_POOL_SMALL"); return (e && e[0] == '1') ? 1 : 0; }();
  if (slab) {
    const int out_box = BM * SLAB * osz;
    switch (d.epi) {
      case EPI_BIAS_ACT:
        g.tma_store = tma_store_enabled;
        g.nbuf = ((g.has_out2 && ctas == 1) || pool_small) ? 2 : 3;  // out + out2 at 3 buffers would starve the mainloop ring
        pool_bytes = EG * (g.has_out2 ? 2 : 1) * (g.tma_store ? g.nbuf : 2) * out_box;
        if (g.tma_store) {
          V4H_TRY(get_map(ctx, ep.out, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out));
          if (g.has_out2) V4H_TRY(get_map(ctx, ep.out2, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out2));
        }
        break;
      case EPI_GATE_RES:
        // TMA-store mode: the residual box is updated in place and stored from its ring slot, each group keeps
        // up to nbuf slots in flight; y (out2) is written from registers, so the pool is the ring alone
        g.tma_store = tma_store_enabled && gate_res_tma_store && ep.vec_ok;
        if (g.tma_store) {
          g.nbuf = 2;
          g.nslots = 5;
          pool_bytes = g.nslots * BM * SLAB * 4;
          V4H_TRY(get_map(ctx, ep.res_out, d.N, d.M, ep.ldo, SLAB, BM, 4, &m.out));
        } else {
          g.nslots = ctas == 2 ? 4 : 3;
          pool_bytes = g.nslots * BM * SLAB * 4 + EG * 2 * out_box;
        }
        V4H_TRY(get_map(ctx, ep.res_in, d.N, d.M, ep.ldo, SLAB, BM, 4, &m.in));
        break;
      case EPI_DACT:
        g.tma_store = tma_store_enabled;
        g.nbuf = pool_small ? 2 : 3;
        g.nslots = g.tma_store ? (pool_small ? 5 : MAX_SLOTS) : 6;  // with in-flight stores each group holds up to nbuf slots
        pool_bytes = g.nslots * out_box;
        V4H_TRY(get_map(ctx, ep.aux, d.N, d.M, ep.ld_aux, SLAB, BM, osz, &m.in));
        if (g.tma_store) V4H_TRY(get_map(ctx, ep.out, d.N, d.M, ep.ldo, SLAB, BM, osz, &m.out));
        break;
    }
  }
  g.stages = (SMEM_LIMIT - 1024 - BAR_BYTES - VEC_BYTES - pool_bytes) / g.stage_bytes;
  static const int max_stages = [] { const char* e = getenv("V4H_GEMM_MAX_STAGES"); const int v = e ? atoi(e) : MAX_STAGES; return v >= 2 && v <= MAX_STAGES ? v : MAX_STAGES; }();
  if (g.stages > max_stages) g.stages = max_stages;
  V4H_REQUIRE(g.stages >= 2, "gemm_umma: internal shared-memory budget error");

  const int total = g.tiles_m * g.tiles_n * g.splits;
  const int grid = (total < units ? total : units) * ctas;
  if (launch_sync_enabled())
    fprintf(stderr, "[v4h] gemm %s layout=%d epi=%d act=%d M=%d N=%d K=%d obf=%d slab=%d ctas=%d bn=%d n_pad=%d tiles=%dx%d splits=%d stages=%d nslots=%d nbuf=%d tma_store=%d out2=%d grid=%d\n",
            d.tag, d.layout, d.epi, d.act, d.M, d.N, d.K, (int)obf, (int)slab, ctas, g.bn, g.n_pad, g.tiles_m, g.tiles_n,
            g.splits, g.stages, g.nslots, g.nbuf, g.tma_store, g.has_out2, grid);
  switch (d.epi) {
    case EPI_BIAS_ACT:
      if (d.act == ACT_NONE)
        return obf ? launch2<EPI_BIAS_ACT, ACT_NONE, bf16>(slab, m, g, ep, ctas, grid, s)
                   : launch2<EPI_BIAS_ACT, ACT_NONE, float>(slab, m, g, ep, ctas, grid, s);
      if (d.act == ACT_GELU_TANH && obf) return launch2<EPI_BIAS_ACT, ACT_GELU_TANH_FAST, bf16>(slab, m, g, ep, ctas, grid, s);
      if (d.act == ACT_SILU)
        return obf ? launch2<EPI_BIAS_ACT, ACT_SILU, bf16>(slab, m, g, ep, ctas, grid, s)
                   : launch2<EPI_BIAS_ACT, ACT_SILU, float>(slab, m, g, ep, ctas, grid, s);
      if (d.act == ACT_RELU && obf) return launch2<EPI_BIAS_ACT, ACT_RELU, bf16>(slab, m, g, ep, ctas, grid, s);
      break;
    case EPI_GATE_RES:
      if (obf) return launch2<EPI_GATE_RES, ACT_NONE, bf16>(slab, m, g, ep, ctas, grid, s);
      break;
    case EPI_DACT:
      if (d.act == ACT_GELU_TANH && obf) return launch2<EPI_DACT, ACT_GELU_TANH_FAST, bf16>(slab, m, g, ep, ctas, grid, s);
      if (d.act == ACT_SILU && obf) return launch2<EPI_DACT, ACT_SILU, bf16>(slab, m, g, ep, ctas, grid, s);
      break;
    case EPI_ATOMIC:
      return launch<EPI_ATOMIC, ACT_NONE, float, false>(m, g, ep, ctas, grid, s);
  }
  return fail(V4H_ERR_UNSUPPORTED, "gemm_umma: epilogue %d / activation %d / output dtype %d is not instantiated", d.epi,
              d.act, d.out_dtype);
}

// ------------------------------------------------------------------------------------------
// gemm_gate_res_ln: host side
// ------------------------------------------------------------------------------------------
// tile / cluster / shared-memory plan of one problem; false: not representable (the caller runs the unfused path)
// *cn_out: 1 = one CTA per row tile, 2 = 2-CTA cluster splitting the columns, 3 = cta_group::2 pair over 256 rows
static bool ln_plan(const GemmDesc& d, int num_sms, LnArgs* out, int* cn_out) {
  LnArgs g;
  memset(&g, 0, sizeof(g));
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.kblocks = (int)ceil_div(d.K, BK);
  const int tiles_m = (int)ceil_div(d.M, BM);
  // one CTA per row tile when that fills the GPU; otherwise a 2-CTA cluster splits the columns so that twice as many
  // SMs work on the problem (V4H_GEMM_LN_CLUSTER = 1 / 2 forces the choice)
  static const int forced = [] { const char* e = getenv("V4H_GEMM_LN_CLUSTER"); return e ? atoi(e) : 0; }();
  int cn = tiles_m >= num_sms ? 3 : 2;
  if (forced >= 1 && forced <= 3) cn = forced;
  if (d.N < 64) cn = 1;
  const bool pair = cn == 3;
  auto split_acc = [&](int wc, int* a0, int* a1) {
    if (wc <= 256) { *a0 = wc; *a1 = 0; }
    else { *a0 = (int)ceil_div(wc / 2, 16) * 16; *a1 = wc - *a0; }
  };
  if (cn == 1 || pair) {
    g.col0[0] = 0; g.wc[0] = d.N;
  } else {
    g.wc[0] = (int)ceil_div(d.N / 2, 32) * 32; g.wc[1] = d.N - g.wc[0];
    g.col0[0] = 0; g.col0[1] = g.wc[0];
  }
  int max_wc = 0, tmem_cols = 0;
  for (int r = 0; r < (cn == 2 ? 2 : 1); ++r) {
    split_acc(g.wc[r], &g.a0w[r], &g.a1w[r]);
    if (g.a1w[r] != 0 && g.a1w[r] != g.a0w[r]) return false;  // both halves come through one tensor map (same box)
    max_wc = std::max(max_wc, g.wc[r]);
    tmem_cols = std::max(tmem_cols, g.a1w[r] ? 256 + g.a1w[r] : g.a0w[r]);
  }
  g.vec_w = (int)align_up(max_wc, 4);
  g.tmem_cols = pair ? 512 : tmem_cols;
  g.stage_bytes = (int)align_up((size_t)A_BYTES + (size_t)(pair ? max_wc / 2 : max_wc) * BK * 2, 1024);
  g.has_y = d.ep.out2 != nullptr;
  g.ld_ln = d.ep.ld_ln;
  g.dbg = d.dbg;
  // shared memory: [stages][dedicated residual slots][vectors + partial sums] ... [barriers]
  const int vec_bytes = (int)align_up((size_t)(1 + 3 * LN_MAX_SAMPLES) * g.vec_w * 4 + (EG + 1) * BM * 2 * 4, 1024);
  const int in_box = BM * SLAB * 4, ln_box = BM * SLAB * 2;
  const int nch_max = max_wc / SLAB;
  const int avail = SMEM_LIMIT - 1024 - BAR_BYTES - vec_bytes;
  for (int st = std::min(MAX_STAGES, std::max(2, g.kblocks)); st >= 2; --st) {
    const int left = avail - st * g.stage_bytes;
    if (left < 2 * in_box) continue;
    const int pre = std::min(std::min(left / in_box, nch_max), LN_MAX_SLOTS - 1);
    int post = 0;
    if (pre < nch_max) {
      // the freed operand stages hold the remaining residual slots (pass 1) and then the pass-2 staging boxes: pass 2
      // starts after every pass-1 store has read its slot, so the two may overlap.  A slot is handed back two chunks
      // of its group (four chunks) after its own, so a ring that wraps needs at least 5 slots
      const int room = st * g.stage_bytes / in_box;
      post = std::min(std::min(room, LN_MAX_SLOTS - pre), nch_max - pre);
      if (post < nch_max - pre && post < 5) continue;
      if (post < 1) continue;
    } else if (st * g.stage_bytes < EG * LN_NBUF * ln_box) {
      continue;
    }
    g.stages = st; g.nslots_pre = pre; g.nslots_post = std::max(post, 1);
    *out = g; *cn_out = cn;
    return true;
  }
  return false;
}

bool gemm_gate_res_ln_supported(const GemmDesc& d) {
  static const int enabled = [] { const char* e = getenv("V4H_GEMM_LN_FUSE"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!enabled) return false;
  const EpiParams& p = d.ep;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (d.epi != EPI_GATE_RES || d.layout != GEMM_NT || d.out_dtype != DT_BF16 || !gemm_umma_supported(d)) return false;
  if (!p.ln_out || !p.ln_shift || !p.ln_scale || !p.gate || !p.res_in || !p.res_out) return false;
  if (d.N % 32 != 0 || d.N > 512 || d.N < 32 || p.ldo != d.N || p.ld_ln < d.N || p.ld_ln - d.N > 64 || p.ld_ln % 8) return false;
  if (p.rows_per_sample < 64) return false;  // a 128-row tile must not span more than LN_MAX_SAMPLES samples
  if (!al(p.res_in) || !al(p.res_out) || !al(p.ln_out) || !al(p.out2)) return false;
  LnArgs g; int cn;
  return ln_plan(d, 148, &g, &cn);
}

int gemm_gate_res_ln(UmmaContext* ctx, const GemmDesc& d, cudaStream_t s) {
  V4H_REQUIRE(ctx != nullptr && gemm_gate_res_ln_supported(d), "gemm_gate_res_ln: unsupported problem");
  LnArgs g;
  int cn = 1;
  V4H_REQUIRE(ln_plan(d, ctx->num_sms, &g, &cn), "gemm_gate_res_ln: shared-memory budget (N = %d)", d.N);
  const int tiles_m = (int)ceil_div(d.M, BM);

  Maps m;
  memset(&m, 0, sizeof(m));
  CUtensorMap mB1, mLn;
  memset(&mB1, 0, sizeof(mB1)); memset(&mLn, 0, sizeof(mLn));
  V4H_TRY(get_map(ctx, d.A, d.K, d.M, d.lda, BK, BM, 2, &m.a));
  // W rows of rank 0: box = accumulator width (half of it per CTA of a pair)
  V4H_TRY(get_map(ctx, d.B, d.K, d.N, d.ldb, BK, cn == 3 ? g.a0w[0] / 2 : g.a0w[0], 2, &m.b));
  if (cn == 2) V4H_TRY(get_map(ctx, d.B, d.K, d.N, d.ldb, BK, g.a0w[1], 2, &mB1));
  V4H_TRY(get_map(ctx, d.ep.res_in, d.N, d.M, d.ep.ldo, SLAB, BM, 4, &m.in));
  V4H_TRY(get_map(ctx, d.ep.res_out, d.N, d.M, d.ep.ldo, SLAB, BM, 4, &m.out));
  V4H_TRY(get_map(ctx, d.ep.ln_out, d.N, d.M, d.ep.ld_ln, SLAB, BM, 2, &mLn));
  if (launch_sync_enabled())
    fprintf(stderr, "[v4h] gemm_ln %s M=%d N=%d K=%d cn=%d wc=%d/%d acc=%d+%d stages=%d pre=%d post=%d tmem=%d y=%d\n", d.tag, d.M,
            d.N, d.K, cn, g.wc[0], g.wc[1], g.a0w[0], g.a1w[0], g.stages, g.nslots_pre, g.nslots_post, g.tmem_cols, g.has_y);
  EpiParams ep = d.ep;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(cn == 3 ? (tiles_m + 1) / 2 * 2 : tiles_m * cn));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_LIMIT;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cn >= 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (cn == 3) {
    static bool configured = false;
    if (!configured) {
      V4H_CUDA(cudaFuncSetAttribute(gemm_gate_res_ln_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      configured = true;
    }
    V4H_CUDA(cudaLaunchKernelEx(&cfg, gemm_gate_res_ln_kernel<1, true>, m.a, m.b, m.b, m.in, m.out, mLn, g, ep));
  } else if (cn == 2) {
    static bool configured = false;
    if (!configured) {
      V4H_CUDA(cudaFuncSetAttribute(gemm_gate_res_ln_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      configured = true;
    }
    V4H_CUDA(cudaLaunchKernelEx(&cfg, gemm_gate_res_ln_kernel<2, false>, m.a, m.b, mB1, m.in, m.out, mLn, g, ep));
  } else {
    static bool configured = false;
    if (!configured) {
      V4H_CUDA(cudaFuncSetAttribute(gemm_gate_res_ln_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      configured = true;
    }
    V4H_CUDA(cudaLaunchKernelEx(&cfg, gemm_gate_res_ln_kernel<1, false>, m.a, m.b, m.b, m.in, m.out, mLn, g, ep));
  }
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace v4h
