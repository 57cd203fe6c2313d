// tcgen05 / TMEM / TMA GEMM for sm_100a: bf16 operands, fp32 accumulation in tensor memory, the
// shared epilogues of common.cuh applied straight out of TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer   (cp.async.bulk.tensor, 128B-swizzled tiles, 4-stage mbarrier ring)
//   warp 1    MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M = 128, N = bn <= 256, K = 16)
//   warp 2    TMEM allocator (512 columns = two accumulator stages of up to 256 columns)
//   warps 4-7 epilogue       (tcgen05.ld 32x32b, one output row per thread, overlaps the next tile's MMAs)
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
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_BYTES = MAX_BN * BK * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 48 KB
constexpr int CHUNK_BYTES = 64 * BK * 2;       // one [64 k][64 mn] box of an MN-major operand (8 KB)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 512;

struct UmmaArgs {
  int M, N, K;
  int bn;                 // UMMA N of this launch (multiple of 16)
  int tiles_m, tiles_n, splits;
  int kblocks;            // ceil(K / BK)
  int kblocks_per_split;
  int a_mn, b_mn;         // 1 = MN-major operand
  uint32_t idesc;
};

template <int EPI, int ACT, typename TOut>
__global__ void __launch_bounds__(THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const UmmaArgs g, const EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;       // [2]  MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = g.tiles_m * g.tiles_n * g.splits;
  const uint32_t b_tile_bytes = g.b_mn ? (uint32_t)((g.bn + 63) / 64) * CHUNK_BYTES : (uint32_t)g.bn * BK * 2;

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
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int tile = work / g.splits;
      const int m0 = (tile % g.tiles_m) * BM, n0 = (tile / g.tiles_m) * g.bn;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + (uint32_t)acc * MAX_BN + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < g.bn; c0 += 16) {
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

template <int EPI, int ACT, typename TOut>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const UmmaArgs& g, const EpiParams& ep, int grid,
           cudaStream_t s) {
  static bool configured = false;  // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    V4H_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<EPI, ACT, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SMEM_BYTES));
    configured = true;
  }
  gemm_umma_kernel<EPI, ACT, TOut><<<grid, THREADS, SMEM_BYTES, s>>>(ta, tb, g, ep);
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

}  // namespace

struct UmmaContext {
  EncodeTiledFn encode = nullptr;
  int num_sms = 148;
  std::mutex mu;
  // (base pointer, inner extent, outer extent, row pitch in elements, box inner, box outer) -> map
  std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> cache;
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

static int get_map(UmmaContext* ctx, const void* base, int inner, int outer, int pitch, int box_inner, int box_outer,
                   CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto key = std::make_tuple(base, inner, outer, pitch, box_inner, box_outer);
  auto it = ctx->cache.find(key);
  if (it != ctx->cache.end()) { *out = it->second; return V4H_OK; }
  if (!ctx->encode) return fail(V4H_ERR_CUDA, "gemm_umma: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = ctx->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
    // auto: enough K ranges to give every SM a work item, at least 4 k-blocks each
    const int tiles = g.tiles_m * g.tiles_n;
    splits = (int)ceil_div(ctx->num_sms, tiles);
    const int cap = g.kblocks / 4 > 0 ? g.kblocks / 4 : 1;
    if (splits > cap) splits = cap;
  }
  if (splits < 1) splits = 1;
  if (splits > g.kblocks) splits = g.kblocks;
  g.kblocks_per_split = (int)ceil_div(g.kblocks, splits);
  g.splits = (int)ceil_div(g.kblocks, g.kblocks_per_split);
  g.idesc = make_idesc_bf16(BM, g.bn, g.a_mn != 0, g.b_mn != 0);

  CUtensorMap ta, tb;
  if (!g.a_mn) V4H_TRY(get_map(ctx, d.A, d.K, d.M, d.lda, BK, BM, &ta));      // A (M, K): box 64 k x 128 rows
  else         V4H_TRY(get_map(ctx, d.A, d.M, d.K, d.lda, 64, BK, &ta));      // A (K, M): box 64 m x 64 k rows
  if (!g.b_mn) V4H_TRY(get_map(ctx, d.B, d.K, d.N, d.ldb, BK, g.bn, &tb));    // B (N, K): box 64 k x bn rows
  else         V4H_TRY(get_map(ctx, d.B, d.N, d.K, d.ldb, 64, BK, &tb));      // B (K, N): box 64 n x 64 k rows

  const int total = g.tiles_m * g.tiles_n * g.splits;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  const bool obf = d.out_dtype == DT_BF16;
  EpiParams ep = d.ep;
  ep.vec_ok = epilogue_vec_ok(ep, d.epi, d.epi == EPI_ATOMIC ? false : obf);
  switch (d.epi) {
    case EPI_BIAS_ACT:
      if (d.act == ACT_NONE)
        return obf ? launch<EPI_BIAS_ACT, ACT_NONE, bf16>(ta, tb, g, ep, grid, s)
                   : launch<EPI_BIAS_ACT, ACT_NONE, float>(ta, tb, g, ep, grid, s);
      if (d.act == ACT_GELU_TANH && obf) return launch<EPI_BIAS_ACT, ACT_GELU_TANH_FAST, bf16>(ta, tb, g, ep, grid, s);
      if (d.act == ACT_SILU && !obf) return launch<EPI_BIAS_ACT, ACT_SILU, float>(ta, tb, g, ep, grid, s);
      break;
    case EPI_GATE_RES:
      if (obf) return launch<EPI_GATE_RES, ACT_NONE, bf16>(ta, tb, g, ep, grid, s);
      break;
    case EPI_DACT:
      if (d.act == ACT_GELU_TANH && obf) return launch<EPI_DACT, ACT_GELU_TANH_FAST, bf16>(ta, tb, g, ep, grid, s);
      break;
    case EPI_ATOMIC:
      return launch<EPI_ATOMIC, ACT_NONE, float>(ta, tb, g, ep, grid, s);
  }
  return fail(V4H_ERR_UNSUPPORTED, "gemm_umma: epilogue %d / activation %d / output dtype %d is not instantiated", d.epi,
              d.act, d.out_dtype);
}

}  // namespace v4h
