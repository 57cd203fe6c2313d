// fp32-accumulate SIMT GEMM with the shared epilogues.  This is the arithmetic of the fp32
// precision mode (the reference computes in fp32 with TF32 off, SURVEY.md section 7 "hard parts")
// and of the few GEMMs whose shapes cannot feed TMA/tcgen05 (patch_dim 48/90/6/75/5, per-sample
// conditioning MLPs with B rows).  64x64x16 tiles, 256 threads, 4x4 register micro-tile.
#include "kernels.cuh"

namespace v4h {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct SimtArgs {
  const void* A;
  const void* B;
  int M, N, K;
  long sam, sak, sbk, sbn;  // element strides: A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
  int k_per_split;
};

template <typename TA, typename TB, int EPI, int ACT, typename TOut>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(SimtArgs g, EpiParams p) {
  pdl_wait();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const TA* __restrict__ A = reinterpret_cast<const TA*>(g.A);
  const TB* __restrict__ B = reinterpret_cast<const TB*>(g.B);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const int ty = tid / 16, tx = tid % 16;
  const bool a_kcontig = (g.sak == 1);
  const bool b_kcontig = (g.sbk == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, mm;
      if (a_kcontig) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = to_f(A[(long)gm * g.sam + (long)gk * g.sak]);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, nn;
      if (b_kcontig) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = to_f(B[(long)gk * g.sbk + (long)gn * g.sbn]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    int col0 = n0 + tx * 4;
    if (row < g.M && col0 < g.N) {
      epilogue_run<EPI, ACT, TOut, 4>(p, row, col0, min(4, g.N - col0), acc[i]);
    }
  }
}

template <typename TA, typename TB, int EPI, int ACT, typename TOut>
int launch(const GemmDesc& d, cudaStream_t s) {
  SimtArgs g;
  g.A = d.A; g.B = d.B; g.M = d.M; g.N = d.N; g.K = d.K;
  switch (d.layout) {
    case GEMM_NT: g.sam = d.lda; g.sak = 1; g.sbk = 1; g.sbn = d.ldb; break;
    case GEMM_NN: g.sam = d.lda; g.sak = 1; g.sbk = d.ldb; g.sbn = 1; break;
    default:      g.sam = 1; g.sak = d.lda; g.sbk = d.ldb; g.sbn = 1; break;
  }
  int splits = 1;
  if (EPI == EPI_ATOMIC) {
    splits = d.splitk;
    if (splits <= 0) {  // auto: about two CTAs per SM, at least 256 of K each
      const long tiles = ceil_div(d.M, BM) * ceil_div(d.N, BN);
      splits = (int)ceil_div(296, tiles);
      const int cap = d.K / 256 > 0 ? d.K / 256 : 1;
      if (splits > cap) splits = cap;
    }
    if (splits < 1) splits = 1;
  }
  int kper = (int)ceil_div(ceil_div(d.K, splits), BK) * BK;
  splits = (int)ceil_div(d.K, kper);
  g.k_per_split = kper;
  dim3 grid((unsigned)ceil_div(d.N, BN), (unsigned)ceil_div(d.M, BM), (unsigned)splits);
  EpiParams ep = d.ep;
  ep.vec_ok = epilogue_vec_ok(ep, EPI, sizeof(TOut) == 2);
  V4H_CUDA(launch_pdl(gemm_simt_kernel<TA, TB, EPI, ACT, TOut>, dim3(grid), dim3(NT), 0, s, g, ep));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

template <typename TA, typename TB, typename TOut>
int dispatch_epi(const GemmDesc& d, cudaStream_t s) {
  switch (d.epi) {
    case EPI_BIAS_ACT:
      if (d.act == ACT_NONE) return launch<TA, TB, EPI_BIAS_ACT, ACT_NONE, TOut>(d, s);
      if (d.act == ACT_SILU) return launch<TA, TB, EPI_BIAS_ACT, ACT_SILU, TOut>(d, s);
      if (d.act == ACT_RELU) return launch<TA, TB, EPI_BIAS_ACT, ACT_RELU, TOut>(d, s);
      return launch<TA, TB, EPI_BIAS_ACT, ACT_GELU_TANH, TOut>(d, s);
    case EPI_GATE_RES:
      return launch<TA, TB, EPI_GATE_RES, ACT_NONE, TOut>(d, s);
    case EPI_DACT:
      if (d.act == ACT_SILU) return launch<TA, TB, EPI_DACT, ACT_SILU, TOut>(d, s);
      return launch<TA, TB, EPI_DACT, ACT_GELU_TANH, TOut>(d, s);
    case EPI_ATOMIC:
      return launch<TA, TB, EPI_ATOMIC, ACT_NONE, float>(d, s);
  }
  return fail(V4H_ERR_INVALID, "gemm_simt: unknown epilogue %d", d.epi);
}

}  // namespace

int gemm_simt(const GemmDesc& d, cudaStream_t s) {
  V4H_REQUIRE(d.A && d.B && d.M > 0 && d.N > 0 && d.K > 0, "gemm_simt: bad arguments (M=%d N=%d K=%d)", d.M, d.N, d.K);
  const int key = d.a_dtype * 4 + d.b_dtype * 2 + d.out_dtype;
  switch (key) {
    case 0: return dispatch_epi<float, float, float>(d, s);
    case 1: return dispatch_epi<float, float, bf16>(d, s);
    case 2: return dispatch_epi<float, bf16, float>(d, s);
    case 3: return dispatch_epi<float, bf16, bf16>(d, s);
    case 4: return dispatch_epi<bf16, float, float>(d, s);
    case 5: return dispatch_epi<bf16, float, bf16>(d, s);
    case 6: return dispatch_epi<bf16, bf16, float>(d, s);
    case 7: return dispatch_epi<bf16, bf16, bf16>(d, s);
  }
  return fail(V4H_ERR_INVALID, "gemm_simt: unsupported dtype combination");
}

}  // namespace v4h
