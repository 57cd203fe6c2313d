// Host-side launchers of every kernel in the library (implemented in the .cu files named
// in the comments).  All of them enqueue on `s` and return a V4H_* code.
#pragma once

#include "common.cuh"

namespace v4h {

enum { DT_F32 = 0, DT_BF16 = 1 };
enum { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

// C(M,N) = sum_k A(m,k) B(k,n), row-major storage:
//   NT: A[m*lda+k], B[n*ldb+k]   (forward Linear:  x W^T)
//   NN: A[m*lda+k], B[k*ldb+n]   (dgrad:           dY W)
//   TN: A[k*lda+m], B[k*ldb+n]   (wgrad:           dY^T X)
struct GemmDesc {
  int layout = GEMM_NT;
  const void* A = nullptr;
  int a_dtype = DT_F32;
  int lda = 0;
  const void* B = nullptr;
  int b_dtype = DT_F32;
  int ldb = 0;
  int M = 0, N = 0, K = 0;
  int epi = EPI_BIAS_ACT;
  int act = ACT_NONE;
  int out_dtype = DT_F32;
  const char* tag = "gemm";  // kernel class for v4h_profile_*
  int splitk = 1;  // EPI_ATOMIC only: number of K ranges, 0 = let the engine choose
  long long* dbg = nullptr;  // tcgen05 engine only: device array of 16 cycle counters (v4h_debug_gemm)
  EpiParams ep;
};

// gemm_simt.cu: fp32-accumulate SIMT GEMM (any shape; fp32 or bf16 operands)
int gemm_simt(const GemmDesc& g, cudaStream_t s);

// gemm_umma.cu: tcgen05 / TMEM / TMA GEMM, bf16 operands, fp32 accumulate.
// Requirements: K-extent row pitches multiple of 8 elements (16 B TMA stride alignment).
struct UmmaContext;  // tensor-map cache + driver entry point
UmmaContext* umma_context_create();
void umma_context_destroy(UmmaContext*);
bool gemm_umma_supported(const GemmDesc& g);
int gemm_umma(UmmaContext* ctx, const GemmDesc& g, cudaStream_t s);
// EPI_GATE_RES GEMM whose epilogue also applies the LayerNorm + modulation that follows it (EpiParams::ln_*): one
// CTA (or a 2-CTA cluster splitting the columns) owns whole output rows, so the row statistics never leave the SM
bool gemm_gate_res_ln_supported(const GemmDesc& g);
int gemm_gate_res_ln(UmmaContext* ctx, const GemmDesc& g, cudaStream_t s);
// measurement only: see gemm_umma.cu
int tma_probe(UmmaContext* ctx, const void* buf, int rows, int cols, int stages, int boxes, int box_rows, int producers,
              int iters, int ctas, long long* cycles, cudaStream_t s);

// attention_simt.cu: softmax(q k^T / sqrt(dh)) v per (batch, head); qkv (B,T,3,H,dh)
template <typename T>
int attention_fwd_simt(const T* qkv, T* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s);
template <typename T>
int attention_bwd_simt(const T* qkv, const T* o, const float* lse, const T* d_o, T* dqkv, int B, int Tn,
                       int H, int dh, cudaStream_t s);

// attention_umma.cu: the same contraction on tcgen05 / TMEM (bf16 only, head_dim a multiple of 8, <= 128).
// The backward also writes delta (B, H, T) = rowsum(dO * O) (caller-provided scratch).
bool attention_umma_supported(int dh);
void attention_debug_counters(long long* dev_counters);  // 10 cycle counters of the forward kernel, or null
int attention_fwd_umma(const bf16* qkv, bf16* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s);
int attention_bwd_umma(const bf16* qkv, const bf16* o, const float* lse, const bf16* d_o, float* delta, bf16* dqkv,
                       int B, int Tn, int H, int dh, cudaStream_t s);

// layernorm.cu
// a = LN(h) * (1 + scale[b]) + shift[b];  stats[row] = (mean, rstd).  ld_a >= D is the row pitch of a; when
// ld_a > D the kernel also writes a[row][D] = 1 and zeros up to the pitch: that "ones" column turns the
// weight-gradient GEMM dY^T a into [dW | column sums of dY], i.e. the bias gradient comes for free.
template <typename T>
int ln_modulate_fwd(const float* h, const float* shift, const float* scale, int mod_stride, T* a, int ld_a,
                    float2* stats, int M, int D, int rows_per_sample, cudaStream_t s);
// dh (+)= LNbwd(da * (1 + scale));  dshift[b] += sum_t da;  dscale[b] += sum_t da * xhat
// optionally fused gate backward of the branch that follows in the backward chain:
//   dy = gate[b] * dh_new;  dgate[b] += sum_t dh_new * y;  dbias[n] += sum_rows dy
template <typename T>
int ln_modulate_bwd(const T* da, const float* h, const float2* stats, const float* scale, int mod_stride,
                    float* dh, bool dh_accumulate, float* dshift, float* dscale, int dmod_stride,
                    const T* y, const float* gate, T* dy, float* dgate, float* dbias, int M, int D,
                    int rows_per_sample, cudaStream_t s);
// gate backward alone (first step of the chain inside a block: there is no LN above it):
//   dy = gate[b] * dh;  dgate[b] += sum_t dh * y;  dbias[n] += sum_rows dy
template <typename T>
int gate_bwd(const float* dh, const T* y, const float* gate, int mod_stride, T* dy, float* dgate,
             int dmod_stride, float* dbias, int M, int D, int rows_per_sample, cudaStream_t s);

// elementwise.cu
template <typename T> int colsum_add(const T* x, int ld, float* out, int M, int N, cudaStream_t s);
// out_bf / (dsilu_mul) out_bf: optional bf16 copies of the result (tensor-core operands)
int timestep_embedding(const float* t, int shared_t, float* out, bf16* out_bf, int B, int dim, cudaStream_t s);
int pos_embedding_fwd(const float* freqs, const float* pz, const float* py, const float* px, float* pe,
                      int Tn, int F, cudaStream_t s);
// dfreqs[f] += sum over (b, t, part) of dh[b,t,part*F+f] * d pe / d freq
int pos_embedding_bwd(const float* dh, const float* freqs, const float* pz, const float* py,
                      const float* px, float* dfreqs, int B, int Tn, int F, cudaStream_t s);
int dsilu_mul(const float* x, const float* pre, float* out, bf16* out_bf, int64_t n, cudaStream_t s);
int silu_to_bf16(const float* x, bf16* out, int64_t n, cudaStream_t s);
int cast_f32_to_bf16(const float* x, bf16* out, int64_t n, cudaStream_t s);
struct CastJob { const float* src; bf16* dst; int64_t n; };
int cast_many_f32_to_bf16(const CastJob* jobs_dev, int njobs, int64_t max_n, cudaStream_t s);
int cfm_prepare(const float* x1, const float* x0, const float* t, const int32_t* table, float* xt_tok,
                float* target_tok, int64_t B, int per_sample, cudaStream_t s);
int cfm_loss(const float* v, const float* target, int64_t n, float grad_scale, float* loss_out, float* dv,
             cudaStream_t s);
int axpy4(float* out, const float* y, const float* k0, float a0, const float* k1, float a1,
          const float* k2, float a2, const float* k3, float a3, int64_t n, cudaStream_t s);

// optim.cu: fused gradient clipping + AdamW + bf16 weight refresh
int grad_norm_sq(const float* flat, int64_t n, float* out, cudaStream_t s);
int adamw_step(const v4h_adamw_job* jobs_dev, int njobs, int64_t max_n, const float* norm_sq, float max_norm, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, const int* step_dev,
               const float* lr_dev, float ema_decay, int ema_updates, const int* ema_updates_dev, cudaStream_t s);
int ema_update(const v4h_adamw_job* jobs_dev, int njobs, int64_t max_n, float decay, int num_updates,
               const int* num_updates_dev, cudaStream_t s);
int counter_increment(int* counter, cudaStream_t s);

// postprocess.cu: reverse transform chain of the CaloChallenge ds2 / ds3 shape model, one fused pass
int postprocess_showers(const float* x, const float* cond, int64_t n, int voxels, int n_layers, const int32_t* bounds_dev,
                        int max_layer, float mean, float std, float delta, float cut, float factor, float e_min, float e_max, float alpha,
                        float eps, float norm_cut, float* out, float* e_out, cudaStream_t s);

// preprocess.cu: forward transform chain of the same model (the data feed), one or two fused passes
int preprocess_showers(const float* showers, const float* e_inc, int64_t n, int voxels, int n_layers,
                       const int32_t* bounds_dev, int max_layer, float eps, float factor, float delta, float alpha,
                       float e_min, float e_max, float* mean_std_dev, int compute_stats, double* stats_dev, float* x,
                       float* cond, cudaStream_t s);

// patchify.cu:  dst[b, j] = src[b, table[j]] staged through shared memory per chunk
int patch_permute(const float* src, float* dst, const int32_t* table, const int32_t* chunk_bounds,
                  int num_chunks, int max_chunk, int64_t B, int per_sample, cudaStream_t s);

}  // namespace v4h
