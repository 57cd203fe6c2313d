// Energy-ratio velocity network: the reference's ParallelTransformer
// (nn/cfm/transformer_cfm.py:12-119, embeds=True) around torch.nn.Transformer (post-norm, ReLU, final LayerNorms),
// forward only -- it is sampled for the same conditions right before every shape-sampling job
// (experiments/calochallenge/experiment.py:225-247).
//
// The condition side (c_embed + encoder + the K/V projections of every decoder layer's cross attention) depends
// only on the condition, not on (x, t): energy_encode() runs it ONCE per batch, energy_forward() is one velocity
// evaluation of the decoder + head (80 of them per sampled shower).  Every Linear runs on the library's GEMM
// engines (tcgen05 in bf16 precision, SIMT fp32 in fp32 precision) with bias / ReLU / SiLU epilogues, the
// self-attention over the dims_in tokens on the attention kernels of the ViT path; the small kernels below cover
// what is specific to this network: the embeddings, residual + LayerNorm(affine), cross attention over a
// handful of memory tokens, and the final 512 -> 1 projection.
#include <vector>

#include "kernels.cuh"

namespace v4h {

struct EnergyPlan {
  v4h_energy_dims d;
  int E = 0;  // d_model = encode_t_dim + dim_embedding
  bool bf16 = false, use_umma = false, use_umma_attn = false;
  UmmaContext* umma = nullptr;
  // bf16 weight arena (element offsets)
  struct Enc { size_t in_w, out_w, l1_w, l2_w; } enc[V4H_ENERGY_MAX_LAYERS];
  struct Dec { size_t sa_in_w, sa_out_w, ca_in_w, ca_out_w, l1_w, l2_w; } dec[V4H_ENERGY_MAX_LAYERS];
  size_t head_w = 0, arena_elems = 0;
  CastJob *jobs_dev = nullptr, *jobs_host = nullptr;
  int njobs = 0;
};

namespace {

// ---------------------------------------------------------------- small kernels
// time embedding (GaussianFourierProjection -> Linear, reference transformer_cfm.py:39-42, :165-176) and the token
// embedding of x (:78-82): tgt[b, j] = [ temb_b | x[b, j] * wx + bx + pos_x[j] ]; temb also goes to the head input.
// One CTA per sample; blockDim = E.
template <typename T>
__global__ void energy_embed_kernel(const float* __restrict__ x, const float* __restrict__ t, int shared_t,
                                    const float* __restrict__ gfp_w, const float* __restrict__ tw,
                                    const float* __restrict__ tb, const float* __restrict__ xw,
                                    const float* __restrict__ xb, const float* __restrict__ posx, float* __restrict__ tgt,
                                    T* __restrict__ tgt_t, T* __restrict__ head_in, int dims_in, int Dt, int De) {
  pdl_wait();
  extern __shared__ float sm[];  // feat[Dt], temb[Dt]
  float* feat = sm;
  float* temb = sm + Dt;
  const int b = blockIdx.x, E = Dt + De, tid = threadIdx.x;
  const float tv = t[shared_t ? 0 : b];
  if (tid < Dt / 2) {
    const float proj = tv * gfp_w[tid] * 2.f * 3.14159265358979323846f;
    feat[tid] = sinf(proj);
    feat[Dt / 2 + tid] = cosf(proj);
  }
  __syncthreads();
  if (tid < Dt) {
    float acc = tb[tid];
    for (int k = 0; k < Dt; ++k) acc = fmaf(feat[k], tw[tid * Dt + k], acc);
    temb[tid] = acc;
  }
  __syncthreads();
  for (int j = 0; j < dims_in; ++j) {
    const size_t row = (size_t)b * dims_in + j;
    if (tid < E) {
      float v;
      if (tid < Dt) v = temb[tid];
      else {
        const int e = tid - Dt;
        v = fmaf(x[row], xw[e], xb[e]) + posx[j * De + e];
      }
      tgt[row * E + tid] = v;
      tgt_t[row * E + tid] = from_f<T>(v);
      if (tid < Dt) head_in[row * (size_t)(Dt + E) + tid] = from_f<T>(v);
    }
  }
}

// src[b, s] = c[b, s] * wc + bc + pos_c[s]   (reference transformer_cfm.py:84-87)
template <typename T>
__global__ void energy_cond_embed_kernel(const float* __restrict__ c, const float* __restrict__ cw,
                                         const float* __restrict__ cb, const float* __restrict__ posc,
                                         float* __restrict__ src, T* __restrict__ src_t, int64_t rows, int S, int E) {
  pdl_wait();
  const int64_t n = rows * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / E;
    const int e = (int)(i % E), s = (int)(row % S);
    const float v = fmaf(c[row], cw[e], cb[e]) + posc[s * E + e];
    src[i] = v;
    src_t[i] = from_f<T>(v);
  }
}

// out = LayerNorm(x + y) * gamma + beta (eps 1e-5, biased variance): the post-norm residual step of
// nn.TransformerEncoderLayer / DecoderLayer; y == nullptr: plain LayerNorm (the stacks' final norms).
// One warp per row, E <= 32 * 8.  Writes the fp32 stream (may alias x) and the GEMM-operand copy with its own pitch.
// y_div > 1: y has one row per y_div rows of x (a per-sample vector broadcast over the sample's tokens).
template <typename T>
__global__ void __launch_bounds__(256) add_ln_kernel(const float* x, const float* __restrict__ y, int y_div,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* out, T* __restrict__ out_t, int ld_t, int64_t rows, int E) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = lane + 32 * i;
    v[i] = e < E ? x[row * E + e] + (y ? y[(row / y_div) * E + e] : 0.f) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = lane + 32 * i;
    const float dlt = e < E ? v[i] - mean : 0.f;
    q += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = lane + 32 * i;
    if (e < E) {
      const float o = (v[i] - mean) * rstd * gamma[e] + beta[e];
      if (out) out[row * E + e] = o;
      if (out_t) out_t[row * (size_t)ld_t + e] = from_f<T>(o);
    }
  }
}

// softmax(q k^T / sqrt(dh)) v over S <= 16 key tokens per sample: the decoder's cross attention over the encoded
// condition (S = dims_c: 1 or 3 in the shipped configs) and the encoder's self attention over the same tokens.
// q rows (B * Tq, ldq), k / v rows (B * S, ldkv) with their own column offsets; one thread per (row, head, 4 dims).
template <typename T>
__global__ void small_attention_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, const T* __restrict__ v,
                                       int ldkv, T* __restrict__ o, int ldo, int64_t rows, int Tq, int S, int H, int dh,
                                       float scale) {
  pdl_wait();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= rows * H) return;
  const int64_t row = idx / H;
  const int h = (int)(idx % H);
  const int64_t b = row / Tq;
  const T* qp = q + row * ldq + h * dh;
  float sc[16];
  float mx = -INFINITY;
  for (int s = 0; s < S; ++s) {
    const T* kp = k + (b * S + s) * ldkv + h * dh;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(to_f(qp[d]), to_f(kp[d]), acc);
    sc[s] = acc * scale;
    mx = fmaxf(mx, sc[s]);
  }
  float den = 0.f;
  for (int s = 0; s < S; ++s) { sc[s] = expf(sc[s] - mx); den += sc[s]; }
  const float inv = 1.f / den;
  T* op = o + row * ldo + h * dh;
  for (int d = 0; d < dh; ++d) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(sc[s], to_f(v[(b * S + s) * ldkv + h * dh + d]), acc);
    op[d] = from_f<T>(acc * inv);
  }
}

// out[row] = dot(h[row, :], w) + b   (the head's Linear(dim_feedforward, 1)); one warp per row
template <typename T>
__global__ void __launch_bounds__(256) rowdot_kernel(const T* __restrict__ h, const float* __restrict__ w,
                                                     const float* __restrict__ b, float* __restrict__ out, int64_t rows,
                                                     int K) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(to_f(h[row * K + k]), w[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc + b[0];
}

// ---------------------------------------------------------------- workspace
template <typename T>
struct EnergyWs {
  // condition side (persists between energy_encode and the forwards that follow)
  float* mem_f;  // (B*S, E)
  T* mem_t;
  T* kv[V4H_ENERGY_MAX_LAYERS];  // (B*S, 2E) per decoder layer
  float* yca[V4H_ENERGY_MAX_LAYERS];  // dims_c == 1: the whole cross-attention branch output per sample (B, E)
  // scratch
  float *x_f, *y_f, *lse;
  T *x_t, *qkv, *att, *q, *ff, *head_in, *head_h;
  size_t bytes = 0;
  void layout(const EnergyPlan& p, char* base, int64_t B) {
    const v4h_energy_dims& d = p.d;
    const size_t E = p.E, M = (size_t)B * std::max(d.dims_in, d.dims_c), Ms = (size_t)B * d.dims_c, F = d.dim_feedforward;
    size_t off = 0;
    auto take = [&](size_t n) -> char* { char* q = base ? base + off : nullptr; off += align_up(n, 256); return q; };
    mem_f = (float*)take(Ms * E * 4); mem_t = (T*)take(Ms * E * sizeof(T));
    for (int l = 0; l < d.n_dec; ++l) kv[l] = (T*)take(Ms * 2 * E * sizeof(T));
    for (int l = 0; l < d.n_dec; ++l) yca[l] = d.dims_c == 1 ? (float*)take((size_t)B * E * 4) : nullptr;
    x_f = (float*)take(M * E * 4); y_f = (float*)take(M * E * 4); lse = (float*)take(M * d.nhead * 4);
    x_t = (T*)take(M * E * sizeof(T)); qkv = (T*)take(M * 3 * E * sizeof(T)); att = (T*)take(M * E * sizeof(T));
    q = (T*)take(M * E * sizeof(T)); ff = (T*)take(M * F * sizeof(T));
    head_in = (T*)take(M * (d.encode_t_dim + E) * sizeof(T)); head_h = (T*)take(M * F * sizeof(T));
    bytes = off;
  }
};

template <typename T> constexpr int dt_of() { return sizeof(T) == 2 ? DT_BF16 : DT_F32; }

int energy_gemm(const EnergyPlan& p, const GemmDesc& g, cudaStream_t s) {
  ProfScope ps(g.tag, 2.0 * g.M * g.N * g.K, 0, s);
  if (p.use_umma && g.a_dtype == DT_BF16 && g.b_dtype == DT_BF16 && gemm_umma_supported(g)) return gemm_umma(p.umma, g, s);
  return gemm_simt(g, s);
}

// out (M, N) = act(A (M, K) W^T + bias); W fp32 or its bf16 arena copy
template <typename T>
int linear(const EnergyPlan& p, const char* tag, const T* A, int lda, const float* W, size_t arena_off, const char* arena,
           const float* bias, void* out, int out_dt, int ldo, int64_t M, int N, int K, int act, cudaStream_t s) {
  GemmDesc g;
  g.tag = tag; g.layout = GEMM_NT; g.A = A; g.a_dtype = dt_of<T>(); g.lda = lda;
  if (p.bf16) { g.B = reinterpret_cast<const bf16*>(arena) + arena_off; g.b_dtype = DT_BF16; }
  else { g.B = W; g.b_dtype = DT_F32; }
  g.ldb = K; g.M = (int)M; g.N = N; g.K = K; g.act = act; g.out_dtype = out_dt;
  g.ep.bias = bias; g.ep.out = out; g.ep.ldo = ldo;
  return energy_gemm(p, g, s);
}

template <typename T>
int self_attention(const EnergyPlan& p, const T* qkv, T* o, float* lse, int64_t B, int Tn, cudaStream_t s);
template <>
int self_attention<float>(const EnergyPlan& p, const float* qkv, float* o, float* lse, int64_t B, int Tn, cudaStream_t s) {
  const int H = p.d.nhead, dh = p.E / H;
  return attention_fwd_simt<float>(qkv, o, lse, (int)B, Tn, H, dh, s);
}
template <>
int self_attention<bf16>(const EnergyPlan& p, const bf16* qkv, bf16* o, float* lse, int64_t B, int Tn, cudaStream_t s) {
  const int H = p.d.nhead, dh = p.E / H;
  if (p.use_umma_attn) return attention_fwd_umma(qkv, o, lse, (int)B, Tn, H, dh, s);
  return attention_fwd_simt<bf16>(qkv, o, lse, (int)B, Tn, H, dh, s);
}

template <typename T>
int add_ln(const float* x, const float* y, const float* gamma, const float* beta, float* out, T* out_t, int ld_t,
           int64_t rows, int E, cudaStream_t s, int y_div = 1) {
  ProfScope ps("energy.ln", 0, (double)rows * E * 12, s);
  V4H_CUDA(launch_pdl(add_ln_kernel<T>, dim3((unsigned)ceil_div(rows, 8)), dim3(256), 0, s, x, y, y_div, gamma, beta, out,
                      out_t, ld_t, rows, E));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

template <typename T>
int small_attention(const T* q, int ldq, const T* k, const T* v, int ldkv, T* o, int ldo, int64_t rows, int Tq, int S,
                    int H, int dh, cudaStream_t s) {
  ProfScope ps("energy.xattn", 4.0 * rows * S * H * dh, 0, s);
  V4H_CUDA(launch_pdl(small_attention_kernel<T>, dim3((unsigned)ceil_div(rows * H, 128)), dim3(128), 0, s, q, ldq, k, v, ldkv,
                      o, ldo, rows, Tq, S, H, dh, 1.f / sqrtf((float)dh)));
  V4H_LAUNCH_CHECK();
  return V4H_OK;
}

template <typename T>
int encode_impl(EnergyPlan& p, const v4h_energy_params& w, const char* arena, const float* c, int64_t B, EnergyWs<T>& ws,
                cudaStream_t s) {
  const v4h_energy_dims& d = p.d;
  const int E = p.E, S = d.dims_c, H = d.nhead, dh = E / H, F = d.dim_feedforward, TD = dt_of<T>();
  const int64_t Ms = B * S;
  V4H_CUDA(launch_pdl(energy_cond_embed_kernel<T>, dim3((unsigned)std::min<int64_t>(ceil_div(Ms * E, 256), 1184)), dim3(256),
                      0, s, c, (const float*)w.c_embed_w, (const float*)w.c_embed_b, (const float*)w.pos_c, ws.x_f, ws.x_t,
                      Ms, S, E));
  V4H_LAUNCH_CHECK();
  for (int l = 0; l < d.n_enc; ++l) {
    const v4h_energy_enc_layer& L = w.enc[l];
    V4H_TRY(linear<T>(p, "energy.enc", ws.x_t, E, L.in_w, p.enc[l].in_w, arena, L.in_b, ws.qkv, TD, 3 * E, Ms, 3 * E, E, ACT_NONE, s));
    V4H_TRY(small_attention<T>(ws.qkv, 3 * E, ws.qkv + E, ws.qkv + 2 * E, 3 * E, ws.att, E, Ms, S, S, H, dh, s));
    V4H_TRY(linear<T>(p, "energy.enc", ws.att, E, L.out_w, p.enc[l].out_w, arena, L.out_b, ws.y_f, DT_F32, E, Ms, E, E, ACT_NONE, s));
    V4H_TRY(add_ln<T>(ws.x_f, ws.y_f, L.n1_w, L.n1_b, ws.x_f, ws.x_t, E, Ms, E, s));
    V4H_TRY(linear<T>(p, "energy.enc", ws.x_t, E, L.l1_w, p.enc[l].l1_w, arena, L.l1_b, ws.ff, TD, F, Ms, F, E, ACT_RELU, s));
    V4H_TRY(linear<T>(p, "energy.enc", ws.ff, F, L.l2_w, p.enc[l].l2_w, arena, L.l2_b, ws.y_f, DT_F32, E, Ms, E, F, ACT_NONE, s));
    V4H_TRY(add_ln<T>(ws.x_f, ws.y_f, L.n2_w, L.n2_b, ws.x_f, ws.x_t, E, Ms, E, s));
  }
  V4H_TRY(add_ln<T>(ws.x_f, nullptr, w.enc_norm_w, w.enc_norm_b, ws.mem_f, ws.mem_t, E, Ms, E, s));
  // K / V of every decoder layer's cross attention: rows E..3E of its packed in_proj
  for (int l = 0; l < d.n_dec; ++l) {
    const v4h_energy_dec_layer& L = w.dec[l];
    V4H_TRY(linear<T>(p, "energy.enc", ws.mem_t, E, L.ca_in_w + (size_t)E * E, p.dec[l].ca_in_w + (size_t)E * E, arena,
                      L.ca_in_b + E, ws.kv[l], TD, 2 * E, Ms, 2 * E, E, ACT_NONE, s));
    // one memory token: the softmax over a single key is 1 whatever the query, so the cross-attention branch of
    // this layer is the same vector out_proj(v) for every token of the sample -- computed here, once per batch
    if (S == 1)
      V4H_TRY(linear<T>(p, "energy.enc", ws.kv[l] + E, 2 * E, L.ca_out_w, p.dec[l].ca_out_w, arena, L.ca_out_b, ws.yca[l],
                        DT_F32, E, Ms, E, E, ACT_NONE, s));
  }
  return V4H_OK;
}

template <typename T>
int forward_impl(EnergyPlan& p, const v4h_energy_params& w, const char* arena, const float* x, const float* t, bool shared_t,
                 float* out, int64_t B, EnergyWs<T>& ws, cudaStream_t s) {
  const v4h_energy_dims& d = p.d;
  const int E = p.E, S = d.dims_c, H = d.nhead, dh = E / H, F = d.dim_feedforward, Tn = d.dims_in, Dt = d.encode_t_dim;
  const int TD = dt_of<T>();
  const int64_t M = B * Tn;
  {
    ProfScope ps("energy.embed", 0, 0, s);
    V4H_CUDA(launch_pdl(energy_embed_kernel<T>, dim3((unsigned)B), dim3((unsigned)((E + 31) / 32 * 32)), (size_t)2 * Dt * 4, s,
                        x, t, shared_t ? 1 : 0, (const float*)w.gfp_w, (const float*)w.time_w, (const float*)w.time_b,
                        (const float*)w.x_embed_w, (const float*)w.x_embed_b, (const float*)w.pos_x, ws.x_f, ws.x_t,
                        ws.head_in, Tn, Dt, d.dim_embedding));
    V4H_LAUNCH_CHECK();
  }
  for (int l = 0; l < d.n_dec; ++l) {
    const v4h_energy_dec_layer& L = w.dec[l];
    // self attention over the dims_in tokens
    V4H_TRY(linear<T>(p, "energy.qkv", ws.x_t, E, L.sa_in_w, p.dec[l].sa_in_w, arena, L.sa_in_b, ws.qkv, TD, 3 * E, M, 3 * E, E, ACT_NONE, s));
    {
      ProfScope ps("energy.attn", 4.0 * B * H * Tn * Tn * dh, 0, s);
      V4H_TRY(self_attention<T>(p, ws.qkv, ws.att, ws.lse, B, Tn, s));
    }
    V4H_TRY(linear<T>(p, "energy.proj", ws.att, E, L.sa_out_w, p.dec[l].sa_out_w, arena, L.sa_out_b, ws.y_f, DT_F32, E, M, E, E, ACT_NONE, s));
    V4H_TRY(add_ln<T>(ws.x_f, ws.y_f, L.n1_w, L.n1_b, ws.x_f, ws.x_t, E, M, E, s));
    // cross attention over the encoded condition (K / V precomputed by energy_encode)
    if (S == 1) {
      V4H_TRY(add_ln<T>(ws.x_f, ws.yca[l], L.n2_w, L.n2_b, ws.x_f, ws.x_t, E, M, E, s, Tn));
    } else {
      V4H_TRY(linear<T>(p, "energy.proj", ws.x_t, E, L.ca_in_w, p.dec[l].ca_in_w, arena, L.ca_in_b, ws.q, TD, E, M, E, E, ACT_NONE, s));
      V4H_TRY(small_attention<T>(ws.q, E, ws.kv[l], ws.kv[l] + E, 2 * E, ws.att, E, M, Tn, S, H, dh, s));
      V4H_TRY(linear<T>(p, "energy.proj", ws.att, E, L.ca_out_w, p.dec[l].ca_out_w, arena, L.ca_out_b, ws.y_f, DT_F32, E, M, E, E, ACT_NONE, s));
      V4H_TRY(add_ln<T>(ws.x_f, ws.y_f, L.n2_w, L.n2_b, ws.x_f, ws.x_t, E, M, E, s));
    }
    // feed forward
    V4H_TRY(linear<T>(p, "energy.ff1", ws.x_t, E, L.l1_w, p.dec[l].l1_w, arena, L.l1_b, ws.ff, TD, F, M, F, E, ACT_RELU, s));
    V4H_TRY(linear<T>(p, "energy.ff2", ws.ff, F, L.l2_w, p.dec[l].l2_w, arena, L.l2_b, ws.y_f, DT_F32, E, M, E, F, ACT_NONE, s));
    V4H_TRY(add_ln<T>(ws.x_f, ws.y_f, L.n3_w, L.n3_b, ws.x_f, ws.x_t, E, M, E, s));
  }
  // decoder.norm -> columns [Dt, Dt + E) of the head input (columns [0, Dt) hold the time embedding); head
  V4H_TRY(add_ln<T>(ws.x_f, nullptr, w.dec_norm_w, w.dec_norm_b, nullptr, ws.head_in + Dt, Dt + E, M, E, s));
  V4H_TRY(linear<T>(p, "energy.head", ws.head_in, Dt + E, w.head0_w, p.head_w, arena, w.head0_b, ws.head_h, TD, F, M, F, Dt + E, ACT_SILU, s));
  {
    ProfScope ps("energy.head", 2.0 * M * F, 0, s);
    V4H_CUDA(launch_pdl(rowdot_kernel<T>, dim3((unsigned)ceil_div(M, 8)), dim3(256), 0, s, (const T*)ws.head_h,
                        (const float*)w.head2_w, (const float*)w.head2_b, out, M, F));
    V4H_LAUNCH_CHECK();
  }
  return V4H_OK;
}

}  // namespace

int energy_plan_create(const v4h_energy_dims* dims, EnergyPlan** out) {
  V4H_REQUIRE(dims && out, "energy_plan_create: null argument");
  const v4h_energy_dims& d = *dims;
  V4H_REQUIRE(d.dims_in > 0 && d.dims_c > 0 && d.dims_c <= 16 && d.dim_embedding > 0 && d.encode_t_dim > 0 &&
                  d.encode_t_dim % 2 == 0 && d.nhead > 0 && d.n_enc >= 0 && d.n_enc <= V4H_ENERGY_MAX_LAYERS &&
                  d.n_dec > 0 && d.n_dec <= V4H_ENERGY_MAX_LAYERS && d.dim_feedforward > 0,
              "energy_plan_create: invalid dimensions");
  const int E = d.encode_t_dim + d.dim_embedding;
  V4H_REQUIRE(E % d.nhead == 0 && E <= 256, "energy_plan_create: d_model %d must be a multiple of nhead and <= 256", E);
  V4H_REQUIRE(d.precision == V4H_FP32 || d.precision == V4H_BF16, "energy_plan_create: unknown precision");
  EnergyPlan* p = new EnergyPlan();
  p->d = d; p->E = E;
  p->bf16 = d.precision == V4H_BF16;
  const char* no_umma = getenv("V4H_DISABLE_UMMA");
  p->use_umma = p->bf16 && !(no_umma && no_umma[0] == '1');
  if (p->use_umma) p->umma = umma_context_create();
  p->use_umma_attn = p->use_umma && attention_umma_supported(E / d.nhead);
  if (p->bf16) {
    const size_t F = d.dim_feedforward;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 128); return o; };
    for (int l = 0; l < d.n_enc; ++l) p->enc[l] = {take(3ull * E * E), take((size_t)E * E), take(F * E), take(E * F)};
    for (int l = 0; l < d.n_dec; ++l)
      p->dec[l] = {take(3ull * E * E), take((size_t)E * E), take(3ull * E * E), take((size_t)E * E), take(F * E), take(E * F)};
    p->head_w = take(F * (size_t)(d.encode_t_dim + E));
    p->arena_elems = off;
    p->njobs = 4 * d.n_enc + 6 * d.n_dec + 1;
    if (cudaMalloc(&p->jobs_dev, sizeof(CastJob) * p->njobs) != cudaSuccess ||
        cudaMallocHost(&p->jobs_host, sizeof(CastJob) * p->njobs) != cudaSuccess) {
      delete p;
      return fail(V4H_ERR_CUDA, "energy_plan_create: cannot allocate the cast job table");
    }
    memset(p->jobs_host, 0, sizeof(CastJob) * p->njobs);
  }
  *out = p;
  return V4H_OK;
}

void energy_plan_destroy(EnergyPlan* p) {
  if (!p) return;
  if (p->jobs_dev) cudaFree(p->jobs_dev);
  if (p->jobs_host) cudaFreeHost(p->jobs_host);
  if (p->umma) umma_context_destroy(p->umma);
  delete p;
}

size_t energy_workspace_bytes(const EnergyPlan* p, int64_t B) {
  if (p->bf16) { EnergyWs<bf16> ws; ws.layout(*p, nullptr, B); return ws.bytes; }
  EnergyWs<float> ws; ws.layout(*p, nullptr, B); return ws.bytes;
}
size_t energy_arena_bytes(const EnergyPlan* p) { return p->arena_elems * 2; }

int energy_prepare_weights(EnergyPlan* p, const v4h_energy_params* w, void* arena, cudaStream_t s) {
  if (!p->bf16) return V4H_OK;
  V4H_REQUIRE(w && arena, "energy_prepare_weights: null argument");
  const v4h_energy_dims& d = p->d;
  const size_t E = p->E, F = d.dim_feedforward;
  bf16* wa = reinterpret_cast<bf16*>(arena);
  std::vector<CastJob> jobs;
  int64_t max_n = 0;
  auto add = [&](const float* src, size_t off, size_t n) {
    jobs.push_back(CastJob{src, wa + off, (int64_t)n});
    if ((int64_t)n > max_n) max_n = (int64_t)n;
  };
  for (int l = 0; l < d.n_enc; ++l) {
    const v4h_energy_enc_layer& L = w->enc[l];
    V4H_REQUIRE(L.in_w && L.out_w && L.l1_w && L.l2_w, "energy_prepare_weights: null encoder weight");
    add(L.in_w, p->enc[l].in_w, 3 * E * E); add(L.out_w, p->enc[l].out_w, E * E);
    add(L.l1_w, p->enc[l].l1_w, F * E); add(L.l2_w, p->enc[l].l2_w, E * F);
  }
  for (int l = 0; l < d.n_dec; ++l) {
    const v4h_energy_dec_layer& L = w->dec[l];
    V4H_REQUIRE(L.sa_in_w && L.sa_out_w && L.ca_in_w && L.ca_out_w && L.l1_w && L.l2_w, "energy_prepare_weights: null decoder weight");
    add(L.sa_in_w, p->dec[l].sa_in_w, 3 * E * E); add(L.sa_out_w, p->dec[l].sa_out_w, E * E);
    add(L.ca_in_w, p->dec[l].ca_in_w, 3 * E * E); add(L.ca_out_w, p->dec[l].ca_out_w, E * E);
    add(L.l1_w, p->dec[l].l1_w, F * E); add(L.l2_w, p->dec[l].l2_w, E * F);
  }
  V4H_REQUIRE(w->head0_w, "energy_prepare_weights: null head weight");
  add(w->head0_w, p->head_w, F * (d.encode_t_dim + E));
  if (memcmp(jobs.data(), p->jobs_host, sizeof(CastJob) * jobs.size()) != 0) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
      return fail(V4H_ERR_INVALID, "energy_prepare_weights: parameter storage moved during CUDA-graph capture");
    V4H_CUDA(cudaStreamSynchronize(s));
    memcpy(p->jobs_host, jobs.data(), sizeof(CastJob) * jobs.size());
    V4H_CUDA(cudaMemcpyAsync(p->jobs_dev, p->jobs_host, sizeof(CastJob) * jobs.size(), cudaMemcpyHostToDevice, s));
  }
  return cast_many_f32_to_bf16(p->jobs_dev, (int)jobs.size(), max_n, s);
}

int energy_encode(EnergyPlan* p, const v4h_energy_params* w, const void* arena, const float* c, int64_t B, void* workspace,
                  size_t workspace_bytes, cudaStream_t s) {
  V4H_REQUIRE(p && w && c && workspace && B > 0, "energy_encode: bad arguments");
  V4H_REQUIRE(!p->bf16 || arena, "energy_encode: bf16 precision needs the weight arena");
  if (p->bf16) {
    EnergyWs<bf16> ws; ws.layout(*p, reinterpret_cast<char*>(workspace), B);
    V4H_REQUIRE(ws.bytes <= workspace_bytes, "energy_encode: workspace too small");
    return encode_impl<bf16>(*p, *w, (const char*)arena, c, B, ws, s);
  }
  EnergyWs<float> ws; ws.layout(*p, reinterpret_cast<char*>(workspace), B);
  V4H_REQUIRE(ws.bytes <= workspace_bytes, "energy_encode: workspace too small");
  return encode_impl<float>(*p, *w, nullptr, c, B, ws, s);
}

int energy_forward(EnergyPlan* p, const v4h_energy_params* w, const void* arena, const float* x, const float* t, bool shared_t,
                   float* out, int64_t B, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  V4H_REQUIRE(p && w && x && t && out && workspace && B > 0, "energy_forward: bad arguments");
  V4H_REQUIRE(!p->bf16 || arena, "energy_forward: bf16 precision needs the weight arena");
  if (p->bf16) {
    EnergyWs<bf16> ws; ws.layout(*p, reinterpret_cast<char*>(workspace), B);
    V4H_REQUIRE(ws.bytes <= workspace_bytes, "energy_forward: workspace too small");
    return forward_impl<bf16>(*p, *w, (const char*)arena, x, t, shared_t, out, B, ws, s);
  }
  EnergyWs<float> ws; ws.layout(*p, reinterpret_cast<char*>(workspace), B);
  V4H_REQUIRE(ws.bytes <= workspace_bytes, "energy_forward: workspace too small");
  return forward_impl<float>(*p, *w, nullptr, x, t, shared_t, out, B, ws, s);
}

}  // namespace v4h
