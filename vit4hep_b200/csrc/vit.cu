// Orchestration of the ViT velocity network forward / backward (reference nn/vit.py:185-206 forward,
// :327-333 DiTBlock, :347-351 FinalLayer; backward derived in SURVEY.md appendix A) on top of the
// kernels in this directory.  Nothing here allocates activations: every buffer is a slice of the
// caller's workspace, laid out by Workspace::layout().
#include <cstdlib>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace v4h {

// Streams of one sub-batch ("lane").  The batch is split into up to two sub-batches whose kernel chains are
// issued on two streams: every kernel of the chain keeps all SMs busy only in the middle of its run (launch
// gap, prologue, ragged last round, exposed last epilogue: about 4 + 2..7 us of a 15..30 us GEMM at batch 64),
// and a second, independent chain fills exactly those holes -- its CTAs are already queued when an SM frees.
// Results do not depend on the split (a GEMM output element sees the same k order whatever M is).
struct Lane {
  cudaStream_t main = nullptr;  // lane 0 runs on the caller's stream (nullptr here)
  // weight-gradient GEMMs and the small embedding / conditioning chains run on side streams, off the
  // dgrad -> LayerNorm -> attention chain (V4H_WGRAD_STREAM=0: same stream)
  cudaStream_t side = nullptr, side2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_done = nullptr;
};

struct Plan {
  v4h_vit_dims d;
  int Nmod = 0;           // depth * 6D + 2D columns of the concatenated adaLN output
  bool bf16 = false;
  bool use_umma = false;  // tcgen05 GEMMs (bf16 precision only)
  bool use_umma_attn = false;
  int ldx = 0;  // row pitch of the LN-modulated activations a / m: D, or D + 8 with the "ones" column
  UmmaContext* umma = nullptr;
  // bf16 weight arena layout (element offsets in bf16 units, then the fp32 bias tail in bytes)
  struct BlockArena { size_t qkv, proj, fc1, fc2; } arena_blocks[V4H_MAX_DEPTH];
  size_t arena_ada = 0;           // bf16 (Nmod, D)
  // bf16 copies of the small Linears so that they run on the tensor cores too
  size_t arena_final = 0, arena_x = 0, arena_t0 = 0, arena_t2 = 0, arena_c2 = 0;
  size_t arena_bf16_elems = 0;
  size_t arena_ada_bias_bytes = 0;  // byte offset of fp32 (Nmod) concatenated adaLN biases
  size_t arena_bytes = 0;
  // device + pinned staging for the multi-tensor cast job table
  CastJob* jobs_dev = nullptr;
  CastJob* jobs_host = nullptr;
  int njobs = 0;
  Lane lanes[2];
  cudaEvent_t ev_split = nullptr;
  int microbatch = 1;  // V4H_MICROBATCH=2: run every batch of >= 2 samples as two sub-batches (default: one)
  // side streams in use (none while the per-kernel profiler wants launches that do not overlap)
  bool forking() const { return lanes[0].side != nullptr && !profiling_enabled(); }
  // number of sub-batches a batch of B samples runs as (fixed per (plan, B): the workspace layout follows it)
  int sub_batches(int64_t B) const {
    if (!use_umma || lanes[0].side == nullptr || B < 2 || microbatch == 1) return 1;
    // Measured (ds2, batch 64, one B200): 2.655 ms per step split against 2.619 ms unsplit -- every half-size kernel
    // pays the full launch / prologue / exposed-epilogue cost again and the overlap only wins that back -- so the
    // split is opt-in (V4H_MICROBATCH=2)
    return microbatch == 2 ? 2 : 1;
  }
};

// q waits for everything issued on s so far
static int fork_to(const Lane& L, cudaStream_t s, cudaStream_t q) {
  V4H_CUDA(cudaEventRecord(L.ev_fork, s));
  V4H_CUDA(cudaStreamWaitEvent(q, L.ev_fork, 0));
  return V4H_OK;
}
// s waits for everything issued on the side stream q so far
static int join_from(const Lane& L, cudaStream_t q, cudaStream_t s) {
  cudaEvent_t ev = q == L.side ? L.ev_join : L.ev_join2;
  V4H_CUDA(cudaEventRecord(ev, q));
  V4H_CUDA(cudaStreamWaitEvent(s, ev, 0));
  return V4H_OK;
}

namespace {

// ------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------
struct BlockBufs {
  void *a, *qkv, *o, *y1, *m, *u, *g, *y2;  // TA
  float* lse;
  float2 *stats1, *stats2;
};

// attention forward / backward of one block: tcgen05 kernels in bf16 mode, SIMT otherwise
template <typename T>
int attn_fwd(const Plan& p, const void* qkv, void* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s);
template <typename T>
int attn_bwd(const Plan& p, const void* qkv, const void* o, const float* lse, const void* d_o, float* delta,
             void* dqkv, int B, int Tn, int H, int dh, cudaStream_t s);

struct Workspace {
  // conditioning
  float *pe, *temb_in, *t_h_pre, *t_h, *te, *c_h_pre, *c_h, *cond, *sc, *mod;
  bf16* sc_bf16;
  // bf16 mode: tensor-core operands of the embedding / conditioning / output Linears
  bf16 *x_bf, *temb_bf, *t_h_bf, *t_hpre_bf, *c_h_bf, *c_hpre_bf;
  bf16 *dout_bf, *dh_bf, *dcond_bf, *dvec_bf, *dvec2_bf;  // backward only
  // finetuning mappers (fp32): SiLU(x Wm^T + bm) / SiLU(c Wm^T + bm), their pre-activations, and the gradients
  // flowing back into them
  float *xm, *xm_pre, *cm, *cm_pre, *dxm, *dcm;
  // residual stream (fp32): training keeps 2*depth+1 copies, inference 1
  std::vector<float*> h;
  std::vector<BlockBufs> blk;  // training: depth entries; inference: 1 shared entry
  void* a_f;
  float2* stats_f;
  // backward scratch
  float *dh, *dmod, *dsc, *dcond, *dvec, *attn_delta;
  bf16* dmod_bf16;
  void *dy_mlp[2], *dy_attn, *du, *dm, *dqkv;  // dy: gate * dh of the MLP branch (by block parity) / attention branch
  size_t bytes = 0;

  void layout(const Plan& p, char* base, int64_t B, bool train) {
    const v4h_vit_dims& d = p.d;
    const size_t M = (size_t)B * d.tokens, D = d.hidden_dim, Hm = d.mlp_hidden;
    const size_t ta = p.bf16 ? 2 : 4;
    size_t off = 0;
    auto take = [&](size_t nbytes) -> char* {
      char* ptr = base ? base + off : nullptr;
      off += align_up(nbytes, 256);
      return ptr;
    };
    pe = (float*)take((size_t)d.tokens * D * 4);
    temb_in = (float*)take((size_t)B * d.freq_dim * 4);
    t_h_pre = (float*)take(B * D * 4); t_h = (float*)take(B * D * 4); te = (float*)take(B * D * 4);
    c_h_pre = (float*)take(B * D * 4); c_h = (float*)take(B * D * 4);
    cond = (float*)take(B * D * 4); sc = (float*)take(B * D * 4);
    sc_bf16 = (bf16*)take(B * D * 2);
    mod = (float*)take((size_t)B * p.Nmod * 4);
    if (p.bf16) {
      x_bf = (bf16*)take(M * d.patch_dim * 2);
      temb_bf = (bf16*)take((size_t)B * d.freq_dim * 2);
      t_h_bf = (bf16*)take(B * D * 2); t_hpre_bf = (bf16*)take(B * D * 2);
      c_h_bf = (bf16*)take(B * D * 2); c_hpre_bf = (bf16*)take(B * D * 2);
    } else {
      x_bf = temb_bf = t_h_bf = t_hpre_bf = c_h_bf = c_hpre_bf = nullptr;
    }
    xm = xm_pre = cm = cm_pre = dxm = dcm = nullptr;
    if (d.x_map_dim > 0) {
      xm = (float*)take(M * d.patch_dim * 4);
      if (train) { xm_pre = (float*)take(M * d.patch_dim * 4); dxm = (float*)take(M * d.patch_dim * 4); }
    }
    if (d.c_map_dim > 0) {
      cm = (float*)take((size_t)B * d.cond_dim * 4);
      if (train) {
        cm_pre = (float*)take((size_t)B * d.cond_dim * 4); dcm = (float*)take((size_t)B * d.cond_dim * 4);
      }
    }
    const int nh = train ? 2 * d.depth + 1 : 1;
    h.resize(nh);
    for (int i = 0; i < nh; ++i) h[i] = (float*)take(M * D * 4);
    const int nb = train ? d.depth : 1;
    blk.resize(nb);
    for (int i = 0; i < nb; ++i) {
      BlockBufs& b = blk[i];
      b.a = take(M * p.ldx * ta); b.qkv = take(M * 3 * D * ta); b.o = take(M * D * ta);
      b.y1 = train ? take(M * D * ta) : nullptr;
      b.m = take(M * p.ldx * ta);
      b.u = train ? take(M * Hm * ta) : nullptr;
      b.g = take(M * Hm * ta);
      b.y2 = train ? take(M * D * ta) : nullptr;
      b.lse = (float*)take((size_t)B * d.num_heads * d.tokens * 4);
      b.stats1 = (float2*)take(M * 8); b.stats2 = (float2*)take(M * 8);
    }
    a_f = take(M * D * ta);
    stats_f = (float2*)take(M * 8);
    if (train) {
      dh = (float*)take(M * D * 4);
      dmod = (float*)take((size_t)B * p.Nmod * 4);
      dmod_bf16 = (bf16*)take((size_t)B * p.Nmod * 2);
      dsc = (float*)take(B * D * 4); dcond = (float*)take(B * D * 4); dvec = (float*)take(B * D * 4);
      dy_mlp[0] = take(M * D * ta); dy_mlp[1] = take(M * D * ta); dy_attn = take(M * D * ta); du = take(M * Hm * ta); dm = take(M * D * ta); dqkv = take(M * 3 * D * ta);
      attn_delta = (float*)take((size_t)B * d.num_heads * d.tokens * 4);
      if (p.bf16) {
        dout_bf = (bf16*)take(M * d.out_dim * 2); dh_bf = (bf16*)take(M * D * 2);
        dcond_bf = (bf16*)take(B * D * 2); dvec_bf = (bf16*)take(B * D * 2); dvec2_bf = (bf16*)take(B * D * 2);
      } else {
        dout_bf = dh_bf = dcond_bf = dvec_bf = dvec2_bf = nullptr;
      }
    } else {
      dout_bf = dh_bf = dcond_bf = dvec_bf = dvec2_bf = nullptr;
      dh = dmod = dsc = dcond = dvec = attn_delta = nullptr; dmod_bf16 = nullptr;
      dy_mlp[0] = dy_mlp[1] = dy_attn = du = dm = dqkv = nullptr;
    }
    bytes = off;
  }
};

inline int hidx(bool train, int i) { return train ? i : 0; }

template <>
int attn_fwd<float>(const Plan&, const void* qkv, void* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s) {
  return attention_fwd_simt<float>((const float*)qkv, (float*)o, lse, B, Tn, H, dh, s);
}
template <>
int attn_fwd<bf16>(const Plan& p, const void* qkv, void* o, float* lse, int B, int Tn, int H, int dh, cudaStream_t s) {
  if (p.use_umma_attn) return attention_fwd_umma((const bf16*)qkv, (bf16*)o, lse, B, Tn, H, dh, s);
  return attention_fwd_simt<bf16>((const bf16*)qkv, (bf16*)o, lse, B, Tn, H, dh, s);
}
template <>
int attn_bwd<float>(const Plan&, const void* qkv, const void* o, const float* lse, const void* d_o, float*,
                    void* dqkv, int B, int Tn, int H, int dh, cudaStream_t s) {
  return attention_bwd_simt<float>((const float*)qkv, (const float*)o, lse, (const float*)d_o, (float*)dqkv, B, Tn, H,
                                   dh, s);
}
template <>
int attn_bwd<bf16>(const Plan& p, const void* qkv, const void* o, const float* lse, const void* d_o, float* delta,
                   void* dqkv, int B, int Tn, int H, int dh, cudaStream_t s) {
  if (p.use_umma_attn)
    return attention_bwd_umma((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)d_o, delta, (bf16*)dqkv, B, Tn, H,
                              dh, s);
  return attention_bwd_simt<bf16>((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)d_o, (bf16*)dqkv, B, Tn, H, dh,
                                  s);
}

// run `f` (a launcher) inside a profiling scope
template <typename F>
int prof(const char* tag, double flops, double bytes, cudaStream_t s, F&& f) {
  ProfScope ps(tag, flops, bytes, s);
  return f();
}

// ------------------------------------------------------------------------------------------
// GEMM dispatch
// ------------------------------------------------------------------------------------------
int run_gemm(const Plan& p, const GemmDesc& g, cudaStream_t s) {
  const double esz_a = g.a_dtype == DT_BF16 ? 2 : 4, esz_b = g.b_dtype == DT_BF16 ? 2 : 4;
  const double esz_o = g.epi == EPI_ATOMIC ? 4 : (g.out_dtype == DT_BF16 ? 2 : 4);
  ProfScope ps(g.tag, 2.0 * g.M * g.N * g.K,
               esz_a * g.M * g.K + esz_b * g.N * g.K + esz_o * (double)g.M * g.N, s);
  if (p.use_umma && g.a_dtype == DT_BF16 && g.b_dtype == DT_BF16 && gemm_umma_supported(g)) {
    if (g.ep.ln_out != nullptr) return gemm_gate_res_ln(p.umma, g, s);
    return gemm_umma(p.umma, g, s);
  }
  if (g.ep.ln_out != nullptr) return fail(V4H_ERR_INVALID, "internal: fused LayerNorm epilogue on the SIMT engine");
  return gemm_simt(g, s);
}

// Gated-residual GEMM followed by the LayerNorm + modulation of the NEXT branch (reference nn/vit.py:331-332: the
// norm + modulate that consumes the new residual stream).  bf16 tensor-core path: ONE kernel, the LayerNorm lives in
// the GEMM epilogue (gemm_gate_res_ln); otherwise the GEMM and the stand-alone LayerNorm kernel.
template <typename T>
int gate_res_then_ln(const Plan& p, GemmDesc g, const float* shift, const float* scale, void* ln_out, int ld_ln,
                     float2* stats, int M, int D, int Tn, cudaStream_t s) {
  GemmDesc f = g;
  f.ep.ln_shift = shift; f.ep.ln_scale = scale; f.ep.ln_out = ln_out; f.ep.ld_ln = ld_ln; f.ep.ln_stats = stats;
  f.ep.ln_eps = 1e-6f;
  if (p.use_umma && g.a_dtype == DT_BF16 && g.b_dtype == DT_BF16 && gemm_gate_res_ln_supported(f)) return run_gemm(p, f, s);
  V4H_TRY(run_gemm(p, g, s));
  return prof("ln.fwd", 0, (double)M * D * (4 + sizeof(T)), s, [&] {
    return ln_modulate_fwd<T>(g.ep.res_out, shift, scale, g.ep.mod_stride, (T*)ln_out, ld_ln, stats, M, D, Tn, s); });
}

GemmDesc linear_fwd(const void* A, int a_dt, int lda, const void* W, int w_dt, int ldw, int M, int N, int K) {
  GemmDesc g;
  g.layout = GEMM_NT; g.A = A; g.a_dtype = a_dt; g.lda = lda; g.B = W; g.b_dtype = w_dt; g.ldb = ldw;
  g.M = M; g.N = N; g.K = K;
  return g;
}

// dW (M_out x N_out) += A^T B with A (K, M_out), B (K, N_out)
// bias_out (optional): B carries a "ones" column at index N_out (ldb > N_out): the extra output column, the
// column sums of A = the bias gradient, is accumulated into bias_out (M_out)
int wgrad(const Plan& p, const void* A, int a_dt, int lda, const void* B, int b_dt, int ldb, float* dW,
          int M_out, int N_out, int K, cudaStream_t s, const char* tag = "wgrad", float* bias_out = nullptr) {
  GemmDesc g;
  g.tag = tag;
  g.layout = GEMM_TN; g.A = A; g.a_dtype = a_dt; g.lda = lda; g.B = B; g.b_dtype = b_dt; g.ldb = ldb;
  g.M = M_out; g.N = bias_out ? N_out + 1 : N_out; g.K = K;
  g.epi = EPI_ATOMIC; g.out_dtype = DT_F32;
  g.ep.extra_out = bias_out; g.ep.extra_col = bias_out ? N_out : -1;
  g.ep.out = dW; g.ep.ldo = N_out;
  g.splitk = 0;  // auto
  return run_gemm(p, g, s);
}

// one sub-batch of a forward / backward call: its slice of the inputs, its own workspace and streams
struct Sub {
  const float *x = nullptr, *t = nullptr, *c = nullptr, *dout = nullptr;
  float* out = nullptr;
  int B = 0;
  Workspace ws;
  cudaStream_t s = nullptr;
  const Lane* lane = nullptr;
  bool side_busy = false;   // backward: a weight gradient is running on lane->side
  bool side2_busy = false;  // backward: an adaLN weight gradient is running on lane->side2
};

// ---- positional embedding + patch embedding, conditioning, every adaLN modulation   (nn/vit.py:192-199, :328-330)
template <typename T>
int forward_prologue(Plan& p, const v4h_vit_params& w, const char* arena, Sub& u, bool shared_t, bool train) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  cudaStream_t s = u.s;
  const Lane& L = *u.lane;
  const int B = u.B, Tn = d.tokens, D = d.hidden_dim, M = B * Tn;
  const int Bt = shared_t ? 1 : B;
  const bf16* wa = reinterpret_cast<const bf16*>(arena);
  const float* pe = w.pos_embed;
  const bool fast = p.bf16 && p.use_umma;  // small Linears on the tensor cores with bf16 operand copies
  // The embedding of x and the two conditioning MLPs are independent chains of small kernels (a few CTAs each,
  // latency-bound): the x chain runs on one side stream, the first c_embedder Linear on a second one, the
  // t_embedder on the caller's stream; they meet at c_embedder.2 (+ te) and before the first block
  const bool forked = fast && p.forking();
  cudaStream_t sx = forked ? L.side : s;   // x chain
  cudaStream_t sc = forked ? L.side2 : s;  // c_embedder.0
  if (forked) { V4H_TRY(fork_to(L, s, sx)); V4H_TRY(fork_to(L, s, sc)); }
  if (d.learn_pos_embed) {
    V4H_TRY(pos_embedding_fwd(w.pos_embed_freqs, w.pos_z, w.pos_y, w.pos_x, ws.pe, Tn, D / 6, sx));
    pe = ws.pe;
  }
  const float* xin = u.x;
  if (d.x_map_dim > 0) {  // finetuning: x_embedder = Sequential(mapper, SiLU, old Linear)
    GemmDesc g = linear_fwd(u.x, DT_F32, d.x_map_dim, w.xm_w, DT_F32, d.x_map_dim, M, d.patch_dim, d.x_map_dim);
    g.tag = "gemm.x_map"; g.act = ACT_SILU; g.ep.bias = w.xm_b; g.ep.out = ws.xm; g.ep.out2 = train ? ws.xm_pre : nullptr;
    g.ep.ldo = d.patch_dim;
    V4H_TRY(run_gemm(p, g, sx));
    xin = ws.xm;
  }
  {
    GemmDesc g = linear_fwd(xin, DT_F32, d.patch_dim, w.x_w, DT_F32, d.patch_dim, M, D, d.patch_dim);
    if (fast) {
      V4H_TRY(cast_f32_to_bf16(xin, ws.x_bf, (int64_t)M * d.patch_dim, sx));
      g = linear_fwd(ws.x_bf, DT_BF16, d.patch_dim, wa + p.arena_x, DT_BF16, d.patch_dim, M, D, d.patch_dim);
    }
    g.ep.bias = w.x_b; g.ep.out = ws.h[0]; g.ep.ldo = D;
    g.ep.addend = pe; g.ep.addend_rows = Tn; g.ep.ld_addend = D;
    g.tag = "gemm.x_embed";
    V4H_TRY(run_gemm(p, g, sx));
  }
  // ---- conditioning: cond = t_embedder(t) + c_embedder(c); sc = SiLU(cond)   (nn/vit.py:197-199)
  const float* cin = u.c;
  if (d.c_map_dim > 0) {  // finetuning: c_embedder = Sequential(mapper, SiLU, old Sequential)
    GemmDesc g = linear_fwd(u.c, DT_F32, d.c_map_dim, w.cm_w, DT_F32, d.c_map_dim, B, d.cond_dim, d.c_map_dim);
    g.tag = "gemm.c_map"; g.act = ACT_SILU; g.ep.bias = w.cm_b; g.ep.out = ws.cm; g.ep.out2 = train ? ws.cm_pre : nullptr;
    g.ep.ldo = d.cond_dim;
    V4H_TRY(run_gemm(p, g, sc));
    cin = ws.cm;
  }
  V4H_TRY(timestep_embedding(u.t, shared_t ? 1 : 0, ws.temb_in, fast ? ws.temb_bf : nullptr, Bt, d.freq_dim, s));
  if (fast) {
    // t_embedder.mlp and c_embedder.2 on the tensor cores; hidden activations kept in bf16
    GemmDesc g = linear_fwd(ws.temb_bf, DT_BF16, d.freq_dim, wa + p.arena_t0, DT_BF16, d.freq_dim, Bt, D, d.freq_dim);
    g.tag = "gemm.cond"; g.act = ACT_SILU; g.out_dtype = DT_BF16;
    g.ep.bias = w.t0_b; g.ep.out = ws.t_h_bf; g.ep.out2 = ws.t_hpre_bf; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, s));
    g = linear_fwd(ws.t_h_bf, DT_BF16, D, wa + p.arena_t2, DT_BF16, D, Bt, D, D); g.tag = "gemm.cond";
    g.ep.bias = w.t2_b; g.ep.out = ws.te; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, s));
    g = linear_fwd(cin, DT_F32, d.cond_dim, w.c0_w, DT_F32, d.cond_dim, B, D, d.cond_dim); g.tag = "gemm.cond";
    g.act = ACT_SILU; g.out_dtype = DT_BF16;
    g.ep.bias = w.c0_b; g.ep.out = ws.c_h_bf; g.ep.out2 = ws.c_hpre_bf; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, sc));
    if (forked) V4H_TRY(join_from(L, sc, s));
    g = linear_fwd(ws.c_h_bf, DT_BF16, D, wa + p.arena_c2, DT_BF16, D, B, D, D); g.tag = "gemm.cond";
    g.act = ACT_SILU; g.ep.bias = w.c2_b; g.ep.out = ws.sc; g.ep.out2 = ws.cond; g.ep.ldo = D;
    g.ep.addend = ws.te; g.ep.addend_rows = shared_t ? 1 : 0; g.ep.ld_addend = D;
    V4H_TRY(run_gemm(p, g, s));
  } else {
    GemmDesc g = linear_fwd(ws.temb_in, DT_F32, d.freq_dim, w.t0_w, DT_F32, d.freq_dim, Bt, D, d.freq_dim);
    g.tag = "gemm.cond"; g.act = ACT_SILU; g.ep.bias = w.t0_b; g.ep.out = ws.t_h; g.ep.out2 = ws.t_h_pre; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, s));
    g = linear_fwd(ws.t_h, DT_F32, D, w.t2_w, DT_F32, D, Bt, D, D); g.tag = "gemm.cond";
    g.ep.bias = w.t2_b; g.ep.out = ws.te; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, s));
    g = linear_fwd(cin, DT_F32, d.cond_dim, w.c0_w, DT_F32, d.cond_dim, B, D, d.cond_dim); g.tag = "gemm.cond";
    g.act = ACT_SILU; g.ep.bias = w.c0_b; g.ep.out = ws.c_h; g.ep.out2 = ws.c_h_pre; g.ep.ldo = D;
    V4H_TRY(run_gemm(p, g, s));
    g = linear_fwd(ws.c_h, DT_F32, D, w.c2_w, DT_F32, D, B, D, D); g.tag = "gemm.cond";
    g.act = ACT_SILU; g.ep.bias = w.c2_b; g.ep.out = ws.sc; g.ep.out2 = ws.cond; g.ep.ldo = D;
    g.ep.addend = ws.te; g.ep.addend_rows = shared_t ? 1 : 0; g.ep.ld_addend = D;
    V4H_TRY(run_gemm(p, g, s));
  }
  // ---- every adaLN modulation of the network in one pass: mod = sc Wada^T + bada   (nn/vit.py:328-330, :348)
  if (fast) {
    V4H_TRY(cast_f32_to_bf16(ws.sc, ws.sc_bf16, (int64_t)B * D, s));
    GemmDesc g = linear_fwd(ws.sc_bf16, DT_BF16, D, wa + p.arena_ada, DT_BF16, D, B, p.Nmod, D);
    g.tag = "gemm.adaln";
    g.ep.bias = reinterpret_cast<const float*>(arena + p.arena_ada_bias_bytes);
    g.ep.out = ws.mod; g.ep.ldo = p.Nmod; g.out_dtype = DT_F32;
    V4H_TRY(run_gemm(p, g, s));
  } else {
    for (int i = 0; i <= d.depth; ++i) {
      const bool fin = i == d.depth;
      const int n = fin ? 2 * D : 6 * D;
      GemmDesc g = linear_fwd(ws.sc, DT_F32, D, fin ? w.final_ada_w : w.blocks[i].ada_w, DT_F32, D, B, n, D);
      g.tag = "gemm.adaln";
      g.ep.bias = fin ? w.final_ada_b : w.blocks[i].ada_b;
      g.ep.out = ws.mod + (size_t)i * 6 * D; g.ep.ldo = p.Nmod;
      V4H_TRY(run_gemm(p, g, s));
    }
  }
  if (forked) V4H_TRY(join_from(L, sx, s));  // h0 is needed from here on
  return V4H_OK;
}

// ---- transformer block i   (nn/vit.py:327-333)
template <typename T>
int forward_block(Plan& p, const v4h_vit_params& w, const char* arena, Sub& u, int i, bool train) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  cudaStream_t s = u.s;
  const int B = u.B, Tn = d.tokens, D = d.hidden_dim, Hm = d.mlp_hidden, M = B * Tn;
  const int TA = p.bf16 ? DT_BF16 : DT_F32;
  const bf16* wa = reinterpret_cast<const bf16*>(arena);
  const v4h_block_params& bw = w.blocks[i];
  BlockBufs& bb = ws.blk[train ? i : 0];
  float* hin = ws.h[hidx(train, 2 * i)];
  float* hmid = ws.h[hidx(train, 2 * i + 1)];
  float* hout = ws.h[hidx(train, 2 * i + 2)];
  const float* mod = ws.mod + (size_t)i * 6 * D;
  const void* Wqkv = p.bf16 ? (const void*)(wa + p.arena_blocks[i].qkv) : (const void*)bw.qkv_w;
  const void* Wproj = p.bf16 ? (const void*)(wa + p.arena_blocks[i].proj) : (const void*)bw.proj_w;
  const void* Wfc1 = p.bf16 ? (const void*)(wa + p.arena_blocks[i].fc1) : (const void*)bw.fc1_w;
  const void* Wfc2 = p.bf16 ? (const void*)(wa + p.arena_blocks[i].fc2) : (const void*)bw.fc2_w;

  // LN1 + modulate of block 0 follows the x_embedder; for the later blocks it was applied where their input was
  // produced (the fc2 step of the block before, below)
  if (i == 0)
    V4H_TRY(prof("ln.fwd", 0, (double)M * D * (4 + sizeof(T)), s, [&] { return ln_modulate_fwd<T>(hin, mod + 0 * D, mod + 1 * D, p.Nmod, (T*)bb.a, p.ldx, bb.stats1, M, D, Tn, s); }));
  {
    GemmDesc g = linear_fwd(bb.a, TA, p.ldx, Wqkv, TA, D, M, 3 * D, D);
    g.tag = "gemm.qkv";
    g.ep.bias = bw.qkv_b; g.ep.out = bb.qkv; g.ep.ldo = 3 * D; g.out_dtype = TA;
    V4H_TRY(run_gemm(p, g, s));
  }
  V4H_TRY(prof("attn.fwd", 4.0 * B * d.num_heads * Tn * Tn * (D / d.num_heads), (double)M * 4 * D * sizeof(T), s, [&] { return attn_fwd<T>(p, bb.qkv, bb.o, bb.lse, B, Tn, d.num_heads, D / d.num_heads, s); }));
  {
    GemmDesc g = linear_fwd(bb.o, TA, D, Wproj, TA, D, M, D, D);
    g.tag = "gemm.proj";
    g.epi = EPI_GATE_RES; g.out_dtype = TA;
    g.ep.bias = bw.proj_b; g.ep.out2 = bb.y1; g.ep.ldo = D;
    g.ep.gate = mod + 2 * D; g.ep.mod_stride = p.Nmod; g.ep.rows_per_sample = Tn;
    g.ep.res_in = hin; g.ep.res_out = hmid;
    V4H_TRY(gate_res_then_ln<T>(p, g, mod + 3 * D, mod + 4 * D, bb.m, p.ldx, bb.stats2, M, D, Tn, s));  // + LN2, modulate
  }
  {
    GemmDesc g = linear_fwd(bb.m, TA, p.ldx, Wfc1, TA, D, M, Hm, D);
    g.tag = "gemm.fc1";
    g.act = ACT_GELU_TANH; g.out_dtype = TA;
    g.ep.bias = bw.fc1_b; g.ep.out = bb.g; g.ep.out2 = bb.u; g.ep.ldo = Hm;
    V4H_TRY(run_gemm(p, g, s));
  }
  {
    GemmDesc g = linear_fwd(bb.g, TA, Hm, Wfc2, TA, Hm, M, D, Hm);
    g.tag = "gemm.fc2";
    g.epi = EPI_GATE_RES; g.out_dtype = TA;
    g.ep.bias = bw.fc2_b; g.ep.out2 = bb.y2; g.ep.ldo = D;
    g.ep.gate = mod + 5 * D; g.ep.mod_stride = p.Nmod; g.ep.rows_per_sample = Tn;
    g.ep.res_in = hmid; g.ep.res_out = hout;
    // + the LayerNorm / modulation that consumes this block's output: LN1 of the next block, or the final layer's
    if (i + 1 < d.depth) {
      const float* modn = ws.mod + (size_t)(i + 1) * 6 * D;
      BlockBufs& bn = ws.blk[train ? i + 1 : 0];
      V4H_TRY(gate_res_then_ln<T>(p, g, modn + 0 * D, modn + 1 * D, bn.a, p.ldx, bn.stats1, M, D, Tn, s));
    } else {
      const float* modf = ws.mod + (size_t)d.depth * 6 * D;
      V4H_TRY(gate_res_then_ln<T>(p, g, modf, modf + D, ws.a_f, D, ws.stats_f, M, D, Tn, s));
    }
  }
  return V4H_OK;
}

// ---- final layer   (nn/vit.py:347-351)
template <typename T>
int forward_final(Plan& p, const v4h_vit_params& w, const char* arena, Sub& u, bool train) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  cudaStream_t s = u.s;
  const int B = u.B, Tn = d.tokens, D = d.hidden_dim, M = B * Tn;
  const int TA = p.bf16 ? DT_BF16 : DT_F32;
  const bf16* wa = reinterpret_cast<const bf16*>(arena);
  const bool fast = p.bf16 && p.use_umma;
  // (the final LayerNorm + modulation was applied by the last block's fc2 step into ws.a_f)
  GemmDesc g = fast ? linear_fwd(ws.a_f, TA, D, wa + p.arena_final, DT_BF16, D, M, d.out_dim, D)
                    : linear_fwd(ws.a_f, TA, D, w.final_w, DT_F32, D, M, d.out_dim, D);
  g.tag = "gemm.final";
  g.ep.bias = w.final_b; g.ep.out = u.out; g.ep.ldo = d.out_dim; g.out_dtype = DT_F32;
  V4H_TRY(run_gemm(p, g, s));
  return V4H_OK;
}

// dX (M x Nin) = dY (M x Nout) W (Nout x Nin)
GemmDesc dgrad(const void* dY, int dy_dt, int ld_dy, const void* W, int w_dt, int ldw, int M, int Nin, int Nout,
               const char* tag = "dgrad") {
  GemmDesc g;
  g.tag = tag;
  g.layout = GEMM_NN; g.A = dY; g.a_dtype = dy_dt; g.lda = ld_dy; g.B = W; g.b_dtype = w_dt; g.ldb = ldw;
  g.M = M; g.N = Nin; g.K = Nout;
  return g;
}

// side-stream weight gradients of one sub-batch: fork after the kernel that produced dY, join before the
// buffers a block's wgrads read are written again (start of the next stage) and before returning
template <typename F>
int on_side(const Plan& p, Sub& u, F&& launch) {
  if (!p.forking()) return launch(u.s);
  V4H_CUDA(cudaEventRecord(u.lane->ev_fork, u.s));
  V4H_CUDA(cudaStreamWaitEvent(u.lane->side, u.lane->ev_fork, 0));
  u.side_busy = true;
  return launch(u.lane->side);
}
int join_side(Sub& u) {
  if (!u.side_busy) return V4H_OK;
  V4H_CUDA(cudaEventRecord(u.lane->ev_join, u.lane->side));
  V4H_CUDA(cudaStreamWaitEvent(u.s, u.lane->ev_join, 0));
  u.side_busy = false;
  return V4H_OK;
}
// second side stream: work that nothing in the chain waits for (its inputs are final and never rewritten), joined
// only when the call returns
template <typename F>
int on_side2(const Plan& p, Sub& u, F&& launch) {
  if (!p.forking()) return launch(u.s);
  V4H_CUDA(cudaEventRecord(u.lane->ev_fork, u.s));
  V4H_CUDA(cudaStreamWaitEvent(u.lane->side2, u.lane->ev_fork, 0));
  u.side2_busy = true;
  return launch(u.lane->side2);
}
int join_side2(Sub& u) {
  if (!u.side2_busy) return V4H_OK;
  V4H_CUDA(cudaEventRecord(u.lane->ev_join2, u.lane->side2));
  V4H_CUDA(cudaStreamWaitEvent(u.s, u.lane->ev_join2, 0));
  u.side2_busy = false;
  return V4H_OK;
}

// adaLN modulation Linear `i` (depth = the final layer's): d W = dmod_i^T sc, d b = colsum(dmod_i).  The
// modulation gradients of a block are complete when its backward stage ends, so its weight gradient is issued
// there (side stream) and lands in that block's slice of the flat gradient buffer: nothing but the small
// embedding / conditioning Linears is left for stage 0, whose data-parallel bucket is the one nothing hides.
int ada_wgrad(Plan& p, const v4h_vit_params& gr, Sub& u, int i) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  const int B = u.B, D = d.hidden_dim;
  const bool fin = i == d.depth;
  const int n = fin ? 2 * D : 6 * D;
  const size_t off = (size_t)i * 6 * D;
  float* dW = fin ? gr.final_ada_w : gr.blocks[i].ada_w;
  float* db = fin ? gr.final_ada_b : gr.blocks[i].ada_b;
  // K = B is one or two k-steps: launch-latency-bound whatever the engine, so the fp32 SIMT GEMM reads the fp32
  // modulation gradients and SiLU(cond) in place (no bf16 copies, exact fp32 like the reference)
  // (second side stream: the modulation gradients of a finished block and SiLU(cond) are never written again, so the
  // chain does not wait for these launches at the next stage)
  V4H_TRY(on_side2(p, u, [&](cudaStream_t q) { return wgrad(p, ws.dmod + off, DT_F32, p.Nmod, ws.sc, DT_F32, D, dW, n, D, B, q, "wgrad.adaln"); }));
  V4H_TRY(on_side2(p, u, [&](cudaStream_t q) { return prof("colsum", 0, 0, q, [&] { return colsum_add<float>(ws.dmod + off, p.Nmod, db, B, n, q); }); }));
  return V4H_OK;
}

// one backward stage of one sub-batch: depth + 1 = final layer, depth .. 1 = blocks depth-1 .. 0
template <typename T>
int backward_stage(Plan& p, const v4h_vit_params& w, const char* arena, const v4h_vit_params& gr, Sub& u, int stage) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  cudaStream_t s = u.s;
  const int B = u.B, Tn = d.tokens, D = d.hidden_dim, Hm = d.mlp_hidden, M = B * Tn;
  const int TA = p.bf16 ? DT_BF16 : DT_F32;
  const bf16* wa = reinterpret_cast<const bf16*>(arena);
  const int H = d.num_heads, dh = D / H;
  const float* dout = u.dout;

  V4H_TRY(join_side(u));
  if (stage == d.depth + 1) {
    // ---------------- final layer
    V4H_CUDA(cudaMemsetAsync(ws.dmod, 0, (size_t)B * p.Nmod * sizeof(float), s));
    const size_t offF = (size_t)d.depth * 6 * D;
    const bool fast = p.bf16 && p.use_umma;
    if (fast) V4H_TRY(cast_f32_to_bf16(dout, ws.dout_bf, (int64_t)M * d.out_dim, s));
    const void* dY = fast ? (const void*)ws.dout_bf : (const void*)dout;
    const int dy_dt = fast ? DT_BF16 : DT_F32;
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, dY, dy_dt, d.out_dim, ws.a_f, TA, D, gr.final_w, d.out_dim, D, M, q, "wgrad.final"); }));
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return prof("colsum", 0, 0, q, [&] { return colsum_add<float>(dout, d.out_dim, gr.final_b, M, d.out_dim, q); }); }));
    {
      GemmDesc g = fast ? dgrad(dY, DT_BF16, d.out_dim, wa + p.arena_final, DT_BF16, D, M, D, d.out_dim, "dgrad.final")
                        : dgrad(dout, DT_F32, d.out_dim, w.final_w, DT_F32, D, M, D, d.out_dim, "dgrad.final");
      g.ep.out = ws.dm; g.ep.ldo = D; g.out_dtype = TA;
      V4H_TRY(run_gemm(p, g, s));
    }
    const int last = d.depth - 1;
    const float* modL = ws.mod + (size_t)last * 6 * D;
    float* dmodL = ws.dmod + (size_t)last * 6 * D;
    V4H_TRY(prof("ln.bwd", 0, (double)M * D * (12 + 3 * sizeof(T)), s, [&] { return ln_modulate_bwd<T>((const T*)ws.dm, ws.h[2 * d.depth], ws.stats_f, ws.mod + offF + D, p.Nmod,
                               ws.dh, false, ws.dmod + offF, ws.dmod + offF + D, p.Nmod,
                               (const T*)ws.blk[last].y2, modL + 5 * D, (T*)ws.dy_mlp[last & 1], dmodL + 5 * D,
                               gr.blocks[last].fc2_b, M, D, Tn, s); }));
    V4H_TRY(ada_wgrad(p, gr, u, d.depth));  // d shift / d scale of the final layer are complete
    return V4H_OK;
  }
  if (stage < 1) return fail(V4H_ERR_INVALID, "internal: stage 0 must be run through backward_stage0");
  // ---------------- transformer block i
  const int i = stage - 1;
  const v4h_block_params& bw = w.blocks[i];
  const v4h_block_params& bg = gr.blocks[i];
  BlockBufs& bb = ws.blk[i];
  const float* mod = ws.mod + (size_t)i * 6 * D;
  float* dmod = ws.dmod + (size_t)i * 6 * D;
  const void* Wqkv = p.bf16 ? (const void*)(wa + p.arena_blocks[i].qkv) : (const void*)bw.qkv_w;
  const void* Wproj = p.bf16 ? (const void*)(wa + p.arena_blocks[i].proj) : (const void*)bw.proj_w;
  const void* Wfc1 = p.bf16 ? (const void*)(wa + p.arena_blocks[i].fc1) : (const void*)bw.fc1_w;
  const void* Wfc2 = p.bf16 ? (const void*)(wa + p.arena_blocks[i].fc2) : (const void*)bw.fc2_w;

  // MLP branch: dy = gate_mlp * dh is ready in dy_mlp[i & 1]; the attention branch uses dy_attn, and the
  // LayerNorm backward that closes this block writes dy_mlp[(i - 1) & 1]: no buffer a side-stream wgrad of
  // this block reads is rewritten before the join at the start of the next block
  void* dy1 = ws.dy_mlp[i & 1];
  void* dy2 = ws.dy_attn;
  V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, dy1, TA, D, bb.g, TA, Hm, bg.fc2_w, D, Hm, M, q, "wgrad.fc2"); }));
  {
    GemmDesc g = dgrad(dy1, TA, D, Wfc2, TA, Hm, M, Hm, D, "dgrad.fc2");
    g.epi = EPI_DACT; g.act = ACT_GELU_TANH; g.out_dtype = TA;
    g.ep.out = ws.du; g.ep.ldo = Hm; g.ep.aux = bb.u; g.ep.ld_aux = Hm;
    V4H_TRY(run_gemm(p, g, s));
  }
  if (p.ldx > D) {  // the ones column of m: fc1 bias gradient out of the same GEMM
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, ws.du, TA, Hm, bb.m, TA, p.ldx, bg.fc1_w, Hm, D, M, q, "wgrad.fc1", bg.fc1_b); }));
  } else {
    V4H_TRY(prof("colsum", 0, 0, s, [&] { return colsum_add<T>((const T*)ws.du, Hm, bg.fc1_b, M, Hm, s); }));
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, ws.du, TA, Hm, bb.m, TA, D, bg.fc1_w, Hm, D, M, q, "wgrad.fc1"); }));
  }
  {
    GemmDesc g = dgrad(ws.du, TA, Hm, Wfc1, TA, D, M, D, Hm, "dgrad.fc1");
    g.ep.out = ws.dm; g.ep.ldo = D; g.out_dtype = TA;
    V4H_TRY(run_gemm(p, g, s));
  }
  V4H_TRY(prof("ln.bwd", 0, (double)M * D * (12 + 3 * sizeof(T)), s, [&] { return ln_modulate_bwd<T>((const T*)ws.dm, ws.h[2 * i + 1], bb.stats2, mod + 4 * D, p.Nmod, ws.dh, true,
                             dmod + 3 * D, dmod + 4 * D, p.Nmod, (const T*)bb.y1, mod + 2 * D, (T*)dy2,
                             dmod + 2 * D, bg.proj_b, M, D, Tn, s); }));
  // attention branch: dy = gate_msa * dh
  V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, dy2, TA, D, bb.o, TA, D, bg.proj_w, D, D, M, q, "wgrad.proj"); }));
  {
    GemmDesc g = dgrad(dy2, TA, D, Wproj, TA, D, M, D, D, "dgrad.proj");
    g.ep.out = ws.dm; g.ep.ldo = D; g.out_dtype = TA;
    V4H_TRY(run_gemm(p, g, s));
  }
  V4H_TRY(prof("attn.bwd", 10.0 * B * H * Tn * Tn * dh, (double)M * 9 * D * sizeof(T), s, [&] { return attn_bwd<T>(p, bb.qkv, bb.o, bb.lse, ws.dm, ws.attn_delta, ws.dqkv, B, Tn, H, dh, s); }));
  if (p.ldx > D) {
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, ws.dqkv, TA, 3 * D, bb.a, TA, p.ldx, bg.qkv_w, 3 * D, D, M, q, "wgrad.qkv", bg.qkv_b); }));
  } else {
    V4H_TRY(prof("colsum", 0, 0, s, [&] { return colsum_add<T>((const T*)ws.dqkv, 3 * D, bg.qkv_b, M, 3 * D, s); }));
    V4H_TRY(on_side(p, u, [&](cudaStream_t q) { return wgrad(p, ws.dqkv, TA, 3 * D, bb.a, TA, D, bg.qkv_w, 3 * D, D, M, q, "wgrad.qkv"); }));
  }
  {
    GemmDesc g = dgrad(ws.dqkv, TA, 3 * D, Wqkv, TA, D, M, D, 3 * D, "dgrad.qkv");
    g.ep.out = ws.dm; g.ep.ldo = D; g.out_dtype = TA;
    V4H_TRY(run_gemm(p, g, s));
  }
  if (i > 0) {
    const float* modP = ws.mod + (size_t)(i - 1) * 6 * D;
    float* dmodP = ws.dmod + (size_t)(i - 1) * 6 * D;
    V4H_TRY(prof("ln.bwd", 0, (double)M * D * (12 + 3 * sizeof(T)), s, [&] { return ln_modulate_bwd<T>((const T*)ws.dm, ws.h[2 * i], bb.stats1, mod + 1 * D, p.Nmod, ws.dh, true,
                               dmod + 0 * D, dmod + 1 * D, p.Nmod, (const T*)ws.blk[i - 1].y2, modP + 5 * D,
                               (T*)ws.dy_mlp[(i - 1) & 1], dmodP + 5 * D, gr.blocks[i - 1].fc2_b, M, D, Tn, s); }));
  } else {
    V4H_TRY(prof("ln.bwd", 0, (double)M * D * (12 + 3 * sizeof(T)), s, [&] { return ln_modulate_bwd<T>((const T*)ws.dm, ws.h[0], bb.stats1, mod + 1 * D, p.Nmod, ws.dh, true,
                               dmod + 0 * D, dmod + 1 * D, p.Nmod, (const T*)nullptr, nullptr, (T*)nullptr,
                               nullptr, nullptr, M, D, Tn, s); }));
  }
  // every modulation gradient of block i is final: d gate_mlp came from the LayerNorm backward that closed the
  // stage above, the other five from this stage
  V4H_TRY(ada_wgrad(p, gr, u, i));
  return V4H_OK;
}

// stage 0: embeddings and conditioning.  ws.dh = d loss / d h0
int backward_stage0(Plan& p, const v4h_vit_params& w, const char* arena, const v4h_vit_params& gr, Sub& u) {
  const v4h_vit_dims& d = p.d;
  Workspace& ws = u.ws;
  cudaStream_t s = u.s;
  const Lane& L = *u.lane;
  const int B = u.B, Tn = d.tokens, D = d.hidden_dim, M = B * Tn;
  const bool fast = p.bf16 && p.use_umma;
  const bf16* wa = reinterpret_cast<const bf16*>(arena);
  V4H_TRY(join_side(u));
  // This stage is a long list of small kernels (a few CTAs each).  Only dgrad.adaln -> d SiLU -> the two
  // MLP dgrads form a chain; the embedding gradients of x and every weight gradient hang off it, so with the
  // side streams they run next to it: sx = x / positional embedding, sw = weight gradients of the second MLP Linears
  const bool forked = fast && p.forking();
  cudaStream_t sx = forked ? L.side2 : s, sw = forked ? L.side : s;
  if (forked) V4H_TRY(fork_to(L, s, sx));
  // x_embedder: dW = dh0^T x, db = colsum(dh0); the network input is only differentiated for a finetuning mapper
  const float* xin = d.x_map_dim > 0 ? ws.xm : u.x;
  if (fast) {
    V4H_TRY(cast_f32_to_bf16(ws.dh, ws.dh_bf, (int64_t)M * D, sx));
    V4H_TRY(wgrad(p, ws.dh_bf, DT_BF16, D, ws.x_bf, DT_BF16, d.patch_dim, gr.x_w, D, d.patch_dim, M, sx, "wgrad.x_embed"));
  } else {
    V4H_TRY(wgrad(p, ws.dh, DT_F32, D, xin, DT_F32, d.patch_dim, gr.x_w, D, d.patch_dim, M, sx, "wgrad.x_embed"));
  }
  V4H_TRY(prof("colsum", 0, 0, sx, [&] { return colsum_add<float>(ws.dh, D, gr.x_b, M, D, sx); }));
  if (d.x_map_dim > 0) {
    // d xm_pre = (dh0 Wx) * SiLU'(xm_pre);  d Wm = d xm_pre^T x;  d bm = colsum(d xm_pre)
    GemmDesc g = dgrad(ws.dh, DT_F32, D, w.x_w, DT_F32, d.patch_dim, M, d.patch_dim, D, "dgrad.x_embed");
    g.epi = EPI_DACT; g.act = ACT_SILU; g.ep.out = ws.dxm; g.ep.ldo = d.patch_dim; g.ep.aux = ws.xm_pre; g.ep.ld_aux = d.patch_dim;
    V4H_TRY(run_gemm(p, g, sx));
    V4H_TRY(wgrad(p, ws.dxm, DT_F32, d.patch_dim, u.x, DT_F32, d.x_map_dim, gr.xm_w, d.patch_dim, d.x_map_dim, M, sx, "wgrad.x_map"));
    V4H_TRY(prof("colsum", 0, 0, sx, [&] { return colsum_add<float>(ws.dxm, d.patch_dim, gr.xm_b, M, d.patch_dim, sx); }));
  }
  if (d.learn_pos_embed) {
    // sum d h0 over the batch first (a column sum of the (B, T*D) view, into the now idle PE buffer), then
    // one pass over (T, D) applies d PE / d freq
    V4H_CUDA(cudaMemsetAsync(ws.pe, 0, (size_t)Tn * D * sizeof(float), sx));
    V4H_TRY(prof("colsum", 0, 0, sx, [&] { return colsum_add<float>(ws.dh, Tn * D, ws.pe, B, Tn * D, sx); }));
    V4H_TRY(pos_embedding_bwd(ws.pe, w.pos_embed_freqs, w.pos_z, w.pos_y, w.pos_x, gr.pos_embed_freqs, 1, Tn,
                              D / 6, sx));
  }
  // adaLN Linears: d sc = dmod Wada (their weight gradients were issued stage by stage, ada_wgrad)
  V4H_CUDA(cudaMemsetAsync(ws.dsc, 0, (size_t)B * D * sizeof(float), s));
  if (fast) {
    V4H_TRY(cast_f32_to_bf16(ws.dmod, ws.dmod_bf16, (int64_t)B * p.Nmod, s));
    GemmDesc g = dgrad(ws.dmod_bf16, DT_BF16, p.Nmod, wa + p.arena_ada, DT_BF16, D, B, D, p.Nmod, "dgrad.adaln");
    g.epi = EPI_ATOMIC; g.ep.out = ws.dsc; g.ep.ldo = D; g.splitk = 0;
    V4H_TRY(run_gemm(p, g, s));
  } else {
    for (int i = 0; i <= d.depth; ++i) {
      const bool fin = i == d.depth;
      const int n = fin ? 2 * D : 6 * D;
      GemmDesc g = dgrad(ws.dmod + (size_t)i * 6 * D, DT_F32, p.Nmod, fin ? w.final_ada_w : w.blocks[i].ada_w, DT_F32, D, B,
                         D, n, "dgrad.adaln");
      g.epi = EPI_ATOMIC; g.ep.out = ws.dsc; g.ep.ldo = D; g.splitk = std::max(1, n / 240);
      V4H_TRY(run_gemm(p, g, s));
    }
  }
  V4H_TRY(dsilu_mul(ws.dsc, ws.cond, ws.dcond, fast ? ws.dcond_bf : nullptr, (int64_t)B * D, s));
  // c_embedder (nn/vit.py:77-81) and t_embedder.mlp (nn/vit.py:361-365): Linear -> SiLU -> Linear
  const float* cin = d.c_map_dim > 0 ? ws.cm : u.c;
  const void* dvec_c = nullptr;  // d c_embedder.0 output (B, D), for the mapper below
  int dvec_c_dt = DT_F32;
  if (fast) {
    struct Mlp { const void* in; int in_dt, in_dim; const bf16 *h_pre, *h; const bf16* w2; float *dw0, *db0, *dw2, *db2; };
    Mlp mlps[2] = {
        {cin, DT_F32, d.cond_dim, ws.c_hpre_bf, ws.c_h_bf, wa + p.arena_c2, gr.c0_w, gr.c0_b, gr.c2_w, gr.c2_b},
        {ws.temb_bf, DT_BF16, d.freq_dim, ws.t_hpre_bf, ws.t_h_bf, wa + p.arena_t2, gr.t0_w, gr.t0_b, gr.t2_w, gr.t2_b}};
    if (forked) V4H_TRY(fork_to(L, s, sw));  // d cond is ready
    for (const Mlp& m : mlps) {
      V4H_TRY(wgrad(p, ws.dcond_bf, DT_BF16, D, m.h, DT_BF16, D, m.dw2, D, D, B, sw, "wgrad.cond"));
      V4H_TRY(prof("colsum", 0, 0, sw, [&] { return colsum_add<float>(ws.dcond, D, m.db2, B, D, sw); }));
    }
    // the two MLPs are independent from here on: the t_embedder's chain (dgrad -> weight gradient) runs on the
    // x side stream next to the c_embedder's on the caller's stream, each with its own d hidden buffer
    if (forked) V4H_TRY(fork_to(L, s, sx));
    for (int k = 0; k < 2; ++k) {
      const Mlp& m = mlps[k];
      cudaStream_t q = k == 1 ? sx : s;
      bf16* dvec = k == 1 ? ws.dvec2_bf : ws.dvec_bf;
      GemmDesc g = dgrad(ws.dcond_bf, DT_BF16, D, m.w2, DT_BF16, D, B, D, D, "dgrad.cond");
      g.epi = EPI_DACT; g.act = ACT_SILU; g.out_dtype = DT_BF16;
      g.ep.out = dvec; g.ep.ldo = D; g.ep.aux = m.h_pre; g.ep.ld_aux = D;
      V4H_TRY(run_gemm(p, g, q));
      V4H_TRY(wgrad(p, dvec, DT_BF16, D, m.in, m.in_dt, m.in_dim, m.dw0, D, m.in_dim, B, q, "wgrad.cond"));
      V4H_TRY(prof("colsum", 0, 0, q, [&] { return colsum_add<bf16>(dvec, D, m.db0, B, D, q); }));
    }
    dvec_c = ws.dvec_bf; dvec_c_dt = DT_BF16;
  } else {
    struct Mlp { const float *in, *h_pre, *h; int in_dim; const float* w2; float *dw0, *db0, *dw2, *db2; };
    Mlp mlps[2] = {
        {ws.temb_in, ws.t_h_pre, ws.t_h, d.freq_dim, w.t2_w, gr.t0_w, gr.t0_b, gr.t2_w, gr.t2_b},
        {cin, ws.c_h_pre, ws.c_h, d.cond_dim, w.c2_w, gr.c0_w, gr.c0_b, gr.c2_w, gr.c2_b}};
    for (const Mlp& m : mlps) {  // c_embedder last: its d hidden stays in ws.dvec for the mapper
      V4H_TRY(wgrad(p, ws.dcond, DT_F32, D, m.h, DT_F32, D, m.dw2, D, D, B, s, "wgrad.cond"));
      V4H_TRY(prof("colsum", 0, 0, s, [&] { return colsum_add<float>(ws.dcond, D, m.db2, B, D, s); }));
      GemmDesc g = dgrad(ws.dcond, DT_F32, D, m.w2, DT_F32, D, B, D, D, "dgrad.cond");
      g.epi = EPI_DACT; g.act = ACT_SILU; g.ep.out = ws.dvec; g.ep.ldo = D; g.ep.aux = m.h_pre; g.ep.ld_aux = D;
      V4H_TRY(run_gemm(p, g, s));
      V4H_TRY(wgrad(p, ws.dvec, DT_F32, D, m.in, DT_F32, m.in_dim, m.dw0, D, m.in_dim, B, s, "wgrad.cond"));
      V4H_TRY(prof("colsum", 0, 0, s, [&] { return colsum_add<float>(ws.dvec, D, m.db0, B, D, s); }));
    }
    dvec_c = ws.dvec;
  }
  if (d.c_map_dim > 0) {
    // d cm_pre = (d c_h_pre Wc0) * SiLU'(cm_pre);  d Wm = d cm_pre^T c;  d bm = colsum(d cm_pre)   (caller's stream:
    // the c_embedder chain above ran there)
    GemmDesc g = dgrad(dvec_c, dvec_c_dt, D, w.c0_w, DT_F32, d.cond_dim, B, d.cond_dim, D, "dgrad.cond");
    g.epi = EPI_DACT; g.act = ACT_SILU; g.ep.out = ws.dcm; g.ep.ldo = d.cond_dim; g.ep.aux = ws.cm_pre; g.ep.ld_aux = d.cond_dim;
    V4H_TRY(run_gemm(p, g, s));
    V4H_TRY(wgrad(p, ws.dcm, DT_F32, d.cond_dim, u.c, DT_F32, d.c_map_dim, gr.cm_w, d.cond_dim, d.c_map_dim, B, s, "wgrad.c_map"));
    V4H_TRY(prof("colsum", 0, 0, s, [&] { return colsum_add<float>(ws.dcm, d.cond_dim, gr.cm_b, B, d.cond_dim, s); }));
  }
  if (forked) { V4H_TRY(join_from(L, sw, s)); V4H_TRY(join_from(L, sx, s)); }
  return V4H_OK;
}

// Split a call over B samples into the plan's sub-batches: slices of the inputs, one workspace each (laid out
// back to back in the caller's buffer), lane 0 on the caller's stream and lane 1 on its own stream.
int make_subs(Plan& p, char* workspace, size_t workspace_bytes, int64_t B, bool train, bool shared_t, const float* x,
              const float* t, const float* c, float* out, const float* dout, cudaStream_t s, Sub (&subs)[2], int* nsub) {
  const v4h_vit_dims& d = p.d;
  const int n = p.sub_batches(B);
  const int in_dim = d.x_map_dim > 0 ? d.x_map_dim : d.patch_dim;
  const int c_dim = d.c_map_dim > 0 ? d.c_map_dim : d.cond_dim;
  size_t off = 0;
  int64_t b0 = 0;
  for (int k = 0; k < n; ++k) {
    Sub& u = subs[k];
    const int64_t bk = n == 1 ? B : (k == 0 ? (B + 1) / 2 : B - (B + 1) / 2);
    u.B = (int)bk;
    u.x = x ? x + (size_t)b0 * d.tokens * in_dim : nullptr;
    u.t = t ? t + (shared_t ? 0 : b0) : nullptr;
    u.c = c ? c + (size_t)b0 * c_dim : nullptr;
    u.out = out ? out + (size_t)b0 * d.tokens * d.out_dim : nullptr;
    u.dout = dout ? dout + (size_t)b0 * d.tokens * d.out_dim : nullptr;
    u.ws.layout(p, workspace + off, bk, train);
    off += align_up(u.ws.bytes, 1024);
    u.lane = &p.lanes[k];
    u.s = (k == 0 || !p.forking()) ? s : p.lanes[k].main;
    b0 += bk;
  }
  V4H_REQUIRE(off <= workspace_bytes, "vit: workspace too small (%zu < %zu)", workspace_bytes, off);
  *nsub = n;
  return V4H_OK;
}
// lane 1 starts after everything issued on the caller's stream so far / the caller's stream waits for lane 1
int split_lanes(Plan& p, Sub (&subs)[2], int nsub, cudaStream_t s) {
  if (nsub < 2 || subs[1].s == s) return V4H_OK;
  V4H_CUDA(cudaEventRecord(p.ev_split, s));
  V4H_CUDA(cudaStreamWaitEvent(subs[1].s, p.ev_split, 0));
  return V4H_OK;
}
int merge_lanes(Plan& p, Sub (&subs)[2], int nsub, cudaStream_t s) {
  if (nsub < 2 || subs[1].s == s) return V4H_OK;
  V4H_CUDA(cudaEventRecord(p.lanes[1].ev_done, subs[1].s));
  V4H_CUDA(cudaStreamWaitEvent(s, p.lanes[1].ev_done, 0));
  return V4H_OK;
}

template <typename T>
int forward_all(Plan& p, const v4h_vit_params& w, const char* arena, Sub (&subs)[2], int nsub, bool shared_t, bool train,
                cudaStream_t s) {
  V4H_TRY(split_lanes(p, subs, nsub, s));
  // issue order alternates between the sub-batches stage by stage, so that both streams have work queued early
  // whether the launches come from the host (eager) or from a replayed graph
  for (int k = 0; k < nsub; ++k) V4H_TRY(forward_prologue<T>(p, w, arena, subs[k], shared_t, train));
  for (int i = 0; i < p.d.depth; ++i)
    for (int k = 0; k < nsub; ++k) V4H_TRY(forward_block<T>(p, w, arena, subs[k], i, train));
  for (int k = 0; k < nsub; ++k) V4H_TRY(forward_final<T>(p, w, arena, subs[k], train));
  return merge_lanes(p, subs, nsub, s);
}

template <typename T>
int backward_all(Plan& p, const v4h_vit_params& w, const char* arena, const v4h_vit_params& gr, Sub (&subs)[2], int nsub,
                 int stage_begin, int stage_end, cudaStream_t s) {
  V4H_TRY(split_lanes(p, subs, nsub, s));
  const int lo = stage_end < 1 ? 1 : stage_end;
  for (int stage = stage_begin; stage >= lo; --stage)
    for (int k = 0; k < nsub; ++k) V4H_TRY(backward_stage<T>(p, w, arena, gr, subs[k], stage));
  if (stage_end == 0)
    for (int k = 0; k < nsub; ++k) V4H_TRY(backward_stage0(p, w, arena, gr, subs[k]));
  for (int k = 0; k < nsub; ++k) { V4H_TRY(join_side(subs[k])); V4H_TRY(join_side2(subs[k])); }
  return merge_lanes(p, subs, nsub, s);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
int plan_create(const v4h_vit_dims* dims, Plan** out) {
  V4H_REQUIRE(dims && out, "plan_create: null argument");
  const v4h_vit_dims& d = *dims;
  V4H_REQUIRE(d.hidden_dim > 0 && d.depth > 0 && d.depth <= V4H_MAX_DEPTH && d.num_heads > 0 &&
                  d.hidden_dim % d.num_heads == 0 && d.mlp_hidden > 0 && d.patch_dim > 0 && d.out_dim > 0 &&
                  d.cond_dim > 0 && d.tokens > 0 && d.freq_dim >= 2,
              "plan_create: invalid dimensions");
  V4H_REQUIRE(!d.learn_pos_embed || d.hidden_dim % 6 == 0, "plan_create: hidden_dim must be a multiple of 6");
  V4H_REQUIRE(d.precision == V4H_FP32 || d.precision == V4H_BF16, "plan_create: unknown precision");
  V4H_REQUIRE(d.x_map_dim >= 0 && d.c_map_dim >= 0, "plan_create: negative mapper width");
  if (d.hidden_dim > 512)
    return fail(V4H_ERR_UNSUPPORTED, "hidden_dim %d > 512 is not supported by the LayerNorm kernels", d.hidden_dim);
  if (d.hidden_dim / d.num_heads > 128)
    return fail(V4H_ERR_UNSUPPORTED, "head_dim %d > 128 is not supported", d.hidden_dim / d.num_heads);
  Plan* p = new Plan();
  p->d = d;
  p->Nmod = d.depth * 6 * d.hidden_dim + 2 * d.hidden_dim;
  p->bf16 = d.precision == V4H_BF16;
  const char* no_umma = getenv("V4H_DISABLE_UMMA");
  p->use_umma = p->bf16 && !(no_umma && no_umma[0] == '1');
  if (p->use_umma) p->umma = umma_context_create();
  {
    const char* e = getenv("V4H_WGRAD_STREAM");
    if (p->use_umma && !(e && e[0] == '0')) {
      bool ok = cudaEventCreateWithFlags(&p->ev_split, cudaEventDisableTiming) == cudaSuccess;
      for (int k = 0; k < 2 && ok; ++k) {
        Lane& L = p->lanes[k];
        ok = (k == 0 || cudaStreamCreateWithFlags(&L.main, cudaStreamNonBlocking) == cudaSuccess) &&
             cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&L.side2, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&L.ev_join, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&L.ev_join2, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming) == cudaSuccess;
      }
      if (!ok) {
        cudaGetLastError();
        p->lanes[0].side = nullptr;  // no side streams: everything stays on the caller's stream
      }
    }
    const char* mb = getenv("V4H_MICROBATCH");
    if (mb && (mb[0] == '0' || mb[0] == '1')) p->microbatch = 1;
    else if (mb && mb[0] == '2') p->microbatch = 2;
  }
  p->ldx = (p->use_umma && d.hidden_dim % 8 == 0) ? d.hidden_dim + 8 : d.hidden_dim;
  const char* no_umma_attn = getenv("V4H_DISABLE_UMMA_ATTN");
  p->use_umma_attn = p->use_umma && attention_umma_supported(d.hidden_dim / d.num_heads) &&
                     !(no_umma_attn && no_umma_attn[0] == '1');
  if (p->bf16) {
    const size_t D = d.hidden_dim, Hm = d.mlp_hidden;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 128); return o; };
    for (int i = 0; i < d.depth; ++i) {
      p->arena_blocks[i].qkv = take(3 * D * D);
      p->arena_blocks[i].proj = take(D * D);
      p->arena_blocks[i].fc1 = take(Hm * D);
      p->arena_blocks[i].fc2 = take(D * Hm);
    }
    p->arena_ada = take((size_t)p->Nmod * D);
    p->arena_final = take((size_t)d.out_dim * D);
    p->arena_x = take(D * (size_t)d.patch_dim);
    p->arena_t0 = take(D * (size_t)d.freq_dim);
    p->arena_t2 = take(D * D);
    p->arena_c2 = take(D * D);
    p->arena_bf16_elems = off;
    p->arena_ada_bias_bytes = align_up(off * 2, 256);
    p->arena_bytes = p->arena_ada_bias_bytes + align_up((size_t)p->Nmod * 4, 256);
    p->njobs = 4 * d.depth + d.depth + 1 + 5;
    if (cudaMalloc(&p->jobs_dev, sizeof(CastJob) * p->njobs) != cudaSuccess ||
        cudaMallocHost(&p->jobs_host, sizeof(CastJob) * p->njobs) != cudaSuccess) {
      delete p;
      return fail(V4H_ERR_CUDA, "plan_create: cannot allocate the cast job table");
    }
    memset(p->jobs_host, 0, sizeof(CastJob) * p->njobs);
  }
  *out = p;
  return V4H_OK;
}

void plan_destroy(Plan* p) {
  if (!p) return;
  if (p->jobs_dev) cudaFree(p->jobs_dev);
  if (p->jobs_host) cudaFreeHost(p->jobs_host);
  if (p->umma) umma_context_destroy(p->umma);
  if (p->ev_split) cudaEventDestroy(p->ev_split);
  for (Lane& L : p->lanes) {
    for (cudaEvent_t e : {L.ev_fork, L.ev_join, L.ev_join2, L.ev_done})
      if (e) cudaEventDestroy(e);
    for (cudaStream_t q : {L.main, L.side, L.side2})
      if (q) cudaStreamDestroy(q);
  }
  delete p;
}

size_t plan_workspace_bytes(const Plan* p, int64_t B, bool train) {
  const int n = p->sub_batches(B);
  size_t total = 0;
  for (int k = 0; k < n; ++k) {
    const int64_t bk = n == 1 ? B : (k == 0 ? (B + 1) / 2 : B - (B + 1) / 2);
    Workspace ws;
    ws.layout(*p, nullptr, bk, train);
    total += align_up(ws.bytes, 1024);
  }
  return total;
}

size_t plan_arena_bytes(const Plan* p) { return p->arena_bytes; }

// byte offset of the bf16 copy of a parameter inside the weight arena (-1: none)
int64_t plan_arena_offset(const Plan* p, const char* field) {
  if (!p->bf16) return -1;
  const size_t D = p->d.hidden_dim;
  const std::string f(field);
  size_t elems;
  if (f == "final_w") elems = p->arena_final;
  else if (f == "x_w") elems = p->arena_x;
  else if (f == "t0_w") elems = p->arena_t0;
  else if (f == "t2_w") elems = p->arena_t2;
  else if (f == "c2_w") elems = p->arena_c2;
  else if (f == "final_ada_w") elems = p->arena_ada + (size_t)p->d.depth * 6 * D * D;
  else if (f == "final_ada_b") return (int64_t)(p->arena_ada_bias_bytes + (size_t)p->d.depth * 6 * D * 4);
  else if (f.rfind("blocks.", 0) == 0) {
    const size_t dot = f.find('.', 7);
    if (dot == std::string::npos) return -1;
    const int i = atoi(f.substr(7, dot - 7).c_str());
    if (i < 0 || i >= p->d.depth) return -1;
    const std::string name = f.substr(dot + 1);
    if (name == "qkv_w") elems = p->arena_blocks[i].qkv;
    else if (name == "proj_w") elems = p->arena_blocks[i].proj;
    else if (name == "fc1_w") elems = p->arena_blocks[i].fc1;
    else if (name == "fc2_w") elems = p->arena_blocks[i].fc2;
    else if (name == "ada_w") elems = p->arena_ada + (size_t)i * 6 * D * D;
    else if (name == "ada_b") return (int64_t)(p->arena_ada_bias_bytes + (size_t)i * 6 * D * 4);
    else return -1;
  } else {
    return -1;
  }
  return (int64_t)(elems * 2);
}

int plan_prepare_weights(Plan* p, const v4h_vit_params* w, void* arena, cudaStream_t s) {
  if (!p->bf16) return V4H_OK;
  V4H_REQUIRE(w && arena, "prepare_weights: null argument");
  const v4h_vit_dims& d = p->d;
  const size_t D = d.hidden_dim, Hm = d.mlp_hidden;
  bf16* wa = reinterpret_cast<bf16*>(arena);
  std::vector<CastJob> jobs;
  int64_t max_n = 0;
  auto add = [&](const float* src, bf16* dst, size_t n) {
    jobs.push_back(CastJob{src, dst, (int64_t)n});
    if ((int64_t)n > max_n) max_n = (int64_t)n;
  };
  for (int i = 0; i < d.depth; ++i) {
    const v4h_block_params& b = w->blocks[i];
    V4H_REQUIRE(b.qkv_w && b.proj_w && b.fc1_w && b.fc2_w && b.ada_w && b.ada_b, "prepare_weights: null block weight");
    add(b.qkv_w, wa + p->arena_blocks[i].qkv, 3 * D * D);
    add(b.proj_w, wa + p->arena_blocks[i].proj, D * D);
    add(b.fc1_w, wa + p->arena_blocks[i].fc1, Hm * D);
    add(b.fc2_w, wa + p->arena_blocks[i].fc2, D * Hm);
    add(b.ada_w, wa + p->arena_ada + (size_t)i * 6 * D * D, 6 * D * D);
  }
  add(w->final_ada_w, wa + p->arena_ada + (size_t)d.depth * 6 * D * D, 2 * D * D);
  V4H_REQUIRE(w->final_w && w->x_w && w->t0_w && w->t2_w && w->c2_w, "prepare_weights: null embedding / output weight");
  add(w->final_w, wa + p->arena_final, (size_t)d.out_dim * D);
  add(w->x_w, wa + p->arena_x, D * (size_t)d.patch_dim);
  add(w->t0_w, wa + p->arena_t0, D * (size_t)d.freq_dim);
  add(w->t2_w, wa + p->arena_t2, D * D);
  add(w->c2_w, wa + p->arena_c2, D * D);
  if (memcmp(jobs.data(), p->jobs_host, sizeof(CastJob) * jobs.size()) != 0) {
    // parameter storage moved: wait for earlier uses of the staging table, then refresh it -- impossible inside
    // a stream capture (the synchronisation would invalidate it): fail with a message instead
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
      return fail(V4H_ERR_INVALID, "prepare_weights: parameter storage moved during CUDA-graph capture; run one "
                                   "eager forward after (re)allocating parameters and before capturing");
    V4H_CUDA(cudaStreamSynchronize(s));
    memcpy(p->jobs_host, jobs.data(), sizeof(CastJob) * jobs.size());
    V4H_CUDA(cudaMemcpyAsync(p->jobs_dev, p->jobs_host, sizeof(CastJob) * jobs.size(), cudaMemcpyHostToDevice, s));
  }
  V4H_TRY(cast_many_f32_to_bf16(p->jobs_dev, (int)jobs.size(), max_n, s));
  float* bias = reinterpret_cast<float*>(reinterpret_cast<char*>(arena) + p->arena_ada_bias_bytes);
  for (int i = 0; i < d.depth; ++i)
    V4H_CUDA(cudaMemcpyAsync(bias + (size_t)i * 6 * D, w->blocks[i].ada_b, 6 * D * 4, cudaMemcpyDeviceToDevice, s));
  V4H_CUDA(cudaMemcpyAsync(bias + (size_t)d.depth * 6 * D, w->final_ada_b, 2 * D * 4, cudaMemcpyDeviceToDevice, s));
  return V4H_OK;
}

int plan_forward(Plan* p, const v4h_vit_params* w, const void* arena, const float* x, const float* t,
                 const float* c, float* out, int64_t B, bool shared_t, bool train, void* workspace,
                 size_t workspace_bytes, cudaStream_t s) {
  V4H_REQUIRE(p && w && x && t && c && out && workspace, "vit_forward: null argument");
  V4H_REQUIRE(B > 0 && B * p->d.tokens < (1ll << 31) / 4096, "vit_forward: batch out of range");
  V4H_REQUIRE(!p->bf16 || arena, "vit_forward: bf16 precision needs the weight arena");
  V4H_REQUIRE(p->d.x_map_dim == 0 || (w->xm_w && w->xm_b), "vit_forward: x mapper weights are null");
  V4H_REQUIRE(p->d.c_map_dim == 0 || (w->cm_w && w->cm_b), "vit_forward: c mapper weights are null");
  Sub subs[2];
  int nsub = 0;
  V4H_TRY(make_subs(*p, reinterpret_cast<char*>(workspace), workspace_bytes, B, train, shared_t, x, t, c, out, nullptr, s,
                    subs, &nsub));
  if (p->bf16) return forward_all<bf16>(*p, *w, (const char*)arena, subs, nsub, shared_t, train, s);
  return forward_all<float>(*p, *w, (const char*)arena, subs, nsub, shared_t, train, s);
}

int plan_backward(Plan* p, const v4h_vit_params* w, const void* arena, const v4h_vit_params* grads,
                  const float* x, const float* c, const float* dout, int64_t B, int stage_begin, int stage_end,
                  void* workspace, size_t workspace_bytes, cudaStream_t s) {
  V4H_REQUIRE(p && w && grads && workspace, "vit_backward: null argument");
  V4H_REQUIRE(stage_begin <= p->d.depth + 1 && stage_end >= 0 && stage_begin >= stage_end,
              "vit_backward: bad stage range [%d, %d]", stage_begin, stage_end);
  V4H_REQUIRE(stage_begin != p->d.depth + 1 || dout, "vit_backward: dout is null");
  V4H_REQUIRE(stage_end != 0 || (x && c), "vit_backward: stage 0 needs the forward inputs x and c");
  Sub subs[2];
  int nsub = 0;
  V4H_TRY(make_subs(*p, reinterpret_cast<char*>(workspace), workspace_bytes, B, true, false, x, nullptr, c, nullptr, dout,
                    s, subs, &nsub));
  if (p->bf16) return backward_all<bf16>(*p, *w, (const char*)arena, *grads, subs, nsub, stage_begin, stage_end, s);
  return backward_all<float>(*p, *w, (const char*)arena, *grads, subs, nsub, stage_begin, stage_end, s);
}

}  // namespace v4h
