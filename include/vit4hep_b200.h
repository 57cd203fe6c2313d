/*
 * vit4hep_b200 - C ABI of the B200-native CFM-ViT hot path.
 *
 * The reference (luigifvr/vit4hep) is 100 % Python and has no FFI of its own; the
 * boundary it exposes for this path is the nn.Module contract of
 *   - nn.vit.ViT.forward(x, t, c)                      reference nn/vit.py:185-206
 *   - CaloChallengeCFM.to_patches / from_patches       reference experiments/calochallenge/calochallenge_cfm/model.py:40-60
 *   - CFM._batch_loss                                  reference models/base_model.py:203-218
 *   - CaloChallengeCFM.sample_batch (torchdiffeq rk4)  reference experiments/calochallenge/calochallenge_cfm/model.py:68-94
 * Each entry point below names the reference interface it stands in for.  The Python
 * host in vit4hep_b200/ binds these with ctypes and mirrors the
 * reference's module/class names above them.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all memory is owned by the caller (PyTorch); nothing here allocates device memory
 *     except the small per-plan tables created by v4h_plan_create/v4h_geometry_create;
 *   - every launch goes to the cudaStream_t passed in; calls are asynchronous;
 *   - return value 0 = success, otherwise a V4H_ERR_* code; v4h_last_error() gives the
 *     message for the calling thread;
 *   - there is no CPU fallback: a device that is not sm_100 fails with V4H_ERR_ARCH.
 */
#ifndef VIT4HEP_B200_H
#define VIT4HEP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* v4h_stream_t; /* cudaStream_t */

enum {
  V4H_OK = 0,
  V4H_ERR_INVALID = 1,     /* bad shape / null pointer / misalignment */
  V4H_ERR_CUDA = 2,        /* a CUDA runtime call failed */
  V4H_ERR_ARCH = 3,        /* device is not sm_100 (B200) */
  V4H_ERR_UNSUPPORTED = 4  /* a knob the reference never enables (dropout, qk_norm, causal mask ...) */
};

enum { V4H_FP32 = 0, V4H_BF16 = 1 }; /* arithmetic of the GEMM operands / saved activations */

#define V4H_MAX_DEPTH 32
#define V4H_MAX_SEGMENTS 16

const char* v4h_last_error(void);
int v4h_version(void);
/* 0 when device `dev` is sm_100 and usable, else V4H_ERR_ARCH / V4H_ERR_CUDA */
int v4h_check_device(int dev);

/* ------------------------------------------------------------------------------------
 * Geometry: 3D patchify / unpatchify over (layer, angular, radial) voxel grids.
 * Replaces einops.rearrange in to_patches/from_patches (reference
 * calochallenge_cfm/model.py:40-60 regular grid; :146-173, experiments/calogan/model.py:60-87,
 * experiments/calohadronic/model.py:59-86 segmented).  Bit-exact copies.
 * ------------------------------------------------------------------------------------ */
typedef struct v4h_geometry v4h_geometry;

/* shapes/patches: n_segments x 3 ints (L, A, R) / (P1, P2, P3); flat_input: the wrapper's
 * input is (B, C, sum V) split at the segment edges instead of (B, C, L, A, R). */
int v4h_geometry_create(const int32_t* shapes_host, const int32_t* patches_host, int32_t n_segments,
                        int32_t in_channels, int32_t flat_input, v4h_geometry** out);
void v4h_geometry_destroy(v4h_geometry* g);
int32_t v4h_geometry_tokens(const v4h_geometry* g);
int32_t v4h_geometry_patch_dim(const v4h_geometry* g);
int32_t v4h_geometry_voxels(const v4h_geometry* g); /* per sample, channels included */
/* copies the host-side gather table (tokens.flat[j] = x.flat[table[j]]) for inspection */
int v4h_geometry_table_host(const v4h_geometry* g, int32_t* table_out_host, int32_t n);

/* to_patches: (B, C, *grid) fp32 -> (B, T, P) fp32 */
int v4h_to_patches(const v4h_geometry* g, const float* x, float* tokens, int64_t batch, v4h_stream_t s);
/* from_patches: (B, T, P) fp32 -> (B, C, *grid) fp32 */
int v4h_from_patches(const v4h_geometry* g, const float* tokens, float* x, int64_t batch, v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * The ViT velocity network (reference nn/vit.py ViT / DiTBlock / FinalLayer /
 * TimestepEmbedder / Attention).  Parameters stay in the caller's nn.Module with the
 * reference's state_dict names; these structs only carry their device addresses.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  int32_t hidden_dim;   /* D   */
  int32_t depth;
  int32_t num_heads;    /* H, D % H == 0 */
  int32_t mlp_hidden;   /* int(D * mlp_ratio) */
  int32_t patch_dim;    /* P   */
  int32_t out_dim;      /* P * out_channels */
  int32_t cond_dim;     /* K   */
  int32_t tokens;       /* T   */
  int32_t freq_dim;     /* TimestepEmbedder.frequency_embedding_size (256) */
  int32_t learn_pos_embed; /* 1: pos_embed_freqs + pos_{z,y,x}; 0: fixed (T, D) table */
  int32_t precision;    /* V4H_FP32 | V4H_BF16 */
  /* Finetuning structures (reference experiments/calochallenge/calochallenge_cfm/experiment_finetuning.py:79-118):
   * x_embedder / c_embedder wrapped as Sequential(mapper Linear, SiLU, old embedder).  > 0: width of the
   * network input x / c that the mapper Linear takes to patch_dim / cond_dim; 0: no mapper. */
  int32_t x_map_dim;
  int32_t c_map_dim;
} v4h_vit_dims;

typedef struct {
  float *qkv_w, *qkv_b;     /* blocks.i.attn.qkv      (3D, D), (3D) */
  float *proj_w, *proj_b;   /* blocks.i.attn.proj     (D, D), (D) */
  float *fc1_w, *fc1_b;     /* blocks.i.mlp.fc1       (Hm, D), (Hm) */
  float *fc2_w, *fc2_b;     /* blocks.i.mlp.fc2       (D, Hm), (D) */
  float *ada_w, *ada_b;     /* blocks.i.adaLN_modulation.1  (6D, D), (6D) */
} v4h_block_params;

/* Used both for the weights (read-only) and for the gradients (written). */
typedef struct {
  float *pos_embed_freqs;           /* (D/6)   - null when learn_pos_embed == 0 */
  float *pos_z, *pos_y, *pos_x;     /* (T) buffers, no gradient */
  float *pos_embed;                 /* (T, D) fixed table when learn_pos_embed == 0 */
  float *x_w, *x_b;                 /* x_embedder (D, P), (D) */
  float *c0_w, *c0_b, *c2_w, *c2_b; /* c_embedder.0 (D, K), c_embedder.2 (D, D) */
  float *t0_w, *t0_b, *t2_w, *t2_b; /* t_embedder.mlp.0 (D, 256), .2 (D, D) */
  float *final_w, *final_b;         /* final_layer.linear (out_dim, D) */
  float *final_ada_w, *final_ada_b; /* final_layer.adaLN_modulation.1 (2D, D) */
  float *xm_w, *xm_b;               /* x_embedder.0 of a mapped embedder (P, x_map_dim), (P); else null */
  float *cm_w, *cm_b;               /* c_embedder.0 of a mapped embedder (K, c_map_dim), (K); else null */
  v4h_block_params blocks[V4H_MAX_DEPTH];
} v4h_vit_params;

typedef struct v4h_plan v4h_plan;

int v4h_plan_create(const v4h_vit_dims* dims, v4h_plan** out);
void v4h_plan_destroy(v4h_plan* p);

/* bytes of caller-provided scratch for one forward (save_for_backward = 0) or one
 * forward + backward (= 1) at this batch size */
size_t v4h_vit_workspace_bytes(const v4h_plan* p, int64_t batch, int32_t save_for_backward);
/* bytes of the bf16 weight arena (0 in fp32 precision) */
size_t v4h_vit_weight_arena_bytes(const v4h_plan* p);
/* (re)build the bf16 operand copies of the GEMM weights; call whenever a parameter changed */
int v4h_vit_prepare_weights(v4h_plan* p, const v4h_vit_params* w, void* arena, v4h_stream_t s);

/* ViT.forward: x (B, T, P), t (B, 1), c (B, K) -> out (B, T, out_dim), all fp32.
 * shared_t != 0: every sample has the same t (ODE sampling, reference
 * calochallenge_cfm/model.py:81-83) - t then points to ONE float. */
int v4h_vit_forward(v4h_plan* p, const v4h_vit_params* w, const void* arena,
                    const float* x, const float* t, const float* c, float* out,
                    int64_t batch, int32_t shared_t, int32_t save_for_backward,
                    void* workspace, size_t workspace_bytes, v4h_stream_t s);

/* Backward of the forward that last wrote `workspace` with save_for_backward = 1.
 * dout (B, T, out_dim) fp32.  Gradients of every parameter are ADDED into `grads`
 * (same layout as the weights; caller zeroes them).  block_hi/block_lo select a range of
 * the backward chain so the caller can overlap gradient all-reduce buckets with the rest:
 * stages run from `stage_begin` down to `stage_end` inclusive where stage depth+1 is the
 * final layer, stages depth..1 are blocks depth-1..0 and stage 0 is embeddings +
 * conditioning (the only stage that reads the forward inputs x (B, T, P) and c (B, K)).
 * Pass (depth+1, 0) for the whole backward. */
int v4h_vit_backward(v4h_plan* p, const v4h_vit_params* w, const void* arena,
                     const v4h_vit_params* grads, const float* x, const float* c,
                     const float* dout, int64_t batch,
                     int32_t stage_begin, int32_t stage_end,
                     void* workspace, size_t workspace_bytes, v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * CFM training step pieces (reference models/base_model.py:203-218, models/trajectories.py:5-8)
 * ------------------------------------------------------------------------------------ */
/* x_t = (1-t) x0 + t x1 and x_t_dot = x1 - x0, both emitted directly in token layout
 * (B, T, P) through the geometry's gather table; x0/x1 are voxel-layout fp32, t is (B). */
int v4h_cfm_prepare(const v4h_geometry* g, const float* x1, const float* x0, const float* t,
                    float* xt_tokens, float* target_tokens, int64_t batch, v4h_stream_t s);
/* loss = mean((v - target)^2) over n elements -> loss_out[0] (fp32, overwritten);
 * dv = 2 (v - target) / n * grad_scale (pass dv = NULL to skip) */
int v4h_cfm_loss(const float* v, const float* target, int64_t n, float grad_scale,
                 float* loss_out, float* dv, v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * Fused tail of the training step (reference experiments/base_experiment.py:573-597:
 * clip_grad_norm_(parameters, max_grad_norm) then AdamW.step()), one multi-tensor pass that also
 * refreshes the bf16 operand copy of the GEMM weights.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  float* p;        /* parameter (updated in place) */
  const float* g;  /* gradient */
  float* m;        /* exp_avg */
  float* v;        /* exp_avg_sq */
  void* bf16_dst;  /* bf16 copy of the updated parameter inside the weight arena, or NULL */
  float* f32_dst;  /* fp32 copy inside the arena (the concatenated adaLN biases), or NULL */
  float* ema;      /* exponential-moving-average shadow of the parameter (torch_ema), or NULL */
  int64_t n;
} v4h_adamw_job;
/* out[0] = sum of squares of n fp32 values (the squared global gradient norm when `flat` is the flat
 * gradient buffer) */
int v4h_grad_norm_sq(const float* flat, int64_t n, float* out, v4h_stream_t s);
/* jobs: DEVICE array of njobs entries; max_n = largest jobs[i].n; norm_sq: device scalar from
 * v4h_grad_norm_sq (NULL = no clipping); step = 1-based step count (bias correction).  step_dev / lr_dev
 * (optional device scalars) override step / lr so that a captured CUDA graph of the training step stays
 * correct across replays; v4h_counter_increment advances the device step counter in stream order. */
/* ema_decay > 0: the same pass also updates jobs[i].ema like torch_ema.ExponentialMovingAverage.update()
 * (reference experiments/base_experiment.py:127-134, :594): shadow -= (1 - d) (shadow - p) with
 * d = min(ema_decay, (1 + n) / (10 + n)), n = ema_updates (1-based count including this update) or
 * *ema_updates_dev when non-NULL. */
int v4h_adamw_step(const v4h_adamw_job* jobs, int32_t njobs, int64_t max_n, const float* norm_sq, float max_norm,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   const int32_t* step_dev, const float* lr_dev, float ema_decay, int32_t ema_updates,
                   const int32_t* ema_updates_dev, v4h_stream_t s);
/* ExponentialMovingAverage.update() alone (jobs[i].p, jobs[i].ema, jobs[i].n are read) */
int v4h_ema_update(const v4h_adamw_job* jobs, int32_t njobs, int64_t max_n, float decay, int32_t num_updates,
                   const int32_t* num_updates_dev, v4h_stream_t s);
int v4h_counter_increment(int32_t* counter, v4h_stream_t s);
/* byte offset inside the weight arena of the bf16 copy of a parameter, by its v4h_vit_params field
 * ("final_w", "x_w", "t0_w", "t2_w", "c2_w", "final_ada_w", "blocks.<i>.{qkv_w,proj_w,fc1_w,fc2_w,ada_w}"), or of the fp32 copy of an adaLN bias
 * ("blocks.<i>.ada_b", "final_ada_b"); -1 when the parameter has no copy */
int64_t v4h_vit_arena_offset(const v4h_plan* p, const char* field);

/* ------------------------------------------------------------------------------------
 * ODE sampling, torchdiffeq fixed-grid 'rk4' = 3/8 rule (reference
 * calochallenge_cfm/model.py:85-92).  out = y + a0*k0 + a1*k1 + a2*k2 + a3*k3 (null k = skipped):
 * one fused kernel per RK stage/combination.
 * ------------------------------------------------------------------------------------ */
int v4h_axpy4(float* out, const float* y, const float* k0, float a0, const float* k1, float a1,
              const float* k2, float a2, const float* k3, float a3, int64_t n, v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * Energy-ratio velocity network, forward only (reference nn/cfm/transformer_cfm.py:12-119 ParallelTransformer with
 * embeds = True around torch.nn.Transformer; sampled before every shape-sampling job, reference
 * experiments/calochallenge/experiment.py:225-247).  Parameters stay in the caller's nn.Module with the
 * reference's state_dict names; the structs carry their device addresses (all fp32).
 * ------------------------------------------------------------------------------------ */
#define V4H_ENERGY_MAX_LAYERS 16
typedef struct {
  int32_t dims_in;        /* tokens of x (45 energy ratios) */
  int32_t dims_c;         /* tokens of the condition (1; 3 for LEMURS), <= 16 */
  int32_t dim_embedding;  /* x embedding width; d_model = encode_t_dim + dim_embedding */
  int32_t encode_t_dim;   /* time embedding width (64) */
  int32_t nhead;
  int32_t n_enc, n_dec;   /* num_encoder_layers, num_decoder_layers */
  int32_t dim_feedforward;
  int32_t precision;      /* V4H_FP32 | V4H_BF16 */
} v4h_energy_dims;
typedef struct {  /* transformer.encoder.layers.i */
  float *in_w, *in_b, *out_w, *out_b;      /* self_attn.in_proj_{weight,bias} (3E, E), self_attn.out_proj (E, E) */
  float *l1_w, *l1_b, *l2_w, *l2_b;        /* linear1 (F, E), linear2 (E, F) */
  float *n1_w, *n1_b, *n2_w, *n2_b;        /* norm1, norm2 (E) */
} v4h_energy_enc_layer;
typedef struct {  /* transformer.decoder.layers.i */
  float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b;  /* self_attn */
  float *ca_in_w, *ca_in_b, *ca_out_w, *ca_out_b;  /* multihead_attn (cross attention over the encoded condition) */
  float *l1_w, *l1_b, *l2_w, *l2_b;
  float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;
} v4h_energy_dec_layer;
typedef struct {
  float *gfp_w;                       /* time_embed.0.W (encode_t_dim / 2) */
  float *time_w, *time_b;             /* time_embed.1 (Dt, Dt) */
  float *x_embed_w, *x_embed_b;       /* x_embed (De, 1) */
  float *c_embed_w, *c_embed_b;       /* c_embed (E, 1) */
  float *pos_x, *pos_c;               /* pos_embed_x.weight (dims_in, De), pos_embed_c.weight (dims_c, E) */
  float *enc_norm_w, *enc_norm_b;     /* transformer.encoder.norm */
  float *dec_norm_w, *dec_norm_b;     /* transformer.decoder.norm */
  float *head0_w, *head0_b;           /* layers.0 (= layer) (F, Dt + E) */
  float *head2_w, *head2_b;           /* layers.2 (1, F) */
  v4h_energy_enc_layer enc[V4H_ENERGY_MAX_LAYERS];
  v4h_energy_dec_layer dec[V4H_ENERGY_MAX_LAYERS];
} v4h_energy_params;
typedef struct v4h_energy_plan v4h_energy_plan;

int v4h_energy_plan_create(const v4h_energy_dims* dims, v4h_energy_plan** out);
void v4h_energy_plan_destroy(v4h_energy_plan* p);
size_t v4h_energy_workspace_bytes(const v4h_energy_plan* p, int64_t batch);
size_t v4h_energy_weight_arena_bytes(const v4h_energy_plan* p); /* bf16 operand copies; 0 in fp32 precision */
int v4h_energy_prepare_weights(v4h_energy_plan* p, const v4h_energy_params* w, void* arena, v4h_stream_t s);
/* condition side, once per batch: c (B, dims_c) -> encoder memory and the cross-attention K / V of every decoder
 * layer, kept in `workspace` for the v4h_energy_forward calls that follow (same workspace, same batch) */
int v4h_energy_encode(v4h_energy_plan* p, const v4h_energy_params* w, const void* arena, const float* c, int64_t batch,
                      void* workspace, size_t workspace_bytes, v4h_stream_t s);
/* one velocity evaluation: x (B, dims_in), t (B) [one value when shared_t] -> out (B, dims_in) */
int v4h_energy_forward(v4h_energy_plan* p, const v4h_energy_params* w, const void* arena, const float* x, const float* t,
                       int32_t shared_t, float* out, int64_t batch, void* workspace, size_t workspace_bytes,
                       v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * Post-processing of sampled showers (reference experiments/calochallenge/experiment.py:286-289: the transforms of
 * configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28 applied in reverse; classes in
 * experiments/calochallenge/transforms.py): Reshape, AddFeaturesToCond, ScaleEnergy(e_min, e_max), LogEnergy(alpha),
 * GlobalStandardizeFromFile(mean, std), ExclusiveLogitTransform(delta, rescale=True), CutValues(cut),
 * ScaleTotalEnergy(factor), NormalizeByElayer(eps, norm_cut), fused: every voxel is read once and written once.
 * x (n, voxels): sampled showers; cond (n, n_layers + 1): the u features then the scaled log incident energy (the
 * conditions the shape network was sampled with); layer_bounds (n_layers + 1) DEVICE int32 voxel offsets of the
 * calorimeter layers (XMLHandler.GetBinEdges in the reference); max_layer_voxels: the largest layer (<= 1024 selects
 * the kernels that keep a layer in registers and touch every voxel once; 0 = unknown, generic two-sweep kernel).
 * out (n, voxels): energy per voxel; e_out (n): incident energies.
 * ------------------------------------------------------------------------------------ */
int v4h_postprocess_showers(const float* x, const float* cond, int64_t n, int32_t voxels, int32_t n_layers,
                            const int32_t* layer_bounds, int32_t max_layer_voxels, float mean, float std, float delta, float cut, float factor,
                            float e_min, float e_max, float alpha, float eps, float norm_cut, float* out,
                            float* e_out, v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * Pre-processing of raw showers = the data feed (reference experiments/calochallenge/datasets.py:44-47 applies the
 * transforms of configs/calochallenge/cfm/calochallenge_ds2.yaml:15-28 forwards, on the CPU, when the dataset is
 * built; classes in experiments/calochallenge/transforms.py): NormalizeByElayer(eps), ScaleTotalEnergy(factor),
 * CutValues (identity forwards), ExclusiveLogitTransform(delta, rescale=True), GlobalStandardizeFromFile,
 * LogEnergy(alpha), ScaleEnergy(e_min, e_max), AddFeaturesToCond, Reshape.
 * showers (n, voxels): raw energies per voxel; e_inc (n): incident energies; layer_bounds as above.
 * mean_std: DEVICE float[2].  compute_stats == 0: it holds the (mean, std) that GlobalStandardizeFromFile loaded and
 * everything happens in one kernel.  compute_stats != 0: the `written == False` branch (transforms.py:55-63) — mean
 * and unbiased std of the non-saturated features of THIS call are computed on the device (stats: DEVICE double[3]
 * scratch), written to mean_std, and applied by a second kernel.
 * x (n, voxels): network-space showers; cond (n, n_layers + 1): the u features then the scaled log incident energy.
 * ------------------------------------------------------------------------------------ */
int v4h_preprocess_showers(const float* showers, const float* e_inc, int64_t n, int32_t voxels, int32_t n_layers,
                           const int32_t* layer_bounds, int32_t max_layer_voxels, float eps, float factor, float delta, float alpha, float e_min,
                           float e_max, float* mean_std, int32_t compute_stats, double* stats, float* x, float* cond,
                           v4h_stream_t s);

/* ------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): how many kernels the library launched, and per-kernel-class device
 * time from CUDA events recorded on the launching stream around each launch.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  char name[32];     /* kernel class, e.g. "gemm.fc1", "attn.fwd", "ln.fwd" */
  int64_t launches;  /* bracketed regions (one or a few kernel launches each) */
  double ms;         /* summed device time */
  double flops;      /* summed algorithmic flops (2 M N K per GEMM, 4 H T^2 dh per attention forward) */
  double bytes;      /* summed algorithmic bytes (operands read + results written once) */
} v4h_profile_entry;

int64_t v4h_launch_count(void);
/* start booking; every later launch is bracketed by two events until v4h_profile_end */
int v4h_profile_begin(void);
/* synchronises the device, aggregates by name into out[0..*n) (at most max entries), stops booking */
int v4h_profile_end(v4h_profile_entry* out, int32_t max, int32_t* n);

/* ------------------------------------------------------------------------------------
 * Kernel-level test hooks (used by tests/ to localise parity failures; not needed by a host).
 * ------------------------------------------------------------------------------------ */
/* C (m, n) fp32 = op(A) op(B) with row-major inputs.  layout: 0 = NT (A (m,k), B (n,k)),
 * 1 = NN (A (m,k), B (k,n)), 2 = TN (A (k,m), B (k,n)).  engine: 0 = SIMT fp32 inputs,
 * 1 = tcgen05 bf16 inputs (A, B are bf16 then). */
int v4h_test_gemm(int32_t engine, int32_t layout, const void* A, const void* B, float* C,
                  int32_t m, int32_t n, int32_t k, v4h_stream_t s);
/* One tcgen05 GEMM with the epilogue of a model call site, for kernel benchmarking (scripts/gemm_bench.py).
 * kind: 0 fc1-like (NT, bias + GELU, out and out2), 1 qkv-like (NT, bias), 2 proj/fc2-like (NT, gate +
 * residual), 3 dgrad through GELU (NN, aux), 4 plain dgrad (NN), 5 wgrad (TN, fp32 atomics into out),
 * 6 final-layer-like (NT, bias, fp32 out).
 * counters: optional device array of 16 int64 cycle counters summed over CTAs: [0,1] producer wait /
 * issue, [2,3,4] MMA wait accumulator / wait operands / issue, [5..11] epilogue wait accumulator, loads +
 * TMEM, wait input box, math + staging, barrier, copy-out, tail. */
int v4h_debug_gemm(int32_t kind, int32_t m, int32_t n, int32_t k, int32_t rows_per_sample, const void* A,
                   const void* B, const float* bias, void* out, void* out2, const float* res_in,
                   float* res_out, const float* gate, const void* aux, int64_t* counters, v4h_stream_t s);

/* The gated-residual GEMM with the following LayerNorm + modulation in its epilogue (csrc/gemm_umma.cu
 * gemm_gate_res_ln; reference nn/vit.py:331-332), stand-alone for kernel tests and benchmarks:
 *   y = A W^T + bias (bf16, optional);  res_out = res_in + gate[b] * y;  stats[row] = (mean, rstd) (optional);
 *   ln_out[row, :n] = LN(res_out[row]) * (1 + scale[b]) + shift[b]   (bf16, row pitch ld_ln, eps 1e-6)
 * A (m, k), W (n, k) bf16; gate / shift / scale (ceil(m / rows_per_sample), n) fp32; res_in / res_out (m, n) fp32.
 * counters: optional device array of 8 int64 cycle counters of the first epilogue thread, summed over CTAs: wait
 * accumulator; pass 1 TMEM loads, wait residual box, math + TMEM store, store / barrier; statistics; pass 2 math,
 * store / barrier. */
int v4h_debug_gemm_ln(int32_t m, int32_t n, int32_t k, int32_t rows_per_sample, const void* A, const void* W,
                      const float* bias, void* y, const float* res_in, float* res_out, const float* gate,
                      const float* shift, const float* scale, void* ln_out, int32_t ld_ln, float* stats, int64_t* counters,
                      v4h_stream_t s);

/* Measurement hook: feed rate of a TMA + mbarrier ring without a consumer.  `ctas` persistent CTAs each pull
 * `iters` stages of `boxes` [box_rows x 64] bf16 boxes (box_rows 64 / 128 / 256: 8 / 16 / 32 KB, 128-byte
 * swizzle) through a ring of `stages` stages out of the (rows x cols) bf16 matrix `buf` (rows % box_rows == 0,
 * cols % 64 == 0; its size sets the working set); `producers` (1..4) warps share the issue of a stage's
 * boxes; cycles[cta] receives the consumer's cycle count.  No reference counterpart. */
int v4h_debug_tma_probe(const void* buf, int32_t rows, int32_t cols, int32_t stages, int32_t boxes, int32_t box_rows,
                        int32_t producers, int32_t iters, int32_t ctas, int64_t* cycles, v4h_stream_t s);
/* device array of 10 int64 cycle counters booked by thread 0 of every tcgen05 attention-forward CTA
 * (prologue, issue loads, wait loads, publish, S MMA, softmax, publish, PV MMA, output, teardown); NULL = off */
int v4h_debug_attention_counters(int64_t* counters);
/* qkv (B, T, 3, H, dh) -> o (B, T, H, dh), lse (B, H, T); precision picks fp32 / bf16 buffers */
int v4h_test_attention_fwd(int32_t precision, int32_t engine, const void* qkv, void* o, float* lse,
                           int32_t batch, int32_t tokens, int32_t heads, int32_t head_dim, v4h_stream_t s);
int v4h_test_attention_bwd(int32_t precision, int32_t engine, const void* qkv, const void* o,
                           const float* lse, const void* d_o, void* dqkv, int32_t batch, int32_t tokens,
                           int32_t heads, int32_t head_dim, v4h_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* VIT4HEP_B200_H */
