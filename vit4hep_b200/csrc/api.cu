// extern "C" surface of libvit4hep_b200.so (declared in include/vit4hep_b200.h).
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace v4h {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

// ---------------------------------------------------------------- launch counter / profiler
static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("V4H_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
bool launch_sync_enabled() {
  static const bool on = [] { const char* e = getenv("V4H_LAUNCH_SYNC"); return e && e[0] == '1'; }();
  return on;
}

namespace {
struct ProfRecord { std::string tag; double flops, bytes; cudaEvent_t e0, e1; };
struct Profiler {
  std::mutex mu;
  std::atomic<bool> on{false};
  std::vector<ProfRecord> records;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
Profiler& profiler() { static Profiler p; return p; }
}  // namespace

bool profiling_enabled() { return profiler().on.load(std::memory_order_relaxed); }
// inside a stream capture the scope's events become EXTERNAL event-record nodes: they are recorded again by
// every replay of the graph and can be timed afterwards (scripts/graph_profile.py)
static void record_scope_event(cudaEvent_t e, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive)
    cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
  else
    cudaEventRecord(e, s);
}
void profile_open(const char* tag, double flops, double bytes, cudaStream_t s, int* slot) {
  Profiler& p = profiler();
  std::lock_guard<std::mutex> lock(p.mu);
  ProfRecord r{tag, flops, bytes, p.get_event(), p.get_event()};
  record_scope_event(r.e0, s);
  p.records.push_back(r);
  *slot = (int)p.records.size() - 1;
}
void profile_close(int slot, cudaStream_t s) {
  Profiler& p = profiler();
  std::lock_guard<std::mutex> lock(p.mu);
  if (slot >= 0 && slot < (int)p.records.size()) record_scope_event(p.records[slot].e1, s);
}

struct EnergyPlan;
int energy_plan_create(const v4h_energy_dims* dims, EnergyPlan** out);
void energy_plan_destroy(EnergyPlan* p);
size_t energy_workspace_bytes(const EnergyPlan* p, int64_t B);
size_t energy_arena_bytes(const EnergyPlan* p);
int energy_prepare_weights(EnergyPlan* p, const v4h_energy_params* w, void* arena, cudaStream_t s);
int energy_encode(EnergyPlan* p, const v4h_energy_params* w, const void* arena, const float* c, int64_t B, void* workspace,
                  size_t workspace_bytes, cudaStream_t s);
int energy_forward(EnergyPlan* p, const v4h_energy_params* w, const void* arena, const float* x, const float* t, bool shared_t,
                   float* out, int64_t B, void* workspace, size_t workspace_bytes, cudaStream_t s);

struct Plan;
int plan_create(const v4h_vit_dims* dims, Plan** out);
void plan_destroy(Plan* p);
size_t plan_workspace_bytes(const Plan* p, int64_t B, bool train);
size_t plan_arena_bytes(const Plan* p);
int plan_prepare_weights(Plan* p, const v4h_vit_params* w, void* arena, cudaStream_t s);
int plan_forward(Plan* p, const v4h_vit_params* w, const void* arena, const float* x, const float* t,
                 const float* c, float* out, int64_t B, bool shared_t, bool train, void* workspace,
                 size_t workspace_bytes, cudaStream_t s);
int64_t plan_arena_offset(const Plan* p, const char* field);
int plan_backward(Plan* p, const v4h_vit_params* w, const void* arena, const v4h_vit_params* grads,
                  const float* x, const float* c, const float* dout, int64_t B, int stage_begin, int stage_end,
                  void* workspace, size_t workspace_bytes, cudaStream_t s);

// ---------------------------------------------------------------- geometry
struct Geometry {
  int tokens = 0, patch_dim = 0, voxels = 0;  // voxels: per sample, channels included
  std::vector<int32_t> table, inverse, chunks;  // host copies
  int max_chunk = 0;
  int32_t *table_dev = nullptr, *inverse_dev = nullptr, *chunks_dev = nullptr;
  int num_chunks = 0;  // 0 -> plain gather kernel
  std::mutex upload_mutex;
  int device = -1;  // device that holds the tables (uploaded on first launch)

  // The host tables are built at creation (no GPU needed, so the index map can be inspected
  // anywhere); the device copies are made on the first launch, on the then-current device.
  int ensure_device() {
    std::lock_guard<std::mutex> lock(upload_mutex);
    int dev = 0;
    V4H_CUDA(cudaGetDevice(&dev));
    if (device == dev) return V4H_OK;
    if (device >= 0)
      return fail(V4H_ERR_INVALID, "geometry was uploaded to device %d but is used on device %d", device, dev);
    V4H_TRY(v4h_check_device(dev));
    const size_t nb = sizeof(int32_t) * voxels;
    V4H_CUDA(cudaMalloc(&table_dev, nb));
    V4H_CUDA(cudaMalloc(&inverse_dev, nb));
    V4H_CUDA(cudaMalloc(&chunks_dev, sizeof(int32_t) * std::max<size_t>(1, chunks.size())));
    V4H_CUDA(cudaMemcpy(table_dev, table.data(), nb, cudaMemcpyHostToDevice));
    V4H_CUDA(cudaMemcpy(inverse_dev, inverse.data(), nb, cudaMemcpyHostToDevice));
    if (!chunks.empty())
      V4H_CUDA(cudaMemcpy(chunks_dev, chunks.data(), sizeof(int32_t) * chunks.size(), cudaMemcpyHostToDevice));
    device = dev;
    return V4H_OK;
  }
};

}  // namespace v4h

using namespace v4h;

extern "C" {

const char* v4h_last_error(void) { return last_error_buffer(); }
int v4h_version(void) { return 100; }

int v4h_check_device(int dev) {
  cudaDeviceProp prop;
  V4H_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(V4H_ERR_ARCH, "device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", dev,
                prop.major, prop.minor);
  return V4H_OK;
}

// ------------------------------------------------------------------------------------ geometry
int v4h_geometry_create(const int32_t* shapes, const int32_t* patches, int32_t n_segments, int32_t in_channels,
                        int32_t flat_input, v4h_geometry** out) {
  V4H_REQUIRE(shapes && patches && out, "geometry_create: null argument");
  V4H_REQUIRE(n_segments >= 1 && n_segments <= V4H_MAX_SEGMENTS, "geometry_create: %d segments", n_segments);
  V4H_REQUIRE(in_channels >= 1, "geometry_create: in_channels must be >= 1");
  V4H_REQUIRE(flat_input || n_segments == 1, "geometry_create: a non-flat input has exactly one segment");
  const int C = in_channels;
  int64_t V = 0;
  int patch_vol = -1;
  for (int k = 0; k < n_segments; ++k) {
    int vol = 1;
    for (int a = 0; a < 3; ++a) {
      const int sdim = shapes[3 * k + a], pdim = patches[3 * k + a];
      V4H_REQUIRE(sdim > 0 && pdim > 0, "geometry_create: non-positive extent");
      // same check as the reference's asserts (calochallenge_cfm/model.py:33-36)
      V4H_REQUIRE(sdim % pdim == 0, "Input size (%d) should be divisible by patch size (%d) in axis %d.", sdim, pdim, a);
      vol *= pdim;
    }
    V4H_REQUIRE(patch_vol < 0 || patch_vol == vol, "geometry_create: patch volume differs between segments");
    patch_vol = vol;
    V += (int64_t)shapes[3 * k] * shapes[3 * k + 1] * shapes[3 * k + 2];
  }
  V4H_REQUIRE(V * C < (1ll << 30), "geometry_create: sample too large");
  Geometry* g = new Geometry();
  g->voxels = (int)(V * C);
  g->patch_dim = patch_vol * C;
  g->table.reserve(g->voxels);
  std::vector<int32_t> slab_bounds;  // token-order offsets where a slab (one patch-layer of a segment) starts
  int64_t seg_off = 0;
  for (int k = 0; k < n_segments; ++k) {
    const int L = shapes[3 * k], A = shapes[3 * k + 1], R = shapes[3 * k + 2];
    const int P1 = patches[3 * k], P2 = patches[3 * k + 1], P3 = patches[3 * k + 2];
    const int nl = L / P1, na = A / P2, nr = R / P3;
    g->tokens += nl * na * nr;
    for (int l = 0; l < nl; ++l) {
      slab_bounds.push_back((int32_t)g->table.size());
      for (int a = 0; a < na; ++a)
        for (int r = 0; r < nr; ++r)
          for (int p1 = 0; p1 < P1; ++p1)
            for (int p2 = 0; p2 < P2; ++p2)
              for (int p3 = 0; p3 < P3; ++p3)
                for (int c = 0; c < C; ++c) {
                  const int64_t vox = ((int64_t)(l * P1 + p1) * A + (a * P2 + p2)) * R + (r * P3 + p3);
                  g->table.push_back((int32_t)(c * V + seg_off + vox));
                }
    }
    seg_off += (int64_t)L * A * R;
  }
  slab_bounds.push_back((int32_t)g->table.size());
  g->inverse.assign(g->voxels, 0);
  for (int j = 0; j < g->voxels; ++j) g->inverse[g->table[j]] = j;
  // chunks = groups of whole slabs; valid only when every slab is a permutation of its own range (C == 1)
  bool block_diagonal = (C == 1);
  const int target = 2048, cap = 12288 - 64;
  if (block_diagonal) {
    g->chunks.push_back(0);
    int begin = 0;
    for (size_t i = 1; i < slab_bounds.size(); ++i) {
      const int here = slab_bounds[i];
      const bool last = i + 1 == slab_bounds.size();
      const int next = last ? here : slab_bounds[i + 1];
      if (last || next - begin > target) {
        g->chunks.push_back(here);
        g->max_chunk = std::max(g->max_chunk, here - begin);
        begin = here;
      }
    }
    if (g->max_chunk > cap) block_diagonal = false;
  }
  g->num_chunks = block_diagonal ? (int)g->chunks.size() - 1 : 0;
  *out = reinterpret_cast<v4h_geometry*>(g);
  return V4H_OK;
}

void v4h_geometry_destroy(v4h_geometry* gg) {
  Geometry* g = reinterpret_cast<Geometry*>(gg);
  if (!g) return;
  if (g->device >= 0) { cudaFree(g->table_dev); cudaFree(g->inverse_dev); cudaFree(g->chunks_dev); }
  delete g;
}
int32_t v4h_geometry_tokens(const v4h_geometry* g) { return reinterpret_cast<const Geometry*>(g)->tokens; }
int32_t v4h_geometry_patch_dim(const v4h_geometry* g) { return reinterpret_cast<const Geometry*>(g)->patch_dim; }
int32_t v4h_geometry_voxels(const v4h_geometry* g) { return reinterpret_cast<const Geometry*>(g)->voxels; }
int v4h_geometry_table_host(const v4h_geometry* gg, int32_t* out, int32_t n) {
  const Geometry* g = reinterpret_cast<const Geometry*>(gg);
  V4H_REQUIRE(g && out && n == g->voxels, "geometry_table_host: bad arguments");
  memcpy(out, g->table.data(), sizeof(int32_t) * n);
  return V4H_OK;
}

int v4h_to_patches(const v4h_geometry* gg, const float* x, float* tokens, int64_t batch, v4h_stream_t s) {
  Geometry* g = const_cast<Geometry*>(reinterpret_cast<const Geometry*>(gg));
  V4H_REQUIRE(g && x && tokens && batch >= 0, "to_patches: bad arguments");
  V4H_TRY(g->ensure_device());
  return patch_permute(x, tokens, g->table_dev, g->chunks_dev, g->num_chunks, g->max_chunk, batch, g->voxels,
                       (cudaStream_t)s);
}
int v4h_from_patches(const v4h_geometry* gg, const float* tokens, float* x, int64_t batch, v4h_stream_t s) {
  Geometry* g = const_cast<Geometry*>(reinterpret_cast<const Geometry*>(gg));
  V4H_REQUIRE(g && x && tokens && batch >= 0, "from_patches: bad arguments");
  V4H_TRY(g->ensure_device());
  return patch_permute(tokens, x, g->inverse_dev, g->chunks_dev, g->num_chunks, g->max_chunk, batch, g->voxels,
                       (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ network
int v4h_plan_create(const v4h_vit_dims* dims, v4h_plan** out) {
  return plan_create(dims, reinterpret_cast<Plan**>(out));
}
void v4h_plan_destroy(v4h_plan* p) { plan_destroy(reinterpret_cast<Plan*>(p)); }
size_t v4h_vit_workspace_bytes(const v4h_plan* p, int64_t batch, int32_t save_for_backward) {
  return plan_workspace_bytes(reinterpret_cast<const Plan*>(p), batch, save_for_backward != 0);
}
size_t v4h_vit_weight_arena_bytes(const v4h_plan* p) { return plan_arena_bytes(reinterpret_cast<const Plan*>(p)); }
int v4h_vit_prepare_weights(v4h_plan* p, const v4h_vit_params* w, void* arena, v4h_stream_t s) {
  return plan_prepare_weights(reinterpret_cast<Plan*>(p), w, arena, (cudaStream_t)s);
}
int v4h_vit_forward(v4h_plan* p, const v4h_vit_params* w, const void* arena, const float* x, const float* t,
                    const float* c, float* out, int64_t batch, int32_t shared_t, int32_t save_for_backward,
                    void* workspace, size_t workspace_bytes, v4h_stream_t s) {
  return plan_forward(reinterpret_cast<Plan*>(p), w, arena, x, t, c, out, batch, shared_t != 0,
                      save_for_backward != 0, workspace, workspace_bytes, (cudaStream_t)s);
}
int v4h_vit_backward(v4h_plan* p, const v4h_vit_params* w, const void* arena, const v4h_vit_params* grads,
                     const float* x, const float* c, const float* dout, int64_t batch, int32_t stage_begin,
                     int32_t stage_end, void* workspace, size_t workspace_bytes, v4h_stream_t s) {
  return plan_backward(reinterpret_cast<Plan*>(p), w, arena, grads, x, c, dout, batch, stage_begin, stage_end,
                       workspace, workspace_bytes, (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ CFM / ODE
int v4h_cfm_prepare(const v4h_geometry* gg, const float* x1, const float* x0, const float* t, float* xt_tokens,
                    float* target_tokens, int64_t batch, v4h_stream_t s) {
  Geometry* g = const_cast<Geometry*>(reinterpret_cast<const Geometry*>(gg));
  V4H_REQUIRE(g && x1 && x0 && t && xt_tokens && target_tokens && batch > 0 && batch <= 65535,
              "cfm_prepare: bad arguments");
  V4H_TRY(g->ensure_device());
  return cfm_prepare(x1, x0, t, g->table_dev, xt_tokens, target_tokens, batch, g->voxels, (cudaStream_t)s);
}
int v4h_cfm_loss(const float* v, const float* target, int64_t n, float grad_scale, float* loss_out, float* dv,
                 v4h_stream_t s) {
  V4H_REQUIRE(v && target && loss_out && n > 0, "cfm_loss: bad arguments");
  return cfm_loss(v, target, n, grad_scale, loss_out, dv, (cudaStream_t)s);
}
int v4h_axpy4(float* out, const float* y, const float* k0, float a0, const float* k1, float a1, const float* k2,
              float a2, const float* k3, float a3, int64_t n, v4h_stream_t s) {
  V4H_REQUIRE(out && y && n > 0, "axpy4: bad arguments");
  return axpy4(out, y, k0, a0, k1, a1, k2, a2, k3, a3, n, (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ energy network
int v4h_energy_plan_create(const v4h_energy_dims* dims, v4h_energy_plan** out) {
  return energy_plan_create(dims, reinterpret_cast<EnergyPlan**>(out));
}
void v4h_energy_plan_destroy(v4h_energy_plan* p) { energy_plan_destroy(reinterpret_cast<EnergyPlan*>(p)); }
size_t v4h_energy_workspace_bytes(const v4h_energy_plan* p, int64_t batch) {
  return energy_workspace_bytes(reinterpret_cast<const EnergyPlan*>(p), batch);
}
size_t v4h_energy_weight_arena_bytes(const v4h_energy_plan* p) { return energy_arena_bytes(reinterpret_cast<const EnergyPlan*>(p)); }
int v4h_energy_prepare_weights(v4h_energy_plan* p, const v4h_energy_params* w, void* arena, v4h_stream_t s) {
  return energy_prepare_weights(reinterpret_cast<EnergyPlan*>(p), w, arena, (cudaStream_t)s);
}
int v4h_energy_encode(v4h_energy_plan* p, const v4h_energy_params* w, const void* arena, const float* c, int64_t batch,
                      void* workspace, size_t workspace_bytes, v4h_stream_t s) {
  return energy_encode(reinterpret_cast<EnergyPlan*>(p), w, arena, c, batch, workspace, workspace_bytes, (cudaStream_t)s);
}
int v4h_energy_forward(v4h_energy_plan* p, const v4h_energy_params* w, const void* arena, const float* x, const float* t,
                       int32_t shared_t, float* out, int64_t batch, void* workspace, size_t workspace_bytes, v4h_stream_t s) {
  return energy_forward(reinterpret_cast<EnergyPlan*>(p), w, arena, x, t, shared_t != 0, out, batch, workspace,
                        workspace_bytes, (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ post-processing
int v4h_postprocess_showers(const float* x, const float* cond, int64_t n, int32_t voxels, int32_t n_layers,
                            const int32_t* layer_bounds, int32_t max_layer_voxels, float mean, float std, float delta, float cut, float factor,
                            float e_min, float e_max, float alpha, float eps, float norm_cut, float* out,
                            float* e_out, v4h_stream_t s) {
  V4H_REQUIRE(x && cond && layer_bounds && out && e_out && n > 0 && voxels > 0, "postprocess_showers: bad arguments");
  V4H_REQUIRE(std != 0.f && factor != 0.f && delta >= 0.f && delta < 0.5f, "postprocess_showers: bad transform parameters");
  return postprocess_showers(x, cond, n, voxels, n_layers, layer_bounds, max_layer_voxels > 0 ? max_layer_voxels : voxels, mean, std, delta, cut, factor, e_min, e_max,
                             alpha, eps, norm_cut, out, e_out, (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ pre-processing
int v4h_preprocess_showers(const float* showers, const float* e_inc, int64_t n, int32_t voxels, int32_t n_layers,
                           const int32_t* layer_bounds, int32_t max_layer_voxels, float eps, float factor, float delta, float alpha, float e_min,
                           float e_max, float* mean_std, int32_t compute_stats, double* stats, float* x, float* cond,
                           v4h_stream_t s) {
  V4H_REQUIRE(showers && e_inc && layer_bounds && mean_std && x && cond && n > 0 && voxels > 0,
              "preprocess_showers: bad arguments");
  V4H_REQUIRE(!compute_stats || stats, "preprocess_showers: compute_stats needs the double[3] scratch");
  V4H_REQUIRE(delta >= 0.f && delta < 0.5f && e_max != e_min, "preprocess_showers: bad transform parameters");
  return preprocess_showers(showers, e_inc, n, voxels, n_layers, layer_bounds, max_layer_voxels > 0 ? max_layer_voxels : voxels, eps, factor, delta, alpha, e_min, e_max,
                            mean_std, compute_stats, stats, x, cond, (cudaStream_t)s);
}

// ------------------------------------------------------------------------------------ optimizer
int v4h_grad_norm_sq(const float* flat, int64_t n, float* out, v4h_stream_t s) {
  V4H_REQUIRE(flat && out && n > 0, "grad_norm_sq: bad arguments");
  return grad_norm_sq(flat, n, out, (cudaStream_t)s);
}
int v4h_adamw_step(const v4h_adamw_job* jobs, int32_t njobs, int64_t max_n, const float* norm_sq, float max_norm,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   const int32_t* step_dev, const float* lr_dev, float ema_decay, int32_t ema_updates,
                   const int32_t* ema_updates_dev, v4h_stream_t s) {
  V4H_REQUIRE(jobs && njobs > 0 && njobs <= 65535 && max_n > 0 && (step >= 1 || step_dev), "adamw_step: bad arguments");
  V4H_REQUIRE(ema_decay < 1.f, "adamw_step: ema_decay must be < 1");
  return adamw_step(jobs, njobs, max_n, norm_sq, max_norm, lr, beta1, beta2, eps, weight_decay, step, step_dev, lr_dev,
                    ema_decay, ema_updates, ema_updates_dev, (cudaStream_t)s);
}
int v4h_ema_update(const v4h_adamw_job* jobs, int32_t njobs, int64_t max_n, float decay, int32_t num_updates,
                   const int32_t* num_updates_dev, v4h_stream_t s) {
  V4H_REQUIRE(jobs && njobs > 0 && njobs <= 65535 && max_n > 0 && decay > 0.f && decay < 1.f, "ema_update: bad arguments");
  return ema_update(jobs, njobs, max_n, decay, num_updates, num_updates_dev, (cudaStream_t)s);
}
int v4h_counter_increment(int32_t* counter, v4h_stream_t s) {
  V4H_REQUIRE(counter, "counter_increment: null");
  return counter_increment(counter, (cudaStream_t)s);
}
int64_t v4h_vit_arena_offset(const v4h_plan* p, const char* field) {
  if (!p || !field) return -1;
  return plan_arena_offset(reinterpret_cast<const Plan*>(p), field);
}

// ------------------------------------------------------------------------------------ measurement
int64_t v4h_launch_count(void) { return g_launches.load(); }

int v4h_profile_begin(void) {
  Profiler& p = profiler();
  std::lock_guard<std::mutex> lock(p.mu);
  for (ProfRecord& r : p.records) { p.pool.push_back(r.e0); p.pool.push_back(r.e1); }
  p.records.clear();
  p.on.store(true);
  return V4H_OK;
}

int v4h_profile_end(v4h_profile_entry* out, int32_t max, int32_t* n) {
  V4H_REQUIRE(out && n && max > 0, "profile_end: bad arguments");
  Profiler& p = profiler();
  p.on.store(false);
  V4H_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(p.mu);
  std::map<std::string, v4h_profile_entry> agg;
  std::vector<std::string> order;
  for (ProfRecord& r : p.records) {
    float ms = 0.f;
    V4H_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    auto it = agg.find(r.tag);
    if (it == agg.end()) {
      v4h_profile_entry e;
      memset(&e, 0, sizeof(e));
      strncpy(e.name, r.tag.c_str(), sizeof(e.name) - 1);
      it = agg.emplace(r.tag, e).first;
      order.push_back(r.tag);
    }
    it->second.launches += 1; it->second.ms += ms; it->second.flops += r.flops; it->second.bytes += r.bytes;
    p.pool.push_back(r.e0); p.pool.push_back(r.e1);
  }
  p.records.clear();
  int32_t k = 0;
  for (const std::string& name : order) {
    if (k >= max) break;
    out[k++] = agg[name];
  }
  *n = k;
  return V4H_OK;
}

// ------------------------------------------------------------------------------------ test hooks
int v4h_test_gemm(int32_t engine, int32_t layout, const void* A, const void* B, float* C, int32_t m, int32_t n,
                  int32_t k, v4h_stream_t s) {
  V4H_REQUIRE(A && B && C && layout >= 0 && layout <= 2, "test_gemm: bad arguments");
  GemmDesc g;
  g.layout = layout; g.A = A; g.B = B; g.M = m; g.N = n; g.K = k;
  g.lda = layout == GEMM_TN ? m : k;
  g.ldb = layout == GEMM_NT ? k : n;
  g.out_dtype = DT_F32; g.ep.out = C; g.ep.ldo = n;
  if (engine == 0) {
    return gemm_simt(g, (cudaStream_t)s);
  }
  g.a_dtype = g.b_dtype = DT_BF16;
  static UmmaContext* ctx = umma_context_create();
  if (layout == GEMM_TN) {  // weight-gradient form: split-K with atomics into a zeroed C
    g.epi = EPI_ATOMIC;
    g.splitk = 0;
    V4H_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * n, (cudaStream_t)s));
  }
  if (!gemm_umma_supported(g)) return fail(V4H_ERR_UNSUPPORTED, "test_gemm: shape not supported by the tcgen05 GEMM");
  return gemm_umma(ctx, g, (cudaStream_t)s);
}

int v4h_debug_gemm(int32_t kind, int32_t m, int32_t n, int32_t k, int32_t rows_per_sample, const void* A, const void* B,
                   const float* bias, void* out, void* out2, const float* res_in, float* res_out, const float* gate,
                   const void* aux, int64_t* counters, v4h_stream_t s) {
  V4H_REQUIRE(A && B && kind >= 0 && kind <= 6, "debug_gemm: bad arguments");
  static UmmaContext* ctx = umma_context_create();
  GemmDesc g;
  g.A = A; g.B = B; g.M = m; g.N = n; g.K = k; g.a_dtype = g.b_dtype = DT_BF16; g.out_dtype = DT_BF16;
  g.dbg = reinterpret_cast<long long*>(counters);
  g.ep.bias = bias; g.ep.ldo = n;
  switch (kind) {
    case 0:  // fc1-like: x W^T + b, GELU; both pre-activation and activation stored
      g.layout = GEMM_NT; g.lda = k; g.ldb = k; g.act = ACT_GELU_TANH; g.ep.out = out; g.ep.out2 = out2; break;
    case 1:  // qkv-like: x W^T + b
      g.layout = GEMM_NT; g.lda = k; g.ldb = k; g.ep.out = out; break;
    case 2:  // proj / fc2-like: gated residual update
      g.layout = GEMM_NT; g.lda = k; g.ldb = k; g.epi = EPI_GATE_RES; g.ep.out2 = out2; g.ep.gate = gate;
      g.ep.mod_stride = n; g.ep.rows_per_sample = rows_per_sample; g.ep.res_in = res_in; g.ep.res_out = res_out; break;
    case 3:  // dgrad through GELU: (dY W) * gelu'(u)
      g.layout = GEMM_NN; g.lda = k; g.ldb = n; g.epi = EPI_DACT; g.act = ACT_GELU_TANH; g.ep.out = out; g.ep.aux = aux;
      g.ep.ld_aux = n; g.ep.bias = nullptr; break;
    case 4:  // plain dgrad
      g.layout = GEMM_NN; g.lda = k; g.ldb = n; g.ep.out = out; g.ep.bias = nullptr; break;
    case 6:  // final-layer-like: x W^T + b into fp32
      g.layout = GEMM_NT; g.lda = k; g.ldb = k; g.ep.out = out; g.out_dtype = DT_F32; break;
    default:  // wgrad: dY^T X with split-K atomics into fp32
      g.layout = GEMM_TN; g.lda = m; g.ldb = n; g.epi = EPI_ATOMIC; g.out_dtype = DT_F32; g.splitk = 0;
      g.ep.out = out; g.ep.bias = nullptr; break;
  }
  V4H_REQUIRE(gemm_umma_supported(g), "debug_gemm: shape not supported by the tcgen05 GEMM");
  return gemm_umma(ctx, g, (cudaStream_t)s);
}

int v4h_debug_gemm_ln(int32_t m, int32_t n, int32_t k, int32_t rows_per_sample, const void* A, const void* W,
                      const float* bias, void* y, const float* res_in, float* res_out, const float* gate,
                      const float* shift, const float* scale, void* ln_out, int32_t ld_ln, float* stats, int64_t* counters,
                      v4h_stream_t s) {
  V4H_REQUIRE(A && W && res_in && res_out && gate && shift && scale && ln_out, "debug_gemm_ln: null argument");
  static UmmaContext* ctx = umma_context_create();
  GemmDesc g;
  g.tag = "gemm.ln"; g.layout = GEMM_NT; g.A = A; g.B = W; g.M = m; g.N = n; g.K = k; g.lda = k; g.ldb = k;
  g.a_dtype = g.b_dtype = DT_BF16; g.out_dtype = DT_BF16; g.epi = EPI_GATE_RES;
  g.ep.bias = bias; g.ep.out2 = y; g.ep.ldo = n; g.ep.gate = gate; g.ep.mod_stride = n; g.ep.rows_per_sample = rows_per_sample;
  g.ep.res_in = res_in; g.ep.res_out = res_out;
  g.ep.ln_shift = shift; g.ep.ln_scale = scale; g.ep.ln_out = ln_out; g.ep.ld_ln = ld_ln;
  g.ep.ln_stats = reinterpret_cast<float2*>(stats); g.ep.ln_eps = 1e-6f;
  g.dbg = reinterpret_cast<long long*>(counters);
  V4H_REQUIRE(gemm_gate_res_ln_supported(g), "debug_gemm_ln: shape not supported by the fused kernel");
  return gemm_gate_res_ln(ctx, g, (cudaStream_t)s);
}

int v4h_debug_tma_probe(const void* buf, int32_t rows, int32_t cols, int32_t stages, int32_t boxes, int32_t box_rows,
                        int32_t producers, int32_t iters, int32_t ctas, int64_t* cycles, v4h_stream_t s) {
  static UmmaContext* ctx = umma_context_create();
  return tma_probe(ctx, buf, rows, cols, stages, boxes, box_rows, producers, iters, ctas,
                   reinterpret_cast<long long*>(cycles), (cudaStream_t)s);
}

int v4h_debug_attention_counters(int64_t* counters) {
  attention_debug_counters(reinterpret_cast<long long*>(counters));
  return V4H_OK;
}

int v4h_test_attention_fwd(int32_t precision, int32_t engine, const void* qkv, void* o, float* lse, int32_t batch,
                           int32_t tokens, int32_t heads, int32_t head_dim, v4h_stream_t s) {
  V4H_REQUIRE(qkv && o && lse, "test_attention_fwd: null argument");
  if (engine != 0) {
    V4H_REQUIRE(precision == V4H_BF16, "test_attention_fwd: the tcgen05 kernels are bf16 only");
    return attention_fwd_umma((const bf16*)qkv, (bf16*)o, lse, batch, tokens, heads, head_dim, (cudaStream_t)s);
  }
  if (precision == V4H_BF16)
    return attention_fwd_simt<bf16>((const bf16*)qkv, (bf16*)o, lse, batch, tokens, heads, head_dim, (cudaStream_t)s);
  return attention_fwd_simt<float>((const float*)qkv, (float*)o, lse, batch, tokens, heads, head_dim, (cudaStream_t)s);
}
int v4h_test_attention_bwd(int32_t precision, int32_t engine, const void* qkv, const void* o, const float* lse,
                           const void* d_o, void* dqkv, int32_t batch, int32_t tokens, int32_t heads,
                           int32_t head_dim, v4h_stream_t s) {
  V4H_REQUIRE(qkv && o && lse && d_o && dqkv, "test_attention_bwd: null argument");
  if (engine != 0) {
    V4H_REQUIRE(precision == V4H_BF16, "test_attention_bwd: the tcgen05 kernels are bf16 only");
    float* delta = nullptr;  // test hook only: scratch for rowsum(dO * O)
    V4H_CUDA(cudaMallocAsync(&delta, sizeof(float) * (size_t)batch * heads * tokens, (cudaStream_t)s));
    const int rc = attention_bwd_umma((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)d_o, delta, (bf16*)dqkv,
                                      batch, tokens, heads, head_dim, (cudaStream_t)s);
    cudaFreeAsync(delta, (cudaStream_t)s);
    return rc;
  }
  if (precision == V4H_BF16)
    return attention_bwd_simt<bf16>((const bf16*)qkv, (const bf16*)o, lse, (const bf16*)d_o, (bf16*)dqkv, batch,
                                    tokens, heads, head_dim, (cudaStream_t)s);
  return attention_bwd_simt<float>((const float*)qkv, (const float*)o, lse, (const float*)d_o, (float*)dqkv, batch,
                                   tokens, heads, head_dim, (cudaStream_t)s);
}

}  // extern "C"
