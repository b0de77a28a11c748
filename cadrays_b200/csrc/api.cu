// api.cu -- C-ABI of libcadrays_b200.so (include/cadrays_b200.h) and the host
// orchestration of the wavefront path tracer.  The reference call site each entry
// point replaces is cited in the header.
#include "../../include/cadrays_b200.h"
#include "host_scene.hpp"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace crt;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg)
{
  g_error = msg;
  return code;
}

#define CRT_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      cudaGetLastError();                                                                       \
      return fail(e_ == cudaErrorMemoryAllocation ? CRT_ERR_OUT_OF_MEMORY : CRT_ERR_CUDA,       \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                          \
    }                                                                                           \
  } while (0)

#define CRT_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) return fail(CRT_ERR_INVALID_ARG, msg);         \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count)
  {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

constexpr size_t kCounterWords = 4 * 64 + 8;   // queue lengths and work counters of one wave part

enum Family { F_GENERATE = 0, F_EXTEND, F_SHADE, F_CONNECT, F_RESOLVE, F_RENDER, F_COUNT };

struct TimedSpan { int family; cudaEvent_t a, b; };

}  // namespace

struct crt_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;

  // host scene (what CADRays hands to OCCT)
  HostScene scene;
  bool geometry_dirty = true;
  std::vector<uint8_t> blob;
  bool has_layout = false;
  bool blob_on_device = false;       // the device layout was converted from `blob` as it stands (top-level patches apply)
  bool quad = false;                 // the uploaded layout is the 4-wide one (crt_params.bvh_width = 4)
  std::vector<crt_bsdf> mats;
  std::vector<float> lights;         // 8 floats per light, shader form
  std::vector<float> env;            // rgba per texel
  uint32_t env_w = 0, env_h = 0;
  std::vector<uint8_t> tex_texels;   // RGBA8 of all textures, back to back
  std::vector<uint32_t> tex_table;   // offset, width, height per texture
  bool mats_dirty = true, lights_dirty = true, env_dirty = true, tex_dirty = true;
  crt_params params;
  crt_camera cam;
  uint32_t width = 0, height = 0;

  // device scene
  DevBuf<float4> d_arena, d_mats, d_lights, d_env;
  DevBuf<uint8_t> d_mat_class;
  DevBuf<float2> d_tri_uv;
  DevBuf<float4> d_top_cache;
  DevBuf<uchar4> d_tex;
  DevBuf<uint32_t> d_tex_table;
  DeviceScene ds{};
  DeviceParams dp{};
  size_t arena_nodes = 0, arena_inst_off = 0;                // float4 offsets of the sections inside d_arena (partial re-upload)
  size_t scene_traversal_bytes = 0, scene_total_bytes = 0;   // crt_scene_bytes
  DeviceLayout layout_info;          // the uploaded layout without its big arrays (counts, depths, mesh_root_ref)
  uint64_t top_patches = 0;          // commits that re-uploaded the top-level nodes + instance records only
  // change counters: crt_group replicates what changed on the other GPUs of the group
  uint64_t gen_scene = 0, gen_mats = 0, gen_lights = 0, gen_env = 0, gen_tex = 0, gen_accum = 0;

  // path state
  DevBuf<float4> ray_o, ray_d, thr, rad, hit, sh_o, sh_d, sh_c;
  DevBuf<int32_t> hit_inst;
  DevBuf<uint32_t> queue0, queue1, counters, seeds;
  uint32_t state_capacity = 0;       // path slots allocated
  uint32_t batch = 0;                // samples per pixel in flight
  std::vector<uint32_t> h_seeds;

  // accumulation
  DevBuf<float4> accum_internal;
  float4* accum = nullptr;
  bool accum_external = false;
  uint64_t first_sample = 0, next_sample = 0;
  uint32_t rng_hi = 0, rng_lo = 0; uint64_t rng_index = 0; bool rng_valid = false;

  // adaptive screen sampling (crt_params.adaptive_sampling): per-tile state, see kernels.cuh
  DevBuf<uint32_t> ad_count, ad_err, ad_cum, ad_qoff, ad_seeds, ad_cum_own;
  DevBuf<float> ad_even;
  std::vector<uint32_t> ad_h_seeds;  // frame seeds of sample indices first_sample + [0, size)
  size_t ad_seeds_uploaded = 0;
  uint64_t ad_bound = 0;             // upper bound of any tile's sample count after the waves enqueued so far
  uint32_t ad_wave = 0;
  bool ad_clear = true;              // state must be zeroed before the next adaptive wave

  // read-back staging
  DevBuf<uint8_t> d_ldr;
  DevBuf<float> d_hdr;

  // traversal driver: per-lane-refill persistent kernels (default) or the static form (A/B knob:
  // environment CRT_TRAVERSAL=static)
  bool persistent = true;
  bool fuse_traversal = true;   // connect(d) + extend(d+1) in one launch (CRT_FUSE=0 disables)
  bool l2_persist = false;      // L2 access-policy window over the scene arena (CRT_L2_PERSIST=1)
  // software pipeline: a wave is split into two half-waves on two streams so that the latency-bound shading of
  // one half overlaps the traversal of the other (CRT_PIPELINE=0/1); traversal CTAs per SM when pipelined
  bool mats_lean = false;       // no material has a coat or transmission: k_shade<.., LEAN> (CRT_SHADE_LEAN=0 disables)
  bool shade_lean = true;
  bool shade_classes = true;    // k_shade files paths by the shading class of the hit material (CRT_SHADE_CLASSES=0: hit / miss only)
  bool smem_mats = true;        // material tables of up to kSmemMats records are staged in shared memory by k_shade (CRT_SMEM_MATS=0 disables)
  bool shade_sort = true;       // hit / miss grouping inside k_shade after the first bounce (CRT_SHADE_SORT=0 disables)
  bool primary_lockstep = true; // camera rays walked in lockstep per 8x4 tile instead of per-lane refill (CRT_PRIMARY_LOCKSTEP=0)
  int primary_grid = 72;        // CTAs per SM of the lockstep kernels' grid-stride grid (measured: 9 / 16 / 36 / 72 / 144 / 576 ->
                                // 18.90 / 18.80 / 18.57 / 18.45 / 18.43 / 18.47 ms of traversal per step)
  int sample_group = 32;        // samples of a pixel block that share a warp (1, 4, 8, 16, 32; the largest that divides the wave's sample count is used; CRT_SAMPLE_GROUP)
  bool fuse_primary = true;     // depth 0 without a generate pass (CRT_FUSE_PRIMARY=0 disables); tile-aligned sizes only
  // per-path kernel for the thin end of a wave (k_tail): from bounce `tail_min_depth` on, once at most `tail_max` paths
  // are active, one launch carries them to their end (CRT_TAIL=0 disables, CRT_TAIL_MAX / CRT_TAIL_MIN_DEPTH tune)
  bool tail = true;
  uint32_t tail_max = 0;        // 0 = one path per resident thread of k_tail (set in crt_create)
  int tail_min_depth = 1;
  // a wave split into `parts` ranges of 8x4 pixel tiles, each on its own stream: every launch of a thin wave ends
  // with a tail in which a few long rays finish while most SMs idle, and the parts fill each other's tails.
  // pipeline: 0 = never, 1 = always, 2 = automatic: waves of at most pipeline_auto_paths paths, unless per-family timing
  // is on (crt_timing_enable: a kernel's duration is only defined while kernels do not overlap, so timed waves run
  // unsplit on one stream).  Measured with two parts, 1080p: C2 +3.4 % at 16 spp per wave and 405 -> 432 frames/s at
  // one sample per Redraw, C5 flattened +6.5 %, C3 +1.5 %, Cornell +-0.  CRT_PIPELINE / CRT_PIPELINE_PARTS / CRT_PIPELINE_AUTO_PATHS.
  int pipeline = 2;
  int pipeline_parts = 2;
  uint64_t pipeline_auto_paths = 1ull << 40;
  int pipeline_trace_ctas = 0;   // cap of the traversal grids (CTAs per SM) while split; 0 = none
  static constexpr int kMaxParts = 4;
  cudaStream_t side_stream[kMaxParts - 1] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxParts - 1] = {};
  struct WavePart { size_t slot0; int index; };
  std::vector<WavePart> last_parts;   // how the last wave was laid out (crt_wavefront_rays)

  // metrics
  DevBuf<Counters> d_counters;
  bool stats_on = false;
  bool timing_on = false;
  std::vector<TimedSpan> spans;
  std::vector<cudaEvent_t> event_pool;
  double family_ms[F_COUNT] = {};
  uint64_t family_launches[F_COUNT] = {};
  uint64_t kernel_launches = 0;   // kernels enqueued since crt_stats_reset (always counted; crt_launch_count)
};

namespace {

int set_device(crt_context* c)
{
  if (c->device < 0) return fail(CRT_ERR_NO_DEVICE, "host-only context: no CUDA device bound (there is no CPU fallback)");
  CRT_CUDA(cudaSetDevice(c->device));
  return CRT_OK;
}

// math_BullardGenerator: seed of sample index k = k-th NextInt() >> 2.
uint32_t frame_seed(crt_context* c, uint64_t index)
{
  if (!c->rng_valid || index < c->rng_index) {
    c->rng_hi = c->params.frame_seed0;
    c->rng_lo = c->params.frame_seed0 ^ 0x49616E42u;
    c->rng_index = 0;
    c->rng_valid = true;
  }
  uint32_t out = 0;
  // state after producing sample (rng_index - 1); step until `index` is produced
  while (c->rng_index <= index) {
    c->rng_hi = (c->rng_hi >> 2) + (c->rng_hi << 2);
    c->rng_hi += c->rng_lo;
    c->rng_lo += c->rng_hi;
    out = c->rng_hi;
    c->rng_index++;
  }
  return out >> 2;
}

cudaEvent_t get_event(crt_context* c)
{
  if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

struct SpanGuard {
  crt_context* c; int family; cudaStream_t s; cudaEvent_t a = nullptr;
  SpanGuard(crt_context* c_, int f, cudaStream_t s_ = nullptr) : c(c_), family(f), s(s_ ? s_ : c_->stream)
  {
    if (family != F_RENDER) c->kernel_launches++;      // every guarded scope holds one kernel launch (k_tail is added where it is launched)
    if (c->timing_on) { a = get_event(c); cudaEventRecord(a, s); }
  }
  ~SpanGuard()
  {
    if (a) { cudaEvent_t b = get_event(c); cudaEventRecord(b, s); c->spans.push_back({ family, a, b }); }
  }
};

void reset_accum_state(crt_context* c)
{
  c->next_sample = c->first_sample;
  c->ad_clear = true;
  c->gen_accum++;
  if (c->device >= 0 && c->accum && c->width && c->height)
    cudaMemsetAsync(c->accum, 0, sizeof(float4) * (size_t)c->width * c->height, c->stream);
}

// camera basis and light table: same formulas as the oracle's orc_set_camera /
// orc_set_lights (host libm tanf/cosf on both sides)
void update_device_params(crt_context* c)
{
  DeviceParams& P = c->dp;
  const crt_params& p = c->params;
  P.max_depth = std::max(1, std::min(p.max_depth, 64));
  P.max_radiance = p.max_radiance;
  P.two_sided = p.two_sided;
  P.rng_radius = p.coherent_rng ? 8u : 1u;
  P.aperture_radius = p.aperture_radius;
  P.focal_dist = p.focal_dist;
  P.env_as_background = (p.env_as_background && c->env_w) ? 1 : 0;
  P.russian_roulette = p.russian_roulette;
  for (int k = 0; k < 3; ++k) P.background[k] = p.background[k];
  const crt_camera& cam = c->cam;
  v3 w = normalize3(V(cam.dir[0], cam.dir[1], cam.dir[2]));
  v3 u = normalize3(cross3(w, V(cam.up[0], cam.up[1], cam.up[2])));
  v3 v = cross3(u, w);
  P.cu[0] = u.x; P.cu[1] = u.y; P.cu[2] = u.z;
  P.cv[0] = v.x; P.cv[1] = v.y; P.cv[2] = v.z;
  P.cw[0] = w.x; P.cw[1] = w.y; P.cw[2] = w.z;
  for (int k = 0; k < 3; ++k) P.eye[k] = cam.eye[k];
  P.hh = cam.is_ortho ? cam.ortho_scale * 0.5f : tanf(cam.fovy_deg * 0.5f * 0.0174532925199433f);
  P.hw = P.hh * cam.aspect;
  P.is_ortho = cam.is_ortho;
  P.width = c->width; P.height = c->height;
  P.tiles_x = (c->width + 7u) / 8u;
  P.tiles_y = (c->height + 3u) / 4u;
  P.tile0 = 0;
  P.n_tiles = P.tiles_x * P.tiles_y;
  P.sample_group = 1;
}

int ensure_path_slots(crt_context* c, uint64_t need)
{
  if (need > 0x1fffffffull) return fail(CRT_ERR_INVALID_ARG, "batch too large");   // k_shade packs (slot, class) in 32 bits
  if (need <= c->state_capacity) return CRT_OK;
  const size_t n = (size_t)need;
  CRT_CUDA(c->ray_o.ensure(n)); CRT_CUDA(c->ray_d.ensure(n)); CRT_CUDA(c->thr.ensure(n));
  CRT_CUDA(c->rad.ensure(n)); CRT_CUDA(c->hit.ensure(n)); CRT_CUDA(c->hit_inst.ensure(n));
  CRT_CUDA(c->queue0.ensure(n)); CRT_CUDA(c->queue1.ensure(n));
  CRT_CUDA(c->sh_o.ensure(n)); CRT_CUDA(c->sh_d.ensure(n)); CRT_CUDA(c->sh_c.ensure(n));
  c->state_capacity = (uint32_t)need;
  return CRT_OK;
}

int ensure_path_state(crt_context* c, uint32_t batch)
{
  return ensure_path_slots(c, (uint64_t)c->dp.tiles_x * c->dp.tiles_y * 32u * batch);
}

uint32_t auto_batch(const crt_context* c)
{
  if (c->params.samples_per_batch > 0) return (uint32_t)c->params.samples_per_batch;
  const uint64_t px = (uint64_t)c->width * c->height;
  if (!px) return 1;
  uint64_t b = (32ull << 20) / px;  // about 32 M paths in flight (5 GB of path state): 16 samples per wave at 1080p, which also
                                    // lets a warp hold 16 samples of a pixel pair (sample_group)
  return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(b, 64));
}

int upload_layout(crt_context* c, const DeviceLayout& L, float scene_eps)
{
  // validate before anything on the device changes: a failed upload must leave the previous layout usable
  if (L.quad ? 3 * (L.max_depth_top + L.max_depth_bottom) + 4 > kStackSizeQuad : L.max_depth_top + L.max_depth_bottom + 4 > kStackSize)
    return fail(CRT_ERR_FORMAT, "BVH deeper than the traversal stack");
  // one arena [nodes | triangle vertices | instances | vertex normals]: the traversal working set is
  // contiguous, so a single L2 access-policy window can cover it
  auto pad = [](size_t n) { return (n + 15u) & ~(size_t)15u; };   // float4 units, 256-byte sections
  const size_t n_nodes = pad(std::max<size_t>(L.nodes.size(), 4)), n_verts = pad(std::max<size_t>(L.tri_verts.size(), 3));
  const size_t n_inst = pad(std::max<size_t>(L.inst.size(), 4)), n_nrm = pad(std::max<size_t>(L.tri_nrm.size(), 3));
  // new buffers are allocated before the old ones are released (DevBuf::ensure frees first, so grow into
  // temporaries and swap): an out-of-memory failure keeps c->ds pointing at live memory
  const size_t n_arena = n_nodes + n_verts + n_inst + n_nrm, n_uv = std::max<size_t>(L.tri_uv.size() / 2, 3);
  const uint32_t n_cache = L.quad ? 0u : std::min<uint32_t>(L.n_top_inner, 1024u);
  const size_t n_cache4 = (size_t)5 * std::max<uint32_t>(n_cache, 1);
  DevBuf<float4> arena_new, cache_new;
  DevBuf<float2> uv_new;
  const bool grow_arena = n_arena > c->d_arena.n || !c->d_arena.p, grow_uv = n_uv > c->d_tri_uv.n || !c->d_tri_uv.p;
  const bool grow_cache = n_cache4 > c->d_top_cache.n || !c->d_top_cache.p;
  cudaError_t ae = cudaSuccess;
  if (grow_arena) ae = arena_new.ensure(n_arena);
  if (ae == cudaSuccess && grow_uv) ae = uv_new.ensure(n_uv);
  if (ae == cudaSuccess && grow_cache) ae = cache_new.ensure(n_cache4);
  if (ae != cudaSuccess) {
    arena_new.release(); uv_new.release(); cache_new.release();
    cudaGetLastError();
    return fail(ae == cudaErrorMemoryAllocation ? CRT_ERR_OUT_OF_MEMORY : CRT_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(ae));
  }
  // from here on the previous layout is being replaced: a failure leaves the context without one
  c->has_layout = false;
  CRT_CUDA(cudaStreamSynchronize(c->stream));       // no kernel still reads the buffers about to be released
  if (grow_arena) { c->d_arena.release(); c->d_arena = arena_new; }
  if (grow_uv) { c->d_tri_uv.release(); c->d_tri_uv = uv_new; }
  if (grow_cache) { c->d_top_cache.release(); c->d_top_cache = cache_new; }
  float4* nodes = c->d_arena.p;
  float4* verts = nodes + n_nodes;
  float4* inst = verts + n_verts;
  float4* nrm = inst + n_inst;
  c->ds.nodes = nodes; c->ds.tri_verts = verts; c->ds.tri_nrm = nrm; c->ds.inst = inst;
  c->ds.tri_uv = c->d_tri_uv.p;
  c->ds.top_cache = c->d_top_cache.p;
  c->ds.n_top_cache = 0;
  c->ds.top_root = kRefNone;
  c->arena_nodes = n_nodes; c->arena_inst_off = n_nodes + n_verts;
  if (!L.nodes.empty()) CRT_CUDA(cudaMemcpyAsync(nodes, L.nodes.data(), L.nodes.size() * 16, cudaMemcpyHostToDevice, c->stream));
  if (!L.tri_verts.empty()) CRT_CUDA(cudaMemcpyAsync(verts, L.tri_verts.data(), L.tri_verts.size() * 16, cudaMemcpyHostToDevice, c->stream));
  if (!L.tri_nrm.empty()) CRT_CUDA(cudaMemcpyAsync(nrm, L.tri_nrm.data(), L.tri_nrm.size() * 16, cudaMemcpyHostToDevice, c->stream));
  if (!L.inst.empty()) CRT_CUDA(cudaMemcpyAsync(inst, L.inst.data(), L.inst.size() * 16, cudaMemcpyHostToDevice, c->stream));
  if (!L.tri_uv.empty()) CRT_CUDA(cudaMemcpyAsync(c->d_tri_uv.p, L.tri_uv.data(), L.tri_uv.size() * 4, cudaMemcpyHostToDevice, c->stream));
  {
    // padded copy of the first top-level nodes (breadth-first = top of the tree) for shared-memory staging
    std::vector<f4> cache(n_cache4, f4{ 0, 0, 0, 0 });
    for (uint32_t k = 0; k < n_cache; ++k)
      for (int q = 0; q < 4; ++q) cache[5 * (size_t)k + q] = L.nodes[4 * (size_t)k + q];
    CRT_CUDA(cudaMemcpyAsync(c->d_top_cache.p, cache.data(), cache.size() * 16, cudaMemcpyHostToDevice, c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    c->ds.n_top_cache = n_cache;
  }
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  c->ds.top_root = L.top_root;
  c->ds.scene_eps = scene_eps;
  c->quad = L.quad;
  c->scene_traversal_bytes = (L.nodes.size() + L.tri_verts.size() + L.inst.size()) * 16;
  c->scene_total_bytes = c->scene_traversal_bytes + L.tri_nrm.size() * 16 + L.tri_uv.size() * 4;
  if (c->l2_persist) {
    // keep nodes + triangle vertices + instance records resident in the 126 MB L2 while gigabytes of
    // path state stream past them (hit ratio scaled to the set-aside the device allows)
    cudaDeviceProp prop;
    CRT_CUDA(cudaGetDeviceProperties(&prop, c->device));
    const size_t want = (n_nodes + n_verts + n_inst) * sizeof(float4);
    const size_t window = std::min<size_t>(want, (size_t)prop.accessPolicyMaxWindowSize);
    const size_t setaside = std::min<size_t>(window, (size_t)prop.persistingL2CacheMaxSize);
    if (window > 0 && setaside > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) == cudaSuccess) {
      cudaStreamAttrValue attr;
      std::memset(&attr, 0, sizeof attr);
      attr.accessPolicyWindow.base_ptr = nodes;
      attr.accessPolicyWindow.num_bytes = window;
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setaside / (double)window);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    } else {
      cudaGetLastError();
    }
  }
  c->has_layout = true;
  c->gen_scene++;
  c->layout_info = DeviceLayout{};
  c->layout_info.top_root = L.top_root; c->layout_info.n_tris = L.n_tris; c->layout_info.n_inst = L.n_inst;
  c->layout_info.n_top_inner = L.n_top_inner; c->layout_info.quad = L.quad;
  c->layout_info.max_depth_top = L.max_depth_top; c->layout_info.max_depth_bottom = L.max_depth_bottom;
  c->layout_info.mesh_root_ref = L.mesh_root_ref;
  return CRT_OK;
}

// Device side of an instance-only edit (SetLocation, ImRaytraceControls.cxx:88; material assignment,
// MaterialEditor.cxx:522-523): `T` holds the re-emitted top-level nodes and the instance records; the bottom
// trees, triangles and normals already on the device are left alone (about 200 KB instead of 120 MB for the
// 1 M-triangle assembly).
int upload_layout_top(crt_context* c, const DeviceLayout& T, float scene_eps)
{
  if (T.max_depth_top + T.max_depth_bottom + 4 > kStackSize) return fail(CRT_ERR_FORMAT, "BVH deeper than the traversal stack");
  if (T.nodes.size() > c->arena_nodes || T.inst.size() > c->d_arena.n - c->arena_inst_off)
    return fail(CRT_ERR_STATE, "top-level patch does not fit the uploaded layout");
  CRT_CUDA(cudaStreamSynchronize(c->stream));       // no kernel still walks the records about to change
  float4* nodes = c->d_arena.p;
  float4* inst = c->d_arena.p + c->arena_inst_off;
  if (!T.nodes.empty()) CRT_CUDA(cudaMemcpyAsync(nodes, T.nodes.data(), T.nodes.size() * 16, cudaMemcpyHostToDevice, c->stream));
  if (!T.inst.empty()) CRT_CUDA(cudaMemcpyAsync(inst, T.inst.data(), T.inst.size() * 16, cudaMemcpyHostToDevice, c->stream));
  const uint32_t n_cache = std::min<uint32_t>(T.n_top_inner, 1024u);
  std::vector<f4> cache((size_t)5 * std::max<uint32_t>(n_cache, 1), f4{ 0, 0, 0, 0 });
  for (uint32_t k = 0; k < n_cache; ++k)
    for (int q = 0; q < 4; ++q) cache[5 * (size_t)k + q] = T.nodes[4 * (size_t)k + q];
  if (cache.size() <= c->d_top_cache.n)
    CRT_CUDA(cudaMemcpyAsync(c->d_top_cache.p, cache.data(), cache.size() * 16, cudaMemcpyHostToDevice, c->stream));
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  c->ds.top_root = T.top_root;
  c->ds.scene_eps = scene_eps;
  c->layout_info.top_root = T.top_root;
  c->layout_info.max_depth_top = T.max_depth_top;
  c->top_patches++;
  c->gen_scene++;
  return CRT_OK;
}

// `patched`: build_blob rewrote only the header, the top-level nodes and the instance records of the blob this
// context's device layout came from, so the device copy is patched the same way; `keep` (optional) receives the
// layout that was uploaded so that crt_group can upload it to the other GPUs without converting the blob again.
int load_blob(crt_context* c, bool patched = false, DeviceLayout* keep = nullptr, bool* keep_is_top = nullptr)
{
  BlobView view;
  std::string err;
  if (!parse_blob(c->blob.data(), c->blob.size(), view, err)) return fail(CRT_ERR_FORMAT, err);
  DeviceLayout local;
  DeviceLayout& L = keep ? *keep : local;
  if (keep_is_top) *keep_is_top = false;
  if (patched && c->has_layout && !c->quad && build_device_layout_top(view, c->layout_info, L, err)) {
    const int rc = upload_layout_top(c, L, view.hdr.scene_eps);
    if (rc == CRT_OK) { if (keep_is_top) *keep_is_top = true; return rc; }
    if (rc != CRT_ERR_STATE) return rc;
  }
  err.clear();
  if (!build_device_layout(view, L, err)) return fail(CRT_ERR_FORMAT, err);
  return upload_layout(c, L, view.hdr.scene_eps);
}

int upload_tables(crt_context* c)
{
  if (c->mats_dirty) {
    CRT_CUDA(c->d_mats.ensure(std::max<size_t>(c->mats.size() * 8, 8)));
    if (!c->mats.empty())
      CRT_CUDA(cudaMemcpyAsync(c->d_mats.p, c->mats.data(), c->mats.size() * sizeof(crt_bsdf), cudaMemcpyHostToDevice, c->stream));
    c->ds.mats = c->d_mats.p; c->ds.n_mats = (uint32_t)c->mats.size();
    c->mats_lean = c->shade_lean;
    std::vector<uint8_t> cls(std::max<size_t>(c->mats.size(), 1), (uint8_t)kClassDiffuse);
    for (size_t i = 0; i < c->mats.size(); ++i) {
      const crt_bsdf& m = c->mats[i];
      bool coat = false, trans = false, spec = false;
      for (int k = 0; k < 3; ++k) { coat = coat || m.Kc[k] != 0.0f; trans = trans || m.Kt[k] != 0.0f; spec = spec || m.Ks[k] != 0.0f; }
      if (coat || trans) c->mats_lean = false;
      cls[i] = (uint8_t)(trans ? kClassTransmissive : coat ? kClassCoat : spec ? kClassGlossy : kClassDiffuse);
    }
    CRT_CUDA(c->d_mat_class.ensure(cls.size()));
    CRT_CUDA(cudaMemcpyAsync(c->d_mat_class.p, cls.data(), cls.size(), cudaMemcpyHostToDevice, c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));      // `cls` is a local
    c->ds.mat_class = c->d_mat_class.p;
    c->ds.mats_in_smem = (c->smem_mats && !c->mats.empty() && c->mats.size() <= kSmemMats) ? 1u : 0u;
    c->mats_dirty = false;
  }
  if (c->lights_dirty) {
    CRT_CUDA(c->d_lights.ensure(std::max<size_t>(c->lights.size() / 4, 2)));
    if (!c->lights.empty())
      CRT_CUDA(cudaMemcpyAsync(c->d_lights.p, c->lights.data(), c->lights.size() * 4, cudaMemcpyHostToDevice, c->stream));
    c->ds.lights = c->d_lights.p; c->ds.n_lights = (uint32_t)(c->lights.size() / 8);
    c->lights_dirty = false;
  }
  if (c->tex_dirty) {
    CRT_CUDA(c->d_tex.ensure(std::max<size_t>(c->tex_texels.size() / 4, 1)));
    CRT_CUDA(c->d_tex_table.ensure(std::max<size_t>(c->tex_table.size(), 3)));
    if (!c->tex_texels.empty())
      CRT_CUDA(cudaMemcpyAsync(c->d_tex.p, c->tex_texels.data(), c->tex_texels.size(), cudaMemcpyHostToDevice, c->stream));
    if (!c->tex_table.empty())
      CRT_CUDA(cudaMemcpyAsync(c->d_tex_table.p, c->tex_table.data(), c->tex_table.size() * 4, cudaMemcpyHostToDevice, c->stream));
    c->ds.tex_data = c->d_tex.p; c->ds.tex_table = c->d_tex_table.p; c->ds.n_tex = (uint32_t)(c->tex_table.size() / 3);
    c->tex_dirty = false;
  }
  if (c->env_dirty) {
    CRT_CUDA(c->d_env.ensure(std::max<size_t>(c->env.size() / 4, 1)));
    if (!c->env.empty())
      CRT_CUDA(cudaMemcpyAsync(c->d_env.p, c->env.data(), c->env.size() * 4, cudaMemcpyHostToDevice, c->stream));
    c->ds.env = c->d_env.p; c->ds.env_w = c->env_w; c->ds.env_h = c->env_h;
    c->env_dirty = false;
  }
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  return CRT_OK;
}

// host BVH build (or in-place patch of the previous blob) + device upload, when the geometry changed
int commit_geometry(crt_context* c, DeviceLayout* keep, bool* keep_is_top)
{
  if (!c->geometry_dirty) return CRT_OK;
  std::string err;
  const uint64_t patched_before = c->scene.blobs_patched;
  const bool same_blob = c->blob_on_device;      // the device layout was made from c->blob as it is now
  if (!build_blob(c->scene, c->blob, err, c->params.bvh_width == 4 ? 4 : 2)) { c->blob_on_device = false; return fail(CRT_ERR_INVALID_ARG, err); }
  const bool patched = same_blob && c->scene.blobs_patched != patched_before;
  c->blob_on_device = false;
  const int rc = load_blob(c, patched, keep, keep_is_top);
  if (rc) return rc;
  c->blob_on_device = true;
  c->geometry_dirty = false;
  reset_accum_state(c);
  return CRT_OK;
}

int grid_for(const crt_context* c, int blocks_per_sm) { return c->sm_count * blocks_per_sm; }

// resident CTAs per SM of a kernel (a property of the kernel and the architecture, so callers may cache it)
template <typename K>
int resident_per_sm(K kernel, int block)
{
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  return per_sm;
}

// grid of a persistent kernel = the CTAs that are resident at once on this context's device
template <typename K>
int resident_grid(const crt_context* c, K kernel, int block) { return c->sm_count * resident_per_sm(kernel, block); }

PathState make_state(crt_context* c, size_t slot0, int half)
{
  PathState st;
  st.ray_o = c->ray_o.p + slot0; st.ray_d = c->ray_d.p + slot0; st.thr = c->thr.p + slot0; st.rad = c->rad.p + slot0;
  st.hit = c->hit.p + slot0; st.hit_inst = c->hit_inst.p + slot0;
  st.queue[0] = c->queue0.p + slot0; st.queue[1] = c->queue1.p + slot0;
  st.sh_o = c->sh_o.p + slot0; st.sh_d = c->sh_d.p + slot0; st.sh_c = c->sh_c.p + slot0;
  const int depth_max = c->dp.max_depth;
  st.n_active = c->counters.p + (size_t)half * kCounterWords;
  st.n_shadow = st.n_active + (depth_max + 1);
  st.work_extend = st.n_shadow + depth_max;
  st.work_connect = st.work_extend + depth_max;
  return st;
}

// generate + all bounces of one (half-)wave on stream s; path state indices are relative to st
template <bool COUNT, bool QUAD>
int enqueue_bounces(crt_context* c, const PathState& st, const DeviceParams& dp, cudaStream_t s, uint32_t n_batch, const uint32_t* d_seeds,
                    const AdaptiveState* adaptive, int trace_ctas)
{
  const int depth_max = dp.max_depth;
  Counters* gc = c->d_counters.p;
  const bool pers = c->persistent || QUAD;      // the 4-wide walk exists in the persistent driver only
  const bool fuse = pers && c->fuse_traversal;
  static const int p_ext = resident_per_sm(k_extend<COUNT, true, QUAD>, CRT_TRACE_BLOCK);
  static const int p_con = resident_per_sm(k_connect<COUNT, true, QUAD>, CRT_TRACE_BLOCK);
  static const int p_dual = resident_per_sm(k_trace_dual<COUNT, QUAD>, CRT_TRACE_BLOCK);
  const int r_ext = c->sm_count * p_ext, r_con = c->sm_count * p_con, r_dual = c->sm_count * p_dual;
  const int cap = trace_ctas > 0 ? c->sm_count * trace_ctas : (1 << 30);
  const int g_ext = std::min(r_ext, cap), g_con = std::min(r_con, cap), g_dual = std::min(r_dual, cap);
  // camera rays computed inside the depth-0 kernels instead of a generate pass
  const bool primary = c->fuse_primary && pers && !adaptive && (c->width & 7u) == 0 && (c->height & 3u) == 0;
  static const int p_pri = resident_per_sm(k_extend_primary<COUNT, QUAD>, CRT_TRACE_BLOCK);
  const int r_pri = c->sm_count * p_pri;
  if (!primary) {
    SpanGuard g(c, F_GENERATE, s);
    if (adaptive) k_generate_adaptive<<<grid_for(c, 8), 256, 0, s>>>(st, dp, *adaptive, d_seeds, n_batch);
    else k_generate<<<grid_for(c, 8), 256, 0, s>>>(st, dp, d_seeds, n_batch);
  }
  for (int depth = 0; depth < depth_max; ++depth) {
    // closest hits of this bounce; when fused, the same launch also resolves the previous bounce's shadow rays
    {
      SpanGuard g(c, F_EXTEND, s);
      if (primary && depth == 0 && !QUAD && c->primary_lockstep)
        k_extend_primary_lockstep<COUNT><<<grid_for(c, c->primary_grid), CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, dp, d_seeds, n_batch, gc);
      else if (primary && depth == 0) k_extend_primary<COUNT, QUAD><<<std::min(r_pri, cap), CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, dp, d_seeds, n_batch, gc);
      else if (fuse && depth > 0) k_trace_dual<COUNT, QUAD><<<g_dual, CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, depth, gc);
      else if (pers && !(depth == 0 && !QUAD && c->primary_lockstep)) k_extend<COUNT, true, QUAD><<<g_ext, CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, depth, gc);
      // camera rays that went through a generate pass (adaptive sampling, ragged sizes): lockstep as well
      else k_extend<COUNT, false, false><<<grid_for(c, depth == 0 ? c->primary_grid : 16), CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, depth, gc);
    }
    {
      SpanGuard g(c, F_SHADE, s);
      const dim3 sg(grid_for(c, 8));
      const bool lean = c->mats_lean && !c->ds.n_tex;
      // the thin end of the wave: k_tail finishes the paths of this bounce when few are left, k_shade then skips them
      const uint32_t tail_max = (c->tail && !QUAD && depth >= std::max(1, c->tail_min_depth)) ? c->tail_max : 0u;   // k_tail walks the binary layout
      if (tail_max) {
        c->kernel_launches++;
        const int tg = c->sm_count * CRT_TAIL_MIN_BLOCKS;
        if (c->ds.n_tex) k_tail<COUNT, true, false><<<tg, 128, 0, s>>>(c->ds, dp, st, depth, tail_max, gc);
        else if (lean) k_tail<COUNT, false, true><<<tg, 128, 0, s>>>(c->ds, dp, st, depth, tail_max, gc);
        else k_tail<COUNT, false, false><<<tg, 128, 0, s>>>(c->ds, dp, st, depth, tail_max, gc);
      }
#define CRT_SHADE(TEXV, FIRSTV, SORTV, LEANV, CLSV) k_shade<COUNT, TEXV, FIRSTV, SORTV, LEANV, CLSV><<<sg, 128, 0, s>>>(c->ds, dp, st, depth, gc, d_seeds, tail_max)
      // filing by shading class pays where warps would otherwise mix coat / transmission / base lobes (Cornell: shade -10 %);
      // with only diffuse and glossy materials the two-pass filing costs more than it returns (C2: shade +7 %)
      const bool classes = c->shade_classes && !lean;
      if (primary && depth == 0) {
        if (c->ds.n_tex) CRT_SHADE(true, true, false, false, false); else if (lean) CRT_SHADE(false, true, false, true, false); else CRT_SHADE(false, true, false, false, false);
      } else if (c->shade_sort && depth > 0) {
        if (c->ds.n_tex) { if (classes) CRT_SHADE(true, false, true, false, true); else CRT_SHADE(true, false, true, false, false); }
        else if (lean) CRT_SHADE(false, false, true, true, false);
        else if (classes) CRT_SHADE(false, false, true, false, true);
        else CRT_SHADE(false, false, true, false, false);
      } else {
        if (c->ds.n_tex) CRT_SHADE(true, false, false, false, false); else if (lean) CRT_SHADE(false, false, false, true, false); else CRT_SHADE(false, false, false, false, false);
      }
#undef CRT_SHADE
    }
    if (!fuse || depth == depth_max - 1) {
      SpanGuard g(c, F_CONNECT, s);
      if (pers) k_connect<COUNT, true, QUAD><<<g_con, CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, depth, gc);
      else k_connect<COUNT, false, false><<<grid_for(c, 16), CRT_TRACE_BLOCK, 0, s>>>(c->ds, st, depth, gc);
    }
  }
  return CRT_OK;
}

// One wave: n_batch samples of every pixel, or (adaptive != nullptr) `n_batch` tile samples dealt out by
// k_adaptive_allocate.  Split form: the frame's 8x4 pixel tiles are divided into `parts` ranges; every part is a
// complete small wave (its own path-state slots, queues and counters, its own k_resolve over its own pixels) on its own
// stream, so the latency-bound end of one part's launches overlaps the other parts' work.  Per-path results do not
// depend on the split (a path's stream is a function of its pixel and sample index only).
template <bool COUNT, bool QUAD>
int launch_batch(crt_context* c, uint32_t n_batch, const uint32_t* d_seeds, const AdaptiveState* adaptive = nullptr)
{
  const int depth_max = c->dp.max_depth;
  Counters* gc = c->d_counters.p;
  const uint32_t n_tiles = c->dp.tiles_x * c->dp.tiles_y;
  const uint64_t paths = (uint64_t)n_tiles * 32u * n_batch;
  DeviceParams base = c->dp;
  base.sample_group = 1;
  if (!adaptive)
    for (uint32_t G : { 32u, 16u, 8u, 4u })
      if ((int)G <= c->sample_group && n_batch % G == 0) { base.sample_group = G; break; }
  int parts = 1;
  if (!adaptive && c->side_stream[0] && (c->pipeline == 1 || (c->pipeline == 2 && paths <= c->pipeline_auto_paths && !c->timing_on)))
    parts = (int)std::min<uint32_t>((uint32_t)std::max(1, std::min(c->pipeline_parts, (int)crt_context::kMaxParts)), std::max(1u, n_tiles / 64u));
  CRT_CUDA(cudaMemsetAsync(c->counters.p, 0, sizeof(uint32_t) * ((size_t)(parts - 1) * kCounterWords + 4 * depth_max + 2), c->stream));
  c->last_parts.clear();
  if (parts == 1) {
    const PathState st = make_state(c, 0, 0);
    c->last_parts.push_back({ 0, 0 });
    enqueue_bounces<COUNT, QUAD>(c, st, base, c->stream, n_batch, d_seeds, adaptive, 0);
    SpanGuard g(c, F_RESOLVE);
    if (adaptive) k_resolve_adaptive<<<grid_for(c, 8), 256, 0, c->stream>>>(st, base, *adaptive, c->accum, COUNT ? gc : nullptr);
    else k_resolve<<<grid_for(c, 8), 256, 0, c->stream>>>(st, base, c->accum, n_batch, COUNT ? gc : nullptr);
  } else {
    CRT_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    size_t slot0 = 0;
    for (int k = 0; k < parts; ++k) {
      cudaStream_t s = k == 0 ? c->stream : c->side_stream[k - 1];
      DeviceParams dp = base;
      dp.tile0 = (uint32_t)((uint64_t)n_tiles * k / parts);
      dp.n_tiles = (uint32_t)((uint64_t)n_tiles * (k + 1) / parts) - dp.tile0;
      const PathState st = make_state(c, slot0, k);
      c->last_parts.push_back({ slot0, k });
      if (k > 0) CRT_CUDA(cudaStreamWaitEvent(s, c->ev_fork, 0));
      enqueue_bounces<COUNT, QUAD>(c, st, dp, s, n_batch, d_seeds, nullptr, c->pipeline_trace_ctas);
      {
        SpanGuard g(c, F_RESOLVE, s);
        k_resolve<<<grid_for(c, 4), 256, 0, s>>>(st, dp, c->accum, n_batch, COUNT ? gc : nullptr);   // the parts own disjoint pixels
      }
      if (k > 0) CRT_CUDA(cudaEventRecord(c->ev_join[k - 1], s));
      slot0 += (size_t)dp.n_tiles * 32u * n_batch;
    }
    for (int k = 1; k < parts; ++k) CRT_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join[k - 1], 0));
  }
  CRT_CUDA(cudaGetLastError());
  return CRT_OK;
}

int dispatch_batch(crt_context* c, uint32_t n, const uint32_t* d_seeds, const AdaptiveState* adaptive)
{
  if (c->quad) return c->stats_on ? launch_batch<true, true>(c, n, d_seeds, adaptive) : launch_batch<false, true>(c, n, d_seeds, adaptive);
  return c->stats_on ? launch_batch<true, false>(c, n, d_seeds, adaptive) : launch_batch<false, false>(c, n, d_seeds, adaptive);
}

// Adaptive screen sampling: n_samples x adaptive_tiles tile samples, in waves of at most
// samples-per-batch x (number of tiles); each wave is dealt out on the device from the error
// estimates the previous wave left (kernels.cuh).
// Adaptive screen sampling, host side.  adaptive_begin() sizes and (after a reset) clears the per-tile state and returns
// the device view of it; adaptive_wave() enqueues one wave of `budget` tile samples (frame seeds, allocation, the wave).
// A crt_group calls the same two steps on every member with (rank, n_members) set and runs the exchange step
// (k_adaptive_error_peers) between the waves; for one context the wave's own resolve pass updates the estimate.
struct AdaptivePlan { uint32_t nt; uint64_t per_unit, wave_cap; };

int adaptive_begin(crt_context* c, uint32_t n_samples, uint32_t rank, uint32_t n_members, AdaptiveState* A, AdaptivePlan* plan)
{
  const uint32_t ntx = (c->width + kAdaptiveTile - 1) / kAdaptiveTile, nty = (c->height + kAdaptiveTile - 1) / kAdaptiveTile;
  const uint32_t nt = ntx * nty;
  const uint64_t per_unit = c->params.adaptive_tiles > 0 ? (uint64_t)c->params.adaptive_tiles : nt;
  const uint64_t wave_cap = (uint64_t)auto_batch(c) * nt;
  // a member renders at most ceil(k / n) samples of a tile that receives k: one wave's share fits wave_cap + nt tile samples
  const uint64_t own_cap = n_members > 1 ? wave_cap + nt : std::min<uint64_t>(wave_cap, per_unit * n_samples);
  if (own_cap * kAdaptiveSlots > 0x1fffffffull) return fail(CRT_ERR_INVALID_ARG, "batch too large");
  int rc = ensure_path_slots(c, own_cap * kAdaptiveSlots);
  if (rc) return rc;
  CRT_CUDA(c->ad_count.ensure(nt)); CRT_CUDA(c->ad_err.ensure(nt));
  CRT_CUDA(c->ad_cum.ensure(nt + 1)); CRT_CUDA(c->ad_qoff.ensure(nt + 1));
  if (n_members > 1) CRT_CUDA(c->ad_cum_own.ensure(nt + 1));
  CRT_CUDA(c->ad_even.ensure((size_t)c->width * c->height));
  if (c->ad_clear) {
    CRT_CUDA(cudaMemsetAsync(c->ad_count.p, 0, sizeof(uint32_t) * nt, c->stream));
    CRT_CUDA(cudaMemsetAsync(c->ad_err.p, 0, sizeof(uint32_t) * nt, c->stream));
    CRT_CUDA(cudaMemsetAsync(c->ad_even.p, 0, sizeof(float) * (size_t)c->width * c->height, c->stream));
    c->ad_h_seeds.clear(); c->ad_seeds_uploaded = 0; c->ad_bound = 0; c->ad_wave = 0;
    c->ad_clear = false;
  }
  A->count = c->ad_count.p; A->err = c->ad_err.p; A->cum = c->ad_cum.p; A->qoff = c->ad_qoff.p; A->even = c->ad_even.p;
  A->ntx = ntx; A->nty = nty; A->nt = nt; A->first_parity = (uint32_t)(c->first_sample & 1u);
  A->cum_own = n_members > 1 ? c->ad_cum_own.p : c->ad_cum.p;
  A->rank = rank; A->n_members = n_members;
  plan->nt = nt; plan->per_unit = per_unit; plan->wave_cap = wave_cap;
  return CRT_OK;
}

int adaptive_wave(crt_context* c, const AdaptiveState& A, uint32_t budget)
{
  // no tile can receive more than 32 * budget / nt + 1 samples in a wave (weights are clamped to a 1:32 range)
  c->ad_bound += 32ull * budget / A.nt + 2;
  if (c->ad_bound > (1ull << 26)) return fail(CRT_ERR_INVALID_ARG, "adaptive accumulation too long; reset it");
  while (c->ad_h_seeds.size() < c->ad_bound) c->ad_h_seeds.push_back(frame_seed(c, c->first_sample + c->ad_h_seeds.size()));
  if (c->ad_seeds.n < c->ad_h_seeds.size()) {
    // the wave in flight may still read the old table
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    CRT_CUDA(c->ad_seeds.ensure(std::max<size_t>(2 * c->ad_h_seeds.size(), 4096)));
    c->ad_seeds_uploaded = 0;
  }
  CRT_CUDA(cudaMemcpyAsync(c->ad_seeds.p + c->ad_seeds_uploaded, c->ad_h_seeds.data() + c->ad_seeds_uploaded,
                           sizeof(uint32_t) * (c->ad_h_seeds.size() - c->ad_seeds_uploaded), cudaMemcpyHostToDevice, c->stream));
  c->ad_seeds_uploaded = c->ad_h_seeds.size();
  k_adaptive_allocate<<<1, 1024, 0, c->stream>>>(A, c->dp, budget, c->ad_wave);
  int rc = dispatch_batch(c, budget, c->ad_seeds.p, &A);
  if (rc) return rc;
  c->ad_wave++;
  return CRT_OK;
}

int render_adaptive(crt_context* c, uint32_t n_samples)
{
  AdaptiveState A;
  AdaptivePlan plan;
  int rc = adaptive_begin(c, n_samples, 0, 1, &A, &plan);
  if (rc) return rc;
  SpanGuard whole(c, F_RENDER);
  for (uint64_t left = plan.per_unit * n_samples; left > 0;) {
    const uint32_t budget = (uint32_t)std::min<uint64_t>(left, plan.wave_cap);
    if ((rc = adaptive_wave(c, A, budget))) return rc;
    left -= budget;
  }
  c->next_sample += n_samples;
  return CRT_OK;
}

// what every render call checks and refreshes before it enqueues waves
int render_prepare(crt_context* c)
{
  if (!c->has_layout || c->geometry_dirty) return fail(CRT_ERR_STATE, "crt_render before crt_commit");
  if (!c->width || !c->height) return fail(CRT_ERR_STATE, "crt_render before crt_resize");
  if (!c->accum) return fail(CRT_ERR_STATE, "no accumulation buffer");
  int rc = set_device(c);
  if (rc) return rc;
  if ((rc = upload_tables(c))) return rc;
  update_device_params(c);
  CRT_CUDA(c->counters.ensure(crt_context::kMaxParts * kCounterWords));
  return CRT_OK;
}

int render_impl(crt_context* c, uint32_t n_samples)
{
  int rc = render_prepare(c);
  if (rc) return rc;
  if (n_samples == 0) return CRT_OK;
  if (c->params.adaptive_sampling) return render_adaptive(c, n_samples);
  const uint32_t batch = std::min(auto_batch(c), n_samples);
  if ((rc = ensure_path_state(c, batch))) return rc;
  CRT_CUDA(c->seeds.ensure(n_samples));
  c->h_seeds.resize(n_samples);
  for (uint32_t k = 0; k < n_samples; ++k) c->h_seeds[k] = frame_seed(c, c->next_sample + k);
  CRT_CUDA(cudaMemcpyAsync(c->seeds.p, c->h_seeds.data(), sizeof(uint32_t) * n_samples, cudaMemcpyHostToDevice, c->stream));
  SpanGuard whole(c, F_RENDER);
  for (uint32_t done = 0; done < n_samples; done += batch) {
    const uint32_t nb = std::min(batch, n_samples - done);
    if ((rc = dispatch_batch(c, nb, c->seeds.p + done, nullptr))) return rc;
  }
  c->next_sample += n_samples;
  return CRT_OK;
}

void collect_spans(crt_context* c)
{
  for (TimedSpan& s : c->spans) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      c->family_ms[s.family] += ms;
      c->family_launches[s.family] += 1;
    }
    c->event_pool.push_back(s.a);
    c->event_pool.push_back(s.b);
  }
  c->spans.clear();
}

}  // namespace

// ------------------------------------------------------------------ C ABI

extern "C" {

const char* crt_last_error(void) { return g_error.c_str(); }
int crt_abi_version(void) { return CRT_ABI_VERSION; }

int crt_params_default(crt_params* p)
{
  CRT_REQUIRE(p, "null params");
  std::memset(p, 0, sizeof *p);
  p->max_depth = 8;             // CADRays' GI slider spans 1..32 (SettingsWidget.cxx:310-316)
  p->max_radiance = 50.0f;      // OCCT default RadianceClampingValue (SURVEY 5.6 range 1..1000)
  p->two_sided = 0;
  p->coherent_rng = 0;
  p->aperture_radius = 0.0f;
  p->focal_dist = 1.0f;
  p->tone_map = 0;
  p->white_point = 1.0f;
  p->exposure = 0.0f;
  p->env_as_background = 1;
  p->frame_seed0 = 1;
  p->russian_roulette = 1;
  p->samples_per_batch = 0;
  return CRT_OK;
}

int crt_create(int device_ordinal, crt_context** out)
{
  CRT_REQUIRE(out, "null out_ctx");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(CRT_ERR_NO_DEVICE, "no CUDA device available (libcadrays_b200 has no CPU fallback)");
  }
  if (device_ordinal < 0 || device_ordinal >= count) return fail(CRT_ERR_INVALID_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  CRT_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  if (prop.major != 10)
    return fail(CRT_ERR_NO_DEVICE, "device is not sm_100-class; this library carries sm_100a code only");
  crt_context* c = new (std::nothrow) crt_context();
  if (!c) return fail(CRT_ERR_OUT_OF_MEMORY, "host allocation failed");
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  if (const char* tv = std::getenv("CRT_TRAVERSAL")) c->persistent = std::string(tv) != "static";
  if (const char* tv = std::getenv("CRT_FUSE")) c->fuse_traversal = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_L2_PERSIST")) c->l2_persist = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_SHADE_LEAN")) c->shade_lean = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_SHADE_CLASSES")) c->shade_classes = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_SMEM_MATS")) c->smem_mats = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_SHADE_SORT")) c->shade_sort = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_PRIMARY_LOCKSTEP")) c->primary_lockstep = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_PRIMARY_GRID")) c->primary_grid = std::max(1, std::atoi(tv));
  if (const char* tv = std::getenv("CRT_FUSE_PRIMARY")) c->fuse_primary = std::atoi(tv) != 0;
  if (const char* tv = std::getenv("CRT_SAMPLE_GROUP")) c->sample_group = std::max(1, std::min(32, std::atoi(tv)));
  if (const char* tv = std::getenv("CRT_TAIL")) c->tail = std::atoi(tv) != 0;
  c->tail_max = (uint32_t)c->sm_count * CRT_TAIL_MIN_BLOCKS * 128u;
  if (const char* tv = std::getenv("CRT_TAIL_MAX")) c->tail_max = (uint32_t)std::max(0, std::atoi(tv));
  if (const char* tv = std::getenv("CRT_TAIL_MIN_DEPTH")) c->tail_min_depth = std::max(1, std::atoi(tv));
  if (const char* tv = std::getenv("CRT_PIPELINE")) c->pipeline = std::max(0, std::min(2, std::atoi(tv)));
  if (const char* tv = std::getenv("CRT_PIPELINE_PARTS")) c->pipeline_parts = std::max(1, std::min((int)crt_context::kMaxParts, std::atoi(tv)));
  if (const char* tv = std::getenv("CRT_PIPELINE_AUTO_PATHS")) c->pipeline_auto_paths = (uint64_t)std::max(0ll, std::atoll(tv));
  if (const char* tv = std::getenv("CRT_PIPELINE_TRACE_CTAS")) c->pipeline_trace_ctas = std::max(0, std::atoi(tv));

  crt_params_default(&c->params);
  std::memset(&c->cam, 0, sizeof c->cam);
  c->cam.dir[1] = 1.0f; c->cam.up[2] = 1.0f; c->cam.fovy_deg = 45.0f; c->cam.aspect = 1.0f; c->cam.ortho_scale = 1.0f;
  cudaError_t se = cudaSetDevice(device_ordinal);
  if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  for (int k = 0; k < crt_context::kMaxParts - 1; ++k) {
    if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&c->side_stream[k], cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming);
  }
  if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (se == cudaSuccess) se = c->d_counters.ensure(1);
  if (se == cudaSuccess) se = cudaMemset(c->d_counters.p, 0, sizeof(Counters));
  if (se != cudaSuccess) {
    std::string m = std::string("context setup: ") + cudaGetErrorString(se);
    delete c;
    return fail(CRT_ERR_CUDA, m);
  }
  *out = c;
  return CRT_OK;
}

int crt_create_host_only(crt_context** out)
{
  CRT_REQUIRE(out, "null out_ctx");
  crt_context* c = new (std::nothrow) crt_context();
  if (!c) return fail(CRT_ERR_OUT_OF_MEMORY, "host allocation failed");
  c->device = -1;
  crt_params_default(&c->params);
  std::memset(&c->cam, 0, sizeof c->cam);
  *out = c;
  return CRT_OK;
}

void crt_destroy(crt_context* c)
{
  if (!c) return;
  if (c->device < 0) { delete c; return; }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  collect_spans(c);
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  c->d_arena.release(); c->d_top_cache.release(); c->d_tri_uv.release(); c->d_tex.release(); c->d_tex_table.release();
  c->d_mats.release(); c->d_mat_class.release(); c->d_lights.release(); c->d_env.release();
  c->ray_o.release(); c->ray_d.release(); c->thr.release(); c->rad.release(); c->hit.release();
  c->sh_o.release(); c->sh_d.release(); c->sh_c.release(); c->hit_inst.release();
  c->queue0.release(); c->queue1.release(); c->counters.release(); c->seeds.release();
  c->accum_internal.release(); c->d_ldr.release(); c->d_hdr.release(); c->d_counters.release();
  c->ad_count.release(); c->ad_err.release(); c->ad_cum.release(); c->ad_qoff.release(); c->ad_seeds.release(); c->ad_even.release(); c->ad_cum_own.release();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int k = 0; k < crt_context::kMaxParts - 1; ++k) {
    if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    if (c->side_stream[k]) cudaStreamDestroy(c->side_stream[k]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int crt_mesh_create(crt_context* c, const float* pos, const float* nrm, const float* uv,
                    uint32_t n_verts, const uint32_t* idx, uint32_t n_tris, uint32_t* out_id)
{
  CRT_REQUIRE(c && pos && idx && out_id, "null argument");
  CRT_REQUIRE(n_verts > 0 && n_tris > 0, "empty mesh");
  for (size_t i = 0; i < (size_t)3 * n_tris; ++i)
    if (idx[i] >= n_verts) return fail(CRT_ERR_INVALID_ARG, "triangle index out of range");
  for (size_t i = 0; i < (size_t)3 * n_verts; ++i)
    if (!std::isfinite(pos[i])) return fail(CRT_ERR_INVALID_ARG, "non-finite vertex position");
  Mesh m;
  m.pos.assign(pos, pos + (size_t)3 * n_verts);
  m.idx.assign(idx, idx + (size_t)3 * n_tris);
  if (nrm) {
    m.nrm.assign(nrm, nrm + (size_t)3 * n_verts);
  } else {
    // area-weighted vertex normals from the faces (Assimp's aiProcess_GenNormals stands in
    // for this on the reference side, MeshImporter.cxx:73-106)
    std::vector<double> acc((size_t)3 * n_verts, 0.0);
    for (uint32_t t = 0; t < n_tris; ++t) {
      const float* a = &pos[3 * (size_t)idx[3 * t]];
      const float* b = &pos[3 * (size_t)idx[3 * t + 1]];
      const float* d = &pos[3 * (size_t)idx[3 * t + 2]];
      double e0[3] = { (double)b[0] - a[0], (double)b[1] - a[1], (double)b[2] - a[2] };
      double e1[3] = { (double)d[0] - a[0], (double)d[1] - a[1], (double)d[2] - a[2] };
      double n[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
      for (int k = 0; k < 3; ++k)
        for (int q = 0; q < 3; ++q) acc[3 * (size_t)idx[3 * t + k] + q] += n[q];
    }
    m.nrm.resize((size_t)3 * n_verts);
    for (uint32_t v = 0; v < n_verts; ++v) {
      double l = std::sqrt(acc[3 * v] * acc[3 * v] + acc[3 * v + 1] * acc[3 * v + 1] + acc[3 * v + 2] * acc[3 * v + 2]);
      if (l > 0.0) { for (int q = 0; q < 3; ++q) m.nrm[3 * (size_t)v + q] = (float)(acc[3 * v + q] / l); }
      else { m.nrm[3 * (size_t)v] = 0.0f; m.nrm[3 * (size_t)v + 1] = 0.0f; m.nrm[3 * (size_t)v + 2] = 1.0f; }
    }
  }
  if (uv) { m.uv.assign(uv, uv + (size_t)2 * n_verts); m.has_uv = true; }
  else m.uv.assign((size_t)2 * n_verts, 0.0f);
  c->scene.meshes.push_back(std::move(m));
  *out_id = (uint32_t)c->scene.meshes.size() - 1;
  return CRT_OK;
}

static const float k_identity[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };

int crt_instance_add(crt_context* c, uint32_t mesh_id, const float xf[12], uint32_t material_id, uint32_t* out_id)
{
  CRT_REQUIRE(c, "null context");
  CRT_REQUIRE(mesh_id < c->scene.meshes.size(), "unknown mesh id");
  Instance in;
  in.mesh = mesh_id; in.material = material_id;
  std::memcpy(in.xf, xf ? xf : k_identity, sizeof in.xf);
  for (int k = 0; k < 12; ++k) if (!std::isfinite(in.xf[k])) return fail(CRT_ERR_INVALID_ARG, "non-finite transform");
  c->scene.instances.push_back(in);
  c->geometry_dirty = true;
  if (out_id) *out_id = (uint32_t)c->scene.instances.size() - 1;
  return CRT_OK;
}

int crt_instance_set_transform(crt_context* c, uint32_t inst, const float xf[12])
{
  CRT_REQUIRE(c, "null context");
  CRT_REQUIRE(inst < c->scene.instances.size(), "unknown instance id");
  const float* src = xf ? xf : k_identity;
  for (int k = 0; k < 12; ++k) if (!std::isfinite(src[k])) return fail(CRT_ERR_INVALID_ARG, "non-finite transform");
  std::memcpy(c->scene.instances[inst].xf, src, sizeof(float) * 12);
  c->geometry_dirty = true;
  return CRT_OK;
}

int crt_instance_set_material(crt_context* c, uint32_t inst, uint32_t material_id)
{
  CRT_REQUIRE(c, "null context");
  CRT_REQUIRE(inst < c->scene.instances.size(), "unknown instance id");
  c->scene.instances[inst].material = material_id;
  c->geometry_dirty = true;
  return CRT_OK;
}

int crt_instance_set_visible(crt_context* c, uint32_t inst, int visible)
{
  CRT_REQUIRE(c, "null context");
  CRT_REQUIRE(inst < c->scene.instances.size(), "unknown instance id");
  if (c->scene.instances[inst].visible != (visible != 0)) {
    c->scene.instances[inst].visible = visible != 0;
    c->geometry_dirty = true;
  }
  return CRT_OK;
}

int crt_scene_clear(crt_context* c)
{
  CRT_REQUIRE(c, "null context");
  c->scene.meshes.clear();
  c->scene.instances.clear();
  c->scene.tree_cache.clear();
  c->scene.blob_signature.clear();      // new meshes may reuse ids and sizes: the next blob is written in full
  c->geometry_dirty = true;
  return CRT_OK;
}

int crt_materials_set(crt_context* c, const crt_bsdf* b, uint32_t n)
{
  CRT_REQUIRE(c && (b || n == 0), "null argument");
  c->mats.assign(b, b + n);
  c->mats_dirty = true;
  c->gen_mats++;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_texture_create(crt_context* c, const uint8_t* rgba8, uint32_t w, uint32_t h, uint32_t* out_id)
{
  CRT_REQUIRE(c && rgba8 && out_id, "null argument");
  CRT_REQUIRE(w > 0 && h > 0 && (uint64_t)w * h < (1ull << 28), "bad texture size");
  const size_t offset = c->tex_texels.size() / 4;
  CRT_REQUIRE(offset + (size_t)w * h < (1ull << 31), "texture storage exhausted");
  c->tex_texels.insert(c->tex_texels.end(), rgba8, rgba8 + (size_t)4 * w * h);
  c->tex_table.push_back((uint32_t)offset); c->tex_table.push_back(w); c->tex_table.push_back(h);
  *out_id = (uint32_t)(c->tex_table.size() / 3) - 1;
  c->tex_dirty = true;
  c->gen_tex++;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_textures_clear(crt_context* c)
{
  CRT_REQUIRE(c, "null context");
  c->tex_texels.clear(); c->tex_table.clear();
  c->tex_dirty = true;
  c->gen_tex++;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_lights_set(crt_context* c, const crt_light* l, uint32_t n)
{
  CRT_REQUIRE(c && (l || n == 0), "null argument");
  c->lights.assign((size_t)8 * n, 0.0f);
  for (uint32_t i = 0; i < n; ++i) {
    float* r = &c->lights[8 * (size_t)i];
    r[0] = l[i].emission[0]; r[1] = l[i].emission[1]; r[2] = l[i].emission[2];
    if (l[i].is_point) {
      r[3] = l[i].smoothness;
      r[4] = l[i].posdir[0]; r[5] = l[i].posdir[1]; r[6] = l[i].posdir[2]; r[7] = 1.0f;
    } else {
      r[3] = cosf(l[i].smoothness);
      v3 d = normalize3(V(l[i].posdir[0], l[i].posdir[1], l[i].posdir[2]));
      r[4] = -d.x; r[5] = -d.y; r[6] = -d.z; r[7] = 0.0f;
    }
  }
  c->lights_dirty = true;
  c->gen_lights++;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_envmap_set_rgb32f(crt_context* c, const float* rgb, uint32_t w, uint32_t h)
{
  CRT_REQUIRE(c, "null context");
  if (!rgb || !w || !h) { c->env.clear(); c->env_w = c->env_h = 0; }
  else {
    c->env.resize((size_t)4 * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      c->env[4 * i] = rgb[3 * i]; c->env[4 * i + 1] = rgb[3 * i + 1]; c->env[4 * i + 2] = rgb[3 * i + 2]; c->env[4 * i + 3] = 0.0f;
    }
    c->env_w = w; c->env_h = h;
  }
  c->env_dirty = true;
  c->gen_env++;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_envmap_set_rgb8(crt_context* c, const uint8_t* rgb, uint32_t w, uint32_t h)
{
  CRT_REQUIRE(c, "null context");
  if (!rgb || !w || !h) return crt_envmap_set_rgb32f(c, nullptr, 0, 0);
  std::vector<float> lin((size_t)3 * w * h);
  for (size_t i = 0; i < lin.size(); ++i) { float v = (float)rgb[i] * (1.0f / 255.0f); lin[i] = v * v; }
  return crt_envmap_set_rgb32f(c, lin.data(), w, h);
}

int crt_params_set(crt_context* c, const crt_params* p)
{
  CRT_REQUIRE(c && p, "null argument");
  CRT_REQUIRE(p->max_depth >= 1 && p->max_depth <= 64, "max_depth out of range");
  CRT_REQUIRE(p->bvh_width == 0 || p->bvh_width == 2 || p->bvh_width == 4, "bvh_width must be 0, 2 or 4");
  CRT_REQUIRE(p->adaptive_tiles >= 0, "adaptive_tiles must not be negative");
  if (p->frame_seed0 != c->params.frame_seed0) c->rng_valid = false;
  if ((p->bvh_width == 4) != (c->params.bvh_width == 4)) c->geometry_dirty = true;   // rebuilt at the next crt_commit
  c->params = *p;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_camera_set(crt_context* c, const crt_camera* cam)
{
  CRT_REQUIRE(c && cam, "null argument");
  c->cam = *cam;
  reset_accum_state(c);
  return CRT_OK;
}

int crt_resize(crt_context* c, uint32_t w, uint32_t h)
{
  CRT_REQUIRE(c, "null context");
  CRT_REQUIRE(w > 0 && h > 0 && (uint64_t)w * h < (1ull << 28), "bad size");
  int rc = set_device(c);
  if (rc) return rc;
  if (w != c->width || h != c->height) {
    if (c->accum_external) return fail(CRT_ERR_STATE, "unbind the external accumulation buffer before resizing");
    c->width = w; c->height = h;
    CRT_CUDA(c->accum_internal.ensure((size_t)w * h));
    c->accum = c->accum_internal.p;
    c->state_capacity = 0;
  }
  reset_accum_state(c);
  return CRT_OK;
}

int crt_commit(crt_context* c)
{
  CRT_REQUIRE(c, "null context");
  if (c->device < 0) {   // host-only: BVH build + blob, nothing to upload
    if (c->geometry_dirty) {
      std::string err;
      if (!build_blob(c->scene, c->blob, err, c->params.bvh_width == 4 ? 4 : 2)) return fail(CRT_ERR_INVALID_ARG, err);
      c->geometry_dirty = false;
    }
    return CRT_OK;
  }
  int rc = set_device(c);
  if (rc) return rc;
  if ((rc = commit_geometry(c, nullptr, nullptr))) return rc;
  return upload_tables(c);
}

int crt_render_async(crt_context* c, uint32_t n)
{
  CRT_REQUIRE(c, "null context");
  return render_impl(c, n);
}

int crt_sync(crt_context* c)
{
  CRT_REQUIRE(c, "null context");
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  return CRT_OK;
}

int crt_render(crt_context* c, uint32_t n, uint64_t* out_total)
{
  CRT_REQUIRE(c, "null context");
  int rc = render_impl(c, n);
  if (rc) return rc;
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  if (out_total) *out_total = c->next_sample - c->first_sample;
  return CRT_OK;
}

int crt_adaptive_tiles_get(crt_context* c, uint32_t* counts, uint32_t* errors, uint32_t capacity, uint32_t* out_tx, uint32_t* out_ty)
{
  CRT_REQUIRE(c, "null context");
  int rc = set_device(c);
  if (rc) return rc;
  const uint32_t ntx = (c->width + kAdaptiveTile - 1) / kAdaptiveTile, nty = (c->height + kAdaptiveTile - 1) / kAdaptiveTile;
  if (out_tx) *out_tx = ntx;
  if (out_ty) *out_ty = nty;
  if (!counts && !errors) return CRT_OK;
  CRT_REQUIRE(capacity >= ntx * nty, "capacity is smaller than the number of tiles");
  const bool live = !c->ad_clear && c->ad_count.p && c->ad_count.n >= (size_t)ntx * nty;
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  if (counts) {
    if (live) CRT_CUDA(cudaMemcpy(counts, c->ad_count.p, sizeof(uint32_t) * ntx * nty, cudaMemcpyDeviceToHost));
    else std::memset(counts, 0, sizeof(uint32_t) * ntx * nty);
  }
  if (errors) {
    if (live) CRT_CUDA(cudaMemcpy(errors, c->ad_err.p, sizeof(uint32_t) * ntx * nty, cudaMemcpyDeviceToHost));
    else std::memset(errors, 0, sizeof(uint32_t) * ntx * nty);
  }
  return CRT_OK;
}

int crt_set_next_sample(crt_context* c, uint64_t next_sample)
{
  CRT_REQUIRE(c, "null context");
  c->next_sample = next_sample;
  return CRT_OK;
}

int crt_reset_accumulation(crt_context* c, uint64_t first_sample)
{
  CRT_REQUIRE(c, "null context");
  int rc = set_device(c);
  if (rc) return rc;
  c->first_sample = first_sample;
  reset_accum_state(c);
  return CRT_OK;
}

static int display_impl(crt_context* c, const float4* accum, uint8_t* rgb8, size_t stride8, float* rgbf, size_t stridef)
{
  if (!c->width || !c->height) return fail(CRT_ERR_STATE, "no render target");
  int rc = set_device(c);
  if (rc) return rc;
  const uint32_t n = c->width * c->height;
  if (rgb8) CRT_CUDA(c->d_ldr.ensure((size_t)3 * n));
  if (rgbf) CRT_CUDA(c->d_hdr.ensure((size_t)3 * n));
  const float ex = exp2f(c->params.exposure);
  float wp = 1.0f;
  if (c->params.tone_map) {
    const float w = c->params.white_point;
    const float f = fmaf(1.425f, w, 0.05f);
    wp = (fmaf(w, f, 0.004f)) / (fmaf(w, f + 0.55f, 0.0491f)) - 0.0821f;
  }
  {
    SpanGuard g(c, F_RESOLVE);
    k_display<<<grid_for(c, 8), 256, 0, c->stream>>>(accum, n, ex, c->params.tone_map, wp,
                                                     rgb8 ? c->d_ldr.p : nullptr, rgbf ? c->d_hdr.p : nullptr);
  }
  CRT_CUDA(cudaGetLastError());
  if (rgb8) {
    const size_t row = (size_t)c->width * 3;
    if (stride8 == 0) stride8 = row;
    CRT_REQUIRE(stride8 >= row, "stride too small");
    CRT_CUDA(cudaMemcpy2DAsync(rgb8, stride8, c->d_ldr.p, row, row, c->height, cudaMemcpyDeviceToHost, c->stream));
  }
  if (rgbf) {
    const size_t row = (size_t)c->width * 12;
    if (stridef == 0) stridef = row;
    CRT_REQUIRE(stridef >= row, "stride too small");
    CRT_CUDA(cudaMemcpy2DAsync(rgbf, stridef, c->d_hdr.p, row, row, c->height, cudaMemcpyDeviceToHost, c->stream));
  }
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  return CRT_OK;
}

int crt_read_ldr(crt_context* c, uint8_t* rgb8, size_t stride)
{
  CRT_REQUIRE(c && rgb8, "null argument");
  if (!c->accum) return fail(CRT_ERR_STATE, "no accumulation buffer");
  return display_impl(c, c->accum, rgb8, stride, nullptr, 0);
}

int crt_read_hdr(crt_context* c, float* rgbf, size_t stride)
{
  CRT_REQUIRE(c && rgbf, "null argument");
  if (!c->accum) return fail(CRT_ERR_STATE, "no accumulation buffer");
  return display_impl(c, c->accum, nullptr, 0, rgbf, stride);
}

int crt_read_ldr_from(crt_context* c, const void* device_accum, uint8_t* rgb8, size_t stride)
{
  CRT_REQUIRE(c && device_accum && rgb8, "null argument");
  return display_impl(c, static_cast<const float4*>(device_accum), rgb8, stride, nullptr, 0);
}

int crt_accum_device_ptr(crt_context* c, void** out_ptr, size_t* out_bytes)
{
  CRT_REQUIRE(c && out_ptr, "null argument");
  *out_ptr = c->accum;
  if (out_bytes) *out_bytes = sizeof(float4) * (size_t)c->width * c->height;
  return CRT_OK;
}

int crt_accum_bind(crt_context* c, void* device_ptr, size_t bytes)
{
  CRT_REQUIRE(c, "null context");
  if (!device_ptr) {
    c->accum = c->accum_internal.p;
    c->accum_external = false;
  } else {
    CRT_REQUIRE(bytes >= sizeof(float4) * (size_t)c->width * c->height, "bound buffer too small");
    CRT_REQUIRE(((uintptr_t)device_ptr & 15u) == 0, "bound buffer must be 16-byte aligned");
    c->accum = static_cast<float4*>(device_ptr);
    c->accum_external = true;
  }
  reset_accum_state(c);
  return CRT_OK;
}

int crt_trace_device(crt_context* c, const void* org4, const void* dir4, uint32_t n, int any_hit, void* hit4, void* inst)
{
  CRT_REQUIRE(c && org4 && dir4 && hit4, "null argument");
  if (!c->has_layout || c->geometry_dirty) return fail(CRT_ERR_STATE, "crt_trace before crt_commit");
  int rc = set_device(c);
  if (rc) return rc;
  if (n == 0) return CRT_OK;
  const float4* o = static_cast<const float4*>(org4);
  const float4* d = static_cast<const float4*>(dir4);
  float4* h = static_cast<float4*>(hit4);
  int32_t* hi = static_cast<int32_t*>(inst);
  Counters* gc = c->d_counters.p;
  CRT_CUDA(c->counters.ensure(4 * 64 + 8));
  uint32_t* work = c->counters.p + 4 * 64 + 4;
  CRT_CUDA(cudaMemsetAsync(work, 0, sizeof(uint32_t), c->stream));
  const bool pers = c->persistent;
  const int sgrid = std::min<int>(grid_for(c, 16), (int)((n + 127u) / 128u));
  SpanGuard g(c, any_hit ? F_CONNECT : F_EXTEND);
#define CRT_LAUNCH_TRACE(ANY, CNT)                                                                                   \
  do {                                                                                                               \
    if (c->quad) {                                                                                                   \
      const int pg = std::min<int>(resident_grid(c, k_trace<ANY, CNT, true, true>, CRT_TRACE_BLOCK), (int)((n + 31u) / 32u)); \
      k_trace<ANY, CNT, true, true><<<pg, CRT_TRACE_BLOCK, 0, c->stream>>>(c->ds, o, d, n, h, hi, work, gc);         \
    } else if (pers) {                                                                                               \
      const int pg = std::min<int>(resident_grid(c, k_trace<ANY, CNT, true, false>, CRT_TRACE_BLOCK), (int)((n + 31u) / 32u)); \
      k_trace<ANY, CNT, true, false><<<pg, CRT_TRACE_BLOCK, 0, c->stream>>>(c->ds, o, d, n, h, hi, work, gc);        \
    } else {                                                                                                         \
      k_trace<ANY, CNT, false, false><<<sgrid, CRT_TRACE_BLOCK, 0, c->stream>>>(c->ds, o, d, n, h, hi, work, gc);    \
    }                                                                                                                \
  } while (0)
  if (any_hit) { if (c->stats_on) CRT_LAUNCH_TRACE(true, true); else CRT_LAUNCH_TRACE(true, false); }
  else { if (c->stats_on) CRT_LAUNCH_TRACE(false, true); else CRT_LAUNCH_TRACE(false, false); }
#undef CRT_LAUNCH_TRACE
  CRT_CUDA(cudaGetLastError());
  return CRT_OK;
}

int crt_trace(crt_context* c, const float* org, const float* dir, const float* tmax, uint32_t n, int any_hit,
              int32_t* prim, int32_t* inst, float* t, float* u, float* v)
{
  CRT_REQUIRE(c && (n == 0 || (org && dir)), "null argument");
  if (!c->has_layout || c->geometry_dirty) return fail(CRT_ERR_STATE, "crt_trace before crt_commit");
  if (n == 0) return CRT_OK;
  int rc = set_device(c);
  if (rc) return rc;
  std::vector<float4> ho(n), hd(n);
  for (uint32_t i = 0; i < n; ++i) {
    ho[i] = make_float4(org[3 * (size_t)i], org[3 * (size_t)i + 1], org[3 * (size_t)i + 2], 0.0f);
    hd[i] = make_float4(dir[3 * (size_t)i], dir[3 * (size_t)i + 1], dir[3 * (size_t)i + 2], tmax ? tmax[i] : CRT_MAXFLOAT);
  }
  DevBuf<float4> d_o, d_d, d_h;
  DevBuf<int32_t> d_i;
  cudaError_t e = d_o.ensure(n);
  if (e == cudaSuccess) e = d_d.ensure(n);
  if (e == cudaSuccess) e = d_h.ensure(n);
  if (e == cudaSuccess) e = d_i.ensure(n);
  auto cleanup = [&]() { d_o.release(); d_d.release(); d_h.release(); d_i.release(); };
  if (e != cudaSuccess) { cleanup(); cudaGetLastError(); return fail(CRT_ERR_OUT_OF_MEMORY, cudaGetErrorString(e)); }
  std::vector<float4> hh(n);
  std::vector<int32_t> hi(n);
  e = cudaMemcpyAsync(d_o.p, ho.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_d.p, hd.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    rc = crt_trace_device(c, d_o.p, d_d.p, n, any_hit, d_h.p, d_i.p);
    if (rc) { cleanup(); return rc; }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(hh.data(), d_h.p, sizeof(float4) * n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hi.data(), d_i.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cleanup();
  if (e != cudaSuccess) { cudaGetLastError(); return fail(CRT_ERR_CUDA, cudaGetErrorString(e)); }
  for (uint32_t i = 0; i < n; ++i) {
    int32_t p;
    std::memcpy(&p, &hh[i].w, 4);
    if (prim) prim[i] = p;
    if (inst) inst[i] = (any_hit && p < 0) ? -1 : hi[i];
    if (t) t[i] = hh[i].x;
    if (u) u[i] = hh[i].y;
    if (v) v[i] = hh[i].z;
  }
  return CRT_OK;
}

int crt_wavefront_rays(crt_context* c, int depth, int kind, float* org, float* dir, float* tmax, uint32_t capacity, uint32_t* out_n)
{
  CRT_REQUIRE(c && out_n, "null argument");
  CRT_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (continuation rays) or 1 (shadow rays)");
  int rc = set_device(c);
  if (rc) return rc;
  if (!c->state_capacity || !c->counters.p) return fail(CRT_ERR_STATE, "crt_wavefront_rays before crt_render");
  CRT_REQUIRE(depth >= 0 && depth < c->dp.max_depth, "depth outside the last wave");
  CRT_REQUIRE(kind == 1 || depth >= 1, "camera rays are not stored (they are recomputed from the pixel and the frame seed)");
  if (c->params.adaptive_sampling) return fail(CRT_ERR_STATE, "crt_wavefront_rays: not available with adaptive sampling");
  if (c->last_parts.empty()) return fail(CRT_ERR_STATE, "crt_wavefront_rays before crt_render");
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  // a wave may have been split into parts (tile ranges on several streams): their rays are returned part after part
  std::vector<uint32_t> counts(c->last_parts.size(), 0);
  uint64_t total = 0;
  for (size_t k = 0; k < c->last_parts.size(); ++k) {
    const PathState st = make_state(c, c->last_parts[k].slot0, c->last_parts[k].index);
    CRT_CUDA(cudaMemcpy(&counts[k], (kind ? st.n_shadow : st.n_active) + depth, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    total += counts[k];
  }
  *out_n = (uint32_t)std::min<uint64_t>(total, 0xffffffffull);
  if (total == 0 || capacity == 0 || (!org && !dir && !tmax)) return CRT_OK;
  uint32_t written = 0;
  for (size_t k = 0; k < c->last_parts.size() && written < capacity; ++k) {
    const uint32_t m = std::min(counts[k], capacity - written);
    if (m == 0) continue;
    const PathState st = make_state(c, c->last_parts[k].slot0, c->last_parts[k].index);
    DevBuf<float4> d_o, d_d;
    cudaError_t e = d_o.ensure(m);
    if (e == cudaSuccess) e = d_d.ensure(m);
    std::vector<float4> ho(m), hd(m);
    if (e == cudaSuccess) {
      k_gather_rays<<<grid_for(c, 8), 256, 0, c->stream>>>(st, st.queue[depth & 1], m, kind, d_o.p, d_d.p);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ho.data(), d_o.p, sizeof(float4) * m, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hd.data(), d_d.p, sizeof(float4) * m, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    d_o.release(); d_d.release();
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? CRT_ERR_OUT_OF_MEMORY : CRT_ERR_CUDA, cudaGetErrorString(e)); }
    for (uint32_t i = 0; i < m; ++i) {
      const size_t o = (size_t)written + i;
      if (org) { org[3 * o] = ho[i].x; org[3 * o + 1] = ho[i].y; org[3 * o + 2] = ho[i].z; }
      if (dir) { dir[3 * o] = hd[i].x; dir[3 * o + 1] = hd[i].y; dir[3 * o + 2] = hd[i].z; }
      if (tmax) tmax[o] = hd[i].w;
    }
    written += m;
  }
  return CRT_OK;
}

int crt_bvh_export(crt_context* c, void* buf, size_t capacity, size_t* out_size)
{
  CRT_REQUIRE(c, "null context");
  if (c->geometry_dirty || c->blob.empty()) return fail(CRT_ERR_STATE, "crt_bvh_export before crt_commit");
  if (out_size) *out_size = c->blob.size();
  if (!buf) return CRT_OK;
  CRT_REQUIRE(capacity >= c->blob.size(), "buffer too small");
  std::memcpy(buf, c->blob.data(), c->blob.size());
  return CRT_OK;
}

int crt_bvh_import(crt_context* c, const void* buf, size_t size)
{
  CRT_REQUIRE(c && buf, "null argument");
  int rc = set_device(c);
  if (rc) return rc;
  std::vector<uint8_t> keep(static_cast<const uint8_t*>(buf), static_cast<const uint8_t*>(buf) + size);
  c->blob.swap(keep);
  c->scene.blob_signature.clear();      // the blob no longer comes from this context's scene
  c->blob_on_device = false;
  rc = load_blob(c);
  if (rc) {
    c->blob.swap(keep);
    // the device may have lost its layout half way (upload_layout clears has_layout): the next crt_commit rebuilds it
    if (!c->has_layout) c->geometry_dirty = true;
    return rc;
  }
  c->geometry_dirty = false;
  c->blob_on_device = false;            // an imported blob is not the scene's: the next crt_commit converts in full
  reset_accum_state(c);
  return CRT_OK;
}

int crt_stats_enable(crt_context* c, int on) { CRT_REQUIRE(c, "null context"); c->stats_on = on != 0; return CRT_OK; }

int crt_stats_reset(crt_context* c)
{
  if (c) c->kernel_launches = 0;
  CRT_REQUIRE(c, "null context");
  int rc = set_device(c);
  if (rc) return rc;
  CRT_CUDA(cudaMemsetAsync(c->d_counters.p, 0, sizeof(Counters), c->stream));
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  collect_spans(c);
  for (int k = 0; k < F_COUNT; ++k) { c->family_ms[k] = 0.0; c->family_launches[k] = 0; }
  return CRT_OK;
}

int crt_stats_get(crt_context* c, crt_stats* out)
{
  CRT_REQUIRE(c && out, "null argument");
  int rc = set_device(c);
  if (rc) return rc;
  static_assert(sizeof(Counters) == sizeof(crt_stats), "counter layout");
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  CRT_CUDA(cudaMemcpy(out, c->d_counters.p, sizeof(Counters), cudaMemcpyDeviceToHost));
  return CRT_OK;
}

int crt_launch_count(crt_context* c, uint64_t* out)
{
  CRT_REQUIRE(c && out, "null argument");
  *out = c->kernel_launches;
  return CRT_OK;
}

int crt_timing_enable(crt_context* c, int on) { CRT_REQUIRE(c, "null context"); c->timing_on = on != 0; return CRT_OK; }

int crt_timing_get(crt_context* c, double ms[6], uint64_t launches[6])
{
  CRT_REQUIRE(c && ms, "null argument");
  int rc = set_device(c);
  if (rc) return rc;
  CRT_CUDA(cudaStreamSynchronize(c->stream));
  collect_spans(c);
  for (int k = 0; k < F_COUNT; ++k) { ms[k] = c->family_ms[k]; if (launches) launches[k] = c->family_launches[k]; }
  return CRT_OK;
}

int crt_scene_bytes(crt_context* c, size_t* out_traversal, size_t* out_total)
{
  CRT_REQUIRE(c, "null context");
  if (!c->has_layout || c->geometry_dirty) return fail(CRT_ERR_STATE, "crt_scene_bytes before crt_commit");
  const size_t trav = c->scene_traversal_bytes;
  if (out_traversal) *out_traversal = trav;
  if (out_total)
    *out_total = c->scene_total_bytes + c->mats.size() * sizeof(crt_bsdf) + c->lights.size() * 4 + c->env.size() * 4 +
                 c->tex_texels.size() + c->tex_table.size() * 4;
  return CRT_OK;
}

int crt_commit_stats(crt_context* c, uint64_t* out_top_level_patches)
{
  CRT_REQUIRE(c, "null context");
  if (out_top_level_patches) *out_top_level_patches = c->top_patches;
  return CRT_OK;
}

int crt_stream(crt_context* c, void** out_stream)
{
  CRT_REQUIRE(c && out_stream, "null argument");
  *out_stream = c->stream;
  return CRT_OK;
}

}  // extern "C"

#include "group.inl"
