// group.inl -- crt_group: one V3d_View::Redraw() driven over several GPUs of one box from ONE host process
// (included at the end of api.cu; needs crt_context and the helpers defined there).
//
// CADRays is a single-process C++ application; its Redraw (src/Launcher/AppViewer.cxx:1047) and BufferDump
// (AppViewer.cxx:1259-1262) must be able to use every GPU of the machine without the host growing a launcher or a
// message-passing layer.  A group wraps the caller's context (member 0, which keeps the host scene) and one
// replica context per further device:
//   crt_group_commit  builds the BVH once on the host, converts it to the device layout once and uploads that
//                     layout to every member (in parallel, one host thread per device); materials, lights,
//                     environment, textures, rendering parameters, camera and size are replicated when they changed;
//   crt_group_render  deals the next n sample indices out in contiguous blocks (the arithmetic of
//                     cadrays_b200/distributed.py sample_range): sample s of pixel p is the same random stream on
//                     whichever GPU renders it, so the union over the members equals the 1-GPU sample set;
//   crt_group_read_*  is the exchange step.  Default: ONE fused kernel per member (k_display_peers) that reads its
//                     rows of every member's accumulation buffer through NVLink peer addresses, adds them in member
//                     order and tone-maps -- reduce-scatter + Display + the copy to the host row block, no
//                     intermediate sum buffer.  CRT_GROUP_REDUCE=nccl (or members without peer access) selects the
//                     library path instead: ncclReduce of the float4 sums to member 0, then the ordinary Display
//                     pass there.  NCCL is bound at run time with dlopen("libnccl.so.2"), so the library has no link
//                     dependency on it and shares the copy a host process (e.g. PyTorch) already loaded.
#include <dlfcn.h>
#include <omp.h>

namespace {

// the few NCCL entry points used, resolved from libnccl.so.2 at run time (types as in nccl.h 2.x)
struct NcclApi {
  void* lib = nullptr;
  typedef struct ncclComm* comm_t;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err)
  {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("loading libnccl.so.2: ") + dlerror(); return false; }
#define CRT_NCCL_SYM(field, name)                                                         \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                          \
    if (!field) { err = std::string("libnccl.so.2 lacks ") + name; dlclose(lib); lib = nullptr; return false; }
    CRT_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    CRT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    CRT_NCCL_SYM(GroupStart, "ncclGroupStart")
    CRT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    CRT_NCCL_SYM(Reduce, "ncclReduce")
    CRT_NCCL_SYM(AllReduce, "ncclAllReduce")
    CRT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CRT_NCCL_SYM
    return true;
  }
};
constexpr int kNcclFloat = 7, kNcclSum = 0;     // ncclFloat32, ncclSum

}  // namespace

struct crt_group {
  std::vector<crt_context*> members;     // [0] = the caller's context, the rest are replicas owned by the group
  bool peer_ok = true;                   // every member can address every other member's memory
  bool use_nccl = false;
  NcclApi nccl;
  std::vector<NcclApi::comm_t> comms;
  DevBuf<float4> nccl_sum;               // on member 0
  // what of member 0 has been replicated so far
  uint64_t seen_scene = ~0ull, seen_mats = ~0ull, seen_lights = ~0ull, seen_env = ~0ull, seen_tex = ~0ull, seen_accum = ~0ull;
  crt_params seen_params;
  crt_camera seen_cam;
  uint32_t seen_w = 0, seen_h = 0;
  uint64_t first_sample = 0, next_sample = 0;   // the group's sample cursor
  std::vector<cudaEvent_t> done;         // per member: its share of the last crt_group_render has finished
  std::vector<cudaEvent_t> ev_wave, ev_estimate;   // adaptive sampling: the member's wave is resolved / its tiles' estimates are written
  double last_reduce_ms = 0.0;           // device time of the last exchange + Display pass (slowest member)
  std::vector<cudaEvent_t> t0, t1;
};

namespace {

// runs fn(rank) for every member on its own host thread (CUDA launches and copies to different devices overlap);
// the first failing rank's status and message are re-raised on the calling thread
template <typename F>
int for_each_member(crt_group* g, F fn)
{
  const int n = (int)g->members.size();
  std::vector<int> rc(n, CRT_OK);
  std::vector<std::string> msg(n);
#pragma omp parallel for num_threads(n) schedule(static, 1) if (n > 1)
  for (int r = 0; r < n; ++r) {
    rc[r] = fn(r);
    if (rc[r]) msg[r] = g_error;
  }
  for (int r = 0; r < n; ++r)
    if (rc[r]) return fail(rc[r], "group member " + std::to_string(r) + ": " + msg[r]);
  return CRT_OK;
}

// materials / lights / environment / textures / parameters / camera / size of member 0 -> replicas
int replicate_state(crt_group* g)
{
  crt_context* c0 = g->members[0];
  const bool mats = g->seen_mats != c0->gen_mats, lights = g->seen_lights != c0->gen_lights;
  const bool env = g->seen_env != c0->gen_env, tex = g->seen_tex != c0->gen_tex;
  const bool params = std::memcmp(&g->seen_params, &c0->params, sizeof(crt_params)) != 0;
  const bool cam = std::memcmp(&g->seen_cam, &c0->cam, sizeof(crt_camera)) != 0;
  const bool size = g->seen_w != c0->width || g->seen_h != c0->height;
  const bool accum = g->seen_accum != c0->gen_accum;
  if (!(mats || lights || env || tex || params || cam || size || accum)) return CRT_OK;
  const int rc = for_each_member(g, [&](int r) -> int {
    if (r == 0) return CRT_OK;
    crt_context* c = g->members[r];
    int e = set_device(c);
    if (e) return e;
    if (mats) { c->mats = c0->mats; c->mats_dirty = true; }
    if (lights) { c->lights = c0->lights; c->lights_dirty = true; }
    if (env) { c->env = c0->env; c->env_w = c0->env_w; c->env_h = c0->env_h; c->env_dirty = true; }
    if (tex) { c->tex_texels = c0->tex_texels; c->tex_table = c0->tex_table; c->tex_dirty = true; }
    if (params) {
      if (c->params.frame_seed0 != c0->params.frame_seed0) c->rng_valid = false;
      c->params = c0->params;
    }
    if (cam) c->cam = c0->cam;
    if (size && c0->width && c0->height && (e = crt_resize(c, c0->width, c0->height))) return e;
    // any of these restarts the accumulation on member 0 (reset_accum_state); the replicas follow
    c->first_sample = c0->first_sample;
    reset_accum_state(c);
    return CRT_OK;
  });
  if (rc) return rc;
  g->seen_mats = c0->gen_mats; g->seen_lights = c0->gen_lights; g->seen_env = c0->gen_env; g->seen_tex = c0->gen_tex;
  g->seen_params = c0->params; g->seen_cam = c0->cam; g->seen_w = c0->width; g->seen_h = c0->height;
  g->seen_accum = c0->gen_accum;
  g->first_sample = g->next_sample = c0->first_sample;
  return CRT_OK;
}

int group_display(crt_group* g, uint8_t* rgb8, size_t stride8, float* rgbf, size_t stridef)
{
  crt_context* c0 = g->members[0];
  if (!c0->width || !c0->height) return fail(CRT_ERR_STATE, "no render target");
  const uint32_t W = c0->width, H = c0->height;
  const int n = (int)g->members.size();
  const size_t row8 = (size_t)W * 3, rowf = (size_t)W * 12;
  if (rgb8) { if (stride8 == 0) stride8 = row8; CRT_REQUIRE(stride8 >= row8, "stride too small"); }
  if (rgbf) { if (stridef == 0) stridef = rowf; CRT_REQUIRE(stridef >= rowf, "stride too small"); }
  const float ex = exp2f(c0->params.exposure);
  float wp = 1.0f;
  if (c0->params.tone_map) {
    const float w = c0->params.white_point;
    const float f = fmaf(1.425f, w, 0.05f);
    wp = (fmaf(w, f, 0.004f)) / (fmaf(w, f + 0.55f, 0.0491f)) - 0.0821f;
  }
  if (g->use_nccl && n > 1) {
    // library path: ncclReduce of the sums to member 0, Display there
    int rc = set_device(c0);
    if (rc) return rc;
    CRT_CUDA(g->nccl_sum.ensure((size_t)W * H));
    int e = g->nccl.GroupStart();
    for (int r = 0; r < n && e == 0; ++r) {
      cudaSetDevice(g->members[r]->device);
      e = g->nccl.Reduce(g->members[r]->accum, r == 0 ? g->nccl_sum.p : nullptr, (size_t)W * H * 4, kNcclFloat, kNcclSum, 0,
                         g->comms[r], g->members[r]->stream);
    }
    const int e2 = g->nccl.GroupEnd();
    if (e || e2) return fail(CRT_ERR_CUDA, std::string("ncclReduce: ") + g->nccl.GetErrorString(e ? e : e2));
    for (int r = 1; r < n; ++r) { cudaSetDevice(g->members[r]->device); CRT_CUDA(cudaStreamSynchronize(g->members[r]->stream)); }
    return display_impl(c0, g->nccl_sum.p, rgb8, stride8, rgbf, stridef);
  }
  // fused path: member r reduces + tone-maps rows [r H / n, (r + 1) H / n) and copies them to the host
  PeerAccums A;
  A.n = n;
  for (int r = 0; r < n; ++r) A.p[r] = g->members[r]->accum;
  std::vector<float> ms(n, 0.0f);
  const int rc = for_each_member(g, [&](int r) -> int {
    crt_context* c = g->members[r];
    int e = set_device(c);
    if (e) return e;
    const uint32_t y0 = (uint32_t)((uint64_t)H * r / n), y1 = (uint32_t)((uint64_t)H * (r + 1) / n);
    if (y1 == y0) return CRT_OK;
    if (rgb8) CRT_CUDA(c->d_ldr.ensure((size_t)3 * W * H));
    if (rgbf) CRT_CUDA(c->d_hdr.ensure((size_t)3 * W * H));
    CRT_CUDA(cudaEventRecord(g->t0[r], c->stream));
    {
      SpanGuard sg(c, F_RESOLVE);
      k_display_peers<<<grid_for(c, 8), 256, 0, c->stream>>>(A, y0 * W, (y1 - y0) * W, ex, c0->params.tone_map, wp,
                                                             rgb8 ? c->d_ldr.p : nullptr, rgbf ? c->d_hdr.p : nullptr, nullptr);
    }
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaEventRecord(g->t1[r], c->stream));
    if (rgb8)
      CRT_CUDA(cudaMemcpy2DAsync(rgb8 + (size_t)y0 * stride8, stride8, c->d_ldr.p + (size_t)y0 * row8, row8, row8, y1 - y0,
                                 cudaMemcpyDeviceToHost, c->stream));
    if (rgbf)
      CRT_CUDA(cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(rgbf) + (size_t)y0 * stridef, stridef,
                                 reinterpret_cast<uint8_t*>(c->d_hdr.p) + (size_t)y0 * rowf, rowf, rowf, y1 - y0,
                                 cudaMemcpyDeviceToHost, c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&ms[r], g->t0[r], g->t1[r]);
    return CRT_OK;
  });
  if (rc) return rc;
  g->last_reduce_ms = *std::max_element(ms.begin(), ms.end());
  return CRT_OK;
}

// Adaptive screen sampling over the members (AdaptiveScreenSampling with several GPUs behind one Redraw).  Every
// member holds the same per-tile state and runs the same allocation; of a tile's new samples -- global sample indices
// count .. count + k - 1 of every pixel of the tile -- member r renders the ones congruent to r modulo N, so the union
// is the sample set one GPU would render for the same allocation.  After each wave the members exchange: member r
// rebuilds the error estimate and count of the tiles j = r (mod N) from all members' sums over peer addresses and
// writes them into every member's arrays (k_adaptive_error_peers); the next wave's allocation waits for that.  All
// waves of a call are enqueued from this one thread without a host synchronisation, ordered by events.
int group_render_adaptive(crt_group* g, uint32_t n_samples)
{
  const int n = (int)g->members.size();
  if (!g->peer_ok) return fail(CRT_ERR_STATE, "adaptive screen sampling over a group needs peer access between its GPUs");
  std::vector<AdaptiveState> A(n);
  AdaptivePlan plan{};
  int rc = CRT_OK;
  for (int r = 0; r < n; ++r) {
    crt_context* c = g->members[r];
    if ((rc = render_prepare(c))) return rc;
    if ((rc = adaptive_begin(c, n_samples, (uint32_t)r, (uint32_t)n, &A[r], &plan))) return rc;
  }
  PeerAdaptive G;
  G.n = n;
  for (int r = 0; r < n; ++r) {
    G.accum[r] = g->members[r]->accum; G.even[r] = g->members[r]->ad_even.p;
    G.err[r] = g->members[r]->ad_err.p; G.count[r] = g->members[r]->ad_count.p;
  }
  const uint64_t cap = plan.wave_cap * (uint64_t)n;
  for (uint64_t left = plan.per_unit * n_samples; left > 0;) {
    const uint32_t budget = (uint32_t)std::min<uint64_t>(left, cap);
    for (int r = 0; r < n; ++r) {            // the wave: allocation (from the estimates every member wrote), trace, resolve
      crt_context* c = g->members[r];
      if ((rc = set_device(c))) return rc;
      for (int m = 0; m < n; ++m) CRT_CUDA(cudaStreamWaitEvent(c->stream, g->ev_estimate[m], 0));   // no-op before the first record
      if ((rc = adaptive_wave(c, A[r], budget))) return rc;
      CRT_CUDA(cudaEventRecord(g->ev_wave[r], c->stream));
    }
    for (int r = 0; r < n; ++r) {            // the exchange: estimates of the tiles j = r (mod n) from all members' sums
      crt_context* c = g->members[r];
      if ((rc = set_device(c))) return rc;
      for (int m = 0; m < n; ++m) CRT_CUDA(cudaStreamWaitEvent(c->stream, g->ev_wave[m], 0));
      {
        SpanGuard sg(c, F_RESOLVE);
        k_adaptive_error_peers<<<grid_for(c, 4), 256, 0, c->stream>>>(G, c->dp, A[r]);
      }
      CRT_CUDA(cudaGetLastError());
      CRT_CUDA(cudaEventRecord(g->ev_estimate[r], c->stream));
    }
    left -= budget;
  }
  // a later call on any member (reads, the next Redraw) must see every member's estimates
  for (int r = 0; r < n; ++r) {
    crt_context* c = g->members[r];
    if ((rc = set_device(c))) return rc;
    for (int m = 0; m < n; ++m) CRT_CUDA(cudaStreamWaitEvent(c->stream, g->ev_estimate[m], 0));
    c->next_sample += n_samples;
  }
  return CRT_OK;
}

}  // namespace

extern "C" {

int crt_group_create(crt_context* primary, const int* devices, int n_devices, crt_group** out)
{
  CRT_REQUIRE(primary && devices && out, "null argument");
  *out = nullptr;
  CRT_REQUIRE(n_devices >= 1 && n_devices <= kMaxGroup, "a group has 1..16 members");
  CRT_REQUIRE(primary->device >= 0, "the primary context is host-only");
  CRT_REQUIRE(devices[0] == primary->device, "devices[0] must be the primary context's device");
  CRT_REQUIRE(!primary->accum_external, "unbind the external accumulation buffer first");
  crt_group* g = new (std::nothrow) crt_group();
  if (!g) return fail(CRT_ERR_OUT_OF_MEMORY, "host allocation failed");
  g->members.push_back(primary);
  std::memset(&g->seen_params, 0xff, sizeof g->seen_params);
  std::memset(&g->seen_cam, 0xff, sizeof g->seen_cam);
  for (int r = 1; r < n_devices; ++r) {
    crt_context* c = nullptr;
    const int rc = crt_create(devices[r], &c);
    if (rc) { const std::string m = g_error; crt_group_destroy(g); return fail(rc, "group member " + std::to_string(r) + ": " + m); }
    // A/B switches follow the primary (they were read from the environment by both, but a caller may have changed them)
    g->members.push_back(c);
  }
  // peer access between every pair of distinct devices
  for (int a = 0; a < n_devices && g->peer_ok; ++a)
    for (int b = 0; b < n_devices; ++b) {
      const int da = g->members[a]->device, db = g->members[b]->device;
      if (da == db) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, da, db) != cudaSuccess || !can) { cudaGetLastError(); g->peer_ok = false; break; }
      cudaSetDevice(da);
      const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); g->peer_ok = false; break; }
      cudaGetLastError();
    }
  const char* mode = std::getenv("CRT_GROUP_REDUCE");
  g->use_nccl = n_devices > 1 && (!g->peer_ok || (mode && std::string(mode) == "nccl"));
  if (g->use_nccl) {
    // NCCL needs distinct devices per rank
    for (int a = 0; a < n_devices; ++a)
      for (int b = a + 1; b < n_devices; ++b)
        if (devices[a] == devices[b]) { crt_group_destroy(g); return fail(CRT_ERR_INVALID_ARG, "the NCCL path needs distinct devices"); }
    std::string err;
    if (!g->nccl.load(err)) { crt_group_destroy(g); return fail(CRT_ERR_NO_DEVICE, err); }
    g->comms.assign(n_devices, nullptr);
    const int e = g->nccl.CommInitAll(g->comms.data(), n_devices, devices);
    if (e) { g->comms.clear(); const std::string m = g->nccl.GetErrorString(e); crt_group_destroy(g); return fail(CRT_ERR_CUDA, "ncclCommInitAll: " + m); }
  }
  g->done.assign(n_devices, nullptr); g->t0.assign(n_devices, nullptr); g->t1.assign(n_devices, nullptr);
  g->ev_wave.assign(n_devices, nullptr); g->ev_estimate.assign(n_devices, nullptr);
  for (int r = 0; r < n_devices; ++r) {
    cudaSetDevice(g->members[r]->device);
    cudaEventCreateWithFlags(&g->done[r], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_wave[r], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&g->ev_estimate[r], cudaEventDisableTiming);
    cudaEventCreate(&g->t0[r]);
    cudaEventCreate(&g->t1[r]);
  }
  cudaSetDevice(primary->device);
  g->first_sample = g->next_sample = primary->first_sample;
  *out = g;
  return CRT_OK;
}

void crt_group_destroy(crt_group* g)
{
  if (!g) return;
  for (size_t r = 0; r < g->comms.size(); ++r) if (g->comms[r]) g->nccl.CommDestroy(g->comms[r]);
  for (size_t r = 0; r < g->members.size(); ++r) {
    cudaSetDevice(g->members[r]->device);
    if (r < g->done.size() && g->done[r]) cudaEventDestroy(g->done[r]);
    if (r < g->ev_wave.size() && g->ev_wave[r]) cudaEventDestroy(g->ev_wave[r]);
    if (r < g->ev_estimate.size() && g->ev_estimate[r]) cudaEventDestroy(g->ev_estimate[r]);
    if (r < g->t0.size() && g->t0[r]) cudaEventDestroy(g->t0[r]);
    if (r < g->t1.size() && g->t1[r]) cudaEventDestroy(g->t1[r]);
  }
  if (!g->members.empty()) { cudaSetDevice(g->members[0]->device); g->nccl_sum.release(); }
  for (size_t r = 1; r < g->members.size(); ++r) crt_destroy(g->members[r]);   // member 0 stays with the caller
  delete g;
}

int crt_group_size(const crt_group* g) { return g ? (int)g->members.size() : 0; }

int crt_group_member(crt_group* g, int rank, crt_context** out)
{
  CRT_REQUIRE(g && out, "null argument");
  CRT_REQUIRE(rank >= 0 && rank < (int)g->members.size(), "rank out of range");
  *out = g->members[rank];
  return CRT_OK;
}

int crt_group_commit(crt_group* g)
{
  CRT_REQUIRE(g, "null group");
  crt_context* c0 = g->members[0];
  int rc = set_device(c0);
  if (rc) return rc;
  if (c0->geometry_dirty || g->seen_scene != c0->gen_scene) {
    DeviceLayout L;
    bool top_only = false;
    float eps = 0.0f;
    if (c0->geometry_dirty) {
      // host BVH build once; member 0 uploads and hands back the converted layout
      if ((rc = commit_geometry(c0, &L, &top_only))) return rc;
      eps = c0->ds.scene_eps;
    } else {
      // member 0 was committed (or a blob imported) outside the group: convert its blob for the replicas
      BlobView view;
      std::string err;
      if (!parse_blob(c0->blob.data(), c0->blob.size(), view, err)) return fail(CRT_ERR_FORMAT, err);
      if (!build_device_layout(view, L, err)) return fail(CRT_ERR_FORMAT, err);
      eps = view.hdr.scene_eps;
    }
    rc = for_each_member(g, [&](int r) -> int {
      if (r == 0) return CRT_OK;
      crt_context* c = g->members[r];
      int e = set_device(c);
      if (e) return e;
      if (top_only && c->has_layout) e = upload_layout_top(c, L, eps);
      else if (top_only) return fail(CRT_ERR_STATE, "replica without a layout");
      else e = upload_layout(c, L, eps);
      if (e) return e;
      c->geometry_dirty = false;
      return CRT_OK;
    });
    if (rc) return rc;
    g->seen_scene = c0->gen_scene;
    g->seen_accum = ~0ull;          // the replicas restart their accumulation below
  }
  if ((rc = upload_tables(c0))) return rc;
  if ((rc = replicate_state(g))) return rc;
  return for_each_member(g, [&](int r) -> int {
    if (r == 0) return CRT_OK;
    int e = set_device(g->members[r]);
    return e ? e : upload_tables(g->members[r]);
  });
}

int crt_group_reset_accumulation(crt_group* g, uint64_t first_sample)
{
  CRT_REQUIRE(g, "null group");
  const int rc = crt_reset_accumulation(g->members[0], first_sample);
  if (rc) return rc;
  return replicate_state(g);
}

int crt_group_render(crt_group* g, uint32_t n_samples, uint64_t* out_total)
{
  CRT_REQUIRE(g, "null group");
  crt_context* c0 = g->members[0];
  if (c0->geometry_dirty || g->seen_scene != c0->gen_scene) return fail(CRT_ERR_STATE, "crt_group_render before crt_group_commit");
  int rc = replicate_state(g);
  if (rc) return rc;
  const int n = (int)g->members.size();
  const uint64_t start = g->next_sample;
  if (c0->params.adaptive_sampling && n > 1) rc = group_render_adaptive(g, n_samples);
  else rc = for_each_member(g, [&](int r) -> int {
    // contiguous blocks, the first (n_samples mod n) members take one sample more (distributed.sample_range)
    const uint32_t base = n_samples / n, extra = n_samples % n;
    const uint32_t cnt = base + ((uint32_t)r < extra ? 1u : 0u);
    const uint64_t off = (uint64_t)base * r + std::min<uint32_t>((uint32_t)r, extra);
    crt_context* c = g->members[r];
    int e = set_device(c);
    if (e) return e;
    c->next_sample = start + off;
    if (cnt && (e = render_impl(c, cnt))) return e;
    CRT_CUDA(cudaEventRecord(g->done[r], c->stream));
    return CRT_OK;
  });
  if (rc) return rc;
  g->next_sample = start + n_samples;
  // the call is synchronous like crt_render: every member's share is finished when it returns
  rc = for_each_member(g, [&](int r) -> int {
    int e = set_device(g->members[r]);
    if (e) return e;
    CRT_CUDA(cudaStreamSynchronize(g->members[r]->stream));
    return CRT_OK;
  });
  if (rc) return rc;
  for (crt_context* c : g->members) c->next_sample = g->next_sample;   // what crt_render's out_total reports per context
  if (out_total) *out_total = g->next_sample - g->first_sample;
  return CRT_OK;
}

int crt_group_read_ldr(crt_group* g, uint8_t* rgb8, size_t stride)
{
  CRT_REQUIRE(g && rgb8, "null argument");
  return group_display(g, rgb8, stride, nullptr, 0);
}

int crt_group_read_hdr(crt_group* g, float* rgbf, size_t stride)
{
  CRT_REQUIRE(g && rgbf, "null argument");
  return group_display(g, nullptr, 0, rgbf, stride);
}

int crt_group_info(crt_group* g, int* out_peer_access, int* out_uses_nccl, double* out_last_reduce_ms)
{
  CRT_REQUIRE(g, "null group");
  if (out_peer_access) *out_peer_access = g->peer_ok ? 1 : 0;
  if (out_uses_nccl) *out_uses_nccl = g->use_nccl ? 1 : 0;
  if (out_last_reduce_ms) *out_last_reduce_ms = g->last_reduce_ms;
  return CRT_OK;
}

}  // extern "C"
