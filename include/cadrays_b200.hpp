// cadrays_b200.hpp -- header-only C++ host mirror over the C-ABI (cadrays_b200.h), using the names
// of the OCCT classes CADRays drives so an adapter inside TKOpenGl reads like the code it replaces:
//   Graphic3d_Fresnel / Graphic3d_BSDF      src/Launcher/MaterialEditor.cxx:177-201,281-331
//   Graphic3d_RenderingParams               src/Launcher/SettingsWidget.cxx:65-90,217-229,263-478
//   V3d_View::Redraw / BufferDump           src/Launcher/AppViewer.cxx:1047,1259-1262; AppGui.cxx:345-350,430
// Errors become crt::Failure (OCCT throws Standard_Failure); nothing here computes.
#pragma once
#include "cadrays_b200.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace crt {

struct Failure : std::runtime_error {
  int code;
  Failure(int c, const char* m) : std::runtime_error(m ? m : ""), code(c) {}
};
inline void check(int rc) { if (rc != CRT_OK) throw Failure(rc, crt_last_error()); }

// Graphic3d_Fresnel: Serialize() is the vec4 the kernels read.
class Fresnel {
public:
  static Fresnel CreateSchlick(float r, float g, float b) { return Fresnel({ clamp01(r), clamp01(g), clamp01(b), 0.f }); }
  static Fresnel CreateConstant(float f) { return Fresnel({ -1.f, 0.f, clamp01(f), 0.f }); }
  static Fresnel CreateConductor(float n, float k) { return Fresnel({ -2.f, n, k, 0.f }); }
  static Fresnel CreateDielectric(float ior) { return Fresnel({ -3.f, ior, 0.f, 0.f }); }
  const std::array<float, 4>& Serialize() const { return myData; }
private:
  explicit Fresnel(std::array<float, 4> d) : myData(d) {}
  static float clamp01(float v) { return std::min(std::max(v, 0.f), 1.f); }
  std::array<float, 4> myData;
};

// Graphic3d_BSDF with CADRays' clamp + normalisation (MaterialEditor.cxx:294-329).
struct BSDF {
  float Kc[4] = { 0, 0, 0, 0 }, Kd[3] = { 0, 0, 0 }, Ks[4] = { 0, 0, 0, 0 }, Kt[3] = { 0, 0, 0 }, Le[3] = { 0, 0, 0 };
  float Absorption[4] = { 0, 0, 0, 0 };
  Fresnel FresnelCoat = Fresnel::CreateConstant(0.f), FresnelBase = Fresnel::CreateConstant(1.f);

  static BSDF CreateDiffuse(float r, float g, float b) { BSDF s; s.Kd[0] = r; s.Kd[1] = g; s.Kd[2] = b; return s; }
  static BSDF CreateMetallic(float r, float g, float b, const Fresnel& f, float roughness)
  { BSDF s; s.Ks[0] = r; s.Ks[1] = g; s.Ks[2] = b; s.Ks[3] = roughness; s.FresnelBase = f; return s; }
  static BSDF CreateGlass(float r, float g, float b, float ar, float ag, float ab, float coeff, float ior)
  {
    BSDF s; s.Kt[0] = r; s.Kt[1] = g; s.Kt[2] = b; s.Kc[0] = s.Kc[1] = s.Kc[2] = 1.f;
    s.Absorption[0] = ar; s.Absorption[1] = ag; s.Absorption[2] = ab; s.Absorption[3] = coeff;
    s.FresnelCoat = Fresnel::CreateDielectric(ior); return s;
  }
  void Normalize()
  {
    auto c01 = [](float& v) { v = std::min(std::max(v, 0.f), 1.f); };
    for (int k = 0; k < 3; ++k) { c01(Kc[k]); c01(Kd[k]); c01(Ks[k]); c01(Kt[k]); c01(Absorption[k]); Le[k] = std::max(Le[k], 0.f); }
    Absorption[3] = std::max(Absorption[3], 0.f);
    float mx = 0.f;
    for (int k = 0; k < 3; ++k) mx = std::max(mx, Kd[k] + Ks[k] + Kt[k]);
    if (mx > 1.f) for (int k = 0; k < 3; ++k) { Kd[k] /= mx; Ks[k] /= mx; Kt[k] /= mx; }
  }
  crt_bsdf Record() const
  {
    crt_bsdf r; std::memset(&r, 0, sizeof r);
    std::memcpy(r.Kc, Kc, 16); std::memcpy(r.Kd, Kd, 12); std::memcpy(r.Ks, Ks, 16); std::memcpy(r.Kt, Kt, 12);
    std::memcpy(r.Le, Le, 12); std::memcpy(r.Absorption, Absorption, 16);
    std::memcpy(r.FresnelCoat, FresnelCoat.Serialize().data(), 16);
    std::memcpy(r.FresnelBase, FresnelBase.Serialize().data(), 16);
    return r;
  }
};

// The Graphic3d_RenderingParams fields CADRays sets.
struct RenderingParams {
  int RaytracingDepth = 8, SamplesPerPixel = 1;
  float RadianceClampingValue = 50.f;
  bool TwoSidedBsdfModels = false, CoherentPathTracingMode = false, UseEnvironmentMapBackground = true;
  bool ToneMappingFilmic = false;
  bool AdaptiveScreenSampling = false;   // SettingsWidget.cxx:70,427-436
  int NbRayTracingTiles = 128;           // SettingsWidget.cxx:72 (CADRays' default)
  float WhitePoint = 1.f, Exposure = 0.f, CameraApertureRadius = 0.f, CameraFocalPlaneDist = 1.f;
  crt_params Record() const
  {
    crt_params p; crt_params_default(&p);
    p.max_depth = RaytracingDepth; p.max_radiance = RadianceClampingValue; p.two_sided = TwoSidedBsdfModels;
    p.coherent_rng = CoherentPathTracingMode; p.aperture_radius = CameraApertureRadius; p.focal_dist = CameraFocalPlaneDist;
    p.tone_map = ToneMappingFilmic; p.white_point = WhitePoint; p.exposure = Exposure;
    p.env_as_background = UseEnvironmentMapBackground;
    p.adaptive_sampling = AdaptiveScreenSampling; p.adaptive_tiles = NbRayTracingTiles;
    return p;
  }
};

// One render target (V3d_View + its OpenGl_View) on one GPU, or -- constructed with a device list -- on every listed
// GPU of the box from this one process (crt_group: Update / Redraw / BufferDump go to the group, the rest is set on
// the view as before and replicated by the library).
class View {
public:
  explicit View(int theDevice) { check(crt_create(theDevice, &myCtx)); }
  explicit View(const std::vector<int>& theDevices)
  {
    if (theDevices.empty()) throw Failure(CRT_ERR_INVALID_ARG, "empty device list");
    check(crt_create(theDevices[0], &myCtx));
    const int rc = crt_group_create(myCtx, theDevices.data(), (int)theDevices.size(), &myGroup);
    if (rc != CRT_OK) { Failure f(rc, crt_last_error()); crt_destroy(myCtx); throw f; }
  }
  struct HostOnly {};
  explicit View(HostOnly) { check(crt_create_host_only(&myCtx)); }
  ~View() { crt_group_destroy(myGroup); crt_destroy(myCtx); }
  View(const View&) = delete;
  View& operator=(const View&) = delete;

  uint32_t AddTriangulation(const float* pos, const float* nrm, const float* uv, uint32_t nVerts, const uint32_t* idx, uint32_t nTris)
  { uint32_t id = 0; check(crt_mesh_create(myCtx, pos, nrm, uv, nVerts, idx, nTris, &id)); return id; }
  uint32_t Display(uint32_t mesh, const float trsf3x4[12], uint32_t material)
  { uint32_t id = 0; check(crt_instance_add(myCtx, mesh, trsf3x4, material, &id)); return id; }
  void SetLocation(uint32_t inst, const float trsf3x4[12]) { check(crt_instance_set_transform(myCtx, inst, trsf3x4)); }
  void SetVisible(uint32_t inst, bool visible) { check(crt_instance_set_visible(myCtx, inst, visible)); }   // Erase / Display
  void Clear() { check(crt_scene_clear(myCtx)); }
  void SetMaterials(const std::vector<BSDF>& m)
  {
    std::vector<crt_bsdf> r; r.reserve(m.size());
    for (const BSDF& b : m) r.push_back(b.Record());
    check(crt_materials_set(myCtx, r.data(), (uint32_t)r.size()));
  }
  void SetLights(const std::vector<crt_light>& l) { check(crt_lights_set(myCtx, l.data(), (uint32_t)l.size())); }
  void SetTextureEnv(const uint8_t* rgb, uint32_t w, uint32_t h) { check(crt_envmap_set_rgb8(myCtx, rgb, w, h)); }
  void SetRenderingParams(const RenderingParams& p) { myParams = p; crt_params r = p.Record(); check(crt_params_set(myCtx, &r)); }
  const RenderingParams& RenderingParameters() const { return myParams; }
  void SetCamera(const crt_camera& c) { myCam = c; myHasCam = true; check(crt_camera_set(myCtx, &c)); }
  void SetWindowSize(uint32_t w, uint32_t h) { check(crt_resize(myCtx, w, h)); myW = w; myH = h; }
  void Update() { check(myGroup ? crt_group_commit(myGroup) : crt_commit(myCtx)); }
  uint64_t Redraw() { return Redraw((uint32_t)std::max(1, myParams.SamplesPerPixel)); }
  uint64_t Redraw(uint32_t samples)
  { uint64_t n = 0; check(myGroup ? crt_group_render(myGroup, samples, &n) : crt_render(myCtx, samples, &n)); return n; }
  // BufferDump(Image_PixMap&, Graphic3d_BT_RGB) returns bool in OCCT (AppGui.cxx:430-433)
  bool BufferDump(std::vector<uint8_t>& rgb8)
  {
    rgb8.resize((size_t)myW * myH * 3);
    return (myGroup ? crt_group_read_ldr(myGroup, rgb8.data(), 0) : crt_read_ldr(myCtx, rgb8.data(), 0)) == CRT_OK;
  }
  bool BufferDumpHdr(std::vector<float>& rgb)
  {
    rgb.resize((size_t)myW * myH * 3);
    return (myGroup ? crt_group_read_hdr(myGroup, rgb.data(), 0) : crt_read_hdr(myCtx, rgb.data(), 0)) == CRT_OK;
  }
  int Members() const { return myGroup ? crt_group_size(myGroup) : 1; }
  crt_group* Group() const { return myGroup; }
  // V3d_View::ToPixMap(Image_PixMap&, width, height): off-screen render at the given size (the camera passed in keeps
  // its pose; its aspect is set from the size), `samples` samples per pixel, RGB8 dump.  Returns false like OCCT.
  bool ToPixMap(std::vector<uint8_t>& rgb8, uint32_t w, uint32_t h, crt_camera cam, uint32_t samples)
  {
    // OCCT renders ToPixMap into its own FBO and leaves the view untouched: window size and camera are restored
    const uint32_t oldW = myW, oldH = myH;
    const crt_camera oldCam = myCam;
    const bool hadCam = myHasCam;
    SetWindowSize(w, h);
    cam.aspect = (float)w / (float)h;
    SetCamera(cam);
    uint64_t n = 0;
    bool ok = (myGroup ? crt_group_render(myGroup, samples, &n) : crt_render(myCtx, samples, &n)) == CRT_OK;
    ok = ok && BufferDump(rgb8);
    if (oldW && oldH && (oldW != w || oldH != h)) SetWindowSize(oldW, oldH);
    if (hadCam) SetCamera(oldCam); else { myHasCam = false; }
    return ok;
  }
  std::vector<uint8_t> ExportBVH()
  {
    size_t n = 0; check(crt_bvh_export(myCtx, nullptr, 0, &n));
    std::vector<uint8_t> b(n); check(crt_bvh_export(myCtx, b.data(), n, &n)); return b;
  }
  crt_context* Handle() const { return myCtx; }
private:
  crt_context* myCtx = nullptr;
  crt_group* myGroup = nullptr;
  RenderingParams myParams;
  uint32_t myW = 0, myH = 0;
  crt_camera myCam{};
  bool myHasCam = false;
};

}  // namespace crt
