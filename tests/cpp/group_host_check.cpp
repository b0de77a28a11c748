// C++ host driving several GPUs through crt::View(device list) -> crt_group (the way an OCCT adapter's
// V3d_View::Redraw / BufferDump would, AppViewer.cxx:1047,1259-1262): the combined frame of the group must equal
// the frame of one context that rendered every sample itself, up to float summation order.
// usage: group_host_check d0 [d1 ...]   (device ordinals; an ordinal may repeat)
#include "cadrays_b200.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

static void fill(crt::View& v, uint32_t w, uint32_t h)
{
  // a floor, a tilted quad and an emissive strip: diffuse + glossy + light sampling, 2 objects sharing one mesh
  const float quad[12] = { -1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0 };
  const uint32_t idx[6] = { 0, 1, 2, 0, 2, 3 };
  const uint32_t m = v.AddTriangulation(quad, nullptr, nullptr, 4, idx, 2);
  const float floor_xf[12] = { 3, 0, 0, 0, 0, 3, 0, 0, 0, 0, 1, 0 };
  const float tilt_xf[12] = { 0.6f, 0, 0, 0, 0, 0.42f, -0.42f, 0.3f, 0, 0.42f, 0.42f, 0.6f };
  const float lamp_xf[12] = { 0.3f, 0, 0, 0, 0, 0.3f, 0, 0, 0, 0, -1, 2.5f };
  v.Display(m, floor_xf, 0);
  v.Display(m, tilt_xf, 1);
  v.Display(m, lamp_xf, 2);
  crt::BSDF lamp = crt::BSDF::CreateDiffuse(0, 0, 0);
  lamp.Le[0] = lamp.Le[1] = lamp.Le[2] = 8.f;
  v.SetMaterials({ crt::BSDF::CreateDiffuse(0.7f, 0.7f, 0.7f),
                   crt::BSDF::CreateMetallic(0.9f, 0.6f, 0.2f, crt::Fresnel::CreateSchlick(0.9f, 0.6f, 0.2f), 0.2f), lamp });
  crt_light sun = {};
  sun.emission[0] = sun.emission[1] = sun.emission[2] = 3.f; sun.smoothness = 0.05f;
  sun.posdir[0] = -0.3f; sun.posdir[1] = 0.4f; sun.posdir[2] = -1.f;
  v.SetLights({ sun });
  crt::RenderingParams p;
  p.RaytracingDepth = 5;
  v.SetRenderingParams(p);
  crt_camera cam = {};
  cam.eye[0] = 0; cam.eye[1] = -4; cam.eye[2] = 1.5f; cam.dir[1] = 1; cam.dir[2] = -0.25f; cam.up[2] = 1;
  cam.fovy_deg = 45; cam.aspect = (float)w / (float)h;
  v.SetCamera(cam);
  v.SetWindowSize(w, h);
  v.Update();
}

int main(int argc, char** argv)
{
  std::vector<int> devices;
  for (int k = 1; k < argc; ++k) devices.push_back(std::atoi(argv[k]));
  if (devices.empty()) devices.push_back(0);
  const uint32_t W = 256, H = 144, SPP = 24;
  try {
    crt::View one(devices[0]);
    fill(one, W, H);
    one.Redraw(SPP);
    std::vector<float> ref;
    std::vector<uint8_t> ref8;
    if (!one.BufferDumpHdr(ref) || !one.BufferDump(ref8)) return 2;

    crt::View many(devices);
    fill(many, W, H);
    if (many.Members() != (int)devices.size()) return 3;
    // three Redraw calls of uneven size: the group's sample cursor must continue across calls
    if (many.Redraw(7) != 7 || many.Redraw(1) != 8 || many.Redraw(SPP - 8) != SPP) return 4;
    std::vector<float> got;
    std::vector<uint8_t> got8;
    if (!many.BufferDumpHdr(got) || !many.BufferDump(got8)) return 5;
    double worst = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
      const double d = std::fabs((double)got[i] - ref[i]) / std::fmax(std::fabs((double)ref[i]), 1e-3);
      if (d > worst) worst = d;
    }
    size_t off8 = 0;
    for (size_t i = 0; i < ref8.size(); ++i) {
      const int d = std::abs((int)got8[i] - (int)ref8[i]);
      if (d > 1) return 6;
      off8 += d != 0;
    }
    int peer = 0, nccl = 0;
    double ms = 0;
    crt_group_info(many.Group(), &peer, &nccl, &ms);
    std::printf("group of %zu: max relative HDR difference %.3g, %zu of %zu RGB8 values differ by one step, peer access %d, nccl %d, "
                "exchange+display %.3f ms\n", devices.size(), worst, off8, ref8.size(), peer, nccl, ms);
    if (worst > 2e-4) return 7;
    // moving an object through the group: top-level patch on every member, image equals a fresh single-GPU build
    const float moved[12] = { 0.6f, 0, 0, 0.5f, 0, 0.42f, -0.42f, 0.1f, 0, 0.42f, 0.42f, 0.8f };
    many.SetLocation(1, moved);
    many.Update();
    many.Redraw(8);
    one.SetLocation(1, moved);
    one.Update();
    one.Redraw(8);
    if (!one.BufferDumpHdr(ref) || !many.BufferDumpHdr(got)) return 8;
    worst = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
      const double d = std::fabs((double)got[i] - ref[i]) / std::fmax(std::fabs((double)ref[i]), 1e-3);
      if (d > worst) worst = d;
    }
    std::printf("after SetLocation: max relative HDR difference %.3g\n", worst);
    if (worst > 2e-4) return 9;
  } catch (const crt::Failure& f) {
    std::fprintf(stderr, "failure %d: %s\n", f.code, f.what());
    return 1;
  }
  return 0;
}
